#!/usr/bin/env python
"""Turn ncu output into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches <launches.csv> [--only omc]   # per-kernel totals and shares of a launch list
  python tools/ncu_summary.py full <report.ncu-rep>                   # key metrics of every launch in a --set full report
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v * 1e6 if unit in ("s", "second") else v


def launches(path, only=None):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for r in csv.DictReader(lines[start:]):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], to_us(r["Metric Value"], r["Metric Unit"]), r["Grid Size"], r["Block Size"]))
    if only:
        rows = [r for r in rows if only in r[0]]
    agg = collections.OrderedDict()
    for k, v, g, b in rows:
        k = re.sub(r"\(.*", "", k)
        a = agg.setdefault(k, [0, 0.0, g, b])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':58s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}  grid / block (first)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:58]:58s} {a[0]:5d} {a[1]:10.1f} {a[1] / a[0]:9.2f} {a[1] / tot:6.3f}  {a[2]} / {a[3]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(out.splitlines())
    hdr, units = next(rd), next(rd)
    print(f"# {path}")
    for i, r in enumerate(rd):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"launch {i}: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"    {k:78s} {d[k]:>14s} {u[k]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
        launches(sys.argv[2], only)
    else:
        full(sys.argv[2])
