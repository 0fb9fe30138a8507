"""Opcode histogram of every compiled kernel object (cuobjdump -sass on omchat_b200/_lib/*.o): the proof, independent of
nvcc being installed where it is read, that the tensor-core kernels are tcgen05 / TMEM / TMA code. Writes one table per
object: total instructions and the counts of the mnemonics that matter (UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = cp.async.bulk, SYNCS = mbarrier, HMMA =
mma.sync, LDSM = ldmatrix, LDG / STG / LDS / STS, MUFU, FFMA2 ...).   python tools/sass_hist.py > profiles/rNN_sass_opcodes.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "LDSM",
       "LDGSTS", "LDG", "STG", "LDS", "STS", "LDL", "STL", "MUFU", "FFMA2", "FFMA", "FADD2", "FMNMX3", "BAR", "ATOM", "RED",
       "SHFL", "UCGABAR", "CCTL", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS"]


def main():
    for obj in sorted(glob.glob(os.path.join(ROOT, "omchat_b200", "_lib", "*.o"))):
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        per_fn, cur = collections.OrderedDict(), None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                per_fn[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and cur is not None:
                per_fn[cur][m.group(1)] += 1
        print(f"=== {os.path.basename(obj)}: {len(per_fn)} kernels")
        for fn, c in per_fn.items():
            total = sum(c.values())
            base = collections.Counter()
            for op, n in c.items():
                base[op.split(".")[0]] += n
            picks = []
            for k in KEY:
                n = c.get(k, 0) if "." in k else base.get(k, 0)
                if "." in k:
                    n = sum(v for op, v in c.items() if op.startswith(k))
                if n:
                    picks.append(f"{k}={n}")
            demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
            print(f"  {demangled}\n      {total} instr: " + " ".join(picks))
    return 0


if __name__ == "__main__":
    sys.exit(main())
