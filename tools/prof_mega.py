"""In-kernel timeline of the persistent decode kernel: per-op %globaltimer stamps of every CTA (OMCHAT_B200_MEGA_PROF=1).
Prints, per op kind, the mean over layers of: barrier wait, activation staging, body, and the max-over-CTAs op duration."""
import os, sys
os.environ.setdefault("OMCHAT_B200_MEGA_PROF", "1")  # OMCHAT_B200_MEGA_PROF=0: only time the step (no stamps)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omchat_b200.config import OmChatQwen2Config
from omchat_b200.model.decoder import Qwen2Decoder
from omchat_b200.model.weights import random_init

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 28
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
cfg = OmChatQwen2Config(num_hidden_layers=layers)
w = random_init(cfg, device="cuda", vision=False)
dec = Qwen2Decoder(cfg, w.llm)
cache = dec.new_cache(B, ctx + 64)
cache.pool.normal_(0, 0.5)
cache.host_lens = [ctx] * B
cache.ctx_lens.fill_(ctx)
toks = torch.randint(0, cfg.vocab_size, (B,), device="cuda")
st = dec._decode_state(B, cache.capacity)
plan = dec._mega_plan(st, cache)
for _ in range(5):
    dec.decode_step(toks, cache)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    plan.step()
e1.record()
torch.cuda.synchronize()
print(f"step time {e0.elapsed_time(e1) / 10 * 1000:.1f} us (profiling stamps {'on' if plan.prof is not None else 'off'})")
if plan.prof is None:
    sys.exit(0)
p = plan.prof.cpu().double()  # [grid, n_ops, 8]
smid = p[:, 0, 3].long().tolist()
cyc = p[:, :, 4:8].clone()    # warp 0: cycles in stage wait, dot products, refill issue, reduce + epilogue
t0 = p[:, 0, 0].min()
p = (p - t0) / 1000.0  # us
n_ops = p.shape[1]
names = ["qkv", "attn", "o", "gate_up", "down"]
kinds = {}
for i in range(n_ops):
    k = names[i % 5] if i < 5 * layers else ("lm_head" if i == 5 * layers else "final")
    kinds.setdefault(k, []).append(i)
print(f"{'op':8s} {'n':>3s} {'wait+stage':>10s} {'body':>9s} {'op span':>9s} {'end skew':>9s}")
for k, idx in kinds.items():
    if k in ("attn", "final"):
        stg = 0.0
        body = (p[:, idx, 2] - p[:, idx, 0]).mean().item()
    else:
        stg = (p[:, idx, 1] - p[:, idx, 0]).mean().item()
        body = (p[:, idx, 2] - p[:, idx, 1]).mean().item()
    span = (p[:, idx, 2].max(dim=0).values - p[:, idx, 0].min(dim=0).values).mean().item()
    skew = (p[:, idx, 2].max(dim=0).values - p[:, idx, 2].min(dim=0).values).mean().item()
    print(f"{k:8s} {len(idx):3d} {stg:10.2f} {body:9.2f} {span:9.2f} {skew:9.2f}")
print("warp 0 cycle breakdown per op (mean over CTAs and layers, us at 1.965 GHz): stage wait / dots / issue / reduce+epilogue")
for k, idx in kinds.items():
    if k in ("attn", "final"):
        continue
    m = cyc[:, idx, :].mean(dim=(0, 1)) / 1965.0
    print(f"  {k:8s} {m[0].item():7.2f} {m[1].item():7.2f} {m[2].item():7.2f} {m[3].item():7.2f}")
ai = kinds["attn"]
raw = plan.prof.cpu().double()
if raw[:, ai, 4].sum() > 0:  # detail build: stamps inside the attention phase (relative to the CTA's phase start)
    st = (raw[:, ai, 4:8] - raw[:, ai, 0:1]) / 1000.0
    for j, nm in enumerate(("q arrived", "own partial published", "warp 0: first K/V row in registers", "warp 0 done with its keys")):
        v = st[:, :, j]
        v = v[v > 0]
        print(f"  attn: {nm:34s} +{v.mean().item():6.2f} us after phase start (max {v.max().item():6.2f}, n={v.numel()})")
print(f"whole step (first op start -> last op end): {(p[:, -1, 2].max() - p[:, 0, 0].min()).item():.1f} us")
# critical path per layer: time between the slowest CTA finishing 'down' of consecutive layers
if layers > 2:
    ends = p[:, [5 * l + 4 for l in range(layers)], 2].max(dim=0).values
    print(f"per-layer period (max-CTA end of down, mean over layers): {(ends[1:] - ends[:-1]).mean().item():.2f} us")
# per-CTA detail of one mid-stack layer
li = min(layers - 1, 3)
for name, off in (("qkv", 0), ("attn", 1), ("o", 2), ("gate_up", 3), ("down", 4)):
    i = 5 * li + off
    total = p[:, i, 2] - p[:, i, 0]
    wait = (p[:, i, 1] - p[:, i, 0]) if name != "attn" else total * 0
    vals = ", ".join(f"{w:.1f}/{t:.1f}" for w, t in list(zip(wait.tolist(), total.tolist()))[:24])
    print(f"layer {li} {name}: wait/total per CTA (first 24) [{vals}]")

# per-CTA body time (mean over layers) against the SM the CTA ran on: is the end skew a property of the SM's position?
if layers > 2:
    import json
    rec = {"smid": smid, "slot_bytes": int(os.environ.get("OMCHAT_B200_MEGA_SLOT", "0"))}
    for name, off in (("qkv", 0), ("o", 2), ("gate_up", 3), ("down", 4)):
        idx = [5 * l + off for l in range(layers)]
        body = (p[:, idx, 2] - p[:, idx, 1])  # [grid, layers]
        mean, std = body.mean(dim=1), body.std(dim=1)
        rec[name] = {"mean": [round(v, 2) for v in mean.tolist()], "std_over_layers": [round(v, 2) for v in std.tolist()]}
        print(f"{name}: per-CTA mean body {mean.min().item():.2f}..{mean.max().item():.2f} us (mean {mean.mean().item():.2f}), "
              f"per-CTA std over layers mean {std.mean().item():.2f} us; spread of per-CTA means {mean.std().item():.2f} us")
    out = os.environ.get("OMCHAT_B200_PROF_OUT")
    if out:
        json.dump(rec, open(out, "w"))
