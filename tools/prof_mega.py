"""In-kernel timeline of the persistent decode kernel: per-op %globaltimer stamps of every CTA (OMCHAT_B200_MEGA_PROF=1).
Prints, per op kind, the mean per layer of: wait + activation staging, body, op span (first CTA in -> last CTA out) and
end skew; K-chunk sub-ops of one matrix are summed. OMCHAT_B200_MEGA_PROF=0: only time the step (no stamps).
With a library built with -DOMC_MEGA_DETAIL=1 (OMCHAT_B200_LIB=...) also warp 0's cycle breakdown of the GEMV loop and
stamps inside the attention phase.  usage: prof_mega.py [layers] [batch] [ctx]"""
import json, os, sys
os.environ.setdefault("OMCHAT_B200_MEGA_PROF", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omchat_b200 import lib
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
print(f"step time {e0.elapsed_time(e1) / 10 * 1000:.1f} us (profiling stamps {'on' if plan.prof is not None else 'off'}), "
      f"{plan.n_ops} ops")
if plan.prof is None:
    sys.exit(0)
raw = plan.prof.cpu().double()  # [grid, n_ops, 8]
smid = raw[:, 0, 3].long().tolist()
p = (raw - raw[:, 0, 0].min()) / 1000.0  # us
names = lib.mega_op_kinds(plan)
order = ["qkv", "attn", "o", "gate_up", "down", "lm_head", "final"]
idx = {k: [i for i, n in enumerate(names) if n == k] for k in order}
nl = max(layers, 1)
print(f"{'op':8s} {'ops/layer':>9s} {'wait+stage':>10s} {'body':>9s} {'op span':>9s} {'end skew':>9s}   (us per layer, sub-ops summed)")
for k in order:
    ii = idx[k]
    if not ii:
        continue
    per = nl if k not in ("lm_head", "final") else 1
    if k in ("attn", "final"):
        stg = 0.0
        body = (p[:, ii, 2] - p[:, ii, 0]).mean(dim=0).sum().item() / per
    else:
        staged = torch.maximum(p[:, ii, 1], p[:, ii, 0])  # ops that keep the previous activation vector have no staging stamp
        stg = (staged - p[:, ii, 0]).mean(dim=0).sum().item() / per
        body = (p[:, ii, 2] - staged).mean(dim=0).sum().item() / per
    span = (p[:, ii, 2].max(dim=0).values - p[:, ii, 0].min(dim=0).values).sum().item() / per
    skew = (p[:, ii, 2].max(dim=0).values - p[:, ii, 2].min(dim=0).values).mean().item()
    print(f"{k:8s} {len(ii) / per:9.1f} {stg:10.2f} {body:9.2f} {span:9.2f} {skew:9.2f}")
cyc = raw[:, :, 4:8]
gi = [i for i, n in enumerate(names) if n in ("qkv", "o", "gate_up", "down", "lm_head")]
if cyc[:, gi, :].sum() > 0:  # detail build
    print("warp 0 cycle breakdown (us at 1.965 GHz, per layer): stage wait / dots / issue / reduce+epilogue")
    for k in ("qkv", "o", "gate_up", "down", "lm_head"):
        per = nl if k != "lm_head" else 1
        m = cyc[:, idx[k], :].mean(dim=0).sum(dim=0) / 1965.0 / per
        print(f"  {k:8s} {m[0].item():7.2f} {m[1].item():7.2f} {m[2].item():7.2f} {m[3].item():7.2f}")
    ai = idx["attn"]
    stamps = (raw[:, ai, 4:8] - raw[:, ai, 0:1]) / 1000.0
    for j, nm in enumerate(("q arrived", "own partial published", "warp 0: first K/V row in registers", "P V done")):
        v = stamps[:, :, j]
        v = v[v > 0]
        if v.numel():
            print(f"  attn: {nm:34s} +{v.mean().item():6.2f} us after phase start (max {v.max().item():6.2f}, n={v.numel()})")
print(f"whole step (first op start -> last op end): {(p[:, -1, 2].max() - p[:, 0, 0].min()).item():.1f} us")
if layers > 2:
    last_of_layer = [i for i in idx["down"] if i + 1 >= len(names) or names[i + 1] != "down"]
    ends = p[:, last_of_layer, 2].max(dim=0).values
    print(f"per-layer period (max-CTA end of the layer's last op, mean over layers): {(ends[1:] - ends[:-1]).mean().item():.2f} us")
    rec = {"smid": smid, "slot_bytes": int(os.environ.get("OMCHAT_B200_MEGA_SLOT", "0")), "names": names}
    out = os.environ.get("OMCHAT_B200_PROF_OUT")
    if out:
        rec["op_start_us"] = p[:, :, 0].tolist()
        rec["op_end_us"] = p[:, :, 2].tolist()
        json.dump(rec, open(out, "w"))
