"""Debug helper: variant-A tiny MoE model, greedy 8 tokens on the per-op / stream paths, eager and graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
from tiny import TINY_MOE as T, tiny_inputs, tiny_state_dict_moe
from test_moe_gpu import moe_cfgs
from omchat_b200.model.moe import OmChatQwen2MoeForCausalLM

g = torch.load(os.path.join(ROOT, "tests/golden/golden_tiny_moe.pt"), weights_only=False)["A"]
sd = {k: v.to(torch.bfloat16).float() for k, v in tiny_state_dict_moe(0, ()).items()}
model = OmChatQwen2MoeForCausalLM.from_state_dict(sd, moe_cfgs(g), device="cuda")
dec = model.get_model().decoder
pixels, _ = tiny_inputs(1)
ids = g["prefill_ids"]
print("reference:", g["greedy_tokens"])
for stream in (False, True):
    for graph in (False, True):
        dec.stream_enabled = stream
        res = model(input_ids=ids, images=pixels[:1], use_cache=True)
        cache = res.past_key_values
        first = res.logits[0, -1].argmax().view(1)
        toks = dec.generate_greedy(first, cache, 7, use_graph=graph)
        print(f"stream={stream} graph={graph}: {[int(first)] + toks[0].tolist()}")
    # eager, teacher-forced with the reference tokens: logits margins + routing of both layers
    res = model(input_ids=ids, images=pixels[:1], use_cache=True)
    cache = res.past_key_values
    for i, t in enumerate(g["greedy_tokens"][:4]):
        lg = dec.decode_step(torch.tensor([t], device="cuda"), cache).clone()
        ws = dec._workspace(1)
        top2 = torch.topk(lg[0], 2)
        print(f"  stream={stream} step {i}: in {t} -> argmax {int(top2.indices[0])} (2nd {int(top2.indices[1])}, margin {float(top2.values[0]-top2.values[1]):.4f}); last layer routing {ws.topk_ids[0].tolist()} w {[round(float(x),4) for x in ws.topk_w[0]]}")
