#!/usr/bin/env python
"""Multi-GPU check of the tensor-parallel model API (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/tp_forward_check.py

Every rank builds the same model sharded N ways (seed 0) and, on rank 0 only, the unsharded model of the same seed. Checks:
  * forward(logits_to_keep=1) and a decode-step forward() return FULL-vocabulary logits on every rank (the vocab shards are
    all-gathered), equal across ranks and equal to the unsharded model's within bf16 tolerance (cosine >= 0.999);
  * encode_images with the crops data-parallel over the ranks + one feature all-gather equals the replicated tower bit for bit;
  * generate() for a batch of 8 (batched step with the all-reduce fused into the GEMM epilogues) gives the same ids on every
    rank and the unsharded model's ids up to near-ties.
Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from omchat_b200.config import InternVisionConfig, OmChatQwen2Config
    from omchat_b200.model.omchat import OmChatQwen2ForCausalLM
    cfg = OmChatQwen2Config(num_hidden_layers=3, eos_token_id=-1, mm_pixel_shuffle_ratio=0.5,
                            vision_config=InternVisionConfig(num_hidden_layers=2))
    tp = OmChatQwen2ForCausalLM(cfg, device=f"cuda:{local}", seed=0, tp_rank=rank, tp_size=world, tp_group=dist.group.WORLD)
    full = OmChatQwen2ForCausalLM(cfg, device=f"cuda:{local}", seed=0)  # every rank: the reference it compares itself with
    g = torch.Generator().manual_seed(3)
    n_img = 5
    pixels = torch.randn(n_img, 3, 448, 448, generator=g)
    B = 8
    ids = torch.randint(0, 151643, (B, 40), generator=g)
    for b in range(n_img):
        ids[b, 7] = -200  # the first five prompts carry one image each; a prompt without placeholder still consumes none here
    ids_img, ids_txt = ids[:n_img], ids[n_img:]
    out = {"world": world}

    def stage(name):
        torch.cuda.synchronize()
        print(f"[rank {rank}] {name}", file=sys.stderr, flush=True)

    # ---- vision data-parallel + all-gather == replicated tower
    f_tp, f_full = tp.encode_images(pixels), full.encode_images(pixels)
    out["encode_images_bit_equal"] = bool(torch.equal(f_tp, f_full))
    stage("encode_images done")
    # ---- forward: full-vocabulary logits on every rank
    r_tp = tp(input_ids=ids_img, images=pixels, logits_to_keep=1, max_cache_len=512)
    r_full = full(input_ids=ids_img, images=pixels, logits_to_keep=1, max_cache_len=512)
    lg_tp, lg_full = r_tp.logits[:, 0], r_full.logits[:, 0]
    assert lg_tp.shape == (n_img, cfg.vocab_size), lg_tp.shape
    cos = torch.nn.functional.cosine_similarity(lg_tp, lg_full, dim=-1).min().item()
    gathered = [torch.empty_like(lg_tp) for _ in range(world)]
    dist.all_gather(gathered, lg_tp.contiguous())
    out["prefill_logits_cos_vs_unsharded"] = cos
    out["prefill_logits_equal_across_ranks"] = all(bool(torch.equal(gathered[0], x)) for x in gathered)
    stage("prefill forward done")
    tok = lg_full.argmax(-1)
    d_tp = tp(input_ids=tok.view(-1, 1), past_key_values=r_tp.past_key_values).logits[:, 0]
    d_full = full(input_ids=tok.view(-1, 1), past_key_values=r_full.past_key_values).logits[:, 0]
    out["decode_logits_cos_vs_unsharded"] = torch.nn.functional.cosine_similarity(d_tp, d_full, dim=-1).min().item()
    assert d_tp.shape == (n_img, cfg.vocab_size)
    stage("decode forward done")
    # ---- generate, batch 8 (text-only rows use a separate call: images are consumed in batch-major order)
    g_tp = tp.generate(ids_img, images=pixels, max_new_tokens=12, do_sample=False, eos_token_id=-1)
    g_full = full.generate(ids_img, images=pixels, max_new_tokens=12, do_sample=False, eos_token_id=-1)
    stage("generate done")
    same = (g_tp == g_full).all(dim=1)
    out["generate_rows_equal_to_unsharded"] = f"{int(same.sum())}/{n_img}"
    allg = [torch.empty_like(g_tp) for _ in range(world)]
    dist.all_gather(allg, g_tp.contiguous())
    out["generate_equal_across_ranks"] = all(bool(torch.equal(allg[0], x)) for x in allg)
    ok = (out["encode_images_bit_equal"] and out["prefill_logits_equal_across_ranks"] and out["generate_equal_across_ranks"]
          and cos >= 0.999 and out["decode_logits_cos_vs_unsharded"] >= 0.999)
    out["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    tp.close()  # captured graphs hold NCCL work: release them before the communicator goes away
    full.close()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
