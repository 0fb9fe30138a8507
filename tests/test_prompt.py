"""Prompt / token contract (omchat_b200/prompt.py) against vectors produced by the reference's own functions
(tests/golden/make_golden_prompt.py: mm_utils.tokenizer_image_token, make_context.make_context, KeywordsStoppingCriteria):
ids must match one for one (integer work: bit-exact)."""
import json
import os

import torch

from omchat_b200 import prompt as P
from toy_tokenizer import ToyTokenizer

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_prompt.json")))


def test_tokenizer_image_token_matches_reference():
    assert len(GOLD["tokenizer_image_token"]) == 14
    for case in GOLD["tokenizer_image_token"]:
        tok = ToyTokenizer(bos_token_id=case["bos"])
        assert P.tokenizer_image_token(case["prompt"], tok) == case["ids"], case
    t = P.tokenizer_image_token("<image>\nhi", ToyTokenizer(), return_tensors="pt")
    assert t.dtype == torch.long and int((t == -200).sum()) == 1


def test_make_context_matches_reference():
    for case in GOLD["make_context"]:
        args = dict(case["args"])
        if args.get("history") is not None:
            args["history"] = [tuple(h) for h in args["history"]]
        text, ids = P.make_context(ToyTokenizer(), **args)
        assert text == case["text"] and ids == case["ids"], case["args"]
    # the window bound drops old turns (third case) and every placeholder survives (first case)
    assert GOLD["make_context"][0]["ids"].count(-200) == 2
    assert len(GOLD["make_context"][2]["ids"]) < 300 + 40


def test_image_prompt_shape():
    s = P.image_prompt(5, "  What is <image> this? ")
    assert s == "<image>\npatch:<image>\npatch:<image>\npatch:<image>\npatch:<image>\nWhat is  this?"
    assert P.tokenizer_image_token(s, ToyTokenizer()).count(-200) == 5


def test_keywords_stopping_matches_reference():
    for case in GOLD["stopping"]:
        tok = ToyTokenizer(bos_token_id=case["bos"])
        prompt = torch.tensor([tok.encode("hello ")])
        crit = P.KeywordsStoppingCriteria(["<|im_end|>", "STOP"], tok, prompt)
        ids = torch.tensor([tok.encode("hello ") + tok.encode_nobos(case["tail"])])
        assert bool(crit(ids, None)) == case["stop"], case


def test_processor_context_matches_reference():
    """OmChatProcessor.__call__ (omchat_b200/processing.py) builds the same ids as the reference's HF processor for one and
    several images (crop counts from a stub image processor: no GPU needed)."""
    from types import SimpleNamespace
    from omchat_b200.processing import BatchFeature, OmChatProcessor
    assert len(GOLD["processor"]) == 4
    for case in GOLD["processor"]:
        nums = case["num_patches"]
        mx = max(nums)
        ip = lambda images, return_tensors=None, n=nums, mx=mx: BatchFeature(  # noqa: E731
            pixel_values=torch.zeros(len(n), mx, 3, 2, 2), num_patches=torch.tensor(n))
        proc = OmChatProcessor(image_processor=ip, tokenizer=ToyTokenizer())
        out = proc(case["text"], images=[object()] * len(nums))
        assert out.input_ids[0].tolist() == case["ids"], case["text"]
        assert out["images"].shape[0] == case["n_images"] == sum(nums)
        assert out.input_ids[0].tolist().count(-200) == sum(nums)
    out = OmChatProcessor(tokenizer=ToyTokenizer())("just <image> text")
    assert "images" not in out and out.input_ids.shape[0] == 1
