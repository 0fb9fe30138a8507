"""Golden vectors for the prompt / token contract: the REAL reference functions (omchat/mm_utils.py tokenizer_image_token,
omchat/make_context.py make_context, mm_utils.KeywordsStoppingCriteria) run in this container with the deterministic toy
tokenizer of tests/golden/toy_tokenizer.py. -> tests/golden/golden_prompt.json"""
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import torch  # noqa: E402
import transformers  # noqa: E402,F401

from toy_tokenizer import ToyTokenizer  # noqa: E402

constants = types.ModuleType("omchat.constants")
for k, v in dict(IMAGE_TOKEN_INDEX=-200, IGNORE_INDEX=-100, DEFAULT_IMAGE_TOKEN="<image>", DEFAULT_IM_START_TOKEN="<im_start>",
                 DEFAULT_IM_END_TOKEN="<im_end>").items():
    setattr(constants, k, v)
sys.modules.setdefault("omchat", types.ModuleType("omchat"))
sys.modules["omchat.constants"] = constants


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


mm = load("omchat.mm_utils", "/root/reference/omchat/mm_utils.py")
mc = load("omchat.make_context", "/root/reference/omchat/make_context.py")

PROMPTS = ["<image>\nWhat is in the picture?", "no image here", "<image>\npatch:<image>\npatch:<image>\nDescribe.",
           "a<image_0>b<image_1>c", "<image>", "tail<image>", ""]
CONTEXTS = [
    dict(query="<image>\npatch:<image>\nWhat is this?", history=None, system="You are a helpful assistant."),
    dict(query="And now?", history=[("<image>\nfirst question", "first answer"), ("second", "reply two")], system="sys"),
    dict(query="short", history=[("q" * 50, "a" * 50)] * 6, system="s", max_window_size=300),
    dict(query="raw text <image> kept", history=None, system="", chat_format="raw"),
]


def main():
    out = {"tokenizer_image_token": [], "make_context": [], "stopping": []}
    for bos in (None, 1):
        tok = ToyTokenizer(bos_token_id=bos)
        for p in PROMPTS:
            out["tokenizer_image_token"].append({"bos": bos, "prompt": p, "ids": mm.tokenizer_image_token(p, tok)})
    tok = ToyTokenizer()
    for c in CONTEXTS:
        text, ids = mc.make_context(tok, **c)
        out["make_context"].append({"args": c, "text": text, "ids": ids})
    for bos in (None, 1):
        tok = ToyTokenizer(bos_token_id=bos)
        prompt = torch.tensor([tok.encode("hello ")])
        crit = mm.KeywordsStoppingCriteria(["<|im_end|>", "STOP"], tok, prompt)
        for tail in ["abc", "abcSTOP", "ab STO", "x<|im_end|>", "STOPx", ""]:
            ids = torch.tensor([tok.encode("hello ") + tok.encode_nobos(tail)])
            out["stopping"].append({"bos": bos, "tail": tail, "stop": bool(crit(ids, None))})
    json.dump(out, open(os.path.join(HERE, "golden_prompt.json"), "w"), indent=0)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()


def processor_golden():
    """The reference OmChatProcessor.__call__ (omchat/hf/processing_omchat.py:169-246) on a stub `self` whose image
    processor only reports crop counts — what is pinned is the id sequence it builds."""
    import numpy as np
    pr = load("ref_processing_omchat", "/root/reference/omchat/hf/processing_omchat.py")
    from types import SimpleNamespace
    cases = []
    for text, nums in [("What's the content of the image?", [5]), ("<image> describe <image> both", [3, 4]),
                       ("A <image> B <image> C <image> D", [2, 10, 3]), ("question<image>", [10])]:
        mx = max(nums)
        fake = SimpleNamespace(
            image_processor=lambda images, return_tensors=None, n=nums, mx=mx: {
                "pixel_values": torch.zeros(len(n), mx, 3, 2, 2), "num_patches": torch.tensor(n)},
            tokenizer=ToyTokenizer())
        out = pr.OmChatProcessor.__call__(fake, text, images=[object()] * len(nums))
        cases.append({"text": text, "num_patches": nums, "ids": out["input_ids"][0].tolist(), "n_images": int(out["images"].shape[0])})
    return cases


if __name__ == "__main__":
    g = json.load(open(os.path.join(HERE, "golden_prompt.json")))
    g["processor"] = processor_golden()
    json.dump(g, open(os.path.join(HERE, "golden_prompt.json"), "w"), indent=0)
    print("processor cases", len(g["processor"]))
