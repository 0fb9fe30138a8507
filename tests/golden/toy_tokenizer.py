"""A deterministic byte-level toy tokenizer with the surface the reference's prompt helpers use (__call__().input_ids,
encode, batch_decode, bos_token_id). Special strings <|im_start|> / <|im_end|> map to the Qwen2 ids 151644 / 151645."""
from types import SimpleNamespace

SPECIALS = {"<|im_start|>": 151644, "<|im_end|>": 151645}


class ToyTokenizer:
    def __init__(self, bos_token_id=None):
        self.bos_token_id = bos_token_id

    def encode_nobos(self, text):
        ids, i = [], 0
        while i < len(text):
            for s, sid in SPECIALS.items():
                if text.startswith(s, i):
                    ids.append(sid)
                    i += len(s)
                    break
            else:
                ids.extend(10 + b for b in text[i].encode("utf-8"))
                i += 1
        return ids

    def encode(self, text):
        ids = self.encode_nobos(text)
        return ([self.bos_token_id] if self.bos_token_id is not None else []) + ids

    def __call__(self, text):
        return SimpleNamespace(input_ids=self.encode(text))

    def batch_decode(self, batch, skip_special_tokens=True):
        outs = []
        inv = {v: k for k, v in SPECIALS.items()}
        for row in batch:
            bs, s = bytearray(), ""
            for t in (row.tolist() if hasattr(row, "tolist") else row):
                if t in inv:
                    s += bs.decode("utf-8", errors="replace") + ("" if skip_special_tokens else inv[t])
                    bs = bytearray()
                elif t == self.bos_token_id:
                    continue
                elif 10 <= t < 266:
                    bs.append(t - 10)
            outs.append(s + bs.decode("utf-8", errors="replace"))
        return outs
