"""Prompt / token contract around the path — host-side integer logic mirroring the reference:

  tokenizer_image_token      omchat/mm_utils.py:197-230   text with <image> / <image_N> tags -> ids with -200 placeholders
  make_context               omchat/make_context.py:66-148 ChatML context (system, bounded history, query) -> (text, ids)
  image_prompt               omchat/make_context.py:25-30,57-62 the "<image>\\npatch:<image>..." prefix for n crops
  KeywordsStoppingCriteria   omchat/mm_utils.py:242-274

Pure Python over a tokenizer object (anything with __call__(text).input_ids, encode(text), batch_decode, bos_token_id);
no tensors on the hot path, results are compared id-for-id with the reference functions in tests/test_prompt.py.
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple

import torch

from .config import IMAGE_TOKEN_INDEX

DEFAULT_IMAGE_TOKEN = "<image>"  # omchat/constants.py:9
IM_START_ID, IM_END_ID = 151644, 151645  # make_context.py:79-80 (Qwen2 <|im_start|>, <|im_end|>)


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX, return_tensors: Optional[str] = None):
    """mm_utils.py:197-230. Numbered tags (<image_0>, <image_1>, ...) each become one placeholder; otherwise the prompt is
    split at <image>, a leading BOS is kept once and the separator is the placeholder id."""
    if "<image_0>" in prompt:
        chunks = re.split(r"<image_[0-9]+>", prompt)
        tags = re.findall(r"<image_(\d+)>", prompt)
        input_ids: List[int] = []
        for i, chunk in enumerate(chunks):
            input_ids.extend(tokenizer(chunk).input_ids)
            if i < len(tags):
                input_ids.append(-200)
    else:
        chunks = [tokenizer(chunk).input_ids for chunk in prompt.split(DEFAULT_IMAGE_TOKEN)]
        input_ids = []
        offset = 0
        if len(chunks) > 0 and len(chunks[0]) > 0 and chunks[0][0] == tokenizer.bos_token_id:
            offset = 1
            input_ids.append(chunks[0][0])
        sep = [image_token_index] * (offset + 1)
        interleaved = [e for pair in zip(chunks, [sep] * len(chunks)) for e in pair][:-1]
        for x in interleaved:
            input_ids.extend(x[offset:])
    if return_tensors is not None:
        if return_tensors == "pt":
            return torch.tensor(input_ids, dtype=torch.long)
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return input_ids


def image_prompt(n_crops: int, text: str) -> str:
    """make_context.py:27,59: one <image> for the whole-image crop, 'patch:<image>' per canvas patch, then the question."""
    return "<image>\n" + "\n".join(["patch:<image>"] * (n_crops - 1)) + "\n" + text.replace("<image>", "").strip()


def make_context(tokenizer, query: str, history: Optional[Sequence[Tuple[str, str]]] = None, system: str = "",
                 max_window_size: int = 6144, chat_format: str = "chatml"):
    """make_context.py:66-148 -> (raw_text, context_tokens)."""
    history = [] if history is None else history
    if chat_format == "raw":
        return query, tokenizer.encode(query)
    if chat_format != "chatml":
        raise NotImplementedError(f"Unknown chat format {chat_format!r}")
    im_start, im_end = "<|im_start|>", "<|im_end|>"
    nl_tokens = tokenizer.encode("\n")

    def tok(role: str, content: str):
        if DEFAULT_IMAGE_TOKEN in content:
            body = tokenizer_image_token(content, tokenizer, IMAGE_TOKEN_INDEX)
        else:
            body = tokenizer.encode(content)
        return f"{role}\n{content}", tokenizer.encode(role) + nl_tokens + body

    system_text, system_part = tok("system", system)
    system_tokens = [IM_START_ID] + system_part + [IM_END_ID]
    raw_text, context_tokens = "", []
    for turn_query, turn_response in reversed(history):
        q_text, q_part = tok("user", turn_query)
        r_text, r_part = tok("assistant", turn_response)
        nxt = nl_tokens + [IM_START_ID] + q_part + [IM_END_ID] + nl_tokens + [IM_START_ID] + r_part + [IM_END_ID]
        if len(system_tokens) + len(nxt) + len(context_tokens) < max_window_size:
            context_tokens = nxt + context_tokens
            raw_text = f"\n{im_start}{q_text}{im_end}\n{im_start}{r_text}{im_end}" + raw_text
        else:
            break
    context_tokens = system_tokens + context_tokens
    raw_text = f"{im_start}{system_text}{im_end}" + raw_text
    context_tokens += (nl_tokens + [IM_START_ID] + tok("user", query)[1] + [IM_END_ID] + nl_tokens + [IM_START_ID]
                       + tokenizer.encode("assistant") + nl_tokens)
    raw_text += f"\n{im_start}user\n{query}{im_end}\n{im_start}assistant\n"
    return raw_text, context_tokens


class KeywordsStoppingCriteria:
    """mm_utils.py:242-274: stop when every sequence ends with (or its decoded tail contains) one of the keywords."""

    def __init__(self, keywords: Sequence[str], tokenizer, input_ids: torch.Tensor):
        self.keywords = list(keywords)
        self.keyword_ids = []
        self.max_keyword_len = 0
        for keyword in self.keywords:
            ids = tokenizer(keyword).input_ids
            if len(ids) > 1 and ids[0] == tokenizer.bos_token_id:
                ids = ids[1:]
            self.max_keyword_len = max(self.max_keyword_len, len(ids))
            self.keyword_ids.append(torch.tensor(ids))
        self.tokenizer = tokenizer
        self.start_len = input_ids.shape[1]

    def call_for_batch(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        offset = min(output_ids.shape[1] - self.start_len, self.max_keyword_len)
        self.keyword_ids = [k.to(output_ids.device) for k in self.keyword_ids]
        for k in self.keyword_ids:
            if torch.equal(output_ids[0, -k.shape[0]:], k):
                return True
        outputs = self.tokenizer.batch_decode(output_ids[:, -offset:], skip_special_tokens=True)[0]
        return any(keyword in outputs for keyword in self.keywords)

    def __call__(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        return all(self.call_for_batch(output_ids[i].unsqueeze(0), scores) for i in range(output_ids.shape[0]))
