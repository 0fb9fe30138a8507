"""Prompt / token contract on either side of the path: what turns chat text into the `input_ids`
(with -200 placeholders) the model consumes, and what decides when generation stops. Host-side integer logic; the ids
it produces are checked one for one against the reference's own functions (tests/golden/make_golden_prompt.py ->
tests/test_prompt.py), the text of this module is not theirs:

  tokenizer_image_token(prompt, tok)     contract of omchat/mm_utils.py:197-230
  image_prompt(n_crops, text)            the crop prefix built at omchat/make_context.py:25-30,57-62
  make_context(tok, query, history, …)   contract of omchat/make_context.py:66-148 (ChatML, bounded history window)
  KeywordsStoppingCriteria               contract of omchat/mm_utils.py:242-274 (usable as generate(stopping_criteria=[…]))

A tokenizer is anything with __call__(text).input_ids, encode(text), batch_decode(ids, skip_special_tokens) and
bos_token_id.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

from .config import IMAGE_TOKEN_INDEX

DEFAULT_IMAGE_TOKEN = "<image>"  # omchat/constants.py:9
IM_START, IM_END = "<|im_start|>", "<|im_end|>"
IM_START_ID, IM_END_ID = 151644, 151645  # Qwen2 ids of the two ChatML markers (make_context.py:79-80)

_PLAIN_TAG = re.compile(re.escape(DEFAULT_IMAGE_TOKEN))
_NUMBERED_TAG = re.compile(r"<image_[0-9]+>")


def _join(parts: Iterable[List[int]], sep: int) -> List[int]:
    out: List[int] = []
    for i, p in enumerate(parts):
        if i:
            out.append(sep)
        out.extend(p)
    return out


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX,
                          return_tensors: Optional[str] = None):
    """Text with image tags -> ids with one placeholder per tag.

    Two tag dialects: numbered (`<image_0>`, `<image_1>`, …; recognised by the presence of `<image_0>`) where every text
    segment is tokenised as it stands and the separator is the constant -200, and plain `<image>` where a tokenizer that
    prepends BOS to every segment contributes that BOS exactly once, at the front."""
    numbered = "<image_0>" in prompt
    segments = [tokenizer(s).input_ids for s in (_NUMBERED_TAG if numbered else _PLAIN_TAG).split(prompt)]
    if numbered:
        ids = _join(segments, IMAGE_TOKEN_INDEX)
    else:
        has_bos = bool(segments[0]) and segments[0][0] == tokenizer.bos_token_id
        lead = 1 if has_bos else 0
        ids = segments[0][:lead] + _join((s[lead:] for s in segments), image_token_index)
    if return_tensors is None:
        return ids
    if return_tensors != "pt":
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return torch.tensor(ids, dtype=torch.long)


def image_prompt(n_crops: int, text: str) -> str:
    """The user text for an any-res image of n_crops crops: `<image>` for the overview crop, `patch:<image>` for every
    canvas patch, one per line, then the question with stray tags removed."""
    lines = [DEFAULT_IMAGE_TOKEN] + ["patch:" + DEFAULT_IMAGE_TOKEN] * (n_crops - 1)
    return lines[0] + "\n" + "\n".join(lines[1:]) + "\n" + text.replace(DEFAULT_IMAGE_TOKEN, "").strip()


@dataclass
class _Rendered:
    text: str
    ids: List[int]

    def __add__(self, other: "_Rendered") -> "_Rendered":
        return _Rendered(self.text + other.text, self.ids + other.ids)


class _ChatML:
    """Renders ChatML pieces to (text, ids) in lock step. A message is <|im_start|>{role}\\n{content}<|im_end|>; the two
    markers are single ids, role and content are tokenised separately, content with image tags through
    tokenizer_image_token."""

    def __init__(self, tokenizer):
        self.tok = tokenizer
        self.newline = _Rendered("\n", list(tokenizer.encode("\n")))

    def _content(self, content: str) -> List[int]:
        if DEFAULT_IMAGE_TOKEN in content:
            return tokenizer_image_token(content, self.tok, IMAGE_TOKEN_INDEX)
        return list(self.tok.encode(content))

    def header(self, role: str) -> _Rendered:
        return _Rendered(IM_START + role, [IM_START_ID] + list(self.tok.encode(role))) + self.newline

    def message(self, role: str, content: str) -> _Rendered:
        return self.header(role) + _Rendered(content + IM_END, self._content(content) + [IM_END_ID])

    def turn(self, question: str, answer: str) -> _Rendered:
        return self.newline + self.message("user", question) + self.newline + self.message("assistant", answer)


def make_context(tokenizer, query: str, history: Optional[Sequence[Tuple[str, str]]] = None, system: str = "",
                 max_window_size: int = 6144, chat_format: str = "chatml"):
    """(raw_text, context_tokens) of a chat request: system message, as many of the MOST RECENT history turns as keep
    system + history strictly under max_window_size tokens (an older turn that does not fit ends the walk), the new user
    message and an open assistant header. chat_format "raw" passes the query through untouched."""
    if chat_format == "raw":
        return query, tokenizer.encode(query)
    if chat_format != "chatml":
        raise NotImplementedError(f"Unknown chat format {chat_format!r}")
    ml = _ChatML(tokenizer)
    head = ml.message("system", system)
    kept: List[_Rendered] = []
    used = len(head.ids)
    for question, answer in reversed(list(history or [])):
        t = ml.turn(question, answer)
        if used + len(t.ids) >= max_window_size:
            break
        kept.append(t)
        used += len(t.ids)
    ctx = head
    for t in reversed(kept):
        ctx = ctx + t
    ctx = ctx + ml.newline + ml.message("user", query) + ml.newline + ml.header("assistant")
    return ctx.text, ctx.ids


class KeywordsStoppingCriteria:
    """Stop once EVERY row of the batch has produced one of the keywords: either the row ends with a keyword's id
    sequence, or the decoded text of its last few generated tokens (as many as the longest keyword has ids; before
    anything was generated: the whole row) contains a keyword. Callable like a transformers StoppingCriteria:
    crit(output_ids [b, len], scores) -> bool."""

    def __init__(self, keywords: Sequence[str], tokenizer, input_ids: torch.Tensor):
        self.keywords = list(keywords)
        self.tokenizer = tokenizer
        self.start_len = int(input_ids.shape[1])
        self.keyword_ids: List[torch.Tensor] = []
        for kw in self.keywords:
            ids = list(tokenizer(kw).input_ids)
            if len(ids) > 1 and ids[0] == tokenizer.bos_token_id:
                del ids[0]
            self.keyword_ids.append(torch.tensor(ids))
        self.max_keyword_len = max((int(k.numel()) for k in self.keyword_ids), default=0)

    def _row_hit(self, row: torch.Tensor) -> bool:
        n = int(row.numel())
        for k in self.keyword_ids:
            m = int(k.numel())
            if 0 < m <= n and row[n - m:].tolist() == k.tolist():
                return True
        window = min(n - self.start_len, self.max_keyword_len)
        tail = row[n - window:] if window > 0 else row
        text = self.tokenizer.batch_decode(tail[None], skip_special_tokens=True)[0]
        return any(kw in text for kw in self.keywords)

    def call_for_batch(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        return self._row_hit(output_ids[0])

    def __call__(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        rows = output_ids.detach().cpu() if output_ids.is_cuda else output_ids
        return all(self._row_hit(rows[i]) for i in range(rows.shape[0]))
