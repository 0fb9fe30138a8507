"""Continuous batching over the paged KV cache (the reference has no serving loop - cli.py serves one
request at a time through generate()).

A fixed number of decode SLOTS share one page pool: up to 4 slots decode with the persistent kernel (one launch per token
for all slots), 5..64 slots with the batched step on the weight-streaming GEMMs (csrc/gemm_stream.cu; one CUDA graph
replay per token). A request is admitted into a free slot as soon as one exists: its prompt is spliced (image features in place of the
-200 placeholders, omchat_arch.py:115-195) and prefilled into pages taken from a free list, its first token is sampled from
the prefill logits, and from the next step on it decodes together with whatever the other slots hold. A finished request
(EOS or max_new_tokens) gives its pages back and frees its slot at the next chunk boundary. Idle slots point at one
scratch page and are reset every chunk, so the kernel always runs the same batch shape.

Every request produces exactly the tokens a stand-alone greedy generate() of the same prompt produces (up to bf16-level
near-ties: the split of the keys over CTAs depends on the batch size) - tests/test_serving_gpu.py checks that against the
oracle.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import lib
from .model.decoder import PagedKVCache

MAX_SLOTS = 64  # rows of the batched decode step (csrc/gemm_stream.cu)


@dataclass
class _Request:
    rid: int
    input_ids: torch.Tensor          # [1, S] int64 with -200 placeholders
    images: Optional[torch.Tensor]   # [n, 3, S, S] or None
    max_new: int
    eos: Optional[int]
    out: List[int] = field(default_factory=list)
    pages: List[int] = field(default_factory=list)
    done: bool = False


class ContinuousBatcher:
    """model: OmChatQwen2ForCausalLM (single GPU). slots: decode batch (1..64). max_ctx: longest prompt + generation a slot
    can hold. total_pages: size of the shared page pool (default: enough for every slot at max_ctx). chunk: decode steps
    between two looks at the host side (admission, EOS, release)."""

    def __init__(self, model, slots: int = 4, max_ctx: int = 4096, total_pages: Optional[int] = None, chunk: int = 8):
        if not 1 <= slots <= MAX_SLOTS:
            raise ValueError(f"slots must be 1..{MAX_SLOTS}")
        self.model, self.dec = model, model.model.decoder
        self.slots, self.chunk = slots, chunk
        ps = model.config.kv_page_size
        # one page more per slot than max_ctx needs: the pool then holds every slot at full length PLUS the scratch page
        self.cache: PagedKVCache = self.dec.new_cache(slots, max_ctx + ps, shuffle_pages=False)
        self.max_pages = self.cache.max_pages
        n_pool = self.cache.pool.shape[1]
        total = n_pool if total_pages is None else min(total_pages, n_pool)
        if total < 2:
            raise ValueError("the page pool needs at least a scratch page and one data page")
        # page 0 = scratch page of the idle slots; the others are handed out from a free list (LIFO: freed pages are
        # reused first, so block tables become non-contiguous as soon as requests of different lengths come and go)
        self.scratch = 0
        self.free_pages: List[int] = list(range(total - 1, 0, -1))
        self.page_size = ps
        self.table_host = torch.zeros(slots, self.max_pages, dtype=torch.int32)
        self.cache.block_table.copy_(self.table_host)
        self.cache.host_lens = [0] * slots
        self.cache.ctx_lens.zero_()
        self.active: List[Optional[_Request]] = [None] * slots
        self.cur = torch.zeros(slots, device=model.device, dtype=torch.int64)  # token each slot feeds next
        self.queue: deque = deque()
        self.results: Dict[int, torch.Tensor] = {}
        self._next_id = 0
        self.steps_run = 0

    # ------------------------------------------------------------------------------------------------ public
    def submit(self, input_ids: torch.Tensor, images: Optional[torch.Tensor] = None, max_new_tokens: int = 64,
               eos_token_id: Optional[int] = None) -> int:
        """Queue one request ([1, S] or [S] ids with -200 where an image goes). Returns its id."""
        ids = input_ids.view(1, -1).to(torch.int64).cpu()
        if max_new_tokens < 1:
            raise ValueError("max_new_tokens must be >= 1")
        rid = self._next_id
        self._next_id += 1
        eos = eos_token_id if eos_token_id is not None and eos_token_id >= 0 else None
        self.queue.append(_Request(rid, ids, images, max_new_tokens, eos))
        return rid

    def run(self) -> Dict[int, torch.Tensor]:
        """Serve until the queue and the slots are empty. Returns {request id: generated ids (LongTensor)}."""
        while self.queue or any(r is not None for r in self.active):
            self.step()
        return self.results

    @torch.no_grad()
    def step(self):
        """Admit what fits, then one chunk of decode steps for all slots, then retire what finished."""
        self._admit()
        live = [r for r in self.active if r is not None]
        if not live:
            if self.queue:
                raise RuntimeError("a queued request does not fit the page pool even when it is empty")
            return
        n = min(self.chunk, min(r.max_new - len(r.out) for r in live))
        if n > 0:
            toks = self.dec.generate_greedy(self.cur, self.cache, n).cpu()  # [slots, n]
            self.steps_run += n
            self.cur = toks[:, -1].to(self.model.device)
            for s, r in enumerate(self.active):
                if r is None:
                    continue
                for t in toks[s].tolist():
                    r.out.append(t)
                    if (r.eos is not None and t == r.eos) or len(r.out) >= r.max_new:
                        r.done = True
                        break
        self._retire()

    # ------------------------------------------------------------------------------------------------ internals
    def _admit(self):
        for s in range(self.slots):
            if self.active[s] is not None or not self.queue:
                continue
            r: _Request = self.queue[0]
            embeds, pos, seq, offsets = self.model._splice_packed(r.input_ids.to(self.model.device), None, r.images)
            T = offsets[-1]
            need = (T + r.max_new + self.page_size - 1) // self.page_size
            if need > self.max_pages:
                raise ValueError(f"request {r.rid}: {T} prompt + {r.max_new} new tokens exceed the slot capacity")
            if need > len(self.free_pages):
                break  # wait for pages (FIFO admission: later requests do not overtake)
            self.queue.popleft()
            r.pages = [self.free_pages.pop() for _ in range(need)]
            self.table_host[s].fill_(self.scratch)
            self.table_host[s, :need] = torch.tensor(r.pages, dtype=torch.int32)
            self.cache.block_table.copy_(self.table_host)
            seq.fill_(s)  # the packed rows of this request belong to cache row s
            logits = self.dec.prefill(embeds, pos, seq, offsets, self.cache, logits="last", slots=[s])
            first = lib.argmax(logits)
            t0 = int(first[0])
            r.out.append(t0)
            if (r.eos is not None and t0 == r.eos) or r.max_new == 1:
                r.done = True
            self.cur[s] = first[0]
            self.active[s] = r
        self._retire()

    def _retire(self):
        changed = False
        for s, r in enumerate(self.active):
            if r is not None and r.done:
                self.results[r.rid] = torch.tensor(r.out, dtype=torch.int64)
                self.free_pages.extend(reversed(r.pages))
                r.pages = []
                self.active[s] = None
                self.table_host[s].fill_(self.scratch)
                changed = True
        # idle slots: scratch page, context 0 (the kernel still runs them: it appends to the scratch page and its
        # samples are ignored); live slots keep their lengths
        for s, r in enumerate(self.active):
            if r is None and self.cache.host_lens[s] != 0:
                self.cache.host_lens[s] = 0
                changed = True
        if changed:
            self.cache.block_table.copy_(self.table_host)
            self.cache.ctx_lens.copy_(torch.tensor(self.cache.host_lens, dtype=torch.int32))
            self.cur = torch.where(torch.tensor([r is not None for r in self.active], device=self.cur.device), self.cur,
                                   torch.zeros_like(self.cur))
