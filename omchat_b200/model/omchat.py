"""The drop-in boundary: OmChatQwen2ForCausalLM with the reference's forward / generate surface, running on the
sm_100a kernels.

Mirrors omchat/model/language_model/omchat_qwen2.py:29-111 (forward :45-89, prepare_inputs_for_generation :92-111),
the multimodal glue of omchat/model/omchat_arch.py:43-209 (encode_images :50-53, prepare_inputs_labels_for_multimodal
:55-209) and the HF GenerationMixin greedy loop as driven by cli.py:60-70. The HF-hub twin
OmChatForConditionalGeneration (omchat/hf/modeling_omchat.py:677-689, forward :1212-1299) is the same path under the
`pixel_values`-free `images=` calling convention of hf_example.py:13-18 and is provided as a thin subclass.

Differences that are deliberate (SURVEY.md §8b "known reference bugs not to replicate"):
  * past_key_values is a PagedKVCache, not a tuple/DynamicCache; the decode-step mask extension of omchat_arch.py:61-70
    (which indexes the cache as a tuple) is replaced by the cache's own per-sequence lengths
  * compute dtype is bf16 (the reference hard-codes fp16, builder.py:28 / internVIT_encoder.py:53)
  * there is no CPU path: a CPU tensor input is moved to the model's device; without CUDA construction fails
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import torch

from .. import lib
from ..config import IGNORE_INDEX, IMAGE_TOKEN_INDEX, OmChatQwen2Config
from .decoder import PagedKVCache, Qwen2Decoder, TPInfo
from .vision import InternVITVisionTower, MMProjector, build_vision_tower
from .weights import OmChatWeights, from_state_dict, random_init


def _on_model_device(fn):
    """Run a model method with the model's GPU as the current CUDA device (kernels launch on the current device's stream;
    a model built with device="cuda:1" must work without the caller calling torch.cuda.set_device first)."""
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapped


@dataclass
class CausalLMOutputWithPast:
    """Same fields as transformers.modeling_outputs.CausalLMOutputWithPast (attribute, key and index access)."""
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[PagedKVCache] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Tuple[torch.Tensor, ...]] = None

    def to_tuple(self):
        return tuple(v for v in (self.loss, self.logits, self.past_key_values, self.hidden_states, self.attentions)
                     if v is not None)

    def __getitem__(self, k):
        return getattr(self, k) if isinstance(k, str) else self.to_tuple()[k]


class _GenerationConfig:
    def __init__(self, cfg: OmChatQwen2Config):
        self.pad_token_id = cfg.pad_token_id
        self.eos_token_id = cfg.eos_token_id
        self.max_new_tokens = 1024
        self.do_sample = False


class OmChatQwen2Model:
    """Holder object mirroring OmChatQwen2Model(OmChatMetaModel, Qwen2Model) (omchat_qwen2.py:22-26,
    omchat_arch.py:21-34): .vision_tower, .mm_projector, .embed_tokens, .layers live here."""
    decoder_class = Qwen2Decoder

    def __init__(self, config: OmChatQwen2Config, weights: OmChatWeights, tp: TPInfo):
        self.config = config
        self.vision_tower = build_vision_tower(config, weights.vit) if config.mm_vision_tower is not None else None
        self.mm_projector = MMProjector(weights.proj) if weights.proj is not None else None
        self.decoder = self.decoder_class(config, weights.llm, tp)
        self.embed_weight = weights.llm.embed

    def get_vision_tower(self):
        return self.vision_tower

    def embed_tokens(self, ids: torch.Tensor) -> torch.Tensor:
        flat = ids.reshape(-1).to(device=self.embed_weight.device, dtype=torch.int64).contiguous()
        return lib.embed_lookup(flat, self.embed_weight).view(*ids.shape, -1)


class OmChatQwen2ForCausalLM:
    config_class = OmChatQwen2Config
    model_class = OmChatQwen2Model

    def __init__(self, config: OmChatQwen2Config, weights: Optional[OmChatWeights] = None, device="cuda", seed: int = 0,
                 tp_rank: int = 0, tp_size: int = 1, tp_group=None):
        if not torch.cuda.is_available():
            raise lib.OmcError("omchat_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        lib.load()
        self.config = config
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = torch.bfloat16
        if weights is None:
            weights = random_init(config, device=device, seed=seed, tp_rank=tp_rank, tp_size=tp_size)
        self.weights = weights
        self.tp = TPInfo(rank=tp_rank, size=tp_size, group=tp_group)
        self.vision_dp = True  # under tensor parallelism: crops data-parallel over the ranks + one feature all-gather
        self.model = self.model_class(config, weights, self.tp)
        self.vocab_size = config.vocab_size
        self.generation_config = _GenerationConfig(config)

    # ------------------------------------------------------------------------------------------------ construction
    @classmethod
    def from_state_dict(cls, sd, config: OmChatQwen2Config, device="cuda", **kw):
        if not torch.cuda.is_available():
            raise lib.OmcError("omchat_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        w = from_state_dict(sd, config, device=device, tp_rank=kw.get("tp_rank", 0), tp_size=kw.get("tp_size", 1))
        return cls(config, w, device=device, **kw)

    @classmethod
    def from_pretrained(cls, path: str, config: Optional[OmChatQwen2Config] = None, device="cuda", **kw):
        """Loads *.safetensors shards under `path` (either the omchat or the HF-hub parameter names;
        builder.py:22-35 / convert_omchat_to_hf.py:26-35). fp16 checkpoints are converted to bf16."""
        from .checkpoint import load_checkpoint
        sd, cfg = load_checkpoint(path, config)
        return cls.from_state_dict(sd, cfg, device=device, **kw)

    # ------------------------------------------------------------------------------------------------ reference helpers
    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.get_model().get_vision_tower()

    def eval(self):
        return self

    def close(self):
        """Release captured graphs and peer memory (see Qwen2Decoder.release); under tensor parallelism call it on every
        rank before destroy_process_group()."""
        self.model.decoder.release()

    def to(self, *a, **k):
        return self

    @_on_model_device
    def encode_images(self, images: torch.Tensor) -> torch.Tensor:
        """omchat_arch.py:50-53: vision tower -> (pixel shuffle) -> mm_projector. [n,3,S,S] -> [n, L, hidden].

        With the decoder tensor-parallel over tp ranks (one process per GPU, every rank holding the whole tower) the crops
        are independent: rank r encodes the r-th contiguous share of them and ONE all-gather of the
        projected features [share, L, hidden] gives every rank the full, identically ordered tensor the splice needs -
        vision data-parallel -> feature all-gather -> decoder tensor-parallel. `vision_dp = False` replicates instead."""
        images = self._to_dev(images)
        n, tp = images.shape[0], self.tp.size
        if tp == 1 or not self.vision_dp or n < 2:
            feats = self.get_vision_tower()(images, self.config.pixel_shuffle_down)
            return self.get_model().mm_projector(feats)
        import torch.distributed as dist
        share = (n + tp - 1) // tp
        lo, hi = min(n, self.tp.rank * share), min(n, (self.tp.rank + 1) * share)
        L, H = self.config.image_tokens_per_crop, self.config.hidden_size
        mine = torch.zeros(share, L, H, device=self.device, dtype=torch.bfloat16)
        if hi > lo:
            feats = self.get_vision_tower()(images[lo:hi], self.config.pixel_shuffle_down)
            mine[:hi - lo] = self.get_model().mm_projector(feats)
        out = torch.empty(tp * share, L, H, device=self.device, dtype=torch.bfloat16)
        dist.all_gather_into_tensor(out, mine, group=self.tp.group)
        return out[:n]

    @_on_model_device
    def process_images(self, images, flatten: bool = True):
        """GPU replacement of omchat.mm_utils.process_images / process_anyres_image (mm_utils.py:119-182) for
        image_aspect_ratio == 'anyres' (Pillow-exact bicubic, csrc/preprocess.cu): PIL images / uint8 [H,W,3] arrays ->
        with flatten (default) ONE tensor [sum of crops, 3, 448, 448] in image order, ready to be passed as `images=`
        together with one placeholder per crop (omchat_b200.prompt.image_prompt(n_crops, text)); with flatten=False the
        reference's own return shape: [n_images, n_crops, 3, 448, 448] when every image yields the same number of crops,
        else a list of [n_crops_i, 3, 448, 448] (mm_utils.py:176-181)."""
        from ..preprocess import AnyResPreprocessor
        pre = getattr(self, "_anyres", None)
        if pre is None:
            pre = AnyResPreprocessor(self.config.image_grid_pinpoints, crop=self.config.vision_config.image_size,
                                     device=self.device, dtype=torch.bfloat16)
            self._anyres = pre
        out = pre.process_images(images if isinstance(images, (list, tuple)) else [images])
        if not flatten:
            return out
        return out.flatten(0, 1) if isinstance(out, torch.Tensor) else torch.cat(list(out), dim=0)

    def _to_dev(self, t):
        if isinstance(t, (list, tuple)):
            # list of [3,S,S] crops (omchat_arch.py:72-75) or of per-image [n_i,3,S,S] stacks (process_images, ragged)
            t = torch.cat([x if x.dim() == 4 else x[None] for x in t], dim=0)
        if t.dim() == 5:  # [n_images, n_crops, 3, S, S] as process_anyres_image stacks them
            t = t.flatten(0, 1)
        if not t.is_cuda:
            t = t.to(self.device, non_blocking=True)
        return t

    # ------------------------------------------------------------------------------------------------ glue
    def _splice_packed(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], images):
        """Device part of prepare_inputs_labels_for_multimodal (omchat_arch.py:72-164): encode images, strip padding,
        place text embeddings and image feature blocks. Returns packed (embeds [T,C], pos_ids, seq_ids, offsets list)."""
        cfg = self.config
        b, S = input_ids.shape
        ids_host = input_ids.detach().cpu()
        if attention_mask is None:
            rows = [ids_host[i] for i in range(b)]
        else:
            mh = attention_mask.detach().cpu().bool()
            rows = [ids_host[i][mh[i]] for i in range(b)]  # omchat_arch.py:115
        feats = None
        n_img, L = 0, 1
        if images is not None:
            feats = self.encode_images(images)
            n_img, L = feats.shape[0], feats.shape[1]
        if cfg.tune_mm_mlp_adapter and cfg.mm_use_im_start_end:
            raise NotImplementedError  # omchat_arch.py:100-101
        # host-side lengths (the reference syncs here too: .sum() :121, .tolist() :131)
        lens, cur = [], 0
        for r in rows:
            k = int((r == IMAGE_TOKEN_INDEX).sum())
            if images is not None:
                if cur + max(k, 1) > n_img:
                    raise IndexError("more image placeholders than images")  # image_features[cur_image_idx] :123,:150
                cur += max(k, 1)
            elif k:
                raise ValueError("input_ids contain image placeholders but no images were given")
            n = r.numel() + k * (L - 1)
            if cfg.tokenizer_model_max_length is not None:
                n = min(n, cfg.tokenizer_model_max_length)
            lens.append(n)
        packed = torch.cat(rows).to(torch.int64)
        seq_off = torch.tensor([0] + list(torch.tensor([r.numel() for r in rows]).cumsum(0)), dtype=torch.int32)
        T = sum(lens)
        cap = max(T, 1)
        embeds, pos_ids, seq_ids, _ = lib.splice(
            packed.to(self.device), seq_off.to(self.device), self.weights.llm.embed,
            feats.reshape(n_img, L, -1) if feats is not None else None, IMAGE_TOKEN_INDEX,
            cfg.tokenizer_model_max_length or 0, cap)
        offsets = [0]
        for n in lens:
            offsets.append(offsets[-1] + n)
        return embeds, pos_ids, seq_ids, offsets

    def _pad(self, packed: torch.Tensor, offsets: Sequence[int], fill=0):
        """[T, ...] packed -> [b, T_max, ...] padded on tokenizer_padding_side (omchat_arch.py:166-195)."""
        b = len(offsets) - 1
        lens = [offsets[i + 1] - offsets[i] for i in range(b)]
        Tm = max(lens)
        out = torch.full((b, Tm) + tuple(packed.shape[1:]), fill, dtype=packed.dtype, device=packed.device)
        left = self.config.tokenizer_padding_side == "left"
        for i, n in enumerate(lens):
            if n == 0:
                continue
            if left:
                out[i, Tm - n:] = packed[offsets[i]:offsets[i + 1]]
            else:
                out[i, :n] = packed[offsets[i]:offsets[i + 1]]
        return out

    @_on_model_device
    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels,
                                             images):
        """Same contract as omchat_arch.py:55-209: returns (None, position_ids, attention_mask, past_key_values,
        inputs_embeds [b,T,C], labels) — or the inputs unchanged for text-only / decode-step calls (:59-70)."""
        if self.get_vision_tower() is None or images is None or input_ids.shape[1] == 1:
            return input_ids, position_ids, attention_mask, past_key_values, None, labels
        input_ids = input_ids.to(self.device)
        embeds, pos, _, offsets = self._splice_packed(input_ids, attention_mask, images)
        b = len(offsets) - 1
        lens = [offsets[i + 1] - offsets[i] for i in range(b)]
        inputs_embeds = self._pad(embeds[:offsets[-1]], offsets)
        new_pos = self._pad(pos[:offsets[-1]].long(), offsets) if position_ids is not None else None
        new_mask = None
        if attention_mask is not None:
            ones = torch.ones(offsets[-1], dtype=attention_mask.dtype, device=self.device)
            new_mask = self._pad(ones, offsets)
        new_labels = None
        if labels is not None:
            new_labels = self._splice_labels(input_ids, attention_mask, labels, lens, offsets)
        return None, new_pos, new_mask, past_key_values, inputs_embeds, new_labels

    def _splice_labels(self, input_ids, attention_mask, labels, lens, offsets):
        """Label placement of omchat_arch.py:116-158: labels are compacted by the attention mask like the ids, image
        positions get IGNORE_INDEX, and the result is re-padded like the logits (training-only; host loop)."""
        L = self.config.image_tokens_per_crop
        ids_h, lab_h = input_ids.cpu(), labels.cpu()
        mh = attention_mask.cpu().bool() if attention_mask is not None else torch.ones_like(ids_h, dtype=torch.bool)
        rows = []
        for i in range(ids_h.shape[0]):
            out = []
            for t, lb in zip(ids_h[i][mh[i]].tolist(), lab_h[i][mh[i]].tolist()):
                out.extend([IGNORE_INDEX] * L if t == IMAGE_TOKEN_INDEX else [lb])
            rows.append(torch.tensor(out[:lens[i]], dtype=labels.dtype))
        packed = torch.cat(rows).to(self.device) if rows else labels.new_zeros(0)
        return self._pad(packed, offsets, fill=IGNORE_INDEX)

    # ------------------------------------------------------------------------------------------------ forward
    @torch.no_grad()
    @_on_model_device
    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, past_key_values: Optional[PagedKVCache] = None,
                inputs_embeds: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                use_cache: Optional[bool] = None, output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None, images: Optional[torch.Tensor] = None,
                return_dict: Optional[bool] = None, max_cache_len: Optional[int] = None,
                logits_to_keep: Union[int, str] = 0):
        """omchat_qwen2.py:45-89. Prefill: input_ids [b,S] (with -200 placeholders) + images [n,3,448,448] -> logits
        [b,T,V] fp32 padded like the reference pads (positions outside the mask are zero). Decode step: input_ids
        [b,1] + past_key_values -> logits [b,1,V]. `max_cache_len` sizes the paged cache of a prefill (default: T +
        generation_config.max_new_tokens); `logits_to_keep=1` computes only each sequence's last position."""
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised by the flash kernels")
        dec = self.model.decoder
        # ---- decode step (omchat_arch.py:59-70 short-circuit)
        if past_key_values is not None and past_key_values.get_seq_length() > 0:
            if inputs_embeds is not None or input_ids is None or input_ids.shape[1] != 1:
                raise ValueError("with a populated cache, forward() takes exactly one new token per sequence")
            logits = self._gather_vocab(dec.decode_step(input_ids.to(self.device).reshape(-1), past_key_values).clone())
            return self._output(logits[:, None, :], past_key_values, None, return_dict)
        # ---- prefill
        if inputs_embeds is None:
            if input_ids is None:
                raise ValueError("You have to specify either input_ids or inputs_embeds")
            input_ids = input_ids.to(self.device)
            embeds, pos, seq, offsets = self._splice_packed(input_ids, attention_mask, images)
            spliced_labels = None
            if labels is not None:
                # always compact labels by the mask and re-pad them like the logits (omchat_arch.py:116), images or not
                lens = [offsets[i + 1] - offsets[i] for i in range(len(offsets) - 1)]
                spliced_labels = self._splice_labels(input_ids, attention_mask, labels, lens, offsets)
        else:
            inputs_embeds = inputs_embeds.to(device=self.device, dtype=torch.bfloat16)
            b, T, _ = inputs_embeds.shape
            if attention_mask is None:
                lens = [T] * b
                embeds = inputs_embeds.reshape(b * T, -1)
            else:
                m = attention_mask.to(self.device).bool()
                lens = m.sum(1).tolist()
                embeds = inputs_embeds[m]
            offsets = [0]
            for n in lens:
                offsets.append(offsets[-1] + n)
            pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(self.device)
            seq = torch.cat([torch.full((n,), i, dtype=torch.int32) for i, n in enumerate(lens)]).to(self.device)
            embeds = embeds.contiguous().clone()
            spliced_labels = None
            if labels is not None:  # same compaction by the mask + re-padding as the logits get
                lab = labels.to(self.device)
                packed = lab.reshape(-1) if attention_mask is None else lab[attention_mask.to(self.device).bool()]
                spliced_labels = self._pad(packed, offsets, fill=IGNORE_INDEX)
        n_seq = len(offsets) - 1
        max_len = max(offsets[i + 1] - offsets[i] for i in range(n_seq))
        if past_key_values is None:
            cap = max_cache_len or (max_len + self.generation_config.max_new_tokens)
            past_key_values = dec.new_cache(n_seq, cap)
        mode = "last" if logits_to_keep == 1 else "all"
        res = dec.prefill(embeds, pos, seq, offsets, past_key_values, logits=mode, collect_hidden=bool(output_hidden_states))
        logits, hiddens = res if output_hidden_states else (res, None)
        logits = self._gather_vocab(logits)
        if mode == "last":
            logits = logits[:, None, :]
        else:
            logits = self._pad(logits, offsets)
        loss = None
        if spliced_labels is not None and mode == "all":
            # Qwen2ForCausalLM loss (modeling_qwen2.py:474-476): shifted cross-entropy; host-level torch, not a hot path
            sl = logits[:, :-1].reshape(-1, logits.shape[-1])
            tl = spliced_labels[:, 1:].reshape(-1)
            loss = torch.nn.functional.cross_entropy(sl, tl, ignore_index=IGNORE_INDEX)
        hs = None
        if hiddens is not None:
            hs = tuple(self._pad(x, offsets) for x in hiddens)
        return self._output(logits, past_key_values if use_cache is not False else None, loss, return_dict, hs)

    __call__ = forward

    def _gather_vocab(self, logits: torch.Tensor) -> torch.Tensor:
        """Under tensor parallelism lm_head is vocab-parallel: every rank holds [rows, V / tp]. forward() returns FULL
        logits like the reference does, so the shards are all-gathered along the vocabulary (rank order = vocab order)."""
        if self.tp.size == 1:
            return logits
        import torch.distributed as dist
        flat = logits.reshape(-1, logits.shape[-1]).contiguous()
        parts = torch.empty(self.tp.size, *flat.shape, device=flat.device, dtype=flat.dtype)
        dist.all_gather_into_tensor(parts, flat, group=self.tp.group)
        return parts.permute(1, 0, 2).reshape(*logits.shape[:-1], self.tp.size * logits.shape[-1])

    @staticmethod
    def _output(logits, cache, loss, return_dict, hidden_states=None):
        out = CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=cache, hidden_states=hidden_states)
        return out if return_dict is not False else out.to_tuple()

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None, inputs_embeds=None,
                                      **kwargs):
        """omchat_qwen2.py:92-111."""
        if past_key_values:
            input_ids = input_ids[:, -1:]
        if inputs_embeds is not None and past_key_values is None:
            model_inputs = {"inputs_embeds": inputs_embeds}
        else:
            model_inputs = {"input_ids": input_ids}
        model_inputs.update({"past_key_values": past_key_values, "use_cache": kwargs.get("use_cache"),
                             "attention_mask": attention_mask, "images": kwargs.get("images", None)})
        return model_inputs

    # ------------------------------------------------------------------------------------------------ generate
    @torch.no_grad()
    @_on_model_device
    def generate(self, input_ids: Optional[torch.Tensor] = None, images: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, max_new_tokens: Optional[int] = None,
                 do_sample: bool = False, temperature: Optional[float] = None, eos_token_id=None, pad_token_id=None,
                 use_cache: bool = True, streamer=None, inputs: Optional[torch.Tensor] = None, use_graph: bool = True,
                 stopping_criteria=None, **unused):
        """Greedy generation with the reference's call shape (cli.py:60-70, hf_example.py:13-18):
        returns LongTensor [b, S + new] = the prompt ids followed by the generated ids (pad_token_id after a row has
        finished). A row finishes at its first EOS token or when `stopping_criteria` (a callable or a list of callables
        `crit(ids [b, len], scores) -> bool | BoolTensor[b]`, e.g. prompt.KeywordsStoppingCriteria, mm_utils.py:242-274;
        HF StoppingCriteriaList semantics: any criterion firing stops) says so; generation ends when every row has."""
        if input_ids is None:
            input_ids = inputs
        if input_ids is None:
            raise ValueError("generate() needs input_ids")
        if do_sample:
            raise NotImplementedError("only greedy decoding (do_sample=False) is on the accelerated path")
        gc = self.generation_config
        max_new = max_new_tokens if max_new_tokens is not None else gc.max_new_tokens
        eos = gc.eos_token_id if eos_token_id is None else eos_token_id
        eos_set = set(eos) if isinstance(eos, (list, tuple)) else ({eos} if eos is not None and eos >= 0 else set())
        pad = pad_token_id if pad_token_id is not None else (gc.pad_token_id if gc.pad_token_id is not None else 0)
        if stopping_criteria is None:
            criteria = []
        elif callable(stopping_criteria):
            criteria = [stopping_criteria]
        else:
            criteria = list(stopping_criteria)
        dec = self.model.decoder
        input_ids = input_ids.to(self.device)
        b = input_ids.shape[0]
        if streamer is not None:
            streamer.put(input_ids.cpu())
        if max_new <= 0:
            return input_ids
        embeds, pos, seq, offsets = self._splice_packed(input_ids, attention_mask, images)
        max_len = max(offsets[i + 1] - offsets[i] for i in range(b))
        cache = dec.acquire_cache(b, max_len + max_new)
        logits = dec.prefill(embeds, pos, seq, offsets, cache, logits="last")
        first = torch.empty(b, device=self.device, dtype=torch.int64)
        if self.tp.size == 1:
            lib.argmax(logits, out=first)
        else:
            st = dec._decode_state(b, cache.capacity)
            st.logits.copy_(logits)
            dec._greedy(st)
            first.copy_(st.tokens)
        hook = None
        if streamer is not None:
            streamer.put(first.cpu())
            hook = lambda i, toks: streamer.put(toks.cpu())  # noqa: E731
        # Decode in chunks; EOS and the stopping criteria are evaluated on the host once per chunk, token by token, and
        # whatever was generated past a row's end is discarded -> the same ids as a per-token loop, without a sync per token.
        check_stop = bool(eos_set or criteria)
        prompt_host = input_ids.cpu() if criteria else None
        new_host = torch.empty(b, 0, dtype=torch.int64)
        ends: List[Optional[int]] = [None] * b  # number of generated tokens row i keeps, once it has finished

        def scan(from_col: int):
            """Mark rows finished by the tokens new_host[:, from_col:]. Returns True when every row has finished."""
            for j in range(from_col, new_host.shape[1]):
                live = [i for i in range(b) if ends[i] is None]
                if not live:
                    break
                for i in live:
                    if int(new_host[i, j]) in eos_set:
                        ends[i] = j + 1
                if criteria:
                    prefix = torch.cat([prompt_host, new_host[:, :j + 1]], dim=1)
                    for crit in criteria:
                        r = crit(prefix, None)
                        hit = [bool(r)] * b if not isinstance(r, torch.Tensor) or r.dim() == 0 else [bool(x) for x in r.tolist()]
                        for i in live:
                            if hit[i] and ends[i] is None:
                                ends[i] = j + 1
            return all(e is not None for e in ends)

        chunks = [first.view(b, 1)]
        produced, cur, done = 1, first, False
        if check_stop:
            new_host = first.view(b, 1).cpu()
            done = scan(0)
        chunk = 1 if streamer is not None else (8 if criteria else 32)
        while produced < max_new and not done:
            n = min(chunk, max_new - produced)
            toks = dec.generate_greedy(cur, cache, n, use_graph=use_graph, on_token=hook)
            chunks.append(toks)
            cur = toks[:, -1].contiguous()
            if check_stop:
                new_host = torch.cat([new_host, toks.cpu()], dim=1)
                done = scan(produced)
            produced += n
        new = torch.cat(chunks, dim=1)
        if check_stop and any(e is not None for e in ends):
            keep = max(e if e is not None else new.shape[1] for e in ends)
            nh = new_host[:, :keep].clone()
            for i, e in enumerate(ends):
                if e is not None:
                    nh[i, e:] = pad
            new = nh.to(self.device)
        if streamer is not None:
            streamer.end()
        return torch.cat([input_ids, new], dim=1)


class OmChatForConditionalGeneration(OmChatQwen2ForCausalLM):
    """HF-hub twin (omchat/hf/modeling_omchat.py:677-689): same forward path (:1212-1299 = splice :769-923 +
    language_model), constructed from the hub-layout state dict (vision_tower.* / multi_modal_projector.linear_{1,2}.* /
    language_model.*, convert_omchat_to_hf.py:26-35). `generate(**processor(text, images))` passes `images=`."""

    def forward(self, input_ids=None, images=None, pixel_values=None, **kw):
        if images is None and pixel_values is not None:
            images = pixel_values
        return super().forward(input_ids=input_ids, images=images, **kw)

    __call__ = forward
