"""Host-side engine: owns the C handle, feeds it the reference state dict, turns a right-padded batch of `input_ids`
into the packed index metadata the kernels need (on the host, from the CPU copy of the ids — the reference's device
syncs `nonzero`, `ids.tolist()`, `0 in attention_mask` disappear), and launches `mjv_score_forward`.

Index semantics are the reference's:
  last token  : ((ids == pad).int().argmax(-1) - 1) % N          moe_reward.py:226-228
  gating token: start of the LAST occurrence of token_pattern     moe_reward.py:45-57  (ValueError if absent)
  image rows  : row-major positions of <IMG_CONTEXT> receive the projector rows in order
                                                                 modeling_internvl_chat.py:176-188
"""
from __future__ import annotations

import ctypes as C
import math
import warnings
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

TOKEN_PATTERN = [92542, 92543, 525, 11353, 364]  # '<|im_end|><|im_start|>assistant\n' (moe_reward.py:48)


def pack_work_entries(blocks, pair_blocks: bool) -> np.ndarray:
    """int32 [n, 2] attention work list from (sequence, query_block) tuples, heaviest entry first (a causal block q reads
    q+1 KV blocks). `pair_blocks`: two blocks of one sequence share an entry - low 16 bits of the second field hold the
    larger block, the high 16 bits the other block + 1 (0 = none) - so one CTA streams their common K/V once
    (csrc/attn2_sm100.cuh). Without it the second field is the plain block index."""
    if not pair_blocks:
        out = sorted(blocks, key=lambda t: -t[1])
    else:
        per_seq: dict[int, list[int]] = {}
        for s, q in blocks:
            per_seq.setdefault(s, []).append(q)
        out = []
        for s, qs in per_seq.items():
            qs = sorted(set(qs), reverse=True)
            if qs and qs[0] >= 0xFFFF:
                raise ValueError("sequence too long for a packed attention work entry")
            for i in range(0, len(qs), 2):
                hi = qs[i]
                lo = qs[i + 1] + 1 if i + 1 < len(qs) else 0
                out.append((s, hi | (lo << 16)))
        out.sort(key=lambda t: -(t[1] & 0xFFFF))
    return np.asarray(out, dtype=np.int32).reshape(-1, 2)


def find_token_for_gating(lst, token_pattern=None) -> int:
    """Index of the last occurrence of `token_pattern` in `lst` (moe_reward.py:50-57)."""
    pat = TOKEN_PATTERN if token_pattern is None else list(token_pattern)
    n = len(pat)
    arr = np.asarray(lst)
    if arr.size >= n:
        hit = np.ones(arr.size - n + 1, dtype=bool)
        for i, p in enumerate(pat):
            hit &= arr[i:arr.size - n + 1 + i] == p
        idx = np.flatnonzero(hit)
        if idx.size:
            return int(idx[-1])
    raise ValueError("Token pattern not found in the list.")


@dataclass
class EngineDims:
    image_size: int
    patch_size: int
    vit_hidden: int
    vit_heads: int
    vit_mlp: int
    vit_layers: int
    vit_eps: float
    llm_arch: int
    llm_hidden: int
    llm_inter: int
    llm_layers: int
    llm_heads: int
    llm_kv_heads: int
    vocab: int
    rms_eps: float
    rope_theta: float
    max_pos: int
    num_objectives: int
    num_aspects: int
    gating_hidden: int
    gating_n_hidden: int
    gating_temperature: float
    aspect2criteria: dict
    downsample_ratio: float = 0.5
    mlp1_eps: float = 1e-5
    rope_scaling_factor: float | None = 2.0   # InternLM2 rope_scaling {'type': 'dynamic', 'factor': f}; None = no scaling
    ps_version: int = 2                        # 1 = 'v1' pixel-shuffle layout (transposed token grid)
    rope_scaling_type: str = "dynamic"         # 'dynamic' (NTK re-basing beyond max_pos) or 'linear' (positions / factor)
    # Phi-3 LongRoPE (rope_scaling {'type': 'su' | 'longrope', 'short_factor': [...], 'long_factor': [...]}; Llama-layout
    # branch only): per-frequency divisors, the long set once a batch is longer than the original context
    rope_short_factor: tuple | None = None
    rope_long_factor: tuple | None = None
    rope_original_max_pos: int | None = None

    @staticmethod
    def from_synth(d) -> "EngineDims":
        """From mj_video_b200.synth.ModelDims."""
        return EngineDims(
            image_size=d.image_size, patch_size=d.patch_size, vit_hidden=d.vit_hidden, vit_heads=d.vit_heads,
            vit_mlp=d.vit_mlp, vit_layers=applied_vit_layers(d.vit_layers, getattr(d, "select_layer", -1)), vit_eps=d.vit_eps,
            llm_arch=1 if getattr(d, "llm_arch", "") == "LlamaForCausalLM" else 0, llm_hidden=d.llm_hidden,
            llm_inter=d.llm_inter, llm_layers=d.llm_layers, llm_heads=d.llm_heads, llm_kv_heads=d.llm_kv_heads,
            vocab=d.vocab, rms_eps=d.rms_eps, rope_theta=d.rope_theta, max_pos=d.max_pos,
            num_objectives=d.num_objectives, num_aspects=d.num_aspects, gating_hidden=d.gating_hidden,
            gating_n_hidden=d.gating_n_hidden, gating_temperature=d.gating_temperature,
            aspect2criteria={k: list(v) for k, v in d.aspect2criteria.items()}, downsample_ratio=d.downsample_ratio,
            rope_scaling_factor=getattr(d, "rope_scaling_factor", 2.0),
            ps_version=1 if getattr(d, "ps_version", "v2") == "v1" else 2,
            rope_scaling_type=getattr(d, "rope_scaling_type", "dynamic"),
            rope_short_factor=getattr(d, "rope_short_factor", None), rope_long_factor=getattr(d, "rope_long_factor", None),
            rope_original_max_pos=getattr(d, "rope_original_max_pos", None))

    @property
    def longrope(self) -> bool:
        return self.llm_arch == 1 and self.rope_scaling_type in ("su", "longrope") and self.rope_short_factor is not None

    @property
    def vit_seq(self) -> int:
        return (self.image_size // self.patch_size) ** 2 + 1

    @property
    def tokens_per_tile(self) -> int:
        return int((self.image_size // self.patch_size) ** 2 * self.downsample_ratio ** 2)


def applied_vit_layers(n_layers: int, select_layer: int) -> int:
    """Encoder layers that produce the hidden state `extract_feature` reads (modeling_internvl_chat.py:244-254):
    select_layer == -1 -> last_hidden_state (all layers); else hidden_states[select_layer] of the n_layers + 1 states."""
    if select_layer == -1:
        return n_layers
    k = n_layers + 1 + select_layer if select_layer < 0 else select_layer
    if not 0 < k <= n_layers:
        raise ValueError(f"select_layer={select_layer} is outside the {n_layers}-layer vision encoder")
    return k


def longrope_inv_freq(d: "EngineDims", hd: int, seq_len: int):
    """Phi-3 LongRoPE (transformers `_compute_longrope_parameters` / Phi3LongRoPEScaledRotaryEmbedding): inverse frequencies
    divided by `long_factor` when the call is longer than the original context, else by `short_factor`; cos/sin scaled by
    sqrt(1 + ln(max_pos / original) / ln(original)) when the context was extended. Returns (inv_freq fp32, scale, long?)."""
    orig = int(d.rope_original_max_pos or d.max_pos)
    use_long = seq_len > orig
    ext = torch.tensor(d.rope_long_factor if use_long else d.rope_short_factor, dtype=torch.float32)
    if ext.numel() != hd // 2:
        raise ValueError(f"LongRoPE factor lists need head_dim/2 = {hd // 2} entries, got {ext.numel()}")
    inv_freq = 1.0 / (ext * d.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    scale = d.max_pos / orig
    mscale = 1.0 if scale <= 1.0 else math.sqrt(1.0 + math.log(scale) / math.log(orig))
    return inv_freq, mscale, use_long


_DTYPES = {torch.float32: _lib.MJV_DTYPE_F32, torch.bfloat16: _lib.MJV_DTYPE_BF16, torch.float16: _lib.MJV_DTYPE_F16}


class Engine:
    """One model on one CUDA device. Not thread-safe (like the reference nn.Module)."""

    def __init__(self, dims: EngineDims, device: torch.device | str = "cuda", rope_inv_freq_bf16: bool = True,
                 precision: str = "bf16"):
        """precision: "bf16" = the product path (tcgen05 kernels, fp32 accumulation); "fp32" = the verification mode of
        north_star ("1e-4 in fp32"): every weight and activation in fp32, one plain CUDA kernel per reference op
        (csrc/engine_fp32.cuh) - what a float32 InternVLChatRewardModeling computes (moe_reward.py:183-297)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.lib = _lib.load()
        self.dims = dims
        self.precision = precision
        self.fp32 = precision == "fp32"
        self.rope_inv_freq_bf16 = rope_inv_freq_bf16 and not self.fp32
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MjvError("mj_video_b200 runs on CUDA devices only (no CPU fallback)")
        if dims.downsample_ratio != 0.5:
            raise _lib.MjvError("only downsample_ratio=0.5 (pixel-shuffle 2x2) is implemented")
        cfg = _lib.Config()
        for f in ("image_size", "patch_size", "vit_hidden", "vit_heads", "vit_mlp", "vit_layers", "vit_eps", "llm_arch",
                  "llm_hidden", "llm_inter", "llm_layers", "llm_heads", "llm_kv_heads", "vocab", "rms_eps", "max_pos",
                  "mlp1_eps", "num_objectives", "num_aspects", "gating_hidden", "gating_n_hidden",
                  "gating_temperature", "ps_version"):
            setattr(cfg, f, getattr(dims, f))
        cfg.precision = 1 if self.fp32 else 0
        off = 0
        for a, (_, items) in enumerate(dims.aspect2criteria.items()):
            cfg.group_offsets[a] = off
            for it in items:
                cfg.group_items[off] = int(it)
                off += 1
        for a in range(len(dims.aspect2criteria), 33):
            cfg.group_offsets[a] = off
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mjv_create(C.byref(cfg), C.byref(self._handle)))
        self._workspace: torch.Tensor | None = None
        self._finalized = False
        self._load_rope()

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                self.lib.mjv_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:  # noqa: BLE001
            pass

    # -------------------------------------------------------------------------------------------- weights
    def _load_rope(self, seq_len: int | None = None, call_len: int = 0):
        """cos/sin tables as the reference's rotary modules produce them in a bf16 model, one column per rotary pair of
        the 128-wide head slot the kernels use (pairs beyond head_dim/2 are padding: cos 1, sin 0).
        InternLM2 (modeling_internlm2.py:147-180): fp32 tables for max_position_embeddings rows built at init, cast with
        the model -> bf16-rounded values. When a padded batch is LONGER than the cached table, the dynamic-NTK class
        (:210-231) re-bases theta for that length (only beyond max_position_embeddings), rebuilds the table and keeps it
        for every later call — `seq_len` does the same here.
        Llama branch (modeling_internvl_chat.py:125-126 -> transformers LlamaRotaryEmbedding): no cache and no scaling;
        `inv_freq` is a buffer, so `model.to(bfloat16)` rounds inv_freq itself to bf16; angles are then computed in fp32
        and cos/sin cast to bf16. `rope_inv_freq_bf16=False` keeps fp32 inv_freq (a bf16 checkpoint loaded without a
        later `.to()`)."""
        d = self.dims
        hd = d.llm_hidden // d.llm_heads
        if hd % 2 or hd > 128:
            raise _lib.MjvError(f"LLM head_dim {hd} unsupported (even, <= 128)")
        rows = d.max_pos if seq_len is None else int(seq_len)
        base = d.rope_theta
        linear = d.llm_arch == 0 and d.rope_scaling_type == "linear" and d.rope_scaling_factor is not None
        if d.llm_arch == 0 and seq_len is not None and not linear:
            if d.rope_scaling_factor is None:
                # the plain InternLM2RotaryEmbedding re-caches with the model's (bf16) inv_freq and bf16 positions
                raise NotImplementedError(f"a batch of {rows} positions exceeds max_position_embeddings={d.max_pos} and "
                                          "the config has no dynamic rope_scaling")
            if rows > d.max_pos:
                f = d.rope_scaling_factor
                base = base * ((f * rows / d.max_pos) - (f - 1)) ** (hd / (hd - 2))
        inv_freq = 1.0 / (base ** (torch.arange(0, hd, 2).float() / hd))
        mscale = 1.0
        if d.longrope:
            inv_freq, mscale, self._rope_long = longrope_inv_freq(d, hd, call_len)   # factor set by the CALL's length
        elif d.llm_arch == 1 and self.rope_inv_freq_bf16:
            inv_freq = inv_freq.to(torch.bfloat16).float()
        t = torch.arange(rows, dtype=inv_freq.dtype)
        if linear:
            t = t / d.rope_scaling_factor   # InternLM2LinearScalingRotaryEmbedding (modeling_internlm2.py:183-199)
        freqs = torch.einsum("i,j->ij", t, inv_freq)
        cos = torch.ones((rows, 64), dtype=torch.float32)
        sin = torch.zeros((rows, 64), dtype=torch.float32)
        if self.fp32:   # a float32 model keeps the tables as computed (modeling_internlm2.py:166-170 casts to the model dtype)
            cos[:, : hd // 2] = freqs.cos() * mscale
            sin[:, : hd // 2] = freqs.sin() * mscale
        else:
            cos[:, : hd // 2] = (freqs.cos() * mscale).to(torch.bfloat16).to(torch.float32)
            sin[:, : hd // 2] = (freqs.sin() * mscale).to(torch.bfloat16).to(torch.float32)
        self.load_weight("rope.cos", cos)
        self.load_weight("rope.sin", sin)
        self._rope_rows = rows

    def _ensure_rope(self, padded_len: int):
        """The reference looks at kv_seq_len = the PADDED batch length (modeling_internlm2.py:172-175, 380). LongRoPE picks
        its factor set per call from that length, so crossing the original context in either direction rebuilds the table."""
        d = self.dims
        flip = d.longrope and (padded_len > (d.rope_original_max_pos or d.max_pos)) != getattr(self, "_rope_long", False)
        if padded_len > self._rope_rows or flip:
            self._load_rope(seq_len=max(padded_len, self._rope_rows) if d.longrope else padded_len, call_len=padded_len)
            self.finalize()

    def load_weight(self, key: str, tensor: torch.Tensor):
        t = tensor.detach()
        if t.dtype not in _DTYPES:
            t = t.to(torch.float32)
        t = t.contiguous()
        shape = (C.c_int64 * max(1, t.dim()))(*([int(s) for s in t.shape] or [1]))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mjv_load_weight(self._handle, key.encode(), t.data_ptr(), shape, max(1, t.dim()),
                                                _DTYPES[t.dtype]))
        self._finalized = False

    def load_state_dict(self, state_dict: dict):
        for k, v in state_dict.items():
            self.load_weight(k, v)
        self.finalize()

    def finalize(self):
        _lib.check(self.lib.mjv_finalize_weights(self._handle))
        self._finalized = True

    # -------------------------------------------------------------------------------------------- batch metadata
    def prepare_batch(self, input_ids: torch.Tensor, attention_mask: torch.Tensor | None, n_tiles: int,
                      pad_token_id, img_context_token_id: int):
        """Host-side indexing of a right-padded batch. Returns (meta dict of numpy int32 arrays, scalars)."""
        ids = input_ids.detach().to("cpu", torch.long).numpy()
        B, N = ids.shape
        if pad_token_id is None and B != 1:
            raise ValueError("Cannot handle batch sizes > 1 if no padding token is defined.")
        if attention_mask is not None:
            mask = attention_mask.detach().to("cpu").numpy() != 0
            lengths = mask.sum(1).astype(np.int64)
            if not all(mask[b, : lengths[b]].all() for b in range(B)):
                raise ValueError("attention_mask must be right-padded (contiguous ones then zeros)")
        else:
            lengths = np.full(B, N, dtype=np.int64)
        if (lengths <= 0).any():
            raise ValueError("empty sequence in batch")
        # last-token index (moe_reward.py:226-228)
        if pad_token_id is None:
            last = np.full(B, N - 1, dtype=np.int64)
        else:
            last = ((ids == pad_token_id).astype(np.int32).argmax(-1) - 1) % N
        gate = np.array([find_token_for_gating(ids[b]) for b in range(B)], dtype=np.int64)
        for b in range(B):
            if last[b] >= lengths[b] or gate[b] >= lengths[b]:
                raise ValueError("the reward/gating token lies in the padded region of the row")
        cu = np.zeros(B + 1, dtype=np.int64)
        cu[1:] = np.cumsum(lengths)
        M = int(cu[-1])
        token_ids = np.concatenate([ids[b, : lengths[b]] for b in range(B)])
        # nn.Embedding raises IndexError for an id outside the table (e.g. a tokenizer with added tokens beyond
        # config.vocab_size, modeling_internvl_chat.py:163); the embed kernel would silently leave a stale row instead
        if token_ids.size and (int(token_ids.min()) < 0 or int(token_ids.max()) >= self.dims.vocab):
            raise IndexError(f"index out of range in self: token id {int(token_ids.max())} / {int(token_ids.min())} "
                             f"outside the embedding table of {self.dims.vocab} rows")
        token_ids = token_ids.astype(np.int32)
        row_pos = np.concatenate([np.arange(lengths[b]) for b in range(B)]).astype(np.int32)
        vt_starts = np.zeros(B, dtype=np.int64)
        off = 0
        for b in range(B):
            vt_starts[b] = off
            off += (int(lengths[b]) + 7) // 8 * 8
        vt_cols = off
        row_vt = np.concatenate([vt_starts[b] + np.arange(lengths[b]) for b in range(B)]).astype(np.int32)
        img_rows = np.flatnonzero(token_ids == img_context_token_id).astype(np.int32)
        n_rows_vit = n_tiles * self.dims.tokens_per_tile
        if img_rows.size > n_rows_vit:
            # the reference's fallback path raises here too (shape mismatch inside the except branch,
            # modeling_internvl_chat.py:180-186)
            raise RuntimeError(f"{img_rows.size} <IMG_CONTEXT> tokens but only {n_rows_vit} visual embeddings")
        if img_rows.size < n_rows_vit:
            warnings.warn(f"warning: {n_rows_vit} visual embeddings for {img_rows.size} <IMG_CONTEXT> tokens; "
                          "truncating like the reference (modeling_internvl_chat.py:178-186)")
        # attention work entries (sequence, query block), heaviest first. With an even GQA group the kernel pairs the two
        # query heads of a kv head itself; otherwise (MHA) two query blocks of one sequence travel in one entry
        pair_blocks = (self.dims.llm_heads // self.dims.llm_kv_heads) % 2 != 0
        work = pack_work_entries([(s, q) for s in range(B) for q in range((int(lengths[s]) + 127) // 128)], pair_blocks)
        gather = np.concatenate([cu[:-1] + last, cu[:-1] + gate]).astype(np.int32)
        # the last decoder layer only needs the attention blocks that contain a gathered row
        work_last = pack_work_entries(
            sorted({(s, int(last[s]) // 128) for s in range(B)} | {(s, int(gate[s]) // 128) for s in range(B)}), pair_blocks)
        meta = dict(token_ids=token_ids, row_pos=row_pos, row_vt=row_vt, img_rows=img_rows,
                    cu_seqlens=cu.astype(np.int32), vt_starts=vt_starts.astype(np.int32), work=work.reshape(-1),
                    gather_rows=gather, work_last=work_last.reshape(-1))
        return meta, dict(B=B, M=M, vt_cols=int(vt_cols), n_work=int(work.shape[0]), n_img=int(img_rows.size),
                          n_work_last=int(work_last.shape[0]),
                          lengths=lengths, last=last, gate=gate)

    def upload_meta(self, meta: dict) -> tuple[torch.Tensor, dict]:
        """One pinned staging buffer, one H2D copy; returns the device buffer and the element offset of each section."""
        offsets, total = {}, 0
        for k, v in meta.items():
            offsets[k] = total
            total += (v.size + 3) // 4 * 4  # keep every section 16-byte aligned
        host = torch.empty(max(total, 4), dtype=torch.int32, pin_memory=True)
        hv = host.numpy()
        for k, v in meta.items():
            hv[offsets[k]: offsets[k] + v.size] = v
        dev = host.to(self.device, non_blocking=True)
        return dev, offsets

    # -------------------------------------------------------------------------------------------- forward
    def workspace(self, B: int, T: int, M: int) -> torch.Tensor:
        need = int(self.lib.mjv_workspace_bytes(self._handle, B, T, M))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    def stage_batch(self, pixel_values: torch.Tensor, input_ids: torch.Tensor, attention_mask: torch.Tensor | None,
                    pad_token_id, img_context_token_id: int) -> "StagedBatch":
        """Index the batch on the host, upload the index metadata (one H2D copy) and allocate outputs. The returned
        object can be launched any number of times with `run` (its inputs then already live in HBM)."""
        if not self._finalized:
            self.finalize()
        self._ensure_rope(int(input_ids.shape[1]))
        d = self.dims
        if pixel_values.dim() != 4:
            raise ValueError(f"wrong pixel_values size: {tuple(pixel_values.shape)}")  # modeling_intern_vit.py:410-413
        # the im2col and ViT kernels index every tile with the configured geometry: anything else would read out of
        # bounds (the reference raises inside nn.Conv2d / the position-embedding add for a wrong channel count or size)
        if tuple(pixel_values.shape[1:]) != (3, d.image_size, d.image_size):
            raise ValueError(f"pixel_values must be [tiles, 3, {d.image_size}, {d.image_size}], got "
                             f"{tuple(pixel_values.shape)}")
        T = int(pixel_values.shape[0])
        meta, info = self.prepare_batch(input_ids, attention_mask, T, pad_token_id, img_context_token_id)
        with torch.cuda.device(self.device):
            px = pixel_values.to(self.device, non_blocking=True)
            if px.dtype not in (torch.bfloat16, torch.float32):
                px = px.to(torch.float32)
            px = px.contiguous()
            dev_meta, off = self.upload_meta(meta)
            B, M = info["B"], info["M"]
            NO, NA, D = d.num_objectives, d.num_aspects, d.llm_hidden
            sb = StagedBatch()
            sb.info = info
            sb.px, sb.dev_meta = px, dev_meta
            sb.scores = torch.empty((B, NO + NA + 1), dtype=torch.float32, device=self.device)
            sb.extras = torch.empty((B, NO + NA + NO + 1), dtype=torch.float32, device=self.device)
            act = torch.float32 if self.fp32 else torch.bfloat16
            sb.hidden = torch.empty((B, D), dtype=act, device=self.device)
            sb.prompt = torch.empty((B, D), dtype=act, device=self.device)
            base = dev_meta.data_ptr()
            b = _lib.Batch()
            b.n_rows, b.n_tiles, b.n_img_tokens, b.n_work = B, T, info["n_img"], info["n_work"]
            b.img_context_token_id = int(img_context_token_id)
            b.n_tokens, b.vt_cols = M, info["vt_cols"]
            hd = d.llm_hidden // d.llm_heads
            b.attn_flops = float(sum(4.0 * d.llm_heads * hd * n * (n + 1) / 2 for n in info["lengths"].tolist()))
            b.pixel_values, b.pixel_dtype = px.data_ptr(), _DTYPES[px.dtype]
            for k in ("token_ids", "row_pos", "row_vt", "img_rows", "cu_seqlens", "vt_starts", "work", "gather_rows",
                      "work_last"):
                setattr(b, k, base + 4 * off[k])
            b.n_work_last = info["n_work_last"]
            b.max_seq_len = int(info["lengths"].max())
            o = _lib.Outputs()
            o.scores, o.extras = sb.scores.data_ptr(), sb.extras.data_ptr()
            o.hidden_state, o.prompt_embedding = sb.hidden.data_ptr(), sb.prompt.data_ptr()
            sb.batch, sb.outputs = b, o
            sb.B, sb.T, sb.M = B, T, M
        return sb

    def gating(self, which: int, x: torch.Tensor, with_softmax: bool = True) -> torch.Tensor:
        """GatingNetwork as a standalone op (moe_reward.py:16-42): which = 0 aspect_gating, 1 criteria_gating.
        x [B, llm_hidden] on this device -> [B, n_out] bf16."""
        if not self._finalized:
            raise _lib.MjvError("weights not loaded")
        d = self.dims
        if x.dim() != 2 or x.shape[1] != d.llm_hidden or x.device != self.device:
            raise ValueError(f"gating: expected a [B, {d.llm_hidden}] tensor on {self.device}")
        act = torch.float32 if self.fp32 else torch.bfloat16
        xb = x.to(act).contiguous()
        n_out = d.num_aspects if which == 0 else d.num_objectives
        out = torch.empty((xb.shape[0], n_out), dtype=act, device=self.device)
        with torch.cuda.device(self.device):
            ws = self.workspace(max(1, xb.shape[0]), 0, 0)
            _lib.check(self.lib.mjv_gating_forward(self._handle, which, xb.data_ptr(), xb.shape[0], int(with_softmax),
                                                   out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.current_stream()))
        return out

    def run(self, sb: "StagedBatch") -> dict:
        """Enqueue the whole forward on the current stream (no synchronisation)."""
        with torch.cuda.device(self.device):
            ws = self.workspace(sb.B, sb.T, sb.M)
            _lib.check(self.lib.mjv_score_forward(self._handle, C.byref(sb.batch), C.byref(sb.outputs), ws.data_ptr(),
                                                  ws.numel(), _lib.current_stream()))
        return dict(scores=sb.scores, extras=sb.extras, hidden_state=sb.hidden, prompt_embedding=sb.prompt,
                    info=sb.info)

    def forward(self, pixel_values: torch.Tensor, input_ids: torch.Tensor, attention_mask: torch.Tensor | None,
                pad_token_id, img_context_token_id: int) -> dict:
        sb = self.stage_batch(pixel_values, input_ids, attention_mask, pad_token_id, img_context_token_id)
        out = self.run(sb)
        # keep the staging buffers alive until the stream has consumed them
        sb.dev_meta.record_stream(torch.cuda.current_stream(self.device))
        sb.px.record_stream(torch.cuda.current_stream(self.device))
        return out


class StagedBatch:
    """Device-resident inputs + index metadata + output buffers of one batch (see Engine.stage_batch)."""
    pass
