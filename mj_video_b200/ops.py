"""Thin torch-facing wrappers over the primitive C-ABI entry points (one per CUDA kernel family).

These exist so that every kernel is separately testable against the oracle; the production forward goes through
`mjv_score_forward` (engine.py). All tensors must live on the current CUDA device.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_GELU, EPI_PATCH, EPI_QKV_LLM, EPI_QKV_VIT, EPI_RESID, EPI_SCATTER, EPI_STORE,  # noqa: F401
                   EPI_SWIGLU)


def _bf16(t: torch.Tensor | None, name: str) -> torch.Tensor | None:
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.bfloat16:
        raise TypeError(f"{name}: expected a CUDA bfloat16 tensor, got {t.device} {t.dtype}")
    return t


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor | None = None, *, epilogue: int = EPI_STORE,
         bias=None, residual=None, scale=None, block_n: int = 0, out_k=None, out_vt=None, rope_cos=None,
         rope_sin=None, row_pos=None, row_map=None, row_vt=None, pos_emb=None, n_q_cols: int = 0, n_k_cols: int = 0,
         vt_stride_head: int = 0, vt_ld: int = 0, seq: int = 0, heads: int = 0, ldo: int | None = None,
         K: int | None = None, cta_group: int = 0, v_head_dim: int = 0, qk_head_slot: int = 0) -> torch.Tensor:
    """D = A · Wᵀ with a fused epilogue (tcgen05 / TMEM / TMA). a: [M, K] bf16, w: [N, K] bf16 (nn.Linear layout)."""
    lib = _lib.load()
    _bf16(a, "a"), _bf16(w, "w")
    M = a.shape[0]
    N = w.shape[0]
    Kdim = K if K is not None else a.shape[1]
    if a.stride(-1) != 1 or w.stride(-1) != 1:
        raise ValueError("gemm operands must have a contiguous innermost dimension")
    if out is None:
        n_out = N // 2 if epilogue == EPI_SWIGLU else N
        out = torch.empty((M, n_out), dtype=torch.bfloat16, device=a.device)
    d = _lib.GemmDesc()
    d.a, d.w, d.out = a.data_ptr(), w.data_ptr(), out.data_ptr()
    d.lda, d.ldw = a.stride(0), w.stride(0)
    d.ldo = ldo if ldo is not None else out.stride(0)
    d.M, d.N, d.K = M, N, Kdim
    d.epilogue, d.block_n = epilogue, block_n
    d.bias = _lib.ptr(_bf16(bias, "bias"))
    d.residual = _lib.ptr(_bf16(residual, "residual"))
    d.ldr = residual.stride(0) if residual is not None else 0
    d.scale = _lib.ptr(_bf16(scale, "scale"))
    d.out_k, d.out_vt = _lib.ptr(out_k), _lib.ptr(out_vt)
    d.rope_cos, d.rope_sin = _lib.ptr(rope_cos), _lib.ptr(rope_sin)
    d.row_pos, d.row_map, d.pos_emb = _lib.ptr(row_pos), _lib.ptr(row_map), _lib.ptr(pos_emb)
    d.row_vt = _lib.ptr(row_vt)
    d.n_q_cols, d.n_k_cols = n_q_cols, n_k_cols
    d.vt_stride_head, d.vt_ld, d.seq, d.heads = vt_stride_head, vt_ld, seq, heads
    d.cta_group = cta_group
    d.v_head_dim = v_head_dim
    d.qk_head_slot = qk_head_slot
    _lib.check(lib.mjv_gemm_bf16(C.byref(d), _lib.current_stream()))
    return out


def attention(q, k, vt, out, *, head_dim, heads, kv_heads, causal, scale, q_col0=0, k_col0=0, nseq=0, seq_len=0,
              cu_seqlens=None, work=None, vt_ld=None, vt_plane_stride=None, vt_starts=None, vt_cols=0, v_head_dim=0,
              out_head_dim=0, qk_head_stride=0):
    """Fused attention. Batched mode: q/k are [nseq*seq_len, ld] and vt is [nseq*heads, head_dim, vt_ld].
    Varlen mode: packed rows with device int32 cu_seqlens[nseq+1] and work[(seq, q_block)] pairs; vt is
    [kv_heads, head_dim, vt_ld]."""
    lib = _lib.load()
    d = _lib.AttnDesc()
    d.q, d.k, d.vt, d.out = q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr()
    d.ldq, d.ldk, d.ldo = q.stride(0), k.stride(0), out.stride(0)
    d.vt_ld = vt_ld if vt_ld is not None else vt.stride(-2)
    d.vt_plane_stride = vt_plane_stride if vt_plane_stride is not None else vt.stride(0)
    d.total_rows = q.shape[0]
    d.head_dim, d.heads, d.kv_heads = head_dim, heads, kv_heads
    d.q_col0, d.k_col0, d.causal = q_col0, k_col0, int(causal)
    d.nseq, d.seq_len = nseq, seq_len
    d.cu_seqlens = _lib.ptr(cu_seqlens)
    d.work = _lib.ptr(work)
    d.vt_starts = _lib.ptr(vt_starts)
    d.vt_cols = vt_cols
    d.n_work = 0 if work is None else work.shape[0]
    d.scale = scale
    d.v_head_dim, d.out_head_dim = v_head_dim, out_head_dim
    d.qk_head_stride = qk_head_stride
    _lib.check(lib.mjv_attention(C.byref(d), _lib.current_stream()))
    return out


def causal_work_list(lengths, block_q: int = 128, pair_blocks: bool = False):
    """(sequence, q_block) entries, heaviest (largest q_block) first so long causal rows start early. `pair_blocks`
    packs two blocks of a sequence into one entry (MHA shapes; see engine.pack_work_entries)."""
    from .engine import pack_work_entries
    items = [(s, b) for s, n in enumerate(lengths) for b in range((n + block_q - 1) // block_q)]
    return torch.from_numpy(pack_work_entries(items, pair_blocks))


def vt_layout(lengths, align: int = 8):
    """Column layout of the transposed-V buffer in varlen mode: each sequence starts at a multiple of `align`
    columns (TMA needs a 16-byte aligned innermost coordinate). Returns (starts, total_cols)."""
    starts, off = [], 0
    for n in lengths:
        starts.append(off)
        off += (n + align - 1) // align * align
    return starts, off
