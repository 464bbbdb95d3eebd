"""ctypes binding of libmjv_b200.so (include/mjv_b200.h). No torch types cross this boundary: tensors are passed as
raw device pointers + the current CUDA stream. The product path has NO fallback: if the library is missing this module
raises at first use.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libmjv_b200.so"
_lib = None

MJV_DTYPE_F32, MJV_DTYPE_BF16, MJV_DTYPE_F16 = 0, 1, 2
EPI_STORE, EPI_GELU, EPI_RESID, EPI_SWIGLU, EPI_QKV_LLM, EPI_QKV_VIT, EPI_PATCH, EPI_SCATTER = range(8)


MJV_ERR_NOMEM = -5   # include/mjv_b200.h


class MjvError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("out", C.c_void_p),
        ("lda", C.c_int64), ("ldw", C.c_int64), ("ldo", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("epilogue", C.c_int32), ("block_n", C.c_int32),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("ldr", C.c_int64), ("scale", C.c_void_p),
        ("out_k", C.c_void_p), ("out_vt", C.c_void_p),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
        ("row_pos", C.c_void_p), ("row_map", C.c_void_p), ("row_vt", C.c_void_p), ("pos_emb", C.c_void_p),
        ("n_q_cols", C.c_int32), ("n_k_cols", C.c_int32),
        ("vt_stride_head", C.c_int64), ("vt_ld", C.c_int32), ("seq", C.c_int32), ("heads", C.c_int32),
        ("cta_group", C.c_int32), ("v_head_dim", C.c_int32), ("qk_head_slot", C.c_int32),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("vt", C.c_void_p), ("out", C.c_void_p),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldo", C.c_int64),
        ("vt_ld", C.c_int64), ("vt_plane_stride", C.c_int64), ("total_rows", C.c_int64),
        ("head_dim", C.c_int32), ("heads", C.c_int32), ("kv_heads", C.c_int32),
        ("q_col0", C.c_int32), ("k_col0", C.c_int32), ("causal", C.c_int32),
        ("nseq", C.c_int32), ("seq_len", C.c_int32),
        ("cu_seqlens", C.c_void_p), ("work", C.c_void_p), ("vt_starts", C.c_void_p), ("vt_cols", C.c_int64),
        ("n_work", C.c_int32), ("scale", C.c_float), ("flops_hint", C.c_double),
        ("v_head_dim", C.c_int32), ("out_head_dim", C.c_int32), ("qk_head_stride", C.c_int32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("image_size", C.c_int32), ("patch_size", C.c_int32),
        ("vit_hidden", C.c_int32), ("vit_heads", C.c_int32), ("vit_mlp", C.c_int32), ("vit_layers", C.c_int32),
        ("vit_eps", C.c_float),
        ("llm_arch", C.c_int32),
        ("llm_hidden", C.c_int32), ("llm_inter", C.c_int32), ("llm_layers", C.c_int32), ("llm_heads", C.c_int32),
        ("llm_kv_heads", C.c_int32), ("vocab", C.c_int32),
        ("rms_eps", C.c_float), ("max_pos", C.c_int32), ("mlp1_eps", C.c_float),
        ("num_objectives", C.c_int32), ("num_aspects", C.c_int32), ("gating_hidden", C.c_int32),
        ("gating_n_hidden", C.c_int32), ("gating_temperature", C.c_float),
        ("group_offsets", C.c_int32 * 33), ("group_items", C.c_int32 * 32),
        ("ps_version", C.c_int32), ("precision", C.c_int32),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int32), ("n_tiles", C.c_int32), ("n_img_tokens", C.c_int32), ("n_work", C.c_int32),
        ("img_context_token_id", C.c_int32),
        ("n_tokens", C.c_int64), ("vt_cols", C.c_int64), ("attn_flops", C.c_double),
        ("pixel_values", C.c_void_p), ("pixel_dtype", C.c_int32),
        ("token_ids", C.c_void_p), ("row_pos", C.c_void_p), ("row_vt", C.c_void_p), ("img_rows", C.c_void_p),
        ("cu_seqlens", C.c_void_p), ("vt_starts", C.c_void_p), ("work", C.c_void_p), ("gather_rows", C.c_void_p),
        ("work_last", C.c_void_p), ("n_work_last", C.c_int32), ("max_seq_len", C.c_int32),
    ]


class Outputs(C.Structure):
    _fields_ = [("scores", C.c_void_p), ("extras", C.c_void_p), ("hidden_state", C.c_void_p),
                ("prompt_embedding", C.c_void_p)]


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise MjvError(
            f"{_LIB_PATH} is missing: build it with `python -m mj_video_b200.build` (or __graft_entry__.build()). "
            "mj_video_b200 has no CPU/PyTorch fallback.")
    lib = C.CDLL(str(_LIB_PATH))
    lib.mjv_abi_version.restype = C.c_int
    lib.mjv_abi_struct_size.argtypes = [C.c_int]
    lib.mjv_abi_struct_size.restype = C.c_int
    for which, mirror in enumerate((GemmDesc, AttnDesc, Config, Batch, Outputs)):
        if lib.mjv_abi_struct_size(which) != C.sizeof(mirror):
            raise MjvError(f"ABI mismatch: {mirror.__name__} is {C.sizeof(mirror)} bytes here but "
                           f"{lib.mjv_abi_struct_size(which)} in {_LIB_PATH.name}; rebuild the library")
    lib.mjv_last_error.restype = C.c_char_p
    lib.mjv_launch_count.restype = C.c_longlong
    lib.mjv_reset_launch_count.restype = None
    lib.mjv_profile_enable.argtypes = [C.c_uint32]
    lib.mjv_profile_enable.restype = None
    lib.mjv_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.mjv_profile_read.restype = C.c_int
    lib.mjv_kernel_tag_name.argtypes = [C.c_int32]
    lib.mjv_kernel_tag_name.restype = C.c_char_p
    lib.mjv_kernel_tag_count.restype = C.c_int
    lib.mjv_gemm_bf16.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    lib.mjv_gemm_bf16.restype = C.c_int
    lib.mjv_attention.argtypes = [C.POINTER(AttnDesc), C.c_void_p]
    lib.mjv_attention.restype = C.c_int
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.mjv_layernorm.argtypes = [vp, vp, vp, vp, i64, i32, f32, vp]
    lib.mjv_pixel_shuffle_layernorm.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, i32, vp]
    lib.mjv_rmsnorm.argtypes = [vp, vp, vp, i64, i32, f32, vp, vp]
    lib.mjv_embed_tokens.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp]
    lib.mjv_im2col_patches.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp]
    lib.mjv_preprocess_frames.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, i32, vp, i64, C.POINTER(i32), vp]
    lib.mjv_choose_grid.argtypes = [i32, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.mjv_bicubic_coeffs.argtypes = [i32, i32, C.POINTER(i32), vp, vp, i64]
    lib.mjv_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.mjv_destroy.argtypes = [vp]
    lib.mjv_destroy.restype = None
    lib.mjv_load_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32, i32]
    lib.mjv_finalize_weights.argtypes = [vp]
    lib.mjv_workspace_bytes.argtypes = [vp, i32, i32, i64]
    lib.mjv_workspace_bytes.restype = i64
    lib.mjv_vit_forward.argtypes = [vp, vp, i32, i32, vp, vp, i64, vp]
    lib.mjv_project_and_splice.argtypes = [vp, vp, C.POINTER(Batch), vp, vp, i64, vp]
    lib.mjv_llm_prefill.argtypes = [vp, vp, C.POINTER(Batch), vp, vp, i64, vp]
    lib.mjv_reward_head.argtypes = [vp, vp, i32, C.POINTER(Outputs), vp, i64, vp]
    lib.mjv_score_forward.argtypes = [vp, C.POINTER(Batch), C.POINTER(Outputs), vp, i64, vp]
    lib.mjv_gating_forward.argtypes = [vp, i32, vp, i32, i32, vp, vp, i64, vp]
    lib.mjv_gating_forward.restype = C.c_int
    lib.mjv_comm_unique_id.argtypes = [vp]
    lib.mjv_comm_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    lib.mjv_comm_destroy.argtypes = [vp]
    lib.mjv_allgather_scores.argtypes = [vp, vp, i64, i32, vp, vp]
    for name in ("mjv_comm_unique_id", "mjv_comm_create", "mjv_comm_destroy", "mjv_allgather_scores"):
        getattr(lib, name).restype = C.c_int
    for name in ("mjv_layernorm", "mjv_pixel_shuffle_layernorm", "mjv_rmsnorm", "mjv_embed_tokens",
                 "mjv_im2col_patches", "mjv_preprocess_frames", "mjv_choose_grid", "mjv_bicubic_coeffs", "mjv_create",
                 "mjv_load_weight", "mjv_finalize_weights", "mjv_vit_forward", "mjv_project_and_splice",
                 "mjv_llm_prefill", "mjv_reward_head", "mjv_score_forward"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().mjv_last_error()
        raise MjvError(f"mjv error {rc}: {msg.decode() if msg else ''}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def profile_collect(capacity: int = 1 << 16):
    import numpy as np
    lib = load()
    tags = np.zeros(capacity, np.int32)
    ms = np.zeros(capacity, np.float32)
    work = np.zeros(capacity, np.float64)
    n = lib.mjv_profile_read(tags.ctypes.data, ms.ctypes.data, work.ctypes.data, capacity)
    n = min(n, capacity)
    names = [lib.mjv_kernel_tag_name(int(t)).decode() for t in tags[:n]]
    return list(zip(names, ms[:n].tolist(), work[:n].tolist()))


def tag_mask(*names) -> int:
    lib = load()
    all_names = [lib.mjv_kernel_tag_name(i).decode() for i in range(lib.mjv_kernel_tag_count())]
    if not names:
        return (1 << len(all_names)) - 1
    return sum(1 << all_names.index(n) for n in names)
