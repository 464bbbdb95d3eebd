"""Timing of the small memory-bound kernels at bench shapes (tools only): im2col, projector gather-LN, LayerNorm, RMSNorm, in
algorithmic GB/s (CUDA events, inputs larger than L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import _lib  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


lib = _lib.load()
st = _lib.current_stream()
dev = "cuda"
T = 256
px = torch.randn(T, 3, 448, 448, device=dev).bfloat16()
o = torch.empty(T * 1024, 640, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: lib.mjv_im2col_patches(px.data_ptr(), _lib.MJV_DTYPE_BF16, o.data_ptr(), T, 448, 14, 640, st))
print(f"im2col        {ms:.4f} ms  {(px.numel() * 2 + o.numel() * 2) / ms / 1e6:.0f} GB/s")
x = torch.randn(T * 1025, 1024, device=dev).bfloat16()
w, b = torch.randn(4096, device=dev).bfloat16(), torch.randn(4096, device=dev).bfloat16()
o2 = torch.empty(T * 256, 4096, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: lib.mjv_pixel_shuffle_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), o2.data_ptr(), T, 1025, 1024, 1e-5, 0, st))
print(f"gather-LN     {ms:.4f} ms  {(T * 1024 * 1024 * 2 + o2.numel() * 2) / ms / 1e6:.0f} GB/s")
y = torch.empty_like(x)
w1, b1 = torch.randn(1024, device=dev).bfloat16(), torch.randn(1024, device=dev).bfloat16()
ms = timeit(lambda: lib.mjv_layernorm(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), y.data_ptr(), x.shape[0], 1024, 1e-6, st))
print(f"layernorm     {ms:.4f} ms  {x.numel() * 4 / ms / 1e6:.0f} GB/s")
for C in (2048, 3072):
    xr = torch.randn(70428, C, device=dev).bfloat16()
    yr = torch.empty_like(xr)
    wr = torch.randn(C, device=dev).bfloat16()
    ms = timeit(lambda: lib.mjv_rmsnorm(xr.data_ptr(), wr.data_ptr(), yr.data_ptr(), xr.shape[0], C, 1e-5, None, st))
    print(f"rmsnorm C={C} {ms:.4f} ms  {xr.numel() * 4 / ms / 1e6:.0f} GB/s")
