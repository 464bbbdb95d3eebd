"""One bench step between cudaProfilerStart/Stop for `ncu --profile-from-start off` (launch list of a real step),
or, with --kernel NAME, a standalone launch of one hot kernel at its bench shape for `ncu --set full`."""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kernel", default=None, choices=[None, "swiglu", "gelu", "resid_fc2", "qkv_vit", "attn_vit", "attn_llm",
                                                   "layernorm", "rmsnorm", "gather_ln", "frames", "im2col", "resize720"])
ap.add_argument("--batch", type=int, default=32)
args = ap.parse_args()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, scale=0.1):
    return (torch.randn(*shape, device=dev, generator=g) * scale).bfloat16()


if args.kernel is None:
    import bench
    from mj_video_b200.data_processor import preprocess_frames
    from mj_video_b200.engine import Engine, EngineDims
    d = synth.dims_2b()
    eng = Engine(EngineDims.from_synth(d), dev)
    eng.load_state_dict(synth.make_state_dict(d, seed=1234, include_lm_head=False))
    frames_np, rows = bench.make_rows(args.batch, 0)
    ids, mask = synth.pad_batch(rows)
    frames = torch.from_numpy(frames_np).to(dev)
    px = torch.empty((frames.shape[0], 3, 448, 448), dtype=torch.bfloat16, device=dev)
    sb = eng.stage_batch(px, ids, mask, synth.PAD, synth.IMG_CONTEXT)
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        preprocess_frames(frames, 448, 1, True, torch.bfloat16, out=px)
        eng.run(sb)
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

Mv, Ml = 256 * 1025, 70428
reps = 3
if args.kernel == "swiglu":
    a, w = rnd(Ml, 2048), rnd(16384, 2048)
    fn = lambda: ops.gemm(a, w, epilogue=ops.EPI_SWIGLU)  # noqa: E731
elif args.kernel == "gelu":
    a, w, b = rnd(Mv, 1024), rnd(4096, 1024), rnd(4096)
    fn = lambda: ops.gemm(a, w, bias=b, epilogue=ops.EPI_GELU)  # noqa: E731
elif args.kernel == "resid_fc2":
    a, w, b, ls, res = rnd(Mv, 4096), rnd(1024, 4096), rnd(1024), rnd(1024), rnd(Mv, 1024)
    fn = lambda: ops.gemm(a, w, out=res, bias=b, scale=ls, residual=res, epilogue=ops.EPI_RESID)  # noqa: E731
elif args.kernel == "qkv_vit":
    a, w, b = rnd(Mv, 1024), rnd(3072, 1024), rnd(3072)
    qk = torch.empty(Mv, 2048, device=dev, dtype=torch.bfloat16)
    vt = torch.empty(256 * 16, 64, 1032, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, out=qk, epilogue=ops.EPI_QKV_VIT, bias=b, out_vt=vt, n_q_cols=1024, n_k_cols=1024,  # noqa: E731
                          vt_stride_head=64 * 1032, vt_ld=1032, seq=1025, heads=16)
elif args.kernel == "attn_vit":
    qk = rnd(Mv, 2048, scale=1.0)
    vt = rnd(256 * 16, 64, 1032, scale=1.0)
    out = torch.empty(Mv, 1024, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.attention(qk, qk, vt, out, head_dim=64, heads=16, kv_heads=16, causal=False, scale=0.125,  # noqa: E731
                               q_col0=0, k_col0=1024, nseq=256, seq_len=1025)
elif args.kernel == "attn_llm":
    lens = [2201] * 32
    tot = sum(lens)
    q, k = rnd(tot, 2048, scale=1.0), rnd(tot, 1024, scale=1.0)
    starts, cols = ops.vt_layout(lens)
    vt = rnd(8, 128, cols, scale=1.0)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    work = ops.causal_work_list(lens).to(dev)
    vts = torch.tensor(starts, dtype=torch.int32, device=dev)
    out = torch.empty(tot, 2048, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.attention(q, k, vt, out, head_dim=128, heads=16, kv_heads=8, causal=True, scale=1 / math.sqrt(128),  # noqa: E731
                               cu_seqlens=cu, work=work, vt_starts=vts, vt_cols=cols)
elif args.kernel == "layernorm":
    from mj_video_b200 import _lib
    x, w, b = rnd(Mv, 1024, scale=1.0), rnd(1024), rnd(1024)
    o = torch.empty_like(x)
    lib = _lib.load()
    fn = lambda: lib.mjv_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), Mv, 1024, 1e-6, _lib.current_stream())  # noqa: E731
elif args.kernel == "rmsnorm":
    from mj_video_b200 import _lib
    x, w = rnd(Ml, 2048, scale=1.0), rnd(2048)
    o = torch.empty_like(x)
    lib = _lib.load()
    fn = lambda: lib.mjv_rmsnorm(x.data_ptr(), w.data_ptr(), o.data_ptr(), Ml, 2048, 1e-5, None, _lib.current_stream())  # noqa: E731
elif args.kernel == "gather_ln":
    from mj_video_b200 import _lib
    x, w, b = rnd(Mv, 1024, scale=1.0), rnd(4096), rnd(4096)
    o = torch.empty(256 * 256, 4096, device=dev, dtype=torch.bfloat16)
    lib = _lib.load()
    fn = lambda: lib.mjv_pixel_shuffle_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), 256, 1025, 1024, 1e-5, 0,  # noqa: E731
                                                 _lib.current_stream())
elif args.kernel == "frames":
    from mj_video_b200.data_processor import preprocess_frames
    fr = torch.randint(0, 256, (256, 448, 448, 3), device=dev, dtype=torch.uint8)
    px = torch.empty(256, 3, 448, 448, device=dev, dtype=torch.bfloat16)
    fn = lambda: preprocess_frames(fr, 448, 1, True, torch.bfloat16, out=px)  # noqa: E731
elif args.kernel == "resize720":   # fused bicubic resize, 256 frames 1280x720 -> 448x448
    from mj_video_b200.data_processor import preprocess_frames
    fr = torch.randint(0, 256, (256, 720, 1280, 3), device=dev, dtype=torch.uint8)
    px = torch.empty(256, 3, 448, 448, device=dev, dtype=torch.bfloat16)
    fn = lambda: preprocess_frames(fr, 448, 1, True, torch.bfloat16, out=px)  # noqa: E731
elif args.kernel == "im2col":
    from mj_video_b200 import _lib
    px = rnd(256, 3, 448, 448, scale=1.0)
    o = torch.empty(256 * 1024, 640, device=dev, dtype=torch.bfloat16)
    lib = _lib.load()
    fn = lambda: lib.mjv_im2col_patches(px.data_ptr(), _lib.MJV_DTYPE_BF16, o.data_ptr(), 256, 448, 14, 640, _lib.current_stream())  # noqa: E731
for _ in range(2):
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(reps):
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
