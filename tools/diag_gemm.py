"""GPU diagnostic for the tcgen05 GEMM: correctness vs torch on several shapes + epilogues, structured error
report on mismatch, and a first throughput number. Run under gpurun; writes gpurun_out/diag_gemm.log."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def report(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-9
    mx = err.max().item()
    ok = mx <= tol * max(1.0, denom)
    print(f"[{name}] max_abs_err={mx:.4e} ref_absmax={denom:.3e} nan={torch.isnan(got).any().item()} "
          f"{'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        bad = err > tol * max(1.0, denom)
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print(f"   bad rows: {rows.numel()}/{bad.shape[0]} first {rows[:16].tolist()}")
        print(f"   bad cols: {cols.numel()}/{bad.shape[1]} first {cols[:16].tolist()}")
        print("   got[0,:8]", got[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
        r = rows[0].item() if rows.numel() else 0
        print(f"   got[{r},:8]", got[r, :8].tolist())
        print(f"   ref[{r},:8]", ref[r, :8].tolist())
        # ratio got/ref to reveal partial-K accumulation
        ratio = (got[:8, :8] / (ref[:8, :8] + 1e-9))
        print("   ratio[:4,:4]", ratio[:4, :4].tolist())
    return ok


def ref_mm(a, w):
    return a.float() @ w.float().t()


all_ok = True
# 1. structured single-tile tests: identity-like patterns expose layout bugs
for (M, N, K, bn) in [(128, 128, 64, 128), (128, 256, 64, 256), (128, 256, 256, 256), (256, 512, 1024, 256),
                      (1000, 384, 576, 128), (4100, 1024, 1024, 256), (333, 256, 640, 256)]:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    try:
        out = ops.gemm(a, w, block_n=bn)
        torch.cuda.synchronize()
        all_ok &= report(f"store M{M} N{N} K{K} bn{bn}", out, ref_mm(a, w).bfloat16().float())
    except Exception as e:  # noqa: BLE001
        print(f"[store M{M} N{N} K{K}] EXCEPTION {e}", flush=True)
        all_ok = False
        break

if all_ok:
    M, N, K = 1000, 1024, 512
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    res = torch.randn(M, N, device=dev).bfloat16()
    ls = torch.randn(N, device=dev).bfloat16()
    acc = ref_mm(a, w)
    # bias
    out = ops.gemm(a, w, bias=bias)
    all_ok &= report("bias", out, (acc + bias.float()).bfloat16())
    # gelu
    out = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GELU)
    all_ok &= report("gelu", out, torch.nn.functional.gelu((acc + bias.float()).bfloat16().float()).bfloat16())
    # resid
    out = ops.gemm(a, w, bias=bias, residual=res, scale=ls, epilogue=ops.EPI_RESID)
    t = (acc + bias.float()).bfloat16().float()
    t = (t * ls.float()).bfloat16().float()
    all_ok &= report("resid", out, (res.float() + t).bfloat16())
    out = ops.gemm(a, w, residual=res, epilogue=ops.EPI_RESID)
    all_ok &= report("resid-plain", out, (res.float() + acc.bfloat16().float()).bfloat16())
    # swiglu: packed weight rows interleaved [32 gate | 32 up]
    F = 512
    w1 = (torch.randn(F, K, device=dev) * 0.1).bfloat16()
    w3 = (torch.randn(F, K, device=dev) * 0.1).bfloat16()
    packed = torch.stack([w1.view(F // 32, 32, K), w3.view(F // 32, 32, K)], dim=1).reshape(2 * F, K).contiguous()
    out = ops.gemm(a, packed, epilogue=ops.EPI_SWIGLU)
    g = ref_mm(a, w1).bfloat16().float()
    u = ref_mm(a, w3).bfloat16().float()
    all_ok &= report("swiglu", out, (torch.nn.functional.silu(g).bfloat16().float() * u).bfloat16())
    # scatter
    perm = torch.randperm(M, device=dev).int()
    outbuf = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out=outbuf, bias=bias, row_map=perm, epilogue=ops.EPI_SCATTER)
    ref = torch.zeros(M, N, device=dev)
    ref[perm.long()] = (acc + bias.float()).bfloat16().float()
    all_ok &= report("scatter", outbuf, ref)

# throughput
if all_ok:
    for (M, N, K) in [(8192, 8192, 8192), (262400, 3072, 1024), (262400, 1024, 4096), (69984, 16384, 2048)]:
        a = (torch.randn(M, K, device=dev) * 0.1).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, w, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            ops.gemm(a, w, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = 2.0 * M * N * K / ms / 1e9
        # cuBLAS for comparison
        for _ in range(3):
            torch.matmul(a, w.t(), out=out)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(a, w.t(), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / iters
        print(f"[perf] M{M} N{N} K{K}: mjv {ms:.3f} ms {tf:.1f} TFLOP/s | cuBLAS {ms2:.3f} ms "
              f"{2.0 * M * N * K / ms2 / 1e9:.1f} TFLOP/s", flush=True)
        del a, w, out

print("ALL_OK" if all_ok else "SOME_FAILED")
