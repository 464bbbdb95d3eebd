"""GEMM perf experiments: where does the per-tile overhead go? (debug 1 = no TMEM ld/st in epilogue, 2 = ld only)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops
dev = "cuda"
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (M, N, K) in [(262400, 3072, 1024), (262400, 1024, 1024), (262400, 1024, 4096), (70428, 16384, 2048), (70428, 2048, 8192)]:
    a = (torch.randn(M, K, device=dev) * 0.1).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    res = []
    for dbg in (0, 1, 2):
        ms = timeit(lambda: ops.gemm(a, w, out=out, debug=dbg))
        res.append(f"dbg{dbg}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.0f} TF/s")
    ms = timeit(lambda: torch.matmul(a, w.t(), out=out))
    print(f"M{M} N{N} K{K}: " + " | ".join(res) + f" | cuBLAS {ms:.3f} ms {2.0*M*N*K/ms/1e9:.0f} TF/s", flush=True)
    del a, w, out
