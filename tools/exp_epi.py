"""Epilogue cost experiment: same GEMM shape, different epilogues."""
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
def r(*s, sc=0.1): return (torch.randn(*s, device=dev) * sc).bfloat16()
for (M, N, K, tag) in [(262400, 1024, 4096, "vit fc2"), (262400, 1024, 1024, "vit proj"), (262400, 4096, 1024, "vit fc1"), (70428, 2048, 8192, "llm w2"), (70428, 2048, 2048, "llm wo")]:
    a, w, b, ls, res = r(M, K), r(N, K), r(N), r(N), r(M, N)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    t = {}
    t["store"] = timeit(lambda: ops.gemm(a, w, out=out))
    t["bias"] = timeit(lambda: ops.gemm(a, w, out=out, bias=b))
    t["gelu"] = timeit(lambda: ops.gemm(a, w, out=out, bias=b, epilogue=ops.EPI_GELU))
    t["resid"] = timeit(lambda: ops.gemm(a, w, out=out, residual=res, epilogue=ops.EPI_RESID))
    t["resid+b+ls"] = timeit(lambda: ops.gemm(a, w, out=out, bias=b, scale=ls, residual=res, epilogue=ops.EPI_RESID))
    t["resid inplace"] = timeit(lambda: ops.gemm(a, w, out=res, bias=b, scale=ls, residual=res, epilogue=ops.EPI_RESID))
    print(f"{tag:9s} M{M} N{N} K{K}: " + " | ".join(f"{k} {fl/v/1e9:.0f}" for k, v in t.items()), flush=True)
    del a, w, res, out
