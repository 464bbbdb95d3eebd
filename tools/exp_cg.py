"""1-CTA vs CTA-pair GEMM: correctness + speed on bench shapes."""
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
torch.manual_seed(0)
# correctness on awkward shapes first
for (M, N, K) in [(256, 256, 64), (300, 256, 128), (1000, 512, 576), (4100, 1024, 1024), (129, 256, 72), (70428, 2048, 2048)]:
    a, w = r(M, K, sc=0.5), r(N, K, sc=0.5)
    ref = (a.float() @ w.float().t())
    o1 = ops.gemm(a, w, cta_group=1); o2 = ops.gemm(a, w, cta_group=2)
    torch.cuda.synchronize()
    e2 = (o2.float() - ref).abs().max().item(); e1 = (o1.float() - ref).abs().max().item()
    print(f"check M{M} N{N} K{K}: cg1 err {e1:.3e} cg2 err {e2:.3e} equal={torch.equal(o1, o2)}", flush=True)
for (M, N, K, tag) in [(262400, 3072, 1024, "qkv"), (262400, 1024, 1024, "proj"), (262400, 4096, 1024, "fc1"), (262400, 1024, 4096, "fc2"), (70428, 16384, 2048, "w13"), (70428, 2048, 8192, "w2"), (70428, 4096, 2048, "wqkv"), (70428, 2048, 2048, "wo")]:
    a, w, b, ls, res = r(M, K), r(N, K), r(N), r(N), r(M, N if N <= 4096 else 8)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    row = []
    for cg in (1, 2):
        t = timeit(lambda: ops.gemm(a, w, out=out, cta_group=cg))
        row.append(f"cg{cg} store {fl/t/1e9:.0f}")
        if N <= 4096:
            t = timeit(lambda: ops.gemm(a, w, out=out, bias=b, scale=ls, residual=res, epilogue=ops.EPI_RESID, cta_group=cg))
            row.append(f"resid {fl/t/1e9:.0f}")
            t = timeit(lambda: ops.gemm(a, w, out=out, bias=b, epilogue=ops.EPI_GELU, cta_group=cg))
            row.append(f"gelu {fl/t/1e9:.0f}")
    t = timeit(lambda: torch.matmul(a, w.t(), out=out))
    print(f"{tag:5s} M{M} N{N} K{K}: " + " | ".join(row) + f" | cuBLAS {fl/t/1e9:.0f}", flush=True)
    del a, w, res, out
