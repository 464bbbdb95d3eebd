"""Attention A/B on one B200: attention2_kernel (default) vs the first-generation kernel (MJV_ATTN_IMPL=1) vs flash-attn 2
(the library kernel the reference calls: modeling_intern_vit.py:72-75, modeling_internlm2.py:498-561), on the bench shapes.

Checks every variant against an fp32 torch reference on a slice first, then times with CUDA events (inputs > L2).
Writes gpurun_out/attn_bench.json. Not part of the product; the product never imports flash_attn or torch ops.
"""
from __future__ import annotations

import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
OUT = {}
POLY_SWEEP = tuple(int(x) for x in os.environ.get("MJV_BENCH_POLY", "").split(",") if x)


def timeit(fn, iters=10, warm=3):
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


def ref_attn(q, k, v, causal, scale):
    """q [n,H,d], k/v [n,Hkv,d] fp32 -> [n,H*d]"""
    n, H, d = q.shape
    g = H // k.shape[1]
    k = k.repeat_interleave(g, 1)
    v = v.repeat_interleave(g, 1)
    s = torch.einsum("qhd,khd->hqk", q, k) * scale
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(n, n, device=q.device, dtype=torch.bool), 1), float("-inf"))
    return torch.einsum("hqk,khd->qhd", s.softmax(-1), v).reshape(n, H * v.shape[-1])


def vit_case(T, S=1025, H=16, check_tiles=2):
    Cv = H * 64
    qk = (torch.randn(T * S, 2 * Cv, device=dev) * 1.0).bfloat16()
    S_pad = (S + 7) // 8 * 8
    v = torch.randn(T, S, H, 64, device=dev).bfloat16()
    vt = torch.zeros(T * H, 64, S_pad, device=dev, dtype=torch.bfloat16)
    vt[:, :, :S] = v.permute(0, 2, 3, 1).reshape(T * H, 64, S)
    flops = 4.0 * H * 64 * S * S * T
    res = {"shape": f"T{T} S{S} H{H} hd64 bidirectional", "flops": flops}
    refs = None
    for impl in ("2", "1"):
        os.environ["MJV_ATTN_IMPL"] = impl
        out = torch.full((T * S, Cv), float("nan"), device=dev, dtype=torch.bfloat16)
        run = lambda: ops.attention(qk, qk, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125,  # noqa: E731
                                    q_col0=0, k_col0=Cv, nseq=T, seq_len=S)
        run()
        torch.cuda.synchronize()
        if refs is None:
            refs = []
            for t in list(range(check_tiles)) + [T - 1]:
                q = qk[t * S:(t + 1) * S, :Cv].view(S, H, 64).float()
                k = qk[t * S:(t + 1) * S, Cv:].view(S, H, 64).float()
                refs.append((t, ref_attn(q, k, v[t].float(), False, 0.125)))
        err = max((out[t * S:(t + 1) * S].float() - r).abs().max().item() for t, r in refs)
        nan = bool(torch.isnan(out.float()).any())
        ms = timeit(run)
        res[f"impl{impl}"] = {"ms": ms, "tflops": flops / ms / 1e9, "max_err": err, "nan": nan}
        print(f"[vit T{T}] impl{impl}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  err {err:.2e} nan {nan}", flush=True)
    os.environ.pop("MJV_ATTN_IMPL", None)
    for poly in POLY_SWEEP:   # share of exponentials evaluated on the FMA pipe (of 8 pairs)
        os.environ["MJV_ATTN_POLY"] = str(poly)
        out = torch.full((T * S, Cv), float("nan"), device=dev, dtype=torch.bfloat16)
        run = lambda: ops.attention(qk, qk, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125,  # noqa: E731
                                    q_col0=0, k_col0=Cv, nseq=T, seq_len=S)
        run()
        torch.cuda.synchronize()
        err = max((out[t * S:(t + 1) * S].float() - r).abs().max().item() for t, r in refs)
        ms = timeit(run)
        res[f"impl2_poly{poly}"] = {"ms": ms, "tflops": flops / ms / 1e9, "max_err": err}
        print(f"[vit T{T}] impl2 poly {poly}/8: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  err {err:.2e}", flush=True)
    os.environ.pop("MJV_ATTN_POLY", None)
    try:
        from flash_attn import flash_attn_varlen_qkvpacked_func
        qkv = torch.cat([qk.view(T * S, 2, H, 64), v.reshape(T * S, 1, H, 64)], 1).contiguous()
        cu = torch.arange(0, (T + 1) * S, S, device=dev, dtype=torch.int32)
        run = lambda: flash_attn_varlen_qkvpacked_func(qkv, cu, S, 0.0, softmax_scale=0.125, causal=False)  # noqa: E731
        o = run()
        err = max((o[t * S:(t + 1) * S].reshape(S, Cv).float() - r).abs().max().item() for t, r in refs)
        ms = timeit(run)
        res["flash_attn2"] = {"ms": ms, "tflops": flops / ms / 1e9, "max_err": err}
        print(f"[vit T{T}] flash-attn2: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  err {err:.2e}", flush=True)
    except Exception as e:  # noqa: BLE001
        res["flash_attn2"] = {"error": repr(e)[:200]}
        print("flash-attn2 failed:", repr(e)[:200], flush=True)
    return res


def llm_case(lens, H=16, Hkv=8, hd=128, check=2, pair_blocks=False, tag="llm"):
    tot = sum(lens)
    q = torch.randn(tot, H, hd, device=dev).bfloat16()
    k = torch.randn(tot, Hkv, hd, device=dev).bfloat16()
    v = torch.randn(tot, Hkv, hd, device=dev).bfloat16()
    starts, cols = ops.vt_layout(lens)
    vt = torch.zeros(Hkv, hd, cols, device=dev, dtype=torch.bfloat16)
    o = 0
    for st, n in zip(starts, lens):
        vt[:, :, st:st + n] = v[o:o + n].permute(1, 2, 0)
        o += n
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    vts = torch.tensor(starts, dtype=torch.int32, device=dev)
    flops = float(sum(4.0 * H * hd * n * (n + 1) / 2 for n in lens))
    res = {"shape": f"{len(lens)} rows x ~{lens[0]} tokens, {H}/{Hkv} heads hd{hd} causal", "flops": flops}
    refs = None
    scale = 1 / math.sqrt(hd)
    for impl in ("2", "1"):
        if impl == "1" and pair_blocks:
            continue
        os.environ["MJV_ATTN_IMPL"] = impl
        work = ops.causal_work_list(lens, pair_blocks=(pair_blocks and impl == "2")).to(dev)
        out = torch.full((tot, H * hd), float("nan"), device=dev, dtype=torch.bfloat16)
        run = lambda: ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=hd, heads=H, kv_heads=Hkv,  # noqa: E731
                                    causal=True, scale=scale, cu_seqlens=cu, work=work, vt_starts=vts, vt_cols=cols)
        run()
        torch.cuda.synchronize()
        if refs is None:
            refs = []
            offs = [0] + list(torch.tensor(lens).cumsum(0).tolist())
            for s in list(range(check)) + [len(lens) - 1]:
                a, b = offs[s], offs[s + 1]
                refs.append((a, b, ref_attn(q[a:b].float(), k[a:b].float(), v[a:b].float(), True, scale)))
        err = max((out[a:b].float() - r).abs().max().item() for a, b, r in refs)
        nan = bool(torch.isnan(out.float()).any())
        ms = timeit(run)
        res[f"impl{impl}"] = {"ms": ms, "tflops": flops / ms / 1e9, "max_err": err, "nan": nan}
        print(f"[{tag}] impl{impl}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  err {err:.2e} nan {nan}", flush=True)
    os.environ.pop("MJV_ATTN_IMPL", None)
    for poly in (POLY_SWEEP if (hd == 128 and not pair_blocks) else ()):
        os.environ["MJV_ATTN_POLY"] = str(poly)
        work = ops.causal_work_list(lens).to(dev)
        out = torch.full((tot, H * hd), float("nan"), device=dev, dtype=torch.bfloat16)
        run = lambda: ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=hd, heads=H, kv_heads=Hkv,  # noqa: E731
                                    causal=True, scale=scale, cu_seqlens=cu, work=work, vt_starts=vts, vt_cols=cols)
        run()
        torch.cuda.synchronize()
        err = max((out[a:b].float() - r).abs().max().item() for a, b, r in refs)
        ms = timeit(run)
        res[f"impl2_poly{poly}"] = {"ms": ms, "tflops": flops / ms / 1e9, "max_err": err}
        print(f"[{tag}] impl2 poly {poly}/8: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  err {err:.2e}", flush=True)
    os.environ.pop("MJV_ATTN_POLY", None)
    try:
        from flash_attn import flash_attn_varlen_func
        mx = max(lens)
        run = lambda: flash_attn_varlen_func(q, k, v, cu, cu, mx, mx, 0.0, softmax_scale=scale, causal=True)  # noqa: E731
        o = run()
        err = max((o[a:b].reshape(b - a, -1).float() - r).abs().max().item() for a, b, r in refs)
        ms = timeit(run)
        res["flash_attn2"] = {"ms": ms, "tflops": flops / ms / 1e9, "max_err": err}
        print(f"[{tag}] flash-attn2: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  err {err:.2e}", flush=True)
    except Exception as e:  # noqa: BLE001
        res["flash_attn2"] = {"error": repr(e)[:200]}
        print("flash-attn2 failed:", repr(e)[:200], flush=True)
    return res


def gemm_vs_cublas():
    """Our tcgen05 GEMM (plain store epilogue) vs torch.matmul (cuBLAS) on the step's main shapes."""
    shapes = [("vit_qkv", 262400, 3072, 1024), ("vit_proj", 262400, 1024, 1024), ("vit_fc1", 262400, 4096, 1024),
              ("vit_fc2", 262400, 1024, 4096), ("llm_wqkv", 70428, 4096, 2048), ("llm_wo", 70428, 2048, 2048),
              ("llm_w13", 70428, 16384, 2048), ("llm_w2", 70428, 2048, 8192)]
    res = {}
    for name, M, N, K in shapes:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16() * 0.02
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        ms_c = timeit(lambda: torch.matmul(a, w.t(), out=out), iters=5)
        ms_o = timeit(lambda: ops.gemm(a, w, out, epilogue=ops.EPI_STORE), iters=5)
        res[name] = {"M": M, "N": N, "K": K, "cublas_ms": ms_c, "cublas_tflops": fl / ms_c / 1e9, "ours_ms": ms_o,
                     "ours_tflops": fl / ms_o / 1e9}
        print(f"[gemm {name}] cuBLAS {fl / ms_c / 1e9:.0f}  ours(store) {fl / ms_o / 1e9:.0f} TFLOP/s", flush=True)
        del a, w, out
    return res


if __name__ == "__main__":
    what = sys.argv[1:] or ["small", "vit", "llm", "mha", "gemm"]
    t0 = time.time()
    if "small" in what:  # correctness on odd shapes before the big ones
        OUT["vit_small"] = vit_case(3, S=1025, H=2, check_tiles=2)
        OUT["vit_short"] = vit_case(5, S=300, H=3, check_tiles=2)
        OUT["llm_small"] = llm_case([300, 77, 513, 128, 129], H=4, Hkv=2, check=4, tag="llm_small")
        OUT["mha_small"] = llm_case([700, 77, 513], H=3, Hkv=3, check=2, pair_blocks=True, tag="mha_small")
        OUT["mha_single"] = llm_case([700, 77, 513], H=3, Hkv=3, check=2, pair_blocks=False, tag="mha_single")
        # several items per CTA (single-tile items followed by two-tile items, odd step counts)
        OUT["vit_multi"] = vit_case(8, S=1025, H=4, check_tiles=3)
        OUT["llm_multi"] = llm_case([2187, 1500, 2100, 640, 129, 2201], H=16, Hkv=8, check=5, tag="llm_multi")
        OUT["mha_multi"] = llm_case([2187, 1500, 2100, 640, 129, 2201], H=8, Hkv=8, check=5, pair_blocks=True, tag="mha_multi")
    if "vit" in what:
        OUT["vit"] = vit_case(256)
    if "llm" in what:
        g = torch.Generator().manual_seed(1)
        lens = (2141 + torch.randint(16, 97, (32,), generator=g)).tolist()
        OUT["llm"] = llm_case(lens, tag="llm")
    if "llmlong" in what:  # BASELINE configs[3] shaped rows (16 frames x 7 tiles): items of up to 225 key blocks
        OUT["llmlong"] = llm_case([28800, 28800, 300], check=0, tag="llmlong")
    if "mha" in what:
        OUT["mha"] = llm_case([2187] * 16, H=32, Hkv=32, hd=128, pair_blocks=True, tag="mha")
    if "gemm" in what:
        OUT["gemm_vs_cublas"] = gemm_vs_cublas()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/attn_bench.json", "w") as f:
        json.dump(OUT, f, indent=1)
    print(f"done in {time.time() - t0:.0f}s")
