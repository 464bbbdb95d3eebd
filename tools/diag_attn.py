"""GPU diagnostic for the tcgen05 attention kernel and the QKV GEMM epilogues that feed it."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def report(name, got, ref, tol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    mx = err.max().item()
    ok = mx <= tol and not torch.isnan(got).any().item()
    print(f"[{name}] max_abs_err={mx:.4e} mean_err={err.mean().item():.3e} ref_absmax={ref.abs().max().item():.3e} "
          f"nan={torch.isnan(got).any().item()} {'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        bad = (err > tol) | torch.isnan(got)
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print(f"   bad rows {rows.numel()}/{bad.shape[0]} first {rows[:12].tolist()} last {rows[-4:].tolist()}")
        print(f"   bad cols {cols.numel()}/{bad.shape[1]} first {cols[:12].tolist()}")
        r = rows[0].item()
        print(f"   got[{r},:6]", got[r, :6].tolist(), "ref", ref[r, :6].tolist())
    return ok


def ref_attn(q, k, v, causal, scale):
    # q [S,H,D] k,v [S,Hkv,D] fp32
    H, Hkv = q.shape[1], k.shape[1]
    k = k.repeat_interleave(H // Hkv, dim=1)
    v = v.repeat_interleave(H // Hkv, dim=1)
    s = torch.einsum("qhd,khd->hqk", q, k) * scale
    if causal:
        S = q.shape[0]
        s = s.masked_fill(torch.ones(S, S, device=q.device, dtype=torch.bool).triu(1), float("-inf"))
    p = s.softmax(-1)
    return torch.einsum("hqk,khd->qhd", p, v).reshape(q.shape[0], -1)


ok_all = True
# ---------------- batched bidirectional (ViT shape): hd 64
for (T, S, H) in [(1, 128, 1), (1, 257, 2), (3, 1025, 2)]:
    C = H * 64
    q = torch.randn(T * S, H, 64, device=dev).bfloat16()
    k = torch.randn(T * S, H, 64, device=dev).bfloat16()
    v = torch.randn(T * S, H, 64, device=dev).bfloat16()
    qk = torch.cat([q.reshape(T * S, C), k.reshape(T * S, C)], dim=1).contiguous()
    S_pad = (S + 7) // 8 * 8
    vt = torch.zeros(T * H, 64, S_pad, device=dev, dtype=torch.bfloat16)
    vt[:, :, :S] = v.reshape(T, S, H, 64).permute(0, 2, 3, 1).reshape(T * H, 64, S)
    out = torch.zeros(T * S, C, device=dev, dtype=torch.bfloat16)
    try:
        ops.attention(qk, qk, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125, q_col0=0,
                      k_col0=C, nseq=T, seq_len=S)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("EXCEPTION", e)
        ok_all = False
        break
    ref = torch.cat([ref_attn(q[t * S:(t + 1) * S].float(), k[t * S:(t + 1) * S].float(), v[t * S:(t + 1) * S].float(),
                              False, 0.125) for t in range(T)])
    ok_all &= report(f"vit-attn T{T} S{S} H{H}", out, ref, 2e-2)

# ---------------- varlen causal GQA (LLM shape): hd 128
if ok_all:
    for lens, H, Hkv in [([128], 1, 1), ([300, 77, 513], 4, 2), ([2187, 2100], 16, 8)]:
        tot = sum(lens)
        q = torch.randn(tot, H, 128, device=dev).bfloat16()
        k = torch.randn(tot, Hkv, 128, device=dev).bfloat16()
        v = torch.randn(tot, Hkv, 128, device=dev).bfloat16()
        starts, tot_pad = ops.vt_layout(lens)
        vt = torch.zeros(Hkv, 128, tot_pad, device=dev, dtype=torch.bfloat16)
        o = 0
        for st, n in zip(starts, lens):
            vt[:, :, st:st + n] = v[o:o + n].permute(1, 2, 0)
            o += n
        vts = torch.tensor(starts, dtype=torch.int32, device=dev)
        cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
        work = ops.causal_work_list(lens).to(dev)
        out = torch.zeros(tot, H * 128, device=dev, dtype=torch.bfloat16)
        try:
            ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=128, heads=H, kv_heads=Hkv,
                          causal=True, scale=1 / math.sqrt(128), cu_seqlens=cu, work=work, vt_starts=vts, vt_cols=tot_pad)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("EXCEPTION", e)
            ok_all = False
            break
        refs, o = [], 0
        for n in lens:
            refs.append(ref_attn(q[o:o + n].float(), k[o:o + n].float(), v[o:o + n].float(), True, 1 / math.sqrt(128)))
            o += n
        ok_all &= report(f"llm-attn lens{lens} H{H}/{Hkv}", out, torch.cat(refs), 2e-2)

# ---------------- QKV epilogues
if ok_all:
    T, S, H = 2, 1025, 2
    C = H * 64
    x = (torch.randn(T * S, C, device=dev) * 0.5).bfloat16()
    w = (torch.randn(3 * C, C, device=dev) * 0.1).bfloat16()
    b = torch.randn(3 * C, device=dev).bfloat16()
    S_pad = (S + 7) // 8 * 8
    qk = torch.zeros(T * S, 2 * C, device=dev, dtype=torch.bfloat16)
    vt = torch.zeros(T * H, 64, S_pad, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, w, out=qk, epilogue=ops.EPI_QKV_VIT, bias=b, out_vt=vt, n_q_cols=C, n_k_cols=C,
             vt_stride_head=64 * S_pad, vt_ld=S_pad, seq=S, heads=H, block_n=128)
    ref = (x.float() @ w.float().t() + b.float()).bfloat16().float()
    ok_all &= report("qkv-vit qk", qk, ref[:, :2 * C], 2e-2)
    vref = ref[:, 2 * C:].reshape(T, S, H, 64).permute(0, 2, 3, 1).reshape(T * H, 64, S)
    ok_all &= report("qkv-vit vT", vt[:, :, :S].reshape(T * H * 64, S), vref.reshape(T * H * 64, S), 2e-2)

    # LLM qkv + rope: packed weight rows [q heads | k heads | v heads]
    tot, D, Hq, Hkv = 777, 256, 2, 1
    x = (torch.randn(tot, D, device=dev) * 0.5).bfloat16()
    w = (torch.randn((Hq + 2 * Hkv) * 128, D, device=dev) * 0.1).bfloat16()
    pos = torch.arange(tot, device=dev, dtype=torch.int32) % 400
    inv = 1.0 / (1e6 ** (torch.arange(0, 128, 2, device=dev).float() / 128))
    fr = torch.outer(torch.arange(512, device=dev).float(), inv)
    cos = fr.cos().bfloat16().float().contiguous()
    sin = fr.sin().bfloat16().float().contiguous()
    tot_pad = (tot + 7) // 8 * 8
    qo = torch.zeros(tot, Hq * 128, device=dev, dtype=torch.bfloat16)
    ko = torch.zeros(tot, Hkv * 128, device=dev, dtype=torch.bfloat16)
    vt = torch.zeros(Hkv, 128, tot_pad, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, w, out=qo, epilogue=ops.EPI_QKV_LLM, out_k=ko, out_vt=vt, rope_cos=cos, rope_sin=sin, row_pos=pos,
             n_q_cols=Hq * 128, n_k_cols=Hkv * 128, vt_stride_head=128 * tot_pad, vt_ld=tot_pad, block_n=128)
    y = (x.float() @ w.float().t()).bfloat16()

    def rope(t):  # t [tot, heads, 128] bf16, mirror bf16 rounding
        c = torch.cat([cos, cos], -1)[pos.long()].bfloat16()[:, None, :]
        s = torch.cat([sin, sin], -1)[pos.long()].bfloat16()[:, None, :]
        rot = torch.cat([-t[..., 64:], t[..., :64]], -1)
        return (t * c) + (rot * s)

    ok_all &= report("qkv-llm q", qo, rope(y[:, :Hq * 128].reshape(tot, Hq, 128)).reshape(tot, -1), 2e-2)
    ok_all &= report("qkv-llm k", ko, rope(y[:, Hq * 128:(Hq + Hkv) * 128].reshape(tot, Hkv, 128)).reshape(tot, -1), 2e-2)
    ok_all &= report("qkv-llm vT", vt[:, :, :tot].reshape(-1, tot),
                     y[:, (Hq + Hkv) * 128:].reshape(tot, Hkv, 128).permute(1, 2, 0).reshape(-1, tot), 2e-2)

# ---------------- perf
if ok_all:
    def timeit(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    T, S, H = 256, 1025, 16
    C = H * 64
    qk = torch.randn(T * S, 2 * C, device=dev).bfloat16()
    S_pad = 1032
    vt = torch.randn(T * H, 64, S_pad, device=dev).bfloat16()
    out = torch.zeros(T * S, C, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(qk, qk, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125,
                                      q_col0=0, k_col0=C, nseq=T, seq_len=S))
    fl = 4.0 * T * H * S * S * 64
    print(f"[perf] vit attention T{T}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  ({T*H*S*S/ms/1e6:.1f} Gexp/s)")
    lens = [2187] * 32
    tot = sum(lens)
    q = torch.randn(tot, 16 * 128, device=dev).bfloat16()
    k = torch.randn(tot, 8 * 128, device=dev).bfloat16()
    starts, tot_pad = ops.vt_layout(lens)
    vt = torch.randn(8, 128, tot_pad, device=dev).bfloat16()
    vts = torch.tensor(starts, dtype=torch.int32, device=dev)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    work = ops.causal_work_list(lens).to(dev)
    out = torch.zeros(tot, 2048, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(q, k, vt, out, head_dim=128, heads=16, kv_heads=8, causal=True,
                                      scale=1 / math.sqrt(128), cu_seqlens=cu, work=work, vt_starts=vts, vt_cols=tot_pad))
    fl = sum(4.0 * 16 * 128 * n * (n + 1) / 2 for n in lens)
    print(f"[perf] llm attention 32x2187: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (causal-exact flops)")

print("ALL_OK" if ok_all else "SOME_FAILED")
