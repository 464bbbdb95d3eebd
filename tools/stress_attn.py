"""Determinism/race stress: run the same attention launch many times, every output must be bit-identical and correct."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops
dev = "cuda"; torch.manual_seed(1)
def run_llm(lens, H, Hkv, reps):
    tot = sum(lens)
    q = torch.randn(tot, H, 128, device=dev).bfloat16(); k = torch.randn(tot, Hkv, 128, device=dev).bfloat16(); v = torch.randn(tot, Hkv, 128, device=dev).bfloat16()
    starts, cols = ops.vt_layout(lens)
    vt = torch.zeros(Hkv, 128, cols, device=dev, dtype=torch.bfloat16); o = 0
    for st, n in zip(starts, lens):
        vt[:, :, st:st + n] = v[o:o + n].permute(1, 2, 0); o += n
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    work = ops.causal_work_list(lens).to(dev); vts = torch.tensor(starts, dtype=torch.int32, device=dev)
    outs = []
    for r in range(reps):
        out = torch.full((tot, H * 128), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=128, heads=H, kv_heads=Hkv, causal=True, scale=1 / math.sqrt(128), cu_seqlens=cu, work=work, vt_starts=vts, vt_cols=cols)
        outs.append(out)
    torch.cuda.synchronize()
    ndiff = sum(int(not torch.equal(outs[0], o)) for o in outs[1:])
    nnan = sum(int(torch.isnan(o.float()).any()) for o in outs)
    worst = max(((o.float() - outs[0].float()).abs().nan_to_num(1e9).max().item()) for o in outs)
    print(f"llm lens{lens[:3]}..x{len(lens)} H{H}/{Hkv}: {ndiff}/{reps-1} runs differ, {nnan} with NaN, worst dev {worst:.3e}", flush=True)
def run_vit(T, S, H, reps):
    Cv = H * 64
    qk = torch.randn(T * S, 2 * Cv, device=dev).bfloat16(); S_pad = (S + 7) // 8 * 8
    vt = torch.randn(T * H, 64, S_pad, device=dev).bfloat16()
    outs = []
    for r in range(reps):
        out = torch.full((T * S, Cv), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.attention(qk, qk, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125, q_col0=0, k_col0=Cv, nseq=T, seq_len=S)
        outs.append(out)
    torch.cuda.synchronize()
    ndiff = sum(int(not torch.equal(outs[0], o)) for o in outs[1:]); nnan = sum(int(torch.isnan(o.float()).any()) for o in outs)
    print(f"vit T{T} S{S} H{H}: {ndiff}/{reps-1} runs differ, {nnan} with NaN", flush=True)
run_llm([2187, 2100], 16, 8, 40)
run_llm([300, 77, 513], 4, 2, 40)
run_llm([2201] * 32, 16, 8, 10)
run_vit(3, 1025, 2, 40)
run_vit(64, 1025, 16, 10)
