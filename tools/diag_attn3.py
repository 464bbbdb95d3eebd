import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops
T, S, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = "cuda"; torch.manual_seed(0); Cv = H * 64
q, k, v = [torch.randn(T * S, H, 64, device=dev).bfloat16() for _ in range(3)]
qk = torch.cat([q.reshape(T * S, Cv), k.reshape(T * S, Cv)], dim=1).contiguous()
S_pad = (S + 7) // 8 * 8
vt = torch.zeros(T * H, 64, S_pad, device=dev, dtype=torch.bfloat16)
vt[:, :, :S] = v.reshape(T, S, H, 64).permute(0, 2, 3, 1).reshape(T * H, 64, S)
out = torch.zeros(T * S, Cv, device=dev, dtype=torch.bfloat16)
ops.attention(qk, qk, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125, q_col0=0, k_col0=Cv, nseq=T, seq_len=S)
torch.cuda.synchronize()
qf, kf, vf = q.float().view(T, S, H, 64), k.float().view(T, S, H, 64), v.float().view(T, S, H, 64)
s = torch.einsum("tqhd,tkhd->thqk", qf, kf) * 0.125
ref = torch.einsum("thqk,tkhd->tqhd", s.softmax(-1), vf).reshape(T * S, Cv)
err = (out.float() - ref).abs()
bad = ~(err < 2e-2)
rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
print(f"T{T} S{S} H{H}: max {err.max().item():.3e} bad rows {rows.numel()} {rows[:10].tolist()}..{rows[-3:].tolist()} bad cols {cols.numel()} {cols[:6].tolist()}")
if rows.numel():
    r = rows[0].item(); print(" got", out[r, :6].tolist(), "ref", ref[r, :6].tolist())
