"""Experiment: ViT attention (256 tiles x 16 heads x 1025 tokens, hd 64) under MJV_ATTN_VARIANT; time + accuracy."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops
T, S, H = 256, 1025, 16
dev = "cuda"; torch.manual_seed(0); Cv = H * 64
sc = float(os.environ.get("QK_SCALE", "1.0"))
q, k, v = [torch.randn(T * S, H, 64, device=dev).bfloat16() for _ in range(3)]
q = (q.float() * sc).bfloat16()
qk = torch.cat([q.reshape(T * S, Cv), k.reshape(T * S, Cv)], dim=1).contiguous()
S_pad = (S + 7) // 8 * 8
vt = torch.zeros(T * H, 64, S_pad, device=dev, dtype=torch.bfloat16)
vt[:, :, :S] = v.reshape(T, S, H, 64).permute(0, 2, 3, 1).reshape(T * H, 64, S)
out = torch.zeros(T * S, Cv, device=dev, dtype=torch.bfloat16)
run = lambda: ops.attention(qk, qk, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125, q_col0=0, k_col0=Cv, nseq=T, seq_len=S)
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(40): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 40
Tc = 4
qf, kf, vf = [x.float().view(T, S, H, 64)[:Tc] for x in (q, k, v)]
s = torch.einsum("tqhd,tkhd->thqk", qf, kf) * 0.125
ref = torch.einsum("thqk,tkhd->tqhd", s.softmax(-1), vf).reshape(Tc * S, Cv)
err = (out[: Tc * S].float() - ref).abs()
print(f"variant {os.environ.get('MJV_ATTN_VARIANT','0')} scale {sc}: {ms:.3f} ms  {4*H*64*S*S*T/ms/1e9:.0f} TFLOP/s  max err {err.max().item():.3e} mean {err.mean().item():.3e}")
