"""Single-config causal varlen attention check (one process per config so a fault does not poison the rest)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops  # noqa: E402


def report(name, got, ref, tol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    mx = err.max().item()
    ok = mx <= tol and not torch.isnan(got).any().item()
    print(f"[{name}] max_abs_err={mx:.4e} nan={torch.isnan(got).any().item()} {'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        bad = (err > tol) | torch.isnan(got)
        rows = bad.any(dim=1).nonzero().flatten()
        print(f"   bad rows {rows.numel()}/{bad.shape[0]} first {rows[:12].tolist()} last {rows[-4:].tolist()}")


def ref_attn(q, k, v, causal, scale):
    H, Hkv = q.shape[1], k.shape[1]
    k = k.repeat_interleave(H // Hkv, dim=1)
    v = v.repeat_interleave(H // Hkv, dim=1)
    s = torch.einsum("qhd,khd->hqk", q, k) * scale
    S = q.shape[0]
    s = s.masked_fill(torch.ones(S, S, device=q.device, dtype=torch.bool).triu(1), float("-inf"))
    return torch.einsum("hqk,khd->qhd", s.softmax(-1), v).reshape(q.shape[0], -1)


lens = [int(x) for x in sys.argv[1].split(",")]
H, Hkv = int(sys.argv[2]), int(sys.argv[3])
dev = "cuda"
torch.manual_seed(0)
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
ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=128, heads=H, kv_heads=Hkv,
              causal=True, scale=1 / math.sqrt(128), cu_seqlens=cu, work=work, vt_starts=vts, vt_cols=tot_pad)
torch.cuda.synchronize()
refs, o = [], 0
for n in lens:
    refs.append(ref_attn(q[o:o + n].float(), k[o:o + n].float(), v[o:o + n].float(), True, 1 / math.sqrt(128)))
    o += n
report(f"llm-attn lens{lens} H{H}/{Hkv}", out, torch.cat(refs), 2e-2)
