"""attention2 against the first-generation kernel on EVERY output row of the bench shapes (tools only): the A/B bench checks a
few sequences against an fp32 reference; this catches an error that only shows after many items per CTA."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def run_vit(T=256, S=1025, H=16):
    Cv = H * 64
    qkv = torch.randn(T * S, 3 * Cv, device=dev).bfloat16()
    S_pad = (S + 7) // 8 * 8
    vt = torch.zeros(T * H, 64, S_pad, device=dev, dtype=torch.bfloat16)
    vt[:, :, :S] = qkv[:, 2 * Cv:].reshape(T, S, H, 64).permute(0, 2, 3, 1).reshape(T * H, 64, S)
    outs = []
    for impl in ("2", "1"):
        os.environ["MJV_ATTN_IMPL"] = impl
        out = torch.full((T * S, Cv), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.attention(qkv, qkv, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125, q_col0=0, k_col0=Cv,
                      nseq=T, seq_len=S)
        torch.cuda.synchronize()
        outs.append(out.float())
    d = (outs[0] - outs[1]).abs()
    d[torch.isnan(d)] = 1e9
    row = d.max(dim=1).values
    bad = torch.nonzero(row > 2e-2).flatten()
    print(f"vit T{T} S{S} H{H}: max diff {d.max().item():.3e}, rows over 2e-2: {bad.numel()}",
          [(int(b) // S, int(b) % S) for b in bad[:10]], flush=True)
    if bad.numel():
        cols = torch.nonzero(d[bad[0]] > 2e-2).flatten()
        print("  first bad row: heads", sorted(set((cols // 64).tolist()))[:16])


def run_llm(lens, H=16, Hkv=8, hd=128):
    tot = sum(lens)
    q = torch.randn(tot, H * hd, device=dev).bfloat16()
    k = torch.randn(tot, Hkv * hd, device=dev).bfloat16()
    v = torch.randn(tot, Hkv, hd, device=dev).bfloat16()
    starts, cols = ops.vt_layout(lens)
    vt = torch.zeros(Hkv, hd, cols, device=dev, dtype=torch.bfloat16)
    o = 0
    for st, n in zip(starts, lens):
        vt[:, :, st:st + n] = v[o:o + n].permute(1, 2, 0)
        o += n
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    vts = torch.tensor(starts, dtype=torch.int32, device=dev)
    work = ops.causal_work_list(lens).to(dev)
    outs = []
    for impl in ("2", "1"):
        os.environ["MJV_ATTN_IMPL"] = impl
        out = torch.full((tot, H * hd), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.attention(q, k, vt, out, head_dim=hd, heads=H, kv_heads=Hkv, causal=True, scale=1 / math.sqrt(hd), cu_seqlens=cu,
                      work=work, vt_starts=vts, vt_cols=cols)
        torch.cuda.synchronize()
        outs.append(out.float())
    d = (outs[0] - outs[1]).abs()
    d[torch.isnan(d)] = 1e9
    bad = torch.nonzero(d.max(dim=1).values > 3e-2).flatten()
    print(f"llm {len(lens)} rows: max diff {d.max().item():.3e}, rows over 3e-2: {bad.numel()}", bad[:10].tolist(), flush=True)


if __name__ == "__main__":
    run_vit(8, 1025, 16)
    run_vit(64, 1025, 16)
    run_vit(256, 1025, 16)
    run_vit(256, 1025, 16)
    g = torch.Generator().manual_seed(1)
    run_llm((2141 + torch.randint(16, 97, (32,), generator=g)).tolist())
    os.environ.pop("MJV_ATTN_IMPL", None)
