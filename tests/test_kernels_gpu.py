"""Per-kernel parity on the B200: every CUDA kernel family, called through the C ABI, against a plain PyTorch fp32
reference of the same op that mirrors the reference's bf16 rounding points.
Tolerances: results are bf16, so one bf16 ulp of the output magnitude (2^-8 relative) is the floor."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def ops(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from mj_video_b200 import ops as _ops
    return _ops


def close(got, ref, ulps=2.0):
    got, ref = got.float(), ref.float()
    tol = ulps * 2.0 ** -8 * max(1.0, ref.abs().max().item())
    err = (got - ref).abs().max().item()
    assert not torch.isnan(got).any() and err <= tol, f"max err {err:.4e} > tol {tol:.4e}"


def mm(a, w):
    return a.float() @ w.float().t()


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (128, 256, 64, 256), (256, 512, 1024, 256),
                                      (1000, 384, 576, 128), (4100, 1024, 1024, 256), (333, 256, 640, 256),
                                      (1, 128, 8, 128), (129, 256, 72, 0)])
def test_gemm_store_shapes(ops, M, N, K, bn):
    """Ragged M (TMA zero-fill + row predicate), K not a multiple of 64, both tile widths."""
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()
    out = ops.gemm(a, w, block_n=bn)
    close(out, mm(a, w).bfloat16())


def test_gemm_epilogues(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 1000, 1024, 512
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g).bfloat16()
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    ls = torch.randn(N, device="cuda", generator=g).bfloat16()
    acc = mm(a, w)
    lin = (acc + bias.float()).bfloat16()
    close(ops.gemm(a, w, bias=bias), lin)
    close(ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GELU), torch.nn.functional.gelu(lin.float()).bfloat16())
    t = (lin.float() * ls.float()).bfloat16()
    close(ops.gemm(a, w, bias=bias, residual=res, scale=ls, epilogue=ops.EPI_RESID), (res.float() + t.float()).bfloat16())
    close(ops.gemm(a, w, residual=res, epilogue=ops.EPI_RESID), (res.float() + acc.bfloat16().float()).bfloat16())
    # in-place residual (out aliases residual), as the engine uses it
    buf = res.clone()
    ops.gemm(a, w, out=buf, residual=buf, epilogue=ops.EPI_RESID)
    close(buf, (res.float() + acc.bfloat16().float()).bfloat16())
    # SwiGLU with [32 gate | 32 up] interleaved rows
    F = 512
    w1 = (torch.randn(F, K, device="cuda", generator=g) * 0.1).bfloat16()
    w3 = (torch.randn(F, K, device="cuda", generator=g) * 0.1).bfloat16()
    packed = torch.stack([w1.view(F // 32, 32, K), w3.view(F // 32, 32, K)], dim=1).reshape(2 * F, K).contiguous()
    gate, up = mm(a, w1).bfloat16(), mm(a, w3).bfloat16()
    # silu(g) * u is evaluated on the fp32 accumulators and rounded once: within ~1.5 bf16 ulp of the reference's chain of
    # three bf16 tensors (w1 x, silu(.), w3 x) and within 1 ulp of the fp32 value
    ref_bf16 = torch.nn.functional.silu(gate.float()).bfloat16() * up
    got = ops.gemm(a, packed, epilogue=ops.EPI_SWIGLU)
    close(got, ref_bf16, ulps=3.0)
    close(got, (torch.nn.functional.silu(mm(a, w1)) * mm(a, w3)).bfloat16(), ulps=1.5)
    # scatter (the <IMG_CONTEXT> splice)
    perm = torch.randperm(M, device="cuda", generator=g).int()
    outbuf = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out=outbuf, bias=bias, row_map=perm, epilogue=ops.EPI_SCATTER)
    ref = torch.zeros(M, N, device="cuda")
    ref[perm.long()] = lin.float()
    close(outbuf, ref)
    # patch-embed: out row = m + m/P + 1, + bias + pos[m%P + 1]
    P, T = 256, 3
    a2 = (torch.randn(T * P, K, device="cuda", generator=g) * 0.5).bfloat16()
    pos = torch.randn(P + 1, N, device="cuda", generator=g).bfloat16()
    hid = torch.zeros(T * (P + 1), N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a2, w, out=hid, bias=bias, pos_emb=pos, seq=P, epilogue=ops.EPI_PATCH)
    ref = ((mm(a2, w) + bias.float()).bfloat16().view(T, P, N) + pos[1:]).float()
    close(hid.view(T, P + 1, N)[:, 1:], ref.bfloat16())
    assert hid.view(T, P + 1, N)[:, 0].abs().max() == 0       # CLS rows untouched by the GEMM


def test_gemm_qkv_epilogues(ops):
    g = torch.Generator(device="cuda").manual_seed(2)
    # ViT: +bias, q|k packed, V transposed per (tile, head)
    T, S, H = 2, 1025, 2
    Cv = H * 64
    x = (torch.randn(T * S, Cv, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(3 * Cv, Cv, device="cuda", generator=g) * 0.1).bfloat16()
    b = torch.randn(3 * Cv, device="cuda", generator=g).bfloat16()
    S_pad = (S + 7) // 8 * 8
    qk = torch.zeros(T * S, 2 * Cv, device="cuda", dtype=torch.bfloat16)
    vt = torch.zeros(T * H, 64, S_pad, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, w, out=qk, epilogue=ops.EPI_QKV_VIT, bias=b, out_vt=vt, n_q_cols=Cv, n_k_cols=Cv,
             vt_stride_head=64 * S_pad, vt_ld=S_pad, seq=S, heads=H, block_n=128)
    ref = (mm(x, w) + b.float()).bfloat16()
    close(qk, ref[:, :2 * Cv])
    close(vt[:, :, :S], ref[:, 2 * Cv:].reshape(T, S, H, 64).permute(0, 2, 3, 1).reshape(T * H, 64, S))
    # LLM: rows [q heads | k heads | v heads] x 128, RoPE with bf16 rounding points, V^T at mapped columns
    tot, D, Hq, Hkv = 777, 256, 2, 1
    x = (torch.randn(tot, D, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn((Hq + 2 * Hkv) * 128, D, device="cuda", generator=g) * 0.1).bfloat16()
    pos = (torch.arange(tot, device="cuda") % 400).int()
    inv = 1.0 / (1e6 ** (torch.arange(0, 128, 2, device="cuda").float() / 128))
    fr = torch.outer(torch.arange(512, device="cuda").float(), inv)
    cos, sin = fr.cos().bfloat16().float().contiguous(), fr.sin().bfloat16().float().contiguous()
    row_vt = (torch.arange(tot, device="cuda") + 8 * (torch.arange(tot, device="cuda") // 300)).int()
    vt_cols = int(row_vt.max().item()) // 8 * 8 + 8
    qo = torch.zeros(tot, Hq * 128, device="cuda", dtype=torch.bfloat16)
    ko = torch.zeros(tot, Hkv * 128, device="cuda", dtype=torch.bfloat16)
    vt = torch.zeros(Hkv, 128, vt_cols, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, w, out=qo, epilogue=ops.EPI_QKV_LLM, out_k=ko, out_vt=vt, rope_cos=cos, rope_sin=sin, row_pos=pos,
             row_vt=row_vt, n_q_cols=Hq * 128, n_k_cols=Hkv * 128, vt_stride_head=128 * vt_cols, vt_ld=vt_cols,
             block_n=128)
    y = mm(x, w).bfloat16()

    def rope(t):
        c = torch.cat([cos, cos], -1)[pos.long()].bfloat16()[:, None, :]
        s = torch.cat([sin, sin], -1)[pos.long()].bfloat16()[:, None, :]
        rot = torch.cat([-t[..., 64:], t[..., :64]], -1)
        return (t * c) + (rot * s)

    close(qo, rope(y[:, :Hq * 128].reshape(tot, Hq, 128)).reshape(tot, -1))
    close(ko, rope(y[:, Hq * 128:(Hq + Hkv) * 128].reshape(tot, Hkv, 128)).reshape(tot, -1))
    close(vt[:, :, row_vt.long()], y[:, (Hq + Hkv) * 128:].reshape(tot, Hkv, 128).permute(1, 2, 0))


def test_gemm_qkv_llm_packed_v_heads(ops):
    """QKV_LLM with V heads narrower than the rotary slot (v_head_dim 96): q/k in 128-wide slots with RoPE, V columns packed
    96 per head in N and stored as V^T planes of 96 rows."""
    g = torch.Generator(device="cuda").manual_seed(4)
    tot, D, Hq, Hkv, hv = 500, 256, 4, 4, 96
    N = (Hq + Hkv) * 128 + Hkv * hv          # 1408 = 11 x 128
    x = (torch.randn(tot, D, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, D, device="cuda", generator=g) * 0.1).bfloat16()
    pos = (torch.arange(tot, device="cuda") % 333).int()
    inv = 1.0 / (1e4 ** (torch.arange(0, 128, 2, device="cuda").float() / 128))
    fr = torch.outer(torch.arange(512, device="cuda").float(), inv)
    cos, sin = fr.cos().bfloat16().float().contiguous(), fr.sin().bfloat16().float().contiguous()
    vt_cols = (tot + 7) // 8 * 8
    qo = torch.zeros(tot, Hq * 128, device="cuda", dtype=torch.bfloat16)
    ko = torch.zeros(tot, Hkv * 128, device="cuda", dtype=torch.bfloat16)
    vt = torch.zeros(Hkv, hv, vt_cols, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, w, out=qo, epilogue=ops.EPI_QKV_LLM, out_k=ko, out_vt=vt, rope_cos=cos, rope_sin=sin, row_pos=pos,
             n_q_cols=Hq * 128, n_k_cols=Hkv * 128, vt_stride_head=hv * vt_cols, vt_ld=vt_cols, v_head_dim=hv)
    y = mm(x, w).bfloat16()

    def rope(t):
        c = torch.cat([cos, cos], -1)[pos.long()].bfloat16()[:, None, :]
        s_ = torch.cat([sin, sin], -1)[pos.long()].bfloat16()[:, None, :]
        return (t * c) + (torch.cat([-t[..., 64:], t[..., :64]], -1) * s_)

    close(qo, rope(y[:, :Hq * 128].reshape(tot, Hq, 128)).reshape(tot, -1))
    close(ko, rope(y[:, Hq * 128:(Hq + Hkv) * 128].reshape(tot, Hkv, 128)).reshape(tot, -1))
    close(vt[:, :, :tot], y[:, (Hq + Hkv) * 128:].reshape(tot, Hkv, hv).permute(1, 2, 0))


@pytest.mark.parametrize("block_n", [0, 128])
def test_gemm_qkv_llm_native_head_dim_96(ops, block_n):
    """QKV_LLM with the native head_dim-96 slot (mjv_gemm_desc.qk_head_slot = 96): q / k heads are 96 columns wide, the
    rotary pair (d, d + 48) sits in adjacent columns (2d, 2d + 1) - the weight rows are stored in that order - and no column
    of N is padding. Compared with rotate-half RoPE on the un-permuted projection."""
    g = torch.Generator(device="cuda").manual_seed(5)
    tot, D, Hq, Hkv, hd = 700, 256, 8, 4, 96
    perm = torch.tensor([(j // 2) + 48 * (j % 2) for j in range(hd)], device="cuda")   # slot column j <- head dim perm[j]
    x = (torch.randn(tot, D, device="cuda", generator=g) * 0.5).bfloat16()
    w_ref = (torch.randn((Hq + 2 * Hkv) * hd, D, device="cuda", generator=g) * 0.1).bfloat16()   # [q heads | k heads | v heads]
    wq = w_ref[:Hq * hd].reshape(Hq, hd, D)[:, perm].reshape(Hq * hd, D)
    wk = w_ref[Hq * hd:(Hq + Hkv) * hd].reshape(Hkv, hd, D)[:, perm].reshape(Hkv * hd, D)
    w = torch.cat([wq, wk, w_ref[(Hq + Hkv) * hd:]]).contiguous()      # N = 1536 = 6 x 256
    pos = (torch.arange(tot, device="cuda") % 333).int()
    inv = 1.0 / (1e4 ** (torch.arange(0, hd, 2, device="cuda").float() / hd))
    fr = torch.outer(torch.arange(512, device="cuda").float(), inv)
    cos = torch.ones(512, 64, device="cuda")
    sin = torch.zeros(512, 64, device="cuda")
    cos[:, :48], sin[:, :48] = fr.cos().bfloat16().float(), fr.sin().bfloat16().float()
    vt_cols = (tot + 7) // 8 * 8
    qo = torch.zeros(tot, Hq * hd, device="cuda", dtype=torch.bfloat16)
    ko = torch.zeros(tot, Hkv * hd, device="cuda", dtype=torch.bfloat16)
    vt = torch.zeros(Hkv, hd, vt_cols, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, w, out=qo, epilogue=ops.EPI_QKV_LLM, out_k=ko, out_vt=vt, rope_cos=cos.contiguous(), rope_sin=sin.contiguous(),
             row_pos=pos, n_q_cols=Hq * hd, n_k_cols=Hkv * hd, vt_stride_head=hd * vt_cols, vt_ld=vt_cols, v_head_dim=hd,
             qk_head_slot=96, block_n=block_n)
    y = mm(x, w_ref).bfloat16()

    def rope(t):   # t [tot, heads, 96] in the model's own order
        c = torch.cat([cos[:, :48], cos[:, :48]], -1)[pos.long()].bfloat16()[:, None, :]
        s_ = torch.cat([sin[:, :48], sin[:, :48]], -1)[pos.long()].bfloat16()[:, None, :]
        return (t * c) + (torch.cat([-t[..., 48:], t[..., :48]], -1) * s_)

    close(qo, rope(y[:, :Hq * hd].reshape(tot, Hq, hd))[..., perm].reshape(tot, -1))
    close(ko, rope(y[:, Hq * hd:(Hq + Hkv) * hd].reshape(tot, Hkv, hd))[..., perm].reshape(tot, -1))
    close(vt[:, :, :tot], y[:, (Hq + Hkv) * hd:].reshape(tot, Hkv, hd).permute(1, 2, 0))


def ref_attn(q, k, v, causal, scale):
    H, Hkv = q.shape[1], k.shape[1]
    k = k.repeat_interleave(H // Hkv, dim=1)
    v = v.repeat_interleave(H // Hkv, dim=1)
    s = torch.einsum("qhd,khd->hqk", q, k) * scale
    if causal:
        S = q.shape[0]
        s = s.masked_fill(torch.ones(S, S, device=q.device, dtype=torch.bool).triu(1), float("-inf"))
    return torch.einsum("hqk,khd->qhd", s.softmax(-1), v).reshape(q.shape[0], -1)


@pytest.mark.parametrize("T,S,H", [(1, 1, 1), (1, 128, 1), (2, 129, 2), (1, 257, 2), (3, 1025, 2),
                                   # S % 128 == 1 paths of attention2: last key on the CUDA cores (S > 512), odd last q-blocks
                                   # of two heads sharing an item (even head count, odd block count)
                                   (1, 385, 2), (1, 513, 2), (2, 641, 2), (2, 1025, 3), (5, 1025, 4), (150, 1025, 2),
                                   # one q-block only / three q-blocks with an even head count: every item (or the last) is a head pair
                                   (2, 100, 2), (3, 128, 4), (2, 300, 4), (40, 300, 8)])
def test_attention_bidirectional(ops, T, S, H):
    """ViT shape (hd 64): ragged S incl. the 1-row tail block of S=1025 and the degenerate S=1."""
    g = torch.Generator(device="cuda").manual_seed(S)
    Cv = H * 64
    q, k, v = [torch.randn(T * S, H, 64, device="cuda", generator=g).bfloat16() for _ in range(3)]
    qk = torch.cat([q.reshape(T * S, Cv), k.reshape(T * S, Cv)], dim=1).contiguous()
    S_pad = (S + 7) // 8 * 8
    vt = torch.zeros(T * H, 64, S_pad, device="cuda", dtype=torch.bfloat16)
    vt[:, :, :S] = v.reshape(T, S, H, 64).permute(0, 2, 3, 1).reshape(T * H, 64, S)
    out = torch.zeros(T * S, Cv, device="cuda", dtype=torch.bfloat16)
    ops.attention(qk, qk, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125, q_col0=0, k_col0=Cv,
                  nseq=T, seq_len=S)
    ref = torch.cat([ref_attn(q[t * S:(t + 1) * S].float(), k[t * S:(t + 1) * S].float(), v[t * S:(t + 1) * S].float(),
                              False, 0.125) for t in range(T)])
    assert (out.float() - ref).abs().max().item() < 1e-2


@pytest.mark.parametrize("lens,H,Hkv", [([1], 1, 1), ([128], 2, 1), ([300, 77, 513], 4, 2), ([2187, 2100], 16, 8),
                                         ([63, 64, 65, 1, 129], 2, 2)])
def test_attention_causal_varlen_gqa(ops, lens, H, Hkv):
    """LLM shape (hd 128): packed ragged sequences, GQA, unaligned sequence starts, single-token rows."""
    g = torch.Generator(device="cuda").manual_seed(sum(lens))
    tot = sum(lens)
    q = torch.randn(tot, H, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(tot, Hkv, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(tot, Hkv, 128, device="cuda", generator=g).bfloat16()
    starts, vt_cols = ops.vt_layout(lens)
    vt = torch.zeros(Hkv, 128, vt_cols, device="cuda", dtype=torch.bfloat16)
    o = 0
    for st, n in zip(starts, lens):
        vt[:, :, st:st + n] = v[o:o + n].permute(1, 2, 0)
        o += n
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    out = torch.zeros(tot, H * 128, device="cuda", dtype=torch.bfloat16)
    ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=128, heads=H, kv_heads=Hkv, causal=True,
                  scale=1 / math.sqrt(128), cu_seqlens=cu, work=ops.causal_work_list(lens).cuda(),
                  vt_starts=torch.tensor(starts, dtype=torch.int32, device="cuda"), vt_cols=vt_cols)
    refs, o = [], 0
    for n in lens:
        refs.append(ref_attn(q[o:o + n].float(), k[o:o + n].float(), v[o:o + n].float(), True, 1 / math.sqrt(128)))
        o += n
    assert (out.float() - torch.cat(refs)).abs().max().item() < 2e-2


def test_attention_narrow_heads_packed_v_and_output(ops):
    """head_dim 96 served by the 128-wide tile: q/k in 128-wide slots, V^T planes of 96 rows (TMA zero-fills the rest of
    the box), only 96 output columns per head written, head h at column 96*h (mjv_attn_desc.v_head_dim / out_head_dim)."""
    lens, H, Hkv, hd = [300, 77, 129], 4, 2, 96
    g = torch.Generator(device="cuda").manual_seed(11)
    tot = sum(lens)
    q = torch.zeros(tot, H, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.zeros(tot, Hkv, 128, device="cuda", dtype=torch.bfloat16)
    q[..., :hd] = torch.randn(tot, H, hd, device="cuda", generator=g).bfloat16()
    k[..., :hd] = torch.randn(tot, Hkv, hd, device="cuda", generator=g).bfloat16()
    v = torch.randn(tot, Hkv, hd, device="cuda", generator=g).bfloat16()
    starts, vt_cols = ops.vt_layout(lens)
    vt = torch.zeros(Hkv, hd, vt_cols, device="cuda", dtype=torch.bfloat16)
    o = 0
    for st, n in zip(starts, lens):
        vt[:, :, st:st + n] = v[o:o + n].permute(1, 2, 0)
        o += n
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    out = torch.full((tot, H * hd), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=128, heads=H, kv_heads=Hkv, causal=True,
                  scale=1 / math.sqrt(hd), cu_seqlens=cu, work=ops.causal_work_list(lens).cuda(),
                  vt_starts=torch.tensor(starts, dtype=torch.int32, device="cuda"), vt_cols=vt_cols, v_head_dim=hd,
                  out_head_dim=hd)
    refs, o = [], 0
    for n in lens:
        refs.append(ref_attn(q[o:o + n, :, :hd].float(), k[o:o + n, :, :hd].float(), v[o:o + n].float(), True, 1 / math.sqrt(hd)))
        o += n
    assert (out.float() - torch.cat(refs)).abs().max().item() < 2e-2


@pytest.mark.parametrize("lens,H,Hkv", [([300, 77, 129], 4, 2), ([2187, 640, 1], 8, 8), ([513], 3, 1)])
def test_attention_packed_head_dim_96(ops, lens, H, Hkv):
    """Native head_dim 96 (mjv_attn_desc.qk_head_stride = 96): q / k heads packed 96 columns apart under the 128-wide tile;
    the tile's last 32 columns hold the NEXT head (or TMA zero fill after the last one) and must not enter Q.K^T."""
    hd = 96
    g = torch.Generator(device="cuda").manual_seed(12)
    tot = sum(lens)
    q = torch.randn(tot, H, hd, device="cuda", generator=g).bfloat16()
    k = torch.randn(tot, Hkv, hd, device="cuda", generator=g).bfloat16()
    v = torch.randn(tot, Hkv, hd, device="cuda", generator=g).bfloat16()
    starts, vt_cols = ops.vt_layout(lens)
    vt = torch.zeros(Hkv, hd, vt_cols, device="cuda", dtype=torch.bfloat16)
    o = 0
    for st, n in zip(starts, lens):
        vt[:, :, st:st + n] = v[o:o + n].permute(1, 2, 0)
        o += n
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    out = torch.full((tot, H * hd), 7.0, device="cuda", dtype=torch.bfloat16)
    pair = (H // Hkv) % 2 == 1
    ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=128, heads=H, kv_heads=Hkv, causal=True,
                  scale=1 / math.sqrt(hd), cu_seqlens=cu, work=ops.causal_work_list(lens, pair_blocks=pair).cuda(),
                  vt_starts=torch.tensor(starts, dtype=torch.int32, device="cuda"), vt_cols=vt_cols, v_head_dim=hd,
                  out_head_dim=hd, qk_head_stride=hd)
    refs, o = [], 0
    for n in lens:
        refs.append(ref_attn(q[o:o + n].float(), k[o:o + n].float(), v[o:o + n].float(), True, 1 / math.sqrt(hd)))
        o += n
    assert (out.float() - torch.cat(refs)).abs().max().item() < 2e-2


def test_attention_large_logits_rescale_path(ops):
    """Scores that grow along the sequence force the lazy in-TMEM rescale of the output accumulator."""
    S, H = 1025, 1
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(S, H, 64, device="cuda", generator=g).bfloat16()
    k = (torch.randn(S, H, 64, device="cuda", generator=g) * torch.linspace(0.2, 6.0, S, device="cuda")[:, None, None]).bfloat16()
    v = torch.randn(S, H, 64, device="cuda", generator=g).bfloat16()
    qk = torch.cat([q.reshape(S, 64), k.reshape(S, 64)], dim=1).contiguous()
    vt = torch.zeros(H, 64, 1032, device="cuda", dtype=torch.bfloat16)
    vt[:, :, :S] = v.permute(1, 2, 0)
    out = torch.zeros(S, 64, device="cuda", dtype=torch.bfloat16)
    ops.attention(qk, qk, vt, out, head_dim=64, heads=1, kv_heads=1, causal=False, scale=0.125, k_col0=64, nseq=1,
                  seq_len=S)
    ref = ref_attn(q.float(), k.float(), v.float(), False, 0.125)
    assert (out.float() - ref).abs().max().item() < 3e-2


def test_rowwise_kernels(ops, built_lib):
    from mj_video_b200 import _lib, synth
    from oracle import port
    lib = built_lib
    g = torch.Generator(device="cuda").manual_seed(3)
    st = _lib.current_stream()
    for rows, Cc in [(1000, 1024), (77, 128), (300, 4096), (5, 256), (200, 3072), (123, 2048), (64, 1536), (9, 2304)]:
        x = torch.randn(rows, Cc, device="cuda", generator=g).bfloat16()
        w = torch.randn(Cc, device="cuda", generator=g).bfloat16()
        b = torch.randn(Cc, device="cuda", generator=g).bfloat16()
        out = torch.empty_like(x)
        _lib.check(lib.mjv_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), rows, Cc, 1e-6, st))
        close(out, torch.nn.functional.layer_norm(x.float(), (Cc,), w.float(), b.float(), 1e-6).bfloat16(), ulps=1.5)
        _lib.check(lib.mjv_rmsnorm(x.data_ptr(), w.data_ptr(), out.data_ptr(), rows, Cc, 1e-5, None, st))
        xf = x.float()
        ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).bfloat16()
        close(out, ref, ulps=1.5)
    # gathered rmsnorm (final norm on 2 rows per sample)
    x = torch.randn(500, 2048, device="cuda", generator=g).bfloat16()
    w = torch.randn(2048, device="cuda", generator=g).bfloat16()
    idx = torch.tensor([3, 499, 0, 250], dtype=torch.int32, device="cuda")
    out = torch.empty(4, 2048, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.mjv_rmsnorm(x.data_ptr(), w.data_ptr(), out.data_ptr(), 4, 2048, 1e-5, idx.data_ptr(), st))
    xf = x[idx.long()].float()
    close(out, w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).bfloat16(), ulps=1.5)
    # pixel-shuffle + LayerNorm against the oracle's own pixel_shuffle (modeling_internvl_chat.py:228-262)
    T, S, Cv = 3, 1025, 128
    hid = torch.randn(T, S, Cv, device="cuda", generator=g).bfloat16()
    w = torch.randn(4 * Cv, device="cuda", generator=g).bfloat16()
    b = torch.randn(4 * Cv, device="cuda", generator=g).bfloat16()
    out = torch.empty(T * 256, 4 * Cv, device="cuda", dtype=torch.bfloat16)
    import types
    for ps_version in ("v2", "v1"):   # 'v1' = the reference's layout without the final permute (:237-239)
        _lib.check(lib.mjv_pixel_shuffle_layernorm(hid.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), T, S, Cv,
                                                   1e-5, int(ps_version == "v1"), st))
        orc = port.Oracle.__new__(port.Oracle)
        orc.d = types.SimpleNamespace(ps_version=ps_version)
        ps = orc.pixel_shuffle(hid[:, 1:].float().reshape(T, 32, 32, Cv), 0.5).reshape(T * 256, 4 * Cv)
        close(out, torch.nn.functional.layer_norm(ps, (4 * Cv,), w.float(), b.float(), 1e-5).bfloat16(), ulps=1.5)
    # embedding gather skips <IMG_CONTEXT> rows
    table = torch.randn(1000, 256, device="cuda", generator=g).bfloat16()
    ids = torch.tensor([5, 999, 7, 7, 0], dtype=torch.int32, device="cuda")
    out = torch.full((5, 256), 9.0, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.mjv_embed_tokens(ids.data_ptr(), table.data_ptr(), out.data_ptr(), 5, 256, 7, 1000, st))
    assert torch.equal(out[[0, 1, 4]], table[[5, 999, 0]]) and (out[2:4] == 9.0).all()
    # im2col == unfold of the conv operand (k = c*196 + ky*14 + kx), zero padded to 640
    px = torch.randn(2, 3, 448, 448, device="cuda", generator=g)
    for t, code in ((px, _lib.MJV_DTYPE_F32), (px.bfloat16(), _lib.MJV_DTYPE_BF16)):
        out = torch.full((2 * 1024, 640), 7.0, device="cuda", dtype=torch.bfloat16)
        _lib.check(lib.mjv_im2col_patches(t.contiguous().data_ptr(), code, out.data_ptr(), 2, 448, 14, 640, st))
        ref = torch.nn.functional.unfold(t.float(), 14, stride=14).transpose(1, 2).reshape(2 * 1024, 588).bfloat16()
        assert torch.equal(out[:, :588], ref) and (out[:, 588:] == 0).all()


def test_attention_is_deterministic_under_stress(ops):
    """Race detector for the barrier pipeline (warps with masked / absent rows must not lap the others, the final
    hand-over must not alias barrier phases): identical launches give bit-identical results."""
    g = torch.Generator(device="cuda").manual_seed(9)
    lens = [2187, 2100, 130, 64]
    H, Hkv = 16, 8
    tot = sum(lens)
    q = torch.randn(tot, H, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(tot, Hkv, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(tot, Hkv, 128, device="cuda", generator=g).bfloat16()
    starts, cols = ops.vt_layout(lens)
    vt = torch.zeros(Hkv, 128, cols, device="cuda", dtype=torch.bfloat16)
    o = 0
    for st, n in zip(starts, lens):
        vt[:, :, st:st + n] = v[o:o + n].permute(1, 2, 0)
        o += n
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    work = ops.causal_work_list(lens).cuda()
    vts = torch.tensor(starts, dtype=torch.int32, device="cuda")
    outs = []
    for _ in range(25):
        out = torch.full((tot, H * 128), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=128, heads=H, kv_heads=Hkv, causal=True,
                      scale=1 / math.sqrt(128), cu_seqlens=cu, work=work, vt_starts=vts, vt_cols=cols)
        outs.append(out)
    torch.cuda.synchronize()
    assert not torch.isnan(outs[0].float()).any()
    assert all(torch.equal(outs[0], x) for x in outs[1:])
    # ViT shape with a partially filled last query block
    T, S, Hh = 8, 1025, 4
    qk = torch.randn(T * S, 2 * Hh * 64, device="cuda", generator=g).bfloat16()
    vt = torch.randn(T * Hh, 64, 1032, device="cuda", generator=g).bfloat16()
    outs = []
    for _ in range(25):
        out = torch.full((T * S, Hh * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.attention(qk, qk, vt, out, head_dim=64, heads=Hh, kv_heads=Hh, causal=False, scale=0.125, q_col0=0,
                      k_col0=Hh * 64, nseq=T, seq_len=S)
        outs.append(out)
    torch.cuda.synchronize()
    assert not torch.isnan(outs[0].float()).any()
    assert all(torch.equal(outs[0], x) for x in outs[1:])


@pytest.mark.parametrize("seed", range(6))
def test_gemm_fuzz_shapes(ops, seed):
    """Random M (ragged), N (multiples of 128), K (multiples of 8) through the STORE and in-place RESID epilogues, both
    tile widths where N allows."""
    import random
    rnd = random.Random(seed)
    M = rnd.choice([1, 7, 127, 128, 129, 255, 257, 1000, 3333, 5001])
    N = 128 * rnd.randint(1, 12)
    K = 8 * rnd.randint(1, 300)
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * (1.0 / math.sqrt(K))).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g).bfloat16()
    ref = mm(a, w)
    for bn in ([128, 256] if N % 256 == 0 else [128]):
        close(ops.gemm(a, w, bias=bias, block_n=bn), (ref + bias.float()).bfloat16())
        res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
        buf = res.clone()
        ops.gemm(a, w, out=buf, residual=buf, epilogue=ops.EPI_RESID, block_n=bn)
        close(buf, (res.float() + ref.bfloat16().float()).bfloat16())


@pytest.mark.parametrize("seed", range(4))
def test_attention_fuzz_lengths(ops, seed):
    """Random ragged sequence sets around the 64 / 128 block boundaries, causal GQA (hd 128) and bidirectional (hd 64)."""
    import random
    rnd = random.Random(100 + seed)
    lens = [rnd.choice([1, 2, 63, 64, 65, 127, 128, 129, 191, 192, 193, 500, 1025]) for _ in range(rnd.randint(1, 6))]
    g = torch.Generator(device="cuda").manual_seed(seed)
    tot, H, Hkv = sum(lens), 4, 2
    q = torch.randn(tot, H, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(tot, Hkv, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(tot, Hkv, 128, device="cuda", generator=g).bfloat16()
    starts, vt_cols = ops.vt_layout(lens)
    vt = torch.zeros(Hkv, 128, vt_cols, device="cuda", dtype=torch.bfloat16)
    o = 0
    for st, n in zip(starts, lens):
        vt[:, :, st:st + n] = v[o:o + n].permute(1, 2, 0)
        o += n
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    out = torch.zeros(tot, H * 128, device="cuda", dtype=torch.bfloat16)
    ops.attention(q.reshape(tot, -1), k.reshape(tot, -1), vt, out, head_dim=128, heads=H, kv_heads=Hkv, causal=True,
                  scale=1 / math.sqrt(128), cu_seqlens=cu, work=ops.causal_work_list(lens).cuda(),
                  vt_starts=torch.tensor(starts, dtype=torch.int32, device="cuda"), vt_cols=vt_cols)
    refs, o = [], 0
    for n in lens:
        refs.append(ref_attn(q[o:o + n].float(), k[o:o + n].float(), v[o:o + n].float(), True, 1 / math.sqrt(128)))
        o += n
    assert (out.float() - torch.cat(refs)).abs().max().item() < 2e-2
    # bidirectional, fixed length per "tile"
    S, T, Hh = rnd.choice([1, 64, 65, 200, 1025]), rnd.randint(1, 3), 2
    Cv = Hh * 64
    qb, kb, vb = [torch.randn(T * S, Hh, 64, device="cuda", generator=g).bfloat16() for _ in range(3)]
    qk = torch.cat([qb.reshape(T * S, Cv), kb.reshape(T * S, Cv)], dim=1).contiguous()
    S_pad = (S + 7) // 8 * 8
    vtb = torch.zeros(T * Hh, 64, S_pad, device="cuda", dtype=torch.bfloat16)
    vtb[:, :, :S] = vb.reshape(T, S, Hh, 64).permute(0, 2, 3, 1).reshape(T * Hh, 64, S)
    outb = torch.zeros(T * S, Cv, device="cuda", dtype=torch.bfloat16)
    ops.attention(qk, qk, vtb, outb, head_dim=64, heads=Hh, kv_heads=Hh, causal=False, scale=0.125, q_col0=0, k_col0=Cv,
                  nseq=T, seq_len=S)
    refb = torch.cat([ref_attn(qb.view(T, S, Hh, 64)[t].float(), kb.view(T, S, Hh, 64)[t].float(),
                               vb.view(T, S, Hh, 64)[t].float(), False, 0.125) for t in range(T)])
    assert (outb.float() - refb).abs().max().item() < 2e-2
