"""Pins the oracle against the reference ITSELF, imported unmodified from /root/reference (build container only)."""
import pytest
import torch

from conftest import reference_available
from mj_video_b200 import synth
from oracle import port

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is not present on this machine")


@pytest.fixture(scope="module")
def ref_model():
    from oracle import reference_harness as rh
    d = synth.dims_tiny()
    sd = synth.make_state_dict(d, seed=3)
    return d, sd, rh.build_reference_model(d, sd), rh


def test_forward_bit_exact_fp32(ref_model):
    d, sd, m, rh = ref_model
    vids = [synth.make_video(3), synth.make_video(4)]
    pv = torch.cat([rh.reference_load_video(v, num_segments=8, max_num=1)[0] for v in vids])
    ids, mask = synth.pad_batch([synth.make_input_ids(0, [1] * 8, n_cap=46), synth.make_input_ids(1, [1] * 8, n_cap=20)])
    ref = rh.reference_forward(m, pv, ids, mask)
    out = port.Oracle(d, sd).forward(pv, ids, mask)
    for k in ref:
        assert torch.equal(ref[k].float(), out[k].float()), k


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_llama_branch_bit_exact(dtype):
    """The reference's Llama LLM branch (modeling_internvl_chat.py:125-126, transformers LlamaForCausalLM) with
    head_dim 96 and grouped KV heads; in bf16 the reference rounds the rotary inv_freq buffer too."""
    from oracle import reference_harness as rh
    d = synth.dims_tiny_llama()
    sd = synth.make_state_dict(d, seed=5)
    m = rh.build_reference_model(d, sd, dtype=dtype)
    pv = torch.cat([rh.reference_load_video(synth.make_video(v), num_segments=8, max_num=1)[0] for v in (3, 4)]).to(dtype)
    ids, mask = synth.pad_batch([synth.make_input_ids(0, [1] * 8, n_cap=46), synth.make_input_ids(1, [1] * 8, n_cap=20)])
    ref = rh.reference_forward(m, pv, ids, mask)
    out = port.Oracle(d, sd, dtype=dtype).forward(pv, ids, mask)
    for k in ref:
        assert torch.equal(ref[k].float(), out[k].float()), k


def test_forward_bit_exact_bf16():
    """The reference's own `.to(bfloat16)` forward (what its GPU runs compute, here on the CPU): the oracle's bf16 mode
    reproduces every rounding point — values AND output dtypes (bf16 tensors, fp32 aspect_scores / score: SURVEY §A.3)."""
    from oracle import reference_harness as rh
    d = synth.dims_tiny()
    sd = synth.make_state_dict(d, seed=3)
    m = rh.build_reference_model(d, sd, dtype=torch.bfloat16)
    pv = torch.cat([rh.reference_load_video(synth.make_video(v), num_segments=8, max_num=1)[0] for v in (3, 4)]).to(torch.bfloat16)
    ids, mask = synth.pad_batch([synth.make_input_ids(0, [1] * 8, n_cap=46), synth.make_input_ids(1, [1] * 8, n_cap=20)])
    ref = rh.reference_forward(m, pv, ids, mask)
    out = port.Oracle(d, sd, dtype=torch.bfloat16).forward(pv, ids, mask)
    assert set(ref) == set(out)
    for k in ref:
        assert ref[k].dtype == out[k].dtype, k
        assert torch.equal(ref[k].float(), out[k].float()), k


def test_first_slot_layout(ref_model):
    """All IMG_CONTEXT tokens in the first <image> slot (what the reference's own callers produce)."""
    d, sd, m, rh = ref_model
    pv = rh.reference_load_video(synth.make_video(9), num_segments=8, max_num=1)[0]
    ids, mask = synth.pad_batch([synth.make_input_ids(5, [1] * 8, n_cap=30, first_slot_layout=True)])
    ref = rh.reference_forward(m, pv, ids, mask)
    out = port.Oracle(d, sd).forward(pv, ids, mask)
    for k in ref:
        assert torch.equal(ref[k].float(), out[k].float()), k


def test_text_only_row_and_surplus_tiles(ref_model, capsys):
    """A batch whose second row has no <image> at all, and a call with one tile more than the prompt has room for (the
    reference prints a warning and keeps the first n_token visual embeddings, modeling_internvl_chat.py:178-186)."""
    d, sd, m, rh = ref_model
    pv = rh.reference_load_video(synth.make_video(21), num_segments=8, max_num=1)[0]
    text_only = [1, 92543] + list(range(2000, 2040)) + synth.GATING_PATTERN
    ids, mask = synth.pad_batch([synth.make_input_ids(2, [1] * 8, n_cap=12), text_only])
    ref = rh.reference_forward(m, pv, ids, mask)
    out = port.Oracle(d, sd).forward(pv, ids, mask)
    for k in ref:
        assert torch.equal(ref[k].float(), out[k].float()), k
    pv9 = torch.cat([pv, pv[:1] * 0.5])
    ids1, mask1 = synth.pad_batch([synth.make_input_ids(2, [1] * 8, n_cap=12)])
    ref = rh.reference_forward(m, pv9, ids1, mask1)
    assert "warning" in capsys.readouterr().out
    out = port.Oracle(d, sd).forward(pv9, ids1, mask1)
    for k in ref:
        assert torch.equal(ref[k].float(), out[k].float()), k


def test_select_layer_and_ps_version_v1():
    """Config knobs of the backbone: the projector reads hidden_states[select_layer] of the ViT
    (modeling_internvl_chat.py:244-254) and ps_version 'v1' skips the final permute of pixel_shuffle (:237-241)."""
    from oracle import reference_harness as rh
    d = synth.dims_tiny()
    d.vit_layers, d.select_layer, d.ps_version = 3, -2, "v1"
    sd = synth.make_state_dict(d, seed=9)
    m = rh.build_reference_model(d, sd)
    pv = rh.reference_load_video(synth.make_video(33), num_segments=8, max_num=1)[0]
    ids, mask = synth.pad_batch([synth.make_input_ids(0, [1] * 8, n_cap=25)])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = rh.reference_forward(m, pv, ids, mask)
    out = port.Oracle(d, sd).forward(pv, ids, mask)
    for k in ref:
        assert torch.equal(ref[k].float(), out[k].float()), k


def test_linear_rope_scaling():
    """rope_scaling {'type': 'linear'}: positions divided by the factor (modeling_internlm2.py:183-199)."""
    from oracle import reference_harness as rh
    d = synth.dims_tiny()
    d.rope_scaling_type, d.rope_scaling_factor = "linear", 4.0
    sd = synth.make_state_dict(d, seed=10)
    m = rh.build_reference_model(d, sd)
    pv = rh.reference_load_video(synth.make_video(34), num_segments=8, max_num=1)[0][:2]
    ids, mask = synth.pad_batch([synth.make_input_ids(0, [1] * 2, n_cap=25)])
    ref = rh.reference_forward(m, pv, ids, mask)
    out = port.Oracle(d, sd).forward(pv, ids, mask)
    for k in ref:
        assert torch.equal(ref[k].float(), out[k].float()), k


def test_dynamic_ntk_rope_beyond_max_positions():
    """A padded batch longer than max_position_embeddings makes InternLM2DynamicNTKScalingRotaryEmbedding re-base theta
    for that length and keep the re-based cache for later (shorter) calls (modeling_internlm2.py:172-175, 210-231)."""
    from oracle import reference_harness as rh
    d = synth.dims_tiny()
    d.max_pos = 1024
    sd = synth.make_state_dict(d, seed=8)
    m = rh.build_reference_model(d, sd)
    orc = port.Oracle(d, sd)
    pv = rh.reference_load_video(synth.make_video(31), num_segments=8, max_num=1)[0]
    long_ids, long_mask = synth.pad_batch([synth.make_input_ids(0, [1] * 8, n_cap=40)])          # 2181 tokens > 1024
    short_ids, short_mask = synth.pad_batch([synth.make_input_ids(1, [1] * 2, n_cap=20)])         # 599 tokens
    for ids, mask, px in ((short_ids, short_mask, pv[:2]), (long_ids, long_mask, pv), (short_ids, short_mask, pv[:2])):
        ref = rh.reference_forward(m, px, ids, mask)
        out = orc.forward(px, ids, mask)
        for k in ref:
            assert torch.equal(ref[k].float(), out[k].float()), (k, ids.shape)
    assert orc.rope_cached == long_ids.shape[1]


def test_reference_errors(ref_model):
    d, sd, m, rh = ref_model
    pv = torch.zeros(8, 3, 448, 448)
    ids, mask = synth.pad_batch([synth.make_input_ids(0, [1] * 8, n_cap=5)])
    bad = ids.clone()
    bad[0, -5:] = 7  # destroy the gating pattern
    with pytest.raises(ValueError, match="Token pattern not found"):
        rh.reference_forward(m, pv, bad, mask)
    with pytest.raises(ValueError, match="Token pattern not found"):
        port.Oracle(d, sd).forward(pv, bad, mask)


@pytest.mark.parametrize("geom", [(896, 1344, 6), (1024, 1024, 6), (720, 1280, 6), (1280, 720, 6), (320, 512, 6),
                                  (100, 77, 1), (500, 300, 4), (448, 448, 1)])
def test_load_video_bit_exact(geom):
    from oracle import reference_harness as rh
    H, W, mx = geom
    v = synth.make_video(7, n_frames=2, height=H, width=W)
    a, c1 = rh.reference_load_video(v, num_segments=2, max_num=mx)
    b, c2 = port.load_video_frames(v, num_segments=2, max_num=mx)
    assert c1 == c2 and torch.equal(a, b)


def test_get_index_matches():
    from oracle import reference_harness as rh
    mod = rh.reference_data_module()
    for (bound, fps, mf, ns) in [(None, 8.0, 15, 8), (None, 30.0, 299, 32), ((1.0, 4.5), 24.0, 200, 16), (None, 8.0, 7, 8)]:
        assert (mod.get_index(bound, fps, mf, 0, ns) == port.get_index(bound, fps, mf, 0, ns)).all()


def test_prepare_chat_input_matches_reference_with_history():
    """The drop-in `prepare_chat_input` (modeling_internvl_chat.py:36-89) token for token, including multi-turn history,
    per-frame tile counts and the mutated generation_config, under the same stub tokenizer."""
    from oracle import reference_harness as rh
    from oracle.make_golden import StubTokenizer
    from mj_video_b200.model import InternVLChatRewardModelingConfig, prepare_chat_input
    _, _, chat = rh.reference_modules()
    d = synth.dims_tiny()
    cfg_ref = rh.build_reference_config(d)
    cfg_mine = InternVLChatRewardModelingConfig(
        vision_config=cfg_ref.vision_config.to_dict(), llm_config=cfg_ref.llm_config.to_dict(), template="internlm2-chat",
        downsample_ratio=0.5, ps_version="v2", select_layer=-1, force_image_size=448, num_objectives=28, num_aspects=5,
        aspect2criteria=d.aspect2criteria, gating_temperature=1.0, gating_hidden_dim=d.gating_hidden, gating_n_hidden=3)
    tok = StubTokenizer()
    pv = torch.zeros(8, 3, 448, 448)
    cases = [dict(question="<image>\nhello", history=[("<image>\nfirst q", "first a")], npl=[4, 4]),
             dict(question="what?", history=[("q1", "a1"), ("q2", "a2")], npl=None),
             dict(question="Frame1: <image>\nFrame2: <image>\ncap", history=None, npl=[3, 5], return_history=True)]
    for c in cases:
        g1, g2 = {}, {}
        kw = dict(history=c["history"], return_history=c.get("return_history", False), num_patches_list=c["npl"])
        ref = chat.prepare_chat_input(cfg_ref, tok, pv if c["npl"] else None, c["question"], g1, **kw)
        got = prepare_chat_input(cfg_mine, tok, pv if c["npl"] else None, c["question"], g2, **kw)
        assert len(ref) == len(got) and g1 == g2
        for a, b in zip(ref, got):
            assert torch.equal(a, b) if torch.is_tensor(a) else a == b


def test_collate_pairs_matches_reference_video_data_collator():
    """`pipeline.collate_pairs` against the unmodified `VideoDataCollator` (dataset.py:407-554) on ragged
    rows. The reference pads every row to max_length (3072) and stacks [B, F, 3, 448, 448] per video side, which its train
    loop flattens to [B*F, 3, 448, 448] (overall_train.py:70-72); the pair scorer interleaves the two sides row by row
    (chosen, rejected, chosen, ...) and pads to the batch maximum. Same tokens, same masks, same frame order; the only
    difference is the amount of trailing padding, which the engine removes anyway (rows are packed)."""
    import numpy as np
    from mj_video_b200.pipeline import collate_pairs
    from oracle import reference_harness as rh
    ds_mod = rh.reference_dataset_module()

    class Tok:
        pad_token_id = 2

    rng = np.random.default_rng(5)
    n_pairs, F = 5, 2
    rows = {}
    for k in range(n_pairs):
        for side in (0, 1):
            n = int(rng.integers(30, 400))
            ids = rng.integers(1000, 90000, size=n)
            frames = rng.integers(0, 256, size=(F, 6 + k, 8, 3), dtype=np.uint8)   # per-pair geometry
            rows[(k, side)] = (frames, ids.tolist())

    class Pairs:
        def __len__(self):
            return n_pairs

        def __getitem__(self, k):
            return rows[(k, 0)], rows[(k, 1)]

    hb = collate_pairs(Pairs(), list(range(n_pairs)), pad_token_id=2, pin=False)

    def side_dict(k, s):
        frames, ids = rows[(k, s)]
        t = torch.tensor(ids, dtype=torch.long)[None]
        # the reference stacks pixel_values; a [F,3,h,w] float view of the same frames stands in for load_video's output
        pv = torch.from_numpy(frames).permute(0, 3, 1, 2).float()
        return dict(input_ids=t, attention_mask=torch.ones_like(t), pixel_values=pv, criteria_score=[0.0], criteria_related=[0],
                    aspect_score=[0.0], aspect_related=[0], total_score=[0.0], total_related=[0])

    for max_length in (3072, None):
        coll = ds_mod.VideoDataCollator(Tok(), max_length=max_length)
        for k in range(n_pairs):     # per pair (the per-pair geometry cannot be stacked across pairs by the reference)
            ref = coll([dict(video_0_dict=side_dict(k, 0), video_1_dict=side_dict(k, 1),
                             preference_dict=dict(aspect_preference=[0], aspect_mask=[0], overall_preference=[0],
                                                  overall_mask=[0]))])
            for s in (0, 1):
                r = 2 * k + s
                rid, rmask = ref[f"video_{s}_input_ids"][0], ref[f"video_{s}_attention_mask"][0]
                n = int(rmask.sum())
                assert n == len(rows[(k, s)][1])
                # identical tokens and mask on the real part, pad id / zero mask on every padded position of both
                assert torch.equal(hb.input_ids[r, :n], rid[:n]) and torch.equal(hb.attention_mask[r, :n], rmask[:n])
                assert bool((rid[n:] == 2).all()) and bool((rmask[n:] == 0).all())
                assert bool((hb.input_ids[r, n:] == 2).all()) and bool((hb.attention_mask[r, n:] == 0).all())
                # frames of the row, in the order the reference's flattened [B*F, 3, H, W] stack has them
                flat = ref[f"video_{s}_pixel_values"].view(-1, 3, 6 + k, 8)
                assert torch.equal(hb.frames[r].permute(0, 3, 1, 2).float(), flat)
    assert hb.input_ids.shape[1] == max(len(v[1]) for v in rows.values()) and hb.pair_ids == list(range(n_pairs))
