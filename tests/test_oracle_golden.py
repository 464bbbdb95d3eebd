"""The oracle (oracle/port.py) against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py). Runs everywhere, including the GPU box where /root/reference does not exist."""
import hashlib

import numpy as np
import pytest
import torch

from mj_video_b200 import synth
from oracle import port


def _sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def _inputs(fix):
    pv = torch.cat([port.load_video_frames(synth.make_video(v), num_segments=8, max_num=1)[0] for v in fix["video_ids"]])
    rows = [synth.make_input_ids(i, [1] * 8, n_cap=n) for i, n in enumerate(fix["n_caps"])]
    ids, mask = synth.pad_batch(rows)
    return pv, ids, mask, rows


def test_forward_matches_reference_golden(golden_forward):
    fix = golden_forward
    d = synth.dims_tiny()
    sd = synth.make_state_dict(d, seed=fix["seed"])
    pv, ids, mask, rows = _inputs(fix)
    assert _sha(pv) == fix["pixel_sha256"], "frame path (identity geometry) is not bit-identical to the reference"
    out = port.Oracle(d, sd).forward(pv, ids, mask)
    for k, ref in fix["outputs"].items():
        assert out[k].shape == ref.shape, k
        assert torch.allclose(out[k].float(), ref, atol=2e-6, rtol=1e-5), (k, (out[k].float() - ref).abs().max())
    # batch-1 call without padding equals the reference's batch-1 call
    n0 = len(rows[0])
    out1 = port.Oracle(d, sd).forward(pv[:8], ids[:1, :n0], mask[:1, :n0])
    for k, ref in fix["outputs_row0_alone"].items():
        assert torch.allclose(out1[k].float(), ref, atol=2e-6, rtol=1e-5), k


def test_llama_branch_matches_reference_golden(golden_llama_forward):
    """Llama-branch LLM (modeling_internvl_chat.py:125-126) with head_dim 96, fp32 and the reference's bf16 forward."""
    fix = golden_llama_forward
    d = synth.dims_tiny_llama()
    sd = synth.make_state_dict(d, seed=fix["seed"])
    pv, ids, mask, _ = _inputs(fix)
    out = port.Oracle(d, sd).forward(pv, ids, mask)
    for k, ref in fix["outputs"].items():
        assert torch.allclose(out[k].float(), ref, atol=2e-6, rtol=1e-5), (k, (out[k].float() - ref).abs().max())
    out = port.Oracle(d, sd, dtype=torch.bfloat16).forward(pv.to(torch.bfloat16), ids, mask)
    for k in ("rewards", "aspect_scores", "score", "aspect_gating_output"):
        # bf16 CPU matmul kernels may differ between machines in summation order: a few bf16 ulp
        assert torch.allclose(out[k].float(), fix["outputs_bf16"][k], atol=2e-2), k


def test_output_contract(golden_forward):
    """Shapes/dtypes of CustomOutput as the reference returns them."""
    o = golden_forward["outputs"]
    B = 3
    assert o["rewards"].shape == (B, 28) and o["aspect_scores"].shape == (B, 5) and o["score"].shape == (B,)
    assert o["aspect_gating_output"].shape == (B, 5) and o["criteria_gating_output"].shape == (B, 28)
    assert o["aspect_weights"].shape == (B, 28) and o["weighted_scores"].shape == (B,)
    # weighted_scores is the leaked loop variable == last aspect's score (moe_reward.py:273,294)
    assert torch.equal(o["weighted_scores"], o["aspect_scores"][:, -1])
    # aspect gate is a softmax * logit_scale(=1); group weights sum to 1 per aspect
    assert torch.allclose(o["aspect_gating_output"].sum(-1), torch.ones(B), atol=1e-5)
    for idx in synth.ASPECT2CRITERIA.values():
        assert torch.allclose(o["aspect_weights"][:, idx].sum(-1), torch.ones(B), atol=1e-5)


@pytest.mark.parametrize("name", ["448x448_m1", "896x1344_m6", "1024x1024_m6", "720x1280_m6", "1280x720_m6",
                                  "320x512_m6", "100x77_m1", "500x300_m4"])
def test_frame_path_matches_reference_golden(golden_frames, name):
    g = golden_frames[name]
    v = synth.make_video(g["video_id"], n_frames=g["n_frames"], height=g["H"], width=g["W"])
    pv, counts = port.load_video_frames(v, num_segments=g["num_segments"], max_num=g["max_num"])
    assert counts == g["counts"] and list(pv.shape) == g["shape"]
    assert _sha(pv) == g["sha256"]
    probe = pv.flatten()[:: max(1, pv.numel() // 64)][:64]
    assert np.array_equal(np.asarray(g["probe"], np.float32), probe.numpy())


def test_grid_choice_cases():
    """SURVEY §A.7 probe cases of dynamic_preprocess."""
    cases = {(448, 448, 1): 1, (1024, 1024, 6): 5, (1344, 896, 6): 7, (1280, 720, 6): 3, (720, 1280, 6): 3,
             (512, 320, 6): 7}
    for (w, h, mx), tiles in cases.items():
        gi, gj = port.choose_grid(w, h, 1, mx, 448)
        n = gi * gj
        assert n + (1 if n > 1 else 0) == tiles, (w, h, mx, gi, gj)


def test_pillow_bicubic_restatement_is_bit_exact():
    """The third-party arithmetic (Pillow's resampler, data.py:101,115) restated in oracle/port.py equals Pillow itself."""
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for (h, w, oh, ow) in [(100, 77, 448, 448), (320, 512, 896, 1344), (720, 1280, 448, 896), (600, 600, 448, 448),
                           (1080, 1920, 448, 448), (50, 50, 50, 120)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(PIL.fromarray(img).resize((ow, oh)))
        assert np.array_equal(port.pil_bicubic_resize(img, ow, oh), ref), (h, w, oh, ow)
