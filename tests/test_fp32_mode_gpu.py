"""fp32 verification mode (north_star: "per-criterion and aspect rewards within 2e-2 absolute in bf16 (1e-4 in fp32)";
SURVEY §A.7 - the float32 run of scripts/model/moe_reward.py:183-297 is the numerical oracle).

Engine(precision="fp32") / a float32 drop-in module computes every op in fp32 with this library's own plain CUDA kernels
(csrc/engine_fp32.cuh). Gate: 1e-4 ABSOLUTE on rewards / aspect_scores / score (and on every other head output) against
  * the golden vectors of the REAL reference run in fp32 on the CPU (tests/golden/*.pt, made by oracle/make_golden.py), and
  * the fp32 oracle port on the GPU (different summation order: cuBLAS / cuDNN),
on the tiny InternLM2 model (ragged batch, both prompt layouts), the tiny Llama-branch model (head_dim 96, GQA) and one
full-size MJ-VIDEO-2B row."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL32 = 1e-4
KEYS = ("rewards", "aspect_scores", "score", "weighted_scores", "aspect_gating_output", "aspect_weights",
        "criteria_gating_output")


@pytest.fixture(scope="module")
def gpu(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def _as_dict(d, out):
    NO, NA = d.num_objectives, d.num_aspects
    s, e = out["scores"], out["extras"]
    return dict(rewards=s[:, :NO], aspect_scores=s[:, NO:NO + NA], score=s[:, NO + NA],
                criteria_gating_output=e[:, :NO], aspect_gating_output=e[:, NO:NO + NA],
                aspect_weights=e[:, NO + NA:NO + NA + NO], weighted_scores=e[:, NO + NA + NO],
                hidden_state=out["hidden_state"], prompt_embedding=out["prompt_embedding"])


def _llama_cfg(d):
    from mj_video_b200.model import InternVLChatRewardModelingConfig
    vision = dict(hidden_size=d.vit_hidden, num_attention_heads=d.vit_heads, intermediate_size=d.vit_mlp,
                  num_hidden_layers=d.vit_layers, image_size=448, patch_size=14, qkv_bias=True, qk_normalization=False,
                  norm_type="layer_norm", layer_norm_eps=d.vit_eps)
    llm = dict(architectures=["LlamaForCausalLM"], hidden_size=d.llm_hidden, intermediate_size=d.llm_inter,
               num_hidden_layers=d.llm_layers, num_attention_heads=d.llm_heads, num_key_value_heads=d.llm_kv_heads,
               vocab_size=d.vocab, rms_norm_eps=d.rms_eps, rope_theta=d.rope_theta, max_position_embeddings=d.max_pos)
    return InternVLChatRewardModelingConfig(
        vision_config=vision, llm_config=llm, template="phi3-chat", downsample_ratio=0.5, ps_version="v2",
        select_layer=-1, force_image_size=448, num_objectives=28, num_aspects=5, aspect2criteria=d.aspect2criteria,
        gating_temperature=1.0, gating_hidden_dim=d.gating_hidden, gating_n_hidden=3)


def _check32(out, ref, tol=TOL32, hidden_tol=None):
    worst = {}
    for k in KEYS:
        worst[k] = (out[k].float().cpu() - ref[k].float().cpu()).abs().max().item()
        assert worst[k] <= tol, f"{k}: {worst[k]:.3e} > {tol}"
    if hidden_tol is not None:
        for k in ("hidden_state", "prompt_embedding"):
            err = (out[k].float().cpu() - ref[k].float().cpu()).abs().max().item()
            assert err <= hidden_tol, f"{k}: {err:.3e}"
    assert torch.equal(out["aspect_gating_output"].float().argmax(-1).cpu(), ref["aspect_gating_output"].float().argmax(-1).cpu())
    return worst


def test_fp32_engine_tiny_vs_real_reference_golden_and_oracle(gpu, golden_forward):
    from mj_video_b200 import synth
    from mj_video_b200.data_processor import load_video
    from mj_video_b200.engine import Engine, EngineDims
    from oracle import port
    fix = golden_forward
    d = synth.dims_tiny()
    sd = synth.make_state_dict(d, seed=fix["seed"])
    eng = Engine(EngineDims.from_synth(d), gpu, precision="fp32")
    eng.load_state_dict(sd)
    pv = torch.cat([load_video(synth.make_video(v), num_segments=8, max_num=1)[0] for v in fix["video_ids"]])  # CPU fp32
    rows = [synth.make_input_ids(i, [1] * 8, n_cap=n) for i, n in enumerate(fix["n_caps"])]
    ids, mask = synth.pad_batch(rows)
    res = eng.forward(pv.cuda(), ids, mask, synth.PAD, synth.IMG_CONTEXT)
    assert res["hidden_state"].dtype == torch.float32 and res["scores"].dtype == torch.float32
    mine = _as_dict(d, res)
    w = _check32(mine, fix["outputs"], hidden_tol=2e-3)          # the unmodified reference, fp32 on the CPU
    print("fp32 tiny vs real reference:", {k: f"{v:.2e}" for k, v in w.items()})
    o32 = port.Oracle(d, sd, dtype=torch.float32, device=gpu).forward(pv.cuda(), ids, mask)
    _check32(mine, o32, hidden_tol=2e-3)
    # a ragged batch with the first-slot prompt layout and the gating token far from the end
    pv2 = torch.cat([load_video(synth.make_video(900 + i), num_segments=8, max_num=1)[0] for i in range(3)])
    rows2 = [synth.make_input_ids(i, [1] * 8, n_cap=n, first_slot_layout=(i == 1)) for i, n in enumerate([5, 120, 33])]
    ids2, mask2 = synth.pad_batch(rows2)
    mine2 = _as_dict(d, eng.forward(pv2.cuda(), ids2, mask2, synth.PAD, synth.IMG_CONTEXT))
    _check32(mine2, port.Oracle(d, sd, dtype=torch.float32, device=gpu).forward(pv2.cuda(), ids2, mask2))
    # the bf16 product path on the same inputs stays within ITS tolerance of the fp32 mode
    eng16 = Engine(EngineDims.from_synth(d), gpu)
    eng16.load_state_dict(sd)
    m16 = _as_dict(d, eng16.forward(pv.cuda(), ids, mask, synth.PAD, synth.IMG_CONTEXT))
    for k in ("rewards", "aspect_scores", "score"):
        assert (m16[k].float() - mine[k].float()).abs().max().item() <= 2e-2
    # standalone GatingNetwork operator in fp32
    x = o32["prompt_embedding"].to(gpu)
    ag = eng.gating(0, x)
    assert ag.dtype == torch.float32 and (ag.cpu() - o32["aspect_gating_output"].cpu()).abs().max().item() <= 1e-5
    cg = eng.gating(1, x, with_softmax=False)
    assert (cg.cpu() - o32["criteria_gating_output"].cpu()).abs().max().item() <= 1e-4


def test_fp32_module_llama_branch_head_dim_96(gpu, golden_llama_forward):
    """A float32 drop-in module selects the fp32 engine by itself; Llama branch (separate q/k/v, head_dim 96, GQA)."""
    from mj_video_b200 import synth
    from mj_video_b200.data_processor import load_video
    from mj_video_b200.model import InternVLChatRewardModeling
    fix = golden_llama_forward
    d = synth.dims_tiny_llama()
    sd = synth.make_state_dict(d, seed=fix["seed"])
    m = InternVLChatRewardModeling("unused", _llama_cfg(d))
    m.load_state_dict(sd, strict=True)
    m.config.pad_token_id = synth.PAD
    m = m.cuda().eval()                      # stays float32
    m.model.img_context_token_id = synth.IMG_CONTEXT
    pv = torch.cat([load_video(synth.make_video(v), num_segments=8, max_num=1)[0] for v in fix["video_ids"]])
    rows = [synth.make_input_ids(i, [1] * 8, n_cap=n) for i, n in enumerate(fix["n_caps"])]
    ids, mask = synth.pad_batch(rows)
    out = m.forward(pv.cuda(), ids.cuda(), mask.cuda())
    assert m.engine.precision == "fp32"
    assert out.rewards.dtype == torch.float32 and out.hidden_state.dtype == torch.float32
    _check32(out, fix["outputs"], hidden_tol=2e-3)
    # the same module cast to bf16 switches to the tensor-core path (and its 2e-2 tolerance)
    m16 = m.to(torch.bfloat16)
    out16 = m16.forward(pv.to(torch.bfloat16).cuda(), ids.cuda(), mask.cuda())
    assert m16.engine.precision == "bf16" and out16.rewards.dtype == torch.bfloat16
    assert (out16.score.float().cpu() - fix["outputs"]["score"].float()).abs().max().item() <= 2e-2


def test_fp32_engine_full_size_2b_row(gpu):
    """One MJ-VIDEO-2B row (8 tiles, ~2.2k tokens, 24 + 24 layers) in fp32 against the fp32 oracle on the GPU: 1e-4."""
    from mj_video_b200 import synth
    from mj_video_b200.data_processor import load_video
    from mj_video_b200.engine import Engine, EngineDims
    from oracle import port
    d = synth.dims_2b()
    sd = synth.make_state_dict(d, seed=1234, include_lm_head=False)
    eng = Engine(EngineDims.from_synth(d), gpu, precision="fp32")
    eng.load_state_dict(sd)
    pv = load_video(synth.make_video(3), num_segments=8, max_num=1)[0].cuda()
    ids, mask = synth.pad_batch([synth.make_input_ids(0, [1] * 8)])
    mine = _as_dict(d, eng.forward(pv, ids, mask, synth.PAD, synth.IMG_CONTEXT))
    del eng
    torch.cuda.empty_cache()
    o32 = port.Oracle(d, sd, dtype=torch.float32, device=gpu).forward(pv, ids, mask)
    w = _check32(mine, o32, hidden_tol=5e-3)
    print("fp32 2B row vs fp32 oracle:", {k: f"{v:.2e}" for k, v in w.items()})


def test_fp32_handle_rejects_bf16_stage_entry_points(gpu):
    import ctypes as C
    from mj_video_b200 import _lib, synth
    from mj_video_b200.engine import Engine, EngineDims
    d = synth.dims_tiny()
    eng = Engine(EngineDims.from_synth(d), gpu, precision="fp32")
    eng.load_state_dict(synth.make_state_dict(d, seed=3, include_lm_head=False))
    ws = eng.workspace(1, 8, 2200)
    rc = eng.lib.mjv_vit_forward(eng._handle, C.c_void_p(ws.data_ptr()), 1, 1, C.c_void_p(ws.data_ptr()),
                                 C.c_void_p(ws.data_ptr()), C.c_int64(ws.numel()), _lib.current_stream())
    assert rc != 0 and b"bf16-only" in eng.lib.mjv_last_error()
