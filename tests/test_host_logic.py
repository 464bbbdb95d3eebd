"""Host-side logic of the product (no GPU): index semantics, prompt building, configs, ABI surface, sharding."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, reference_available
from mj_video_b200 import synth
from mj_video_b200.engine import find_token_for_gating
from oracle import port


def test_find_token_for_gating_semantics():
    pat = synth.GATING_PATTERN
    seq = [5, 6] + pat + [9, 9] + pat + [1]
    assert find_token_for_gating(seq) == port.find_token_for_gating(seq) == 9  # LAST occurrence
    assert find_token_for_gating(pat) == 0
    with pytest.raises(ValueError, match="Token pattern not found"):
        find_token_for_gating([1, 2, 3])
    with pytest.raises(ValueError):
        find_token_for_gating(pat[:-1])


def _engine_stub():
    """Engine.prepare_batch is pure host code; build an instance without touching CUDA."""
    from mj_video_b200.engine import Engine, EngineDims
    e = Engine.__new__(Engine)
    e.dims = EngineDims.from_synth(synth.dims_tiny())
    return e


def test_prepare_batch_indices_match_reference_semantics():
    e = _engine_stub()
    rows = [synth.make_input_ids(i, [1] * 8, n_cap=n) for i, n in enumerate([46, 17, 90])]
    ids, mask = synth.pad_batch(rows)
    meta, info = e.prepare_batch(ids, mask, 24, synth.PAD, synth.IMG_CONTEXT)
    last = port.sequence_end_index(ids, synth.PAD)
    assert info["last"].tolist() == last.tolist() == [len(r) - 1 for r in rows]
    assert info["gate"].tolist() == [port.find_token_for_gating(r) for r in rows]
    cu = meta["cu_seqlens"]
    assert cu.tolist() == [0] + list(np.cumsum([len(r) for r in rows]))
    assert meta["token_ids"].tolist() == sum(rows, [])
    assert (meta["vt_starts"] % 8 == 0).all()                      # TMA needs 16-byte aligned V^T starts
    assert meta["img_rows"].size == 24 * 256
    assert (meta["token_ids"][meta["img_rows"]] == synth.IMG_CONTEXT).all()
    assert meta["gather_rows"].tolist() == [cu[b] + len(rows[b]) - 1 for b in range(3)] + \
        [cu[b] + port.find_token_for_gating(rows[b]) for b in range(3)]
    # work list: every (sequence, 128-row block) once, heaviest first
    work = meta["work"].reshape(-1, 2)
    assert sorted(map(tuple, work.tolist())) == sorted((s, q) for s in range(3) for q in range((len(rows[s]) + 127) // 128))
    assert (np.diff(work[:, 1]) <= 0).all()
    # last-layer shortcut: exactly the query blocks that hold a gathered row
    wl = set(map(tuple, meta["work_last"].reshape(-1, 2).tolist()))
    exp = {(b, (len(rows[b]) - 1) // 128) for b in range(3)} | {(b, port.find_token_for_gating(rows[b]) // 128) for b in range(3)}
    assert wl == exp and info["n_work_last"] == len(exp)


def test_prepare_batch_errors():
    e = _engine_stub()
    rows = [synth.make_input_ids(0, [1] * 8, n_cap=20), synth.make_input_ids(1, [1] * 8, n_cap=30)]
    ids, mask = synth.pad_batch(rows)
    with pytest.raises(ValueError, match="Cannot handle batch sizes > 1"):       # moe_reward.py:218-219
        e.prepare_batch(ids, mask, 16, None, synth.IMG_CONTEXT)
    bad = ids.clone()
    bad[0, len(rows[0]) - 5:len(rows[0])] = 11
    with pytest.raises(ValueError, match="Token pattern not found"):             # moe_reward.py:57
        e.prepare_batch(bad, mask, 16, synth.PAD, synth.IMG_CONTEXT)
    with pytest.raises(RuntimeError, match="visual embeddings"):                  # more image tokens than ViT rows
        e.prepare_batch(ids, mask, 15, synth.PAD, synth.IMG_CONTEXT)
    with pytest.warns(UserWarning, match="truncating"):                           # modeling_internvl_chat.py:178-186
        meta, info = e.prepare_batch(ids, mask, 17, synth.PAD, synth.IMG_CONTEXT)
    assert info["n_img"] == 16 * 256


def test_single_row_without_pad_token():
    e = _engine_stub()
    rows = [synth.make_input_ids(0, [1] * 8, n_cap=20)]
    ids, mask = synth.pad_batch(rows)
    meta, info = e.prepare_batch(ids, None, 8, None, synth.IMG_CONTEXT)
    assert info["last"].tolist() == [ids.shape[1] - 1]


def test_prepare_chat_input_matches_reference_golden(golden_chat):
    import hashlib
    from mj_video_b200.model import InternVLChatRewardModelingConfig, prepare_chat_input
    from oracle.make_golden import StubTokenizer
    d = synth.dims_tiny()
    cfg = InternVLChatRewardModelingConfig(
        vision_config=dict(image_size=448, patch_size=14, hidden_size=d.vit_hidden), llm_config=dict(hidden_size=256),
        template="internlm2-chat", downsample_ratio=0.5, force_image_size=448)
    for name, g in golden_chat.items():
        gen = {}
        ids, mask = prepare_chat_input(cfg, StubTokenizer(), torch.zeros(g["n_tiles"], 3, 448, 448), g["question"], gen,
                                       num_patches_list=g["num_patches_list"])
        assert gen["eos_token_id"] == g["eos"]
        assert ids.shape[1] == g["n"] and int((ids == 92546).sum()) == g["n_img"], name
        assert ids[0, :60].tolist() == g["head"] and ids[0, -12:].tolist() == g["tail"]
        assert hashlib.sha256(ids.contiguous().numpy().tobytes()).hexdigest() == g["sha256"]
        assert torch.equal(mask, torch.ones_like(ids))
    with pytest.raises(AssertionError):   # tile-count mismatch (modeling_internvl_chat.py:49)
        prepare_chat_input(cfg, StubTokenizer(), torch.zeros(3, 3, 448, 448), "<image>", {}, num_patches_list=[1, 1])


def test_config_and_module_surface():
    from mj_video_b200.model import InternVLChatRewardModeling, InternVLChatRewardModelingConfig
    with pytest.raises(ValueError, match="Unsupported architecture"):      # configuration_internvl_chat.py:50-55
        InternVLChatRewardModelingConfig(llm_config={"architectures": ["Phi3ForCausalLM"]})
    d = synth.dims_tiny()
    vision = dict(hidden_size=d.vit_hidden, num_attention_heads=d.vit_heads, intermediate_size=d.vit_mlp,
                  num_hidden_layers=d.vit_layers, image_size=448, patch_size=14, qkv_bias=True, qk_normalization=False,
                  norm_type="layer_norm", layer_norm_eps=1e-6)
    llm = dict(architectures=["InternLM2ForCausalLM"], hidden_size=d.llm_hidden, intermediate_size=d.llm_inter,
               num_hidden_layers=d.llm_layers, num_attention_heads=d.llm_heads, num_key_value_heads=d.llm_kv_heads,
               vocab_size=d.vocab, rms_norm_eps=1e-5, bias=False, rope_theta=1e6, max_position_embeddings=32768)
    cfg = InternVLChatRewardModelingConfig(vision_config=vision, llm_config=llm, template="internlm2-chat",
                                           ps_version="v2", force_image_size=448, num_objectives=28, num_aspects=5,
                                           aspect2criteria=d.aspect2criteria, gating_hidden_dim=d.gating_hidden)
    assert all(k in cfg.to_dict() for k in ("num_objectives", "num_aspects", "aspect2criteria", "gating_temperature",
                                            "gating_hidden_dim", "gating_n_hidden"))
    m = InternVLChatRewardModeling("unused", cfg)
    sd = synth.make_state_dict(d, seed=1)
    assert set(m.state_dict().keys()) == set(sd.keys())                    # SURVEY §A.2 key space
    m.config.pad_token_id = 2
    m.model.img_context_token_id = synth.IMG_CONTEXT
    ids, mask = synth.pad_batch([synth.make_input_ids(0, [1] * 8, n_cap=10)])
    # the reference's from_pretrained(name) loads the backbone or raises (moe_reward.py:142); here a backbone that was never
    # loaded, or only partly (strict=False with missing keys), must not score uninitialised memory
    with pytest.raises(RuntimeError, match="never loaded"):
        m.forward(torch.zeros(8, 3, 448, 448), ids, mask)
    partial = {k: v for k, v in sd.items() if "vision_model.encoder.layers.0." not in k}
    m.load_state_dict(partial, strict=False)
    with pytest.raises(RuntimeError, match=r"backbone parameters were never loaded .*vision_model\.encoder\.layers\.0\."):
        m.forward(torch.zeros(8, 3, 448, 448), ids, mask)
    m.load_state_dict(sd, strict=True)
    assert not m.reward_transform_matrix.requires_grad
    with pytest.raises(RuntimeError, match="CUDA"):                        # no CPU fallback
        m.forward(torch.zeros(8, 3, 448, 448), ids, mask)
    with pytest.raises(AssertionError):                                    # moe_reward.py:153-157
        bad = InternVLChatRewardModelingConfig(vision_config=vision, llm_config=llm, template="internlm2-chat",
                                               num_objectives=28, num_aspects=4, aspect2criteria=d.aspect2criteria)
        InternVLChatRewardModeling("unused", bad)


def test_abi_exports_every_declared_symbol(built_lib):
    """include/mjv_b200.h is the drop-in boundary: every function it declares must be exported by the library."""
    hdr = open(os.path.join(ROOT, "include", "mjv_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mjv_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(built_lib, n), f"{n} declared in include/mjv_b200.h but not exported"
    assert built_lib.mjv_abi_version() == 1


def test_ctypes_mirrors_match_the_compiled_structs(built_lib):
    """Every ctypes.Structure in _lib.py has the size (hence field layout up to padding) the C compiler gave the struct of
    include/mjv_b200.h; _lib.load() itself refuses to continue otherwise."""
    from mj_video_b200 import _lib
    lib = _lib.load()
    for which, mirror in enumerate((_lib.GemmDesc, _lib.AttnDesc, _lib.Config, _lib.Batch, _lib.Outputs)):
        assert lib.mjv_abi_struct_size(which) == C.sizeof(mirror), mirror.__name__
    assert lib.mjv_abi_struct_size(99) == -1


def test_host_helpers_of_the_library(built_lib):
    """Pure-host entry points: grid choice and the Pillow coefficient tables equal the oracle's."""
    gi, gj = C.c_int32(), C.c_int32()
    for (w, h, mx) in [(448, 448, 1), (1344, 896, 6), (1024, 1024, 6), (1280, 720, 6), (720, 1280, 6), (512, 320, 6),
                       (300, 500, 4), (1920, 1080, 12), (77, 100, 1)]:
        assert built_lib.mjv_choose_grid(w, h, 1, mx, 448, C.byref(gi), C.byref(gj)) == 0
        assert (gi.value, gj.value) == port.choose_grid(w, h, 1, mx, 448)
    for (i, o) in [(1280, 896), (720, 448), (100, 448), (320, 896), (1920, 1344), (333, 448), (2160, 448)]:
        ks = C.c_int32()
        assert built_lib.mjv_bicubic_coeffs(i, o, C.byref(ks), None, None, 0) == 0
        b = np.zeros(2 * o, np.int32)
        kk = np.zeros(o * ks.value, np.int32)
        assert built_lib.mjv_bicubic_coeffs(i, o, C.byref(ks), b.ctypes.data, kk.ctypes.data, kk.size) == 0
        xmin, cnt, pk = port.bicubic_coeffs(i, o)
        assert (b[0::2] == xmin).all() and (b[1::2] == cnt).all() and (pk == kk.reshape(o, ks.value)).all()
    assert built_lib.mjv_choose_grid(0, 5, 1, 1, 448, C.byref(gi), C.byref(gj)) < 0
    assert b"bad args" in built_lib.mjv_last_error()


def test_compute_entry_points_fail_loudly_without_gpu(built_lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mj_video_b200 import _lib
    from mj_video_b200.engine import Engine, EngineDims
    with pytest.raises(_lib.MjvError):
        Engine(EngineDims.from_synth(synth.dims_tiny()), "cpu")
    cfg = _lib.Config()
    cfg.image_size, cfg.patch_size, cfg.vit_hidden, cfg.llm_hidden = 448, 14, 128, 256
    cfg.num_objectives, cfg.num_aspects, cfg.gating_n_hidden = 28, 5, 3
    h = C.c_void_p()
    assert built_lib.mjv_create(C.byref(cfg), C.byref(h)) == -2        # MJV_ERR_CUDA: no device, no fallback
    assert b"no CUDA device" in built_lib.mjv_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CUDA extension -> an error that says so; nothing falls back to PyTorch or the oracle."""
    from mj_video_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_PATH", tmp_path / "libmjv_b200.so")
    with pytest.raises(_lib.MjvError, match="no CPU/PyTorch fallback"):
        _lib.load()
    from mj_video_b200 import ops
    with pytest.raises(_lib.MjvError):
        ops.gemm(torch.zeros(1, 8, dtype=torch.bfloat16), torch.zeros(128, 8, dtype=torch.bfloat16))


def test_pair_sharding():
    from mj_video_b200.dist import pair_rows, preference_decisions, shard_pairs
    for n, w in [(4096, 8), (4096, 2), (10, 4), (3, 8), (0, 2)]:
        blocks = [list(shard_pairs(n, r, w)) for r in range(w)]
        assert sum(blocks, []) == list(range(n))
        assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= (n + w - 1) // w
    assert pair_rows(range(2, 4)) == [(4, 5), (6, 7)]                      # chosen/rejected stay together
    s = torch.tensor([0.3, 0.1, 0.2, 0.2, -1.0, 0.5])
    assert preference_decisions(s).tolist() == [0, 1, 1]                    # ~(s0 > s1): ties -> video_1


def test_safetensors_checkpoint_roundtrip(tmp_path):
    """weights travel as *.safetensors named by state-dict key (overall_train.py:503-522)."""
    pytest.importorskip("safetensors")
    import torch
    from mj_video_b200 import checkpoint, synth
    d = synth.dims_tiny()
    sd = synth.make_state_dict(d, seed=4)
    keys = sorted(sd)
    checkpoint.save_state_dict({k: sd[k] for k in keys[: len(keys) // 2]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    checkpoint.save_state_dict({k: sd[k] for k in keys[len(keys) // 2:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    got = dict(checkpoint.iter_tensors(str(tmp_path)))
    assert sorted(got) == keys and all(torch.equal(got[k], sd[k]) for k in keys)

    class FakeEngine:
        def __init__(self):
            self.keys, self.final = [], False

        def load_weight(self, k, t):
            self.keys.append(k)

        def finalize(self):
            self.final = True
    fe = FakeEngine()
    n = checkpoint.load_into_engine(fe, str(tmp_path))
    assert fe.final and n == len(keys) - 1 and "model.language_model.output.weight" not in fe.keys
    with pytest.raises(FileNotFoundError):
        checkpoint.checkpoint_files(str(tmp_path / "missing"))


def test_opencv_reader_decodes_a_real_file(tmp_path):
    """`load_video(path)` on an actual video file: the OpenCV reader (the image has no decord / PyAV) returns the frames
    `get_index` selects, in RGB, identical to decoding the whole file and indexing."""
    cv2 = pytest.importorskip("cv2")
    from mj_video_b200.data_processor import data as dp
    path = str(tmp_path / "clip.avi")
    vid = synth.make_video(5, n_frames=12, height=96, width=128)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 8.0, (128, 96))
    if not wr.isOpened():
        pytest.skip("this OpenCV build cannot write MJPG/AVI")
    for f in vid:
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    wr.release()
    rd = dp._OpenCVReader(path)
    assert len(rd) == 12 and abs(rd.get_avg_fps() - 8.0) < 1e-3
    cap = cv2.VideoCapture(path)
    allf = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        allf.append(cv2.cvtColor(fr, cv2.COLOR_BGR2RGB))
    cap.release()
    idx = dp.get_index(None, rd.get_avg_fps(), len(rd) - 1, first_idx=0, num_segments=5)
    got = rd.take(idx)
    assert got.shape == (5, 96, 128, 3) and got.dtype == np.uint8
    assert all(np.array_equal(got[k], allf[int(i)]) for k, i in enumerate(idx))
    assert np.abs(got[0].astype(int) - vid[int(idx[0])].astype(int)).mean() < 12     # MJPG is lossy but close
    assert isinstance(dp.open_video(path), (dp._OpenCVReader, dp._DecordReader, dp._AvReader))


def test_prepare_chat_input_memoises_identical_prompts():
    """Chosen and rejected rows of a pair share the prompt: the second call must not tokenise again, and must return
    independent tensors with the same ids."""
    import torch
    from mj_video_b200.model import InternVLChatRewardModelingConfig, prepare_chat_input
    from oracle.make_golden import StubTokenizer

    class Counting(StubTokenizer):
        calls = 0

        def __call__(self, text, return_tensors="pt"):
            Counting.calls += 1
            return super().__call__(text, return_tensors)
    d = synth.dims_tiny()
    cfg = InternVLChatRewardModelingConfig(
        vision_config=dict(hidden_size=d.vit_hidden, num_attention_heads=d.vit_heads, intermediate_size=d.vit_mlp,
                           num_hidden_layers=d.vit_layers, image_size=448, patch_size=14),
        llm_config=dict(architectures=["InternLM2ForCausalLM"], hidden_size=d.llm_hidden, intermediate_size=d.llm_inter,
                        num_hidden_layers=d.llm_layers, num_attention_heads=d.llm_heads,
                        num_key_value_heads=d.llm_kv_heads, vocab_size=d.vocab),
        template="internlm2-chat", downsample_ratio=0.5, ps_version="v2", select_layer=-1, force_image_size=448,
        num_objectives=28, num_aspects=5, aspect2criteria=d.aspect2criteria)
    tok = Counting()
    pv = torch.zeros(8, 3, 448, 448)
    q = "".join(f"Frame{i+1}: <image>\n" for i in range(8)) + "a red kite"
    a_ids, a_mask = prepare_chat_input(cfg, tok, pv, q, {}, num_patches_list=[1] * 8)
    b_ids, b_mask = prepare_chat_input(cfg, tok, pv, q, {}, num_patches_list=[1] * 8)
    # the first call of a tokenizer object tokenises twice (collapsed prompt + the full text that verifies it); the repeat
    # of the same prompt does not tokenise at all
    first = Counting.calls
    assert first == 2 and torch.equal(a_ids, b_ids) and torch.equal(a_mask, b_mask)
    assert a_ids.data_ptr() != b_ids.data_ptr()
    a_ids[0, 0] = -5                                   # callers may edit their tensors without poisoning the cache
    c_ids, _ = prepare_chat_input(cfg, tok, pv, q, {}, num_patches_list=[1] * 8)
    assert torch.equal(c_ids, b_ids)
    prepare_chat_input(cfg, tok, pv, q + "!", {}, num_patches_list=[1] * 8)
    assert Counting.calls == first + 2       # a new prompt, still inside the verification window
    prepare_chat_input(cfg, tok, pv, q + "?", {}, num_patches_list=[1] * 8)
    assert Counting.calls == first + 3       # trusted now: one short tokenisation per new prompt


def test_product_never_imports_the_oracle_or_the_reference():
    """The oracle is test infrastructure: no module of the product package (nor its CUDA sources) may import, call or
    read anything under oracle/ or /root/reference; only tests/, __graft_entry__.smoke() and bench.py's CPU legs do."""
    import ast
    import pathlib
    pkg = pathlib.Path(ROOT) / "mj_video_b200"
    for f in pkg.rglob("*.py"):
        tree = ast.parse(f.read_text())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), f"{f} imports the oracle"
        assert "/root/reference" not in f.read_text(), f"{f} reads the reference tree"
    for f in list((pkg / "csrc").glob("*")) + [pathlib.Path(ROOT) / "include" / "mjv_b200.h"]:
        assert "/root/reference" not in f.read_text(errors="ignore")
    bench = (pathlib.Path(ROOT) / "bench.py").read_text()
    assert "reference_harness" not in bench and "REFERENCE_ROOT" not in bench   # the GPU box has no reference tree


def test_plain_c_host_links_against_the_abi(built_lib, tmp_path):
    """include/mjv_b200.h is plain C and examples/abi_host.c — a non-Python caller — compiles, links against the shared
    library and runs its host-only part (struct sizes, grid choice; mjv_create fails loudly without a GPU)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from mj_video_b200 import _lib
    libdir = str(_lib.lib_path().parent)
    exe = str(tmp_path / "abi_host")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c",
                    os.path.join(ROOT, "include", "mjv_b200.h")], check=True)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "abi_host.c"),
                    "-o", exe, "-L", libdir, "-lmjv_b200", f"-Wl,-rpath,{libdir}"], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "grid rc 0: 3 x 2" in res.stdout


def _chat_cfg():
    from mj_video_b200.model import InternVLChatRewardModelingConfig
    d = synth.dims_tiny()
    vision = dict(hidden_size=d.vit_hidden, num_attention_heads=d.vit_heads, intermediate_size=d.vit_mlp,
                  num_hidden_layers=d.vit_layers, image_size=448, patch_size=14, qkv_bias=True, qk_normalization=False,
                  norm_type="layer_norm", layer_norm_eps=1e-6)
    llm = dict(architectures=["InternLM2ForCausalLM"], hidden_size=d.llm_hidden, intermediate_size=d.llm_inter,
               num_hidden_layers=d.llm_layers, num_attention_heads=d.llm_heads, num_key_value_heads=d.llm_kv_heads,
               vocab_size=d.vocab, rms_norm_eps=1e-5, bias=False, rope_theta=1e6, max_position_embeddings=32768)
    return InternVLChatRewardModelingConfig(vision_config=vision, llm_config=llm, template="internlm2-chat", ps_version="v2",
                                            force_image_size=448, num_objectives=28, num_aspects=5,
                                            aspect2criteria=d.aspect2criteria, gating_hidden_dim=d.gating_hidden)


def test_prepare_chat_input_collapsed_context_runs_are_token_identical():
    """the prompt is tokenised with ONE <IMG_CONTEXT> per image and the id repeated afterwards; the result
    must equal tokenizer(full prompt) (modeling_internvl_chat.py:36-89), the tokenizer must only ever see the short text
    once the path is trusted, and a tokenizer for which the shortcut is NOT valid must be detected and left alone."""
    from oracle.make_golden import StubTokenizer
    from mj_video_b200.model import prepare_chat_input
    import mj_video_b200.model.internvl2.modeling_internvl_chat as mod
    cfg = _chat_cfg()

    class Counting(StubTokenizer):
        def __init__(self):
            self.lengths = []

        def __call__(self, text, return_tensors="pt"):
            self.lengths.append(len(text))
            return super().__call__(text, return_tensors)

    def full(tok, question, npl):
        """The reference's statement sequence, with the full text handed to the tokenizer."""
        from mj_video_b200.model.internvl2.conversation import get_conv_template
        t = get_conv_template(cfg.template)
        q = question if "<image>" in question else "<image>\n" + question
        t.append_message(t.roles[0], q)
        t.append_message(t.roles[1], None)
        query = t.get_prompt()
        for n in npl:
            query = query.replace("<image>", "<img>" + "<IMG_CONTEXT>" * 256 * n + "</img>", 1)
        return StubTokenizer()(query)["input_ids"]

    tok = Counting()
    cases = [("a caption", [8]), ("Frame1: <image>\nFrame2: <image>\nwhat happens?", [1, 3]),
             ("<image> zero tiles <image>", [0, 2]), ("other caption", [8]), ("third", [2]), ("fourth", [2])]
    for question, npl in cases:
        pv = torch.zeros(sum(npl), 3, 448, 448)
        ids, mask = prepare_chat_input(cfg, tok, pv, question, {}, num_patches_list=npl)
        assert torch.equal(ids, full(tok, question, npl)) and torch.equal(mask, torch.ones_like(ids))
    # the first _VERIFY_FIRST calls also tokenised the full text; afterwards only the collapsed prompt (< 1.5 KB) is seen
    trusted = tok.lengths[2 * mod._VERIFY_FIRST:]
    assert trusted and max(trusted) < 1500, tok.lengths
    # a question that itself contains the context token: the count check sends it down the full-text path, same ids
    q = "why <IMG_CONTEXT> here? <image>"
    ids, _ = prepare_chat_input(cfg, tok, torch.zeros(1, 3, 448, 448), q, {}, num_patches_list=[1])
    assert torch.equal(ids, full(tok, q, [1]))

    class Hostile(StubTokenizer):
        """Appends a marker after every run of two or more context tokens: collapsing the run changes the ids."""
        def __call__(self, text, return_tensors="pt"):
            out = super().__call__(text, return_tensors)["input_ids"][0].tolist()
            res, i = [], 0
            while i < len(out):
                j = i
                while j < len(out) and out[j] == 92546:
                    j += 1
                res += out[i:max(j, i + 1)]
                if j - i >= 2:
                    res.append(77)
                i = max(j, i + 1)
            t = torch.tensor([res], dtype=torch.long)
            return {"input_ids": t, "attention_mask": torch.ones_like(t)}

    bad = Hostile()
    for question in ("x", "y", "z"):
        ids, _ = prepare_chat_input(cfg, bad, torch.zeros(2, 3, 448, 448), question, {}, num_patches_list=[2])
        t = __import__("mj_video_b200.model.internvl2.conversation", fromlist=["get_conv_template"]).get_conv_template(cfg.template)
        t.append_message(t.roles[0], "<image>\n" + question)
        t.append_message(t.roles[1], None)
        want = bad(t.get_prompt().replace("<image>", "<img>" + "<IMG_CONTEXT>" * 512 + "</img>", 1))["input_ids"]
        assert torch.equal(ids, want)
    assert mod._COLLAPSE_STATE[id(bad)][2] is False      # detected on the first call, never collapsed again


def test_clip_decode_pool_matches_serial_decode(tmp_path):
    """SURVEY §8(f) N1 (host part): clips decoded by the thread pool are bit-identical to serial `sample_clip` and come
    back in submission order."""
    cv2 = pytest.importorskip("cv2")
    from mj_video_b200.data_processor import ClipDecodePool, sample_clip
    paths = []
    for v in range(5):
        path = str(tmp_path / f"clip{v}.avi")
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 8.0, (96, 64))
        vid = synth.make_video(300 + v, n_frames=10 + v, height=64, width=96)
        for f in vid:
            wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
        wr.release()
        paths.append(path)
    serial = [sample_clip(p, None, 4) for p in paths]
    with ClipDecodePool(workers=3, num_segments=4) as pool:
        par = list(pool.map(paths))
        arr = list(pool.map([synth.make_video(7, n_frames=9, height=32, width=48)]))   # in-memory arrays work as well
    assert len(par) == 5 and all(np.array_equal(a, b) for a, b in zip(serial, par))
    assert all(c.shape == (4, 64, 96, 3) and c.dtype == np.uint8 for c in par)
    assert arr[0].shape == (4, 32, 48, 3)
    with ClipDecodePool(workers=2, num_segments=4) as pool, pytest.raises(Exception):
        list(pool.map([paths[0], str(tmp_path / "missing.avi")]))


def test_longrope_tables_match_transformers():
    """N4 (Phi-3 / InternVL2-4B LongRoPE): the factor-set choice, inverse frequencies and the attention scale agree with
    transformers' own `_compute_longrope_parameters` (the implementation the reference's Llama-layout branch would call)."""
    import types
    import torch
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    from mj_video_b200 import synth
    from mj_video_b200.engine import EngineDims, longrope_inv_freq
    d0 = synth.dims_tiny_llama()
    hd = d0.llm_hidden // d0.llm_heads
    g = torch.Generator().manual_seed(0)
    short = (1.0 + 0.1 * torch.rand(hd // 2, generator=g)).tolist()
    long_ = (1.0 + 30.0 * torch.rand(hd // 2, generator=g)).tolist()
    import dataclasses
    d = dataclasses.replace(EngineDims.from_synth(d0), rope_scaling_type="longrope", rope_short_factor=tuple(short),
                            rope_long_factor=tuple(long_), rope_original_max_pos=64, max_pos=2048)
    assert d.longrope
    cfg = types.SimpleNamespace(
        hidden_size=d.llm_hidden, num_attention_heads=d.llm_heads, max_position_embeddings=d.max_pos,
        rope_parameters=dict(rope_type="longrope", rope_theta=d.rope_theta, short_factor=short, long_factor=long_,
                             original_max_position_embeddings=64))
    cfg.standardize_rope_params = lambda: None
    for seq_len, want_long in ((10, False), (64, False), (65, True), (2000, True)):
        inv, scale, is_long = longrope_inv_freq(d, hd, seq_len)
        ref_inv, ref_scale = ROPE_INIT_FUNCTIONS["longrope"](cfg, None, seq_len=seq_len)
        assert is_long == want_long
        assert torch.equal(inv, ref_inv) and abs(scale - ref_scale) < 1e-12 and scale > 1.0
    # no extension of the context -> no scale
    d1 = dataclasses.replace(d, max_pos=64)
    assert longrope_inv_freq(d1, hd, 10)[1] == 1.0
    import pytest
    with pytest.raises(ValueError):
        longrope_inv_freq(dataclasses.replace(d, rope_short_factor=(1.0,)), hd, 10)


def test_bench_rank0_only_sections_issue_no_collective():
    """bench.py at N > 1: every rank must reach every collective. The per-kernel breakdown step and the CPU baseline run on
    rank 0 only, so they must use the rank-local step (no all-gather) - a breakdown step that ended in the NCCL all-gather
    hung every multi-GPU run of one build of this round."""
    import ast
    import pathlib
    src = pathlib.Path(__file__).resolve().parents[1].joinpath("bench.py").read_text()
    tree = ast.parse(src)
    run_mjv = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "run_mjv")
    collective_names = {"step_resident", "step_e2e", "barrier", "gather", "allgather_scores", "all_gather_into_tensor",
                        "all_reduce", "broadcast"}

    def is_rank0_test(t):
        return (isinstance(t, ast.Compare) and isinstance(t.left, ast.Name) and t.left.id == "rank") or \
               (isinstance(t, ast.BoolOp) and any(is_rank0_test(v) for v in t.values))

    checked = 0
    for node in ast.walk(run_mjv):
        if isinstance(node, ast.If) and is_rank0_test(node.test):
            checked += 1
            for sub in node.body:
                for call in (c for c in ast.walk(sub) if isinstance(c, ast.Call)):
                    name = call.func.id if isinstance(call.func, ast.Name) else getattr(call.func, "attr", "")
                    assert name not in collective_names, f"bench.py line {call.lineno}: {name}() inside a rank-0-only block"
    assert checked >= 2   # the breakdown block and the JSON print at least
