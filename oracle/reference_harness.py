"""TEST INFRASTRUCTURE — not product code. Imports the UNMODIFIED reference from /root/reference (this container only;
the GPU box has no /root/reference) so that
  (a) the torch restatement in oracle/port.py can be validated against the reference's own forward, and
  (b) golden input/output vectors can be generated (oracle/make_golden.py -> tests/golden/).
Nothing under mj_video_b200/ imports this module.

Shims (all outside the reference tree, nothing edited):
  1. stub `timm.models.layers.DropPath` (modeling_intern_vit.py:13) — registered AFTER transformers is imported;
  2. `transformers.models.llama.modeling_llama.LLAMA_INPUTS_DOCSTRING = ''` (imported-but-unused at moe_reward.py:11);
  3. `InternVLChatModel.from_pretrained` patched to return a locally built model (moe_reward.py:142 hits the hub);
  4. data.py is exec'd with stub `decord`/`av` and a writable `__file__` (data.py:4,16,20-22).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("MJV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scripts", "model"))


_mods = None


def _import_reference():
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import transformers
    from transformers import AutoModel  # noqa: F401  (forces the lazy module)
    import transformers.models.llama.modeling_llama as ml
    if not hasattr(ml, "LLAMA_INPUTS_DOCSTRING"):
        ml.LLAMA_INPUTS_DOCSTRING = ""
    if "timm" not in sys.modules:
        class DropPath(torch.nn.Module):
            def __init__(self, p=0.0):
                super().__init__()

            def forward(self, x):
                return x
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm_layers = types.ModuleType("timm.models.layers")
        timm_layers.DropPath = DropPath
        timm.models = timm_models
        timm_models.layers = timm_layers
        sys.modules["timm"], sys.modules["timm.models"], sys.modules["timm.models.layers"] = timm, timm_models, timm_layers
    model_dir = os.path.join(REFERENCE_ROOT, "scripts", "model")
    if model_dir not in sys.path:
        sys.path.insert(0, model_dir)
    import moe_reward
    import internvl2
    from internvl2 import modeling_internvl_chat
    _mods = (moe_reward, internvl2, modeling_internvl_chat)
    return _mods


def ensure_process_group():
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("gloo", rank=0, world_size=1)


def build_reference_config(d):
    """InternVLChatRewardModelingConfig for synthetic dims `d` (mj_video_b200.synth.ModelDims)."""
    moe_reward, _, _ = _import_reference()
    vision = dict(hidden_size=d.vit_hidden, num_attention_heads=d.vit_heads, intermediate_size=d.vit_mlp,
                  num_hidden_layers=d.vit_layers, image_size=d.image_size, patch_size=d.patch_size,
                  qkv_bias=True, qk_normalization=False, norm_type="layer_norm", layer_norm_eps=d.vit_eps,
                  hidden_act="gelu", drop_path_rate=0.0, initializer_factor=d.vit_ls_init, use_flash_attn=False)
    arch = getattr(d, "llm_arch", "InternLM2ForCausalLM")
    llm = dict(architectures=[arch], hidden_size=d.llm_hidden, intermediate_size=d.llm_inter,
               num_hidden_layers=d.llm_layers, num_attention_heads=d.llm_heads, num_key_value_heads=d.llm_kv_heads,
               vocab_size=d.vocab, rms_norm_eps=d.rms_eps, hidden_act="silu", rope_theta=d.rope_theta,
               max_position_embeddings=d.max_pos, attn_implementation="eager", use_cache=True, pad_token_id=2)
    if arch == "InternLM2ForCausalLM":
        f = getattr(d, "rope_scaling_factor", 2.0)
        llm.update(bias=False, rope_scaling=None if f is None else {"type": getattr(d, "rope_scaling_type", "dynamic"), "factor": f})
    cfg = moe_reward.InternVLChatRewardModelingConfig(
        vision_config=vision, llm_config=llm, template=getattr(d, "template", "internlm2-chat"), downsample_ratio=d.downsample_ratio,
        ps_version=getattr(d, "ps_version", "v2"), select_layer=getattr(d, "select_layer", -1), force_image_size=d.image_size, num_objectives=d.num_objectives,
        num_aspects=d.num_aspects, aspect2criteria={k: list(v) for k, v in d.aspect2criteria.items()},
        gating_temperature=d.gating_temperature, gating_hidden_dim=d.gating_hidden, gating_n_hidden=d.gating_n_hidden)
    return cfg


def build_reference_model(d, state_dict, dtype=torch.float32, pad_token_id=2, img_context_token_id=92546):
    """The reference's own InternVLChatRewardModeling with `state_dict` loaded strictly (eval_genai_mjvideo.py:98-115)."""
    moe_reward, internvl2, _ = _import_reference()
    ensure_process_group()
    cfg = build_reference_config(d)
    base = internvl2.InternVLChatModel(cfg, use_flash_attn=False)
    orig = internvl2.InternVLChatModel.from_pretrained
    internvl2.InternVLChatModel.from_pretrained = classmethod(lambda cls, name, *a, **k: base)
    try:
        m = moe_reward.InternVLChatRewardModeling("unused", cfg)
    finally:
        internvl2.InternVLChatModel.from_pretrained = orig
    m.load_state_dict(state_dict, strict=True)
    m.config.pad_token_id = pad_token_id
    m.model.img_context_token_id = img_context_token_id
    m = m.to(dtype).eval()
    return m


@torch.no_grad()
def reference_forward(model, pixel_values, input_ids, attention_mask):
    out = model.forward(pixel_values, input_ids, attention_mask)
    return {k: v for k, v in out.items()}


def reference_modules():
    return _import_reference()


# ----------------------------------------------------------------------------------------- frame path
class _StubVideoReader:
    """In-memory stand-in for decord.VideoReader (data.py:164-172): len, get_avg_fps, vr[i].asnumpy()."""
    registry: dict[str, np.ndarray] = {}

    class _Frame:
        def __init__(self, a):
            self._a = a

        def asnumpy(self):
            return self._a

    def __init__(self, path, ctx=None, num_threads=1):
        self.frames = _StubVideoReader.registry[str(path)]

    def __len__(self):
        return len(self.frames)

    def get_avg_fps(self):
        return 8.0

    def __getitem__(self, i):
        return _StubVideoReader._Frame(self.frames[int(i)])


_data_mod = None


def reference_data_module():
    """scripts/data_processor/data.py exec'd unmodified with stub decord/av and a writable __file__."""
    global _data_mod
    if _data_mod is not None:
        return _data_mod
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import transformers  # noqa: F401
    for name in ("decord", "av", "cv2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # noqa: BLE001
                stub = types.ModuleType(name)
                if name == "decord":
                    stub.VideoReader = _StubVideoReader
                    stub.cpu = lambda i=0: None
                sys.modules[name] = stub
    sys.modules["decord"].VideoReader = _StubVideoReader
    src_path = os.path.join(REFERENCE_ROOT, "scripts", "data_processor", "data.py")
    with open(src_path) as f:
        src = f.read()
    mod = types.ModuleType("reference_data")
    tmp = tempfile.mkdtemp(prefix="mjv_refdata_")
    mod.__file__ = os.path.join(tmp, "data.py")
    exec(compile(src, src_path, "exec"), mod.__dict__)
    _data_mod = mod
    return mod


def reference_load_video(frames: np.ndarray, **kw):
    """Run the reference's load_video (data.py:158-179) on an in-memory uint8 [F,H,W,3] 'video'."""
    mod = reference_data_module()
    key = f"/mem/video_{id(frames)}.mp4"
    _StubVideoReader.registry[key] = frames
    try:
        return mod.load_video(key, **kw)
    finally:
        _StubVideoReader.registry.pop(key, None)


_dataset_mod = None


def reference_dataset_module():
    """scripts/data_processor/dataset.py exec'd unmodified (VideoDataCollator :407-554) with stub `decord` and `boto3`
    (neither is installed here, and the collator uses neither) and scripts/model on sys.path for `internvl2.conversation`."""
    global _dataset_mod
    if _dataset_mod is not None:
        return _dataset_mod
    _import_reference()          # puts scripts/model on sys.path, registers the timm stub
    reference_data_module()      # registers the decord stub
    if "boto3" not in sys.modules:
        try:
            __import__("boto3")
        except Exception:  # noqa: BLE001
            b3 = types.ModuleType("boto3")
            sess = types.ModuleType("boto3.session")
            sess.Session = object
            b3.session = sess
            sys.modules["boto3"] = b3
            sys.modules["boto3.session"] = sess
    src_path = os.path.join(REFERENCE_ROOT, "scripts", "data_processor", "dataset.py")
    with open(src_path) as f:
        src = f.read()
    mod = types.ModuleType("reference_dataset")
    mod.__file__ = src_path
    mod.__name__ = "reference_dataset"      # the `if __name__ == "__main__"` block must not run
    exec(compile(src, src_path, "exec"), mod.__dict__)
    _dataset_mod = mod
    return mod
