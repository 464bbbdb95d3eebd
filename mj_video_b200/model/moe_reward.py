"""Drop-in for scripts/model/moe_reward.py: `InternVLChatRewardModeling`, its config and output container.

Same constructor, attributes, state-dict keys, `forward` signature, output fields/dtypes and exceptions as the reference
(moe_reward.py:60-89, 92-133, 138-297); the forward itself runs in hand-written sm_100a kernels behind the C ABI
(include/mjv_b200.h -> mjv_score_forward). There is no PyTorch fallback: without the CUDA library or a CUDA device,
`forward` raises.

Compute dtype: bf16 with fp32 accumulation, mirroring the rounding points of the reference's `.to(bfloat16)` path
(SURVEY §A.3). Outputs follow the reference dtypes: `aspect_scores`/`score` fp32, everything else the module dtype.
"""
from __future__ import annotations

import glob
import os
import weakref
from collections import OrderedDict
from typing import Optional

import torch
from torch import nn

from ..engine import Engine, EngineDims, applied_vit_layers, find_token_for_gating  # noqa: F401  (re-exported like the reference)
from .internvl2 import InternVLChatModel, InternVLChatRewardModelingConfig  # noqa: F401

token_pattern = [92542, 92543, 525, 11353, 364]  # moe_reward.py:48


class GatingNetwork(nn.Module):
    """Parameter container of the reference's gating MLP (moe_reward.py:16-42): `layers.{i}.weight/bias`,
    `logit_scale`, `temperature`. Its math runs inside the fused head kernels of `InternVLChatRewardModeling.forward`."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, temperature: float = 10,
                 logit_scale: float = 1., hidden_dim: int = 1024, n_hidden: int = 3):
        super().__init__()
        self.temperature = temperature
        self.logit_scale = nn.Parameter(torch.ones(1) * logit_scale)
        layers = []
        for _ in range(n_hidden):
            layers.append(nn.Linear(in_features, hidden_dim))
            in_features = hidden_dim
        layers.append(nn.Linear(in_features, out_features, bias=bias))
        self.layers = nn.ModuleList(layers)

    def _bind(self, owner, which: int):
        """Attach to the reward model whose engine holds this network's weights (0 = aspect, 1 = criteria)."""
        object.__setattr__(self, "_owner", weakref.ref(owner))
        object.__setattr__(self, "_which", which)

    def _run(self, x, with_softmax: bool):
        owner = getattr(self, "_owner", lambda: None)()
        if owner is None:
            raise RuntimeError("GatingNetwork runs through the engine of its InternVLChatRewardModeling (CUDA head "
                               "kernels); a free-standing instance has no CPU path")
        return owner.engine.gating(self._which, x, with_softmax).to(x.dtype)

    def forward(self, x):
        """softmax(MLP(x) / temperature, dim=1) * logit_scale[0] (moe_reward.py:29-36) — mjv_gating_forward."""
        return self._run(x, True)

    def forward_wo_softmax(self, x):
        """The raw logits of the last layer (moe_reward.py:38-42)."""
        return self._run(x, False)


class CustomOutput(OrderedDict):
    """Same fields, attribute access, key access and None-dropping as the reference's `ModelOutput` dataclass
    (moe_reward.py:60-89)."""
    _FIELDS = ("rewards", "hidden_state", "prompt_embedding", "criteria_gating_output", "aspect_gating_output",
               "aspect_weights", "score", "weighted_scores", "aspect_scores")

    def __init__(self, **kw):
        super().__init__()
        for k in kw:
            if k not in self._FIELDS:
                raise TypeError(f"unexpected field {k}")
        for k in self._FIELDS:
            v = kw.get(k)
            object.__setattr__(self, k, v)
            if v is not None:
                OrderedDict.__setitem__(self, k, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return OrderedDict.__getitem__(self, k)
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


def _engine_dims(config) -> EngineDims:
    vc, lc = config.vision_config, config.llm_config
    arch = lc.architectures[0]
    if arch not in ("InternLM2ForCausalLM", "LlamaForCausalLM"):
        raise NotImplementedError(f"{arch} is not implemented.")  # modeling_internvl_chat.py:125-130
    if vc.qk_normalization or vc.norm_type != "layer_norm" or not vc.qkv_bias:
        raise NotImplementedError("only the InternViT-300M flavour (layer_norm, qkv_bias, no qk_normalization) is built")
    if config.ps_version not in ("v1", "v2"):
        raise NotImplementedError(f"ps_version {config.ps_version!r} is unknown")
    rs = lc.rope_scaling
    if arch == "InternLM2ForCausalLM" and rs is not None and rs.get("type") not in ("dynamic", "linear"):
        # the reference raises here as well (modeling_internlm2.py:334-335)
        raise ValueError("Currently we only support rotary embedding's type being 'dynamic' or 'linear'.")
    return EngineDims(
        image_size=config.force_image_size or vc.image_size, patch_size=vc.patch_size, vit_hidden=vc.hidden_size,
        vit_heads=vc.num_attention_heads, vit_mlp=vc.intermediate_size, vit_layers=applied_vit_layers(vc.num_hidden_layers, config.select_layer),
        vit_eps=vc.layer_norm_eps, llm_arch=0 if arch == "InternLM2ForCausalLM" else 1, llm_hidden=lc.hidden_size,
        llm_inter=lc.intermediate_size, llm_layers=lc.num_hidden_layers, llm_heads=lc.num_attention_heads,
        llm_kv_heads=lc.num_key_value_heads, vocab=lc.vocab_size, rms_eps=lc.rms_norm_eps, rope_theta=lc.rope_theta,
        max_pos=lc.max_position_embeddings, num_objectives=config.num_objectives, num_aspects=config.num_aspects,
        gating_hidden=config.gating_hidden_dim, gating_n_hidden=config.gating_n_hidden,
        gating_temperature=config.gating_temperature,
        aspect2criteria={k: list(v) for k, v in config.aspect2criteria.items()},
        downsample_ratio=config.downsample_ratio,
        rope_scaling_factor=(rs or {}).get("factor") if arch == "InternLM2ForCausalLM" else None,
        ps_version=1 if config.ps_version == "v1" else 2,
        rope_scaling_type=(rs or {}).get("type", "dynamic") if arch == "InternLM2ForCausalLM" else "dynamic")


class InternVLChatRewardModeling(nn.Module):
    def __init__(self, name: str, config):
        super().__init__()
        self.num_labels = getattr(config, "num_labels", 2)
        self.model = InternVLChatModel(config)
        config_dict = config.to_dict()
        self.num_objectives = config_dict['num_objectives']
        self.num_aspects = config_dict['num_aspects']
        self.aspect2criteria = config_dict['aspect2criteria']
        # sanity checks of the reference (moe_reward.py:153-157)
        assert len(self.aspect2criteria) == self.num_aspects
        assert sum(len(v) for v in self.aspect2criteria.values()) == self.num_objectives
        temp = []
        for k in self.aspect2criteria.values():
            temp += k
        assert sum(len(set(v)) for v in self.aspect2criteria.values()) == len(set(temp))

        hidden_size = config.llm_config.hidden_size
        self.regression_layer = nn.Linear(hidden_size, self.num_objectives, bias=False)
        eye = torch.zeros(self.num_objectives, self.num_objectives)
        eye[range(self.num_objectives), range(self.num_objectives)] = 1.
        self.reward_transform_matrix = nn.Parameter(eye)
        self.reward_transform_matrix.requires_grad = False
        self.aspect_gating = GatingNetwork(hidden_size, self.num_aspects, temperature=config_dict["gating_temperature"],
                                           hidden_dim=config_dict["gating_hidden_dim"],
                                           n_hidden=config_dict["gating_n_hidden"])
        self.criteria_gating = GatingNetwork(hidden_size, config.num_objectives,
                                             temperature=config_dict["gating_temperature"],
                                             hidden_dim=config_dict["gating_hidden_dim"],
                                             n_hidden=config_dict["gating_n_hidden"])
        self.aspect_gating._bind(self, 0)
        self.criteria_gating._bind(self, 1)
        self.config = config
        self._engine: Engine | None = None
        self._engine_key = None
        # The reference builds the backbone with InternVLChatModel.from_pretrained(name) (moe_reward.py:142): it either
        # holds the checkpoint's weights or the constructor raises. Here the backbone parameters start as uninitialised
        # memory (there is no hub offline); every key stays in this set until a load_state_dict delivers it, and forward
        # refuses to run on a backbone that was never (or only partly) loaded instead of scoring garbage.
        self._backbone_unloaded = {"model." + k for k, _ in self.model.named_parameters()}
        self.register_load_state_dict_post_hook(self._note_loaded_keys(""))
        self.model.register_load_state_dict_post_hook(self._note_loaded_keys("model."))
        if name and os.path.isdir(str(name)):
            self._load_backbone_checkpoint(str(name))

    def _note_loaded_keys(self, prefix: str):
        def hook(module, incompatible):
            # the backbone's hook sees keys relative to it when it is loaded directly and absolute ones when the load came
            # through this module (its own children are vision_model / language_model / mlp1, never "model.")
            missing = {k if (not prefix or k.startswith(prefix)) else prefix + k for k in incompatible.missing_keys}
            own = {prefix + k for k, _ in module.named_parameters()}
            self._backbone_unloaded -= (own - missing)
        return hook

    def mark_backbone_initialised(self) -> None:
        """Declare that the backbone parameters were filled in place (not through load_state_dict)."""
        self._backbone_unloaded.clear()

    # ------------------------------------------------------------------ checkpoints (InternVLChatModel.from_pretrained)
    def _load_backbone_checkpoint(self, path: str):
        """`name` is a local directory: its *.safetensors must hold the complete backbone, as from_pretrained requires
        (a directory with only *.bin shards, or a partial checkpoint, raises instead of leaving parameters unset)."""
        files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
        if not files:
            raise FileNotFoundError(f"no *.safetensors backbone weights under {path}")
        from safetensors.torch import load_file
        sd = {}
        for f in files:
            sd.update(load_file(f))
        missing, unexpected = self.model.load_state_dict(sd, strict=False)
        if unexpected:
            raise RuntimeError(f"unexpected keys in backbone checkpoint: {unexpected[:5]}")
        if missing:
            raise RuntimeError(f"backbone checkpoint {path} lacks {len(missing)} tensors, e.g. {missing[:5]}")

    # ------------------------------------------------------------------ engine sync
    def _weights_key(self):
        ps = list(self.parameters())
        return (ps[0].device, ps[0].dtype, sum(p._version for p in ps), tuple(p.data_ptr() for p in ps[:4]))

    def release_module_weights(self) -> None:
        """Inference-only memory saver. The engine holds its own re-packed copy of every weight, so after the first forward
        a CUDA module carries the parameters twice (about 2 x 4.4 GB for MJ-VIDEO-2B). This moves the module's own
        parameters and buffers to host memory and keeps scoring through the engine that was built from them; `state_dict()`
        keeps working (CPU tensors). Moving the module back with `.cuda()` / loading new weights returns to the normal,
        parameter-tracking behaviour. The reference has no equivalent: its module IS the only copy."""
        eng = self._get_engine()
        for t in list(self.parameters()) + list(self.buffers()):
            t.data = t.data.to("cpu")
        self._released_engine = eng
        self._released_version = sum(p._version for p in self.parameters())

    def _get_engine(self) -> Engine:
        released = getattr(self, "_released_engine", None)
        if released is not None:
            if self.model.device.type == "cpu" and sum(p._version for p in self.parameters()) == self._released_version:
                return released
            self._released_engine = None   # moved back to a device or modified: rebuild from the parameters
        dev = self.model.device
        if self._backbone_unloaded:
            some = sorted(self._backbone_unloaded)[:3]
            raise RuntimeError(f"{len(self._backbone_unloaded)} backbone parameters were never loaded (e.g. {some}): load a "
                               "checkpoint with load_state_dict / a local `name` directory first (the reference's "
                               "from_pretrained would have loaded them or raised)")
        if dev.type != "cuda":
            raise RuntimeError("InternVLChatRewardModeling (mj_video_b200) needs a CUDA device: call .cuda() first; "
                               "there is no CPU path")
        key = self._weights_key()
        # a float32 module computes in fp32, as the reference's would (north_star's 1e-4 verification mode); bf16 / fp16
        # modules run the bf16 tensor-core path
        precision = "fp32" if self.model.dtype == torch.float32 else "bf16"
        if self._engine is None or self._engine.device != dev or self._engine.precision != precision:
            # a bf16 module mirrors the reference's `.to(bfloat16)` Llama branch, whose rotary inv_freq buffer is rounded too
            self._engine = Engine(_engine_dims(self.config), dev,
                                  rope_inv_freq_bf16=self.model.dtype == torch.bfloat16, precision=precision)
            self._engine_key = None
        if self._engine_key != key:
            self._engine.load_state_dict(self.state_dict())
            self._engine_key = key
        return self._engine

    @property
    def engine(self) -> Engine:
        """The CUDA engine holding this module's current weights (built / refreshed on demand)."""
        return self._get_engine()

    def forward(self, pixel_values: torch.FloatTensor, input_ids: torch.LongTensor = None,
                attention_mask: Optional[torch.Tensor] = None, position_ids=None, image_flags=None,
                past_key_values=None, labels=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict=None) -> CustomOutput:
        if input_ids is None:
            raise ValueError("input_ids is required")
        if position_ids is not None or past_key_values is not None:
            raise NotImplementedError("position_ids / past_key_values are not used by the scoring path")
        batch_size = input_ids.shape[0]
        pad_id = getattr(self.config, "pad_token_id", None)
        if pad_id is None and batch_size != 1:
            raise ValueError("Cannot handle batch sizes > 1 if no padding token is defined.")
        engine = self._get_engine()
        res = engine.forward(pixel_values, input_ids, attention_mask, pad_id, self.model.img_context_token_id)
        NO, NA = self.num_objectives, self.num_aspects
        dt = self.model.dtype
        s, e = res["scores"], res["extras"]
        return CustomOutput(
            rewards=s[:, :NO].to(dt),
            hidden_state=res["hidden_state"].to(dt),
            prompt_embedding=res["prompt_embedding"].to(dt),
            criteria_gating_output=e[:, :NO].to(dt),
            aspect_gating_output=e[:, NO:NO + NA].to(dt),
            aspect_weights=e[:, NO + NA:NO + NA + NO].to(dt),
            weighted_scores=e[:, NO + NA + NO].to(dt),
            aspect_scores=s[:, NO:NO + NA].contiguous(),
            score=s[:, NO + NA].contiguous(),
        )
