"""Configuration objects carrying the fields the kernels are specialised on. Same attribute names as the reference's
HF config classes (scripts/model/internvl2/configuration_{intern_vit,internlm2,internvl_chat}.py and
scripts/model/moe_reward.py:92-133) so caller code that touches `config.llm_config.hidden_size`,
`config.vision_config.*`, `config.pad_token_id`, `config.to_dict()` keeps working. They are plain Python objects (no
hub access); `from_pretrained` reads a local `config.json`."""
from __future__ import annotations

import copy
import json
import os


class _Cfg:
    def to_dict(self) -> dict:
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.to_dict() if isinstance(v, _Cfg) else copy.deepcopy(v)
        return out

    def __repr__(self):
        return f"{type(self).__name__}({json.dumps(self.to_dict(), default=str, ensure_ascii=False)[:400]}...)"


class InternVisionConfig(_Cfg):
    """Defaults follow configuration_intern_vit.py:64-84 (6B ViT); real values come from the checkpoint config."""
    model_type = "intern_vit_6b"

    def __init__(self, num_channels=3, patch_size=14, image_size=224, qkv_bias=False, hidden_size=3200,
                 num_attention_heads=25, intermediate_size=12800, qk_normalization=True, num_hidden_layers=48,
                 use_flash_attn=True, hidden_act="gelu", norm_type="rms_norm", layer_norm_eps=1e-6, dropout=0.0,
                 drop_path_rate=0.0, attention_dropout=0.0, initializer_range=0.02, initializer_factor=0.1, **kwargs):
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.dropout = dropout
        self.drop_path_rate = drop_path_rate
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.num_channels = num_channels
        self.patch_size = patch_size
        self.image_size = image_size
        self.initializer_range = initializer_range
        self.initializer_factor = initializer_factor
        self.attention_dropout = attention_dropout
        self.layer_norm_eps = layer_norm_eps
        self.hidden_act = hidden_act
        self.norm_type = norm_type
        self.qkv_bias = qkv_bias
        self.qk_normalization = qk_normalization
        self.use_flash_attn = use_flash_attn


class InternLM2Config(_Cfg):
    """Defaults follow configuration_internlm2.py:77-98 (7B); also used for LlamaForCausalLM-shaped LLMs."""
    model_type = "internlm2"

    def __init__(self, vocab_size=103168, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                 num_attention_heads=32, num_key_value_heads=None, hidden_act="silu", max_position_embeddings=2048,
                 initializer_range=0.02, rms_norm_eps=1e-6, use_cache=True, pad_token_id=0, bos_token_id=1,
                 eos_token_id=2, tie_word_embeddings=False, bias=True, rope_theta=10000, rope_scaling=None,
                 attn_implementation="eager", architectures=None, **kwargs):
        self.vocab_size = vocab_size
        self.max_position_embeddings = max_position_embeddings
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.bias = bias
        self.num_key_value_heads = num_attention_heads if num_key_value_heads is None else num_key_value_heads
        self.hidden_act = hidden_act
        self.initializer_range = initializer_range
        self.rms_norm_eps = rms_norm_eps
        self.use_cache = use_cache
        self.rope_theta = rope_theta
        self.rope_scaling = rope_scaling
        self.attn_implementation = attn_implementation
        self.pad_token_id = pad_token_id
        self.bos_token_id = bos_token_id
        self.eos_token_id = eos_token_id
        self.tie_word_embeddings = tie_word_embeddings
        self.architectures = architectures or ["InternLM2ForCausalLM"]


class InternVLChatConfig(_Cfg):
    """configuration_internvl_chat.py:19-96."""
    model_type = "internvl_chat"

    def __init__(self, vision_config=None, llm_config=None, use_backbone_lora=0, use_llm_lora=0, select_layer=-1,
                 force_image_size=None, downsample_ratio=0.5, template=None, dynamic_image_size=False,
                 use_thumbnail=False, ps_version="v1", min_dynamic_patch=1, max_dynamic_patch=6, pad_token_id=None,
                 **kwargs):
        vision_config = vision_config if vision_config is not None else {}
        llm_config = llm_config if llm_config is not None else {}
        if isinstance(vision_config, _Cfg):
            vision_config = vision_config.to_dict()
        if isinstance(llm_config, _Cfg):
            llm_config = llm_config.to_dict()
        arch = (llm_config.get("architectures") or ["InternLM2ForCausalLM"])[0]
        if arch not in ("LlamaForCausalLM", "InternLM2ForCausalLM"):
            raise ValueError("Unsupported architecture: {}".format(arch))  # configuration_internvl_chat.py:50-55
        self.vision_config = InternVisionConfig(**vision_config)
        self.llm_config = InternLM2Config(**llm_config)
        self.use_backbone_lora = use_backbone_lora
        self.use_llm_lora = use_llm_lora
        self.select_layer = select_layer
        self.force_image_size = force_image_size
        self.downsample_ratio = downsample_ratio
        self.template = template
        self.dynamic_image_size = dynamic_image_size
        self.use_thumbnail = use_thumbnail
        self.ps_version = ps_version
        self.min_dynamic_patch = min_dynamic_patch
        self.max_dynamic_patch = max_dynamic_patch
        self.pad_token_id = pad_token_id
        self.num_labels = kwargs.get("num_labels", 2)
        self.use_return_dict = True

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, **kwargs):
        path = os.path.join(str(pretrained_model_name_or_path), "config.json")
        if not os.path.isfile(path):
            raise OSError(f"{path} not found: mj_video_b200 reads local checkpoints only (no hub access)")
        with open(path) as f:
            data = json.load(f)
        data.update(kwargs)
        return cls(**data)


class InternVLChatRewardModelingConfig(InternVLChatConfig):
    """moe_reward.py:92-133 — adds the six reward-head hyper-parameters."""

    def __init__(self, internVLChatConfigName_or_path=None, **kwargs):
        super().__init__(**kwargs)
        self.num_objectives = kwargs.get("num_objectives", 0)
        self.num_aspects = kwargs.get("num_aspects", 0)
        self.aspect2criteria = kwargs.get("aspect2criteria", {})
        self.gating_temperature = kwargs.get("gating_temperature", 1.0)
        self.gating_hidden_dim = kwargs.get("gating_hidden_dim", 1024)
        self.gating_n_hidden = kwargs.get("gating_n_hidden", 3)
