"""Seeded synthetic configs, weights, frames and token ids.

There is no network: no checkpoint, tokenizer or video is available, so every run (tests, smoke, bench, oracle) uses
these generators. Weights follow the reference's own initialisers (nn.Linear/Conv2d default uniform, randn CLS/pos-emb
`modeling_intern_vit.py:141-152`, layer-scale = initializer_factor, LLM N(0, 0.02²) `modeling_internlm2.py:714-723`)
with the state-dict key space of `InternVLChatRewardModeling`, so the very same tensors can be loaded
into the reference with `load_state_dict(strict=True)`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

# special token ids of InternLM2 / InternVL2-2B (moe_reward.py:46-48 pins 92542/92543; the rest are [ext])
IM_END, IM_START = 92542, 92543
IMG_START, IMG_END, IMG_CONTEXT = 92544, 92545, 92546
BOS, PAD = 1, 2
GATING_PATTERN = [92542, 92543, 525, 11353, 364]  # '<|im_end|><|im_start|>assistant\n' (moe_reward.py:48)

ASPECT2CRITERIA = {0: [0, 1, 2, 3, 4], 1: [5, 6, 7, 8, 9, 10], 2: [11, 12, 13, 14, 15],
                   3: [16, 17, 18, 19, 20, 21, 22], 4: [23, 24, 25, 26, 27]}  # eval_genai_mjvideo.py:38-44


@dataclass
class ModelDims:
    """Shapes of one synthetic model."""
    name: str = "mjvideo-2b"
    # vision (InternViT)
    image_size: int = 448
    patch_size: int = 14
    vit_hidden: int = 1024
    vit_heads: int = 16
    vit_mlp: int = 4096
    vit_layers: int = 24
    vit_eps: float = 1e-6
    vit_ls_init: float = 0.1
    # LLM (InternLM2, or the HF-Llama branch of modeling_internvl_chat.py:125-126)
    llm_arch: str = "InternLM2ForCausalLM"
    template: str = "internlm2-chat"
    llm_hidden: int = 2048
    llm_inter: int = 8192
    llm_layers: int = 24
    llm_heads: int = 16
    llm_kv_heads: int = 8
    vocab: int = 92553
    rms_eps: float = 1e-5
    rope_theta: float = 1000000.0
    max_pos: int = 32768
    rope_scaling_factor: float | None = 2.0   # rope_scaling {'type': 'dynamic', 'factor': 2.0} (InternLM2 branch only)
    rope_scaling_type: str = "dynamic"        # or 'linear' (modeling_internlm2.py:318-333); 'longrope' on the Llama-layout branch
    rope_short_factor: tuple | None = None    # Phi-3 LongRoPE (InternVL2-4B): per-frequency divisors, head_dim/2 entries each
    rope_long_factor: tuple | None = None
    rope_original_max_pos: int | None = None
    # projector
    downsample_ratio: float = 0.5
    select_layer: int = -1        # ViT hidden state fed to the projector (modeling_internvl_chat.py:244-254)
    ps_version: str = "v2"        # pixel-shuffle layout (modeling_internvl_chat.py:237-241)
    # head (eval_genai_mjvideo.py:36-47)
    num_objectives: int = 28
    num_aspects: int = 5
    gating_hidden: int = 1024
    gating_n_hidden: int = 3
    gating_temperature: float = 1.0
    aspect2criteria: dict = field(default_factory=lambda: {k: list(v) for k, v in ASPECT2CRITERIA.items()})

    @property
    def llm_head_dim(self) -> int:
        return self.llm_hidden // self.llm_heads

    @property
    def vit_head_dim(self) -> int:
        return self.vit_hidden // self.vit_heads

    @property
    def patches_per_tile(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def vit_seq(self) -> int:
        return self.patches_per_tile + 1

    @property
    def tokens_per_tile(self) -> int:
        return int(self.patches_per_tile * self.downsample_ratio ** 2)


def dims_2b() -> ModelDims:
    """MJ-VIDEO-2B = InternVL2-2B (InternViT-300M + InternLM2-1.8B) + MoE head (BASELINE.json configs 1-4)."""
    return ModelDims()


def dims_4b() -> ModelDims:
    """BASELINE.json configs[4]: InternViT-300M + a Phi-3-mini-*shaped* LLM run through the reference's Llama branch
    (SURVEY §A.1, §8(a) A12): d 3072, 32 layers, 32 heads x 96, MHA, plain RoPE theta 1e4. The vocabulary stays at
    92 553 because moe_reward.py:45-48 hard-codes InternLM2 token ids for the gating pattern."""
    return ModelDims(name="mjvideo-4b-shaped", llm_arch="LlamaForCausalLM", template="phi3-chat", llm_hidden=3072,
                     llm_inter=8192, llm_layers=32, llm_heads=32, llm_kv_heads=32, rope_theta=10000.0)


def dims_tiny_llama() -> ModelDims:
    """Small Llama-branch model with head_dim 96 (the config-5 head shape) and grouped KV heads."""
    return ModelDims(name="tiny-llama", llm_arch="LlamaForCausalLM", template="phi3-chat", vit_hidden=128, vit_heads=2,
                     vit_mlp=256, vit_layers=2, llm_hidden=384, llm_inter=512, llm_layers=2, llm_heads=4,
                     llm_kv_heads=2, gating_hidden=128, rope_theta=10000.0)


def dims_tiny_llama_mha() -> ModelDims:
    """dims_tiny_llama with 4 kv heads: head counts that are multiples of 4 take the native 96-wide q/k slot of the engine."""
    d = dims_tiny_llama()
    d.name, d.llm_kv_heads = "tiny-llama-mha", 4
    return d


def dims_tiny() -> ModelDims:
    """Small model with the same structure and the same ragged shapes (S=1025, head dims 64/128) for fast parity."""
    return ModelDims(name="tiny", vit_hidden=128, vit_heads=2, vit_mlp=256, vit_layers=2, llm_hidden=256,
                     llm_inter=512, llm_layers=2, llm_heads=2, llm_kv_heads=1, gating_hidden=128)


# --------------------------------------------------------------------------------------------- weights
def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def _linear(sd, g, prefix, out_f, in_f, bias=True):
    b = 1.0 / math.sqrt(in_f)
    sd[prefix + ".weight"] = _uniform(g, (out_f, in_f), b)
    if bias:
        sd[prefix + ".bias"] = _uniform(g, (out_f,), b)


def _norm(sd, g, prefix, n, bias=True):
    sd[prefix + ".weight"] = 1.0 + 0.1 * torch.randn((n,), generator=g)
    if bias:
        sd[prefix + ".bias"] = 0.1 * torch.randn((n,), generator=g)


def make_state_dict(d: ModelDims, seed: int = 1234, include_lm_head: bool = True) -> dict[str, torch.Tensor]:
    """fp32 CPU state dict with the reference's key space. Deterministic in (dims, seed, include_lm_head):
    the LM head draws from the same generator, so the projector / head tensors generated after it differ between the two
    settings — compare a model only with one built from the SAME dict."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: dict[str, torch.Tensor] = {}
    no = d.num_objectives
    sd["reward_transform_matrix"] = torch.eye(no)
    v = "model.vision_model."
    C = d.vit_hidden
    sd[v + "embeddings.class_embedding"] = torch.randn((1, 1, C), generator=g)
    sd[v + "embeddings.position_embedding"] = torch.randn((1, d.vit_seq, C), generator=g)
    fan_in = 3 * d.patch_size * d.patch_size
    sd[v + "embeddings.patch_embedding.weight"] = _uniform(g, (C, 3, d.patch_size, d.patch_size), 1 / math.sqrt(fan_in))
    sd[v + "embeddings.patch_embedding.bias"] = _uniform(g, (C,), 1 / math.sqrt(fan_in))
    for i in range(d.vit_layers):
        p = f"{v}encoder.layers.{i}."
        sd[p + "ls1"] = d.vit_ls_init * (1.0 + 0.2 * torch.randn((C,), generator=g))
        sd[p + "ls2"] = d.vit_ls_init * (1.0 + 0.2 * torch.randn((C,), generator=g))
        _linear(sd, g, p + "attn.qkv", 3 * C, C)
        _linear(sd, g, p + "attn.proj", C, C)
        _linear(sd, g, p + "mlp.fc1", d.vit_mlp, C)
        _linear(sd, g, p + "mlp.fc2", C, d.vit_mlp)
        _norm(sd, g, p + "norm1", C)
        _norm(sd, g, p + "norm2", C)
    lm = "model.language_model."
    D = d.llm_hidden
    hd = d.llm_head_dim
    llama = d.llm_arch == "LlamaForCausalLM"
    # key names of the two LLM branches (modeling_internlm2.py vs transformers LlamaForCausalLM)
    names = dict(emb="model.embed_tokens", wo="self_attn.o_proj", w1="mlp.gate_proj", w3="mlp.up_proj",
                 w2="mlp.down_proj", n1="input_layernorm", n2="post_attention_layernorm", head="lm_head") if llama else \
        dict(emb="model.tok_embeddings", wo="attention.wo", w1="feed_forward.w1", w3="feed_forward.w3",
             w2="feed_forward.w2", n1="attention_norm", n2="ffn_norm", head="output")
    sd[lm + names["emb"] + ".weight"] = 0.02 * torch.randn((d.vocab, D), generator=g)
    for i in range(d.llm_layers):
        p = f"{lm}model.layers.{i}."
        if llama:
            sd[p + "self_attn.q_proj.weight"] = 0.02 * torch.randn((d.llm_heads * hd, D), generator=g)
            sd[p + "self_attn.k_proj.weight"] = 0.02 * torch.randn((d.llm_kv_heads * hd, D), generator=g)
            sd[p + "self_attn.v_proj.weight"] = 0.02 * torch.randn((d.llm_kv_heads * hd, D), generator=g)
        else:
            sd[p + "attention.wqkv.weight"] = 0.02 * torch.randn(((d.llm_heads + 2 * d.llm_kv_heads) * hd, D), generator=g)
        sd[p + names["wo"] + ".weight"] = 0.02 * torch.randn((D, d.llm_heads * hd), generator=g)
        sd[p + names["w1"] + ".weight"] = 0.02 * torch.randn((d.llm_inter, D), generator=g)
        sd[p + names["w3"] + ".weight"] = 0.02 * torch.randn((d.llm_inter, D), generator=g)
        sd[p + names["w2"] + ".weight"] = 0.02 * torch.randn((D, d.llm_inter), generator=g)
        _norm(sd, g, p + names["n1"], D, bias=False)
        _norm(sd, g, p + names["n2"], D, bias=False)
    _norm(sd, g, lm + "model.norm", D, bias=False)
    if include_lm_head:
        sd[lm + names["head"] + ".weight"] = 0.02 * torch.randn((d.vocab, D), generator=g)
    c4 = C * int(1 / d.downsample_ratio) ** 2
    _norm(sd, g, "model.mlp1.0", c4)
    _linear(sd, g, "model.mlp1.1", D, c4)
    _linear(sd, g, "model.mlp1.3", D, D)
    # head: nn.Linear defaults; regression scaled so |rewards| <~ 0.45 (a trained +-1 regressor; SURVEY §8(d).1: x0.25), last gating layers
    # scaled so logits are O(1) and the softmax/argmax is well separated (SURVEY §7 hard parts, §8(d).1)
    sd["regression_layer.weight"] = _uniform(g, (no, D), 1.0 / math.sqrt(D)) * 0.25
    for name, n_out in (("aspect_gating", d.num_aspects), ("criteria_gating", no)):
        sd[f"{name}.logit_scale"] = torch.ones(1)
        in_f = D
        for li in range(d.gating_n_hidden):
            _linear(sd, g, f"{name}.layers.{li}", d.gating_hidden, in_f)
            in_f = d.gating_hidden
        _linear(sd, g, f"{name}.layers.{d.gating_n_hidden}", n_out, in_f)
        sd[f"{name}.layers.{d.gating_n_hidden}.weight"] *= 24.0
    return sd


# --------------------------------------------------------------------------------------------- frames
def make_video(video_id: int, n_frames: int = 16, height: int = 448, width: int = 448,
               structured: bool = True) -> np.ndarray:
    """uint8 [n_frames, H, W, 3]. Structured mode: base colour + drifting discs + sinusoidal grating, so that two
    videos differ visibly to the model (iid noise videos are nearly indistinguishable: SURVEY §8(d).1)."""
    rng = np.random.default_rng(video_id)
    if not structured:
        return rng.integers(0, 256, size=(n_frames, height, width, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    base = rng.uniform(30, 220, size=3).astype(np.float32)
    n_disc = 6
    cx = rng.uniform(0, width, n_disc)
    cy = rng.uniform(0, height, n_disc)
    rad = rng.uniform(0.05, 0.2, n_disc) * min(height, width)
    col = rng.uniform(0, 255, (n_disc, 3)).astype(np.float32)
    vel = rng.uniform(-6, 6, (n_disc, 2))
    fx, fy = rng.uniform(0.01, 0.08, 2)
    phase = rng.uniform(0, 2 * np.pi)
    amp = rng.uniform(10, 40)
    out = np.empty((n_frames, height, width, 3), dtype=np.uint8)
    for f in range(n_frames):
        img = np.broadcast_to(base, (height, width, 3)).copy()
        img += (amp * np.sin(fx * xx + fy * yy + phase + 0.2 * f))[..., None]
        for k in range(n_disc):
            m = (xx - (cx[k] + vel[k, 0] * f)) ** 2 + (yy - (cy[k] + vel[k, 1] * f)) ** 2 < rad[k] ** 2
            img[m] = col[k]
        img += rng.normal(0, 2.0, img.shape).astype(np.float32)
        out[f] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


# --------------------------------------------------------------------------------------------- token ids
def make_input_ids(sample_id: int, tiles_per_frame: list[int], tokens_per_tile: int = 256,
                   n_cap: int | None = None, first_slot_layout: bool = False) -> list[int]:
    """Synthetic prompt with the structure `prepare_chat_input` produces:
    <s><|im_start|> sys(42) <|im_end|><|im_start|> usr(2) { lab(2) <img> <IMG_CONTEXT>*256*tiles </img> \\n }*frames
    caption(n_cap) <|im_end|><|im_start|>assistant\\n.
    first_slot_layout=True puts every IMG_CONTEXT token into the first frame slot (what the reference's own callers do,
    modeling_internvl_chat.py:47-48) — same counts, same FLOPs."""
    rng = np.random.default_rng(1_000_003 * 7 + sample_id)
    free = lambda n: rng.integers(1000, 90000, size=n).tolist()  # noqa: E731
    if n_cap is None:
        n_cap = int(rng.integers(16, 97))
    ids = [BOS, IM_START] + free(42) + [IM_END, IM_START] + free(2)
    total_tiles = sum(tiles_per_frame)
    for f, tiles in enumerate(tiles_per_frame):
        ids += free(2) + [IMG_START]
        if first_slot_layout:
            ids += [IMG_CONTEXT] * (tokens_per_tile * total_tiles) if f == 0 else free(3)
        else:
            ids += [IMG_CONTEXT] * (tokens_per_tile * tiles)
        ids += [IMG_END, 364]
    ids += free(n_cap) + GATING_PATTERN
    return ids


def pad_batch(rows: list[list[int]], pad_id: int = PAD) -> tuple[torch.Tensor, torch.Tensor]:
    """Right-pad to the batch max (notebook `pad_to_batch`, dataset.py:464-469): ids padded with pad_id, mask 0."""
    n = max(len(r) for r in rows)
    ids = torch.full((len(rows), n), pad_id, dtype=torch.long)
    mask = torch.zeros((len(rows), n), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = torch.tensor(r, dtype=torch.long)
        mask[i, : len(r)] = 1
    return ids, mask
