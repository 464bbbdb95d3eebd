"""`prepare_chat_input` and the parameter containers of the VLM backbone.

Reference: scripts/model/internvl2/modeling_internvl_chat.py:36-89 (prepare_chat_input), :100-144 (module tree).
The containers reproduce the reference's state-dict key space so `load_state_dict(strict=True)`,
`.to(dtype)`, `.cuda()` and `state_dict()` behave like the reference module; they hold parameters only — all compute is
in the CUDA library behind `InternVLChatRewardModeling.forward`.
"""
from __future__ import annotations

import weakref

import torch
from torch import nn

from .conversation import get_conv_template


def prepare_chat_input(config, tokenizer, pixel_values, question, generation_config, history=None,
                       return_history=False, num_patches_list=None, IMG_START_TOKEN='<img>', IMG_END_TOKEN='</img>',
                       IMG_CONTEXT_TOKEN='<IMG_CONTEXT>', verbose=False, device='cpu'):
    """Build the chat prompt, expand each `<image>` into `<img>` + `<IMG_CONTEXT>`*(tokens*patches) + `</img>` and
    tokenise it. Same signature, return value and side effect (`generation_config['eos_token_id']`) as the reference."""
    if history is None and pixel_values is not None and '<image>' not in question:
        question = '<image>\n' + question
    if num_patches_list is None:
        num_patches_list = [pixel_values.shape[0]] if pixel_values is not None else []
    assert pixel_values is None or len(pixel_values) == sum(num_patches_list)

    template = get_conv_template(config.template)
    eos_token_id = tokenizer.convert_tokens_to_ids(template.sep.strip())
    for (old_question, old_answer) in ([] if history is None else history):
        template.append_message(template.roles[0], old_question)
        template.append_message(template.roles[1], old_answer)
    template.append_message(template.roles[0], question)
    template.append_message(template.roles[1], None)
    query = template.get_prompt()

    image_size = config.force_image_size or config.vision_config.image_size
    patch_size = config.vision_config.patch_size
    num_image_token = int((image_size // patch_size) ** 2 * (config.downsample_ratio ** 2))
    counts = [num_image_token * num_patches for num_patches in num_patches_list]

    def expand(text, n_ctx_of):
        for i in range(len(num_patches_list)):
            text = text.replace('<image>', IMG_START_TOKEN + IMG_CONTEXT_TOKEN * n_ctx_of(i) + IMG_END_TOKEN, 1)
        return text

    # SURVEY §8(f) N3, host-side only. (1) The expanded prompt is ~30 KB of text for 8 frames, nearly all of it the
    # <IMG_CONTEXT> runs. Special tokens are split off atomically before a tokenizer sees the text between them, so a run of
    # n <IMG_CONTEXT> tokens yields n copies of one id whatever its neighbours are: the prompt is tokenised with ONE context
    # token per image and the id is repeated afterwards. This is verified, per tokenizer object, against the full
    # tokenisation on its first calls (and on every 64th after that); one mismatch switches that tokenizer back to the
    # reference's full-text call for good. (2) The chosen and the rejected row of a pair share the prompt verbatim: the last
    # few results are memoised per tokenizer. Both give exactly the ids of modeling_internvl_chat.py:36-89.
    short = expand(query, lambda i: 1)
    key = (id(tokenizer), short, tuple(counts))
    hit = _TOKENIZED.get(key)
    if hit is not None and hit[2] is not None and hit[2]() is not tokenizer:
        hit = None  # the id was recycled by another tokenizer object
    if hit is None:
        ids_mask = _tokenize_collapsed(tokenizer, short, counts, IMG_CONTEXT_TOKEN, lambda: expand(query, lambda i: counts[i]))
        try:
            ref = weakref.ref(tokenizer)
        except TypeError:
            ref = None
        hit = (ids_mask[0], ids_mask[1], ref)
        if len(_TOKENIZED) >= _TOKENIZED_MAX:
            _TOKENIZED.pop(next(iter(_TOKENIZED)))
        _TOKENIZED[key] = hit
    input_ids = hit[0].clone().to(device)
    attention_mask = hit[1].clone().to(device)
    generation_config['eos_token_id'] = eos_token_id
    return input_ids, attention_mask


_TOKENIZED: dict = {}      # (id(tokenizer), collapsed prompt, counts) -> (input_ids, attention_mask), insertion-ordered, bounded
_TOKENIZED_MAX = 32
_COLLAPSE_STATE: dict = {}  # id(tokenizer) -> [weakref | None, calls, trusted]
_VERIFY_FIRST, _VERIFY_EVERY = 2, 64


def _tokenize_collapsed(tokenizer, short, counts, ctx_token, full_text):
    """ids / mask of the full prompt. Fast path: tokenise `short` (one context token per image) and repeat the context id
    `counts[i]` times; falls back to tokenizer(full_text()) whenever that cannot be shown to be identical."""
    def slow():
        mi = tokenizer(full_text(), return_tensors='pt')
        return mi['input_ids'].clone(), mi['attention_mask'].clone()

    st = _COLLAPSE_STATE.get(id(tokenizer))
    if st is not None and st[0] is not None and st[0]() is not tokenizer:
        st = None
    if st is None:
        try:
            ref = weakref.ref(tokenizer)
        except TypeError:
            ref = None
        st = _COLLAPSE_STATE[id(tokenizer)] = [ref, 0, True]
    if not st[2] or not counts:
        return slow()
    ctx_id = tokenizer.convert_tokens_to_ids(ctx_token)
    mi = tokenizer(short, return_tensors='pt')
    ids, mask = mi['input_ids'], mi['attention_mask']
    pos = (ids[0] == ctx_id).nonzero().flatten()
    if ids.dim() != 2 or ids.shape[0] != 1 or pos.numel() != len(counts):
        return slow()  # e.g. the question itself contains the context token: not our case to be clever about
    rep = torch.ones(ids.shape[1], dtype=torch.long)
    rep[pos] = torch.tensor(counts, dtype=torch.long)
    fast = (torch.repeat_interleave(ids[0], rep)[None].clone(), torch.repeat_interleave(mask[0], rep)[None].clone())
    st[1] += 1
    if st[1] <= _VERIFY_FIRST or st[1] % _VERIFY_EVERY == 0:
        ref_ids, ref_mask = slow()
        if ref_ids.shape != fast[0].shape or not torch.equal(ref_ids, fast[0]) or not torch.equal(ref_mask, fast[1]):
            st[2] = False  # this tokenizer lets the run interact with its surroundings: never collapse for it again
            return ref_ids, ref_mask
    return fast


# ------------------------------------------------------------------------------------------------ parameter containers
class _Linear(nn.Module):
    def __init__(self, in_f, out_f, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_f, in_f), requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_f), requires_grad=False)


class _Norm(nn.Module):
    def __init__(self, n, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n), requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.empty(n), requires_grad=False)


class _Empty(nn.Module):
    pass


class _PatchEmbedding(nn.Module):
    def __init__(self, c_out, ps):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c_out, 3, ps, ps), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(c_out), requires_grad=False)


def _vision_model(vc) -> nn.Module:
    C = vc.hidden_size
    vm = _Empty()
    emb = _Empty()
    n_pos = (vc.image_size // vc.patch_size) ** 2 + 1
    emb.class_embedding = nn.Parameter(torch.empty(1, 1, C), requires_grad=False)
    emb.position_embedding = nn.Parameter(torch.empty(1, n_pos, C), requires_grad=False)
    emb.patch_embedding = _PatchEmbedding(C, vc.patch_size)
    vm.embeddings = emb
    enc = _Empty()
    layers = []
    for _ in range(vc.num_hidden_layers):
        layer = _Empty()
        layer.ls1 = nn.Parameter(torch.empty(C), requires_grad=False)
        layer.ls2 = nn.Parameter(torch.empty(C), requires_grad=False)
        layer.attn = _Empty()
        layer.attn.qkv = _Linear(C, 3 * C, bias=vc.qkv_bias)
        layer.attn.proj = _Linear(C, C)
        layer.mlp = _Empty()
        layer.mlp.fc1 = _Linear(C, vc.intermediate_size)
        layer.mlp.fc2 = _Linear(vc.intermediate_size, C)
        layer.norm1 = _Norm(C)
        layer.norm2 = _Norm(C)
        layers.append(layer)
    enc.layers = nn.ModuleList(layers)
    vm.encoder = enc
    return vm


def _language_model(lc) -> nn.Module:
    D = lc.hidden_size
    hd = D // lc.num_attention_heads
    lm = _Empty()
    core = _Empty()
    layers = []
    llama = lc.architectures[0] == "LlamaForCausalLM"
    if llama:
        core.embed_tokens = _Linear(D, lc.vocab_size, bias=False)
    else:
        core.tok_embeddings = _Linear(D, lc.vocab_size, bias=False)
    for _ in range(lc.num_hidden_layers):
        layer = _Empty()
        if llama:
            layer.self_attn = _Empty()
            layer.self_attn.q_proj = _Linear(D, lc.num_attention_heads * hd, bias=False)
            layer.self_attn.k_proj = _Linear(D, lc.num_key_value_heads * hd, bias=False)
            layer.self_attn.v_proj = _Linear(D, lc.num_key_value_heads * hd, bias=False)
            layer.self_attn.o_proj = _Linear(lc.num_attention_heads * hd, D, bias=False)
            layer.mlp = _Empty()
            layer.mlp.gate_proj = _Linear(D, lc.intermediate_size, bias=False)
            layer.mlp.up_proj = _Linear(D, lc.intermediate_size, bias=False)
            layer.mlp.down_proj = _Linear(lc.intermediate_size, D, bias=False)
            layer.input_layernorm = _Norm(D, bias=False)
            layer.post_attention_layernorm = _Norm(D, bias=False)
        else:
            layer.attention = _Empty()
            layer.attention.wqkv = _Linear(D, (lc.num_attention_heads + 2 * lc.num_key_value_heads) * hd, bias=False)
            layer.attention.wo = _Linear(lc.num_attention_heads * hd, D, bias=False)
            layer.feed_forward = _Empty()
            layer.feed_forward.w1 = _Linear(D, lc.intermediate_size, bias=False)
            layer.feed_forward.w3 = _Linear(D, lc.intermediate_size, bias=False)
            layer.feed_forward.w2 = _Linear(lc.intermediate_size, D, bias=False)
            layer.attention_norm = _Norm(D, bias=False)
            layer.ffn_norm = _Norm(D, bias=False)
        layers.append(layer)
    core.layers = nn.ModuleList(layers)
    core.norm = _Norm(D, bias=False)
    lm.model = core
    # the LM head is part of the reference state dict (strict loading) but never used by the reward path
    if llama:
        lm.lm_head = _Linear(D, lc.vocab_size, bias=False)
    else:
        lm.output = _Linear(D, lc.vocab_size, bias=False)
    return lm


class InternVLChatModel(nn.Module):
    """Parameter container with the reference's attribute surface: `.vision_model`, `.language_model`, `.mlp1`,
    `.img_context_token_id`, `.device`, `.template`, `.num_image_token`, `.downsample_ratio`, `.ps_version`."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        image_size = config.force_image_size or config.vision_config.image_size
        patch_size = config.vision_config.patch_size
        self.patch_size = patch_size
        self.select_layer = config.select_layer
        self.template = config.template
        self.num_image_token = int((image_size // patch_size) ** 2 * (config.downsample_ratio ** 2))
        self.downsample_ratio = config.downsample_ratio
        self.ps_version = config.ps_version
        vc = config.vision_config
        if config.force_image_size:
            vc.image_size = config.force_image_size
        self.vision_model = _vision_model(vc)
        self.language_model = _language_model(config.llm_config)
        C4 = vc.hidden_size * int(1 / self.downsample_ratio) ** 2
        D = config.llm_config.hidden_size
        self.mlp1 = nn.Sequential(_Norm(C4), _Linear(C4, D), nn.GELU(), _Linear(D, D))
        self.img_context_token_id = None
        self.conv_template = get_conv_template(self.template) if self.template else None
        self.system_message = self.conv_template.system_message if self.conv_template else None

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype
