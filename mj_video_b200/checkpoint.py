"""Checkpoint ingestion (SURVEY §8(f) N4, §5 "Checkpoint / resume"): the reference hands weights between training stages
and to the eval driver as `*.safetensors` files whose tensor names are the module's state-dict keys
(scripts/train/overall_train.py:503-522 loads the first *.safetensors of a directory with strict=True;
scripts/eval/eval_genai_mjvideo.py:91-110 does the same for scoring). These helpers stream such files into the drop-in
module or straight into an Engine, tensor by tensor (no second full copy of the model in host memory)."""
from __future__ import annotations

import glob
import os


def checkpoint_files(path: str) -> list[str]:
    """A file, or every *.safetensors shard of a directory in name order (the reference takes the first one only; sharded
    checkpoints simply contribute disjoint key sets)."""
    if os.path.isdir(path):
        files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
        if not files:
            raise FileNotFoundError(f"no *.safetensors under {path}")
        return files
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return [path]


def iter_tensors(path: str):
    """Yield (key, CPU tensor) for every tensor of the checkpoint, one at a time."""
    from safetensors import safe_open
    for f in checkpoint_files(path):
        with safe_open(f, framework="pt", device="cpu") as sf:
            for k in sf.keys():
                yield k, sf.get_tensor(k)


def load_into_engine(engine, path: str, skip_lm_head: bool = True) -> int:
    """Stream a checkpoint into an Engine (repacked to bf16 on the device as it goes); finalises the weights.
    Returns the number of tensors consumed."""
    n = 0
    for k, t in iter_tensors(path):
        if skip_lm_head and (k.endswith("language_model.output.weight") or k.endswith("language_model.lm_head.weight")):
            continue  # never read by the reward path (moe_reward.py:211 uses hidden_states only)
        engine.load_weight(k, t)
        n += 1
    engine.finalize()
    return n


def load_into_module(model, path: str, strict: bool = True):
    """`model.load_state_dict(load_file(ckpt), strict=True)` as the reference's drivers do."""
    sd = {}
    for k, t in iter_tensors(path):
        sd[k] = t
    return model.load_state_dict(sd, strict=strict)


def save_state_dict(state_dict: dict, path: str) -> None:
    """Write a state dict in the reference's interchange format (used by tests and examples)."""
    from safetensors.torch import save_file
    save_file({k: v.contiguous() for k, v in state_dict.items()}, path)
