"""TEST INFRASTRUCTURE. Generates tests/golden/* by running the UNMODIFIED reference (imported from /root/reference)
on seeded synthetic inputs. Run in the build container:  python -m oracle.make_golden
Inputs are NOT stored: they are regenerated from the seeds by mj_video_b200.synth (weights seed / video ids / sample ids
are recorded in each fixture)."""
from __future__ import annotations

import hashlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_video_b200 import synth  # noqa: E402
from oracle import reference_harness as rh  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class StubTokenizer:
    """Deterministic stand-in for the (unavailable) InternLM2 tokenizer: special tokens get their real ids, every other
    character maps to 1000 + ord(c) % 80000."""
    SPECIAL = {"<|im_end|>": 92542, "<|im_start|>": 92543, "<img>": 92544, "</img>": 92545, "<IMG_CONTEXT>": 92546}

    def convert_tokens_to_ids(self, tok):
        return self.SPECIAL.get(tok, 0)

    def __call__(self, text, return_tensors="pt"):
        ids, i = [1], 0
        while i < len(text):
            for s, v in self.SPECIAL.items():
                if text.startswith(s, i):
                    ids.append(v)
                    i += len(s)
                    break
            else:
                ids.append(1000 + ord(text[i]) % 80000)
                i += 1
        t = torch.tensor([ids], dtype=torch.long)
        return {"input_ids": t, "attention_mask": torch.ones_like(t)}


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def main():
    os.makedirs(GOLD, exist_ok=True)
    assert rh.available(), "reference tree missing"
    # ---------------- forward goldens (tiny dims, B = 3 with ragged lengths)
    d = synth.dims_tiny()
    seed = 1
    sd = synth.make_state_dict(d, seed=seed)
    m = rh.build_reference_model(d, sd)
    video_ids = [41, 42, 43]
    n_caps = [46, 17, 90]
    pvs = [rh.reference_load_video(synth.make_video(v), num_segments=8, max_num=1)[0] for v in video_ids]
    pv = torch.cat(pvs)
    rows = [synth.make_input_ids(i, [1] * 8, n_cap=n) for i, n in enumerate(n_caps)]
    ids, mask = synth.pad_batch(rows)
    out = rh.reference_forward(m, pv, ids, mask)
    fix = dict(dims="tiny", seed=seed, video_ids=video_ids, n_caps=n_caps, pixel_sha256=sha(pv),
               outputs={k: v.float().clone() for k, v in out.items()})
    # batch-1 call of the first row (exercises the no-padding path)
    out1 = rh.reference_forward(m, pvs[0], ids[:1, : len(rows[0])], mask[:1, : len(rows[0])])
    fix["outputs_row0_alone"] = {k: v.float().clone() for k, v in out1.items()}
    torch.save(fix, os.path.join(GOLD, "tiny_forward.pt"))

    # ---------------- Llama branch (modeling_internvl_chat.py:125-126), head_dim 96 (BASELINE configs[4] head shape):
    # fp32 forward and the reference's own bf16 forward (whose `.to(bfloat16)` also rounds the rotary inv_freq buffer)
    dl = synth.dims_tiny_llama()
    sdl = synth.make_state_dict(dl, seed=seed)
    fixl = dict(dims="tiny-llama", seed=seed, video_ids=video_ids, n_caps=n_caps, pixel_sha256=sha(pv))
    for name, dt in (("outputs", torch.float32), ("outputs_bf16", torch.bfloat16)):
        ml = rh.build_reference_model(dl, sdl, dtype=dt)
        o = rh.reference_forward(ml, pv.to(dt), ids, mask)
        fixl[name] = {k: v.float().clone() for k, v in o.items()}
    torch.save(fixl, os.path.join(GOLD, "tiny_llama_forward.pt"))

    # ---------------- frame-path goldens: checksums + a few probes of the reference's load_video output
    frames = {}
    for name, (H, W, mx, nseg) in {"448x448_m1": (448, 448, 1, 8), "896x1344_m6": (896, 1344, 6, 2),
                                    "1024x1024_m6": (1024, 1024, 6, 2), "720x1280_m6": (720, 1280, 6, 2),
                                    "1280x720_m6": (1280, 720, 6, 2), "320x512_m6": (320, 512, 6, 2),
                                    "100x77_m1": (100, 77, 1, 2), "500x300_m4": (500, 300, 4, 2)}.items():
        v = synth.make_video(7, n_frames=16 if nseg == 8 else 4, height=H, width=W)
        pvx, counts = rh.reference_load_video(v, num_segments=nseg, max_num=mx)
        probe = pvx.flatten()[:: max(1, pvx.numel() // 64)][:64].tolist()
        frames[name] = dict(H=H, W=W, max_num=mx, num_segments=nseg, video_id=7, n_frames=16 if nseg == 8 else 4,
                            counts=counts, shape=list(pvx.shape), sha256=sha(pvx), probe=probe)
    with open(os.path.join(GOLD, "frames.json"), "w") as f:
        json.dump(frames, f, indent=1)

    # ---------------- prepare_chat_input golden (stub tokenizer)
    _, _, chat = rh.reference_modules()
    cfg = rh.build_reference_config(d)
    tok = StubTokenizer()
    cases = {}
    for name, (n_tiles, npl, q) in {"first_slot": (8, None, "".join(f"Frame{i+1}: <image>\n" for i in range(8)) + "a cat"),
                                     "per_frame": (8, [1] * 8, "".join(f"Frame{i+1}: <image>\n" for i in range(8)) + "a dog runs"),
                                     "no_image_tag": (2, None, "describe")}.items():
        gen = {}
        ids_, mask_ = chat.prepare_chat_input(cfg, tok, torch.zeros(n_tiles, 3, 448, 448), q, gen, num_patches_list=npl)
        cases[name] = dict(n_tiles=n_tiles, num_patches_list=npl, question=q, eos=gen["eos_token_id"],
                           n=int(ids_.shape[1]), sha256=sha(ids_), head=ids_[0, :60].tolist(),
                           tail=ids_[0, -12:].tolist(), n_img=int((ids_ == 92546).sum()))
    with open(os.path.join(GOLD, "prepare_chat_input.json"), "w") as f:
        json.dump(cases, f, indent=1, ensure_ascii=False)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
