"""End-to-end GPU diagnostic: drop-in module forward vs the oracle port (fp32 and bf16), stage by stage."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import _lib, synth  # noqa: E402
from mj_video_b200.data_processor import load_video, preprocess_frames  # noqa: E402
from mj_video_b200.engine import Engine, EngineDims  # noqa: E402
from oracle import port  # noqa: E402

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
d = synth.dims_tiny() if which == "tiny" else synth.dims_2b()
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2


def cmp(name, got, ref, tol=None):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    msg = f"[{name}] max|d|={err.max().item():.4e} mean|d|={err.mean().item():.3e} ref|max|={ref.abs().max().item():.3e}"
    if tol is not None:
        msg += "  OK" if err.max().item() <= tol else "  MISMATCH"
    print(msg, flush=True)
    return err.max().item()


# ---------------- frame path (GPU kernel vs oracle restatement of PIL/torchvision)
for (H, W, mx) in [(448, 448, 1), (896, 1344, 6), (720, 1280, 6), (100, 77, 1), (500, 300, 4)]:
    v = synth.make_video(11, n_frames=4, height=H, width=W)
    ref, cnt = port.load_video_frames(v, num_segments=2, max_num=mx)
    got, cnt2 = load_video(v, num_segments=2, max_num=mx)
    ok = cnt == cnt2 and got.shape == ref.shape
    e = cmp(f"frames {H}x{W} max_num={mx} tiles {cnt2}", got, ref, 0.0) if ok else print("SHAPE MISMATCH", cnt, cnt2)
    got16, _ = load_video(v, num_segments=2, max_num=mx, device="cuda", dtype=torch.bfloat16)
    cmp("   bf16 variant vs RN(ref)", got16, ref.bfloat16(), 0.0)

# ---------------- model
t0 = time.time()
sd = synth.make_state_dict(d, seed=1, include_lm_head=False)
print(f"state dict: {time.time() - t0:.1f}s", flush=True)
eng = Engine(EngineDims.from_synth(d), dev)
t0 = time.time()
eng.load_state_dict(sd)
torch.cuda.synchronize()
print(f"engine load: {time.time() - t0:.1f}s", flush=True)

vids = [synth.make_video(100 + i) for i in range(B)]
pvs = [load_video(v, num_segments=8, max_num=1, device="cuda", dtype=torch.bfloat16)[0] for v in vids]
pv = torch.cat(pvs)
rows = [synth.make_input_ids(i, [1] * 8) for i in range(B)]
ids, mask = synth.pad_batch(rows)
print("batch", ids.shape, "tiles", pv.shape[0], flush=True)

orc32 = port.Oracle(d, sd, dtype=torch.float32, device=dev)
orc16 = port.Oracle(d, sd, dtype=torch.bfloat16, device=dev)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
r32 = orc32.forward(pv.float(), ids, mask, return_intermediates=True)
r16 = orc16.forward(pv, ids, mask, return_intermediates=True)

# stage: ViT
lib = _lib.load()
T = pv.shape[0]
S, Cv = d.vit_seq, d.vit_hidden
ws = eng.workspace(B, T, int(mask.sum()))
vit_hidden = torch.empty(T * S, Cv, dtype=torch.bfloat16, device=dev)
_lib.check(lib.mjv_vit_forward(eng._handle, pv.data_ptr(), _lib.MJV_DTYPE_BF16, T, vit_hidden.data_ptr(), ws.data_ptr(),
                               ws.numel(), _lib.current_stream()))
torch.cuda.synchronize()
cmp("vit_hidden vs oracle fp32", vit_hidden.view(T, S, Cv), r32["_vit_hidden"])
cmp("vit_hidden vs oracle bf16", vit_hidden.view(T, S, Cv), r16["_vit_hidden"])
cmp("   (oracle bf16 vs fp32)", r16["_vit_hidden"], r32["_vit_hidden"])

out = eng.forward(pv, ids, mask, 2, synth.IMG_CONTEXT)
torch.cuda.synchronize()
NO, NA = d.num_objectives, d.num_aspects
s, e = out["scores"], out["extras"]
mine = dict(rewards=s[:, :NO], aspect_scores=s[:, NO:NO + NA], score=s[:, NO + NA], hidden_state=out["hidden_state"],
            prompt_embedding=out["prompt_embedding"], criteria_gating_output=e[:, :NO],
            aspect_gating_output=e[:, NO:NO + NA], aspect_weights=e[:, NO + NA:NO + NA + NO],
            weighted_scores=e[:, NO + NA + NO])
for k in mine:
    cmp(f"{k:24s} vs fp32", mine[k], r32[k])
    cmp(f"{k:24s} vs bf16", mine[k], r16[k])
    cmp(f"{'':24s} (bf16 oracle vs fp32)", r16[k], r32[k])
print("gate argmax mine/fp32/bf16:", mine["aspect_gating_output"].argmax(-1).tolist(),
      r32["aspect_gating_output"].argmax(-1).tolist(), r16["aspect_gating_output"].argmax(-1).tolist())
print("launches:", lib.mjv_launch_count())

# timing
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    eng.forward(pv, ids, mask, 2, synth.IMG_CONTEXT)
e0.record()
n = 5
for _ in range(n):
    eng.forward(pv, ids, mask, 2, synth.IMG_CONTEXT)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"forward B={B}: {ms:.2f} ms/step  {B / ms * 1e3:.2f} rows/s")
