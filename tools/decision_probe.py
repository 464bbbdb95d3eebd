"""Margins and drift of preference decisions at full size (MJ-VIDEO-2B): CUDA engine vs the oracle port in bf16 and fp32
on the GPU (the checker), for two pair generators: structured-vs-structured and structured-vs-noise videos.
Writes gpurun_out/decision_probe.json (numbers behind tests/test_decisions_gpu.py)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import synth  # noqa: E402
from mj_video_b200.data_processor import load_video  # noqa: E402
from mj_video_b200.dist import preference_decisions  # noqa: E402
from mj_video_b200.engine import Engine, EngineDims  # noqa: E402
from oracle import port  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
P32 = int(sys.argv[2]) if len(sys.argv) > 2 else 32
d = synth.dims_2b()
sd = synth.make_state_dict(d, seed=1234, include_lm_head=False)
eng = Engine(EngineDims.from_synth(d), dev)
eng.load_state_dict(sd)
NO, NA = d.num_objectives, d.num_aspects


def clips(kind, k):
    a = synth.make_video(9000 + 2 * k)
    if kind == "contrast":   # same generator, but the two clips of a pair sit at opposite ends of the brightness range
        b = synth.make_video(9000 + 2 * k + 1)
        a = (a.astype(np.float32) * 0.45).astype(np.uint8)
        b = (b.astype(np.float32) * 0.45 + 140).astype(np.uint8)
        return (a, b) if k % 2 == 0 else (b, a)
    b = synth.make_video(9000 + 2 * k + 1, structured=(kind == "struct"))
    return a, b


def run(kind, n_pairs, n32):
    out = {}
    mine, o16, o32 = [], [], []
    orc16 = port.Oracle(d, sd, dtype=torch.bfloat16, device=dev)
    t0 = time.time()
    for p0 in range(0, n_pairs, 8):
        pvs, rows = [], []
        for k in range(p0, min(p0 + 8, n_pairs)):
            for v in clips(kind, k):
                pvs.append(load_video(v, num_segments=8, max_num=1, device="cuda", dtype=torch.bfloat16)[0])
                rows.append(synth.make_input_ids(k, [1] * 8))
        pv = torch.cat(pvs)
        ids, mask = synth.pad_batch(rows)
        mine.append(eng.forward(pv, ids, mask, synth.PAD, synth.IMG_CONTEXT)["scores"].float().cpu())
        r = orc16.forward(pv, ids, mask)
        o16.append(torch.cat([r["rewards"].float(), r["aspect_scores"].float(), r["score"].float()[:, None]], 1).cpu())
    del orc16
    torch.cuda.empty_cache()
    t16 = time.time() - t0
    orc32 = port.Oracle(d, sd, dtype=torch.float32, device=dev)
    t0 = time.time()
    for p0 in range(0, n32, 4):
        pvs, rows = [], []
        for k in range(p0, min(p0 + 4, n32)):
            for v in clips(kind, k):
                pvs.append(load_video(v, num_segments=8, max_num=1, device="cuda", dtype=torch.bfloat16)[0])
                rows.append(synth.make_input_ids(k, [1] * 8))
        pv = torch.cat(pvs)
        ids, mask = synth.pad_batch(rows)
        r = orc32.forward(pv.float(), ids, mask)
        o32.append(torch.cat([r["rewards"].float(), r["aspect_scores"].float(), r["score"].float()[:, None]], 1).cpu())
    del orc32
    torch.cuda.empty_cache()
    t32 = time.time() - t0
    mine, o16 = torch.cat(mine), torch.cat(o16)
    o32 = torch.cat(o32) if o32 else None
    sm, s16 = mine[:, -1], o16[:, -1]
    m16 = (s16.view(-1, 2)[:, 0] - s16.view(-1, 2)[:, 1])
    out["pairs"] = n_pairs
    out["seconds_engine_plus_bf16_oracle"] = t16
    out["seconds_fp32_oracle"] = t32
    out["margin_ref_bf16_abs"] = dict(mean=float(m16.abs().mean()), median=float(m16.abs().median()), min=float(m16.abs().min()),
                                     hist=np.histogram(m16.abs().numpy(), bins=[0, 1e-3, 2e-3, 5e-3, 1e-2, 2e-2, 5e-2, 0.1, 0.2, 0.5, 10])[0].tolist())
    out["drift_new_vs_ref_bf16"] = dict(rewards=float((mine[:, :NO] - o16[:, :NO]).abs().max()),
                                        aspect_scores=float((mine[:, NO:NO + NA] - o16[:, NO:NO + NA]).abs().max()),
                                        score_max=float((sm - s16).abs().max()), score_mean=float((sm - s16).abs().mean()))
    out["agree_new_vs_ref_bf16"] = float((preference_decisions(sm) == preference_decisions(s16)).float().mean())
    if o32 is not None:
        n = o32.shape[0]
        s32 = o32[:, -1]
        out["pairs_fp32"] = n // 2
        out["drift_new_vs_fp32"] = dict(rewards=float((mine[:n, :NO] - o32[:, :NO]).abs().max()),
                                        aspect_scores=float((mine[:n, NO:NO + NA] - o32[:, NO:NO + NA]).abs().max()),
                                        score_max=float((sm[:n] - s32).abs().max()), score_mean=float((sm[:n] - s32).abs().mean()))
        out["drift_ref_bf16_vs_fp32"] = dict(rewards=float((o16[:n, :NO] - o32[:, :NO]).abs().max()),
                                             aspect_scores=float((o16[:n, NO:NO + NA] - o32[:, NO:NO + NA]).abs().max()),
                                             score_max=float((s16[:n] - s32).abs().max()), score_mean=float((s16[:n] - s32).abs().mean()))
        out["agree_new_vs_fp32"] = float((preference_decisions(sm[:n]) == preference_decisions(s32)).float().mean())
        out["agree_ref_bf16_vs_fp32"] = float((preference_decisions(s16[:n]) == preference_decisions(s32)).float().mean())
    return out


res = {}
for kind in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("struct", "noise")):
    res[kind] = run(kind, P, P32)
    print(kind, json.dumps(res[kind]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/decision_probe.json", "w"), indent=1)
