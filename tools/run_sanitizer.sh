#!/bin/bash
# compute-sanitizer (memcheck, initcheck, synccheck, racecheck) over one small forward that launches every kernel of the
# path: resize + tiling frame kernels (4 tiles + thumbnail), ViT, projector/splice, LLM prefill with ragged rows, head.
# Logs -> gpurun_out/sanitizer_<tool>.log (copied to profiles/r0N_compute_sanitizer_<tool>.log).
mkdir -p gpurun_out
cat > /tmp/tiny_fwd.py <<'PY'
import torch, sys
sys.path.insert(0, ".")
from mj_video_b200 import synth
from mj_video_b200.engine import Engine, EngineDims
from mj_video_b200.data_processor import load_video
d = synth.dims_tiny()
eng = Engine(EngineDims.from_synth(d), "cuda")
eng.load_state_dict(synth.make_state_dict(d, seed=2, include_lm_head=False))
pv = torch.cat([load_video(synth.make_video(i, n_frames=4, height=300, width=500), num_segments=2, max_num=4, device="cuda", dtype=torch.bfloat16)[0] for i in range(2)])
tiles = pv.shape[0] // 2
rows = [synth.make_input_ids(i, [tiles // 2] * 2, n_cap=9 + 30 * i) for i in range(2)]
ids, mask = synth.pad_batch(rows)
out = eng.forward(pv, ids, mask, synth.PAD, synth.IMG_CONTEXT)
torch.cuda.synchronize()
print("scores", out["scores"][:, -1].tolist(), "tiles", pv.shape[0])
# round 2: head_dim-128 causal attention with P in shared memory (GQA pairs and packed MHA blocks), the ViT tail paths on an
# even head count, and the fused resize on an odd width whose last row ends at the end of the allocation
import math
from mj_video_b200 import ops
from mj_video_b200.data_processor import preprocess_frames
for H, Hkv, pair in ((4, 2, False), (3, 3, True)):
    lens = [300, 77, 513]
    tot = sum(lens)
    q = torch.randn(tot, H * 128, device="cuda").bfloat16(); k = torch.randn(tot, Hkv * 128, device="cuda").bfloat16()
    starts, cols = ops.vt_layout(lens)
    vt = torch.randn(Hkv, 128, cols, device="cuda").bfloat16()
    cu = torch.tensor([0, 300, 377, 890], dtype=torch.int32, device="cuda")
    o = torch.empty(tot, H * 128, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, vt, o, head_dim=128, heads=H, kv_heads=Hkv, causal=True, scale=1 / math.sqrt(128), cu_seqlens=cu,
                  work=ops.causal_work_list(lens, pair_blocks=pair).cuda(), vt_starts=torch.tensor(starts, dtype=torch.int32, device="cuda"), vt_cols=cols)
qk = torch.randn(3 * 1025, 2 * 128, device="cuda").bfloat16()
vt = torch.randn(3 * 2, 64, 1032, device="cuda").bfloat16()
o = torch.empty(3 * 1025, 128, device="cuda", dtype=torch.bfloat16)
ops.attention(qk, qk, vt, o, head_dim=64, heads=2, kv_heads=2, causal=False, scale=0.125, q_col0=0, k_col0=128, nseq=3, seq_len=1025)
fr = torch.randint(0, 256, (3, 449, 447, 3), dtype=torch.uint8, device="cuda")
t, tpf = preprocess_frames(fr, 448, 2, True, torch.bfloat16)
torch.cuda.synchronize()
print("extra ok", float(o.float().abs().mean()), tpf)
PY
for tool in memcheck initcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 0 python /tmp/tiny_fwd.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_$tool.log
done
