#!/bin/bash
# compute-sanitizer (memcheck, initcheck, synccheck, racecheck) over one small forward that launches every kernel of the
# path: resize + tiling frame kernels (4 tiles + thumbnail), ViT, projector/splice, LLM prefill with ragged rows, head.
# Logs -> gpurun_out/sanitizer_<tool>.log (copied to profiles/r01_compute_sanitizer_<tool>.log).
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
PY
for tool in memcheck initcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 0 python /tmp/tiny_fwd.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_$tool.log
done
