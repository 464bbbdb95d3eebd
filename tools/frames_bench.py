"""Frame-path timing on one B200 (tools only): the fused resize kernel against the two-pass path (MJV_RESIZE_TWO_PASS=1) and
the identity path, in algorithmic GB/s = (frame bytes read once + tile bytes written once) / time. Inputs exceed L2.
Writes gpurun_out/frames_bench.json."""
from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200.data_processor import preprocess_frames  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    for (F, H, W, max_num) in ((256, 448, 448, 1), (256, 720, 1280, 1), (64, 720, 1280, 6), (128, 1080, 1920, 1), (256, 360, 640, 1),
                               (256, 480, 854, 1)):
        frames = torch.randint(0, 256, (F, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
        res = {}
        for mode in ("fused", "two_pass"):
            if mode == "two_pass":
                os.environ["MJV_RESIZE_TWO_PASS"] = "1"
            else:
                os.environ.pop("MJV_RESIZE_TWO_PASS", None)
            buf, tpf = preprocess_frames(frames, 448, max_num, True, torch.bfloat16)
            ms = timeit(lambda: preprocess_frames(frames, 448, max_num, True, torch.bfloat16, out=buf))
            passes = 1 if tpf == 1 else 2   # tiles + thumbnail each read the frame
            bytes_ = F * (passes * H * W * 3 + tpf * 3 * 448 * 448 * 2)
            res[mode] = {"ms": ms, "GBps": bytes_ / ms / 1e6, "tiles_per_frame": tpf}
            if H == 448 and W == 448:
                break
        os.environ.pop("MJV_RESIZE_TWO_PASS", None)
        out[f"{F}x{H}x{W}_m{max_num}"] = res
        print(f"{F}x{H}x{W} max_num={max_num}:", {k: (round(v['ms'], 3), round(v['GBps'])) for k, v in res.items()}, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/frames_bench.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
