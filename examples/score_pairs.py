#!/usr/bin/env python
"""BASELINE configs[2]: chosen/rejected preference scoring of N synthetic pairs, sharded over the GPUs of one box with ONE
all-gather of the score table at the end.

  python examples/score_pairs.py --pairs 256                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      examples/score_pairs.py --pairs 4096                                     # 8 GPUs

Mirrors the loop of the reference's driver (scripts/eval/eval_genai_mjvideo.py:125-163: two forwards per pair, compare
`.score`), through mj_video_b200.pipeline.PairScorer. Prints one JSON line on rank 0 (wall clock, end to end: synthetic
frame generation on the host, collation, H2D, kernels, all-gather)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--rows-per-step", type=int, default=32)
    ap.add_argument("--model", default="2b", choices=["2b", "4b", "tiny"])
    ap.add_argument("--pool", type=int, default=64, help="distinct synthetic clips generated up front (0: one per row, host-bound)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from mj_video_b200 import synth
    from mj_video_b200.engine import Engine, EngineDims
    from mj_video_b200.pipeline import PairScorer, SyntheticPairs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    d = {"2b": synth.dims_2b, "4b": synth.dims_4b, "tiny": synth.dims_tiny}[args.model]()
    eng = Engine(EngineDims.from_synth(d), dev)
    eng.load_state_dict(synth.make_state_dict(d, seed=1234, include_lm_head=False))
    scorer = PairScorer.for_engine(eng, synth.PAD, synth.IMG_CONTEXT, rows_per_step=args.rows_per_step)
    data = SyntheticPairs(args.pairs, pool=args.pool)
    warm = SyntheticPairs(min(args.pairs, world * args.rows_per_step // 2))
    warm.pool = data.pool
    scorer.score(warm)   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = scorer.score(data)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    el = time.perf_counter() - t0
    if rank == 0:
        s = out["scores"][:, -1].view(-1, 2)
        print(json.dumps(dict(pairs=args.pairs, n_gpus=world, seconds=el, preference_pairs_per_s=args.pairs / el,
                              rows_per_s=2 * args.pairs / el, prefer_video1_frac=float(out["prefer"].float().mean()),
                              mean_abs_margin=float((s[:, 0] - s[:, 1]).abs().mean()), model=d.name)), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
