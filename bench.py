#!/usr/bin/env python
"""Headline benchmark: text-video pairs scored per second on MJ-VIDEO-2B (BASELINE.json), 8 frames of 448² per pair.

  python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path (one process per GPU via torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on the host CPU (oracle port)

A step = one pass of the hot path over one batch of 32 synthetic text-video rows (= 16 chosen/rejected preference
pairs) per GPU: uint8 frames -> frame kernel -> InternViT-300M -> projector/splice -> InternLM2 prefill -> MoE head
(-> all-gather of the score rows when N > 1). Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_OUT = sys.stdout
METRIC = "video pairs scored/sec @8x448^2 frames"
UNIT = "pairs/s"
WORKLOAD = "MJ-VIDEO-2B bf16, batch of 32 synthetic text-video pairs per GPU, 8 frames 448^2, max_num=1 (BASELINE configs[1])"
# other BASELINE configs, selectable for measurement (the default line is configs[1]):
WORKLOADS = {
    ("2b", "pairs8"): WORKLOAD,
    ("2b", "long16x7"): "MJ-VIDEO-2B bf16, long-visual-sequence rows: 16 frames 896x1344, dynamic_preprocess max_num=6 -> "
                        "7 tiles per frame, 112 tiles and ~28.9k tokens per row (BASELINE configs[3])",
    ("4b", "pairs8"): "InternViT-300M + Phi-3-mini-shaped LLM (d 3072, 32 layers, 32 heads x 96; reference Llama branch), "
                      "full 28-criterion MoE head, bf16, 32 synthetic pairs per GPU, 8 frames 448^2 (BASELINE configs[4])",
    ("4b", "long16x7"): "config-5 model on config-4 rows",
    ("2b", "pairs4096"): "MJ-VIDEO-2B bf16, chosen/rejected preference scoring of 4096 synthetic pairs (8192 rows, 8 frames 448^2) "
                         "sharded over the GPUs, micro-batches of 32 rows through pipeline.PairScorer, ONE all-gather "
                         "(BASELINE configs[2]; strong scaling)",
}


def model_dims(name):
    from mj_video_b200 import synth
    return synth.dims_2b() if name == "2b" else synth.dims_4b()


def step_flops(d, n_tiles, lengths):
    """Algorithmic FLOPs of one step: ViT + projector per tile, LLM linear + causal-exact attention per
    row; LM head, padding and the non-causal half of attention excluded."""
    C, S, P, Fv, Lv = d.vit_hidden, d.vit_seq, d.patches_per_tile, d.vit_mlp, d.vit_layers
    k_patch = 3 * d.patch_size ** 2
    f_tile = 2 * P * k_patch * C + Lv * (2 * S * C * 3 * C + 2 * S * C * C + 4 * S * C * Fv + 4 * S * S * C)
    D, F, L, H, KV, hd = d.llm_hidden, d.llm_inter, d.llm_layers, d.llm_heads, d.llm_kv_heads, d.llm_head_dim
    tpt = d.tokens_per_tile
    f_proj = 2 * tpt * 4 * C * D + 2 * tpt * D * D
    f_tok = L * 2 * (D * (H + 2 * KV) * hd + H * hd * D + 3 * D * F)
    return n_tiles * (f_tile + f_proj) + sum(n * f_tok + L * 4 * H * hd * n * (n + 1) / 2 for n in lengths)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], src="measured")
    except Exception:  # noqa: BLE001
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


def make_rows(n_rows: int, first_row: int, n_frames: int = 8, workload: str = "pairs8"):
    """Synthetic workload rows: row r = video id r, caption ids seeded by the PAIR (chosen/rejected share a caption).
    long16x7: 32-frame 896x1344 videos, 16 sampled frames, 7 tiles per frame."""
    import numpy as np
    from mj_video_b200 import synth
    from mj_video_b200.data_processor import get_index
    frames, rows = [], []
    long_ = workload == "long16x7"
    if long_:
        n_frames = 16
    for r in range(first_row, first_row + n_rows):
        vid = synth.make_video(r, n_frames=32, height=896, width=1344) if long_ else synth.make_video(r, n_frames=16)
        idx = get_index(None, 8.0, len(vid) - 1, 0, n_frames)
        frames.append(vid[idx])
        rows.append(synth.make_input_ids(r // 2, [7 if long_ else 1] * n_frames, n_cap=46 if long_ else None))
    return np.concatenate(frames), rows


def cpu_forward_seconds(d, sd, n_threads: int):
    """One fp32 forward of the oracle port on the host CPU for ONE row (BASELINE configs[0]); returns seconds."""
    import torch
    from mj_video_b200 import synth
    from oracle import port
    torch.set_num_threads(n_threads)
    frames, rows = make_rows(1, 0)
    pv, _ = port.load_video_frames(frames, num_segments=8, max_num=1)
    ids, mask = synth.pad_batch(rows)
    orc = port.Oracle(d, sd, dtype=torch.float32, device="cpu")
    t0 = time.perf_counter()
    orc.forward(pv, ids, mask)
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference algorithm (oracle port of the unmodified PyTorch forward; the reference itself is
    Python and /root/reference does not travel to the GPU box) on all host cores, one row per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from mj_video_b200 import synth
    from oracle import port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d = synth.dims_2b()
    sd = synth.make_state_dict(d, seed=1234, include_lm_head=True)
    frames, rows = make_rows(1, 0)
    ids, mask = synth.pad_batch(rows)
    orc = port.Oracle(d, sd, dtype=torch.float32, device="cpu")

    def step():
        pv, _ = port.load_video_frames(frames, num_segments=8, max_num=1)
        out = orc.forward(pv, ids, mask)
        return out["score"]

    budget_s = float(os.environ.get("MJV_REF_BUDGET_S", "300"))
    t0 = time.perf_counter()
    step()  # warm-up (also gives the per-step cost)
    t1 = time.perf_counter() - t0
    warm = 1
    steps = max(1, min(args.steps, int((budget_s - t1) // max(t1, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    el = time.perf_counter() - t0
    val = steps / el
    sample = (f"1 row (8 frames 448^2, N={ids.shape[1]} tokens) per step, fp32, eager attention, {cores} torch threads; "
              f"{steps} of the requested {args.steps} steps timed after {warm} warm-up (budget {budget_s:.0f}s)")
    line = dict(impl="reference", metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=el / steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=WORKLOAD, sample=sample),
                cpu_baseline=dict(value=val, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), file=_OUT, flush=True)


def run_pairs(args):
    """--workload pairs4096 (BASELINE configs[2]): the whole job through the public pipeline (reference loop:
    scripts/eval/eval_genai_mjvideo.py:125-163, decision rule overall_train.py:427) - host collation of ragged prompts,
    pinned H2D of uint8 clips on a copy stream, frame kernels, forward, and ONE all-gather of the score table. STRONG
    scaling: the total number of pairs is fixed, every rank scores a contiguous block with chosen/rejected co-located.
    The timed region is the wall clock of scorer.score() between barriers (max over ranks): there is no device-resident
    variant of a data-loading job, so `value` and `e2e` are the same measurement."""
    import hashlib
    import torch
    import torch.distributed as dist
    from mj_video_b200 import _lib, synth
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
    lib = _lib.load()
    d = model_dims(args.model)
    eng = Engine(EngineDims.from_synth(d), dev)
    eng.load_state_dict(synth.make_state_dict(d, seed=1234, include_lm_head=False))
    scorer = PairScorer.for_engine(eng, synth.PAD, synth.IMG_CONTEXT, rows_per_step=args.batch)
    n_pairs = args.pairs
    data = SyntheticPairs(n_pairs, pool=args.pool)          # `pool` distinct clips, generated before the timed region
    warm = SyntheticPairs(min(n_pairs, world * args.batch * max(1, args.warmup) // 2))
    warm.pool = data.pool

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scorer.score(warm)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    lib.mjv_reset_launch_count()
    t0 = time.perf_counter()
    out = scorer.score(data)
    barrier()
    el = time.perf_counter() - t0
    launches = int(lib.mjv_launch_count())
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([el], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = t.item()
    if rank == 0:
        table = out["scores"].cpu()
        margin = (table[:, -1].view(-1, 2)[:, 0] - table[:, -1].view(-1, 2)[:, 1]).abs()
        rows = 2 * n_pairs
        steps = -(-rows // (world * args.batch))
        h2d = rows * 8 * 448 * 448 * 3 // steps
        line = dict(metric=METRIC, value=rows / el, unit=UNIT, n_gpus=world, steps=steps, warmup=args.warmup,
                    ms_per_step=el / steps * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16",
                    data="synthetic",
                    config=dict(workload=WORKLOADS[(args.model, "pairs4096")], preference_pairs=n_pairs, rows=rows,
                                rows_per_micro_batch=args.batch, distinct_clips=args.pool,
                                l2="every micro-batch streams 0.15 GB of uint8 frames and >5 GB of activations (L2 is 126 MB)",
                                parallelism=f"dp{world}: contiguous pair blocks, chosen/rejected on one rank, 1 all-gather"),
                    e2e=dict(value=rows / el, unit=UNIT, h2d_bytes_per_step=int(h2d),
                             d2h_bytes_per_step=int(table.numel() * 4 // steps)),
                    preference_pairs_per_s=n_pairs / el, seconds=el, gpu_launches=launches, clocks=clocks,
                    decisions_sha256=hashlib.sha256(out["prefer"].cpu().numpy().tobytes()).hexdigest(),
                    scores_sha256=hashlib.sha256(table.numpy().tobytes()).hexdigest(),
                    prefer_video1_frac=float(out["prefer"].float().mean()), mean_abs_margin=float(margin.mean()),
                    roofline=None, cpu_baseline=None)
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_mjv(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mj_video_b200 import _lib, synth
    from mj_video_b200.data_processor import preprocess_frames
    from mj_video_b200.dist import NcclScoreGather, allgather_scores
    from mj_video_b200.engine import Engine, EngineDims

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    lib = _lib.load()
    pk = peaks()

    d = model_dims(args.model)
    sd = synth.make_state_dict(d, seed=1234, include_lm_head=False)
    eng = Engine(EngineDims.from_synth(d), dev)
    eng.load_state_dict(sd)
    del sd
    workload = WORKLOADS[(args.model, args.workload)]
    max_num = 6 if args.workload == "long16x7" else 1

    B = args.batch
    frames_np, rows = make_rows(B, rank * B, workload=args.workload)   # rank r scores pairs [r*B/2, (r+1)*B/2)
    ids, mask = synth.pad_batch(rows)
    frames_host = torch.from_numpy(frames_np).pin_memory()
    frames_dev = frames_host.to(dev)
    T = frames_dev.shape[0] * (7 if max_num == 6 else 1)
    px = torch.empty((T, 3, 448, 448), dtype=torch.bfloat16, device=dev)
    sb = eng.stage_batch(px, ids, mask, synth.PAD, synth.IMG_CONTEXT)
    n_pairs_total = world * B // 2
    lengths = sb.info["lengths"].tolist()

    # the data-path collective goes through the library's own C ABI (mjv_allgather_scores -> ncclAllGather over NVLink)
    gather = NcclScoreGather(dev) if world > 1 else None

    def step_local():   # this rank's kernels only: no collective (the rank-0-only breakdown step below must not hang the others)
        preprocess_frames(frames_dev, 448, max_num, True, torch.bfloat16, out=px)
        return eng.run(sb)["scores"]

    def step_resident():
        scores = step_local()
        if world > 1:
            return gather.gather(scores, n_pairs_total)
        return scores

    if world > 1:  # the C-ABI collective agrees with torch.distributed's on the same table
        chk = eng.run(sb)["scores"]
        if not torch.equal(gather.gather(chk, n_pairs_total), allgather_scores(chk, n_pairs_total)):
            raise RuntimeError("mjv_allgather_scores disagrees with torch.distributed.all_gather_into_tensor")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident throughput (`value`)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    lib.mjv_profile_enable(_lib.tag_mask("gemm_swiglu"))
    lib.mjv_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    launches = int(lib.mjv_launch_count())
    lib.mjv_profile_enable(0)
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    recs = _lib.profile_collect()
    value = world * B * args.steps / (ms * 1e-3)

    # dominant kernel: the fused gate/up SwiGLU GEMM (27% of the algorithmic FLOPs).
    # Rate = sum(FLOPs) / sum(time) over its full-size launches only: 1 launch in 24 is the last decoder layer's shortcut on
    # the 2*B gathered rows (M = 64, ~0.03 ms), and a plain average over launches would overstate the rate by ~4 %.
    M_tokens = sum(lengths)
    k_flops = 2.0 * M_tokens * (2 * d.llm_inter) * d.llm_hidden
    big = [(t_ms, w) for name, t_ms, w in recs if name == "gemm_swiglu" and w >= 0.5 * k_flops]
    sum_ms = sum(t for t, _ in big)
    achieved = sum(w for _, w in big) / (sum_ms * 1e-3) / 1e12 if big else None
    avg_ms = sum_ms / max(1, len(big))
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass
    roofline = dict(bound="tensor", kernel="gemm_bf16_tn_kernel<256,EPI_SWIGLU,4> (LLM w1|w3 + SiLU*mul)",
                    achieved=achieved, peak=pk["sustained"], unit="TFLOP/s",
                    frac=(achieved / pk["sustained"]) if achieved else None, traffic=traffic,
                    traffic_source="profiles/ncu_dominant_kernel.json (one ncu --set full capture of this kernel, per launch)",
                    peak_source=f"{pk['src']} sustained bf16 (kernel timed inside a long step); burst {pk['burst']}",
                    launches_timed=len(big), launches_excluded=len([1 for n, _, _ in recs if n == "gemm_swiglu"]) - len(big),
                    avg_launch_ms=avg_ms, flops_per_launch=k_flops)

    # whole-step algorithmic FLOPs -> end-to-end fraction of the tensor roofline
    F_step = step_flops(d, T, lengths)
    step_tflops = F_step * args.steps / (ms * 1e-3) / 1e12

    # ---------------------------------------------------------------- end to end through the public API (`e2e`)
    from mj_video_b200.data_processor import preprocess_frames as pf
    ids_cpu, mask_cpu = ids, mask
    h2d = frames_host.numel() + 4 * sum(v.size for v in eng.prepare_batch(ids, mask, T, synth.PAD, synth.IMG_CONTEXT)[0].values())
    d2h = B * (d.num_objectives + d.num_aspects + 1) * 4

    def step_e2e():
        fdev = frames_host.to(dev, non_blocking=True)              # H2D of this step's uint8 frames (pinned)
        pv, _ = pf(fdev, 448, max_num, True, torch.bfloat16)
        out = eng.forward(pv, ids_cpu, mask_cpu, synth.PAD, synth.IMG_CONTEXT)  # host indexing + meta H2D + kernels
        scores = out["scores"]
        if world > 1:
            scores = gather.gather(scores, n_pairs_total)
        return scores.cpu()                                       # D2H read of the result (synchronises)

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_val = world * B * args.steps / e2e_s

    # ---------------------------------------------------------------- per-kernel breakdown (one extra, untimed step)
    breakdown = None
    roofline_kernels = None
    if rank == 0:
        lib.mjv_profile_enable(_lib.tag_mask())
        step_local()
        torch.cuda.synchronize()
        lib.mjv_profile_enable(0)
        agg, by_tag = {}, {}
        for name, t_ms, work in _lib.profile_collect():
            tagname = name
            if name.startswith("gemm_"):
                name = f"{name}[{work / 1e12:.2f}TF]"   # one line per GEMM shape
            for key, table in ((name, agg), (tagname, by_tag)):
                a = table.setdefault(key, dict(launches=0, ms=0.0, work=0.0))
                a["launches"] += 1
                a["ms"] += t_ms
                a["work"] += work
        tot = sum(a["ms"] for a in agg.values())
        for a in agg.values():
            a["share"] = a["ms"] / tot
            a["rate_T_per_s"] = a["work"] / (a["ms"] * 1e-3) / 1e12 if a["ms"] > 0 else None
        breakdown = dict(total_kernel_ms=tot, kernels=agg)
        if args.breakdown:
            with open(args.breakdown, "w") as f:
                json.dump(breakdown, f, indent=1)
        # in-step rate of every kernel family against its own roofline (work = FLOPs for GEMM / attention, bytes otherwise)
        roofline_kernels = []
        for tagname, a in sorted(by_tag.items(), key=lambda kv: -kv[1]["ms"]):
            if a["ms"] <= 0 or a["work"] <= 0:
                continue
            tensor = tagname.startswith("gemm_") or tagname.startswith("attn_")
            rate = a["work"] / (a["ms"] * 1e-3) / 1e12
            roofline_kernels.append(dict(
                kernel=tagname, bound="tensor" if tensor else "hbm", launches=a["launches"], ms_per_step=round(a["ms"], 3),
                share_of_step=round(a["ms"] / tot, 4), achieved=round(rate if tensor else rate * 1e3, 1),
                unit="TFLOP/s" if tensor else "GB/s",
                frac=round(rate / pk["sustained"] if tensor else rate * 1e3 / pk["hbm"], 3)))

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        d2 = synth.dims_2b()
        sd_full = synth.make_state_dict(d2, seed=1234, include_lm_head=True)
        secs = cpu_forward_seconds(d2, sd_full, cores)
        cpu_baseline = dict(value=1.0 / secs, unit=UNIT, cores=cores, kind="port",
                            sample=f"1 row (BASELINE configs[0]: 8 frames 448^2, fp32, eager attention; the reference's LM-head "
                                   f"matmul, whose logits nobody reads, is NOT computed by the port: about -6 % FLOPs in "
                                   f"the CPU's favour), one forward = {secs:.1f} s on {cores} threads")

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic",
                    config=dict(workload=workload, rows_per_gpu=B, tiles_per_gpu=T, tokens_per_gpu=M_tokens,
                                l2="inputs and activations per step (0.15 GB uint8 frames, >5 GB activations) exceed the 126 MB L2",
                                parallelism=f"dp{world} over preference pairs (chosen/rejected on one rank) + 1 all-gather"),
                    e2e=dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
                    gpu_launches=launches, clocks=clocks, roofline=roofline, roofline_kernels=roofline_kernels,
                    cpu_baseline=cpu_baseline,
                    step_tflops=step_tflops, step_frac_of_sustained_peak=step_tflops / world / pk["sustained"])
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """Route everything libraries print on fd 1 (NCCL prints its version banner there) to stderr; the ONE JSON line of
    the contract is written to the original stdout through the returned file object."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mjv", choices=["mjv", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="rows (text-video pairs) per GPU per step (default 32; 2 for long16x7)")
    ap.add_argument("--model", default="2b", choices=["2b", "4b"], help="2b = MJ-VIDEO-2B (BASELINE configs[0-3]); 4b = configs[4] shapes")
    ap.add_argument("--workload", default="pairs8", choices=["pairs8", "long16x7", "pairs4096"],
                    help="pairs8 = configs[1]; long16x7 = configs[3]; pairs4096 = configs[2] (whole job, strong scaling)")
    ap.add_argument("--pairs", type=int, default=4096, help="pairs4096: number of preference pairs of the job")
    ap.add_argument("--pool", type=int, default=64, help="pairs4096: distinct synthetic clips (generated before timing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", default=None, help="write a per-kernel time breakdown JSON here (rank 0)")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 2 if args.workload == "long16x7" else 32
    args.warmup = max(args.warmup, 3) if args.impl == "mjv" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun when called directly with --gpus N
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), __file__] + sys.argv[1:]
            sys.exit(subprocess.call(cmd, stdout=_OUT))
        if args.workload == "pairs4096":
            run_pairs(args)
        else:
            run_mjv(args)


if __name__ == "__main__":
    main()
