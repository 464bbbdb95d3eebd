"""Preference-pair scoring driver: BASELINE configs[2] ("chosen/rejected scoring of N pairs sharded over the GPUs with a
score all-gather") and the batching / collation layer of SURVEY §8(f) N2.

What the reference does (scripts/eval/eval_genai_mjvideo.py:125-163): one sample at a time, two forwards per pair, a
device->host sync per forward, single process. Here:
  * pairs are sharded contiguously over the ranks, chosen and rejected rows of a pair on the same rank (dist.shard_pairs),
  * each rank walks its shard in micro-batches of `rows_per_step` rows; rows are PACKED by the engine (no padding work),
    so no length bucketing is needed to keep the GPU efficient — the right-padding convention of the reference's collator
    (dataset.py:464-469, notebook `pad_to_batch`) is only the interface,
  * a producer thread collates micro-batch i+1 into pinned host memory while the GPU works on micro-batch i; the uint8
    frames go host->device on a copy stream and are normalised / tiled by the frame kernels on the device,
  * per-row results (rewards | aspect_scores | score) stay on the device; ONE all-gather at the end assembles the table
    on every rank, and the decision rule `~(s0 > s1)` (overall_train.py:427) is applied to it.
The engine and the frame kernel are injected, so the host logic is testable without a GPU (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import queue
import threading
from dataclasses import dataclass

import numpy as np
import torch

from . import dist as mdist


@dataclass
class HostBatch:
    """One collated micro-batch in host memory."""
    pair_ids: list            # global pair index of each pair in the batch
    frames: list              # per row: uint8 tensor [F, H, W, 3] (pinned when CUDA is available)
    input_ids: torch.Tensor   # [rows, N] right-padded with pad_token_id
    attention_mask: torch.Tensor


def collate_pairs(dataset, pair_ids, pad_token_id: int, pin: bool) -> HostBatch:
    """Rows (chosen, rejected) of the given pairs -> right-padded ids/mask (reference collator convention) + frame list."""
    frames, rows = [], []
    for k in pair_ids:
        for frames_u8, ids in dataset[k]:
            t = torch.from_numpy(np.ascontiguousarray(frames_u8))
            if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
                raise ValueError(f"pair {k}: frames must be uint8 [F,H,W,3], got {t.dtype} {tuple(t.shape)}")
            frames.append(t.pin_memory() if pin else t)
            rows.append(list(ids))
    n = max(len(r) for r in rows)
    ids = torch.full((len(rows), n), pad_token_id, dtype=torch.long)
    mask = torch.zeros((len(rows), n), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = torch.tensor(r, dtype=torch.long)
        mask[i, : len(r)] = 1
    return HostBatch(list(pair_ids), frames, ids, mask)


class PairScorer:
    """score(dataset) -> {"scores": [2*n_pairs, W] fp32 on every rank, "prefer": [n_pairs] int64 (1 = video_1 preferred)}.

    dataset: len() = number of pairs; dataset[k] = ((frames0, ids0), (frames1, ids1)) with uint8 frames [F,H,W,3] (the
    sampled frames of the clip, i.e. what `load_video` feeds to dynamic_preprocess) and prompt token ids."""

    def __init__(self, forward_fn, preprocess_fn, pad_token_id: int, rows_per_step: int = 32, max_num: int = 1,
                 input_size: int = 448, device=None, group=None, prefetch: int = 2, gather_fn=None,
                 preprocess_batch_fn=None, score_width: int | None = None):
        """forward_fn(pixel_values, input_ids, attention_mask) -> [rows, W] fp32 score table on `device`;
        preprocess_fn(frames_u8_device, input_size, max_num) -> (tiles, tiles_per_frame)."""
        if rows_per_step < 2 or rows_per_step % 2:
            raise ValueError("rows_per_step must be even (chosen/rejected stay together)")
        self.forward_fn, self.preprocess_fn = forward_fn, preprocess_fn
        self.pad_token_id = pad_token_id
        self.rows_per_step, self.max_num, self.input_size = rows_per_step, max_num, input_size
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.group = group
        self.prefetch = max(1, prefetch)
        # columns of the score table (num_objectives + num_aspects + 1). A rank that owns no pair never sees a forward
        # result, and its all-gather contribution must still have the same width as everyone else's.
        self.score_width = score_width
        # (local table, n_pairs) -> full table; default: torch.distributed all-gather (gloo on CPU, NCCL on GPU)
        self.gather_fn = gather_fn or (lambda local, n: mdist.allgather_scores(local, n, group))
        # optional: (list of device uint8 clips) -> all tiles of the micro-batch in ONE buffer (no per-row concat copy)
        self.preprocess_batch_fn = preprocess_batch_fn

    @classmethod
    def for_engine(cls, engine, pad_token_id: int, img_context_token_id: int, **kw):
        """Wire to a mj_video_b200.engine.Engine and the CUDA frame kernels."""
        from .data_processor import preprocess_frames

        def fwd(pv, ids, mask):
            return engine.forward(pv, ids, mask, pad_token_id, img_context_token_id)["scores"]

        def pre(frames_dev, input_size, max_num):
            return preprocess_frames(frames_dev, input_size, max_num, True, torch.bfloat16)

        def pre_batch(clips, input_size, max_num):
            """Tile counts are known on the host (grid choice, data.py:66-99), so every clip's tiles are written straight
            into their slice of one micro-batch buffer."""
            from .data_processor.data import find_closest_aspect_ratio
            counts = []
            for f in clips:
                gi, gj = find_closest_aspect_ratio(f.shape[2], f.shape[1], 1, max_num, input_size)
                counts.append(f.shape[0] * (gi * gj + (1 if gi * gj > 1 else 0)))
            pv = torch.empty((sum(counts), 3, input_size, input_size), dtype=torch.bfloat16, device=engine.device)
            off = 0
            for f, n in zip(clips, counts):
                preprocess_frames(f, input_size, max_num, True, torch.bfloat16, out=pv[off: off + n])
                off += n
            return pv

        import torch.distributed as tdist
        if "gather_fn" not in kw and tdist.is_available() and tdist.is_initialized() and \
                tdist.get_world_size(kw.get("group")) > 1:
            kw["gather_fn"] = mdist.NcclScoreGather(engine.device, kw.get("group")).gather  # C ABI: mjv_allgather_scores
        kw.setdefault("preprocess_batch_fn", pre_batch)
        kw.setdefault("score_width", engine.dims.num_objectives + engine.dims.num_aspects + 1)
        return cls(fwd, pre, pad_token_id, device=engine.device, **kw)

    # ------------------------------------------------------------------------------------------------
    def _producer(self, dataset, batches, q, pin):
        try:
            for pair_ids in batches:
                q.put(collate_pairs(dataset, pair_ids, self.pad_token_id, pin))
            q.put(None)
        except BaseException as e:  # noqa: BLE001 - surfaced in the consumer
            q.put(e)

    def score(self, dataset) -> dict:
        world, rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(self.group)
            rank = torch.distributed.get_rank(self.group)
        n_pairs = len(dataset)
        mine = mdist.shard_pairs(n_pairs, rank, world)
        ppb = self.rows_per_step // 2
        batches = [list(mine[i: i + ppb]) for i in range(0, len(mine), ppb)]
        cuda = self.device.type == "cuda"
        q: queue.Queue = queue.Queue(maxsize=self.prefetch)
        th = threading.Thread(target=self._producer, args=(dataset, batches, q, cuda), daemon=True)
        th.start()

        copy_stream = torch.cuda.Stream(self.device) if cuda else None
        local = None
        row0 = 0
        staged = self._stage(q.get(), copy_stream)            # H2D of the first micro-batch
        while staged is not None:
            hb, frames_dev, ready = staged
            nxt = q.get()                                     # collated by the producer while the GPU was busy
            if cuda:
                torch.cuda.current_stream(self.device).wait_event(ready)
            if self.preprocess_batch_fn is not None:
                pv = self.preprocess_batch_fn(frames_dev, self.input_size, self.max_num)
            else:
                tiles = [self.preprocess_fn(f, self.input_size, self.max_num)[0] for f in frames_dev]
                pv = torch.cat(tiles) if len(tiles) > 1 else tiles[0]
            if cuda:
                for f in frames_dev:                          # clips were copied on the side stream
                    f.record_stream(torch.cuda.current_stream(self.device))
            scores = self.forward_fn(pv, hb.input_ids, hb.attention_mask).to(torch.float32)
            if local is None:
                local = torch.zeros((2 * len(mine), scores.shape[1]), dtype=torch.float32, device=scores.device)
            local[row0: row0 + scores.shape[0]] = scores
            row0 += scores.shape[0]
            staged = self._stage(nxt, copy_stream)            # overlaps with the forward just enqueued
        th.join()
        if local is None:                                      # this rank owns no pair
            width = self.score_width if self.score_width is not None else self._width_hint(dataset)
            local = torch.zeros((0, width), dtype=torch.float32, device=self.device)
        table = self.gather_fn(local, n_pairs)
        return dict(scores=table, prefer=mdist.preference_decisions(table[:, -1]), rows_local=row0)

    def _stage(self, hb, copy_stream):
        if hb is None:
            return None
        if isinstance(hb, BaseException):
            raise hb
        if copy_stream is None:
            return hb, [f.to(self.device) for f in hb.frames], None
        with torch.cuda.stream(copy_stream):
            frames_dev = [f.to(self.device, non_blocking=True) for f in hb.frames]
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return hb, frames_dev, ready

    @staticmethod
    def _width_hint(dataset) -> int:
        return int(getattr(dataset, "score_width", 34))


class SyntheticPairs:
    """Seeded synthetic preference pairs: pair k = videos (2k, 2k+1) under one caption.
    `pool` > 0 generates that many distinct clips once and cycles through them (row r -> clip r % pool), so that a
    throughput run measures the scoring pipeline rather than the numpy video generator (~0.3 s per clip)."""
    score_width = 34

    def __init__(self, n_pairs: int, n_frames: int = 8, height: int = 448, width: int = 448, tiles_per_frame: int = 1,
                 first_video: int = 0, pool: int = 0):
        self.n, self.n_frames, self.h, self.w, self.tpf, self.v0 = n_pairs, n_frames, height, width, tiles_per_frame, first_video
        self.pool = [self._clip(i) for i in range(pool)] if pool > 0 else None

    def _clip(self, r):
        from . import synth
        from .data_processor import get_index
        vid = synth.make_video(self.v0 + r, n_frames=2 * self.n_frames, height=self.h, width=self.w)
        return np.ascontiguousarray(vid[get_index(None, 8.0, len(vid) - 1, 0, self.n_frames)])

    def __len__(self):
        return self.n

    def __getitem__(self, k):
        from . import synth
        out = []
        for r in (2 * k, 2 * k + 1):
            clip = self.pool[r % len(self.pool)] if self.pool else self._clip(r)
            out.append((clip, synth.make_input_ids(k, [self.tpf] * self.n_frames)))
        return tuple(out)
