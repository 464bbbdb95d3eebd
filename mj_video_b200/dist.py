"""Data-parallel preference scoring: preference pairs are sharded across ranks with chosen and rejected
rows of a pair on the SAME rank (the decision `~(s0 > s1)`, scripts/train/overall_train.py:427, stays local); weights
are replicated; the only collective is ONE all-gather of the per-row score table
[rows_local, num_objectives + num_aspects + 1] (= rewards | aspect_scores | score) over NCCL/NVLink (gloo on CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_pairs(n_pairs: int, rank: int, world: int) -> range:
    """Contiguous block of ceil(n/world) pairs per rank (the last ranks may get fewer)."""
    per = (n_pairs + world - 1) // world
    return range(min(rank * per, n_pairs), min((rank + 1) * per, n_pairs))


def pair_rows(pairs: range) -> list[tuple[int, int]]:
    """Row ids of a block of pairs: pair k -> rows (2k, 2k+1) = (chosen, rejected)."""
    return [(2 * k, 2 * k + 1) for k in pairs]


def preference_decisions(scores: torch.Tensor) -> torch.Tensor:
    """scores [2P] ordered (video_0, video_1) per pair -> 1 where video_1 is preferred, ties go to video_1
    (`prefer_predict = ~(score_video_0 > score_video_1)`, overall_train.py:427)."""
    s = scores.view(-1, 2)
    return (~(s[:, 0] > s[:, 1])).to(torch.int64)


class NcclScoreGather:
    """The C-ABI all-gather (`mjv_allgather_scores`, csrc/comm.cu): an NCCL communicator created by the library from a
    unique id that rank 0 generates and torch.distributed (any backend) distributes once. `gather` is enqueued on the
    current CUDA stream; it never synchronises the host."""

    def __init__(self, device, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            _lib.check(self.lib.mjv_comm_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        box = [ident]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = (C.c_ubyte * 128)(*box[0].tolist())
        self._comm = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mjv_comm_create(raw, self.rank, self.world, C.byref(self._comm)))

    def gather(self, local: torch.Tensor, n_pairs: int) -> torch.Tensor:
        """local [2*pairs_local, W] fp32 CUDA -> [2*n_pairs, W] in pair order (same contract as allgather_scores)."""
        per = (n_pairs + self.world - 1) // self.world
        width = local.shape[1]
        padded = torch.zeros((2 * per, width), dtype=torch.float32, device=self.device)
        padded[: local.shape[0]] = local
        out = torch.empty((self.world * 2 * per, width), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._lib.check(self.lib.mjv_allgather_scores(self._comm, padded.data_ptr(), 2 * per, width, out.data_ptr(),
                                                          self._lib.current_stream()))
        padded.record_stream(torch.cuda.current_stream(self.device))
        rows = [out[r * 2 * per: r * 2 * per + 2 * len(shard_pairs(n_pairs, r, self.world))] for r in range(self.world)]
        return torch.cat(rows)

    def close(self):
        if getattr(self, "_comm", None) is not None and self._comm.value:
            self.lib.mjv_comm_destroy(self._comm)
            self._comm = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def allgather_scores(local: torch.Tensor, n_pairs: int, group=None) -> torch.Tensor:
    """All-gather the per-row score table. `local` is [2*pairs_local, W] fp32; returns [2*n_pairs, W] in pair order.
    Equal-count collective: every rank pads its block to 2*ceil(n_pairs/world) rows."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[: 2 * n_pairs]
    per = (n_pairs + world - 1) // world
    width = local.shape[1]
    padded = torch.zeros((2 * per, width), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * 2 * per, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    rows = []
    for r in range(world):
        n_r = len(shard_pairs(n_pairs, r, world))
        rows.append(out[r * 2 * per: r * 2 * per + 2 * n_r])
    return torch.cat(rows)
