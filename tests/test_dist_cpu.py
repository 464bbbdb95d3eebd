"""world_size-2 gloo test of the only collective on the path: the all-gather of per-row score tables."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mj_video_b200.dist import allgather_scores, pair_rows, preference_decisions, shard_pairs
    mine = shard_pairs(n_pairs, rank, world)
    rows = [r for pr in pair_rows(mine) for r in pr]
    # "scores" of row r are a deterministic function of r, so the gathered table can be checked exactly
    local = torch.tensor([[float(r), r * 0.5] + [0.0] * 31 + [((r * 7919) % 13) / 13.0] for r in rows],
                         dtype=torch.float32).reshape(len(rows), 34)
    full = allgather_scores(local, n_pairs)
    dec = preference_decisions(full[:, -1])
    if rank == 0:
        q.put((full[:, 0].tolist(), dec.tolist()))
    dist.destroy_process_group()


def test_allgather_scores_world2():
    n_pairs = 7  # odd: the last rank holds fewer pairs, exercising the padded equal-count collective
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    row_ids, dec = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert row_ids == [float(r) for r in range(2 * n_pairs)]
    exp = [int(not ((((2 * k) * 7919) % 13) > (((2 * k + 1) * 7919) % 13))) for k in range(n_pairs)]
    assert dec == exp


class _ToyPairs:
    """Pairs with tiny frames of different geometry per row and ragged prompts; scores are a pure function of the row."""
    score_width = 34

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, k):
        import numpy as np
        out = []
        for r in (2 * k, 2 * k + 1):
            h, w = 4 + (r % 3), 6 + (r % 2)
            frames = np.full((2, h, w, 3), (r * 37) % 251, dtype=np.uint8)
            out.append((frames, [1, 5 + r] + [7] * (r % 5)))
        return tuple(out)


def _toy_fns():
    """Stand-ins for the frame kernel and the engine: one 'tile' per frame whose value carries the row id."""
    def pre(frames_dev, input_size, max_num):
        return frames_dev.float().mean(dim=(1, 2, 3)).reshape(-1, 1), 1

    def fwd(pv, ids, mask):
        per_row = pv.view(ids.shape[0], -1).mean(1)
        n_tok = mask.sum(1).float()
        s = per_row * 0.01 + n_tok * 0.001 + ids[:, 1].float() * 1e-4
        t = torch.zeros((ids.shape[0], 34))
        t[:, 0], t[:, 1], t[:, -1] = ids[:, 1].float(), n_tok, s
        return t
    return fwd, pre


def _scorer_worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mj_video_b200.pipeline import PairScorer
    fwd, pre = _toy_fns()
    out = PairScorer(fwd, pre, pad_token_id=2, rows_per_step=4).score(_ToyPairs(n_pairs))
    q.put((rank, out["scores"].tolist(), out["prefer"].tolist(), out["rows_local"]))
    dist.destroy_process_group()


def test_pair_scorer_world2_matches_single_process():
    """BASELINE configs[2] driver: sharding, micro-batching, collation and the final all-gather give every rank the same
    table as one process scoring all pairs (odd pair count, ragged prompts, per-row frame geometry)."""
    from mj_video_b200.pipeline import PairScorer
    n_pairs = 7
    fwd, pre = _toy_fns()
    single = PairScorer(fwd, pre, pad_token_id=2, rows_per_step=6).score(_ToyPairs(n_pairs))
    assert single["scores"].shape == (2 * n_pairs, 34) and single["rows_local"] == 2 * n_pairs
    assert single["scores"][:, 0].tolist() == [5.0 + r for r in range(2 * n_pairs)]          # row order = pair order
    assert single["scores"][:, 1].tolist() == [2.0 + (r % 5) for r in range(2 * n_pairs)]    # padding never counted
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scorer_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows_local = {}
    for rank, table, prefer, n_local in got:
        assert torch.allclose(torch.tensor(table), single["scores"])
        assert prefer == single["prefer"].tolist()
        rows_local[rank] = n_local
    assert rows_local == {0: 8, 1: 6}   # chosen/rejected co-located: 4 + 3 pairs


def test_pair_scorer_propagates_dataset_errors():
    from mj_video_b200.pipeline import PairScorer
    import pytest

    class Bad(_ToyPairs):
        def __getitem__(self, k):
            if k == 2:
                raise RuntimeError("decode failed")
            return super().__getitem__(k)
    fwd, pre = _toy_fns()
    with pytest.raises(RuntimeError, match="decode failed"):
        PairScorer(fwd, pre, pad_token_id=2, rows_per_step=2).score(Bad(4))


class _NoWidthPairs(_ToyPairs):
    score_width = None  # a dataset that does not know the table width (only the engine does)


def _empty_rank_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mj_video_b200.pipeline import PairScorer
    _, pre = _toy_fns()

    def fwd7(pv, ids, mask):  # a head with 4 objectives + 2 aspects + 1 score = 7 columns
        t = torch.zeros((ids.shape[0], 7))
        t[:, 0], t[:, -1] = ids[:, 1].float(), ids[:, 1].float() * 0.5
        return t
    ds = _NoWidthPairs(1)
    del type(ds).score_width
    out = PairScorer(fwd7, pre, pad_token_id=2, rows_per_step=2, score_width=7).score(ds)
    q.put((rank, out["scores"].tolist(), out["rows_local"]))
    dist.destroy_process_group()


def test_pair_scorer_rank_without_pairs_uses_engine_width():
    """ceil-sharding leaves trailing ranks empty (1 pair on 2 ranks); the empty rank's all-gather contribution
    must have the engine's table width, not a hard-coded 34, or the collective mismatches and hangs."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_empty_rank_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = {rank: n for rank, _, n in got}
    assert rows == {0: 2, 1: 0}
    for _, table, _ in got:
        assert torch.tensor(table).shape == (2, 7)
        assert [r[0] for r in table] == [5.0, 6.0]
