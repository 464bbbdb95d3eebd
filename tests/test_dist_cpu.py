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
