"""The C-ABI collective (mjv_comm_* / mjv_allgather_scores, csrc/comm.cu) on one GPU: an NCCL communicator of size 1
created by the library itself; the 2-GPU agreement with torch.distributed is asserted inside bench.py at N > 1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu


def test_nccl_score_gather_world1(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from mj_video_b200.dist import NcclScoreGather
    created = False
    if not dist.is_initialized():
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=0, world_size=1)
        created = True
    try:
        g = NcclScoreGather(torch.device("cuda:0"))
        local = torch.arange(10 * 34, dtype=torch.float32, device="cuda").view(10, 34) * 0.25
        out = g.gather(local, 5)
        torch.cuda.synchronize()
        assert torch.equal(out, local)
        out3 = g.gather(local[:6], 3)       # fewer rows than the previous call: new padded staging buffer
        assert torch.equal(out3, local[:6])
        g.close()
    finally:
        if created:
            dist.destroy_process_group()
