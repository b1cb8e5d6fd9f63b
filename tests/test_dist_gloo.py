"""World-size-2 gloo run of the multi-GPU plumbing (sharding + the final packed gather) on CPU."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, K, ret):
    sys.path.insert(0, ROOT)
    import einx

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = einx.shard_range(total, rank, world)
        g = torch.Generator().manual_seed(1)
        all_m = torch.randint(-1, K, (total, K), generator=g)
        all_n = (all_m > -1).sum(1).to(torch.int32)
        packed = einx.pack_matches(all_m[lo:hi], all_n[lo:hi])
        per_rank = (total + world - 1) // world
        full = einx.gather_matches(packed, per_rank, total)
        raw = einx.gather_matches(packed, per_rank)  # no trimming: one collective, padding rows (-2) kept
        assert raw.shape == (world * per_rank, K + 1)
        ok = full.shape == (total, K + 1) and torch.equal(full[:, 1:].long(), all_m) and torch.equal(full[:, 0], all_n)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 8])
def test_shard_and_gather_world2(total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, total, 16, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
