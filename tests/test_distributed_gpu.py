"""GPU: the sharded enumeration through the real sibgpu_dist_* phases.  The box of `pytest -m gpu` has one GPU, so
the two (three) ranks share cuda:0 and talk over gloo (records staged through host memory); the NCCL/NVLink variant
of the same code path is what bench.py runs under torchrun on N GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from oracle import restate

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(seed, big):
    if big:
        return helpers.strain_case(4, 400_000, p_sub=0.004, inv_len=20_000, seed=seed)
    return helpers.strain_case(3, 9_000, p_sub=0.02, inv_len=700, seed=seed)


def _worker(rank, world, port, k, seed, big, part, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if part:
        os.environ["SIBGPU_PART_RECORDS"] = str(part)
    import torch
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sibelia_b200 as sb
    from sibelia_b200 import distributed as D
    ctx = sb.Context(0)
    count, pos_part, neg_part = D.enumerate_sharded(D.GpuShard(ctx), _case(seed, big), k)
    count, pos, neg = D.gather_tables(count, pos_part, neg_part)
    if rank == 0:
        q.put((count, pos, neg))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,k,seed,big,part", [
    (2, 12, 1, False, 0), (2, 25, 2, False, 0), (3, 31, 3, False, 0), (2, 25, 4, True, 65536), (4, 29, 5, True, 0),
    (4, 9, 6, False, 0),
])
def test_sharded_enumeration_matches_oracle(built, world, k, seed, big, part):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, seed, big, part, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    helpers.assert_tables_equal(got, restate.enumerate_bifurcations(_case(seed, big), k), "world=%d k=%d" % (world, k))
