"""GPU: the sharded enumeration through the real sibgpu_dist_* phases.  The box of `pytest -m gpu` has one GPU, so
the two (three, four) ranks are separate processes sharing cuda:0 and the small collectives go over gloo.  The peer
strategy (records read straight out of the other ranks' send buffers through CUDA IPC mappings) works the same way
between processes on one device as between GPUs over NVLink; the staged strategy (all-to-all through host memory
under gloo) and the overflow fallback from peer to staged are covered too.  bench.py runs the NCCL/NVLink variant
under torchrun on N GPUs."""
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
    if big == "poly":
        # 40 000 identical k-mers land in one hash partition: fixed-capacity segments overflow -> staged fallback
        from sibelia_b200 import synth
        rnd = synth.random_genome(90_000, seed)
        poly = np.full(40_000, ord("A"), dtype=np.uint8)
        return [np.concatenate([rnd[:50_000], poly, rnd[50_000:]]), synth.revcomp(rnd[20_000:60_000])]
    if big:
        return helpers.strain_case(4, 400_000, p_sub=0.004, inv_len=20_000, seed=seed)
    return helpers.strain_case(3, 9_000, p_sub=0.02, inv_len=700, seed=seed)


def _worker(rank, world, port, k, seed, big, part, q, env=None):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if part:
        os.environ["SIBGPU_PART_RECORDS"] = str(part)
    os.environ.update(env or {})
    import torch
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sibelia_b200 as sb
    from sibelia_b200 import distributed as D
    ctx = sb.Context(0)
    count, pos_part, neg_part = D.enumerate_sharded(D.GpuShard(ctx), _case(seed, big), k)
    count, pos, neg = D.gather_tables(count, pos_part, neg_part)
    if rank == 0:
        q.put((count, pos, neg, ctx.partition_fallbacks()))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def _run(world, k, seed, big, part, env=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, seed, big, part, q, env)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    helpers.assert_tables_equal(got[:3], restate.enumerate_bifurcations(_case(seed, big), k), "world=%d k=%d" % (world, k))
    return got[3]


@pytest.mark.parametrize("world,k,seed,big,part", [
    (2, 12, 1, False, 0), (2, 25, 2, False, 0), (3, 31, 3, False, 0), (2, 25, 4, True, 65536), (4, 29, 5, True, 0),
    (4, 9, 6, False, 0),
])
def test_sharded_enumeration_matches_oracle(built, world, k, seed, big, part):
    """peer strategy (default): segments read out of the other ranks' send buffers"""
    assert _run(world, k, seed, big, part) == 0


@pytest.mark.parametrize("world,k,seed,big,part", [(2, 25, 2, False, 0), (3, 30, 4, True, 65536)])
def test_sharded_staged_strategy(built, world, k, seed, big, part):
    _run(world, k, seed, big, part, {"SIBGPU_DIST_PEER": "0"})


def test_sharded_peer_overflow_falls_back_to_staged(built):
    fallbacks = _run(2, 25, 7, "poly", 4096, {"SIBGPU_PART_SLACK": "16"})
    assert fallbacks >= 1
