"""GPU: the sharded enumeration through the real sibgpu_fused_* / sibgpu_dist_* entry points.  The box of
`pytest -m gpu` has one GPU, so the two (three, four) ranks are separate processes sharing cuda:0 (time-sliced: a
kernel spinning on a peer's step counter is preempted) and the set-up collectives go over gloo.  The fused strategy
(k <= 28: step counters in the exported buffers, segments pulled out of the other ranks' buffers by TMA inside the
split kernel, keys pulled by a kernel; CUDA IPC mappings) works the same way between processes on one device as
between GPUs over NVLink; the peer strategy of round 1 (k <= 32), the staged strategy (all-to-all through host memory
under gloo) and the overflow fallbacks fused -> peer -> staged are covered too.  With two or more GPUs visible
(gpurun --gpus 2) the ranks get their own devices and NCCL.  bench.py runs the NCCL/NVLink variant under torchrun."""
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


def _worker(rank, world, port, k, seed, big, part, q, env=None, own_gpu=False):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if part:
        os.environ["SIBGPU_PART_RECORDS"] = str(part)
    os.environ.update(env or {})
    import torch
    dev = rank if own_gpu else 0
    torch.cuda.set_device(dev)
    if own_gpu:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import sibelia_b200 as sb
    from sibelia_b200 import distributed as D
    ctx = sb.Context(dev)
    shard = D.GpuShard(ctx)
    chrs = _case(seed, big)
    for _ in range(2):                                   # the second step reuses the exported buffers (step counters)
        count, pos_part, neg_part = D.enumerate_sharded(shard, chrs, k)
    count, pos, neg = D.gather_tables_device(count, pos_part, neg_part) if own_gpu else D.gather_tables(count, pos_part, neg_part)
    if rank == 0:
        q.put((count, pos, neg, ctx.partition_fallbacks(), shard.last_strategy.split()[0]))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def _run(world, k, seed, big, part, env=None, own_gpu=False, want=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, seed, big, part, q, env, own_gpu)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    helpers.assert_tables_equal(got[:3], restate.enumerate_bifurcations(_case(seed, big), k), "world=%d k=%d" % (world, k))
    if want:
        assert got[4] == want, "exchange strategy %s, expected %s" % (got[4], want)
    return got[3]


@pytest.mark.parametrize("world,k,seed,big,part", [
    (2, 12, 1, False, 0), (2, 25, 2, False, 0), (3, 31, 3, False, 0), (2, 25, 4, True, 65536), (4, 29, 5, True, 0),
    (4, 9, 6, False, 0),
])
def test_sharded_enumeration_matches_oracle(built, world, k, seed, big, part):
    """default strategies: fused for k <= 28, peer (segments read out of the other ranks' send buffers) above"""
    assert _run(world, k, seed, big, part, want="fused" if k <= 28 else "peer") == 0


@pytest.mark.parametrize("world,k,seed,big,part", [
    (2, 33, 11, False, 0), (3, 100, 12, False, 0), (2, 64, 13, True, 65536), (4, 500, 14, True, 0), (2, 5000, 15, True, 0),
    (8, 40, 18, False, 0),                               # 7 text tiles over 8 ranks: one rank owns partitions but no text
])
def test_sharded_fingerprint_k(built, world, k, seed, big, part):
    """k > 32 (the later -s loose stages) through the fused path: packed text replicated by peer pulls, fingerprint
    records exchanged like the exact ones, class representatives min-reduced, ids = string ranks on every rank"""
    assert _run(world, k, seed, big, part, want="fused") == 0


def test_sharded_fingerprint_k_falls_back_to_replicas(built):
    """no phased exchange exists for k > 32: without the fused path (here: switched off; in production: a bucket
    overflow or missing peer access) every rank indexes the whole input and keeps the rows of its own text range"""
    _run(2, 64, 16, False, 0, {"SIBGPU_DIST_FUSED": "0"}, want="replicated")
    _run(2, 40, 7, "poly", 4096, {"SIBGPU_PART_SLACK": "16"}, want="replicated")


@pytest.mark.parametrize("world,k,seed,big,part", [(2, 25, 2, False, 0), (4, 9, 6, False, 0), (2, 25, 4, True, 65536)])
def test_sharded_peer_strategy(built, world, k, seed, big, part):
    assert _run(world, k, seed, big, part, {"SIBGPU_DIST_FUSED": "0"}, want="peer") == 0


@pytest.mark.parametrize("world,k,seed,big,part", [(2, 25, 2, False, 0), (3, 30, 4, True, 65536)])
def test_sharded_staged_strategy(built, world, k, seed, big, part):
    _run(world, k, seed, big, part, {"SIBGPU_DIST_PEER": "0", "SIBGPU_DIST_FUSED": "0"}, want="staged")


def test_sharded_overflow_falls_back_to_staged(built):
    """40 000 identical k-mers: the fused step reports the overflow on every rank, so does the peer strategy"""
    fallbacks = _run(2, 25, 7, "poly", 4096, {"SIBGPU_PART_SLACK": "16"}, want="staged")
    assert fallbacks >= 1


def test_fused_key_regions_regrow(built):
    """vertex-key regions far too small: every rank learns it from the headers, the buffers are regrown collectively"""
    _run(2, 25, 8, True, 0, {"SIBGPU_CKEYS_INIT": "16"}, want="fused")


@pytest.mark.parametrize("world,k,seed,big,part", [(2, 25, 9, True, 0), (2, 31, 10, True, 0), (2, 100, 17, True, 0)])
def test_sharded_on_separate_gpus(built, world, k, seed, big, part):
    """one GPU per rank, NCCL for the set-up, NVLink for the pulls (needs gpurun --gpus 2)"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _run(world, k, seed, big, part, own_gpu=True, want="peer" if 28 < k <= 32 else "fused")
