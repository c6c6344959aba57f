"""CPU, world_size 2 and 3 over gloo: the sharded-enumeration orchestration (sibelia_b200/distributed.py: partition
counts, all-to-all splits and layout, vertex-key all-gather, assembly of the per-rank tables) with the numpy test
double standing in for the GPU phases; the assembled result must equal the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from oracle import restate

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, k, seed, q, peer_dir=None, seg_cap=None):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sibelia_b200 import distributed as D
    from fake_shard import NumpyShard, NumpyPeerShard
    chrs = helpers.strain_case(3, 9_000, p_sub=0.02, inv_len=700, seed=seed)
    shard = NumpyPeerShard(peer_dir, seg_cap=seg_cap) if peer_dir else NumpyShard()
    count, pos_part, neg_part = D.enumerate_sharded(shard, chrs, k)
    count, pos, neg = D.gather_tables(count, pos_part, neg_part)
    if rank == 0:
        q.put((count, pos, neg, getattr(shard, "fallbacks", 0), getattr(shard, "peer", False)))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, k, seed, peer_dir=None, seg_cap=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, seed, q, peer_dir, seg_cap)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    chrs = helpers.strain_case(3, 9_000, p_sub=0.02, inv_len=700, seed=seed)
    helpers.assert_tables_equal(got[:3], restate.enumerate_bifurcations(chrs, k), "world=%d k=%d" % (world, k))
    return got[3], got[4]


@pytest.mark.parametrize("world,k,seed", [(2, 12, 1), (2, 25, 2), (3, 17, 3)])
def test_sharded_orchestration_matches_oracle(world, k, seed):
    """staged strategy: histogram, all-to-all of the records, key all-gather"""
    _run(world, k, seed)


@pytest.mark.parametrize("world,k,seed", [(2, 25, 2), (3, 14, 4)])
def test_peer_orchestration_matches_oracle(tmp_path, world, k, seed):
    """peer strategy: counts + capacity + overflow flag + 64-byte handle in one all-gather, owners read the peers' segments"""
    fallbacks, still_peer = _run(world, k, seed, str(tmp_path))
    assert fallbacks == 0 and still_peer


def test_peer_overflow_takes_the_staged_path_on_every_rank(tmp_path):
    fallbacks, _ = _run(2, 25, 5, str(tmp_path), seg_cap=16)       # far too small: every rank reports overflow
    assert fallbacks == 1


def _worker_fp(rank, world, port, k, seed, q, script):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sibelia_b200 import distributed as D
    from fake_shard import fake_fused_fp_shard
    chrs = helpers.strain_case(3, 9_000, p_sub=0.02, inv_len=700, seed=seed)
    shard = fake_fused_fp_shard(rank, world, **script)
    count, pos_part, neg_part = D.enumerate_sharded(shard, chrs, k)
    count, pos, neg = D.gather_tables(count, pos_part, neg_part)
    logs = [None] * world if rank == 0 else None
    dist.gather_object(shard.ctx.log, logs, dst=0)
    if rank == 0:
        q.put((count, pos, neg, logs, shard.last_strategy.split()[0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,k,script", [
    (2, 40, {}), (3, 64, {"collide_rank": 1}), (2, 100, {"regrow": True}), (3, 33, {"regrow": True, "collide_rank": 2}),
    (2, 25, {}), (3, 17, {"regrow": True}),              # k <= 28: sibgpu_fused_run, one call per step
])
def test_fused_orchestration(world, k, script):
    """fused strategy, k > 32 (sibgpu_fused_run_fp / _finish_fp) and k <= 28 (sibgpu_fused_run): collective allocation of the exported buffers, the class
    representatives min-reduced between the two halves, a verification failure on ONE rank repeats the step with other
    hash bases on ALL ranks, a key region that is too small is regrown collectively"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fp, args=(r, world, port, k, 7, q, script)) for r in range(world)]
    for p in procs:
        p.start()
    count, pos, neg, logs, strategy = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    chrs = helpers.strain_case(3, 9_000, p_sub=0.02, inv_len=700, seed=7)
    helpers.assert_tables_equal((count, pos, neg), restate.enumerate_bifurcations(chrs, k), "world=%d k=%d" % (world, k))
    assert strategy == "fused"
    want = ["plan", "release", "alloc", "import"]
    first = "run0" if k > 32 else "run"
    if script.get("regrow"):
        want += [first, "plan", "release", "alloc", "import"]
    want += [first, "finish0"] if k > 32 else [first]
    want += ["run1", "finish1"] if "collide_rank" in script else []
    assert all(log == want for log in logs), logs
