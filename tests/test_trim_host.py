"""CPU: the host part of sibgpu_trim_blocks (per-vertex nearest instances on other sequences + the reduction over the
marks of each sequence) fed with the ORACLE's instance tables, against BlockFinder::TrimBlocks of the unmodified
reference (src/synteny.cpp:31-122) and the committed golden fixtures.  The GPU tests run the same function behind the
device enumeration (tests/test_trim_gpu.py)."""
import glob
import os
import sys

import numpy as np
import pytest

from oracle import ref, restate
from sibelia_b200 import binding

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "trim_*.npz")))


def finish(trim, dirs, k, min_size):
    res, drop = [], False
    for c, (found, start, end) in enumerate(trim):
        if not found:
            drop = True
            continue
        start, end = int(start), int(end)
        if abs(start - end) + k >= min_size:
            end = end + (k - 1) if dirs[c] == 0 else end - (k - 1)
            res.append((c, min(start, end), max(start, end) + 1 - min(start, end)))
    return res, drop


def host_trim(built, seqs, dirs, k):
    count, pos, neg = restate.enumerate_bifurcations(seqs, k)
    return binding.debug_trim_from_tables(count, pos, neg, [len(s) for s in seqs], dirs)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_golden_trim_host(built, path):
    z = np.load(path)
    n, k, ms = int(z["n"]), int(z["k"]), int(z["min_size"])
    seqs = [z["seq_%d" % i] for i in range(n)]
    got = finish(host_trim(built, seqs, z["dirs"], k), z["dirs"], k, ms)
    assert got == ([tuple(int(x) for x in r) for r in z["result"]], bool(z["drop"]))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_random_blocks_against_reference(built):
    sys.path.insert(0, GOLD)
    from make_golden_trim import block_case
    rng = np.random.default_rng(78)
    for it in range(40):
        n = int(rng.integers(2, 7))
        bl = int(rng.integers(100, 3000))
        k = int(rng.choice([8, 12, 30, 33, 64]))
        seqs, dirs = block_case(2000 + it, n, bl, p_sub=float(rng.choice([0.0, 0.01, 0.05])))
        if rng.random() < 0.3:
            seqs.append(seqs[0][: max(1, len(seqs[0]) // 3)].copy())
            dirs.append(int(rng.integers(0, 2)))
        ms = int(rng.integers(1, bl))
        got = finish(host_trim(built, seqs, dirs, k), dirs, k, ms)
        assert got == ref.trim_blocks(seqs, dirs, k, ms), "case %d (n=%d k=%d minSize=%d)" % (it, n, k, ms)
