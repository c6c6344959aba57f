"""GPU parity: sibgpu_trim_blocks (the index-and-search part of BlockFinder::TrimBlocks, src/synteny.cpp:31-122,
through the C ABI) against golden fixtures generated from the unmodified reference and against the reference itself."""
import glob
import os

import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "trim_*.npz")))


def finish(trim, dirs, k, min_size):
    """the caller's part of TrimBlocks (synteny.cpp:103-120) for blocks whose original position is 0"""
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


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_golden_trim(ctx, path):
    z = np.load(path)
    n, k, ms = int(z["n"]), int(z["k"]), int(z["min_size"])
    seqs = [z["seq_%d" % i] for i in range(n)]
    got = finish(ctx.trim_blocks(seqs, z["dirs"], k), z["dirs"], k, ms)
    assert got == ([tuple(int(x) for x in r) for r in z["result"]], bool(z["drop"]))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel")
def test_against_reference_random_blocks(ctx):
    import sys
    sys.path.insert(0, GOLD)
    from make_golden_trim import block_case
    rng = np.random.default_rng(77)
    for it in range(60):
        n = int(rng.integers(2, 7))
        bl = int(rng.integers(100, 4000))
        k = int(rng.choice([8, 12, 30, 33, 64]))
        seqs, dirs = block_case(1000 + it, n, bl, p_sub=float(rng.choice([0.0, 0.01, 0.05])))
        if rng.random() < 0.3:
            seqs.append(seqs[0][: max(1, len(seqs[0]) // 3)].copy())       # a partial extra copy
            dirs.append(int(rng.integers(0, 2)))
        ms = int(rng.integers(1, bl))
        want = ref.trim_blocks(seqs, dirs, k, ms)
        got = finish(ctx.trim_blocks(seqs, dirs, k), dirs, k, ms)
        assert got == want, "case %d (n=%d k=%d minSize=%d)" % (it, n, k, ms)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel")
def test_degenerate_blocks(ctx):
    A = lambda s: np.frombuffer(s, dtype=np.uint8)
    cases = [([A(b"ACGTACGTAC"), A(b"ACGTACGTAC")], [0, 0], 4, 1), ([A(b"ACGTACGTAC"), A(b"GTACGTACGT")], [0, 1], 4, 1),
             ([A(b"AC"), A(b"ACGTACGT")], [0, 0], 4, 1), ([A(b"AAAAAAAAAAAA"), A(b"AAAAAAAAAAAA")], [1, 0], 3, 2),
             ([A(b"ACGTTGCAAC")], [0], 3, 1)]
    for seqs, dirs, k, ms in cases:
        assert finish(ctx.trim_blocks(seqs, dirs, k), dirs, k, ms) == ref.trim_blocks(seqs, dirs, k, ms)
