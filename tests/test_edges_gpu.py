"""GPU parity: sibgpu_list_edges (index + BlockFinder::ListEdges through the C ABI, no host-side index) against the
committed golden fixtures (generated from the unmodified reference by tests/golden/make_golden_edges.py) and, where
oracle/_ref travelled to the box, against the reference itself (src/synteny.cpp:238-241, src/serialization.cpp:56-86)."""
import glob
import os

import numpy as np
import pytest

import helpers
from oracle import ref, restate

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "edges_*.npz")))


def assert_edges_equal(got, want, what):
    assert len(got) == len(want), "%s: %d edges != %d" % (what, len(got), len(want))
    for f in want.dtype.names:
        assert np.array_equal(got[f], want[f]), "%s: field %s differs (first at %d)" % (
            what, f, int(np.flatnonzero(got[f] != want[f])[0]))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_golden_edges(ctx, path):
    z = np.load(path)
    n = int(z["n"])
    chrs = [z["seq_%d" % i] for i in range(n)]
    op = [z["op_%d" % i] for i in range(n)]
    assert_edges_equal(ctx.list_edges(chrs, op, int(z["k"])), z["edges"], os.path.basename(path))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("seed,ns,bl,stages,k", [
    (41, 4, 30_000, [], 25), (42, 3, 40_000, [(30, 150)], 30), (43, 4, 20_000, [(30, 150), (100, 1000)], 1000),
    (44, 2, 50_000, [], 33), (45, 5, 8_000, [(20, 100)], 5000),
])
def test_against_reference(ctx, seed, ns, bl, stages, k):
    chrs = [c.tobytes() for c in helpers.strain_case(ns, bl, p_sub=0.01, inv_len=max(200, bl // 20), seed=seed)]
    op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
    for (sk, D) in stages:
        chrs, op, _, _ = ref.simplify(chrs, op, sk, D, 4)
    want, _ = ref.list_edges(chrs, op, k)
    assert_edges_equal(ctx.list_edges(chrs, op, k), want, "seed %d k=%d" % (seed, k))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel")
def test_tiny_and_degenerate(ctx):
    rng = np.random.default_rng(9)
    for it in range(150):
        chrs, k = helpers.random_case(rng, max_rec=5, max_len=80, kmax=10)
        chrs = [c.tobytes() for c in chrs]
        op = [rng.permutation(len(c)).astype(np.uint32) for c in chrs]    # arbitrary original positions
        want, _ = ref.list_edges(chrs, op, k)
        assert_edges_equal(ctx.list_edges(chrs, op, k), want, "tiny %d k=%d" % (it, k))
    assert len(ctx.list_edges([], [], 5)) == 0
    assert len(ctx.list_edges([b"ACG"], None, 5)) == 0


def test_against_restatement(ctx):
    """the numpy restatement of ListEdges (oracle/restate.py, pinned against the reference in tests/test_oracle.py) needs
    no oracle/_ref on the box"""
    rng = np.random.default_rng(10)
    st = helpers.strain_case(4, 50_000, p_sub=0.01, inv_len=3_000, seed=48)
    op = [rng.permutation(len(c)).astype(np.uint32) for c in st]
    for k in (12, 25, 32, 40):
        assert_edges_equal(ctx.list_edges(st, op, k), restate.list_edges(st, op, k), "restatement k=%d" % k)


def test_identity_origpos_when_null(ctx):
    st = helpers.strain_case(3, 20_000, seed=46)
    op = [np.arange(len(c), dtype=np.uint32) for c in st]
    assert_edges_equal(ctx.list_edges(st, None, 25), ctx.list_edges(st, op, 25), "NULL origpos")


def test_edges_are_consecutive_instance_pairs(ctx):
    """size-independent property: per strand, edges = pairs of neighbouring rows of the enumeration table with equal chr"""
    st = helpers.strain_case(4, 200_000, p_sub=0.005, inv_len=10_000, seed=47)
    k = 25
    count, pos, neg = ctx.enumerate(st, k)
    e = ctx.list_edges(st, None, k)
    for strand, tab in ((0, pos), (1, neg)):
        same = tab["chr"][:-1] == tab["chr"][1:]
        es = e[e["direction"] == strand]
        assert len(es) == int(same.sum())
        assert np.array_equal(es["start_vertex"], tab["bifId"][:-1][same])
        assert np.array_equal(es["end_vertex"], tab["bifId"][1:][same])
        assert np.array_equal(es["actual_length"], (tab["pos"][1:] - tab["pos"][:-1])[same] + k)
