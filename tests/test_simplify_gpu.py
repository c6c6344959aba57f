"""GPU parity: sibgpu_simplify (one BlockFinder::PerformGraphSimplifications stage through the C ABI) against the
committed golden fixtures (generated from the unmodified reference) and, where oracle/_ref travelled to the box,
against the reference itself on seeded inputs."""
import glob
import os

import numpy as np
import pytest

import helpers
from oracle import ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "simplify_*.npz")))


def assert_state_equal(got, want, what):
    gc, go, gb = got
    wc, wo, wb = want
    assert gb == wb, "%s bulges %d != %d" % (what, gb, wb)
    assert [len(c) for c in gc] == [len(c) for c in wc], "%s lengths differ" % what
    for i in range(len(wc)):
        assert bytes(gc[i]) == bytes(wc[i]), "%s sequence of chr %d differs" % (what, i)
        assert np.array_equal(go[i], wo[i]), "%s original positions of chr %d differ" % (what, i)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_golden_stages(ctx, path):
    z = np.load(path)
    n = int(z["n"])
    chrs = [z["in_seq_%d" % i].tobytes() for i in range(n)]
    op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
    for s, (k, D) in enumerate(z["stages"]):
        chrs, op, bulges = ctx.simplify(chrs, op, int(k), int(D), 4)
        want = ([z["seq_%d_%d" % (s, i)].tobytes() for i in range(n)], [z["op_%d_%d" % (s, i)] for i in range(n)],
                int(z["bulges_%d" % s]))
        assert_state_equal((chrs, op, bulges), want, "%s stage %d (k=%d, D=%d)" % (os.path.basename(path), s, k, D))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("seed,ns,bl,ps,stages", [
    (21, 4, 30_000, 0.01, [(25, 150)]),
    (22, 4, 30_000, 0.002, [(30, 150), (100, 1000), (1000, 5000), (5000, 15000)]),
    (23, 2, 60_000, 0.03, [(20, 80), (50, 500)]),
    (24, 6, 10_000, 0.05, [(15, 100)]),
    (25, 3, 20_000, 0.01, [(33, 200), (64, 700)]),
])
def test_against_reference(ctx, seed, ns, bl, ps, stages):
    chrs = [c.tobytes() for c in helpers.strain_case(ns, bl, p_sub=ps, inv_len=max(200, bl // 20), seed=seed)]
    op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
    rchrs, rop = chrs, op
    for (k, D) in stages:
        rchrs, rop, rb, _ = ref.simplify(rchrs, rop, k, D, 4)
        chrs, op, b = ctx.simplify(chrs, op, k, D, 4)
        assert_state_equal((chrs, op, b), (rchrs, rop, rb), "seed %d stage (%d,%d)" % (seed, k, D))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel")
def test_max_iterations_and_tiny_inputs(ctx):
    rng = np.random.default_rng(8)
    for it in range(40):
        chrs, k = helpers.random_case(rng, max_rec=4, max_len=120, kmax=6)
        chrs = [c.tobytes() for c in chrs]
        k = max(k, 2)
        op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
        iters = int(rng.integers(1, 5))
        D = int(rng.integers(k + 1, 40))
        want = ref.simplify(chrs, op, k, D, iters)[:3]
        got = ctx.simplify(chrs, op, k, D, iters)
        assert_state_equal(got, want, "tiny %d (k=%d D=%d iters=%d)" % (it, k, D, iters))


def _run_digest_fixture(ctx, name):
    import hashlib
    import json
    from sibelia_b200 import synth
    z = json.load(open(os.path.join(GOLD, name)))
    chrs = [c.tobytes() for c in synth.strains(z["n_strains"], z["base_len"], base_seed=z["base_seed"],
                                               strain_seed=z["strain_seed"], p_sub=z["p_sub"])]
    assert [hashlib.sha256(c).hexdigest() for c in chrs] == z["input"]["seq_sha256"], "generator drifted"
    op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
    for st in z["after"]:
        chrs, op, bulges = ctx.simplify(chrs, op, st["k"], st["D"], z["iters"])
        what = "stage (%d,%d)" % (st["k"], st["D"])
        assert bulges == st["bulges"], "%s bulges %d != %d" % (what, bulges, st["bulges"])
        assert [len(c) for c in chrs] == st["len"], what + " lengths differ"
        assert [hashlib.sha256(bytes(c)).hexdigest() for c in chrs] == st["seq_sha256"], what + " sequences differ"
        assert [hashlib.sha256(np.ascontiguousarray(o, dtype=np.uint32).tobytes()).hexdigest() for o in op] == st["origpos_sha256"], \
            what + " original positions differ"


def test_large_strain_set_against_reference_digests(ctx):
    """SURVEY 8(d) strain recipe at 4 x 12.5 Mb through the four loose stages: 143 410 collapses in stage 1 (Boost
    rehashes past 16 buckets, multi-group vertices, the overlap guard and the double accumulation en masse).  The
    reference's states are pinned by sha256 digests (tests/golden/make_golden_simplify_large.py, ~4 min of CPU)."""
    _run_digest_fixture(ctx, "simplify_large_digests.json")


def test_c3_500mb_pipeline_against_reference_digests(ctx):
    """BASELINE configs[2] itself: the 500 MB 4-strain set through the `-s loose` stages, 1 434 694 collapses in stage 1.
    The unmodified reference ran once in the authoring container (`make_golden_simplify_large.py c3`: 22 min for stage
    1, 11 min per later stage, ~30 GB of RAM); its states are pinned by sha256 digests per chromosome."""
    _run_digest_fixture(ctx, "simplify_c3_digests.json")
