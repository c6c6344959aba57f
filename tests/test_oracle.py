"""CPU: the oracle restatement (oracle/enum_restate.c + oracle/restate.py) against the committed golden fixtures
(generated from the unmodified reference by tests/golden/make_golden.py) and, where oracle/_ref exists, against
the reference itself."""
import glob
import os

import numpy as np
import pytest

import helpers
from oracle import ref, restate

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(path):
    z = np.load(path)
    lens = z["lens"]
    seq = z["seq"]
    chrs, at = [], 0
    for L in lens:
        chrs.append(seq[at:at + L].copy())
        at += L
    return chrs, int(z["k"]), z


GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLD, "enumerate_*.npz")))


def test_golden_present():
    assert len(GOLDEN_FILES) >= 30


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=[os.path.basename(p)[:-4] for p in GOLDEN_FILES])
def test_restatement_matches_golden(path):
    chrs, k, z = load_golden(path)
    count, pos, neg = restate.enumerate_bifurcations(chrs, k)
    helpers.assert_tables_equal((count, pos, neg), (int(z["count"]), z["pos"], z["neg"]), "golden")
    off, gidx, strand = restate.list_positions(count, pos, neg, [len(c) for c in chrs])
    assert np.array_equal(off, z["lp_off"])
    assert np.array_equal(gidx, z["lp_gidx"])
    assert np.array_equal(strand, z["lp_strand"])


def test_known_answer_vector():
    """k=3, ACGTACGGA / TTACGTC (ids are lexicographic ranks; w and revcomp(w) differ)."""
    chrs = [b"ACGTACGGA", b"TTACGTC"]
    count, pos, neg = restate.enumerate_bifurcations(chrs, 3)
    assert count == 10
    off, gidx, strand = restate.list_positions(count, pos, neg, [9, 7])
    want = {0: "+12 +4 +0 -15 -3", 1: "+13 +1 -14 -2 -6", 2: "-16", 3: "+6", 4: "+2 -13 -5", 5: "+14", 6: "-12",
            7: "+11 +3 -4", 8: "-8", 9: "+10", 10: ""}
    for i in range(11):
        got = " ".join("+-"[s] + str(g) for g, s in zip(gidx[off[i]:off[i + 1]], strand[off[i]:off[i + 1]]))
        assert got == want[i], (i, got)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_restatement_matches_reference_random():
    rng = np.random.default_rng(11)
    for _ in range(150):
        chrs, k = helpers.random_case(rng)
        r = ref.index(chrs, k)
        got = restate.enumerate_bifurcations(chrs, k)
        helpers.assert_tables_equal(got, (r["maxId"], r["pos"], r["neg"]), "ref")
        off, gidx, strand = restate.list_positions(got[0], got[1], got[2], [len(c) for c in chrs])
        assert np.array_equal(off, r["lp_off"]) and np.array_equal(gidx, r["lp_gidx"])
        assert np.array_equal(strand, r["lp_strand"])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_restatement_matches_reference_strains():
    st = helpers.strain_case(4, 30_000, seed=4242)
    for k in (25, 40, 1000):
        r = ref.index(st, k)
        helpers.assert_tables_equal(restate.enumerate_bifurcations(st, k), (r["maxId"], r["pos"], r["neg"]), "ref k=%d" % k)


HP = "/root/reference/examples/Sibelia/Helicobacter_pylori/Helicobacter_pylori.fasta"


@pytest.mark.skipif(not os.path.exists(HP), reason="reference example genome not present")
def test_restatement_on_reference_example_genome():
    """BASELINE.json configs[0] input: V=74 442, I=152 956 at k=25 + committed digest."""
    import hashlib
    from sibelia_b200 import synth
    hp = synth.read_fasta(HP)
    want = {}
    for line in open(os.path.join(GOLD, "hpylori_index_digests.txt")):
        if not line.startswith("#"):
            k, v, i, h = line.split()
            want[int(k)] = (int(v), int(i), h)
    count, pos, neg = restate.enumerate_bifurcations(hp, 25)
    assert (count, len(pos) + len(neg)) == (74442, 152956)
    assert hashlib.sha256(pos.tobytes() + neg.tobytes()).hexdigest() == want[25][2]


# ---- edge list (IndexedSequence + BlockFinder::ListEdges): numpy restatement vs golden fixtures and the reference
EDGE_FILES = sorted(glob.glob(os.path.join(GOLD, "edges_*.npz")))


def _edges_equal(got, want, what):
    assert len(got) == len(want), "%s: %d edges != %d" % (what, len(got), len(want))
    for f in want.dtype.names:
        assert np.array_equal(got[f], want[f]), "%s: field %s differs" % (what, f)


@pytest.mark.parametrize("path", EDGE_FILES, ids=[os.path.basename(p)[:-4] for p in EDGE_FILES])
def test_edge_restatement_matches_golden(path):
    z = np.load(path)
    n = int(z["n"])
    _edges_equal(restate.list_edges([z["seq_%d" % i] for i in range(n)], [z["op_%d" % i] for i in range(n)], int(z["k"])),
                 z["edges"], os.path.basename(path))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_edge_restatement_matches_reference():
    rng = np.random.default_rng(12)
    for it in range(60):
        chrs, k = helpers.random_case(rng, max_rec=5, max_len=80, kmax=10)
        chrs = [c.tobytes() for c in chrs]
        op = [rng.permutation(len(c)).astype(np.uint32) for c in chrs]
        _edges_equal(restate.list_edges(chrs, op, k), ref.list_edges(chrs, op, k)[0], "tiny %d k=%d" % (it, k))
    st = [c.tobytes() for c in helpers.strain_case(3, 20_000, seed=77)]
    op = [np.arange(len(c), dtype=np.uint32) for c in st]
    st2, op2, _, _ = ref.simplify(st, op, 30, 150, 4)
    for k in (30, 100):
        _edges_equal(restate.list_edges(st2, op2, k), ref.list_edges(st2, op2, k)[0], "simplified k=%d" % k)
