"""CPU: the host-side exact committer of sibgpu_simplify (sibelia_b200/csrc/simplifier.h: array-backed DNASequence /
BifurcationStorage / bulgeremoval.cpp restatement) against the committed golden stages and, where oracle/_ref exists,
the reference itself.  The GPU is replaced by "every vertex flagged" and the vertex tables by the oracle's, so this
checks exactly the host logic."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

import helpers
from oracle import ref, restate

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
LIB = os.path.join(HERE, "_build", "libhostcommit.so")


@pytest.fixture(scope="module")
def hc():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    src = os.path.join(HERE, "host_commit.cpp")
    hdr = os.path.join(HERE, "..", "sibelia_b200", "csrc", "simplifier.h")
    if not os.path.exists(LIB) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(LIB):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-fPIC", "-shared", "-o", LIB, src])
    return C.CDLL(LIB)


def host_simplify(lib, chrs, origpos, k, D, iters=4, dirty_mode=0):
    count, pos, neg = restate.enumerate_bifurcations(chrs, k)
    n = len(chrs)
    bufs = [np.frombuffer(bytes(c), dtype=np.uint8).copy() for c in chrs]
    ops = [np.ascontiguousarray(o, dtype=np.uint32) for o in origpos]
    seq = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in ops])
    lens = (C.c_uint64 * n)(*[len(b) for b in bufs])
    bulges, calls = C.c_uint64(), C.c_uint64()
    pos = np.ascontiguousarray(pos)
    neg = np.ascontiguousarray(neg)
    rc = lib.host_simplify(C.c_uint32(n), seq, op, lens, C.c_uint32(k), C.c_uint32(D), C.c_uint32(iters),
                           C.c_void_p(pos.ctypes.data), C.c_uint64(len(pos)), C.c_void_p(neg.ctypes.data),
                           C.c_uint64(len(neg)), C.c_uint32(count), C.byref(bulges), C.byref(calls), C.c_int(dirty_mode))
    assert rc == 0
    out_c, out_o = [], []
    lib.host_free.argtypes = [C.c_void_p]
    for i in range(n):
        m = lens[i]
        out_c.append(bytes((C.c_char * m).from_address(seq[i])) if m else b"")
        out_o.append(np.frombuffer(bytes((C.c_char * (4 * m)).from_address(op[i])), dtype=np.uint32).copy() if m
                     else np.zeros(0, np.uint32))
        lib.host_free(C.c_void_p(seq[i]))
        lib.host_free(C.c_void_p(op[i]))
    return out_c, out_o, bulges.value


def assert_state_equal(got, want, what):
    assert got[2] == want[2], "%s bulges %d != %d" % (what, got[2], want[2])
    for i in range(len(want[0])):
        assert bytes(got[0][i]) == bytes(want[0][i]), "%s sequence of chr %d differs" % (what, i)
        assert np.array_equal(got[1][i], want[1][i]), "%s original positions of chr %d differ" % (what, i)


FILES = sorted(glob.glob(os.path.join(GOLD, "simplify_*.npz")))


@pytest.mark.parametrize("dirty_mode", [0, 1, 2])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_golden_stages(hc, path, dirty_mode):
    z = np.load(path)
    n = int(z["n"])
    chrs = [z["in_seq_%d" % i].tobytes() for i in range(n)]
    op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
    for s, (k, D) in enumerate(z["stages"]):
        chrs, op, bulges = host_simplify(hc, chrs, op, int(k), int(D), 4, dirty_mode)
        want = ([z["seq_%d_%d" % (s, i)].tobytes() for i in range(n)], [z["op_%d_%d" % (s, i)] for i in range(n)],
                int(z["bulges_%d" % s]))
        assert_state_equal((chrs, op, bulges), want, "%s stage %d" % (os.path.basename(path), s))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_random_small_against_reference(hc):
    rng = np.random.default_rng(17)
    for it in range(60):
        chrs, k = helpers.random_case(rng, max_rec=4, max_len=150, kmax=7)
        chrs = [c.tobytes() for c in chrs]
        k = max(k, 2)
        op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
        iters = int(rng.integers(1, 5))
        D = int(rng.integers(k + 1, 50))
        want = ref.simplify(chrs, op, k, D, iters)[:3]
        for dm in (0, 1, 2):
            assert_state_equal(host_simplify(hc, chrs, op, k, D, iters, dm), want,
                               "case %d k=%d D=%d iters=%d dirty_mode=%d" % (it, k, D, iters, dm))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dirty_mode,seed,ps", [(0, 91, 0.02), (1, 91, 0.02), (1, 92, 0.05), (1, 93, 0.005), (2, 91, 0.02), (2, 92, 0.05),
                                                (2, 94, 0.01)])
def test_strains_against_reference(hc, dirty_mode, seed, ps):
    """dirty_mode = 1 is the sweep policy of sibgpu_simplify: later sweeps visit only vertices dirtied since their
    last visit (no snapshot, no renumbering); dirty_mode = 2 adds the parallel read-only screen of the dirty vertices at
    the start of those sweeps; both must give the reference's result exactly like visiting everything."""
    chrs = [c.tobytes() for c in helpers.strain_case(4, 40_000, p_sub=ps, inv_len=3000, seed=seed)]
    op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
    rc, ro = chrs, op
    for (k, D) in [(25, 150), (100, 1000)]:
        rc, ro, rb, _ = ref.simplify(rc, ro, k, D, 4)
        chrs, op, b = host_simplify(hc, chrs, op, k, D, 4, dirty_mode)
        assert_state_equal((chrs, op, b), (rc, ro, rb), "stage (%d,%d)" % (k, D))
