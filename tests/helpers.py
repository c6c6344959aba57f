"""Shared input generators / comparators for the parity tests."""
import numpy as np

from sibelia_b200 import synth

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_case(rng, max_rec=4, max_len=40, kmax=8):
    """Tiny adversarial inputs in the spirit of SURVEY.md section 3.3: few letters, records shorter than k, empty
    records, exact and reverse-complement copies."""
    nrec = int(rng.integers(1, max_rec + 1))
    k = int(rng.integers(1, kmax + 1))
    alpha = int(rng.integers(1, 5))
    chrs = []
    for i in range(nrec):
        L = int(rng.integers(0, max_len))
        a = ACGT[rng.integers(0, alpha, L)]
        if i and rng.random() < 0.3:
            a = synth.revcomp(chrs[-1]) if rng.random() < 0.5 else chrs[-1].copy()
        chrs.append(a)
    return chrs, k


def strain_case(n_strains=4, base_len=50_000, p_sub=0.01, inv_len=3000, seed=1000):
    return synth.strains(n_strains, base_len, base_seed=seed, strain_seed=seed + 1000, p_sub=p_sub, inv_len=inv_len)


def assert_tables_equal(got, want, what=""):
    gc, gp, gn = got
    wc, wp, wn = want
    assert gc == wc, "%s vertex count %d != %d" % (what, gc, wc)
    assert len(gp) == len(wp) and len(gn) == len(wn), "%s instance counts (%d,%d) != (%d,%d)" % (
        what, len(gp), len(gn), len(wp), len(wn))
    assert np.array_equal(gp, wp), "%s positive-strand table differs" % what
    assert np.array_equal(gn, wn), "%s negative-strand table differs" % what
