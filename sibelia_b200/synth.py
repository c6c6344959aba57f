"""Synthetic genome generators (SURVEY.md section 8(d) recipe).  Pure numpy; used by tests and bench.py."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = np.arange(256)
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


def random_genome(length, seed):
    """C2-style contig: uniform random ACGT, numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, int(length), dtype=np.uint8)]


def revcomp(a):
    return _COMP[a[::-1]]


def mutate_strain(base, seed, p_sub=0.002, n_inv=4, inv_len=200_000):
    """One strain of the C3/C4 recipe: substitutions (p_sub), n_inv in-place reverse-complemented segments,
    1-base deletions and insertions (p_sub/20 each), in that order."""
    rng = np.random.default_rng(seed)
    s = base.copy()
    n = len(s)
    code = np.zeros(256, dtype=np.uint8)
    code[_ACGT] = np.arange(4, dtype=np.uint8)
    sub = rng.random(n) < p_sub
    shift = rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)
    s[sub] = _ACGT[(code[s[sub]] + shift) & 3]
    inv_len = min(inv_len, max(1, n // 8))
    for _ in range(n_inv):
        o = int(rng.integers(0, n - inv_len + 1))
        s[o:o + inv_len] = revcomp(s[o:o + inv_len])
    keep = rng.random(n) >= p_sub / 20
    s = s[keep]
    ins = np.flatnonzero(rng.random(len(s)) < p_sub / 20)
    if len(ins):
        s = np.insert(s, ins, _ACGT[rng.integers(0, 4, len(ins), dtype=np.uint8)])
    return s


def strains(n_strains, base_len, base_seed=1000, strain_seed=2000, p_sub=0.002, inv_len=200_000):
    """C3 (n_strains=4, base_len=125e6) / C4 (8 strains): list of uint8 arrays."""
    base = random_genome(base_len, base_seed)
    return [mutate_strain(base, strain_seed + s, p_sub=p_sub, inv_len=inv_len) for s in range(n_strains)]


def read_fasta(path):
    """Minimal FASTA reader with the reference's normalisation (upper-case; fasta.cpp:93-106)."""
    recs, cur = [], []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if cur:
                    recs.append(b"".join(cur))
                cur = []
            else:
                cur.append(line.strip().upper())
    if cur:
        recs.append(b"".join(cur))
    return [np.frombuffer(r, dtype=np.uint8).copy() for r in recs]
