"""Generates tests/golden/trim_*.npz from the UNMODIFIED reference (oracle/_ref): BlockFinder::TrimBlocks
(src/synteny.cpp:31-122) on blocks of related sequences in mixed directions.  Authoring container only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
import helpers  # noqa: E402
from sibelia_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def block_case(seed, n, base_len, p_sub=0.02):
    """n diverged copies of one segment with random flanks, some reverse-complemented and read on the negative strand"""
    rng = np.random.default_rng(seed)
    st = helpers.strain_case(n, base_len, p_sub=p_sub, inv_len=max(50, base_len // 10), seed=seed)
    seqs, dirs = [], []
    for i, s in enumerate(st):
        left = synth.random_genome(int(rng.integers(0, 60)), seed * 100 + i)
        right = synth.random_genome(int(rng.integers(0, 60)), seed * 100 + 50 + i)
        s = np.concatenate([left, s, right])
        d = int(rng.integers(0, 2))
        seqs.append(synth.revcomp(s) if d else s)
        dirs.append(d)
    return seqs, dirs


CASES = {"trim_a": (51, 3, 2_000, 30, 100), "trim_b": (52, 5, 6_000, 30, 500), "trim_c": (53, 2, 900, 12, 50),
         "trim_d": (54, 4, 3_000, 30, 2_900)}


def main():
    for name, (seed, n, bl, k, min_size) in CASES.items():
        seqs, dirs = block_case(seed, n, bl)
        if name == "trim_b":
            seqs.append(synth.random_genome(700, 999))   # unrelated sequence: drop = true
            dirs.append(0)
        res, drop = ref.trim_blocks(seqs, dirs, k, min_size)
        out = {"n": np.int64(len(seqs)), "k": np.int64(k), "min_size": np.int64(min_size), "dirs": np.array(dirs, dtype=np.uint8),
               "result": np.array(res, dtype=np.int64).reshape(-1, 3), "drop": np.int64(drop)}
        for i, s in enumerate(seqs):
            out["seq_%d" % i] = s
        print(name, res, drop)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


if __name__ == "__main__":
    main()
