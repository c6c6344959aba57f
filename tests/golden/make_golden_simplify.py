"""Generates tests/golden/simplify_*.npz from the UNMODIFIED reference (oracle/_ref): chained
BlockFinder::PerformGraphSimplifications stages on small synthetic strain sets.  Authoring container only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
import helpers  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (n_strains, base_len, p_sub, inv_len, seed, stages)
    "simplify_a": (4, 8_000, 0.01, 600, 11, [(25, 150), (100, 1000)]),
    "simplify_b": (3, 12_000, 0.02, 800, 12, [(30, 150), (100, 1000), (1000, 5000)]),
    "simplify_c": (5, 5_000, 0.03, 400, 13, [(12, 60), (40, 300)]),
}


def main():
    for name, (ns, bl, ps, il, seed, stages) in CASES.items():
        chrs = [c.tobytes() for c in helpers.strain_case(ns, bl, p_sub=ps, inv_len=il, seed=seed)]
        op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
        out = {"n": np.int64(len(chrs)), "stages": np.array(stages, dtype=np.int64)}
        for i, c in enumerate(chrs):
            out["in_seq_%d" % i] = np.frombuffer(c, dtype=np.uint8)
        for s, (k, D) in enumerate(stages):
            chrs, op, bulges, sec = ref.simplify(chrs, op, k, D, 4)
            out["bulges_%d" % s] = np.int64(bulges)
            for i, c in enumerate(chrs):
                out["seq_%d_%d" % (s, i)] = np.frombuffer(c, dtype=np.uint8)
                out["op_%d_%d" % (s, i)] = op[i]
            print(name, "stage", (k, D), "bulges", bulges, "len", sum(len(c) for c in chrs), "%.2fs" % sec)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


if __name__ == "__main__":
    main()
