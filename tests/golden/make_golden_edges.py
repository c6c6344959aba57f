"""Generates tests/golden/edges_*.npz from the UNMODIFIED reference (oracle/_ref): the edge list
GenerateSyntenyBlocks starts from (IndexedSequence + BlockFinder::ListEdges, src/synteny.cpp:238-241) for states
before and after simplification stages.  Authoring container only."""
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
    # name: (n_strains, base_len, p_sub, inv_len, seed, stages run before listing, k of the listing)
    "edges_a": (4, 8_000, 0.01, 600, 31, [], 25),
    "edges_b": (3, 12_000, 0.02, 800, 32, [(30, 150), (100, 1000)], 100),
    "edges_c": (5, 5_000, 0.03, 400, 33, [(12, 60)], 40),
}


def main():
    for name, (ns, bl, ps, il, seed, stages, k) in CASES.items():
        chrs = [c.tobytes() for c in helpers.strain_case(ns, bl, p_sub=ps, inv_len=il, seed=seed)]
        chrs.append(b"ACGTAC")                           # shorter than any k used: no vertices, no edges
        op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
        for (sk, D) in stages:
            chrs, op, _, _ = ref.simplify(chrs, op, sk, D, 4)
        edges, sec = ref.list_edges(chrs, op, k)
        out = {"n": np.int64(len(chrs)), "k": np.int64(k), "edges": edges}
        for i, c in enumerate(chrs):
            out["seq_%d" % i] = np.frombuffer(c, dtype=np.uint8)
            out["op_%d" % i] = op[i]
        print(name, "k", k, "edges", len(edges), "%.3fs" % sec)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


if __name__ == "__main__":
    main()
