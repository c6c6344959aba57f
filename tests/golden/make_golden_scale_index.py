"""Generates tests/golden/scale_index_digests.json from the UNMODIFIED reference (oracle/_ref): the index
(IndexedSequence ctor, in-RAM suffix-array path) of BASELINE configs[2]/[4]'s input (4 strains x 125 Mb) and of configs[3]'s
input (8 strains x 125 Mb = 10^9 bases) at k = 25.  The tables are far too large to commit: each case is pinned by
sha256(vertex count as uint64 || positive table || negative table) -- the `result_digest` of bench.py -- plus the counts.
Authoring container only: ~10 min and ~20 GB for the 500 Mb case, ~25 min and ~45 GB for the 10^9 case.

    python make_golden_scale_index.py [c3] [c4]"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from sibelia_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scale_index_digests.json")
CASES = {"c3": {"n_strains": 4, "base_len": 125_000_000, "k": 25}, "c4": {"n_strains": 8, "base_len": 125_000_000, "k": 25}}


def digest(count, pos, neg):
    h = hashlib.sha256()
    h.update(np.uint64(count).tobytes())
    h.update(np.ascontiguousarray(pos).tobytes())
    h.update(np.ascontiguousarray(neg).tobytes())
    return h.hexdigest()


def main():
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in (sys.argv[1:] or ["c3", "c4"]):
        case = dict(CASES[name])
        chrs = synth.strains(case["n_strains"], case["base_len"])
        t0 = time.perf_counter()
        r = ref.index(chrs, case["k"], dump=True)
        case.update({"bases": int(sum(len(c) for c in chrs)), "vertices": int(r["maxId"]), "instances_per_strand": int(len(r["pos"])),
                     "result_digest": digest(r["maxId"], r["pos"], r["neg"]), "reference_seconds": r["seconds"]})
        out[name] = case
        json.dump(out, open(OUT, "w"), indent=1)
        print(name, case, "wall %.0f s" % (time.perf_counter() - t0), flush=True)
        del r


if __name__ == "__main__":
    main()
