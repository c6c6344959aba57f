"""Generates tests/golden/simplify_large_digests.json from the UNMODIFIED reference (oracle/_ref): the SURVEY 8(d)
strain recipe at 4 x 12.5 Mb (p_sub = 0.002) through the four `-s loose`-like stages (25,150), (100,1000), (1000,5000),
(5000,15000) of BlockFinder::PerformGraphSimplifications, maxIterations = 4.  The states are far too large to commit,
so every stage is pinned by sha256 digests of rawSeq_ / originalPos_ per chromosome plus the bulge count; the input is
regenerated from the seeds (sibelia_b200.synth.strains) and pinned by its own digest.  Authoring container only
(about 15 minutes of CPU).

    python make_golden_simplify_large.py c3 [nstages]   BASELINE configs[2] itself: 4 x 125 Mb -> simplify_c3_digests.json
                                                        (hours of CPU and ~30 GB of RAM; the JSON is rewritten after
                                                        every stage, so a partial run still pins the stages it finished)"""
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

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simplify_large_digests.json")
CASE = {"n_strains": 4, "base_len": 12_500_000, "p_sub": 0.002, "base_seed": 1000, "strain_seed": 2000,
        "stages": [[25, 150], [100, 1000], [1000, 5000], [5000, 15000]], "iters": 4}


if len(sys.argv) > 1 and sys.argv[1] == "c3":
    OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simplify_c3_digests.json")
    CASE["base_len"] = 125_000_000
    if len(sys.argv) > 2:
        CASE["stages"] = CASE["stages"][:int(sys.argv[2])]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    chrs = [c.tobytes() for c in synth.strains(CASE["n_strains"], CASE["base_len"], base_seed=CASE["base_seed"],
                                               strain_seed=CASE["strain_seed"], p_sub=CASE["p_sub"])]
    op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
    out = dict(CASE)
    out["input"] = {"len": [len(c) for c in chrs], "seq_sha256": [hashlib.sha256(c).hexdigest() for c in chrs]}
    out["after"] = []
    for (k, D) in CASE["stages"]:
        t0 = time.perf_counter()
        chrs, op, bulges, sec = ref.simplify(chrs, op, k, D, CASE["iters"])
        out["after"].append({"k": k, "D": D, "bulges": int(bulges), "len": [len(c) for c in chrs],
                             "seq_sha256": [hashlib.sha256(c).hexdigest() for c in chrs],
                             "origpos_sha256": [sha(o) for o in op], "reference_seconds": sec})
        print("stage", (k, D), "bulges", bulges, "len", sum(len(c) for c in chrs), "%.1fs (wall %.1fs)" % (sec, time.perf_counter() - t0),
              flush=True)
        json.dump(out, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
