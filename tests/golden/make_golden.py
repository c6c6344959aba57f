"""Generates tests/golden/enumerate_*.npz from the UNMODIFIED reference (oracle/_ref/libsibelia_ref.so, built by
`make -C oracle ref` from /root/reference/src).  Run in the authoring container only; the fixtures are committed.

    python tests/golden/make_golden.py
"""
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


def dump(name, chrs, k):
    r = ref.index(chrs, k)
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        k=np.int64(k), lens=np.array([len(c) for c in chrs], dtype=np.int64),
                        seq=np.concatenate([np.asarray(c, dtype=np.uint8) for c in chrs] + [np.zeros(0, np.uint8)]),
                        count=np.int64(r["maxId"]), pos=r["pos"], neg=r["neg"],
                        lp_off=r["lp_off"], lp_gidx=r["lp_gidx"], lp_strand=r["lp_strand"])
    print(name, "k=%d V=%d I=%d" % (k, r["maxId"], len(r["pos"]) + len(r["neg"])))


def main():
    # SURVEY.md section 4 known-answer vector
    dump("enumerate_kat_k3", [np.frombuffer(b"ACGTACGGA", np.uint8), np.frombuffer(b"TTACGTC", np.uint8)], 3)
    rng = np.random.default_rng(7)
    for i in range(24):
        chrs, k = helpers.random_case(rng)
        dump("enumerate_tiny_%02d" % i, chrs, k)
    st = helpers.strain_case(4, 20_000, p_sub=0.01, inv_len=1500, seed=1000)
    for k in (15, 25, 28, 29, 32, 33, 64, 100, 500):
        dump("enumerate_strains20k_k%d" % k, st, k)
    # the reference's own example genome: only digests are committed (the FASTA itself is not ours to copy)
    hp_path = "/root/reference/examples/Sibelia/Helicobacter_pylori/Helicobacter_pylori.fasta"
    if os.path.exists(hp_path):
        import hashlib
        hp = synth.read_fasta(hp_path)
        lines = []
        for k in (25, 30):
            r = ref.index(hp, k)
            h = hashlib.sha256(r["pos"].tobytes() + r["neg"].tobytes()).hexdigest()
            lines.append("%d %d %d %s" % (k, r["maxId"], len(r["pos"]) + len(r["neg"]), h))
            print("hpylori", lines[-1])
        with open(os.path.join(OUT, "hpylori_index_digests.txt"), "w") as f:
            f.write("# k maxId instances sha256(pos||neg) -- reference IndexedSequence on examples/Sibelia/Helicobacter_pylori\n")
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
