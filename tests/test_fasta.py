"""FASTA ingest: the oracle's restatement of FASTAReader::GetSequences against the unmodified
reference (CPU) and against the committed golden answers; sibgpu_fasta_parse against both (GPU)."""
import json
import os

import numpy as np
import pytest

import fasta_cases
from oracle import ref, restate

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fasta_cases.json")


def restated(data):
    try:
        return [[d.decode("latin-1"), s.decode("latin-1")] for d, s in restate.fasta_get_sequences(data)]
    except restate.FastaError as e:
        return {"line": e.line, "what": e.what}


def reference(data, tmp_path, name="in.fasta"):
    p = os.path.join(str(tmp_path), name)
    with open(p, "wb") as f:
        f.write(data)
    try:
        return [[d.decode("latin-1"), s.decode("latin-1")] for d, s in ref.fasta_parse(p)[0]]
    except RuntimeError as e:
        msg = str(e)
        assert msg.startswith("parse error in " + p + " on line "), msg
        line, what = msg[len("parse error in " + p + " on line "):].split(": ", 1)
        return {"line": int(line), "what": what}


def gpu(ctx, data):
    import sibelia_b200 as sb
    try:
        return [[d.decode("latin-1"), s.tobytes().decode("latin-1")] for d, s in ctx.fasta_parse(data)]
    except sb.binding.FastaParseError as e:
        return {"line": e.line, "what": e.what}


def test_restatement_matches_golden():
    gold = json.load(open(GOLD))
    assert len(gold) == len(fasta_cases.HAND)
    for i, data in enumerate(fasta_cases.HAND):
        assert restated(data) == gold[i], "hand case %d" % i


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_restatement_matches_reference(tmp_path):
    for i, data in enumerate(fasta_cases.HAND):
        assert restated(data) == reference(data, tmp_path), "hand case %d" % i
    rng = np.random.default_rng(41)
    for i in range(300):
        data = fasta_cases.random_case(rng)
        assert restated(data) == reference(data, tmp_path), "random case %d" % i


@pytest.mark.gpu
def test_gpu_ingest_matches_oracle(ctx, tmp_path):
    gold = json.load(open(GOLD))
    for i, data in enumerate(fasta_cases.HAND):
        assert gpu(ctx, data) == gold[i], "hand case %d" % i
    rng = np.random.default_rng(42)
    for i in range(400):
        data = fasta_cases.random_case(rng)
        assert gpu(ctx, data) == restated(data), "random case %d" % i


@pytest.mark.gpu
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel")
def test_gpu_ingest_matches_reference_on_large_files(ctx, tmp_path):
    """wrapped at 80, unwrapped (one 30 MB line) and mixed-case multi-record files against the reference's reader"""
    from sibelia_b200 import synth
    rng = np.random.default_rng(43)
    g = synth.random_genome(30_000_000, 7).tobytes()
    wrapped = b">big one\n" + b"\n".join(g[o:o + 80] for o in range(0, len(g), 80)) + b"\n"
    unwrapped = b">big\n" + g + b"\n>second\n" + g[:1000].lower() + b"\n"
    for name, data in (("wrapped", wrapped), ("unwrapped", unwrapped)):
        assert gpu(ctx, data) == reference(data, tmp_path, name + ".fasta"), name
    for i in range(6):
        data = fasta_cases.random_case(rng, big=True)
        assert gpu(ctx, data) == reference(data, tmp_path, "r%d.fasta" % i), "big random case %d" % i
