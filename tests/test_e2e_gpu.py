"""GPU end-to-end parity (BASELINE.json configs[0]): the reference CLI with its two hot-path seams served by
libsibgpu.so (oracle/_ref/Sibelia_gpu = the reference's own translation units minus vertexenumeration.cpp /
blockfinder.cpp / bulgeremoval.cpp / libdivsufsort, plus sibelia_b200/csrc/facade/*_gpu.cpp) must write byte-identical
blocks_coords.txt / genomes_permutations.txt / coverage_report.txt to the unmodified reference CLI (oracle/_ref/Sibelia).
Both binaries are built in the authoring container by `make -C oracle sibelia sibelia_gpu data` and travel to the box."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import helpers
from sibelia_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "Sibelia")
GPU_BIN = os.path.join(ROOT, "oracle", "_ref", "Sibelia_gpu")
DATA = os.path.join(ROOT, "oracle", "_ref", "data")
FILES = ["blocks_coords.txt", "genomes_permutations.txt", "coverage_report.txt"]
have = os.path.exists(REF_BIN) and os.path.exists(GPU_BIN)


def write_fasta(path, chrs, names=None):
    with open(path, "wb") as f:
        for i, c in enumerate(chrs):
            f.write((">%s\n" % (names[i] if names else "strain%d" % i)).encode())
            b = bytes(c)
            for o in range(0, len(b), 80):
                f.write(b[o:o + 80] + b"\n")


def run(binary, args, outdir):
    os.makedirs(outdir, exist_ok=True)
    subprocess.run([binary] + args + ["-o", outdir], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                   timeout=1500)
    return {f: open(os.path.join(outdir, f), "rb").read() for f in FILES}


def compare(args, tmp_path):
    want = run(REF_BIN, args, str(tmp_path / "ref"))
    got = run(GPU_BIN, args, str(tmp_path / "gpu"))
    for f in FILES:
        assert got[f] == want[f], "%s differs (%d vs %d bytes)" % (f, len(got[f]), len(want[f]))
    return got


@pytest.mark.skipif(not have, reason="oracle/_ref/Sibelia{,_gpu} did not travel")
@pytest.mark.parametrize("params", ["loose", "fine"])
def test_synthetic_strains(tmp_path, params):
    chrs = helpers.strain_case(3, 150_000, p_sub=0.01, inv_len=20_000, seed=31)
    fa = str(tmp_path / "in.fasta")
    write_fasta(fa, chrs)
    got = compare(["-s", params, "-m", "2000", fa], tmp_path)
    assert got["blocks_coords.txt"].count(b"Block #") >= 2


@pytest.mark.skipif(not have, reason="oracle/_ref/Sibelia{,_gpu} did not travel")
def test_non_acgt_input_consumes_rand_like_the_reference(tmp_path):
    """Ns are replaced with rand() % 4 on the host in the reference's order (indexedsequence.cpp:31-37); --inram keeps
    the temp-file names from consuming rand() in the reference."""
    rng = np.random.default_rng(5)
    chrs = [c.copy() for c in helpers.strain_case(3, 60_000, p_sub=0.01, inv_len=5_000, seed=32)]
    for c in chrs:
        for o in rng.integers(0, len(c) - 50, 30):
            c[o:o + int(rng.integers(1, 40))] = ord("N")
    fa = str(tmp_path / "in.fasta")
    write_fasta(fa, chrs)
    compare(["-s", "loose", "-m", "1000", "-r", fa], tmp_path)


@pytest.mark.skipif(not have, reason="oracle/_ref/Sibelia{,_gpu} did not travel")
def test_non_acgt_input_without_inram(tmp_path):
    """Without --inram every index of the reference creates two temp files whose names draw 24 values from the same
    rand() stream (platform.cpp:50-58) before the next index replaces its Ns: the GPU-served CLI builds no files but
    must leave the stream in the same state (facade/gpu_session.h, ConsumeTempFileSideEffects)."""
    rng = np.random.default_rng(6)
    chrs = [c.copy() for c in helpers.strain_case(3, 60_000, p_sub=0.01, inv_len=5_000, seed=33)]
    for c in chrs:
        for o in rng.integers(0, len(c) - 50, 30):
            c[o:o + int(rng.integers(1, 40))] = ord("N")
    fa = str(tmp_path / "in.fasta")
    write_fasta(fa, chrs)
    compare(["-s", "loose", "-m", "1000", fa], tmp_path)


@pytest.mark.skipif(not have, reason="oracle/_ref/Sibelia{,_gpu} did not travel")
@pytest.mark.parametrize("defect", [b">a\nACGT\nACJT\n", b">a\n>b\nACGT\n", b"> x\nACGT\n"])
def test_cli_reports_parse_errors_like_the_reference(tmp_path, defect):
    """the bound CLI reads its FASTA files through sibgpu_fasta_parse (facade/fasta_gpu.cpp): same exception text, same
    exit code as FASTAReader::GetSequences"""
    fa = str(tmp_path / "bad.fasta")
    with open(fa, "wb") as f:
        f.write(b">ok\n" + b"ACGT" * 500 + b"\n" + defect)
    outs = []
    for binary in (REF_BIN, GPU_BIN):
        p = subprocess.run([binary, "-s", "loose", "-o", str(tmp_path / "o"), fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        outs.append((p.returncode, [ln for ln in (p.stdout + p.stderr).splitlines() if b"parse error" in ln]))
    assert outs[0] == outs[1] and outs[0][0] != 0 and outs[0][1], outs


@pytest.mark.skipif(not have, reason="oracle/_ref/Sibelia{,_gpu} did not travel")
def test_condensed_graph_output(tmp_path):
    """-g: BlockFinder::SerializeCondensedGraph (serialization.cpp:88-110) writes the condensed de Bruijn graph of every
    stage (--allstages, sibelia.cpp:256-262) and of the final index (:333-344); in the bound CLI its index + ListEdges
    pair is one sibgpu_list_edges call."""
    chrs = helpers.strain_case(3, 80_000, p_sub=0.01, inv_len=8_000, seed=34)
    fa = str(tmp_path / "in.fasta")
    write_fasta(fa, chrs)
    for sub, extra in (("one", []), ("all", ["--allstages"])):
        outs = {}
        for name, binary in (("ref", REF_BIN), ("gpu", GPU_BIN)):
            outs[name] = str(tmp_path / sub / name)
            os.makedirs(outs[name], exist_ok=True)
            subprocess.run([binary, "-s", "loose", "-m", "1000", "-g"] + extra + [fa, "-o", outs[name]], check=True,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500)
        names = sorted(f for f in os.listdir(outs["ref"]) if f.endswith(".dot") or f.startswith("blocks_coords"))
        assert sum(f.endswith(".dot") for f in names) == (5 if extra else 1), names
        for f in names:
            want = open(os.path.join(outs["ref"], f), "rb").read()
            got = open(os.path.join(outs["gpu"], f), "rb").read()
            assert got == want, "%s differs" % f


@pytest.mark.skipif(not (have and os.path.exists(os.path.join(DATA, "Helicobacter_pylori.fasta"))),
                    reason="reference example genome did not travel")
def test_helicobacter_pylori_loose(tmp_path):
    """BASELINE configs[0]; digests pinned in BASELINE.md section 4 / SURVEY.md section 8(c)."""
    got = compare(["-s", "loose", os.path.join(DATA, "Helicobacter_pylori.fasta")], tmp_path)
    assert hashlib.md5(got["blocks_coords.txt"]).hexdigest() == "9cf97c63809c08c961a5f30036e3dc28"
    assert hashlib.md5(got["genomes_permutations.txt"]).hexdigest() == "a1d4765580a36622839e9065873304d4"


@pytest.mark.skipif(not (have and os.path.exists(os.path.join(DATA, "Staphylococcus.fasta"))),
                    reason="reference example genome did not travel")
def test_staphylococcus_aureus_loose(tmp_path):
    """The reference's second example dataset (4 records, 11.6 Mb; SURVEY.md section 4 known answers: 37 380 / 821 / 255 /
    46 bulges over the four `-s loose` stages): all three output files byte-identical to the unmodified CLI."""
    compare(["-s", "loose", os.path.join(DATA, "Staphylococcus.fasta")], tmp_path)
