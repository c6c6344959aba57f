"""Generates tests/golden/fasta_cases.json from the UNMODIFIED reference (oracle/_ref, FASTAReader::GetSequences) on the
hand-written cases of tests/fasta_cases.py.  Authoring container only."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fasta_cases  # noqa: E402
from test_fasta import reference  # noqa: E402

out = []
with tempfile.TemporaryDirectory() as d:
    for data in fasta_cases.HAND:
        out.append(reference(data, d))
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "fasta_cases.json"), "w"), indent=0)
print(len(out), "cases")
