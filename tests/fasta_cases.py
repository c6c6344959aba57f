"""FASTA inputs for the ingest parity tests (FASTAReader::GetSequences, /root/reference/src/fasta.cpp:22-106): hand-written
quirks and seeded random files."""
import numpy as np

HAND = [
    b">chr1 some description\nACGT\nacgtn\n>chr2\nTTTT\n",
    b">a\nACGT",                                            # no trailing newline
    b">a\r\nAC\r\nGT\r\n>b\r\nNNNN\r\n",                    # CRLF
    b"\n\n  >x  y z\n\n  ACGT  \n\t\nACGT\n\n",             # blank lines, indented header and sequence
    b"ACGT\nACGT\n>late\nTT\n",                             # sequence before the first header: glued to the first record
    b"ACGT\nAC\n",                                          # no header at all: one record, empty description
    b">a\n>b\nACGT\n",                                      # empty sequence (two headers in a row)
    b">a\nACGT\n>b\n",                                      # empty sequence at the end
    b"",                                                    # empty file
    b"\n\n\n",
    b">\nACGT\n",                                           # empty header
    b"> name\nACGT\n",                                      # blank right after '>': empty header
    b">a\tb c\nACGT\n",                                     # TAB does not end the description
    b">a\nAC GT\n",                                         # interior blank: illegal character
    b">a\nACGTJ\n",                                         # illegal letter
    b">a\nACGT\nAC1T\n>b\nEE\n",                            # first of several errors wins
    b">a\nacgturykmswbdhwnx-\n",                            # every legal character, lower case
    b">a\nACGT\n>b c\n" + b"ACGTACGTAC\n" * 50 + b">c\n" + b"T" * 5000 + b"\n",
    b">a\nAC\x00GT\n",
    b">a\nAC\xe9GT\n",
    b">a b\n>\nACGT\n",                                     # empty sequence is checked before the empty header
    b">only\n",
    b">a\nACGT\n\n\n>b\n\nGG\n \n",
]


def random_case(rng, big=False):
    """A random FASTA with the occasional defect."""
    out = []
    nrec = int(rng.integers(1, 6))
    alphabet = np.frombuffer(b"ACGTacgtNnRYKMSWBDHXU-", dtype=np.uint8)
    eol = b"\r\n" if rng.random() < 0.2 else b"\n"
    for r in range(nrec):
        if rng.random() < 0.92 or r:
            name = b"seq%d" % r if rng.random() < 0.95 else b""
            desc = b" description %d" % r if rng.random() < 0.5 else b""
            out.append(b" " * int(rng.integers(0, 2)) + b">" + name + desc + eol)
        L = int(rng.integers(0, 200_000 if big else 400))
        if rng.random() < 0.06:
            L = 0
        seq = alphabet[rng.integers(0, len(alphabet), L)].tobytes()
        if L and rng.random() < 0.08:
            at = int(rng.integers(0, L))
            seq = seq[:at] + bytes([int(rng.integers(33, 127))]) + seq[at + 1:]
        width = int(rng.choice([60, 70, 80, 7, 1_000_000]))
        for o in range(0, len(seq), width):
            out.append(seq[o:o + width] + (b" " if rng.random() < 0.02 else b"") + eol)
            if rng.random() < 0.02:
                out.append(eol)
    data = b"".join(out)
    if data.endswith(eol) and rng.random() < 0.3:
        data = data[:-len(eol)]
    return data
