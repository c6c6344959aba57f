"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_build/liboracle_enum.so (oracle/enum_restate.c, the plain-C
restatement of /root/reference/src/vertexenumeration.cpp:263-364) plus the numpy restatement of the
BifurcationStorage list order (/root/reference/src/indexedsequence.cpp:51-67 + bifurcationstorage.cpp:113-126).
Never imported by sibelia_b200/.  PARITY PINNED against oracle/_ref (tests/test_oracle.py, tests/golden/).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle_enum.so")
INST_DTYPE = np.dtype([("bifId", "<u4"), ("chr", "<u4"), ("pos", "<u4")])
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "restate"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_enumerate.restype = C.c_uint64
        _lib.orc_free.argtypes = [C.c_void_p]
    return _lib


def _take(ptr, n):
    if n:
        buf = (C.c_char * (n * INST_DTYPE.itemsize)).from_address(ptr.value)
        out = np.frombuffer(buf, dtype=INST_DTYPE, count=n).copy()
    else:
        out = np.zeros(0, dtype=INST_DTYPE)
    lib().orc_free(ptr)
    return out


def enumerate_bifurcations(chrs, k):
    """-> (count, pos, neg): the reference's `bifurcationCount` and its two (chr,pos)-sorted instance vectors."""
    L = lib()
    chrs = [c if isinstance(c, (bytes, bytearray)) else (c.tobytes() if isinstance(c, np.ndarray) else c.encode())
            for c in chrs]
    n = len(chrs)
    arr = (C.c_char_p * n)(*chrs)
    lens = (C.c_uint64 * n)(*[len(c) for c in chrs])
    pos, neg = C.c_void_p(), C.c_void_p()
    npos, nneg = C.c_uint64(), C.c_uint64()
    cnt = L.orc_enumerate(C.c_uint32(n), arr, lens, C.c_uint32(k), C.byref(pos), C.byref(npos),
                          C.byref(neg), C.byref(nneg))
    if cnt == 2 ** 64 - 1:
        raise MemoryError("orc_enumerate")
    return int(cnt), _take(pos, npos.value), _take(neg, nneg.value)


def list_positions(count, pos, neg, lens):
    """ListPositions(id) order for id in 0..count (inclusive: the storage holds maxId+1 lists,
    bifurcationstorage.cpp:53-59) as CSR (off, gidx, strand).

    IndexedSequence::Init walks strand 0 then strand 1, chromosomes ascending, positions ascending, and AddPoint
    pushes to the FRONT of the per-id slist (bifurcationstorage.cpp:122), so each list is the reverse of that walk;
    ListPositions emits the positive list and then the negative list (bifurcationstorage.h:59-72).
    gidx = DNASequence::GlobalIndex of the instance's base element (dnasequence.cpp:269-272): elements are numbered
    chr0 '$' chr1 '$' ... from 0; a negative-strand instance at rc-position p sits on element start+len-1-p."""
    lens = np.asarray(lens, dtype=np.int64)
    start = np.concatenate([[0], np.cumsum(lens + 1)[:-1]])
    ids, gidx, strand = [], [], []
    for s, inst in ((0, pos), (1, neg)):
        if len(inst) == 0:
            continue
        c = inst["chr"].astype(np.int64)
        p = inst["pos"].astype(np.int64)
        g = start[c] + (p if s == 0 else lens[c] - 1 - p)
        ids.append(inst["bifId"].astype(np.int64)[::-1])          # reverse insertion order
        gidx.append(g[::-1])
        strand.append(np.full(len(inst), s, dtype=np.uint8))
    if not ids:
        return np.zeros(count + 2, dtype=np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint8)
    ids = np.concatenate(ids)
    gidx = np.concatenate(gidx)
    strand = np.concatenate(strand)
    order = np.lexsort((np.arange(len(ids)), strand, ids))       # stable: by id, then strand, then list order
    off = np.zeros(count + 2, dtype=np.uint64)
    np.cumsum(np.bincount(ids, minlength=count + 1), out=off[1:])
    return off, gidx[order].astype(np.uint32), strand[order]


EDGE_DTYPE = np.dtype([("chr", "<u4"), ("direction", "<u4"), ("start_vertex", "<u4"), ("end_vertex", "<u4"),
                       ("actual_position", "<u4"), ("actual_length", "<u4"), ("original_position", "<u4"),
                       ("original_length", "<u4"), ("first_char", "<u4")])
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def list_edges(chrs, origpos, k):
    """numpy restatement of `IndexedSequence iseq(rawSeq_, originalPos_, k, ""); ListEdges(...)`
    (/root/reference/src/synteny.cpp:238-241, src/serialization.cpp:56-86), pinned against oracle/_ref in
    tests/test_oracle.py: per strand and chromosome the walk emits one Edge per pair of consecutive vertex marks --
    start/end vertex = the two ids, step = distance between the marks, actualPosition = pos (positive strand) or
    length - (pos + step + k) (negative), actualLength = step + k, firstChar = the base k steps after the first mark read
    along the strand (:75), original coordinates = min/max of the stored positions of the first and the last element
    spelled by the edge (DNASequence::SpellOriginal, src/dnasequence.cpp:254-260)."""
    chrs = [np.frombuffer(c, dtype=np.uint8) if isinstance(c, (bytes, bytearray)) else np.asarray(c, dtype=np.uint8)
            for c in chrs]
    count, pos, neg = enumerate_bifurcations(chrs, k)
    lens = np.array([len(c) for c in chrs], dtype=np.int64)
    out = []
    for strand, tab in ((0, pos), (1, neg)):
        if len(tab) < 2:
            continue
        same = tab["chr"][:-1] == tab["chr"][1:]
        a, b = tab[:-1][same], tab[1:][same]
        c = a["chr"].astype(np.int64)
        p = a["pos"].astype(np.int64)
        step = b["pos"].astype(np.int64) - p
        L = lens[c]
        e = np.zeros(len(a), dtype=EDGE_DTYPE)
        e["chr"], e["direction"], e["start_vertex"], e["end_vertex"] = c, strand, a["bifId"], b["bifId"]
        e["actual_position"] = p if strand == 0 else L - (p + step + k)
        e["actual_length"] = step + k
        first_el = p if strand == 0 else L - 1 - p                 # element indices in positive coordinates
        last_el = p + step + k - 1 if strand == 0 else L - 1 - (p + step + k - 1)
        char_el = p + k if strand == 0 else L - 1 - (p + k)
        fc = np.zeros(len(a), dtype=np.uint8)
        o1 = np.zeros(len(a), dtype=np.int64)
        o2 = np.zeros(len(a), dtype=np.int64)
        for ci in np.unique(c):
            m = c == ci
            seq = chrs[ci]
            fc[m] = seq[char_el[m]] if strand == 0 else _COMP[seq[char_el[m]]]
            op = np.arange(len(seq), dtype=np.int64) if origpos is None else np.asarray(origpos[ci], dtype=np.int64)
            o1[m], o2[m] = op[first_el[m]], op[last_el[m]]
        e["first_char"] = fc
        e["original_position"] = np.minimum(o1, o2)
        e["original_length"] = np.maximum(o1, o2) + 1 - np.minimum(o1, o2)
        out.append(e)
    return np.concatenate(out) if out else np.zeros(0, dtype=EDGE_DTYPE)


_FASTA_WS = b" \t\n\v\f\r"
_FASTA_VALID = set(b"ACGTURYKMSWBDHWNX-")


class FastaError(Exception):
    def __init__(self, line, what):
        super().__init__("on line %d: %s" % (line, what))
        self.line, self.what = line, what


def fasta_get_sequences(data):
    """FASTAReader::GetSequences restated (/root/reference/src/fasta.cpp:22-106): list of (description, sequence) bytes, or
    FastaError(line, what) with the reference's line counter (non-empty lines) and message.  Pure Python: small inputs."""
    header, sequence, line, out = b"", bytearray(), 1, []
    for buf in data.split(b"\n"):                        # std::getline; a trailing newline yields one more, empty, line
        buf = buf.strip(_FASTA_WS)                       # boost::algorithm::trim (C locale)
        if not buf:
            continue
        if buf[:1] == b">":
            if header:
                if not sequence:
                    raise FastaError(line, "empty sequence")
                out.append((header, bytes(sequence)))
                sequence = bytearray()
                header = b""
            delim = buf.find(b" ")                       # ValidateHeader :75-90
            delim = len(buf) - 1 if delim < 0 else delim - 1
            name = buf[1:1 + delim] if delim > 0 else b""
            if not name:
                raise FastaError(line, "empty header")
            header = name
        else:
            up = bytearray(buf)
            for i, c in enumerate(buf):                  # ValidateSequence :92-106
                u = c - 32 if 97 <= c <= 122 else c
                if u not in _FASTA_VALID:
                    # the reference builds the message through c_str(): a NUL byte ends it
                    raise FastaError(line, "illegal character: " + (chr(c) if c else ""))
                up[i] = u
            sequence += up
        line += 1
    if not sequence:
        raise FastaError(line, "empty sequence")
    out.append((header, bytes(sequence)))
    return out
