"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libsibelia_ref.so (the UNMODIFIED reference hot path,
built by oracle/Makefile from /root/reference/src).  Never imported by sibelia_b200/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsibelia_ref.so")

INST_DTYPE = np.dtype([("bifId", "<u4"), ("chr", "<u4"), ("pos", "<u4")])


class _Inst(C.Structure):
    _fields_ = [("bifId", C.c_uint32), ("chr", C.c_uint32), ("pos", C.c_uint32)]


_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_free.argtypes = [C.c_void_p]
        _lib.ref_index.restype = C.c_int
        _lib.ref_simplify.restype = C.c_int
    return _lib


def _as_bytes_list(chrs):
    return [c if isinstance(c, (bytes, bytearray)) else (c.tobytes() if isinstance(c, np.ndarray) else c.encode())
            for c in chrs]


def _take(ptr, n, dtype):
    """Copy n items of dtype from a malloc'ed C pointer and free it."""
    if n:
        buf = (C.c_char * (n * dtype.itemsize)).from_address(C.cast(ptr, C.c_void_p).value)
        out = np.frombuffer(buf, dtype=dtype, count=n).copy()
    else:
        out = np.zeros(0, dtype=dtype)
    lib().ref_free(C.cast(ptr, C.c_void_p))
    return out


def index(chrs, k, dump=True):
    """IndexedSequence(record, k, "") of the reference.  Returns dict(maxId, pos, neg, lp_off, lp_gidx, lp_strand,
    seconds); pos/neg are structured arrays (bifId, chr, pos) sorted by (chr, pos)."""
    L = lib()
    chrs = _as_bytes_list(chrs)
    n = len(chrs)
    arr = (C.c_char_p * n)(*chrs)
    lens = (C.c_uint64 * n)(*[len(c) for c in chrs])
    maxId = C.c_uint32()
    pos = C.POINTER(_Inst)()
    neg = C.POINTER(_Inst)()
    npos = C.c_uint64()
    nneg = C.c_uint64()
    lp_off = C.POINTER(C.c_uint64)()
    lp_g = C.POINTER(C.c_uint32)()
    lp_s = C.POINTER(C.c_uint8)()
    sec = C.c_double()
    rc = L.ref_index(C.c_uint32(n), arr, lens, C.c_uint32(k), C.c_int(1 if dump else 0), C.byref(maxId),
                     C.byref(pos), C.byref(npos), C.byref(neg), C.byref(nneg),
                     C.byref(lp_off), C.byref(lp_g), C.byref(lp_s), C.byref(sec))
    if rc != 0:
        raise RuntimeError(L.ref_last_error().decode())
    out = {"maxId": maxId.value, "seconds": sec.value}
    if dump:
        out["pos"] = _take(pos, npos.value, INST_DTYPE)
        out["neg"] = _take(neg, nneg.value, INST_DTYPE)
        off = _take(lp_off, maxId.value + 2, np.dtype("<u8"))
        out["lp_off"] = off
        out["lp_gidx"] = _take(lp_g, int(off[-1]), np.dtype("<u4"))
        out["lp_strand"] = _take(lp_s, int(off[-1]), np.dtype("u1"))
    return out


def simplify(chrs, origpos, k, D, iters=4):
    """One BlockFinder::PerformGraphSimplifications(k, D, iters) stage seeded with (rawSeq_, originalPos_).
    Returns (new_chrs [bytes], new_origpos [uint32 arrays], bulges, seconds)."""
    L = lib()
    chrs = _as_bytes_list(chrs)
    n = len(chrs)
    keep = [C.create_string_buffer(c, len(c) + 1) for c in chrs]
    seq = (C.c_void_p * n)(*[C.cast(b, C.c_void_p).value for b in keep])
    ops = [np.ascontiguousarray(o, dtype=np.uint32) for o in origpos]
    op = (C.c_void_p * n)(*[o.ctypes.data for o in ops])
    lens = (C.c_uint64 * n)(*[len(c) for c in chrs])
    bulges = C.c_uint64()
    sec = C.c_double()
    rc = L.ref_simplify(C.c_uint32(n), seq, op, lens, C.c_uint32(k), C.c_uint32(D), C.c_uint32(iters),
                        C.byref(bulges), C.byref(sec))
    if rc != 0:
        raise RuntimeError(L.ref_last_error().decode())
    new_chrs, new_op = [], []
    for i in range(n):
        m = lens[i]
        new_chrs.append(_take(C.c_void_p(seq[i]), m, np.dtype("u1")).tobytes())
        new_op.append(_take(C.c_void_p(op[i]), m, np.dtype("<u4")))
    return new_chrs, new_op, bulges.value, sec.value


EDGE_DTYPE = np.dtype([("chr", "<u4"), ("direction", "<u4"), ("start_vertex", "<u4"), ("end_vertex", "<u4"),
                       ("actual_position", "<u4"), ("actual_length", "<u4"), ("original_position", "<u4"),
                       ("original_length", "<u4"), ("first_char", "<u4")])


def list_edges(chrs, origpos, k):
    """IndexedSequence(rawSeq_, originalPos_, k, "") + BlockFinder::ListEdges of the reference (synteny.cpp:238-241).
    Returns (edges structured array in the reference's order, seconds)."""
    L = lib()
    chrs = _as_bytes_list(chrs)
    n = len(chrs)
    arr = (C.c_char_p * n)(*chrs)
    ops = [np.ascontiguousarray(o, dtype=np.uint32) for o in origpos]
    op = (C.c_void_p * n)(*[o.ctypes.data for o in ops])
    lens = (C.c_uint64 * n)(*[len(c) for c in chrs])
    edges = C.c_void_p()
    ne = C.c_uint64()
    sec = C.c_double()
    L.ref_list_edges.restype = C.c_int
    rc = L.ref_list_edges(C.c_uint32(n), arr, op, lens, C.c_uint32(k), C.byref(edges), C.byref(ne), C.byref(sec))
    if rc != 0:
        raise RuntimeError(L.ref_last_error().decode())
    return _take(edges, ne.value, EDGE_DTYPE), sec.value


def trim_blocks(chrs, directions, trim_k, min_size):
    """BlockFinder::TrimBlocks of the reference on a block made of whole sequences (block[i] = sequence i read along
    directions[i]).  Returns (list of (chr, originalPosition, originalLength), drop)."""
    L = lib()
    chrs = _as_bytes_list(chrs)
    n = len(chrs)
    arr = (C.c_char_p * max(n, 1))(*chrs)
    lens = (C.c_uint64 * max(n, 1))(*[len(c) for c in chrs])
    d = np.ascontiguousarray(directions, dtype=np.uint8)
    out = np.zeros(3 * max(n, 1), dtype=np.uint32)
    nout, drop = C.c_uint32(), C.c_int()
    L.ref_trim_blocks.restype = C.c_int
    rc = L.ref_trim_blocks(C.c_uint32(n), arr, lens, C.c_void_p(d.ctypes.data), C.c_uint32(trim_k), C.c_uint32(min_size),
                           C.c_void_p(out.ctypes.data), C.byref(nout), C.byref(drop))
    if rc != 0:
        raise RuntimeError(L.ref_last_error().decode())
    return [tuple(int(x) for x in out[3 * i:3 * i + 3]) for i in range(nout.value)], bool(drop.value)


def boost_order(keys):
    """Iteration order of the vendored boost::unordered_map<size_t,int> after inserting distinct keys in order."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.zeros(len(keys), dtype=np.uint64)
    lib().ref_boost_order(C.c_void_p(keys.ctypes.data), C.c_uint64(len(keys)), C.c_void_p(out.ctypes.data))
    return out


def fasta_parse(path):
    """FASTAReader(path).GetSequences of the reference.  Returns ([(description bytes, sequence bytes)], seconds) or raises
    RuntimeError with the reference's exception text."""
    L = lib()
    L.ref_fasta_parse.restype = C.c_int
    nrec, nb = C.c_uint32(), C.c_uint64()
    names, seqs, lens, sec = C.c_void_p(), C.c_void_p(), C.POINTER(C.c_uint64)(), C.c_double()
    rc = L.ref_fasta_parse(C.c_char_p(path.encode()), C.byref(nrec), C.byref(names), C.byref(nb), C.byref(seqs), C.byref(lens),
                           C.byref(sec))
    if rc != 0:
        raise RuntimeError(L.ref_last_error().decode("latin-1"))
    out, na, sa = [], 0, 0
    nbuf = C.string_at(names, nb.value)
    for i in range(nrec.value):
        end = nbuf.index(b"\0", na)
        n = lens[i]
        out.append((nbuf[na:end], C.string_at(seqs.value + sa, n)))
        na = end + 1
        sa += n
    for p in (names, seqs, C.cast(lens, C.c_void_p)):
        L.ref_free(p)
    return out, sec.value
