"""ctypes binding of libsibgpu.so (include/sibgpu.h).  Fails loudly when the library or a GPU is missing."""
import ctypes as C
import os
import subprocess
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
INST_DTYPE = np.dtype([("bifId", "<u4"), ("chr", "<u4"), ("pos", "<u4")])
EDGE_DTYPE = np.dtype([("chr", "<u4"), ("direction", "<u4"), ("start_vertex", "<u4"), ("end_vertex", "<u4"),
                       ("actual_position", "<u4"), ("actual_length", "<u4"), ("original_position", "<u4"),
                       ("original_length", "<u4"), ("first_char", "<u4")])

# every symbol include/sibgpu.h declares (tests/test_abi.py cross-checks this list against the header)
SYMBOLS = [
    "sibgpu_last_error", "sibgpu_version", "sibgpu_device_count", "sibgpu_create", "sibgpu_destroy", "sibgpu_free",
    "sibgpu_enumerate", "sibgpu_list_edges", "sibgpu_trim_blocks", "sibgpu_upload", "sibgpu_enumerate_resident", "sibgpu_download", "sibgpu_set_profiling",
    "sibgpu_kernel_stats", "sibgpu_last_launches", "sibgpu_partition_fallbacks", "sibgpu_bucket_fallbacks", "sibgpu_last_device_ms", "sibgpu_simplify", "sibgpu_debug_unordered_order", "sibgpu_debug_trim_from_tables", "sibgpu_dist_upload", "sibgpu_dist_scan", "sibgpu_dist_record_bytes",
    "sibgpu_dist_scatter", "sibgpu_dist_group", "sibgpu_dist_keys", "sibgpu_dist_finish",
    "sibgpu_dist_scatter_local", "sibgpu_dist_upload_scatter", "sibgpu_dist_export_send", "sibgpu_dist_import_peers", "sibgpu_dist_group_peer",
    "sibgpu_fasta_parse", "sibgpu_fasta_free",
    "sibgpu_fused_plan", "sibgpu_fused_release_peers", "sibgpu_fused_alloc", "sibgpu_fused_import", "sibgpu_fused_run",
    "sibgpu_fused_run_fp", "sibgpu_fused_finish_fp",
]


class SibgpuError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("sibgpu status %d: %s" % (status, msg))
        self.status = status


class _FastaRec(C.Structure):
    _fields_ = [("name", C.c_char_p), ("name_len", C.c_uint64), ("seq", C.c_void_p), ("len", C.c_uint64)]


class _Fasta(C.Structure):
    _fields_ = [("nrec", C.c_uint32), ("rec", C.POINTER(_FastaRec)), ("text_block", C.c_void_p), ("name_block", C.c_void_p),
                ("total", C.c_uint64)]


class FastaParseError(RuntimeError):
    """The reference's ParseException: `line` is its line counter, `what` its message (fasta.cpp:66-70)."""
    def __init__(self, line, what):
        super().__init__("parse error on line %d: %s" % (line, what))
        self.line = line
        self.what = what


class _KStat(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_uint32), ("ms", C.c_float), ("algo_bytes", C.c_uint64)]


def lib_path():
    # SIBGPU_LIB: a differently compiled build of the same sources (tools/build_variant.sh, kernel tuning only)
    return os.environ.get("SIBGPU_LIB") or os.path.join(_HERE, "libsibgpu.so")


def build(verbose=False):
    """Compile libsibgpu.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", _HERE, "-j8"] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)
    return lib_path()


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path()):
            raise ImportError("libsibgpu.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(lib_path())
        L.sibgpu_last_error.restype = C.c_char_p
        L.sibgpu_version.restype = C.c_char_p
        L.sibgpu_last_launches.restype = C.c_uint64
        L.sibgpu_last_launches.argtypes = [C.c_void_p]
        L.sibgpu_partition_fallbacks.restype = C.c_uint64
        L.sibgpu_partition_fallbacks.argtypes = [C.c_void_p]
        L.sibgpu_bucket_fallbacks.restype = C.c_uint64
        L.sibgpu_bucket_fallbacks.argtypes = [C.c_void_p]
        L.sibgpu_last_device_ms.restype = C.c_float
        L.sibgpu_last_device_ms.argtypes = [C.c_void_p]
        L.sibgpu_destroy.argtypes = [C.c_void_p]
        L.sibgpu_destroy.restype = None
        L.sibgpu_free.argtypes = [C.c_void_p]
        L.sibgpu_free.restype = None
        _lib = L
    return _lib


def device_count():
    return load().sibgpu_device_count()


def _check(rc):
    if rc != 0:
        raise SibgpuError(rc, load().sibgpu_last_error().decode())


def _chr_args(chrs):
    bufs = []
    for c in chrs:
        if isinstance(c, np.ndarray):
            bufs.append(np.ascontiguousarray(c, dtype=np.uint8))
        else:
            bufs.append(np.frombuffer(bytes(c) if not isinstance(c, str) else c.encode(), dtype=np.uint8))
    n = len(bufs)
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data if len(b) else None for b in bufs])
    lens = (C.c_uint64 * max(n, 1))(*[len(b) for b in bufs])
    return bufs, ptrs, lens, n


def _take(ptr, n):
    """The library's table as a numpy array WITHOUT a copy: the array owns the (pinned, pooled) buffer and hands it back
    with sibgpu_free when the last view dies."""
    L = load()
    if n:
        buf = (C.c_char * (n * INST_DTYPE.itemsize)).from_address(ptr.value)
        weakref.finalize(buf, L.sibgpu_free, C.c_void_p(ptr.value))
        return np.frombuffer(buf, dtype=INST_DTYPE, count=n)
    if ptr.value:
        L.sibgpu_free(ptr)
    return np.zeros(0, dtype=INST_DTYPE)


class Context:
    """One sibgpu context (= one GPU)."""

    def __init__(self, device=0):
        L = load()
        self._h = C.c_void_p()
        _check(L.sibgpu_create(C.c_int(device), C.byref(self._h)))
        self._keep = None

    def close(self):
        if self._h:
            load().sibgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- sibgpu_enumerate: host buffers in, host tables out
    def enumerate(self, chrs, k):
        L = load()
        bufs, ptrs, lens, n = _chr_args(chrs)
        pos, neg = C.c_void_p(), C.c_void_p()
        npos, nneg, cnt = C.c_uint64(), C.c_uint64(), C.c_uint32()
        _check(L.sibgpu_enumerate(self._h, ptrs, lens, C.c_uint32(n), C.c_uint32(k), C.byref(pos), C.byref(npos),
                                  C.byref(neg), C.byref(nneg), C.byref(cnt)))
        return cnt.value, _take(pos, npos.value), _take(neg, nneg.value)

    # -- sibgpu_list_edges: index + ListEdges without a host-side index
    def list_edges(self, chrs, origpos, k):
        L = load()
        bufs, ptrs, lens, n = _chr_args(chrs)
        if origpos is None:
            op = None
        else:
            ops = [np.ascontiguousarray(o, dtype=np.uint32) for o in origpos]
            op = (C.c_void_p * max(n, 1))(*[o.ctypes.data if len(o) else None for o in ops])
        edges, ne = C.c_void_p(), C.c_uint64()
        _check(L.sibgpu_list_edges(self._h, ptrs, op, lens, C.c_uint32(n), C.c_uint32(k), C.byref(edges), C.byref(ne)))
        if ne.value:
            buf = (C.c_char * (ne.value * EDGE_DTYPE.itemsize)).from_address(edges.value)
            out = np.frombuffer(buf, dtype=EDGE_DTYPE, count=ne.value).copy()
        else:
            out = np.zeros(0, dtype=EDGE_DTYPE)
        if edges.value:
            L.sibgpu_free(edges)
        return out

    # -- sibgpu_trim_blocks: the index-and-search part of BlockFinder::TrimBlocks
    def trim_blocks(self, chrs, directions, trim_k):
        bufs, ptrs, lens, n = _chr_args(chrs)
        d = np.ascontiguousarray(directions, dtype=np.uint8)
        out = np.zeros((max(n, 1), 3), dtype=np.uint32)
        _check(load().sibgpu_trim_blocks(self._h, ptrs, lens, C.c_void_p(d.ctypes.data), C.c_uint32(n), C.c_uint32(trim_k),
                                         C.c_void_p(out.ctypes.data)))
        return out[:n]

    # -- staged form
    def upload(self, chrs):
        bufs, ptrs, lens, n = _chr_args(chrs)
        _check(load().sibgpu_upload(self._h, ptrs, lens, C.c_uint32(n)))

    def enumerate_resident(self, k):
        ninst, cnt = C.c_uint64(), C.c_uint32()
        _check(load().sibgpu_enumerate_resident(self._h, C.c_uint32(k), C.byref(ninst), C.byref(cnt)))
        return cnt.value, ninst.value

    def download(self):
        pos, neg = C.c_void_p(), C.c_void_p()
        npos, nneg = C.c_uint64(), C.c_uint64()
        _check(load().sibgpu_download(self._h, C.byref(pos), C.byref(npos), C.byref(neg), C.byref(nneg)))
        return _take(pos, npos.value), _take(neg, nneg.value)

    def set_profiling(self, on):
        _check(load().sibgpu_set_profiling(self._h, C.c_int(1 if on else 0)))

    def kernel_stats(self):
        arr = (_KStat * 64)()
        n = load().sibgpu_kernel_stats(self._h, arr, C.c_int(64))
        return [dict(name=arr[i].name.decode(), launches=arr[i].launches, ms=arr[i].ms, algo_bytes=arr[i].algo_bytes)
                for i in range(min(n, 64))]

    def last_device_ms(self):
        return float(load().sibgpu_last_device_ms(self._h))

    def last_launches(self):
        return int(load().sibgpu_last_launches(self._h))

    def partition_fallbacks(self):
        return int(load().sibgpu_partition_fallbacks(self._h))

    def bucket_fallbacks(self):
        return int(load().sibgpu_bucket_fallbacks(self._h))

    # -- sharded enumeration phases (see sibelia_b200/distributed.py for the orchestration)
    def dist_upload(self, chrs, rank, world):
        bufs, ptrs, lens, n = _chr_args(chrs)
        _check(load().sibgpu_dist_upload(self._h, ptrs, lens, C.c_uint32(n), C.c_uint32(rank), C.c_uint32(world)))

    def dist_scan(self, k):
        hist = np.zeros(1024, dtype=np.uint32)
        nparts, nrec = C.c_uint32(), C.c_uint64()
        _check(load().sibgpu_dist_scan(self._h, C.c_uint32(k), C.byref(nparts), C.c_void_p(hist.ctypes.data), C.byref(nrec)))
        return nparts.value, hist[:nparts.value].copy(), nrec.value

    def dist_record_bytes(self):
        f = load().sibgpu_dist_record_bytes
        f.restype = C.c_uint32
        return int(f(self._h))

    def dist_scatter(self, send_ptr):
        _check(load().sibgpu_dist_scatter(self._h, C.c_void_p(send_ptr)))

    def dist_group(self, recv_ptr, counts):
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        nkeys = C.c_uint64()
        _check(load().sibgpu_dist_group(self._h, C.c_void_p(recv_ptr), C.c_void_p(counts.ctypes.data), C.byref(nkeys)))
        return nkeys.value

    # -- peer variant: fixed-capacity segments in the own send buffer, peers read them over NVLink (CUDA IPC)
    def dist_scatter_local(self, k):
        counts = np.zeros(1024, dtype=np.uint64)
        nparts, cap, ovf = C.c_uint32(), C.c_uint64(), C.c_int()
        _check(load().sibgpu_dist_scatter_local(self._h, C.c_uint32(k), C.byref(nparts), C.c_void_p(counts.ctypes.data),
                                                C.byref(cap), C.byref(ovf)))
        return nparts.value, counts[:nparts.value].copy(), cap.value, bool(ovf.value)

    def dist_upload_scatter(self, chrs, rank, world, k):
        bufs, ptrs, lens, n = _chr_args(chrs)
        counts = np.zeros(1024, dtype=np.uint64)
        nparts, cap, ovf = C.c_uint32(), C.c_uint64(), C.c_int()
        _check(load().sibgpu_dist_upload_scatter(self._h, ptrs, lens, C.c_uint32(n), C.c_uint32(rank), C.c_uint32(world),
                                                 C.c_uint32(k), C.byref(nparts), C.c_void_p(counts.ctypes.data),
                                                 C.byref(cap), C.byref(ovf)))
        return nparts.value, counts[:nparts.value].copy(), cap.value, bool(ovf.value)

    def dist_export_send(self):
        h = np.zeros(64, dtype=np.uint8)
        _check(load().sibgpu_dist_export_send(self._h, C.c_void_p(h.ctypes.data)))
        return h

    def dist_import_peers(self, handles):
        handles = np.ascontiguousarray(handles, dtype=np.uint8)
        _check(load().sibgpu_dist_import_peers(self._h, C.c_void_p(handles.ctypes.data)))

    def dist_group_peer(self, counts, seg_caps):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        seg_caps = np.ascontiguousarray(seg_caps, dtype=np.uint64)
        nkeys = C.c_uint64()
        _check(load().sibgpu_dist_group_peer(self._h, C.c_void_p(counts.ctypes.data), C.c_void_p(seg_caps.ctypes.data),
                                             C.byref(nkeys)))
        return nkeys.value

    # -- fused variant: one exported buffer per rank, exchange inside the kernels (see include/sibgpu.h)
    def dist2_plan(self, chrs, rank, world, k, resident=False):
        bufs, ptrs, lens, n = _chr_args(chrs)
        need = C.c_int()
        _check(load().sibgpu_fused_plan(self._h, ptrs, lens, C.c_uint32(n), C.c_uint32(rank), C.c_uint32(world), C.c_uint32(k),
                                        C.c_int(1 if resident else 0), C.byref(need)))
        return need.value

    def dist2_release_peers(self):
        _check(load().sibgpu_fused_release_peers(self._h))

    def dist2_alloc(self):
        h = np.zeros(64, dtype=np.uint8)
        _check(load().sibgpu_fused_alloc(self._h, C.c_void_p(h.ctypes.data)))
        return h

    def dist2_import(self, handles):
        handles = np.ascontiguousarray(handles, dtype=np.uint8)
        _check(load().sibgpu_fused_import(self._h, C.c_void_p(handles.ctypes.data)))

    def dist2_run(self, chrs, resident=False):
        bufs, ptrs, lens, n = _chr_args(chrs)
        cnt, ninst, status = C.c_uint32(), C.c_uint64(), C.c_int()
        _check(load().sibgpu_fused_run(self._h, ptrs, lens, C.c_uint32(n), C.c_int(1 if resident else 0), C.byref(cnt),
                                       C.byref(ninst), C.byref(status)))
        return status.value, cnt.value, ninst.value

    def dist2_run_fp(self, chrs, resident=False, attempt=0):
        """k > 32, first half: (status, classes, device pointer of the class representatives to min-reduce)"""
        bufs, ptrs, lens, n = _chr_args(chrs)
        ncls, rep, status = C.c_uint64(), C.c_void_p(), C.c_int()
        _check(load().sibgpu_fused_run_fp(self._h, ptrs, lens, C.c_uint32(n), C.c_int(1 if resident else 0), C.c_uint32(attempt),
                                          C.byref(ncls), C.byref(rep), C.byref(status)))
        return status.value, ncls.value, rep.value or 0

    def dist2_finish_fp(self):
        cnt, ninst, coll = C.c_uint32(), C.c_uint64(), C.c_int()
        _check(load().sibgpu_fused_finish_fp(self._h, C.byref(cnt), C.byref(ninst), C.byref(coll)))
        return cnt.value, ninst.value, coll.value

    def dist_keys(self, keys_ptr):
        _check(load().sibgpu_dist_keys(self._h, C.c_void_p(keys_ptr)))

    def dist_finish(self, allkeys_ptr, nkeys_total):
        ninst, cnt = C.c_uint64(), C.c_uint32()
        _check(load().sibgpu_dist_finish(self._h, C.c_void_p(allkeys_ptr), C.c_uint64(nkeys_total), C.byref(ninst), C.byref(cnt)))
        return cnt.value, ninst.value

    # -- sibgpu_fasta_parse: FASTAReader::GetSequences on the GPU
    def fasta_parse(self, data):
        """bytes of a FASTA file -> list of (description: bytes, sequence: uint8 array); raises FastaParseError."""
        L = load()
        data = bytes(data)
        f = _Fasta()
        line = C.c_uint64()
        rc = L.sibgpu_fasta_parse(self._h, C.c_char_p(data), C.c_uint64(len(data)), C.byref(f), C.byref(line))
        if rc == 3:
            raise FastaParseError(line.value, L.sibgpu_last_error().decode("latin-1"))
        _check(rc)
        out = []
        for i in range(f.nrec):
            r = f.rec[i]
            seq = np.frombuffer((C.c_char * r.len).from_address(r.seq), dtype=np.uint8).copy() if r.len else np.zeros(0, np.uint8)
            out.append((C.string_at(r.name, r.name_len), seq))
        L.sibgpu_fasta_free(C.byref(f))
        return out

    # -- sibgpu_simplify: one PerformGraphSimplifications stage
    def simplify(self, chrs, origpos, k, min_branch_size, max_iterations=4):
        L = load()
        bufs, _, lens, n = _chr_args(chrs)
        seq = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in bufs])
        ops = [np.ascontiguousarray(o, dtype=np.uint32) for o in origpos]
        op = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in ops])
        bulges = C.c_uint64()
        _check(L.sibgpu_simplify(self._h, seq, op, lens, C.c_uint32(n), C.c_uint32(k), C.c_uint32(min_branch_size),
                                 C.c_uint32(max_iterations), None, None, C.byref(bulges)))
        new_chrs, new_op = [], []
        for i in range(n):
            m = lens[i]
            if seq[i] == (bufs[i].ctypes.data if len(bufs[i]) else None) or (not seq[i] and not len(bufs[i])):
                # the stage found no bulge at all: the library left the caller's arrays untouched (include/sibgpu.h)
                new_chrs.append(chrs[i])
                new_op.append(origpos[i])
                continue
            sbuf = (C.c_char * m).from_address(seq[i]) if m else b""
            new_chrs.append(bytes(sbuf))
            obuf = (C.c_char * (4 * m)).from_address(op[i]) if m else b""
            new_op.append(np.frombuffer(bytes(obuf), dtype=np.uint32).copy())
            L.sibgpu_free(C.c_void_p(seq[i]))
            L.sibgpu_free(C.c_void_p(op[i]))
        return new_chrs, new_op, bulges.value


def debug_trim_from_tables(count, pos, neg, lens, directions):
    """Host-only test hook: trim points from given instance tables (see sibgpu_trim_blocks)."""
    pos = np.ascontiguousarray(pos, dtype=INST_DTYPE)
    neg = np.ascontiguousarray(neg, dtype=INST_DTYPE)
    lens = np.ascontiguousarray(lens, dtype=np.uint64)
    d = np.ascontiguousarray(directions, dtype=np.uint8)
    out = np.zeros((max(len(lens), 1), 3), dtype=np.uint32)
    f = load().sibgpu_debug_trim_from_tables
    f.restype = None
    f(C.c_void_p(pos.ctypes.data), C.c_uint64(len(pos)), C.c_void_p(neg.ctypes.data), C.c_uint64(len(neg)), C.c_uint32(count),
      C.c_void_p(lens.ctypes.data), C.c_void_p(d.ctypes.data), C.c_uint32(len(lens)), C.c_void_p(out.ctypes.data))
    return out[:len(lens)]


def debug_unordered_order(keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.zeros(len(keys), dtype=np.uint64)
    f = load().sibgpu_debug_unordered_order
    f.restype = None
    f(C.c_void_p(keys.ctypes.data), C.c_uint64(len(keys)), C.c_void_p(out.ctypes.data))
    return out
