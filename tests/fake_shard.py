"""Test double for sibelia_b200.distributed: the per-rank phases of the sharded enumeration restated in numpy, so the
orchestration (splits, all-to-all layout, key all-gather, assembly) can run under gloo on CPU with world_size > 1.
It follows the same contract as GpuShard: same tile split, records ordered by partition, partition p owned by rank
p // (nparts / world).  k <= 28.

NumpyPeerShard additionally stands in for the peer strategy: fixed-capacity segments in a per-rank "send buffer" (a
file under a shared temp directory plays the CUDA IPC mapping: its 64-byte "handle" is the file name), counts + handle
swapped in one all-gather, the owner of a partition reads the peers' segments itself; a segment that outgrows its
capacity reports overflow, upon which enumerate_sharded must take the staged path on every rank."""
import os
import numpy as np
import torch

TILE = 4096
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x):
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return x


def comp_sym(s):
    return np.where(s == 4, 4, 3 - s)


class NumpyShard:
    words = 1

    def __init__(self, nparts_local=3):
        self.pl = nparts_local

    def upload(self, chrs, rank, world):
        self.chrs = [np.asarray(c, dtype=np.uint8) for c in chrs]
        self.rank, self.world = rank, world
        lens = np.array([len(c) for c in self.chrs], dtype=np.int64)
        self.lens = lens
        self.start = 1 + np.concatenate([[0], np.cumsum(lens + 1)[:-1]]) if len(lens) else np.zeros(0, np.int64)
        self.M = int(lens.sum() + len(lens) + 1)
        ntiles = (self.M + TILE - 1) // TILE
        self.lo = (ntiles * rank // world) * TILE
        self.hi = (ntiles * (rank + 1) // world) * TILE

    def _occurrences(self, k):
        code = np.zeros(256, dtype=np.int64)
        for i, ch in enumerate(b"ACGT"):
            code[ch] = i
        P, C, KEY, RKEY, PREV, NEXT = [], [], [], [], [], []
        for c, (s, L) in enumerate(zip(self.start, self.lens)):
            if L < k:
                continue
            cs = code[self.chrs[c]]
            n = L - k + 1
            pos = np.arange(n)
            tp = s + pos
            sel = (tp >= self.lo) & (tp < self.hi)
            if not sel.any():
                continue
            pos = pos[sel]
            f = np.zeros(len(pos), dtype=np.uint64)
            r = np.zeros(len(pos), dtype=np.uint64)
            for j in range(k):
                f = (f << np.uint64(2)) | cs[pos + j].astype(np.uint64)
                r = (r << np.uint64(2)) | (3 - cs[pos + k - 1 - j]).astype(np.uint64)
            prev = np.where(pos == 0, 4, cs[np.maximum(pos - 1, 0)])
            nxt = np.where(pos + k == L, 4, cs[np.minimum(pos + k, L - 1)])
            P.append(pos); C.append(np.full(len(pos), c)); KEY.append(f); RKEY.append(r); PREV.append(prev); NEXT.append(nxt)
        if not P:
            z = np.zeros(0, dtype=np.int64)
            return z, z, z.astype(np.uint64), z.astype(np.uint64), z, z
        return tuple(np.concatenate(x) for x in (P, C, KEY, RKEY, PREV, NEXT))

    def scan(self, k):
        self.k = k
        pos, chr_, f, r, prev, nxt = self._occurrences(k)
        fw = f <= r
        canon = np.where(fw, f, r)
        cp = np.where(fw, prev, comp_sym(nxt))
        cn = np.where(fw, nxt, comp_sym(prev))
        ctx = (np.where(f == r, 64, 0) | (cp << 3) | cn).astype(np.uint64)
        rec = (canon << np.uint64(7)) | ctx
        nparts = self.pl * self.world
        part = ((mix64(canon) >> np.uint64(32)) * np.uint64(nparts) >> np.uint64(32)).astype(np.int64)
        order = np.argsort(part, kind="stable")
        self.send = rec[order]
        hist = np.bincount(part, minlength=nparts).astype(np.uint32)
        self.occ = (pos, chr_, f, r)
        return nparts, hist

    def scatter(self):
        return torch.from_numpy(self.send.view(np.int64).copy())

    def group(self, recv, counts):
        rec = recv.numpy().view(np.uint64)
        key = rec >> np.uint64(7)
        ctx = (rec & np.uint64(127)).astype(np.int64)
        p, n, pal = (ctx >> 3) & 7, ctx & 7, (ctx & 64) != 0
        bits = (1 << p) | (32 << n)
        bits = np.where(pal, bits | (1 << comp_sym(n)) | (32 << comp_sym(p)), bits)
        uk, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        pay = np.zeros(len(uk), dtype=np.int64)
        np.bitwise_or.at(pay, inv, bits)
        multi = (cnt > 1)
        np.logical_or.at(multi, inv, pal)
        P, N = pay & 31, (pay >> 5) & 31
        pop = lambda x: np.array([bin(int(v)).count("1") for v in x])
        sep = ((P | N) & 16) != 0
        bif = np.where(multi, (pop(P) > 1) | (pop(N) > 1) | sep, sep)
        return torch.from_numpy(uk[bif].view(np.int64).copy())

    def _rc(self, key):
        k = self.k
        out = np.zeros(len(key), dtype=np.uint64)
        x = key.copy()
        for _ in range(k):
            out = (out << np.uint64(2)) | (np.uint64(3) - (x & np.uint64(3)))
            x >>= np.uint64(2)
        return out

    def finish(self, allkeys):
        ck = allkeys.numpy().view(np.uint64)
        inst = np.dtype([("bifId", "<u4"), ("chr", "<u4"), ("pos", "<u4")])
        if len(ck) == 0:
            return 0, np.zeros(0, inst), np.zeros(0, inst)
        v = np.unique(np.concatenate([ck, self._rc(ck)]))
        pos, chr_, f, r = self.occ
        canon = np.minimum(f, r)
        hit = np.isin(canon, ck)
        pos, chr_, f, r = pos[hit], chr_[hit], f[hit], r[hit]
        tp = self.start[chr_] + pos
        o = np.argsort(tp, kind="stable")
        pos, chr_, f, r = pos[o], chr_[o], f[o], r[o]
        P = np.zeros(len(pos), inst)
        P["bifId"], P["chr"], P["pos"] = np.searchsorted(v, f), chr_, pos
        Nt = np.zeros(len(pos), inst)
        Nt["bifId"], Nt["chr"], Nt["pos"] = np.searchsorted(v, r), chr_, self.lens[chr_] - pos - self.k
        return len(v), P, Nt


class NumpyPeerShard(NumpyShard):
    peer = True
    device = "cpu"

    def __init__(self, tmpdir, nparts_local=3, seg_cap=None):
        super().__init__(nparts_local)
        self.tmpdir = tmpdir
        self.seg_cap = seg_cap
        self.fallbacks = 0

    def scatter_local(self, k):
        nparts, hist = self.scan(k)                       # self.send = records ordered by partition
        cap = int(self.seg_cap) if self.seg_cap is not None else int(hist.max()) + 8
        ovf = bool((hist > cap).any())
        buf = np.zeros(nparts * cap, dtype=np.uint64)     # segment p at p * cap, like the device send buffer
        if not ovf:
            off = np.concatenate([[0], np.cumsum(hist)]).astype(np.int64)
            for p in range(nparts):
                buf[p * cap:p * cap + hist[p]] = self.send[off[p]:off[p + 1]]
        else:
            self.fallbacks += 1
        self.path = os.path.join(self.tmpdir, "send_rank%d.npy" % self.rank)
        np.save(self.path, buf)
        self.nparts = nparts
        return nparts, hist.astype(np.uint64), cap, ovf

    def export_send(self):
        h = np.zeros(64, dtype=np.uint8)
        name = os.path.basename(self.path).encode()
        h[:len(name)] = np.frombuffer(name, dtype=np.uint8)
        return h

    def import_peers(self, handles):
        self.peer_files = [os.path.join(self.tmpdir, bytes(h[h != 0]).decode()) for h in np.asarray(handles, dtype=np.uint8)]
        return all(os.path.exists(f) for f in self.peer_files)

    def group_peer(self, counts, seg_caps):
        pl = self.nparts // self.world
        segs = []
        for p in range(self.rank * pl, (self.rank + 1) * pl):
            for s in range(self.world):
                buf = np.load(self.peer_files[s], mmap_mode="r")
                cap, n = int(seg_caps[s]), int(counts[s, p])
                segs.append(np.asarray(buf[p * cap:p * cap + n]))
        rec = np.concatenate(segs) if segs else np.zeros(0, dtype=np.uint64)
        return self.group(torch.from_numpy(rec.view(np.int64).copy()), None)


class FakeFusedFpContext:
    """Stand-in for the library behind GpuShard's fused k > 32 path (sibgpu_fused_plan / _alloc / _import / _run_fp /
    _finish_fp), so that the orchestration of sibelia_b200.distributed -- collective (re)allocation of the exported
    buffers, the min-reduction of the class representatives between the two halves of a step, the max-reduction of the
    verification flag and the repeat with other hash bases -- runs under gloo on CPU.  The tables come from the oracle;
    a "class" is a vertex id and its representative the smallest text position of its positive-strand instances.
    script: regrow = the first run reports a key region too small (status 2); collide_rank = that rank reports a
    verification failure on attempt 0."""
    SENTINEL = 0x7F7F7F7F7F7F7F7F

    def __init__(self, rank, world, regrow=False, collide_rank=None):
        self.rank, self.world = rank, world
        self.regrow_pending, self.collide_rank = regrow, collide_rank
        self.alloc_needed = True
        self.log = []

    def dist2_plan(self, chrs, rank, world, k, resident):
        assert (rank, world) == (self.rank, self.world)
        self.chrs, self.k = [np.asarray(c, dtype=np.uint8) for c in chrs], k
        self.log.append("plan")
        return 1 if self.alloc_needed else 0

    def dist2_release_peers(self):
        self.log.append("release")

    def dist2_alloc(self):
        self.alloc_needed = False
        self.log.append("alloc")
        return np.full(64, self.rank, dtype=np.uint8)

    def dist2_import(self, handles):
        assert handles.shape == (self.world, 64) and all((handles[s] == s).all() for s in range(self.world))
        self.log.append("import")

    def _tables(self):
        from oracle import restate
        count, pos, neg = restate.enumerate_bifurcations(self.chrs, self.k)
        lens = np.array([len(c) for c in self.chrs], dtype=np.int64)
        start = 1 + np.concatenate([[0], np.cumsum(lens + 1)[:-1]])
        M = int(lens.sum() + len(lens) + 1)
        ntiles = (M + TILE - 1) // TILE
        lo, hi = ntiles * self.rank // self.world * TILE, ntiles * (self.rank + 1) // self.world * TILE
        tp = start[pos["chr"]] + pos["pos"]
        tn = start[neg["chr"]] + (lens[neg["chr"]] - neg["pos"] - self.k)
        keep_n = np.flatnonzero((tn >= lo) & (tn < hi))
        keep_n = keep_n[np.argsort(tn[keep_n], kind="stable")]
        own = (tp >= lo) & (tp < hi)
        glob = np.full(count, self.SENTINEL, dtype=np.int64)
        np.minimum.at(glob, pos["bifId"].astype(np.int64), tp)
        mine = np.full(count, self.SENTINEL, dtype=np.int64)
        np.minimum.at(mine, pos["bifId"][own].astype(np.int64), tp[own])
        return count, pos[own], neg[keep_n], mine, glob

    def dist2_run_fp(self, chrs, resident, attempt):
        self.log.append("run%d" % attempt)
        if self.regrow_pending:
            self.regrow_pending = False
            self.alloc_needed = True                     # the next plan asks for the collective reallocation
            return 2, 0, 0
        self.attempt = attempt
        self.count, self.pos, self.negtext, self.rep, self.want_rep = self._tables()
        return 0, len(self.rep), 1

    def dist2_run(self, chrs, resident):
        """k <= 28: the whole step is one call (sibgpu_fused_run)"""
        self.log.append("run")
        if self.regrow_pending:
            self.regrow_pending = False
            self.alloc_needed = True
            return 2, 0, 0
        self.count, self.pos, self.negtext, _, _ = self._tables()
        return 0, self.count, len(self.pos)

    def dist2_finish_fp(self):
        assert np.array_equal(self.rep, self.want_rep), "the representatives were not min-reduced over the ranks"
        self.log.append("finish%d" % self.attempt)
        collision = 1 if (self.attempt == 0 and self.collide_rank == self.rank) else 0
        return self.count, len(self.pos), collision

    def download(self):
        return self.pos, self.negtext


def fake_fused_fp_shard(rank, world, **script):
    """a GpuShard whose context is the fake above and whose representative array lives in host memory"""
    from sibelia_b200.distributed import GpuShard

    class Shard(GpuShard):
        def __init__(self):
            self.ctx = FakeFusedFpContext(rank, world, **script)
            self.device = torch.device("cpu")
            self.fused = True

        def _rep_tensor(self, ptr, n):
            return torch.from_numpy(self.ctx.rep)        # shares memory: the all-reduce lands in the fake's array
    return Shard()
