"""Sharded bifurcation enumeration over several GPUs, one process per GPU (torch.distributed is only the plumbing).

    count, pos_part, negtext_part = enumerate_sharded(GpuShard(ctx), chrs, k)      # on every rank
    count, pos, neg = gather_tables(count, pos_part, negtext_part)                  # full tables on rank 0

The compute phases live in libsibgpu (sibgpu_dist_*, include/sibgpu.h).  Two exchange strategies:

* peer (default on GPUs): every rank scatters its records into fixed-capacity segments of its own send buffer (no
  histogram pass); the ranks all-gather the segment counts + the CUDA IPC handle of the buffer (one small collective,
  which is also the barrier), and the owner of a hash partition reads that partition's segments straight out of the
  peers' send buffers over NVLink inside its insert kernel -- the all-to-all is fused into the grouping kernel.
* staged (fallback when a segment overflows, and the path of the CPU test double): histogram, scatter into an exactly
  sized send buffer, ONE all-to-all of the k-mer records (NCCL on the device buffers, or gloo through host memory).

Either way one all-gather of the (few) vertex keys follows, so every rank can compute the global ids.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

_TRACE = os.environ.get("SIBGPU_TRACE_DIST") is not None


class GpuShard:
    """Backend of enumerate_sharded that runs the phases on this rank's GPU through the C ABI."""

    def __init__(self, ctx, device=None):
        self.ctx = ctx
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def upload(self, chrs, rank, world):
        # deferred: the peer strategy pipelines the copy of the own text range with its pack and scatter kernels
        self._pending = (chrs, rank, world)

    def _upload_now(self):
        if getattr(self, "_pending", None) is not None:
            self.ctx.dist_upload(*self._pending)
            self._pending = None

    # -- fused strategy (k <= 28): the whole step is one C call, no collective inside (include/sibgpu.h, sibgpu_fused_*)
    fused = os.environ.get("SIBGPU_DIST_FUSED", "1") != "0"
    resident = False                                 # the text range is already in HBM (ctx.dist_upload)

    def fused_enumerate(self, chrs, rank, world, k, group=None, download=True):
        """(count, pos, negtext) through the fused path, or None when it does not apply here (k > 28, no peer access) or
        asked every rank to take the phased path (a segment or bucket overflowed somewhere).  download=False leaves the
        local tables in HBM (ctx.download() fetches them) and returns (count, local instance count, None)."""
        for _ in range(3):
            need = self.ctx.dist2_plan(chrs, rank, world, k, self.resident)
            if need < 0:
                return None
            if need:
                # (re)allocation of the exported buffers: collective, and only when the input shape outgrows them
                dist.barrier(group)
                self.ctx.dist2_release_peers()
                dist.barrier(group)
                mine = torch.from_numpy(self.ctx.dist2_alloc().view(np.int64).copy()).to(self.device)
                m = _comm(mine, group)
                allh = torch.empty(world * 8, dtype=torch.int64, device=m.device)
                dist.all_gather_into_tensor(allh, m, group=group)
                handles = np.ascontiguousarray(allh.cpu().numpy()).view(np.uint8).reshape(world, 64)
                ok = 1
                try:
                    self.ctx.dist2_import(handles)
                except Exception as e:                   # noqa: BLE001 -- reported once, the phased paths take over
                    print("sibelia_b200.distributed: peer mapping unavailable (%s); using the phased exchange" % e, flush=True)
                    ok = 0
                flag = _comm(torch.tensor([ok], dtype=torch.int64, device=self.device), group)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                if not int(flag.cpu()[0]):
                    self.ctx.dist2_release_peers()
                    type(self).fused = False
                    return None
            if k > 32:
                out = self._fused_fp(chrs, group, download)
                if out != "regrow":
                    return out
                continue
            status, count, ninst = self.ctx.dist2_run(chrs, self.resident)
            if status == 0:
                if not download:
                    return count, ninst, None
                pos, negtext = self.ctx.download()
                return count, pos, negtext
            if status == 1:
                return None
        raise RuntimeError("sibelia_b200.distributed: the vertex key regions kept overflowing")

    def _rep_tensor(self, ptr, n):
        """the library's array of class representatives as a tensor (zero copy)"""
        return torch.as_tensor(_DeviceArray(ptr, n), device=self.device)

    def _fused_fp(self, chrs, group, download):
        """k > 32: the two halves of the fingerprint step with the min-reduction of the class representatives between
        them and the max-reduction of the verification flag behind them (include/sibgpu.h, sibgpu_fused_run_fp)."""
        for attempt in range(3):
            status, ncls, rep_ptr = self.ctx.dist2_run_fp(chrs, self.resident, attempt)
            if status == 2:
                return "regrow"
            if status == 1:
                return None
            if ncls:
                rep = self._rep_tensor(rep_ptr, ncls)
                if dist.get_backend(group) == "nccl":
                    dist.all_reduce(rep, op=dist.ReduceOp.MIN, group=group)
                else:
                    h = rep.cpu()
                    dist.all_reduce(h, op=dist.ReduceOp.MIN, group=group)
                    rep.copy_(h)
                if rep.is_cuda:
                    torch.cuda.synchronize(self.device)
            count, ninst, collision = self.ctx.dist2_finish_fp()
            flag = _comm(torch.tensor([collision], dtype=torch.int64, device=self.device), group)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            if int(flag.cpu()[0]):
                continue                                 # a fingerprint collision inside a vertex class: other hash bases
            if not download:
                return count, ninst, None
            pos, negtext = self.ctx.download()
            return count, pos, negtext
        raise RuntimeError("sibelia_b200.distributed: fingerprint verification failed three times")

    # -- peer strategy
    peer = os.environ.get("SIBGPU_DIST_PEER", "1") != "0"

    def scatter_local(self, k):
        if getattr(self, "_pending", None) is not None:
            chrs, rank, world = self._pending
            self._pending = None
            out = self.ctx.dist_upload_scatter(chrs, rank, world, k)
        else:
            out = self.ctx.dist_scatter_local(k)
        self.words = self.ctx.dist_record_bytes() // 8
        return out

    def export_send(self):
        return self.ctx.dist_export_send()

    def import_peers(self, handles):
        """Maps the other ranks' send buffers (CUDA IPC); False if this rank cannot (no peer access, IPC unavailable)."""
        try:
            self.ctx.dist_import_peers(handles)
            return True
        except Exception as e:                       # noqa: BLE001 -- reported once, the staged exchange takes over
            if not getattr(self, "_warned", False):
                print("sibelia_b200.distributed: peer mapping unavailable (%s); using the staged all-to-all" % e, flush=True)
                self._warned = True
            return False

    def group_peer(self, counts, seg_caps):
        n = self.ctx.dist_group_peer(counts, seg_caps)
        keys = torch.empty(max(n * self.words, 1), dtype=torch.int64, device=self.device)
        self.ctx.dist_keys(keys.data_ptr())
        return keys[:n * self.words]

    # -- staged strategy
    def scan(self, k):
        self._upload_now()
        nparts, hist, self.nrec = self.ctx.dist_scan(k)
        self.words = self.ctx.dist_record_bytes() // 8
        return nparts, hist

    def scatter(self):
        send = torch.empty(max(self.nrec * self.words, 1), dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        self.ctx.dist_scatter(send.data_ptr())
        return send[:self.nrec * self.words]

    def group(self, recv, counts):
        torch.cuda.synchronize(self.device)
        recv = recv.contiguous()
        n = self.ctx.dist_group(recv.data_ptr() if recv.numel() else 0, counts)
        keys = torch.empty(max(n * self.words, 1), dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        self.ctx.dist_keys(keys.data_ptr())
        return keys[:n * self.words]

    def finish(self, allkeys):
        torch.cuda.synchronize(self.device)
        allkeys = allkeys.contiguous()
        count, ninst = self.ctx.dist_finish(allkeys.data_ptr() if allkeys.numel() else 0, allkeys.numel() // self.words)
        pos, negtext = self.ctx.download()
        return count, pos, negtext


class _DeviceArray:
    """n int64 words of library-owned device memory, as torch.as_tensor takes it (zero copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def _comm(t, group):
    """Tensor as the process group wants it: device tensors for NCCL, host tensors for gloo."""
    return t if dist.get_backend(group) == "nccl" else t.cpu()


def enumerate_sharded(shard, chrs, k, group=None, download=True):
    """Runs the sharded enumeration on every rank of `group`; returns (global vertex count, this rank's positive table,
    this rank's negative table in TEXT order).  `shard` is a backend with upload/scan/scatter/group/finish.
    download=False (fused strategy only): the tables stay in HBM, the second item is the local instance count."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    marks = []

    def lap(what):
        if _TRACE:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            marks.append((what, time.perf_counter()))
    lap("start")
    if getattr(shard, "fused", False) and world <= 16:
        out = shard.fused_enumerate(chrs, rank, world, k, group, download)
        if out is not None:
            lap("fused step")
            shard.last_strategy = ("fused (device-side step counters, TMA peer pulls in k_split, key pull kernel; no collective)"
                                   if k <= 32 else
                                   "fused (k > 32: packed text, k-mer records and keys pulled from the peers' buffers; one "
                                   "all-reduce of the class representatives)")
            return out
    if k > 32:
        # no phased exchange exists for fingerprint classes: a segment or bucket overflowed (a k-mer repeated hundreds of
        # times) or peer mappings are unavailable -- every rank indexes the whole input and keeps its own text range
        out = _replicated(shard, chrs, k, rank, world)
        shard.last_strategy = "replicated (k > 32 fallback: every rank runs the single-GPU enumeration, keeps its range)"
        return out
    shard.upload(chrs, rank, world)
    lap("upload")
    keys = None
    if getattr(shard, "peer", False) and world <= 16:
        # --- peer strategy: scatter into the own send buffer, swap counts + IPC handles, read the peers' segments
        nparts, cnt, cap, ovf = shard.scatter_local(k)
        lap("pack+scatter")
        handle = shard.export_send()
        mine = np.concatenate([cnt.astype(np.int64), np.array([cap, int(ovf)], dtype=np.int64), handle.view(np.int64)])
        dev = shard.device
        m = _comm(torch.from_numpy(mine).to(dev), group)
        allm = torch.empty(world * len(mine), dtype=torch.int64, device=m.device)
        dist.all_gather_into_tensor(allm, m, group=group)
        allm = allm.cpu().numpy().reshape(world, len(mine))
        lap("allgather counts+handles")
        if not allm[:, nparts + 1].any():
            handles = np.ascontiguousarray(allm[:, nparts + 2:]).view(np.uint8).reshape(world, 64)
            ok = shard.import_peers(handles)
            if getattr(shard, "_peer_handles", None) is None or not np.array_equal(shard._peer_handles, handles):
                # new mappings: every rank must have succeeded, or all take the staged path together (one tiny
                # collective, only when a send buffer was (re)allocated -- normally the first call)
                flag = _comm(torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev), group)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                ok = bool(int(flag.cpu()[0]))
                shard._peer_handles = handles.copy() if ok else None
                if not ok:
                    shard.peer = False
            if ok:
                keys = shard.group_peer(allm[:, :nparts].astype(np.uint64), allm[:, nparts].astype(np.uint64))
                lap("group (peer reads)")
                words = shard.words
                shard.last_strategy = "peer (counts all-gather, peer reads in the insert kernel, key all-gather)"
    if keys is None:
        keys, words, dev = _staged_exchange(shard, k, rank, world, group, lap)
        shard.last_strategy = "staged (histogram, one all-to-all of the records, key all-gather)"
    # --- vertex keys of all ranks (variable sizes: pad to the maximum)
    nk = _comm(torch.tensor([keys.numel()], dtype=torch.int64, device=dev), group)
    all_nk = torch.empty(world, dtype=torch.int64, device=nk.device)
    dist.all_gather_into_tensor(all_nk, nk, group=group)
    all_nk = [int(x) for x in all_nk.cpu()]
    mx = max(max(all_nk), 1)
    padded = torch.zeros(mx, dtype=torch.int64, device=dev)
    padded[:keys.numel()] = keys
    p_c = _comm(padded, group)
    gathered = torch.empty(world * mx, dtype=torch.int64, device=p_c.device)
    dist.all_gather_into_tensor(gathered, p_c, group=group)
    allkeys = torch.cat([gathered[s * mx:s * mx + all_nk[s]] for s in range(world)]).to(dev)
    lap("allgather keys")
    out = shard.finish(allkeys)
    lap("finish")
    if _TRACE and rank == 0:
        print("[sharded] " + "  ".join("%s %.2f ms" % (marks[i][0], (marks[i][1] - marks[i - 1][1]) * 1e3)
                                       for i in range(1, len(marks))), flush=True)
    return out


def _replicated(shard, chrs, k, rank, world):
    """(count, pos rows of the own tile range, neg rows of the own tile range in text order) from a full enumeration"""
    count, pos, neg = shard.ctx.enumerate(chrs, k)
    lens = np.array([len(c) for c in chrs], dtype=np.int64)
    start = 1 + np.concatenate([[0], np.cumsum(lens + 1)])[:-1] if len(chrs) else np.zeros(0, dtype=np.int64)
    M = int(lens.sum()) + len(chrs) + 1
    ntiles = (M + 4095) // 4096
    lo, hi = ntiles * rank // world * 4096, ntiles * (rank + 1) // world * 4096
    tp = start[pos["chr"]] + pos["pos"] if len(pos) else np.zeros(0, dtype=np.int64)
    tn = start[neg["chr"]] + (lens[neg["chr"]] - neg["pos"] - k) if len(neg) else np.zeros(0, dtype=np.int64)
    keep_n = np.flatnonzero((tn >= lo) & (tn < hi))
    keep_n = keep_n[np.argsort(tn[keep_n], kind="stable")]
    return count, pos[(tp >= lo) & (tp < hi)], neg[keep_n]


def _staged_exchange(shard, k, rank, world, group, lap):
    """histogram -> exactly sized send buffer -> ONE all-to-all of the records -> grouping of the received records"""
    nparts, hist = shard.scan(k)
    lap("scan")
    words = shard.words
    send = shard.scatter()
    lap("scatter")
    dev = send.device
    # --- partition counts of every rank
    h = _comm(torch.from_numpy(hist.astype(np.int64)).to(dev), group)
    allh = torch.empty(world * nparts, dtype=torch.int64, device=h.device)
    dist.all_gather_into_tensor(allh, h, group=group)
    counts = allh.cpu().numpy().reshape(world, nparts)
    pl = nparts // world
    in_splits = [int(counts[rank, r * pl:(r + 1) * pl].sum()) * words for r in range(world)]
    out_splits = [int(counts[s, rank * pl:(rank + 1) * pl].sum()) * words for s in range(world)]
    lap("allgather counts")
    # --- THE exchange: every record goes to the rank that owns its hash partition
    s_c = _comm(send, group)
    recv = torch.empty(sum(out_splits), dtype=torch.int64, device=s_c.device)
    dist.all_to_all_single(recv, s_c, out_splits, in_splits, group=group)
    lap("all_to_all records")
    keys = shard.group(recv.to(dev), counts.astype(np.uint32))
    lap("group")
    return keys, words, dev


def assemble_tables(pos_parts, negtext_parts):
    """Ranks in order -> the reference's two tables: positive = concatenation (text order = (chr,pos) order);
    negative = sorted by (chr, position in the reverse complement) (vertexenumeration.cpp:361-362)."""
    pos = np.concatenate(pos_parts) if pos_parts else np.zeros(0)
    neg = np.concatenate(negtext_parts) if negtext_parts else np.zeros(0)
    order = np.lexsort((neg["pos"], neg["chr"]))
    return pos, neg[order]


def gather_tables(count, pos_part, negtext_part, group=None):
    """Collects the per-rank tables on rank 0 (returns (count, pos, neg) there, (count, None, None) elsewhere)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    objs = [None] * world if rank == 0 else None
    dist.gather_object((pos_part, negtext_part), objs, dst=0, group=group)
    if rank != 0:
        return count, None, None
    pos, neg = assemble_tables([o[0] for o in objs], [o[1] for o in objs])
    return count, pos, neg


def gather_tables_device(count, pos_part, negtext_part, group=None):
    """gather_tables for large tables under NCCL: the parts travel as byte tensors over NVLink (point to point to rank 0)
    instead of pickled objects; the negative table is assembled per chromosome (descending text order inside each)."""
    from .binding import INST_DTYPE
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if dist.get_backend(group) != "nccl":
        return gather_tables(count, pos_part, negtext_part, group)
    dev = torch.device("cuda", torch.cuda.current_device())
    n = torch.tensor([len(pos_part)], dtype=torch.int64, device=dev)
    alln = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(alln, n, group=group)
    alln = [int(x) for x in alln.cpu()]

    def as_bytes(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)
    if rank != 0:
        for part in (pos_part, negtext_part):
            if len(part):
                dist.send(as_bytes(part), dst=0, group=group)
        return count, None, None
    pos_parts, neg_parts = [pos_part], [negtext_part]
    for src in range(1, world):
        for parts in (pos_parts, neg_parts):
            if alln[src]:
                buf = torch.empty(alln[src] * INST_DTYPE.itemsize, dtype=torch.uint8, device=dev)
                dist.recv(buf, src=src, group=group)
                parts.append(buf.cpu().numpy().view(INST_DTYPE))
            else:
                parts.append(np.zeros(0, dtype=INST_DTYPE))
    pos = np.concatenate(pos_parts)
    negtext = np.concatenate(neg_parts)
    # text order is ascending (chr, text position): reverse every chromosome's run
    bounds = np.flatnonzero(np.diff(negtext["chr"].astype(np.int64))) + 1
    edges = np.concatenate([[0], bounds, [len(negtext)]])
    neg = np.concatenate([negtext[a:b][::-1] for a, b in zip(edges[:-1], edges[1:])]) if len(negtext) else negtext
    return count, pos, neg
