#!/usr/bin/env python
"""bench.py -- Mbases/s indexed (k=25) on B200, next to the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mbases 100] [--k 25]

A "step" is one construction of the de Bruijn-graph index of one synthetic genome: sibgpu_enumerate_resident
(pack -> scan/histogram -> scatter -> L2-resident hash grouping -> vertex ranking -> instance tables), the GPU
replacement of IndexedSequence's EnumerateBifurcationsSArrayInRAM (/root/reference/src/vertexenumeration.cpp:263-364).
Workload at N=1 is BASELINE.json configs[1]: 100 MB random-ACGT single contig, numpy default_rng(12345), k=25.
With N>1 ranks (torchrun, one process per GPU) the workload is ONE genome of N such contigs (seed 12345+c), sharded by
contiguous text range over the ranks (weak scaling: 100 Mbases per GPU): every rank scatters its k-mer records, bucketed
by hash prefix, into its own send buffer; one small NCCL all-gather swaps the bucket counts; the owner of a bucket
reads the bucket's segments straight out of the peers' send buffers over NVLink inside its grouping kernel (TMA bulk
copies into a shared-memory ring -- the all-to-all is fused into the kernel); all-gather of the vertex keys; local
instance tables (sibelia_b200/distributed.py); `value` = total bases / max-over-ranks step time.

One JSON line is printed by rank 0 (see the task contract): value = device-resident throughput, e2e = the same
metric through sibgpu_enumerate with pinned HOST buffers (H2D + D2H inside the timed region), roofline = dominant
kernel vs the measured HBM peak, cpu_baseline = the unmodified reference (oracle/_ref) timed on this box's host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mbases/s indexed (k=25)"
UNIT = "Mbases/s"
REF_SAMPLE_BASES = 8_000_000


def genome(mbases, seed):
    from sibelia_b200 import synth
    return synth.random_genome(int(mbases * 1_000_000), seed)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock/throttle sampling during the timed region (the profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref = unmodified /root/reference sources compiled
    by oracle/Makefile), single thread (the reference has no threading), on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsibelia_ref.so is not built"}))
        return
    sample = min(REF_SAMPLE_BASES, int(args.mbases * 1_000_000))
    g = genome(args.mbases, 12345)[:sample]
    for _ in range(args.warmup_ref):
        ref.index([g], args.k, dump=False)
    t = []
    for _ in range(args.steps):
        t.append(ref.index([g], args.k, dump=False)["seconds"])
    sec = float(np.mean(t))
    v = sample / 1e6 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup_ref, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "synthetic %g MB random-ACGT single contig, k=%d (BASELINE configs[1])" % (args.mbases, args.k),
                   "k": args.k},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                         "sample": "first %d bases of the workload genome, IndexedSequence ctor (in-RAM SA path), "
                                   "libdivsufsort 32-bit, 1 thread (the reference is single-threaded); %d host cores present"
                                   % (sample, os.cpu_count())},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mbases", type=float, default=100.0)
    ap.add_argument("--k", type=int, default=25)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup_ref = min(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    args.warmup = max(args.warmup, 3)
    import torch
    import torch.distributed as dist
    import sibelia_b200 as sb

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N = int(args.mbases * 1_000_000)
    ctx = sb.Context(local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def l2_flush():
        flush.fill_(1)
        torch.cuda.synchronize()

    kstats = {}

    def add_stats():
        for s in ctx.kernel_stats():
            a = kstats.setdefault(s["name"], {"launches": 0, "ms": 0.0, "algo_bytes": 0})
            a["launches"] += s["launches"]
            a["ms"] += s["ms"]
            a["algo_bytes"] += s["algo_bytes"]

    if world == 1:
        g = genome(args.mbases, 12345)
        host = torch.empty(N, dtype=torch.uint8, pin_memory=True)
        host.numpy()[:] = g
        hview = [host.numpy()]
        ctx.upload(hview)

        def step_resident():
            count, ninst = ctx.enumerate_resident(args.k)
            return count, ninst, ctx.last_device_ms()

        def step_e2e():
            c2, pos, neg = ctx.enumerate(hview, args.k)
            return pos.nbytes + neg.nbytes + 64
        h2d = N + 8
    else:
        from sibelia_b200 import distributed as D
        hosts = []
        for c in range(world):
            h = torch.empty(N, dtype=torch.uint8, pin_memory=(c == rank))
            h.numpy()[:] = genome(args.mbases, 12345 + c)
            hosts.append(h)
        hview = [h.numpy() for h in hosts]
        g = hview[0]
        shard = D.GpuShard(ctx)

        class Resident(D.GpuShard):
            """the shard's text is already in HBM: skip the upload phase of enumerate_sharded"""
            def upload(self, chrs, rank, world):
                self._pending = None
        resident = Resident(ctx)
        ctx.dist_upload(hview, rank, world)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def step_resident():
            torch.cuda.synchronize()
            ev0.record()
            count, pos, neg = D.enumerate_sharded(resident, hview, args.k)
            torch.cuda.synchronize()
            ev1.record()
            ev1.synchronize()
            return count, len(pos), ev0.elapsed_time(ev1)

        def step_e2e():
            count, pos, neg = D.enumerate_sharded(shard, hview, args.k)
            return pos.nbytes + neg.nbytes + 64
        h2d = N + 8

    # ---- device-resident arm: inputs already in HBM
    ctx.set_profiling(True)
    for _ in range(args.warmup):
        count, ninst, _ms = step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, launches = 0.0, 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        l2_flush()
        if world > 1:
            dist.barrier()
        count, ninst, ms = step_resident()
        dev_ms += ms
        launches += ctx.last_launches()
        add_stats()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    ctx.set_profiling(False)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * N / 1e6 / (ms_per_step / 1e3)

    # ---- end-to-end arm: pinned host buffers in, host tables out, through the public API
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        d2h = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N / 1e6 / float(te.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events on the library's stream, live over the timed steps)
    peak, peak_src = peaks()
    dom = max(kstats.items(), key=lambda kv: kv[1]["ms"])
    dname, d = dom
    achieved = d["algo_bytes"] / 1e9 / (d["ms"] / 1e3) if d["ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8TBps": achieved / 8000.0, "traffic": None, "peak_source": peak_src,
                "launches_per_step": d["launches"] / args.steps, "ms_per_step": d["ms"] / args.steps,
                "share_of_step": d["ms"] / dev_ms if dev_ms else None,
                "kernels": {n: {"ms_per_step": s["ms"] / args.steps, "launches_per_step": s["launches"] / args.steps,
                                "algo_GBps": (s["algo_bytes"] / 1e9 / (s["ms"] / 1e3)) if s["ms"] > 0 else None}
                            for n, s in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])}}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
        # (profiles/r1_ncu_*.txt); for the overlapped insert+scan phase: both kernels, all partitions of one step
        try:
            tj = json.load(open(tr))
            if dname in tj:
                roofline["traffic"] = tj[dname]
            elif dname == "k_insert+k_table_scan":
                per_launch_records = 1048576.0      # the captures were taken with the default 1 Mi-record partitions
                nrec = d["algo_bytes"] / 8.0 / args.steps
                roofline["traffic"] = (tj["k_insert"] + tj["k_table_scan"]) * nrec / per_launch_records
                roofline["traffic_note"] = ("whole phase per step = per-launch dram bytes of the two kernels (ncu --set full, warm L2: the "
                                            "table and most of the freshly scattered records are L2-resident) x partitions per step")
        except Exception:
            pass
    roofline["note"] = ("k_insert is bound by scattered L2 atomics, not by HBM: tools/ubench/atomics.cu measures 90-104 G CAS/s "
                        "on this part for an L2-resident table; the phase issues ~1.5 CAS per record (linear probing at load 0.5)")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref
        if ref.available():
            sample = min(REF_SAMPLE_BASES * 2, N)
            sec = ref.index([g[:sample]], args.k, dump=False)["seconds"]
            cpu = {"value": sample / 1e6 / sec, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": "first %d bases of the workload genome, unmodified reference IndexedSequence ctor "
                             "(oracle/_ref, libdivsufsort in-RAM path), 1 thread of %d host cores, %.1f s"
                             % (sample, os.cpu_count(), sec)}
        else:
            from oracle import restate
            sample = min(2_000_000, N)
            t0 = time.perf_counter()
            restate.enumerate_bifurcations([g[:sample]], args.k)
            sec = time.perf_counter() - t0
            cpu = {"value": sample / 1e6 / sec, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "first %d bases, oracle/enum_restate.c (sort-based restatement), %.1f s" % (sample, sec)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": ("synthetic %g MB random-ACGT single contig, numpy default_rng(12345), k=%d (BASELINE configs[1])"
                                % (args.mbases, args.k)) if world == 1 else
                               ("one synthetic genome of %d random-ACGT contigs x %g MB (default_rng(12345+c)), k=%d, sharded by "
                                "text range over %d GPUs; k-mer records exchanged by peer reads over NVLink fused into the "
                                "grouping kernel (NCCL only for the bucket counts and the vertex keys)" % (world, args.mbases, args.k, world)),
                   "k": args.k, "bases_per_gpu": N, "vertices": int(count), "instances_per_strand": int(ninst),
                   "l2": "256 MB L2 flush between timed iterations (outside the event-timed region)",
                   "timing": ("CUDA events on the library stream around each whole step" if world == 1 else
                              "CUDA events around each whole sharded step (device-synchronised on both sides)")
                             + "; wall %.1f ms/step incl. flush" % (wall / args.steps * 1e3)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": float(te.item()) * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
