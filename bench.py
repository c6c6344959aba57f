#!/usr/bin/env python
"""bench.py -- Mbases/s indexed (k=25) on B200, next to the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mbases 100] [--k 25]

A "step" is one construction of the de Bruijn-graph index of one synthetic genome: sibgpu_enumerate_resident
(pack -> scan + hash partition -> bucket split -> shared-memory grouping -> vertex ranking -> instance tables), the GPU
replacement of IndexedSequence's EnumerateBifurcationsSArrayInRAM (/root/reference/src/vertexenumeration.cpp:263-364).
Workload at N=1 is BASELINE.json configs[1]: 100 MB random-ACGT single contig, numpy default_rng(12345), k=25.
With N>1 ranks (torchrun, one process per GPU) the headline workload is its N-fold -- ONE random-ACGT genome of N contigs
x 100 MB sharded by contiguous text range over the ranks (weak scaling: 100 Mbases per GPU): every rank scatters its
k-mer records, bucketed by hash prefix, into its own exported buffer and publishes a step counter; the owner of a
bucket waits for the counters on the device and pulls the bucket's segments straight out of the peers' buffers over
NVLink with TMA bulk copies inside its split kernel (the all-to-all is fused into the kernel), groups them in shared
memory, publishes its vertex keys the same way; a pull kernel concatenates all ranks' keys (the all-gather, fused);
local instance tables (sibelia_b200/distributed.py).  `value` = total bases / max-over-ranks step time.
`c4` (N>1) = the same measurement on the SURVEY 8(d) strain recipe with N strains of 125 MB (N = 8: BASELINE
configs[3], 10^9 bases; millions of vertices, so the id ranking and the instance tables count), with
`same_genome_on_1_gpu`: that genome indexed by rank 0 alone (speed-up of N GPUs on it, and the check that N GPUs produced
the same tables).  `result_digest` = sha256 over (vertex count, positive table, negative table) assembled on rank 0
outside the timed region.

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
CPU_SAMPLE_BASES = 16_000_000
STRAIN_BASES = 125_000_000


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


def workload(world, mbases):
    """The chromosomes of the N-GPU headline workload and its description: N random contigs of `mbases` MB, i.e. the
    N-fold of the N=1 workload (weak scaling; north_star: "throughput on a synthetic random-ACGT genome ... at 1, 2, 4 and
    8 GPUs")."""
    if world == 1:
        return [genome(mbases, 12345)], ("synthetic %g MB random-ACGT single contig, numpy default_rng(12345) "
                                         "(BASELINE configs[1])" % mbases)
    from sibelia_b200 import synth
    return [synth.random_genome(int(mbases * 1_000_000), 12345 + c) for c in range(world)], (
        "synthetic random-ACGT genome of %d contigs x %g MB (default_rng(12345 + c); contig 0 = BASELINE configs[1]), one "
        "genome sharded by text range over %d GPUs" % (world, mbases, world))


def strain_workload(world):
    from sibelia_b200 import synth
    return synth.strains(world, STRAIN_BASES), (
        "SURVEY 8(d) strain recipe, %d strains x 125 MB (base default_rng(1000), strain s default_rng(2000+s): p_sub 0.002, "
        "4 x 200 kb inversions, indels)%s, one genome sharded by text range over %d GPUs"
        % (world, " = BASELINE configs[3]" if world == 8 else "", world))


def result_digest(count, pos, neg):
    import hashlib
    h = hashlib.sha256()
    h.update(np.uint64(count).tobytes())
    h.update(np.ascontiguousarray(pos).tobytes())
    h.update(np.ascontiguousarray(neg).tobytes())
    return h.hexdigest()


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref = unmodified /root/reference sources compiled
    by oracle/Makefile), single thread (the reference has no threading).  N=1: the FULL 100 MB workload, one step
    (40-120 s; warm-up would only repeat it).  N>1: one rank's shard of the N-contig genome (contig 0, 100 MB): the
    whole set needs tens of GB and ~N x 45 s (and more: suffix sorting is superlinear) per index on the CPU."""
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsibelia_ref.so is not built"}))
        return
    if world == 1:
        g = genome(args.mbases, 12345)
        what = "synthetic %g MB random-ACGT single contig, numpy default_rng(12345), k=%d (BASELINE configs[1]), FULL size" % (args.mbases, args.k)
        sample = "the whole workload genome (%d bases)" % len(g)
    else:
        g = genome(args.mbases, 12345)
        what = ("contig 0 (%g MB, = the N=1 workload) of the %d-contig random-ACGT genome of the GPU arm, k=%d: one rank's "
                "shard -- a bounded sample, the reference gets slower per base with size" % (args.mbases, world, args.k))
        sample = "contig 0 of the %d-contig genome (%d bases)" % (world, len(g))
    steps = max(1, min(args.steps, args.ref_steps))
    t = []
    for _ in range(steps):
        t.append(ref.index([g], args.k, dump=False)["seconds"])
    sec = float(np.mean(t))
    v = len(g) / 1e6 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 0, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": what, "k": args.k,
                   "steps_note": "one index construction takes %.0f s on this host: %d step(s) run, no warm-up (requested "
                                 "--steps %d --warmup %d)" % (sec, steps, args.steps, args.warmup)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                         "sample": "%s, IndexedSequence ctor (in-RAM SA path), libdivsufsort 32-bit, 1 thread (the reference "
                                   "is single-threaded); %d host cores present" % (sample, os.cpu_count())},
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
    ap.add_argument("--ref-steps", type=int, default=1, help="steps the reference arm actually runs (each is 40-170 s)")
    ap.add_argument("--no-c4", action="store_true", help="N>1: skip the strain-set workload (BASELINE configs[3] at N = 8)")
    ap.add_argument("--no-single", action="store_true", help="N>1: skip the one-GPU run of the strain set on rank 0")
    args = ap.parse_args()
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

    ctx = sb.Context(local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def l2_flush():
        flush.fill_(1)
        torch.cuda.synchronize()

    def pinned(a):
        t = torch.empty(len(a), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = a
        return t

    def measure(chrs, profile):
        """Device-resident and end-to-end arms on one workload (list of uint8 arrays, the same on every rank)."""
        total = int(sum(len(c) for c in chrs))
        kstats = {}

        def add_stats():
            for s in ctx.kernel_stats():
                a = kstats.setdefault(s["name"], {"launches": 0, "ms": 0.0, "algo_bytes": 0})
                a["launches"] += s["launches"]
                a["ms"] += s["ms"]
                a["algo_bytes"] += s["algo_bytes"]

        if world == 1:
            keep = [pinned(c) for c in chrs]
            hview = [t.numpy() for t in keep]
            ctx.upload(hview)

            def step_resident():
                count, ninst = ctx.enumerate_resident(args.k)
                return count, ninst, ctx.last_device_ms()

            def step_e2e():
                c2, pos, neg = ctx.enumerate(hview, args.k)
                return c2, pos, neg
            h2d = total + 8
        else:
            from sibelia_b200 import distributed as D
            # a rank copies only the bytes of its own text range (+ halo): pin the chromosomes that range touches
            starts = np.concatenate([[1], 1 + np.cumsum([len(c) + 1 for c in chrs])])[:-1]
            M = total + len(chrs) + 1
            lo, hi = M * rank // world - 8192, M * (rank + 1) // world + 8192
            keep = [pinned(c) if (s < hi and s + len(c) > lo) else None for c, s in zip(chrs, starts)]
            hview = [t.numpy() if t is not None else c for t, c in zip(keep, chrs)]
            shard = D.GpuShard(ctx)
            resident = D.GpuShard(ctx)
            resident.resident = True                     # fused path: the text range stays in HBM

            def _skip_upload(chrs_, rank_, world_):      # phased paths: likewise
                resident._pending = None
            resident.upload = _skip_upload
            ctx.dist_upload(hview, rank, world)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def step_resident():
                torch.cuda.synchronize()
                ev0.record()
                count, pos, neg = D.enumerate_sharded(resident, hview, args.k, download=False)   # tables stay in HBM (fused path)
                torch.cuda.synchronize()
                ev1.record()
                ev1.synchronize()
                return count, pos if neg is None else len(pos), ev0.elapsed_time(ev1)

            def step_e2e():
                return D.enumerate_sharded(shard, hview, args.k)
            h2d = (hi - lo if world > 1 else total) + 8

        # ---- device-resident arm: inputs already in HBM
        ctx.set_profiling(profile)
        # clocks: nvidia-smi needs ~0.2 s to deliver its first sample and a sharded timed region can be shorter than that,
        # so the sampler runs from the warm-up steps (the same kernels, back to back) to the end of the timed region
        sampler = ClockSampler(local_rank)
        sampler.start()
        for _ in range(args.warmup):
            count, ninst, _ms = step_resident()
        barrier()
        dev_ms, launches = 0.0, 0
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            l2_flush()
            if world > 1:
                dist.barrier()
            count, ninst, ms = step_resident()
            dev_ms += ms
            launches += ctx.last_launches()
            if profile:
                add_stats()
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop()
        ctx.set_profiling(False)
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item()) / args.steps

        # ---- end-to-end arm: pinned host buffers in, host tables out, through the public API
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            c2, pos, neg = step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        d2h = pos.nbytes + neg.nbytes + 64
        pageable_s = None
        if world == 1:
            # the same call from PAGEABLE host memory -- what the C++ facade hands over (std::string::data(),
            # sibelia_b200/csrc/facade/vertexenumeration_gpu.cpp): the driver stages the copy through its own pinned buffer
            for _ in range(2):
                ctx.enumerate(chrs, args.k)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ctx.enumerate(chrs, args.k)
            pageable_s = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        ti = torch.tensor([len(pos)], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(ti, op=dist.ReduceOp.SUM)

        # ---- result digest, outside every timed region: the full tables on rank 0
        if world > 1:
            c2, pos, neg = D.gather_tables_device(c2, pos, neg)
        digest = result_digest(c2, pos, neg) if rank == 0 else None
        return {"total": total, "ms_per_step": ms_per_step, "value": total / 1e6 / (ms_per_step / 1e3), "dev_ms": dev_ms,
                "e2e_ms": float(te.item()) * 1e3, "e2e_value": total / 1e6 / float(te.item()), "h2d": int(h2d), "d2h": int(d2h),
                "pageable_ms": pageable_s * 1e3 if pageable_s else None,
                "launches": int(launches), "count": int(c2), "ninst": int(ti.item()), "clocks": clocks, "wall": wall,
                "kstats": kstats, "digest": digest, "strategy": getattr(shard, "last_strategy", None) if world > 1 else None}

    chrs, what = workload(world, args.mbases)
    r = measure(chrs, True)
    c4 = None
    if world > 1 and not args.no_c4:
        # BASELINE configs[3] (N = 8: 10^9 bases) and its smaller siblings: N strains x 125 MB, with the output check
        schrs, swhat = strain_workload(world)
        a = measure(schrs, False)
        c4 = {"workload": swhat + ", k=%d" % args.k, "value": a["value"], "ms_per_step": a["ms_per_step"], "e2e_value": a["e2e_value"],
              "e2e_ms_per_step": a["e2e_ms"], "vertices": a["count"], "instances_per_strand": a["ninst"], "exchange": a["strategy"],
              "result_digest": a["digest"]}
        # the unmodified reference indexed the 4- and the 8-strain set once in the authoring container
        # (tests/golden/make_golden_scale_index.py: 730 s and 1524 s on one CPU thread); its digests are committed
        gold = os.path.join(ROOT, "tests", "golden", "scale_index_digests.json")
        if rank == 0 and args.k == 25 and os.path.exists(gold):
            for z in json.load(open(gold)).values():
                if z["n_strains"] == world and z["base_len"] == STRAIN_BASES and z["k"] == args.k:
                    c4["reference_digest"] = z["result_digest"]
                    c4["digest_equals_reference"] = z["result_digest"] == a["digest"]
                    c4["reference_seconds_1_cpu_thread"] = z["reference_seconds"]
        # the denominator this figure should be read against: the SAME genome indexed by ONE GPU (rank 0 alone, the others
        # wait), and the proof that N GPUs computed the same tables
        if rank == 0 and not args.no_single:
            ctx.upload(schrs)
            for _ in range(2):
                ctx.enumerate_resident(args.k)
            ms1 = []
            for _ in range(3):
                l2_flush()
                ctx.enumerate_resident(args.k)
                ms1.append(ctx.last_device_ms())
            t1 = float(np.mean(ms1))
            c1, _n1 = ctx.enumerate_resident(args.k)
            p1, n1 = ctx.download()
            d1 = result_digest(c1, p1, n1)               # outside every timed region
            c4["same_genome_on_1_gpu"] = {"ms_per_step": t1, "value": a["total"] / 1e6 / (t1 / 1e3),
                                          "speedup_of_n_gpus": t1 / a["ms_per_step"], "result_digest": d1,
                                          "digest_equals_n_gpu_result": d1 == a["digest"]}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events on the library's stream, live over the timed steps)
    kstats, dev_ms = r["kstats"], r["dev_ms"]
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
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of that kernel on this workload, from the committed
        # `ncu --set full --cache-control all` captures (profiles/r2_ncu_*.txt)
        try:
            tj = json.load(open(tr))
            if dname in tj and world == 1:
                roofline["traffic"] = tj[dname]
        except Exception:
            pass
    whole = sum(s["algo_bytes"] for s in kstats.values()) / 1e9 / (dev_ms / 1e3) if dev_ms else 0.0
    roofline["whole_step_algo_GBps"] = whole
    roofline["note"] = ("every kernel of the step is listed with its algorithmic bytes / its own device time; the scan kernels "
                        "(k_scatter, k_mark) are instruction-issue bound (rolling canonical 2-bit k-mers, mixing, shared-memory "
                        "counting sort), k_split / k_group move each 8-byte record once more")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref
        g = chrs[0]
        if ref.available():
            sample = min(CPU_SAMPLE_BASES, len(g))
            sec = ref.index([g[:sample]], args.k, dump=False)["seconds"]
            cpu = {"value": sample / 1e6 / sec, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": "first %d bases of the workload genome, unmodified reference IndexedSequence ctor "
                             "(oracle/_ref, libdivsufsort in-RAM path), 1 thread of %d host cores, %.1f s; the reference gets "
                             "slower per base with size (suffix sorting): the full-size figure is the --impl reference arm"
                             % (sample, os.cpu_count(), sec)}
        else:
            from oracle import restate
            sample = min(2_000_000, len(g))
            t0 = time.perf_counter()
            restate.enumerate_bifurcations([g[:sample]], args.k)
            sec = time.perf_counter() - t0
            cpu = {"value": sample / 1e6 / sec, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "first %d bases, oracle/enum_restate.c (sort-based restatement), %.1f s" % (sample, sec)}

    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": what + ", k=%d" % args.k, "k": args.k, "bases_total": r["total"], "bases_per_gpu": r["total"] // world,
                   "vertices": r["count"], "instances_per_strand": r["ninst"], "exchange": r["strategy"],
                   "l2": "256 MB L2 flush between timed iterations (outside the event-timed region)",
                   "timing": ("CUDA events on the library stream around each whole step" if world == 1 else
                              "CUDA events around each whole sharded step (device-synchronised on both sides), max over ranks")
                             + "; wall %.1f ms/step incl. flush" % (r["wall"] / args.steps * 1e3)},
        "clocks": r["clocks"],
        "e2e": {"value": r["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                "ms_per_step": r["e2e_ms"], "note": "pinned host buffers -> sibgpu_enumerate / the sharded step -> host tables; "
                                                    "bytes are per rank",
                "pageable_ms_per_step": r["pageable_ms"],
                "pageable_value": (r["total"] / 1e6 / (r["pageable_ms"] / 1e3)) if r["pageable_ms"] else None,
                "pageable_note": "same call from pageable host memory, as the C++ facade of the reference CLI passes it"},
        "gpu_launches": r["launches"],
        "result_digest": r["digest"],
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if c4:
        line["c4"] = c4
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
