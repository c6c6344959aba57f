#!/bin/bash
# Run on the GPU box (one GPU): launch list + full ncu captures of the round-2 hot kernels with COLD caches
# (--cache-control all: every launch starts with flushed L2, so dram__bytes is the traffic of a kernel whose input is
# not L2-resident -- what the step sees for everything but the smallest kernels).  Outputs land in gpurun_out/;
# tools/make_profile_summaries.py r2 (authoring container) turns them into the tracked summaries under profiles/.
set -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.log 2>&1
# k = 25, 100 MB random contig: one launch of every kernel of the step after the warm-up enumerations
ncu --set full --clock-control none --cache-control all --import-source on -k regex:"^k_pack|^k_scatter|^k_split|^k_group|^k_mark|^k_emit" -s 18 -c 6 -o gpurun_out/r2_step $B > gpurun_out/r2_step.log 2>&1
# k = 100 on 4 x 12.5 Mb strains: the fingerprint path (checkpoints, fused rolling scans, string ranking)
KSWEEP_K=100 ncu --set full --clock-control none --cache-control all --import-source on -k regex:"^k_fp_ckpt|^k_scatter|^k_mark|^k_emit" -s 4 -c 4 -o gpurun_out/r2_fp python tools/k_sweep.py 4 12.5e6 > gpurun_out/r2_fp.log 2>&1
ncu --set full --clock-control none --cache-control all --import-source on -k regex:"k_bulge_detect|k_list_edges" -c 3 -o gpurun_out/r2_simplify python tools/simplify_bench.py 4 1e6 0.002 --noref > gpurun_out/r2_simplify.log 2>&1
ls -la gpurun_out | tail -12
