#!/bin/bash
# Run on the GPU box (one GPU): launch list + full ncu captures of the hot kernels (warm caches: --cache-control none).
# Outputs land in gpurun_out/; tools/make_profile_summaries.py (run in the authoring container) turns them into the
# tracked summaries under profiles/.
set -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1_launches.csv $B > gpurun_out/r1_launches.log 2>&1
# one warm launch of each scan-type kernel (the first 3 enumerations are warm-up: skip their 4 launches each)
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"^k_pack|^k_scatter|^k_mark|^k_emit" -s 12 -c 4 -o gpurun_out/r1_scan $B > gpurun_out/r1_scan.log 2>&1
# two warm launches of the grouping kernels
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_insert_compact|k_table_scan_compact" -s 600 -c 4 -o gpurun_out/r1_group $B > gpurun_out/r1_group.log 2>&1
# sharded grouping kernel (TMA ring over CUDA IPC mappings): two processes on this GPU, peers read each other's buffers
SIBGPU_STREAMS=1 EXP_MBASES=50 ncu --target-processes all --set full --clock-control none --cache-control none --import-source on -k regex:k_insert_seg -s 100 -c 2 -o gpurun_out/r1_seg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/exp_peer.py > gpurun_out/r1_seg.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_bulge_detect|k_list_edges" -c 3 -o gpurun_out/r1_simplify python tools/simplify_bench.py 4 1e6 0.002 --noref > gpurun_out/r1_simplify.log 2>&1
ls -la gpurun_out
