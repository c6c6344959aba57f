#!/bin/bash
# Run on the GPU box: launch list + full ncu captures of the hot kernels (warm caches: --cache-control none).
set -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r1_launches.csv $B > gpurun_out/r1_launches.log 2>&1
for kern in k_insert_compact k_scatter k_scan_hist k_pack k_table_scan_compact k_mark; do
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:$kern -s 6 -c 2 -o gpurun_out/r1_$kern $B > gpurun_out/r1_$kern.log 2>&1
done
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_bulge_detect -c 2 -o gpurun_out/r1_k_bulge_detect python tools/simplify_bench.py 4 1e6 0.002 --noref > gpurun_out/r1_k_bulge_detect.log 2>&1
ls -la gpurun_out
