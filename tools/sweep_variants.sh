#!/bin/bash
# dev: k_insert_compact variants x table factor x partition size
for v in 0 1 2 3; do for f in 2 4; do for p in 2097152 4194304; do
  echo "variant=$v factor=$f"; SIBGPU_INSERT_VARIANT=$v SIBGPU_TABLE_FACTOR=$f python tools/sweep_parts.py 100 25 ${1:-random} $p | cut -c1-190
done; done; done
