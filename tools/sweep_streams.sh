#!/bin/bash
for s in ${2:-1 4 6 8}; do for p in ${3:-524288 1048576}; do
  echo "streams=$s"; SIBGPU_STREAMS=$s python tools/sweep_parts.py 100 25 ${1:-random} $p | cut -c1-60
done; done
