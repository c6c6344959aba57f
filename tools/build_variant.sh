#!/bin/bash
# Dev tool: builds sibelia_b200/variants/libsibgpu_<name>.so from the same sources with extra nvcc flags
# (e.g. tools/build_variant.sh u4 -DSIBGPU_MARK_UNROLL=4); run it with SIBGPU_LIB=<that file>.
set -e
name=$1; shift
cd "$(dirname "$0")/../sibelia_b200"
mkdir -p variants/build_$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 20012"
for f in api enumerate fingerprint edges fasta simplify; do
	$NV "$@" -c csrc/$f.cu -o variants/build_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libsibgpu_$name.so variants/build_$name/*.o -cudart shared
rm -rf variants/build_$name
echo variants/libsibgpu_$name.so
