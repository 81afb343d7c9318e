#!/usr/bin/env bash
# Builds libdcvgan_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
OUT=../libdcvgan_b200.so
mkdir -p build
pids=()
for f in api simt_conv elementwise misc tc_conv; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ conv_geom.cuh -nt build/$f.o ] || [ ../../include/dcvgan_b200.h -nt build/$f.o ]; then
    ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/api.o build/simt_conv.o build/elementwise.o build/misc.o build/tc_conv.o -cudart static
echo "built $OUT"
