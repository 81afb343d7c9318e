#!/usr/bin/env bash
# Builds libdcvgan_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
#   build.sh        product library
#   build.sh exp    libdcvgan_b200_exp.so with -DDCV_EXPERIMENTS (timing-experiment hooks for tools/exp_*.py; loaded when
#                   DCV_EXPERIMENTS_LIB=1 is set - never by the tests, bench.py or the package by default)
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
OUT=../libdcvgan_b200.so
BUILD=build
if [ "${1:-}" = "exp" ]; then
  FLAGS="$FLAGS -DDCV_EXPERIMENTS"
  OUT=../libdcvgan_b200_exp.so
  BUILD=build_exp
fi
mkdir -p $BUILD
SRCS="api simt_conv elementwise misc tc_conv image_conv"
pids=()
for f in $SRCS; do
  if [ ! -f $BUILD/$f.o ] || [ $f.cu -nt $BUILD/$f.o ] || [ common.cuh -nt $BUILD/$f.o ] || [ conv_geom.cuh -nt $BUILD/$f.o ] || [ ../../include/dcvgan_b200.h -nt $BUILD/$f.o ]; then
    ( $NVCC $FLAGS -c $f.cu -o $BUILD/$f.o > $BUILD/$f.log 2>&1 || { cat $BUILD/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
OBJS=""
for f in $SRCS; do OBJS="$OBJS $BUILD/$f.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $OBJS -cudart static
echo "built $OUT"
