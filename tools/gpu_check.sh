#!/usr/bin/env bash
# Runs on the GPU box (via gpurun): op-level parity in two processes so that a hang in the
# tcgen05 kernels cannot take the CUDA-core results down with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -rA -k "not tcgen05 and not channel_slices and not padded" -p no:cacheprovider > gpurun_out/ops_simt.log 2>&1
echo "simt exit $?" >> gpurun_out/ops_simt.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -rA -k "tcgen05 or channel_slices or padded" -p no:cacheprovider --timeout 120 --timeout-method thread > gpurun_out/ops_tc.log 2>&1
echo "tc exit $?" >> gpurun_out/ops_tc.log
tail -5 gpurun_out/ops_simt.log; tail -40 gpurun_out/ops_tc.log
