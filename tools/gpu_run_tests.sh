#!/bin/bash
# Run the GPU test groups in separate processes (a sticky CUDA error in one group must not mask the others).
mkdir -p gpurun_out
rc=0
for sel in "test_gemm" "test_layernorm" "test_attention" "test_cross_entropy or test_embedding" "test_conv or test_small" "test_logmel"; do
  echo "=== $sel"
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$sel" 2>&1 | tail -25
  [ ${PIPESTATUS[0]} -ne 0 ] && rc=1
done
exit $rc
