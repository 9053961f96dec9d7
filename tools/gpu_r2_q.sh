#!/bin/bash
# tensor-map cache: parity (kernels + model), host enqueue time per step with the cache on / off
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/q_pytest.txt
cat gpurun_out/q_pytest.txt
echo "== cache on" > gpurun_out/q_host.txt
timeout 300 python tools/profile_step.py --serial 2>&1 | grep -E "host enqueue|sum of kernel" >> gpurun_out/q_host.txt
echo "== OASR_TMAP_CACHE=0" >> gpurun_out/q_host.txt
OASR_TMAP_CACHE=0 timeout 300 python tools/profile_step.py --serial 2>&1 | grep -E "host enqueue|sum of kernel" >> gpurun_out/q_host.txt
cat gpurun_out/q_host.txt
