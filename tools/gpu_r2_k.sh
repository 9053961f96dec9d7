#!/bin/bash
# attention backward with the TMA-store dK/dV epilogue: parity (kernel + model level), timing, per-phase trace
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/k_pytest.txt
cat gpurun_out/k_pytest.txt
timeout 200 python tools/time_attention.py > gpurun_out/k_attn_times3.txt 2>&1
cat gpurun_out/k_attn_times3.txt
OASR_B200_LIB=olmoasr_b200/csrc/_ab/attn_trace.so timeout 200 python tools/trace_attention.py > gpurun_out/k_trace3.txt 2>&1
echo "trace rc=$?"
grep -n "==== backward encoder" -A400 gpurun_out/k_trace3.txt | grep -E "role|last dQ|dK,dV|start|landed|S,dP\(0\)" | head -40
