#!/bin/bash
# attention backward with the dQ-drain warpgroup: parity, timing table (new / old structure), per-phase trace of one CTA
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -5 > gpurun_out/k_pytest.txt
cat gpurun_out/k_pytest.txt
echo "== drain warpgroup (default)" > gpurun_out/k_attn_times.txt
timeout 300 python tools/time_attention.py >> gpurun_out/k_attn_times.txt 2>&1
echo "== without (-DOASR_BWD_DRAIN_WG=0)" >> gpurun_out/k_attn_times.txt
OASR_B200_LIB=olmoasr_b200/csrc/_ab/bwd_nodrain.so timeout 300 python tools/time_attention.py >> gpurun_out/k_attn_times.txt 2>&1
cat gpurun_out/k_attn_times.txt
OASR_B200_LIB=olmoasr_b200/csrc/_ab/attn_trace.so timeout 300 python tools/trace_attention.py > gpurun_out/k_trace.txt 2>&1
echo "trace rc=$?"
head -60 gpurun_out/k_trace.txt
