#!/bin/bash
# persistent attention forward (+ TMA-store epilogue) and backward (+ L2 prefetch of the next item): parity, A/B timing, traces
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k attention 2>&1 | tail -5 > gpurun_out/k_pytest.txt
cat gpurun_out/k_pytest.txt
if grep -q passed gpurun_out/k_pytest.txt && ! grep -q failed gpurun_out/k_pytest.txt; then
  timeout 900 python -m pytest tests/test_model_gpu.py tests/test_decode_gpu.py -x -q -m gpu 2>&1 | tail -3
  echo "== persistent fwd + bwd" > gpurun_out/k_attn_times5.txt
  timeout 200 python tools/time_attention.py >> gpurun_out/k_attn_times5.txt 2>&1
  echo "== OASR_FWD_PERSISTENT=0 OASR_BWD_PERSISTENT=0" >> gpurun_out/k_attn_times5.txt
  OASR_FWD_PERSISTENT=0 OASR_BWD_PERSISTENT=0 timeout 200 python tools/time_attention.py >> gpurun_out/k_attn_times5.txt 2>&1
  cat gpurun_out/k_attn_times5.txt
  OASR_B200_LIB=olmoasr_b200/csrc/_ab/attn_trace.so timeout 200 python tools/trace_attention.py > gpurun_out/k_trace5.txt 2>&1
  echo "trace rc=$?"
fi
