#!/bin/bash
# attention forward with P in TMEM (TMEM-A MMA): parity + timing for the variants
set -u
mkdir -p gpurun_out
: > gpurun_out/l_attn.txt
for v in fwd_ptmem fwd_ptmem_poly1 fwd_ptmem_poly2; do
  echo "== $v" >> gpurun_out/l_attn.txt
  OASR_B200_LIB=olmoasr_b200/csrc/_ab/$v.so timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -4 >> gpurun_out/l_attn.txt
  OASR_B200_LIB=olmoasr_b200/csrc/_ab/$v.so timeout 300 python tools/time_attention.py >> gpurun_out/l_attn.txt 2>&1
done
echo "== default" >> gpurun_out/l_attn.txt
timeout 300 python tools/time_attention.py >> gpurun_out/l_attn.txt 2>&1
cat gpurun_out/l_attn.txt
