#!/bin/bash
# attention A/B: forward ping-pong of the softmax warpgroups, backward exponentials partly on the FMA pipe
set -u
mkdir -p gpurun_out
: > gpurun_out/l_attn.txt
for v in fwd_pingpong bwd_poly1 bwd_poly2; do
  echo "== $v" >> gpurun_out/l_attn.txt
  OASR_B200_LIB=olmoasr_b200/csrc/_ab/$v.so timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -4 >> gpurun_out/l_attn.txt
  OASR_B200_LIB=olmoasr_b200/csrc/_ab/$v.so timeout 200 python tools/time_attention.py 2>&1 | tail -5 >> gpurun_out/l_attn.txt
done
echo "== default" >> gpurun_out/l_attn.txt
timeout 200 python tools/time_attention.py >> gpurun_out/l_attn.txt 2>&1
cat gpurun_out/l_attn.txt
