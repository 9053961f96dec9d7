#!/bin/bash
# Round-2 GPU call C (1 GPU): re-run of the adjusted tests, decode step profile, attention exp2-polynomial A/B, short bench.
mkdir -p gpurun_out
rm -f gpurun_out/c_*.txt gpurun_out/c_*.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "gemm_production or logmel or padding_mask" 2>&1 | tail -15 > gpurun_out/c_tests.log
timeout 600 python -m pytest tests/test_optimizer_gpu.py tests/test_decode_engine_gpu.py tests/test_decode_gpu.py -m gpu -q -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -40 >> gpurun_out/c_tests.log
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "tiny or forward_logits or fused_loss or benchmark_widths" 2>&1 | tail -5 >> gpurun_out/c_tests.log
timeout 400 python tools/profile_decode.py > gpurun_out/c_decode_profile.txt 2>&1
for v in default attn_poly1 attn_poly2; do
  if [ $v != default ]; then export OASR_B200_LIB=$PWD/olmoasr_b200/csrc/_ab/$v.so; else unset OASR_B200_LIB; fi
  echo "== $v" >> gpurun_out/c_attn_ab.txt
  timeout 200 python tools/time_attention.py >> gpurun_out/c_attn_ab.txt 2>&1
done
unset OASR_B200_LIB
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
grep -E "passed|failed|rel-L2|margin|Error" gpurun_out/c_tests.log
cat gpurun_out/c_attn_ab.txt
head -60 gpurun_out/c_decode_profile.txt
python -c "
import json; d=json.load(open('gpurun_out/c_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
