#!/bin/bash
# Round-2 GPU call A: parity suite (per file, separate processes), bench with gpu_baseline, attention poly A/B, step profile.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
for f in test_kernels_gpu test_optimizer_gpu test_model_gpu test_decode_gpu test_decode_engine_gpu; do
  echo "=== $f" >> gpurun_out/a_tests.log
  timeout 900 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -70 >> gpurun_out/a_tests.log
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -5 gpurun_out/a_bench.err
timeout 900 python bench.py --metric rtf --steps 2 --warmup 1 > gpurun_out/a_rtf.json 2> gpurun_out/a_rtf.err
tail -5 gpurun_out/a_rtf.err
for v in default attn_poly1 attn_poly2; do
  if [ $v != default ]; then export OASR_B200_LIB=$PWD/olmoasr_b200/csrc/_ab/$v.so; else unset OASR_B200_LIB; fi
  echo "== $v" >> gpurun_out/a_attn_ab.txt
  timeout 200 python tools/time_attention.py >> gpurun_out/a_attn_ab.txt 2>&1
done
unset OASR_B200_LIB
timeout 300 python tools/profile_step.py --serial > gpurun_out/a_profile_serial.txt 2>&1
timeout 300 python tools/profile_step.py > gpurun_out/a_profile.txt 2>&1
cat gpurun_out/a_tests.log | grep -E "passed|failed|error|===" 
head -c 3000 gpurun_out/a_bench.json
head -c 3000 gpurun_out/a_rtf.json
