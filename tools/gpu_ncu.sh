#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) --set full captures of the dominant kernels (training and
# decode).  Run under gpurun (one GPU).  Numbers printed by a run under ncu are never bench values.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# (1) every launch of a short bench run with its device time (shares, not absolutes); baselines off: only our kernels matter here
timeout 1500 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/launches_bench.stdout 2>&1
echo "launch list rc=$?"
# (2) full sets on a 2+2-layer medium-width model (same per-layer shapes as the bench)
for k in gemm_tcgen05 attention_fwd attention_bwd layernorm_bwd logmel_frames colsum; do
  timeout 900 $NCU --set full -k regex:$k -s 6 -c 2 -f -o gpurun_out/ncu_$k \
      python tools/profile_step.py --layers 2 --no-profiler > gpurun_out/ncu_$k.stdout 2>&1
  echo "$k rc=$?"
done
# (3) the decode step at 64 sequences (graph replay is profiled kernel by kernel)
for k in dec_linear dec_attn_fused dec_sample; do
  timeout 900 $NCU --set full -k regex:$k -s 40 -c 2 -f -o gpurun_out/ncu_$k \
      python tools/profile_decode.py --batches 64 --pos 8 > gpurun_out/ncu_$k.stdout 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out | grep ncu | tail -12
