#!/bin/bash
# ncu --set full of the reworked attention kernels (2+2-layer medium-width model, same per-layer shapes as the bench)
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
for k in attention_fwd attention_bwd; do
  timeout 600 $NCU --set full -k regex:$k -s 6 -c 3 -f -o gpurun_out/ncu2_$k python tools/profile_step.py --layers 2 --no-profiler > gpurun_out/ncu2_$k.stdout 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out | grep ncu2
