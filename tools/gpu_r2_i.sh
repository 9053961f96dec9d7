#!/bin/bash
# Round-2 GPU call I (1 GPU): ncu evidence (launch list + full sets), compute-sanitizer over the new kernels.
mkdir -p gpurun_out
rm -f gpurun_out/ncu_* gpurun_out/launches_* gpurun_out/i_*
bash tools/gpu_ncu.sh > gpurun_out/i_ncu.log 2>&1
tail -12 gpurun_out/i_ncu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_new_kernels.py > gpurun_out/i_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/i_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_new_kernels.py > gpurun_out/i_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/i_racecheck.txt
ls -la gpurun_out | grep -E "ncu_|launches" | awk '{print $5, $9}'
