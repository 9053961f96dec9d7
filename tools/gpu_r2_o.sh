#!/bin/bash
# multi-item parity of the persistent attention kernels (grid capped) + compute-sanitizer memcheck / racecheck on them
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -5 > gpurun_out/o_pytest.txt
cat gpurun_out/o_pytest.txt
for tool in memcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_attention.py > gpurun_out/o_san_$tool.txt 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|hazard" gpurun_out/o_san_$tool.txt | sort | uniq -c | head -12
done
