#!/bin/bash
# Round-2 GPU call H (2 GPUs): FSDP parity (pytest launcher) and config 2 (base, 8 clips per GPU) at 2 GPUs.
mkdir -p gpurun_out
rm -f gpurun_out/h_*
timeout 600 python -m pytest tests/test_fsdp_gpu.py -m gpu -q -p no:cacheprovider -s 2>&1 | grep -E "rank|passed|failed|PASS|FAIL" > gpurun_out/h_fsdp_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus 2 --variant base --batch-per-gpu 8 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/h_err.log | grep "^{" > gpurun_out/h_base2.json
cat gpurun_out/h_fsdp_pytest.txt
python -c "
import json; d=json.loads(open('gpurun_out/h_base2.json').read()); print('base x2:', d['value'], d['ms_per_step'], d.get('gpu_baseline',{}).get('value'))"
