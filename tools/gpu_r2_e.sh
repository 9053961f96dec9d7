#!/bin/bash
# Round-2 GPU call E (2 GPUs): FSDP (config 4) parity against the plain model, FSDP large step time at 2 GPUs.
mkdir -p gpurun_out
rm -f gpurun_out/e_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29506 tools/fsdp_check.py parity > gpurun_out/e_fsdp_parity.txt 2> gpurun_out/e_fsdp_parity.err
grep -v "NCCL INFO" gpurun_out/e_fsdp_parity.err | grep -E "Error|error|rank0\]:" | tail -12
timeout 900 $TR --master-port 29507 tools/fsdp_check.py bench --variant large --batch 16 --steps 3 --warmup 2 > gpurun_out/e_fsdp_large2.txt 2> gpurun_out/e_fsdp_large2.err
grep -v "NCCL INFO" gpurun_out/e_fsdp_large2.err | grep -E "Error|error|rank0\]:" | tail -12
timeout 600 python -m pytest tests/test_fsdp_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/e_fsdp_pytest.txt
cat gpurun_out/e_fsdp_parity.txt gpurun_out/e_fsdp_large2.txt gpurun_out/e_fsdp_pytest.txt
