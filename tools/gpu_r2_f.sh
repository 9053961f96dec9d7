#!/bin/bash
# Round-2 GPU call F (8 GPUs): the headline configuration (medium, 256 clips) with the stock-PyTorch arm, a rank-0 profile of the
# 8-GPU step naming the NCCL kernels, and config 4 (large under FSDP, 16 clips per GPU).
mkdir -p gpurun_out
rm -f gpurun_out/f_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29601 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/f_bench8.json 2> gpurun_out/f_bench8.err
grep -v "NCCL INFO" gpurun_out/f_bench8.err | grep -E "Error|error" | tail -5
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 $TR --master-port 29602 tools/profile_step_ddp.py > gpurun_out/f_profile_ddp8.txt 2> gpurun_out/f_profile_ddp8.err
grep -h -m 30 -E "NVLS|Channel 00|Connected all|Ring 00|Trees" gpurun_out/f_profile_ddp8.err gpurun_out/f_profile_ddp8.txt | grep "\[0\]" | head -12 > gpurun_out/f_nccl_info.txt
timeout 900 $TR --master-port 29603 tools/fsdp_check.py bench --variant large --batch 16 --steps 5 --warmup 3 > gpurun_out/f_fsdp_large8.txt 2> gpurun_out/f_fsdp_large8.err
grep -v "NCCL INFO" gpurun_out/f_fsdp_large8.err | grep -E "Error|error" | tail -5
python - <<'PY'
import json
d = json.load(open("gpurun_out/f_bench8.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_baseline", {}).get("value"), d["config"]["grad_sync"])
PY
grep -v "NCCL INFO" gpurun_out/f_profile_ddp8.txt | head -40
cat gpurun_out/f_fsdp_large8.txt gpurun_out/f_nccl_info.txt
