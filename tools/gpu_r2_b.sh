#!/bin/bash
# Round-2 GPU call B (2 GPUs): data-parallel A/B (SlabGradSync vs torch DDP, NCCL channel caps), rank-0 step profile, FSDP parity.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== slab (default)" > gpurun_out/b_ddp_ab.txt
timeout 600 $TR --master-port 29501 bench.py --gpus 2 --steps 5 --warmup 3 --no-gpu-baseline 2>gpurun_out/b_err1.log | tail -1 >> gpurun_out/b_ddp_ab.txt
echo "== slab, NCCL_MAX_NCHANNELS=8" >> gpurun_out/b_ddp_ab.txt
NCCL_MAX_NCHANNELS=8 timeout 600 $TR --master-port 29502 bench.py --gpus 2 --steps 5 --warmup 3 --no-gpu-baseline 2>gpurun_out/b_err2.log | tail -1 >> gpurun_out/b_ddp_ab.txt
echo "== slab, 64 MB segments" >> gpurun_out/b_ddp_ab.txt
OASR_BUCKET_MB=64 timeout 600 $TR --master-port 29503 bench.py --gpus 2 --steps 5 --warmup 3 --no-gpu-baseline 2>gpurun_out/b_err3.log | tail -1 >> gpurun_out/b_ddp_ab.txt
echo "== torch DDP + stock baseline" >> gpurun_out/b_ddp_ab.txt
OASR_DDP_IMPL=torch timeout 900 $TR --master-port 29504 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/b_err4.log | tail -1 >> gpurun_out/b_ddp_ab.txt
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 $TR --master-port 29505 tools/profile_step_ddp.py > gpurun_out/b_profile_ddp2.txt 2> gpurun_out/b_profile_ddp2.err
grep -m 20 -E "NCCL INFO (Channel|Connected|.*algo|.*NVLS|.*Ring|.*Tree)" gpurun_out/b_profile_ddp2.txt gpurun_out/b_profile_ddp2.err | head -20 > gpurun_out/b_nccl_info.txt
timeout 600 $TR --master-port 29506 tools/fsdp_check.py parity > gpurun_out/b_fsdp_parity.txt 2> gpurun_out/b_fsdp_parity.err
tail -3 gpurun_out/b_fsdp_parity.err
timeout 600 $TR --master-port 29507 tools/fsdp_check.py bench --variant large --batch 16 --steps 3 --warmup 2 > gpurun_out/b_fsdp_large2.txt 2> gpurun_out/b_fsdp_large2.err
tail -3 gpurun_out/b_fsdp_large2.err
cat gpurun_out/b_fsdp_parity.txt gpurun_out/b_fsdp_large2.txt
python - <<'PY'
import json
for ln in open("gpurun_out/b_ddp_ab.txt"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print(f"  {d['value']:.1f} clips/s  {d['ms_per_step']:.1f} ms/step  e2e {d['e2e']['value']:.1f}  sync={d['config']['grad_sync'][:40]}  stock={d.get('gpu_baseline', {}).get('value')}")
    else:
        print(ln.strip())
PY
head -30 gpurun_out/b_profile_ddp2.txt
