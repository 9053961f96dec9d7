#!/bin/bash
# Round-2 GPU call G (8 GPUs): tail-segment A/B of SlabGradSync and NCCL protocol choice, same box.
mkdir -p gpurun_out
rm -f gpurun_out/g_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { echo "== $1" >> gpurun_out/g_ab8.txt; shift; env "$@" timeout 600 $TR --master-port $PORT bench.py --gpus 8 --steps 5 --warmup 3 --no-gpu-baseline --no-cpu-baseline 2>>gpurun_out/g_err.log | grep "^{" >> gpurun_out/g_ab8.txt; }
PORT=29611 run "tail segments 32 MB (default)" X=1
PORT=29612 run "tail segments 256 MB (round-2 first version)" OASR_TAIL_BUCKET_MB=256
PORT=29613 run "tail 32 MB + NCCL_PROTO=Simple" NCCL_PROTO=Simple
PORT=29614 run "tail 32 MB + NCCL_ALGO=NVLS" NCCL_ALGO=NVLS
python - <<'PY'
import json
for ln in open("gpurun_out/g_ab8.txt"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print(f"   {d['value']:.1f} clips/s  {d['ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']:.1f}")
    else:
        print(ln.strip())
PY
grep -v "NCCL INFO" gpurun_out/g_err.log | grep -E "Error|error" | tail -5
