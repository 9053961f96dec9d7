"""In-process A/B of the GEMM SM budget under DDP (launch with torchrun, 2+ GPUs): budgets alternate step by step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        tools/ab_step_ddp.py 0,140,132,116
"""
import os
import statistics
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import olmoasr_b200 as ob
    from olmoasr_b200 import kernels as K
    from olmoasr_b200 import synthetic as synth
    from olmoasr_b200.model import OLMoASR
    from olmoasr_b200.optim import FusedAdamW

    budgets = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,132").split(",")]
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    with torch.device(dev):
        model = OLMoASR(ob.VARIANT_TO_DIMS["medium"])
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = FusedAdamW(model.parameters())
    B = 32
    wav = synth.waveforms(B, int16=True).to(dev)
    ti, ty, pm, _ = (t.to(dev) for t in synth.text_batch(B))

    def step():
        mel = ob.log_mel_spectrogram(wav)
        loss = net(mel, ti, pm, targets=ty)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    times = {b: [] for b in budgets}
    for _ in range(4):
        for b in budgets:
            K.set_gemm_sm_budget(b)
            step()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(); step()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 2], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times[b].append(float(t.item()))
    if rank == 0:
        for b in budgets:
            print(f"gemm sm budget {b:4d}: median {statistics.median(times[b]):.2f} ms/step (min {min(times[b]):.2f}, max {max(times[b]):.2f})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
