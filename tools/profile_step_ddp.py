"""Per-kernel device-time breakdown of one DDP training step on rank 0 (launch with torchrun, 2+ GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/profile_step_ddp.py > gpurun_out/step_profile_ddp.txt
"""
import os
import sys
from collections import defaultdict
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import olmoasr_b200 as ob
    from olmoasr_b200 import synthetic as synth
    from olmoasr_b200.model import OLMoASR
    from olmoasr_b200.optim import FusedAdamW

    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    with torch.device(dev):
        model = OLMoASR(ob.VARIANT_TO_DIMS["medium"])
    impl = os.environ.get("OASR_DDP_IMPL", "slab")          # "slab": SlabGradSync (default); "torch": DistributedDataParallel
    slabs = model.use_slabs(direct_grads=(impl == "slab"))
    net, sync = model, None
    if impl == "slab":
        from olmoasr_b200.ddp import SlabGradSync
        sync = SlabGradSync(model, slabs, bucket_bytes=int(os.environ.get("OASR_BUCKET_MB", "256")) << 20,
                            tail_bucket_bytes=int(os.environ.get("OASR_TAIL_BUCKET_MB", "32")) << 20)
    else:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = FusedAdamW(model.parameters(), slabs=slabs)
    B = 32
    wav = synth.waveforms(B, int16=True).to(dev)
    ti, ty, pm, _ = (t.to(dev) for t in synth.text_batch(B))

    def step():
        mel = ob.log_mel_spectrogram(wav)
        loss = net(mel, ti, pm, targets=ty)
        opt.zero_grad()
        loss.backward()
        opt.step(inv_scale=sync.finish() if sync is not None else 1.0)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"# data parallel x{dist.get_world_size()} ({impl}; segments {len(sync.segments) if sync else '-'}; NCCL_MAX_NCHANNELS="
              f"{os.environ.get('NCCL_MAX_NCHANNELS', 'default')}): {e0.elapsed_time(e1) / 3:.2f} ms/step")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    if rank == 0:
        agg = defaultdict(lambda: [0.0, 0])
        evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
        for ev in evs:
            name = ev.name
            for key in ("gemm_tcgen05_kernel", "attention_fwd_kernel", "attention_bwd_kernel"):
                if key in name:
                    name = key
            agg[name[:80]][0] += ev.device_time_total
            agg[name[:80]][1] += 1
        span = max(ev.time_range.end for ev in evs) - min(ev.time_range.start for ev in evs)
        print(f"# device span {span / 1e3:.2f} ms, sum of kernel time {sum(v[0] for v in agg.values()) / 1e3:.2f} ms")
        for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
            print(f"{name:82s} {n:6d} {t / 1e3:9.3f} ms")
        nccl = sorted((ev for ev in evs if "nccl" in ev.name.lower()), key=lambda ev: ev.time_range.start)
        if nccl:
            t0 = min(ev.time_range.start for ev in evs)
            print("# NCCL kernels (start offset ms, duration ms):")
            for ev in nccl[:60]:
                print(f"   {(ev.time_range.start - t0) / 1e3:8.2f}  {ev.device_time_total / 1e3:7.3f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
