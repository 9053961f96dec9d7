"""Config 4 of BASELINE.json on >= 2 GPUs: the model under torch FSDP exactly as the reference wraps it
(scripts/training/train_fsdp_timestamps.py:2588-2615 MixedPrecision(bf16, bf16, bf16); :2665-2678 FULL_SHARD with
transformer_auto_wrap_policy({ResidualAttentionBlock}) and BACKWARD_PRE; :2711-2719 non-reentrant activation checkpointing
of every ResidualAttentionBlock; model.clip_grad_norm_ :1598).

    torchrun --nproc-per-node 2 tools/fsdp_check.py parity            # tiny: loss / gradient parity against the plain model
    torchrun --nproc-per-node 8 tools/fsdp_check.py bench --variant large --batch 16 --steps 5

`parity` prints PASS / FAIL lines (tests/test_fsdp_gpu.py runs it when two GPUs are visible)."""
import argparse
import functools
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def wrap(model, local_rank, act_ckpt=True):
    from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import (CheckpointImpl, apply_activation_checkpointing,
                                                                             checkpoint_wrapper)
    from torch.distributed.fsdp import BackwardPrefetch, FullyShardedDataParallel as FSDP, MixedPrecision, ShardingStrategy
    from torch.distributed.fsdp.wrap import transformer_auto_wrap_policy

    from olmoasr_b200.model import ResidualAttentionBlock

    mp = MixedPrecision(param_dtype=torch.bfloat16, reduce_dtype=torch.bfloat16, buffer_dtype=torch.bfloat16)
    policy = functools.partial(transformer_auto_wrap_policy, transformer_layer_cls={ResidualAttentionBlock})
    fs = FSDP(model, device_id=local_rank, auto_wrap_policy=policy, mixed_precision=mp, backward_prefetch=BackwardPrefetch.BACKWARD_PRE,
              sharding_strategy=ShardingStrategy.FULL_SHARD)
    if act_ckpt:
        # (the reference also passes offload_to_cpu=False, an argument torch >= 2.1 no longer has: it would be forwarded
        # into ResidualAttentionBlock.forward)
        apply_activation_checkpointing(fs, checkpoint_wrapper_fn=functools.partial(checkpoint_wrapper,
                                                                                  checkpoint_impl=CheckpointImpl.NO_REENTRANT),
                                       check_fn=lambda m: isinstance(m, ResidualAttentionBlock))
    return fs


def parity(args, rank, world, dev):
    import olmoasr_b200 as ob
    from olmoasr_b200 import synthetic as synth
    from olmoasr_b200.model import OLMoASR
    from torch.distributed.fsdp import FullyShardedDataParallel as FSDP

    dims = ob.VARIANT_TO_DIMS["tiny"]
    torch.manual_seed(0)
    plain = OLMoASR(dims).to(dev)
    torch.manual_seed(0)
    fs = wrap(OLMoASR(dims).to(dev), dev.index, act_ckpt=True)
    B = 2
    mel = ob.log_mel_spectrogram(synth.waveforms(B, rank=rank).to(dev))
    ti, ty, pm, _ = (t.to(dev) for t in synth.text_batch(B, rank=rank))
    loss_p = plain(mel, ti, pm, targets=ty)
    loss_p.backward()
    loss_f = fs(mel, ti, pm, targets=ty)
    loss_f.backward()
    ok = abs(loss_p.item() - loss_f.item()) <= 1e-3 * abs(loss_p.item())
    print(f"[rank {rank}] loss plain {loss_p.item():.6f} fsdp {loss_f.item():.6f} {'PASS' if ok else 'FAIL'} loss", flush=True)
    # gradients: FSDP holds shards of the rank-averaged gradient (bf16 reduce-scatter).  summon_full_params(with_grads=True)
    # does not exist for the reference's use_orig_params=False wrapping, so the check is on the global gradient norm
    # (FSDP's own clip_grad_norm_ returns it) against the norm of the all-reduced gradients of the plain model.
    for p in plain.parameters():
        dist.all_reduce(p.grad)
        p.grad /= world
    norm_plain = float(torch.sqrt(sum((p.grad.float() ** 2).sum() for p in plain.parameters())))
    norm_fsdp = float(fs.clip_grad_norm_(1e9))
    okg = abs(norm_fsdp - norm_plain) <= 3e-2 * norm_plain      # bf16 gradient reduction + bf16-rounded parameter views
    print(f"[rank {rank}] global gradient norm plain {norm_plain:.5f} fsdp {norm_fsdp:.5f} {'PASS' if okg else 'FAIL'} grads", flush=True)
    # one optimizer step through FSDP's own clip + our fused AdamW on the flat shards, then the loss must move
    from olmoasr_b200.optim import FusedAdamW
    opt = FusedAdamW(fs.parameters(), lr=1e-3, max_grad_norm=0.0)
    fs.clip_grad_norm_(1.0)
    opt.step()
    opt.zero_grad()
    with torch.no_grad():
        l2 = fs(mel, ti, pm, targets=ty).item()
    print(f"[rank {rank}] loss after one step {l2:.6f} {'PASS' if l2 < loss_f.item() else 'FAIL'} step", flush=True)


def bench(args, rank, world, dev):
    import olmoasr_b200 as ob
    from olmoasr_b200 import synthetic as synth
    from olmoasr_b200.model import OLMoASR
    from olmoasr_b200.optim import FusedAdamW

    dims = ob.VARIANT_TO_DIMS[args.variant]
    torch.manual_seed(0)
    fs = wrap(OLMoASR(dims).to(dev), dev.index, act_ckpt=True)
    opt = FusedAdamW(fs.parameters(), lr=1.5e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.1, max_grad_norm=0.0)
    B = args.batch
    wav = synth.waveforms(B, rank=rank, int16=True).to(dev)
    ti, ty, pm, _ = (t.to(dev) for t in synth.text_batch(B, rank=rank))

    def step():
        mel = ob.log_mel_spectrogram(wav)
        loss = fs(mel, ti, pm, targets=ty)
        opt.zero_grad()
        loss.backward()
        fs.clip_grad_norm_(1.0)
        opt.step()
        return loss.detach()

    for _ in range(args.warmup):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t)
        print(f"FSDP {args.variant} bf16 act-ckpt, {B} clips/GPU x {world} GPUs: {ms:.1f} ms/step = {B * world / ms * 1e3:.1f} clips/s, "
              f"loss {float(loss):.4f}, peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["parity", "bench"])
    ap.add_argument("--variant", default="large")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    (parity if args.mode == "parity" else bench)(args, rank, world, dev)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
