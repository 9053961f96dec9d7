"""Headline benchmark: training throughput in 30 s-clips/s (BASELINE.json metric) for the waveform -> log-mel ->
encoder/decoder fwd+bwd -> token CE -> (DDP all-reduce) -> clip + AdamW step.

    python bench.py --gpus 1 --steps 5 --warmup 3                     # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1    # the reference's CPU path (oracle port) on host cores

One JSON line on stdout (rank 0).  See the bench contract in DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TRAIN_GFLOP_PER_CLIP = {"tiny": 213.2, "base": 453.2, "small": 1629.8, "medium": 5214.4, "large": 10342.1}
METRIC = "30s-clips/sec training"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"], "hbm_gbs": p["hbm_gbs"],
                "source": "MEASURED_PEAKS.json"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU PyTorch path (restated in oracle/, pinned bit-exact to it)
# ------------------------------------------------------------------------------------------------------------------
def usable_cpus() -> int:
    """Cores this process may actually use: affinity mask capped by the cgroup CPU quota (a 128-thread pool on a
    16-core quota is an order of magnitude slower than 16 threads)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def run_cpu_reference(variant: str, clips: int, steps: int, warmup: int, layers=None):
    """The reference's CPU PyTorch path (oracle port, bit-identical model code): fp32, all usable host cores:
    log-mel (numpy) -> model fwd -> CE -> bwd -> clip_grad_norm_ -> AdamW, `clips` clips per step.
    `layers=(Le, Ld)` truncates the encoder / decoder depth (used for the bounded, extrapolated sample)."""
    from dataclasses import replace

    from oracle import logmel
    from oracle import model as OM
    from olmoasr_b200 import synthetic as synth

    threads = usable_cpus()
    torch.set_num_threads(threads)
    dims = OM.variant_dims(variant)
    if layers is not None:
        dims = replace(dims, n_audio_layer=layers[0], n_text_layer=layers[1])
    sd = OM.init_state_dict(dims, seed=0, train=True)
    params = {k: v.requires_grad_(k != "encoder.positional_embedding") for k, v in sd.items()}
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=1.5e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.1)
    wav = synth.waveforms(clips).numpy()
    ti, ty, pm, _ = synth.text_batch(clips)

    def step():
        mel = torch.from_numpy(logmel.log_mel_spectrogram(wav))
        loss = OM.token_ce(OM.model_forward(params, dims, mel, ti, pm, train_model=True), ty)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params.values() if p.requires_grad], 1.0)
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return clips * steps / dt, dt / steps, threads


def bounded_cpu_sample(variant: str):
    """~10-60 s of CPU work: one clip through the full-width model at depth 1+1 and 3+3; the per-layer-pair cost is
    the difference and the full-depth step time is extrapolated linearly (stem, embedding, logits/CE and the
    embedding's optimizer state are in the depth-1 term)."""
    from oracle import model as OM

    L = OM.variant_dims(variant).n_audio_layer
    _, t1, threads = run_cpu_reference(variant, 1, 1, 1, layers=(1, 1))
    _, t2, _ = run_cpu_reference(variant, 1, 1, 1, layers=(3, 3))
    per_pair = max(t2 - t1, 0.0) / 2.0
    full = t1 + (L - 1) * per_pair
    return 1.0 / full, threads, (f"1 clip, {variant} width, fp32, torch CPU ({threads} threads): timed depth 1+1 ({t1:.2f} s) and "
                                 f"3+3 ({t2:.2f} s) steps incl. log-mel/CE/bwd/clip/AdamW, extrapolated linearly to {L}+{L} layers "
                                 f"({full:.1f} s per clip)")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips = 1
    # each "step" of this arm is the bounded sample (extrapolated full-depth step, see bounded_cpu_sample)
    vals = []
    sample = ""
    for _ in range(max(1, min(args.steps, 3))):
        v, threads, sample = bounded_cpu_sample(args.variant)
        vals.append(v)
    value = sum(vals) / len(vals)
    sec_per_step = 1.0 / value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.variant} training step, 30 s synthetic clips (reference CPU path, {clips} clip per step)",
                   "global_batch": clips, "parallelism": "cpu"},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def our_arm(args):
    import torch.distributed as dist

    import olmoasr_b200 as ob
    from olmoasr_b200 import _lib
    from olmoasr_b200 import kernels as K
    from olmoasr_b200.model import OLMoASR
    from olmoasr_b200.optim import FusedAdamW
    from olmoasr_b200 import synthetic as synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"

    dims = ob.VARIANT_TO_DIMS[args.variant]
    B = args.batch_per_gpu
    torch.manual_seed(0)
    with torch.device(dev):
        model = OLMoASR(dims)
    net = model
    reducer = None
    ddp_impl = os.environ.get("OASR_DDP_IMPL", "torch")   # "blockwise": olmoasr_b200.ddp (opt-in until measured on NVLink)
    if world > 1 and ddp_impl == "blockwise":
        from olmoasr_b200.ddp import BlockwiseGradReducer
        reducer = BlockwiseGradReducer(model)
    elif world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], output_device=local_rank)
    opt = FusedAdamW(model.parameters(), lr=1.5e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.1, max_grad_norm=1.0)

    # synthetic batch of SURVEY.md 8(d); host copies are pinned (what a DataLoader with pin_memory hands over)
    wav_h = synth.waveforms(B, rank=rank, int16=True).pin_memory()
    ti_h, ty_h, pm_h, _ = synth.text_batch(B, rank=rank)
    ti_h, ty_h, pm_h = ti_h.pin_memory(), ty_h.pin_memory(), pm_h.pin_memory()
    h2d_bytes = sum(t.numel() * t.element_size() for t in (wav_h, ti_h, ty_h, pm_h))
    dev_in = [t.to(dev) for t in (wav_h, ti_h, ty_h, pm_h)]

    def train_step(wav, ti, ty, pm):
        mel = ob.log_mel_spectrogram(wav)
        loss = net(mel, ti, pm, targets=ty)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if reducer is not None:
            opt.step(inv_scale=reducer.finish())
        else:
            opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        last = None
        for _ in range(n_steps):
            if e2e:
                ins = [t.to(dev, non_blocking=True) for t in (wav_h, ti_h, ty_h, pm_h)]
                last = train_step(*ins).item()  # device->host read of the step's loss, every step
            else:
                last = train_step(*dev_in)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = max(ms, wall_ms) if e2e else ms
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    for _ in range(args.warmup):
        train_step(*dev_in)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.LAUNCH_COUNT
    ms, loss = timed(args.steps, e2e=False)
    launches = _lib.LAUNCH_COUNT - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, loss_e2e = timed(args.steps, e2e=True)
    # roofline pass: the same K steps with every GEMM launch bracketed by CUDA events.  The weight-gradient side stream
    # is switched off for this pass so that launches do not overlap and a launch's event time is its own duration.
    from olmoasr_b200 import _core
    side_prev, _core.SIDE_STREAM = _core.SIDE_STREAM, False
    K.GEMM_PROFILE = []
    ms_serial, _ = timed(args.steps, e2e=False)
    prof, K.GEMM_PROFILE = K.GEMM_PROFILE, None
    _core.SIDE_STREAM = side_prev
    gemm_flops = sum(p[0] for p in prof)
    gemm_ms = sum(p[1].elapsed_time(p[2]) for p in prof)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    global_batch = B * world
    value = global_batch * args.steps / (ms / 1e3)
    e2e_value = global_batch * args.steps / (ms_e2e / 1e3)
    peaks = _peaks()
    gemm_tflops = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    step_tflops = value * TRAIN_GFLOP_PER_CLIP[args.variant] / 1e3 / world
    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"{args.variant} model DDP training bf16, {B} x 30 s synthetic clips per GPU "
                               f"(BASELINE.json configs[2] at {world} GPU(s)): int16 waveform -> log-mel -> fwd/bwd -> CE -> "
                               f"{'NCCL all-reduce -> ' if world > 1 else ''}clip + AdamW",
                   "global_batch": global_batch, "parallelism": f"dp{world}",
                   "grad_sync": ("none" if world == 1 else ("olmoasr_b200.ddp.BlockwiseGradReducer" if reducer is not None
                                                            else "torch DistributedDataParallel")),
                   "l2": "no flush needed: per-step working set (~60 GB of activations at medium/32) is >> the 126 MB L2",
                   "loss": float(loss)},
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "loss": loss_e2e},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all GEMM launches of K steps)",
                     "achieved": gemm_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": gemm_tflops / peaks["bf16_tflops_sustained"], "traffic": None,
                     "peak_source": peaks["source"] + " (sustained: kernel timed inside a long step)",
                     "measured_in": "a second pass of the same K steps with the weight-gradient side stream off (launches "
                                    "serialised, so each launch's CUDA-event time is its own duration)",
                     "serial_ms_per_step": ms_serial / args.steps,
                     "gemm_share_of_step": gemm_ms / ms_serial, "gemm_launches": len(prof),
                     "whole_step_tflops_per_gpu": step_tflops, "whole_step_frac": step_tflops / peaks["bf16_tflops_sustained"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        v, threads, sample = bounded_cpu_sample(args.variant)
        line["cpu_baseline"] = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="medium", choices=list(TRAIN_GFLOP_PER_CLIP))
    ap.add_argument("--batch-per-gpu", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
