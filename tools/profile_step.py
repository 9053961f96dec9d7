"""Per-kernel device-time breakdown of one training step (torch.profiler / CUPTI), for profiles/.

    python tools/profile_step.py [--variant medium] [--batch 32] > gpurun_out/step_profile.txt
"""
import argparse
import sys
from collections import defaultdict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="medium")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--layers", type=int, default=0, help="truncate encoder/decoder depth (for ncu captures)")
    ap.add_argument("--gemm-table", action="store_true", help="per-shape GEMM time / TFLOP/s (CUDA events per launch)")
    ap.add_argument("--serial", action="store_true", help="side stream off: per-kernel device times do not overlap")
    ap.add_argument("--no-profiler", action="store_true", help="just run the steps (when wrapped in ncu)")
    ap.add_argument("--no-slabs", action="store_true", help="plain parameter storage + pointer-table optimizer")
    args = ap.parse_args()
    import olmoasr_b200 as ob
    from olmoasr_b200 import synthetic as synth
    from olmoasr_b200.model import OLMoASR
    from olmoasr_b200.optim import FusedAdamW

    if args.serial:
        from olmoasr_b200 import _core as _c
        _c.SIDE_STREAM = False
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    dims = ob.VARIANT_TO_DIMS[args.variant]
    if args.layers:
        from dataclasses import replace
        dims = replace(dims, n_audio_layer=args.layers, n_text_layer=args.layers)
    with torch.device(dev):
        model = OLMoASR(dims)
    slabs = None if args.no_slabs else model.use_slabs()
    opt = FusedAdamW(model.parameters(), slabs=slabs)
    B = args.batch
    wav = synth.waveforms(B, int16=True).to(dev)
    ti, ty, pm, _ = (t.to(dev) for t in synth.text_batch(B))

    def step():
        mel = ob.log_mel_spectrogram(wav)
        loss = model(mel, ti, pm, targets=ty)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    t_enq = (time.perf_counter() - t0) / 3
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 3
    print(f"# host enqueue time per step {t_enq * 1e3:.1f} ms; wall per step {t_all * 1e3:.1f} ms (CPU-bound if these are close)")
    if args.gemm_table:
        from olmoasr_b200 import _core
        from olmoasr_b200 import kernels as K
        _core.SIDE_STREAM = False   # serialise launches: a launch's event time must be its own duration
        K.GEMM_PROFILE = []
        step()
        torch.cuda.synchronize()
        prof, K.GEMM_PROFILE = K.GEMM_PROFILE, None
        tab = defaultdict(lambda: [0.0, 0.0, 0])
        for fl, a, b, key in prof:
            t = tab[key]
            t[0] += fl; t[1] += a.elapsed_time(b); t[2] += 1
        epi_names = ["bf16", "gelu", "resid", "gelu_bwd", "f32", "f32_atomic"]
        print(f"# per-shape GEMM table ({args.variant} B={B}); layouts 0=K-major 1=MN-major")
        print(f"{'M':>7s} {'N':>6s} {'K':>6s} A B {'epilogue':>10s} splitk bn {'calls':>5s} {'ms':>8s} {'TFLOP/s':>8s}")
        for key, (fl, ms, n) in sorted(tab.items(), key=lambda kv: -kv[1][1]):
            M, N, Kd, am, bm, epi, sk, bn = key
            print(f"{M:7d} {N:6d} {Kd:6d} {am} {bm} {epi_names[epi]:>10s} {sk:6d} {bn:3d} {n:5d} {ms:8.3f} {fl / ms / 1e9:8.1f}")
        print(f"# total {sum(v[1] for v in tab.values()):.2f} ms, {sum(v[0] for v in tab.values()) / sum(v[1] for v in tab.values()) / 1e9:.1f} TFLOP/s")
        return
    if args.no_profiler:
        step()
        torch.cuda.synchronize()
        return
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name
            for key in ("gemm_tcgen05_kernel", "attention_fwd_kernel", "attention_bwd_kernel"):
                if key in name:
                    name = key + ("" if key != "gemm_tcgen05_kernel" else name[name.find("<"):name.find(">") + 1])
            agg[name][0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            agg[name][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"# {args.variant} B={B}: one training step, sum of kernel device time {total / 1e3:.2f} ms")
    print(f"{'kernel':90s} {'calls':>6s} {'ms':>9s} {'share':>7s}")
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{name[:90]:90s} {n:6d} {t / 1e3:9.3f} {100 * t / total:6.2f}%")
    # idle gaps between consecutive device activities, attributed to the kernel that FOLLOWS the gap
    evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda ev: ev.time_range.start)
    gaps = defaultdict(lambda: [0.0, 0])
    span = evs[-1].time_range.end - evs[0].time_range.start
    tot_gap = 0.0
    for a, b in zip(evs, evs[1:]):
        g = b.time_range.start - a.time_range.end
        if g > 0:
            nm = b.name
            for key in ("gemm_tcgen05_kernel", "attention_fwd_kernel", "attention_bwd_kernel"):
                if key in nm:
                    nm = key
            gaps[nm[:60]][0] += g
            gaps[nm[:60]][1] += 1
            tot_gap += g
    print(f"# device span {span / 1e3:.2f} ms, idle between activities {tot_gap / 1e3:.2f} ms; gap preceding each kernel type:")
    for nm, (g, n) in sorted(gaps.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"{nm:62s} {n:6d} {g / 1e3:8.3f} ms  avg {g / n:6.2f} us")


if __name__ == "__main__":
    main()
