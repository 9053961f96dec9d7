"""Per-kernel device time of one greedy-decode step (graph replay) at several batch sizes, for profiles/.

    python tools/profile_decode.py [--variant small] [--dtype fp16] [--batches 1,8,64] > gpurun_out/decode_profile.txt
"""
import argparse
import sys
from collections import defaultdict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="small")
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--batches", default="1,8,64")
    ap.add_argument("--pos", type=int, default=112, help="self-attention length at which the step is profiled")
    args = ap.parse_args()
    import olmoasr_b200 as ob
    from olmoasr_b200.inf_model import OLMoASR
    from torch.profiler import ProfilerActivity, profile

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):
        model = OLMoASR(ob.VARIANT_TO_DIMS[args.variant])
    with torch.no_grad():
        model.decoder.positional_embedding.normal_(0, 0.01)
    dtype = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    eng = model.decode_engine(dtype)
    d = model.dims.n_audio_state
    for n in [int(b) for b in args.batches.split(",")]:
        xa = torch.randn(n, 1500, d, device=dev).bfloat16()
        eng.prepare(xa)
        eng.reset(n, torch.tensor([[50257, 50362]], dtype=torch.int32).repeat(n, 1))
        for _ in range(args.pos):
            eng.replay(n)
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.replay(n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(reps):
                eng.replay(n)
            torch.cuda.synchronize()
        agg = defaultdict(lambda: [0.0, 0])
        evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
        for ev in evs:
            nm = ev.name
            for key in ("dec_linear_kernel", "dec_attn_scores_kernel", "dec_attn_pv_kernel", "dec_sample_kernel", "dec_embed_kernel",
                        "dec_advance_kernel"):
                if key in nm:
                    nm = key + (nm[nm.find("<"):nm.find(">") + 1] if "<" in nm else "")
            agg[nm[:70]][0] += ev.device_time_total
            agg[nm[:70]][1] += 1
        busy = sum(v[0] for v in agg.values()) / reps / 1e3
        t = int(eng.pos.item())
        gb = eng.step_bytes(n, t) / 1e9
        print(f"# {args.variant} {args.dtype} N={n} at t={t}: {ms:.3f} ms per step ({eng.launches_per_step} launches), kernels busy {busy:.3f} ms, "
              f"algorithmic {gb:.3f} GB -> {gb / ms * 1e3:.0f} GB/s")
        print(f"{'kernel':72s} {'per step':>9s} {'us each':>9s} {'us/step':>9s}")
        for nm, (tt, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"{nm:72s} {c / reps:9.1f} {tt / c:9.2f} {tt / reps:9.1f}")
        # one step in launch order (the last replay): embed, then per layer LN+qkv, self scores, self pv, out, LN+q, cross scores,
        # cross pv, out, LN+fc1+GELU, fc2, ... , LN+logits, sample, advance
        last = sorted(evs, key=lambda ev: ev.time_range.start)[-eng.launches_per_step:]
        cross = ("c_att",) if eng.cross_splits(n) == 1 else ("c_sc", "c_pv")
        names = ["embed"] + [f"L{l}.{k}" for l in range(eng.L) for k in ("qkv", "s_att", "out", "cq") + cross + ("cout", "fc1", "fc2")] + \
                ["logits", "sample", "advance"]
        per = defaultdict(float)
        for nm, ev in zip(names, last):
            per[nm.split(".")[-1]] += ev.device_time_total
        gaps = sum(max(0.0, b.time_range.start - a.time_range.end) for a, b in zip(last, last[1:]))
        print("# by call site (us per step, summed over layers): " + "  ".join(f"{k} {v:.1f}" for k, v in per.items()) + f"  | gaps between kernels {gaps:.1f}")
        print()


if __name__ == "__main__":
    main()
