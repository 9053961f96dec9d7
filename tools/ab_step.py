"""In-process A/B of an environment switch on the full training step: the variants alternate step by step on the same
box / model / clocks, so thermal and power drift cancels.  Only switches that the library re-reads on every call work
(OASR_GEMM_CLUSTER) plus the Python-level olmoasr_b200._core.SIDE_STREAM ("SIDE_STREAM=0,1").

    python tools/ab_step.py OASR_GEMM_CLUSTER=2,4 [--rounds 6]
"""
import argparse
import os
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("switch")
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--variant", default="medium")
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    name, vals = args.switch.split("=")
    vals = vals.split(",")
    import olmoasr_b200 as ob
    from olmoasr_b200 import _core
    from olmoasr_b200 import synthetic as synth
    from olmoasr_b200.model import OLMoASR
    from olmoasr_b200.optim import FusedAdamW

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):
        model = OLMoASR(ob.VARIANT_TO_DIMS[args.variant])
    opt = FusedAdamW(model.parameters())
    B = args.batch
    wav = synth.waveforms(B, int16=True).to(dev)
    ti, ty, pm, _ = (t.to(dev) for t in synth.text_batch(B))

    def step():
        mel = ob.log_mel_spectrogram(wav)
        loss = model(mel, ti, pm, targets=ty)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    def set_variant(v):
        if name == "SIDE_STREAM":
            _core.SIDE_STREAM = v != "0"
        else:
            os.environ[name] = v

    for v in vals:
        set_variant(v)
        for _ in range(2):
            step()
    torch.cuda.synchronize()
    times = {v: [] for v in vals}
    for _ in range(args.rounds):
        for v in vals:
            set_variant(v)
            step()   # settle
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(); step()
            e1.record()
            torch.cuda.synchronize()
            times[v].append(e0.elapsed_time(e1) / 2)
    for v in vals:
        t = times[v]
        print(f"{name}={v}: median {statistics.median(t):.2f} ms/step  (min {min(t):.2f}, max {max(t):.2f}, n={len(t)})")


if __name__ == "__main__":
    main()
