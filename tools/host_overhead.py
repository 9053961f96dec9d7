"""Host-side cost of enqueuing one training step: run the medium model at batch 1 (device work is a few ms, so the
wall time per step is the Python / ctypes / allocator time for the ~2400 launches) and print a cProfile of it.

    python tools/host_overhead.py [--variant medium] > gpurun_out/host_overhead.txt
"""
import argparse
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="medium")
    ap.add_argument("--batch", type=int, default=1)
    args = ap.parse_args()
    import olmoasr_b200 as ob
    from olmoasr_b200 import _lib
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
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n0 = _lib.LAUNCH_COUNT
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    t_enq = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 5
    n = (_lib.LAUNCH_COUNT - n0) // 5
    print(f"# {args.variant} B={B}: host enqueue {t_enq * 1e3:.1f} ms/step, wall {t_all * 1e3:.1f} ms/step, {n} launches "
          f"-> {t_enq / n * 1e6:.1f} us per launch")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr, stream=sys.stdout)
    st.sort_stats("tottime").print_stats(35)
    st.sort_stats("cumulative").print_stats(35)


if __name__ == "__main__":
    main()
