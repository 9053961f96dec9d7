"""HBM-bound kernels at the model's shapes: time, GB/s and fraction of the measured copy bandwidth.

    python tools/time_membound.py            # medium width (d = 1024), 48000 / 14336 rows
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


FLUSH = None


def timeit(fn, iters=20):
    """Per-call device time with the 126 MB L2 flushed before EVERY timed call (a 256 MB memset: round 1's version
    allocated the flush buffer and never wrote it, so small inputs were served from L2 -- profiles/r01_membound_times.txt
    shows >100 % lines).  Each call is bracketed by its own event pair; the flush is outside the pair."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    pairs = []
    for _ in range(iters):
        FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    for e0, e1 in pairs:
        total += e0.elapsed_time(e1)
    return total / iters


def main():
    from olmoasr_b200 import kernels as K
    dev = torch.device("cuda", 0)
    peaks = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
    hbm = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6500.0
    d = 1024
    g = torch.Generator(device=dev).manual_seed(0)
    global FLUSH
    FLUSH = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

    def report(name, ms, nbytes):
        gbs = nbytes / ms / 1e6
        print(f"{name:44s} {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  {100 * gbs / hbm:5.1f}% of {hbm:.0f}")

    for rows in (48000, 14336):
        x = torch.randn(rows, d, device=dev, generator=g).bfloat16()
        dy = torch.randn(rows, d, device=dev, generator=g).bfloat16()
        res = torch.randn(rows, d, device=dev, generator=g).bfloat16()
        w = torch.randn(d, device=dev, generator=g)
        b = torch.randn(d, device=dev, generator=g)
        y, mean, rstd = K.layernorm_fwd(x, w, b)
        dw = torch.zeros(d, device=dev); db = torch.zeros(d, device=dev)
        report(f"layernorm_fwd  {rows}x{d}", timeit(lambda: K.layernorm_fwd(x, w, b)), rows * d * 4)
        report(f"layernorm_bwd  {rows}x{d} (+residual)", timeit(lambda: K.layernorm_bwd(dy, x, w, mean, rstd, dw, db, res)), rows * d * 8)
        report(f"layernorm_bwd  {rows}x{d}", timeit(lambda: K.layernorm_bwd(dy, x, w, mean, rstd, dw, db, None)), rows * d * 6)
        for n in (1024, 3072, 4096):
            t = torch.randn(rows, n, device=dev, generator=g).bfloat16()
            dbn = torch.zeros(n, device=dev)
            report(f"colsum         {rows}x{n}", timeit(lambda: K.colsum_(t, dbn)), rows * n * 2)
            del t
    # log-mel at the bench's batch (32 clips, int16 in, fp32 out: 0.96 + 0.96 MB per clip)
    from olmoasr_b200 import audio, synthetic
    wav = synthetic.waveforms(32, int16=True).to(dev)
    report("log_mel_spectrogram 32 x 480000 int16", timeit(lambda: audio.log_mel_spectrogram(wav)), 32 * (480000 * 2 + 80 * 3000 * 4))
    # fused optimizer at medium size is timed by the step profile (one launch per step)


if __name__ == "__main__":
    main()
