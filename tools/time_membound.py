"""HBM-bound kernels at the model's shapes: time, GB/s and fraction of the measured copy bandwidth.

    python tools/time_membound.py            # medium width (d = 1024), 48000 / 14336 rows
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    from olmoasr_b200 import kernels as K
    dev = torch.device("cuda", 0)
    peaks = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
    hbm = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6500.0
    d = 1024
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

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
    # fused optimizer at medium size is timed by the step profile (one launch per step)
    del flush


if __name__ == "__main__":
    main()
