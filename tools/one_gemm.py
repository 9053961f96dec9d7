"""Run one GEMM shape / epilogue a few times (ncu target, or quick timing).

    python tools/one_gemm.py M N K [--epi gelu|resid|gelu_bwd|bf16|f32] [--b-mn] [--a-mn] [--iters 20]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("M", type=int); ap.add_argument("N", type=int); ap.add_argument("K", type=int)
    ap.add_argument("--epi", default="bf16")
    ap.add_argument("--a-mn", action="store_true"); ap.add_argument("--b-mn", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--seconds", type=float, default=0.0, help="loop for this long and report SM clock / power (nvidia-smi)")
    args = ap.parse_args()
    from olmoasr_b200 import kernels as K
    dev = torch.device("cuda", 0)
    M, N, Kd = args.M, args.N, args.K
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn((Kd, M) if args.a_mn else (M, Kd), device=dev, generator=g).bfloat16()
    b = (torch.randn((Kd, N) if args.b_mn else (N, Kd), device=dev, generator=g) * 0.03).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    aux = torch.randn(M, N, device=dev, generator=g).bfloat16()
    epi = {"bf16": K.EPI_BF16, "gelu": K.EPI_BF16_GELU, "resid": K.EPI_BF16_RESIDUAL, "gelu_bwd": K.EPI_BF16_GELU_BWD,
           "f32": K.EPI_F32}[args.epi]
    kw = dict(a_mn=args.a_mn, b_mn=args.b_mn, epi=epi)
    if args.epi in ("resid", "gelu_bwd"):
        kw["aux"] = aux
    if args.epi != "gelu_bwd":
        kw["bias"] = bias
    for _ in range(3):
        K.gemm(a, b, M, N, Kd, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        K.gemm(a, b, M, N, Kd, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    if args.seconds > 0:
        import subprocess
        import time
        n = max(1, int(args.seconds * 1e3 / ms))
        pr = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"],
                              stdout=subprocess.PIPE, text=True)
        e0.record()
        for _ in range(n):
            K.gemm(a, b, M, N, Kd, **kw)
        e1.record()
        torch.cuda.synchronize()
        pr.terminate()
        rows = [ln.split(",") for ln in pr.stdout.read().strip().splitlines()]
        rows = rows[len(rows) // 3:]   # steady state
        clk = sorted(float(r[0]) for r in rows)
        pw = sorted(float(r[1]) for r in rows)
        ms = e0.elapsed_time(e1) / n
        print(f"   sustained {args.seconds:.0f} s: sm clock median {clk[len(clk) // 2]:.0f} MHz, power median {pw[len(pw) // 2]:.0f} W, ", end="")
    print(f"M={M} N={N} K={Kd} epi={args.epi} a_mn={args.a_mn} b_mn={args.b_mn}: {ms:.4f} ms  {2.0 * M * N * Kd / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
