"""Bring-up harness for the tcgen05 GEMM: every configuration runs in its own process under a
timeout so that a trap / hang in one variant cannot take the others down.

    python tools/gpu_gemm_bringup.py            # all variants
    python tools/gpu_gemm_bringup.py one <json> # (internal) one variant
"""
import json
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def run_one(cfg):
    import torch
    from olmoasr_b200 import kernels as K

    torch.manual_seed(0)
    M, N, Kd = cfg["M"], cfg["N"], cfg["K"]
    a_mn, b_mn, epi, bn, split = cfg["a_mn"], cfg["b_mn"], cfg["epi"], cfg["bn"], cfg.get("split", 1)
    dev = "cuda"
    pat = cfg.get("pattern", "rand")
    if pat == "rand":
        A = torch.randn(M, Kd, device=dev)
        B = torch.randn(N, Kd, device=dev)
    elif pat == "a_ones":
        A = torch.ones(M, Kd, device=dev)
        B = torch.randn(N, Kd, device=dev)
    else:
        A = torch.randn(M, Kd, device=dev)
        B = torch.ones(N, Kd, device=dev)
    A = A.bfloat16()
    B = B.bfloat16()
    a_st = A.t().contiguous() if a_mn else A
    b_st = B.t().contiguous() if b_mn else B
    ref = A.float() @ B.float().t()
    bias = torch.randn(N, device=dev) if cfg.get("bias") else None
    aux = torch.randn(M, N, device=dev).bfloat16() if epi in (K.EPI_BF16_RESIDUAL, K.EPI_BF16_GELU_BWD) else None
    out = None
    if epi == K.EPI_F32_ATOMIC_ADD:
        out = torch.full((M, N), 1.0, device=dev, dtype=torch.float32)
    res = K.gemm(a_st, b_st, M, N, Kd, a_mn=a_mn, b_mn=b_mn, out=out, bias=bias, aux=aux, epi=epi,
                 split_k=split, block_n=bn)
    torch.cuda.synchronize()
    bb = bias.bfloat16().float() if bias is not None else 0.0
    if epi == K.EPI_BF16:
        want = (ref + bb).bfloat16().float(); got = res.float()
    elif epi == K.EPI_BF16_GELU:
        h = (ref + bb).bfloat16().float()
        want = torch.nn.functional.gelu(h).bfloat16().float(); got = res[1].float()
        assert (res[0].float() - h).abs().max().item() <= 0.02 * h.abs().max().item()
    elif epi == K.EPI_BF16_RESIDUAL:
        want = (aux.float() + (ref + bb).bfloat16().float()).bfloat16().float(); got = res.float()
    elif epi == K.EPI_BF16_GELU_BWD:
        x = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(x).backward(ref.bfloat16().float())
        want = x.grad.bfloat16().float(); got = res.float()
    elif epi == K.EPI_F32:
        want = ref + (bias if bias is not None else 0.0); got = res
    else:
        want = ref + 1.0; got = res
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    bad = ((got - want).abs() > 0.02 * scale + 1e-3).nonzero()
    info = {"max_err": err, "scale": scale, "n_bad": int(bad.shape[0])}
    if bad.shape[0]:
        info["first_bad"] = bad[:6].tolist()
        i, j = bad[0].tolist()
        info["got"] = got[i, j].item(); info["want"] = want[i, j].item()
        rows = sorted(set(bad[:, 0].tolist()))[:8]; cols = sorted(set(bad[:, 1].tolist()))[:8]
        info["bad_rows_head"] = rows; info["bad_cols_head"] = cols
    # quick timing
    if cfg.get("time"):
        for _ in range(3):
            K.gemm(a_st, b_st, M, N, Kd, a_mn=a_mn, b_mn=b_mn, out=out, bias=bias, aux=aux, epi=epi, split_k=split, block_n=bn)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            K.gemm(a_st, b_st, M, N, Kd, a_mn=a_mn, b_mn=b_mn, out=out, bias=bias, aux=aux, epi=epi, split_k=split, block_n=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        info["ms"] = ms; info["tflops"] = 2.0 * M * N * Kd / ms / 1e9
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        Bt = B.t().contiguous()
        for _ in range(3): torch.matmul(A, Bt)
        t0.record()
        for _ in range(10): torch.matmul(A, Bt)
        t1.record(); torch.cuda.synchronize()
        info["cublas_tflops"] = 2.0 * M * N * Kd / (t0.elapsed_time(t1) / 10) / 1e9
    print("RESULT " + json.dumps(info))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        run_one(json.loads(sys.argv[2]))
        return
    cfgs = []
    base = dict(a_mn=0, b_mn=0, epi=0, bias=0)
    # smallest possible: one tile, one k-block
    for bn in (64, 128, 256):
        cfgs.append(dict(base, M=128, N=bn, K=64, bn=bn, name=f"single-tile bn{bn}"))
    cfgs.append(dict(base, M=128, N=128, K=64, bn=128, pattern="a_ones", name="A=1 (B read check)"))
    cfgs.append(dict(base, M=128, N=128, K=64, bn=128, pattern="b_ones", name="B=1 (A read check)"))
    cfgs.append(dict(base, M=128, N=128, K=256, bn=128, name="4 k-blocks"))
    cfgs.append(dict(base, M=128, N=128, K=1024, bn=128, name="16 k-blocks (ring wrap)"))
    cfgs.append(dict(base, M=4096, N=1024, K=1024, bn=256, name="multi-tile persistent bn256"))
    cfgs.append(dict(base, M=4096, N=1024, K=1024, bn=128, name="multi-tile persistent bn128"))
    cfgs.append(dict(base, M=1000, N=200, K=240, bn=128, name="ragged M/N/K"))
    cfgs.append(dict(base, M=300, N=51865 % 1000 + 1000, K=128, bn=256, name="odd N (scalar store path)"))
    for epi in (1, 2, 3, 4):
        cfgs.append(dict(base, M=512, N=512, K=512, bn=128, epi=epi, bias=int(epi != 3), name=f"epilogue {epi}"))
    # MN-major operands
    cfgs.append(dict(base, M=128, N=128, K=64, bn=128, b_mn=1, name="B MN-major single"))
    cfgs.append(dict(base, M=128, N=128, K=64, bn=128, a_mn=1, b_mn=1, name="A,B MN-major single"))
    cfgs.append(dict(base, M=128, N=128, K=64, bn=128, a_mn=1, name="A MN-major single"))
    cfgs.append(dict(base, M=1024, N=1024, K=1024, bn=256, b_mn=1, name="dgrad-like (K,MN) bn256"))
    cfgs.append(dict(base, M=1024, N=1024, K=4096, bn=256, a_mn=1, b_mn=1, epi=4, name="wgrad-like (MN,MN) f32"))
    cfgs.append(dict(base, M=1024, N=1024, K=4096, bn=256, a_mn=1, b_mn=1, epi=5, split=4, name="wgrad split-k atomic"))
    cfgs.append(dict(base, M=1000, N=520, K=1000, bn=128, a_mn=1, b_mn=1, epi=4, name="wgrad ragged"))
    # timing
    cfgs.append(dict(base, M=48000, N=1024, K=1024, bn=256, bias=1, time=1, name="perf enc proj bn256"))
    cfgs.append(dict(base, M=48000, N=1024, K=1024, bn=128, bias=1, time=1, name="perf enc proj bn128"))
    cfgs.append(dict(base, M=48000, N=4096, K=1024, bn=256, bias=1, epi=1, time=1, name="perf fc1+gelu"))
    cfgs.append(dict(base, M=48000, N=1024, K=4096, bn=256, bias=1, epi=2, time=1, name="perf fc2+residual"))
    cfgs.append(dict(base, M=48000, N=1024, K=1024, bn=256, b_mn=1, time=1, name="perf dgrad"))
    cfgs.append(dict(base, M=1024, N=1024, K=48000, bn=256, a_mn=1, b_mn=1, epi=5, split=5, time=1, name="perf wgrad split5"))
    cfgs.append(dict(base, M=8192, N=8192, K=8192, bn=256, time=1, name="perf 8192^3"))
    cfgs.append(dict(base, M=14336, N=1024, K=1024, bn=128, bias=1, time=1, name="perf dec proj bn128"))
    cfgs.append(dict(base, M=14336, N=51865, K=1024, bn=256, time=1, name="perf logits"))
    cfgs.append(dict(base, M=2048 + 128, N=1024, K=512, bn=256, name="odd number of M tiles (17)"))
    cfgs.append(dict(base, M=2048 + 128, N=1024, K=512, bn=128, b_mn=1, name="odd M tiles, B MN-major"))
    for c in list(cfgs):
        if c.get("time"):
            cfgs.append(dict(c, cluster=1, name=c["name"] + " [no multicast]"))
    only = sys.argv[1] if len(sys.argv) > 1 else None
    n_fail = 0
    for c in cfgs:
        if only and only not in c["name"]:
            continue
        t = time.time()
        try:
            import os
            env = dict(os.environ)
            if "cluster" in c:
                env["OASR_GEMM_CLUSTER"] = str(c["cluster"])
            r = subprocess.run([sys.executable, __file__, "one", json.dumps(c)], capture_output=True, text=True, timeout=180, env=env)
            res = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if r.returncode != 0 or not res:
                n_fail += 1
                print(f"[CRASH] {c['name']}: rc={r.returncode}\n  stdout: {r.stdout[-600:]}\n  stderr: {r.stderr[-1200:]}")
                continue
            info = json.loads(res[0][7:])
            ok = info["n_bad"] == 0
            n_fail += (not ok)
            print(f"[{'PASS' if ok else 'FAIL'}] {c['name']}: {info}  ({time.time()-t:.1f}s)")
        except subprocess.TimeoutExpired:
            n_fail += 1
            print(f"[TIMEOUT] {c['name']}")
        sys.stdout.flush()
    print(f"failures: {n_fail}")


if __name__ == "__main__":
    main()
