"""Thin tensor-level wrappers over the C ABI (one Python function per entry point).

Shapes / dtypes are validated here so that a bad call fails with a Python exception before any
kernel is launched; the kernels themselves never allocate.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream

K_MAJOR, MN_MAJOR = 0, 1
EPI_BF16, EPI_BF16_GELU, EPI_BF16_RESIDUAL, EPI_BF16_GELU_BWD, EPI_F32, EPI_F32_ATOMIC_ADD = range(6)


def _req(cond, msg):
    if not cond:
        raise ValueError(msg)


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, out=None, out2=None, bias=None, aux=None,
         epi=EPI_BF16, split_k=1, block_n=0):
    """D[M,N] = epi(sum_k A[m,k] B[n,k]) on tcgen05.  `a` is (M,K) (or (K,M) when a_mn), `b` is (N,K)
    (or (K,N) when b_mn); both bf16 with unit inner stride.  See include/oasr_b200.h."""
    _req(a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, "gemm: operands must be bf16")
    _req(a.is_cuda and b.is_cuda, "gemm: operands must be CUDA tensors")
    _req(a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1, "gemm: 2-D, unit inner stride")
    _req(tuple(a.shape) == ((K, M) if a_mn else (M, K)), f"gemm: A shape {tuple(a.shape)} vs M={M} K={K} a_mn={a_mn}")
    _req(tuple(b.shape) == ((K, N) if b_mn else (N, K)), f"gemm: B shape {tuple(b.shape)} vs N={N} K={K} b_mn={b_mn}")
    f32_out = epi in (EPI_F32, EPI_F32_ATOMIC_ADD)
    if out is None:
        _req(epi != EPI_F32_ATOMIC_ADD, "gemm: atomic-add epilogue needs a pre-initialised `out`")
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    _req(out.dtype == (torch.float32 if f32_out else torch.bfloat16), "gemm: out dtype")
    _req(out.dim() == 2 and out.shape[0] == M and out.shape[1] >= N and out.stride(1) == 1, "gemm: out shape")
    if epi == EPI_BF16_GELU and out2 is None:
        out2 = torch.empty_like(out)
    if aux is not None:
        _req(aux.dtype == torch.bfloat16 and aux.stride(1) == 1 and aux.shape[0] == M, "gemm: aux")
    if bias is not None:
        _req(bias.dtype == torch.float32 and bias.numel() >= N and bias.is_contiguous(), "gemm: bias must be f32")
    call("oasr_gemm_bf16", ptr(a), a.stride(0), int(a_mn), ptr(b), b.stride(0), int(b_mn),
         ptr(out), out.stride(0), ptr(out2), ptr(bias), ptr(aux), aux.stride(0) if aux is not None else 0,
         M, N, K, epi, split_k, block_n, stream())
    if epi == EPI_BF16_GELU:
        return out, out2
    return out
