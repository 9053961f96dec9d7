"""TEST INFRASTRUCTURE: plain-torch stand-ins for the functions of olmoasr_b200/kernels.py (the thin wrappers over the C ABI),
installed by monkeypatching, so that the HOST logic above the kernels -- the per-block autograd Functions, the operand /
shadow dictionaries, slab mode with direct gradient accumulation, the tied-embedding loss head, the conv-stem sequencing
-- can be executed and checked against the oracle on a machine without a GPU.

This is not a CPU path of the product: nothing under olmoasr_b200/ imports it, it lives under tests/, and the GPU tests
never use it.  Each stand-in states the contract of the kernel it replaces (same argument order, same in-place /
accumulate behaviour, bf16 rounding where the kernel rounds).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BF = torch.bfloat16


def _r(x):  # round to bf16, keep computing in fp32
    return x.to(BF).float()


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, out=None, out2=None, bias=None, aux=None, epi=0, split_k=1, block_n=0):
    A = (a.t() if a_mn else a).float()
    B = (b.t() if b_mn else b).float()
    assert A.shape == (M, K) and B.shape == (N, K), (A.shape, B.shape, M, N, K)
    acc = A @ B.t()
    bz = _r(bias[:N].float()) if bias is not None else 0.0
    f32_out = epi in (4, 5)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32 if f32_out else BF)
    assert out.dtype == (torch.float32 if f32_out else BF) and out.shape[0] == M and out.shape[1] >= N
    if epi == 0:
        out[:, :N] = (acc + bz).to(BF)
    elif epi == 1:
        h = (acc + bz).to(BF)
        out[:, :N] = h
        if out2 is None:
            out2 = torch.empty_like(out)
        out2[:, :N] = F.gelu(h.float()).to(BF)
        return out, out2
    elif epi == 2:
        out[:, :N] = (aux.float() + _r(acc + bz)).to(BF)
    elif epi == 3:
        x = aux.float()
        cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
        pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
        out[:, :N] = (_r(acc) * (cdf + x * pdf)).to(BF)
    elif epi == 4:
        out[:, :N] = acc + (bias[:N].float() if bias is not None else 0.0)
    elif epi == 5:
        out[:, :N] += acc
    return out


def layernorm_fwd(x, weight, bias, eps=1e-5, want_stats=True):
    xf = x.float()
    mean = xf.mean(-1)
    var = xf.var(-1, unbiased=False)
    rstd = torch.rsqrt(var + eps)
    y = (((xf - mean[:, None]) * rstd[:, None]) * weight.float() + bias.float()).to(BF)
    return (y, mean, rstd) if want_stats else y


def layernorm_bwd(dy, x, weight, mean, rstd, dweight, dbias, dresidual=None):
    xf, g = x.float(), dy.float()
    xhat = (xf - mean[:, None]) * rstd[:, None]
    dweight += (g * xhat).sum(0)
    dbias += g.sum(0)
    gw = g * weight.float()
    d = xf.shape[1]
    dx = (gw - gw.mean(-1, keepdim=True) - xhat * (gw * xhat).mean(-1, keepdim=True)) * rstd[:, None]
    dx = _r(dx)
    if dresidual is not None:
        dx = dresidual.float() + dx
    return dx.to(BF)


def _heads(t, B, T, H):
    return t.float().reshape(B, T, H, 64).permute(0, 2, 1, 3)


def _mask(B, Tq, Tkv, causal, kv_len):
    m = torch.zeros(B, 1, Tq, Tkv)
    if causal:
        m = m + torch.full((Tq, Tkv), -math.inf).triu_(1)
    if kv_len is not None:
        for i in range(B):
            m[i, :, :, int(kv_len[i]):] = -math.inf
    return m


def attention_fwd(q, k, v, B, H, Tq, Tkv, causal=False, kv_len=None, scale=None, want_lse=True, out=None):
    qh, kh, vh = _heads(q, B, Tq, H), _heads(k, B, Tkv, H), _heads(v, B, Tkv, H)
    s = qh @ kh.transpose(-1, -2) * (scale or 64 ** -0.5) + _mask(B, Tq, Tkv, causal, kv_len)
    p = torch.softmax(s, -1)
    o = (_r(p) @ vh).permute(0, 2, 1, 3).reshape(B * Tq, H * 64).to(BF)
    if out is not None:
        out.copy_(o)
        o = out
    return o, (torch.logsumexp(s, -1) if want_lse else None)


def attention_bwd(q, k, v, o, dout, lse, B, H, Tq, Tkv, causal=False, kv_len=None, scale=None, dq=None, dk=None, dv=None):
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    with torch.enable_grad():
        qh, kh, vh = _heads(qf, B, Tq, H), _heads(kf, B, Tkv, H), _heads(vf, B, Tkv, H)
        s = qh @ kh.transpose(-1, -2) * (scale or 64 ** -0.5) + _mask(B, Tq, Tkv, causal, kv_len)
        oo = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B * Tq, H * 64)
        oo.backward(dout.float())
    res = []
    for g, dst in ((qf.grad, dq), (kf.grad, dk), (vf.grad, dv)):
        g = g.to(BF)
        if dst is not None:
            dst.copy_(g)
            g = dst
        res.append(g)
    return tuple(res)


def ce_fwd(logits, targets, V, ignore_index):
    x = logits[:, :V].float()
    lse = torch.logsumexp(x, -1)
    valid = (targets != ignore_index) & (targets >= 0) & (targets < V)
    lsc = torch.zeros(4)
    t = targets.clamp(0, V - 1)
    lsc[0] = ((lse - x.gather(1, t[:, None])[:, 0]) * valid).sum()
    lsc[1] = valid.sum()
    lsc[2] = ((targets != ignore_index) & ~valid).sum()
    return torch.where(valid, lse, torch.zeros_like(lse)), lsc


def ce_finalize(lsc):
    return (lsc[0] / lsc[1]).reshape(())


def ce_bwd_(logits, targets, lse, lsc, grad_out, V, ignore_index):
    x = logits[:, :V].float()
    valid = (targets != ignore_index) & (targets >= 0) & (targets < V)
    g = grad_out.reshape(()) / lsc[1]
    p = torch.exp(x - lse[:, None])
    p[torch.arange(x.shape[0]), targets.clamp(0, V - 1)] -= 1.0
    logits[:, :V] = (g * p * valid[:, None]).to(BF)
    return logits


def logits_to_f32(logits, V):
    return logits[:, :V].float().contiguous()


def embed_fwd(ids, emb, pos, pos_offset=0):
    B, S = ids.shape
    return (emb[ids].float() + pos[pos_offset:pos_offset + S].float()).reshape(B * S, -1).to(BF)


def embed_bwd(ids, dx, demb, dpos, padding_idx):
    B, S = ids.shape
    g = dx.float().reshape(B, S, -1)
    flat = ids.reshape(-1)
    keep = flat != padding_idx
    demb.index_add_(0, flat[keep], g.reshape(B * S, -1)[keep])
    dpos[:S] += g.sum(0)


def cast_bf16(src, dst=None):
    if dst is None:
        return src.to(BF)
    dst.copy_(src.to(BF).reshape(dst.shape))
    return dst


def convert(src, dst=None, dtype=None):
    if dst is None:
        return src.to(dtype)
    dst.copy_(src.to(dst.dtype).reshape(dst.shape))
    return dst


def cast_conv_weight(w, dst=None):
    out = w.permute(0, 2, 1).reshape(w.shape[0], -1).to(BF)       # column = k * C_in + c
    if dst is not None:
        dst.copy_(out)
        return dst
    return out


def unpermute_conv_wgrad(g, c_out, c_in, out=None):
    r = g.reshape(c_out, 3, c_in).permute(0, 2, 1).contiguous()
    if out is not None:
        out += r
        return out
    return r


def im2col_conv1(mel, kpad):
    B, C, T = mel.shape
    x = F.pad(mel, (1, 1))                                          # (B, C, T + 2)
    cols = torch.stack([x[:, :, k:k + T] for k in range(3)], dim=1)  # (B, 3, C, T)
    A = torch.zeros(B * T, kpad)
    A[:, : 3 * C] = cols.permute(0, 3, 1, 2).reshape(B * T, 3 * C)
    return A.to(BF)


def im2col_conv2(h, B, T_in, d):
    T_out = (T_in + 2 - 3) // 2 + 1
    x = F.pad(h.reshape(B, T_in, d), (0, 0, 1, 1))                  # rows -1 .. T_in
    rows = torch.stack([x[:, k:k + 2 * T_out:2] for k in range(3)], dim=2)   # (B, T_out, 3, d): rows 2t-1, 2t, 2t+1
    return rows.reshape(B * T_out, 3 * d).contiguous()


def col2im_conv2_gelu_bwd(dA, pre1, B, T_in, T_out, d):
    cols = dA.float().reshape(B, T_out, 3, d)
    acc = torch.zeros(B, T_in + 2, d)
    for k in range(3):
        acc[:, k:k + 2 * T_out:2] += cols[:, :, k]
    g = _r(acc[:, 1:T_in + 1]).reshape(B * T_in, d)
    x = pre1.float()
    cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
    pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    return (g * (cdf + x * pdf)).to(BF)


def add_pos(x, pos, T):
    return (x.float().reshape(-1, T, x.shape[1]) + pos.float()).reshape(x.shape).to(BF)


def gelu_bwd(dy, pre):
    x = pre.float()
    cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
    pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    return (dy.float() * (cdf + x * pdf)).to(BF)


def colsum_(dy, db, N=None):
    N = dy.shape[1] if N is None else N
    db += dy[:, :N].float().sum(0)
    return db


def mask_to_kvlen(mask, err_flag):
    kv = (mask[:, 0, :] == 0).sum(-1).to(torch.int32)
    S = mask.shape[1]
    want = torch.where(torch.arange(S)[None, None, :] < kv[:, None, None], 0.0, -math.inf).expand_as(mask)
    ok = ((mask == 0) == (want == 0)) & ((mask < -1e30) == (want < -1e30))
    if not bool(ok.all()):
        err_flag |= 1
    return kv


def install(monkeypatch):
    """Route olmoasr_b200.kernels (and the two host helpers that insist on a CUDA device) to the stand-ins above."""
    from olmoasr_b200 import _core, kernels

    for name, fn in globals().items():
        if callable(fn) and not name.startswith("_") and name not in ("install",) and hasattr(kernels, name):
            monkeypatch.setattr(kernels, name, fn)

    def kv_len_cpu(padding_mask):
        if padding_mask.dim() == 1:
            return padding_mask.to(torch.int32)
        flag = torch.zeros(1, dtype=torch.int32)
        kv = mask_to_kvlen(padding_mask.float(), flag)
        if int(flag):
            raise ValueError("padding_mask is not of the form [0]*len + [-inf]*(n_ctx-len)")
        return kv

    monkeypatch.setattr(_core, "kv_len_from_padding_mask", kv_len_cpu)
    monkeypatch.setattr(_core, "_sm_count", lambda: 148)
