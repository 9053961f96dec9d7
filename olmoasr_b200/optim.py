"""Fused AdamW with global-norm clipping (the optimizer step right after the hot path:
scripts/training/train_timestamps.py:1508-1522, defaults :2110-2116).  Two kernel launches per step for the
whole model; state (`exp_avg`, `exp_avg_sq`, `step`) uses torch.optim.AdamW's names so `state_dict()` round-trips
with the reference's `optimizer_state_dict` checkpoints (train_timestamps.py:930-955)."""
from __future__ import annotations

import math

import numpy as np
import torch

from ._lib import call, lib, ptr, stream


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1.5e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.1, max_grad_norm=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self._tables = None
        self._norm_sq = None
        self._found_inf = None

    def _build(self):
        self._groups = []
        chunk = lib().oasr_optim_chunk_elems()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            for p in ps:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            dev = ps[0].device
            recs = np.zeros((len(ps), 5), dtype=np.int64)
            chunks = []
            for i, p in enumerate(ps):
                assert p.dtype == torch.float32 and p.is_contiguous(), "FusedAdamW: fp32 contiguous parameters only"
                st = self.state[p]
                recs[i] = (p.data_ptr(), 0, st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
                chunks += [(i, c) for c in range(math.ceil(p.numel() / chunk))]
            host = torch.from_numpy(recs).pin_memory()
            self._groups.append(dict(group=group, params=ps, host=host, dev=torch.empty_like(host, device=dev),
                                     chunks=torch.tensor(chunks, dtype=torch.int32, device=dev), n_chunks=len(chunks)))
        dev = self._groups[0]["params"][0].device
        self._norm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self._found_inf = torch.zeros(1, device=dev, dtype=torch.float32)
        self._tables = True

    @torch.no_grad()
    def step(self, closure=None, inv_scale: float = 1.0):
        assert closure is None
        if self._tables is None:
            self._build()
        self._found_inf.zero_()
        for g in self._groups:
            host = g["host"]
            for i, p in enumerate(g["params"]):
                grad = p.grad
                assert grad is not None and grad.dtype == torch.float32, "FusedAdamW: every parameter needs an fp32 grad"
                if not grad.is_contiguous():
                    grad = p.grad = grad.contiguous()
                host[i, 1] = grad.data_ptr()
            g["dev"].copy_(host, non_blocking=True)
        # single global norm across all groups (clip_grad_norm_(model.parameters()))
        assert len(self._groups) == 1, "FusedAdamW: one param group (the reference uses one: train_timestamps.py:706-735)"
        g = self._groups[0]
        group = g["group"]
        call("oasr_grad_sqnorm", ptr(g["dev"]), ptr(g["chunks"]), g["n_chunks"], ptr(self._norm_sq), stream())
        st0 = self.state[g["params"][0]]
        step = int(st0["step"].item()) + 1 if isinstance(st0["step"], torch.Tensor) else int(st0["step"]) + 1
        b1, b2 = group["betas"]
        call("oasr_adamw_step", ptr(g["dev"]), ptr(g["chunks"]), g["n_chunks"], ptr(self._norm_sq), ptr(self._found_inf),
             float(inv_scale), float(self.max_grad_norm or 0.0), float(group["lr"]), b1, b2, group["eps"],
             group["weight_decay"], step, stream())
        step_t = torch.tensor(float(step))      # one host scalar shared by every entry (torch.optim.AdamW keeps one per param)
        for p in g["params"]:
            self.state[p]["step"] = step_t
        return None

    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the last step's (scaled) gradients, on the device."""
        return self._norm_sq.sqrt()
