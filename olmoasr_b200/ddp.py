"""Block-wise gradient all-reduce: a leaner stand-in for DistributedDataParallel on this path (opt-in).

The reference wraps the model in `DDP(model, device_ids=[rank])` (scripts/training/train_timestamps.py:2329-2330): 25 MB
buckets, one flatten-copy and one divide kernel per parameter (946 + 948 launches per step at medium, 7 ms of device
time, profiles/r01_step_profile_ddp_2gpu.txt) and 87 all-reduce launches wherever a bucket happens to fill.  Here the
gradients stay where the backward kernels wrote them: when the last parameter of a ResidualAttentionBlock has its
gradient, the block's tensors are summed across ranks by ONE coalesced NCCL call (ncclGroupStart / ncclAllReduce per
tensor / ncclGroupEnd -- no flattening), and the 1 / world_size average is folded into the fused optimizer's
`inv_scale` (the clip threshold then applies to the averaged gradient norm, exactly like clip_grad_norm_ after DDP).

    reducer = BlockwiseGradReducer(model)          # after torch.distributed.init_process_group
    loss.backward()
    opt.step(inv_scale=reducer.finish())           # waits for the collectives; returns 1 / world_size

Same collective order on every rank (the autograd graph is the same), same result as DDP up to fp32 summation order.
Status: host logic covered by the 2-rank gloo test; not yet measured on NVLink (written after this round's GPU
budget was spent), so bench.py keeps DistributedDataParallel unless OASR_DDP_IMPL=blockwise.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn


def default_buckets(model: nn.Module) -> List[List[nn.Parameter]]:
    """One bucket per module that owns a `_param_names` list (the fused ResidualAttentionBlocks), one for the rest."""
    seen, buckets = set(), []
    for m in model.modules():
        if hasattr(m, "_param_names") and hasattr(m, "attn"):
            ps = [p for p in m.parameters() if p.requires_grad and id(p) not in seen]
            if ps:
                buckets.append(ps)
                seen.update(id(p) for p in ps)
    rest = [p for p in model.parameters() if p.requires_grad and id(p) not in seen]
    if rest:
        buckets.append(rest)
    return buckets


class BlockwiseGradReducer:
    def __init__(self, model: nn.Module, process_group=None, buckets: Optional[Sequence[Iterable[nn.Parameter]]] = None,
                 coalesce: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("BlockwiseGradReducer needs an initialised process group")
        self.group = process_group
        # torch.distributed._coalescing_manager is a private API: without it every tensor gets its own async all-reduce
        self.coalesce = coalesce and hasattr(dist, "_coalescing_manager")
        self.world = dist.get_world_size(process_group)
        self.buckets = [list(b) for b in (buckets if buckets is not None else default_buckets(model))]
        self._left = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._hooks = []
        for i, b in enumerate(self.buckets):
            for p in b:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
        # every rank must start from the same weights, as DDP's constructor guarantees
        with torch.no_grad():
            for p in model.parameters():
                dist.broadcast(p.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
            for b in model.buffers():
                dist.broadcast(b.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)

    def _make_hook(self, i: int):
        def hook(_param):
            self._left[i] -= 1
            if self._left[i] == 0:
                self._launch(i)
        return hook

    def _launch(self, i: int):
        grads = [p.grad for p in self.buckets[i] if p.grad is not None]
        self._launched[i] = True
        if not grads or self.world == 1:
            return
        if self.coalesce:
            with dist._coalescing_manager(group=self.group, async_ops=True) as cm:
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            self._works.append(cm)
        else:
            self._works.extend(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for g in grads)

    def finish(self) -> float:
        """Launch whatever did not complete through the hooks (parameters without a gradient this step), wait for every
        collective in the order it was issued, re-arm, and return the factor that turns the sums into averages."""
        for i in range(len(self.buckets)):
            if not self._launched[i]:
                self._launch(i)
        for w in self._works:
            w.wait()
        self._works.clear()
        self._left = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        return 1.0 / self.world

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()
