"""Checkpoint files in the reference's on-disk format.

Writer / reader for what `save_ckpt` / `load_ckpt` of scripts/training/train_timestamps.py:894-1074 exchange -- one
`torch.save`d dict with the keys `global_step, local_step, epoch, best_eval_wer, model_state_dict, optimizer_state_dict,
scaler_state_dict, scheduler_state_dict, dims` (:930-955), in the `*_ddp.pt` flavour (`module.`-prefixed parameter
names, the state dict of the DDP wrapper) or the `*_non_ddp.pt` flavour -- and the training -> inference conversion of
scripts/eval/gen_inf_ckpt.py:4-11 (drop the padding row 51864 of the tied embedding, `dims` as a plain dict), which
`olmoasr.load_model(..., inference=True)` (olmoasr/__init__.py:147-161) expects.

State dicts are storage-agnostic: a model in slab mode (olmoasr_b200/slab.py) saves ordinary, independent tensors.
"""
from __future__ import annotations

import os
from dataclasses import asdict, is_dataclass
from typing import Any, Dict, Optional

import torch

from .config.model_dims import ModelDimensions

PAD_ROW_VOCAB = 51865  # n_vocab + 1 rows in the training model's embedding (olmoasr/model.py:665-667)


def _plain_state_dict(module: torch.nn.Module, prefix: str = "") -> Dict[str, torch.Tensor]:
    inner = module.module if hasattr(module, "module") and isinstance(module.module, torch.nn.Module) else module
    return {prefix + k: v.detach().to("cpu", copy=True).contiguous() for k, v in inner.state_dict().items()}


def save_ckpt(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, scaler: Any = None,
              scheduler: Any = None, *, global_step: int = 0, local_step: int = 0, epoch: int = 0,
              best_eval_wer: Optional[float] = None, ddp_names: bool = False, dims: Optional[ModelDimensions] = None) -> str:
    """Write one checkpoint file with the reference's keys.  `ddp_names=True` writes the `*_ddp.pt` flavour (every
    parameter name prefixed with `module.`); `model` may be the bare model or a DDP-style wrapper with `.module`."""
    inner = model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model
    ckpt = {
        "global_step": global_step, "local_step": local_step, "epoch": epoch, "best_eval_wer": best_eval_wer,
        "model_state_dict": _plain_state_dict(model, "module." if ddp_names else ""),
        "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else None,
        "scaler_state_dict": scaler.state_dict() if scaler is not None else None,
        "scheduler_state_dict": scheduler.state_dict() if scheduler is not None else None,
        "dims": dims if dims is not None else inner.dims,
    }
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(ckpt, path)
    return path


def load_ckpt(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, scaler: Any = None,
              scheduler: Any = None, map_location="cpu") -> dict:
    """Restore a checkpoint of either flavour into `model` (+ optimizer / scaler / scheduler when given); returns the dict
    (global_step, epoch, ... for the caller's loop, train_timestamps.py:1045-1074)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    state = {k[len("module."):] if k.startswith("module.") else k: v for k, v in ckpt["model_state_dict"].items()}
    inner = model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model
    inner.load_state_dict(state)
    if optimizer is not None and ckpt.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    if scaler is not None and ckpt.get("scaler_state_dict") is not None:
        scaler.load_state_dict(ckpt["scaler_state_dict"])
    if scheduler is not None and ckpt.get("scheduler_state_dict") is not None:
        scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    return ckpt


def gen_inf_ckpt(ckpt_path: str, save_path: str) -> str:
    """Training checkpoint -> inference checkpoint (scripts/eval/gen_inf_ckpt.py:4-11): the tied embedding loses its
    last (padding) row, `dims` becomes a plain dict, parameter names lose a `module.` prefix if there is one."""
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    state = {k[len("module."):] if k.startswith("module.") else k: v for k, v in ckpt["model_state_dict"].items()}
    emb = state["decoder.token_embedding.weight"]
    dims = ckpt["dims"]
    dims = dict(dims) if isinstance(dims, dict) else (asdict(dims) if is_dataclass(dims) else dict(vars(dims)))
    if emb.shape[0] != dims["n_vocab"] + 1:
        raise ValueError(f"not a training checkpoint: embedding has {emb.shape[0]} rows, expected n_vocab + 1 = {dims['n_vocab'] + 1}")
    state["decoder.token_embedding.weight"] = emb[:-1].clone()
    ckpt["model_state_dict"] = state
    ckpt["dims"] = dims
    os.makedirs(os.path.dirname(os.path.abspath(save_path)), exist_ok=True)
    torch.save(ckpt, save_path)
    return save_path
