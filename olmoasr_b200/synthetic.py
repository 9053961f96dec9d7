"""Synthetic batch of SURVEY.md section 8(d): the tensors AudioTextDataset.__getitem__ hands to the training loop
(scripts/training/train_timestamps.py:295-329 -- text_input / text_y padded to 448 with 51864, the dense additive
padding mask) built from random token ids and Gaussian-noise waveforms.  Host-side input generator for bench.py,
smoke() and the tests: pure tensor construction, no model code, nothing computed on the data."""
from __future__ import annotations

import numpy as np
import torch

SOT, NO_TIMESTAMPS, EOT, PAD = 50257, 50362, 50256, 51864
N_TEXT_CTX = 448
N_SAMPLES = 480000


def waveforms(batch: int, rank: int = 0, n_samples: int = N_SAMPLES, int16: bool = False):
    g = torch.Generator().manual_seed(1234 + rank)
    w = torch.clamp(torch.randn(batch, n_samples, generator=g) * 0.1, -1.0, 1.0)
    if int16:  # mirror np.load(int16) / 32768 (train_timestamps.py:196)
        return (w * 32767.0).round().to(torch.int16)
    return w


def text_batch(batch: int, rank: int = 0, n_text_ctx: int = N_TEXT_CTX):
    """Returns text_input (B,448) i64, text_y (B,448) i64, padding_mask (B,448,448) f32, lengths (B,) i64."""
    g = torch.Generator().manual_seed(4321 + rank)
    ti = torch.full((batch, n_text_ctx), PAD, dtype=torch.long)
    ty = torch.full((batch, n_text_ctx), PAD, dtype=torch.long)
    pm = torch.zeros(batch, n_text_ctx, n_text_ctx)
    lens = torch.zeros(batch, dtype=torch.long)
    for i in range(batch):
        L = 32 + (37 * i) % 192
        row = torch.cat([torch.tensor([SOT, NO_TIMESTAMPS]), torch.randint(0, EOT, (L - 3,), generator=g),
                         torch.tensor([EOT])])
        n = L - 1  # len(text_input) == len(text_y)
        ti[i, :n] = row[:-1]
        ty[i, :n] = row[1:]
        pm[i, :, n:] = -np.inf  # train_timestamps.py:314-315
        lens[i] = n
    return ti, ty, pm, lens
