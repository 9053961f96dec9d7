"""CPU restatement of the greedy path of whisper/decoding.py over the oracle model (TEST INFRASTRUCTURE).

openai-whisper is an un-vendored, un-pinned dependency of the reference (requirements.txt:21) and is absent here, so
this loop is "parity unpinned": it is written from the published algorithm (DecodingTask._main_loop,
PyTorchInference.logits, GreedyDecoder.update, SuppressBlank, SuppressTokens) and anchored on the reference's call
sites (scripts/eval/eval.py:1846-1847: DecodingOptions(language="en", without_timestamps=True)).  The MODEL calls
inside it (encoder, decoder with the kv-cache protocol of inf_model.py:422-453) are pinned bit-exactly through
oracle/model.py.  Kept deliberately independent of olmoasr_b200/decoding.py (different structure: functional, cache
dict keyed by parameter prefix, no hooks).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import model as OM

EOT, SOT, NO_TIMESTAMPS, NO_SPEECH = 50256, 50257, 50362, 50361
BLANK_IDS = (220, EOT)
SPECIAL_SUPPRESS = (50358, 50357, 50257, 50360, 50359, 50361)  # transcribe, translate, sot, sot_prev, sot_lm, no_speech


def non_speech_tokens() -> Sequence[int]:
    """tokenizer.non_speech_tokens for the gpt2 vocabulary, taken from the copy transformers ships
    (configuration_whisper.NON_SPEECH_TOKENS, entries below the special-token range)."""
    from transformers.models.whisper.configuration_whisper import NON_SPEECH_TOKENS

    return tuple(t for t in NON_SPEECH_TOKENS if t < EOT)


def greedy_decode(sd, dims: OM.Dims, mel: torch.Tensor, sample_len: Optional[int] = None, dtype=torch.float32,
                  without_timestamps: bool = True, return_margins: bool = False):
    """Returns the list of generated token-id lists (up to, not including, eot) for each clip."""
    assert without_timestamps, "the oracle restates the short-form eval configuration only"
    # DecodingTask feeds a half-precision mel into fp32 weights; Linear / Conv1d cast their weights per call
    # (inf_model.py:56-60), LayerNorm stays fp32 -- so only the input is cast here
    sdc = sd
    xa = OM.encoder_forward(sdc, dims, mel.to(dtype), sdpa=False)
    n = mel.shape[0]
    initial = [SOT, NO_TIMESTAMPS]
    tokens = torch.tensor([initial]).repeat(n, 1)
    suppress = sorted(set(non_speech_tokens()) | set(SPECIAL_SUPPRESS))
    cache: dict = {}
    sample_len = sample_len or dims.n_text_ctx // 2
    margins = []
    for i in range(sample_len):
        inp = tokens if tokens.shape[1] <= len(initial) else tokens[:, -1:]
        logits = OM.decoder_forward(sdc, dims, inp, xa, None, train_model=False, cache=cache)[:, -1]
        if tokens.shape[1] == len(initial):
            logits[:, list(BLANK_IDS)] = -np.inf
        logits[:, suppress] = -np.inf
        top2 = logits.float().topk(2, dim=-1).values
        margins.append((top2[:, 0] - top2[:, 1]).min().item())
        nxt = logits.argmax(dim=-1)
        nxt[tokens[:, -1] == EOT] = EOT
        tokens = torch.cat([tokens, nxt[:, None]], dim=-1)
        if bool((tokens[:, -1] == EOT).all()) or tokens.shape[1] > dims.n_text_ctx:
            break
    out: List[List[int]] = []
    for k in range(n):
        seq = tokens[k, len(initial):].tolist()
        out.append(seq[: seq.index(EOT)] if EOT in seq else seq)
    return (out, margins) if return_margins else out
