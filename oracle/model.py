"""torch-CPU restatement of the reference model (olmoasr/model.py, olmoasr/inf_model.py).

TEST INFRASTRUCTURE.  Written as plain functions over a state dict plus a thin module that reproduces the
reference's parameter creation ORDER (so `torch.manual_seed(s)` gives bit-identical initial weights):

  * every Linear / Conv1d draws kaiming_normal_(fan_in, relu) in its constructor (model.py:81, :171);
    the TRAIN model's MultiHeadAttention re-draws query/key/value/out a second time (model.py:258-264),
    the inference model does not (inf_model.py:131-138);
  * token_embedding: nn.Embedding init then kaiming (model.py:665-670); decoder positional embedding:
    kaiming in the train model (model.py:671-675), torch.empty in the inference model (inf_model.py:307).

Numerics restated (and checked bit-for-bit against the unmodified reference in tests/test_oracle_pin.py):
  LayerNorm in fp32 then cast back (model.py:25-39); Linear/Conv1d cast weights to the activation dtype
  (model.py:97-101,193-195); conv stem + exact-erf GELU + sinusoid add (model.py:592-602); 3-D mask -> SDPA
  branch, 2-D mask -> manual qkv_attention with hd**-0.25 on q and k and fp32 softmax
  (model.py:317-340, 347-442; inf_model.py:172-196); tied logits `.float()` (model.py:768-770).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn


@dataclass
class Dims:  # olmoasr/config/model_dims.py:4-25
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


def variant_dims(name: str) -> Dims:  # model_dims.py:28-89
    d, h, l = {"tiny": (384, 6, 4), "base": (512, 8, 6), "small": (768, 12, 12),
               "medium": (1024, 16, 24), "large": (1280, 20, 32)}[name]
    return Dims(80, 1500, d, h, l, 51864, 448, d, h, l)


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> Tensor:  # model.py:199-230
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1)


# ------------------------------------------------------------------------------------------------
# parameter creation in reference order
# ------------------------------------------------------------------------------------------------
def _kaiming(t: Tensor) -> Tensor:
    return nn.init.kaiming_normal_(t, mode="fan_in", nonlinearity="relu")


def _linear(sd, prefix, n_in, n_out, bias=True):
    m = nn.Linear(n_in, n_out, bias=bias)  # default init consumes RNG exactly like the reference's super().__init__
    _kaiming(m.weight)
    sd[prefix + ".weight"] = m.weight.detach()
    if bias:
        sd[prefix + ".bias"] = m.bias.detach()


def _mha(sd, prefix, d, train: bool):
    for name, bias in (("query", True), ("key", False), ("value", True), ("out", True)):
        _linear(sd, f"{prefix}.{name}", d, d, bias)
        if train:  # model.py:258-264 re-initialises each projection right after constructing it
            _kaiming(sd[f"{prefix}.{name}.weight"])


def _ln(sd, prefix, d):
    sd[prefix + ".weight"] = torch.ones(d)
    sd[prefix + ".bias"] = torch.zeros(d)


def _block(sd, prefix, d, cross: bool, train: bool):
    _mha(sd, prefix + ".attn", d, train)
    _ln(sd, prefix + ".attn_ln", d)
    if cross:
        _mha(sd, prefix + ".cross_attn", d, train)
        _ln(sd, prefix + ".cross_attn_ln", d)
    _linear(sd, prefix + ".mlp.0", d, 4 * d)
    _linear(sd, prefix + ".mlp.2", 4 * d, d)
    _ln(sd, prefix + ".mlp_ln", d)


def init_state_dict(dims: Dims, seed: int = 0, train: bool = True) -> Dict[str, Tensor]:
    """State dict with the reference's names, shapes and (for a given seed) VALUES."""
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    d = dims.n_audio_state
    for name, cin, k in (("conv1", dims.n_mels, 3), ("conv2", d, 3)):
        m = nn.Conv1d(cin, d, kernel_size=k, padding=1)
        _kaiming(m.weight)
        sd[f"encoder.{name}.weight"] = m.weight.detach()
        sd[f"encoder.{name}.bias"] = m.bias.detach()
    sd["encoder.positional_embedding"] = sinusoids(dims.n_audio_ctx, d)
    for i in range(dims.n_audio_layer):
        _block(sd, f"encoder.blocks.{i}", d, cross=False, train=train)
    _ln(sd, "encoder.ln_post", d)
    dt = dims.n_text_state
    n_rows = dims.n_vocab + 1 if train else dims.n_vocab
    emb = nn.Embedding(n_rows, dt, padding_idx=(51864 if dims.n_vocab == 51864 else 51865) if train else None)
    _kaiming(emb.weight)  # overwrites the zeroed padding row too (model.py:668-670)
    sd["decoder.token_embedding.weight"] = emb.weight.detach()
    pos = torch.empty(dims.n_text_ctx, dt)
    if train:
        _kaiming(pos)
    else:
        pos.zero_()  # inf_model.py:307 leaves it uninitialised; checkpoints always overwrite it
    sd["decoder.positional_embedding"] = pos
    for i in range(dims.n_text_layer):
        _block(sd, f"decoder.blocks.{i}", dt, cross=True, train=train)
    _ln(sd, "decoder.ln", dt)
    return sd


# ------------------------------------------------------------------------------------------------
# forward restatement
# ------------------------------------------------------------------------------------------------
def _lin(sd, p, x):
    w = sd[p + ".weight"].to(x.dtype)
    b = sd.get(p + ".bias")
    return F.linear(x, w, None if b is None else b.to(x.dtype))


def _layer_norm(sd, p, x):
    return F.layer_norm(x.float(), (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5).type(x.dtype)


def _split_heads(t, n_head):
    return t.view(*t.shape[:2], n_head, -1).permute(0, 2, 1, 3)


def _manual_attention(q, k, v, n_head, mask):
    """qkv_attention: scale q and k by hd**-0.25 in the activation dtype, fp32 softmax, probabilities cast
    back (model.py:347-442 / inf_model.py:172-196)."""
    n_ctx = q.shape[1]
    scale = (q.shape[-1] // n_head) ** -0.25
    qh = _split_heads(q, n_head) * scale
    kh = k.view(*k.shape[:2], n_head, -1).permute(0, 2, 3, 1) * scale
    vh = _split_heads(v, n_head)
    qk = qh @ kh
    if mask is not None:
        if mask.dim() == 2:
            qk = qk + mask
        else:
            qk = qk + mask.unsqueeze(1)[:, :, :n_ctx, :n_ctx]
    w = F.softmax(qk.float(), dim=-1).to(qh.dtype)
    return (w @ vh).permute(0, 2, 1, 3).flatten(start_dim=2)


def _attention(sd, p, n_head, x, xa=None, mask=None, sdpa=True, cache=None):
    q = _lin(sd, p + ".query", x)
    src = x if xa is None else xa
    if cache is not None and xa is not None and (p + ".key") in cache:
        k, v = cache[p + ".key"], cache[p + ".value"]
    else:
        k, v = _lin(sd, p + ".key", src), _lin(sd, p + ".value", src)
        if cache is not None:  # install_kv_cache_hooks semantics (model.py:925-964)
            if xa is not None or (p + ".key") not in cache:
                cache[p + ".key"], cache[p + ".value"] = k, v
            else:
                cache[p + ".key"] = k = torch.cat([cache[p + ".key"], k], dim=1)
                cache[p + ".value"] = v = torch.cat([cache[p + ".value"], v], dim=1)
    manual = (not sdpa) or (mask is not None and mask.dim() == 2)
    if manual:
        wv = _manual_attention(q, k, v, n_head, mask)
    else:  # training branch: F.scaled_dot_product_attention with the (B,1,S,S) additive mask (model.py:327-340)
        m = None if mask is None else mask.unsqueeze(1)
        wv = F.scaled_dot_product_attention(_split_heads(q, n_head), _split_heads(k, n_head),
                                            _split_heads(v, n_head), attn_mask=m)
        wv = wv.permute(0, 2, 1, 3).flatten(start_dim=2)
    return _lin(sd, p + ".out", wv)


def _res_block(sd, p, n_head, x, xa=None, mask=None, sdpa=True, cache=None):
    x = x + _attention(sd, p + ".attn", n_head, _layer_norm(sd, p + ".attn_ln", x), mask=mask, sdpa=sdpa, cache=cache)
    if xa is not None:
        x = x + _attention(sd, p + ".cross_attn", n_head, _layer_norm(sd, p + ".cross_attn_ln", x), xa, sdpa=sdpa,
                           cache=cache)
    h = _lin(sd, p + ".mlp.0", _layer_norm(sd, p + ".mlp_ln", x))
    return x + _lin(sd, p + ".mlp.2", F.gelu(h))


def encoder_forward(sd, dims: Dims, mel: Tensor, sdpa: bool = True) -> Tensor:
    """AudioEncoder.forward (model.py:571-623)."""
    w1, b1 = sd["encoder.conv1.weight"], sd["encoder.conv1.bias"]
    w2, b2 = sd["encoder.conv2.weight"], sd["encoder.conv2.bias"]
    x = F.gelu(F.conv1d(mel, w1.to(mel.dtype), b1.to(mel.dtype), padding=1))
    x = F.gelu(F.conv1d(x, w2.to(x.dtype), b2.to(x.dtype), stride=2, padding=1))
    x = x.permute(0, 2, 1)
    assert x.shape[1:] == sd["encoder.positional_embedding"].shape, "incorrect audio shape"
    x = (x + sd["encoder.positional_embedding"]).to(x.dtype)
    for i in range(dims.n_audio_layer):
        x = _res_block(sd, f"encoder.blocks.{i}", dims.n_audio_head, x, sdpa=sdpa)
    return _layer_norm(sd, "encoder.ln_post", x)


def causal_mask(n: int) -> Tensor:
    return torch.empty(n, n).fill_(-np.inf).triu_(1)


def decoder_forward(sd, dims: Dims, tokens: Tensor, xa: Tensor, padding_mask: Optional[Tensor] = None,
                    train_model: bool = True, cache: Optional[dict] = None) -> Tensor:
    """TextDecoder.forward.  train_model=True follows model.py:688-775 (3-D mask -> SDPA, else the 2-D manual
    branch); False follows inf_model.py:320-362 (always manual)."""
    offset = 0
    if cache:
        offset = next(iter(cache.values())).shape[1]
    x = F.embedding(tokens, sd["decoder.token_embedding.weight"]) + \
        sd["decoder.positional_embedding"][offset: offset + tokens.shape[-1]]
    x = x.to(xa.dtype)
    n_ctx = x.shape[1]
    if padding_mask is not None:
        mask = padding_mask + causal_mask(dims.n_text_ctx if train_model else n_ctx)
    else:
        mask = causal_mask(dims.n_text_ctx)[:n_ctx, :n_ctx] if train_model else causal_mask(n_ctx)
    mask = mask.to(x.device)   # the reference moves its mask to the activations' device too (inf_model.py:341-352)
    for i in range(dims.n_text_layer):
        x = _res_block(sd, f"decoder.blocks.{i}", dims.n_text_head, x, xa, mask=mask, sdpa=train_model, cache=cache)
    x = _layer_norm(sd, "decoder.ln", x)
    return (x @ sd["decoder.token_embedding.weight"].to(x.dtype).t()).float()


def model_forward(sd, dims: Dims, mel, tokens, padding_mask=None, train_model=True, autocast_dtype=None):
    """OLMoASR.forward (model.py:856-887) optionally under torch.autocast('cpu', dtype) -- the CPU twin of the
    `with autocast(device_type="cuda", dtype=precision)` in train_timestamps.py:1414."""
    if autocast_dtype is None:
        xa = encoder_forward(sd, dims, mel, sdpa=train_model)
        return decoder_forward(sd, dims, tokens, xa, padding_mask, train_model)
    with torch.autocast("cpu", dtype=autocast_dtype):
        xa = encoder_forward(sd, dims, mel, sdpa=train_model)
        return decoder_forward(sd, dims, tokens, xa, padding_mask, train_model)


def token_ce(logits: Tensor, targets: Tensor, ignore_index: int = 51864) -> Tensor:
    """train_timestamps.py:1444-1448."""
    return F.cross_entropy(logits.view(-1, logits.shape[-1]), targets.view(-1), ignore_index=ignore_index)
