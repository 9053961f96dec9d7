"""The reference's GPU path: the same training step written with stock PyTorch ops only (nn.Linear / Conv1d /
F.layer_norm in fp32 / F.scaled_dot_product_attention with the reference's dense float mask / F.cross_entropy on fp32
logits / clip_grad_norm_ / fused torch AdamW / DistributedDataParallel), under torch.autocast(bf16) -- i.e. what running
the reference's model code (olmoasr/model.py:445-528,560-602,688-775 under scripts/training/train_timestamps.py:1414,
1440-1454,1508-1522,2330) costs on this GPU with cuBLASLt / cuDNN / ATen SDPA.  SURVEY 8(d) last row: the number the
headline has to beat.  Independent of oracle/ and of olmoasr_b200's kernels (it only borrows the synthetic batch and the
mel filterbank table); `bench.py` reports it as `gpu_baseline` and `bench.py --impl torch_gpu` prints it alone.

    python -m baseline.torch_stock [--variant medium] [--batch 32] [--steps 3]
"""
import argparse
import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F
from torch import nn

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def _f(x):
    return float(x.detach()) if torch.is_tensor(x) else float(x)


def run(variant: str, batch: int, steps: int, warmup: int, device, rank: int = 0, world: int = 1, e2e: bool = False):
    """Timed stock-PyTorch training steps on `device` (DDP when world > 1).  Returns dict(ms_per_step, clips_per_s (this
    rank's batch / step time x world), loss, peak_mem_gib).  Inputs are resident on the device unless e2e."""
    import olmoasr_b200 as ob
    from olmoasr_b200 import audio as A
    from olmoasr_b200 import synthetic as synth

    torch.manual_seed(0)
    dims = ob.VARIANT_TO_DIMS[variant]
    model = Model(dims).to(device)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], output_device=device.index)
    opt = torch.optim.AdamW(model.parameters(), lr=1.5e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.1, fused=device.type == "cuda")
    wav_h = synth.waveforms(batch, rank=rank, int16=True)
    ti_h, ty_h, pm_h, _ = synth.text_batch(batch, rank=rank)
    if device.type == "cuda":
        wav_h, ti_h, ty_h, pm_h = (t.pin_memory() for t in (wav_h, ti_h, ty_h, pm_h))
    dev_in = [t.to(device) for t in (wav_h, ti_h, ty_h, pm_h)]
    filters = torch.as_tensor(A.mel_filterbank(80), dtype=torch.float32, device=device)

    def step(wav, ti, ty, pm):
        mel = log_mel_torch(wav.float() / 32768.0, filters)     # the reference runs this on CPU workers; here on the GPU
        with torch.autocast(device.type, dtype=torch.bfloat16):
            logits = net(mel, ti, pm)
            loss = F.cross_entropy(logits.view(-1, logits.shape[-1]), ty.view(-1), ignore_index=51864)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    def one():
        if e2e:
            return step(*[t.to(device, non_blocking=True) for t in (wav_h, ti_h, ty_h, pm_h)]).item()
        return step(*dev_in)

    for _ in range(warmup):
        loss = one()
    if device.type != "cuda":
        import time
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = one()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        return dict(ms_per_step=ms, clips_per_s=batch * world / ms * 1e3, loss=_f(loss), peak_mem_gib=0.0)
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = one()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = dict(ms_per_step=ms, clips_per_s=batch * world / ms * 1e3, loss=_f(loss), peak_mem_gib=torch.cuda.max_memory_allocated() / 2**30)
    del net, model, opt
    return out


class LN(nn.LayerNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class MHA(nn.Module):
    def __init__(self, d, h):
        super().__init__()
        self.h = h
        self.query, self.key, self.value, self.out = nn.Linear(d, d), nn.Linear(d, d, bias=False), nn.Linear(d, d), nn.Linear(d, d)

    def forward(self, x, xa=None, mask=None):
        B, T, d = x.shape
        src = x if xa is None else xa
        q = self.query(x).view(B, T, self.h, -1).permute(0, 2, 1, 3)
        k = self.key(src).view(B, src.shape[1], self.h, -1).permute(0, 2, 1, 3)
        v = self.value(src).view(B, src.shape[1], self.h, -1).permute(0, 2, 1, 3)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
        return self.out(o.permute(0, 2, 1, 3).reshape(B, T, d))


class Block(nn.Module):
    def __init__(self, d, h, cross):
        super().__init__()
        self.attn, self.attn_ln = MHA(d, h), LN(d)
        self.cross_attn = MHA(d, h) if cross else None
        self.cross_attn_ln = LN(d) if cross else None
        self.mlp = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Linear(4 * d, d))
        self.mlp_ln = LN(d)

    def forward(self, x, xa=None, mask=None):
        x = x + self.attn(self.attn_ln(x), mask=mask)
        if self.cross_attn is not None:
            x = x + self.cross_attn(self.cross_attn_ln(x), xa)
        return x + self.mlp(self.mlp_ln(x))


def sinusoids(length, channels, max_timescale=10000):
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


class Model(nn.Module):
    def __init__(self, dims):
        super().__init__()
        d = dims.n_audio_state
        self.conv1, self.conv2 = nn.Conv1d(dims.n_mels, d, 3, padding=1), nn.Conv1d(d, d, 3, stride=2, padding=1)
        self.register_buffer("pos_a", sinusoids(dims.n_audio_ctx, d))
        self.enc = nn.ModuleList(Block(d, dims.n_audio_head, False) for _ in range(dims.n_audio_layer))
        self.ln_post = LN(d)
        dt = dims.n_text_state
        self.tok = nn.Embedding(dims.n_vocab + 1, dt)
        self.pos_t = nn.Parameter(torch.randn(dims.n_text_ctx, dt) * 0.01)
        self.dec = nn.ModuleList(Block(dt, dims.n_text_head, True) for _ in range(dims.n_text_layer))
        self.ln = LN(dt)
        nn.init.normal_(self.tok.weight, std=(2.0 / dt) ** 0.5)   # kaiming_normal_(fan_in) like the reference (model.py:668-675)
        self.register_buffer("causal", torch.full((dims.n_text_ctx, dims.n_text_ctx), float("-inf")).triu_(1), persistent=False)

    def forward(self, mel, tokens, padding_mask):
        x = F.gelu(self.conv2(F.gelu(self.conv1(mel)))).permute(0, 2, 1)
        x = (x + self.pos_a).to(x.dtype)
        for b in self.enc:
            x = b(x)
        xa = self.ln_post(x)
        S = tokens.shape[1]
        h = (self.tok(tokens) + self.pos_t[:S]).to(xa.dtype)
        full = (padding_mask + self.causal)[:, None, :S, :S]      # dense float mask, as the reference builds it
        for b in self.dec:
            h = b(h, xa, mask=full)
        h = self.ln(h)
        return (h @ self.tok.weight.to(h.dtype).T).float()


def log_mel_torch(wav, filters):
    """whisper.audio.log_mel_spectrogram on the device (the reference runs it in CPU workers)."""
    st = torch.stft(wav, 400, 160, window=torch.hann_window(400, device=wav.device), return_complex=True)
    mag = st[..., :-1].abs() ** 2
    spec = torch.clamp(filters @ mag, min=1e-10).log10()
    spec = torch.maximum(spec, spec.amax(dim=(1, 2), keepdim=True) - 8.0)
    return (spec + 4.0) / 4.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="medium")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cpu" if args.cpu else "cuda:0")
    r = run(args.variant, args.batch, args.steps, args.warmup, dev)
    print(f"stock PyTorch ({torch.__version__}) {args.variant} B={args.batch} bf16 autocast: {r['ms_per_step']:.1f} ms/step = "
          f"{r['clips_per_s']:.1f} clips/s, loss {r['loss']:.4f}, peak memory {r['peak_mem_gib']:.1f} GiB")


if __name__ == "__main__":
    main()
