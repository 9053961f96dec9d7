"""Event-timed attention forward / backward at the bench's three shapes (medium, 32 clips), for A/B experiments
(e.g. OASR_DEBUG_ATTN_BWD=1 skips the dQ reduce-add to expose its cost)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from olmoasr_b200 import kernels as K  # noqa: E402

B, H = 32, 16
d = H * 64


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, Tq, Tkv, causal in (("encoder self", 1500, 1500, False), ("cross", 448, 1500, False), ("decoder self", 448, 448, True)):
    torch.manual_seed(0)
    q = torch.randn(B * Tq, d, device="cuda").bfloat16()
    kv = torch.randn(B * Tkv, 2 * d, device="cuda").bfloat16()
    k, v = kv[:, :d], kv[:, d:]
    dout = torch.randn(B * Tq, d, device="cuda").bfloat16()
    kv_len = torch.full((B,), 200, device="cuda", dtype=torch.int32) if causal else None
    o, lse = K.attention_fwd(q, k, v, B, H, Tq, Tkv, causal=causal, kv_len=kv_len)
    fwd_ms = timeit(lambda: K.attention_fwd(q, k, v, B, H, Tq, Tkv, causal=causal, kv_len=kv_len))
    bwd_ms = timeit(lambda: K.attention_bwd(q, k, v, o, dout, lse, B, H, Tq, Tkv, causal=causal, kv_len=kv_len))
    fl = 4.0 * B * H * Tq * Tkv * 64 * (0.5 if causal else 1.0)
    print(f"{name:13s} fwd {fwd_ms:7.3f} ms ({fl / fwd_ms / 1e9:5.0f} TFLOP/s)   bwd {bwd_ms:7.3f} ms ({2.5 * fl / bwd_ms / 1e9:5.0f} TFLOP/s)")
