"""Small multi-item invocations of the persistent attention kernels for
`compute-sanitizer --tool memcheck|racecheck python tools/sanitize_attention.py` (grid capped at 2 CTAs so that every CTA crosses
item boundaries: TMA-store staging reuse, barrier parities carried across items, fully masked items)."""
import os
import sys
from pathlib import Path

import torch

os.environ["OASR_ATTN_MAX_CTAS"] = "2"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from olmoasr_b200 import kernels as K  # noqa: E402


def main():
    torch.manual_seed(0)
    for B, H, Tq, Tkv, causal, use_len in ((2, 1, 300, 300, True, True), (1, 2, 200, 400, False, False)):
        d = H * 64
        q = torch.randn(B * Tq, d, device="cuda").bfloat16()
        kv = torch.randn(B * Tkv, 2 * d, device="cuda").bfloat16()
        k, v = kv[:, :d], kv[:, d:]
        dout = torch.randn(B * Tq, d, device="cuda").bfloat16()
        kv_len = torch.tensor([40 + 200 * i for i in range(B)], device="cuda", dtype=torch.int32).clamp(max=Tkv) if use_len else None
        o, lse = K.attention_fwd(q, k, v, B, H, Tq, Tkv, causal=causal, kv_len=kv_len)
        dq, dk, dv = K.attention_bwd(q, k, v, o, dout, lse, B, H, Tq, Tkv, causal=causal, kv_len=kv_len)
        torch.cuda.synchronize()
        print("ok", B, H, Tq, Tkv, float(o.float().abs().mean()), float(dq.float().abs().mean()), float(dk.float().abs().mean()))


if __name__ == "__main__":
    main()
