"""Phase-by-phase cycle trace of ONE attention-backward CTA (debug build only).

    python -m olmoasr_b200.build --variant attn_trace -DOASR_ATTN_TRACE
    OASR_B200_LIB=olmoasr_b200/csrc/_ab/attn_trace.so python tools/trace_attention.py

Roles: 0 = compute warp 0 (lane quarter 0, key columns 0..63), 1 = MMA warp, 2 = TMA warp, 3 = dQ-drain warp 12.  Every line
is (event id, cycles since the previous event of that role, cycles since the first event of the CTA)."""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from olmoasr_b200 import _lib, kernels as K  # noqa: E402

NAMES = {
    0: {1: "start", 10: "loop top", 11: "S,dP ready", 12: "S,dP in regs", 13: "P,dS math done", 20: " dq: MMAs retired", 21: " dq: in regs",
        22: " dq: bar1", 23: " dq: staged", 14: "dQ drained", 15: "P,dS stored", 30: "last dQ out", 31: "dK,dV ready", 32: "dK,dV stored"},
    1: {1: "start", 2: "K,V landed", 3: "Q0 landed", 4: "S,dP(0) issued", 5: "Q(next) landed", 6: "S,dP regs free", 7: "S,dP(next) issued",
        8: "P,dS in smem", 9: "dV,dK,dQ issued"},
    2: {1: "start", 2: "stage free"},
}
NAMES[3] = {20: "dQ MMAs retired", 21: "dQ staged in smem", 22: "reduce-add issued"}
# forward: roles 0 / 3 = softmax warps 0 / 4 (query tiles 0 / 1), 1 = MMA warp
FWD = {
    0: {1: "start", 10: "loop top", 11: "S ready", 12: "S in regs", 13: "row max done", 14: " P V(j-1) retired", 15: "O rescaled", 16: "P stored",
        30: "O written"},
    1: {1: "start", 2: "QK(0) issued", 3: "K,V(next) landed", 4: "S0 regs free", 5: "QK0(next) issued", 6: "S1 regs free", 7: "QK1(next) issued",
        8: "P0 ready", 9: "PV0 issued", 10: "P1 ready", 11: "PV1 issued"},
    2: {},
}
FWD[3] = FWD[0]


def dump(fn, names):
    buf = (ctypes.c_ulonglong * (4 * 512))()
    assert fn(buf) == 0
    ev = {}
    for role in range(4):
        n = buf[role * 512]
        ev[role] = [((buf[role * 512 + 1 + i] >> 48), buf[role * 512 + 1 + i] & ((1 << 48) - 1)) for i in range(n)]
    t0 = min(e[0][1] for e in ev.values() if e)
    for role in (0, 1, 2, 3):
        if not ev[role]:
            continue
        print(f"-- role {role}")
        prev = None
        for eid, t in ev[role]:
            print(f"   {names[role].get(eid, eid):24s} +{(t - prev) if prev is not None else 0:6d}   @{t - t0:7d}")
            prev = t


B, H = 32, 16
d = H * 64
lib = ctypes.CDLL(str(_lib._LIB_PATH))
for name, Tq, Tkv, causal in (("encoder self", 1500, 1500, False), ("cross", 448, 1500, False)):
    torch.manual_seed(0)
    q = torch.randn(B * Tq, d, device="cuda").bfloat16()
    kv = torch.randn(B * Tkv, 2 * d, device="cuda").bfloat16()
    k, v = kv[:, :d], kv[:, d:]
    dout = torch.randn(B * Tq, d, device="cuda").bfloat16()
    o, lse = K.attention_fwd(q, k, v, B, H, Tq, Tkv, causal=causal, kv_len=None)
    for _ in range(3):
        K.attention_bwd(q, k, v, o, dout, lse, B, H, Tq, Tkv, causal=causal, kv_len=None)
    torch.cuda.synchronize()
    print(f"==== backward {name}: Tq {Tq} Tkv {Tkv}")
    dump(lib.oasr_debug_bwd_trace, NAMES)
    print(f"==== forward {name}: Tq {Tq} Tkv {Tkv}")
    dump(lib.oasr_debug_fwd_trace, FWD)
