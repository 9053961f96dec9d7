"""Summarise an ncu launch list (gpu__time_duration.sum per launch) of bench.py into per-kernel shares of ONE training
step (delimited by consecutive logmel_frames_kernel launches).

    python tools/ncu_launch_summary.py gpurun_out/launches_bench.csv > profiles/r01_ncu_launch_list_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
names = [r[4] for r in rows]
vals = [float(r[14].replace(",", "")) for r in rows]
unit = rows[0][13]
starts = [i for i, n in enumerate(names) if "logmel_frames_kernel" in n]
print(f"# {len(rows)} launches captured, unit {unit}; logmel launches (step starts) at {starts}")
if len(starts) >= 2:
    lo, hi = starts[-2], starts[-1]
else:
    lo, hi = (starts[0] if starts else 0), len(rows)
print(f"# step = launches [{lo}, {hi}) -> {hi - lo} launches")


def short(n):
    m = re.search(r"(gemm_tcgen05_kernel<[^>]*>|attention_fwd_kernel|attention_bwd_kernel|[a-z0-9_]+_kernel)", n)
    return m.group(1) if m else n[:60]


agg = defaultdict(lambda: [0.0, 0])
for n, v in zip(names[lo:hi], vals[lo:hi]):
    k = short(n)
    agg[k][0] += v
    agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1e-6)
print(f"# sum of kernel durations in the step: {tot * scale:.2f} ms (cold-cache, serialised: compare shares, not absolutes)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k[:70]:70s} {c:8d} {t * scale:10.3f} {100 * t / tot:6.2f}%")
