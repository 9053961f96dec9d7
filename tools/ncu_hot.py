"""Top stall sites of a kernel from an ncu report's source page (SASS view).

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [N] [section]

The csv holds one section per profiled launch ("Kernel Name" row, header row, one row per SASS instruction).
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
lo, hi = starts[want], starts[want + 1]
print("section", want, "of", len(starts) - 1, ":", rows[lo][1][:100])
hdr = rows[lo + 1]
body = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ci[h]] or 0) for r in body) for h in stall_cols}
print("total samples", tot)
print("stall breakdown:", ", ".join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
top = sorted(enumerate(body), key=lambda ir: -int(ir[1][ci["# Samples"]] or 0))[:n]
for idx, r in sorted(top):
    s = int(r[ci["# Samples"]])
    reasons = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{idx:5d} {100*s/tot:5.1f}%  {r[ci['Source']].strip()[:90]:90s} {reasons}")
