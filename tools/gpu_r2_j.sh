#!/bin/bash
# Round-2 GPU call J (1 GPU): what the driver runs at round end -- the whole GPU test suite in ONE process, smoke(), the default bench
# line, the reference arm -- plus the RTF line.
mkdir -p gpurun_out
rm -f gpurun_out/j_*
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/j_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err
timeout 900 python bench.py --metric rtf > gpurun_out/j_rtf.json 2> gpurun_out/j_rtf.err
tail -4 gpurun_out/j_pytest_gpu.txt; cat gpurun_out/j_smoke.txt | tail -2
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/j_bench.json") if l.startswith("{")][-1])
print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_baseline"]["value"], d["cpu_baseline"]["value"], d["clocks"])
r = json.loads([l for l in open("gpurun_out/j_bench_ref.json") if l.startswith("{")][-1])
print("ref", r["value"], r["steps"], r["ms_per_step"])
t = json.loads([l for l in open("gpurun_out/j_rtf.json") if l.startswith("{")][-1])
print("rtf", t["value"], [(s["n_clips"], round(s["ms_per_step"], 3)) for s in t["config"]["sweep"]], t["roofline"]["frac"])
PY
