#!/bin/bash
# validation after the attention rework: the whole GPU suite in ONE process, smoke(), the default bench line, the RTF line, step profile
mkdir -p gpurun_out
rm -f gpurun_out/m_*
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/m_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
timeout 900 python bench.py --metric rtf > gpurun_out/m_rtf.json 2> gpurun_out/m_rtf.err
timeout 600 python tools/profile_step.py --serial > gpurun_out/m_step_profile_serial.txt 2>&1
tail -4 gpurun_out/m_pytest_gpu.txt; cat gpurun_out/m_smoke.txt | tail -2
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/m_bench.json") if l.startswith("{")][-1])
print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_baseline"]["value"], d["cpu_baseline"]["value"], d["clocks"])
t = json.loads([l for l in open("gpurun_out/m_rtf.json") if l.startswith("{")][-1])
print("rtf", t["value"], [(s["n_clips"], round(s["ms_per_step"], 3)) for s in t["config"]["sweep"]], t["roofline"]["frac"])
PY
head -16 gpurun_out/m_step_profile_serial.txt
