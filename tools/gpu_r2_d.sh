#!/bin/bash
# Round-2 GPU call D (1 GPU): decode-engine iteration -- tests, step profile with and without programmatic dependent launch, RTF sweep.
mkdir -p gpurun_out
rm -f gpurun_out/d_*
timeout 900 python -m pytest tests/test_decode_engine_gpu.py tests/test_decode_gpu.py tests/test_optimizer_gpu.py -m gpu -q -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/d_tests.log
OASR_DEC_PDL=0 timeout 400 python tools/profile_decode.py > gpurun_out/d_decode_profile_nopdl.txt 2>&1
timeout 400 python tools/profile_decode.py > gpurun_out/d_decode_profile_pdl.txt 2>&1
timeout 900 python bench.py --metric rtf --steps 2 --warmup 1 > gpurun_out/d_rtf.json 2> gpurun_out/d_rtf.err
OASR_DEC_PDL=0 timeout 900 python bench.py --metric rtf --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/d_rtf_nopdl.json 2> gpurun_out/d_rtf_nopdl.err
timeout 900 python bench.py --metric rtf --dtype bf16 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/d_rtf_bf16.json 2> gpurun_out/d_rtf_bf16.err
timeout 300 python tools/time_membound.py > gpurun_out/d_membound.txt 2>&1
grep -E "passed|failed|rel-L2|margin|Error" gpurun_out/d_tests.log
grep -E "^#" gpurun_out/d_decode_profile_nopdl.txt gpurun_out/d_decode_profile_pdl.txt
python - <<'PY'
import json
for f in ("d_rtf", "d_rtf_nopdl", "d_rtf_bf16"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, [(s["n_clips"], round(s["ms_per_step"], 3), round(s["rtf"], 5), round(s["hbm_frac"], 3)) for s in d["config"]["sweep"]])
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/d_membound.txt | tail -20
