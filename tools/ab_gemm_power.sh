export OASR_GEMM_EW=${OASR_GEMM_EW:-8}
for e in bf16 gelu resid; do timeout 120 python tools/one_gemm.py 48000 4096 1024 --epi $e --seconds 3; done
timeout 120 python tools/one_gemm.py 48000 4096 1024 --epi gelu_bwd --b-mn --seconds 3
timeout 120 python tools/one_gemm.py 48000 1024 4096 --epi bf16 --seconds 3
timeout 120 python tools/one_gemm.py 8192 8192 8192 --epi bf16 --seconds 3
