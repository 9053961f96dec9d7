for cl in 2 4; do
  echo "== OASR_GEMM_CLUSTER=$cl"
  export OASR_GEMM_CLUSTER=$cl OASR_GEMM_VERBOSE=1
  for e in bf16 gelu resid; do timeout 120 python tools/one_gemm.py 48000 4096 1024 --epi $e 2>&1 | tail -2; done
  timeout 120 python tools/one_gemm.py 48000 4096 1024 --epi gelu_bwd --b-mn 2>&1 | tail -2
  timeout 120 python tools/one_gemm.py 48000 1024 4096 --epi resid 2>&1 | tail -1
  timeout 120 python tools/one_gemm.py 14336 1024 1024 --epi resid 2>&1 | tail -1
  timeout 120 python tools/one_gemm.py 4096 1024 48000 --epi f32 --a-mn --b-mn 2>&1 | tail -2
  timeout 120 python tools/one_gemm.py 8192 8192 8192 --epi bf16 --seconds 3 2>&1 | tail -1
done
