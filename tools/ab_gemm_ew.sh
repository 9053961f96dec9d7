for ew in 8 16; do
  echo "== EW=$ew"
  export OASR_GEMM_EW=$ew
  for e in bf16 gelu resid; do timeout 120 python tools/one_gemm.py 48000 4096 1024 --epi $e; done
  timeout 120 python tools/one_gemm.py 48000 4096 1024 --epi gelu_bwd --b-mn
  timeout 120 python tools/one_gemm.py 48000 1024 4096 --epi resid
  timeout 120 python tools/one_gemm.py 48000 1024 1024 --epi resid
  timeout 120 python tools/one_gemm.py 14336 1024 1024 --epi resid
  timeout 120 python tools/one_gemm.py 4096 1024 48000 --epi f32 --a-mn --b-mn
done
