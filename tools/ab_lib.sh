# A/B of two builds of the library (olmoasr_b200/csrc/_ab/*.so) on sustained GEMMs and on the training step
for v in "$@"; do
  export OASR_B200_LIB=$PWD/olmoasr_b200/csrc/_ab/$v.so
  echo "== $v"
  timeout 120 python tools/one_gemm.py 8192 8192 8192 --epi bf16 --seconds 3 | tail -1
  timeout 120 python tools/one_gemm.py 48000 4096 1024 --epi bf16 --seconds 3 | tail -1
  timeout 120 python tools/one_gemm.py 48000 4096 1024 --epi gelu --seconds 3 | tail -1
done
for rep in 1 2; do for v in "$@"; do
  export OASR_B200_LIB=$PWD/olmoasr_b200/csrc/_ab/$v.so
  echo "== $v step"; timeout 300 python tools/ab_step.py SIDE_STREAM=1 --rounds 5 | tail -1
done; done
