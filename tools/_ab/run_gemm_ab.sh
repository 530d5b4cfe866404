#!/bin/bash
# same-box A/B of the fp32 GEMM: round-1 FFMA kernel vs the FFMA2 kernel
set -x
for lib in tools/_ab/libllo_cuda_r1sgemm.so ""; do
  export LLO_CUDA_LIB=$lib
  echo "=== lib=${lib:-current}"
  python tools/gemm_bench.py --size 8192 --layouts nn,nt,tn,tt --iters 5
  python tools/gemm_bench.py --mnk 65536,1024,784 --layouts nn --iters 5
  python tools/gemm_bench.py --mnk 65536,1024,1024 --layouts nn,nt --iters 5
  python tools/gemm_bench.py --mnk 784,1024,65536 --layouts tn --iters 5
  python tools/gemm_bench.py --mnk 1024,1024,65536 --layouts tn --iters 5
  python tools/gemm_bench.py --mnk 8192,1024,784 --layouts nn --iters 5
  python tools/gemm_bench.py --mnk 8192,1024,1024 --layouts nn,nt --iters 5
  python tools/gemm_bench.py --mnk 784,1024,8192 --layouts tn --iters 5
  python tools/gemm_bench.py --mnk 1024,1024,8192 --layouts tn --iters 5
done
unset LLO_CUDA_LIB
python tools/gemm_bench.py --check --size 256 --iters 1
