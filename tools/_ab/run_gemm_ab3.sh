#!/bin/bash
for lib in tools/_ab/libllo_cuda_prev.so ""; do
  export LLO_CUDA_LIB=$lib
  echo "=== lib=${lib:-current}"
  for mnk in 65536,1024,784:nn 65536,1024,1024:nn,nt 784,1024,65536:tn 1024,1024,65536:tn 8192,1024,784:nn 8192,1024,1024:nn,nt 784,1024,8192:tn 1024,1024,8192:tn 16384,1024,1024:nn 32768,1024,784:nn 784,1024,32768:tn 1024,1024,16384:tn 8192,8192,8192:nn; do
    python tools/gemm_bench.py --mnk ${mnk%%:*} --layouts ${mnk##*:} --iters 5 | grep TFLOP
  done
done
unset LLO_CUDA_LIB
python tools/gemm_bench.py --check --size 256 --iters 1 | tail -2
