#!/bin/bash
for rl in 2048 512; do for ps in "" 0 1; do
  export LLO_GEMM_RELAYOUT_MIN=$rl
  if [ -n "$ps" ]; then export LLO_GEMM_PERSIST=$ps; else unset LLO_GEMM_PERSIST; fi
  echo "=== relayout_min=$rl persist=${ps:-auto}"
  python tools/gemm_bench.py --mnk 65536,1024,784 --layouts nn --iters 5 | grep TFLOP
  python tools/gemm_bench.py --mnk 65536,1024,1024 --layouts nn,nt --iters 5 | grep TFLOP
  python tools/gemm_bench.py --mnk 784,1024,65536 --layouts tn --iters 5 | grep TFLOP
  python tools/gemm_bench.py --mnk 8192,1024,784 --layouts nn --iters 5 | grep TFLOP
  python tools/gemm_bench.py --mnk 8192,1024,1024 --layouts nn,nt --iters 5 | grep TFLOP
  python tools/gemm_bench.py --mnk 784,1024,8192 --layouts tn --iters 5 | grep TFLOP
  python tools/gemm_bench.py --mnk 1024,1024,8192 --layouts tn --iters 5 | grep TFLOP
done; done
