#!/bin/bash
# A/B: rolling-A SGEMM (no transposition pass for a k-contiguous A) and the wide DGEMM
for r in 0 1; do
  echo "== LLO_GEMM_ROLL_A=$r"
  for mnk in 65536,1024,784 65536,1024,1024 8192,1024,784 8192,1024,1024 8192,8192,8192 4096,4096,4096 1000,3000,5000; do
    LLO_GEMM_ROLL_A=$r python tools/gemm_bench.py --mnk $mnk --layouts nn,nt --iters 5 2>&1 | grep "^f32"
  done
done
for w in 0 1; do
  echo "== LLO_DGEMM_WIDE=$w"
  LLO_DGEMM_WIDE=$w python tools/gemm_bench.py --dtype f64 --size 8192 --layouts nn,nt,tn,tt --iters 3 2>&1 | grep "^f64"
  LLO_DGEMM_WIDE=$w python tools/gemm_bench.py --dtype f64 --mnk 4096,4096,4096 --layouts nn --iters 3 2>&1 | grep "^f64"
  LLO_DGEMM_WIDE=$w python tools/gemm_bench.py --dtype f64 --mnk 3000,5000,1000 --layouts nn,tn --iters 3 2>&1 | grep "^f64"
done
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm or matmul" 2>&1 | tail -2
python tools/mlp_step.py 2>&1 | tail -3
