#!/bin/bash
for r in 0 1; do
  echo "== LLO_GEMM_ROLL_A=$r"
  for mnk in 65536,1024,784 65536,1024,1024 8192,1024,784 8192,1024,1024 8192,8192,8192 1000,3000,5000; do
    LLO_GEMM_ROLL_A=$r python tools/gemm_bench.py --mnk $mnk --layouts nn --iters 5 2>&1 | grep "^f32"
  done
done
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm or matmul" 2>&1 | tail -2
