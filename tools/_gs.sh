mkdir -p gpurun_out/s7
python tools/gemm_bench.py --check --size 1024 --iters 2 > gpurun_out/s7/check.log 2>&1; tail -2 gpurun_out/s7/check.log
LLO_GEMM_PERSIST=1 python tools/gemm_bench.py --check --size 1024 --iters 2 > gpurun_out/s7/check_p.log 2>&1; tail -2 gpurun_out/s7/check_p.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for mode in 0 x; do
for b in 8192 65536; do
if [ $mode = 0 ]; then export LLO_GEMM_PERSIST=0; else unset LLO_GEMM_PERSIST; fi
echo "== persist=$mode batch=$b"
python bench.py --skip-cpu --skip-side --mlp-batch $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']; print(w['mlp_grad_eval']['ms'], w['mlp_grad_eval']['launches_per_eval'], w['mlp_grad_eval_3xtf32']['ms'])"
done; done | tee gpurun_out/s7/mlp_ab.log
