mkdir -p gpurun_out/s9
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
LLO_RECOMPUTE=8 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rc in 0 6 8 12 0 8; do
for b in 8192 65536; do
echo "== recompute=$rc batch=$b"
LLO_RECOMPUTE=$rc python bench.py --skip-cpu --skip-side --mlp-batch $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']; print(w['mlp_grad_eval']['ms'], w['mlp_grad_eval']['launches_per_eval'], w['mlp_grad_eval_3xtf32']['ms'])"
done; done | tee gpurun_out/s9/mlp_rc.log
LLO_RECOMPUTE=8 LLO_PLAN_TRACE=1 python bench.py --skip-cpu --skip-side --mlp-batch 8192 --steps 1 --warmup 1 2> gpurun_out/s9/trace.log >/dev/null
