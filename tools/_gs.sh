mkdir -p gpurun_out/s16
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s16/bench.log 2> gpurun_out/s16/bench.err; tail -c 600 gpurun_out/s16/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s16/bench.log') if l.startswith('{')][0])
w=d['workloads']
print('chain', d['value'], 'e2e', d['e2e']['value'])
for k in ('matmul_f32','matmul_f32_3xtf32','matmul_f64','mlp_grad_eval','mlp_grad_eval_3xtf32'):
    print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in w[k].items() if a in ('TFLOPs','ms','launches_per_eval','host_enqueue_ms')})
PY
