mkdir -p gpurun_out/s13
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --skip-side --skip-mlp --skip-cpu > gpurun_out/s13/bench_e2e.log 2> gpurun_out/s13/bench_e2e.err; tail -c 800 gpurun_out/s13/bench_e2e.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/s13/bench_e2e.log') if l.startswith('{')][0])
print(d['value'], json.dumps(d['e2e'], indent=1))"
