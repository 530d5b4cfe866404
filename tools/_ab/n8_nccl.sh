#!/bin/bash
# N-rank MLP step under different NCCL algorithm choices
N=${1:-8}
run() {
  echo "== $*"
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29771 \
    bench.py --gpus $N --steps 10 --warmup 3 --mlp-only --skip-verify > gpurun_out/n8_nccl.out 2>gpurun_out/n8_nccl.err
  tail -1 gpurun_out/n8_nccl.out | python -c "
import json,sys
d=json.loads(sys.stdin.read())
m=d['mlp_grad_eval']; print({k:m.get(k) for k in ('ms','n1_ms_same_run','speedup_vs_n1_same_run')}, d.get('mlp_grad_eval_3xtf32',{}).get('ms'))
" || tail -3 gpurun_out/n8_nccl.err
  grep -h -i -E "nvls|algo|channels" gpurun_out/n8_nccl.err gpurun_out/n8_nccl.out | grep -v "^{" | sort | uniq -c | sort -rn | head -8
}
run NCCL_DEBUG=INFO
run NCCL_ALGO=NVLS
