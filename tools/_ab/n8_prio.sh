#!/bin/bash
# N-rank MLP step: priority of the collective lane, NCCL CTA cap, overlap off
N=${1:-8}
run() {
  echo "== $*"
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29771 \
    bench.py --gpus $N --steps 20 --warmup 5 --mlp-only --skip-verify 2>gpurun_out/n8_prio.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
m=d['mlp_grad_eval']; print({k:m.get(k) for k in ('ms','n1_ms_same_run','speedup_vs_n1_same_run','launches')}, d.get('mlp_grad_eval_3xtf32',{}).get('ms'))
" || tail -3 gpurun_out/n8_prio.err
}
run LLO_COMM_PRIORITY=0
run LLO_COMM_PRIORITY=1
run LLO_COMM_PRIORITY=1 NCCL_MAX_CTAS=8
run LLO_COMM_PRIORITY=1 LLO_SHARD_OVERLAP=0
