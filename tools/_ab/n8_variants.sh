#!/bin/bash
run () { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 8 --steps 20 --warmup 5 --mlp-only --skip-verify 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); m=d['mlp_grad_eval']; print({k:m[k] for k in ('ms','launches','host_enqueue_ms','n1_ms_same_run','speedup_vs_n1_same_run')}, d['mlp_grad_eval_3xtf32']['ms'])"; }
run X=1
run NCCL_MAX_NCHANNELS=4
run NCCL_MAX_NCHANNELS=8
run LLO_SHARD_OVERLAP=0
run LLO_SHARD_GRAPH=0
run NCCL_ALGO=Tree
