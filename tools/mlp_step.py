"""A few MLP gradient evaluations (BASELINE config 5) on one GPU, for timing and ncu launch lists.
usage: python tools/mlp_step.py [--batch 65536] [--iters 5] [--tf32x3]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads  # noqa: E402
from cortenn_b200 import age, capi, llo  # noqa: E402
from cortenn_b200.sharded import BatchShardedExecutor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--tf32x3", action="store_true")
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--recompute", type=int, default=-1)
    args = ap.parse_args()
    if args.tf32x3:
        llo.set_option("tf32x3", 1)
    if args.recompute >= 0:
        llo.set_option("recompute", args.recompute)
    X, Y, params = workloads.mlp_data(args.batch)
    ex = BatchShardedExecutor(args.batch, rank=0, world=1)
    loss, pvars, _ = workloads.mlp_graph(age, llo, ex.shard(X), ex.shard(Y), params, global_batch=args.batch)
    ctx = capi.Context(handle=llo.context())
    f32 = np.dtype(np.float32)
    import time
    start, stop = ctx.event(), ctx.event()
    for it in range(3 + args.iters):
        t0 = time.perf_counter()
        before = llo.launch_count()
        ctx.record(start)
        ex.grad_eval_device(loss, pvars, f32)
        t1 = time.perf_counter()
        ctx.record(stop)
        one = ctx.elapsed_ms(start, stop)
        t2 = time.perf_counter()
        print("  iter %d: device %.3f ms, host enqueue %.3f ms, host wait %.3f ms, %d launches" % (
            it, one, (t1 - t0) * 1e3, (t2 - t1) * 1e3, llo.launch_count() - before))
    llo.sync()
    ctx.record(start)
    for _ in range(args.iters):
        ex.grad_eval_device(loss, pvars, f32)
    ctx.record(stop)
    print("batch %d: %.3f ms per grad-eval back to back, specialise %s" % (
        args.batch, ctx.elapsed_ms(start, stop) / args.iters, llo.specialize_stats()))


if __name__ == "__main__":
    main()
