"""One rank of a multi-GPU batch-sharded gradient evaluation (launched by
tests/test_gpu_sharded.py through torch.distributed.run).  Rank 0 writes the
summed gradients to argv[1]."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path, batch, ordered = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ["LLO_CUDA_DEVICE"] = str(local)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import workloads
    from cortenn_b200 import age, llo
    from cortenn_b200.sharded import BatchShardedExecutor
    sizes = (48, 64, 40, 10)
    X, Y, params = workloads.mlp_data(batch, sizes, np.float32)
    ex = BatchShardedExecutor(batch, ordered=bool(ordered))
    assert llo.comm_size() == world and llo.comm_rank() == rank
    loss, pvars, _ = workloads.mlp_graph(age, llo, ex.shard(X), ex.shard(Y), params, global_batch=batch)
    first = ex.grad_eval(loss, pvars, np.float32, with_loss=True)
    again = ex.grad_eval(loss, pvars, np.float32, with_loss=True)
    for a, b in zip(first, again):
        np.testing.assert_array_equal(a, b)      # run-to-run deterministic
    # every rank holds the same sums
    flat = np.concatenate([np.asarray(g).reshape(-1) for g in first])
    t = torch.from_numpy(flat.copy()).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if ordered:
        assert torch.equal(lo, hi), "ordered allreduce must be bit-identical on every rank"
    if rank == 0:
        np.savez(out_path, *first)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
