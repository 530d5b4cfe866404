"""CPU tests of the batch-sharded executor's HOST logic: shard
bounds, the packed gradient layout, and a world_size-2 run over gloo in which
the oracle stands in for the device evaluator (test infrastructure only; the
product backend is DeviceBackend and has no CPU path)."""
import os
import socket
import sys

import numpy as np
import pytest

import util
from util import ROOT

sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from cortenn_b200.sharded import BatchShardedExecutor, GradientPack, ShardPlan  # noqa: E402


def test_shard_plan_covers_the_batch_contiguously():
    for batch, world in ((8, 1), (8, 2), (10, 4), (65536, 8), (7, 7), (13, 5)):
        plan = ShardPlan(batch, world)
        edges = [plan.bounds(r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == batch
        for (lo, hi), (nlo, _) in zip(edges, edges[1:]):
            assert hi == nlo and hi > lo
        sizes = [hi - lo for lo, hi in edges]
        assert max(sizes) - min(sizes) <= 1
        assert sum(sizes) == batch
    with pytest.raises(ValueError):
        ShardPlan(3, 4)
    with pytest.raises(ValueError):
        ShardPlan(8, 2).bounds(2)
    x = np.arange(20).reshape(10, 2)
    plan = ShardPlan(10, 4)
    np.testing.assert_array_equal(np.concatenate([plan.slice(x, r) for r in range(4)]), x)
    with pytest.raises(ValueError):
        plan.slice(np.zeros((9, 2)), 0)


def test_gradient_pack_layout_round_trips():
    shapes = [(), (784, 1024), (1024,), (1024, 10), (10,), (3, 5, 7)]
    pack = GradientPack(shapes, np.float32)
    assert all(off % 4 == 0 for off in pack.offsets)          # 16-byte aligned
    assert pack.count >= sum(pack.sizes())
    rng = np.random.default_rng(0)
    arrays = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    flat = pack.pack(arrays)
    for want, got in zip(arrays, pack.unpack(flat)):
        np.testing.assert_array_equal(want, got)
    # the MLP of config 5: 1 863 690 gradient elements
    mlp = GradientPack([(784, 1024), (1024,), (1024, 1024), (1024,), (1024, 10), (10,)], np.float32)
    assert sum(mlp.sizes()) == 1863690
    assert mlp.nbytes == mlp.count * 4


class OracleGlooBackend:
    """Stand-in backend: oracle evaluation + torch.distributed (gloo) sum"""

    def __init__(self, dist):
        self.dist = dist
        self.comm = None

    def init_comm(self, rank, world, exchange):
        ident = b"x" * 128 if rank == 0 else None
        self.comm = (rank, world, exchange(ident) if world > 1 else ident)

    def evaluate_many(self, roots, dtype):
        return [np.asarray(util.oracle.evaluate(r, np.dtype(dtype))) for r in roots]

    def reduce_packed(self, results, pack):
        import torch
        flat = torch.from_numpy(pack.pack(results))
        if self.dist is not None:
            self.dist.all_reduce(flat)
        return flat.numpy()

    def to_host(self, buf, pack):
        return pack.unpack(buf)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, sizes, out_path):
    import torch.distributed as dist
    from cortenn_b200 import age, llo
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        X, Y, params = workloads.mlp_data(batch, sizes, np.float64)
        ex = BatchShardedExecutor(batch, rank=rank, world=world, backend=OracleGlooBackend(dist))
        assert ex.backend.comm[2] == b"x" * 128                 # the id reached every rank
        loss, pvars, _ = workloads.mlp_graph(age, llo, ex.shard(X), ex.shard(Y), params, global_batch=batch)
        got = ex.grad_eval(loss, pvars, np.float64, with_loss=True)
        if rank == 0:
            np.savez(out_path, *got)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo_shard_gradients_sum_to_the_full_gradient(tmp_path):
    import torch.multiprocessing as mp
    batch, sizes = 10, (6, 8, 5, 3)
    out_path = str(tmp_path / "grads.npz")
    mp.spawn(_worker, args=(2, _free_port(), batch, sizes, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    got = [got["arr_%d" % i] for i in range(len(got.files))]
    X, Y, params = workloads.mlp_data(batch, sizes, np.float64)
    want_loss, want = workloads.mlp_reference_gradients(X, Y, params)
    assert abs(float(got[0]) - want_loss) <= 1e-12 * max(1.0, abs(want_loss))
    for g, w in zip(got[1:], want):
        assert g.shape == w.shape
        assert np.abs(g - w).max() <= 1e-12 * max(1.0, np.abs(w).max())
    # and equals the single-rank evaluation of the whole batch
    from cortenn_b200 import age, llo
    ex1 = BatchShardedExecutor(batch, rank=0, world=1, backend=OracleGlooBackend(None))
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params)
    single = ex1.grad_eval(loss, pvars, np.float64, with_loss=True)
    for g, s in zip(got, single):
        assert np.abs(np.asarray(g) - np.asarray(s)).max() <= 1e-12 * max(1.0, np.abs(s).max())


def test_device_backend_refuses_to_run_without_a_gpu():
    from cortenn_b200 import llo
    from cortenn_b200.sharded import DeviceBackend
    if llo.available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceBackend()
