"""GPU tests of multi-root evaluation and the batch-sharded executor (K8).
The multi-rank cases need >= 2 GPUs on the box and are skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from util import ROOT, age, llo, oracle

sys.path.insert(0, ROOT)
import workloads  # noqa: E402

pytestmark = pytest.mark.gpu

F4 = np.float32


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def test_evaluate_many_shares_the_forward_pass():
    X, Y, params = workloads.mlp_data(32, (20, 24, 16, 6), F4)
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params)
    roots, _ = llo.optimize([loss] + [llo.derive(loss, p) for p in pvars])   # as the executor does
    llo.sync()
    before = llo.launch_count()
    singles = [llo.evaluate(r, np.dtype(F4)) for r in roots]
    separate = llo.launch_count() - before
    before = llo.launch_count()
    many = llo.evaluate_device_many(roots, np.dtype(F4))
    together = llo.launch_count() - before
    assert together < separate, (together, separate)
    for one, res in zip(singles, many):
        got = res.numpy(np.dtype(F4))
        assert got.shape == one.shape
        np.testing.assert_array_equal(one, got)     # same kernels, same order: bit-identical


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-5), (np.float64, 1e-12)])
def test_world1_executor_matches_oracle_and_numpy(dtype, tol):
    from cortenn_b200.sharded import BatchShardedExecutor
    batch, sizes = 24, (12, 16, 10, 4)
    X, Y, params = workloads.mlp_data(batch, sizes, dtype)
    ex = BatchShardedExecutor(batch, rank=0, world=1)
    loss, pvars, _ = workloads.mlp_graph(age, llo, ex.shard(X), ex.shard(Y), params, global_batch=batch)
    got = ex.grad_eval(loss, pvars, dtype, with_loss=True)
    want_loss, want = workloads.mlp_reference_gradients(X, Y, params)
    assert abs(float(got[0]) - want_loss) <= tol * max(1.0, abs(want_loss))
    for g, w, p in zip(got[1:], want, pvars):
        ref = oracle.evaluate(llo.derive(loss, p), np.dtype(dtype))
        assert g.shape == w.shape == ref.shape
        assert np.abs(g - w).max() <= tol * max(1.0, np.abs(w).max())
        assert np.abs(g - ref).max() <= tol * max(1.0, np.abs(ref).max())


def test_native_pack_layout_matches_the_python_mirror():
    from cortenn_b200.sharded import GradientPack
    X, Y, params = workloads.mlp_data(16, (12, 16, 10, 4), F4)
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params)
    ev = llo.ShardedEvaluator(loss, pvars)
    for with_loss in (False, True):
        lay = ev.pack_layout(np.dtype(F4), with_loss)
        shapes = ([()] if with_loss else []) + [tuple(int(d) for d in p.shape()) for p in pvars]
        mirror = GradientPack(shapes, F4)
        assert list(lay["offsets"]) == mirror.offsets and lay["count"] == mirror.count
        assert [tuple(s) for s in lay["shapes"]] == [tuple(s) for s in mirror.shapes]
    rep = ev.opt_report()
    assert rep["div_cancel"] > 0 and rep["functors_after"] < rep["functors_before"]
    # gradients are written straight into the packed buffer: no device-to-device copies,
    # and the results equal the one-root-at-a-time evaluation of the same rewritten graphs
    got = ev.grad_eval(np.dtype(F4), True)
    for root, g in zip(ev.roots(), got):
        want = llo.evaluate(root, np.dtype(F4))
        np.testing.assert_array_equal(np.asarray(g).reshape(want.shape), want)


def test_steps_replay_from_a_cuda_graph_until_a_leaf_changes():
    """from the second evaluation on the whole step is one cudaGraphLaunch; assigning to a
    leaf (new version, new mirror) makes the next step plan and capture again"""
    X, Y, params = workloads.mlp_data(64, (24, 32, 16, 6), F4)
    loss, pvars, (vx, vy) = workloads.mlp_graph(age, llo, X, Y, params)
    eager = llo.ShardedEvaluator(loss, pvars, graph=False)
    want = eager.grad_eval(np.dtype(F4), True)
    ev = llo.ShardedEvaluator(loss, pvars)
    first = ev.grad_eval(np.dtype(F4), True)           # eager warm-up
    assert ev.replays() == 0
    second = ev.grad_eval(np.dtype(F4), True)          # captured, launched once
    before = llo.launch_count()
    third = ev.grad_eval(np.dtype(F4), True)           # replayed
    assert ev.replays() == 1
    assert llo.launch_count() - before > 10            # the replay's kernels are counted
    for w, a, b, c in zip(want, first, second, third):
        np.testing.assert_array_equal(w, a)
        np.testing.assert_array_equal(w, b)
        np.testing.assert_array_equal(w, c)
    # new inputs: the graph is stale, the step is planned again and matches a fresh evaluator
    X2 = (X * 0.5 + 0.1).astype(F4)
    vx.assign(X2)
    got = ev.grad_eval(np.dtype(F4), True)
    assert ev.replays() == 1
    again = ev.grad_eval(np.dtype(F4), True)
    assert ev.replays() == 2
    fresh = eager.grad_eval(np.dtype(F4), True)
    want_loss, want_grads = workloads.mlp_reference_gradients(X2, Y, params)
    assert abs(float(got[0]) - want_loss) <= 2e-5 * abs(want_loss)
    for f, g, a in zip(fresh, got, again):
        np.testing.assert_array_equal(f, g)
        np.testing.assert_array_equal(f, a)


def test_device_buffer_round_trip_and_world1_allreduce():
    buf = llo.DeviceBuffer(64 * 4)
    data = np.arange(64, dtype=F4)
    buf.upload(0, data)
    llo.comm_init(0, 1)
    buf.allreduce(np.dtype(F4), "SUM", False)        # single rank: identity
    buf.allreduce(np.dtype(F4), "SUM", True)
    np.testing.assert_array_equal(buf.numpy(np.dtype(F4), 0, 64), data)
    with pytest.raises(RuntimeError):
        buf.numpy(np.dtype(F4), 0, 65)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("ordered", [0, 1])
@pytest.mark.timeout(600)
def test_multi_gpu_gradients_equal_single_gpu(world, ordered, tmp_path):
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    batch = 96
    out_path = str(tmp_path / "grads.npz")
    port = 29500 + world * 2 + ordered
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sharded_rank.py"), out_path, str(batch), str(ordered)]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    got = np.load(out_path)
    got = [got["arr_%d" % i] for i in range(len(got.files))]
    X, Y, params = workloads.mlp_data(batch, (48, 64, 40, 10), F4)
    want_loss, want = workloads.mlp_reference_gradients(X, Y, params)
    assert abs(float(got[0]) - want_loss) <= 2e-5 * max(1.0, abs(want_loss))
    # single-GPU evaluation of the whole batch: 1e-6 relative
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params)
    single = [llo.evaluate(llo.derive(loss, p), np.dtype(F4)) for p in pvars]
    for g, w, s in zip(got[1:], want, single):
        assert np.abs(g - w).max() <= 2e-5 * max(1.0, np.abs(w).max())
        assert np.abs(g - s).max() <= 2e-6 * max(1.0, np.abs(s).max())
