"""CPU: the host logic of the native batch-sharded evaluator (llo/sharded.hpp) that needs
no GPU -- the shard split, the derivation + optimisation of the gradient roots -- and its
agreement with the Python mirror the gloo tests drive (cortenn_b200/sharded.py)."""
import numpy as np
import pytest

import workloads
from cortenn_b200 import age, llo
from cortenn_b200.sharded import ShardPlan

F4 = np.dtype(np.float32)


@pytest.mark.parametrize("batch,world", [(65536, 8), (65536, 3), (10, 4), (7, 7), (96, 1)])
def test_native_shard_split_matches_the_python_mirror(batch, world):
    plan = ShardPlan(batch, world)
    covered = 0
    for rank in range(world):
        lo, hi = llo.shard_bounds(batch, world, rank)
        assert (lo, hi) == plan.bounds(rank)
        assert lo == covered and hi > lo
        covered = hi
    assert covered == batch
    with pytest.raises(RuntimeError):
        llo.shard_bounds(batch, world, world)


def test_native_evaluator_derives_and_optimises_once_without_a_gpu():
    X, Y, params = workloads.mlp_data(32, sizes=(16, 32, 32, 10))
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params, global_batch=32)
    ev = llo.ShardedEvaluator(loss, list(pvars))
    roots = ev.roots()
    assert len(roots) == 1 + len(pvars)                  # loss, then one gradient per parameter
    for root, p in zip(roots[1:], pvars):
        assert list(root.shape()) == list(p.shape())
    report = ev.opt_report()
    assert report["functors_after"] < report["functors_before"]
    assert report["div_cancel"] > 0 and report["shared"] > 0
    assert llo.plan_preview(roots, F4)["generic"] == 0
    assert ev.replays() == 0
    # no passes: the derived graphs as llo::derive returns them
    plain = llo.ShardedEvaluator(loss, list(pvars), passes=0)
    assert plain.opt_report()["functors_after"] in (0, plain.opt_report()["functors_before"])


def test_native_evaluator_fails_loudly_without_a_gpu():
    if llo.available():
        pytest.skip("a GPU is present")
    X, Y, params = workloads.mlp_data(8, sizes=(4, 6, 6, 3))
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params, global_batch=8)
    ev = llo.ShardedEvaluator(loss, list(pvars))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ev.grad_eval_device(F4, True)
