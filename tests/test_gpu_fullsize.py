"""Parity at the sizes bench.py times (BASELINE.json configs 2-5), not at toy sizes.

The size-dependent code paths -- the pair / tile kernels' work-unit counters,
8-wave column folds, the [k][row] relayout, split-K, the persistent GEMM's tail
split, the skinny kernels -- only switch on at these shapes, so they are
compared here against float64 numpy on the full tensors (a few seconds of host
time each), with the tolerances of BASELINE.json's north star written out:
1e-6 relative for fp32 elementwise results, 1e-12 for fp64, sums against the
sum |terms| bound, integer results bit-exact.  The oracle (an O(150 ns/element)
CPU interpreter) checks slabs: the same kernels at a size it finishes in seconds.
"""
import numpy as np
import pytest

import util
from util import age, llo, oracle

import workloads

pytestmark = pytest.mark.gpu

F4, F8 = np.dtype(np.float32), np.dtype(np.float64)
SIDE = 16384


def rel_err(got, want64):
    return float((np.abs(got.astype(np.float64) - want64) / np.abs(want64)).max())


@pytest.fixture(scope="module")
def big_x():
    return np.random.default_rng(2).uniform(0.5, 2.0, (SIDE, SIDE)).astype(np.float32)


def test_chain_2pow28_matches_float64_numpy(big_x):
    """config 2: add(mul(exp(x), extend(b)), permute(x)) over [16384,16384] fp32:
    every one of the 2^28 outputs within 1e-6 relative of float64"""
    b = np.random.default_rng(3).uniform(0.5, 2.0, (SIDE,)).astype(np.float32)
    vx, vb, root = workloads.chain_graph(age, llo, big_x, b)
    got = llo.evaluate(root, F4)
    assert llo.plan_stats()["programs"] == 1
    worst = 0.0
    for lo in range(0, SIDE, 2048):            # float64 reference in row blocks (memory)
        x64 = big_x[lo:lo + 2048].astype(np.float64)
        want = np.exp(x64) * b.astype(np.float64)[None, :] + big_x[:, lo:lo + 2048].T.astype(np.float64)
        worst = max(worst, rel_err(got[lo:lo + 2048], want))
    assert worst <= 1e-6, worst


def test_chain_128x4_matches_float64_numpy(big_x):
    """config 2 in the reference-representable shape [128,128,128,128]"""
    x4 = big_x.reshape(128, 128, 128, 128)
    b = np.random.default_rng(3).uniform(0.5, 2.0, (128,)).astype(np.float32)
    vx, vb = llo.variable(x4, "x"), llo.variable(b, "b")
    root = age.add(age.mul(age.exp(vx), age.extend(vb, 1, [128, 128, 128])), age.permute(vx, [1, 0, 2, 3]))
    got = llo.evaluate(root, F4)
    worst = 0.0
    for i in range(0, 128, 16):
        x64 = x4[i:i + 16].astype(np.float64)
        want = np.exp(x64) * b.astype(np.float64) + np.swapaxes(x64, 2, 3)
        worst = max(worst, rel_err(got[i:i + 16], want))
    assert worst <= 1e-6, worst


@pytest.mark.parametrize("op", ["sum", "max"])
def test_reductions_16384_squared_all_axes(big_x, op):
    """config 3: reduce_sum / reduce_max of [16384,16384] along each axis; sums within
    1e-6 of float64 (all terms positive: the sum IS the sum |terms| bound), max exact"""
    vx = llo.variable(big_x, "x")
    red = {"sum": (age.reduce_sum, age.reduce_sum0), "max": (age.reduce_max, age.reduce_max0)}[op]
    x64 = big_x.astype(np.float64)
    cases = {
        "cols": (red[0](vx, 1), x64.sum(axis=0) if op == "sum" else x64.max(axis=0)),
        "rows": (red[0](age.permute(vx, [1, 0]), 1), x64.sum(axis=1) if op == "sum" else x64.max(axis=1)),
        "all": (red[1](vx), np.array(x64.sum() if op == "sum" else x64.max())),
    }
    for name, (root, want) in cases.items():
        got = llo.evaluate(root, F4).reshape(want.shape)
        if op == "max":
            np.testing.assert_array_equal(got.astype(np.float64), want, err_msg=name)
        else:
            assert rel_err(got, want) <= 1e-6, (name, rel_err(got, want))
        # fixed order: a second evaluation is bit-identical
        np.testing.assert_array_equal(got, llo.evaluate(root, F4).reshape(want.shape), err_msg=name)


def test_sequential_reduce_is_bit_exact_against_the_oracle_on_a_slab(big_x):
    """LLO_REDUCE_SEQUENTIAL reproduces the reference's accumulation order
    (llo/operator.hpp:376-393): bit-identical fp32 sums on a [16384, 256] slab"""
    slab = np.ascontiguousarray(big_x[:256])                     # ade shape [16384, 256]
    vx = llo.variable(slab, "x")
    llo.set_option("sequential_reduce", 1)
    try:
        for root in (age.reduce_sum(vx, 1), age.reduce_sum0(vx), age.reduce_sum(age.permute(vx, [1, 0]), 1)):
            got = llo.evaluate(root, F4)
            want = oracle.evaluate(root, F4)
            np.testing.assert_array_equal(want, got)
    finally:
        llo.set_option("sequential_reduce", 0)


@pytest.mark.parametrize("dtype,option,tol", [(np.float32, 0, 1e-6), (np.float32, 1, 1e-6), (np.float64, 0, 1e-12)],
                         ids=["f32-ffma", "f32-3xtf32", "f64-dfma"])
def test_matmul_8192_cubed_sampled_rows(dtype, option, tol):
    """config 4: 64 sampled rows x all 8192 columns against float64, error within
    tol * sum_k |a||b| (the forward error bound of a length-8192 dot product)"""
    m = 8192
    a = np.random.default_rng(5).uniform(-1, 1, (m, m)).astype(dtype)
    b = np.random.default_rng(6).uniform(-1, 1, (m, m)).astype(dtype)
    root = age.matmul(llo.variable(a, "a"), llo.variable(b, "b"))
    llo.set_option("tf32x3", option)
    try:
        res = llo.evaluate_device(root, np.dtype(dtype))
        assert llo.plan_stats()["gemms"] == 1
        got = res.numpy(np.dtype(dtype))
    finally:
        llo.set_option("tf32x3", 0)
    rows = np.random.default_rng(7).choice(m, 64, replace=False)
    a64, b64 = a[rows].astype(np.float64), b.astype(np.float64)
    want = a64 @ b64
    bound = np.abs(a64) @ np.abs(b64)
    err = np.abs(got[rows].astype(np.float64) - want) / bound
    assert err.max() <= tol, err.max()


def mlp_reference(X, Y, params, batch):
    """float64 gradients plus the sum |terms| bound of every gradient element"""
    X, Y = X.astype(np.float64), Y.astype(np.float64)
    P = [p.astype(np.float64) for p in params]
    hs = [X]
    for i in range(len(P) // 2):
        hs.append(1.0 / (1.0 + np.exp(-(hs[-1] @ P[2 * i] + P[2 * i + 1]))))
    loss = ((hs[-1] - Y) ** 2).sum() / batch
    grads, bounds = [None] * len(P), [None] * len(P)
    delta = 2.0 * (hs[-1] - Y) / batch
    for i in reversed(range(len(P) // 2)):
        dz = delta * hs[i + 1] * (1.0 - hs[i + 1])
        grads[2 * i], bounds[2 * i] = hs[i].T @ dz, np.abs(hs[i]).T @ np.abs(dz)
        grads[2 * i + 1], bounds[2 * i + 1] = dz.sum(axis=0), np.abs(dz).sum(axis=0)
        delta = dz @ P[2 * i].T
    return loss, grads, bounds


def test_mlp_batch_65536_gradients_match_float64():
    """config 5 at N = 1: all six gradients of the 784-1024-1024-10 MLP at batch 65536
    (the graph bench.py times, through llo.derive + llo.optimize) against float64.
    A gradient element is a 65536-term sum of signed products that were themselves
    computed in fp32 through three layers: the bound is 2e-5 * sum |terms| (measured
    ~3e-6; the north star's 1e-6 is per operation, not per 65536-term sum)."""
    batch = 65536
    X, Y, params = workloads.mlp_data(batch)
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params)
    grads = [llo.derive(loss, p) for p in pvars]
    roots, report = llo.optimize([loss] + grads)
    assert report["div_cancel"] > 0
    res = llo.evaluate_device_many(roots, F4)
    st = llo.plan_stats()
    assert st["generic"] == 0 and st["gemms"] >= 8, st
    got = [r.numpy(F4) for r in res]
    want_loss, want, bounds = mlp_reference(X, Y, params, batch)
    assert abs(float(got[0].reshape(-1)[0]) - want_loss) <= 1e-5 * want_loss
    for g, w, bnd, p in zip(got[1:], want, bounds, params):
        err = np.abs(g.reshape(w.shape).astype(np.float64) - w) / bnd
        assert err.max() <= 2e-5, (p.shape, err.max())
