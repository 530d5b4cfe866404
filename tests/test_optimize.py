"""llo.optimize: graph -> graph passes over ade graphs.

CPU part: structure of the rewritten graphs and their values through the ORACLE
(the oracle evaluates any ade graph, so "the rewritten graph computes what the
graph as derived computes" is checked without a GPU).  GPU part: the rewritten
graph on the B200 against the oracle on the un-rewritten graph."""
import numpy as np
import pytest

import util
from util import age, llo, oracle

import workloads

F4, F8, I4 = np.float32, np.float64, np.int32


def functors(root):
    # (wrappers are kept alive while walking: a freed wrapper's id() may be handed to the next)
    seen, todo, n = {}, [root], 0
    while todo:
        t = todo.pop()
        if id(t) in seen:
            continue
        seen[id(t)] = t
        kids = t.children()
        if t.opname():
            n += 1
        todo += kids
    return n


def test_neutral_elements_and_pow_one_return_the_argument_itself():
    x = llo.variable(np.arange(6, dtype=F8).reshape(2, 3) + 1, "x")
    one = llo.scalar(1.0, [2, 3])
    zero = llo.scalar(0.0, [2, 3])
    for node, field in ((age.mul(x, one), "one_mul"), (age.mul(one, x), "one_mul"),
                        (age.add(x, zero), "zero_add"), (age.pow(x, one), "pow_one")):
        (out,), rep = llo.optimize([node])
        assert rep[field] == 1 and rep["functors_after"] == 0, rep
        assert str(out) == str(x) and not out.opname()
    # every argument neutral: a constant of the functor's shape
    (out,), rep = llo.optimize([age.mul(one, one)])
    assert not out.opname() and list(out.shape()) == [2, 3]
    np.testing.assert_array_equal(oracle.evaluate(out, np.dtype(F8)), np.ones((2, 3)))
    # a pass that is not selected does not fire
    (out,), rep = llo.optimize([age.mul(x, one)], llo.OPT_CSE)
    assert out.opname() == "PROD" and rep["one_mul"] == 0


def test_extended_argument_keeps_its_map():
    b = llo.variable(np.array([1.0, 2.0, 3.0]), "b")
    one = llo.scalar(1.0, [4, 3])
    node = age.mul(age.extend(b, 1, [4]), one)
    (out,), rep = llo.optimize([node])
    assert rep["one_mul"] == 1
    np.testing.assert_array_equal(oracle.evaluate(out, np.dtype(F8)), oracle.evaluate(node, np.dtype(F8)))


def test_cse_merges_equal_functors_but_never_random_ones():
    x = llo.variable(np.linspace(0.5, 2, 12).reshape(3, 4), "x")
    a = age.exp(age.sin(x))
    b = age.exp(age.sin(x))          # structurally equal, distinct nodes
    root = age.add(a, b)
    (out,), rep = llo.optimize([root], llo.OPT_CSE)
    assert rep["shared"] == 2 and rep["functors_before"] == 5 and rep["functors_after"] == 3, rep
    kids = out.children()
    assert kids[0] is kids[1] or str(kids[0]) == str(kids[1])
    np.testing.assert_array_equal(oracle.evaluate(out, np.dtype(F8)), oracle.evaluate(root, np.dtype(F8)))
    lo, hi = llo.scalar(0.0, [3, 4]), llo.scalar(1.0, [3, 4])
    r1, r2 = age.rand_unif(lo, hi), age.rand_unif(lo, hi)
    (out,), rep = llo.optimize([age.sub(age.neg(r1), age.neg(r2))], llo.OPT_CSE)
    assert rep["shared"] == 0 and rep["functors_after"] == 5, rep


def test_div_cancel_is_floating_point_only_and_flagged_inexact():
    a = llo.variable(np.array([[0.0, 2.0], [3.0, 4.0]]), "a")
    b = llo.variable(np.array([[5.0, 6.0], [7.0, 8.0]]), "b")
    grad = llo.derive(age.mul(a, b), a)
    (exact,), rep = llo.optimize([grad], llo.OPT_EXACT)
    assert rep["div_cancel"] == 0
    want = oracle.evaluate(grad, np.dtype(F8))
    got = oracle.evaluate(exact, np.dtype(F8))
    np.testing.assert_array_equal(np.isnan(want), np.isnan(got))      # NaN at a == 0 survives OPT_EXACT
    (fast,), rep = llo.optimize([grad], llo.OPT_ALL)
    assert rep["div_cancel"] == 1
    np.testing.assert_array_equal(oracle.evaluate(fast, np.dtype(F8)), [[5.0, 6.0], [7.0, 8.0]])
    # integer leaves: (100 * 3) / 100 wraps and truncates in int8, so no rewrite
    ia = llo.variable(np.array([[100, 2]], dtype=np.int8), "ia")
    ib = llo.variable(np.array([[3, 4]], dtype=np.int8), "ib")
    igrad = llo.derive(age.mul(ia, ib), ia)
    (same,), rep = llo.optimize([igrad], llo.OPT_ALL)
    assert rep["div_cancel"] == 0
    np.testing.assert_array_equal(oracle.evaluate(same, np.dtype(np.int8)), oracle.evaluate(igrad, np.dtype(np.int8)))


@pytest.mark.parametrize("dtype,tol", [(F4, 2e-5), (F8, 1e-11)])
def test_optimized_mlp_gradients_match_the_graph_as_derived_on_the_oracle(dtype, tol):
    X, Y, params = workloads.mlp_data(16, sizes=(6, 8, 8, 3), dtype=dtype)
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params)
    grads = [llo.derive(loss, p) for p in pvars]
    exact, rep_exact = llo.optimize(grads, llo.OPT_EXACT)
    fast, rep = llo.optimize(grads)
    assert rep_exact["functors_after"] < rep_exact["functors_before"]
    assert rep["div_cancel"] > 0 and rep["functors_after"] <= rep_exact["functors_after"]
    for g, e, f in zip(grads, exact, fast):
        want = oracle.evaluate(g, np.dtype(dtype))
        np.testing.assert_array_equal(want, oracle.evaluate(e, np.dtype(dtype)))   # exact passes: bit-identical
        got = oracle.evaluate(f, np.dtype(dtype)).astype(np.float64)
        scale = np.maximum(np.abs(want).max(), 1e-30)
        assert np.abs(got - want).max() <= tol * scale
    # an optimized graph is an ordinary ade graph: it derives and optimizes again
    again, rep2 = llo.optimize(fast)
    assert rep2["functors_after"] == rep2["functors_before"]
    assert functors(again[0]) == functors(fast[0])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [F4, F8])
def test_gpu_evaluates_optimized_graphs_like_the_oracle_evaluates_the_original(dtype):
    X, Y, params = workloads.mlp_data(48, sizes=(10, 16, 12, 4), dtype=dtype)
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params)
    grads = [llo.derive(loss, p) for p in pvars]
    fast, rep = llo.optimize([loss] + grads)
    got = [d.numpy(np.dtype(dtype)) for d in llo.evaluate_device_many(fast, np.dtype(dtype))]
    st = llo.plan_stats()
    assert st["gemms"] >= 7 and st["generic"] == 0, st
    tol = 2e-4 if dtype == F4 else 1e-10
    for root, g in zip([loss] + grads, got):
        want = oracle.evaluate(root, np.dtype(dtype)).astype(np.float64)
        assert np.abs(g.reshape(want.shape) - want).max() <= tol * max(np.abs(want).max(), 1e-30)
