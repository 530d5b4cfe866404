"""CPU property test: on random expression graphs the EXACT passes of llo.optimize
(x*1, x+0, pow(x,1), single-argument SUM/PROD, CSE) never change a value -- the oracle
evaluates the graph as written and the rewritten graph to the same bits, in float64,
float32 and int32 -- and they never add functors.  Gradients of the same graphs
(llo.derive, which is where the 1- and 0-labelled leaves come from) included."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from util import age, llo, oracle

ROWS, COLS = 3, 4          # numpy shape; ade shape [COLS, ROWS]


def leaves():
    rng = np.random.default_rng(7)
    return [llo.variable(rng.uniform(0.5, 2.0, (ROWS, COLS)), "v%d" % i) for i in range(3)]


def build(draw, pool, depth):
    """a random expression over `pool` (repeats allowed: shared sub-graphs feed CSE)"""
    if depth == 0 or draw(st.integers(0, 9)) < 2:
        return pool[draw(st.integers(0, len(pool) - 1))]
    kind = draw(st.sampled_from(["add", "sub", "mul", "div", "neg", "abs", "sqrt", "one_mul", "mul_one",
                                 "zero_add", "pow_one", "twice", "bcast", "min", "max", "single", "sum3"]))
    a = build(draw, pool, depth - 1)
    if kind in ("neg", "abs", "sqrt"):
        return getattr(age, kind)(age.abs(a) if kind == "sqrt" else a)
    if kind == "one_mul":
        return age.mul(llo.scalar(1.0, [ROWS, COLS]), a)
    if kind == "mul_one":
        return age.mul(a, llo.scalar(1.0, [ROWS, COLS]))
    if kind == "zero_add":
        return age.add(a, llo.scalar(0.0, [ROWS, COLS]))
    if kind == "pow_one":
        return age.pow(a, llo.scalar(1.0, [ROWS, COLS]))
    if kind == "single":
        return age.sum([a])                                    # one-argument SUM: OPT_SINGLE
    if kind == "sum3":
        return age.sum([a, llo.scalar(0.0, [ROWS, COLS]), a])
    if kind == "twice":
        return age.add(age.mul(a, a), age.mul(a, a))          # equal functors, built twice
    if kind == "bcast":
        return age.mul(a, age.extend(age.reduce_sum(a, 1), 1, [ROWS]))
    b = build(draw, pool, depth - 1)
    if kind in ("min", "max"):
        return getattr(age, kind)([a, b])
    return getattr(age, kind)(a, b)


def opnames(root):
    # (wrappers are kept alive while walking: a freed wrapper's id() may be handed to the next)
    seen, todo, names = {}, [root], []
    while todo:
        t = todo.pop()
        if id(t) in seen:
            continue
        seen[id(t)] = t
        if t.opname():
            names.append(t.opname())
        todo += t.children()
    return names


def count(root):
    return len(opnames(root))


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.data())
def test_exact_passes_preserve_every_bit(data):
    pool = leaves()
    root = build(data.draw, pool, data.draw(st.integers(1, 4)))
    roots = [root]
    if root.opname():
        roots.append(llo.derive(root, pool[0]))
    new, report = llo.optimize(roots, llo.OPT_EXACT)
    assert report["div_cancel"] == 0
    assert report["functors_after"] <= report["functors_before"]
    for before, after in zip(roots, new):
        assert list(after.shape()) == list(before.shape())
        # (an integer x / 0 traps in the reference's `a / b`, operator.hpp:257-261, hence in
        # the oracle: integer runs only for graphs without a division)
        dtypes = [np.float64, np.float32]
        if "DIV" not in opnames(before) and "POW" not in opnames(before):
            dtypes.append(np.int32)
        for dtype in dtypes:
            want = oracle.evaluate(before, np.dtype(dtype))
            got = oracle.evaluate(after, np.dtype(dtype))
            assert want.dtype == got.dtype
            np.testing.assert_array_equal(want.view(np.uint8), got.view(np.uint8))


@settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.data())
def test_optimize_is_idempotent(data):
    pool = leaves()
    root = build(data.draw, pool, data.draw(st.integers(1, 4)))
    (once,), first = llo.optimize([root], llo.OPT_EXACT)
    (twice,), second = llo.optimize([once], llo.OPT_EXACT)
    assert second["functors_after"] == second["functors_before"] == first["functors_after"]
    assert count(twice) == count(once)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.data())
def test_random_graphs_survive_pbm_save_and_load(data):
    """pbm::GraphSaver -> bytes -> pbm::load_graph (pbm/src/save.cpp, load.cpp): the loaded
    graph -- rewritten or as derived -- evaluates to the same bits and saves to the same bytes"""
    pool = leaves()
    root = build(data.draw, pool, data.draw(st.integers(1, 4)))
    roots = [root, llo.derive(root, pool[1])] if root.opname() else [root]
    if data.draw(st.booleans()):
        roots = list(llo.optimize(roots, llo.OPT_EXACT)[0])
    blob = llo.save_graph(roots)
    loaded = llo.load_graph(blob)
    assert len(loaded["roots"]) >= 1
    # roots that are sub-graphs of other roots are not roots of the file: compare by value
    want = {oracle.evaluate(r, np.dtype(np.float64)).tobytes() for r in roots}
    got = {oracle.evaluate(r, np.dtype(np.float64)).tobytes() for r in loaded["roots"]}
    assert got <= want and len(got) >= 1
    # one round trip is a fixed point of the byte form
    blob2 = llo.save_graph(loaded["roots"])
    again = llo.load_graph(blob2)
    assert len(again["nodes"]) == len(loaded["nodes"])
    assert llo.save_graph(again["roots"]) == blob2
