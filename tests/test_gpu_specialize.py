"""GPU: the run-time specialised build of the elementwise engine (NVRTC, LLO_VM_STATIC)
against the interpreted build of the same kernels -- bit-identical, since both apply the
same operator code in the same order -- and against the oracle."""
import numpy as np
import pytest

import util
from util import age, llo, oracle

pytestmark = pytest.mark.gpu

F4, F8, I4, I2, U1 = np.float32, np.float64, np.int32, np.int16, np.uint8
BIG = 1 << 62


@pytest.fixture
def both():
    """evaluate(node, dtype) -> (specialised, interpreted); checks that the specialised
    kernels really ran and none failed to build"""
    def run(node, dtype):
        llo.set_specialize_min(0)
        before = llo.specialize_stats()
        spec = llo.evaluate(node, np.dtype(dtype))
        after = llo.specialize_stats()
        assert after["failures"] == before["failures"], after
        assert after["launches"] > before["launches"], after
        llo.set_specialize_min(BIG)
        mid = llo.specialize_stats()
        interp = llo.evaluate(node, np.dtype(dtype))
        assert llo.specialize_stats()["launches"] == mid["launches"]
        return spec, interp
    yield run
    llo.set_specialize_min(1 << 20)


def rnd(rng, shape, dtype, lo=0.5, hi=2.0):
    if np.issubdtype(dtype, np.floating):
        return rng.uniform(lo, hi, shape).astype(dtype)
    return rng.integers(1, 9, shape).astype(dtype)


@pytest.mark.parametrize("dtype", [F4, F8, I4, I2, U1])
@pytest.mark.parametrize("side", [64, 200])
def test_chain_pair_and_tile_kernels(both, dtype, side):
    rng = np.random.default_rng(31)
    x = llo.variable(rnd(rng, (side, side), dtype), "x")
    b = llo.variable(rnd(rng, (side,), dtype), "b")
    first = age.exp(x) if np.issubdtype(dtype, np.floating) else age.abs(x)
    root = age.add(age.mul(first, age.extend(b, 1, [side])), age.permute(x, [1, 0]))
    spec, interp = both(root, dtype)
    np.testing.assert_array_equal(spec, interp)
    want = oracle.evaluate(root, np.dtype(dtype))
    if np.issubdtype(dtype, np.floating):
        assert np.abs(spec - want).max() <= (1e-6 if dtype == F4 else 1e-12) * np.abs(want).max()
    else:
        np.testing.assert_array_equal(spec, want)


def test_chain_rank4_pair_kernel(both):
    rng = np.random.default_rng(32)
    x4 = rnd(rng, (3, 2, 128, 128), F4)
    vx = llo.variable(x4, "x")
    vb = llo.variable(rnd(rng, (128,), F4), "b")
    root = age.add(age.mul(age.exp(vx), age.extend(vb, 1, [128, 2, 3])), age.permute(vx, [1, 0, 2, 3]))
    spec, interp = both(root, F4)
    np.testing.assert_array_equal(spec, interp)
    want = np.exp(x4.astype(np.float64)) * vb_value(vb, x4) + np.swapaxes(x4, 2, 3)
    assert (np.abs(spec - want) <= 1e-6 * np.abs(want)).all()


def vb_value(vb, like):
    return llo.evaluate(vb, np.dtype(np.float64)).reshape(1, 1, 1, -1)


UNARY = ["abs", "neg", "sin", "cos", "tan", "exp", "log", "sqrt", "round"]
BINARY = ["pow", "add", "sub", "mul", "div", "eq", "neq", "lt", "gt"]


@pytest.mark.parametrize("dtype", [F4, F8, I4])
def test_every_operator_flat_kernel(both, dtype):
    rng = np.random.default_rng(33)
    a = llo.variable(rnd(rng, (37, 53), dtype), "a")
    b = llo.variable(rnd(rng, (37, 53), dtype), "b")
    for name in UNARY:
        spec, interp = both(getattr(age, name)(a), dtype)
        np.testing.assert_array_equal(spec, interp, err_msg=name)
    for name in BINARY:
        spec, interp = both(getattr(age, name)(a, b), dtype)
        np.testing.assert_array_equal(spec, interp, err_msg=name)
    for name in ("min", "max", "sum", "prod"):
        spec, interp = both(getattr(age, name)([a, b, a]), dtype)
        np.testing.assert_array_equal(spec, interp, err_msg=name)


def test_program_with_spills_constants_and_broadcasts(both):
    rng = np.random.default_rng(34)
    shape = (24, 40)
    a, b, c = (llo.variable(rnd(rng, shape, F8), n) for n in "abc")
    row = llo.variable(rnd(rng, (40,), F8), "row")
    two = llo.scalar(2.0, list(shape))
    # (a + b) * (c - row) / (a * 2 + log(b)): both operands of * and / are sub-expressions
    root = age.div(age.mul(age.add(a, b), age.sub(c, age.extend(row, 1, [24]))),
                   age.add(age.mul(a, two), age.log(b)))
    spec, interp = both(root, F8)
    np.testing.assert_array_equal(spec, interp)
    want = oracle.evaluate(root, np.dtype(F8))
    assert np.abs(spec - want).max() <= 1e-12 * np.abs(want).max()


def test_kernels_are_cached_in_memory_and_on_disk(both):
    rng = np.random.default_rng(35)
    a = llo.variable(rnd(rng, (16, 16), F4), "a")
    root = age.sqrt(age.add(a, a))
    both(root, F4)
    llo.set_specialize_min(0)
    st0 = llo.specialize_stats()
    llo.evaluate(root, np.dtype(F4))
    st1 = llo.specialize_stats()
    assert st1["compiles"] == st0["compiles"] and st1["disk_hits"] == st0["disk_hits"]   # in-memory hit
    assert st1["launches"] == st0["launches"] + 1
