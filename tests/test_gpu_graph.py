"""GPU parity, graph level: llo.evaluate (planner: views, fused programs,
reductions from views, GEMM lowering) and the literal one-call-per-functor
mode, against the CPU oracle on the same ade graphs."""
import numpy as np
import pytest

import util
from util import age, llo, oracle, var

pytestmark = pytest.mark.gpu

F4, F8, I4, I8 = np.float32, np.float64, np.int32, np.int64


def tol_of(dtype):
    return {F4: 1e-6, F8: 1e-12}.get(dtype, 0)


def check(node, dtype, scale_tol=1.0, what="", floor=True):
    """floor=False: true relative error (graphs of + - * / exp sqrt pow abs neg);
    floor=True keeps an absolute floor at unit scale (sums with cancellation, sin / cos /
    log near their zeros), see tests/test_gpu_ops.py ABSOLUTE_FLOOR_OPS"""
    got = llo.evaluate(node, np.dtype(dtype))
    want = oracle.evaluate(node, np.dtype(dtype))
    assert got.shape == want.shape, what
    tol = tol_of(dtype) * scale_tol
    if tol == 0:
        np.testing.assert_array_equal(want, got, err_msg=what)
    else:
        w = want.astype(np.float64)
        g = got.astype(np.float64)
        assert (np.isnan(w) == np.isnan(g)).all(), what
        fin = np.isfinite(w)
        err = np.abs(w[fin] - g[fin])
        scale = np.maximum(np.abs(w[fin]), 1.0) if floor else np.abs(w[fin])
        assert (err <= tol * scale).all(), \
            "%s: max rel err %g" % (what, (err / np.maximum(scale, 1e-300)).max())
    return got


@pytest.fixture(params=[1, 0], ids=["planned", "literal"])
def mode(request):
    llo.set_option("plan", request.param)
    yield request.param
    llo.set_option("plan", 1)


def rnd(rng, shape, dtype, lo=0.5, hi=2.0):
    if np.issubdtype(dtype, np.floating):
        return rng.uniform(lo, hi, shape).astype(dtype)
    return rng.integers(1, 9, shape).astype(dtype)


@pytest.mark.parametrize("dtype", [F4, F8, I4])
def test_chain_with_broadcast_and_permute(mode, dtype):
    rng = np.random.default_rng(3)
    d = 96
    x = llo.variable(rnd(rng, (d, d), dtype), "x")
    b = llo.variable(rnd(rng, (d,), dtype), "b")
    first = age.exp(x) if dtype != I4 else age.abs(x)
    root = age.add(age.mul(first, age.extend(b, 1, [d])), age.permute(x, [1, 0]))
    check(root, dtype, what="chain", floor=False)   # exp, *, +: true relative error
    if mode == 1:
        st = llo.plan_stats()
        assert st["programs"] == 1 and st["views"] >= 2 and st["generic"] == 0, st


@pytest.mark.parametrize("dtype", [F4, F8, I8])
def test_reductions_of_views(mode, dtype):
    rng = np.random.default_rng(5)
    x = llo.variable(rnd(rng, (40, 72), dtype, 0.0, 1.0), "x")   # ade [72, 40]
    for op in (age.reduce_sum, age.reduce_max, age.reduce_min):
        check(op(x, 1), dtype, what="columns")
        check(op(x, 0), dtype, what="all")
        check(op(age.permute(x, [1, 0]), 1), dtype, what="rows via permute")
    if mode == 1:
        st = llo.plan_stats()
        assert st["reduces"] == 1 and st["programs"] == 0, st  # permute never materialised
    check(age.reduce_prod(age.add(x, x), 1), dtype, scale_tol=50, what="prod of fused")
    check(age.reduce_mean(x), dtype, what="mean")


@pytest.mark.parametrize("dtype", [F4, F8, I4, I8])
@pytest.mark.parametrize("mnk", [(7, 5, 3), (64, 48, 80), (130, 70, 257)])
def test_matmul_lowers_to_gemm(dtype, mnk):
    m, n, k = mnk
    rng = np.random.default_rng(7)
    if np.issubdtype(dtype, np.integer):
        a = rng.integers(-8, 9, (n, k)).astype(dtype)
        b = rng.integers(-8, 9, (k, m)).astype(dtype)
    else:
        a = rng.uniform(-1, 1, (n, k)).astype(dtype)
        b = rng.uniform(-1, 1, (k, m)).astype(dtype)
    va, vb = llo.variable(a, "a"), llo.variable(b, "b")
    root = age.matmul(va, vb)
    llo.set_option("plan", 1)
    got = llo.evaluate(root, np.dtype(dtype))
    st = llo.plan_stats()
    assert st["gemms"] == 1 and st["programs"] == 0, st
    if np.issubdtype(dtype, np.integer):
        np.testing.assert_array_equal(a.astype(np.int64) @ b.astype(np.int64), got.astype(np.int64))
    else:
        exact = a.astype(np.float64) @ b.astype(np.float64)
        scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
        assert (np.abs(exact - got) <= tol_of(dtype) * scale).all()
    if m * n * k <= 64 * 48 * 80:
        # the reference's own materialised [M,N,C] evaluation
        want = oracle.evaluate(root, np.dtype(dtype))
        if np.issubdtype(dtype, np.integer):
            np.testing.assert_array_equal(want, got)
        else:
            scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
            assert (np.abs(want.astype(np.float64) - got) <= 4 * tol_of(dtype) * scale).all()


def test_matmul_int32_1024_cubed_is_bit_exact():
    """int32 randint(-8, 9) at 1024^3 (the sizes of
    test_api.cpp:829-878 scaled up); every sum stays below 2^17, so a float64
    BLAS product is the exact integer answer"""
    rng = np.random.default_rng(8)
    a = rng.integers(-8, 9, (1024, 1024)).astype(I4)
    b = rng.integers(-8, 9, (1024, 1024)).astype(I4)
    root = age.matmul(llo.variable(a, "a"), llo.variable(b, "b"))
    got = llo.evaluate(root, np.dtype(I4))
    want = (a.astype(np.float64) @ b.astype(np.float64)).astype(np.int64)
    np.testing.assert_array_equal(want, got.astype(np.int64))


@pytest.mark.parametrize("dtype", [F4, F8, I4])
def test_matmul_gradients(mode, dtype):
    rng = np.random.default_rng(11)
    n, k, m = 12, 9, 7
    if np.issubdtype(dtype, np.integer):
        a = rng.integers(1, 9, (n, k)).astype(dtype)
        b = rng.integers(1, 9, (k, m)).astype(dtype)
    else:
        a = rng.uniform(0.5, 2, (n, k)).astype(dtype)
        b = rng.uniform(0.5, 2, (k, m)).astype(dtype)
    va, vb = llo.variable(a, "a"), llo.variable(b, "b")
    root = age.matmul(va, vb)
    for target in (va, vb):
        grad = llo.derive(root, target)
        check(grad, dtype, scale_tol=8, what="d matmul")
        if mode == 1:
            st = llo.plan_stats()
            assert st["gemms"] >= 1 or st["reduces"] >= 1, st
    # matmul(c, c): both paths to the same leaf
    c = llo.variable(rnd(rng, (6, 6), dtype), "c")
    check(llo.derive(age.matmul(c, c), c), dtype, scale_tol=8, what="d matmul(c,c)")


def mlp(dtype, batch=24, sizes=(10, 16, 12, 4), seed=13):
    rng = np.random.default_rng(seed)
    x = llo.variable(rng.uniform(0.05, 1, (batch, sizes[0])).astype(dtype), "X")
    y = llo.variable(rng.uniform(0, 1, (batch, sizes[-1])).astype(dtype), "Y")
    params = []
    h = x
    for i in range(len(sizes) - 1):
        w = llo.variable((rng.standard_normal((sizes[i], sizes[i + 1])) / np.sqrt(sizes[i])).astype(dtype), "W%d" % i)
        b = llo.variable((0.1 * rng.standard_normal((sizes[i + 1],))).astype(dtype), "b%d" % i)
        params += [w, b]
        z = age.add(age.matmul(h, w), age.extend(b, 1, [batch]))
        one = llo.scalar(1.0, [batch, sizes[i + 1]])
        h = age.div(one, age.add(one, age.exp(age.neg(z))))
    diff = age.sub(h, y)
    loss = age.div(age.reduce_sum0(age.pow(diff, llo.scalar(2.0, [batch, sizes[-1]]))),
                   llo.scalar(float(batch), []))
    return loss, params


@pytest.mark.parametrize("dtype", [F4, F8])
def test_mlp_loss_and_gradients(mode, dtype):
    loss, params = mlp(dtype)
    check(loss, dtype, scale_tol=20, what="loss")
    for p in params:
        grad = llo.derive(loss, p)
        check(grad, dtype, scale_tol=200, what="grad")
    if mode == 1:
        st = llo.plan_stats()
        assert st["generic"] == 0, st


def test_mlp_gradients_use_gemm_not_the_materialised_product():
    """llo.optimize (OPT_DIV_CANCEL) turns the product rule's (a*b)/a into b, so the
    matmul gradients lower to GEMMs; the rewritten graph is checked against the
    oracle on the graph as derived"""
    loss, params = mlp(F4, batch=32, sizes=(20, 24, 16, 6))
    llo.set_option("plan", 1)
    grad = llo.derive(loss, params[0])
    (better,), report = llo.optimize([grad])
    assert report["div_cancel"] > 0 and report["functors_after"] < report["functors_before"], report
    got = llo.evaluate(better, np.dtype(F4))
    st = llo.plan_stats()
    assert st["gemms"] >= 5, st          # 3 forward + backward contractions
    want = oracle.evaluate(grad, np.dtype(F4)).astype(np.float64)
    assert (np.abs(got - want) <= 2e-4 * np.maximum(np.abs(want), 1.0)).all()


def test_default_evaluation_keeps_the_product_rule(mode):
    """the evaluator's default is the graph as written: (x*y)/x is NaN at x == 0 like
    the reference (llo/src/helper.cpp:18-33); the explicit llo.optimize pass (or the
    planner's legacy `simplify` option) rewrites it to y"""
    a = llo.variable(np.array([[0.0, 2.0], [3.0, 4.0]]), "a")
    b = llo.variable(np.array([[5.0, 6.0], [7.0, 8.0]]), "b")
    grad = llo.derive(age.mul(a, b), a)
    assert llo.get_option("simplify") == 0
    got = llo.evaluate(grad, np.dtype(F8))
    want = oracle.evaluate(grad, np.dtype(F8))
    np.testing.assert_array_equal(np.isnan(want), np.isnan(got))
    np.testing.assert_array_equal(want[~np.isnan(want)], got[~np.isnan(got)])
    assert np.isnan(got[0, 0])
    (better,), report = llo.optimize([grad])
    assert report["div_cancel"] == 1
    np.testing.assert_array_equal(llo.evaluate(better, np.dtype(F8)), [[5.0, 6.0], [7.0, 8.0]])
    if mode == 1:
        llo.set_option("simplify", 1)
        try:
            np.testing.assert_array_equal(llo.evaluate(grad, np.dtype(F8)), [[5.0, 6.0], [7.0, 8.0]])
        finally:
            llo.set_option("simplify", 0)


def test_shared_subgraphs_are_evaluated_once():
    x = llo.variable(np.random.default_rng(17).uniform(0.5, 2, (64, 64)).astype(F4), "x")
    shared = age.exp(age.sin(x))
    root = age.add(age.mul(shared, shared), age.sub(shared, x))
    llo.set_option("plan", 1)
    before = llo.launch_count()
    check(root, F4, what="shared")
    # one program for `shared` (two consumers), one for the root
    assert llo.plan_stats()["programs"] == 2
    assert llo.launch_count() - before == 2


def test_constant_subexpressions_fold_on_the_host():
    """neg(1), 2 - 1, 1 / n of the derivative rules become immediates: no
    launch fills a tensor with a constant, results unchanged"""
    rng = np.random.default_rng(29)
    data = rng.uniform(0.5, 2, (48, 40)).astype(F4)
    x = llo.variable(data, "x")
    shape = [48, 40]
    minus_one = age.neg(llo.scalar(1.0, shape))
    inv_n = age.div(llo.scalar(1.0, shape), llo.scalar(8.0, shape))
    expo = age.sub(llo.scalar(2.0, shape), llo.scalar(1.0, shape))
    root = age.mul(age.mul(age.pow(x, expo), minus_one), inv_n)
    llo.set_option("plan", 1)
    before = llo.launch_count()
    got = check(root, F4, what="folded constants")
    assert llo.launch_count() - before == 1
    assert llo.plan_stats()["programs"] == 1
    np.testing.assert_array_equal(got, -data / np.float32(8.0))   # pow(x, 1) = x, * -1, * 0.125: exact
    # a graph that is constant altogether still evaluates (materialised once)
    check(age.add(minus_one, inv_n), F4, what="constant root")
    check(age.add(minus_one, inv_n), F8, what="constant root f64")


def test_random_expression_graphs(mode):
    rng = np.random.default_rng(19)
    unary = [age.abs, age.neg, age.sin, age.cos, age.exp, age.sqrt, age.round, age.log]
    binary = [age.add, age.sub, age.mul, age.div, age.pow, age.lt, age.gt,
              lambda a, b: age.min([a, b]), lambda a, b: age.max([a, b])]
    for trial in range(12):
        shape = [int(rng.integers(2, 9)) for _ in range(int(rng.integers(1, 4)))]
        leaves = [llo.variable(rng.uniform(0.6, 1.8, shape), "v%d" % i) for i in range(3)]
        pool = list(leaves)
        for _ in range(int(rng.integers(3, 10))):
            if rng.random() < 0.4:
                node = unary[int(rng.integers(len(unary)))](age.abs(pool[int(rng.integers(len(pool)))]))
            else:
                a = pool[int(rng.integers(len(pool)))]
                b = pool[int(rng.integers(len(pool)))]
                node = binary[int(rng.integers(len(binary)))](age.abs(a), age.abs(b))
            # keep magnitudes tame so pow / exp stay finite
            pool.append(age.add(age.div(node, age.add(age.abs(node), leaves[0])), leaves[1]))
        root = pool[-1]
        check(root, F8, scale_tol=100, what="random graph %d" % trial)
        check(llo.derive(root, leaves[0]), F8, scale_tol=1e4, what="random grad %d" % trial)


def test_int_types_through_graphs(mode):
    rng = np.random.default_rng(23)
    for dtype in (np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64):
        a = llo.variable(rng.integers(1, 7, (5, 6)).astype(dtype), "a")
        b = llo.variable(rng.integers(1, 7, (5, 6)).astype(dtype), "b")
        root = age.add(age.mul(a, age.permute(age.permute(b, [1, 0]), [1, 0])),
                       age.extend(age.reduce_max(age.sub(age.add(a, b), b), 1), 1, [5]))
        check(root, dtype, what=str(dtype))
        check(age.lt(a, b), dtype)
        check(age.pow(a, age.eq(a, b)), dtype)


def test_leaf_retype_and_constants(mode):
    data = np.array([[1.5, 2.5, -3.5], [4.25, 5.75, 6.0]])
    v = llo.variable(data, "v")
    for dtype in (np.uint16, np.int32, np.float32):
        check(age.add(v, llo.scalar(2.0, [2, 3])), dtype)
        check(v, dtype)
    check(age.n_elems(v), F8)
    check(age.mul(age.n_dims(v, 0), llo.scalar(3.0, [])), F8)


def test_rand_nodes_draw_per_visit_and_stay_in_range(mode):
    lo = llo.variable(np.array(2.0), "lo")
    hi = llo.variable(np.array(5.0), "hi")
    r = age.functor_extended("RAND_UNIF", [lo, hi], 0, [50, 40])
    llo.seed(5)
    both = llo.evaluate(age.sub(r, r), np.dtype(F8))   # two visits, two draws (llo/eval.hpp:77-87)
    assert np.abs(both).max() > 0.1
    vals = llo.evaluate(r, np.dtype(F8))
    assert vals.min() > 2.0 and vals.max() < 5.0


def direct_convolution(img, kernel):
    """numpy VALID cross-correlation: img [batch, w, h, in], kernel [kw, kh, in, out]
    (what llo/test/ptest.py:441-490 checks against tf.nn.convolution)"""
    nb, iw, ih, _ = img.shape
    kw, kh, _, oc = kernel.shape
    out = np.zeros((nb, iw - kw + 1, ih - kh + 1, oc))
    for a in range(kw):
        for b in range(kh):
            out += np.einsum("nwhc,co->nwho", img[:, a:a + iw - kw + 1, b:b + ih - kh + 1, :], kernel[a, b])
    return out


def test_convolution_lowers_to_one_gathered_gemm(mode):
    """age.convolution (llo/src/helper.cpp:102-202): PROD of two mapped arguments over a
    7-d box, summed over (in, kh, kw).  The image's map has two sources per coordinate
    (h = oh + kh); the planner keeps it as an integer view, gathers the image once into a
    [rows][k] matrix (im2col) and runs ONE GEMM: no generic per-element functor is left."""
    rng = np.random.default_rng(29)
    img = rng.uniform(0, 1, (2, 5, 5, 3))          # [batch, w, h, in]
    kernel = rng.uniform(0, 1, (3, 3, 3, 4))       # [kw, kh, in, out]
    vimg, vk = llo.variable(img, "img"), llo.variable(kernel, "k")
    root = age.convolution(vimg, vk)
    got = check(root, F8, scale_tol=100, what="convolution")
    if mode == 1:
        st = llo.plan_stats()
        assert st["generic"] == 0 and st["gemms"] == 1 and st["gathers"] == 1, st
    assert np.abs(got - direct_convolution(img, kernel)).max() <= 1e-12 * np.abs(got).max()
    check(llo.derive(root, vk), F8, scale_tol=100, what="d convolution / d kernel")
    if mode == 1:
        assert llo.plan_stats()["generic"] == 0, llo.plan_stats()
    # the gradient w.r.t. the image scatters through the reversed two-source map (col2im):
    # it stays on the generic per-element path, checked against the oracle all the same
    check(llo.derive(root, vimg), F8, scale_tol=100, what="d convolution / d image")


@pytest.mark.parametrize("dtype", [F4, F8])
def test_convolution_at_a_real_size_matches_a_direct_convolution(dtype):
    """64 x 64 images, 16 -> 32 channels, 3 x 3 window, batch 8 (wide DimT: the reference's
    uint8_t dims could not hold the 7-d box): 6.1e8 flop through im2col + the TMA GEMM"""
    rng = np.random.default_rng(31)
    img = rng.uniform(-1, 1, (8, 64, 64, 16)).astype(dtype)
    kernel = rng.uniform(-1, 1, (3, 3, 16, 32)).astype(dtype)
    root = age.convolution(llo.variable(img, "img"), llo.variable(kernel, "k"))
    llo.set_option("plan", 1)
    got = llo.evaluate(root, np.dtype(dtype))
    st = llo.plan_stats()
    assert st["generic"] == 0 and st["gemms"] == 1 and st["gathers"] == 1, st
    want = direct_convolution(img.astype(np.float64), kernel.astype(np.float64))
    bound = direct_convolution(np.abs(img).astype(np.float64), np.abs(kernel).astype(np.float64))
    assert got.shape == want.shape
    assert (np.abs(got - want) <= (1e-6 if dtype == F4 else 1e-12) * bound).all()


def test_tf32x3_option_routes_matmul_to_the_tensor_cores():
    rng = np.random.default_rng(61)
    a = rng.uniform(-1, 1, (256, 320)).astype(F4)
    b = rng.uniform(-1, 1, (320, 384)).astype(F4)
    root = age.matmul(llo.variable(a, "a"), llo.variable(b, "b"))
    exact = llo.evaluate(root, np.dtype(F4))
    llo.set_option("tf32x3", 1)
    try:
        before = llo.launch_count()
        fast = llo.evaluate(root, np.dtype(F4))
        assert llo.launch_count() - before == 3     # 2 splits + tcgen05 kernel
    finally:
        llo.set_option("tf32x3", 0)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    assert (np.abs(ref - exact) <= 1e-6 * scale).all()
    assert (np.abs(ref - fast) <= 1e-6 * scale).all()


@pytest.mark.parametrize("side,extra", [(200, ()), (256, ()), (130, (3,)), (64, ()), (65, (2,)), (192, (2,)),
                                        (320, (2, 2))])
def test_program_with_a_tensor_and_its_transpose(side, extra):
    """f(x, permute(x)): the tile engine computes output tiles (i,j) and (j,i)
    from the two input tiles it staged once (vm_pair4_kernel); ragged sizes,
    extra dims and the single-tile case included"""
    rng = np.random.default_rng(71)
    shape = extra + (side, side)                      # numpy order: ade dims 0, 1 are the last two
    x = rng.uniform(0.5, 2.0, shape).astype(F4)
    b = rng.uniform(0.5, 2.0, (side,)).astype(F4)
    vx, vb = llo.variable(x, "x"), llo.variable(b, "b")
    ext = age.extend(vb, 1, list(reversed(shape[:-1])))
    order = [1, 0] + list(range(2, len(shape)))
    root = age.add(age.mul(age.exp(vx), ext), age.permute(vx, order))
    got = llo.evaluate(root, np.dtype(F4))
    want = oracle.evaluate(root, np.dtype(F4))
    assert got.shape == want.shape == x.shape
    assert np.abs(got.astype(np.float64) - want).max() <= 1e-6 * np.abs(want).max()
    expect = (np.exp(x.astype(np.float64)) * b) + np.swapaxes(x, -1, -2)
    assert np.abs(got - expect).max() <= 2e-6 * np.abs(expect).max()
    # integer: bit-exact
    xi = rng.integers(-50, 50, shape).astype(np.int32)
    vi = llo.variable(xi, "xi")
    rooti = age.sub(age.mul(vi, vi), age.permute(vi, order))
    np.testing.assert_array_equal(llo.evaluate(rooti, np.dtype(np.int32)),
                                  xi * xi - np.swapaxes(xi, -1, -2))


def test_pipelined_assign_and_evaluate_match_the_synchronous_path():
    """Variable.assign_async + llo.evaluate_async: uploads on the copy lane into
    fresh buffers, downloads on the other lane; results of a pipelined sequence
    over CHANGING inputs equal the synchronous ones, step by step"""
    rng = np.random.default_rng(83)
    shape = (512, 768)
    b = rng.uniform(0.5, 2.0, (shape[1],)).astype(F4)
    vx = llo.variable(np.zeros(shape, F4), "x")
    vb = llo.variable(b, "b")
    root = age.add(age.mul(age.exp(vx), age.extend(vb, 1, [shape[0]])), age.neg(vx))
    inputs = []
    for k in range(5):
        xp = llo.pinned_empty(list(shape), np.dtype(F4))
        xp[...] = rng.uniform(0.5, 2.0, shape).astype(F4)
        inputs.append(xp)
    want = []
    for xp in inputs:
        vx.assign(xp)
        want.append(llo.evaluate(root, np.dtype(F4)).copy())
    pending, got = None, []
    for xp in inputs:
        vx.assign_async(xp)
        nxt = llo.evaluate_async(root, np.dtype(F4))
        if pending is not None:
            got.append(pending.result())
        pending = nxt
    got.append(pending.result())
    assert len(got) == len(want)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    # the variable's host copy follows the last asynchronous assignment
    llo.sync()
    vx.assign(inputs[0])
    np.testing.assert_array_equal(llo.evaluate(vx, np.dtype(F4)), inputs[0])
    with pytest.raises(RuntimeError):
        vx.assign_async(np.zeros((3, 3), F4))


def test_random_subgraphs_are_drawn_per_visit_in_both_modes(mode):
    """a RAND_* node reached through two parents is drawn twice (the reference re-evaluates
    every child, llo/eval.hpp:77-87): neither the planner nor the literal evaluator's memo
    may share it or anything above it"""
    lo, hi = llo.scalar(0.0, [64, 64]), llo.scalar(1.0, [64, 64])
    r = age.rand_unif(lo, hi)
    diff = age.sub(age.neg(r), age.neg(r))          # zero everywhere iff r were shared
    llo.seed(5)
    got = llo.evaluate(diff, np.dtype(F8))
    assert np.count_nonzero(got) > 0.99 * got.size
    assert np.abs(got).max() < 1.0
