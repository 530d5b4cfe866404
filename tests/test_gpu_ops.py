"""GPU parity, operator level: the CUDA C-ABI against the CPU oracle on the
same seeded inputs.  Integer / index / comparison results must be bit-exact;
fp32 within 1e-6 relative, fp64 within 1e-12 relative (north star)."""
import numpy as np
import pytest

import util
from util import capi

pytestmark = pytest.mark.gpu

ALL_DTYPES = [np.float64, np.float32, np.int8, np.uint8, np.int16, np.uint16,
              np.int32, np.uint32, np.int64, np.uint64]
INT_DTYPES = [t for t in ALL_DTYPES if np.issubdtype(t, np.integer)]
UNARY = ["ABS", "NEG", "SIN", "COS", "TAN", "EXP", "LOG", "SQRT", "ROUND"]
BINARY = ["POW", "SUB", "DIV", "EQ", "NEQ", "LT", "GT"]
NNARY = ["SUM", "PROD", "MIN", "MAX"]


def tolerance(dtype):
    if dtype == np.float32:
        return 1e-6
    if dtype == np.float64:
        return 1e-12
    return 0


# operators whose result can pass through zero with a finite slope keep an absolute floor
# at unit scale (the north star's 1e-6 / 1e-12 is relative; sin / cos / tan / log of a value
# next to one of their zeros has an unbounded RELATIVE condition number, so two correctly
# rounded libm's differ by more than that there); everything else is held to true relative
# error (IEEE + - * / sqrt are bit-identical, exp / pow within their libm ulps)
ABSOLUTE_FLOOR_OPS = ("SIN", "COS", "TAN", "LOG")


def assert_parity(expect, got, dtype, what="", floor=True):
    tol = tolerance(dtype)
    if tol == 0:
        np.testing.assert_array_equal(expect, got, err_msg=what)
        return
    expect = expect.astype(np.float64)
    got = got.astype(np.float64)
    same_special = (np.isnan(expect) == np.isnan(got)) & (np.isinf(expect) == np.isinf(got))
    assert same_special.all(), what
    fin = np.isfinite(expect)
    # relative to the reference value, with an absolute floor of the same
    # factor at unit scale (sin / cos / log near their zeros)
    err = np.abs(expect[fin] - got[fin])
    scale = np.maximum(np.abs(expect[fin]), 1.0) if floor else np.abs(expect[fin])
    bound = tol * scale
    assert (err <= bound).all(), "%s: max err %g" % (what, (err / np.maximum(scale, 1e-300)).max())


def sample(rng, dtype, n, lo=0.5, hi=3.0, positive=True):
    if np.issubdtype(dtype, np.floating):
        return rng.uniform(lo, hi, n).astype(dtype)
    info = np.iinfo(dtype)
    top = min(info.max, 12)
    low = 1 if (positive or info.min == 0) else -min(12, -info.min)
    return rng.integers(low, top + 1, n).astype(dtype)


@pytest.mark.parametrize("dst", ALL_DTYPES)
@pytest.mark.parametrize("src", ALL_DTYPES)
def test_convert(ctx, src, dst):
    rng = np.random.default_rng(7)
    n = 1031
    if np.issubdtype(src, np.floating):
        data = rng.uniform(0, 100, n).astype(src)
    else:
        data = rng.integers(0, min(np.iinfo(src).max, 100) + 1, n).astype(src)
    expect = util.oracle_convert(data, dst)
    ptr = ctx.upload(data)
    out = ctx.alloc(n * np.dtype(dst).itemsize)
    capi.check(ctx.lib.llo_cuda_convert(ctx.handle, out, capi.dtype_code(dst), ptr, capi.dtype_code(src), n))
    got = ctx.download(out, n, dst)
    ctx.free(ptr); ctx.free(out)
    np.testing.assert_array_equal(expect, got)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("op", UNARY)
def test_unary_identity(ctx, op, dtype):
    rng = np.random.default_rng(11)
    shape = [5, 7, 3]
    data = sample(rng, dtype, util.nelems(shape), positive=(op in ("LOG", "SQRT")))
    args = [(data, shape)]
    if op == "NEG" and np.issubdtype(dtype, np.unsignedinteger):
        with pytest.raises(util.OracleUnsupported):
            util.oracle_op_exec(op, dtype, shape, args)
        with pytest.raises(capi.UnsupportedTypedOp):
            util.gpu_op_exec(ctx, op, dtype, shape, args)
        return
    expect = util.oracle_op_exec(op, dtype, shape, args)
    got = util.gpu_op_exec(ctx, op, dtype, shape, args)
    assert_parity(expect, got, dtype, "%s %s" % (op, dtype.__name__), floor=op in ABSOLUTE_FLOOR_OPS)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("op", BINARY)
def test_binary_identity(ctx, op, dtype):
    rng = np.random.default_rng(13)
    shape = [6, 5, 4]
    n = util.nelems(shape)
    a = sample(rng, dtype, n)
    b = sample(rng, dtype, n)
    if op == "POW" and np.issubdtype(dtype, np.integer):
        a = (a % 5 + 1).astype(dtype)
        b = (b % 4).astype(dtype)
    if op in ("EQ", "NEQ"):
        b[::3] = a[::3]
    args = [(a, shape), (b, shape)]
    expect = util.oracle_op_exec(op, dtype, shape, args)
    got = util.gpu_op_exec(ctx, op, dtype, shape, args)
    assert_parity(expect, got, dtype, "%s %s" % (op, dtype.__name__), floor=False)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("op", NNARY)
@pytest.mark.parametrize("nargs", [1, 3, 11])
def test_nnary_identity(ctx, op, dtype, nargs):
    rng = np.random.default_rng(17)
    shape = [9, 4, 2]
    n = util.nelems(shape)
    args = [(sample(rng, dtype, n, 0.5, 1.5, positive=False) if op != "PROD"
             else (sample(rng, dtype, n, 0.8, 1.2) if np.issubdtype(dtype, np.floating)
                   else rng.integers(1, 3, n).astype(dtype)), shape) for _ in range(nargs)]
    expect = util.oracle_op_exec(op, dtype, shape, args)
    got = util.gpu_op_exec(ctx, op, dtype, shape, args)
    assert_parity(expect, got, dtype, "%s x%d %s" % (op, nargs, dtype.__name__), floor=False)


# ---- coordinate maps --------------------------------------------------------

def extend_pull(rank, ext):
    """coorder of ade::extend_map (pull): diag(1/ext) on dims rank.."""
    m = np.eye(9)
    for i, e in enumerate(ext):
        m[rank + i][rank + i] = 1.0 / e
    return m


def reduce_push(rank, red):
    m = np.eye(9)
    for i, r in enumerate(red):
        m[rank + i][rank + i] = 1.0 / r
    return m


def permute_pull(order):
    """coorder of ade::permute_map: inverse (= transpose) of M[order[i]][i] = 1"""
    full = list(order) + [d for d in range(8) if d not in order]
    m = np.zeros((9, 9))
    for i in range(8):
        m[full[i]][i] = 1
    m[8][8] = 1
    return m.T.copy()


def flip_pull(dim):
    m = np.eye(9)
    m[dim][dim] = -1
    m[8][dim] = -1
    return m


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.uint8, np.int64])
def test_broadcast_permute_flip_views(ctx, dtype):
    rng = np.random.default_rng(19)
    # extend: [6,5] -> [6,5,4,3]
    src = sample(rng, dtype, 30)
    outshape = [6, 5, 4, 3]
    args = [(src, [6, 5], extend_pull(2, [4, 3]), False)]
    np.testing.assert_array_equal(util.oracle_op_exec("SUM", dtype, outshape, args),
                                  util.gpu_op_exec(ctx, "SUM", dtype, outshape, args))
    # extend in the middle of a binary: out = a - broadcast(b)
    a = sample(rng, dtype, 360)
    args = [(a, outshape), (src, [6, 5], extend_pull(2, [4, 3]), False)]
    np.testing.assert_array_equal(util.oracle_op_exec("SUB", dtype, outshape, args),
                                  util.gpu_op_exec(ctx, "SUB", dtype, outshape, args))
    # permute {2,0,1,3}: arg [6,5,4,3] -> out [4,6,5,3]
    order = [2, 0, 1]
    pout = [4, 6, 5, 3]
    args = [(a, outshape, permute_pull(order), False)]
    np.testing.assert_array_equal(util.oracle_op_exec("SUM", dtype, pout, args),
                                  util.gpu_op_exec(ctx, "SUM", dtype, pout, args))
    # flip dim 1 fused with a product against a permuted view of itself
    args = [(a, outshape, flip_pull(1), False), (a, outshape)]
    np.testing.assert_array_equal(util.oracle_op_exec("PROD", dtype, outshape, args),
                                  util.gpu_op_exec(ctx, "PROD", dtype, outshape, args))
    # transpose of a matrix with awkward sizes
    mat = sample(rng, dtype, 37 * 53)
    args = [(mat, [37, 53], permute_pull([1, 0]), False)]
    np.testing.assert_array_equal(util.oracle_op_exec("SUM", dtype, [53, 37], args),
                                  util.gpu_op_exec(ctx, "SUM", dtype, [53, 37], args))


@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("op", NNARY)
@pytest.mark.parametrize("first", [0, 1, 2])
def test_reduce_push(ctx, op, dtype, first):
    rng = np.random.default_rng(23)
    shape = [7, 6, 5]
    n = util.nelems(shape)
    if op == "PROD":
        data = sample(rng, dtype, n, 0.9, 1.1) if np.issubdtype(dtype, np.floating) \
            else rng.integers(1, 3, n).astype(dtype)
    else:
        data = sample(rng, dtype, n, 0.0, 1.0, positive=False)
    outshape = shape[:first] + [1] * (3 - first)
    args = [(data, shape, reduce_push(first, shape[first:]), True)]
    expect = util.oracle_op_exec(op, dtype, outshape, args)
    got = util.gpu_op_exec(ctx, op, dtype, outshape, args)
    assert_parity(expect, got, dtype, "reduce %s dim>=%d %s" % (op, first, dtype.__name__))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,first", [([300, 70], 1), ([300, 70], 0), ([9, 1100], 1), ([64, 33, 17], 1)])
def test_reduce_sum_sequential_mode_is_bit_exact(ctx, dtype, shape, first):
    """mode 1 reproduces the reference's summation order bit for bit"""
    rng = np.random.default_rng(29)
    data = rng.uniform(0, 1, util.nelems(shape)).astype(dtype)
    outshape = shape[:first] + [1] * (len(shape) - first)
    args = [(data, shape, reduce_push(first, shape[first:]), True)]
    expect = util.oracle_op_exec("SUM", dtype, outshape, args)
    capi.check(ctx.lib.llo_cuda_set_reduce_mode(ctx.handle, 1))
    try:
        got = util.gpu_op_exec(ctx, "SUM", dtype, outshape, args)
    finally:
        capi.check(ctx.lib.llo_cuda_set_reduce_mode(ctx.handle, 0))
    np.testing.assert_array_equal(expect, got)
    # the fixed-shape tree is compared with the exactly-rounded sum: a long
    # sequential fp32 sum itself drifts by more than 1e-6
    exact = util.oracle_op_exec("SUM", np.float64, outshape,
                                [(data.astype(np.float64), shape, args[0][2], True)])
    tree = util.gpu_op_exec(ctx, "SUM", dtype, outshape, args)
    assert_parity(exact, tree, dtype, "tree mode")
    again = util.gpu_op_exec(ctx, "SUM", dtype, outshape, args)
    np.testing.assert_array_equal(tree, again)  # deterministic


def test_reduce_entry_point(ctx):
    rng = np.random.default_rng(31)
    shape = [33, 20, 9]
    data = rng.integers(-50, 50, util.nelems(shape)).astype(np.int64)
    ptr = ctx.upload(data)
    for first in (0, 1, 2):
        nout = util.nelems(shape[:first])
        out = ctx.alloc(nout * 8)
        capi.check(ctx.lib.llo_cuda_reduce(ctx.handle, capi.OPCODES["SUM"], capi.dtype_code(np.int64),
                                           out, ptr, capi.shape8(shape), first, 0))
        got = ctx.download(out, nout, np.int64)
        ctx.free(out)
        expect = data.reshape(9, 20, 33).sum(axis=tuple(range(0, 3 - first))).reshape(-1)
        np.testing.assert_array_equal(expect, got)
    ctx.free(ptr)


def test_mixed_push_pull_sum(ctx):
    """SUM{reduce(x), y}: a push argument followed by a pull argument"""
    rng = np.random.default_rng(37)
    x = rng.integers(-9, 9, 6 * 5 * 4).astype(np.int32)
    y = rng.integers(-9, 9, 6).astype(np.int32)
    args = [(x, [6, 5, 4], reduce_push(1, [5, 4]), True), (y, [6])]
    np.testing.assert_array_equal(util.oracle_op_exec("SUM", np.int32, [6], args),
                                  util.gpu_op_exec(ctx, "SUM", np.int32, [6], args))
    args = [(y, [6]), (x, [6, 5, 4], reduce_push(1, [5, 4]), True)]
    np.testing.assert_array_equal(util.oracle_op_exec("SUM", np.int32, [6], args),
                                  util.gpu_op_exec(ctx, "SUM", np.int32, [6], args))


def test_general_map_repeat_each(ctx):
    """a fractional scale onto a real dimension (nearest-neighbour upsample)
    is not a stride pattern: general path, still bit-exact"""
    rng = np.random.default_rng(41)
    src = rng.integers(0, 100, 4 * 3).astype(np.int32)
    m = np.eye(9)
    m[1][1] = 1.0 / 3
    args = [(src, [4, 3], m, False)]
    np.testing.assert_array_equal(util.oracle_op_exec("SUM", np.int32, [4, 9], args),
                                  util.gpu_op_exec(ctx, "SUM", np.int32, [4, 9], args))
    # and its adjoint: a push that folds triples of rows
    big = rng.integers(0, 100, 4 * 9).astype(np.int32)
    args = [(big, [4, 9], m, True)]
    np.testing.assert_array_equal(util.oracle_op_exec("SUM", np.int32, [4, 3], args),
                                  util.gpu_op_exec(ctx, "SUM", np.int32, [4, 3], args))


def test_bad_coordinate_is_fatal(ctx):
    src = np.arange(12, dtype=np.float64)
    m = np.eye(9)
    m[8][0] = 2  # shifts dim 0 out of range
    with pytest.raises(util.OracleError):
        util.oracle_op_exec("SUM", np.float64, [4, 3], [(src, [4, 3], m, False)])
    with pytest.raises(capi.LloError):
        util.gpu_op_exec(ctx, "SUM", np.float64, [4, 3], [(src, [4, 3], m, False)])


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
def test_rand_statistics(ctx, dtype):
    shape = [31, 27, 14]
    n = util.nelems(shape)
    capi.check(ctx.lib.llo_cuda_seed(ctx.handle, 1234))
    lo = np.array([2], dtype=dtype)
    hi = np.array([9], dtype=dtype)
    ext = extend_pull(0, shape)
    got = util.gpu_op_exec(ctx, "RAND_UNIF", dtype, shape, [(lo, [], ext, False), (hi, [], ext, False)])
    assert got.min() >= 2 and got.max() <= 9
    if np.issubdtype(dtype, np.floating):
        assert got.max() < 9
        assert abs(got.mean() - 5.5) < 0.1
    else:
        assert set(np.unique(got)) == set(range(2, 10))  # inclusive, like uniform_int_distribution
    if np.issubdtype(dtype, np.floating):
        mu = np.array([3.2234], dtype=dtype)
        sd = np.array([1.2547977589], dtype=dtype)
        got = util.gpu_op_exec(ctx, "RAND_NORM", dtype, shape, [(mu, [], ext, False), (sd, [], ext, False)])
        assert abs(got.mean() - 3.2234) / 3.2234 < 0.05
        assert abs(got.var() - 1.2547977589 ** 2) / 1.2547977589 ** 2 < 0.05
    else:
        with pytest.raises(capi.UnsupportedTypedOp):
            util.gpu_op_exec(ctx, "RAND_NORM", dtype, shape, [(lo, [], ext, False), (hi, [], ext, False)])
    trials = np.array([20], dtype=dtype)
    prob = np.array([0.3], dtype=np.float64)
    got = util.gpu_op_exec(ctx, "RAND_BINO", dtype, shape, [(trials, [], ext, False), (prob, [], ext, False)])
    assert abs(got.mean() - 6.0) / 6.0 < 0.05
    assert abs(got.var() - 4.2) / 4.2 < 0.1
    # same seed, same stream
    capi.check(ctx.lib.llo_cuda_seed(ctx.handle, 99))
    a = util.gpu_op_exec(ctx, "RAND_UNIF", dtype, shape, [(lo, [], ext, False), (hi, [], ext, False)])
    capi.check(ctx.lib.llo_cuda_seed(ctx.handle, 99))
    b = util.gpu_op_exec(ctx, "RAND_UNIF", dtype, shape, [(lo, [], ext, False), (hi, [], ext, False)])
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("layout", ["nn", "nt", "tn", "tt"])
def test_gemm(ctx, dtype, layout):
    rng = np.random.default_rng(43)
    M, N, K = 67, 45, 83
    if np.issubdtype(dtype, np.integer):
        A = rng.integers(-8, 9, (M, K)).astype(dtype)
        B = rng.integers(-8, 9, (K, N)).astype(dtype)
    else:
        A = rng.uniform(-1, 1, (M, K)).astype(dtype)
        B = rng.uniform(-1, 1, (K, N)).astype(dtype)
    a_store = A if layout[0] == "n" else np.ascontiguousarray(A.T)
    b_store = B if layout[1] == "n" else np.ascontiguousarray(B.T)
    a_rs, a_cs = (K, 1) if layout[0] == "n" else (1, M)
    b_rs, b_cs = (N, 1) if layout[1] == "n" else (1, K)
    pa = ctx.upload(a_store); pb = ctx.upload(b_store)
    pc = ctx.alloc(M * N * np.dtype(dtype).itemsize)
    capi.check(ctx.lib.llo_cuda_gemm(ctx.handle, capi.dtype_code(dtype), M, N, K,
                                     pa, a_rs, a_cs, pb, b_rs, b_cs, pc, N, 1, 0))
    got = ctx.download(pc, M * N, dtype).reshape(M, N)
    for p in (pa, pb, pc):
        ctx.free(p)
    if np.issubdtype(dtype, np.integer):
        np.testing.assert_array_equal(A.astype(np.int64) @ B.astype(np.int64), got.astype(np.int64))
    else:
        expect = A.astype(np.float64) @ B.astype(np.float64)
        tol = 1e-6 if dtype == np.float32 else 1e-12
        scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
        assert (np.abs(expect - got) <= tol * scale).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("layout", ["nn", "nt", "tn", "tt"])
@pytest.mark.parametrize("mnk", [
    (2048, 1280, 96),     # 160 short tiles: the persistent kernel, tail tiles split along k
    (2000, 1300, 100),    # the same, ragged edges
    (2432, 2432, 40),     # 361 tiles: two full rounds + a split tail
    (128, 256, 4096),     # two tiles, long k: split-K with the fixed-order fold
    (1000, 10, 3000),     # skinny output: warp-per-rows / k-split column kernels (skinny.cu)
    (10, 1000, 3000),     # skinny the other way round (transposed problem)
    (4100, 16, 300),      # widest skinny output, ragged rows
    (3000, 1, 1000),      # matrix-vector
    (784, 1024, 4096),    # a 16-row last tile row: its dead warps skip the math, split-K by the cost model
    (4099, 10, 1024),     # persistent skinny rows kernel with B staged whole, ragged last row group
    (1024, 10, 8191),     # skinny columns, k ranges that do not divide
])
def test_gemm_schedules(ctx, dtype, layout, mnk):
    """every work decomposition of the TMA GEMMs (tile grid, split-K, persistent
    with a split tail) against a float64 product"""
    rng = np.random.default_rng(47)
    M, N, K = mnk
    A = rng.uniform(-1, 1, (M, K)).astype(dtype)
    B = rng.uniform(-1, 1, (K, N)).astype(dtype)
    a_store = A if layout[0] == "n" else np.ascontiguousarray(A.T)
    b_store = B if layout[1] == "n" else np.ascontiguousarray(B.T)
    a_rs, a_cs = (K, 1) if layout[0] == "n" else (1, M)
    b_rs, b_cs = (N, 1) if layout[1] == "n" else (1, K)
    pa = ctx.upload(a_store); pb = ctx.upload(b_store)
    # C between canary bands (compute-sanitizer is closed on this pool): edge tiles, tail
    # splits and folds must not write outside it
    nbytes = M * N * np.dtype(dtype).itemsize
    padded = (nbytes + 15) // 16 * 16
    block = ctx.alloc(padded + 2 * util.GUARD)
    capi.check(ctx.lib.llo_cuda_memset(ctx.handle, block, util.GUARD_BYTE, padded + 2 * util.GUARD))
    pc = block + util.GUARD
    got = []
    for _ in range(2):
        capi.check(ctx.lib.llo_cuda_gemm(ctx.handle, capi.dtype_code(dtype), M, N, K,
                                         pa, a_rs, a_cs, pb, b_rs, b_cs, pc, N, 1, 0))
        whole = ctx.download(block, padded + 2 * util.GUARD, np.uint8)
        assert (whole[:util.GUARD] == util.GUARD_BYTE).all() and \
            (whole[util.GUARD + nbytes:] == util.GUARD_BYTE).all(), "gemm wrote outside C"
        got.append(whole[util.GUARD:util.GUARD + nbytes].view(dtype).reshape(M, N).copy())
    for p in (pa, pb, block):
        ctx.free(p)
    expect = A.astype(np.float64) @ B.astype(np.float64)
    tol = 1e-6 if dtype == np.float32 else 1e-12
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    assert (np.abs(expect - got[0]) <= tol * scale).all()
    np.testing.assert_array_equal(got[0], got[1])   # fixed summation order: run-to-run identical


@pytest.mark.parametrize("layout", ["nn", "nt", "tn", "tt"])
@pytest.mark.parametrize("mnk", [(256, 384, 512), (300, 200, 1000), (129, 131, 2085)])
def test_gemm_3xtf32_tensor_core_path(ctx, layout, mnk):
    """opt-in precision 1 (tcgen05 kind::tf32, 3 products per k): within 1e-6
    of sum |a||b| of the float64 product, i.e. as close as the exact FFMA path
    needs to be (north star: 1e-6 relative for fp32)"""
    rng = np.random.default_rng(53)
    M, N, K = mnk
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    a_store = A if layout[0] == "n" else np.ascontiguousarray(A.T)
    b_store = B if layout[1] == "n" else np.ascontiguousarray(B.T)
    a_rs, a_cs = (K, 1) if layout[0] == "n" else (1, M)
    b_rs, b_cs = (N, 1) if layout[1] == "n" else (1, K)
    pa = ctx.upload(a_store); pb = ctx.upload(b_store)
    pc = ctx.alloc(M * N * 4)
    before = ctx.launch_count()
    capi.check(ctx.lib.llo_cuda_gemm(ctx.handle, capi.dtype_code(np.float32), M, N, K,
                                     pa, a_rs, a_cs, pb, b_rs, b_cs, pc, N, 1, 1))
    assert ctx.launch_count() - before == 3          # 2 operand splits + the tensor-core kernel
    got = ctx.download(pc, M * N, np.float32).reshape(M, N)
    for p in (pa, pb, pc):
        ctx.free(p)
    expect = A.astype(np.float64) @ B.astype(np.float64)
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    assert (np.abs(expect - got) <= 1e-6 * scale).all()


def test_elementwise_program_entry(ctx):
    """llo_cuda_elementwise: out = exp(x) * b[broadcast] + x[transposed]"""
    import ctypes as C
    rng = np.random.default_rng(47)
    d0, d1 = 48, 40
    x = rng.uniform(0.5, 2, d0 * d0).astype(np.float32)   # square so the transpose has out's shape
    b = rng.uniform(0.5, 2, d0).astype(np.float32)
    px = ctx.upload(x); pb = ctx.upload(b)
    out = ctx.alloc(d0 * d0 * 4)
    views = (capi.View * 3)()
    views[0].data = px; views[0].stride = (C.c_int64 * 8)(1, d0, 0, 0, 0, 0, 0, 0)
    views[1].data = pb; views[1].stride = (C.c_int64 * 8)(1, 0, 0, 0, 0, 0, 0, 0)
    views[2].data = px; views[2].stride = (C.c_int64 * 8)(d0, 1, 0, 0, 0, 0, 0, 0)
    prog = (capi.Insn * 7)(*[capi.Insn(*p) for p in [
        (capi.VM_LOAD, 0), (capi.OPCODES["EXP"], 0), (capi.VM_LOAD, 1), (capi.OPCODES["PROD"], 0),
        (capi.VM_LOAD, 2), (capi.OPCODES["SUM"], 0)]] + [capi.Insn(0, 0)])
    capi.check(ctx.lib.llo_cuda_elementwise(ctx.handle, capi.dtype_code(np.float32), out,
                                            capi.shape8([d0, d0]), views, 3, None, 0, prog, 6))
    got = ctx.download(out, d0 * d0, np.float32).reshape(d0, d0)
    X = x.reshape(d0, d0)
    e = util.oracle_op_exec("EXP", np.float32, [d0, d0], [(x, [d0, d0])]).reshape(d0, d0)
    expect = (e * b[None, :]).astype(np.float32) + X.T
    assert_parity(expect.reshape(-1), got.reshape(-1), np.float32, "fused chain")
    for p in (px, pb, out):
        ctx.free(p)
    del d1
