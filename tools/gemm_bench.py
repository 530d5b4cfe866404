"""GEMM through the C-ABI: correctness against a float64 numpy product for every
operand layout and ragged sizes, then timings (CUDA events).
usage: python tools/gemm_bench.py [--check] [--size 8192] [--dtype f32|f64] [--iters 5] [--precision 0|1]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cortenn_b200 import capi  # noqa: E402


def run_gemm(ctx, dtype, A, B, layout, precision=0, c_colmajor=False):
    M, K = A.shape
    N = B.shape[1]
    a_store = A if layout[0] == "n" else np.ascontiguousarray(A.T)
    b_store = B if layout[1] == "n" else np.ascontiguousarray(B.T)
    a_rs, a_cs = (K, 1) if layout[0] == "n" else (1, M)
    b_rs, b_cs = (N, 1) if layout[1] == "n" else (1, K)
    pa, pb = ctx.upload(a_store), ctx.upload(b_store)
    # compute-sanitizer is closed on this pool: C sits between two canary bands that
    # must come back untouched (out-of-bounds writes of edge tiles / split-K folds)
    guard = 256
    nbytes = M * N * np.dtype(dtype).itemsize
    padded = (nbytes + 15) // 16 * 16
    block = ctx.alloc(padded + 2 * guard)
    capi.check(ctx.lib.llo_cuda_memset(ctx.handle, block, 0xA5, padded + 2 * guard))
    pc = block + guard
    c_rs, c_cs = (1, M) if c_colmajor else (N, 1)
    capi.check(ctx.lib.llo_cuda_gemm(ctx.handle, capi.dtype_code(dtype), M, N, K,
                                     pa, a_rs, a_cs, pb, b_rs, b_cs, pc, c_rs, c_cs, precision))
    whole = ctx.download(block, padded + 2 * guard, np.uint8)
    assert (whole[:guard] == 0xA5).all() and (whole[guard + nbytes:] == 0xA5).all(), \
        "gemm wrote outside C (M=%d N=%d K=%d %s)" % (M, N, K, layout)
    got = whole[guard:guard + nbytes].view(dtype).copy()
    got = got.reshape(N, M).T if c_colmajor else got.reshape(M, N)
    for p in (pa, pb, block):
        ctx.free(p)
    return got


def check(ctx, dtype, precision):
    rng = np.random.default_rng(7)
    worst = 0.0
    sizes = [(128, 128, 32), (256, 128, 64), (200, 136, 100), (129, 260, 36), (64, 32, 1000),
             (1000, 36, 64), (333, 444, 555), (32, 32, 32), (1024, 1024, 1024),
             (4096, 10, 256), (512, 10, 8192), (4096, 256, 10), (100, 100, 10000), (1, 64, 1024),
             (64, 1, 1024), (300, 200, 1), (784, 1024, 4096), (7, 5, 3000)]
    for (M, N, K) in sizes:
        A = rng.uniform(-1, 1, (M, K)).astype(dtype)
        B = rng.uniform(-1, 1, (K, N)).astype(dtype)
        expect = A.astype(np.float64) @ B.astype(np.float64)
        scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
        for layout in ("nn", "nt", "tn", "tt"):
            for colmajor in (False, True):
                before = ctx.launch_count()
                got = run_gemm(ctx, dtype, A, B, layout, precision, colmajor)
                err = float((np.abs(expect - got) / scale).max())
                worst = max(worst, err)
                tol = (1e-6 if precision == 0 else 2e-6) if dtype == np.float32 else 1e-12
                status = "ok" if err <= tol else "FAIL"
                print("%-5s M=%4d N=%4d K=%4d %s colmajor=%d  rel err %.3e  launches %d" % (
                    status, M, N, K, layout, colmajor, err, ctx.launch_count() - before), flush=True)
                if err > tol:
                    bad = np.argwhere(np.abs(expect - got) / scale > tol)
                    print("   first bad", bad[:5].tolist(), "of", len(bad))
                    return False
    print("worst relative error", worst)
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--mnk", type=str, default="")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", type=int, default=0)
    ap.add_argument("--layouts", default="nn")
    args = ap.parse_args()
    dtype = np.float32 if args.dtype == "f32" else np.float64
    ctx = capi.Context(0)
    if args.check:
        ok = check(ctx, dtype, args.precision)
        if not ok:
            raise SystemExit(1)
    if args.mnk:
        M, N, K = [int(v) for v in args.mnk.split(",")]
    else:
        M = N = K = args.size
    rng = np.random.default_rng(5)
    A = rng.uniform(-1, 1, (M, K)).astype(dtype)
    B = rng.uniform(-1, 1, (K, N)).astype(dtype)
    results = {}
    for layout in args.layouts.split(","):
        a_store = A if layout[0] == "n" else np.ascontiguousarray(A.T)
        b_store = B if layout[1] == "n" else np.ascontiguousarray(B.T)
        a_rs, a_cs = (K, 1) if layout[0] == "n" else (1, M)
        b_rs, b_cs = (N, 1) if layout[1] == "n" else (1, K)
        pa, pb = ctx.upload(a_store), ctx.upload(b_store)
        pc = ctx.alloc(M * N * np.dtype(dtype).itemsize)

        def call():
            capi.check(ctx.lib.llo_cuda_gemm(ctx.handle, capi.dtype_code(dtype), M, N, K,
                                             pa, a_rs, a_cs, pb, b_rs, b_cs, pc, N, 1, args.precision))
        for _ in range(2):
            call()
        start, stop = ctx.event(), ctx.event()
        times = []
        for _ in range(args.iters):
            ctx.record(start)
            call()
            ctx.record(stop)
            times.append(ctx.elapsed_ms(start, stop))
        ms = float(np.median(times))
        tf = 2.0 * M * N * K / ms / 1e9
        # spot check 64 random outputs against float64
        got = ctx.download(pc, M * N, dtype).reshape(M, N)
        idx = rng.integers(0, M, 64), rng.integers(0, N, 64)
        exp = np.einsum("ik,ki->i", A[idx[0]].astype(np.float64), B[:, idx[1]].astype(np.float64))
        scale = np.einsum("ik,ki->i", np.abs(A[idx[0]]).astype(np.float64), np.abs(B[:, idx[1]]).astype(np.float64))
        err = float((np.abs(exp - got[idx]) / scale).max())
        print("%s %s M=%d N=%d K=%d: %.3f ms (min %.3f)  %.2f TFLOP/s  spot rel err %.2e" % (
            args.dtype, layout, M, N, K, ms, min(times), tf, err), flush=True)
        results[layout] = {"ms": ms, "tflops": tf, "err": err}
        for p in (pa, pb, pc):
            ctx.free(p)
    print(json.dumps(results))


if __name__ == "__main__":
    main()
