"""Run one named kernel case a few times (for ncu captures).
usage: python tools/prof_cases.py --case exp|add|extend|mulext|permute|chain|rsum_cols|rsum_rows|rsum_all [--log2n 28] [--iters 3]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cortenn_b200 import capi  # noqa: E402
from microbench import view, program  # noqa: E402

F4 = capi.dtype_code(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    ctx = capi.Context(0)
    lib = ctx.lib
    side = 1 << (args.log2n // 2)
    n = side * side
    nbytes = n * 4
    x, y, out, b = ctx.alloc(nbytes), ctx.alloc(nbytes), ctx.alloc(nbytes), ctx.alloc(side * 4)
    host = np.random.default_rng(0).uniform(0.5, 2.0, 1 << 20).astype(np.float32)
    hp = ctx.upload(host)
    for dst in (x, y):
        for off in range(0, nbytes, host.nbytes):
            capi.check(lib.llo_cuda_d2d(ctx.handle, dst + off, hp, min(host.nbytes, nbytes - off)))
    capi.check(lib.llo_cuda_d2d(ctx.handle, b, hp, side * 4))
    shape2 = capi.shape8([side, side])
    dense = [1, side]
    cases = {
        "copy": ([view(x, dense)], program((capi.VM_LOAD, 0))),
        "exp": ([view(x, dense)], program((capi.VM_LOAD, 0), ("EXP", 0))),
        "add": ([view(x, dense), view(y, dense)], program((capi.VM_LOAD, 0), (capi.VM_LOAD, 1), ("SUM", 0))),
        "extend": ([view(b, [1, 0])], program((capi.VM_LOAD, 0))),
        "mulext": ([view(x, dense), view(b, [1, 0])], program((capi.VM_LOAD, 0), (capi.VM_LOAD, 1), ("PROD", 0))),
        "permute": ([view(x, [side, 1])], program((capi.VM_LOAD, 0))),
        "pairadd": ([view(x, dense), view(x, [side, 1])], program((capi.VM_LOAD, 0), (capi.VM_LOAD, 1), ("SUM", 0))),
        "pairexp": ([view(x, dense), view(x, [side, 1])],
                    program((capi.VM_LOAD, 0), ("EXP", 0), (capi.VM_LOAD, 1), ("SUM", 0))),
        "pairmulb": ([view(x, dense), view(b, [1, 0]), view(x, [side, 1])],
                     program((capi.VM_LOAD, 0), (capi.VM_LOAD, 1), ("PROD", 0), (capi.VM_LOAD, 2), ("SUM", 0))),
        "chain": ([view(x, dense), view(b, [1, 0]), view(x, [side, 1])],
                  program((capi.VM_LOAD, 0), ("EXP", 0), (capi.VM_LOAD, 1), ("PROD", 0), (capi.VM_LOAD, 2), ("SUM", 0))),
    }
    small = ctx.alloc(side * 16)
    # the chain over the reference-representable shape [128,128,128,128] (needs log2n = 28)
    shape4 = capi.shape8([128, 128, 128, 128])
    dense4 = [1, 128, 128 * 128, 128 * 128 * 128]
    cases4 = {
        "chain4": ([view(x, dense4), view(b, [1, 0, 0, 0]), view(x, [128, 1, 128 * 128, 128 * 128 * 128])],
                   program((capi.VM_LOAD, 0), ("EXP", 0), (capi.VM_LOAD, 1), ("PROD", 0), (capi.VM_LOAD, 2), ("SUM", 0))),
    }
    if args.case in cases or args.case in cases4:
        views, (prog, nprog) = (cases if args.case in cases else cases4)[args.case]
        arr = (capi.View * len(views))(*views)
        oshape = shape2 if args.case in cases else shape4

        def run():
            capi.check(lib.llo_cuda_elementwise(ctx.handle, F4, out, oshape, arr, len(views), None, 0, prog, nprog))
    elif args.case == "rsum_cols":
        def run():
            capi.check(lib.llo_cuda_reduce(ctx.handle, capi.OPCODES["SUM"], F4, small, x, shape2, 1, 0))
    elif args.case == "rsum_all":
        def run():
            capi.check(lib.llo_cuda_reduce(ctx.handle, capi.OPCODES["SUM"], F4, small, x, shape2, 0, 0))
    elif args.case == "rsum_rows":
        m = np.eye(9)
        m[0][0] = 1.0 / side
        arg = capi.make_arg(x, [side, side], m, True)
        oshape = capi.shape8([1, side])

        def run():
            capi.check(lib.llo_cuda_nnary(ctx.handle, capi.OPCODES["SUM"], F4, small, oshape, C.byref(arg), 1))
    elif args.case == "convert":
        # GenericData::copyover: f32 -> f64 of 2^26 elements
        n26 = min(n, 1 << 26)
        wide = ctx.alloc(n26 * 8)

        def run():
            capi.check(lib.llo_cuda_convert(ctx.handle, wide, capi.dtype_code(np.float64), x, F4, n26))
    elif args.case == "rand":
        # rand_unif over 2^26 fp32 outputs, scalar bounds broadcast by an extend map
        n26 = min(n, 1 << 26)
        ext = np.zeros((9, 9))
        ext[8][8] = 1.0
        lo_h, hi_h = ctx.upload(np.array([2.0], dtype=np.float32)), ctx.upload(np.array([9.0], dtype=np.float32))
        a_lo, a_hi = capi.make_arg(lo_h, [], ext, False), capi.make_arg(hi_h, [], ext, False)
        oshape = capi.shape8([8192, n26 // 8192])

        def run():
            capi.check(lib.llo_cuda_rand(ctx.handle, capi.OPCODES["RAND_UNIF"], F4, out, oshape, C.byref(a_lo), C.byref(a_hi)))
    elif args.case == "general":
        # a fractional coordinate map (nearest-neighbour upsample by 3 along dim 1): the
        # per-element double-precision path of general.cu, 2^24 outputs
        m = np.eye(9)
        m[1][1] = 1.0 / 3
        rows = 1 << 10
        arg = capi.make_arg(x, [4096, rows], m, False)
        oshape = capi.shape8([4096, rows * 3])

        def run():
            capi.check(lib.llo_cuda_nnary(ctx.handle, capi.OPCODES["SUM"], F4, out, oshape, C.byref(arg), 1))
    else:
        raise SystemExit("unknown case " + args.case)
    start, stop = ctx.event(), ctx.event()
    for _ in range(args.iters):
        ctx.record(start)
        run()
        ctx.record(stop)
        print("%s: %.3f ms" % (args.case, ctx.elapsed_ms(start, stop)))
    ctx.sync()


if __name__ == "__main__":
    main()
