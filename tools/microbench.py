"""Kernel-level timings through the C-ABI (CUDA events on the context's
stream).  Development aid: prints achieved algorithmic GB/s per kernel."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cortenn_b200 import capi  # noqa: E402

F4 = capi.dtype_code(np.float32)


def timeit(ctx, fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    start, stop = ctx.event(), ctx.event()
    times = []
    for _ in range(iters):
        ctx.record(start)
        fn()
        ctx.record(stop)
        times.append(ctx.elapsed_ms(start, stop))
    return float(np.median(times)), float(np.min(times))


def view(ptr, strides, offset=0):
    v = capi.View()
    v.data = ptr
    v.stride = (C.c_int64 * 8)(*(list(strides) + [0] * (8 - len(strides))))
    v.offset = offset
    return v


def program(*insns):
    arr = (capi.Insn * len(insns))()
    for i, (op, arg) in enumerate(insns):
        arr[i].op = capi.OPCODES[op] if isinstance(op, str) else op
        arr[i].arg = arg
    return arr, len(insns)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    ctx = capi.Context(0)
    lib = ctx.lib
    side = 1 << (args.log2n // 2)
    n = side * side
    nbytes = n * 4
    x = ctx.alloc(nbytes)
    y = ctx.alloc(nbytes)
    out = ctx.alloc(nbytes)
    b = ctx.alloc(side * 4)
    # fill with something finite
    host = np.random.default_rng(0).uniform(0.5, 2.0, 1 << 20).astype(np.float32)
    hp = ctx.upload(host)
    for dst in (x, y):
        for off in range(0, nbytes, host.nbytes):
            capi.check(lib.llo_cuda_d2d(ctx.handle, dst + off, hp, min(host.nbytes, nbytes - off)))
    capi.check(lib.llo_cuda_d2d(ctx.handle, b, hp, side * 4))
    shape2 = capi.shape8([side, side])
    results = {}

    def report(name, bytes_moved, fn):
        med, best = timeit(ctx, fn, args.iters)
        results[name] = {"ms": med, "ms_min": best, "GBps": bytes_moved / med / 1e6}
        print("%-28s %8.3f ms (min %8.3f)  %8.1f GB/s" % (name, med, best, bytes_moved / med / 1e6), flush=True)

    report("d2d copy (cudaMemcpy)", 2 * nbytes, lambda: capi.check(lib.llo_cuda_d2d(ctx.handle, out, x, nbytes)))

    def elementwise(views, prog):
        arr = (capi.View * len(views))(*views)
        p, np_ = prog
        return lambda: capi.check(lib.llo_cuda_elementwise(ctx.handle, F4, out, shape2, arr, len(views), None, 0, p, np_))

    dense = [1, side]
    report("copy (vm)", 2 * nbytes, elementwise([view(x, dense)], program((capi.VM_LOAD, 0))))
    report("exp", 2 * nbytes, elementwise([view(x, dense)], program((capi.VM_LOAD, 0), ("EXP", 0))))
    report("sin", 2 * nbytes, elementwise([view(x, dense)], program((capi.VM_LOAD, 0), ("SIN", 0))))
    report("add", 3 * nbytes, elementwise([view(x, dense), view(y, dense)],
                                          program((capi.VM_LOAD, 0), (capi.VM_LOAD, 1), ("SUM", 0))))
    report("pow", 3 * nbytes, elementwise([view(x, dense), view(y, dense)],
                                          program((capi.VM_LOAD, 0), (capi.VM_LOAD, 1), ("POW", 0))))
    report("extend b (broadcast dim1)", nbytes + side * 4, elementwise([view(b, [1, 0])], program((capi.VM_LOAD, 0))))
    report("mul x, extend(b)", 2 * nbytes + side * 4, elementwise(
        [view(x, dense), view(b, [1, 0])], program((capi.VM_LOAD, 0), (capi.VM_LOAD, 1), ("PROD", 0))))
    report("permute {1,0}", 2 * nbytes, elementwise([view(x, [side, 1])], program((capi.VM_LOAD, 0))))
    report("chain exp*b+permute (fused)", 2 * nbytes + side * 4, elementwise(
        [view(x, dense), view(b, [1, 0]), view(x, [side, 1])],
        program((capi.VM_LOAD, 0), ("EXP", 0), (capi.VM_LOAD, 1), ("PROD", 0), (capi.VM_LOAD, 2), ("SUM", 0))))

    small = ctx.alloc(side * 4 * 4)
    for name, op in (("SUM", "SUM"), ("MAX", "MAX")):
        report("reduce_%s(x,1) columns" % name.lower(), nbytes + side * 4, lambda op=op: capi.check(
            lib.llo_cuda_reduce(ctx.handle, capi.OPCODES[op], F4, small, x, shape2, 1, 0)))
        report("reduce_%s(x,0) scalar" % name.lower(), nbytes + 4, lambda op=op: capi.check(
            lib.llo_cuda_reduce(ctx.handle, capi.OPCODES[op], F4, small, x, shape2, 0, 0)))
    # rows: reduce(permute(x)) = push with the transposed layout
    m = np.eye(9)
    m[0][0] = 1.0 / side  # fold dim 0 (contiguous), keep dim 1
    arg = capi.make_arg(x, [side, side], m, True)
    oshape = capi.shape8([1, side])
    report("reduce_sum rows (contiguous)", nbytes + side * 4, lambda: capi.check(
        lib.llo_cuda_nnary(ctx.handle, capi.OPCODES["SUM"], F4, small, oshape, C.byref(arg), 1)))
    print(json.dumps(results))
    ctx.sync()


if __name__ == "__main__":
    main()
