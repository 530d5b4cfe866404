"""Shared helpers of the test-suite: access to the ORACLE (CPU restatement of
the reference, test infrastructure only) and shape conventions."""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import build_oracle  # noqa: E402
import cortenn_b200  # noqa: E402,F401  (registers ade::iTensor with pybind11 first)
from cortenn_b200 import age, capi, llo  # noqa: E402

_spec = importlib.util.spec_from_file_location("_oracle", build_oracle.module_path())
oracle = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(oracle)

_olib = C.CDLL(build_oracle.capi_path())
_olib.oracle_last_error.restype = C.c_char_p
_olib.oracle_convert.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint64]
_olib.oracle_unary.argtypes = [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(capi.Arg)]
_olib.oracle_op_exec.argtypes = [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(capi.Arg), C.c_int]
_olib.oracle_seed.argtypes = [C.c_uint64]


class OracleError(RuntimeError):
    pass


class OracleUnsupported(OracleError):
    pass


def _ocheck(status):
    if status == 0:
        return
    msg = _olib.oracle_last_error().decode()
    if status == capi.ERR_UNSUPPORTED:
        raise OracleUnsupported(msg)
    raise OracleError(msg)


def nelems(shape):
    n = 1
    for d in shape:
        n *= int(d)
    return n


def host_arg(array, shape, coorder=None, push=False):
    """llo_cuda_arg over a numpy buffer (kept alive by the caller)"""
    return capi.make_arg(array.ctypes.data, shape, coorder, push)


def oracle_op_exec(opcode, np_dtype, outshape, args):
    """args: list of (array, shape, coorder|None, push) -> flat result array.
    The probability argument of RAND_BINO must be float64."""
    out = np.zeros(nelems(outshape), dtype=np_dtype)
    keep = [np.ascontiguousarray(a[0]) for a in args]
    cargs = (capi.Arg * len(args))(*[
        host_arg(k, a[1], a[2] if len(a) > 2 else None, a[3] if len(a) > 3 else False)
        for k, a in zip(keep, args)])
    _ocheck(_olib.oracle_op_exec(capi.OPCODES[opcode], capi.dtype_code(np_dtype),
                                 out.ctypes.data, capi.shape8(outshape), cargs, len(args)))
    return out


def oracle_unary(opcode, np_dtype, outshape, arg):
    out = np.zeros(nelems(outshape), dtype=np_dtype)
    keep = np.ascontiguousarray(arg[0])
    carg = host_arg(keep, arg[1], arg[2] if len(arg) > 2 else None,
                    arg[3] if len(arg) > 3 else False)
    _ocheck(_olib.oracle_unary(capi.OPCODES[opcode], capi.dtype_code(np_dtype),
                               out.ctypes.data, capi.shape8(outshape), C.byref(carg)))
    return out


def oracle_convert(array, out_dtype):
    array = np.ascontiguousarray(array)
    out = np.zeros(array.size, dtype=out_dtype)
    _ocheck(_olib.oracle_convert(out.ctypes.data, capi.dtype_code(out_dtype),
                                 array.ctypes.data, capi.dtype_code(array.dtype), array.size))
    return out


GUARD = 256          # bytes of canary on each side of every output buffer
GUARD_BYTE = 0xA5


def gpu_op_exec(ctx, opcode, np_dtype, outshape, args, entry="op_exec"):
    """Same contract as oracle_op_exec, through the CUDA C-ABI.  compute-sanitizer is
    closed on this pool, so every call checks for out-of-bounds WRITES itself: the output
    sits between two canary bands that must come back untouched."""
    ptrs = [ctx.upload(np.ascontiguousarray(a[0])) for a in args]
    nout = nelems(outshape)
    nbytes = nout * np.dtype(np_dtype).itemsize
    padded = (max(1, nbytes) + 15) // 16 * 16
    block = ctx.alloc(padded + 2 * GUARD)
    out_ptr = block + GUARD
    capi.check(ctx.lib.llo_cuda_memset(ctx.handle, block, GUARD_BYTE, padded + 2 * GUARD))
    capi.check(ctx.lib.llo_cuda_memset(ctx.handle, out_ptr, 0, nbytes))
    cargs = (capi.Arg * len(args))(*[
        capi.make_arg(p, a[1], a[2] if len(a) > 2 else None, a[3] if len(a) > 3 else False)
        for p, a in zip(ptrs, args)])
    try:
        if entry == "unary":
            capi.check(ctx.lib.llo_cuda_unary(ctx.handle, capi.OPCODES[opcode], capi.dtype_code(np_dtype),
                                              out_ptr, capi.shape8(outshape), cargs))
        else:
            capi.check(ctx.lib.llo_cuda_op_exec(ctx.handle, capi.OPCODES[opcode], capi.dtype_code(np_dtype),
                                                out_ptr, capi.shape8(outshape), cargs, len(args)))
        whole = ctx.download(block, padded + 2 * GUARD, np.uint8)
        assert (whole[:GUARD] == GUARD_BYTE).all(), "%s wrote below its output buffer" % opcode
        assert (whole[GUARD + nbytes:] == GUARD_BYTE).all(), "%s wrote past its output buffer" % opcode
        return whole[GUARD:GUARD + nbytes].view(np_dtype).copy()
    finally:
        for p in ptrs:
            ctx.free(p)
        ctx.free(block)


def var(data, ade_shape, dtype=np.float64, label=""):
    """llo variable from flat data in ade order (dim 0 fastest)"""
    arr = np.asarray(data, dtype=dtype).reshape(list(reversed(ade_shape)))
    return llo.variable(arr, label)


def flat(result):
    return np.asarray(result).reshape(-1)


def ade_shape_of(result, rank=None):
    shape = list(reversed(np.asarray(result).shape))
    return shape
