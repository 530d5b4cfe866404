"""ctypes binding of the C-ABI (include/llo_cuda.h).  Used by the parity tests
and bench.py to call the CUDA library exactly as a foreign host would."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LLO_CUDA_LIB: same-box A/B runs of a variant build (tools/), never set in production
LIB_PATH = os.environ.get("LLO_CUDA_LIB") or os.path.join(HERE, "libllo_cuda.so")

RANK = 8
MAT = 81

DTYPES = {
    "DOUBLE": 1, "FLOAT": 2, "INT8": 3, "UINT8": 4, "INT16": 5, "UINT16": 6,
    "INT32": 7, "UINT32": 8, "INT64": 9, "UINT64": 10,
}
NP_DTYPES = {
    1: np.float64, 2: np.float32, 3: np.int8, 4: np.uint8, 5: np.int16,
    6: np.uint16, 7: np.int32, 8: np.uint32, 9: np.int64, 10: np.uint64,
}
OPCODES = {name: i + 1 for i, name in enumerate(
    "ABS NEG SIN COS TAN EXP LOG SQRT ROUND POW SUM SUB PROD DIV MIN MAX "
    "EQ NEQ LT GT RAND_BINO RAND_UNIF RAND_NORM".split())}
VM_LOAD = 64
VM_CONST = 65

OK, ERR_FATAL, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3


def dtype_code(np_dtype):
    np_dtype = np.dtype(np_dtype)
    for code, t in NP_DTYPES.items():
        if np.dtype(t) == np_dtype:
            return code
    raise ValueError("unsupported dtype %s" % np_dtype)


class Arg(C.Structure):
    """llo_cuda_arg"""
    _fields_ = [("data", C.c_void_p), ("shape", C.c_uint64 * RANK),
                ("coorder", C.c_double * MAT), ("push", C.c_int32),
                ("reserved", C.c_int32)]


class View(C.Structure):
    """llo_cuda_view"""
    _fields_ = [("data", C.c_void_p), ("stride", C.c_int64 * RANK),
                ("offset", C.c_int64)]


class Insn(C.Structure):
    """llo_cuda_insn"""
    _fields_ = [("op", C.c_uint8), ("arg", C.c_uint8)]


EXPORTS = {
    "llo_cuda_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "llo_cuda_destroy": (C.c_int, [C.c_void_p]),
    "llo_cuda_last_error": (C.c_char_p, []),
    "llo_cuda_abi_version": (C.c_int, []),
    "llo_cuda_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "llo_cuda_get_stream": (C.c_void_p, [C.c_void_p]),
    "llo_cuda_device": (C.c_int, [C.c_void_p]),
    "llo_cuda_sm_count": (C.c_int, [C.c_void_p]),
    "llo_cuda_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "llo_cuda_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "llo_cuda_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "llo_cuda_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "llo_cuda_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "llo_cuda_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "llo_cuda_h2d_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "llo_cuda_d2h_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "llo_cuda_event_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "llo_cuda_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "llo_cuda_memset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    "llo_cuda_wait_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "llo_cuda_sync": (C.c_int, [C.c_void_p]),
    "llo_cuda_graph_begin": (C.c_int, [C.c_void_p]),
    "llo_cuda_graph_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "llo_cuda_graph_abort": (C.c_int, [C.c_void_p]),
    "llo_cuda_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "llo_cuda_graph_destroy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "llo_cuda_launch_count": (C.c_uint64, [C.c_void_p]),
    "llo_cuda_event_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "llo_cuda_event_record": (C.c_int, [C.c_void_p, C.c_void_p]),
    "llo_cuda_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "llo_cuda_event_destroy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "llo_cuda_set_specialize_min": (C.c_int, [C.c_void_p, C.c_uint64]),
    "llo_cuda_specialize_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "llo_cuda_convert": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint64]),
    "llo_cuda_unary": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(Arg)]),
    "llo_cuda_binary": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(Arg), C.POINTER(Arg)]),
    "llo_cuda_nnary": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(Arg), C.c_int]),
    "llo_cuda_op_exec": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(Arg), C.c_int]),
    "llo_cuda_seed": (C.c_int, [C.c_void_p, C.c_uint64]),
    "llo_cuda_rand": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(Arg), C.POINTER(Arg)]),
    "llo_cuda_reduce": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int]),
    "llo_cuda_set_reduce_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "llo_cuda_reduce_view": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(View),
                                       C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "llo_cuda_elementwise": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(View), C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(Insn), C.c_int]),
    "llo_cuda_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                C.c_void_p, C.c_int64, C.c_int64, C.c_int]),
    "llo_cuda_comm_unique_id": (C.c_int, [C.c_void_p]),
    "llo_cuda_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "llo_cuda_comm_destroy": (C.c_int, [C.c_void_p]),
    "llo_cuda_comm_rank": (C.c_int, [C.c_void_p]),
    "llo_cuda_comm_size": (C.c_int, [C.c_void_p]),
    "llo_cuda_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int]),
    "llo_cuda_allreduce_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]),
    "llo_cuda_comm_join": (C.c_int, [C.c_void_p]),
}

_lib = None


def lib():
    """Load libllo_cuda.so and declare every exported symbol.  Fails loudly
    when the library was not built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError("%s not built; run `python cortenn_b200/build.py`" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)  # AttributeError if not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class LloError(RuntimeError):
    pass


class UnsupportedTypedOp(LloError):
    """std::bad_function_call of the reference"""


def check(status):
    if status == OK:
        return
    msg = lib().llo_cuda_last_error().decode()
    if status == ERR_UNSUPPORTED:
        raise UnsupportedTypedOp(msg)
    raise LloError(msg)


def shape8(shape):
    """ade shape (dim 0 fastest) padded to rank 8"""
    dims = list(shape) + [1] * (RANK - len(shape))
    return (C.c_uint64 * RANK)(*dims)


def identity_matrix():
    return np.eye(9, dtype=np.float64)


def make_arg(ptr, shape, coorder=None, push=False):
    arg = Arg()
    arg.data = ptr
    arg.shape = shape8(shape)
    mat = identity_matrix() if coorder is None else np.asarray(coorder, dtype=np.float64)
    arg.coorder = (C.c_double * MAT)(*mat.reshape(-1))
    arg.push = 1 if push else 0
    return arg


class Context:
    """One llo_cuda_ctx plus small numpy <-> device helpers."""

    def __init__(self, device=0, handle=None):
        """handle: adopt an existing llo_cuda_ctx* (e.g. the evaluator's,
        from cortenn_b200.llo.context()) instead of creating one"""
        self.lib = lib()
        self.owned = handle is None
        if handle is None:
            handle = C.c_void_p()
            check(self.lib.llo_cuda_init(device, C.byref(handle)))
        else:
            handle = C.c_void_p(handle)
        self.handle = handle

    def close(self):
        if self.handle and self.owned:
            self.lib.llo_cuda_destroy(self.handle)
        self.handle = None

    def alloc(self, nbytes):
        ptr = C.c_void_p()
        check(self.lib.llo_cuda_alloc(self.handle, nbytes, C.byref(ptr)))
        return ptr.value

    def free(self, ptr):
        check(self.lib.llo_cuda_free(self.handle, ptr))

    def upload(self, array):
        array = np.ascontiguousarray(array)
        ptr = self.alloc(max(array.nbytes, 1))
        check(self.lib.llo_cuda_h2d(self.handle, ptr, array.ctypes.data, array.nbytes))
        return ptr

    def download(self, ptr, count, np_dtype):
        out = np.empty(count, dtype=np_dtype)
        check(self.lib.llo_cuda_d2h(self.handle, out.ctypes.data, ptr, out.nbytes))
        return out

    def sync(self):
        check(self.lib.llo_cuda_sync(self.handle))

    def launch_count(self):
        return self.lib.llo_cuda_launch_count(self.handle)

    def event(self):
        ev = C.c_void_p()
        check(self.lib.llo_cuda_event_create(self.handle, C.byref(ev)))
        return ev

    def record(self, ev):
        check(self.lib.llo_cuda_event_record(self.handle, ev))

    def elapsed_ms(self, start, stop):
        ms = C.c_float()
        check(self.lib.llo_cuda_event_elapsed_ms(self.handle, start, stop, C.byref(ms)))
        return ms.value

    def op_exec(self, opcode, np_dtype, out_ptr, outshape, args):
        arr = (Arg * len(args))(*args)
        check(self.lib.llo_cuda_op_exec(self.handle, OPCODES[opcode], dtype_code(np_dtype),
                                        out_ptr, shape8(outshape), arr, len(args)))
