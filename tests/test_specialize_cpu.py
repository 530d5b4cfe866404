"""CPU: the device sources of the elementwise engine compile with NVRTC for sm_100a as
specialised (LLO_VM_STATIC) translation units -- the build libllo_cuda.so makes at run
time for large launches (csrc/cuda/specialize.cu).  NVRTC needs no GPU."""
import ctypes as C
import os

import pytest

import util

CUDA = os.path.join(util.ROOT, "cortenn_b200", "csrc", "cuda")
SOURCES = [("vm_kernels.cuh", "vm_kernels.cuh"), ("vm.cuh", "vm.cuh"), ("ops.cuh", "ops.cuh"),
           ("llo_cuda.h", "rtc/llo_cuda.h"), ("cstdint", "rtc/cstdint"), ("type_traits", "rtc/type_traits"),
           ("stdint.h", "rtc/stdint.h"), ("stddef.h", "rtc/stddef.h")]


def nvrtc():
    for name in ("libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    pytest.skip("libnvrtc not present")


def tables(sel, arg, mode, s0, s1, vec, ndims):
    def row(name, vals):
        return "static constexpr int %s[] = {%s};\n" % (name, ", ".join(map(str, vals)))
    return ("#define LLO_VM_STATIC 1\n#define LLO_VM_S_IDX32 1\n#define LLO_VM_S_ROW_UNIFORM 0\n#define LLO_VM_S_ROWS_ALIGNED 1\n#define LLO_VM_S_OUT_VEC 1\n#define LLO_VM_PAIR_NBUF 2\n#define LLO_VM_PAIR_MINCTAS 3\n#define LLO_VM_STATIC_NDIMS %d\n#define LLO_VM_S_NINSNS %d\n#define LLO_VM_S_NINPUTS %d\n" % (ndims, len(sel), len(mode)) +
            row("LLO_VM_S_SEL", sel) + row("LLO_VM_S_ARG", arg) + row("LLO_VM_S_MODE", mode) +
            row("LLO_VM_S_S0", s0) + row("LLO_VM_S_S1", s1) + row("LLO_VM_S_VEC", vec))


def compile_tu(lib, prefix, expr):
    src = (prefix + '#include "vm_kernels.cuh"\n').encode()
    texts = [open(os.path.join(CUDA, rel), "rb").read() for _, rel in SOURCES]
    hs = (C.c_char_p * len(texts))(*texts)
    ns = (C.c_char_p * len(texts))(*[n.encode() for n, _ in SOURCES])
    prog = C.c_void_p()
    assert lib.nvrtcCreateProgram(C.byref(prog), src, b"spec.cu", len(texts), hs, ns) == 0
    lib.nvrtcAddNameExpression(prog, expr.encode())
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"--fmad=false"]
    rc = lib.nvrtcCompileProgram(prog, len(opts), (C.c_char_p * len(opts))(*opts))
    n = C.c_size_t()
    lib.nvrtcGetProgramLogSize(prog, C.byref(n))
    log = C.create_string_buffer(n.value)
    lib.nvrtcGetProgramLog(prog, log)
    assert rc == 0, log.value.decode()[:2000]
    size = C.c_size_t()
    assert lib.nvrtcGetCUBINSize(prog, C.byref(size)) == 0 and size.value > 0
    return size.value


# accumulator-form selectors (vm.cuh VmSel) and operand sources (arg = source << 4 | index)
SET, PUSH, EXP, LOG, SUM, PROD, POW, RPOW = 0, 1, 8, 9, 11, 12, 24, 25
IN, CONST, POP = 0x10, 0x20, 0x30
LINEAR, STRIDED, TILE, ALIAS = 0, 1, 3, 4


@pytest.mark.parametrize("expr,prefix", [
    # the chain of BASELINE config 2 in the pair kernel: exp(x) * extend(b) + permute(x)
    ("llo_cuda::vm_pair4_kernel<float>",
     tables([SET, EXP, PROD, SUM], [IN | 0, 0, IN | 1, IN | 2], [ALIAS, STRIDED, TILE], [1, 1, 2], [2, 0, 1], [1, 1, 0], 2)),
    ("llo_cuda::vm_flat_kernel<float>", tables([SET, EXP], [IN | 0, 0], [LINEAR], [1], [2], [1], 1)),
    # a program with a spill (push / pop) and a constant operand, 8-byte type
    ("llo_cuda::vm_flat_kernel<double>",
     tables([SET, SUM, PUSH, SET, LOG, PROD], [IN | 0, CONST | 0, 0, IN | 1, 0, POP], [LINEAR, STRIDED], [1, 0], [2, 1], [1, 0], 2)),
    ("llo_cuda::vm_tile_kernel<int16_t>", tables([SET, SUM], [IN | 0, IN | 1], [TILE, STRIDED], [2, 1], [1, 2], [0, 1], 2)),
    ("llo_cuda::vm_tile4_kernel<float,true>", tables([SET], [IN | 0], [TILE], [2], [1], [0], 2)),
    ("llo_cuda::vm_tile4_kernel<int32_t,false>", tables([SET, SUM], [IN | 0, IN | 1], [TILE, STRIDED], [2, 1], [1, 2], [0, 0], 3)),
    ("llo_cuda::vm_flat_kernel<uint8_t>", tables([SET, PROD], [IN | 0, IN | 1], [LINEAR, LINEAR], [1, 1], [2, 2], [1, 1], 0)),
])
def test_specialised_translation_units_compile(expr, prefix):
    assert compile_tu(nvrtc(), prefix, expr) > 1000
