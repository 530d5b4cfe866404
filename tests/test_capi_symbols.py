"""CPU: the C-ABI library loads and exports every symbol include/llo_cuda.h
declares (no compute calls without a GPU)."""
import os
import re
import shutil
import subprocess

import pytest

import util
from util import capi

HEADER = os.path.join(util.ROOT, "include", "llo_cuda.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(llo_cuda_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = declared_symbols()
    assert len(names) >= 30
    lib = capi.lib()
    for name in names:
        assert hasattr(lib, name), "libllo_cuda.so does not export %s" % name


def test_binding_covers_the_header():
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_abi_version_and_error_channel():
    lib = capi.lib()
    assert lib.llo_cuda_abi_version() == 1
    assert isinstance(lib.llo_cuda_last_error(), bytes)


def test_struct_layouts_match_the_header():
    import ctypes as C
    assert C.sizeof(capi.Arg) == 8 + 8 * 8 + 81 * 8 + 4 + 4
    assert C.sizeof(capi.View) == 8 + 8 * 8 + 8
    assert C.sizeof(capi.Insn) == 2
    assert capi.OPCODES["ABS"] == 1 and capi.OPCODES["RAND_NORM"] == 23
    assert capi.DTYPES["DOUBLE"] == 1 and capi.DTYPES["UINT64"] == 10


def test_no_gpu_means_loud_failure_not_fallback():
    import numpy as np
    import pytest
    from util import llo
    if llo.available():
        pytest.skip("a GPU is present")
    v = llo.variable(np.ones((2, 3)), "v")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        llo.evaluate(v)


def test_nvrtc_stand_in_header_matches_the_real_one():
    """csrc/cuda/rtc/llo_cuda.h repeats the constants the device code uses (a JIT
    translation unit may not declare host functions): same values as include/llo_cuda.h"""
    stub = open(os.path.join(util.ROOT, "cortenn_b200", "csrc", "cuda", "rtc", "llo_cuda.h")).read()
    real = open(HEADER).read()
    for name in ("LLO_RANK", "LLO_MAT", "LLO_VM_MAX_INPUTS", "LLO_VM_MAX_INSNS", "LLO_VM_MAX_CONSTS",
                 "LLO_VM_MAX_DEPTH"):
        pat = r"#define\s+%s\s+(\d+)" % name
        assert re.search(pat, stub).group(1) == re.search(pat, real).group(1), name

    def opcodes(text):
        body = re.search(r"enum\s*\{\s*(LLO_OP_BAD.*?)\}", text, flags=re.S).group(1)
        return [t.strip().split("=")[0].strip() for t in body.replace("\n", " ").split(",") if t.strip()]
    assert opcodes(stub) == opcodes(real)


def test_header_is_plain_c_and_links(tmp_path):
    """the boundary is a C ABI: include/llo_cuda.h compiles as C99 (-pedantic) and a C
    program links against the library and reads its ABI version"""
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    root = util.ROOT
    src = tmp_path / "abi.c"
    src.write_text('#include "llo_cuda.h"\n'
                   'int main(void) { return llo_cuda_abi_version() == LLO_CUDA_ABI_VERSION ? 0 : 1; }\n')
    lib_dir = os.path.join(root, "cortenn_b200")
    exe = str(tmp_path / "abi")
    subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                    str(src), "-L", lib_dir, "-lllo_cuda", "-Wl,-rpath," + lib_dir, "-o", exe],
                   check=True, capture_output=True, timeout=120)
    assert subprocess.run([exe], timeout=60).returncode == 0
