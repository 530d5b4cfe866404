"""In-tree build of the native parts.

  libllo_cuda.so   cortenn_b200/csrc/cuda/*.cu      nvcc, sm_100a only
  _C.*.so          host C++ (ade/age/llo/opt) + pybind11 bindings, links the above
  oracle (separate, test infrastructure): see oracle/build_oracle.py

Objects go to build/ (git-ignored); the shared objects land next to this file
so they travel to the GPU box with the source snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(ROOT, "build")
INCLUDE = os.path.join(ROOT, "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# the image exports CXX=/opt/gcc/bin/g++, a wrapper whose link step omits
# libstdc++; use the system compiler unless told otherwise
CXX = os.environ.get("CORTENN_CXX", "/usr/bin/g++")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",              # no multiply-add contraction: CPU parity
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE, "-I", os.path.join(CSRC, "cuda"),
]
# kernels whose FMAs are written out explicitly keep contraction on
NVCC_FMAD_OK = {"gemm.cu"}  # (sgemm.cu / dgemm.cu spell their FMAs as fmaf / fma)

CXX_FLAGS = [
    "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden",
    "-ffp-contract=off", "-Wall", "-Wno-unused-function",
]


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError("build step failed: " + cmd[0])
    return proc.stdout + proc.stderr


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _headers(*dirs):
    out = []
    for d in dirs:
        for base, _, files in os.walk(d):
            out += [os.path.join(base, f) for f in files
                    if f.endswith((".h", ".hpp", ".cuh"))]
    return out


def pybind_includes():
    import pybind11
    return [pybind11.get_include(), sysconfig.get_paths()["include"]]


def ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def cuda_lib_path():
    return os.path.join(HERE, "libllo_cuda.so")


def host_lib_path():
    return os.path.join(HERE, "_C" + ext_suffix())


def build_cuda(verbose=False, jobs=None):
    """nvcc -> libllo_cuda.so (static cudart, no torch, no python)."""
    src_dir = os.path.join(CSRC, "cuda")
    objdir = os.path.join(BUILD, "cuda")
    os.makedirs(objdir, exist_ok=True)
    sources = sorted(f for f in os.listdir(src_dir) if f.endswith(".cu"))
    headers = _headers(src_dir, INCLUDE)
    if not verbose and not _stale(cuda_lib_path(), [os.path.join(src_dir, f) for f in sources] + headers):
        return ""  # prebuilt library is current (e.g. on the GPU box)
    todo = []
    objs = []
    for name in sources:
        src = os.path.join(src_dir, name)
        obj = os.path.join(objdir, name[:-3] + ".o")
        objs.append(obj)
        if _stale(obj, [src] + headers):
            flags = list(NVCC_FLAGS)
            if name in NVCC_FMAD_OK:
                flags.remove("-fmad=false")
            if verbose:
                flags += ["-Xptxas", "-v"]
            todo.append([NVCC] + flags + ["-c", src, "-o", obj])
    logs = []
    if todo:
        with ThreadPoolExecutor(max_workers=jobs or os.cpu_count()) as pool:
            logs = list(pool.map(_run, todo))
    lib = cuda_lib_path()
    if todo or _stale(lib, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
              "-o", lib] + objs + ["-cudart", "static"])
    return "\n".join(logs)


def build_host():
    """g++ -> _C extension (ade/age/llo host layer + pybind11), links libllo_cuda.so."""
    objdir = os.path.join(BUILD, "host")
    os.makedirs(objdir, exist_ok=True)
    host = os.path.join(CSRC, "host")
    sources = [os.path.join(host, "llo", "src", f)
               for f in sorted(os.listdir(os.path.join(host, "llo", "src")))
               if f.endswith(".cpp")]
    sources.append(os.path.join(CSRC, "pybind", "module.cpp"))
    headers = _headers(host, INCLUDE)
    if not _stale(host_lib_path(), sources + headers + [cuda_lib_path()]):
        return
    incs = ["-I", host, "-I", INCLUDE]
    for inc in pybind_includes():
        incs += ["-I", inc]
    todo = []
    objs = []
    for src in sources:
        obj = os.path.join(objdir, os.path.basename(src)[:-4] + ".o")
        objs.append(obj)
        if _stale(obj, [src] + headers):
            todo.append([CXX] + CXX_FLAGS + incs + ["-c", src, "-o", obj])
    if todo:
        with ThreadPoolExecutor(max_workers=os.cpu_count()) as pool:
            list(pool.map(_run, todo))
    lib = host_lib_path()
    if todo or _stale(lib, objs + [cuda_lib_path()]):
        _run([CXX, "-shared", "-o", lib] + objs +
             ["-L", HERE, "-lllo_cuda", "-lstdc++", "-lm",
              "-Wl,-rpath,$ORIGIN"])


def build_all(verbose=False):
    log = build_cuda(verbose=verbose)
    build_host()
    return log


if __name__ == "__main__":
    out = build_all(verbose="-v" in sys.argv)
    if out.strip():
        print(out)
    print("built", cuda_lib_path(), host_lib_path())
