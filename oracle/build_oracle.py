"""Build the ORACLE (test infrastructure): CPU restatement of the reference's
llo evaluator.  Outputs go to oracle/_build/ (git-ignored, travels to the GPU
box):  liboracle.so (C entry points, host pointers)  and  _oracle.*.so (pybind).

/root/reference cannot be compiled here (it needs bazel, the un-vendored
Tenncor@959d51e sources and gtest), so there is no oracle/_ref: the oracle is a
port, pinned by the reference's golden vectors (tests/test_golden.py, tests/test_pbm.py, tests/test_shear_cpp.py).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
HOST = os.path.join(ROOT, "cortenn_b200", "csrc", "host")
INCLUDE = os.path.join(ROOT, "include")

# the image exports CXX=/opt/gcc/bin/g++, a wrapper whose link step omits
# libstdc++; use the system compiler unless told otherwise
CXX = os.environ.get("CORTENN_CXX", "/usr/bin/g++")
FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall",
         "-I", HERE, "-I", HOST, "-I", INCLUDE]


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError("oracle build failed")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def capi_path():
    return os.path.join(OUT, "liboracle.so")


def module_path():
    return os.path.join(OUT, "_oracle" + sysconfig.get_config_var("EXT_SUFFIX"))


def build():
    import pybind11
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE)
            if f.endswith((".cpp", ".hpp"))]
    for base, _, files in os.walk(HOST):
        deps += [os.path.join(base, f) for f in files if f.endswith(".hpp")]
    core = os.path.join(HERE, "cpu_llo.cpp")
    if _stale(capi_path(), deps):
        _run([CXX] + FLAGS + ["-shared", "-o", capi_path(), core,
                              os.path.join(HERE, "oracle_capi.cpp")])
    if _stale(module_path(), deps):
        _run([CXX] + FLAGS + ["-fvisibility=hidden",
                              "-I", pybind11.get_include(),
                              "-I", sysconfig.get_paths()["include"],
                              "-shared", "-o", module_path(), core,
                              os.path.join(HERE, "pyoracle.cpp")])


if __name__ == "__main__":
    build()
    print("built", capi_path(), module_path())
