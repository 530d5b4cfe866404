"""The reference's C++ API as a C++ library: tests/cpp/host_api.cpp is a C++ caller (no
Python) of `age::*`, `llo::derive`, `llo::optimize`, `llo::plan_preview`,
`llo::ShardedEvaluator` and `llo::eval`, linked against cortenn_b200/libllo.so +
libllo_cuda.so.  Without a GPU it must fail loudly at evaluation; on a B200 the gradients it
evaluates are checked against central differences of the loss."""
import os
import shutil
import subprocess

import pytest

import util

LIB_DIR = os.path.join(util.ROOT, "cortenn_b200")


def build(tmp_path):
    cxx = os.environ.get("CORTENN_CXX") or shutil.which("g++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    if not os.path.exists(os.path.join(LIB_DIR, "libllo.so")):
        pytest.fail("cortenn_b200/libllo.so is not built (python __graft_entry__.py)")
    exe = str(tmp_path / "host_api")
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(LIB_DIR, "csrc", "host"),
                    "-I", os.path.join(util.ROOT, "include"), os.path.join(util.ROOT, "tests", "cpp", "host_api.cpp"),
                    "-L", LIB_DIR, "-lllo", "-lllo_cuda", "-Wl,-rpath," + LIB_DIR, "-o", exe],
                   check=True, capture_output=True, timeout=600)
    return exe


def test_cpp_caller_builds_plans_and_fails_loudly_without_a_gpu(tmp_path):
    exe = build(tmp_path)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "plan preview" in run.stdout and "generic 0" in run.stdout
    assert "no CPU fallback" in run.stdout


@pytest.mark.gpu
def test_cpp_caller_gradients_match_central_differences(tmp_path):
    exe = build(tmp_path)
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "gradients match central differences" in run.stdout
