"""CPU: opt::TargetPruner against the golden tree of the reference's own test
(opt/test/test_shear.cpp:100-178, SHEAR.Prune), through a small C++ program compiled here
(the pruner is host C++ with no Python surface)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_target_pruner_reproduces_the_reference_golden_tree(tmp_path):
    cxx = os.environ.get("CORTENN_CXX") or shutil.which("g++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "shear_prune")
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "cortenn_b200", "csrc", "host"),
                    os.path.join(ROOT, "tests", "cpp", "shear_prune.cpp"), "-o", exe],
                   check=True, capture_output=True, timeout=300)
    run = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0, "pruned tree differs from opt/test/test_shear.cpp:160-176:\n" + run.stdout
    assert run.stdout.count("killable") == 6 and "leaf2=" in run.stdout
