"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints the
one JSON line the driver parses, with the keys of the bench contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*extra):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "1", "--warmup", "1", "--ref-log2n", "12", *extra],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def check_contract(line):
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "elementwise_chain_hbm_gbps" and line["unit"] == "GB/s"
    assert line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"],
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["value"] == line["value"] and base["cores"] >= 1
    assert base["single_thread"]["value"] > 0


def test_reference_arm_all_cores():
    line = run()
    check_contract(line)
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["config"]["processes"] == line["cpu_baseline"]["cores"]


def test_reference_arm_single_thread():
    line = run("--ref-single")
    check_contract(line)
    assert line["cpu_baseline"]["cores"] == 1
    assert line["value"] == line["cpu_baseline"]["single_thread"]["value"]
