#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 LLO evaluator (driver contract).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Headline workload (BASELINE.json configs[1]): the elementwise chain
    out = add(mul(exp(x), extend(b, 1, {D})), permute(x, {1,0}))
over N = 2^28 fp32 elements, x as [D, D] with D = 16384, built with the
reference's `age` graph API and evaluated through llo.evaluate_device (value,
inputs resident in HBM) and llo.evaluate with host buffers (e2e).  A "step" is
one evaluation of that graph.  Metric: algorithmic HBM GB/s =
4 * (N + D + N) bytes / step time, whole job over all
ranks.  With --gpus N > 1 every rank evaluates its own replica of the chain
(the single-operator configs do not shard: "replicas only", weak scaling);
the batch-sharded MLP gradient evaluation with its NCCL allreduce is reported
under "workloads" for every N.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "elementwise_chain_hbm_gbps"
UNIT = "GB/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            data = json.load(fh)
        return float(data["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs"""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        clocks, reasons, maxes = [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            with open(self.path) as fh:
                for line in fh:
                    cols = [c.strip() for c in line.split(",")]
                    if len(cols) < 7:
                        continue
                    try:
                        clocks.append(float(cols[0]))
                        maxes.append(float(cols[1]))
                    except ValueError:
                        continue
                    for name, val in zip(names, cols[3:7]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if clocks:
            out["sm_mhz"] = float(np.median(clocks))
            out["sm_max_mhz"] = float(max(maxes))
            out["samples"] = len(clocks)
        out["reasons"] = sorted(reasons)
        return out


# ---- the chain graph ------------------------------------------------------------

def chain_inputs(side, seed_x=2, seed_b=3):
    x = np.random.default_rng(seed_x).uniform(0.5, 2.0, (side, side)).astype(np.float32)
    b = np.random.default_rng(seed_b).uniform(0.5, 2.0, (side,)).astype(np.float32)
    return x, b


def chain_graph(age, llo, x, b):
    side = x.shape[0]
    vx = llo.variable(x, "x")
    vb = llo.variable(b, "b")
    root = age.add(age.mul(age.exp(vx), age.extend(vb, 1, [side])), age.permute(vx, [1, 0]))
    return vx, vb, root


def chain_bytes(side):
    n = side * side
    return 4 * (n + side + n)


def cpu_chain(side, repeats):
    """The oracle (CPU restatement of the reference evaluator) on a bounded
    sample of the chain.  Returns (seconds per evaluation, GB/s)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util  # oracle access
    from cortenn_b200 import age, llo
    x, b = chain_inputs(side)
    _, _, root = chain_graph(age, llo, x, b)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        util.oracle.evaluate(root, np.dtype(np.float32))
        times.append(time.perf_counter() - t0)
    sec = float(np.median(times))
    return sec, chain_bytes(side) / sec / 1e9


# ---- reference arm ----------------------------------------------------------------

def run_reference(args, rank, world):
    if rank != 0:
        return
    side = 1 << (args.ref_log2n // 2)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from cortenn_b200 import age, llo
    x, b = chain_inputs(side)
    _, _, root = chain_graph(age, llo, x, b)
    for _ in range(args.warmup):
        util.oracle.evaluate(root, np.dtype(np.float32))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        util.oracle.evaluate(root, np.dtype(np.float32))
    total = time.perf_counter() - t0
    per_step = total / args.steps
    value = chain_bytes(side) / per_step / 1e9
    sample = ("the same chain on x=[%d,%d] (2^%d elements) per step; the reference's evaluator "
              "is single-threaded by construction" % (side, side, args.ref_log2n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "elementwise chain add(mul(exp(x),extend(b)),permute(x)), fp32, "
                               "bounded CPU sample", "sample_elements": side * side,
                   "full_elements": 1 << args.log2n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- own arm ------------------------------------------------------------------------

def run_own(args, rank, local_rank, world):
    os.environ.setdefault("LLO_CUDA_DEVICE", str(local_rank))
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import cortenn_b200  # noqa: F401  (fails loudly when the extension is missing)
    from cortenn_b200 import age, capi, llo
    if not llo.available():
        raise RuntimeError("bench.py: no usable B200 and there is no CPU fallback")
    ctx = capi.Context(handle=llo.context())

    def barrier():
        llo.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    side = 1 << (args.log2n // 2)
    f32 = np.dtype(np.float32)
    x, b = chain_inputs(side, seed_x=2 + rank)
    vx, vb, root = chain_graph(age, llo, x, b)
    nbytes = chain_bytes(side)

    # ---- value: inputs resident in HBM, device-timed --------------------------
    for _ in range(max(args.warmup, 3)):
        res = llo.evaluate_device(root, f32)
    del res
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = llo.launch_count()
    start, stop = ctx.event(), ctx.event()
    ctx.record(start)
    for _ in range(args.steps):
        res = llo.evaluate_device(root, f32)
    ctx.record(stop)
    ms_total = ctx.elapsed_ms(start, stop)
    barrier()
    launches = llo.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = world * nbytes / ms_step / 1e6
    del res

    # ---- e2e: host buffers in, host buffer out, copies inside the timed region --
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    llo.evaluate(root, f32)  # warm the staging buffers
    barrier()
    ctx.record(start)
    for _ in range(e2e_steps):
        vx.assign(x)                     # host -> pinned -> device at the next use
        out = llo.evaluate(root, f32)    # device -> host numpy
    ctx.record(stop)
    e2e_ms = max_over_ranks(ctx.elapsed_ms(start, stop)) / e2e_steps
    e2e = {"value": world * nbytes / e2e_ms / 1e6, "unit": UNIT,
           "h2d_bytes_per_step": int(x.nbytes), "d2h_bytes_per_step": int(out.nbytes),
           "steps": e2e_steps, "ms_per_step": e2e_ms}
    del out

    peak, peak_src = measured_peaks()
    roofline = {"bound": "hbm", "achieved": nbytes / ms_step / 1e6, "peak": peak, "unit": "GB/s",
                "frac": nbytes / ms_step / 1e6 / peak, "traffic": None, "peak_source": peak_src,
                "kernel": "whole step (%d launch(es) per evaluation)" % (launches // args.steps)}
    traffic_path = os.path.join(ROOT, "profiles", "chain_traffic.json")
    if os.path.exists(traffic_path):
        with open(traffic_path) as fh:
            roofline["traffic"] = json.load(fh).get("dram_bytes_per_launch")

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cside = 1 << (args.cpu_log2n // 2)
        sec, gbps = cpu_chain(cside, 2)
        cpu_baseline = {"value": gbps, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": "oracle (CPU restatement of llo::Evaluator) on the same chain at "
                                  "x=[%d,%d] (2^%d elements), %.2f s per evaluation, single thread "
                                  "like the reference" % (cside, cside, args.cpu_log2n, sec)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "elementwise chain add(mul(exp(x),extend(b,1,{D})),permute(x,{1,0})), "
                                   "x=[%d,%d] fp32 (2^%d elements), ade graph via age API -> llo.evaluate_device"
                                   % (side, side, args.log2n),
                       "bytes_per_step": nbytes,
                       "l2": "inputs larger than L2 (x is %d MiB, L2 is 126 MB)" % (x.nbytes >> 20),
                       "multi_gpu": "independent replicas per rank (single-operator configs do not shard)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--log2n", type=int, default=28, help="elements of the chain (2^k, k even)")
    ap.add_argument("--cpu-log2n", type=int, default=22, help="elements of the CPU-baseline sample")
    ap.add_argument("--ref-log2n", type=int, default=20, help="elements per step of the reference arm")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_own(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
