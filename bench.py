#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 LLO evaluator (driver contract).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Headline workload (BASELINE.json configs[1]): the elementwise chain
    out = add(mul(exp(x), extend(b, 1, {D})), permute(x, {1,0}))
over N = 2^28 fp32 elements, x as [D, D] with D = 16384, built with the
reference's `age` graph API and evaluated through llo.evaluate_device (value,
inputs resident in HBM) and llo.evaluate with pinned host buffers (e2e).  A
"step" is one evaluation of that graph.  Metric: algorithmic HBM GB/s =
4 * (N + D + N) bytes / step time, whole job over all
ranks.  With --gpus N > 1 every rank evaluates its own replica of the chain
(the single-operator configs do not shard: "replicas only", weak scaling); the
workload that shards is the MLP gradient evaluation (`scaling_workload`).

The other BASELINE.json configurations are under "workloads", each with its own
`roofline` block and a `verified` flag (the timed output is compared, outside
the timed region, with float64 numpy at the full size), and once more, in a few
hundred bytes, under "summary" at the very end of the line.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import workloads  # noqa: E402

METRIC = "elementwise_chain_hbm_gbps"
UNIT = "GB/s"
F32 = np.dtype(np.float32)
F64 = np.dtype(np.float64)


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


def fma_peaks():
    """FFMA / DFMA issue peaks: 148 SMs x 128 (64) lanes x 2 flop x the SM clock the
    driver recorded (MEASURED_PEAKS.json sm_max_mhz); tools/probes/fma_peak.cu measured
    72.5 / 37.0 TFLOP/s on this pool (profiles/fma_peak.json), 97 % of that product"""
    mhz = 1965.0
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            mhz = float(json.load(fh).get("sm_max_mhz", mhz))
    nominal = 148 * 128 * 2 * mhz * 1e6 / 1e12
    probe = os.path.join(ROOT, "profiles", "fma_peak.json")
    if os.path.exists(probe):
        with open(probe) as fh:
            data = json.load(fh)
        return float(data["ffma_tflops"]), float(data["dfma_tflops"]), "probe (profiles/fma_peak.json)", nominal
    return nominal, nominal / 2, "148 SM x 128 lanes x 2 x sm_max_mhz", nominal


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
            time.sleep(0.3)    # let the first samples land before the timed region
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
        clocks, reasons, maxes, power = [], set(), [], []
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
                        power.append(float(cols[2]))
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
            out["power_w_max"] = float(max(power)) if power else None
            out["samples"] = len(clocks)
        out["reasons"] = sorted(reasons)
        return out


# ---- CPU baseline / reference arm ------------------------------------------------------

def oracle_module():
    """the oracle (CPU restatement of the reference evaluator): checker and
    CPU baseline only, never on the product path"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    return util.oracle


def cpu_chain(side, repeats):
    """The oracle on a bounded sample of the chain: (seconds per evaluation, GB/s)."""
    from cortenn_b200 import age, llo
    oracle = oracle_module()
    x, b = workloads.chain_inputs(side)
    _, _, root = workloads.chain_graph(age, llo, x, b)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        oracle.evaluate(root, F32)
        times.append(time.perf_counter() - t0)
    sec = float(np.median(times))
    return sec, workloads.chain_bytes(side) / sec / 1e9


def cpu_worker(args):
    """one of the `nproc` independent oracle processes of the all-core figures.
    With --cpu-steps K (the reference arm): build the graph, warm up, print "ready", wait for
    a line on stdin (all workers start their timed loop together), time K evaluations."""
    side = 1 << (args.cpu_log2n // 2)
    if args.cpu_steps <= 0:
        sec, gbps = cpu_chain(side, max(1, args.cpu_repeats))
        print(json.dumps({"sec": sec, "gbps": gbps}), flush=True)
        return
    from cortenn_b200 import age, llo
    oracle = oracle_module()
    x, b = workloads.chain_inputs(side)
    _, _, root = workloads.chain_graph(age, llo, x, b)
    for _ in range(max(0, args.cpu_warmup)):
        oracle.evaluate(root, F32)
    print(json.dumps({"ready": 1}), flush=True)
    sys.stdin.readline()
    t0 = time.perf_counter()
    for _ in range(args.cpu_steps):
        oracle.evaluate(root, F32)
    print(json.dumps({"sec": time.perf_counter() - t0, "steps": args.cpu_steps}), flush=True)


def cpu_all_cores(args):
    """the reference has no threading, so the all-core figure is
    `nproc` independent single-threaded oracle processes, each on the same chain sample"""
    procs = os.cpu_count() or 1
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", "--cpu-log2n", str(args.cpu_log2n),
           "--cpu-repeats", "2"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    t0 = time.perf_counter()
    running = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env, text=True)
               for _ in range(procs)]
    rates = []
    for p in running:
        try:
            out, _ = p.communicate(timeout=180)
            rates.append(json.loads(out.strip().splitlines()[-1])["gbps"])
        except Exception:
            p.kill()
    if not rates:
        return None
    return {"GBps": float(sum(rates)), "processes": len(rates), "wall_s": time.perf_counter() - t0}


def cpu_side_baselines(age, llo):
    """the oracle (one thread, like the reference) on bounded samples
    of configs 1, 3, 4 and 5; extrapolations are linear in elements / flops and say so"""
    oracle = oracle_module()

    def timed(root, dtype=F32, repeats=1):
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            oracle.evaluate(root, dtype)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best

    out = {}
    # config 1: the reference's CPU-runnable unit graphs, float [64,64]
    rng = np.random.default_rng(1)
    x = rng.uniform(0.5, 1.5, (64, 64)).astype(np.float32)
    y = rng.uniform(0.5, 1.5, (64, 64)).astype(np.float32)
    vx, vy = llo.variable(x, "x"), llo.variable(y, "y")
    graphs = {"add": age.add(vx, vy), "mul": age.mul(vx, vy), "exp": age.exp(vx),
              "rsum1": age.reduce_sum(vx, 1), "rsum0": age.reduce_sum0(vx), "matmul": age.matmul(vx, vy)}
    out["c1_us"] = {k: round(timed(r, repeats=10) * 1e6, 1) for k, r in graphs.items()}
    # config 3: reductions on a 2^20 sample
    side = 1024
    xs = np.random.default_rng(4).uniform(0.0, 1.0, (side, side)).astype(np.float32)
    vs = llo.variable(xs, "x")
    red = {"cols": age.reduce_sum(vs, 1), "all": age.reduce_sum0(vs),
           "rows": age.reduce_sum(age.permute(vs, [1, 0]), 1)}
    out["c3_GBps"] = {k: round(4.0 * side * side / timed(r) / 1e9, 4) for k, r in red.items()}
    out["c3_sample"] = "reduce_sum of [1024,1024] fp32 (2^20 of 2^28 elements)"
    # config 4: materialised [M,N,C] matmul of the reference at 64^3 and 128^3
    mm = {}
    for m in (64, 128):
        a = llo.variable(np.random.default_rng(5).uniform(-1, 1, (m, m)).astype(np.float32), "a")
        b = llo.variable(np.random.default_rng(6).uniform(-1, 1, (m, m)).astype(np.float32), "b")
        mm[str(m)] = round(2.0 * m ** 3 / timed(age.matmul(a, b)) / 1e6, 2)
    out["c4_MFLOPs"] = mm
    out["c4_note"] = "O(MNC) mapped elements: 8192^3 extrapolates to ~2.8e5 s and needs 5 x 2.2 TB"
    # config 5: down-scaled MLP 16-32-32-10, batch 32, six derived gradients
    sizes, batch = (16, 32, 32, 10), 32
    X, Y, params = workloads.mlp_data(batch, sizes=sizes)
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params)
    grads = [llo.derive(loss, p) for p in pvars]
    t0 = time.perf_counter()
    for g in grads:
        oracle.evaluate(g, F32)
    sec = time.perf_counter() - t0
    flops = workloads.mlp_flops(batch, sizes)
    out["c5"] = {"grad_evals_per_s": round(1.0 / sec, 3), "sample": "MLP 16-32-32-10, batch 32",
                 "MFLOPs": round(flops / sec / 1e6, 3),
                 "extrapolated_s_per_grad_eval_at_65536": round(workloads.mlp_flops(65536) / (flops / sec), 0)}
    return out


def reference_all_cores(args, side):
    """`nproc` single-threaded oracle processes, each evaluating the chain sample args.steps
    times after args.warmup warm-ups, started together: (seconds for the slowest, processes)"""
    procs = os.cpu_count() or 1
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", "--cpu-log2n", str(args.ref_log2n),
           "--cpu-steps", str(args.steps), "--cpu-warmup", str(args.warmup)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for key in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(key, None)
    running = [subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                env=env, text=True) for _ in range(procs)]
    try:
        for p in running:
            if "ready" not in json.loads(p.stdout.readline()):
                raise RuntimeError("worker did not come up")
        for p in running:
            p.stdin.write("go\n")
            p.stdin.flush()
        secs = [json.loads(p.stdout.readline())["sec"] for p in running]
        for p in running:
            p.wait(timeout=60)
        return max(secs), procs
    except Exception:
        for p in running:
            p.kill()
        return None, 0


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (the oracle port: the reference does
    not build here) on the box's host cores.  The reference evaluator is
    single-threaded (no threads, SIMD or BLAS: SURVEY.md 2.2), so "all the host threads it can
    use" is one evaluation per core: `nproc` independent processes, like this repo's own
    replicas of the chain at N > 1.  The single-thread figure rides along."""
    if rank != 0:
        return
    side = 1 << (args.ref_log2n // 2)
    from cortenn_b200 import age, llo
    oracle = oracle_module()
    x, b = workloads.chain_inputs(side)
    _, _, root = workloads.chain_graph(age, llo, x, b)
    for _ in range(args.warmup):
        oracle.evaluate(root, F32)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.evaluate(root, F32)
    single_step = (time.perf_counter() - t0) / args.steps
    single = workloads.chain_bytes(side) / single_step / 1e9
    slowest, procs = (None, 0) if args.ref_single else reference_all_cores(args, side)
    if slowest is None:
        per_step, value, cores = single_step, single, 1
        sample = ("the same chain on x=[%d,%d] (2^%d elements) per step; 1 thread: the reference "
                  "evaluator has no threads, SIMD or BLAS (SURVEY.md 2.2)" % (side, side, args.ref_log2n))
    else:
        per_step = slowest / args.steps
        value = procs * workloads.chain_bytes(side) / per_step / 1e9
        cores = procs
        sample = ("%d independent single-threaded evaluations (one process per core: the reference "
                  "evaluator has no threads, SIMD or BLAS, SURVEY.md 2.2) of the same chain on "
                  "x=[%d,%d] (2^%d elements) per step, started together, slowest process timed"
                  % (procs, side, side, args.ref_log2n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "elementwise chain add(mul(exp(x),extend(b,1,{D})),permute(x,{1,0})), fp32, "
                               "bounded CPU sample of x=[16384,16384]", "sample_elements": side * side,
                   "full_elements": 1 << args.log2n, "processes": max(cores, 1)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "single_thread": {"value": single, "unit": UNIT, "ms_per_step": single_step * 1e3}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- own arm ------------------------------------------------------------------------------

class Timer:
    """CUDA events on the evaluator's stream + barrier / max over ranks"""

    def __init__(self, llo, capi, dist, torch):
        self.llo = llo
        self.ctx = capi.Context(handle=llo.context())
        self.dist = dist
        self.torch = torch
        self.start, self.stop = self.ctx.event(), self.ctx.event()

    def barrier(self):
        self.llo.sync()
        if self.dist is not None:
            self.torch.cuda.synchronize()
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.dist is None:
            return ms
        t = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup, finish=None, collective=True):
        """ms per step of `fn`, max over ranks; also returns the launches.
        `finish` (pipelined callers) completes whatever the last step left in
        flight; it runs inside the timed region and after the warm-up.
        collective=False: this rank alone (no barrier, no max over ranks)."""
        for _ in range(warmup):
            fn()
        if finish is not None:
            finish()
        if collective:
            self.barrier()
        else:
            self.llo.sync()
        launches0 = self.llo.launch_count()
        self.ctx.record(self.start)
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        self.ctx.record(self.stop)
        ms = self.ctx.elapsed_ms(self.start, self.stop)
        if collective:
            self.barrier()
        launches = self.llo.launch_count() - launches0
        return (self.max_over_ranks(ms) if collective else ms) / steps, launches


def hbm_block(nbytes, ms, kernel):
    peak, _ = measured_peaks()
    gbps = nbytes / ms / 1e6
    return {"bound": "hbm", "achieved": round(gbps, 1), "peak": peak, "unit": "GB/s",
            "frac": round(gbps / peak, 4), "kernel": kernel}


def chain_verify(got, x, b):
    """max relative error of the chain's full output against float64 numpy"""
    side = x.shape[0]
    worst = 0.0
    b64 = b.astype(np.float64)[None, :]
    for lo in range(0, side, 2048):
        hi = min(side, lo + 2048)
        want = np.exp(x[lo:hi].astype(np.float64)) * b64 + x[:, lo:hi].T.astype(np.float64)
        worst = max(worst, float((np.abs(got[lo:hi].astype(np.float64) - want) / np.abs(want)).max()))
    return worst


def mlp_reference(X, Y, params, batch):
    """float64 gradients of the MLP loss and the sum |terms| bound of every element"""
    X, Y = X.astype(np.float64), Y.astype(np.float64)
    P = [p.astype(np.float64) for p in params]
    hs = [X]
    for i in range(len(P) // 2):
        hs.append(1.0 / (1.0 + np.exp(-(hs[-1] @ P[2 * i] + P[2 * i + 1]))))
    grads, bounds = [None] * len(P), [None] * len(P)
    delta = 2.0 * (hs[-1] - Y) / batch
    for i in reversed(range(len(P) // 2)):
        dz = delta * hs[i + 1] * (1.0 - hs[i + 1])
        grads[2 * i], bounds[2 * i] = hs[i].T @ dz, np.abs(hs[i]).T @ np.abs(dz)
        grads[2 * i + 1], bounds[2 * i + 1] = dz.sum(axis=0), np.abs(dz).sum(axis=0)
        delta = dz @ P[2 * i].T
    return grads, bounds


def traffic_probe(args):
    """DRAM bytes of one launch of the headline kernel, measured now: one short ncu pass
    (dram__bytes_read.sum + dram__bytes_write.sum) over tools/prof_cases.py --case chain,
    which launches the same kernel on the same shapes.  Returns (bytes, source)."""
    fallback = os.path.join(ROOT, "profiles", "chain_traffic.json")

    def from_file(why):
        if os.path.exists(fallback):
            with open(fallback) as fh:
                return json.load(fh).get("dram_bytes_per_launch"), "profiles/chain_traffic.json (%s)" % why
        return None, why
    if args.no_traffic_probe:
        return from_file("probe disabled")
    ncu = None
    for cand in ("ncu", "/usr/local/cuda/bin/ncu"):
        try:
            subprocess.run([cand, "--version"], capture_output=True, timeout=20, check=True)
            ncu = cand
            break
        except Exception:
            continue
    if ncu is None:
        return from_file("ncu not available")
    case = [sys.executable, os.path.join(ROOT, "tools", "prof_cases.py"), "--case", "chain",
            "--log2n", str(args.log2n), "--iters", "2"]
    try:
        plain = subprocess.run(case, capture_output=True, timeout=120)
        if plain.returncode != 0:
            return from_file("probe case failed")
        fd, log = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        subprocess.run([ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
                        "-k", "regex:vm_pair4", "-s", "1", "-c", "1", "--csv", "--log-file", log] + case,
                       capture_output=True, timeout=180)
        total = 0.0
        import csv
        with open(log) as fh:
            rows = [r for r in csv.reader(fh) if len(r) > 5]
        os.unlink(log)
        hdr = rows[0]
        i_name, i_unit, i_val = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        seen = 0
        for r in rows[1:]:
            if r[i_name].startswith("dram__bytes_"):
                total += float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
                seen += 1
        if seen < 2:
            return from_file("ncu returned no dram metrics")
        return total, "ncu dram__bytes_read.sum + dram__bytes_write.sum, measured in this run"
    except Exception as exc:  # noqa: BLE001
        return from_file("probe error: %s" % type(exc).__name__)


def side_workloads(args, timer, age, llo, rank, world):
    """BASELINE.json configs 2 (per operator), 3, 4, 5"""
    out = {}
    steps, warm = max(3, min(args.steps, 10)), 3
    check = not args.skip_verify
    if world == 1 and not args.skip_side:
        side = 1 << (args.log2n // 2)
        n = side * side
        x = np.random.default_rng(4).uniform(0.0, 1.0, (side, side)).astype(np.float32)
        vx = llo.variable(x, "x")
        x64 = x.astype(np.float64) if check else None
        # config 3: reductions of [D, D], fixed tree order.  "cols" keeps the fastest dim
        # (strided columns), "rows" = reduce(permute(x)) with the permute fused, "all" = scalar
        cases = {
            "reduce_sum_cols": (age.reduce_sum(vx, 1), "col_fold_kernel", lambda: x64.sum(axis=0)),
            "reduce_max_cols": (age.reduce_max(vx, 1), "col_fold_kernel", lambda: x64.max(axis=0)),
            "reduce_sum_all": (age.reduce_sum0(vx), "row_fold_kernel", lambda: np.array(x64.sum())),
            "reduce_max_all": (age.reduce_max0(vx), "row_fold_kernel", lambda: np.array(x64.max())),
            "reduce_sum_rows": (age.reduce_sum(age.permute(vx, [1, 0]), 1), "row_fold_kernel", lambda: x64.sum(axis=1)),
            "reduce_max_rows": (age.reduce_max(age.permute(vx, [1, 0]), 1), "row_fold_kernel", lambda: x64.max(axis=1)),
        }
        for name, (root, kernel, ref) in cases.items():
            holder = {}

            def step(r=root):
                holder["res"] = llo.evaluate_device(r, F32)
            ms, launches = timer.timed(step, steps, warm)
            entry = {"ms": round(ms, 4), "roofline": hbm_block(4 * n, ms, kernel), "launches": launches // steps}
            if check:
                want = ref()
                got = holder["res"].numpy(F32).reshape(want.shape).astype(np.float64)
                err = float((np.abs(got - want) / np.abs(want)).max())
                entry["max_rel_err"] = err
                entry["verified"] = bool(err <= 1e-6)
            out[name] = entry
        del cases, x64
        x4 = x.reshape(128, 128, n // (128 ** 3), 128) if n >= 128 ** 4 else None
        if x4 is not None:
            vx4 = llo.variable(x4, "x4")
            for name, root in (("reduce_sum_128x4_dim2", age.reduce_sum(vx4, 2)),
                               ("reduce_max_128x4_dim2", age.reduce_max(vx4, 2))):
                ms, launches = timer.timed(lambda r=root: llo.evaluate_device(r, F32), steps, warm)
                out[name] = {"ms": round(ms, 4), "roofline": hbm_block(4 * n, ms, "col_fold_kernel")}
            del vx4
        # config 2, every operator stand-alone: algorithmic bytes =
        # 4 * (distinct input elements + output elements)
        y = np.random.default_rng(5).uniform(0.5, 2.0, (side, side)).astype(np.float32)
        bb = np.random.default_rng(3).uniform(0.5, 2.0, (side,)).astype(np.float32)
        vy, vbb = llo.variable(y, "y"), llo.variable(bb, "b")
        ops = {
            "exp": (age.exp(vx), 2 * n), "sin": (age.sin(vx), 2 * n), "sqrt": (age.sqrt(vx), 2 * n),
            "abs": (age.abs(vx), 2 * n), "neg": (age.neg(vx), 2 * n), "log": (age.log(vy), 2 * n),
            "add": (age.add(vx, vy), 3 * n), "sub": (age.sub(vx, vy), 3 * n),
            "mul": (age.mul(vx, vy), 3 * n), "div": (age.div(vx, vy), 3 * n),
            "pow": (age.pow(vy, vx), 3 * n), "lt": (age.lt(vx, vy), 3 * n), "eq": (age.eq(vx, vy), 3 * n),
            "extend": (age.extend(vbb, 1, [side]), n + side),
            "mul_extend": (age.mul(vx, age.extend(vbb, 1, [side])), 2 * n + side),
            "permute": (age.permute(vx, [1, 0]), 2 * n),
        }
        peak, _ = measured_peaks()
        elem = {}
        for name, (root, elems) in ops.items():
            ms, launches = timer.timed(lambda r=root: llo.evaluate_device(r, F32), steps, warm)
            gbps = 4.0 * elems / ms / 1e6
            elem[name] = [round(gbps, 1), round(gbps / peak, 3)]
        out["elementwise_ops"] = {"GBps_and_frac_of_measured_hbm": elem,
                                  "kernels": "vm_flat_kernel (permute: vm_tile4_kernel)"}
        del ops, vy, vbb, y
        if x4 is not None:
            vx4 = llo.variable(x4, "x4")
            b4 = llo.variable(bb[:128].copy(), "b4")
            rest = list(x4.shape[:-1])[::-1]
            root = age.add(age.mul(age.exp(vx4), age.extend(b4, 1, rest)), age.permute(vx4, [1, 0, 2, 3]))
            ms, launches = timer.timed(lambda: llo.evaluate_device(root, F32), steps, warm)
            out["chain_128x4"] = {"ms": round(ms, 4), "roofline": hbm_block(4 * (2 * n + 128), ms, "vm_pair4_kernel")}
            del vx4, b4, root
        del vx, x, x4, bb
        # config 4: matmul 8192^3 — fp32 on the FFMA pipe (exact path, FFMA2 issue), fp64 on
        # the DFMA pipe, and the opt-in 3xTF32 tensor-core path (tcgen05)
        m = args.matmul
        gsteps = max(3, steps // 2)
        fpeak, dpeak, fsrc, _ = fma_peaks()
        for name, dt, option, kernel in (("matmul_f32", np.float32, 0, "sgemm_tma_kernel<0,0> (FFMA2)"),
                                         ("matmul_f32_3xtf32", np.float32, 1, "tf32x3_kernel (tcgen05)"),
                                         ("matmul_f64", np.float64, 0, "dgemm_tma_kernel")):
            a = np.random.default_rng(5).uniform(-1, 1, (m, m)).astype(dt)
            b = np.random.default_rng(6).uniform(-1, 1, (m, m)).astype(dt)
            va, vb = llo.variable(a, "a"), llo.variable(b, "b")
            root = age.matmul(va, vb)
            holder = {}

            def step():
                holder["res"] = llo.evaluate_device(root, np.dtype(dt))
            llo.set_option("tf32x3", option)
            try:
                ms, launches = timer.timed(step, gsteps, 2)
            finally:
                llo.set_option("tf32x3", 0)
            tf = 2.0 * m ** 3 / ms / 1e9
            entry = {"ms": round(ms, 3), "mnk": [m, m, m], "launches": launches // gsteps}
            if option:
                entry["TFLOPs_fp32_equivalent"] = round(tf, 2)
                entry["speedup_vs_exact"] = round(tf / out["matmul_f32"]["roofline"]["achieved"], 2)
                entry["kernel"] = kernel
            else:
                peak_tf = fpeak if dt == np.float32 else dpeak
                entry["roofline"] = {"bound": "ffma" if dt == np.float32 else "dfma", "achieved": round(tf, 2),
                                     "peak": peak_tf, "unit": "TFLOP/s", "frac": round(tf / peak_tf, 4),
                                     "kernel": kernel, "peak_source": fsrc}
            if check:
                rows = np.random.default_rng(7).choice(m, 64, replace=False)
                got = holder["res"].numpy(np.dtype(dt))[rows].astype(np.float64)
                a64, b64 = a[rows].astype(np.float64), b.astype(np.float64)
                err = float((np.abs(got - a64 @ b64) / (np.abs(a64) @ np.abs(b64))).max())
                entry["max_err_over_sum_abs"] = err
                entry["verified"] = bool(err <= (1e-6 if dt == np.float32 else 1e-12))
            out[name] = entry
            del va, vb, a, b, root, holder
        # age.convolution lowered to im2col gather + one TMA GEMM
        nb, side_i, ic, oc, kw = 32, 128, 32, 64, 3
        img = np.random.default_rng(8).uniform(-1, 1, (nb, side_i, side_i, ic)).astype(np.float32)
        ker = np.random.default_rng(9).uniform(-1, 1, (kw, kw, ic, oc)).astype(np.float32)
        croot = age.convolution(llo.variable(img, "img"), llo.variable(ker, "k"))
        holder = {}

        def conv_step():
            holder["res"] = llo.evaluate_device(croot, F32)
        ms, launches = timer.timed(conv_step, steps, warm)
        st = llo.plan_stats()
        so = side_i - kw + 1
        cflops = 2.0 * nb * so * so * ic * kw * kw * oc
        entry = {"ms": round(ms, 4), "TFLOPs": round(cflops / ms / 1e9, 2), "launches": launches // steps,
                 "shape": "img [%d,%d,%d,%d] x kernel [%d,%d,%d,%d], VALID" % (nb, side_i, side_i, ic, kw, kw, ic, oc),
                 "plan": {k: st[k] for k in ("gemms", "gathers", "generic", "programs")}}
        if check:
            got = holder["res"].numpy(F32)[:2].astype(np.float64)
            want = np.zeros_like(got)
            for a in range(kw):
                for b in range(kw):
                    want += np.einsum("nwhc,co->nwho", img[:2, a:a + so, b:b + so, :].astype(np.float64),
                                      ker[a, b].astype(np.float64))
            err = float(np.abs(got - want).max() / np.abs(want).max())
            entry["max_err_over_max_abs"] = err
            entry["verified"] = bool(err <= 1e-5)
        out["convolution_f32"] = entry
        del img, ker, croot, holder
    if not args.skip_mlp:
        out.update(mlp_workload(args, timer, age, llo, rank, world, steps, warm, check))
    return out


def mlp_workload(args, timer, age, llo, rank, world, steps, warm, check):
    """config 5: MLP gradient evaluation, batch sharded over the ranks (strong scaling)"""
    from cortenn_b200.sharded import BatchShardedExecutor
    out = {}
    batch = args.mlp_batch
    X, Y, params = workloads.mlp_data(batch)
    ex = BatchShardedExecutor(batch, rank=rank, world=world)
    loss, pvars, _ = workloads.mlp_graph(age, llo, ex.shard(X), ex.shard(Y), params, global_batch=batch)
    holder = {}

    def grad_eval():
        holder["buf"], holder["pack"] = ex.grad_eval_device(loss, pvars, F32)
    ms, launches = timer.timed(grad_eval, steps, warm)
    # host time to plan + enqueue one evaluation (no synchronisation): it
    # must stay below the device time or the GPU starves
    timer.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        grad_eval()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    timer.barrier()
    flops = workloads.mlp_flops(batch)
    fpeak, _, fsrc, _ = fma_peaks()
    entry = {
        "ms": round(ms, 4), "grad_evals_per_s": round(1e3 / ms, 2), "batch": batch, "rows_per_rank": ex.local_rows(),
        "roofline": {"bound": "ffma", "achieved": round(flops / ms / 1e9, 2), "peak": fpeak * world,
                     "unit": "TFLOP/s (GEMM flops of one grad-eval, all ranks)",
                     "frac": round(flops / ms / 1e9 / (fpeak * world), 4), "kernel": "sgemm_* (FFMA2) + vm_flat + folds"},
        "allreduce_bytes": 4 * 1863690, "launches": launches // steps, "host_enqueue_ms": round(host_ms, 3),
        "graph_passes": "llo.optimize(OPT_ALL) on [loss] + llo.derive(...) (div_cancel %d, shared %d, functors %d -> %d)"
                        % (ex.opt_report["div_cancel"], ex.opt_report["shared"], ex.opt_report["functors_before"],
                           ex.opt_report["functors_after"]),
        "scaling": "strong"}
    got = None
    if check:
        got = ex.backend.to_host(holder["buf"], holder["pack"])
    if check and world == 1:
        want, bounds = mlp_reference(X, Y, params, batch)
        worst = max(float((np.abs(g.reshape(w.shape).astype(np.float64) - w) / bnd).max())
                    for g, w, bnd in zip(got, want, bounds))
        entry["max_err_over_sum_abs_terms"] = worst
        entry["verified"] = bool(worst <= 2e-5)
    if world > 1:
        # the gradient the ranks agreed on against the single-GPU gradient of the whole batch,
        # evaluated now by rank 0 alone on its own GPU (and timed: the N = 1 time of this run)
        n1 = {"ms": None}
        if rank == 0:
            solo = BatchShardedExecutor(batch, rank=0, world=1, backend=ex.backend.solo())
            loss1, pvars1, _ = workloads.mlp_graph(age, llo, X, Y, params, global_batch=batch)
            solo_holder = {}

            def solo_eval():
                solo_holder["out"] = solo.grad_eval_device(loss1, pvars1, F32)
            n1["ms"], _ = timer.timed(solo_eval, max(3, steps // 2), 2, collective=False)
            if check:
                # fp32 sums of 65536 signed terms, associated differently by N ranks (and by
                # the shard-sized split-K): the two results differ by rounding of the partial
                # sums, i.e. relative to sum |terms| (float64 reference below), not to the
                # possibly cancelled sum itself
                ref = solo.backend.to_host(*solo_holder["out"])
                want, bounds = mlp_reference(X, Y, params, batch)
                worst = max(float((np.abs(g.astype(np.float64) - r.astype(np.float64)).reshape(b.shape) / b).max())
                            for g, r, b in zip(got, ref, bounds))
                entry["grad_max_diff_vs_n1_over_sum_abs_terms"] = worst
                entry["grad_matches_n1"] = bool(worst <= 2e-6)
                exact = max(float((np.abs(g.reshape(w.shape).astype(np.float64) - w) / b).max())
                            for g, w, b in zip(got, want, bounds))
                entry["max_err_over_sum_abs_terms"] = exact
                entry["verified"] = bool(exact <= 2e-5)
            del solo, loss1, pvars1, solo_holder
        timer.barrier()
        if rank == 0 and n1["ms"]:
            entry["n1_ms_same_run"] = round(n1["ms"], 4)
            entry["speedup_vs_n1_same_run"] = round(n1["ms"] / ms, 3)
    out["mlp_grad_eval"] = entry
    del X, Y
    llo.set_option("tf32x3", 1)
    try:
        ms_tc, _ = timer.timed(grad_eval, steps, warm)
    finally:
        llo.set_option("tf32x3", 0)
    out["mlp_grad_eval_3xtf32"] = {"ms": round(ms_tc, 4), "grad_evals_per_s": round(1e3 / ms_tc, 2),
                                   "path": "opt-in tensor-core GEMMs (tf32x3 = 1); same graph, same allreduce"}
    return out


def summarise(side, value_frac, verified):
    """configs 2-5 in a few hundred bytes, last in the line (the driver keeps the tail)"""
    def rf(name):
        e = side.get(name)
        if not e:
            return None
        r = e.get("roofline", {})
        return [r.get("achieved"), r.get("frac"), e.get("verified")]
    s = {"c2_chain_frac_verified": [round(value_frac, 4), verified]}
    c3 = {k.replace("reduce_", ""): rf(k) for k in side if k.startswith("reduce_") and "128x4" not in k}
    if c3:
        s["c3_reduce_GBps_frac_verified"] = c3
    c4 = {k.replace("matmul_", ""): rf(k) for k in ("matmul_f32", "matmul_f64") if k in side}
    if "matmul_f32_3xtf32" in side:
        e = side["matmul_f32_3xtf32"]
        c4["3xtf32"] = [e.get("TFLOPs_fp32_equivalent"), None, e.get("verified")]
    if c4:
        s["c4_matmul_TFLOPs_frac_verified"] = c4
    m = side.get("mlp_grad_eval")
    if m:
        s["c5_mlp"] = {"ms": m["ms"], "evals_per_s": m["grad_evals_per_s"], "frac_ffma": m["roofline"]["frac"],
                       "verified": m.get("verified"), "grad_matches_n1": m.get("grad_matches_n1"),
                       "speedup_vs_n1": m.get("speedup_vs_n1_same_run")}
    return s


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned
    buffer is allocated or touched: pinned pages then come from that node's memory and the
    PCIe copies of the e2e steps do not cross the socket interconnect (round 1: all eight
    ranks streamed through one node, 8.2 GB/s per GPU at N = 8).  Host plumbing only; a
    no-op where sysfs does not say."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id",
                              "--format=csv,noheader"], capture_output=True, text=True, timeout=20)
        bus = out.stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"numa_node": node, "cpus": len(allowed)}
    except Exception:  # noqa: BLE001
        pass
    return None


def run_own(args, rank, local_rank, world):
    os.environ.setdefault("LLO_CUDA_DEVICE", str(local_rank))
    placement = bind_to_gpu_numa_node(local_rank)
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
    timer = Timer(llo, capi, dist, torch)
    if args.mlp_only:
        res = mlp_workload(args, timer, age, llo, rank, world, max(3, min(args.steps, 10)), 3,
                           not args.skip_verify)
        if rank == 0:
            print(json.dumps(res), flush=True)
        if dist is not None:
            timer.barrier()
            llo.comm_destroy()
            dist.destroy_process_group()
        return

    side = 1 << (args.log2n // 2)
    x, b = workloads.chain_inputs(side, seed_x=2 + rank)
    vx, vb, root = workloads.chain_graph(age, llo, x, b)
    nbytes = workloads.chain_bytes(side)
    warmup = max(args.warmup, 3)

    # ---- value: inputs resident in HBM, device-timed --------------------------
    last = {}

    def chain_step():
        # the result handle of the last step is kept for the parity check: two output
        # blocks alternate in the stream-ordered pool, in the warm-up as in the timed steps
        last["res"] = llo.evaluate_device(root, F32)
    for _ in range(warmup):
        chain_step()
    timer.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, launches = timer.timed(chain_step, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else {}
    value = world * nbytes / ms_step / 1e6
    verified = None
    chain_err = None
    if rank == 0 and not args.skip_verify:
        chain_err = chain_verify(last["res"].numpy(F32), x, b)
        verified = bool(chain_err <= 1e-6)
    last.clear()

    # ---- e2e: pinned host buffers in, host buffer out, copies inside the timed region --
    # Every step uploads its input (1 GiB) from pinned host memory, evaluates the graph and
    # reads the whole result (1 GiB) back to the host.  The public API pipelines the steps:
    # Variable.assign_async + llo.evaluate_async put the upload of step i + 1 on the
    # host->device copy engine while the result of step i drains over the device->host one
    # (PCIe is full duplex); .result() of step i is taken after step i + 1 has been queued.
    # The strictly synchronous form (assign; evaluate) is timed too and reported beside it.
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    xs = [llo.pinned_empty(list(x.shape), F32) for _ in range(2)]
    for xp in xs:
        xp[...] = x
    holder = {"pending": None, "out": None, "i": 0}

    def e2e_step():
        xp = xs[holder["i"] & 1]                    # inputs alternate between two pinned blocks
        holder["i"] += 1
        vx.assign_async(xp)                         # pinned host -> device, upload lane
        nxt = llo.evaluate_async(root, F32)         # kernels + device -> pinned host, download lane
        if holder["pending"] is not None:
            holder["out"] = holder["pending"].result()   # the previous step's result, on the host
        holder["pending"] = nxt

    def e2e_drain():
        if holder["pending"] is not None:
            holder["out"] = holder["pending"].result()
            holder["pending"] = None

    # warm-up: pinned result blocks (two alive at a time) come from a pool
    e2e_ms, _ = timer.timed(e2e_step, e2e_steps, 3, finish=e2e_drain)
    out = holder["out"]

    def sync_step():
        vx.assign(xs[0])                            # pinned host -> device (DMA), waits
        holder["out"] = llo.evaluate(root, F32)     # device -> pinned host numpy, waits
    sync_ms, _ = timer.timed(sync_step, max(2, e2e_steps // 2), 2)
    e2e = {"value": world * nbytes / e2e_ms / 1e6, "unit": UNIT,
           "h2d_bytes_per_step": int(x.nbytes), "d2h_bytes_per_step": int(out.nbytes),
           "steps": e2e_steps, "ms_per_step": e2e_ms,
           "path": "Variable.assign_async(pinned) + llo.evaluate_async -> .result(), pipelined",
           "host_placement": placement,
           "pcie_GBps_each_way": (x.nbytes + out.nbytes) / 2 / e2e_ms / 1e6,
           "synchronous": {"value": world * nbytes / sync_ms / 1e6, "ms_per_step": sync_ms,
                           "path": "Variable.assign + llo.evaluate, copies back to back"}}
    del out, holder, xs

    peak, peak_src = measured_peaks()
    per_eval = max(1, launches // args.steps)
    roofline = {"bound": "hbm", "achieved": nbytes / ms_step / 1e6, "peak": peak, "unit": "GB/s",
                "frac": nbytes / ms_step / 1e6 / peak, "traffic": None, "peak_source": peak_src,
                "frac_of_nominal_8TBs": nbytes / ms_step / 1e6 / 8000.0,
                "kernel": "vm_pair4_kernel<float> (%d launch per step)" % per_eval,
                "algorithmic_bytes_per_launch": nbytes}

    cpu_baseline = None
    cpu_more = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cside = 1 << (args.cpu_log2n // 2)
        sec, gbps = cpu_chain(cside, args.cpu_repeats)
        cpu_baseline = {"value": gbps, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": "oracle on the same chain at x=[%d,%d] (2^%d of 2^%d elements), %d x %.2f s, 1 thread "
                                  "like the reference; host has %d cores"
                                  % (cside, cside, args.cpu_log2n, args.log2n, args.cpu_repeats, sec, os.cpu_count() or 0)}
        cpu_more = cpu_side_baselines(age, llo)
        cpu_more["all_cores_chain"] = cpu_all_cores(args)

    del vx, vb, root
    side_results = side_workloads(args, timer, age, llo, rank, world)
    if rank == 0 and world == 1:
        # after every timed region: one short ncu pass for the DRAM bytes of the headline kernel
        roofline["traffic"], roofline["traffic_source"] = traffic_probe(args)
    del x, b

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "chain add(mul(exp(x),extend(b,1,{D})),permute(x,{1,0})), x=[%d,%d] fp32, age graph -> "
                                   "llo.evaluate_device" % (side, side),
                       "bytes_per_step": nbytes,
                       "l2": "inputs larger than L2 (x is %d MiB, L2 is 126 MB)" % ((4 * side * side) >> 20),
                       "multi_gpu": "replicas per rank (single-operator configs do not shard); see scaling_workload"},
            "verified": verified, "max_rel_err_vs_float64": chain_err,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        if cpu_more is not None:
            line["cpu_baselines"] = cpu_more
        line["workloads"] = side_results
        mlp = side_results.get("mlp_grad_eval")
        if mlp is not None:
            line["scaling_workload"] = {
                "name": "mlp_grad_eval (MLP 784-1024-1024-10, batch %d split over the ranks, NCCL allreduce)" % args.mlp_batch,
                "scaling": "strong", "n_gpus": world, "ms": mlp["ms"], "grad_evals_per_s": mlp["grad_evals_per_s"],
                "n1_ms_same_run": mlp.get("n1_ms_same_run", mlp["ms"] if world == 1 else None),
                "speedup_vs_n1_same_run": mlp.get("speedup_vs_n1_same_run", 1.0 if world == 1 else None),
                "grad_matches_n1": mlp.get("grad_matches_n1", mlp.get("verified") if world == 1 else None)}
        line["summary"] = summarise(side_results, roofline["frac"], verified)
        print(json.dumps(line), flush=True)
    if dist is not None:
        timer.barrier()
        llo.comm_destroy()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--log2n", type=int, default=28, help="elements of the chain (2^k, k even)")
    ap.add_argument("--cpu-log2n", type=int, default=22, help="elements of the CPU-baseline sample")
    ap.add_argument("--cpu-repeats", type=int, default=3)
    ap.add_argument("--ref-log2n", type=int, default=20, help="elements per step of the reference arm")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--matmul", type=int, default=8192)
    ap.add_argument("--mlp-batch", type=int, default=65536)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-side", action="store_true", help="skip the reduce / matmul workloads")
    ap.add_argument("--skip-mlp", action="store_true")
    ap.add_argument("--mlp-only", action="store_true",
                    help="time only the batch-sharded MLP (experiments; prints its block, not the contract line)")
    ap.add_argument("--skip-verify", action="store_true", help="do not compare the timed outputs with float64 numpy")
    ap.add_argument("--no-traffic-probe", action="store_true", help="do not run the ncu DRAM-bytes pass")
    ap.add_argument("--cpu-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-steps", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-warmup", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--ref-single", action="store_true",
                    help="reference arm on one thread only (default: one evaluation per host core)")
    args = ap.parse_args()
    if args.cpu_worker:
        cpu_worker(args)
        return
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_own(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
