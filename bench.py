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
(the single-operator configs do not shard: "replicas only", weak scaling).

The other BASELINE.json configurations are reported under "workloads":
reductions of [D, D] (config 3), matmul 8192^3 fp32 (config 4) and the
batch-sharded MLP gradient evaluation with its NCCL allreduce (config 5, batch
65536 split over the N ranks: strong scaling of that workload).

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
        return float(data["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ffma_peak():
    """fp32 FFMA issue peak measured by tools/probes/fma_peak.cu on this pool"""
    path = os.path.join(ROOT, "profiles", "fma_peak.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["ffma_tflops"]), "measured (profiles/fma_peak.json)"
    return 148 * 128 * 2 * 1.965e9 / 1e12, "nominal 148 SM x 128 lanes x 2 x 1965 MHz"


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


def cpu_config1(age, llo):
    """BASELINE.json configs[0]: the reference's CPU-runnable unit graphs over float
    [64,64] through the oracle (single thread): ns per output element, median of 10"""
    oracle = oracle_module()
    rng = np.random.default_rng(1)
    x = rng.uniform(0.5, 1.5, (64, 64)).astype(np.float32)
    y = rng.uniform(0.5, 1.5, (64, 64)).astype(np.float32)
    vx, vy = llo.variable(x, "x"), llo.variable(y, "y")
    graphs = {"add": age.add(vx, vy), "mul": age.mul(vx, vy), "exp": age.exp(vx),
              "reduce_sum_dim1": age.reduce_sum(vx, 1), "reduce_sum_all": age.reduce_sum0(vx),
              "matmul": age.matmul(vx, vy)}
    out = {}
    for name, root in graphs.items():
        times = []
        for _ in range(10):
            t0 = time.perf_counter()
            oracle.evaluate(root, F32)
            times.append(time.perf_counter() - t0)
        sec = float(np.median(times))
        entry = {"us": sec * 1e6, "ns_per_input_elem": sec * 1e9 / 4096}
        if name == "matmul":
            entry["MFLOPs"] = 2.0 * 64 ** 3 / sec / 1e6
        out[name] = entry
    out["note"] = "oracle = CPU restatement of llo::Evaluator, 1 thread, ade shape [64,64] fp32"
    return out


def run_reference(args, rank, world):
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
    per_step = (time.perf_counter() - t0) / args.steps
    value = workloads.chain_bytes(side) / per_step / 1e9
    sample = ("the same chain on x=[%d,%d] (2^%d elements) per step; the reference's evaluator "
              "is single-threaded by construction (no threads, SIMD or BLAS: SURVEY.md 2.2), so "
              "one core is all it can use" % (side, side, args.ref_log2n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "elementwise chain add(mul(exp(x),extend(b,1,{D})),permute(x,{1,0})), fp32, "
                               "bounded CPU sample of x=[16384,16384]", "sample_elements": side * side,
                   "full_elements": 1 << args.log2n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
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

    def timed(self, fn, steps, warmup, finish=None):
        """ms per step of `fn`, max over ranks; also returns the launches.
        `finish` (pipelined callers) completes whatever the last step left in
        flight; it runs inside the timed region and after the warm-up"""
        for _ in range(warmup):
            fn()
        if finish is not None:
            finish()
        self.barrier()
        launches0 = self.llo.launch_count()
        self.ctx.record(self.start)
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        self.ctx.record(self.stop)
        ms = self.ctx.elapsed_ms(self.start, self.stop)
        self.barrier()
        launches = self.llo.launch_count() - launches0
        return self.max_over_ranks(ms) / steps, launches


def side_workloads(args, timer, age, llo, rank, world):
    """BASELINE.json configs 3-5 (reported beside the headline)"""
    out = {}
    hbm_peak, _ = measured_peaks()
    steps, warm = max(3, min(args.steps, 10)), 3
    if world == 1 and not args.skip_side:
        # config 3: reductions of [D, D], deterministic tree order
        side = 1 << (args.log2n // 2)
        x = np.random.default_rng(4).uniform(0.0, 1.0, (side, side)).astype(np.float32)
        vx = llo.variable(x, "x")
        nbytes = 4 * side * side
        cases = {
            "reduce_sum_cols": age.reduce_sum(vx, 1),                      # keeps the fastest dim
            "reduce_max_cols": age.reduce_max(vx, 1),
            "reduce_sum_all": age.reduce_sum0(vx),
            "reduce_max_all": age.reduce_max0(vx),
            "reduce_sum_rows": age.reduce_sum(age.permute(vx, [1, 0]), 1),  # contiguous rows, permute fused
            "reduce_max_rows": age.reduce_max(age.permute(vx, [1, 0]), 1),
        }
        for name, root in cases.items():
            ms, launches = timer.timed(lambda r=root: llo.evaluate_device(r, F32), steps, warm)
            gbps = nbytes / ms / 1e6
            out[name] = {"GBps": gbps, "ms": ms, "frac_of_measured_hbm": gbps / hbm_peak,
                         "launches_per_eval": launches // steps}
        del cases
        # config 3, reference-representable shape: [128,128,128,128], reduce_*(x, 2)
        x4 = x.reshape(128, 128, side * side // (128 * 128 * 128), 128) if side * side >= 128 ** 4 else None
        if x4 is not None:
            vx4 = llo.variable(x4, "x4")
            for name, root in (("reduce_sum_128x4_dim2", age.reduce_sum(vx4, 2)),
                               ("reduce_max_128x4_dim2", age.reduce_max(vx4, 2))):
                ms, launches = timer.timed(lambda r=root: llo.evaluate_device(r, F32), steps, warm)
                gbps = nbytes / ms / 1e6
                out[name] = {"GBps": gbps, "ms": ms, "frac_of_measured_hbm": gbps / hbm_peak,
                             "launches_per_eval": launches // steps}
            del vx4
        # config 2, every operator stand-alone: algorithmic bytes =
        # 4 * (distinct input elements + output elements)
        y = np.random.default_rng(5).uniform(0.5, 2.0, (side, side)).astype(np.float32)
        bb = np.random.default_rng(3).uniform(0.5, 2.0, (side,)).astype(np.float32)
        vy, vbb = llo.variable(y, "y"), llo.variable(bb, "b")
        n = side * side
        ops = {
            "exp": (age.exp(vx), 2 * n), "sin": (age.sin(vx), 2 * n), "sqrt": (age.sqrt(vx), 2 * n),
            "abs": (age.abs(vx), 2 * n), "neg": (age.neg(vx), 2 * n), "log": (age.log(vx), 2 * n),
            "add": (age.add(vx, vy), 3 * n), "sub": (age.sub(vx, vy), 3 * n),
            "mul": (age.mul(vx, vy), 3 * n), "div": (age.div(vx, vy), 3 * n),
            "pow": (age.pow(vx, vy), 3 * n), "lt": (age.lt(vx, vy), 3 * n), "eq": (age.eq(vx, vy), 3 * n),
            "extend": (age.extend(vbb, 1, [side]), n + side),
            "mul_extend": (age.mul(vx, age.extend(vbb, 1, [side])), 2 * n + side),
            "permute": (age.permute(vx, [1, 0]), 2 * n),
        }
        elem = {}
        for name, (root, elems) in ops.items():
            ms, launches = timer.timed(lambda r=root: llo.evaluate_device(r, F32), steps, warm)
            gbps = 4.0 * elems / ms / 1e6
            elem[name] = {"GBps": round(gbps, 1), "ms": round(ms, 4), "frac_of_measured_hbm": round(gbps / hbm_peak, 3)}
        out["elementwise_ops"] = elem
        del ops, vy, vbb, y
        # config 2, reference-representable shape: the chain over [128,128,128,128]
        if x4 is not None:
            vx4 = llo.variable(x4, "x4")
            b4 = llo.variable(bb[:128].copy(), "b4")
            rest = list(x4.shape[:-1])[::-1]
            root = age.add(age.mul(age.exp(vx4), age.extend(b4, 1, rest)), age.permute(vx4, [1, 0, 2, 3]))
            ms, launches = timer.timed(lambda: llo.evaluate_device(root, F32), steps, warm)
            gbps = 4.0 * (2 * n + 128) / ms / 1e6
            out["chain_128x4"] = {"GBps": gbps, "ms": ms, "frac_of_measured_hbm": gbps / hbm_peak,
                                  "launches_per_eval": launches // steps}
            del vx4, b4, root
        del vx, x, x4, bb
        # config 4: matmul 8192^3 — fp32 on the FFMA pipe (exact path), fp64 on
        # the DFMA pipe, and the opt-in 3xTF32 tensor-core path (tcgen05)
        m = args.matmul
        gsteps = max(3, steps // 2)
        fpeak, fsrc = ffma_peak()
        dpeak = fpeak / 2.0
        ppath = os.path.join(ROOT, "profiles", "fma_peak.json")
        if os.path.exists(ppath):
            with open(ppath) as fh:
                dpeak = float(json.load(fh).get("dfma_tflops", dpeak))
        for name, dt, option in (("matmul_f32", np.float32, 0), ("matmul_f32_3xtf32", np.float32, 1),
                                 ("matmul_f64", np.float64, 0)):
            a = np.random.default_rng(5).uniform(-1, 1, (m, m)).astype(dt)
            b = np.random.default_rng(6).uniform(-1, 1, (m, m)).astype(dt)
            va, vb = llo.variable(a, "a"), llo.variable(b, "b")
            root = age.matmul(va, vb)
            llo.set_option("tf32x3", option)
            try:
                ms, launches = timer.timed(lambda: llo.evaluate_device(root, np.dtype(dt)), gsteps, 2)
            finally:
                llo.set_option("tf32x3", 0)
            tf = 2.0 * m ** 3 / ms / 1e9
            entry = {"TFLOPs": tf, "ms": ms, "mnk": [m, m, m], "launches_per_eval": launches // gsteps}
            if option:
                entry["path"] = ("opt-in: tcgen05.mma kind::tf32, 3 products per k, TMEM accumulators promoted "
                                 "every 128 k (results within 1e-6 of sum|a||b|, not bit-identical to FFMA)")
                entry["speedup_vs_exact"] = tf / out["matmul_f32"]["TFLOPs"]
            elif dt == np.float32:
                entry.update({"frac_of_ffma_peak": tf / fpeak, "ffma_peak_TFLOPs": fpeak, "peak_source": fsrc})
            else:
                entry.update({"frac_of_dfma_peak": tf / dpeak, "dfma_peak_TFLOPs": dpeak})
            out[name] = entry
            del va, vb, a, b, root
    if not args.skip_mlp:
        # config 5: MLP gradient evaluation, batch sharded over the ranks
        from cortenn_b200.sharded import BatchShardedExecutor
        batch = args.mlp_batch
        X, Y, params = workloads.mlp_data(batch)
        ex = BatchShardedExecutor(batch, rank=rank, world=world)
        loss, pvars, _ = workloads.mlp_graph(age, llo, ex.shard(X), ex.shard(Y), params, global_batch=batch)
        del X, Y

        def grad_eval():
            ex.grad_eval_device(loss, pvars, F32)
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
        peak, _ = ffma_peak()
        llo.set_option("tf32x3", 1)
        try:
            ms_tc, _ = timer.timed(grad_eval, steps, warm)
        finally:
            llo.set_option("tf32x3", 0)
        out["mlp_grad_eval_3xtf32"] = {"grad_evals_per_s": 1e3 / ms_tc, "ms": ms_tc,
                                       "path": "opt-in tensor-core GEMMs (tf32x3 = 1); same graph, same allreduce"}
        out["mlp_grad_eval"] = {
            "grad_evals_per_s": 1e3 / ms, "ms": ms, "batch": batch, "rows_per_rank": ex.local_rows(),
            "gemm_TFLOPs_all_ranks": flops / ms / 1e9,
            "frac_of_ffma_peak": flops / ms / 1e9 / (peak * world),
            "allreduce_bytes": 4 * 1863690, "launches_per_eval": launches // steps,
            "host_enqueue_ms": host_ms,
            "scaling": "strong (batch 65536 split over the ranks; NCCL allreduce of the packed gradients)"}
    return out


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
    timer = Timer(llo, capi, dist, torch)

    side = 1 << (args.log2n // 2)
    x, b = workloads.chain_inputs(side, seed_x=2 + rank)
    vx, vb, root = workloads.chain_graph(age, llo, x, b)
    nbytes = workloads.chain_bytes(side)
    warmup = max(args.warmup, 3)

    # ---- value: inputs resident in HBM, device-timed --------------------------
    for _ in range(warmup):
        llo.evaluate_device(root, F32)
    timer.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, launches = timer.timed(lambda: llo.evaluate_device(root, F32), args.steps, 0)
    clocks = sampler.stop() if rank == 0 else {}
    value = world * nbytes / ms_step / 1e6

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
           "path": "Variable.assign_async(pinned ndarray) + llo.evaluate_async -> .result() ndarray, every "
                   "step; upload of step i+1 overlaps the download of step i",
           "pcie_GBps_each_way": (x.nbytes + out.nbytes) / 2 / e2e_ms / 1e6,
           "synchronous": {"value": world * nbytes / sync_ms / 1e6, "ms_per_step": sync_ms,
                           "path": "Variable.assign + llo.evaluate, copies back to back"}}
    del out, holder, xs

    peak, peak_src = measured_peaks()
    per_eval = max(1, launches // args.steps)
    roofline = {"bound": "hbm", "achieved": nbytes / ms_step / 1e6, "peak": peak, "unit": "GB/s",
                "frac": nbytes / ms_step / 1e6 / peak, "traffic": None, "peak_source": peak_src,
                "frac_of_nominal_8TBs": nbytes / ms_step / 1e6 / 8000.0,
                "kernel": "vm_pair4_kernel<float> (the whole step is %d launch per evaluation)" % per_eval,
                "algorithmic_bytes_per_launch": nbytes}
    traffic_path = os.path.join(ROOT, "profiles", "chain_traffic.json")
    if os.path.exists(traffic_path):
        with open(traffic_path) as fh:
            roofline["traffic"] = json.load(fh).get("dram_bytes_per_launch")

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cside = 1 << (args.cpu_log2n // 2)
        sec, gbps = cpu_chain(cside, args.cpu_repeats)
        cpu_baseline = {"value": gbps, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": "oracle (CPU restatement of llo::Evaluator) on the same chain at "
                                  "x=[%d,%d] (2^%d of the 2^%d elements), %d evaluations of %.2f s, "
                                  "single thread like the reference; host has %d cores"
                                  % (cside, cside, args.cpu_log2n, args.log2n, args.cpu_repeats, sec,
                                     os.cpu_count() or 0)}

    del vx, vb, root, x, b
    side_results = side_workloads(args, timer, age, llo, rank, world)
    if cpu_baseline is not None:
        side_results["cpu_config1"] = cpu_config1(age, llo)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "elementwise chain add(mul(exp(x),extend(b,1,{D})),permute(x,{1,0})), "
                                   "x=[%d,%d] fp32 (2^%d elements), ade graph via age API -> llo.evaluate_device"
                                   % (side, side, args.log2n),
                       "bytes_per_step": nbytes,
                       "l2": "inputs larger than L2 (x is %d MiB, L2 is 126 MB)" % ((4 * side * side) >> 20),
                       "multi_gpu": "independent replicas per rank (single-operator configs do not shard); "
                                    "the batch-sharded MLP is under workloads.mlp_grad_eval"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "workloads": side_results,
        }
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
    ap.add_argument("--cpu-repeats", type=int, default=10)
    ap.add_argument("--ref-log2n", type=int, default=20, help="elements per step of the reference arm")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--matmul", type=int, default=8192)
    ap.add_argument("--mlp-batch", type=int, default=65536)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-side", action="store_true", help="skip the reduce / matmul workloads")
    ap.add_argument("--skip-mlp", action="store_true")
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
