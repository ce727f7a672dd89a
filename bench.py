#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched ODE-solve hot path (BASELINE.json configs[1]).

Workload: README Michaelis-Menten model (states [E,P,S], parameters [K_m,k_cat]), 2^20
initial-condition x parameter sets PER GPU, Tsit5, rtol=atol=1e-6, dt0=0.1, max_steps=4096,
saved on linspace(0,100,T) with T=1000 ("same" grid as configs[0]).  One "step" = one pass of
the hot path (ctx_solve) over the whole batch.  Multi-GPU: rows are sharded by rank, no
data-path collective (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f32|f64] [--T 1000]
  python bench.py --impl reference ...   # the reference ALGORITHM on the host cores (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np

METRIC = "ODE trajectory-steps/s"
UNIT = "steps/s"
ROWS_PER_GPU = 1 << 20
N_INPUT_SETS = 8          # rotate input sets so inputs (8 x 20 MB) are not L2-resident either
CPU_SAMPLE_ROWS = 1 << 18


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU)
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nuts", action="store_true")
    ap.add_argument("--no-other", action="store_true")
    ap.add_argument("--nuts-chains", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"michaelis_menten_ode[E,P,S]_B{a.rows}_per_gpu_tsit5_rtol1e-6_atol1e-6_dt0.1_"
            f"T{a.T}_linspace(0,100)")


def np_dtype(a):
    return np.float32 if a.dtype == "f32" else np.float64


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_oracle_run(a, rows, n_threads=None, repeats=1):
    """Time the oracle's C port (the reference ALGORITHM) on host cores.  Returns steps/s."""
    from oracle import c_oracle
    from tests import cases

    dt = np_dtype(a)
    m = cases.michaelis_menten()
    co = c_oracle.COracle(*m[:4], m[4], dtype=dt)
    y0, th, c = cases.mm_inputs(rows, seed=0, dtype=dt)
    ts = np.linspace(0, 100, a.T).astype(dt)
    n_threads = n_threads or os.cpu_count() or 1
    best = None
    steps = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = co.solve(y0, th, c, ts, rtol=1e-6, atol=1e-6, dt0=0.1, max_steps=4096,
                     n_threads=n_threads)
        dt_s = time.perf_counter() - t0
        steps = int((r.n_accepted.astype(np.int64) + r.n_rejected).sum())
        best = dt_s if best is None else min(best, dt_s)
    return steps / best, best, steps, n_threads


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = min(CPU_SAMPLE_ROWS, a.rows)
    from oracle import c_oracle  # noqa: F401  (compiles on first use)

    for _ in range(max(a.warmup, 0)):
        cpu_oracle_run(a, min(rows, 1 << 14))
    t_total, steps_total, threads = 0.0, 0, os.cpu_count() or 1
    for _ in range(a.steps):
        _, sec, steps, threads = cpu_oracle_run(a, rows)
        t_total += sec
        steps_total += steps
    value = steps_total / t_total
    sample = (f"first {rows} rows of the {a.rows}-row batch per step, C/OpenMP port of the "
              f"reference algorithm (oracle/ctx_oracle_core.h), {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_total / max(a.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype,
        "data": "synthetic",
        "config": {"workload": workload_name(a), "rows_per_step": rows},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- NUTS leg
def nuts_inputs(chains, M, T, dt, seed):
    """BASELINE configs[3]: chains x 64 series x 20 points x 3 observables; data = trajectories at
    theta* + noise; per-chain theta ~ theta* LogN(0, 0.2), sigma ~ LogN(log 0.5, 0.3).
    theta* = (K_m=40, k_cat=0.3): with the README values the discrete tangents of the reference
    algorithm blow up after substrate depletion (tests/test_gpu_parity_loglik.py)."""
    import torch

    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from tests import cases

    rng = np.random.default_rng(seed)
    y0s = np.stack([rng.uniform(5, 20, M), np.zeros(M), rng.uniform(50, 300, M)], axis=1)
    ts = np.tile(np.linspace(0, 100, T), (M, 1))
    theta_star = np.array([40.0, 0.3])
    # noise-free trajectories at theta*: the engine itself in x64 at a tight tolerance (the oracle
    # is test infrastructure and stays out of the GPU arm)
    gen = engine.CompiledModel(cg.generate_field(cases.field_spec(cases.michaelis_menten())), "float64")
    dev = torch.device("cuda", torch.cuda.current_device())
    tt64 = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
    clean = gen.solve(tt64(y0s), tt64(theta_star), tt64(np.zeros((M, 0))), tt64(ts),
                      in_axes=(0, None, 0, 0), rtol=1e-8, atol=1e-8).ys.cpu().numpy()
    data = clean + rng.normal(size=clean.shape) * (0.05 * np.abs(clean) + 1e-3)
    theta = theta_star * np.exp(rng.normal(0, 0.2, (chains, 2)))
    sigma = np.exp(rng.normal(np.log(0.5), 0.3, (chains, 3)))
    c = lambda x: np.ascontiguousarray(x, dtype=dt)
    return c(y0s), c(theta), np.zeros((M, 0), dtype=dt), c(ts), c(data), c(sigma)


def run_nuts_leg(a, model_field, dev, rank, world, dist, fma_peak=None):
    """Chain-sharded (no collective): every rank evaluates its own `nuts_chains` chains over all
    64 series.  value = chain-level (log-density, gradient) evaluations per second, all ranks."""
    import torch

    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from tests import cases

    dt = np_dtype(a)
    CH, M, T = a.nuts_chains, 64, 20
    field = cg.generate_field(cases.field_spec(cases.michaelis_menten()), sens_theta=True)
    model = engine.CompiledModel(field, np.dtype(dt).name)
    y0s, theta, consts, ts, data, sigma = nuts_inputs(CH, M, T, dt, seed=2 + 100 * rank)
    tt = lambda x: torch.as_tensor(x, device=dev)
    args = [tt(x) for x in (y0s, theta, consts, ts, data, sigma)]
    kw = dict(likelihood="softlaplace", rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096,
              kernel=int(os.environ.get("CTX_BENCH_LL_KERNEL", "0")))
    n_iter = max(3, min(a.steps, 20))
    for _ in range(3):
        out, partial, status = model.loglik_grad(*args, **kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_iter):
        out, partial, status = model.loglik_grad(*args, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    fails = int((status != 0).sum().item())
    finite = bool(torch.isfinite(out).all().item())
    # compute roofline of K2 (rank 0's numbers): algorithmic flops = sympy count_ops of the
    # CSE'd field + tangent field, Tsit5 stage / error / controller arithmetic, dense output and
    # likelihood terms vs the MEASURED FFMA/DFMA peak of this GPU.
    roof = None
    if rank == 0:
        S, D = field.S, field.D
        N = S * (1 + D)
        F = field.flops_rhs + field.flops_tan
        nacc, nrej = model.last_loglik_counts
        steps = int((nacc.sum(dtype=torch.int64) + nrej.sum(dtype=torch.int64)).item())
        flops_step = 6 * F + 56 * N + 8 * S + 11
        flops_point = 40 + 14 * N + 25 * S
        flops = steps * flops_step + CH * M * T * flops_point
        peak = fma_peak or engine.measure_fma_peak(np.dtype(dt).name)
        achieved = flops / (ms / n_iter * 1e-3) / 1e12
        roof = {"bound": "fp32" if a.dtype == "f32" else "fp64", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "ctx_measure_fma_peak (FMA loop measured on this GPU)",
                "flops_per_step": flops_step, "flops_per_saved_point": flops_point,
                "steps_per_eval_batch": steps}
    # series-sharded form of the SAME posterior (SURVEY 8d config 4 asks for both): all ranks
    # evaluate the same CH chains (rank 0's parameters) over M/world series each, followed by the
    # path's only collective, an NCCL all-reduce(sum) of the [CH, 1+P+S] block.
    series = None
    if world > 1 and M % world == 0:
        y0s0, theta0, consts0, ts0, data0, sigma0 = nuts_inputs(CH, M, T, dt, seed=2)
        lo, hi = rank * (M // world), (rank + 1) * (M // world)
        sargs = [tt(x) for x in (y0s0[lo:hi], theta0, consts0[lo:hi], ts0[lo:hi], data0[lo:hi], sigma0)]

        def series_eval():
            o, _, st = model.loglik_grad(*sargs, **kw)
            dist.all_reduce(o, op=dist.ReduceOp.SUM)
            return o, st

        for _ in range(3):
            series_eval()
        torch.cuda.synchronize()
        dist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n_iter):
            so, sst = series_eval()
        s1.record()
        torch.cuda.synchronize()
        st_ms = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        dist.all_reduce(st_ms, op=dist.ReduceOp.MAX)
        sms = float(st_ms.item())
        series = {"value": CH * n_iter / (sms * 1e-3), "unit": "evals/s",
                  "ms_per_eval_batch": sms / n_iter, "series_per_rank": M // world,
                  "collective": f"NCCL all-reduce(sum) of [{CH}, {1 + 2 + 3}] {a.dtype} per evaluation",
                  "scaling": "strong (one posterior, series split over ranks)",
                  "all_finite": bool(torch.isfinite(so).all().item())}
    return {"metric": "NUTS logp+grad evals/s", "value": CH * world * n_iter / (ms * 1e-3),
            "series_sharded": series,
            "roofline": roof,
            "unit": "evals/s", "ms_per_eval_batch": ms / n_iter,
            "series_solves_with_gradient_per_s": CH * world * M * n_iter / (ms * 1e-3),
            "config": {"chains_per_gpu": CH, "series": M, "points": T, "observables": 3,
                       "likelihood": "SoftLaplace", "rtol": 1e-5, "atol": 1e-5,
                       "sharding": "by chain (no collective)"},
            "failed_series": fails, "all_finite": finite}


# ----------------------------------------------------------------------------- other configs
def run_other_configs(a, model, sets, dev, rank, world, dist, fma_peak=None):
    """Short device-timed legs for the other BASELINE configs (rank 0's GPU only, reported beside
    the headline, never part of `value`): configs[1] in the integrator-bound regime (T=16),
    configs[2] (32-state mass-action network, warp-per-trajectory kernel) and configs[4]
    (NeuralODE field 5-64-64-64-4 on the tcgen05 tensor cores)."""
    import torch

    from catalax_b200 import engine
    from catalax_b200.neural import NeuralODE
    from catalax_b200.tools import codegen as cg
    from tests import cases

    if rank != 0:
        return None
    dt = np_dtype(a)
    tdt = torch.float32 if a.dtype == "f32" else torch.float64

    def timed(fn, n=5):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    res = {}
    # configs[1], T=16: the output is negligible, the integrator is the bound
    y0, th, c = sets[0]
    ts16 = torch.linspace(0, 100, 16, dtype=tdt, device=dev)
    out16 = torch.empty((y0.shape[0], 16, 3), dtype=tdt, device=dev)
    ms, o = timed(lambda: model.solve(y0, th, c, ts16, solver="tsit5", in_axes=(0, 0, 0, None),
                                      rtol=1e-6, atol=1e-6, dt0=0.1, max_steps=4096, out=out16))
    steps = int((o.n_accepted.sum(dtype=torch.int64) + o.n_rejected.sum(dtype=torch.int64)).item())
    fl = steps * (6 * 6 + 64 * 3 + 11) / (ms * 1e-3) / 1e12
    res["config2_T16"] = {"ms": ms, "steps_per_s": steps / (ms * 1e-3), "rows": int(y0.shape[0]),
                          "roofline": {"bound": a.dtype.replace("f", "fp"), "achieved": fl,
                                       "unit": "TFLOP/s", "flops_per_step": 6 * 6 + 64 * 3 + 11}}
    # configs[2]: 32 states / 64 reactions, 262144 rows, T=32 on [0, 10]
    B3 = 262144
    f3 = cg.generate_field(cases.field_spec(cases.mass_action_network(32)))
    m3 = engine.CompiledModel(f3, np.dtype(dt).name)
    rng = np.random.default_rng(1)
    y3 = torch.as_tensor(rng.uniform(0.1, 2.0, (B3, 32)).astype(dt), device=dev)
    k3 = torch.as_tensor(np.exp(rng.uniform(np.log(0.01), np.log(1.0), (B3, 64))).astype(dt), device=dev)
    c3 = torch.empty((B3, 0), dtype=tdt, device=dev)
    ts3 = torch.linspace(0, 10, 32, dtype=tdt, device=dev)
    out3 = torch.empty((B3, 32, 32), dtype=tdt, device=dev)    # (preallocated: no allocator traffic
    ms, o = timed(lambda: m3.solve(y3, k3, c3, ts3, in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6,
                                   out=out3))                   #  inside the timed region)
    steps = int((o.n_accepted.sum(dtype=torch.int64) + o.n_rejected.sum(dtype=torch.int64)).item())
    fps = 6 * f3.flops_rhs + 64 * 32 + 11
    fl = steps * fps / (ms * 1e-3) / 1e12
    res["config3_mass_action_32x64"] = {
        "ms": ms, "steps_per_s": steps / (ms * 1e-3), "rows": B3, "kernel": "ctx_solve_w (warp per trajectory)",
        "failed_rows": int((o.status != 0).sum().item()),
        "roofline": {"bound": a.dtype.replace("f", "fp"), "achieved": fl, "unit": "TFLOP/s",
                     "flops_per_step": fps}}
    # configs[4]: NeuralODE 5-64-64-64-4 (tanh), 524288 rows, T=16 on [0, 10], rtol 1e-3 (f32 only:
    # the tensor-core form; f64 runs the FFMA form)
    if a.dtype == "f32":
        B5 = 1 << 19
        net = NeuralODE(4, 64, 3, ["a", "b", "c", "d"], activation="tanh", max_time=10.0, key=3)
        y5 = torch.as_tensor(np.random.default_rng(3).uniform(0, 2, (B5, 4)).astype(np.float32), device=dev)
        ts5 = torch.linspace(0, 10, 16, dtype=torch.float32, device=dev)
        ms, _ = timed(lambda: net.predict_arrays(ts5, y5), n=3)
        o = net.last
        steps = int((o.n_accepted.sum(dtype=torch.int64) + o.n_rejected.sum(dtype=torch.int64)).item())
        mlp_flops = steps * 6 * 17536            # 6 stage evaluations per step attempt, 8768 MAC each
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        tpeak = float(peaks.get("bf16_tflops_sustained", 1384.0)) / 2.0   # tf32 dense = bf16 / 2
        ach = 3.0 * mlp_flops / (ms * 1e-3) / 1e12   # 3xTF32 split: three MMAs per useful product
        res["config5_neuralode_5x64x64x64x4"] = {
            "ms": ms, "steps_per_s": steps / (ms * 1e-3), "rows": B5,
            "kernel": "ctx_solve_mlp_tc (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators)",
            "useful_mlp_tflops": mlp_flops / (ms * 1e-3) / 1e12,
            "failed_rows": int((o.status != 0).sum().item()),
            "roofline": {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s",
                         "frac": ach / tpeak,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32)"}}
    peak = fma_peak or engine.measure_fma_peak(np.dtype(dt).name)
    for key in ("config2_T16", "config3_mass_action_32x64"):
        r = res[key]["roofline"]
        r["peak"] = peak
        r["frac"] = r["achieved"] / peak
        r["peak_source"] = "ctx_measure_fma_peak (FMA loop measured on this GPU)"
    return res


# ----------------------------------------------------------------------------- GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist

    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from tests import cases

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {a.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: catalax_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    dt = np_dtype(a)
    tdt = torch.float32 if a.dtype == "f32" else torch.float64
    w = 4 if a.dtype == "f32" else 8
    B, T, S, P = a.rows, a.T, 3, 2
    field = cg.generate_field(cases.field_spec(cases.michaelis_menten()))
    model = engine.CompiledModel(field, np.dtype(dt).name)
    # FMA peak (denominator of the compute rooflines of the side legs) first: the micro-benchmark
    # loads the GPU heavily, nothing timed may run in its wake (input generation + warm-up follow)
    need_peak = rank == 0 and not (a.no_nuts and (a.no_other or world > 1))
    fma_peak = engine.measure_fma_peak(np.dtype(dt).name) if need_peak else None

    # rows are sharded by rank: every rank draws its own seeds (no data-path collective)
    sets = []
    for k in range(N_INPUT_SETS):
        y0, th, c = cases.mm_inputs(B, seed=1000 * rank + k, dtype=dt)
        sets.append(tuple(torch.as_tensor(x, device=dev) for x in (y0, th, c)))
    ts = torch.linspace(0, 100, T, dtype=tdt, device=dev)
    out = torch.empty((B, T, S), dtype=tdt, device=dev)
    kw = dict(solver="tsit5", in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, dt0=0.1,
              max_steps=4096, kernel=a.kernel, out=out)

    def step(i):
        y0, th, c = sets[i % N_INPUT_SETS]
        return model.solve(y0, th, c, ts, **kw)

    # steps per input set (deterministic) -- counted outside the timed region
    steps_of_set, fails = [], 0
    for k in range(N_INPUT_SETS):
        o = step(k)
        steps_of_set.append(int((o.n_accepted.sum(dtype=torch.int64) + o.n_rejected.sum(dtype=torch.int64)).item()))
        fails += int((o.status != 0).sum().item())
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()           # nvidia-smi needs ~100 ms to start: begin before the warm-up
    t_w = time.perf_counter()
    i = 0
    while i < max(a.warmup, 3) or time.perf_counter() - t_w < 0.3:   # >= 3 steps and >= 0.3 s
        step(i)
        i += 1
        if i % 8 == 0:
            torch.cuda.synchronize()
    n_warm = i
    torch.cuda.synchronize()
    if sampler:
        sampler.lines.clear()      # keep only samples taken from here on (timed region)
    if world > 1:
        dist.barrier()
    launches0 = engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = engine.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    my_steps = sum(steps_of_set[i % N_INPUT_SETS] for i in range(a.steps))

    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    tst = torch.tensor([float(my_steps), float(fails)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tst, op=dist.ReduceOp.SUM)
    ms_max, total_steps, total_fails = float(tms.item()), float(tst[0].item()), int(tst[1].item())
    value = total_steps / (ms_max * 1e-3)

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (copies inside the timed region)
    e2e = None
    if not a.no_e2e:
        import psutil

        out_bytes = B * T * S * w
        e2e_rows = B
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        while e2e_rows > 4096 and (e2e_rows * T * S * w) * local_world * 3 > psutil.virtual_memory().available:
            e2e_rows //= 2
        y0h, thh, ch = cases.mm_inputs(e2e_rows, seed=1000 * rank, dtype=dt)
        pin = lambda arr: torch.as_tensor(arr).pin_memory()
        y0p, thp = pin(y0h), pin(thh)
        chp = torch.empty((e2e_rows, 0), dtype=tdt).pin_memory()
        tsp = torch.linspace(0, 100, T, dtype=tdt).pin_memory()
        ysp = torch.empty((e2e_rows, T, S), dtype=tdt).pin_memory()
        n_e2e = max(1, min(a.steps, 3))
        run = lambda: model.solve_host(y0p.numpy(), thp.numpy(), chp.numpy(), tsp.numpy(),
                                       in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, dt0=0.1,
                                       max_steps=4096, kernel=a.kernel, out=ysp.numpy())
        o = run()  # warm-up (allocates device staging)
        e2e_steps = int(o.n_accepted.astype(np.int64).sum() + o.n_rejected.astype(np.int64).sum())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            run()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        tt = torch.tensor([sec], dtype=torch.float64, device=dev)
        ts_ = torch.tensor([float(e2e_steps * n_e2e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ts_, op=dist.ReduceOp.SUM)
        h2d = e2e_rows * (S + P) * w + T * w
        d2h = e2e_rows * T * S * w + 3 * 4 * e2e_rows
        e2e = {"value": float(ts_.item()) / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "rows_per_step": e2e_rows,
               "steps_timed": n_e2e, "api": "ctx_solve_host (C ABI, pinned host buffers)"}
        del ysp

    # ---- second headline metric: NUTS log-density + gradient evaluations / s (configs[3])
    nuts = None
    if not a.no_nuts:
        nuts = run_nuts_leg(a, model_field=None, dev=dev, rank=rank, world=world, dist=dist,
                            fma_peak=fma_peak)

    other = None
    if not a.no_other and world == 1:
        try:
            other = run_other_configs(a, model, sets, dev, rank, world, dist, fma_peak=fma_peak)
        except Exception as exc:       # never lose the headline line to a side leg
            other = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    algo_bytes = B * (T * S * w + (S + P) * w + 12) + T * w     # per launch (per rank)
    launch_ms = ms / a.steps                                     # rank 0's own average
    achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
    prof = ROOT / "profiles" / f"r1_solve_{a.dtype}_T{T}.json"
    traffic = None
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # informational: a pure write stream (memset of the same 12.6 GB) measured on this pool's B200s
    write_only = None
    wo = ROOT / "profiles" / "r1q_write_only_bw.json"
    if wo.exists():
        try:
            write_only = json.loads(wo.read_text())["memset_12.6GB"]["GBps"]
        except Exception:
            write_only = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "warmup_done": n_warm, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "rows_total": B * world, "saved_points": T,
                   "l2": f"outputs {B*T*S*w/1e9:.1f} GB/step >> 126 MB L2; inputs rotate over "
                         f"{N_INPUT_SETS} sets",
                   "kernel": "ctx_solve_v2 (producer warps step with lane refill, consumer warps evaluate "
                             "the dense output: 8-point groups, 256-bit stores)",
                   "failed_rows": total_fails, "steps_per_row": total_steps / (B * world * a.steps)},
        "gpu_launches": launches, "clocks": clocks,
        "saved_points_per_s": B * world * T * a.steps / (ms_max * 1e-3),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "write_only_peak_gbs": write_only},
    }
    if e2e:
        line["e2e"] = e2e
    if nuts:
        line["nuts"] = nuts
    if other:
        line["other_configs"] = other
    if not a.no_cpu_baseline and world == 1:
        rows = B                       # the whole batch once: ~1.5 s on 16 cores (~20 core-seconds)
        v, sec, steps, threads = cpu_oracle_run(a, rows)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {rows} rows of the batch, one pass ({sec:.2f} s), C/OpenMP port of "
                      f"the reference algorithm (oracle/ctx_oracle_core.h)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
