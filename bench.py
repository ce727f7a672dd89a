#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched ODE-solve hot path (BASELINE.json configs[1]).

Workload: README Michaelis-Menten model (states [E,P,S], parameters [K_m,k_cat]), 2^20
initial-condition x parameter sets PER GPU, Tsit5, rtol=atol=1e-6, dt0=0.1, max_steps=4096,
saved on linspace(0,100,T) with T=1000 ("same" grid as configs[0]).  One "step" = one pass of
the hot path (ctx_solve) over the whole batch.  Multi-GPU: rows are sharded by rank, no
data-path collective (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f32|f64] [--T 1000]
  python bench.py --impl reference ...   # the reference ALGORITHM on the host cores (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.  The GPU arm imports
nothing from tests/ or oracle/ except inside the `cpu_baseline` legs (the oracle is the CPU
baseline, never the thing measured).
"""
from __future__ import annotations

import argparse
import json
import math
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
SIDE_MIN_ITERS = 20       # timed iterations of every side leg (after >= 0.3 s of warm-up)


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
    from catalax_b200 import workloads as wl

    lo, hi = wl.KM_RANGE_SOLVABLE
    return (f"michaelis_menten_ode[E,P,S]_B{a.rows}_per_gpu_tsit5_rtol1e-6_atol1e-6_dt0.1_"
            f"T{a.T}_linspace(0,100)_S0~U(50,300)_E0~U(5,20)_K_m~LogU({lo},{hi})_k_cat~LogU(0.02,0.5)")


def np_dtype(a):
    return np.float32 if a.dtype == "f32" else np.float64


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while a timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []          # (monotonic time, csv line)

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
            self.lines.append((time.monotonic(), line.strip()))

    def mark(self):
        return time.monotonic()

    def summary(self, t0=None, t1=None):
        """Median SM clock / reasons over the samples taken in [t0, t1] (all when None)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for ts, ln in list(self.lines):
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1 + 0.06):
                continue
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

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()


class Ctx:
    """What every leg needs: device, rank / world, the clock sampler, max-over-ranks timing."""

    def __init__(self, a, dev, rank, world, dist, sampler):
        self.a, self.dev, self.rank, self.world, self.dist, self.sampler = a, dev, rank, world, dist, sampler

    def max_over_ranks(self, ms):
        import torch

        if self.world == 1:
            return float(ms)
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, *vals):
        import torch

        if self.world == 1:
            return [float(v) for v in vals]
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def timed(self, fn, iters=SIDE_MIN_ITERS, warm_s=0.3, warm_min=3):
        """Device time of `iters` back-to-back calls (CUDA events on the launch stream), after
        >= warm_min calls and >= warm_s seconds of warm-up, bracketed by barriers; returns
        (ms per call: max over ranks, last result, clocks during the timed region)."""
        import torch

        # The number of calls must be the SAME on every rank (fn may contain a collective): warm_min
        # calls, one timed call to estimate the cost, then counts agreed upon by all ranks.
        out = None
        for _ in range(warm_min):
            out = fn()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        est = self.max_over_ranks(max(time.perf_counter() - t1, 1e-5))
        n_warm = int(min(2000, math.ceil(warm_s / est)))
        # long enough for the 50 ms clock sampler to see the leg: >= 0.25 s of timed work
        iters = int(max(iters, min(2000, math.ceil(0.25 / est))))
        for k in range(n_warm):
            out = fn()
            if k % 8 == 7:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        m0 = self.sampler.mark() if self.sampler else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        m1 = self.sampler.mark() if self.sampler else None
        if self.world > 1:
            self.dist.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1)) / iters
        clocks = self.sampler.summary(m0, m1) if self.sampler else None
        if clocks is not None:
            clocks["timed_iterations"] = iters
        return ms, out, clocks


def steps_of(o):
    import torch

    return int((o.n_accepted.sum(dtype=torch.int64) + o.n_rejected.sum(dtype=torch.int64)).item())


# ----------------------------------------------------------------------------- CPU arm
def cpu_oracle_run(a, rows, n_threads=None, repeats=1):
    """Time the oracle's C port (the reference ALGORITHM) on host cores.  Returns steps/s."""
    from catalax_b200 import workloads as wl
    from oracle import c_oracle

    dt = np_dtype(a)
    m = wl.michaelis_menten()
    co = c_oracle.COracle(*m[:4], m[4], dtype=dt)
    y0, th, c = wl.mm_inputs(rows, seed=0, dtype=dt)
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
              f"reference algorithm (oracle/ctx_oracle_core.h; NOT JAX: jax/diffrax are not "
              f"installable in this image), {threads} threads")
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
def nuts_inputs(chains, M, T, dt, seed, dev):
    """BASELINE configs[3]: chains x 64 series x 20 points x 3 observables (workloads.py).  The
    noise-free trajectories at theta* come from the engine itself in x64 at a tight tolerance
    (the oracle is test infrastructure and stays out of the GPU arm)."""
    import torch

    from catalax_b200 import engine, workloads as wl
    from catalax_b200.tools import codegen as cg

    gen = engine.CompiledModel(cg.generate_field(wl.field_spec(wl.michaelis_menten())), "float64")
    tt64 = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)

    def clean(y0s, theta_star, ts):
        return gen.solve(tt64(y0s), tt64(theta_star), tt64(np.zeros((len(y0s), 0))), tt64(ts),
                         in_axes=(0, None, 0, 0), rtol=1e-8, atol=1e-8).ys.cpu().numpy()

    return tuple(np.ascontiguousarray(x, dtype=dt) for x in wl.nuts_operands(chains, M, T, seed, clean))


def run_nuts_leg(cx, fma_peak=None):
    """Chain-sharded (no collective): every rank evaluates its own `nuts_chains` chains over all
    64 series.  value = chain-level (log-density, gradient) evaluations per second, all ranks."""
    import torch

    from catalax_b200 import engine, workloads as wl
    from catalax_b200.tools import codegen as cg

    a, dev, rank, world, dist = cx.a, cx.dev, cx.rank, cx.world, cx.dist
    dt = np_dtype(a)
    CH, M, T = a.nuts_chains, 64, 20
    field = cg.generate_field(wl.field_spec(wl.michaelis_menten()), sens_theta=True)
    model = engine.CompiledModel(field, np.dtype(dt).name)
    host = nuts_inputs(CH, M, T, dt, 2 + 100 * rank, dev)
    y0s, theta, consts, ts, data, sigma = host
    tt = lambda x: torch.as_tensor(x, device=dev)
    args = [tt(x) for x in host]
    base_kw = dict(rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096,
                   kernel=int(os.environ.get("CTX_BENCH_LL_KERNEL", "0")))
    n_iter = max(SIDE_MIN_ITERS, min(a.steps, 40))
    res = {}
    for lik in ("softlaplace", "normal"):
        kw = dict(base_kw, likelihood=lik)
        ms, (out, partial, status), clocks = cx.timed(lambda: model.loglik_grad(*args, **kw), iters=n_iter)
        res[lik] = dict(ms=ms, out=out, status=status, clocks=clocks,
                        counts=model.last_loglik_counts)
    # The sampler regime: NUTS evaluates the SAME posterior at slowly moving parameters, so the step
    # counts of the previous evaluation are a scheduling hint for the next one (BayesianModel keeps
    # such an order, refreshed every 16 evaluations).  Here the order comes from an evaluation at
    # PERTURBED parameters (each theta / sigma multiplied by LogN(0, 0.02): a leapfrog-sized move),
    # never from the timed inputs themselves; results are bit-identical with and without it.
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    pert = lambda x: x * torch.exp(0.02 * torch.randn(x.shape, generator=g, device=dev, dtype=x.dtype))
    hint_args = list(args)
    hint_args[1], hint_args[5] = pert(args[1]), pert(args[5])
    model.loglik_grad(*hint_args, **dict(base_kw, likelihood="softlaplace"))
    order = engine.cost_order(*model.last_loglik_counts, heavy_fraction=1.0)
    kw_h = dict(base_kw, likelihood="softlaplace", order=order)
    ms_h, (out_h, _, _), clocks_h = cx.timed(lambda: model.loglik_grad(*args, **kw_h), iters=n_iter)
    hinted_same_bits = bool(torch.equal(out_h, res["softlaplace"]["out"]))
    ms, out, status = res["softlaplace"]["ms"], res["softlaplace"]["out"], res["softlaplace"]["status"]
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
        nacc, nrej = res["softlaplace"]["counts"]
        steps = int((nacc.sum(dtype=torch.int64) + nrej.sum(dtype=torch.int64)).item())
        flops_step = 6 * F + 56 * N + 8 * S + 11
        flops_point = 40 + 14 * N + 25 * S
        flops = steps * flops_step + CH * M * T * flops_point
        peak = fma_peak or engine.measure_fma_peak(np.dtype(dt).name)
        achieved = flops / (ms * 1e-3) / 1e12
        roof = {"bound": "fp32" if a.dtype == "f32" else "fp64", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "ctx_measure_fma_peak (FMA loop measured on this GPU)",
                "flops_per_step": flops_step, "flops_per_saved_point": flops_point,
                "steps_per_eval_batch": steps}
    # e2e: one evaluation through the C ABI with HOST buffers (ctx_loglik_grad_host): per step the
    # operands (theta, sigma and the posterior's constant data block) go host->device from pinned
    # memory and the [CH, 1+P+S] block comes back -- the round trip a host-side sampler makes
    pin = lambda arr: torch.as_tensor(arr).pin_memory()
    hp = [pin(x) for x in host]
    outp = torch.empty((CH, 1 + 2 + 3), dtype=torch.float32 if a.dtype == "f32" else torch.float64).pin_memory()
    e_kw = dict(base_kw, likelihood="softlaplace")
    run_h = lambda: model.loglik_grad_host(*[x.numpy() for x in hp], out=outp.numpy(), **e_kw)
    for _ in range(3):
        run_h()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n_e2e = max(5, min(a.steps, 20))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        run_h()
    sec = cx.max_over_ranks(time.perf_counter() - t0)
    np.testing.assert_allclose(outp.numpy()[:64], out[:64].cpu().numpy(), rtol=1e-4, atol=1e-3)
    w = np.dtype(dt).itemsize
    e2e = {"value": CH * world * n_e2e / sec, "unit": "evals/s",
           "h2d_bytes_per_step": int(sum(x.numel() for x in hp) * w),
           "d2h_bytes_per_step": int(outp.numel() * w + 4 * CH * M), "steps_timed": n_e2e,
           "api": "ctx_loglik_grad_host (C ABI, pinned host buffers)"}
    # CPU baseline (rank 0, N=1): the oracle's C / OpenMP port on a bounded sample of the chains
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import loglik_oracle as lo
        from oracle.c_oracle import COracle

        mm = wl.michaelis_menten()
        orc = COracle(*mm[:4], mm[4], sens_theta=True, dtype=dt)
        n_s = min(CH, 2048)
        lo.loglik_grad_c(mm, y0s, theta[:64], consts, ts, data, sigma[:64], dtype=dt, oracle=orc,
                         rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096)          # warm-up
        t0 = time.perf_counter()
        o_cpu, _, _ = lo.loglik_grad_c(mm, y0s, theta[:n_s], consts, ts, data, sigma[:n_s], dtype=dt,
                                       oracle=orc, rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096)
        sec_c = time.perf_counter() - t0
        agree = float(np.abs(o_cpu[:, 0] - out[:n_s, 0].double().cpu().numpy()).max()
                      / np.abs(o_cpu[:, 0]).max())
        cpu = {"value": n_s / sec_c, "unit": "evals/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"first {n_s} of the {CH} chains x {M} series once ({sec_c:.2f} s): C/OpenMP "
                         f"port of the reference algorithm with parameter tangents + numpy "
                         f"likelihood terms (oracle/loglik_oracle.py loglik_grad_c); NOT numpyro",
               "max_rel_diff_logp_vs_gpu": agree}
    # series-sharded form of the SAME posterior (SURVEY 8d config 4 asks for both): all ranks
    # evaluate the same CH chains (rank 0's parameters) over M/world series each, followed by the
    # path's only collective, an all-reduce(sum) of the [CH, 1+P+S] block.  The evaluation
    # (likelihood kernel + reduce kernel + NCCL all-reduce) is captured in ONE CUDA graph, as a
    # sampler would replay it every leapfrog step; kernels and collective are also timed apart.
    series = None
    if world > 1 and M % world == 0:
        s_host = nuts_inputs(CH, M, T, dt, 2, dev)
        lo_, hi_ = rank * (M // world), (rank + 1) * (M // world)
        y0s0, theta0, consts0, ts0, data0, sigma0 = s_host
        sargs = [tt(x) for x in (y0s0[lo_:hi_], theta0, consts0[lo_:hi_], ts0[lo_:hi_], data0[lo_:hi_], sigma0)]
        s_kw = dict(base_kw, likelihood="softlaplace")

        from catalax_b200.parallel import PeerAllReduce

        try:
            p2p = PeerAllReduce(CH * (1 + 2 + 3) * 8)
        except Exception as exc:
            p2p = None
            p2p_err = f"{type(exc).__name__}: {exc}"[:200]

        def series_eval(reduce=True):
            o, _, st = model.loglik_grad(*sargs, **s_kw)
            if reduce == "nccl" or (reduce is True and p2p is None):
                dist.all_reduce(o, op=dist.ReduceOp.SUM)
            elif reduce:
                p2p(o)
            return o, st

        def graphed(reduce):
            for _ in range(3):
                series_eval(reduce)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    so_, sst_ = series_eval(reduce)
            torch.cuda.current_stream().wait_stream(side)
            return g, so_, sst_

        breakdown = {}
        try:
            g_full, so, sst = graphed(True)
            sms, _, _ = cx.timed(lambda: g_full.replay(), iters=n_iter)
            g_k, _, _ = graphed(False)
            kms, _, _ = cx.timed(lambda: g_k.replay(), iters=n_iter)
            g_n, so_n, _ = graphed("nccl")
            nms, _, _ = cx.timed(lambda: g_n.replay(), iters=n_iter)
            breakdown["graph_with_nccl_allreduce_ms"] = nms
            breakdown["p2p_equals_nccl"] = bool(torch.allclose(so, so_n, rtol=1e-4, atol=1e-3))
            mode = ("one CUDA graph per evaluation: likelihood kernel + reduce kernel + "
                    + ("one-shot NVLink all-reduce kernel (ctx_p2p_allreduce)" if p2p is not None
                       else f"NCCL all-reduce (peer all-reduce unavailable: {p2p_err})"))
        except Exception as exc:      # graph capture unavailable: eager launches (all ranks agree)
            flag = cx.sum_over_ranks(1.0)[0]
            sms, (so, sst), _ = cx.timed(lambda: series_eval(True), iters=n_iter)
            kms, _, _ = cx.timed(lambda: series_eval(False), iters=n_iter)
            mode = f"eager launches (graph capture failed on {int(flag)} rank(s): {type(exc).__name__}: {exc})"[:200]
        else:
            cx.sum_over_ranks(0.0)
        ems, _, _ = cx.timed(lambda: series_eval(True), iters=n_iter)
        blk = torch.zeros((CH, 1 + 2 + 3), dtype=sargs[1].dtype, device=dev)
        cms, _, _ = cx.timed(lambda: dist.all_reduce(blk, op=dist.ReduceOp.SUM), iters=n_iter)
        breakdown.update({"graph_ms": sms, "kernels_only_graph_ms": kms, "eager_python_ms": ems,
                          "nccl_allreduce_alone_ms": cms})
        if p2p is not None:
            pms, _, _ = cx.timed(lambda: p2p(blk), iters=n_iter)
            breakdown["p2p_allreduce_alone_ms"] = pms
        series = {"value": CH / (sms * 1e-3), "unit": "evals/s",
                  "ms_per_eval_batch": sms, "series_per_rank": M // world, "launch": mode,
                  "breakdown_ms": breakdown,
                  "collective": f"all-reduce(sum) of [{CH}, {1 + 2 + 3}] {a.dtype} per evaluation: "
                                + ("ctx_p2p_allreduce (one kernel over NVLink peer memory)" if p2p is not None else "NCCL"),
                  "scaling": "strong (one posterior, series split over ranks)",
                  "all_finite": bool(torch.isfinite(so).all().item())}
    nm = res["normal"]
    return {"metric": "NUTS logp+grad evals/s", "value": CH * world / (ms * 1e-3),
            "cost_ordered": {"value": CH * world / (ms_h * 1e-3), "unit": "evals/s", "ms_per_eval_batch": ms_h,
                             "same_bits_as_unhinted": hinted_same_bits, "clocks": clocks_h,
                             "schedule": "trajectories handed out most-expensive-first, order taken from an "
                                         "evaluation at parameters perturbed by LogN(0, 0.02) (the previous "
                                         "leapfrog step of a sampler); BayesianModel.log_density does this "
                                         "automatically"},
            "series_sharded": series,
            "roofline": roof, "e2e": e2e, "cpu_baseline": cpu,
            "unit": "evals/s", "ms_per_eval_batch": ms, "timed_iterations": n_iter,
            "clocks": res["softlaplace"]["clocks"],
            "series_solves_with_gradient_per_s": CH * world * M / (ms * 1e-3),
            "normal_likelihood": {"value": CH * world / (nm["ms"] * 1e-3), "unit": "evals/s",
                                  "ms_per_eval_batch": nm["ms"],
                                  "all_finite": bool(torch.isfinite(nm["out"]).all().item())},
            "config": {"chains_per_gpu": CH, "series": M, "points": T, "observables": 3,
                       "likelihood": "SoftLaplace", "rtol": 1e-5, "atol": 1e-5,
                       "theta_star": "(K_m, k_cat) = (40, 0.3); per-chain theta ~ theta* LogN(0, 0.2), "
                                     "sigma ~ LogN(log 0.5, 0.3) (README values deplete: DESIGN section 2)",
                       "sharding": "by chain (no collective)"},
            "failed_series": fails, "all_finite": finite}


# ----------------------------------------------------------------------------- other configs
def run_other_configs(cx, model, sets, fma_peak=None):
    """Device-timed legs for the other BASELINE configs, at every N (rows sharded by rank, max over
    ranks, never part of `value`): configs[1] in the integrator-bound regime (T=16) and on
    SURVEY 8d's original K_m range with throw=False; configs[2] (32-state mass-action network,
    262 144 rows split over the ranks); configs[4] (NeuralODE / RateFlowODE / UniversalODE fields
    5-64-64-64-4 on the tcgen05 tensor cores).  Each with >= 20 timed iterations after >= 0.3 s of
    warm-up, its own clock sample, its roofline, and (N=1) a CPU baseline on a bounded sample."""
    import torch

    import catalax_b200 as ctx
    from catalax_b200 import engine, workloads as wl
    from catalax_b200.neural import NeuralODE, RateFlowODE, UniversalODE
    from catalax_b200.tools import codegen as cg

    a, dev, rank, world = cx.a, cx.dev, cx.rank, cx.world
    dt = np_dtype(a)
    tdt = torch.float32 if a.dtype == "f32" else torch.float64
    want_cpu = rank == 0 and world == 1 and not a.no_cpu_baseline
    peak = fma_peak
    res = {}

    def fp_roof(flops, ms):
        ach = flops / (ms * 1e-3) / 1e12
        r = {"bound": a.dtype.replace("f", "fp"), "achieved": ach, "unit": "TFLOP/s"}
        if peak:
            r.update(peak=peak, frac=ach / peak,
                     peak_source="ctx_measure_fma_peak (FMA loop measured on rank 0's GPU)")
        return r

    # ---- configs[1], T=16: the output is negligible, the integrator is the bound (weak scaling)
    y0, th, c = sets[0]
    ts16 = torch.linspace(0, 100, 16, dtype=tdt, device=dev)
    out16 = torch.empty((y0.shape[0], 16, 3), dtype=tdt, device=dev)
    ms, o, clocks = cx.timed(lambda: model.solve(y0, th, c, ts16, solver="tsit5", in_axes=(0, 0, 0, None),
                                                 rtol=1e-6, atol=1e-6, dt0=0.1, max_steps=4096, out=out16))
    (steps,) = cx.sum_over_ranks(steps_of(o))
    fps = 6 * 6 + 64 * 3 + 11
    res["config2_T16"] = {"ms": ms, "steps_per_s": steps / (ms * 1e-3), "rows_per_gpu": int(y0.shape[0]),
                          "scaling": "weak", "kernel": engine.last_solve_kernel(), "clocks": clocks,
                          "roofline": dict(fp_roof(steps / world * fps, ms), flops_per_step=fps)}
    if want_cpu:
        from oracle.c_oracle import COracle

        mm = wl.michaelis_menten()
        co = COracle(*mm[:4], mm[4], dtype=dt)
        hy, hth, hc = wl.mm_inputs(1 << 18, seed=0, dtype=dt)
        t0 = time.perf_counter()
        r = co.solve(hy, hth, hc, np.linspace(0, 100, 16).astype(dt), rtol=1e-6, atol=1e-6, dt0=0.1, max_steps=4096)
        sec = time.perf_counter() - t0
        res["config2_T16"]["cpu_baseline"] = {
            "value": int((r.n_accepted.astype(np.int64) + r.n_rejected).sum()) / sec, "unit": UNIT,
            "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{1 << 18} rows once ({sec:.2f} s), C/OpenMP port of the reference algorithm"}

    # ---- configs[1] on SURVEY 8d's K_m ~ LogU(0.01, 1): ~3.6 % of the rows exhaust max_steps
    # (throw=False semantics: inf fill); reported beside the headline, same kernel and grid
    lo_km, hi_km = wl.KM_RANGE_SURVEY
    hy, hth, hc = wl.mm_inputs(a.rows, seed=5000 + rank, dtype=dt, km_range=wl.KM_RANGE_SURVEY)
    sy0, sth, sc = (torch.as_tensor(x, device=dev) for x in (hy, hth, hc))
    tsT = torch.linspace(0, 100, a.T, dtype=tdt, device=dev)
    outT = torch.empty((a.rows, a.T, 3), dtype=tdt, device=dev)
    ms, o, clocks = cx.timed(lambda: model.solve(sy0, sth, sc, tsT, solver="tsit5", in_axes=(0, 0, 0, None),
                                                 rtol=1e-6, atol=1e-6, dt0=0.1, max_steps=4096,
                                                 kernel=a.kernel, out=outT), iters=SIDE_MIN_ITERS)
    steps, failed = cx.sum_over_ranks(steps_of(o), int((o.status != 0).sum().item()))
    res["config2_survey_km_range"] = {
        "workload": f"same as the headline but K_m~LogU({lo_km},{hi_km}) (SURVEY 8d), throw=False",
        "ms": ms, "steps_per_s": steps / (ms * 1e-3), "rows_per_gpu": a.rows, "scaling": "weak",
        "failed_rows": int(failed), "failed_fraction": failed / (a.rows * world),
        "steps_per_row": steps / (a.rows * world), "kernel": engine.last_solve_kernel(), "clocks": clocks,
        "hbm_gbs": a.rows * a.T * 3 * np.dtype(dt).itemsize / (ms * 1e-3) / 1e9}
    del outT, sy0, sth, sc

    # ---- configs[2]: 32 states / 64 reactions, 262144 rows in total, T=32 on [0, 10]; the rows
    # are split evenly over the ranks ("sharded over 1/2/4/8 B200": strong scaling)
    B3_total = 262144
    B3 = B3_total // world
    f3 = cg.generate_field(wl.field_spec(wl.mass_action_network(32)))
    m3 = engine.CompiledModel(f3, np.dtype(dt).name)
    hy3, hk3, _ = wl.mass_action_inputs(B3_total, seed=1, dtype=dt)
    y3 = torch.as_tensor(hy3[rank * B3:(rank + 1) * B3], device=dev)
    k3 = torch.as_tensor(hk3[rank * B3:(rank + 1) * B3], device=dev)
    c3 = torch.empty((B3, 0), dtype=tdt, device=dev)
    ts3 = torch.linspace(0, 10, 32, dtype=tdt, device=dev)
    out3 = torch.empty((B3, 32, 32), dtype=tdt, device=dev)
    ms, o, clocks = cx.timed(lambda: m3.solve(y3, k3, c3, ts3, in_axes=(0, 0, 0, None), rtol=1e-6,
                                              atol=1e-6, out=out3))
    steps, failed = cx.sum_over_ranks(steps_of(o), int((o.status != 0).sum().item()))
    fps = 6 * f3.flops_rhs + 64 * 32 + 11
    res["config3_mass_action_32x64"] = {
        "ms": ms, "steps_per_s": steps / (ms * 1e-3), "rows_total": B3_total, "rows_per_gpu": B3,
        "scaling": "strong", "kernel": engine.last_solve_kernel() + " (warp per trajectory)",
        "failed_rows": int(failed), "clocks": clocks,
        "roofline": dict(fp_roof(steps / world * fps, ms), flops_per_step=fps)}
    if want_cpu:
        from oracle.c_oracle import COracle

        ma = wl.mass_action_network(32)
        co = COracle(*ma[:4], ma[4], dtype=dt)
        n_s = 16384
        t0 = time.perf_counter()
        r = co.solve(hy3[:n_s], hk3[:n_s], np.zeros((n_s, 0), dtype=dt), np.linspace(0, 10, 32).astype(dt),
                     rtol=1e-6, atol=1e-6, dt0=0.1, max_steps=4096)
        sec = time.perf_counter() - t0
        res["config3_mass_action_32x64"]["cpu_baseline"] = {
            "value": int((r.n_accepted.astype(np.int64) + r.n_rejected).sum()) / sec, "unit": UNIT,
            "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"first {n_s} rows once ({sec:.2f} s), C/OpenMP port of the reference algorithm"}
    del out3, y3, k3

    # ---- configs[4]: neural vector fields 5-64-64-64-4, 524288 rows per GPU, T=16 on [0, 10],
    # rtol 1e-3 / atol 1e-6 (f32 only: the tensor-core form; f64 runs the FFMA form)
    if a.dtype == "f32":
        B5 = 1 << 19
        ctx.enable_x64(False)
        names = ["a", "b", "c", "d"]
        uode_model = ctx.Model(name="config5 mechanistic part")
        uode_model.add_state(", ".join(names))
        uode_model.add_ode("a", "-k1 * a")
        uode_model.add_ode("b", "k1 * a - k2 * b / (K + b)")
        uode_model.add_ode("c", "k2 * b / (K + b) - k3 * c")
        uode_model.add_ode("d", "k3 * c")
        for pn, v in (("k1", 0.3), ("k2", 0.5), ("k3", 0.2), ("K", 1.0)):
            uode_model.parameters[pn].value = v
        nets = {
            "neuralode": NeuralODE(4, 64, 3, names, activation="tanh", max_time=10.0, key=3),
            "rateflow": RateFlowODE(4, 3, 64, 3, names, activation="celu", max_time=10.0, key=3),
            "universalode": UniversalODE(4, 64, 3, uode_model, activation="celu", max_time=10.0, key=3),
        }
        y5 = torch.as_tensor(np.random.default_rng(3 + rank).uniform(0, 2, (B5, 4)).astype(np.float32), device=dev)
        ts5 = torch.linspace(0, 10, 16, dtype=torch.float32, device=dev)
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        tpeak = float(peaks.get("bf16_tflops_sustained", 1384.0)) / 2.0   # tf32 dense = bf16 / 2
        for kind, net in nets.items():
            try:
                ms, _, clocks = cx.timed(lambda: net.predict_arrays(ts5, y5), iters=SIDE_MIN_ITERS)
                o = net.last
                steps, failed = cx.sum_over_ranks(steps_of(o), int((o.status != 0).sum().item()))
                # 7 MLP evaluations per step attempt in the tile kernel (the first-stage slope is
                # re-evaluated), 6 algorithmically; 8768 MAC each
                useful = steps / world * 6 * 2 * (5 * 64 + 64 * 64 + 64 * 64 + 64 * net.spec.out_size)
                entry = {
                    "ms": ms, "steps_per_s": steps / (ms * 1e-3), "rows_per_gpu": B5, "scaling": "weak",
                    "kernel": engine.last_solve_kernel() + " (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators)",
                    "useful_mlp_tflops": useful / (ms * 1e-3) / 1e12, "failed_rows": int(failed),
                    "clocks": clocks,
                    "roofline": {"bound": "tensor", "achieved": useful / (ms * 1e-3) / 1e12, "peak": tpeak,
                                 "unit": "TFLOP/s", "frac": useful / (ms * 1e-3) / 1e12 / tpeak,
                                 "issued_tflops_3xtf32": 3.0 * useful / (ms * 1e-3) / 1e12,
                                 "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32); "
                                                "achieved counts USEFUL MLP flops (each product is "
                                                "issued as three tf32 MMAs)"}}
                res[f"config5_{kind}_5x64x64x64x{net.spec.out_size}"] = entry
            except Exception as exc:
                if world > 1:          # a rank that skips a leg would desynchronise the collectives
                    raise
                res[f"config5_{kind}"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if want_cpu:
            # the reference algorithm on the CPU: numpy MLP field (BLAS) in the lock-step oracle
            try:
                from oracle import diffrax_oracle as do, mlp_oracle as mo

                net = nets["neuralode"]
                n_s = 4096
                rhs = mo.make_rhs("neuralode", [w.astype(np.float32) for w in net.weights],
                                  [None if b is None else b.astype(np.float32) for b in net.biases],
                                  "tanh", "identity", net.max_time)
                y5h = y5[:n_s].cpu().numpy()
                t0 = time.perf_counter()
                r = do.solve(rhs, y5h, np.zeros(0, np.float32), np.zeros((n_s, 0), np.float32),
                             np.linspace(0, 10, 16).astype(np.float32), rtol=1e-3, atol=1e-6,
                             dt0=10.0 / 15, dtype=np.float32, t0_from_ts=True)
                sec = time.perf_counter() - t0
                key = next(k for k in res if k.startswith("config5_neuralode"))
                res[key]["cpu_baseline"] = {
                    "value": int((r.n_accepted.astype(np.int64) + r.n_rejected).sum()) / sec, "unit": UNIT,
                    "cores": os.cpu_count() or 1, "kind": "port",
                    "sample": f"first {n_s} rows once ({sec:.2f} s): numpy restatement (lock-step vmap "
                              f"semantics, BLAS matmuls) of the reference algorithm"}
            except Exception as exc:
                res["config5_cpu_baseline_error"] = f"{type(exc).__name__}: {exc}"[:200]
    return res


# ----------------------------------------------------------------------------- GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist

    from catalax_b200 import engine, workloads as wl
    from catalax_b200.tools import codegen as cg

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
    field = cg.generate_field(wl.field_spec(wl.michaelis_menten()))
    model = engine.CompiledModel(field, np.dtype(dt).name)
    # FMA peak (denominator of the compute rooflines of the side legs) first: the micro-benchmark
    # loads the GPU heavily, nothing timed may run in its wake (input generation + warm-up follow)
    need_peak = rank == 0 and not (a.no_nuts and a.no_other)
    fma_peak = engine.measure_fma_peak(np.dtype(dt).name) if need_peak else None

    # rows are sharded by rank: every rank draws its own seeds (no data-path collective)
    sets = []
    for k in range(N_INPUT_SETS):
        y0, th, c = wl.mm_inputs(B, seed=1000 * rank + k, dtype=dt)
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
        steps_of_set.append(steps_of(o))
        fails += int((o.status != 0).sum().item())
    headline_kernel = engine.last_solve_kernel()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()           # nvidia-smi needs ~100 ms to start: begin before the warm-up
    cx = Ctx(a, dev, rank, world, dist, sampler)
    t_w = time.perf_counter()
    i = 0
    while i < max(a.warmup, 3) or time.perf_counter() - t_w < 0.3:   # >= 3 steps and >= 0.3 s
        step(i)
        i += 1
        if i % 8 == 0:
            torch.cuda.synchronize()
    n_warm = i
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    m0 = sampler.mark() if sampler else None
    launches0 = engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    m1 = sampler.mark() if sampler else None
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = engine.launch_count() - launches0
    clocks = sampler.summary(m0, m1) if sampler else None
    my_steps = sum(steps_of_set[i % N_INPUT_SETS] for i in range(a.steps))

    # secondary: the same K passes with a scheduling hint per input set -- rows handed out
    # most-expensive-first, the order taken from the step counts of the (untimed) counting pass
    # above.  Never part of `value`: a one-shot simulate has no previous solve to learn from.
    hinted = None
    if not a.no_other:
        orders = []
        for k in range(N_INPUT_SETS):
            orders.append(step(k).cost_order())
        torch.cuda.synchronize()

        def step_h(i):
            y0, th, c = sets[i % N_INPUT_SETS]
            return model.solve(y0, th, c, ts, order=orders[i % N_INPUT_SETS], **kw)

        for i in range(max(a.warmup, 3)):
            step_h(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for i in range(a.steps):
            step_h(i)
        h1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        hinted = cx.max_over_ranks(h0.elapsed_time(h1))
        # ... and with an IMPERFECT hint: the order learnt at parameters a step of an optimiser away
        # (K_m, k_cat scaled by LogN(0, 0.02) per row), applied to the unperturbed solve
        gen = torch.Generator(device=dev).manual_seed(4242 + rank)
        orders_p = []
        for k in range(N_INPUT_SETS):
            y0, th, c = sets[k]
            th_p = th * torch.exp(0.02 * torch.randn(th.shape, generator=gen, device=dev, dtype=th.dtype))
            orders_p.append(model.solve(y0, th_p, c, ts, **kw).cost_order())
        torch.cuda.synchronize()
        orders, orders_exact = orders_p, orders
        for i in range(max(a.warmup, 3)):
            step_h(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for i in range(a.steps):
            step_h(i)
        h1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        hinted_p = cx.max_over_ranks(h0.elapsed_time(h1))
        orders = orders_exact
    # secondary: the same K passes issued alternately on TWO streams into two output buffers.  One
    # launch ends with 0.7-0.9 ms in which only its longest trajectories are still running (DESIGN
    # section 4); a caller with independent solves in flight (chunks of a larger data set, the chains
    # of an optimiser) fills that tail with the next launch.  Never part of `value`.
    pipelined = {}
    if not a.no_other:
        out2 = torch.empty_like(out)
        streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        main_stream = torch.cuda.current_stream(dev)

        def step_p(i, use_order):
            with torch.cuda.stream(streams[i & 1]):
                y0, th, c = sets[i % N_INPUT_SETS]
                model.solve(y0, th, c, ts, order=orders[i % N_INPUT_SETS] if use_order else None,
                            **{**kw, "out": (out, out2)[i & 1]})

        for name, use_order in (("pipelined", False), ("pipelined_cost_ordered", True)):
            for i in range(max(a.warmup, 3) + 1):
                step_p(i, use_order)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for st in streams:
                st.wait_stream(main_stream)
            for i in range(a.steps):
                step_p(i, use_order)
            for st in streams:
                main_stream.wait_stream(st)
            p1.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            pipelined[name] = cx.max_over_ranks(p0.elapsed_time(p1))
        del out2
    ms_max = cx.max_over_ranks(ms)
    total_steps, total_fails = cx.sum_over_ranks(my_steps, fails)
    total_fails = int(total_fails)
    value = total_steps / (ms_max * 1e-3)

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (copies inside the timed region)
    e2e = None
    if not a.no_e2e:
        import psutil

        e2e_rows = B
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        while e2e_rows > 4096 and (e2e_rows * T * S * w) * local_world * 3 > psutil.virtual_memory().available:
            e2e_rows //= 2
        y0h, thh, ch = wl.mm_inputs(e2e_rows, seed=1000 * rank, dtype=dt)
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
        sec = cx.max_over_ranks(time.perf_counter() - t0)
        (e2e_total,) = cx.sum_over_ranks(e2e_steps * n_e2e)
        # the same transfer alone: one cudaMemcpyAsync per 256 MB chunk of the pinned result
        # buffer, all ranks at the same time -- the box's ceiling for this step's D2H traffic
        if world > 1:
            dist.barrier()
        chunk = (256 << 20) // (T * S * w)
        t0 = time.perf_counter()
        for r0 in range(0, e2e_rows, chunk):
            ysp[r0:r0 + chunk].copy_(out[r0:r0 + chunk], non_blocking=True)
        torch.cuda.synchronize()
        d2h_sec = cx.max_over_ranks(time.perf_counter() - t0)
        h2d = e2e_rows * (S + P) * w + T * w
        d2h = e2e_rows * T * S * w + 3 * 4 * e2e_rows
        e2e = {"value": e2e_total / sec, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "rows_per_step": e2e_rows,
               "steps_timed": n_e2e, "ms_per_step": 1e3 * sec / n_e2e,
               "api": "ctx_solve_host (C ABI, pinned host buffers; chunks of the batch are solved "
                      "while the previous chunk's trajectories copy back)",
               "d2h_ceiling": {"ms_per_step": 1e3 * d2h_sec,
                               "gbps_per_gpu": e2e_rows * T * S * w / d2h_sec / 1e9,
                               "frac_of_ceiling": d2h_sec * n_e2e / sec,
                               "how": f"{world} rank(s) copying their [rows,T,S] result to pinned host "
                                      f"memory at the same time, nothing else running"}}
        del ysp

    # ---- second headline metric: NUTS log-density + gradient evaluations / s (configs[3])
    nuts = None
    if not a.no_nuts:
        try:
            nuts = run_nuts_leg(cx, fma_peak=fma_peak)
        except Exception as exc:
            if world > 1:
                raise
            nuts = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    other = None
    if not a.no_other:
        try:
            other = run_other_configs(cx, model, sets, fma_peak=fma_peak)
        except Exception as exc:       # never lose the headline line to a side leg
            if world > 1:
                raise
            other = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if sampler:
        sampler.stop()

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
    traffic = None
    for name in (f"r2_solve_{a.dtype}_T{T}.json", f"r1_solve_{a.dtype}_T{T}.json"):
        prof = ROOT / "profiles" / name
        if prof.exists():
            try:
                traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
                break
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
                   "kernel": headline_kernel,
                   "inputs_note": "K_m range differs from SURVEY 8d's LogU(0.01,1) (3.6 % of those rows "
                                  "exhaust max_steps): that range is the other_configs."
                                  "config2_survey_km_range leg",
                   "failed_rows": total_fails, "steps_per_row": total_steps / (B * world * a.steps)},
        "gpu_launches": launches, "clocks": clocks,
        "saved_points_per_s": B * world * T * a.steps / (ms_max * 1e-3),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "write_only_peak_gbs": write_only},
    }
    if hinted is not None:
        line["cost_ordered"] = {
            "ms_per_step": hinted / a.steps, "value": total_steps / (hinted * 1e-3), "unit": UNIT,
            "hbm_frac": algo_bytes / (hinted / a.steps * 1e-3) / 1e9 / hbm_peak,
            "ms_per_step_order_from_perturbed_parameters": hinted_p / a.steps,
            "schedule": "rows handed out most-expensive-first (ctx_solve_cfg.order), the order taken from the "
                        "step counts of a previous solve of the same input set; results bit-identical; "
                        "not part of `value`"}
    for name, t_ms in pipelined.items():
        line[name] = {
            "ms_per_step": t_ms / a.steps, "value": total_steps / (t_ms * 1e-3), "unit": UNIT,
            "hbm_frac": algo_bytes / (t_ms / a.steps * 1e-3) / 1e9 / hbm_peak,
            "schedule": "the same K passes issued alternately on two CUDA streams (two output buffers): the "
                        "next launch fills the long-trajectory tail of the previous one"
                        + ("; rows most-expensive-first as in cost_ordered" if "cost" in name else "")
                        + "; not part of `value`"}
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
                      f"the reference algorithm (oracle/ctx_oracle_core.h); NOT JAX"}
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
