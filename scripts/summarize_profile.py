"""Turn gpurun_out/<tag>_* ncu captures into tracked summaries under profiles/ (runs on the CPU
box: `ncu -i` needs no GPU).  usage: summarize_profile.py <tag>"""
import csv, io, json, subprocess, sys
from pathlib import Path

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = Path("profiles"); out.mkdir(exist_ok=True)
g = Path("gpurun_out")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max"]


def launches():
    f = g / f"{tag}_launches.csv"
    if not f.exists():
        return
    lines = f.read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    agg, tot = {}, 0.0
    for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)
        a = agg.setdefault(r["Kernel Name"], [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
    with open(out / f"{tag}_bench_launches.md", "w") as fo:
        fo.write(f"# ncu launch list ({tag}): `python bench.py --steps 4 --warmup 3 --no-e2e "
                 "--no-cpu-baseline --no-nuts`\n\n`ncu --metrics gpu__time_duration.sum "
                 "--clock-control none -c 400` -- per-launch times are cold-cache and serialised; "
                 "compare SHARES, not absolutes.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fo.write(f"| `{k[:90]}` | {n} | {ms:.3f} | {100*ms/tot:.1f}% |\n")
    print("launch list ok")


def full(rep_name, kernel, label):
    rep = g / f"{tag}_{rep_name}.ncu-rep"
    if not rep.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = r[0], r[1], r[2]
    m = {w: (vals[hdr.index(w)], units[hdr.index(w)]) for w in WANT if w in hdr}
    def num(k):
        v, u = m[k]; v = float(v.replace(",", ""))
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
    # stall reasons (source page)
    src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h2 = rows[1]
    stalls = {}
    for c in h2:
        if c.startswith("stall_") and "Not Issued" not in c:
            i = h2.index(c)
            stalls[c] = sum(int(x[i]) for x in rows[2:] if len(x) > i and x[i].isdigit())
    tot = sum(stalls.values()) or 1
    summary = {"kernel": kernel, "workload": label,
               "metrics": {k: f"{v} {u}".strip() for k, (v, u) in m.items()},
               "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
               "stall_share": {k: round(v / tot, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}}
    (out / f"{tag}_{rep_name}.json").write_text(json.dumps(summary, indent=1))
    print(rep_name, summary["metrics"].get("gpu__time_duration.sum"), summary["stall_share"])


launches()
full("solve_full", "ctx_solve_v2" if tag >= "r1n" else "ctx_solve_v1",
     "bench.py f32 T=1000 B=2^20 (Michaelis-Menten, config 2)")
full("loglik_full", "ctx_loglik_v2" if tag >= "r1i" else "ctx_loglik_v1", "config 4: 16384 chains x 64 series x 20 points, f32, SoftLaplace")
full("mlp_tc_full", "ctx_solve_mlp_tc", "config 5 (B=262144): NeuralODE 5-64-64-64-4 tanh, f32, tcgen05")
full("solve_w_full", "ctx_solve_w", "config 3: mass-action network 32 states / 64 reactions, 262144 rows, T=32, f32 (scripts/dev_bench3_stab.py)")
full("v1_T16", "ctx_solve_v1", "config 2 in the integrator-bound regime: f32 T=16 B=2^20 (scripts/dev_bench.py --T 16 --kernel 2)")
full("v2_T16", "ctx_solve_v2", "config 2 in the integrator-bound regime: f32 T=16 B=2^20, stepping-heavy layout")
