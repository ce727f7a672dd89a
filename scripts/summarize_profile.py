"""Turn gpurun_out/<tag>_launches.csv and <tag>_solve_full.ncu-rep into tracked summaries
under profiles/ (run on the CPU box: `ncu -i` needs no GPU)."""
import csv, io, json, subprocess, sys
from pathlib import Path

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
name = sys.argv[2] if len(sys.argv) > 2 else "solve_f32_T1000"
out = Path("profiles"); out.mkdir(exist_ok=True)
g = Path("gpurun_out")

# ---- launch list
rows = []
lines = (g / f"{tag}_launches.csv").read_text().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], ms))
tot = sum(ms for _, ms in rows)
agg = {}
for k, ms in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
with open(out / f"{tag}_{name}_launches.md", "w") as f:
    f.write(f"# ncu launch list ({tag}, `python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-nuts`)\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` — per-launch times are cold-cache and serialised; compare SHARES.\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k[:90]}` | {n} | {ms:.3f} | {100*ms/tot:.1f}% |\n")
print("launch list:", len(rows), "launches")

# ---- full capture
rep = g / f"{tag}_solve_full.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = r[0], r[1], r[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max"]
m = {}
for w in want:
    if w in hdr:
        i = hdr.index(w); m[w] = (vals[i], units[i])
def num(k):
    v, u = m[k]; v = float(v.replace(",", ""))
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
summary = {"kernel": "ctx_solve_v1", "metrics": {k: f"{v} {u}" for k, (v, u) in m.items()},
           "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum")}
(out / f"{tag}_{name}.json").write_text(json.dumps(summary, indent=1))
print(json.dumps(summary, indent=1))
