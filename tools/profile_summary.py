#!/usr/bin/env python
"""Turn the ncu artefacts of a gpurun call into the committed summary under profiles/.

  python tools/profile_summary.py <launches.csv> <full.ncu-rep> <bench.json> <out.md> [<traffic.json>]
"""
import csv, json, subprocess, sys, io, collections

launch_csv, rep, bench_json, out_md = sys.argv[1:5]
traffic_json = sys.argv[5] if len(sys.argv) > 5 else None

rows = [r for r in csv.reader(l for l in open(launch_csv) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0]
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += float(r[vi].replace(",", ""))
allns = sum(v[1] for v in tot.values())

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
names, units, vals = rr[0], rr[1], rr[2]
m = {n: (v, u) for n, u, v in zip(names, units, vals)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]

def gb(x):
    v, u = m[x]
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

b = json.loads(open(bench_json).read().strip().splitlines()[-1])
with open(out_md, "w") as f:
    f.write("# Round 1 (final) -- B200, astrobeeSE3 B=1024 N=50\n\n")
    f.write(f"Sources: `{launch_csv.split('/')[-1]}` (ncu launch list of `bench.py --steps 2 --warmup 1`), an `ncu --set full` capture of one "
            f"`ipm_kernel<2>` launch (the .ncu-rep stays in gpurun_out/, 20 MB), `{bench_json.split('/')[-1]}` (bench.py --steps 20 --warmup 5).\n\n")
    f.write("## ncu launch list (cold-cache, serialised; compare shares)\n\n| kernel | launches | total ns | share |\n|---|---|---|---|\n")
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {ns:.0f} | {100 * ns / allns:.2f}% |\n")
    f.write("\n## ncu --set full, ipm_kernel<2> (B=1024, one launch)\n\n")
    for w in want:
        if w in m:
            f.write(f"* `{w}` = {m[w][0]} {m[w][1]}\n")
    tr = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    f.write(f"\nDRAM traffic of the launch: {tr / 1e9:.2f} GB = {tr / 1024 / 1e6:.2f} MB per instance-solve "
            f"(algorithmic: {b['kernels']['solve']['algorithmic_bytes'] / 1024 / 1e6:.3f} MB).\n")
    f.write("\n## bench.py\n\n")
    f.write(f"* value {b['value']:.0f} {b['unit']} ({b['ms_per_step']:.2f} ms/step), e2e {b['e2e']['value']:.0f}\n")
    f.write("* kernels ms: " + ", ".join(f"{k} {v['ms']:.3f}" for k, v in b["kernels"].items()) + "\n")
    f.write(f"* roofline (dominant kernel, algorithmic bytes / measured time): {b['roofline']['achieved']:.1f} GB/s of {b['roofline']['peak']:.0f} "
            f"({100 * b['roofline']['frac']:.2f} %)\n")
    f.write(f"* trajectories/s {b['trajectories_per_sec']:.0f} ({b['full_solve']['converged']}/{b['full_solve']['instances']} converged in "
            f"{b['full_solve']['batch_iterations']} iterations)\n")
    if "cpu_baseline" in b:
        f.write(f"* cpu_baseline {b['cpu_baseline']['value']:.1f} {b['cpu_baseline']['unit']} on {b['cpu_baseline']['cores']} cores ({b['cpu_baseline']['sample']})\n")
    f.write(f"* clocks {b['clocks']}\n")
if traffic_json:
    json.dump({"workload": "astrobeeSE3 B=1024 N=50", "solve_kernel": tr, "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch"},
              open(traffic_json, "w"), indent=1)
print("wrote", out_md)
