#!/usr/bin/env python
"""Round-2 evidence: turn the ncu artefacts of a gpurun call into the committed summaries under profiles/.

  python tools/profile_summary_r02.py <launches.csv> <full.ncu-rep> [<more.ncu-rep> ...] --bench <bench.json> --out profiles/r02
writes <out>_summary.md, <out>_traffic.json, <out>_sass_opcodes.md, <out>_ncu_lines_solve.md
"""
import collections, csv, io, json, os, re, subprocess, sys

args = sys.argv[1:]
bench_json = args[args.index("--bench") + 1]; out = args[args.index("--out") + 1]
pos = [a for i, a in enumerate(args) if not a.startswith("--") and (i == 0 or args[i - 1] not in ("--bench", "--out"))]
launch_csv, reps = pos[0], pos[1:]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gusto.jl_b200", "libgusto_b200.so")

rows = [r for r in csv.reader(l for l in open(launch_csv) if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0]
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += float(r[vi].replace(",", ""))
allns = sum(v[1] for v in tot.values())

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
kern = collections.OrderedDict()
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    names, units = rr[0], rr[1]
    kn = names.index("Kernel Name")
    for vals in rr[2:]:
        k = vals[kn].split("(")[0].replace("void ", "")
        if k in kern:
            continue
        kern[k] = {n: (v, u) for n, u, v in zip(names, units, vals)}

def gbytes(m, x):
    v, u = m[x]; v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

b = json.loads(open(bench_json).read().strip().splitlines()[-1])
traffic = {}
with open(out + "_summary.md", "w") as f:
    f.write(f"# Round 2 -- B200, {b['config']['workload'].split(',')[0]}\n\n")
    f.write(f"Sources: `{os.path.basename(launch_csv)}` (ncu launch list, `--metrics gpu__time_duration.sum --clock-control none`, of `bench.py --steps 12 --warmup 3`), "
            f"`ncu --set full --clock-control none --import-source on` captures ({', '.join(os.path.basename(r) for r in reps)}; the .ncu-rep files stay in gpurun_out/), "
            f"`{os.path.basename(bench_json)}` (the default `bench.py` run of the same library).\n\n")
    f.write("## ncu launch list (cold-cache, serialised by the profiler: compare SHARES, not absolute times)\n\n| kernel | launches | total ns | share |\n|---|---|---|---|\n")
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {ns:.0f} | {100 * ns / allns:.2f}% |\n")
    sk = b["kernels"]; ssum = sum(v["ms"] for v in sk.values())
    f.write("\nCUDA-event kernel times of the bench (e2e pass): " + ", ".join(f"{k} {v['ms']:.3f} ms ({100 * v['ms'] / ssum:.1f} %)" for k, v in sk.items()) + "\n")
    for k, m in kern.items():
        f.write(f"\n## ncu --set full: `{k}` (one launch, B = 1024)\n\n")
        for w in WANT:
            if w in m:
                f.write(f"* `{w}` = {m[w][0]} {m[w][1]}\n")
        tr = gbytes(m, "dram__bytes_read.sum") + gbytes(m, "dram__bytes_write.sum")
        key = "solve" if "ipm" in k else "linearize" if "linearize" in k else "evaluate" if "evaluate" in k else k
        traffic[key + "_kernel"] = tr
        if key in sk:
            f.write(f"\nDRAM traffic of the launch: {tr / 1e9:.3f} GB = {tr / 1024 / 1e6:.3f} MB per instance; algorithmic bytes {sk[key]['algorithmic_bytes'] / 1e9:.3f} GB "
                    f"({tr / sk[key]['algorithmic_bytes']:.1f}x).\n")
    f.write("\n## bench.py (default run)\n\n")
    f.write(f"* value {b['value']:.0f} {b['unit']} ({b['ms_per_step']:.3f} ms/step over {b['steps']} steps of real solves), e2e {b['e2e']['value']:.0f} ({b['e2e']['ms_per_step']:.3f} ms/step)\n")
    f.write(f"* value / summed kernel time = {b['value_vs_kernel_sum']:.3f}; launches in the timed region {b['gpu_launches']}\n")
    f.write(f"* Newton iterations per solve {b['newton_iters_per_solve']:.2f}; solve success fraction {b['instance_iterations']['solve_success_fraction']:.4f}\n")
    f.write(f"* roofline (dominant kernel, algorithmic bytes / measured time): {b['roofline']['achieved']:.1f} GB/s of {b['roofline']['peak']:.0f} ({100 * b['roofline']['frac']:.2f} %)\n")
    f.write(f"* trajectories/s {b['trajectories_per_sec']:.0f}; full solves: {b['full_solve']}\n")
    for k in ("forced_steady_state", "c3_hard", "cpu_baseline", "cpu_baseline_compiled"):
        if k in b:
            f.write(f"* {k}: {json.dumps(b[k])}\n")
    f.write(f"* clocks {b['clocks']}\n")
json.dump({"workload": f"{b['config']['model']} B={b['config']['B_per_gpu']} N={b['config']['N']}", **traffic,
           "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch each"}, open(out + "_traffic.json", "w"), indent=1)

# SASS opcode summary per kernel of the in-tree library
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, ops = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); ops[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        ops[cur][m.group(1)] += 1
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = dict(re.findall(r"Function (\S+):\n\s+(REG:\d+ STACK:\d+ SHARED:\d+)", res))
SHOW = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "LDL", "STL", "SHFL", "BAR", "ATOMS", "ATOMG", "RED", "CCTL"]
with open(out + "_sass_opcodes.md", "w") as f:
    f.write("# SASS opcode counts per kernel of gusto.jl_b200/libgusto_b200.so (`cuobjdump -sass`, static counts; `cuobjdump -res-usage`)\n\n")
    f.write("DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64), UBLKCP / SYNCS = TMA bulk copy + mbarrier, LDGSTS = cp.async, LDL / STL = local-memory (spill / stack) accesses.\n\n")
    f.write("| kernel | instructions | " + " | ".join(SHOW) + " | resources |\n|---|---|" + "---|" * (len(SHOW) + 1) + "\n")
    for k, c in ops.items():
        short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        f.write(f"| `{short}` | {sum(c.values())} | " + " | ".join(str(c.get(o, 0)) for o in SHOW) + f" | {usage.get(k, '')} |\n")
print("wrote", out + "_summary.md", out + "_traffic.json", out + "_sass_opcodes.md")
