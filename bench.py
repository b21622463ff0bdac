#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched GuSTO SCP hot path (BASELINE.json: SCP iterations/s and
trajectories/s, astrobeeSE3 B=1024 N=50, 1/2/4/8 GPUs, next to the CPU restatement on the same host).

  python bench.py --gpus N --steps K --warmup W            one rank per GPU (torchrun for N > 1), weak scaling
  python bench.py --impl reference --gpus N --steps K ...  the CPU restatement (oracle) on the host cores

A "step" is one outer GuSTO SCP iteration over the whole batch: linearize (K1+K2) -> convex solve (K3) -> evaluate (K4) ->
accept/reject + Delta/omega schedule + convergence test (scp_gusto.jl:119-174).  The steps are the iterations of REAL solves
(force = false): every solve starts from the straight-line initialisation and runs until every instance of every rank has
converged or failed (3 iterations on the headline batch), then the next solve starts; the last solve is cut so that exactly K
steps are timed.  `value` counts the instance-iterations whose convex solve succeeded, per second, with the outer loop
resident on the device (gusto_scp_run: one CUDA graph per iteration, the host reads one counter per iteration); `e2e` runs the
same steps through the host-language loop of the public API with the trajectory uploaded from / downloaded to pinned host
buffers every step.  Extra keys: the forced steady state round 1 reported (30 forced iterations from one initialisation), the
hard tier of the same workload, per-kernel times, both CPU baselines.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "scp_instance_iterations_per_sec"
UNIT = "instance-iterations/s"
MAX_ITER = 30


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="astrobeeSE3")
    ap.add_argument("--batch", type=int, default=1024, help="instances per GPU")
    ap.add_argument("--knots", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances per step of the CPU arms (0 = one per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the forced-steady-state and hard-tier runs")
    return ap.parse_args()


def make_problem(pkg, name, B, N, seed, **extra):
    fn = pkg.problems.CONFIGS[name]
    kw = dict(B=B, seed=seed, **extra)
    if N:
        kw["N"] = N
    return fn(**kw)


A_NNZ = {"dubins": 2, "freeflyerSE2": 3, "astrobeeSE3": 27, "astrobeeSE3manifold": 33}    # Traits<M>::ANZ (csrc/common.cuh)


def algorithmic_bytes(bp):
    """Bytes one (instance, SCP iteration) must move at minimum, per kernel: the trajectory, the
    linearization blocks (A on its sparsity pattern, f, g) and the obstacle rows, each touched once per kernel."""
    nx, nu, N = bp.model.x_dim, bp.model.u_dim, bp.N
    no = int(bp.obstacle_table()[0].shape[0])
    traj = N * (nx + nu)
    anz = A_NNZ[bp.model.name]
    blocks = N * (anz + 2 * nx)              # A (pattern), f, g
    rows = N * no * 5
    k12 = traj + blocks + rows               # read traj, write blocks + rows
    k3 = traj + N * (anz + nx) + rows + traj   # read Xp/Up, A, g, rows; write candidate
    k4 = 2 * traj + N * (anz + nx) + rows      # read both trajectories, A, f, rows
    return dict(linearize=8 * k12, solve=8 * k3, evaluate=8 * k4)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, f"/tmp/gusto_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arms
# Both CPU arms run REAL solves (force = false) of instances of the same seeded batch, one worker process per host core,
# `width` instances side by side per round; a round's step count is the largest SCP iteration count among its instances
# (the batch semantics of the GPU arm) and the last round is cut so that exactly `steps` steps are timed.
def _cpu_worker(args):
    kind, name, N, seed, Btot, idx, iters = args
    pkg = entry.load_package()
    bp = make_problem(pkg, name, Btot, N, seed)
    if kind == "port":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import gusto_oracle as orc
        from gusto_oracle.subproblem import Problem
        m = orc.get_model(bp.model.name)
        p = Problem(m, bp.N, float(bp.tf[idx]), bp.x_init[idx], bp.goal_type, bp.goal_lo[idx], bp.goal_hi[idx], bp.obstacle_table())
        t = time.perf_counter()
        S = orc.solve_gusto(p, max_iter=iters)
        return S.iterations, int(S.converged), S.iterations, time.perf_counter() - t
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import compiled_baseline as cb
    return cb.solve_instance(pkg, bp.instance(idx), max_iter=iters)


def cpu_steps(kind, name, N, seed, Btot, width, steps, warmup, cores):
    """Runs `warmup` untimed + exactly `steps` timed steps.  Returns (instance-iterations/s, instance-iterations, seconds, solves)."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    if kind == "compiled":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import compiled_baseline as cb
        cb.build()
    nxt = [0]

    def round_(pool, cap):
        idx = [(nxt[0] + i) % Btot for i in range(width)]
        nxt[0] += width
        res = pool.map(_cpu_worker, [(kind, name, N, seed, Btot, i, cap) for i in idx])
        return sum(r[0] for r in res), max(r[2] for r in res), sum(r[1] for r in res)

    with mp.get_context("spawn").Pool(cores) as pool:
        left = warmup
        while left > 0:
            left -= round_(pool, min(MAX_ITER, left))[1]
        t = time.perf_counter()
        done = st = conv = solves = 0
        while st < steps:
            d, s, c = round_(pool, min(MAX_ITER, steps - st))
            done += d; st += s; conv += c; solves += width
        wall = time.perf_counter() - t
    return done / wall, done, wall, solves, conv, st


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    Btot = args.batch * args.gpus
    width = args.cpu_sample or cores
    pkg = entry.load_package()
    bp = make_problem(pkg, args.config, 2, args.knots, Btot)
    val, done, wall, solves, conv, st = cpu_steps("port", args.config, args.knots, Btot, Btot, width, args.steps, args.warmup, cores)
    sample = (f"{st} steps = the SCP iterations of real solves (force = false, straight-line start) of {solves} instances of the seeded "
              f"B={Btot} batch, {width} side by side ({done} instance-iterations in {wall:.1f} s, {conv} converged); NumPy/SciPy "
              "restatement, one process per core; the reference's Julia/JuMP/Gurobi path cannot run here (no Julia, no solver licences)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": st,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, st), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config} B={Btot} N={bp.N} (BASELINE configs[{2 if args.gpus == 1 else 3}]), GuSTO SCP iterations of real solves; "
                               f"CPU restatement on a bounded sample: {width} instances per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_cpu_baseline:
        v2, d2, w2, s2, c2, st2 = cpu_steps("compiled", args.config, args.knots, Btot, Btot, 8 * width, args.steps, args.warmup, cores)
        line["cpu_baseline_compiled"] = {
            "value": v2, "unit": UNIT, "cores": cores, "kind": "compiled",
            "sample": f"{st2} steps, {s2} instances ({d2} instance-iterations in {w2:.1f} s, {c2} converged): the GPU path's own kernel sources "
                      "compiled for the host (g++ -O3 -march=native, single thread per instance, one process per core)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gusto-b200 has no CPU path (use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.build()
    host = pkg.engine()
    Btot = args.batch * world
    bp_all = make_problem(pkg, args.config, Btot, args.knots, Btot)
    bp = bp_all.shard(rank, world) if world > 1 else bp_all
    B, sp = bp.B, bp.model.scp_params
    eng = host.Engine(bp, device=local)
    X0, U0 = bp.init_traj_straightline()
    if world > 1:
        # the library's own communicator: rank 0 creates the NCCL id, torch.distributed only carries the 128 bytes
        uid = torch.from_numpy(host.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(uid, 0)
        eng.comm_init(rank, world, uid.cpu().numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        tt = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def allsum(a):
        if world == 1:
            return np.asarray(a, dtype=np.float64)
        tt = torch.tensor(np.asarray(a, dtype=np.float64), device="cuda")
        dist.all_reduce(tt)
        return tt.cpu().numpy()

    # ---------------- device-resident steps: real solves, restarted from the initial trajectory kept on the device
    eng.scp_begin(X0, U0, False)                      # uploads X0 / U0 once; later begins restart from the device copy

    def device_steps(nsteps, force=False, stats=None):
        left = nsteps
        while left > 0:
            eng.scp_begin(None, None, force)
            n, unfinished = eng.scp_run(min(MAX_ITER, left))
            left -= n
            if stats is not None:
                _, cv, su, _, cnt = eng.scp_get(n + 1, want_hist=False)
                stats["ran"] += int(cnt[:, 0].sum()); stats["solved"] += int(cnt[:, 1].sum()); stats["accepted"] += int(cnt[:, 2].sum())
                stats["solves"] += 1
                if unfinished == 0:
                    stats["full_solves"] += 1; stats["converged"] += int(cv.sum()); stats["successful"] += int(su.sum())
                    stats["iters_per_solve"].append(n)
            if n == 0:
                break

    def timed(fn):
        barrier()
        launches0 = eng.launch_count()
        eng.timer_start()
        t0 = time.perf_counter()
        fn()
        dev_ms = eng.timer_stop()
        barrier()
        wall = time.perf_counter() - t0
        # the device stopwatch spans the same steps; take the slower of the two clocks, max over ranks
        return allmax(max(dev_ms, 1e3 * wall)), eng.launch_count() - launches0

    device_steps(max(3, args.warmup))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    stats = dict(ran=0, solved=0, accepted=0, solves=0, full_solves=0, converged=0, successful=0, iters_per_solve=[])
    t_ms, launches = timed(lambda: device_steps(args.steps, False, stats))
    clk = clocks.stop() if rank == 0 else None
    tot = allsum([stats["ran"], stats["solved"], stats["accepted"], stats["full_solves"] * B, stats["converged"], stats["successful"]])
    value = tot[1] / (t_ms / 1e3)

    extras = {}
    if not args.no_extras:
        # round 1's headline: 30 forced iterations from ONE initialisation (iterations 4..30 run on converged trajectories)
        fs = dict(ran=0, solved=0, accepted=0, solves=0, full_solves=0, converged=0, successful=0, iters_per_solve=[])
        device_steps(MAX_ITER, True)
        f_ms, _ = timed(lambda: device_steps(MAX_ITER, True, fs))
        ft = allsum([fs["ran"], fs["solved"]])
        extras["forced_steady_state"] = {"value": ft[1] / (f_ms / 1e3), "unit": UNIT, "steps": MAX_ITER, "ms_per_step": f_ms / MAX_ITER,
                                         "note": "force = true, 30 iterations from one straight-line start (round-1 definition of the headline)"}

    # ---------------- end to end through the host-language loop with pinned host buffers; per-kernel CUDA events
    kernel_ms = {"linearize": [], "solve": [], "evaluate": [], "accept": []}
    newton = []
    e2e = None
    if not args.no_e2e:
        nx, nu, N = bp.model.x_dim, bp.model.u_dim, bp.N
        pin = lambda *s: torch.empty(s, dtype=torch.float64).pin_memory()
        hb = {"Xh": pin(B, N, nx), "Uh": pin(B, N, nu), "Xc": pin(B, N, nx), "Uc": pin(B, N, nu), "X0": pin(B, N, nx), "U0": pin(B, N, nu)}
        hb["X0"].numpy()[...] = X0; hb["U0"].numpy()[...] = U0      # every solve starts from this pinned copy (uploaded in its first step)
        om, de = pin(B), pin(B)
        oute, infoe = pin(B, host.EVAL_NOUT), pin(B, host.SOLVE_NINFO)
        act8 = torch.empty(B, dtype=torch.uint8).pin_memory()
        st8 = {}

        def e2e_reset():
            st8.update(Delta=np.full(B, sp[0]), omega=np.full(B, sp[1]), iters=np.zeros(B, np.int64), conv=np.zeros(B),
                       active=np.ones(B, bool), k=0)

        def e2e_step(record):
            om.numpy()[...] = st8["omega"]; de.numpy()[...] = st8["Delta"]; act8.numpy()[...] = st8["active"]
            # H2D: this step's accepted trajectory, penalties, active flags; kernels; D2H: scalars + the step's result (candidate)
            first = st8["k"] == 0
            xin, uin = (hb["X0"], hb["U0"]) if first else (hb["Xh"], hb["Uh"])
            eng.iterate_host(xin.numpy(), uin.numpy(), om.numpy(), de.numpy(), act8.numpy(), oute.numpy(), infoe.numpy(),
                             hb["Xc"].numpy(), hb["Uc"].numpy())
            o = oute.numpy()
            s = host.gusto_update(o, host.solver_status_ok(infoe.numpy()[:, 0]), st8["active"], st8["Delta"], st8["omega"], st8["iters"],
                                  st8["conv"], sp, False)
            acc = s["accept"]
            if acc.all():                                               # every candidate accepted: swap the pinned buffers
                hb["Xh"], hb["Xc"] = hb["Xc"], hb["Xh"]; hb["Uh"], hb["Uc"] = hb["Uc"], hb["Uh"]
            else:
                if first:                                               # the accepted trajectory is still the initial one
                    hb["Xh"].numpy()[...] = hb["X0"].numpy(); hb["Uh"].numpy()[...] = hb["U0"].numpy()
                np.copyto(hb["Xh"].numpy(), hb["Xc"].numpy(), where=acc[:, None, None])
                np.copyto(hb["Uh"].numpy(), hb["Uc"].numpy(), where=acc[:, None, None])
            e2e_cnt[0] += int(s["run"].sum())
            st8.update(Delta=s["Delta"], omega=s["omega"], iters=s["iterations"], conv=np.where(s["run"], o[:, 0], st8["conv"]),
                       active=st8["active"] & ~s["done"], k=st8["k"] + 1)
            if record:
                ms = eng.kernel_ms()
                for k in ("linearize", "solve", "evaluate"):
                    kernel_ms[k].append(ms[k])
                newton.append(float(infoe.numpy()[st8["active"] | s["done"], 1].mean()) if (st8["active"] | s["done"]).any() else 0.0)
            _, unfinished = eng.allgather_status(~st8["active"])       # the path's one collective (in-library NCCL)
            return unfinished

        e2e_cnt = [0]

        def e2e_steps(nsteps, record=False):
            left = nsteps
            while left > 0:
                e2e_reset()
                while left > 0 and st8["k"] < MAX_ITER:
                    left -= 1
                    if e2e_step(record) == 0:
                        break

        e2e_steps(max(3, args.warmup))
        e2e_cnt[0] = 0
        te_ms, _ = timed(lambda: e2e_steps(args.steps, True))
        ec = allsum([e2e_cnt[0]])
        h2d = 8 * (B * N * (nx + nu) + 2 * B) + B
        d2h = 8 * (B * N * (nx + nu) + B * (host.EVAL_NOUT + host.SOLVE_NINFO))
        e2e = {"value": ec[0] / (te_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": te_ms / args.steps, "steps": args.steps,
               "note": "host-language outer loop over the C ABI (gusto_iterate_host: upload trajectory + penalties, three kernels, download "
                       "scalars + candidate; then gusto_allgather_status), pinned host buffers, same real-solve iterations as `value`"}

    # ---------------- hard tier of the same workload (reported separately with its convergence rate)
    if not args.no_extras and args.config == "astrobeeSE3" and world == 1:
        bh = make_problem(pkg, args.config, B, args.knots, Btot, hard=True)
        eh = host.Engine(bh, device=local)
        host.solve_gusto_batch_device(eh, max_iter=MAX_ITER)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Sh = host.solve_gusto_batch_device(eh, max_iter=MAX_ITER)
        th = time.perf_counter() - t0
        eh.close()
        solved = int(Sh.counters[:, 1].sum())
        extras["c3_hard"] = {"workload": bh.name, "value": solved / th, "unit": UNIT, "seconds": th, "batch_iterations": Sh.batch_iterations,
                             "converged": int(Sh.converged.sum()), "successful": int(Sh.successful.sum()), "instances": B,
                             "convergence_rate": float(Sh.successful.mean()), "mean_scp_iterations": float(Sh.iterations.mean()),
                             "trajectories_per_sec": B / th,
                             "note": "endpoints anywhere in the ISS corner (no line of sight): omega escalation and rejected steps occur"}

    if rank == 0:
        ab = algorithmic_bytes(bp)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            pk = json.load(open(peaks_path))
            peak, peak_src = float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kern = {}
        for k in ("linearize", "solve", "evaluate"):
            ms = float(np.mean(kernel_ms[k])) if kernel_ms[k] else 0.0
            gbs = (ab[k] * B / 1e9) / (ms / 1e3) if ms > 0 else 0.0
            kern[k] = {"ms": ms, "algorithmic_bytes": ab[k] * B, "achieved_gbs": gbs, "frac": gbs / peak}
        dom = max(kern, key=lambda k: kern[k]["ms"])
        # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of this workload
        # (profiles/r02c_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum); null when the workload differs
        traffic = None
        for tname in ("r02c_traffic.json", "r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                if tj.get("workload") == f"{bp.model.name} B={B} N={bp.N}":
                    traffic = tj.get(f"{dom}_kernel")
                break
        step_kernel_ms = sum(kern[k]["ms"] for k in kern)
        full = max(1, stats["full_solves"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{bp_all.name} (BASELINE configs[{2 if world == 1 else 3}]), {B} instances/GPU, GuSTO SCP iterations of real solves "
                                   f"(force = false), every solve restarted from the straight-line initialisation",
                       "model": bp.model.name, "B_total": Btot, "B_per_gpu": B, "N": bp.N, "n_obs": int(bp.obstacle_table()[0].shape[0]),
                       "l2": "working set (blocks + solver scratch) >> 126 MB L2, no flush needed",
                       "solver": "structured primal-dual IPM (Riccati recursion), FP64",
                       "outer_loop": "device-resident (gusto_scp_run), one CUDA graph per iteration",
                       "parallelism": f"batch-sharded x{world}, 1 in-library status all-gather (NCCL) per iteration"},
            "clocks": clk,
            "gpu_launches": int(launches),
            "e2e": e2e,
            "roofline": {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": kern[dom]["achieved_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                         "share_of_step_kernel_time": kern[dom]["ms"] / step_kernel_ms if step_kernel_ms else None,
                         "note": "solve kernel is latency/FP64-bound (sequential Riccati sweeps), see DESIGN.md section 5; kernel times are "
                                 "CUDA events on the context stream over the e2e pass (the graph-replayed value pass has no events inside)"},
            "kernels": kern,
            "step_kernel_ms": step_kernel_ms,
            "value_vs_kernel_sum": (t_ms / args.steps) / step_kernel_ms if step_kernel_ms else None,
            "newton_iters_per_solve": float(np.mean(newton)) if newton else None,
            "instance_iterations": {"ran": int(tot[0]), "solved": int(tot[1]), "accepted": int(tot[2]),
                                    "solve_success_fraction": float(tot[1] / max(1.0, tot[0]))},
            "trajectories_per_sec": tot[3] / (t_ms / 1e3),
            "full_solve": {"solves": stats["full_solves"], "instances": Btot, "converged_per_solve": tot[4] / full,
                           "successful_per_solve": tot[5] / full,
                           "batch_iterations": float(np.mean(stats["iters_per_solve"])) if stats["iters_per_solve"] else None},
        }
        line.update(extras)
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            width = args.cpu_sample or cores
            v, done, wall_c, solves, conv, st = cpu_steps("port", args.config, args.knots, Btot, Btot, width, 30, 0, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{st} steps of real solves of {solves} instances of the same batch ({done} instance-iterations in "
                                              f"{wall_c:.1f} s, {conv} converged), NumPy/SciPy oracle, one process per core"}
            v2, d2, w2, s2, c2, st2 = cpu_steps("compiled", args.config, args.knots, Btot, Btot, 8 * width, 30, 3, cores)
            line["cpu_baseline_compiled"] = {"value": v2, "unit": UNIT, "cores": cores, "kind": "compiled",
                                             "sample": f"{st2} steps of real solves of {s2} instances ({d2} instance-iterations in {w2:.1f} s, {c2} converged): "
                                                       "the kernel sources compiled for the host (g++ -O3 -march=native), one process per core"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
