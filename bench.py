#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched GuSTO SCP hot path (BASELINE.json: SCP iterations/s and
trajectories/s, astrobeeSE3 B=1024 N=50, 1/2/4/8 GPUs, next to the CPU restatement on the same host).

  python bench.py --gpus N --steps K --warmup W            one rank per GPU (torchrun for N > 1), weak scaling
  python bench.py --impl reference --gpus N --steps K ...  the CPU restatement (oracle) on the host cores

A "step" is one outer SCP iteration over the whole batch with every instance live (the reference's `force=true`,
scp_gusto.jl:55,173): linearize (K1+K2) -> convex solve (K3) -> evaluate (K4) -> host accept/reject + Delta/omega
schedule (scp_gusto.jl:119-174) -> accept.  After 30 iterations (solve_SCP!'s max_iter, traj_opt.jl:47) the batch is
reset to its straight-line initialisation.  `value` counts instance-iterations per second with the trajectories
resident in HBM (only the 2x8 scalars per instance cross PCIe for the host-side decision); `e2e` runs the same step
through the public host API with the trajectory uploaded from / downloaded to pinned host buffers every step.
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
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="astrobeeSE3")
    ap.add_argument("--batch", type=int, default=1024, help="instances per GPU")
    ap.add_argument("--knots", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def make_problem(pkg, name, B, N, seed):
    fn = pkg.problems.CONFIGS[name]
    kw = dict(B=B, seed=seed)
    if N:
        kw["N"] = N
    return fn(**kw)


def algorithmic_bytes(bp):
    """Bytes one (instance, SCP iteration) must move at minimum, per kernel."""
    nx, nu, N = bp.model.x_dim, bp.model.u_dim, bp.N
    no = int(bp.obstacle_table()[0].shape[0])
    traj = N * (nx + nu)
    blocks = N * (nx * nx + 2 * nx)          # A, f, g
    rows = N * no * 5
    k12 = traj + blocks + rows               # read traj, write blocks + rows
    k3 = traj + N * (nx * nx + nx) + rows + traj   # read Xp/Up, A, g, rows; write candidate
    k4 = 2 * traj + N * (nx * nx + nx) + rows      # read both trajectories, A, f, rows
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


# ------------------------------------------------------------------------------------------ CPU restatement
def _cpu_worker(args):
    name, N, seed, Btot, idx, iters = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    pkg = entry.load_package()
    import gusto_oracle as orc
    from gusto_oracle.subproblem import Problem
    bp = make_problem(pkg, name, Btot, N, seed)
    m = orc.get_model(bp.model.name)
    p = Problem(m, bp.N, float(bp.tf[idx]), bp.x_init[idx], bp.goal_type, bp.goal_lo[idx], bp.goal_hi[idx], bp.obstacle_table())
    t = time.perf_counter()
    S = orc.solve_gusto(p, max_iter=iters, force=True)
    return S.iterations, time.perf_counter() - t


def cpu_baseline(name, N, seed, Btot, n_sample, iters, cores):
    """Oracle (kind 'port') on `cores` worker processes: n_sample instances x `iters` forced SCP iterations each."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    jobs = [(name, N, seed, Btot, i, iters) for i in range(n_sample)]
    t = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t
    done = sum(r[0] for r in res)
    return done / wall, done, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    Btot = args.batch * args.gpus
    n_sample = args.cpu_sample or min(Btot, 48 * cores)
    N = args.knots
    # one "step" = one forced SCP iteration over the bounded sample
    val, done, wall = cpu_baseline(args.config, N, Btot, Btot, n_sample, max(1, min(args.steps, 3)), cores)
    pkg = entry.load_package()
    bp = make_problem(pkg, args.config, 2, N, Btot)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, min(args.steps, 3)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config} B={Btot} N={bp.N} (GuSTO SCP iteration, CPU restatement on a {n_sample}-instance sample)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_sample} instances x {max(1, min(args.steps, 3))} forced SCP iterations ({done} instance-iterations in {wall:.1f} s); "
                                   "the reference's Julia/JuMP/Gurobi path cannot run here (no Julia, no solver licences)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
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

    flags_dev = torch.zeros(B, dtype=torch.uint8, device="cuda")
    flags_all = torch.zeros(B * world, dtype=torch.uint8, device="cuda") if world > 1 else None

    def allgather_status(done_local):
        """The path's only collective: one all-gather of per-instance status bytes per outer iteration (NCCL)."""
        if world == 1:
            return bool(done_local.all())
        flags_dev.copy_(torch.from_numpy(done_local.astype(np.uint8)), non_blocking=False)
        dist.all_gather_into_tensor(flags_all, flags_dev)
        return bool(flags_all.all().item())

    state = {}

    def reset():
        eng.set_trajectory(X0, U0)
        state["Delta"] = np.full(B, sp[0]); state["omega"] = np.full(B, sp[1])
        state["iters"] = np.zeros(B, np.int64); state["conv_prev"] = np.zeros(B); state["k"] = 0
        eng.set_penalties(state["omega"], state["Delta"])
        eng.set_active(np.ones(B, np.uint8))

    out = np.empty((B, host.EVAL_NOUT)); info = np.empty((B, host.SOLVE_NINFO))
    active = np.ones(B, bool)
    kernel_ms = {"linearize": [], "solve": [], "evaluate": [], "accept": []}
    newton = []

    def step(record=False):
        if state["k"] >= MAX_ITER:
            reset()
        eng.iterate(out, info)
        st = host.gusto_update(out, host.solver_status_ok(info[:, 0]), active, state["Delta"], state["omega"], state["iters"],
                               state["conv_prev"], sp, force=True)
        eng.accept(st["accept"], st["omega"], st["Delta"])
        state["Delta"], state["omega"], state["iters"] = st["Delta"], st["omega"], st["iterations"]
        state["conv_prev"] = out[:, 0].copy(); state["k"] += 1
        allgather_status(~active)
        if record:
            ms = eng.kernel_ms()
            for k in kernel_ms:
                kernel_ms[k].append(ms[k])
            newton.append(float(info[:, 1].mean()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    reset()
    for _ in range(args.warmup):
        step()
    # -------- timed region: device-resident trajectories
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count()
    barrier()
    eng.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(record=True)
    dev_ms = eng.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    t_ms = max(dev_ms, 1e3 * wall)        # the device stopwatch spans the same steps; take the slower of the two clocks
    if world > 1:
        tt = torch.tensor([t_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt.item())
    value = Btot * args.steps / (t_ms / 1e3)

    # -------- end to end through the host API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        nx, nu, N = bp.model.x_dim, bp.model.u_dim, bp.N
        Xh = torch.empty((B, N, nx), dtype=torch.float64).pin_memory(); Uh = torch.empty((B, N, nu), dtype=torch.float64).pin_memory()
        Xc = torch.empty((B, N, nx), dtype=torch.float64).pin_memory(); Uc = torch.empty((B, N, nu), dtype=torch.float64).pin_memory()
        Xh.numpy()[...] = X0; Uh.numpy()[...] = U0
        hb = {"Xh": Xh, "Uh": Uh, "Xc": Xc, "Uc": Uc}
        om = torch.empty(B, dtype=torch.float64).pin_memory(); de = torch.empty(B, dtype=torch.float64).pin_memory()
        oute = torch.empty((B, host.EVAL_NOUT), dtype=torch.float64).pin_memory()
        infoe = torch.empty((B, host.SOLVE_NINFO), dtype=torch.float64).pin_memory()
        st8 = {"Delta": np.full(B, sp[0]), "omega": np.full(B, sp[1]), "iters": np.zeros(B, np.int64), "conv": np.zeros(B), "k": 0}

        def e2e_step():
            if st8["k"] >= MAX_ITER:
                hb["Xh"].numpy()[...] = X0; hb["Uh"].numpy()[...] = U0
                st8.update(Delta=np.full(B, sp[0]), omega=np.full(B, sp[1]), iters=np.zeros(B, np.int64), conv=np.zeros(B), k=0)
            om.numpy()[...] = st8["omega"]; de.numpy()[...] = st8["Delta"]
            eng.set_trajectory(hb["Xh"].numpy(), hb["Uh"].numpy())     # H2D: this step's accepted trajectory
            eng.set_penalties(om.numpy(), de.numpy())                   # H2D
            eng.iterate(oute.numpy(), infoe.numpy())                    # kernels + D2H of the scalars
            eng.get_candidate(hb["Xc"].numpy(), hb["Uc"].numpy())       # D2H: the step's result
            o = oute.numpy()
            s = host.gusto_update(o, host.solver_status_ok(infoe.numpy()[:, 0]), active, st8["Delta"], st8["omega"], st8["iters"], st8["conv"], sp, force=True)
            acc = s["accept"]
            if acc.all():                                               # every candidate accepted: swap the pinned buffers
                hb["Xh"], hb["Xc"] = hb["Xc"], hb["Xh"]; hb["Uh"], hb["Uc"] = hb["Uc"], hb["Uh"]
            elif acc.any():
                np.copyto(hb["Xh"].numpy(), hb["Xc"].numpy(), where=acc[:, None, None])
                np.copyto(hb["Uh"].numpy(), hb["Uc"].numpy(), where=acc[:, None, None])
            st8.update(Delta=s["Delta"], omega=s["omega"], iters=s["iterations"], conv=o[:, 0].copy(), k=st8["k"] + 1)
            allgather_status(~active)

        for _ in range(max(3, args.warmup)):
            e2e_step()
        barrier()
        t1 = time.perf_counter()
        ksteps = args.steps
        for _ in range(ksteps):
            e2e_step()
        barrier()
        te = time.perf_counter() - t1
        if world > 1:
            tt = torch.tensor([te], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt.item())
        h2d = 8 * (B * N * (nx + nu) + 2 * B)
        d2h = 8 * (B * N * (nx + nu) + B * (host.EVAL_NOUT + host.SOLVE_NINFO))
        e2e = {"value": Btot * ksteps / te, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * te / ksteps}

    # -------- trajectories/s: full solves (force = false) from the straight-line initialisation
    barrier()
    t2 = time.perf_counter()
    S = host.solve_gusto_batch(eng, X0, U0, max_iter=MAX_ITER, all_done=allgather_status)
    barrier()
    tsolve = time.perf_counter() - t2
    conv = np.array([S.converged.sum(), S.successful.sum(), S.iterations.sum()], dtype=np.float64)
    if world > 1:
        tt = torch.tensor([tsolve], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tsolve = float(tt.item())
        cc = torch.tensor(conv, device="cuda")
        dist.all_reduce(cc)
        conv = cc.cpu().numpy()

    if rank == 0:
        ab = algorithmic_bytes(bp)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kern = {}
        for k in ("linearize", "solve", "evaluate"):
            ms = float(np.mean(kernel_ms[k])) if kernel_ms[k] else 0.0
            gbs = (ab[k] * B / 1e9) / (ms / 1e3) if ms > 0 else 0.0
            kern[k] = {"ms": ms, "algorithmic_bytes": ab[k] * B, "achieved_gbs": gbs, "frac": gbs / peak}
        dom = max(kern, key=lambda k: kern[k]["ms"])
        # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of this workload
        # (profiles/r01_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum); null when the workload differs
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("workload") == f"{bp.model.name} B={B} N={bp.N}":
                traffic = tj.get(f"{dom}_kernel")
        step_kernel_ms = sum(kern[k]["ms"] for k in kern)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{bp_all.name} (BASELINE configs[{2 if world == 1 else 3}]), {B} instances/GPU, one forced GuSTO SCP iteration per step, "
                                   f"restart from straight-line init every {MAX_ITER} steps",
                       "model": bp.model.name, "B_total": Btot, "B_per_gpu": B, "N": bp.N, "n_obs": int(bp.obstacle_table()[0].shape[0]),
                       "l2": "working set (blocks + solver scratch) >> 126 MB L2, no flush needed",
                       "solver": "structured primal-dual IPM, FP64", "parallelism": f"batch-sharded x{world}, 1 status all-gather/iteration"},
            "clocks": clk,
            "gpu_launches": int(launches),
            "e2e": e2e,
            "roofline": {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": kern[dom]["achieved_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                         "share_of_step_kernel_time": kern[dom]["ms"] / step_kernel_ms if step_kernel_ms else None,
                         "note": "solve kernel is latency/FP64-bound (sequential block-tridiagonal sweeps), see DESIGN.md section 5"},
            "kernels": kern,
            "newton_iters_per_solve": float(np.mean(newton)) if newton else None,
            "trajectories_per_sec": Btot / tsolve,
            "full_solve": {"seconds": tsolve, "converged": int(conv[0]), "successful": int(conv[1]), "instances": Btot,
                           "scp_iterations_total": int(conv[2]), "batch_iterations": S.batch_iterations},
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n_sample = args.cpu_sample or min(Btot, 48 * cores)
            v, done, wall_c = cpu_baseline(args.config, args.knots, Btot, Btot, n_sample, 4, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_sample} instances x 4 forced SCP iterations of the same workload "
                                              f"({done} instance-iterations in {wall_c:.1f} s, NumPy/SciPy oracle, one process per core)"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
