"""Developer probe: Newton-iteration counts of the convex solves along real SCP runs, on the CPU build of the kernel sources
(oracle/compiled_baseline.py).  usage: python tools/newton_probe.py [config] [n_instances] [hard]"""
import ctypes
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as entry
import compiled_baseline as cb


def run(name="astrobeeSE3", n=16, hard=False, lib=None, max_iter=30, seed=1024, Btot=1024):
    pkg = entry.load_package()
    host = pkg.engine()
    kw = dict(B=Btot, seed=seed)
    if hard:
        kw["hard"] = True
    bp_all = pkg.problems.CONFIGS[name](**kw)
    if lib:
        cb._LIB = ctypes.CDLL(lib)
    out = []
    for idx in range(n):
        bp = bp_all.instance(idx)
        cfg, (kind, a, b) = host.make_config(bp, 0)
        B, N, nx, nu = 1, bp.N, bp.model.x_dim, bp.model.u_dim
        no = int(kind.shape[0]) if bp.model.model_id != pkg.models.DUBINS else 0
        dp = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        sp = bp.model.scp_params
        X0, U0 = bp.init_traj_straightline()
        Xp = X0.copy(); Up = U0.copy(); Xn = np.zeros((B, N, nx)); Un = np.zeros((B, N, nu))
        f = np.zeros((B, N, nx)); A = np.zeros((B, N, nx, nx)); g = np.zeros((B, N, nx)); rows = np.zeros((B, N, max(no, 1), 5))
        info = np.zeros((B, 8)); ev = np.zeros((B, 8))
        x_init = np.ascontiguousarray(bp.x_init); glo = np.ascontiguousarray(bp.goal_lo); ghi = np.ascontiguousarray(bp.goal_hi); tf = np.ascontiguousarray(bp.tf)
        Delta = np.full(B, sp[0]); omega = np.full(B, sp[1]); its = np.zeros(B, np.int64); cprev = np.zeros(B); active = np.ones(B, bool)
        conv = False; newt = []; stat = []; acc_h = []
        lib_ = cb._lib()
        if hasattr(lib_, "hostsim_reset_warm"):
            lib_.hostsim_reset_warm()
        for _ in range(max_iter):
            om = omega.copy(); de = Delta.copy()
            lib_.hostsim_iterate(ctypes.byref(cfg), kind.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), dp(a), dp(b), dp(x_init), dp(glo), dp(ghi), dp(tf),
                                 dp(Xp), dp(Up), dp(Xn), dp(Un), dp(om), dp(de), dp(f), dp(A), dp(g), dp(rows), ctypes.c_int(7), dp(info), dp(ev))
            newt.append(int(info[0, 1])); stat.append(int(info[0, 0]))
            st = host.gusto_update(ev, host.solver_status_ok(info[:, 0]), active, Delta, omega, its, cprev, sp, False)
            acc = st["accept"]; acc_h.append(bool(acc[0]))
            Xp[acc] = Xn[acc]; Up[acc] = Un[acc]
            cprev = np.where(st["run"], ev[:, 0], cprev)
            Delta, omega, its = st["Delta"], st["omega"], st["iterations"]
            conv = conv or bool(st["converged_now"][0])
            active = active & ~st["done"]
            if not active.any():
                break
        out.append(dict(newton=newt, status=stat, accept=acc_h, iterations=int(its[0]), converged=conv, J=float(ev[0, 4])))
    return out


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "astrobeeSE3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    hard = len(sys.argv) > 3 and sys.argv[3] == "hard"
    lib = sys.argv[4] if len(sys.argv) > 4 else None
    res = run(name, n, hard, lib)
    tot = sum(sum(r["newton"]) for r in res); cnt = sum(len(r["newton"]) for r in res)
    for i, r in enumerate(res[:12]):
        print(i, r["newton"], r["status"], "conv" if r["converged"] else "NOT", f"J={r['J']:.6f}")
    print(f"{name}{' hard' if hard else ''}: {n} instances, {cnt} solves, mean newton {tot / cnt:.2f}, converged {sum(r['converged'] for r in res)}/{n}, "
          f"mean SCP iterations {np.mean([r['iterations'] for r in res]):.2f}")
