import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
name = sys.argv[1]; B = int(sys.argv[2]); N = int(sys.argv[3]); idx = int(sys.argv[4])
bp = pkg.problems.CONFIGS[name](B=B, N=N)
eng = host.Engine(bp)
X0, U0 = bp.init_traj_straightline()
eng.set_trajectory(X0, U0)
sp = bp.model.scp_params
Delta = np.full(B, sp[0]); omega = np.full(B, sp[1]); its = np.zeros(B, np.int64); cprev = np.zeros(B); active = np.ones(B, bool)
for it in range(12):
    eng.set_active(active)
    out, info = eng.iterate()
    print(it, "omega", omega[idx], "Delta", Delta[idx], "info", info[idx, :5], "eval", out[idx, :4])
    st = host.gusto_update(out, host.solver_status_ok(info[:, 0]), active, Delta, omega, its, cprev, sp)
    eng.accept(st["accept"], st["omega"], st["Delta"])
    cprev = np.where(st["run"], out[:, 0], cprev)
    Delta, omega, its = st["Delta"], st["omega"], st["iterations"]
    active = active & ~st["done"]
    if not active[idx]:
        break
