import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
name = sys.argv[1] if len(sys.argv) > 1 else "astrobeeSE3manifold"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
bp = pkg.problems.CONFIGS[name](B=B)
opts = {}
if os.environ.get('NARROW_BOX_AS_POINT'):
    w = bp.goal_hi - bp.goal_lo
    narrow = (bp.goal_type == 2) & np.all(w < 1e-3, axis=0)
    mid = 0.5 * (bp.goal_lo + bp.goal_hi)
    bp.goal_type = np.where(narrow, 1, bp.goal_type).astype(bp.goal_type.dtype)
    bp.goal_lo = np.where(narrow[None, :], mid, bp.goal_lo); bp.goal_hi = np.where(narrow[None, :], mid, bp.goal_hi)
    print('narrow boxes -> point goals:', narrow)
if len(sys.argv) > 3: opts['ipm_delta_p'] = float(sys.argv[3])
if len(sys.argv) > 4: opts['ipm_nref'] = int(sys.argv[4])
if len(sys.argv) > 5: opts['ipm_tol'] = float(sys.argv[5])
eng = host.Engine(bp, **opts)
S = host.solve_gusto_batch(eng, max_iter=30)
print(name, opts, "B", B, "converged", int(S.converged.sum()), "successful", int(S.successful.sum()), "iterations mean", S.iterations.mean(), "max", S.iterations.max())
st = np.array(S.solver_status)
print("solver status counts (0 ok, 1 iter limit, 2 numerical, -1 inactive):", {int(v): int((st == v).sum()) for v in np.unique(st)})
print("newton mean", np.mean([x[x>0].mean() for x in S.newton_iters if (x>0).any()]))
print("omega max per instance:", np.unique(np.max(np.array(S.omega_vec), axis=0), return_counts=True) if hasattr(S, "omega_vec") else "n/a")
eng.close()
import numpy as np
S2 = S
print("conv measure inst0:", [float(f"{c[0]:.2e}") for c in S2.convergence_measure[:14]])
print("newton inst0:", [int(x[0]) for x in S2.newton_iters[:14]])
print("accept inst0:", [bool(a[0]) for a in S2.accept_solution[:14]])
print("rho inst0:", [float(f"{r[0]:.2e}") for r in S2.rho_vec[:14]] if hasattr(S2, "rho_vec") else None)
