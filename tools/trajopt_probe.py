#!/usr/bin/env python
"""GPU probe of the TrajOpt variant (solve_trajopt_jump!): batched full solves through host.solve_trajopt_batch
(host-language loops, one kernel launch per request kind) and the per-kernel times of one subproblem iteration."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
print("| configuration | convex solves (batch) | wall s | converged (ctol) | solver statuses OPTIMAL / other | Newton iterations per solve | subproblem kernel ms (first iteration) | evaluate ms |")
print("|---|---|---|---|---|---|---|---|")
for name, kw in [("freeflyerSE2", dict(B=256, N=40)), ("astrobeeSE3", dict(B=1024, N=50))]:
    bp = pkg.problems.CONFIGS[name](**kw)
    eng = host.Engine(bp)
    eng.trajopt_enable()
    X0, U0 = bp.init_traj_straightline()
    prm = pkg.models.TRAJOPT_PARAMS[bp.model.model_id]
    eng.set_trajectory(X0, U0)
    eng.trajopt_iterate(np.full(bp.B, prm[0]), np.full(bp.B, prm[1]))
    eng.set_trajectory(X0, U0)
    ev, info = eng.trajopt_iterate(np.full(bp.B, prm[0]), np.full(bp.B, prm[1]))
    ms = eng.kernel_ms()
    t = time.perf_counter(); S = host.solve_trajopt_batch(eng); dt = time.perf_counter() - t
    st = np.concatenate([np.array(s[1:]) for s in S.solver_status])
    nw = np.concatenate([np.array(s) for s in S.newton_iters])
    print(f"| {name} B={bp.B} N={bp.N} | {S.batch_solves} | {dt:.3f} | {int(S.converged.sum())}/{bp.B} | {int((st == 0).sum())} / {int((st != 0).sum())} | {nw.mean():.2f} | "
          f"{ms['solve']:.3f} | {ms['evaluate']:.3f} |", flush=True)
    eng.close()
