#!/usr/bin/env python
"""Full batched GuSTO solves (device-resident loop gusto_scp_run, max_iter 30) of every BASELINE.json configuration that fits
one GPU, plus the hard tier of configs[2] and the literal-C5 tier of configs[4]: wall time after a warm-up solve, converged /
successful counts, SCP iterations, trajectories/s and instance-iterations/s."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
CFG = [("dubins", dict(B=1024, N=30), "configs[0] dubins N=30 (batched here; the reference runs one instance on the CPU)"),
       ("freeflyerSE2", dict(B=256, N=40), "configs[1] freeflyerSE2 B=256 N=40"),
       ("astrobeeSE3", dict(B=1024, N=50), "configs[2] astrobeeSE3 B=1024 N=50"),
       ("astrobeeSE3", dict(B=1024, N=50, hard=True), "configs[2] hard tier (endpoints anywhere in zones 8 / 5, no line of sight)"),
       ("astrobeeSE3manifold", dict(B=1024, N=60), "configs[4] astrobeeSE3manifold B=1024 N=60 (notebook endpoints, BoxGoal q +- 1e-4, 32 obstacles)"),
       ("astrobeeSE3manifold", dict(B=1024, N=60, tier="zones"), "configs[4] 'zones' tier (SURVEY C5 read literally: zone 8 -> zone 5)")]
print("| configuration | wall s | converged | successful | SCP iterations (mean / max) | batch iterations | trajectories/s | instance-iterations/s | Newton iterations per solve | solver statuses over all solves (OPTIMAL / ITERATION_LIMIT / NUMERICAL / ALMOST_OPTIMAL) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for name, kw, label in CFG:
    bp = pkg.problems.CONFIGS[name](**kw)
    eng = host.Engine(bp)
    host.solve_gusto_batch_device(eng, max_iter=30)
    t = time.perf_counter(); S = host.solve_gusto_batch_device(eng, max_iter=30); dt = time.perf_counter() - t
    its = int(S.iterations.sum())
    nw = np.concatenate([x[x > 0] for x in S.newton_iters]) if S.newton_iters else np.zeros(1)
    print(f"| {label} | {dt:.4f} | {int(S.converged.sum())}/{bp.B} | {int(S.successful.sum())}/{bp.B} | {S.iterations.mean():.2f} / {S.iterations.max()} | "
          f"{S.batch_iterations} | {bp.B / dt:.0f} | {its / dt:.0f} | {nw.mean():.2f} | "
          + " / ".join(str(int(sum((st == c).sum() for st in S.solver_status[1:]))) for c in (0, 1, 2, 3)) + " |", flush=True)
    eng.close()
