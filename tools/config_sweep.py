#!/usr/bin/env python
"""Full batched GuSTO solves (solve_gusto_batch, max_iter 30) of every BASELINE.json configuration that fits one GPU:
wall time after a warm-up solve, converged / successful counts, SCP iterations, trajectories/s and instance-iterations/s."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
CFG = [("dubins", dict(B=1024, N=30), "configs[0] dubins N=30 (batched here; the reference runs one instance on the CPU)"),
       ("freeflyerSE2", dict(B=256, N=40), "configs[1] freeflyerSE2 B=256 N=40"),
       ("astrobeeSE3", dict(B=1024, N=50), "configs[2] astrobeeSE3 B=1024 N=50"),
       ("astrobeeSE3manifold", dict(B=1024, N=60), "configs[4] astrobeeSE3manifold B=1024 N=60")]
print("| configuration | wall s | converged | successful | SCP iterations (mean / max) | batch iterations | trajectories/s | instance-iterations/s | Newton iterations per solve |")
print("|---|---|---|---|---|---|---|---|---|")
for name, kw, label in CFG:
    bp = pkg.problems.CONFIGS[name](**kw)
    eng = host.Engine(bp)
    host.solve_gusto_batch(eng, max_iter=30)
    t = time.perf_counter(); S = host.solve_gusto_batch(eng, max_iter=30); dt = time.perf_counter() - t
    its = int(S.iterations.sum())
    nw = np.concatenate([x[x > 0] for x in S.newton_iters]) if S.newton_iters else np.zeros(1)
    print(f"| {label} | {dt:.4f} | {int(S.converged.sum())}/{bp.B} | {int(S.successful.sum())}/{bp.B} | {S.iterations.mean():.2f} / {S.iterations.max()} | "
          f"{S.batch_iterations} | {bp.B / dt:.0f} | {its / dt:.0f} | {nw.mean():.2f} |", flush=True)
    eng.close()
