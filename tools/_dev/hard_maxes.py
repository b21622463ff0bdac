import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
bp = pkg.problems.CONFIGS["astrobeeSE3"](B=1024, hard=True)
eng = host.Engine(bp)
S = host.solve_gusto_batch_device(eng, max_iter=30)
mx = [int(x.max()) for x in S.newton_iters]
mean = float(np.concatenate([x[x > 0] for x in S.newton_iters]).mean())
print("hard tier B=1024: batch iterations", S.batch_iterations, "per-launch max newton", mx, "sum", sum(mx), "mean", round(mean, 2), "converged", int(S.converged.sum()))
