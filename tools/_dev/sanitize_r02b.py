"""compute-sanitizer target: one small GuSTO solve and one small TrajOpt solve per model through the C ABI."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
for name, kw in [("astrobeeSE3", dict(B=9, N=50)), ("freeflyerSE2", dict(B=5, N=40)), ("astrobeeSE3manifold", dict(B=3, N=60)), ("dubins", dict(B=4, N=30))]:
    bp = pkg.problems.CONFIGS[name](**kw)
    eng = host.Engine(bp)
    S = host.solve_gusto_batch_device(eng, max_iter=4)
    print(name, "gusto iterations", S.iterations.tolist(), flush=True)
    if name in ("astrobeeSE3", "freeflyerSE2"):
        T = host.solve_trajopt_batch(eng)
        print(name, "trajopt solves", T.batch_solves, "iterations", T.iterations.tolist(), flush=True)
    eng.close()
