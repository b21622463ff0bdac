import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
bp = pkg.problems.CONFIGS["astrobeeSE3"](B=3, N=30)
eng = host.Engine(bp)
S = host.solve_gusto_batch_device(eng, max_iter=2)
print("gusto", S.iterations.tolist(), flush=True)
eng.trajopt_enable()
X0, U0 = bp.init_traj_straightline()
eng.set_trajectory(X0, U0)
ev, info = eng.trajopt_iterate(np.full(3, 1.0), np.full(3, 10.0))
print("trajopt", info[:, :2].tolist(), flush=True)
eng.close()
