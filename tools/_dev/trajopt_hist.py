"""Newton-iteration histogram and kernel time of the TrajOpt subproblem kernel over the launches of a batched solve (dev probe)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
bp = pkg.problems.CONFIGS["astrobeeSE3"](B=1024, N=50)
eng = host.Engine(bp); eng.trajopt_enable()
X0, U0 = bp.init_traj_straightline()
prm = pkg.models.TRAJOPT_PARAMS[bp.model.model_id]
for mu, s in [(prm[0], prm[1]), (prm[0], prm[1] * 0.5), (prm[0] * 10, prm[1])]:
    for rep in range(2):
        eng.set_trajectory(X0, U0)
        ev, info = eng.trajopt_iterate(np.full(bp.B, mu), np.full(bp.B, s))
    ms = eng.kernel_ms()
    it = info[:, 1].astype(int)
    print(f"mu {mu:g} s {s:g}: solve {ms['solve']:.3f} ms, newton mean {it.mean():.2f} max {it.max()}, hist {np.bincount(it).tolist()}", flush=True)
