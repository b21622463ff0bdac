#!/usr/bin/env python
"""GPU probe of the shooting kernel (K7): device time of one attempt over a batch, after a few forced GuSTO iterations."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
name = sys.argv[1] if len(sys.argv) > 1 else "astrobeeSE3manifold"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
bp = pkg.problems.CONFIGS[name](B=B)
eng = host.Engine(bp)
X0, U0 = bp.init_traj_straightline()
eng.set_trajectory(X0, U0)
for _ in range(warm):
    eng.iterate(); eng.accept(np.ones(B, np.uint8))
xg = 0.5 * (bp.goal_lo + bp.goal_hi)
for rep in range(3):
    eng.timer_start()
    out = eng.shoot(None, xg)
    ms = eng.timer_stop()
    ok = out[:, 0] == 0
    print(f"{name} B={B} N={bp.N} rep{rep}: shoot {ms:.3f} ms (incl. D2H of {out.nbytes} B) | optimal {int(ok.sum())}/{B} | LM iters mean {out[:,1].mean():.2f} max {out[:,1].max():.0f} "
          f"| |F| max over optimal {out[ok,2].max() if ok.any() else float('nan'):.2e} | {B / ms * 1e3:.0f} attempts/s")
eng.close()
