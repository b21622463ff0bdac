#!/usr/bin/env python
"""GPU probe: per-kernel milliseconds and the IPM kernel's per-phase SM-cycle counters for one outer iteration.

GUSTO_PROBE_LIB=<path to a developer build of the library, e.g. compiled with -DGUSTO_PROF_MODE=1> selects another .so;
the meaning of the three counters then follows ipm.cuh (GUSTO_PROF_MODE)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
alt = os.environ.get("GUSTO_PROBE_LIB")
if alt:
    pkg = entry.load_package(); host = pkg.engine(); host.load_library(alt)
else:
    pkg = entry.build(); host = pkg.engine()
name = sys.argv[1] if len(sys.argv) > 1 else "astrobeeSE3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bp = pkg.problems.CONFIGS[name](B=B)
eng = host.Engine(bp)
X0, U0 = bp.init_traj_straightline()
eng.set_trajectory(X0, U0)
for rep in range(reps):
    out, info = eng.iterate()
    ms = eng.kernel_ms()
    it = info[:, 1]
    cyc = info[:, 5:8]
    print(f"{name} B={B} rep{rep}: ms lin {ms['linearize']:.3f} solve {ms['solve']:.3f} eval {ms['evaluate']:.3f} | newton mean {it.mean():.2f} max {it.max():.0f} | "
          f"ok {int((info[:,0]==0).sum())}/{B} | info3/iter {np.mean(info[:,3]/it):.3g} | cycles/newton-iter c5 {np.mean(cyc[:,0]/it):.0f} c6 {np.mean(cyc[:,1]/it):.0f} c7 {np.mean(cyc[:,2]/it):.0f}")
eng.close()
