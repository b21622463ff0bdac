#!/usr/bin/env python
"""GPU probe: per-kernel milliseconds and the IPM kernel's per-phase SM-cycle counters for one outer iteration."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
name = sys.argv[1] if len(sys.argv) > 1 else "astrobeeSE3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
bp = pkg.problems.CONFIGS[name](B=B)
eng = host.Engine(bp)
X0, U0 = bp.init_traj_straightline()
eng.set_trajectory(X0, U0)
for rep in range(3):
    out, info = eng.iterate()
    ms = eng.kernel_ms()
    it = info[:, 1]
    cyc = info[:, 5:8]
    print(f"{name} B={B} rep{rep}: ms {ms} | newton mean {it.mean():.2f} max {it.max():.0f} | status ok {int((info[:,0]==0).sum())}/{B} | "
          f"cycles/newton-iter (mean over CTAs): assemble+slots {np.mean(cyc[:,0]/it):.0f} factorize {np.mean(cyc[:,1]/it):.0f} kkt-solves {np.mean(cyc[:,2]/it):.0f}")
eng.close()
