import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
pkg = entry.load_package(); host = pkg.engine(); host.load_library(os.environ.get("GUSTO_PROBE_LIB", host.LIB_PATH))
bp = pkg.problems.CONFIGS["astrobeeSE3"](B=1024)
X0, U0 = bp.init_traj_straightline()
for nref in (2, 1):
    eng = host.Engine(bp, ipm_nref=nref)
    eng.set_trajectory(X0, U0)
    for rep in range(2):
        out, info = eng.iterate()
    ms = eng.kernel_ms()
    print(f"nref={nref}: solve {ms['solve']:.3f} ms newton mean {info[:,1].mean():.2f} ok {int((info[:,0]==0).sum())}")
    eng.close()
