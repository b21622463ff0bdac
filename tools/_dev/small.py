import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
pkg = entry.load_package(); host = pkg.engine(); host.load_library(os.environ["GUSTO_PROBE_LIB"])
name = sys.argv[1]; B=int(sys.argv[2])
bp = pkg.problems.CONFIGS[name](B=B)
eng = host.Engine(bp)
X0, U0 = bp.init_traj_straightline()
eng.set_trajectory(X0, U0)
out, info = eng.iterate()
print(info[:, :5])
eng.close()
