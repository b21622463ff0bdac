import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
bp = pkg.problems.CONFIGS["astrobeeSE3"](B=1024)
eng = host.Engine(bp)
X0, U0 = bp.init_traj_straightline()
host.solve_gusto_batch(eng, X0, U0)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
t = time.perf_counter(); S = host.solve_gusto_batch(eng, X0, U0); dt = time.perf_counter() - t
pr.disable()
print("total", dt, "batch its", S.batch_iterations, "iter times", S.iter_elapsed_times)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
