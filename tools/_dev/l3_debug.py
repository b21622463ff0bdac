import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import __graft_entry__ as entry
pkg = entry.build(); host = pkg.engine()
name = sys.argv[1]; B = int(sys.argv[2]); N = int(sys.argv[3])
bp = pkg.problems.CONFIGS[name](B=B, N=N)
eng = host.Engine(bp)
S = host.solve_gusto_batch(eng, max_iter=30)
for b in range(B):
    print(b, "its", int(S.iterations[b]), "conv", bool(S.converged[b]), "succ", bool(S.successful[b]),
          "solver", [int(s[b]) for s in S.solver_status[1:]], "newton", [int(n[b]) for n in S.newton_iters],
          "scp", [int(s[b]) for s in S.scp_status[1:]])
