#!/usr/bin/env python
"""L3 (end-to-end) parity evidence: full GuSTO SCP runs of the CUDA path vs the CPU oracle, instance by instance.

  GPU box :  python tools/parity_l3.py gpu  <config> <B> <out.npz>        (solve_gusto_batch, max_iter 30)
  CPU     :  python tools/parity_l3.py cpu  <config> <B> <n> <in.npz>     (oracle on the first n instances, compares)
"""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry

mode, name, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
KW = {}
if name.endswith(":hard"):                       # e.g. astrobeeSE3:hard -- the C3-hard tier of the generator
    name, KW = name[:-5], dict(hard=True)
if mode == "gpu":
    pkg = entry.build(); host = pkg.engine()
    bp = pkg.problems.CONFIGS[name](B=B, **KW)
    eng = host.Engine(bp)
    S = host.solve_gusto_batch_device(eng, max_iter=30)          # the device-resident loop (same decisions as the host loop: GPU test)
    np.savez(sys.argv[4], converged=S.converged, successful=S.successful, iterations=S.iterations,
             J_true=np.array(S.J_true)[-1], omega=np.array(S.omega_vec)[-1])
    print(name, "B", B, "converged", int(S.converged.sum()), "successful", int(S.successful.sum()))
    eng.close()
else:
    from util import gb, to_oracle
    from gusto_oracle.scp import solve_gusto
    n = int(sys.argv[4]); G = np.load(sys.argv[5])
    bp = gb.problems.CONFIGS[name](B=B, **KW)
    agree = 0; rows = []
    for b in range(n):
        R = solve_gusto(to_oracle(bp, b), max_iter=30)
        same = (bool(G["converged"][b]) == R.converged and bool(G["successful"][b]) == R.successful and int(G["iterations"][b]) == R.iterations)
        dj = abs(G["J_true"][b] - R.J_true[-1]) / max(1e-12, abs(R.J_true[-1]))
        agree += same
        rows.append((b, R.converged, R.successful, R.iterations, bool(G["converged"][b]), bool(G["successful"][b]), int(G["iterations"][b]), dj))
        print(f"{name} inst {b:3d}: oracle conv={R.converged} succ={R.successful} it={R.iterations:2d} | gpu conv={bool(G['converged'][b])} "
              f"succ={bool(G['successful'][b])} it={int(G['iterations'][b]):2d} | rel dJ_true {dj:.2e} {'' if same else '  <-- differs'}", flush=True)
    print(f"{name}: {agree}/{n} instances with identical (converged, successful, iterations); max rel dJ_true over agreeing = "
          f"{max([r[7] for r in rows if r[1:4] == r[4:7]] or [0]):.2e}")
