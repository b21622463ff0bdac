#!/usr/bin/env python
"""Generate tests/golden/oracle_golden.npz from the oracle (run in the build container; commit the output).

The reference cannot be executed (no Julia / JuMP / Gurobi / Bullet here), so these are NOT reference outputs: they
freeze the oracle's own outputs so that accidental drift of the restatement shows up in the CPU suite.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from util import gb, orc, to_oracle  # noqa: E402
from gusto_oracle.models import f_dyn, A_dyn, get_model  # noqa: E402
from gusto_oracle.scp import solve_gusto  # noqa: E402

out = {}
rng = np.random.default_rng(2026)
for name in orc.MODELS:
    m = get_model(name)
    x = rng.normal(size=(8, m.n_x)) * 0.4
    u = rng.normal(size=(8, m.n_u)) * 0.4
    out[f"{name}_x"], out[f"{name}_u"] = x, u
    out[f"{name}_f"], out[f"{name}_A"] = f_dyn(m, x, u), A_dyn(m, x)
bp = gb.problems.config_astrobee_se3(B=2, N=20, seed=11)
S = solve_gusto(to_oracle(bp, 0))
out["se3_scp_J"], out["se3_scp_iters"] = S.J_true[-1], S.iterations
out["se3_scp_X"] = S.X
np.savez(os.path.join(HERE, "oracle_golden.npz"), **out)
print("wrote oracle_golden.npz", {k: np.shape(v) for k, v in out.items()})
