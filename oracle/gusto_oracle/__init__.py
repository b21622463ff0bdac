"""CPU oracle: float64 NumPy restatement of the GuSTO SCP hot path of StanfordASL/GuSTO.jl.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  The product path (gusto.jl_b200/) never does.

PARITY UNPINNED: the reference ships no tests or golden vectors (test/runtests.jl:9 is `@test 1 == 2`),
cannot be executed here (no Julia / JuMP / Gurobi / Bullet), and the arithmetic of its convex solve and
signed-distance queries lives in un-vendored third-party packages (JuMP 0.19.2 + Gurobi.jl 0.6.0 /
Ipopt.jl 0.5.4, Manifest.toml:356-360,420-424,454-458; BulletCollision.jl, unpinned).  This oracle
restates the reference's own Julia math line by line (each function cites file:line under
/root/reference/src) and replaces the two third-party pieces with their published algorithms:
a primal-dual interior-point method for the convex QCQP (ipm.py) and closed-form sphere/box signed
distance (sdf.py).  The only recorded reference output (examples/freeflyerSE2.ipynb cell 3) is used as
a loose sanity band in tests/test_oracle_notebook.py.
"""
from .models import MODELS, ModelSpec, get_model, f_dyn, A_dyn, B_dyn  # noqa: F401
from .sdf import signed_distance, Obstacle  # noqa: F401
from .subproblem import Problem, linearize, build_qcqp, obstacle_rows  # noqa: F401
from .ipm import solve_qcqp  # noqa: F401
from .scp import solve_gusto, evaluate, SCPResult  # noqa: F401
