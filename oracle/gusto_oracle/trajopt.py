"""Oracle (test infrastructure): the TrajOpt SCP variant, solve_trajopt_jump! (/root/reference/src/scp/scp_trajopt.jl:33-157).

The reference routine cannot run as written (PARITY UNPINNED, and more so than for GuSTO):
  * scp_trajopt.jl:265-270 penalises the dynamics with  mu*f - v1 <= 0,  v2 - mu*f <= 0,  cost v1 + v2  and no bound on v2 from
    below: the subproblem is unbounded.  Restated here in the form the same file uses for its other equality penalties (:236-246):
    mu*f <= v,  -mu*f <= v,  cost v   (= mu |f|_1, the penalty the paper [Schulman et al.] describes);
  * :69 binds `old_penalty_traj, old_convex_traj = SCPS.traj, SCPS.traj` (aliases, not copies; `copy!` of types.jl:247-252 then
    copies a trajectory onto itself).  Inside the trust loop that makes old_convex_traj the PREVIOUS iterate, which is also the
    linearisation point model.f / model.A belong to: kept (xtol and rho of :112,:121 compare the new iterate with the previous one).
    After the trust loop it makes ftol identically 0, so the loops would end after the first trust loop: :140-141 and :148 are
    restated with real copies taken at :73 / :76 -- the evident intent;
  * :140 tests `xtol[end] < xtol` on the scalar parameter (never true): restated as `xtol_vec[end] < xtol`;
  * evaluate_ctol (:288-312) indexes a Dict{Symbol,...} with 1 (KeyError) and depends on Dict iteration order: restated as
    sum over the constraint classes of max |c(traj) - c(traj_prev)|  over  sum over the classes of max |c(traj)|, classes in the
    order of the merge call (:294) that exist for the model: norm rows (one class each), obstacle signed distance, goal box,
    dynamics (nonlinear trapezoid defect, per-knot norm).
Everything else follows the file line by line: mu-penalised hinge rows for convex_state_ineq / nonconvex_state_convexified_ineq /
convex_control_ineq (:222-233), the HARD trust region |X_k - Xp_k|^2 <= s (:165-173), hard boundary conditions (:175-195),
obstacle_toggle_distance = clearance + 1 (:65), every step accepted (:128), s grown when rho > c else shrunk (:121-126),
trust_region_ratio_trajopt (astrobee_se3.jl:419-459, freeflyer_se2.jl:429-469) with its quirks: the finite-difference term is
(X[:,k] - X[:,k]) / dt = 0, the obstacle model is linearised at the NEW point, all obstacles count.
Slack scaling as in subproblem.py: v = mu t, i.e.  t >= g,  t >= 0,  cost mu t.
"""
from dataclasses import dataclass, field
import numpy as np
import scipy.sparse as sp

from .models import f_dyn, FREEFLYER_SE2, ASTROBEE_SE3
from .sdf import signed_distance
from .subproblem import Problem, linearize, obstacle_rows, workspace_location, QCQP, GOAL_POINT, GOAL_BOX
from .ipm import solve_qcqp
from .scp import cost_true, convergence_metric, FREEFLYER_ARM_OFFSET

# SCPParam_TrajOpt(model): mu0, s0, c, tau_plus, tau_minus, k, ftol, xtol, ctol, max_penalty, max_convex, max_trust
TRAJOPT_PARAMS = {
    ASTROBEE_SE3: np.array([1.0, 10.0, 10.0, 2.0, 0.5, 5.0, 0.01, 0.01, 0.01, 5, 5, 5]),      # astrobee_se3.jl:50-64
    FREEFLYER_SE2: np.array([1.0, 1.0, 10.0, 2.0, 0.5, 5.0, 0.01, 0.1, 0.01, 5, 5, 5]),       # freeflyer_se2.jl:49-63
}


def dynamics_rows(p: Problem, lin):
    """Sparse trapezoid rows (scp_gusto.jl:197-205 form): returns (rows, cols, vals, const) with row index k-1 (k = 1..N-1), so that
    d = M z + const is the linearised defect  dynamics_constraints(traj, traj_prev, SCPP, k)."""
    m = p.model
    N, nx, nu = p.N, m.n_x, m.n_u
    h = p.dt
    A, B, g = lin["A"], lin["B"], lin["g"]
    nX = N * nx
    I = np.eye(nx)
    er, ec, ev, const = [], [], [], []
    for k in range(1, N):
        E = I + 0.5 * h * A[k - 1]
        F = I - 0.5 * h * A[k]
        Gm = 0.5 * h * B
        c = 0.5 * h * (g[k - 1] + g[k])
        for i in range(nx):
            r = (k - 1) * nx + i
            for j in range(nx):
                if E[i, j] != 0.0:
                    er.append(r); ec.append((k - 1) * nx + j); ev.append(E[i, j])
                if F[i, j] != 0.0:
                    er.append(r); ec.append(k * nx + j); ev.append(-F[i, j])
            for j in range(nu):
                if Gm[i, j] != 0.0:
                    er += [r, r]; ec += [nX + (k - 1) * nu + j, nX + k * nu + j]; ev += [Gm[i, j], Gm[i, j]]
            const.append(c[i])
    return er, ec, ev, np.array(const)


def build_qcqp_trajopt(p: Problem, Xp, Up, mu, s, toggle, lin=None, rows=None):
    """add_constraints_trajopt_jump! / add_objective_trajopt_jump! (scp_trajopt.jl:159-279)."""
    m = p.model
    N, nx, nu = p.N, m.n_x, m.n_u
    h = p.dt
    lin = lin or linearize(p, Xp, Up)
    rows = rows or obstacle_rows(p, Xp, toggle)
    nX, nU = N * nx, N * nu
    xi = lambda k, i: k * nx + i
    ui = lambda k, j: nX + k * nu + j
    wk = np.full(N, h); wk[0] = wk[-1] = 0.5 * h
    Pd = [np.zeros(nX), np.repeat(2.0 * wk, nu)]                     # cost_true_convexified (:218)

    er, ec, ev, beq = [], [], [], []
    nrow = 0

    def add_eq(cols, vals, rhs):
        nonlocal nrow
        er.extend([nrow] * len(cols)); ec.extend(cols); ev.extend(vals); beq.append(rhs); nrow += 1

    for i in range(nx):                                               # init (:175-183)
        add_eq([xi(0, i)], [1.0], p.x_init[i])
    for i in range(nx):                                               # PointGoal
        if p.goal_type[i] == GOAL_POINT:
            add_eq([xi(N - 1, i)], [1.0], p.goal_lo[i])

    gr, gc, gv, qr, qc, qvv, hh = [], [], [], [], [], [], []
    nin = 0
    nslack = 0
    slack_kind = []

    def new_slack(kind):
        nonlocal nslack
        idx = nX + nU + nslack
        nslack += 1
        slack_kind.append(kind)
        return idx

    def add_in(lin_cols, lin_vals, rhs, quad_cols=(), quad_vals=()):
        nonlocal nin
        gr.extend([nin] * len(lin_cols)); gc.extend(lin_cols); gv.extend(lin_vals)
        qr.extend([nin] * len(quad_cols)); qc.extend(quad_cols); qvv.extend(quad_vals)
        hh.append(rhs); nin += 1

    # hard state trust region  |X_k - Xp_k|^2 - s <= 0  (:165-173)
    if m.has_trust_region:
        for k in range(N):
            cols = [xi(k, j) for j in range(nx)]
            add_in(cols, list(-2.0 * Xp[k]), s - float(Xp[k] @ Xp[k]), cols, [2.0] * nx)
    # BoxGoal, hard (:185-195)
    for i in range(nx):
        if p.goal_type[i] == GOAL_BOX:
            add_in([xi(N - 1, i)], [1.0], p.goal_hi[i])
            add_in([xi(N - 1, i)], [-1.0], -p.goal_lo[i])
    # mu-penalised inequality rows (:222-233): convex_state_ineq, nonconvex_state_convexified_ineq, convex_control_ineq
    for (idx, lim) in m.soft_norm_rows:
        for k in range(N):
            t = new_slack("norm")
            cols = [xi(k, j) for j in idx]
            add_in([t], [-1.0], lim ** 2, cols, [2.0] * len(cols))
            add_in([t], [-1.0], 0.0)
    for (i, sign, bound) in m.soft_lin_rows:
        for k in range(N):
            t = new_slack("lin")
            add_in([xi(k, i), t], [sign, -1.0], bound)
            add_in([t], [-1.0], 0.0)
    D = m.ws_dim
    for k in range(N):
        for i in range(p.n_obs):
            if rows["active"][k, i]:
                t = new_slack("obs")
                add_in([xi(k, j) for j in range(D)] + [t], list(-rows["nhat"][k, i, :D]) + [-1.0], -rows["off"][k, i])
                add_in([t], [-1.0], 0.0)
    for (idx, scale, rad) in m.ctrl_balls:                            # k = 1..N-1 (the registry's ind_time)
        for k in range(N - 1):
            t = new_slack("ball")
            add_in([t], [-1.0], rad ** 2, [ui(k, j) for j in idx], [2.0 * sc * sc for sc in scale])
            add_in([t], [-1.0], 0.0)
    # mu |dynamics row|_1 (:257-275 as repaired in the module docstring):  t >= d_i,  t >= -d_i
    dr, dc, dv, dconst = dynamics_rows(p, lin)
    Md = sp.csr_matrix((dv, (dr, dc)), shape=((N - 1) * nx, nX + nU))
    for r in range((N - 1) * nx):
        t = new_slack("dyn")
        cols = list(Md.indices[Md.indptr[r]:Md.indptr[r + 1]])
        vals = list(Md.data[Md.indptr[r]:Md.indptr[r + 1]])
        add_in(cols + [t], vals + [-1.0], -dconst[r])
        add_in(cols + [t], [-v for v in vals] + [-1.0], dconst[r])

    n = nX + nU + nslack
    Pd.append(np.zeros(nslack))
    q = np.concatenate([np.zeros(nX + nU), np.full(nslack, float(mu))])
    Aeq = sp.csr_matrix((ev, (er, ec)), shape=(nrow, n))
    G = sp.csr_matrix((gv, (gr, gc)), shape=(nin, n))
    Qd = sp.csr_matrix((qvv, (qr, qc)), shape=(nin, n))
    z0 = np.concatenate([Xp.ravel(), Up.ravel(), np.zeros(nslack)])
    lam0 = np.full(nin, 1e-2)
    hv = np.array(hh)
    if nslack:
        c0 = 0.5 * (Qd @ (z0 * z0)) + G @ z0 - hv
        Gs = G[:, nX + nU:].tocsc()
        for j in range(nslack):
            rws = Gs.indices[Gs.indptr[j]:Gs.indptr[j + 1]]
            z0[nX + nU + j] = max(float(np.max(c0[rws])), 0.0) + 1.0
            lam0[rws] = mu / len(rws)
    return QCQP(n, np.concatenate(Pd), q, Aeq, np.array(beq), Qd, G, hv, nX, nU, z0,
                meta=dict(slack_kind=slack_kind, lin=lin, rows=rows, lam0=lam0, dyn=(Md, dconst)))


def solve_trajopt_subproblem(p: Problem, Xp, Up, mu, s, tol=1e-8):
    m = p.model
    toggle = m.robot_params[9] + 1.0                                  # :65
    lin = linearize(p, Xp, Up)
    rows = obstacle_rows(p, Xp, toggle)
    qp = build_qcqp_trajopt(p, Xp, Up, mu, s, toggle, lin, rows)
    r = solve_qcqp(qp, tol=tol)
    X = r.z[:qp.nX].reshape(p.N, m.n_x)
    U = r.z[qp.nX:qp.nX + qp.nU].reshape(p.N, m.n_u)
    return X, U, r.obj, r.status, lin, rows, r


def linearized_defect(p: Problem, X, U, lin):
    """dynamics_constraints(traj, traj_prev, SCPP, k) for k = 1..N-1 at (X, U), linearisation `lin` of traj_prev: [N-1, n_x]."""
    h = p.dt
    A, B, g = lin["A"], lin["B"], lin["g"]
    fl = np.einsum("kij,kj->ki", A, X) + U @ B.T + g                  # linearised f at every knot
    return X[:-1] - X[1:] + 0.5 * h * (fl[:-1] + fl[1:])


def nonlinear_defect(p: Problem, X, U):
    h = p.dt
    f = f_dyn(p.model, X, U)
    return X[:-1] - X[1:] + 0.5 * h * (f[:-1] + f[1:])


def trust_region_ratio_trajopt(p: Problem, X, U, Xp, Up, lin):
    """astrobee_se3.jl:419-459 / freeflyer_se2.jl:429-469 (see the module docstring for the quirks kept)."""
    m = p.model
    N = p.N
    fp = lin["f"]
    fn = f_dyn(m, X, U)
    dl = linearized_defect(p, X, U, lin)
    phi_old = np.sum(np.abs(fp[:N - 1]), axis=-1)
    phi_new = np.sum(np.abs(fn[:N - 1]), axis=-1)
    phi_hat = np.sum(np.abs(dl), axis=-1)
    num = float(np.sum(phi_old - phi_new))
    den = float(np.sum(phi_old - phi_hat))
    if p.n_obs:
        cl, R = m.robot_params[9], m.robot_params[4]
        r0, r = workspace_location(m, Xp), workspace_location(m, X)
        offsets = [np.zeros(3)] + ([FREEFLYER_ARM_OFFSET] if m.model_id == FREEFLYER_SE2 else [])   # env_.convex_robot_components
        for off in offsets:
            d0, _ = signed_distance(r0 + off, p.obstacles, R, m.ws_dim)
            d1, n1 = signed_distance(r + off, p.obstacles, R, m.ws_dim)
            hat = cl - (d1 + np.einsum("kij,kj->ki", n1, r - r0))
            num += float(np.sum((cl - d0) - (cl - d1)))
            den += float(np.sum((cl - d0) - hat))
    return num / den


def penalized_cost_trajopt(p: Problem, X, U, mu, lin, rows):
    """Objective of the TrajOpt subproblem at (X, U) with every slack at its optimal value (scp_trajopt.jl:213-279)."""
    m = p.model
    J = cost_true(p, U)
    for (idx, lim) in m.soft_norm_rows:
        J += mu * float(np.sum(np.maximum(np.sum(X[:, idx] ** 2, axis=-1) - lim ** 2, 0.0)))
    for (i, sign, bound) in m.soft_lin_rows:
        J += mu * float(np.sum(np.maximum(sign * X[:, i] - bound, 0.0)))
    if p.n_obs:
        v = rows["off"] - np.einsum("kij,kj->ki", rows["nhat"], workspace_location(m, X))
        J += mu * float(np.sum(np.maximum(np.where(rows["active"], v, 0.0), 0.0)))
    for (idx, scale, rad) in m.ctrl_balls:
        J += mu * float(np.sum(np.maximum(np.sum((U[:-1][:, idx] * np.asarray(scale)) ** 2, axis=-1) - rad ** 2, 0.0)))
    J += mu * float(np.sum(np.abs(linearized_defect(p, X, U, lin))))
    return J


def constraint_classes(p: Problem, X, U):
    """Values of the constraint classes evaluate_ctol walks (restated order, module docstring): list of [n_items, width] arrays."""
    m = p.model
    out = []
    for (idx, lim) in m.soft_norm_rows:
        out.append((np.sum(X[:, idx] ** 2, axis=-1) - lim ** 2)[:, None])
    for (i, sign, bound) in m.soft_lin_rows:
        out.append((sign * X[:, i] - bound)[:, None])
    if p.n_obs:
        d, _ = signed_distance(workspace_location(m, X), p.obstacles, m.robot_params[4], m.ws_dim)
        out.append((m.robot_params[9] - d).reshape(-1, 1))            # ncsi_obstacle_avoidance_signed_distance
    box = [i for i in range(m.n_x) if p.goal_type[i] == GOAL_BOX]
    if box:
        out.append(np.array([[X[-1, i] - p.goal_hi[i], p.goal_lo[i] - X[-1, i]] for i in box]).reshape(-1, 1))
    out.append(nonlinear_defect(p, X, U))                             # dynamics, :array (norm per knot)
    return out


def evaluate_ctol(p: Problem, X, U, Xr, Ur):
    num = den = 0.0
    for a, b in zip(constraint_classes(p, X, U), constraint_classes(p, Xr, Ur)):
        num += float(np.max(np.linalg.norm(a - b, axis=-1)))
        den += float(np.max(np.linalg.norm(a, axis=-1)))
    return num / den


def evaluate_ftol(p: Problem, U, Ur):
    return abs(cost_true(p, U) - cost_true(p, Ur)) / abs(cost_true(p, U))


@dataclass
class TrajOptResult:
    X: np.ndarray
    U: np.ndarray
    converged: bool = False
    iterations: int = 0
    J_true: list = field(default_factory=list)
    J_full: list = field(default_factory=list)
    solver_status: list = field(default_factory=lambda: ["NA"])
    convergence_measure: list = field(default_factory=lambda: [0.0])
    rho_vec: list = field(default_factory=lambda: [0.0])
    mu_vec: list = field(default_factory=list)
    s_vec: list = field(default_factory=list)
    xtol_vec: list = field(default_factory=lambda: [0.0])
    ftol_vec: list = field(default_factory=lambda: [0.0])
    ctol_vec: list = field(default_factory=lambda: [0.0])


def solve_trajopt(p: Problem, X0=None, U0=None, params=None, subproblem=solve_trajopt_subproblem, verbose=False):
    """solve_trajopt_jump! (scp_trajopt.jl:33-157) for one instance, repaired as stated in the module docstring."""
    m = p.model
    prm = TRAJOPT_PARAMS[m.model_id] if params is None else params
    mu0, s0, c, tp, tm, kfac, ftol, xtol, ctol = prm[:9]
    max_pen, max_cvx, max_tr = int(prm[9]), int(prm[10]), int(prm[11])
    if X0 is None:
        X0, U0 = p.init_traj_straightline()
    S = TrajOptResult(X0.copy(), U0.copy(), mu_vec=[mu0], s_vec=[s0])
    S.J_true.append(cost_true(p, S.U))                                # :64
    constraints_satisfied = False
    xtol_satisfied = False
    for _pen in range(max_pen):                                       # :71
        if constraints_satisfied:
            break
        Xpen, Upen = S.X.copy(), S.U.copy()
        for _cvx in range(max_cvx):                                   # :75
            Xcvx, Ucvx = S.X.copy(), S.U.copy()
            if constraints_satisfied:
                break
            if xtol_satisfied:
                xtol_satisfied = False
                break
            for _tr in range(max_tr):                                 # :83
                Xn, Un, obj, status, lin, rows, r = subproblem(p, S.X, S.U, S.mu_vec[-1], S.s_vec[-1])
                S.solver_status.append(status)
                if status != "OPTIMAL":                               # the reference only warns (:108-111); a failed solve has no iterate
                    return S
                S.xtol_vec.append(convergence_metric(Xn, S.X))        # :112 (old_convex_traj aliases the previous iterate)
                S.convergence_measure.append(S.xtol_vec[-1])
                S.J_full.append(obj)
                S.rho_vec.append(trust_region_ratio_trajopt(p, Xn, Un, S.X, S.U, lin))
                S.s_vec.append((tp if S.rho_vec[-1] > c else tm) * S.s_vec[-1])      # :122-126
                S.X, S.U = Xn, Un                                     # :128
                S.J_true.append(cost_true(p, Un))
                S.iterations += 1
                if verbose:
                    print(f"it {S.iterations:2d} mu={S.mu_vec[-1]:g} s={S.s_vec[-2]:g} J={S.J_true[-1]:.6f} xtol={S.xtol_vec[-1]:.3e} rho={S.rho_vec[-1]:.3e}")
                if S.s_vec[-1] < xtol:                                # :134
                    xtol_satisfied = True
                    break
            S.ftol_vec.append(evaluate_ftol(p, S.U, Ucvx))            # :140-141
            S.xtol_vec.append(convergence_metric(S.X, Xcvx))
            if S.ftol_vec[-1] < ftol or S.xtol_vec[-1] < xtol:        # :142
                constraints_satisfied = True
                break
        S.ctol_vec.append(evaluate_ctol(p, S.X, S.U, Xpen, Upen))
        if S.ctol_vec[-1] < ctol:                                     # :148
            constraints_satisfied = True
            S.converged = True
            break
        S.mu_vec.append(S.mu_vec[-1] * kfac)                          # :154
    return S
