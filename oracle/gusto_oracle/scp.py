"""Oracle (test infrastructure): GuSTO outer loop and per-iteration evaluation scalars.

Follows /root/reference/src/scp/scp_gusto.jl:49-176 (solve_gusto_jump!), :34-44
(trust_region_satisfied_gusto), :316-343 (convex_ineq_satisfied_gusto_jump), src/traj_opt.jl:74-85
(convergence_metric) and the per-model trust_region_ratio_gusto / cost_true
(dynamics/astrobee_se3.jl:383-417, astrobee_se3_manifold.jl:73-100,610-642, freeflyer_se2.jl:66-80,392-427,
dubins_car.jl:54-69,229-241), including the quirks listed in SURVEY.md App. D (q2, q3, q8, q9, q10, q11).
"""
from dataclasses import dataclass, field
import numpy as np

from .models import f_dyn, FREEFLYER_SE2
from .sdf import signed_distance
from .subproblem import Problem, linearize, obstacle_rows, build_qcqp, workspace_location
from .ipm import solve_qcqp

FREEFLYER_ARM_OFFSET = np.array([0.0, 0.15, 0.0])       # robot/freeflyer.jl:48 (xb), second hull of the compound


def cost_true(p: Problem, U):
    """sum_{k=2..N} 1/2 dtp (|U_{k-1}|^2 + |U_k|^2)  (astrobee_se3_manifold.jl:73-100)."""
    uu = np.sum(U * U, axis=-1)
    return float(np.sum(0.5 * p.dt * (uu[:-1] + uu[1:])))


def convergence_metric(X, Xp):
    """traj_opt.jl:74-85: max_k |X_k - Xp_k| / max_k |X_k|."""
    return float(np.max(np.linalg.norm(X - Xp, axis=-1)) / np.max(np.linalg.norm(X, axis=-1)))


def trust_region_satisfied(X, Xp, Delta):
    """scp_gusto.jl:34-44: max_k |X_k - Xp_k|^2 - Delta <= 0."""
    return bool(np.max(np.sum((X - Xp) ** 2, axis=-1)) - Delta <= 0)


def trust_region_ratio(p: Problem, X, U, Xp, lin):
    """astrobee_se3.jl:383-417: dynamics part over k=1..N-1 WITHOUT the B(U-Up) term (quirk q2); obstacle part
    over k=1..N and ALL obstacles regardless of the toggle distance.  Freeflyer loops over both hulls of the
    compound robot, each placed at the body translation (freeflyer_se2.jl:407-424, quirk q11)."""
    m = p.model
    N = p.N
    linz = lin["f"] + np.einsum("kij,kj->ki", lin["A"], X - Xp)
    fnew = f_dyn(m, X, U)
    num = float(np.sum(np.linalg.norm(fnew[:N - 1] - linz[:N - 1], axis=-1)))
    den = float(np.sum(np.linalg.norm(linz[:N - 1], axis=-1)))
    if p.n_obs:
        cl, R = m.robot_params[9], m.robot_params[4]
        offsets = [np.zeros(3)] + ([FREEFLYER_ARM_OFFSET] if m.model_id == FREEFLYER_SE2 else [])
        r0, r = workspace_location(m, Xp), workspace_location(m, X)
        for off in offsets:
            d0, n0 = signed_distance(r0 + off, p.obstacles, R, m.ws_dim)
            d1, _ = signed_distance(r + off, p.obstacles, R, m.ws_dim)
            linr = cl - (d0 + np.einsum("kij,kj->ki", n0, r - r0))
            num += float(np.sum(np.abs((cl - d1) - linr)))
            den += float(np.sum(np.abs(linr)))
    return num / den


def soft_row_values(p: Problem, X, Xp, rows):
    """Values of every convex_state_ineq / nonconvex_state_convexified_ineq row and convex_state_eq row at X."""
    m = p.model
    ineq = []
    for (idx, lim) in m.soft_norm_rows:
        ineq.append(np.sum(X[:, idx] ** 2, axis=-1) - lim ** 2)
    for (i, sign, bound) in m.soft_lin_rows:
        ineq.append(sign * X[:, i] - bound)
    if p.n_obs:
        r = workspace_location(m, X)
        v = rows["off"] - np.einsum("kij,kj->ki", rows["nhat"], r)
        ineq.append(np.where(rows["active"], v, 0.0).ravel())
    eq = []
    if m.quat_idx is not None:
        qp = Xp[:, m.quat_idx]
        nq = np.linalg.norm(qp, axis=-1)
        eq.append(nq + np.sum(qp * (X[:, m.quat_idx] - qp), axis=-1) / nq - 1.0)
    return (np.concatenate(ineq) if ineq else np.zeros(0)), (np.concatenate(eq) if eq else np.zeros(0))


def convex_ineq_satisfied(p: Problem, X, Xp, rows, eps):
    """scp_gusto.jl:316-343: any soft ineq row >= eps, or soft eq row outside (-eps, eps) -> false."""
    ineq, eq = soft_row_values(p, X, Xp, rows)
    return bool(np.all(ineq < eps) and np.all(np.abs(eq) < eps))


def penalized_cost(p: Problem, X, U, Xp, rows, omega, Delta, eps):
    """Objective of the convex subproblem at (X,U) with slacks at their optimal values (scp_gusto.jl:253-314)."""
    m = p.model
    J = cost_true(p, U)
    if m.has_trust_region:
        J += float(np.sum(np.maximum(omega * np.sum((X - Xp) ** 2, axis=-1) - Delta, 0.0)))
    ineq, eq = soft_row_values(p, X, Xp, rows)
    J += float(np.sum(np.maximum(omega * ineq, 0.0)))
    J += float(np.sum(np.maximum(omega * eq - eps, 0.0)))
    return J


def evaluate(p: Problem, X, U, Xp, Up, omega, Delta, toggle, eps, lin=None, rows=None):
    """All per-iteration scalars of scp_gusto.jl:115-124 for a candidate (X,U) against the previous (Xp,Up)."""
    lin = lin or linearize(p, Xp, Up)
    rows = rows or obstacle_rows(p, Xp, toggle)
    return dict(conv=convergence_metric(X, Xp),
                tr_ok=trust_region_satisfied(X, Xp, Delta),
                ineq_ok=convex_ineq_satisfied(p, X, Xp, rows, eps),
                rho=trust_region_ratio(p, X, U, Xp, lin),
                J_true=cost_true(p, U),
                J_full=penalized_cost(p, X, U, Xp, rows, omega, Delta, eps))


def solve_subproblem(p: Problem, Xp, Up, omega, Delta, toggle, eps, tol=1e-8):
    """One convex subproblem: assemble (scp_gusto.jl:95-102) and solve (:104).  Returns X, U, obj, status."""
    lin = linearize(p, Xp, Up)
    rows = obstacle_rows(p, Xp, toggle)
    qp = build_qcqp(p, Xp, Up, omega, Delta, toggle, eps, lin, rows)
    r = solve_qcqp(qp, tol=tol)
    m = p.model
    X = r.z[:qp.nX].reshape(p.N, m.n_x)
    U = r.z[qp.nX:qp.nX + qp.nU].reshape(p.N, m.n_u)
    return X, U, r.obj, r.status, lin, rows, r


@dataclass
class SCPResult:
    X: np.ndarray
    U: np.ndarray
    converged: bool = False
    successful: bool = False
    iterations: int = 0
    J_true: list = field(default_factory=list)
    J_full: list = field(default_factory=list)
    solver_status: list = field(default_factory=lambda: ["NA"])
    scp_status: list = field(default_factory=lambda: ["NA"])
    accept_solution: list = field(default_factory=lambda: [True])
    convergence_measure: list = field(default_factory=lambda: [0.0])
    Delta_vec: list = field(default_factory=list)
    omega_vec: list = field(default_factory=list)
    rho_vec: list = field(default_factory=lambda: [0.0])
    tr_ok_vec: list = field(default_factory=lambda: [False])
    ineq_ok_vec: list = field(default_factory=lambda: [False])
    dual: np.ndarray = None          # -JuMP.dual of the init constraints of the last solve (scp_gusto.jl:116, get_dual_jump)


def solve_gusto(p: Problem, X0=None, U0=None, max_iter=30, force=False, subproblem=solve_subproblem, verbose=False):
    """solve_gusto_jump! (scp_gusto.jl:49-176) for one instance."""
    m = p.model
    D0, w0, w_max, eps, rho0, rho1, b_succ, b_fail, g_fail, conv_thr = m.scp_params
    if X0 is None:
        X0, U0 = p.init_traj_straightline()
    S = SCPResult(X0.copy(), U0.copy(), Delta_vec=[D0], omega_vec=[w0])
    lin0 = linearize(p, S.X, S.U)                                     # initialize_model_params! :72
    S.J_true.append(cost_true(p, S.U))                                # :73
    S.J_full.append(S.J_true[-1])
    S.rho_vec.append(trust_region_ratio(p, S.X, S.U, S.X, lin0))      # :75
    toggle = S.Delta_vec[-1] / 8 + m.robot_params[9]                  # :76
    iter_cap = S.iterations + max_iter
    while S.iterations < iter_cap:
        Delta, omega = S.Delta_vec[-1], S.omega_vec[-1]
        Xn, Un, obj, status, lin, rows, r = subproblem(p, S.X, S.U, omega, Delta, toggle, eps)
        S.solver_status.append(status)
        if status != "OPTIMAL":                                       # :107-111
            return S
        if getattr(r, "nu", None) is not None:                        # :116; init rows follow the (N-1) n_x dynamics rows
            S.dual = np.array(r.nu[(p.N - 1) * m.n_x:p.N * m.n_x])
        S.convergence_measure.append(convergence_metric(Xn, S.X))     # :115
        S.J_full.append(obj)
        S.tr_ok_vec.append(trust_region_satisfied(Xn, S.X, Delta))    # :120
        S.ineq_ok_vec.append(convex_ineq_satisfied(p, Xn, S.X, rows, eps))   # :121
        if S.tr_ok_vec[-1]:
            S.rho_vec.append(trust_region_ratio(p, Xn, Un, S.X, lin))  # :124
            if S.rho_vec[-1] > rho1:
                S.scp_status.append("InaccurateModel"); S.accept_solution.append(False)
                S.Delta_vec.append(b_fail * Delta); S.omega_vec.append(omega)
            else:
                S.accept_solution.append(True)
                S.Delta_vec.append(min(b_succ * Delta, D0) if S.rho_vec[-1] < rho0 else Delta)
                if not S.ineq_ok_vec[-1]:
                    S.scp_status.append("ViolatesConstraints"); S.omega_vec.append(g_fail * omega)
                else:
                    S.scp_status.append("OK"); S.omega_vec.append(omega)
        else:
            S.scp_status.append("TrustRegionViolated"); S.accept_solution.append(False)
            S.Delta_vec.append(Delta); S.omega_vec.append(g_fail * omega)
        if S.accept_solution[-1]:
            S.J_true.append(cost_true(p, Un))                         # :146-148
            S.X, S.U = Xn, Un
        else:
            S.J_true.append(S.J_true[-1])
        toggle = S.Delta_vec[-1] / 8 + m.robot_params[9]              # :156
        S.iterations += 1
        if verbose:
            print(f"it {S.iterations:2d} {S.scp_status[-1]:20s} J={S.J_true[-1]:.6f} conv={S.convergence_measure[-1]:.3e} "
                  f"rho={S.rho_vec[-1]:.3e} D={S.Delta_vec[-1]:.4g} w={S.omega_vec[-1]:.4g}")
        if S.omega_vec[-1] > w_max:                                   # :163-166
            break
        if not S.accept_solution[-1]:
            continue
        if S.iterations > 2 and sum(S.convergence_measure[-2:]) <= conv_thr:   # :169-174
            S.converged = True
            if S.ineq_ok_vec[-1]:
                S.successful = True
            if not force:
                break
    return S
