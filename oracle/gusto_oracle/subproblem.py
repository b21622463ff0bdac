"""Oracle (test infrastructure): per-knot linearization and assembly of the GuSTO convex subproblem.

Follows /root/reference/src/scp/scp_gusto.jl:178-314 (add_variables_jump!, add_constraints_gusto_jump!,
add_objective_gusto_jump!) with the per-model constraint functions of src/dynamics/*.jl and
src/dynamics.jl:24-42.  Decision vector z = [X (N*n_x, knot-major) ; U (N*n_u) ; t (one slack per
non-trivial soft row)].  Tf is fixed (fixed_final_time) and appears in no constraint (scp_gusto.jl:187).

Slack scaling: the reference writes  s >= omega*g(X) [- Delta],  s >= 0,  cost += s  (:265-295).  The
oracle substitutes s = omega*t, i.e.  t >= g(X) [- Delta/omega],  t >= 0,  cost += omega*t  -- the same
feasible set and objective, better conditioned for the interior-point method.  Soft rows that are
identically zero (obstacle farther than the toggle distance, astrobee_se3.jl:293-304) only give s >= 0,
s = 0 at the optimum and are skipped.
"""
from dataclasses import dataclass, field
import numpy as np
import scipy.sparse as sp

from .models import ModelSpec, f_dyn, A_dyn, B_dyn
from .sdf import signed_distance, pack_obstacles

GOAL_FREE, GOAL_POINT, GOAL_BOX = 0, 1, 2


@dataclass
class Problem:
    """One SCPProblem instance (types.jl:78-86) in flat form (mirrors the C-ABI arguments)."""
    model: ModelSpec
    N: int
    tf: float
    x_init: np.ndarray                 # [n_x]
    goal_type: np.ndarray              # [n_x] int: 0 free / 1 PointGoal (eq) / 2 BoxGoal (ineq), goals.jl
    goal_lo: np.ndarray                # [n_x]  point value or box lower bound
    goal_hi: np.ndarray                # [n_x]  point value or box upper bound
    obstacles: tuple = None            # pack_obstacles(...) = (kind, a, b); order keepout..., obstacle_set... (types.jl:19)

    def __post_init__(self):
        if self.obstacles is None:
            self.obstacles = pack_obstacles([])
        elif not isinstance(self.obstacles, tuple):
            self.obstacles = pack_obstacles(self.obstacles)

    @property
    def n_obs(self):
        return 0 if self.model.ws_dim == 0 else int(self.obstacles[0].shape[0])

    @property
    def dt(self):
        return self.tf / (self.N - 1)                    # types.jl:235

    def init_traj_straightline(self):
        """astrobee_se3.jl:99-113 (same in every model file): X = range(x_init, x_goal, N), U = 0."""
        m = self.model
        x_goal = np.zeros(m.n_x)
        sel = self.goal_type != GOAL_FREE
        x_goal[sel] = 0.5 * (self.goal_lo[sel] + self.goal_hi[sel])      # center(goal), goals.jl:45-47
        s = np.linspace(0.0, 1.0, self.N)[:, None]
        X = self.x_init[None, :] + s * (x_goal - self.x_init)[None, :]
        X[-1] = x_goal
        return X, np.zeros((self.N, m.n_u))


def workspace_location(m: ModelSpec, X):
    """get_workspace_location: astrobee_se3.jl:319-321 (X[1:3,k]); freeflyer_se2.jl:334-336 ([X[1:2,k];0])."""
    r = np.zeros(X.shape[:-1] + (3,))
    if m.ws_dim == 3:
        r[...] = X[..., 0:3]
    elif m.ws_dim == 2:
        r[..., 0:2] = X[..., 0:2]
    return r


def linearize(p: Problem, Xp, Up):
    """update_model_params! (astrobee_se3.jl:130-138) + the affine part of dynamics_constraints (:151-165).

    Returns dict with f[N,n_x], A[N,n_x,n_x], B[n_x,n_u] and g[N,n_x] = f_k - A_k Xp_k - B Up_k, so that row
    block k (k=1..N-1, zero-based pair (k-1,k)) of the dynamics constraint reads
        (I + h/2 A_{k-1}) X_{k-1} + h/2 B U_{k-1} - (I - h/2 A_k) X_k + h/2 B U_k + h/2 (g_{k-1} + g_k) = 0.
    """
    m = p.model
    f = f_dyn(m, Xp, Up)
    A = A_dyn(m, Xp)
    B = B_dyn(m)
    g = f - np.einsum("kij,kj->ki", A, Xp) - Up @ B.T
    return dict(f=f, A=A, B=B, g=g)


def obstacle_rows(p: Problem, Xp, toggle):
    """ncsi_*_convexified (astrobee_se3.jl:282-305): per (knot, obstacle) dist0, nhat, active, and the affine
    row  val(r) = clearance - (dist0 + nhat.(r - r0))  = off - nhat.r  with off = clearance - dist0 + nhat.r0."""
    m = p.model
    r0 = workspace_location(m, Xp)
    dist0, nhat = signed_distance(r0, p.obstacles, m.robot_params[4], max(m.ws_dim, 1)) if p.n_obs else (
        np.zeros((p.N, 0)), np.zeros((p.N, 0, 3)))
    active = dist0 < toggle
    off = m.robot_params[9] - dist0 + np.einsum("kij,kj->ki", nhat, r0)
    return dict(dist0=dist0, nhat=nhat, active=active, off=off)


@dataclass
class QCQP:
    """min 1/2 z'Pz + q'z  s.t.  Aeq z = beq ;  1/2 sum_j Qd[i,j] z_j^2 + G[i,:] z <= h[i]."""
    n: int
    P: np.ndarray          # diagonal of P  [n]
    q: np.ndarray
    Aeq: sp.csr_matrix
    beq: np.ndarray
    Qd: sp.csr_matrix
    G: sp.csr_matrix
    h: np.ndarray
    nX: int
    nU: int
    z0: np.ndarray
    meta: dict = field(default_factory=dict)


def build_qcqp(p: Problem, Xp, Up, omega, Delta, toggle, eps, lin=None, rows=None):
    m = p.model
    N, nx, nu = p.N, m.n_x, m.n_u
    h = p.dt
    lin = lin or linearize(p, Xp, Up)
    rows = rows or obstacle_rows(p, Xp, toggle)
    A, B, g = lin["A"], lin["B"], lin["g"]
    nX, nU = N * nx, N * nu
    xi = lambda k, i: k * nx + i
    ui = lambda k, j: nX + k * nu + j

    # ---- objective: cost_true_convexified = sum_{k=2..N} h/2 (|U_{k-1}|^2 + |U_k|^2)  (astrobee_se3_manifold.jl:73-100)
    wk = np.full(N, h)
    wk[0] = wk[-1] = 0.5 * h
    Pd = [np.zeros(nX), np.repeat(2.0 * wk, nu)]
    qv = [np.zeros(nX + nU)]

    # ---- equalities
    er, ec, ev, beq = [], [], [], []
    nrow = 0

    def add_eq(cols, vals, rhs):
        nonlocal nrow
        er.extend([nrow] * len(cols)); ec.extend(cols); ev.extend(vals); beq.append(rhs); nrow += 1

    # dynamics, k = 2..N  (scp_gusto.jl:197-205)
    I = np.eye(nx)
    for k in range(1, N):
        E = I + 0.5 * h * A[k - 1]
        F = I - 0.5 * h * A[k]
        Gm = 0.5 * h * B
        c = 0.5 * h * (g[k - 1] + g[k])
        for i in range(nx):
            cols, vals = [], []
            for j in range(nx):
                if E[i, j] != 0.0:
                    cols.append(xi(k - 1, j)); vals.append(E[i, j])
                if F[i, j] != 0.0:
                    cols.append(xi(k, j)); vals.append(-F[i, j])
            for j in range(nu):
                if Gm[i, j] != 0.0:
                    cols += [ui(k - 1, j), ui(k, j)]; vals += [Gm[i, j], Gm[i, j]]
            add_eq(cols, vals, -c[i])
    # init  X[i,1] = x_init[i]  (dynamics.jl:24-27, scp_gusto.jl:207-215)
    for i in range(nx):
        add_eq([xi(0, i)], [1.0], p.x_init[i])
    # PointGoal  X[ind,N] = point  (dynamics.jl:30-35, scp_gusto.jl:227-235)
    for i in range(nx):
        if p.goal_type[i] == GOAL_POINT:
            add_eq([xi(N - 1, i)], [1.0], p.goal_lo[i])

    # ---- inequalities
    gr, gc, gv, qr, qc, qvv, hh = [], [], [], [], [], [], []
    nin = 0
    slack_cost = []          # objective coefficient (omega) per slack
    slack_kind = []
    nslack = 0

    def new_slack(kind):
        nonlocal nslack
        idx = nX + nU + nslack
        nslack += 1
        slack_cost.append(omega); slack_kind.append(kind)
        return idx

    def add_in(lin_cols, lin_vals, rhs, quad_cols=(), quad_vals=()):
        nonlocal nin
        gr.extend([nin] * len(lin_cols)); gc.extend(lin_cols); gv.extend(lin_vals)
        qr.extend([nin] * len(quad_cols)); qc.extend(quad_cols); qvv.extend(quad_vals)
        hh.append(rhs); nin += 1

    # hard control balls, k = 1..N-1  (astrobee_se3.jl:255-263,370-371; scp_gusto.jl:217-225)
    for (idx, scale, rad) in m.ctrl_balls:
        for k in range(N - 1):
            add_in([], [], rad ** 2, [ui(k, j) for j in idx], [2.0 * s * s for s in scale])
    # BoxGoal  X[ind,N]-ub <= 0, lb-X[ind,N] <= 0  (dynamics.jl:37-42, scp_gusto.jl:237-245)
    for i in range(nx):
        if p.goal_type[i] == GOAL_BOX:
            add_in([xi(N - 1, i)], [1.0], p.goal_hi[i])
            add_in([xi(N - 1, i)], [-1.0], -p.goal_lo[i])
    # soft state trust region  (astrobee_se3.jl:308-311; scp_gusto.jl:265-279): t >= |X_k-Xp_k|^2 - Delta/omega
    if m.has_trust_region:
        for k in range(N):
            t = new_slack("tr")
            cols = [xi(k, j) for j in range(nx)]
            add_in(cols + [t], list(-2.0 * Xp[k]) + [-1.0], Delta / omega - float(Xp[k] @ Xp[k]), cols, [2.0] * nx)
            add_in([t], [-1.0], 0.0)
    # soft convex state ineq, quadratic  (astrobee_se3.jl:244-252; scp_gusto.jl:281-295)
    for (idx, lim) in m.soft_norm_rows:
        for k in range(N):
            t = new_slack("norm")
            cols = [xi(k, j) for j in idx]
            add_in([t], [-1.0], lim ** 2, cols, [2.0] * len(cols))
            add_in([t], [-1.0], 0.0)
    # soft convex state ineq, linear  (astrobee_se3_manifold.jl:316-319; dynamics.jl:56-64)
    for (i, sign, bound) in m.soft_lin_rows:
        for k in range(N):
            t = new_slack("lin")
            add_in([xi(k, i), t], [sign, -1.0], bound)
            add_in([t], [-1.0], 0.0)
    # soft convexified obstacle rows  (astrobee_se3.jl:282-305): t >= off - nhat.r
    D = m.ws_dim
    for k in range(N):
        for i in range(p.n_obs):
            if rows["active"][k, i]:
                t = new_slack("obs")
                add_in([xi(k, j) for j in range(D)] + [t], list(-rows["nhat"][k, i, :D]) + [-1.0], -rows["off"][k, i])
                add_in([t], [-1.0], 0.0)
    # soft convex state eq (manifold quaternion norm, astrobee_se3_manifold.jl:308-313; scp_gusto.jl:297-311):
    #   e = |qp| + qp.(q-qp)/|qp| - 1 = a.q - 1, a = qp/|qp|.
    #   j=1:  v1 <= omega e + eps, v1 >= 0, cost v1  -> v1 = 0 and the hard row  e >= -eps/omega
    #   j=2:  v2 >= omega e - eps, v2 >= 0, cost v2  -> t >= e - eps/omega
    if m.quat_idx is not None:
        for k in range(N):
            qp = Xp[k, m.quat_idx]
            a = qp / np.linalg.norm(qp)
            cols = [xi(k, j) for j in m.quat_idx]
            add_in(cols, list(-a), -1.0 + eps / omega)
            t = new_slack("eq")
            add_in(cols + [t], list(a) + [-1.0], 1.0 + eps / omega)
            add_in([t], [-1.0], 0.0)

    n = nX + nU + nslack
    Pd.append(np.zeros(nslack))
    qfull = np.concatenate([qv[0], np.array(slack_cost, dtype=np.float64)])
    Aeq = sp.csr_matrix((ev, (er, ec)), shape=(nrow, n))
    G = sp.csr_matrix((gv, (gr, gc)), shape=(nin, n))
    Qd = sp.csr_matrix((qvv, (qr, qc)), shape=(nin, n))
    # Start point: X, U <- previous trajectory (set_start_value, scp_gusto.jl:100-102).  The reference starts the
    # slacks at 0; an interior-point method wants them strictly inside, so each slack starts one unit above its
    # hinge argument and its two multipliers at omega/2 (zero dual residual for the slack column).
    z0 = np.concatenate([Xp.ravel(), Up.ravel(), np.zeros(nslack)])
    lam0 = np.full(nin, 1e-2)
    if nslack:
        c0 = 0.5 * (Qd @ (z0 * z0)) + G @ z0 - np.array(hh)
        Gs = G[:, nX + nU:].tocsc()
        for j in range(nslack):
            rws = Gs.indices[Gs.indptr[j]:Gs.indptr[j + 1]]
            z0[nX + nU + j] = max(float(np.max(c0[rws])), 0.0) + 1.0
            lam0[rws] = omega / len(rws)
    return QCQP(n, np.concatenate(Pd), qfull, Aeq, np.array(beq), Qd, G, np.array(hh), nX, nU, z0,
                meta=dict(slack_kind=slack_kind, lin=lin, rows=rows, lam0=lam0))
