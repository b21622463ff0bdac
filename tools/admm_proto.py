#!/usr/bin/env python
"""NumPy prototype of the ADMM iteration the CUDA subproblem kernel (csrc/solve.cu) implements.

Development tool only (used to choose rho / relaxation / stopping rules before writing the kernel); it is
neither product code nor the oracle.  Uses generic sparse algebra instead of the kernel's block-tridiagonal
Schur factorization, but the splitting, prox operators and update order are the same.
"""
import sys, os, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gusto_oracle.subproblem import linearize, obstacle_rows, GOAL_POINT, GOAL_BOX  # noqa: E402


class Split:
    """Blocks w_i = C_i z + d_i with a prox-friendly g_i."""

    def __init__(self):
        self.rows, self.cols, self.vals, self.d = [], [], [], []
        self.blocks = []          # (kind, row_start, row_len, param, rho_class)
        self.n = 0

    def add(self, kind, entries, d, param, cls):
        """entries: list over rows of [(col, val), ...]"""
        s = self.n
        for e, di in zip(entries, d):
            for c, v in e:
                self.rows.append(self.n); self.cols.append(c); self.vals.append(v)
            self.d.append(di); self.n += 1
        self.blocks.append((kind, s, len(d), param, cls))


def build_split(p, Xp, Up, omega, Delta, rows, eps):
    m = p.model
    N, nx, nu = p.N, m.n_x, m.n_u
    nX = N * nx
    xi = lambda k, i: k * nx + i
    ui = lambda k, j: nX + k * nu + j
    S = Split()
    for (idx, scale, rad) in m.ctrl_balls:
        for k in range(N - 1):
            S.add("ball", [[(ui(k, j), s)] for j, s in zip(idx, scale)], [0.0] * len(idx), rad, "u")
    for i in range(nx):
        if p.goal_type[i] == GOAL_BOX:
            S.add("box", [[(xi(N - 1, i), 1.0)]], [0.0], (p.goal_lo[i], p.goal_hi[i]), "x")
    if m.has_trust_region:
        for k in range(N):
            S.add("qhinge", [[(xi(k, j), 1.0)] for j in range(nx)], list(-Xp[k]), (omega, Delta / omega), "x")
    for (idx, lim) in m.soft_norm_rows:
        for k in range(N):
            S.add("qhinge", [[(xi(k, j), 1.0)] for j in idx], [0.0] * len(idx), (omega, lim ** 2), "x")
    for (i, sign, bound) in m.soft_lin_rows:
        for k in range(N):
            S.add("hinge", [[(xi(k, i), sign)]], [-bound], omega, "x")
    D = m.ws_dim
    for k in range(N):
        for i in range(p.n_obs):
            if rows["active"][k, i]:
                S.add("hinge", [[(xi(k, j), -rows["nhat"][k, i, j]) for j in range(D)]], [rows["off"][k, i]], omega, "x")
    if m.quat_idx is not None:
        for k in range(N):
            qp = Xp[k, m.quat_idx]; a = qp / np.linalg.norm(qp)
            S.add("eqhinge", [[(xi(k, j), a[t]) for t, j in enumerate(m.quat_idx)]], [-1.0], (omega, eps), "x")
    return S


def prox_blocks(S, v, rho):
    w = v.copy()
    for (kind, s, n, prm, cls) in S.blocks:
        vi = v[s:s + n]
        t = 1.0 / rho[s]
        if kind == "ball":
            r = np.linalg.norm(vi)
            if r > prm:
                w[s:s + n] = vi * (prm / r)
        elif kind == "box":
            w[s:s + n] = np.clip(vi, prm[0], prm[1])
        elif kind == "qhinge":
            om, lim2 = prm
            r2 = float(vi @ vi)
            if r2 > lim2:
                wi = vi / (1.0 + 2.0 * t * om)
                if float(wi @ wi) < lim2:
                    wi = vi * np.sqrt(lim2 / r2)
                w[s:s + n] = wi
        elif kind == "hinge":
            om = prm
            x = vi[0]
            if x > 0:
                w[s] = x - t * om if x >= t * om else 0.0
        elif kind == "eqhinge":
            om, e = prm
            x = vi[0]
            lo, hi = -e / om, e / om
            if x < lo:
                w[s] = lo
            elif x > hi:
                w[s] = max(x - t * om, hi)
    return w


def admm_solve(p, Xp, Up, omega, Delta, toggle, eps, max_iter=2000, rho_x=1.0, rho_u=1.0, sigma=1e-6, alpha=1.6,
               adapt=True, tol_p=1e-7, tol_d=1e-7, verbose=False, ref=None, check_every=25):
    m = p.model
    N, nx, nu = p.N, m.n_x, m.n_u
    h = p.dt
    lin = linearize(p, Xp, Up)
    rows = obstacle_rows(p, Xp, toggle)
    A, B, g = lin["A"], lin["B"], lin["g"]
    nX, nU = N * nx, N * nu
    n = nX + nU
    wk = np.full(N, h); wk[0] = wk[-1] = 0.5 * h
    P = np.concatenate([np.zeros(nX), np.repeat(2 * wk, nu)])
    # equalities (same as oracle)
    er, ec, ev, beq = [], [], [], []
    I = np.eye(nx)
    r = 0
    for k in range(1, N):
        E = I + 0.5 * h * A[k - 1]; F = I - 0.5 * h * A[k]; Gm = 0.5 * h * B; c = 0.5 * h * (g[k - 1] + g[k])
        for i in range(nx):
            for j in range(nx):
                if E[i, j]: er.append(r); ec.append((k - 1) * nx + j); ev.append(E[i, j])
                if F[i, j]: er.append(r); ec.append(k * nx + j); ev.append(-F[i, j])
            for j in range(nu):
                if Gm[i, j]:
                    er += [r, r]; ec += [nX + (k - 1) * nu + j, nX + k * nu + j]; ev += [Gm[i, j]] * 2
            beq.append(-c[i]); r += 1
    for i in range(nx):
        er.append(r); ec.append(i); ev.append(1.0); beq.append(p.x_init[i]); r += 1
    for i in range(nx):
        if p.goal_type[i] == GOAL_POINT:
            er.append(r); ec.append((N - 1) * nx + i); ev.append(1.0); beq.append(p.goal_lo[i]); r += 1
    Aeq = sp.csr_matrix((ev, (er, ec)), shape=(r, n)); beq = np.array(beq)
    S = build_split(p, Xp, Up, omega, Delta, rows, eps)
    C = sp.csr_matrix((S.vals, (S.rows, S.cols)), shape=(S.n, n)); d = np.array(S.d)
    cls = np.concatenate([[b[4]] * b[2] for b in S.blocks]) if S.blocks else np.zeros(0, dtype=str)
    rho = np.where(cls == "u", rho_u, rho_x).astype(float)

    def factor(rho):
        H = sp.diags(P + sigma) + C.T @ sp.diags(rho) @ C
        K = sp.bmat([[H, Aeq.T], [Aeq, None]], format="csc")
        return spla.splu(K)
    lu = factor(rho)
    z = np.concatenate([Xp.ravel(), Up.ravel()])
    w = prox_blocks(S, C @ z + d, rho)
    y = np.zeros(S.n)
    hist = []
    nfac = 1
    for it in range(1, max_iter + 1):
        rhs = sigma * z + C.T @ (rho * (w - d) - y)
        sol = lu.solve(np.concatenate([rhs, beq]))
        zn = sol[:n]
        Cz = C @ zn + d
        zh = alpha * Cz + (1 - alpha) * w
        wn = prox_blocks(S, zh + y / rho, rho)
        y = y + rho * (zh - wn)
        rp = np.max(np.abs(Cz - wn), initial=0.0)
        rd = np.max(np.abs(C.T @ (rho * (wn - w))), initial=0.0)
        z, w = zn, wn
        if it % check_every == 0 or it == max_iter:
            msg = f"  admm {it:5d} rp={rp:.2e} rd={rd:.2e}"
            if ref is not None:
                msg += f" |z-z*|={np.max(np.abs(z - ref)):.2e}"
            if verbose:
                print(msg)
            hist.append((it, rp, rd))
            if rp <= tol_p and rd <= tol_d:
                break
            if adapt:
                # residual balancing on normalized residuals
                np_ = max(np.max(np.abs(Cz)), np.max(np.abs(wn)), 1e-12)
                nd_ = max(np.max(np.abs(C.T @ y)), np.max(np.abs(P * z)), 1e-12)
                ratio = np.sqrt((rp / np_) / max(rd / nd_, 1e-30))
                if ratio > 5 or ratio < 0.2:
                    ratio = min(max(ratio, 0.1), 10.0)
                    rho = rho * ratio
                    lu = factor(rho); nfac += 1
    X = z[:nX].reshape(N, nx); U = z[nX:].reshape(N, nu)
    return X, U, dict(iters=it, rp=rp, rd=rd, nfac=nfac, lin=lin, rows=rows)


if __name__ == "__main__":
    import importlib.util
    spec = importlib.util.spec_from_file_location("gusto_b200", os.path.join(ROOT, "gusto.jl_b200", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(ROOT, "gusto.jl_b200")])
    g = importlib.util.module_from_spec(spec); sys.modules["gusto_b200"] = g; spec.loader.exec_module(g)
    from gusto_oracle import get_model
    from gusto_oracle.subproblem import Problem
    from gusto_oracle.scp import solve_subproblem, penalized_cost, soft_row_values

    def to_oracle(bp, b):
        mm = get_model(bp.model.name)
        return Problem(mm, bp.N, float(bp.tf[b]), bp.x_init[b], bp.goal_type, bp.goal_lo[b], bp.goal_hi[b], bp.obstacle_table())

    cfg = sys.argv[1] if len(sys.argv) > 1 else "astrobeeSE3"
    omega = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    bp = g.problems.CONFIGS[cfg](B=4) if cfg != "nb" else g.problems.config_astrobee_se3_notebook(50)
    for b in range(min(bp.B, 2)):
        p = to_oracle(bp, b)
        mm = p.model
        Xp, Up = p.init_traj_straightline()
        Delta = mm.scp_params[0]; eps = mm.scp_params[3]
        toggle = Delta / 8 + mm.robot_params[9]
        Xs, Us, obj, st, lin, rows, r = solve_subproblem(p, Xp, Up, omega, Delta, toggle, eps)
        zref = np.concatenate([Xs.ravel(), Us.ravel()])
        t = time.time()
        X, U, info = admm_solve(p, Xp, Up, omega, Delta, toggle, eps, verbose=True, ref=zref)
        Jad = penalized_cost(p, X, U, Xp, rows, omega, Delta, eps)
        ineq, eq = soft_row_values(p, X, Xp, rows)
        print(f"b={b} ipm obj={obj:.8f} admm obj={Jad:.8f} rel={abs(Jad-obj)/abs(obj):.2e} iters={info['iters']} nfac={info['nfac']} "
              f"max soft row={ineq.max():.2e} dX={np.max(np.abs(X-Xs)):.2e} dU={np.max(np.abs(U-Us)):.2e} t={time.time()-t:.2f}")
