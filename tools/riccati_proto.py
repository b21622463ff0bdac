"""Developer prototype (NumPy): the KKT solve of the convex subproblem by a primal Riccati recursion, checked against
the oracle's sparse-LU direction inside the oracle's own interior-point loop.  This is the arithmetic csrc/ipm.cuh
implements (round 2); it is a design tool, not product code and not part of the oracle.

    python tools/riccati_proto.py [model] [omega] [Delta]

Newton-step QP (per knot k: dx_k, du_k):
    min  sum_k 1/2 dx'Hx dx + 1/2 du'Hu du - rx'dx - ru'du
    s.t. dx_0 = rho_0
         E_j dx_{j-1} + G du_{j-1} - F_j dx_j + G du_j = rho_j   (j = 1..N-1; E_j = I + h/2 A_{j-1}, F_j = I - h/2 A_j, G = h/2 B)
         M dx_{N-1} = rho_N                                       (PointGoal coordinates)
The trapezoid row is implicit in x_j and couples u_{j-1} AND u_j.  With  s_j = dx_j - Gam_j du_j,  Gam_j = F_j^-1 G
(Gam_0 = 0) it becomes the explicit recursion  s_{j} = Ah_{j-1} s_{j-1} + Bh_{j-1} du_{j-1} + ch_{j-1}  with
Ah_{j-1} = F_j^-1 E_j,  Bh_{j-1} = Ah_{j-1} Gam_{j-1} + Gam_j,  ch_{j-1} = -F_j^-1 rho_j,  and the stage cost picks up the
cross term  S = Gam'Hx,  R = Hu + Gam'Hx Gam.  Only R + Bh'P Bh (n_u x n_u, >= Hu > 0) is ever factorised: Hx may be
singular (it is, on astrobeeSE3manifold).  The terminal equality is a quadratic penalty w_N |M dx_{N-1} - rho_N|^2 / 2
(= dual regularisation 1/w_N of those rows only), multiplier dnu_N = w_N (M dx_{N-1} - rho_N).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def riccati_solve(Hx, Hu, rx, ru, A, Bm, h, rho0, rho, rhoN, pmask, wN):
    """Returns dx[N,nx], du[N,nu], dnu0[nx], dnu[N-1,nx] (rows 1..N-1), dnuN[nx] (masked)."""
    N, nx = rx.shape
    nu = ru.shape[1]
    hh = 0.5 * h
    I = np.eye(nx)
    G = hh * Bm
    Fi = [np.linalg.inv(I - hh * A[k]) for k in range(N)]
    Gam = [np.zeros((nx, nu))] + [Fi[k] @ G for k in range(1, N)]
    Ah = [Fi[k + 1] @ (I + hh * A[k]) for k in range(N - 1)]
    Bh = [Ah[k] @ Gam[k] + Gam[k + 1] for k in range(N - 1)]
    ch = [-Fi[k + 1] @ rho[k] for k in range(N - 1)]          # rho[k] is row j = k + 1
    Mm = np.diag(pmask.astype(float))
    Q = [Hx[k].copy() for k in range(N)]
    q = [rx[k].copy() for k in range(N)]
    Q[N - 1] = Q[N - 1] + wN * Mm
    q[N - 1] = q[N - 1] + wN * (Mm @ rhoN)
    P = [None] * N; p = [None] * N; K = [None] * N; kap = [None] * N; Acl = [None] * N; dvec = [None] * N
    for k in range(N - 1, -1, -1):
        S = Gam[k].T @ Q[k]
        R = Hu[k] + Gam[k].T @ Q[k] @ Gam[k]
        ru_s = ru[k] + Gam[k].T @ q[k]
        if k == N - 1:
            Lam, Mk, rt, pit = R, S, ru_s, None
        else:
            Pn = P[k + 1]
            pit = p[k + 1] - Pn @ ch[k]
            Lam = R + Bh[k].T @ Pn @ Bh[k]
            Mk = S + Bh[k].T @ Pn @ Ah[k]
            rt = ru_s + Bh[k].T @ pit
        L = np.linalg.cholesky(Lam)
        Y = np.linalg.solve(L, Mk)
        K[k] = np.linalg.solve(L.T, Y)
        kap[k] = np.linalg.solve(L.T, np.linalg.solve(L, rt))
        if k == N - 1:
            P[k] = Q[k] - Y.T @ Y
            p[k] = q[k] - K[k].T @ rt
        else:
            P[k] = Q[k] + Ah[k].T @ Pn @ Ah[k] - Y.T @ Y
            p[k] = q[k] + Ah[k].T @ pit - K[k].T @ rt
            Acl[k] = Ah[k] - Bh[k] @ K[k]
            dvec[k] = Bh[k] @ kap[k] + ch[k]
        P[k] = 0.5 * (P[k] + P[k].T)
    s = np.zeros((N, nx)); du = np.zeros((N, nu)); dx = np.zeros((N, nx))
    s[0] = rho0
    for k in range(N):
        du[k] = -K[k] @ s[k] + kap[k]
        dx[k] = s[k] + Gam[k] @ du[k]
        if k < N - 1:
            s[k + 1] = Acl[k] @ s[k] + dvec[k]
    dnu = np.zeros((N - 1, nx))
    for j in range(1, N):
        dnu[j - 1] = Fi[j].T @ (P[j] @ s[j] - p[j])
    dnuN = wN * (Mm @ dx[N - 1] - Mm @ rhoN)
    E1 = I + hh * A[0]
    dnu0 = rx[0] - Hx[0] @ dx[0] - E1.T @ dnu[0]
    return dx, du, dnu0, dnu, dnuN


def solve_qcqp_riccati(qp, p, lin, tol=1e-8, max_iter=200, verbose=False, compare=True, wN_scale=1e8):
    """oracle/gusto_oracle/ipm.py::solve_qcqp with the linear solve replaced by riccati_solve (slack columns eliminated
    first).  Returns (z, iters, status, worst relative direction difference vs the LU solve)."""
    m_ = p.model
    N, nx, nu = p.N, m_.n_x, m_.n_u
    nX, nU = qp.nX, qp.nU
    nz = nX + nU
    n, m, me = qp.n, qp.h.shape[0], qp.beq.shape[0]
    z = qp.z0.copy()
    A = qp.Aeq.tocsr(); AT = A.T.tocsr()
    c = 0.5 * (qp.Qd @ (z * z)) + qp.G @ z - qp.h
    s = np.maximum(-c, 1e-2)
    lam = np.array(qp.meta["lam0"], dtype=np.float64)
    nu_ = np.zeros(me)
    sc_d = 1.0 + float(np.max(np.abs(qp.q)))
    omega = float(np.max(np.abs(qp.q))) if qp.q.size else 0.0
    pmask = (p.goal_type == 1)
    npt = int(pmask.sum())
    worst = 0.0
    status = "ITERATION_LIMIT"
    for it in range(1, max_iter + 1):
        J = (qp.Qd @ sp.diags(z) + qp.G).tocsr(); JT = J.T.tocsr()
        c = 0.5 * (qp.Qd @ (z * z)) + qp.G @ z - qp.h
        r_d = qp.P * z + qp.q + AT @ nu_ + JT @ lam
        r_p = A @ z - qp.beq
        r_c = c + s
        mu = float(s @ lam) / max(m, 1)
        res = max(np.max(np.abs(r_d)) / sc_d, np.max(np.abs(r_p)), np.max(np.abs(r_c)), mu)
        if verbose:
            print(f"  ipm {it:3d} rd={np.max(np.abs(r_d)):.2e} rp={np.max(np.abs(r_p)):.2e} rc={np.max(np.abs(r_c)):.2e} mu={mu:.2e}")
        if res <= tol:
            status = "OPTIMAL"; break
        w = lam / s
        H = (sp.diags(qp.P + qp.Qd.T @ lam) + JT @ sp.diags(w) @ J).tocsr()
        Hzt = H[:nz, nz:].tocsc(); Htt = H[nz:, nz:].diagonal()
        # slack columns eliminated analytically (as the CUDA slot algebra does): a hinge row with weight w1 whose slack has
        # the second row t >= 0 with weight w2 acts on z with weight w1 w2 / (w1 + w2); numerically Hzz - Hzt Htt^-1 Htz
        # cancels catastrophically when w1 ~ 1e14
        Jt = J[:, nz:].tocsc()
        w_eff = w.copy()
        for j in range(n - nz):
            rws = Jt.indices[Jt.indptr[j]:Jt.indptr[j + 1]]
            for i in rws:
                w_eff[i] = 0.0 if J[i, :nz].nnz == 0 else w[i] * (Htt[j] - w[i]) / Htt[j]
        Jz = J[:, :nz]
        Hzz_red = (sp.diags((qp.P + qp.Qd.T @ lam)[:nz]) + Jz.T @ sp.diags(w_eff) @ Jz).toarray()
        K = sp.bmat([[H + 1e-10 * sp.eye(n), AT], [A, -1e-10 * sp.eye(me)]], format="csc")
        lu = spla.splu(K) if compare else None

        def direction(r_sl):
            nonlocal worst
            rhs1 = -r_d - JT @ ((lam * r_c - r_sl) / s)
            rz, rt = rhs1[:nz], rhs1[nz:]
            Hred = Hzz_red
            rred = rz - Hzt @ (rt / Htt) if Htt.size else rz
            Hx = [Hred[k * nx:(k + 1) * nx, k * nx:(k + 1) * nx] for k in range(N)]
            Hu = [Hred[nX + k * nu:nX + (k + 1) * nu, nX + k * nu:nX + (k + 1) * nu] for k in range(N)]
            rx = rred[:nX].reshape(N, nx); ru = rred[nX:].reshape(N, nu)
            rho = (-r_p[:(N - 1) * nx]).reshape(N - 1, nx)
            rho0 = -r_p[(N - 1) * nx:N * nx]
            rhoN = np.zeros(nx); rhoN[pmask] = -r_p[N * nx:N * nx + npt]
            wN = wN_scale * (1.0 + omega)
            dx, du, dnu0, dnud, dnuN = riccati_solve(Hx, Hu, rx, ru, lin["A"], lin["B"], p.dt, rho0, rho, rhoN, pmask, wN)
            dzz = np.concatenate([dx.ravel(), du.ravel()])
            dt = (rt - Hzt.T @ dzz) / Htt if Htt.size else np.zeros(0)
            dz = np.concatenate([dzz, dt])
            dnu = np.concatenate([dnud.ravel(), dnu0, dnuN[pmask]])
            if compare:
                rhs = np.concatenate([rhs1, -r_p])
                sol = lu.solve(rhs); sol += lu.solve(rhs - K @ sol)
                e1 = np.max(np.abs(sol[:n] - dz)) / max(1e-300, np.max(np.abs(sol[:n])))
                e2 = np.max(np.abs(sol[n:] - dnu)) / max(1e-300, np.max(np.abs(sol[n:])))
                worst = max(worst, e1, e2)
                if verbose:
                    Kx = sp.bmat([[H, AT], [A, None]], format="csr")
                    rr = rhs - Kx @ np.concatenate([dz, dnu])
                    print(f"      dir: |dz-lu|/|lu| {e1:.1e}  |dnu-lu|/|lu| {e2:.1e}  kkt res {np.max(np.abs(rr[:n])):.1e} / {np.max(np.abs(rr[n:])):.1e}  (|rhs| {np.max(np.abs(rhs)):.1e})")
            ds = -r_c - J @ dz
            dlam = (-r_sl - lam * ds) / s
            return dz, dnu, ds, dlam

        def max_step(v, dv, tau):
            neg = dv < 0
            return min(1.0, tau * float(np.min(-v[neg] / dv[neg]))) if np.any(neg) else 1.0

        dz, dnu, ds, dlam = direction(s * lam)
        a_aff = min(max_step(s, ds, 1.0), max_step(lam, dlam, 1.0))
        mu_aff = float((s + a_aff * ds) @ (lam + a_aff * dlam)) / max(m, 1)
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        dz, dnu, ds, dlam = direction(s * lam - max(sigma * mu, 0.1 * tol) + ds * dlam)
        tau = min(max(0.995, 1.0 - mu), 0.999999) if mu < 1 else 0.995
        a_p, a_d = max_step(s, ds, tau), max_step(lam, dlam, tau)
        z = z + a_p * dz; s = s + a_p * ds; nu_ = nu_ + a_d * dnu; lam = lam + a_d * dlam
        mu_new = float(s @ lam) / max(m, 1)
        low = s * lam < 1e-4 * mu_new
        lam[low] = 1e-4 * mu_new / s[low]
    return z, it, status, worst


if __name__ == "__main__":
    from util import gb, to_oracle
    from gusto_oracle.subproblem import build_qcqp, linearize, obstacle_rows
    from gusto_oracle.ipm import solve_qcqp
    name = sys.argv[1] if len(sys.argv) > 1 else "astrobeeSE3"
    omega = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    kw = dict(dubins=dict(B=2, N=30), freeflyerSE2=dict(B=2, N=40), astrobeeSE3=dict(B=2, N=50), astrobeeSE3manifold=dict(B=2, N=60))[name]
    bp = gb.problems.CONFIGS[name](**kw)
    sp_ = bp.model.scp_params
    Delta = float(sys.argv[3]) if len(sys.argv) > 3 else sp_[0]
    X0, U0 = bp.init_traj_straightline()
    for b in range(bp.B):
        p = to_oracle(bp, b)
        toggle = Delta / 8 + bp.model.clearance
        lin = linearize(p, X0[b], U0[b]); rows = obstacle_rows(p, X0[b], toggle)
        qp = build_qcqp(p, X0[b], U0[b], omega, Delta, toggle, sp_[3], lin, rows)
        r = solve_qcqp(qp)
        z, it, st, worst = solve_qcqp_riccati(qp, p, lin, verbose="-v" in sys.argv)
        print(f"{name} b={b} omega={omega} Delta={Delta}: oracle {r.status} {r.iters} it | riccati {st} {it} it, |z - z_oracle| = {np.max(np.abs(z - r.z)):.2e}, worst direction diff {worst:.1e}")
