"""Oracle (test infrastructure): primal-dual interior-point solver for the convex QCQP of subproblem.py.

Stands in for `JuMP.optimize!` (scp_gusto.jl:104), i.e. for the barrier methods of Gurobi / Ipopt that the
reference reaches through JuMP 0.19.2 (third-party, source absent from /root/reference; Manifest.toml:
356-360,420-424,454-458).  Published algorithm restated here: infeasible-start primal-dual path following
with Mehrotra's predictor-corrector (Mehrotra 1992; Nocedal & Wright, Numerical Optimization, Alg. 14.3 /
19.2 for the nonlinear-inequality form), slack formulation  c(z) + s = 0, s > 0,  fraction-to-boundary
0.995, sparse LU of the reduced KKT matrix
    [ P + diag(Qd' lam) + J' (lam/s) J + delta I     Aeq' ]
    [ Aeq                                          -delta I ].
Convergence: |r_dual|_inf <= tol*(1+|q|_inf), |r_eq|_inf, |r_ineq|_inf, mu <= tol.  The best iterate is kept; if
the method stalls (complementarity far below the residuals) the best iterate is returned as OPTIMAL when its
scaled residual is <= 1e3*tol, mirroring the ALMOST_LOCALLY_SOLVED statuses the reference accepts (scp_gusto.jl:107).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class IPMResult:
    def __init__(self, z, obj, status, iters, res, lam, nu):
        self.z, self.obj, self.status, self.iters, self.res, self.lam, self.nu = z, obj, status, iters, res, lam, nu


def _cvals(qp, z):
    return 0.5 * (qp.Qd @ (z * z)) + qp.G @ z - qp.h


def solve_qcqp(qp, tol=1e-8, max_iter=200, verbose=False):
    n, m, me = qp.n, qp.h.shape[0], qp.beq.shape[0]
    z = qp.z0.copy()
    A = qp.Aeq.tocsr()
    AT = A.T.tocsr()
    c = _cvals(qp, z)
    s = np.maximum(-c, 1e-2)
    lam = np.array(qp.meta["lam0"], dtype=np.float64) if "lam0" in qp.meta else np.ones(m)
    nu = np.zeros(me)
    delta = 1e-10
    status = "ITERATION_LIMIT"
    it = 0
    res = np.inf
    sc_d = 1.0 + (float(np.max(np.abs(qp.q))) if qp.q.size else 0.0)
    best = (np.inf, z, lam, nu)
    stall = 0
    for it in range(1, max_iter + 1):
        J = (qp.Qd @ sp.diags(z) + qp.G).tocsr()
        JT = J.T.tocsr()
        c = _cvals(qp, z)
        r_d = qp.P * z + qp.q + AT @ nu + JT @ lam
        r_p = A @ z - qp.beq
        r_c = c + s
        mu = float(s @ lam) / max(m, 1)
        res = max(np.max(np.abs(r_d), initial=0.0) / sc_d, np.max(np.abs(r_p), initial=0.0),
                  np.max(np.abs(r_c), initial=0.0), mu)
        if res < best[0]:
            best = (res, z.copy(), lam.copy(), nu.copy())
            stall = 0
        else:
            stall += 1
        if stall >= 8 and best[0] <= 1e3 * tol:
            break
        if verbose:
            print(f"  ipm {it:3d} rd={np.max(np.abs(r_d)):.2e} rp={np.max(np.abs(r_p), initial=0):.2e} "
                  f"rc={np.max(np.abs(r_c), initial=0):.2e} mu={mu:.2e}")
        if res <= tol:
            status = "OPTIMAL"
            break
        w = lam / s
        H = sp.diags(qp.P + qp.Qd.T @ lam + delta) + JT @ sp.diags(w) @ J
        K = sp.bmat([[H, AT], [A, -delta * sp.eye(me)]], format="csc")
        try:
            lu = spla.splu(K)
        except RuntimeError:
            delta *= 100
            continue

        def direction(r_sl):
            rhs1 = -r_d - JT @ ((lam * r_c - r_sl) / s)
            rhs = np.concatenate([rhs1, -r_p])
            sol = lu.solve(rhs)
            sol += lu.solve(rhs - K @ sol)                 # one step of iterative refinement
            dz, dnu = sol[:n], sol[n:]
            ds = -r_c - J @ dz
            dlam = (-r_sl - lam * ds) / s
            return dz, dnu, ds, dlam

        def max_step(v, dv, tau):
            neg = dv < 0
            return min(1.0, tau * float(np.min(-v[neg] / dv[neg]))) if np.any(neg) else 1.0

        # predictor
        dz, dnu, ds, dlam = direction(s * lam)
        a_aff = min(max_step(s, ds, 1.0), max_step(lam, dlam, 1.0))
        mu_aff = float((s + a_aff * ds) @ (lam + a_aff * dlam)) / max(m, 1)
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        # corrector
        # centering target never below 0.1*tol: keeps lam/s bounded once the iterate is converged in mu
        dz, dnu, ds, dlam = direction(s * lam - max(sigma * mu, 0.1 * tol) + ds * dlam)
        tau = min(max(0.995, 1.0 - mu), 0.999999) if mu < 1 else 0.995
        a_p, a_d = max_step(s, ds, tau), max_step(lam, dlam, tau)
        z = z + a_p * dz
        s = s + a_p * ds
        nu = nu + a_d * dnu
        lam = lam + a_d * dlam
        # keep the iterate in a wide neighbourhood of the central path (s_i*lam_i >= 1e-4*mu)
        mu_new = float(s @ lam) / max(m, 1)
        low = s * lam < 1e-4 * mu_new
        if np.any(low):
            lam[low] = 1e-4 * mu_new / s[low]
    if status != "OPTIMAL" and best[0] <= 1e3 * tol:
        status = "OPTIMAL"
    if best[0] < res or status != "OPTIMAL":
        res, z, lam, nu = best
    obj = 0.5 * float(z @ (qp.P * z)) + float(qp.q @ z)
    return IPMResult(z, obj, status, it, res, lam, nu)
