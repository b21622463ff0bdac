"""Oracle (test infrastructure): indirect shooting refinement of the reference, restated.

  solve!(SS, SP)                     shooting.jl:4-49     Newton on the initial costate p0:  F(p0) = x_goal - x(tf; x_init, p0)
  parameterized_shooting_eval!       shooting.jl:51-66
  shooting_ode! / get_control        dynamics/dubins_car.jl:259-280, dynamics/astrobee_se3_manifold.jl:831-895
  ShootingProblem (x_goal, p0)       types.jl:219-226     p0 = SCPS.dual = -JuMP.dual(init constraints) (get_dual_jump,
                                                          dubins_car.jl:254-257, freeflyer_se2.jl:486-489)
  solve_SCPshooting!                 traj_opt.jl:4-45     alternate one SCP iteration / one shooting attempt

Third-party arithmetic the reference delegates and this file replaces by stated algorithms (PARITY UNPINNED):
  * DifferentialEquations.solve (adaptive Tsit5, dtmin = dt forced)  ->  classical RK4, `nsub` equal sub-steps per knot interval;
  * NLsolve.nlsolve (trust region, finite-difference Jacobian, ftol = 1e-3 on |F|_inf, 100 iterations)  ->
    Levenberg-Marquardt on the exact Jacobian of the discrete flow (complex-step here, forward-mode dual numbers in the
    kernel), same ftol / iteration cap.  LM rather than plain Newton because dx(tf)/dp0 is singular for the quaternion
    model (norm gauge); an exact Jacobian because NLsolve's absolute finite-difference step cbrt(eps) = 6e-6 is larger than
    the attitude costates themselves (~1e-6) while dx(tf)/dp0 ~ 1e5: the difference quotient is noise on that model.
The shooting ODEs are only defined for DubinsCar and AstrobeeSE3Manifold in the reference; so here.
"""
import numpy as np

from .models import DUBINS, ASTROBEE_SE3_MANIFOLD, ModelSpec

CSTEP = 1e-30        # complex-step size: the imaginary part carries the exact tangent


def get_control(m: ModelSpec, X, P):
    """dubins_car.jl:277-280, astrobee_se3_manifold.jl:888-895.  X, P: [..., n_x] -> U [..., n_u]."""
    rp = m.robot_params
    if m.model_id == DUBINS:
        return (0.5 * rp[11] * P[..., 2])[..., None]
    if m.model_id == ASTROBEE_SE3_MANIFOLD:
        F = P[..., 3:6] / (2.0 * rp[0])
        M = P[..., 10:13] / (2.0 * rp[1:4])            # Jinv' * p_omega / 2, J diagonal
        return np.concatenate([F, M], axis=-1)
    raise NotImplementedError("the reference defines shooting_ode! only for DubinsCar and AstrobeeSE3Manifold")


def shooting_ode(m: ModelSpec, Y):
    """d/dt [x; p] (dubins_car.jl:259-275, astrobee_se3_manifold.jl:831-886).  Y: [..., 2 n_x]."""
    rp = m.robot_params
    n = m.n_x
    X, P = Y[..., :n], Y[..., n:]
    U = get_control(m, X, P)
    D = np.zeros_like(Y)
    if m.model_id == DUBINS:
        v, k = rp[10], rp[11]
        th = X[..., 2]
        D[..., 0] = v * np.cos(th)
        D[..., 1] = v * np.sin(th)
        D[..., 2] = k * U[..., 0]
        D[..., 5] = P[..., 0] * v * np.sin(th) - P[..., 1] * v * np.cos(th)
        return D
    mass, J = rp[0], rp[1:4]
    vel = X[..., 3:6]
    qw, qx, qy, qz = X[..., 6], X[..., 7], X[..., 8], X[..., 9]
    w = X[..., 10:13]
    wx, wy, wz = w[..., 0], w[..., 1], w[..., 2]
    pr = P[..., 0:3]
    pqw, pqx, pqy, pqz = P[..., 6], P[..., 7], P[..., 8], P[..., 9]
    F, M = U[..., 0:3], U[..., 3:6]
    D[..., 0:3] = vel
    D[..., 3:6] = F / mass
    D[..., 6] = 0.5 * (-wx * qx - wy * qy - wz * qz)
    D[..., 7] = 0.5 * (wx * qw - wz * qy + wy * qz)
    D[..., 8] = 0.5 * (wy * qw + wz * qx - wx * qz)
    D[..., 9] = 0.5 * (wz * qw - wy * qx + wx * qy)
    Jw = J * w
    cr = np.stack([wy * Jw[..., 2] - wz * Jw[..., 1], wz * Jw[..., 0] - wx * Jw[..., 2], wx * Jw[..., 1] - wy * Jw[..., 0]], axis=-1)
    D[..., 10:13] = (M - cr) / J
    # costates (p_r constant; the gyroscopic part of p_omega is commented out in the reference, :848-862)
    D[..., n + 3:n + 6] = -pr
    D[..., n + 6] = -0.5 * (pqx * wx + pqy * wy + pqz * wz)
    D[..., n + 7] = -0.5 * (-pqw * wx + pqy * wz - pqz * wy)
    D[..., n + 8] = -0.5 * (-pqw * wy - pqx * wz + pqz * wx)
    D[..., n + 9] = -0.5 * (-pqw * wz + pqx * wy - pqy * wx)
    D[..., n + 10] = -0.5 * (-pqw * qx + pqx * qw - pqy * qz + pqz * qy)
    D[..., n + 11] = -0.5 * (-pqw * qy + pqx * qz + pqy * qw - pqz * qx)
    D[..., n + 12] = -0.5 * (-pqw * qz - pqx * qy + pqy * qx + pqz * qw)
    return D


def rk4_step(m, Y, h):
    k = shooting_ode(m, Y)
    acc = Y + (h / 6.0) * k
    k = shooting_ode(m, Y + (0.5 * h) * k)
    acc = acc + (h / 3.0) * k
    k = shooting_ode(m, Y + (0.5 * h) * k)
    acc = acc + (h / 3.0) * k
    k = shooting_ode(m, Y + h * k)
    return acc + (h / 6.0) * k


def integrate(m, x_init, P0, tf, N, nsub, save=False):
    """x(tf) for a stack of initial costates P0 [V, n_x]; with save: the knot values [N, V, 2 n_x]."""
    P0 = np.atleast_2d(P0)
    Y = np.concatenate([np.broadcast_to(x_init, P0.shape), P0], axis=-1)
    Y = Y.astype(np.result_type(Y.dtype, np.float64))          # complex costates stay complex (complex-step Jacobian)
    h = tf / (N - 1) / nsub
    out = [Y.copy()] if save else None
    for _ in range(N - 1):
        for _ in range(nsub):
            Y = rk4_step(m, Y, h)
        if save:
            out.append(Y.copy())
    return np.stack(out) if save else Y


def _solve_spd(Am, b):
    """Gaussian elimination without pivoting (the matrix is J'J + damping, SPD) -- the kernel's exact operation order."""
    n = len(b)
    Am = Am.copy(); b = b.copy()
    for q in range(n):
        ip = 1.0 / Am[q, q]
        for i in range(q + 1, n):
            mlt = Am[i, q] * ip
            Am[i, q + 1:] -= mlt * Am[q, q + 1:]
            b[i] -= mlt * b[q]
    x = np.zeros(n)
    for i in range(n - 1, -1, -1):
        x[i] = (b[i] - Am[i, i + 1:] @ x[i + 1:]) / Am[i, i]
    return x


def solve_shooting(m: ModelSpec, x_init, x_goal, p0, tf, N, nsub=4, max_iter=100, ftol=1e-3):
    """Returns dict(status 'Optimal'|'Diverged', iters, fnorm, p0, X, U, P).  Mirrors csrc/shooting.cuh step by step."""
    n = m.n_x
    p = np.asarray(p0, float).copy()
    xf = integrate(m, x_init, p, tf, N, nsub)[0, :n]
    F = x_goal - xf
    lam = 1e-3
    it = 0
    status = "Diverged"
    if not np.all(np.isfinite(F)):
        return dict(status=status, iters=0, fnorm=np.inf, p0=p, X=None, U=None, P=None)
    while True:
        fn = float(np.max(np.abs(F)))
        if fn <= ftol:
            status = "Optimal"
            break
        if it >= max_iter:
            break
        it += 1
        Pv = p[None].astype(complex) + 1j * CSTEP * np.eye(n)
        Jx = (integrate(m, x_init.astype(complex), Pv, tf, N, nsub)[:, :n].imag / CSTEP).T     # Jx[i, j] = d x_i(tf) / d p_j
        JtJ = Jx.T @ Jx
        JtF = Jx.T @ F
        f2 = float(F @ F)
        accepted = False
        for _ in range(12):
            Am = JtJ + np.diag(lam * np.diag(JtJ) + 1e-14)
            d = _solve_spd(Am, JtF)
            xt = integrate(m, x_init, p + d, tf, N, nsub)[0, :n]
            Ft = x_goal - xt
            f2t = float(Ft @ Ft)
            if np.isfinite(f2t) and f2t < f2:
                p = p + d; F = Ft
                lam = max(lam * 0.1, 1e-12)
                accepted = True
                break
            lam *= 10.0
        if not accepted:
            break
    out = dict(status=status, iters=it, fnorm=float(np.max(np.abs(F))), p0=p, X=None, U=None, P=None)
    Y = integrate(m, x_init, p, tf, N, nsub, save=True)[:, 0]
    out["X"], out["P"] = Y[:, :n], Y[:, n:]
    out["U"] = get_control(m, out["X"], out["P"])
    return out
