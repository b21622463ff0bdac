"""Oracle (test infrastructure): trajectory post-processing of the reference, restated.

  dynamics_constraint_satisfaction   dynamics/astrobee_se3.jl:529-540
  verify_collision_free              dynamics/astrobee_se3.jl:542-562 (BulletCollision.distance -> closed-form SDF, sdf.py)
  interpolate_traj                   dynamics/astrobee_se3.jl:495-527 (removed repmat / Matrix(u,n) calls restated)
  trapezoid_defect                   dynamics_constraints (:151-165) evaluated at the trajectory itself (nonlinear)
"""
import math
import numpy as np

from .models import f_dyn
from .sdf import signed_distance
from .subproblem import Problem, workspace_location


def dynamics_constraint_satisfaction(p: Problem, X, U):
    f = f_dyn(p.model, X, U)
    return float(np.sum(np.abs((X[1:] - X[:-1]) / p.dt - f[:-1])))


def trapezoid_defect(p: Problem, X, U):
    f = f_dyn(p.model, X, U)
    return float(np.max(np.abs(X[1:] - X[:-1] - 0.5 * p.dt * (f[:-1] + f[1:]))))


def verify_collision_free(p: Problem, X):
    """Returns (ok, k, obstacle, dist): first violation with obstacles in the outer loop, knots in the inner one."""
    if not p.n_obs:
        return True, -1, -1, 0.0
    m = p.model
    dist, _ = signed_distance(workspace_location(m, X), p.obstacles, m.robot_params[4], max(m.ws_dim, 1))   # [N, n_obs]
    for i in range(dist.shape[1]):
        for k in range(dist.shape[0]):
            if dist[k, i] < 0:
                return False, k, i, float(dist[k, i])
    return True, -1, -1, 0.0


def min_distance(p: Problem, X):
    if not p.n_obs:
        return 0.0
    m = p.model
    dist, _ = signed_distance(workspace_location(m, X), p.obstacles, m.robot_params[4], max(m.ws_dim, 1))
    return float(dist.min())


def nstep_of(p: Problem, dt_min=0.1):
    return int(math.ceil(p.dt / dt_min))


def interpolate_traj(p: Problem, X, U, nstep):
    m = p.model
    N, nx, nu = p.N, X.shape[1], U.shape[1]
    dt = p.dt / nstep
    Xf = np.zeros((nstep * (N - 1) + 1, nx)); Uf = np.zeros((nstep * (N - 1), nu))
    for k in range(N - 1):
        x = X[k].copy(); u = U[k]
        for s in range(nstep):
            i = nstep * k + s
            Xf[i] = x; Uf[i] = u
            k1 = f_dyn(m, x[None], u[None])[0]
            k2 = f_dyn(m, (x + 0.5 * dt * k1)[None], u[None])[0]
            k3 = f_dyn(m, (x + 0.5 * dt * k2)[None], u[None])[0]
            k4 = f_dyn(m, (x + dt * k3)[None], u[None])[0]
            x = x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    Xf[-1] = X[-1]
    return Xf, Uf
