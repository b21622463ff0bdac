"""Oracle (test infrastructure): per-model dynamics f, A=df/dx, B=df/du and parameter tables.

All citations are relative to /root/reference/src.  Arrays are knot-major: X[k] is the state at knot k
(== column k of the reference's x_dim x N Julia matrix, same memory order).

Flat parameter layout shared with the C ABI (include/gusto_b200.h, `robot_params[16]`):
  0 mass | 1..3 Jxx,Jyy,Jzz | 4 robot radius | 5 v_max | 6 a_max | 7 omega_max | 8 alpha_max |
  9 clearance | 10 dubins v | 11 dubins k | 12..14 dubins x_max | 15 dubins u_max
SCP hyper-parameters `scp_params[10]`:
  Delta0, omega0, omega_max, eps, rho0, rho1, beta_succ, beta_fail, gamma_fail, convergence_threshold
"""
from dataclasses import dataclass, field
import numpy as np

DUBINS, FREEFLYER_SE2, ASTROBEE_SE3, ASTROBEE_SE3_MANIFOLD = 0, 1, 2, 3


@dataclass
class ModelSpec:
    model_id: int
    name: str
    n_x: int
    n_u: int
    robot_params: np.ndarray            # [16] float64, layout above
    scp_params: np.ndarray              # [10] float64, layout above
    has_trust_region: bool              # stri_state_trust_region registered?
    ws_dim: int                         # workspace dimension of the obstacle query (3, 2, or 0 = none)
    # soft ||x[idx]||^2 - lim^2 rows (convex_state_ineq, quadratic)
    soft_norm_rows: list = field(default_factory=list)     # [(idx array, limit)]
    # soft linear rows  sign*x[i] - bound  (convex_state_ineq, linear)
    soft_lin_rows: list = field(default_factory=list)      # [(i, sign, bound)]
    # hard control balls  ||scale*u[idx]||^2 <= rad^2, k = 1..N-1 (convex_control_ineq)
    ctrl_balls: list = field(default_factory=list)         # [(idx array, scale array, radius)]
    quat_idx: np.ndarray = None         # manifold: indices of the quaternion for cse_quaternion_norm


def _params(**kw):
    p = np.zeros(16)
    names = dict(mass=0, Jxx=1, Jyy=2, Jzz=3, r=4, v_max=5, a_max=6, w_max=7, al_max=8, clearance=9,
                 dub_v=10, dub_k=11, dub_xmax0=12, dub_xmax1=13, dub_xmax2=14, dub_umax=15)
    for k, v in kw.items():
        p[names[k]] = v
    return p


def _astrobee3d(clearance):
    # robot/astrobee3D.jl:16-30
    s = 0.5 * 0.305
    return _params(mass=7.0, Jxx=0.1083, Jyy=0.1083, Jzz=0.1083, r=np.sqrt(3.0) * s, v_max=0.5, a_max=0.1,
                   w_max=45 * np.pi / 180, al_max=50 * np.pi / 180, clearance=clearance)


def _freeflyer(clearance):
    # robot/freeflyer.jl:29-50
    mass_ff = 0.5 * (15.36 + 18.08)
    J_ff = 0.184
    J_w = J_ff / 6.43
    return _params(mass=mass_ff, Jxx=J_ff, Jyy=J_ff, Jzz=J_ff, r=0.157, v_max=0.2, a_max=2 * 0.185 / mass_ff,
                   w_max=20 * np.pi / 180, al_max=(1.0 / J_w) * 0.593, clearance=clearance)


def get_model(name_or_id) -> ModelSpec:
    key = name_or_id if isinstance(name_or_id, str) else {0: "dubins", 1: "freeflyerSE2", 2: "astrobeeSE3",
                                                          3: "astrobeeSE3manifold"}[int(name_or_id)]
    if key == "astrobeeSE3":
        # dynamics/astrobee_se3.jl:16-40 (x_dim 12, u_dim 6, clearance 0.03; SCPParam/SCPParam_GuSTO),
        # registry :324-380 (with the intended semantics of SURVEY App. D-1).
        rp = _astrobee3d(0.03)
        return ModelSpec(ASTROBEE_SE3, key, 12, 6, rp,
                         np.array([10., 1., 1e10, 1e-6, 0.01, 0.05, 2., 0.5, 5., 0.01]),
                         has_trust_region=True, ws_dim=3,
                         soft_norm_rows=[(np.arange(3, 6), rp[5]), (np.arange(9, 12), rp[7])],
                         ctrl_balls=[(np.arange(0, 3), np.full(3, 1 / rp[0]), rp[6]),
                                     (np.arange(3, 6), 1 / rp[1:4], rp[8])])
    if key == "astrobeeSE3manifold":
        # dynamics/astrobee_se3_manifold.jl:18-46, registry :533-607 (no state trust region, :601)
        rp = _astrobee3d(0.03)
        return ModelSpec(ASTROBEE_SE3_MANIFOLD, key, 13, 6, rp,
                         np.array([1000., 1., 1e10, 1e-1, 0.01, 100., 2., 0.5, 5., 1e-4]),
                         has_trust_region=False, ws_dim=3,
                         soft_lin_rows=[(6, -1.0, 0.0)],            # csi_orientation_sign :316-319
                         soft_norm_rows=[(np.arange(3, 6), rp[5]), (np.arange(10, 13), rp[7])],
                         ctrl_balls=[(np.arange(0, 3), np.full(3, 1 / rp[0]), rp[6]),
                                     (np.arange(3, 6), 1 / rp[1:4], rp[8])],
                         quat_idx=np.arange(6, 10))
    if key == "freeflyerSE2":
        # dynamics/freeflyer_se2.jl:14-39, registry :338-390
        rp = _freeflyer(0.05)
        return ModelSpec(FREEFLYER_SE2, key, 6, 3, rp,
                         np.array([3., 1., 1e10, 1e-2, 0.1, 0.3, 2., 0.5, 10., 1e-2]),
                         has_trust_region=True, ws_dim=2,
                         soft_norm_rows=[(np.arange(3, 5), rp[5]), (np.arange(5, 6), rp[7])],
                         ctrl_balls=[(np.arange(0, 2), np.full(2, 1 / rp[0]), rp[6]),
                                     (np.arange(2, 3), np.array([1 / rp[1]]), rp[8])])
    if key == "dubins":
        # dynamics/dubins_car.jl:20-52, registry :184-226 (init condition taken as hard)
        rp = _params(mass=1.0, dub_v=2.0, dub_k=1.0, dub_xmax0=100., dub_xmax1=100., dub_xmax2=2 * np.pi,
                     dub_umax=10., clearance=0.01)
        lin = [(i, +1.0, rp[12 + i]) for i in range(3)] + [(i, -1.0, rp[12 + i]) for i in range(3)]
        return ModelSpec(DUBINS, key, 3, 1, rp,
                         np.array([1e4, 1., 1e10, 1e-6, 0.4, 1.5, 2., 0.5, 5., 1e-4]),
                         has_trust_region=False, ws_dim=0,
                         soft_lin_rows=lin,
                         ctrl_balls=[(np.arange(0, 1), np.ones(1), rp[15])])
    raise KeyError(key)


MODELS = ("dubins", "freeflyerSE2", "astrobeeSE3", "astrobeeSE3manifold")


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def f_dyn(m: ModelSpec, X, U):
    """Continuous dynamics xdot = f(x,u) for every knot.  X:(...,n_x) U:(...,n_u) -> (...,n_x)."""
    X = np.asarray(X, dtype=np.float64)
    U = np.asarray(U, dtype=np.float64)
    rp = m.robot_params
    f = np.zeros(X.shape)
    if m.model_id == ASTROBEE_SE3:
        # astrobee_se3.jl:180-190, utils/quat_functions.jl:253-256 (mrp_derivative)
        v, p, w = X[..., 3:6], X[..., 6:9], X[..., 9:12]
        F, M = U[..., 0:3], U[..., 3:6]
        J = rp[1:4]
        f[..., 0:3] = v
        f[..., 3:6] = F / rp[0]
        pp = np.sum(p * p, axis=-1, keepdims=True)
        wp = np.sum(w * p, axis=-1, keepdims=True)
        f[..., 6:9] = 0.25 * ((1 - pp) * w - 2 * _cross(w, p) + 2 * wp * p)
        f[..., 9:12] = (M - _cross(w, J * w)) / J
    elif m.model_id == ASTROBEE_SE3_MANIFOLD:
        # astrobee_se3_manifold.jl:230-246
        v, w = X[..., 3:6], X[..., 10:13]
        qw, qx, qy, qz = (X[..., 6], X[..., 7], X[..., 8], X[..., 9])
        wx, wy, wz = w[..., 0], w[..., 1], w[..., 2]
        F, M = U[..., 0:3], U[..., 3:6]
        J = rp[1:4]
        f[..., 0:3] = v
        f[..., 3:6] = F / rp[0]
        f[..., 6] = 0.5 * (-wx * qx - wy * qy - wz * qz)
        f[..., 7] = 0.5 * (wx * qw - wz * qy + wy * qz)
        f[..., 8] = 0.5 * (wy * qw + wz * qx - wx * qz)
        f[..., 9] = 0.5 * (wz * qw - wy * qx + wx * qy)
        f[..., 10:13] = (M - _cross(w, J * w)) / J
    elif m.model_id == FREEFLYER_SE2:
        # freeflyer_se2.jl:189-194
        f[..., 0:3] = X[..., 3:6]
        f[..., 3:5] = U[..., 0:2] / rp[0]
        f[..., 5] = U[..., 2] / rp[1]
    elif m.model_id == DUBINS:
        # dubins_car.jl:161-165
        f[..., 0] = rp[10] * np.cos(X[..., 2])
        f[..., 1] = rp[10] * np.sin(X[..., 2])
        f[..., 2] = rp[11] * U[..., 0]
    else:
        raise ValueError(m.model_id)
    return f


def A_dyn(m: ModelSpec, X):
    """Jacobian df/dx per knot.  X:(...,n_x) -> (...,n_x,n_x) with A[...,i,j] = d f_i / d x_j."""
    X = np.asarray(X, dtype=np.float64)
    rp = m.robot_params
    n = m.n_x
    A = np.zeros(X.shape[:-1] + (n, n))
    if m.model_id == ASTROBEE_SE3:
        # astrobee_se3.jl:192-233
        A[..., 0, 3] = A[..., 1, 4] = A[..., 2, 5] = 1.0        # kron([0 1;0 0], I3) :196
        Jxx, Jyy, Jzz = rp[1:4]
        px, py, pz = X[..., 6], X[..., 7], X[..., 8]
        wx, wy, wz = X[..., 9], X[..., 10], X[..., 11]
        d = (px * wx) / 2 + (py * wy) / 2 + (pz * wz) / 2
        A[..., 6, 6] = d
        A[..., 6, 7] = wz / 2 + (px * wy) / 2 - (py * wx) / 2
        A[..., 6, 8] = (px * wz) / 2 - wy / 2 - (pz * wx) / 2
        A[..., 6, 9] = px ** 2 / 4 - py ** 2 / 4 - pz ** 2 / 4 + 0.25
        A[..., 6, 10] = (px * py) / 2 - pz / 2
        A[..., 6, 11] = py / 2 + (px * pz) / 2
        A[..., 7, 6] = (py * wx) / 2 - (px * wy) / 2 - wz / 2
        A[..., 7, 7] = d
        A[..., 7, 8] = wx / 2 + (py * wz) / 2 - (pz * wy) / 2
        A[..., 7, 9] = pz / 2 + (px * py) / 2
        A[..., 7, 10] = -px ** 2 / 4 + py ** 2 / 4 - pz ** 2 / 4 + 0.25
        A[..., 7, 11] = (py * pz) / 2 - px / 2
        A[..., 8, 6] = wy / 2 - (px * wz) / 2 + (pz * wx) / 2
        A[..., 8, 7] = (pz * wy) / 2 - (py * wz) / 2 - wx / 2
        A[..., 8, 8] = d
        A[..., 8, 9] = (px * pz) / 2 - py / 2
        A[..., 8, 10] = px / 2 + (py * pz) / 2
        A[..., 8, 11] = -px ** 2 / 4 - py ** 2 / 4 + pz ** 2 / 4 + 0.25
        A[..., 9, 10] = (Jyy - Jzz) * wz / Jxx
        A[..., 9, 11] = (Jyy - Jzz) * wy / Jxx
        A[..., 10, 9] = -(Jxx - Jzz) * wz / Jyy
        A[..., 10, 11] = -(Jxx - Jzz) * wx / Jyy
        A[..., 11, 9] = (Jxx - Jyy) * wy / Jzz
        A[..., 11, 10] = (Jxx - Jyy) * wx / Jzz
    elif m.model_id == ASTROBEE_SE3_MANIFOLD:
        # astrobee_se3_manifold.jl:248-296
        A[..., 0, 3] = A[..., 1, 4] = A[..., 2, 5] = 1.0
        Jxx, Jyy, Jzz = rp[1:4]
        qw, qx, qy, qz = X[..., 6], X[..., 7], X[..., 8], X[..., 9]
        wx, wy, wz = X[..., 10], X[..., 11], X[..., 12]
        A[..., 6, 7], A[..., 6, 8], A[..., 6, 9] = -wx / 2, -wy / 2, -wz / 2
        A[..., 6, 10], A[..., 6, 11], A[..., 6, 12] = -qx / 2, -qy / 2, -qz / 2
        A[..., 7, 6], A[..., 7, 8], A[..., 7, 9] = wx / 2, -wz / 2, wy / 2
        A[..., 7, 10], A[..., 7, 11], A[..., 7, 12] = qw / 2, qz / 2, -qy / 2
        A[..., 8, 6], A[..., 8, 7], A[..., 8, 9] = wy / 2, wz / 2, -wx / 2
        A[..., 8, 10], A[..., 8, 11], A[..., 8, 12] = -qz / 2, qw / 2, qx / 2
        A[..., 9, 6], A[..., 9, 7], A[..., 9, 8] = wz / 2, -wy / 2, wx / 2
        A[..., 9, 10], A[..., 9, 11], A[..., 9, 12] = qy / 2, -qx / 2, qw / 2
        A[..., 10, 11] = (Jyy - Jzz) * wz / Jxx
        A[..., 10, 12] = (Jyy - Jzz) * wy / Jxx
        A[..., 11, 10] = -(Jxx - Jzz) * wz / Jyy
        A[..., 11, 12] = -(Jxx - Jzz) * wx / Jyy
        A[..., 12, 10] = (Jxx - Jyy) * wy / Jzz
        A[..., 12, 11] = (Jxx - Jyy) * wx / Jzz
    elif m.model_id == FREEFLYER_SE2:
        # freeflyer_se2.jl:196-198
        A[..., 0, 3] = A[..., 1, 4] = A[..., 2, 5] = 1.0
    elif m.model_id == DUBINS:
        # dubins_car.jl:174-177
        A[..., 0, 2] = -rp[10] * np.sin(X[..., 2])
        A[..., 1, 2] = rp[10] * np.cos(X[..., 2])
    else:
        raise ValueError(m.model_id)
    return A


def B_dyn(m: ModelSpec):
    """Constant df/du, (n_x, n_u)."""
    rp = m.robot_params
    B = np.zeros((m.n_x, m.n_u))
    if m.model_id == ASTROBEE_SE3:
        # astrobee_se3.jl:235-241
        B[3:6, 0:3] = np.eye(3) / rp[0]
        B[9:12, 3:6] = np.diag(1 / rp[1:4])
    elif m.model_id == ASTROBEE_SE3_MANIFOLD:
        # astrobee_se3_manifold.jl:298-304
        B[3:6, 0:3] = np.eye(3) / rp[0]
        B[10:13, 3:6] = np.diag(1 / rp[1:4])
    elif m.model_id == FREEFLYER_SE2:
        # freeflyer_se2.jl:200-206
        B[3, 0] = B[4, 1] = 1 / rp[0]
        B[5, 2] = 1 / rp[1]
    elif m.model_id == DUBINS:
        # dubins_car.jl:179-181
        B[2, 0] = rp[11]
    return B
