"""Host-side mirror of the reference's Robot / DynamicsModel / Environment / Goal plugin surface.

These are the *configuration carriers* the reference user already builds (Astrobee3D(), AstrobeeSE3(),
ISSCorner(), GoalSet ...; /root/reference/src/robot/*.jl, src/dynamics/*.jl, src/environment/*.jl,
src/goals.jl).  Their numerical content is flattened into the plain arrays the C ABI takes
(include/gusto_b200.h): `robot_params[16]`, `scp_params[10]`, an obstacle table and per-instance
init/goal vectors.  The math itself lives in the CUDA kernels (csrc/), never here.
"""
import json
import os
from dataclasses import dataclass, field
import numpy as np

DUBINS, FREEFLYER_SE2, ASTROBEE_SE3, ASTROBEE_SE3_MANIFOLD = 0, 1, 2, 3
OBS_BOX, OBS_SPHERE = 0, 1
GOAL_FREE, GOAL_POINT, GOAL_BOX = 0, 1, 2

# robot_params slots (shared with csrc/gusto_types.h)
RP_MASS, RP_JXX, RP_JYY, RP_JZZ, RP_RADIUS, RP_VMAX, RP_AMAX, RP_WMAX, RP_ALMAX, RP_CLEAR = range(10)
RP_DUB_V, RP_DUB_K, RP_DUB_XMAX0, RP_DUB_XMAX1, RP_DUB_XMAX2, RP_DUB_UMAX = range(10, 16)
# scp_params slots
SP_DELTA0, SP_OMEGA0, SP_OMEGAMAX, SP_EPS, SP_RHO0, SP_RHO1, SP_BSUCC, SP_BFAIL, SP_GFAIL, SP_CONVTHR = range(10)


# ----------------------------------------------------------------------------------------------- robots
@dataclass
class Astrobee3D:
    """robot/astrobee3D.jl:16-30."""
    mass: float = 7.0
    J: tuple = (0.1083, 0.1083, 0.1083)          # diagonal inertia (the reference's A_dyn assumes diagonal J)
    r: float = float(np.sqrt(3.0) * 0.5 * 0.305)
    hard_limit_vel: float = 0.5
    hard_limit_accel: float = 0.1
    hard_limit_omega: float = 45 * np.pi / 180
    hard_limit_alpha: float = 50 * np.pi / 180


@dataclass
class Freeflyer:
    """robot/freeflyer.jl:29-50."""
    mass_ff: float = 0.5 * (15.36 + 18.08)
    J_ff: float = 0.184
    r: float = 0.157
    hard_limit_vel: float = 0.2
    hard_limit_accel: float = 2 * 0.185 / (0.5 * (15.36 + 18.08))
    hard_limit_omega: float = 20 * np.pi / 180
    hard_limit_alpha: float = (6.43 / 0.184) * 0.593


@dataclass
class Car:
    """robot/car.jl:3-6 (a point)."""
    r: float = 0.0


# ----------------------------------------------------------------------------------------------- models
@dataclass
class DynamicsModel:
    model_id: int
    name: str
    x_dim: int
    u_dim: int
    clearance: float
    scp_params: np.ndarray              # SCPParam_GuSTO(model) + SCPParam(model).convergence_threshold
    extra: dict = field(default_factory=dict)


def AstrobeeSE3():
    """dynamics/astrobee_se3.jl:16-40."""
    return DynamicsModel(ASTROBEE_SE3, "astrobeeSE3", 12, 6, 0.03,
                         np.array([10., 1., 1e10, 1e-6, 0.01, 0.05, 2., 0.5, 5., 0.01]))


def AstrobeeSE3Manifold():
    """dynamics/astrobee_se3_manifold.jl:18-46."""
    return DynamicsModel(ASTROBEE_SE3_MANIFOLD, "astrobeeSE3manifold", 13, 6, 0.03,
                         np.array([1000., 1., 1e10, 1e-1, 0.01, 100., 2., 0.5, 5., 1e-4]))


def FreeflyerSE2():
    """dynamics/freeflyer_se2.jl:14-39."""
    return DynamicsModel(FREEFLYER_SE2, "freeflyerSE2", 6, 3, 0.05,
                         np.array([3., 1., 1e10, 1e-2, 0.1, 0.3, 2., 0.5, 10., 1e-2]))


def DubinsCar():
    """dynamics/dubins_car.jl:20-52."""
    return DynamicsModel(DUBINS, "dubins", 3, 1, 0.01,
                         np.array([1e4, 1., 1e10, 1e-6, 0.4, 1.5, 2., 0.5, 5., 1e-4]),
                         extra=dict(v=2.0, k=1.0, x_max=(100., 100., 2 * np.pi), u_max=10.))


def robot_params(robot, model: DynamicsModel) -> np.ndarray:
    p = np.zeros(16)
    p[RP_CLEAR] = model.clearance
    if isinstance(robot, Astrobee3D):
        p[RP_MASS] = robot.mass
        p[RP_JXX:RP_JZZ + 1] = robot.J
        p[RP_RADIUS], p[RP_VMAX], p[RP_AMAX] = robot.r, robot.hard_limit_vel, robot.hard_limit_accel
        p[RP_WMAX], p[RP_ALMAX] = robot.hard_limit_omega, robot.hard_limit_alpha
    elif isinstance(robot, Freeflyer):
        p[RP_MASS] = robot.mass_ff
        p[RP_JXX:RP_JZZ + 1] = robot.J_ff
        p[RP_RADIUS], p[RP_VMAX], p[RP_AMAX] = robot.r, robot.hard_limit_vel, robot.hard_limit_accel
        p[RP_WMAX], p[RP_ALMAX] = robot.hard_limit_omega, robot.hard_limit_alpha
    elif isinstance(robot, Car):
        p[RP_MASS] = 1.0
        p[RP_DUB_V], p[RP_DUB_K] = model.extra["v"], model.extra["k"]
        p[RP_DUB_XMAX0:RP_DUB_XMAX2 + 1] = model.extra["x_max"]
        p[RP_DUB_UMAX] = model.extra["u_max"]
    else:
        raise TypeError(type(robot))
    return p


# ------------------------------------------------------------------------------------------ environments
@dataclass
class Environment:
    name: str
    keepin_zones: list = field(default_factory=list)      # [(lo[3], hi[3])]
    keepout_zones: list = field(default_factory=list)     # [(lo[3], hi[3])]
    obstacle_set: list = field(default_factory=list)      # [("box", lo, hi) | ("sphere", c, r)]

    def obstacle_table(self):
        """Collision components in the reference's order keepout_zones..., obstacle_set... (types.jl:19).
        -> (kind int32[n], a float64[n,3], b float64[n,3])"""
        kind, a, b = [], [], []
        for lo, hi in self.keepout_zones:
            kind.append(OBS_BOX); a.append(lo); b.append(hi)
        for o in self.obstacle_set:
            if o[0] == "box":
                kind.append(OBS_BOX); a.append(o[1]); b.append(o[2])
            else:
                kind.append(OBS_SPHERE); a.append(o[1]); b.append([o[2], 0.0, 0.0])
        n = len(kind)
        return (np.array(kind, dtype=np.int32).reshape(n), np.array(a, dtype=np.float64).reshape(n, 3),
                np.array(b, dtype=np.float64).reshape(n, 3))


_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _f32(v):
    return np.asarray(v, dtype=np.float32).astype(np.float64)


def ISSCorner(add_obstacles=False, mat_path=None):
    """environment/iss_corner.jl:11-39 (+ add_obstacles! :52-63).  With `mat_path` (or $GUSTO_ISS_CORNER_MAT) the geometry is read
    at run time from the reference's own src/environment/iss_corner.mat, as ISSCorner{T}() does with matread; otherwise from the
    packaged data/iss_corner.json, the same numbers extracted once by tools/extract_iss_corner.py (float32-rounded, quirk q7)."""
    mat_path = mat_path or os.environ.get("GUSTO_ISS_CORNER_MAT")
    if mat_path:
        from .trajio import load_iss_corner_mat
        d = load_iss_corner_mat(mat_path)
    else:
        with open(os.path.join(_DATA, "iss_corner.json")) as f:
            d = json.load(f)
    env = Environment("ISSCorner",
                      keepin_zones=[(np.array(z["lo"]), np.array(z["hi"])) for z in d["keepin_zones"]],
                      keepout_zones=[(np.array(z["lo"]), np.array(z["hi"])) for z in d["keepout_zones"]])
    if add_obstacles:
        for z in d["obstacle_rectangles"]:
            env.obstacle_set.append(("box", np.array(z["lo"]), np.array(z["hi"])))
        for z in d["obstacle_spheres"]:
            env.obstacle_set.append(("sphere", np.array(z["center"]), float(z["radius"])))
    return env


def Table(room="stanford"):
    """environment/table.jl:11-55: four 10 m keep-out slabs around the table."""
    if room == "ames":
        radius = 0.15 * np.sqrt(2)
        wmin, wmax = np.array([-0.5, -0.75, 0.]) - radius, np.array([0.75, 0.75, 0.001]) + radius
    else:
        wmin, wmax = np.zeros(3), np.array([12., 9., 0.001]) * 0.3048
    a = 10.0
    koz = [(np.array([wmax[0], -a, -a]), np.array([wmax[0] + a, a, a])),
           (np.array([wmin[0] - a, -a, -a]), np.array([wmin[0], a, a])),
           (np.array([-a, wmax[1], -a]), np.array([a, wmax[1] + a, a])),
           (np.array([-a, wmin[1] - a, -a]), np.array([a, wmin[1], a]))]
    return Environment("Table", keepin_zones=[(wmin, wmax)], keepout_zones=koz)


def BlankEnv():
    """environment/blankenv.jl:11-18."""
    return Environment("BlankEnv")


FREEFLYER_NOTEBOOK_CENTERS = [(0.460, 0.315), (0.201, 1.085), (0.540, 2.020), (1.374, 0.196), (1.063, 1.354),
                              (1.365, 2.322), (2.221, 0.548), (2.077, 1.443), (3.098, 1.186), (2.837, 2.064)]


def add_freeflyer_notebook_obstacles(env):
    """The ten inflated boxes of examples/freeflyerSE2.ipynb cell 2 (Vec3f0-rounded origin/widths)."""
    widths = np.array([0.27, 0.27, 0.127])
    infl = 0.05 * np.ones(3)
    for cx, cy in FREEFLYER_NOTEBOOK_CENTERS:
        origin = np.array([cx, cy, 0.0]) - 0.5 * widths - infl + np.array([0., 0., 0.5 * widths[0]])
        o32 = np.asarray(origin, dtype=np.float32)
        w32 = np.asarray(widths + 2 * infl, dtype=np.float32)
        hi = (o32 + w32).astype(np.float32)
        env.obstacle_set.append(("box", o32.astype(np.float64), hi.astype(np.float64)))
    return env


# ------------------------------------------------------------------------------------------------ goals
@dataclass
class PointGoal:
    point: np.ndarray


@dataclass
class BoxGoal:
    lower_bound: np.ndarray
    upper_bound: np.ndarray


@dataclass
class Goal:
    """goals.jl:1-16.  `ind_coordinates` is zero-based here."""
    params: object
    t_guess: float
    ind_coordinates: np.ndarray


class GoalSet:
    """goals.jl / types.jl:27-30.  Only goals at the final time take part in the GuSTO subproblem
    (SCPConstraints registries, e.g. astrobee_se3.jl:339-345)."""

    def __init__(self):
        self.goals = []

    def add_goal(self, goal: Goal):
        self.goals.append(goal)
        self.goals.sort(key=lambda g: g.t_guess)

    def flatten(self, x_dim, tf_guess):
        gtype = np.zeros(x_dim, dtype=np.int32)
        lo = np.zeros(x_dim)
        hi = np.zeros(x_dim)
        for g in self.goals:
            if g.t_guess != tf_guess:
                continue
            idx = np.asarray(g.ind_coordinates)
            if isinstance(g.params, PointGoal):
                gtype[idx], lo[idx], hi[idx] = GOAL_POINT, g.params.point, g.params.point
            else:
                gtype[idx], lo[idx], hi[idx] = GOAL_BOX, g.params.lower_bound, g.params.upper_bound
        return gtype, lo, hi


# SCPParam_TrajOpt(model): mu0, s0, c, tau_plus, tau_minus, k, ftol, xtol, ctol, max_penalty_iteration, max_convex_iteration,
# max_trust_iteration (dynamics/astrobee_se3.jl:50-64, freeflyer_se2.jl:49-63)
TRAJOPT_PARAMS = {
    ASTROBEE_SE3: np.array([1.0, 10.0, 10.0, 2.0, 0.5, 5.0, 0.01, 0.01, 0.01, 5, 5, 5]),
    FREEFLYER_SE2: np.array([1.0, 1.0, 10.0, 2.0, 0.5, 5.0, 0.01, 0.1, 0.01, 5, 5, 5]),
}
