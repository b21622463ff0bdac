"""Batched problem definitions and the synthetic-input generators of the BASELINE.json configs.

`BatchProblem` is the batched counterpart of the reference's ProblemDefinition + TrajectoryOptimizationProblem
(/root/reference/src/types.jl:32-63): one robot/model/environment shared by B independent instances, each
with its own x_init, goal set and tf_guess.  Generators follow SURVEY.md section 8(d) (configs C1-C5).
"""
from dataclasses import dataclass
import numpy as np

from . import models as M


@dataclass
class BatchProblem:
    robot: object
    model: M.DynamicsModel
    env: M.Environment
    N: int
    tf: np.ndarray            # [B]
    x_init: np.ndarray        # [B, x_dim]
    goal_type: np.ndarray     # [x_dim] int32 (shared by the batch)
    goal_lo: np.ndarray       # [B, x_dim]
    goal_hi: np.ndarray       # [B, x_dim]
    name: str = ""

    @property
    def B(self):
        return int(self.x_init.shape[0])

    def robot_params(self):
        return M.robot_params(self.robot, self.model)

    def obstacle_table(self):
        if self.model.model_id == M.DUBINS:           # dubins registers no obstacle rows (dubins_car.jl:184-226)
            return np.zeros(0, np.int32), np.zeros((0, 3)), np.zeros((0, 3))
        return self.env.obstacle_table()

    def init_traj_straightline(self):
        """init_traj_straightline (astrobee_se3.jl:99-113 and siblings): X = range(x_init, x_goal, N), U = 0,
        x_goal = centres of the final-time goals, zero where no goal is set."""
        x_goal = np.where(self.goal_type[None, :] != M.GOAL_FREE, 0.5 * (self.goal_lo + self.goal_hi), 0.0)
        s = np.linspace(0.0, 1.0, self.N)[None, :, None]
        X = self.x_init[:, None, :] + s * (x_goal - self.x_init)[:, None, :]
        X[:, -1, :] = x_goal
        U = np.zeros((self.B, self.N, self.model.u_dim))
        return np.ascontiguousarray(X), U

    def shard(self, rank, world):
        """Static contiguous shard of the batch: rank g owns [g*B/G, (g+1)*B/G)."""
        lo, hi = rank * self.B // world, (rank + 1) * self.B // world
        return BatchProblem(self.robot, self.model, self.env, self.N, self.tf[lo:hi].copy(), self.x_init[lo:hi].copy(),
                            self.goal_type.copy(), self.goal_lo[lo:hi].copy(), self.goal_hi[lo:hi].copy(), self.name)

    def instance(self, b):
        return BatchProblem(self.robot, self.model, self.env, self.N, self.tf[b:b + 1].copy(),
                            self.x_init[b:b + 1].copy(), self.goal_type.copy(), self.goal_lo[b:b + 1].copy(),
                            self.goal_hi[b:b + 1].copy(), self.name)


def quat2mrp(q):
    """utils/quat_functions.jl:218-226, q = [vector; scalar]."""
    q = np.asarray(q, dtype=np.float64)
    return q[..., :3] / (1.0 + q[..., 3:4])


def _random_quat(rng, n, min_scalar=0.3):
    out = np.zeros((n, 4))
    i = 0
    while i < n:
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        if q[3] < 0:
            q = -q
        if q[3] > min_scalar:
            out[i] = q
            i += 1
    return out


def _sdf_min(pts, table, R):
    """Smallest signed distance of spheres (radius R) centred at pts[n,3] to the obstacle table (generator-side
    rejection sampling only; the engine's signed distance lives in csrc/)."""
    kind, a, b = table
    out = np.full(pts.shape[0], np.inf)
    for k, lo, hi in zip(kind, a, b):
        if k == M.OBS_BOX:
            q = np.maximum(lo - pts, pts - hi)
            d = np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(q.max(axis=1), 0)
        else:
            d = np.linalg.norm(pts - lo, axis=1) - hi[0]
        out = np.minimum(out, d - R)
    return out


def _sample_box(rng, B, lo, hi, table, R, min_sd):
    out = np.zeros((B, 3))
    n = 0
    while n < B:
        pts = rng.uniform(lo, hi, size=(B, 3))
        pts = pts[_sdf_min(pts, table, R) >= min_sd]
        k = min(B - n, pts.shape[0])
        out[n:n + k] = pts[:k]
        n += k
    return out


def _iss_endpoints(rng, B, margin, hard):
    env = M.ISSCorner()
    z8, z5 = env.keepin_zones[7], env.keepin_zones[4]
    def sample(zone):
        lo, hi = zone[0] + margin, zone[1] - margin
        if not hard:                      # line-of-sight tier: stay inside the hatch cross-section
            lo = np.maximum(lo, [10.64, -np.inf, 4.48])
            hi = np.minimum(hi, [11.25, np.inf, 5.15])
        return rng.uniform(lo, hi, size=(B, 3))
    return sample(z8), sample(z5)


def config_astrobee_se3(B=1024, N=50, seed=None, hard=False, tf=70.0):
    """C3 / C4: Astrobee3D in ISSCorner, start in keep-in zone 8, goal in zone 5, random goal attitude."""
    rng = np.random.default_rng(B if seed is None else seed)
    robot, model, env = M.Astrobee3D(), M.AstrobeeSE3(), M.ISSCorner()
    margin = robot.r + model.clearance + 0.05
    r0, r1 = _iss_endpoints(rng, B, margin, hard)
    x_init = np.zeros((B, 12)); x_goal = np.zeros((B, 12))
    x_init[:, 0:3] = r0
    x_goal[:, 0:3] = r1
    x_goal[:, 6:9] = quat2mrp(_random_quat(rng, B))
    gs_type = np.full(12, M.GOAL_POINT, dtype=np.int32)
    return BatchProblem(robot, model, env, N, np.full(B, tf), x_init, gs_type, x_goal, x_goal.copy(),
                        name=f"astrobeeSE3{'-hard' if hard else ''} B={B} N={N}")


def config_astrobee_se3_manifold(B=1024, N=60, seed=None, hard=False, tf=70.0, eps_q=1e-4, tier="notebook"):
    """C5: quaternion state [r v qw qx qy qz w]; goals as examples/astrobeeSE3manifold.ipynb cell 1
    (PointGoal r, v, w; BoxGoal q +- 1e-4); add_obstacles! => 26 + 4 boxes + 2 spheres.
    tier "notebook" (default): endpoints around the notebook's own [11.2,-0.8,5.6] -> [10.9,3.0,5.0];
    tier "zones": SURVEY C5 read literally, "as C3": start anywhere in keep-in zone 8, goal anywhere in zone 5, both inside the
    hatch cross-section (line of sight), at least 0.1 m from every one of the 32 collision components."""
    rng = np.random.default_rng(B + 5 if seed is None else seed)
    robot, model, env = M.Astrobee3D(), M.AstrobeeSE3Manifold(), M.ISSCorner(add_obstacles=True)
    table = env.obstacle_table()
    if tier == "zones":
        margin = robot.r + model.clearance + 0.05
        z8, z5 = env.keepin_zones[7], env.keepin_zones[4]
        sect_lo, sect_hi = np.array([10.64, -np.inf, 4.48]), np.array([11.25, np.inf, 5.15])
        r0 = _sample_box(rng, B, np.maximum(z8[0] + margin, sect_lo), np.minimum(z8[1] - margin, sect_hi), table, robot.r, 0.1)
        r1 = _sample_box(rng, B, np.maximum(z5[0] + margin, sect_lo), np.minimum(z5[1] - margin, sect_hi), table, robot.r, 0.1)
    else:
        # Notebook-like difficulty (examples/astrobeeSE3manifold.ipynb: [11.2,-0.8,5.6] -> [10.9,3.0,5.0]): start above
        # the first obstacle box in module 8, goal just past the hatch and before the box that blocks module 5;
        # both endpoints at least 0.1 m (signed distance) from every collision component.
        r0 = _sample_box(rng, B, [10.70, -0.90, 5.30], [11.25, 0.90, 5.70], table, robot.r, 0.1)
        r1 = _sample_box(rng, B, [10.64, 2.90, 4.48], [11.25, 3.20, 5.15], table, robot.r, 0.1)
    x_init = np.zeros((B, 13)); lo = np.zeros((B, 13)); hi = np.zeros((B, 13))
    x_init[:, 0:3] = r0
    x_init[:, 6] = 1.0
    q = _random_quat(rng, B)                      # [vector; scalar] -> state order [qw qx qy qz]
    qg = np.concatenate([q[:, 3:4], q[:, 0:3]], axis=1)
    lo[:, 0:3] = hi[:, 0:3] = r1
    lo[:, 6:10], hi[:, 6:10] = qg - eps_q, qg + eps_q
    gtype = np.full(13, M.GOAL_POINT, dtype=np.int32)
    gtype[6:10] = M.GOAL_BOX
    return BatchProblem(robot, model, env, N, np.full(B, tf), x_init, gtype, lo, hi,
                        name=f"astrobeeSE3manifold{'-zones' if tier == 'zones' else ''} B={B} N={N}")


def config_freeflyer_se2(B=256, N=40, seed=None):
    """C2: Table(:stanford) + the ten notebook boxes; endpoints uniform on the table shrunk by 0.3 m with
    signed distance >= 0.1 m to every obstacle; theta ~ U(-pi, pi); tf = max(60, 20*|dr|)."""
    rng = np.random.default_rng(B if seed is None else seed)
    robot, model = M.Freeflyer(), M.FreeflyerSE2()
    env = M.add_freeflyer_notebook_obstacles(M.Table("stanford"))
    wmin, wmax = env.keepin_zones[0]
    kind, a, b = env.obstacle_table()

    def sd2(pt):                                   # circle-vs-rectangle signed distance (host-side rejection only)
        q = np.maximum(a[:, :2] - pt, pt - b[:, :2])
        out = np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(q.max(axis=1), 0)
        return out.min() - robot.r

    def sample():
        while True:
            pt = rng.uniform(wmin[:2] + 0.3, wmax[:2] - 0.3)
            if sd2(pt) >= 0.1:
                return pt
    x_init = np.zeros((B, 6)); x_goal = np.zeros((B, 6))
    for i in range(B):
        x_init[i, :2], x_goal[i, :2] = sample(), sample()
        x_init[i, 2], x_goal[i, 2] = rng.uniform(-np.pi, np.pi, size=2)
    tf = np.maximum(60.0, 20.0 * np.linalg.norm(x_goal[:, :2] - x_init[:, :2], axis=1))
    gtype = np.full(6, M.GOAL_POINT, dtype=np.int32)
    return BatchProblem(robot, model, env, N, tf, x_init, gtype, x_goal, x_goal.copy(), name=f"freeflyerSE2 B={B} N={N}")


def config_dubins(B=1, N=30, seed=None):
    """C1: examples/dubins_car.ipynb cell 1 (x_init = [2,2,2], PointGoal 0, tf = 10), N reduced to 30.
    For B > 1 the extra instances perturb x_init."""
    rng = np.random.default_rng(30 if seed is None else seed)
    robot, model, env = M.Car(), M.DubinsCar(), M.BlankEnv()
    x_init = np.tile(np.array([2.0, 2.0, 2.0]), (B, 1))
    if B > 1:
        x_init[1:] += rng.uniform(-0.5, 0.5, size=(B - 1, 3))
    goal = np.zeros((B, 3))
    return BatchProblem(robot, model, env, N, np.full(B, 10.0), x_init, np.full(3, M.GOAL_POINT, dtype=np.int32),
                        goal, goal.copy(), name=f"dubins B={B} N={N}")


def config_freeflyer_notebook(N=200):
    """The one recorded reference run: examples/freeflyerSE2.ipynb cell 2 (N=200, tf=200)."""
    robot, model = M.Freeflyer(), M.FreeflyerSE2()
    env = M.add_freeflyer_notebook_obstacles(M.Table("stanford"))
    x_init = np.array([[0.2, 2.4, 0, 0, 0, 0.]])
    x_goal = np.array([[3., 0.5, 0, 0.05, -0.05, 0]])
    return BatchProblem(robot, model, env, N, np.array([200.0]), x_init, np.full(6, M.GOAL_POINT, dtype=np.int32),
                        x_goal, x_goal.copy(), name="freeflyerSE2 notebook")


def config_astrobee_se3_notebook(N=30):
    """examples/astrobeeSE3.ipynb cell 1."""
    robot, model, env = M.Astrobee3D(), M.AstrobeeSE3(), M.ISSCorner()
    x_init = np.zeros((1, 12)); x_goal = np.zeros((1, 12))
    x_init[0, :3] = [11.2, -0.8, 5.6]
    x_goal[0, :3] = [10.2, 6.9, 4.2]
    return BatchProblem(robot, model, env, N, np.array([70.0]), x_init, np.full(12, M.GOAL_POINT, dtype=np.int32),
                        x_goal, x_goal.copy(), name="astrobeeSE3 notebook")


CONFIGS = {
    "dubins": config_dubins,
    "freeflyerSE2": config_freeflyer_se2,
    "astrobeeSE3": config_astrobee_se3,
    "astrobeeSE3manifold": config_astrobee_se3_manifold,
}
