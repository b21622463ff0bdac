"""Oracle (test infrastructure): closed-form signed distance replacing BulletCollision.distance.

The reference calls `BulletCollision.distance(env, rb_idx, r, env_idx) -> (dist, xbody, xobs)` (39 call
sites, e.g. dynamics/astrobee_se3.jl:291,403,409) through the un-vendored, unpinned BulletCollision.jl
(README.md:7).  Its contract, inferred from the call sites: the robot component is a
convex shape translated to `r` (never rotated); `dist` is the signed separation to obstacle `env_idx`
(<0 = penetration depth); callers build the outward normal as
    nhat = dist > 0 ? (xbody-xobs)/|.| : (xobs-xbody)/|.|      (astrobee_se3.jl:296-298)
For a sphere of radius R (Astrobee3D, robot/astrobee3D.jl:17-18,30) against an axis-aligned box or a
sphere the GJK/EPA answer has the closed form below; both branches of the callers' sign rule give the
same `nhat`, which is what this function returns directly.  For the Freeflyer body (vertical cylinder,
robot/freeflyer.jl:53-57) against boxes whose z-extent overlaps the cylinder the query reduces to a
circle against a rectangle in the table plane (ws_dim=2, nhat_z = 0).
"""
from dataclasses import dataclass
import numpy as np

BOX, SPHERE = 0, 1


@dataclass
class Obstacle:
    kind: int            # BOX: a=lo, b=hi ; SPHERE: a=center, b[0]=radius
    a: np.ndarray
    b: np.ndarray


def pack_obstacles(obstacles):
    """-> (kind int32[n], a float64[n,3], b float64[n,3])"""
    n = len(obstacles)
    kind = np.array([o.kind for o in obstacles], dtype=np.int32).reshape(n)
    a = np.array([o.a for o in obstacles], dtype=np.float64).reshape(n, 3)
    b = np.array([o.b for o in obstacles], dtype=np.float64).reshape(n, 3)
    return kind, a, b


def signed_distance(r, obstacles, R, ws_dim=3):
    """r:(...,3) robot centre(s).  Returns dist:(...,n_obs), nhat:(...,n_obs,3).

    Box, centre outside : c = clamp(r, lo, hi); d = |r-c|; nhat = (r-c)/d; dist = d - R.
    Box, centre inside  : q = max(lo-r, r-hi) (<=0); j = argmax q (first max on ties);
                          nhat = +-e_j toward that face; dist = q_j - R.
    Sphere (c, rho)     : d = |r-c|; nhat = (r-c)/d; dist = d - rho - R.
    """
    r = np.asarray(r, dtype=np.float64)
    kind, a, b = obstacles if isinstance(obstacles, tuple) else pack_obstacles(obstacles)
    n = kind.shape[0]
    D = ws_dim
    rr = r[..., None, :D]                                   # (...,1,D)
    dist = np.zeros(r.shape[:-1] + (n,))
    nhat = np.zeros(r.shape[:-1] + (n, 3))
    if n == 0:
        return dist, nhat
    lo, hi = a[:, :D], b[:, :D]
    # --- boxes
    q = np.maximum(lo - rr, rr - hi)                        # (...,n,D)
    outside = np.any(q > 0, axis=-1)
    c = np.clip(rr, lo, hi)
    diff = rr - c
    d = np.sqrt(np.sum(diff * diff, axis=-1))
    with np.errstate(invalid="ignore", divide="ignore"):
        n_out = diff / d[..., None]
    j = np.argmax(q, axis=-1)                               # first max on ties
    qj = np.take_along_axis(q, j[..., None], axis=-1)[..., 0]
    hi_side = np.take_along_axis(rr - hi >= lo - rr, j[..., None], axis=-1)[..., 0]
    n_in = np.zeros(q.shape)
    np.put_along_axis(n_in, j[..., None], np.where(hi_side, 1.0, -1.0)[..., None], axis=-1)
    box_dist = np.where(outside, d - R, qj - R)
    box_n = np.where(outside[..., None], n_out, n_in)
    # --- spheres
    cs = a[:, :D]
    ds = rr - cs
    dd = np.sqrt(np.sum(ds * ds, axis=-1))
    with np.errstate(invalid="ignore", divide="ignore"):
        sph_n = ds / dd[..., None]
    sph_dist = dd - b[:, 0] - R
    is_box = (kind == BOX)
    dist[...] = np.where(is_box, box_dist, sph_dist)
    nhat[..., :D] = np.where(is_box[:, None], box_n, sph_n)
    return dist, nhat
