"""Shared helpers for the test-suite: package loader, oracle bridge and the kernel host-simulation harness."""
import ctypes
import importlib.util
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def load_package():
    if "gusto_b200" in sys.modules:
        return sys.modules["gusto_b200"]
    d = os.path.join(ROOT, "gusto.jl_b200")
    spec = importlib.util.spec_from_file_location("gusto_b200", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gusto_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


gb = load_package()
import gusto_oracle as orc  # noqa: E402
from gusto_oracle.subproblem import Problem  # noqa: E402


def to_oracle(bp, b):
    """Instance b of a BatchProblem as an oracle Problem (checks that both sides carry the same parameter tables)."""
    m = orc.get_model(bp.model.name)
    assert np.array_equal(m.robot_params, bp.robot_params()), "robot parameter tables differ"
    assert np.array_equal(m.scp_params, bp.model.scp_params), "SCP parameter tables differ"
    return Problem(m, bp.N, float(bp.tf[b]), bp.x_init[b].copy(), bp.goal_type.copy(), bp.goal_lo[b].copy(),
                   bp.goal_hi[b].copy(), bp.obstacle_table())


# ------------------------------------------------------------------------------------ host simulation of the kernels
_HS = None


def hostsim_lib():
    global _HS
    if _HS is not None:
        return _HS
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    out = os.path.join(ROOT, "tests", "hostsim", "_build", "libgusto_hostsim.so")
    deps = [src] + [os.path.join(ROOT, "gusto.jl_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "gusto.jl_b200", "csrc"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", out])
    _HS = ctypes.CDLL(out)
    return _HS


def hostsim_iterate(bp, Xp, Up, omega, delta, stages=7, Xn=None, Un=None, **ipm_opts):
    """Run linearize (1) | solve (2) | evaluate (4) of the kernel bodies on the host.  Returns a dict of arrays."""
    host = gb.engine()
    cfg, (kind, a, b) = host.make_config(bp, 0, **ipm_opts)
    B, N, nx, nu = bp.B, bp.N, bp.model.x_dim, bp.model.u_dim
    no = int(kind.shape[0]) if bp.model.model_id != gb.models.DUBINS else 0
    dp = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    Xp = np.ascontiguousarray(Xp, dtype=np.float64).copy(); Up = np.ascontiguousarray(Up, dtype=np.float64).copy()
    Xn = np.zeros((B, N, nx)) if Xn is None else np.ascontiguousarray(Xn, dtype=np.float64).copy()
    Un = np.zeros((B, N, nu)) if Un is None else np.ascontiguousarray(Un, dtype=np.float64).copy()
    omega = np.ascontiguousarray(np.broadcast_to(omega, (B,)), dtype=np.float64).copy()
    delta = np.ascontiguousarray(np.broadcast_to(delta, (B,)), dtype=np.float64).copy()
    f = np.zeros((B, N, nx)); A = np.zeros((B, N, nx, nx)); g = np.zeros((B, N, nx)); rows = np.zeros((B, N, max(no, 1), 5))
    info = np.zeros((B, 8)); ev = np.zeros((B, 8))
    x_init = np.ascontiguousarray(bp.x_init); glo = np.ascontiguousarray(bp.goal_lo); ghi = np.ascontiguousarray(bp.goal_hi)
    tf = np.ascontiguousarray(bp.tf)
    rc = hostsim_lib().hostsim_iterate(ctypes.byref(cfg), kind.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), dp(a), dp(b),
                                       dp(x_init), dp(glo), dp(ghi), dp(tf), dp(Xp), dp(Up), dp(Xn), dp(Un), dp(omega),
                                       dp(delta), dp(f), dp(A), dp(g), dp(rows), ctypes.c_int(stages), dp(info), dp(ev))
    assert rc == 0
    return dict(f=f, A=A, g=g, rows=rows[:, :, :no], Xn=Xn, Un=Un, info=info, eval=ev)


def hostsim_postprocess(bp, X, U, nstep):
    """check_instance / interpolate_interval kernel bodies on the host.  Returns (chk[B,8], Xfull, Ufull)."""
    host = gb.engine()
    cfg, (kind, a, b) = host.make_config(bp, 0)
    B, N, nx, nu = bp.B, bp.N, bp.model.x_dim, bp.model.u_dim
    dp = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    X = np.ascontiguousarray(X, dtype=np.float64); U = np.ascontiguousarray(U, dtype=np.float64)
    tf = np.ascontiguousarray(bp.tf)
    nf = nstep * (N - 1)
    chk = np.zeros((B, 8)); Xf = np.zeros((B, nf + 1, nx)); Uf = np.zeros((B, nf, nu))
    rc = hostsim_lib().hostsim_postprocess(ctypes.byref(cfg), kind.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), dp(a), dp(b),
                                           dp(tf), dp(X), dp(U), ctypes.c_int(nstep), dp(chk), dp(Xf), dp(Uf))
    assert rc == 0
    return chk, Xf, Uf


def hostsim_iterate_with_duals(bp, Xp, Up, omega, delta, **kw):
    """hostsim_iterate that also returns SCPS.dual (the init-row multipliers written by the IPM body), [B, n_x]."""
    dual = np.zeros((bp.B, bp.model.x_dim))
    lib = hostsim_lib()
    lib.hostsim_set_dual_buffer(dual.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    try:
        hs = hostsim_iterate(bp, Xp, Up, omega, delta, **kw)
    finally:
        lib.hostsim_set_dual_buffer(None)
    hs["dual"] = dual
    return hs


def hostsim_shoot(bp, p0, x_goal, Xs_prev, nsub=4, max_iter=100, ftol=1e-3):
    """shoot_instance kernel body on the host.  Returns (out[B,8], Xs, Us, Ps)."""
    host = gb.engine()
    cfg, _ = host.make_config(bp, 0)
    B, N, nx, nu = bp.B, bp.N, bp.model.x_dim, bp.model.u_dim
    dp = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    p0 = np.ascontiguousarray(p0, dtype=np.float64); x_goal = np.ascontiguousarray(x_goal, dtype=np.float64)
    x_init = np.ascontiguousarray(bp.x_init); tf = np.ascontiguousarray(bp.tf)
    Xs = np.ascontiguousarray(Xs_prev, dtype=np.float64).copy(); Us = np.zeros((B, N, nu)); Ps = np.zeros((B, N, nx))
    out = np.zeros((B, 8))
    rc = hostsim_lib().hostsim_shoot(ctypes.byref(cfg), dp(x_init), dp(tf), dp(p0), dp(x_goal), ctypes.c_int(nsub), ctypes.c_int(max_iter),
                                     ctypes.c_double(ftol), dp(Xs), dp(Us), dp(Ps), dp(out))
    assert rc == 0
    return out, Xs, Us, Ps


def hostsim_solve_gusto(bp, max_iter=30):
    """solve_gusto_batch (host.py) with every kernel replaced by its host-simulated body: the L3 check of the kernel
    arithmetic in the GPU-less container.  Returns dict(converged, successful, iterations, J_true, accept[list], X, U)."""
    host = gb.engine()
    M = gb.models
    B, sp = bp.B, bp.model.scp_params
    X, U = bp.init_traj_straightline()
    Delta = np.full(B, sp[M.SP_DELTA0]); omega = np.full(B, sp[M.SP_OMEGA0])
    ev0 = hostsim_iterate(bp, X, U, omega, Delta, stages=5, Xn=X, Un=U)["eval"]
    J = ev0[:, 4].copy()
    iterations = np.zeros(B, np.int64); converged = np.zeros(B, bool); successful = np.zeros(B, bool)
    active = np.ones(B, bool); conv_prev = np.zeros(B); accept_hist = [np.ones(B, bool)]; newton = []
    for _ in range(max_iter):
        hs = hostsim_iterate(bp, X, U, omega, Delta)
        st = host.gusto_update(hs["eval"], host.solver_status_ok(hs["info"][:, 0]), active, Delta, omega, iterations, conv_prev, sp)
        acc = st["accept"]
        X = np.where(acc[:, None, None], hs["Xn"], X); U = np.where(acc[:, None, None], hs["Un"], U)
        J = np.where(acc, hs["eval"][:, 4], J)
        conv_prev = np.where(st["run"], hs["eval"][:, 0], conv_prev)
        accept_hist.append(acc.copy()); newton.append(np.where(active, hs["info"][:, 1], 0))
        Delta, omega, iterations = st["Delta"], st["omega"], st["iterations"]
        converged |= st["converged_now"]; successful |= st["successful_now"]
        active = active & ~st["done"]
        if not active.any():
            break
    return dict(converged=converged, successful=successful, iterations=iterations, J_true=J, accept=accept_hist, X=X, U=U,
                newton=newton)


# --------------------------------------------------------------------------------- TrajOpt variant on the host simulation
def hostsim_trajopt_iterate(bp, Xp, Up, mu, s, stages=7):
    """linearize | TrajOpt subproblem | TrajOpt evaluation kernel bodies on the host (omega carries mu, delta carries s)."""
    lib = hostsim_lib()
    lib.hostsim_set_algorithm(1)
    try:
        return hostsim_iterate(bp, Xp, Up, mu, s, stages=stages)
    finally:
        lib.hostsim_set_algorithm(0)


def hostsim_trajopt_ctol(bp, X, U, Xr, Ur):
    host = gb.engine()
    cfg, (kind, a, b) = host.make_config(bp, 0)
    dp = lambda arr: np.ascontiguousarray(arr, dtype=np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    out = np.zeros((bp.B, 2))
    keep = [np.ascontiguousarray(v, dtype=np.float64) for v in (bp.goal_lo, bp.goal_hi, bp.tf, X, U, Xr, Ur)]
    rc = hostsim_lib().hostsim_trajopt_ctol(ctypes.byref(cfg), kind.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), dp(a), dp(b),
                                            *[v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) for v in keep], dp(out))
    assert rc == 0
    return out


class HostsimTrajOptEngine:
    """Stands in for host.Engine under host.solve_trajopt_batch with every kernel replaced by its host-simulated body: the L3 check
    of the TrajOpt kernels' arithmetic and of the host loop in the GPU-less container."""

    def __init__(self, bp):
        self.bp, self.B = bp, bp.B

    def trajopt_enable(self):
        self.ref = [None, None]

    def set_trajectory(self, X, U):
        self.X, self.U = np.array(X, dtype=np.float64), np.array(U, dtype=np.float64)

    def set_candidate(self, X, U):
        self.Xn, self.Un = np.array(X, dtype=np.float64), np.array(U, dtype=np.float64)

    def set_penalties(self, mu, s):
        self.mu, self.s = np.array(mu, dtype=np.float64), np.array(s, dtype=np.float64)

    def linearize(self):
        pass

    def evaluate(self):
        return hostsim_iterate(self.bp, self.X, self.U, 1.0, 1.0, stages=5, Xn=self.Xn, Un=self.Un)["eval"]

    def trajopt_mark(self, slot, which=None):
        w = np.ones(self.B, bool) if which is None else np.asarray(which, bool)
        if self.ref[slot] is None:
            self.ref[slot] = (self.X.copy(), self.U.copy())
        self.ref[slot][0][w] = self.X[w]; self.ref[slot][1][w] = self.U[w]

    def trajopt_compare(self, slot):
        Xr, Ur = self.ref[slot]
        c = hostsim_trajopt_ctol(self.bp, self.X, self.U, Xr, Ur)
        out = np.zeros((self.B, 5))
        out[:, :2] = c
        h = self.bp.tf / (self.bp.N - 1)
        for b in range(self.B):
            out[b, 2] = np.max(np.linalg.norm(self.X[b] - Xr[b], axis=-1)) / np.max(np.linalg.norm(self.X[b], axis=-1))
            for col, Uv in ((3, self.U[b]), (4, Ur[b])):
                uu = np.sum(Uv * Uv, axis=-1)
                out[b, col] = np.sum(0.5 * h[b] * (uu[:-1] + uu[1:]))
        return out

    def trajopt_iterate(self, mu, s, active=None, out=None, info=None):
        hs = hostsim_trajopt_iterate(self.bp, self.X, self.U, mu, s)
        self.Xn, self.Un = hs["Xn"], hs["Un"]
        return hs["eval"], hs["info"]

    def accept(self, accept, omega=None, delta=None):
        a = np.asarray(accept, bool)
        self.X[a] = self.Xn[a]; self.U[a] = self.Un[a]

    def get_trajectory(self):
        return self.X.copy(), self.U.copy()
