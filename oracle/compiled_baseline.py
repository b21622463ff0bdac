"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- a compiled single-thread CPU build of the kernel bodies as a second CPU baseline.

tests/hostsim/hostsim.cpp compiles the CUDA kernel sources (gusto.jl_b200/csrc/*.cuh) with -DGUSTO_HOSTSIM into a sequential
host program (G_TID = 0, barriers are no-ops).  Built here with `g++ -O3 -march=native` into oracle/_build/, it is the same
structured interior-point algorithm as the GPU path running on one host core: an honest compiled CPU number next to the
NumPy/SciPy restatement (`cpu_baseline.kind = "port"`).  It is never linked into libgusto_b200.so and never reachable from the
product path; only bench.py's cpu_baseline legs call it.

The outer loop below is the reference's solve_gusto_jump! (/root/reference/src/scp/scp_gusto.jl:49-176) driven through
gusto.jl_b200/host.py::gusto_update, one instance per call.
"""
import ctypes
import os
import subprocess
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
OUT = os.path.join(ROOT, "oracle", "_build", "libgusto_hostsim_native.so")
_LIB = None


def build(force=False):
    csrc = os.path.join(ROOT, "gusto.jl_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if force or not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        tmp = OUT + f".{os.getpid()}.tmp"
        subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC, "-o", tmp])
        os.replace(tmp, OUT)
    return OUT


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def solve_instance(pkg, bp, max_iter=30, force=False):
    """Full GuSTO SCP of a 1-instance BatchProblem on one host core.  Returns (iterations, converged, seconds)."""
    host = pkg.engine()
    cfg, (kind, a, b) = host.make_config(bp, 0)
    B, N, nx, nu = bp.B, bp.N, bp.model.x_dim, bp.model.u_dim
    no = int(kind.shape[0]) if bp.model.model_id != pkg.models.DUBINS else 0
    dp = lambda arr: arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    sp = bp.model.scp_params
    X0, U0 = bp.init_traj_straightline()
    Xp = np.ascontiguousarray(X0, dtype=np.float64).copy(); Up = np.ascontiguousarray(U0, dtype=np.float64).copy()
    Xn = np.zeros((B, N, nx)); Un = np.zeros((B, N, nu))
    f = np.zeros((B, N, nx)); A = np.zeros((B, N, nx, nx)); g = np.zeros((B, N, nx)); rows = np.zeros((B, N, max(no, 1), 5))
    info = np.zeros((B, 8)); ev = np.zeros((B, 8))
    x_init = np.ascontiguousarray(bp.x_init); glo = np.ascontiguousarray(bp.goal_lo); ghi = np.ascontiguousarray(bp.goal_hi)
    tf = np.ascontiguousarray(bp.tf)
    Delta = np.full(B, sp[0]); omega = np.full(B, sp[1])
    its = np.zeros(B, np.int64); conv_prev = np.zeros(B); active = np.ones(B, bool); converged = np.zeros(B, bool)
    lib = _lib()
    t0 = time.perf_counter()
    steps = 0
    for _ in range(max_iter):
        om = omega.copy(); de = Delta.copy()
        rc = lib.hostsim_iterate(ctypes.byref(cfg), kind.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), dp(a), dp(b), dp(x_init), dp(glo),
                                 dp(ghi), dp(tf), dp(Xp), dp(Up), dp(Xn), dp(Un), dp(om), dp(de), dp(f), dp(A), dp(g), dp(rows),
                                 ctypes.c_int(7), dp(info), dp(ev))
        assert rc == 0
        steps += 1
        st = host.gusto_update(ev, host.solver_status_ok(info[:, 0]), active, Delta, omega, its, conv_prev, sp, force)
        acc = st["accept"]
        Xp[acc] = Xn[acc]; Up[acc] = Un[acc]
        conv_prev = np.where(st["run"], ev[:, 0], conv_prev)
        Delta, omega, its = st["Delta"], st["omega"], st["iterations"]
        converged |= st["converged_now"]
        active = active & ~st["done"]
        if not active.any():
            break
    return int(its.sum()), int(converged.sum()), steps, time.perf_counter() - t0
