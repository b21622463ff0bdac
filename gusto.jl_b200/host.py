"""ctypes binding over the C ABI (include/gusto_b200.h) and the host-side GuSTO outer loop.

`solve_gusto_batch` is the batched counterpart of `solve_gusto_jump!` (/root/reference/src/scp/scp_gusto.jl:49-176):
the accept/reject logic, the Delta/omega schedule and the convergence test run here on the host, per instance,
exactly as in the reference (:119-174); everything numerical runs in the CUDA kernels behind the C ABI.  The same
loop is mirrored in Julia in julia/GuSTOB200.jl.

There is no CPU path: if libgusto_b200.so is missing or no GPU is usable every entry point raises.
"""
import ctypes
import os
import time
from dataclasses import dataclass, field
import numpy as np

from . import models as M
from .problems import BatchProblem

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GUSTO_B200_LIB", os.path.join(_HERE, "libgusto_b200.so"))
EVAL_NOUT = 8
CHECK_NOUT = 8
SHOOT_NOUT = 8
SH_STATUS, SH_ITERS, SH_FNORM, SH_JTRUE, SH_CONV, SH_LAMBDA = range(6)
SOLVE_NINFO = 8
EV_CONV, EV_TR_OK, EV_INEQ_OK, EV_RHO, EV_JTRUE, EV_JFULL, EV_MAXDX2, EV_MAXSOFT = range(8)
HIST_W = 12
TRAJOPT_NOUT = 8
TO_XTOL, TO_RHO, TO_JTRUE, TO_JPREV, TO_MAXDX2, TO_NUM, TO_DEN, TO_DEFECT = range(8)
H_JTRUE, H_JFULL, H_SCP_STATUS, H_SOLVER_STATUS, H_ACCEPT, H_CONV, H_DELTA, H_OMEGA, H_RHO, H_TR_OK, H_INEQ_OK, H_NEWTON = range(12)
SCP_MAX_HIST = 64
SCP_STATUS = ("NA", "OK", "InaccurateModel", "ViolatesConstraints", "TrustRegionViolated", "SolverFailed", "Inactive")
ST_NA, ST_OK, ST_INACCURATE, ST_VIOLATES, ST_TRVIOLATED, ST_SOLVERFAIL, ST_INACTIVE = range(7)


class GustoConfig(ctypes.Structure):
    _fields_ = [("model_id", ctypes.c_int32), ("N", ctypes.c_int32), ("B", ctypes.c_int32), ("n_obs", ctypes.c_int32),
                ("robot_params", ctypes.c_double * 16), ("scp_params", ctypes.c_double * 10),
                ("goal_type", ctypes.c_int32 * 16), ("device", ctypes.c_int32),
                ("ipm_max_iter", ctypes.c_int32), ("ipm_nref", ctypes.c_int32),
                ("ipm_tol", ctypes.c_double), ("ipm_delta_p", ctypes.c_double), ("ipm_delta_d", ctypes.c_double)]


def make_config(bp: BatchProblem, device=0, ipm_max_iter=0, ipm_nref=0, ipm_tol=0.0, ipm_delta_p=0.0, ipm_delta_d=0.0):
    kind, a, b = bp.obstacle_table()
    cfg = GustoConfig()
    cfg.model_id, cfg.N, cfg.B, cfg.n_obs = bp.model.model_id, bp.N, bp.B, int(kind.shape[0])
    cfg.robot_params[:] = list(bp.robot_params())
    cfg.scp_params[:] = list(bp.model.scp_params)
    gt = np.zeros(16, dtype=np.int32)
    gt[:bp.model.x_dim] = bp.goal_type          # BoxGoal rows go to the solver as they are (no presolve since round 2)
    cfg.goal_type[:] = list(gt)
    cfg.device = device
    cfg.ipm_max_iter, cfg.ipm_nref = ipm_max_iter, ipm_nref
    cfg.ipm_tol, cfg.ipm_delta_p, cfg.ipm_delta_d = ipm_tol, ipm_delta_p, ipm_delta_d
    return cfg, (np.ascontiguousarray(kind, dtype=np.int32), np.ascontiguousarray(a), np.ascontiguousarray(b))


_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int32)
_BP = ctypes.POINTER(ctypes.c_uint8)


def _dp(a):
    return None if a is None else a.ctypes.data_as(_DP)


_lib = None


def load_library(path=LIB_PATH):
    """Load libgusto_b200.so and declare every symbol of include/gusto_b200.h.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(gusto-b200 has no CPU fallback)")
    lib = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.gusto_create.argtypes = [ctypes.POINTER(GustoConfig), _IP, _DP, _DP, ctypes.POINTER(vp)]
    lib.gusto_destroy.argtypes = [vp]
    lib.gusto_last_error.argtypes = [vp]
    lib.gusto_last_error.restype = ctypes.c_char_p
    lib.gusto_version.restype = i32
    lib.gusto_set_problems.argtypes = [vp, _DP, _DP, _DP, _DP]
    for name in ("gusto_set_trajectory", "gusto_get_trajectory", "gusto_get_candidate", "gusto_set_candidate",
                 "gusto_set_penalties"):
        getattr(lib, name).argtypes = [vp, _DP, _DP]
    lib.gusto_linearize.argtypes = [vp]
    lib.gusto_get_blocks.argtypes = [vp, _DP, _DP, _DP, _DP]
    lib.gusto_solve_subproblem.argtypes = [vp, _DP]
    lib.gusto_evaluate.argtypes = [vp, _DP]
    lib.gusto_accept.argtypes = [vp, _BP, _DP, _DP]
    lib.gusto_set_active.argtypes = [vp, _BP]
    lib.gusto_iterate.argtypes = [vp, _DP, _DP]
    lib.gusto_iterate_host.argtypes = [vp, _DP, _DP, _DP, _DP, _BP, _DP, _DP, _DP, _DP]
    lib.gusto_check_trajectory.argtypes = [vp, _DP]
    lib.gusto_interpolate_trajectory.argtypes = [vp, i32, _DP, _DP]
    lib.gusto_get_duals.argtypes = [vp, _DP]
    lib.gusto_shoot.argtypes = [vp, _DP, _DP, i32, i32, ctypes.c_double, _DP]
    lib.gusto_get_shooting_trajectory.argtypes = [vp, _DP, _DP, _DP]
    lib.gusto_set_shooting_trajectory.argtypes = [vp, _DP, _DP]
    lib.gusto_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.gusto_timer_start.argtypes = [vp]
    lib.gusto_timer_stop.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.gusto_launch_count.argtypes = [vp]
    lib.gusto_launch_count.restype = i64
    lib.gusto_device_ptr.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(i64)]
    lib.gusto_iterate_device.argtypes = [vp]
    lib.gusto_accept_device.argtypes = [vp, vp, vp, vp]
    lib.gusto_stream_handle.argtypes = [vp]
    lib.gusto_stream_handle.restype = i64
    lib.gusto_scp_begin.argtypes = [vp, _DP, _DP, i32]
    lib.gusto_scp_run.argtypes = [vp, i32, _IP, _IP]
    lib.gusto_scp_get.argtypes = [vp, _IP, _BP, _BP, _DP, i32, _IP]
    lib.gusto_comm_unique_id.argtypes = [_BP]
    lib.gusto_comm_init.argtypes = [vp, i32, i32, _BP]
    lib.gusto_allgather_status.argtypes = [vp, _BP, _BP, _IP]
    lib.gusto_trajopt_enable.argtypes = [vp]
    lib.gusto_trajopt_iterate.argtypes = [vp, _DP, _DP, _BP, _DP, _DP]
    lib.gusto_trajopt_mark.argtypes = [vp, i32, _BP]
    lib.gusto_trajopt_compare.argtypes = [vp, i32, _DP]
    for name in ("gusto_create", "gusto_destroy", "gusto_set_problems", "gusto_set_trajectory", "gusto_get_trajectory",
                 "gusto_get_candidate", "gusto_set_candidate", "gusto_set_penalties", "gusto_linearize",
                 "gusto_get_blocks", "gusto_solve_subproblem", "gusto_evaluate", "gusto_accept", "gusto_set_active",
                 "gusto_iterate", "gusto_iterate_host", "gusto_last_kernel_ms", "gusto_device_ptr", "gusto_iterate_device",
                 "gusto_accept_device", "gusto_timer_start", "gusto_timer_stop", "gusto_check_trajectory",
                 "gusto_interpolate_trajectory", "gusto_get_duals", "gusto_shoot", "gusto_get_shooting_trajectory",
                 "gusto_set_shooting_trajectory", "gusto_scp_begin", "gusto_scp_run", "gusto_scp_get", "gusto_comm_unique_id", "gusto_comm_init",
                 "gusto_allgather_status", "gusto_trajopt_enable", "gusto_trajopt_iterate", "gusto_trajopt_mark", "gusto_trajopt_compare"):
        getattr(lib, name).restype = i32
    _lib = lib
    return lib


class GustoError(RuntimeError):
    pass


class Engine:
    """One gusto_ctx: a batch of B instances resident on one GPU."""

    def __init__(self, bp: BatchProblem, device=0, **ipm_opts):
        self.lib = load_library()
        self.bp = bp
        self.B, self.N, self.nx, self.nu = bp.B, bp.N, bp.model.x_dim, bp.model.u_dim
        cfg, (kind, a, b) = make_config(bp, device, **ipm_opts)
        self.n_obs = int(kind.shape[0]) if bp.model.model_id != M.DUBINS else 0
        self._ctx = ctypes.c_void_p()
        rc = self.lib.gusto_create(ctypes.byref(cfg), kind.ctypes.data_as(_IP), _dp(a), _dp(b), ctypes.byref(self._ctx))
        if rc != 0:
            raise GustoError(f"gusto_create failed ({rc}): {self.lib.gusto_last_error(None).decode()}")
        self._chk(self.lib.gusto_set_problems(self._ctx, _dp(np.ascontiguousarray(bp.x_init)),
                                              _dp(np.ascontiguousarray(bp.goal_lo)), _dp(np.ascontiguousarray(bp.goal_hi)),
                                              _dp(np.ascontiguousarray(bp.tf))))

    def _chk(self, rc):
        if rc != 0:
            raise GustoError(f"gusto call failed ({rc}): {self.lib.gusto_last_error(self._ctx).decode()}")

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.gusto_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data movement
    def set_trajectory(self, X, U):
        self._chk(self.lib.gusto_set_trajectory(self._ctx, _dp(np.ascontiguousarray(X)), _dp(np.ascontiguousarray(U))))

    def set_candidate(self, X, U):
        self._chk(self.lib.gusto_set_candidate(self._ctx, _dp(np.ascontiguousarray(X)), _dp(np.ascontiguousarray(U))))

    def _get(self, fn, X=None, U=None):
        X = np.empty((self.B, self.N, self.nx)) if X is None else X
        U = np.empty((self.B, self.N, self.nu)) if U is None else U
        self._chk(fn(self._ctx, _dp(X), _dp(U)))
        return X, U

    def get_trajectory(self, X=None, U=None):
        return self._get(self.lib.gusto_get_trajectory, X, U)

    def get_candidate(self, X=None, U=None):
        return self._get(self.lib.gusto_get_candidate, X, U)

    def set_penalties(self, omega, delta):
        self._chk(self.lib.gusto_set_penalties(self._ctx, _dp(np.ascontiguousarray(omega, dtype=np.float64)),
                                               _dp(np.ascontiguousarray(delta, dtype=np.float64))))

    def set_active(self, active):
        a = np.ascontiguousarray(active, dtype=np.uint8)
        self._chk(self.lib.gusto_set_active(self._ctx, a.ctypes.data_as(_BP)))

    # -- kernels
    def linearize(self):
        self._chk(self.lib.gusto_linearize(self._ctx))

    def get_blocks(self):
        BN = self.B * self.N
        f = np.empty((self.B, self.N, self.nx)); A = np.empty((self.B, self.N, self.nx, self.nx))
        g = np.empty((self.B, self.N, self.nx)); rows = np.zeros((self.B, self.N, self.n_obs, 5))
        self._chk(self.lib.gusto_get_blocks(self._ctx, _dp(f), _dp(A), _dp(g), _dp(rows) if self.n_obs else None))
        return f, A, g, rows

    def solve_subproblem(self, info=None):
        info = np.empty((self.B, SOLVE_NINFO)) if info is None else info
        self._chk(self.lib.gusto_solve_subproblem(self._ctx, _dp(info)))
        return info

    def evaluate(self, out=None):
        out = np.empty((self.B, EVAL_NOUT)) if out is None else out
        self._chk(self.lib.gusto_evaluate(self._ctx, _dp(out)))
        return out

    def accept(self, accept, omega=None, delta=None):
        a = np.ascontiguousarray(accept, dtype=np.uint8)
        self._chk(self.lib.gusto_accept(self._ctx, a.ctypes.data_as(_BP),
                                        _dp(None if omega is None else np.ascontiguousarray(omega, dtype=np.float64)),
                                        _dp(None if delta is None else np.ascontiguousarray(delta, dtype=np.float64))))

    def iterate(self, out=None, info=None):
        out = np.empty((self.B, EVAL_NOUT)) if out is None else out
        info = np.empty((self.B, SOLVE_NINFO)) if info is None else info
        self._chk(self.lib.gusto_iterate(self._ctx, _dp(out), _dp(info)))
        return out, info

    def iterate_host(self, X, U, omega, delta, active, out, info, Xn, Un):
        """One outer iteration with host-resident trajectories: H2D of (X, U, omega, delta, active), the three kernels, D2H of
        (out, info, Xn, Un), one synchronisation.  All arrays must be C-contiguous (float64; active uint8); any input may be None."""
        ab = None if active is None else active.ctypes.data_as(_BP)
        self._chk(self.lib.gusto_iterate_host(self._ctx, _dp(X), _dp(U), _dp(omega), _dp(delta), ab, _dp(out), _dp(info), _dp(Xn), _dp(Un)))

    def iterate_device(self):
        self._chk(self.lib.gusto_iterate_device(self._ctx))

    # ---- TrajOpt variant (solve_trajopt_jump!, scp_trajopt.jl)
    def trajopt_enable(self):
        self._chk(self.lib.gusto_trajopt_enable(self._ctx))

    def trajopt_iterate(self, mu, s, active=None, out=None, info=None):
        """linearize -> TrajOpt subproblem -> evaluation of the candidate: out[B, TRAJOPT_NOUT], info[B, SOLVE_NINFO]."""
        out = np.empty((self.B, TRAJOPT_NOUT)) if out is None else out
        info = np.empty((self.B, SOLVE_NINFO)) if info is None else info
        ab = None if active is None else np.ascontiguousarray(active, dtype=np.uint8).ctypes.data_as(_BP)
        self._chk(self.lib.gusto_trajopt_iterate(self._ctx, _dp(np.ascontiguousarray(mu, dtype=np.float64)),
                                                 _dp(np.ascontiguousarray(s, dtype=np.float64)), ab, _dp(out), _dp(info)))
        return out, info

    def trajopt_mark(self, slot, which=None):
        wb = None if which is None else np.ascontiguousarray(which, dtype=np.uint8).ctypes.data_as(_BP)
        self._chk(self.lib.gusto_trajopt_mark(self._ctx, slot, wb))

    def trajopt_compare(self, slot):
        out = np.empty((self.B, 5))
        self._chk(self.lib.gusto_trajopt_compare(self._ctx, slot, _dp(out)))
        return out

    def accept_device(self, accept_ptr, omega_ptr, delta_ptr):
        self._chk(self.lib.gusto_accept_device(self._ctx, accept_ptr, omega_ptr, delta_ptr))

    def device_ptr(self, which):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        self._chk(self.lib.gusto_device_ptr(self._ctx, which, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def check_trajectory(self):
        """Post-processing scalars of the accepted trajectory, [B, CHECK_NOUT] (see include/gusto_b200.h)."""
        out = np.empty((self.B, CHECK_NOUT))
        self._chk(self.lib.gusto_check_trajectory(self._ctx, _dp(out)))
        return out

    def interpolate(self, nstep):
        """interpolate_traj: RK4 upsampling of the accepted trajectory, nstep sub-steps per knot interval."""
        nf = int(nstep) * (self.N - 1)
        Xf = np.empty((self.B, nf + 1, self.nx)); Uf = np.empty((self.B, nf, self.nu))
        self._chk(self.lib.gusto_interpolate_trajectory(self._ctx, int(nstep), _dp(Xf), _dp(Uf)))
        return Xf, Uf

    def get_duals(self):
        """SCPS.dual of the last solve: minus the JuMP dual of the init constraints (scp_gusto.jl:116), [B, n_x]."""
        d = np.empty((self.B, self.nx))
        self._chk(self.lib.gusto_get_duals(self._ctx, _dp(d)))
        return d

    def shoot(self, p0=None, x_goal=None, nsub=4, max_iter=100, ftol=1e-3):
        """One shooting attempt per instance (solve!(SS, SP), shooting.jl:4-49).  Returns out[B, SHOOT_NOUT]."""
        out = np.empty((self.B, SHOOT_NOUT))
        p0 = None if p0 is None else np.ascontiguousarray(p0, dtype=np.float64)
        x_goal = None if x_goal is None else np.ascontiguousarray(x_goal, dtype=np.float64)
        self._chk(self.lib.gusto_shoot(self._ctx, _dp(p0), _dp(x_goal), int(nsub), int(max_iter), float(ftol), _dp(out)))
        return out

    def set_shooting_trajectory(self, X, U):
        """Seed SS.traj (the reference starts it as deepcopy(traj_init), traj_opt.jl:16)."""
        self._chk(self.lib.gusto_set_shooting_trajectory(self._ctx, _dp(np.ascontiguousarray(X, dtype=np.float64)),
                                                         _dp(np.ascontiguousarray(U, dtype=np.float64))))

    def get_shooting_trajectory(self):
        X = np.empty((self.B, self.N, self.nx)); U = np.empty((self.B, self.N, self.nu)); P = np.empty((self.B, self.N, self.nx))
        self._chk(self.lib.gusto_get_shooting_trajectory(self._ctx, _dp(X), _dp(U), _dp(P)))
        return X, U, P

    def kernel_ms(self):
        ms = (ctypes.c_float * 4)()
        self._chk(self.lib.gusto_last_kernel_ms(self._ctx, ms))
        return dict(linearize=ms[0], solve=ms[1], evaluate=ms[2], accept=ms[3])

    def timer_start(self):
        self._chk(self.lib.gusto_timer_start(self._ctx))

    def timer_stop(self):
        ms = ctypes.c_float()
        self._chk(self.lib.gusto_timer_stop(self._ctx, ctypes.byref(ms)))
        return float(ms.value)

    def launch_count(self):
        return int(self.lib.gusto_launch_count(self._ctx))

    # -- device-resident outer loop and the status all-gather (include/gusto_b200.h, "Device-resident outer loop")
    def scp_begin(self, X0=None, U0=None, force=False):
        X0 = None if X0 is None else np.ascontiguousarray(X0, dtype=np.float64)
        U0 = None if U0 is None else np.ascontiguousarray(U0, dtype=np.float64)
        self._chk(self.lib.gusto_scp_begin(self._ctx, _dp(X0), _dp(U0), int(bool(force))))

    def scp_run(self, max_iter):
        """Up to max_iter outer iterations on the device; returns (iterations that ran, unfinished instances over all ranks)."""
        n, u = ctypes.c_int32(), ctypes.c_int32()
        self._chk(self.lib.gusto_scp_run(self._ctx, int(max_iter), ctypes.byref(n), ctypes.byref(u)))
        return int(n.value), int(u.value)

    def scp_get(self, n_hist, want_hist=True):
        """(iterations[B], converged[B], successful[B], hist[n_hist, B, HIST_W] or None, counters[n_hist - 1, 4])."""
        it = np.empty(self.B, np.int32); cv = np.empty(self.B, np.uint8); su = np.empty(self.B, np.uint8)
        hist = np.empty((n_hist, self.B, HIST_W)) if want_hist else None
        cnt = np.zeros((max(n_hist - 1, 0), 4), np.int32)
        self._chk(self.lib.gusto_scp_get(self._ctx, it.ctypes.data_as(_IP), cv.ctypes.data_as(_BP), su.ctypes.data_as(_BP), _dp(hist),
                                         int(n_hist), cnt.ctypes.data_as(_IP) if n_hist > 1 else None))
        return it.astype(np.int64), cv.astype(bool), su.astype(bool), hist, cnt

    def comm_init(self, rank, nranks, uid):
        u = np.ascontiguousarray(uid, dtype=np.uint8)
        assert u.size == 128
        self._chk(self.lib.gusto_comm_init(self._ctx, int(rank), int(nranks), u.ctypes.data_as(_BP)))
        self.nranks = int(nranks)

    def allgather_status(self, done_local):
        """One all-gather of this rank's B status bytes; returns (done_all[nranks * B], unfinished instances over all ranks)."""
        d = np.ascontiguousarray(done_local, dtype=np.uint8)
        out = np.empty(self.B * getattr(self, "nranks", 1), np.uint8)
        n = ctypes.c_int32()
        self._chk(self.lib.gusto_allgather_status(self._ctx, d.ctypes.data_as(_BP), out.ctypes.data_as(_BP), ctypes.byref(n)))
        return out, int(n.value)


# ------------------------------------------------------------------------------------------- outer loop
@dataclass
class BatchSCPSolution:
    """Per-instance SCPSolution histories (types.jl:150-173) + SCPParam_GuSTO vectors (scp_gusto.jl:15-19)."""
    X: np.ndarray
    U: np.ndarray
    converged: np.ndarray
    successful: np.ndarray
    iterations: np.ndarray
    J_true: list = field(default_factory=list)            # list of [B] arrays, one per history entry
    J_full: list = field(default_factory=list)
    scp_status: list = field(default_factory=list)        # int codes, SCP_STATUS
    solver_status: list = field(default_factory=list)
    accept_solution: list = field(default_factory=list)
    convergence_measure: list = field(default_factory=list)
    Delta_vec: list = field(default_factory=list)
    omega_vec: list = field(default_factory=list)
    rho_vec: list = field(default_factory=list)
    tr_ok_vec: list = field(default_factory=list)
    ineq_ok_vec: list = field(default_factory=list)
    newton_iters: list = field(default_factory=list)
    iter_elapsed_times: list = field(default_factory=list)
    total_time: float = 0.0
    batch_iterations: int = 0
    counters: np.ndarray = None                           # device loop only: per iteration [ran, solved, accepted, still live]


IPM_OPTIMAL, IPM_ITERATION_LIMIT, IPM_NUMERICAL, IPM_ALMOST_OPTIMAL = range(4)


def solver_status_ok(status):
    """scp_gusto.jl:107: OPTIMAL / LOCALLY_SOLVED / ALMOST_LOCALLY_SOLVED continue, anything else returns."""
    return (status == IPM_OPTIMAL) | (status == IPM_ALMOST_OPTIMAL)


def gusto_update(ev, solver_ok, active, Delta, omega, iterations, conv_prev, sp, force=False):
    """Vectorised accept/reject + Delta/omega schedule + convergence test of scp_gusto.jl:119-174 for one outer
    iteration.  All arguments are [B] arrays; returns the new state.  Pure host logic (also used by the gloo tests)."""
    D0, w_max, rho0, rho1 = sp[M.SP_DELTA0], sp[M.SP_OMEGAMAX], sp[M.SP_RHO0], sp[M.SP_RHO1]
    b_succ, b_fail, g_fail, thr = sp[M.SP_BSUCC], sp[M.SP_BFAIL], sp[M.SP_GFAIL], sp[M.SP_CONVTHR]
    conv, tr_ok, ineq_ok, rho = ev[:, EV_CONV], ev[:, EV_TR_OK] > 0.5, ev[:, EV_INEQ_OK] > 0.5, ev[:, EV_RHO]
    run = active & solver_ok
    failed = active & ~solver_ok                                  # :107-111 early return
    inaccurate = run & tr_ok & (rho > rho1)
    accepted = run & tr_ok & ~(rho > rho1)
    tr_viol = run & ~tr_ok
    status = np.full(active.shape, ST_INACTIVE, dtype=np.int32)
    status[failed] = ST_SOLVERFAIL
    status[inaccurate] = ST_INACCURATE
    status[accepted & ineq_ok] = ST_OK
    status[accepted & ~ineq_ok] = ST_VIOLATES
    status[tr_viol] = ST_TRVIOLATED
    Delta_n, omega_n = Delta.copy(), omega.copy()
    Delta_n[inaccurate] = b_fail * Delta[inaccurate]
    grow = accepted & (rho < rho0)
    Delta_n[grow] = np.minimum(b_succ * Delta[grow], D0)
    esc = (accepted & ~ineq_ok) | tr_viol
    omega_n[esc] = g_fail * omega[esc]
    iterations_n = iterations + run.astype(iterations.dtype)
    omega_exceeded = run & (omega_n > w_max)                      # :163-166
    conv_test = accepted & ~omega_exceeded & (iterations_n > 2) & (conv + conv_prev <= thr)   # :167-174
    converged_now = conv_test
    successful_now = conv_test & ineq_ok
    done = failed | omega_exceeded | (converged_now & (not force))
    return dict(accept=accepted, status=status, Delta=Delta_n, omega=omega_n, iterations=iterations_n,
                converged_now=converged_now, successful_now=successful_now, done=done, run=run)


def solve_gusto_batch(engine: Engine, X0=None, U0=None, max_iter=30, force=False, verbose=False, all_done=None):
    """Batched solve_gusto_jump!.  `all_done(local_done_flags) -> bool` lets a multi-GPU caller plug in the status
    all-gather (one collective per outer iteration); default is the local decision."""
    bp = engine.bp
    B = bp.B
    sp = bp.model.scp_params
    if X0 is None:
        X0, U0 = bp.init_traj_straightline()
    t0 = time.perf_counter()
    engine.set_trajectory(X0, U0)
    Delta = np.full(B, sp[M.SP_DELTA0]); omega = np.full(B, sp[M.SP_OMEGA0])
    engine.set_penalties(omega, Delta)
    # :72-75  initialize_model_params!, J_true[1] = cost_true(traj), rho_vec[2] = ratio(traj, traj)
    engine.set_candidate(X0, U0)
    engine.linearize()
    ev0 = engine.evaluate()
    S = BatchSCPSolution(X0, U0, np.zeros(B, bool), np.zeros(B, bool), np.zeros(B, np.int64))
    S.J_true.append(ev0[:, EV_JTRUE].copy()); S.J_full.append(ev0[:, EV_JTRUE].copy())
    S.scp_status.append(np.full(B, ST_NA, np.int32)); S.solver_status.append(np.full(B, -1, np.int32))
    S.accept_solution.append(np.ones(B, bool)); S.convergence_measure.append(np.zeros(B))
    S.Delta_vec.append(Delta.copy()); S.omega_vec.append(omega.copy())
    S.rho_vec += [np.zeros(B), ev0[:, EV_RHO].copy()]
    S.tr_ok_vec.append(np.zeros(B, bool)); S.ineq_ok_vec.append(np.zeros(B, bool))
    active = np.ones(B, bool)
    iter_cap = max_iter
    out = np.empty((B, EVAL_NOUT)); info = np.empty((B, SOLVE_NINFO))
    for it in range(iter_cap):
        ti = time.perf_counter()
        engine.set_active(active)
        engine.iterate(out, info)
        solver_ok = solver_status_ok(info[:, 0])
        st = gusto_update(out, solver_ok, active, Delta, omega, S.iterations, S.convergence_measure[-1], sp, force)
        engine.accept(st["accept"], st["omega"], st["Delta"])
        J_prev = S.J_true[-1]
        S.J_true.append(np.where(st["accept"], out[:, EV_JTRUE], J_prev))
        S.J_full.append(np.where(st["run"], info[:, 4], S.J_full[-1]))
        S.scp_status.append(st["status"]); S.solver_status.append(np.where(active, info[:, 0], -1).astype(np.int32))
        S.accept_solution.append(st["accept"])
        S.convergence_measure.append(np.where(st["run"], out[:, EV_CONV], S.convergence_measure[-1]))
        S.rho_vec.append(np.where(st["run"] & (out[:, EV_TR_OK] > 0.5), out[:, EV_RHO], np.nan))
        S.tr_ok_vec.append(out[:, EV_TR_OK] > 0.5); S.ineq_ok_vec.append(out[:, EV_INEQ_OK] > 0.5)
        S.newton_iters.append(np.where(active, info[:, 1], 0))
        Delta, omega = st["Delta"], st["omega"]
        S.Delta_vec.append(Delta.copy()); S.omega_vec.append(omega.copy())
        S.iterations = st["iterations"]
        S.converged |= st["converged_now"]; S.successful |= st["successful_now"]
        active = active & ~st["done"]
        S.iter_elapsed_times.append(time.perf_counter() - ti)
        S.batch_iterations += 1
        if verbose:
            print(f"[gusto] it {it + 1:2d} active {int(active.sum()):5d} accepted {int(st['accept'].sum()):5d} "
                  f"converged {int(S.converged.sum()):5d} newton {info[:, 1].mean():.1f} "
                  f"ms/it {1e3 * S.iter_elapsed_times[-1]:.2f}")
        finished = (not active.any()) if all_done is None else all_done(~active)
        if finished:
            break
    S.X, S.U = engine.get_trajectory()
    S.total_time = time.perf_counter() - t0
    return S


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0 calls it; the host broadcasts the 128 bytes to the other ranks)."""
    uid = np.zeros(128, np.uint8)
    rc = load_library().gusto_comm_unique_id(uid.ctypes.data_as(_BP))
    if rc != 0:
        raise GustoError(f"gusto_comm_unique_id failed ({rc}): {load_library().gusto_last_error(None).decode()}")
    return uid


def solve_gusto_batch_device(engine: Engine, X0=None, U0=None, max_iter=30, force=False):
    """solve_gusto_batch with the outer loop resident on the device (gusto_scp_begin / gusto_scp_run): same decisions, same
    histories, no host round trip per iteration; with a communicator (Engine.comm_init) every rank stops together."""
    bp = engine.bp
    B = bp.B
    if max_iter > SCP_MAX_HIST:
        raise ValueError(f"max_iter <= {SCP_MAX_HIST}")
    if X0 is None:
        X0, U0 = bp.init_traj_straightline()
    t0 = time.perf_counter()
    engine.scp_begin(X0, U0, force)
    n_it, _ = engine.scp_run(max_iter)
    it, cv, su, hist, cnt = engine.scp_get(n_it + 1)
    X, U = engine.get_trajectory()
    S = BatchSCPSolution(X, U, cv, su, it)
    S.total_time = time.perf_counter() - t0
    S.batch_iterations = n_it
    S.counters = cnt
    S.rho_vec.append(np.zeros(B))
    for h in range(n_it + 1):
        r = hist[h]
        S.J_true.append(r[:, H_JTRUE].copy()); S.J_full.append(r[:, H_JFULL].copy())
        S.scp_status.append(r[:, H_SCP_STATUS].astype(np.int32)); S.solver_status.append(r[:, H_SOLVER_STATUS].astype(np.int32))
        S.accept_solution.append(r[:, H_ACCEPT] > 0.5); S.convergence_measure.append(r[:, H_CONV].copy())
        S.Delta_vec.append(r[:, H_DELTA].copy()); S.omega_vec.append(r[:, H_OMEGA].copy()); S.rho_vec.append(r[:, H_RHO].copy())
        S.tr_ok_vec.append(r[:, H_TR_OK] > 0.5); S.ineq_ok_vec.append(r[:, H_INEQ_OK] > 0.5)
        if h:
            S.newton_iters.append(r[:, H_NEWTON].copy())
    return S


# ------------------------------------------------------------------------------------- SCP + shooting
@dataclass
class BatchShootingSolution:
    """Per-instance ShootingSolution histories (types.jl:197-208) next to the SCP solution (TrajectoryOptimizationSolution)."""
    SCPS: BatchSCPSolution
    X: np.ndarray                       # TOS.traj: the shooting trajectory where shooting converged, else SCPS.traj
    U: np.ndarray
    converged: np.ndarray               # SS.converged
    prob_status: list = field(default_factory=list)          # per attempt [B]: 0 :Optimal, 1 :Diverged, -1 not attempted
    J_true: list = field(default_factory=list)
    convergence_measure: list = field(default_factory=list)
    newton_iters: list = field(default_factory=list)
    attempts: int = 0


def solve_scp_shooting_batch(engine: Engine, X0=None, U0=None, max_iter=30, nsub=4, shoot_max_iter=100, ftol=1e-3, verbose=False):
    """Batched solve_SCPshooting! (traj_opt.jl:4-45): one GuSTO iteration, then alternate {shooting attempt started from the
    init-constraint duals of the last convex solve, one more GuSTO iteration} until the last two successful shooting runs
    of an instance moved less than the convergence threshold (SS.converged) or its SCP converged / ran out of iterations.
    Host logic only; every number comes from the kernels behind the C ABI."""
    bp = engine.bp
    B = bp.B
    thr = bp.model.scp_params[M.SP_CONVTHR]
    # solve_method!(SCPS, SCPP, solver, 1): the batched loop below is solve_gusto_batch unrolled one iteration at a time
    if X0 is None:
        X0, U0 = bp.init_traj_straightline()
    engine.set_shooting_trajectory(X0, U0)                               # TOS.SS = ShootingSolution(SP, deepcopy(traj_init))
    st = _ScpStepper(engine, X0, U0)
    st.step(np.ones(B, bool))
    S = st.S
    SS = BatchShootingSolution(S, S.X, S.U, np.zeros(B, bool))
    SS.J_true.append(S.J_true[1].copy())
    conv_hist = [np.full(B, np.nan)]                                     # SS.convergence_measure starts as [NaN]
    x_goal = 0.5 * (bp.goal_lo + bp.goal_hi)                             # types.jl:219-224
    while True:
        run = ~S.converged & (S.iterations < max_iter) & ~SS.converged & ~st.dead
        if not run.any():
            break
        out = engine.shoot(None, x_goal, nsub, shoot_max_iter, ftol)     # SP = ShootingProblem(TOP, SCPS); solve!(SS, SP)
        ok = run & (out[:, SH_STATUS] == 0)
        SS.prob_status.append(np.where(run, out[:, SH_STATUS], -1).astype(np.int32))
        SS.J_true.append(np.where(ok, out[:, SH_JTRUE], np.nan))
        conv_hist.append(np.where(run, np.where(ok, out[:, SH_CONV], np.nan), conv_hist[-1]))
        SS.newton_iters.append(np.where(run, out[:, SH_ITERS], 0))
        SS.attempts += 1
        with np.errstate(invalid="ignore"):
            two = conv_hist[-1] + conv_hist[-2]
            done = run & (S.iterations > 2) & (two <= thr)               # traj_opt.jl:32-38 (NaN compares false)
        SS.converged |= done
        if verbose:
            print(f"[shooting] attempt {SS.attempts}: optimal {int(ok.sum())}/{int(run.sum())} converged {int(SS.converged.sum())}")
        cont = run & ~done
        if not cont.any():
            break
        st.step(cont)                                                    # solve_method!(SCPS, SCPP, solver, 1)
    SS.convergence_measure = conv_hist
    Xs, Us, _ = engine.get_shooting_trajectory() if SS.attempts else (None, None, None)
    S.X, S.U = engine.get_trajectory()
    SS.X, SS.U = S.X.copy(), S.U.copy()
    if SS.attempts:
        SS.X[SS.converged] = Xs[SS.converged]; SS.U[SS.converged] = Us[SS.converged]
    return SS


class _ScpStepper:
    """solve_gusto_batch one outer iteration at a time (the reference calls solve_method! with max_iter = 1 repeatedly and
    keeps Delta / omega / histories in SCPS between the calls -- here they live in this object)."""

    def __init__(self, engine, X0=None, U0=None):
        bp = engine.bp
        self.engine, self.B, self.sp = engine, bp.B, bp.model.scp_params
        if X0 is None:
            X0, U0 = bp.init_traj_straightline()
        B, sp = self.B, self.sp
        engine.set_trajectory(X0, U0)
        self.Delta = np.full(B, sp[M.SP_DELTA0]); self.omega = np.full(B, sp[M.SP_OMEGA0])
        engine.set_penalties(self.omega, self.Delta)
        engine.set_candidate(X0, U0)
        engine.linearize()
        ev0 = engine.evaluate()
        S = BatchSCPSolution(X0, U0, np.zeros(B, bool), np.zeros(B, bool), np.zeros(B, np.int64))
        S.J_true.append(ev0[:, EV_JTRUE].copy()); S.convergence_measure.append(np.zeros(B))
        self.S = S
        self.dead = np.zeros(B, bool)          # solver failure or omega > omega_max: the reference's loop would spin; we stop

    def step(self, active):
        e, S = self.engine, self.S
        e.set_active(active)
        out, info = e.iterate()
        st = gusto_update(out, solver_status_ok(info[:, 0]), active, self.Delta, self.omega, S.iterations, S.convergence_measure[-1], self.sp)
        e.accept(st["accept"], st["omega"], st["Delta"])
        S.J_true.append(np.where(st["accept"], out[:, EV_JTRUE], S.J_true[-1]))
        S.convergence_measure.append(np.where(st["run"], out[:, EV_CONV], S.convergence_measure[-1]))
        S.scp_status.append(st["status"]); S.accept_solution.append(st["accept"])
        self.Delta, self.omega = st["Delta"], st["omega"]
        S.iterations = st["iterations"]
        S.converged |= st["converged_now"]; S.successful |= st["successful_now"]
        self.dead |= st["done"] & ~st["converged_now"]
        S.batch_iterations += 1
        S.X, S.U = e.get_trajectory()
        return st


# ------------------------------------------------------------------------------------------------ TrajOpt variant
@dataclass
class BatchTrajOptSolution:
    """Per-instance SCPSolution histories + SCPParam_TrajOpt vectors (scp_trajopt.jl:3-22) of a batched solve_trajopt_jump!."""
    X: np.ndarray
    U: np.ndarray
    converged: np.ndarray
    iterations: np.ndarray
    J_true: list = field(default_factory=list)            # per instance: list of floats
    J_full: list = field(default_factory=list)
    solver_status: list = field(default_factory=list)
    convergence_measure: list = field(default_factory=list)
    rho_vec: list = field(default_factory=list)
    mu_vec: list = field(default_factory=list)
    s_vec: list = field(default_factory=list)
    xtol_vec: list = field(default_factory=list)
    ftol_vec: list = field(default_factory=list)
    ctol_vec: list = field(default_factory=list)
    newton_iters: list = field(default_factory=list)
    batch_solves: int = 0
    total_time: float = 0.0


def _trajopt_instance(prm, H):
    """The three nested loops of solve_trajopt_jump! (scp_trajopt.jl:71-155) for ONE instance, as a coroutine: it yields the device
    work it needs next -- ("mark", slot) = copy!(old_*_traj, SCPS.traj) (:73, :76), ("solve", mu, s) = one convex subproblem with
    the step installed (:83-133), ("compare", slot) = evaluate_ftol / evaluate_xtol / evaluate_ctol against a marked trajectory
    (:140-141, :148) -- and is resumed with the result.  The batch driver runs B of them in lockstep, one kernel launch per
    request kind.  Repairs of the reference routine (it cannot run as written) are listed in oracle/gusto_oracle/trajopt.py and
    DESIGN.md: real copies for old_penalty_traj / old_convex_traj after the trust loop, xtol_vec[end] in :142."""
    mu0, s0, c, tp, tm, kfac, ftol, xtol, ctol = prm[:9]
    max_pen, max_cvx, max_tr = int(prm[9]), int(prm[10]), int(prm[11])
    H["mu_vec"].append(mu0); H["s_vec"].append(s0)
    cs = xs = False
    for _pen in range(max_pen):
        if cs:
            break
        yield ("mark", 0)
        for _cvx in range(max_cvx):
            yield ("mark", 1)
            if cs:
                break
            if xs:
                xs = False
                break
            for _tr in range(max_tr):
                ok, ev, info = yield ("solve", H["mu_vec"][-1], H["s_vec"][-1])
                H["solver_status"].append(int(info[0])); H["newton_iters"].append(int(info[1]))
                if not ok:                                            # no iterate to continue from (the reference only warns, :108-111)
                    return
                H["xtol_vec"].append(ev[TO_XTOL]); H["convergence_measure"].append(ev[TO_XTOL]); H["J_full"].append(info[4])
                H["rho_vec"].append(ev[TO_RHO])
                H["s_vec"].append((tp if ev[TO_RHO] > c else tm) * H["s_vec"][-1])          # :122-126
                H["J_true"].append(ev[TO_JTRUE]); H["iterations"] += 1
                if H["s_vec"][-1] < xtol:                             # :134
                    xs = True
                    break
            cmp = yield ("compare", 1)
            H["ftol_vec"].append(abs(cmp[3] - cmp[4]) / abs(cmp[3])); H["xtol_vec"].append(cmp[2])
            if H["ftol_vec"][-1] < ftol or H["xtol_vec"][-1] < xtol:  # :142
                cs = True
                break
        cmp = yield ("compare", 0)
        H["ctol_vec"].append(cmp[0] / cmp[1])
        if H["ctol_vec"][-1] < ctol:                                  # :148
            cs = True
            H["converged"] = True
            break
        H["mu_vec"].append(H["mu_vec"][-1] * kfac)                    # :154


def solve_trajopt_batch(engine: Engine, X0=None, U0=None, params=None, verbose=False):
    """Batched solve_trajopt_jump! (scp_trajopt.jl:33-157): the loops above per instance, every subproblem solve / evaluation of
    the whole batch in one kernel launch."""
    bp = engine.bp
    B = bp.B
    prm = M.TRAJOPT_PARAMS[bp.model.model_id] if params is None else params
    if X0 is None:
        X0, U0 = bp.init_traj_straightline()
    t0 = time.perf_counter()
    engine.trajopt_enable()
    engine.set_trajectory(X0, U0)
    engine.set_candidate(X0, U0)
    mu = np.full(B, prm[0]); s = np.full(B, prm[1])
    engine.set_penalties(mu, s)
    engine.linearize()
    J0 = engine.evaluate()[:, EV_JTRUE]                               # :64
    hist = [dict(mu_vec=[], s_vec=[], solver_status=[-1], newton_iters=[], xtol_vec=[0.0], convergence_measure=[0.0], J_full=[], rho_vec=[0.0],
                 J_true=[float(J0[b])], ftol_vec=[0.0], ctol_vec=[0.0], iterations=0, converged=False) for b in range(B)]
    gens = [_trajopt_instance(prm, hist[b]) for b in range(B)]
    req = [None] * B
    for b in range(B):
        req[b] = next(gens[b])
    nsolve = 0

    def advance(b, value):
        try:
            req[b] = gens[b].send(value)
        except StopIteration:
            req[b] = None

    while any(r is not None for r in req):
        # bookkeeping requests first, one device call per kind, until every live instance waits for a solve
        for _ in range(16):
            pending = [b for b in range(B) if req[b] is not None and req[b][0] != "solve"]
            if not pending:
                break
            for slot in (0, 1):
                which = np.zeros(B, np.uint8)
                for b in pending:
                    if req[b] == ("mark", slot):
                        which[b] = 1
                if which.any():
                    engine.trajopt_mark(slot, which)
                    for b in np.nonzero(which)[0]:
                        advance(b, None)
            for slot in (1, 0):
                idx = [b for b in range(B) if req[b] == ("compare", slot)]
                if idx:
                    cmpv = engine.trajopt_compare(slot)
                    for b in idx:
                        advance(b, cmpv[b])
        live = np.array([r is not None for r in req])
        if not live.any():
            break
        for b in range(B):
            if live[b]:
                mu[b], s[b] = req[b][1], req[b][2]
        ev, info = engine.trajopt_iterate(mu, s, live.astype(np.uint8))
        ok = solver_status_ok(info[:, 0])
        engine.accept((live & ok).astype(np.uint8))                   # copy!(SCPS.traj, new_traj), :128
        nsolve += 1
        if verbose:
            print(f"[trajopt] solve {nsolve:2d} live {int(live.sum()):5d} newton {info[live, 1].mean():.1f} mu {mu[live].max():g} s in [{s[live].min():g}, {s[live].max():g}]")
        for b in range(B):
            if live[b]:
                advance(b, (bool(ok[b]), ev[b].copy(), info[b].copy()))
    X, U = engine.get_trajectory()
    S = BatchTrajOptSolution(X, U, np.array([h["converged"] for h in hist]), np.array([h["iterations"] for h in hist]))
    for key in ("J_true", "J_full", "solver_status", "convergence_measure", "rho_vec", "mu_vec", "s_vec", "xtol_vec", "ftol_vec", "ctol_vec", "newton_iters"):
        setattr(S, key, [h[key] for h in hist])
    S.batch_solves = nsolve
    S.total_time = time.perf_counter() - t0
    return S
