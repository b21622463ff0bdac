/* gusto_b200.h -- C ABI of the B200-native batched GuSTO SCP hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / CUDA types.  A Julia host binds it with
 * `ccall` (gusto.jl_b200/julia/GuSTOB200.jl), the tests and bench bind it with ctypes (gusto.jl_b200/host.py).
 *
 * It replaces, inside `solve_gusto_jump!` (reference: /root/reference/src/scp/scp_gusto.jl:49-176), everything
 * that is not the outer accept/reject logic:
 *     gusto_linearize          <- update_model_params! (:95) + the cc.func row evaluations of
 *                                 add_constraints_gusto_jump! (:192-251): f_dyn/A_dyn/B_dyn, dynamics_constraints,
 *                                 ncsi_*_convexified + BulletCollision.distance (dynamics/astrobee_se3.jl:130-305)
 *     gusto_solve_subproblem   <- Model(with_optimizer(...)) (:82-92), add_variables_jump! (:178-190),
 *                                 add_objective_gusto_jump! (:253-314), JuMP.optimize! (:104), JuMP.value (:114)
 *     gusto_evaluate           <- convergence_metric (:115, traj_opt.jl:74-85), trust_region_satisfied_gusto (:120),
 *                                 convex_ineq_satisfied_gusto_jump (:121), trust_region_ratio_gusto (:124),
 *                                 cost_true (:146), JuMP.objective_value (:116)
 *     gusto_accept             <- copy!(SCPS.traj, new_traj) (:147) and the Delta/omega pushes (:125-145,156)
 * The trust-region update and convergence test themselves (:119-174) stay in the host language -- or, for a caller that wants a
 * whole solve without a host round trip per iteration, run in the library's own update kernel (gusto_scp_run below; same table).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (GUSTO_E_*); it never throws.  gusto_last_error() returns a
 *     message for the last failure on that context (or a global one when ctx is NULL).
 *   - all arrays are Float64, dense, instance-major / knot-major: X is [B][N][n_x], which is exactly the memory of a
 *     Julia Array{Float64,3} of size (n_x, N, B); U is [B][N][n_u].
 *   - the library never keeps a caller pointer after the call returns (Julia arrays may move after GC.@preserve).
 *   - a context owns one CUDA stream; calls on one context are serialised and block until host outputs are written.
 */
#ifndef GUSTO_B200_H
#define GUSTO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gusto_ctx gusto_ctx;

/* model_id: which DynamicsModel plugin (src/dynamics/*.jl) the batch uses */
enum { GUSTO_DUBINS = 0, GUSTO_FREEFLYER_SE2 = 1, GUSTO_ASTROBEE_SE3 = 2, GUSTO_ASTROBEE_SE3_MANIFOLD = 3 };
/* obstacle kinds: HyperRectangle (a = min corner, b = max corner) / HyperSphere (a = centre, b[0] = radius) */
enum { GUSTO_OBS_BOX = 0, GUSTO_OBS_SPHERE = 1 };
/* goal_type per state coordinate: none / PointGoal (equality) / BoxGoal (inequality)  (src/goals.jl) */
enum { GUSTO_GOAL_FREE = 0, GUSTO_GOAL_POINT = 1, GUSTO_GOAL_BOX = 2 };
/* per-instance convex-solver status (MOI-like): OPTIMAL / ITERATION_LIMIT / NUMERICAL_ERROR / ALMOST_OPTIMAL (the solve
 * stalled or broke down within 1e3 * tol of the tolerance and below 0.1 * eps of the SCP's own soft-row threshold and answers
 * with its best iterate: the MOI.ALMOST_LOCALLY_SOLVED the reference accepts next to OPTIMAL, scp_gusto.jl:107).  For a
 * NUMERICAL status info[3] = -1 marks a factorisation breakdown (otherwise a NaN / Inf appeared). */
enum { GUSTO_SOLVER_OPTIMAL = 0, GUSTO_SOLVER_ITERATION_LIMIT = 1, GUSTO_SOLVER_NUMERICAL = 2, GUSTO_SOLVER_ALMOST_OPTIMAL = 3 };
enum {
  GUSTO_OK = 0, GUSTO_E_ARG = -1, GUSTO_E_CUDA = -2, GUSTO_E_ALLOC = -3, GUSTO_E_STATE = -4, GUSTO_E_NODEVICE = -5
};

/* number of doubles per instance written by gusto_evaluate / gusto_solve_subproblem */
#define GUSTO_EVAL_NOUT 8 /* conv, tr_ok, ineq_ok, rho, J_true, J_full, max_k|dX_k|^2, max soft-row value */
#define GUSTO_SOLVE_NINFO 8 /* status, newton iterations, residual, mu, objective, SM cycles: assembly, factorisation, KKT solves */

typedef struct {
  int32_t model_id;        /* GUSTO_* model                                                        */
  int32_t N;               /* knots            (SCPProblem.N, types.jl:78-86)                      */
  int32_t B;               /* instances in this context (this rank's shard)                        */
  int32_t n_obs;           /* collision components: keepout_zones..., obstacle_set... (types.jl:19) */
  double robot_params[16]; /* 0 mass | 1-3 Jxx,Jyy,Jzz | 4 radius | 5 v_max | 6 a_max | 7 w_max | 8 alpha_max |
                              9 clearance | 10 dubins v | 11 dubins k | 12-14 dubins x_max | 15 dubins u_max
                              (robot/astrobee3D.jl:16-30, robot/freeflyer.jl:29-50, dynamics/dubins_car.jl:20-33) */
  double scp_params[10];   /* Delta0, omega0, omega_max, eps, rho0, rho1, beta_succ, beta_fail, gamma_fail,
                              convergence_threshold  (SCPParam_GuSTO(model), SCPParam(model))      */
  int32_t goal_type[16];   /* per state coordinate, shared by the batch                                          */
  int32_t device;          /* CUDA device ordinal                                                                */
  /* convex-solver controls (0 selects the default) */
  int32_t ipm_max_iter;    /* default 60   */
  int32_t ipm_nref;        /* accepted and ignored since round 2 (the Riccati solve needs no iterative refinement) */
  double ipm_tol;          /* default 1e-8 */
  double ipm_delta_p;      /* accepted and ignored since round 2 (no primal regularisation) */
  double ipm_delta_d;      /* accepted and ignored since round 2 (PointGoal rows: penalty 1e8 + 1e4 omega, csrc/ipm.cuh) */
} gusto_config;

/* Create a context for a batch of B instances sharing robot / model / environment.
 * obs_kind[n_obs], obs_a[n_obs*3], obs_b[n_obs*3] describe the collision components (may be NULL when n_obs == 0). */
int32_t gusto_create(const gusto_config* cfg, const int32_t* obs_kind, const double* obs_a, const double* obs_b,
                     gusto_ctx** out);
int32_t gusto_destroy(gusto_ctx* ctx);
const char* gusto_last_error(const gusto_ctx* ctx);
int32_t gusto_version(void);

/* Per-instance problem data: x_init[B*n_x], goal_lo/goal_hi[B*n_x] (PointGoal: lo == hi == point), tf[B]
 * (ProblemDefinition.x_init / goal_set, TrajectoryOptimizationProblem.tf_guess; types.jl:32-63). */
int32_t gusto_set_problems(gusto_ctx* ctx, const double* x_init, const double* goal_lo, const double* goal_hi,
                           const double* tf);
/* Accepted ("previous") trajectory SCPS.traj: X[B*N*n_x], U[B*N*n_u]. */
int32_t gusto_set_trajectory(gusto_ctx* ctx, const double* X, const double* U);
int32_t gusto_get_trajectory(gusto_ctx* ctx, double* X, double* U);
/* Candidate trajectory of the last gusto_solve_subproblem (new_traj, scp_gusto.jl:114). */
int32_t gusto_get_candidate(gusto_ctx* ctx, double* X, double* U);
int32_t gusto_set_candidate(gusto_ctx* ctx, const double* X, const double* U);
/* Current penalty weight and trust-region size per instance: omega[B], delta[B] (param.alg.omega_vec[end], Delta_vec[end]). */
int32_t gusto_set_penalties(gusto_ctx* ctx, const double* omega, const double* delta);

/* K1+K2: linearize dynamics and obstacle rows about the accepted trajectory, for all B*N knots. */
int32_t gusto_linearize(gusto_ctx* ctx);
/* Test hook: copy the blocks out.  Any pointer may be NULL.  f[B*N*n_x], A[B*N*n_x*n_x] (row-major per knot),
 * g[B*N*n_x] (= f - A Xp - B Up), rows[B*N*n_obs*5] (nhat_x, nhat_y, nhat_z, off, dist0). */
int32_t gusto_get_blocks(gusto_ctx* ctx, double* f, double* A, double* g, double* rows);

/* K3: solve the convex subproblem of every instance; info[B*GUSTO_SOLVE_NINFO] (may be NULL). */
int32_t gusto_solve_subproblem(gusto_ctx* ctx, double* info);
/* K4: evaluate the candidate against the accepted trajectory; out[B*GUSTO_EVAL_NOUT]. */
int32_t gusto_evaluate(gusto_ctx* ctx, double* out);
/* accept[b] != 0: candidate becomes the accepted trajectory of instance b.  omega/delta (may be NULL) are the values
 * for the NEXT iteration (the obstacle toggle distance Delta/8 + clearance follows, scp_gusto.jl:156). */
int32_t gusto_accept(gusto_ctx* ctx, const uint8_t* accept, const double* omega, const double* delta);

/* active[b] == 0 freezes instance b: solve / evaluate skip it (converged or failed instances, scp_gusto.jl:172-173).
 * All instances are active after gusto_create. */
int32_t gusto_set_active(gusto_ctx* ctx, const uint8_t* active);

/* One fused outer iteration for callers that keep everything on the device between iterations:
 * linearize -> solve -> evaluate, a single D2H copy of out[B*GUSTO_EVAL_NOUT] and info[B*GUSTO_SOLVE_NINFO]. */
int32_t gusto_iterate(gusto_ctx* ctx, double* out, double* info);

/* The same iteration for a host-language loop that keeps the trajectories on the HOST (pinned buffers recommended): uploads
 * the accepted trajectory X, U (NULL, NULL: keep the device copy), the penalties omega / delta and the active flags (each may be
 * NULL), runs linearize -> solve -> evaluate, downloads out / info (info may be NULL) and the candidate Xn, Un (NULL, NULL: skip)
 * -- all enqueued on the context stream with one synchronisation at the end. */
int32_t gusto_iterate_host(gusto_ctx* ctx, const double* X, const double* U, const double* omega, const double* delta,
                           const uint8_t* active, double* out, double* info, double* Xn, double* Un);

/* Post-processing of the ACCEPTED trajectory (SCPS.traj), SURVEY.md section 8(f):
 * out[B*GUSTO_CHECK_NOUT] = { dynamics_constraint_satisfaction (dynamics/astrobee_se3.jl:529-540: sum_k |(X_{k+1}-X_k)/dt -
 * f(X_k,U_k)|_1), max |X_{k+1} - X_k - h/2 (f_k + f_{k+1})| (nonlinear trapezoid defect), verify_collision_free (:542-562)
 * as collision_free (0/1), first violating knot k (0-based, -1), its obstacle index (-1), its signed distance, the
 * minimum signed distance over all (knot, obstacle) pairs, max_k |scale .* U_k| / bound over the control balls }. */
#define GUSTO_CHECK_NOUT 8
int32_t gusto_check_trajectory(gusto_ctx* ctx, double* out);
/* interpolate_traj (dynamics/astrobee_se3.jl:495-527): RK4 upsampling of the accepted trajectory with nstep sub-steps per
 * knot interval (the reference uses nstep = ceil(dt/dt_min), dt_min = 0.1) under the held control U_k.
 * Xfull[B*(nstep*(N-1)+1)*n_x], Ufull[B*nstep*(N-1)*n_u]. */
int32_t gusto_interpolate_trajectory(gusto_ctx* ctx, int32_t nstep, double* Xfull, double* Ufull);

/* Indirect shooting refinement (reference shooting.jl:4-66, solve_SCPshooting! traj_opt.jl:4-45).
 * gusto_get_duals: dual[B*n_x] = SCPS.dual of the last gusto_solve_subproblem / gusto_iterate, i.e. minus the JuMP dual of
 * the init constraints X[:,1] = x_init (scp_gusto.jl:116, get_dual_jump dynamics/dubins_car.jl:254-257).
 * gusto_shoot: one shooting attempt per instance (solve!(SS, SP)): find the initial costate p0 with
 * x(tf; x_init, p0) = x_goal under shooting_ode! (dynamics/dubins_car.jl:259-280, astrobee_se3_manifold.jl:831-895; only
 * these two models -- GUSTO_E_ARG otherwise).  p0[B*n_x] start (NULL: the duals of the last solve, types.jl:225),
 * x_goal[B*n_x] (NULL: centres of the goal sets, types.jl:219-224), RK4 with nsub sub-steps per knot interval,
 * Levenberg-Marquardt with at most max_iter iterations (reference: 100) until |F|_inf <= ftol (reference: 1e-3).
 * out[B*GUSTO_SHOOT_NOUT] = { status (0 :Optimal, 1 :Diverged), LM iterations, |x_goal - x(tf)|_inf, J_true (cost_true of the
 * new trajectory), convergence_metric(new, SS.traj) (traj_opt.jl:74-85), final damping, 0, 0 }; J_true and the metric are NaN
 * for a diverged attempt (shooting.jl:16-22,41-47).  A converged attempt replaces the shooting trajectory SS.traj kept on
 * the device; gusto_get_shooting_trajectory downloads it (X[B*N*n_x], U[B*N*n_u] = get_control, costates P[B*N*n_x]; any
 * pointer may be NULL).  gusto_set_shooting_trajectory seeds SS.traj (the reference: ShootingSolution(SP, deepcopy(traj_init)),
 * traj_opt.jl:16); without it the first gusto_shoot seeds SS.traj with the context's accepted trajectory at that moment. */
#define GUSTO_SHOOT_NOUT 8
int32_t gusto_get_duals(gusto_ctx* ctx, double* dual);
int32_t gusto_shoot(gusto_ctx* ctx, const double* p0, const double* x_goal, int32_t nsub, int32_t max_iter, double ftol, double* out);
int32_t gusto_get_shooting_trajectory(gusto_ctx* ctx, double* X, double* U, double* P);
int32_t gusto_set_shooting_trajectory(gusto_ctx* ctx, const double* X, const double* U);

/* Timing of the last call of each kernel on this context, in milliseconds (CUDA events on the context's stream):
 * ms[0] linearize, ms[1] solve, ms[2] evaluate, ms[3] accept. */
int32_t gusto_last_kernel_ms(gusto_ctx* ctx, float* ms);
/* Device-clock stopwatch: two CUDA events recorded on the context's stream (start synchronises the stream first). */
int32_t gusto_timer_start(gusto_ctx* ctx);
int32_t gusto_timer_stop(gusto_ctx* ctx, float* ms);
/* Number of kernel launches issued by this context since creation. */
int64_t gusto_launch_count(const gusto_ctx* ctx);

/* Device pointers for zero-copy plumbing (torch.distributed all_gather of the evaluation/status block).
 * which: 0 eval_out[B*8], 1 solve_info[B*8], 2 Xp, 3 Up, 4 Xn, 5 Un, 6 omega, 7 delta, 8 active (B bytes). */
int32_t gusto_device_ptr(gusto_ctx* ctx, int32_t which, void** ptr, int64_t* n_doubles);
/* Device-only variants (no host copies): run on the context stream and return after the stream is synchronised. */
int32_t gusto_iterate_device(gusto_ctx* ctx);
int32_t gusto_accept_device(gusto_ctx* ctx, const uint8_t* accept_dev, const double* omega_dev, const double* delta_dev);
/* Handle of the context's CUDA stream (cudaStream_t as integer) so that a torch stream can wait on it. */
int64_t gusto_stream_handle(gusto_ctx* ctx);

/* ---- Device-resident outer loop: a whole solve_gusto_jump! (scp_gusto.jl:49-176) without a host round trip per iteration.
 * The accept/reject test, the Delta/omega schedule and the convergence test (:119-174) run in one small kernel per iteration
 * (the same decision table as julia/GuSTOB200.jl and host.py::gusto_update, which stay available for a host that wants the
 * loop in its own language); one iteration = K1 -> K3 -> K4 -> update(+copy!(SCPS.traj, new_traj), :147), replayed as a CUDA
 * graph.  The host only reads one counter per iteration (unfinished instances over all ranks), and it reads it while the NEXT
 * iteration already runs -- when the counter is 0 that extra iteration found every instance frozen and did nothing.
 *
 * gusto_scp_begin   traj <- X0,U0 (NULL,NULL: the X0,U0 of the previous begin, kept on the device), Delta0/omega0, every
 *                   instance live, J_true[1] = cost_true(traj) (:60-75).  force != 0: the reference's force flag (:55,173).
 * gusto_scp_run     up to max_iter more outer iterations; stops as soon as every instance of every rank is finished
 *                   (converged, omega > omega_max, or solver failure :107-111).  batch_iterations: iterations that ran with
 *                   live instances; n_unfinished: instances (all ranks) still live afterwards.
 * gusto_scp_get     iterations[B], converged[B], successful[B] (SCPS.iterations/converged/successful), and the histories:
 *                   hist[n_hist][B][GUSTO_HIST_W], slot 0 = initial entry, slot i = after outer iteration i, each record
 *                   { J_true, J_full, scp_status, solver_status, accept, convergence_measure, Delta, omega, rho, tr_ok, ineq_ok,
 *                   newton iterations } (types.jl:150-173, scp_gusto.jl:15-19); counters[(n_hist-1)][4] = per iteration
 *                   { instances that ran, usable solves, accepted, still live afterwards } of THIS rank.  Any pointer may be
 *                   NULL.  n_hist <= GUSTO_SCP_MAX_HIST + 1.  The trajectory itself: gusto_get_trajectory. */
#define GUSTO_HIST_W 12
#define GUSTO_SCP_MAX_HIST 64
/* scp_status codes of the history records (SCPS.scp_status, scp_gusto.jl:125-145) */
enum { GUSTO_SCP_NA = 0, GUSTO_SCP_OK, GUSTO_SCP_INACCURATE_MODEL, GUSTO_SCP_VIOLATES_CONSTRAINTS, GUSTO_SCP_TR_VIOLATED,
       GUSTO_SCP_SOLVER_FAILED, GUSTO_SCP_INACTIVE };
int32_t gusto_scp_begin(gusto_ctx* ctx, const double* X0, const double* U0, int32_t force);
int32_t gusto_scp_run(gusto_ctx* ctx, int32_t max_iter, int32_t* batch_iterations, int32_t* n_unfinished);
int32_t gusto_scp_get(gusto_ctx* ctx, int32_t* iterations, uint8_t* converged, uint8_t* successful, double* hist, int32_t n_hist,
                      int32_t* counters);

/* ---- Multi-GPU: one context per GPU / process, each owning a contiguous shard of the batch; the path's only exchange is one
 * all-gather of B status bytes per outer iteration so that every rank stops together.  NCCL is
 * loaded at run time (dlopen libnccl.so.2; GUSTO_NCCL_LIB overrides) -- a single-GPU host needs none.
 * gusto_comm_unique_id  rank 0 creates the 128-byte ncclUniqueId; the host language broadcasts it (MPI / a file / torch).
 * gusto_comm_init       collective: every rank calls it with the same id.  From then on gusto_scp_run gathers the status bytes
 *                       the update kernel wrote (straight from that buffer, on a side stream, overlapping the next K1).
 * gusto_allgather_status  the same collective for a host-language loop: done_local[B] (host) -> done_all[nranks*B] (host,
 *                       may be NULL), n_unfinished = number of zeros over all ranks.  Works without a communicator (1 rank). */
int32_t gusto_comm_unique_id(uint8_t* id128);
int32_t gusto_comm_init(gusto_ctx* ctx, int32_t rank, int32_t nranks, const uint8_t* id128);
int32_t gusto_allgather_status(gusto_ctx* ctx, const uint8_t* done_local, uint8_t* done_all, int32_t* n_unfinished);

/* ---- TrajOpt SCP variant: solve_trajopt_jump! (/root/reference/src/scp/scp_trajopt.jl:33-157) behind the
 * same solve_method! slot.  The convex subproblem (:159-279) is solved on the device by a second compilation of the solve kernel:
 * HARD state trust region |X_k - Xp_k|^2 <= s (:165-173), hard boundary conditions (:175-195), every other inequality -- the control
 * balls included -- as a mu-penalised hinge (:222-233), the dynamics rows l1-penalised with weight mu (:257-275; the literal code leaves
 * one of its two slack vectors unbounded below, the restated form is the file's own equality penalty :236-246), toggle distance
 * clearance + 1 (:65).  The three nested loops (:71-155) stay in the host language (julia/GuSTOB200.jl solve_trajopt_b200!, Python twin
 * host.solve_trajopt_batch); they need per iteration:
 * gusto_trajopt_enable   allocates the variant's solver scratch and reference-trajectory buffers (freeflyerSE2 and astrobeeSE3: the
 *                        in-scope models with a SCPParam_TrajOpt; GUSTO_E_ARG otherwise).
 * gusto_trajopt_iterate  mu[B], s[B], active[B] (each may be NULL: keep) -> linearize, solve, evaluate the candidate against the
 *                        accepted trajectory: out[B*GUSTO_TRAJOPT_NOUT] = { evaluate_xtol (convergence_metric, :281-283),
 *                        trust_region_ratio_trajopt (astrobee_se3.jl:419-459), cost_true(candidate), cost_true(accepted), max_k |dX_k|^2,
 *                        ratio numerator, ratio denominator, sum of |linearised dynamics rows| }, info as gusto_solve_subproblem.
 *                        The step is installed with gusto_accept (the reference accepts every step, :128).
 * gusto_trajopt_mark     copies the accepted trajectory into reference slot 0 (old_penalty_traj, :73) or 1 (old_convex_traj, :76)
 *                        for the instances flagged in which[B] (NULL: all).
 * gusto_trajopt_compare  accepted trajectory against reference slot: out[B*5] = { evaluate_ctol numerator, denominator (:288-312 as
 *                        restated in oracle/gusto_oracle/trajopt.py), evaluate_xtol, cost_true(accepted), cost_true(reference) }. */
#define GUSTO_TRAJOPT_NOUT 8
int32_t gusto_trajopt_enable(gusto_ctx* ctx);
int32_t gusto_trajopt_iterate(gusto_ctx* ctx, const double* mu, const double* s, const uint8_t* active, double* out, double* info);
int32_t gusto_trajopt_mark(gusto_ctx* ctx, int32_t slot, const uint8_t* which);
int32_t gusto_trajopt_compare(gusto_ctx* ctx, int32_t slot, double* out);

#ifdef __cplusplus
}
#endif
#endif /* GUSTO_B200_H */
