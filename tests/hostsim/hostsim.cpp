// TEST INFRASTRUCTURE ONLY -- single-thread host simulation of the CUDA kernel bodies.
//
// Compiles gusto.jl_b200/csrc/{linearize,ipm,evaluate}.cuh with -DGUSTO_HOSTSIM (G_TID = 0, G_NTHR = 1, barriers are
// no-ops) so that the kernels' arithmetic can be checked against the oracle in the GPU-less build container
// (pytest -m "not gpu").  It is built into tests/hostsim/_build/, is never linked into libgusto_b200.so and is never
// reachable from the product path; it proves nothing about races or memory spaces -- the -m gpu tests do that.
#define GUSTO_HOSTSIM 1
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/gusto_b200.h"
#include "../../gusto.jl_b200/csrc/common.cuh"
#include "../../gusto.jl_b200/csrc/linearize.cuh"
#include "../../gusto.jl_b200/csrc/ipm.cuh"
#undef GUSTO_IPM_ALG
#define GUSTO_IPM_ALG 1                      // second inclusion: the TrajOpt subproblem (namespace gusto::ipm_trajopt)
#include "../../gusto.jl_b200/csrc/ipm.cuh"
#include "../../gusto.jl_b200/csrc/evaluate.cuh"
#include "../../gusto.jl_b200/csrc/trajopt.cuh"
#include "../../gusto.jl_b200/csrc/postprocess.cuh"
#include "../../gusto.jl_b200/csrc/shooting.cuh"
#include "../../gusto.jl_b200/csrc/scp.cuh"
#include "../../gusto.jl_b200/csrc/blocks.cuh"

using namespace gusto;

// optional [B][n_x] buffer that receives SCPS.dual (the init-row multipliers) of the next hostsim_iterate call
static double* g_dual = nullptr;
extern "C" void hostsim_set_dual_buffer(double* dual) { g_dual = dual; }
// 0 = GuSTO subproblem, 1 = TrajOpt subproblem (omega carries mu, delta carries s) for the next hostsim_iterate calls
static int g_alg = 0;
extern "C" void hostsim_set_algorithm(int alg) { g_alg = alg; }

template <int M>
static void run(const BatchDesc& d, BatchPtrs& p, const IpmParams& prm, int stages, double* info, double* eval) {
  using T = Traits<M>;
  using L = IpmLayout<M>;
  const int N = d.N, B = d.B;
  if (stages & 1) {
    std::vector<double> ws(T::NX * T::NX + T::NX);
    double bv[T::NU > 0 ? T::NU : 1];
    dyn_B_columns<M>(d.rp, bv);
    for (int b = 0; b < B; ++b)
      for (int k = 0; k < N; ++k)
        linearize_knot_global<M>(d, p, b, k, p.Xp + ((size_t)b * N + k) * T::NX, p.Up + ((size_t)b * N + k) * T::NU, ws.data(), bv);
  }
  if ((stages & 2) && g_alg == 1) {
    if constexpr (M == FREEFLYER_SE2 || M == ASTROBEE_SE3) {
      using LT = ipm_trajopt::IpmLayout<M>;
      std::vector<double> scratch(LT::scratch_doubles(N, d.n_obs)), smem(LT::smem_doubles(N, 1));
      for (int b = 0; b < B; ++b) ipm_trajopt::ipm_solve_instance<M>(d, p, prm, b, scratch.data(), smem.data(), info + (size_t)b * IPM_NINFO);
    }
  } else if (stages & 2) {
    std::vector<double> scratch(L::scratch_doubles(N, d.n_obs)), smem(L::smem_doubles(N, 1));
    for (int b = 0; b < B; ++b) ipm_solve_instance<M>(d, p, prm, b, scratch.data(), smem.data(), info + (size_t)b * IPM_NINFO);
  }
  if (stages & 4) {
    double red[4];
    for (int b = 0; b < B; ++b) {
      if (g_alg == 1) trajopt_evaluate_instance<M>(d, p, b, p.Xn + (size_t)b * N * T::NX, p.Un + (size_t)b * N * T::NU, eval + (size_t)b * TRAJOPT_NOUT, red);
      else evaluate_instance<M>(d, p, b, p.Xn + (size_t)b * N * T::NX, p.Un + (size_t)b * N * T::NU, eval + (size_t)b * EVAL_NOUT, red);
    }
  }
}

extern "C" int hostsim_iterate(const gusto_config* cfg, const int32_t* obs_kind, const double* obs_a, const double* obs_b,
                               const double* x_init, const double* goal_lo, const double* goal_hi, const double* tf,
                               double* Xp, double* Up, double* Xn, double* Un, double* omega, double* delta,
                               double* f, double* A, double* g, double* rows, int stages, double* info, double* eval) {
  BatchDesc d;
  memset(&d, 0, sizeof(d));
  d.model_id = cfg->model_id; d.N = cfg->N; d.B = cfg->B; d.n_obs = cfg->model_id == DUBINS ? 0 : cfg->n_obs;
  for (int i = 0; i < 16; ++i) d.rp[i] = cfg->robot_params[i];
  for (int i = 0; i < 10; ++i) d.sp[i] = cfg->scp_params[i];
  for (int i = 0; i < MAX_NX; ++i) d.goal_type[i] = cfg->goal_type[i];
  for (int i = 0; i < d.n_obs; ++i) {
    d.obs_kind[i] = obs_kind[i];
    for (int a = 0; a < 3; ++a) { d.obs_a[i][a] = obs_a[i * 3 + a]; d.obs_b[i][a] = obs_b[i * 3 + a]; }
  }
  BatchPtrs p;
  p.active = nullptr;
  p.dual = g_dual;
  p.tf = tf; p.x_init = x_init; p.goal_lo = goal_lo; p.goal_hi = goal_hi;
  p.Xp = Xp; p.Up = Up; p.Xn = Xn; p.Un = Un; p.omega = omega; p.delta = delta;
  // the kernels keep the blocks knot-minor (A on its sparsity pattern); callers of this harness pass / receive the dense arrays
  const size_t np = g_np(d.N), no1 = d.n_obs > 0 ? d.n_obs : 1;
  int nx = 0, anz = 0;
  switch (cfg->model_id) {
    case DUBINS: nx = Traits<DUBINS>::NX; anz = Traits<DUBINS>::ANZ; break;
    case FREEFLYER_SE2: nx = Traits<FREEFLYER_SE2>::NX; anz = Traits<FREEFLYER_SE2>::ANZ; break;
    case ASTROBEE_SE3: nx = Traits<ASTROBEE_SE3>::NX; anz = Traits<ASTROBEE_SE3>::ANZ; break;
    case ASTROBEE_SE3_MANIFOLD: nx = Traits<ASTROBEE_SE3_MANIFOLD>::NX; anz = Traits<ASTROBEE_SE3_MANIFOLD>::ANZ; break;
    default: return -1;
  }
  std::vector<double> fc((size_t)d.B * nx * np), gc(fc.size()), Ac((size_t)d.B * anz * np), rc((size_t)d.B * 5 * no1 * np);
  p.f = fc.data(); p.A = Ac.data(); p.g = gc.data(); p.rows = rc.data();
  IpmParams prm;
  prm.max_iter = cfg->ipm_max_iter > 0 ? cfg->ipm_max_iter : 60;
  prm.tol = cfg->ipm_tol > 0 ? cfg->ipm_tol : 1e-8;
  prm.wn_base = 1e8; prm.wn_omega = 1e4;
  ipm_default_start(prm);
  if (const char* ev = getenv("GUSTO_IPM_MU0")) {      // developer override: "A[,B[,cap[,rp[,lo[,smin]]]]]" (rp = 0: no scaling with the start point's infeasibility), "0" = the tuned default start everywhere
    double a = 0, b2 = prm.mu0_b, cp = prm.mu0_cap, rp = prm.mu0_rp, lo = prm.mu0_lo, sm = prm.mu0_smin;
    const int n = sscanf(ev, "%lf,%lf,%lf,%lf,%lf,%lf", &a, &b2, &cp, &rp, &lo, &sm);
    if (n >= 1) { prm.mu0_a = a; prm.mu0_b = b2; prm.mu0_cap = cp; prm.mu0_rp = rp; prm.mu0_lo = lo; prm.mu0_smin = sm; }
  }
#define HS_CASE(MM)                                                                                                         \
  case MM:                                                                                                                  \
    if (!(stages & 1)) blocks_pack<MM>(d.B, d.N, d.n_obs, f, A, g, rows, fc.data(), Ac.data(), gc.data(), rc.data());          \
    run<MM>(d, p, prm, stages, info, eval);                                                                                 \
    blocks_unpack<MM>(d.B, d.N, d.n_obs, fc.data(), Ac.data(), gc.data(), d.n_obs > 0 ? rc.data() : nullptr, f, A, g, rows); \
    break;
  switch (cfg->model_id) {
    HS_CASE(DUBINS)
    HS_CASE(FREEFLYER_SE2)
    HS_CASE(ASTROBEE_SE3)
    HS_CASE(ASTROBEE_SE3_MANIFOLD)
    default: return -1;
  }
#undef HS_CASE
  return 0;
}

// TrajOpt: (X, U) against a reference trajectory (trajopt_ctol_instance), out[B][2]
template <int M>
static void run_ctol(const BatchDesc& d, BatchPtrs& p, const double* X, const double* U, const double* Xr, const double* Ur, double* out) {
  using T = Traits<M>;
  double red[4];
  const size_t sx = (size_t)d.N * T::NX, su = (size_t)d.N * T::NU;
  for (int b = 0; b < d.B; ++b) trajopt_ctol_instance<M>(d, p, b, X + b * sx, U + b * su, Xr + b * sx, Ur + b * su, out + (size_t)b * 2, red);
}
extern "C" int hostsim_trajopt_ctol(const gusto_config* cfg, const int32_t* obs_kind, const double* obs_a, const double* obs_b,
                                    const double* goal_lo, const double* goal_hi, const double* tf, const double* X, const double* U,
                                    const double* Xr, const double* Ur, double* out) {
  BatchDesc d;
  memset(&d, 0, sizeof(d));
  d.model_id = cfg->model_id; d.N = cfg->N; d.B = cfg->B; d.n_obs = cfg->model_id == DUBINS ? 0 : cfg->n_obs;
  for (int i = 0; i < 16; ++i) d.rp[i] = cfg->robot_params[i];
  for (int i = 0; i < 10; ++i) d.sp[i] = cfg->scp_params[i];
  for (int i = 0; i < MAX_NX; ++i) d.goal_type[i] = cfg->goal_type[i];
  for (int i = 0; i < d.n_obs; ++i) {
    d.obs_kind[i] = obs_kind[i];
    for (int a = 0; a < 3; ++a) { d.obs_a[i][a] = obs_a[i * 3 + a]; d.obs_b[i][a] = obs_b[i * 3 + a]; }
  }
  BatchPtrs p;
  memset(&p, 0, sizeof(p));
  p.tf = tf; p.goal_lo = goal_lo; p.goal_hi = goal_hi;
  switch (cfg->model_id) {
    case FREEFLYER_SE2: run_ctol<FREEFLYER_SE2>(d, p, X, U, Xr, Ur, out); break;
    case ASTROBEE_SE3: run_ctol<ASTROBEE_SE3>(d, p, X, U, Xr, Ur, out); break;
    default: return -1;
  }
  return 0;
}

// post-processing bodies (check_instance / interpolate_interval) on the trajectory (X, U)
template <int M>
static void run_post(const BatchDesc& d, BatchPtrs& p, const double* X, const double* U, int nstep, double* chk, double* Xfull, double* Ufull) {
  using T = Traits<M>;
  const int N = d.N;
  double red[4];
  for (int b = 0; b < d.B; ++b) {
    const double* Xb = X + (size_t)b * N * T::NX;
    const double* Ub = U + (size_t)b * N * T::NU;
    check_instance<M>(d, p, b, Xb, Ub, chk + (size_t)b * CHECK_NOUT, red);
    const size_t nf = (size_t)nstep * (N - 1);
    for (int k = 0; k < N - 1; ++k)
      interpolate_interval<M>(d, p, b, k, nstep, Xb, Ub, Xfull + (size_t)b * (nf + 1) * T::NX, Ufull + (size_t)b * nf * T::NU);
  }
}

extern "C" int hostsim_postprocess(const gusto_config* cfg, const int32_t* obs_kind, const double* obs_a, const double* obs_b,
                                   const double* tf, const double* X, const double* U, int nstep, double* chk, double* Xfull, double* Ufull) {
  BatchDesc d;
  memset(&d, 0, sizeof(d));
  d.model_id = cfg->model_id; d.N = cfg->N; d.B = cfg->B; d.n_obs = cfg->model_id == DUBINS ? 0 : cfg->n_obs;
  for (int i = 0; i < 16; ++i) d.rp[i] = cfg->robot_params[i];
  for (int i = 0; i < 10; ++i) d.sp[i] = cfg->scp_params[i];
  for (int i = 0; i < d.n_obs; ++i) {
    d.obs_kind[i] = obs_kind[i];
    for (int a = 0; a < 3; ++a) { d.obs_a[i][a] = obs_a[i * 3 + a]; d.obs_b[i][a] = obs_b[i * 3 + a]; }
  }
  BatchPtrs p;
  memset(&p, 0, sizeof(p));
  p.tf = tf;
  switch (cfg->model_id) {
    case DUBINS: run_post<DUBINS>(d, p, X, U, nstep, chk, Xfull, Ufull); break;
    case FREEFLYER_SE2: run_post<FREEFLYER_SE2>(d, p, X, U, nstep, chk, Xfull, Ufull); break;
    case ASTROBEE_SE3: run_post<ASTROBEE_SE3>(d, p, X, U, nstep, chk, Xfull, Ufull); break;
    case ASTROBEE_SE3_MANIFOLD: run_post<ASTROBEE_SE3_MANIFOLD>(d, p, X, U, nstep, chk, Xfull, Ufull); break;
    default: return -1;
  }
  return 0;
}

// shooting body (shoot_instance), lanes run one after the other
template <int M>
static void run_shoot(const BatchDesc& d, BatchPtrs& p, const double* p0, const double* x_goal, int nsub, int max_iter, double ftol,
                      double* Xs, double* Us, double* Ps, double* out) {
  using T = Traits<M>;
  std::vector<double> sm(ShootLayout<M>::TOTAL);
  const size_t N = d.N;
  for (int b = 0; b < d.B; ++b)
    shoot_instance<M>(d, p, b, p0 + (size_t)b * T::NX, x_goal + (size_t)b * T::NX, nsub, max_iter, ftol, sm.data(),
                      Xs + (size_t)b * N * T::NX, Us + (size_t)b * N * T::NU, Ps + (size_t)b * N * T::NX, out + (size_t)b * SHOOT_NOUT);
}

extern "C" int hostsim_shoot(const gusto_config* cfg, const double* x_init, const double* tf, const double* p0, const double* x_goal,
                             int nsub, int max_iter, double ftol, double* Xs, double* Us, double* Ps, double* out) {
  BatchDesc d;
  memset(&d, 0, sizeof(d));
  d.model_id = cfg->model_id; d.N = cfg->N; d.B = cfg->B; d.n_obs = 0;
  for (int i = 0; i < 16; ++i) d.rp[i] = cfg->robot_params[i];
  BatchPtrs p;
  memset(&p, 0, sizeof(p));
  p.tf = tf; p.x_init = x_init;
  switch (cfg->model_id) {
    case DUBINS: run_shoot<DUBINS>(d, p, p0, x_goal, nsub, max_iter, ftol, Xs, Us, Ps, out); break;
    case ASTROBEE_SE3_MANIFOLD: run_shoot<ASTROBEE_SE3_MANIFOLD>(d, p, p0, x_goal, nsub, max_iter, ftol, Xs, Us, Ps, out); break;
    default: return -1;
  }
  return 0;
}

// device-resident outer step (scp.cuh::scp_update_instance), one call per instance: the kernel's decision table on the host
extern "C" int hostsim_scp_update(int B, const double* ev, const double* info, const double* sp, int force, const uint8_t* active,
                                  double* delta, double* omega, int* iterations, double* conv_prev, double* j_true, double* j_full,
                                  uint8_t* converged, uint8_t* successful, double* rec, int* flags) {
  for (int b = 0; b < B; ++b)
    flags[b] = scp_update_instance(ev + (size_t)b * EVAL_NOUT, info + (size_t)b * IPM_NINFO, sp, force != 0, active[b] != 0, delta + b, omega + b,
                                   iterations + b, conv_prev + b, j_true + b, j_full + b, converged + b, successful + b, rec + (size_t)b * HIST_W);
  return 0;
}
