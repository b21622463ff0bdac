// libgusto_b200.so: CUDA kernels (sm_100a) + the C ABI declared in include/gusto_b200.h.
// Product path only: there is no CPU fallback; every entry point fails with GUSTO_E_NODEVICE / GUSTO_E_CUDA when
// no B200 is usable.
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/gusto_b200.h"
#include "common.cuh"
#include "linearize.cuh"
#include "ipm.cuh"
#undef GUSTO_IPM_ALG
#define GUSTO_IPM_ALG 1            // second inclusion: the TrajOpt subproblem, namespace gusto::ipm_trajopt (see ipm.cuh)
#include "ipm.cuh"
#include "evaluate.cuh"
#include "trajopt.cuh"
#include "postprocess.cuh"
#include "shooting.cuh"
#include "scp.cuh"
#include "blocks.cuh"
#include <dlfcn.h>

using namespace gusto;

// ------------------------------------------------------------------------------------------ TMA helpers
namespace {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk tensor-memory-accelerator copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
}  // namespace

// ------------------------------------------------------------------------------------------------ kernels
constexpr int LIN_KNOTS_PER_CTA = 8;      // one warp per knot
#ifndef GUSTO_IPM_THREADS
#define GUSTO_IPM_THREADS 64
#endif
#ifndef GUSTO_IPM_MINBLOCKS
#define GUSTO_IPM_MINBLOCKS 7
#endif
constexpr int IPM_THREADS = GUSTO_IPM_THREADS;
#ifndef GUSTO_IPM_MAX_PACK
#define GUSTO_IPM_MAX_PACK 7
#endif
constexpr int IPM_MAX_PACK = GUSTO_IPM_MAX_PACK;   // 7 groups x 64 threads x 144 registers fit the register file of an SM (8 x 128 spilled more)
static_assert(GUSTO_IPM_THREADS == GUSTO_IPM_GROUP, "ipm.cuh's group size and the launch configuration must agree");
constexpr int EVAL_THREADS = 128;
// K4 resident CTAs per SM: 7 x 148 = 1036 slots hold the headline batch (1024 CTAs) in ONE wave; with 4 (128 registers) the second
// wave was 73 % full
#ifndef GUSTO_EVAL_MINBLOCKS
#define GUSTO_EVAL_MINBLOCKS 7
#endif

// K1+K2.  Grid: ceil(B*N/8) CTAs of 8 warps.  The 8 knots' states and controls are contiguous in HBM
// ([B][N][NX] knot-major) and are staged into shared memory with two TMA bulk copies per CTA.
template <int M>
__global__ void __launch_bounds__(LIN_KNOTS_PER_CTA * 32) linearize_kernel(const BatchDesc* __restrict__ dp, BatchPtrs p) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU, KPC = LIN_KNOTS_PER_CTA;
  __shared__ alignas(16) double sx[KPC * NX];
  __shared__ alignas(16) double su[KPC * NU];
  __shared__ double ws[KPC][NX * NX + NX];
  __shared__ double outs[(T::ANZ + 2 * NX + 5 * (T::WS > 0 ? MAX_OBS : 0)) * KPC];     // the CTA's blocks, [field][knot of the CTA]
  __shared__ alignas(8) unsigned long long mbar;
  __shared__ double sbv[NU > 0 ? NU : 1];
  __shared__ unsigned char pat[T::ANZ];                    // a_row(e) * NX + a_col(e)
  __shared__ int kn_b[KPC], kn_k[KPC];                     // instance / knot of the CTA's knots (-1: none or frozen instance)
  const BatchDesc& d = *dp;
  if (threadIdx.x == 0) dyn_B_columns<M>(d.rp, sbv);     // visible after the barrier / mbarrier wait below
  const int total = d.B * d.N;
  const int k0 = blockIdx.x * KPC;
  const int nk = (total - k0) < KPC ? (total - k0) : KPC;
  const uint32_t bx = (uint32_t)(nk * NX * sizeof(double)), bu = (uint32_t)(nk * NU * sizeof(double));
  const double* gx = p.Xp + (size_t)k0 * NX;
  const double* gu = p.Up + (size_t)k0 * NU;
  const bool use_tma = (bx % 16u == 0u) && (bu % 16u == 0u);
  if (use_tma) {
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&mbar, bx + bu);
      tma_bulk_g2s(sx, gx, bx, &mbar);
      tma_bulk_g2s(su, gu, bu, &mbar);
    }
    mbar_wait(&mbar, 0);
  } else {   // ragged tail CTA whose byte count is not a multiple of 16
    for (int i = threadIdx.x; i < nk * NX; i += blockDim.x) sx[i] = gx[i];
    for (int i = threadIdx.x; i < nk * NU; i += blockDim.x) su[i] = gu[i];
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_obs = T::WS > 0 ? d.n_obs : 0, nfield = T::ANZ + 2 * NX + 5 * n_obs;
  if (threadIdx.x < T::ANZ) pat[threadIdx.x] = (unsigned char)(T::a_row(threadIdx.x) * NX + T::a_col(threadIdx.x));
  if (threadIdx.x >= 32 && threadIdx.x < 32 + KPC) {       // (the write-out below used to divide by N and test `active` per element)
    const int kk = threadIdx.x - 32, gk = k0 + kk;
    int b = -1, k = 0;
    if (kk < nk) { b = gk / d.N; k = gk - b * d.N; if (p.active && !p.active[b]) b = -1; }
    kn_b[kk] = b; kn_k[kk] = k;
  }
  for (int i = threadIdx.x; i < KPC * (NX * NX + NX); i += blockDim.x) (&ws[0][0])[i] = 0.0;
  __syncthreads();
  const bool live = warp < nk && !(p.active && !p.active[(k0 + warp) / d.N]);     // (frozen instance: its blocks are never read again)
  // phase 1: the serial part (f and the 24 closed-form entries of A) of ALL the CTA's knots on the lanes of warp 0 -- one
  // instruction stream for 8 knots instead of 8 streams with one active lane each; meanwhile every warp does the obstacle rows
  // of its own knot, which need the state only
  if (warp == 0 && lane < nk && !(p.active && !p.active[(k0 + lane) / d.N])) linearize_fA<M>(d, sx + lane * NX, su + lane * NU, ws[lane]);
  if (live) linearize_rows<M>(d, sx + warp * NX, outs + (T::ANZ + 2 * NX) * KPC + warp, KPC);
  __syncthreads();
  // phase 2: A on its pattern, f and g of the warp's knot
  if (live) linearize_emit<M>(sx + warp * NX, su + warp * NU, ws[warp], sbv, pat, outs + warp, outs + T::ANZ * KPC + warp,
                              outs + (T::ANZ + NX) * KPC + warp, KPC);
  __syncthreads();
  // write-out: every field row of the knot-minor layout receives the CTA's 8 consecutive knots as 64 contiguous bytes
  const size_t np = g_np(d.N);
  for (int idx = threadIdx.x; idx < nfield * KPC; idx += blockDim.x) {
    const int fld = idx / KPC, kk = idx - fld * KPC;
    const int b = kn_b[kk], k = kn_k[kk];
    if (b < 0) continue;
    const double v = outs[idx];
    if (fld < T::ANZ) p.A[((size_t)b * T::ANZ + fld) * np + k] = v;
    else if (fld < T::ANZ + NX) p.f[((size_t)b * NX + fld - T::ANZ) * np + k] = v;
    else if (fld < T::ANZ + 2 * NX) p.g[((size_t)b * NX + fld - T::ANZ - NX) * np + k] = v;
    else p.rows[((size_t)b * 5 * n_obs + fld - T::ANZ - 2 * NX) * np + k] = v;
  }
}

// K3.  One instance per GROUP of IPM_THREADS threads; a CTA packs blockDim.x / IPM_THREADS groups (ipm_pack(), chosen so that
// B instances fill the SMs in one wave), each with its own slice of the dynamic shared memory.
template <int M>
__global__ void __launch_bounds__(IPM_THREADS * IPM_MAX_PACK, 1) ipm_kernel(const BatchDesc* __restrict__ dp, BatchPtrs p, IpmParams prm,
                                                          double* scratch, size_t stride, double* info, int smem_doubles) {
  extern __shared__ __align__(16) double smem[];
  const int sub = threadIdx.x / IPM_THREADS;
  const int b = blockIdx.x * (blockDim.x / IPM_THREADS) + sub;
  if (b >= dp->B) return;                    // group-uniform exits: an exited group no longer counts in the CTA barrier
  if (p.active && !p.active[b]) return;      // converged / failed instance: frozen
  ipm_solve_instance<M>(*dp, p, prm, b, scratch + (size_t)b * stride, smem + (size_t)sub * smem_doubles, info + (size_t)b * IPM_NINFO);
}

// K3, TrajOpt subproblem (scp_trajopt.jl:159-279): same launch shape, the second compilation of ipm.cuh.
template <int M>
__global__ void __launch_bounds__(IPM_THREADS * IPM_MAX_PACK, 1) ipm_trajopt_kernel(const BatchDesc* __restrict__ dp, BatchPtrs p, IpmParams prm,
                                                                  double* scratch, size_t stride, double* info, int smem_doubles) {
  extern __shared__ __align__(16) double smem[];
  const int sub = threadIdx.x / IPM_THREADS;
  const int b = blockIdx.x * (blockDim.x / IPM_THREADS) + sub;
  if (b >= dp->B) return;
  if (p.active && !p.active[b]) return;
  ipm_trajopt::ipm_solve_instance<M>(*dp, p, prm, b, scratch + (size_t)b * stride, smem + (size_t)sub * smem_doubles, info + (size_t)b * IPM_NINFO);
}
// TrajOpt evaluation scalars (trajopt.cuh).  Grid: B CTAs.
template <int M>
__global__ void __launch_bounds__(EVAL_THREADS) trajopt_evaluate_kernel(const BatchDesc* __restrict__ dp, BatchPtrs p, double* out) {
  using T = Traits<M>;
  __shared__ double red[EVAL_THREADS];
  const int b = blockIdx.x;
  if (p.active && !p.active[b]) return;
  trajopt_evaluate_instance<M>(*dp, p, b, p.Xn + (size_t)b * dp->N * T::NX, p.Un + (size_t)b * dp->N * T::NU, out + (size_t)b * TRAJOPT_NOUT, red);
}
// accepted trajectory against a marked reference trajectory: evaluate_ctol numerator / denominator, evaluate_xtol, cost_true of both
template <int M>
__global__ void __launch_bounds__(EVAL_THREADS) trajopt_compare_kernel(const BatchDesc* __restrict__ dp, BatchPtrs p, const double* Xr, const double* Ur, double* out) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU;
  __shared__ double red[EVAL_THREADS];
  const int b = blockIdx.x, N = dp->N;
  const double* X = p.Xp + (size_t)b * N * NX;
  const double* U = p.Up + (size_t)b * N * NU;
  const double* Xb = Xr + (size_t)b * N * NX;
  const double* Ub = Ur + (size_t)b * N * NU;
  double* o = out + (size_t)b * 5;
  trajopt_ctol_instance<M>(*dp, p, b, X, U, Xb, Ub, o, red);
  const double h = p.tf[b] / (N - 1);
  double mdx2 = 0, mnx2 = 0, J = 0, Jr = 0;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    double dx2 = 0, nx2 = 0, uu = 0, ur = 0;
    for (int i = 0; i < NX; ++i) { const double dxi = X[k * NX + i] - Xb[k * NX + i]; dx2 += dxi * dxi; nx2 += X[k * NX + i] * X[k * NX + i]; }
    for (int i = 0; i < NU; ++i) { uu += U[k * NU + i] * U[k * NU + i]; ur += Ub[k * NU + i] * Ub[k * NU + i]; }
    mdx2 = dx2 > mdx2 ? dx2 : mdx2; mnx2 = nx2 > mnx2 ? nx2 : mnx2;
    const double wk = (k == 0 || k == N - 1) ? 0.5 * h : h;
    J += wk * uu; Jr += wk * ur;
  }
  mdx2 = block_max(mdx2, red); mnx2 = block_max(mnx2, red); J = block_sum(J, red); Jr = block_sum(Jr, red);
  if (threadIdx.x == 0) { o[2] = sqrt(mdx2) / sqrt(mnx2); o[3] = J; o[4] = Jr; }
}

// K4.  Grid: B CTAs.
template <int M>
__global__ void __launch_bounds__(EVAL_THREADS, GUSTO_EVAL_MINBLOCKS) evaluate_kernel(const BatchDesc* __restrict__ dp, BatchPtrs p, double* out) {
  using T = Traits<M>;
  __shared__ double red[EVAL_THREADS];
  const int b = blockIdx.x;
  if (p.active && !p.active[b]) return;
  evaluate_instance<M>(*dp, p, b, p.Xn + (size_t)b * dp->N * T::NX, p.Un + (size_t)b * dp->N * T::NU,
                       out + (size_t)b * EVAL_NOUT, red);
}

// K5.  Grid: B CTAs: post-processing scalars of the ACCEPTED trajectory.
template <int M>
__global__ void __launch_bounds__(EVAL_THREADS) check_kernel(const BatchDesc* __restrict__ dp, BatchPtrs p, double* out) {
  using T = Traits<M>;
  __shared__ double red[EVAL_THREADS];
  const int b = blockIdx.x;
  check_instance<M>(*dp, p, b, p.Xp + (size_t)b * dp->N * T::NX, p.Up + (size_t)b * dp->N * T::NU, out + (size_t)b * CHECK_NOUT, red);
}

// K6.  One thread per (instance, knot interval): RK4 upsampling of the accepted trajectory.
template <int M>
__global__ void __launch_bounds__(128) interp_kernel(const BatchDesc* __restrict__ dp, BatchPtrs p, int nstep, double* Xfull, double* Ufull) {
  using T = Traits<M>;
  const int N = dp->N, nseg = N - 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= dp->B * nseg) return;
  const int b = t / nseg, k = t - b * nseg;
  const size_t nf = (size_t)nstep * nseg;
  interpolate_interval<M>(*dp, p, b, k, nstep, p.Xp + (size_t)b * N * T::NX, p.Up + (size_t)b * N * T::NU,
                          Xfull + (size_t)b * (nf + 1) * T::NX, Ufull + (size_t)b * nf * T::NU);
}

// K7.  One WARP per instance, SHOOT_WARPS instances per CTA (the warps of a CTA share instruction fetches: the kernel is
// ~150 KB of SASS for the quaternion model): indirect shooting, lanes = Jacobian columns.
constexpr int SHOOT_WARPS = 4;
template <int M>
__global__ void __launch_bounds__(32 * SHOOT_WARPS) shoot_kernel(const BatchDesc* __restrict__ dp, BatchPtrs p, const double* p0, const double* x_goal,
                                                   int nsub, int max_iter, double ftol, double* Xs, double* Us, double* Ps, double* out) {
  using T = Traits<M>;
  __shared__ double sm[SHOOT_WARPS][ShootLayout<M>::TOTAL];
  const int w = threadIdx.x >> 5;
  const int b = blockIdx.x * SHOOT_WARPS + w;
  if (b >= dp->B) return;
  const size_t N = dp->N;
  shoot_instance<M>(*dp, p, b, p0 + (size_t)b * T::NX, x_goal + (size_t)b * T::NX, nsub, max_iter, ftol, sm[w],
                    Xs + (size_t)b * N * T::NX, Us + (size_t)b * N * T::NU, Ps + (size_t)b * N * T::NX, out + (size_t)b * SHOOT_NOUT);
}

// accept: candidate -> accepted trajectory for flagged instances; install next omega / Delta.
__global__ void accept_kernel(BatchPtrs p, int N, int nx, int nu, const uint8_t* accept, const double* omega,
                              const double* delta) {
  const int b = blockIdx.x;
  if (accept[b]) {
    const size_t ox = (size_t)b * N * nx, ou = (size_t)b * N * nu;
    for (int i = threadIdx.x; i < N * nx; i += blockDim.x) p.Xp[ox + i] = p.Xn[ox + i];
    for (int i = threadIdx.x; i < N * nu; i += blockDim.x) p.Up[ou + i] = p.Un[ou + i];
  }
  if (threadIdx.x == 0) {
    if (omega) p.omega[b] = omega[b];
    if (delta) p.delta[b] = delta[b];
  }
}

// ------------------------------------------------------------------------------------------------ context
struct gusto_ctx {
  gusto_config cfg;
  int nx = 0, nu = 0;
  BatchDesc hdesc;
  BatchDesc* ddesc = nullptr;
  BatchPtrs p;
  double *d_tf = nullptr, *d_xinit = nullptr, *d_glo = nullptr, *d_ghi = nullptr;
  double *d_scratch = nullptr, *d_info = nullptr, *d_eval = nullptr, *d_omega_in = nullptr, *d_delta_in = nullptr;
  double *d_dual = nullptr, *d_Xs = nullptr, *d_Us = nullptr, *d_Ps = nullptr, *d_p0 = nullptr, *d_xgoal = nullptr, *d_shoot = nullptr;
  uint8_t* d_accept = nullptr;
  uint8_t* d_active = nullptr;
  double *d_check = nullptr, *d_interp = nullptr;     // outputs of gusto_check_trajectory / gusto_interpolate_trajectory
  size_t interp_cap = 0;                              // doubles allocated behind d_interp
  size_t scratch_stride = 0;
  int ipm_smem = 0;          // dynamic shared memory of ONE instance group (bytes)
  int ipm_pack_max = 1;      // groups per CTA allowed by shared memory / registers
  int ipm_pack_force = 0;    // test hook (GUSTO_IPM_FORCE_PACK): use exactly this many groups per CTA
  int nsm = 1;
  IpmParams prm;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  cudaEvent_t tev[2] = {};
  bool timed[4] = {false, false, false, false};
  int64_t launches = 0;
  std::string err;
  // ---- device-resident outer loop (gusto_scp_*) and the status all-gather
  ScpState scp = {};
  bool scp_ready = false, scp_begun = false, capturing = false;
  double *d_X0 = nullptr, *d_U0 = nullptr;       // initial trajectory of the last gusto_scp_begin (restart without H2D)
  uint8_t* d_done_all = nullptr;                 // [nranks][B] gathered status bytes
  int* d_nunf = nullptr;                         // [SCP_RING] unfinished instances over all ranks after an iteration
  int* h_nunf = nullptr;                         // pinned mirror
  cudaStream_t cstream = nullptr;                // collective + counter read-back, beside the next iteration's kernels
  cudaEvent_t sev[4] = {}, fev[4] = {};          // iteration finished on the main stream / its counter is on the host
  cudaGraphExec_t step_graph = nullptr;
  int scp_iter = 0;                              // outer iterations enqueued since gusto_scp_begin (incl. a speculative one)
  int scp_real = 0;                              // ... that ran with at least one live instance anywhere
  void* comm = nullptr;                          // ncclComm_t
  int rank = 0, nranks = 1;
  // ---- TrajOpt variant (gusto_trajopt_*): own solver scratch / launch shape, two reference trajectories, evaluation outputs
  bool to_ready = false;
  double *d_to_scratch = nullptr, *d_to_eval = nullptr, *d_to_cmp = nullptr;
  double* d_to_ref[4] = {nullptr, nullptr, nullptr, nullptr};    // X, U of slot 0 (old_penalty_traj) and slot 1 (old_convex_traj)
  size_t to_stride = 0;
  int to_smem = 0, to_pack_max = 1;
};

static std::string g_err;

#define CK(call)                                                                                 \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
      return GUSTO_E_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

static void dims_of(int model, int* nx, int* nu) {
  switch (model) {
    case DUBINS: *nx = 3; *nu = 1; break;
    case FREEFLYER_SE2: *nx = 6; *nu = 3; break;
    case ASTROBEE_SE3: *nx = 12; *nu = 6; break;
    case ASTROBEE_SE3_MANIFOLD: *nx = 13; *nu = 6; break;
    default: *nx = 0; *nu = 0;
  }
}
static int anz_of(int model) {
  switch (model) {
    case DUBINS: return blocks_anz<DUBINS>();
    case FREEFLYER_SE2: return blocks_anz<FREEFLYER_SE2>();
    case ASTROBEE_SE3: return blocks_anz<ASTROBEE_SE3>();
    default: return blocks_anz<ASTROBEE_SE3_MANIFOLD>();
  }
}
static size_t scratch_doubles_of(int model, int N, int n_obs) {
  switch (model) {
    case DUBINS: return IpmLayout<DUBINS>::scratch_doubles(N, 0);
    case FREEFLYER_SE2: return IpmLayout<FREEFLYER_SE2>::scratch_doubles(N, n_obs);
    case ASTROBEE_SE3: return IpmLayout<ASTROBEE_SE3>::scratch_doubles(N, n_obs);
    default: return IpmLayout<ASTROBEE_SE3_MANIFOLD>::scratch_doubles(N, n_obs);
  }
}
static int ipm_smem_doubles_raw(int model, int N) {
  switch (model) {
    case DUBINS: return IpmLayout<DUBINS>::smem_doubles(N, IPM_THREADS);
    case FREEFLYER_SE2: return IpmLayout<FREEFLYER_SE2>::smem_doubles(N, IPM_THREADS);
    case ASTROBEE_SE3: return IpmLayout<ASTROBEE_SE3>::smem_doubles(N, IPM_THREADS);
    default: return IpmLayout<ASTROBEE_SE3_MANIFOLD>::smem_doubles(N, IPM_THREADS);
  }
}
// per-group slice of the CTA's dynamic shared memory: a multiple of 128 bytes, so that every group's tiles keep the
// 16-byte alignment LDS.128 and the TMA ring need
static int ipm_smem_doubles_of(int model, int N) { return (ipm_smem_doubles_raw(model, N) + 15) & ~15; }

static void scp_release(gusto_ctx* ctx);

extern "C" {

int32_t gusto_version(void) { return 200; }

const char* gusto_last_error(const gusto_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int32_t gusto_create(const gusto_config* cfg, const int32_t* obs_kind, const double* obs_a, const double* obs_b,
                     gusto_ctx** out) {
  if (!cfg || !out) { g_err = "gusto_create: null argument"; return GUSTO_E_ARG; }
  *out = nullptr;
  int nx, nu;
  dims_of(cfg->model_id, &nx, &nu);
  if (nx == 0) { g_err = "gusto_create: unknown model_id"; return GUSTO_E_ARG; }
  if (cfg->N < 3 || cfg->B < 1) { g_err = "gusto_create: need N >= 3 and B >= 1"; return GUSTO_E_ARG; }
  if (cfg->n_obs < 0 || cfg->n_obs > MAX_OBS) { g_err = "gusto_create: n_obs out of range (max 64)"; return GUSTO_E_ARG; }
  if (cfg->n_obs > 0 && (!obs_kind || !obs_a || !obs_b)) { g_err = "gusto_create: obstacle table missing"; return GUSTO_E_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_err = "gusto_create: no CUDA device (this library has no CPU path)";
    return GUSTO_E_NODEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_err = "gusto_create: bad device ordinal"; return GUSTO_E_ARG; }
  gusto_ctx* ctx = new (std::nothrow) gusto_ctx();
  if (!ctx) { g_err = "gusto_create: out of host memory"; return GUSTO_E_ALLOC; }
  ctx->cfg = *cfg;
  ctx->nx = nx; ctx->nu = nu;
  auto fail = [&](int code) { g_err = ctx->err; gusto_destroy(ctx); return code; };
  if (cudaSetDevice(cfg->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return fail(GUSTO_E_CUDA); }
  BatchDesc& h = ctx->hdesc;
  memset(&h, 0, sizeof(h));
  h.model_id = cfg->model_id; h.N = cfg->N; h.B = cfg->B; h.n_obs = (cfg->model_id == DUBINS) ? 0 : cfg->n_obs;
  for (int i = 0; i < 16; ++i) h.rp[i] = cfg->robot_params[i];
  for (int i = 0; i < 10; ++i) h.sp[i] = cfg->scp_params[i];
  for (int i = 0; i < MAX_NX; ++i) h.goal_type[i] = i < nx ? cfg->goal_type[i] : 0;
  for (int i = 0; i < h.n_obs; ++i) {
    h.obs_kind[i] = obs_kind[i];
    for (int a = 0; a < 3; ++a) { h.obs_a[i][a] = obs_a[i * 3 + a]; h.obs_b[i][a] = obs_b[i * 3 + a]; }
  }
  ctx->prm.max_iter = cfg->ipm_max_iter > 0 ? cfg->ipm_max_iter : 60;
  ctx->prm.tol = cfg->ipm_tol > 0 ? cfg->ipm_tol : 1e-8;
  // PointGoal rows: penalty weight w_N = wn_base + wn_omega * omega (ipm.cuh); ipm_nref / ipm_delta_p / ipm_delta_d of the
  // configuration are accepted and ignored since round 2 (the Riccati solve has no regularisation and no refinement)
  ctx->prm.wn_base = 1e8; ctx->prm.wn_omega = 1e4;
  ipm_default_start(ctx->prm);
  if (const char* ev = getenv("GUSTO_IPM_MU0")) {      // developer override: "A[,B[,cap[,rp[,lo[,smin]]]]]" (rp = 0: no scaling with the start point's infeasibility), "0" = the tuned default start everywhere
    double a = 0, b2 = ctx->prm.mu0_b, cp = ctx->prm.mu0_cap, rp = ctx->prm.mu0_rp, lo = ctx->prm.mu0_lo, sm = ctx->prm.mu0_smin;
    const int n = sscanf(ev, "%lf,%lf,%lf,%lf,%lf,%lf", &a, &b2, &cp, &rp, &lo, &sm);
    if (n >= 1) { ctx->prm.mu0_a = a; ctx->prm.mu0_b = b2; ctx->prm.mu0_cap = cp; ctx->prm.mu0_rp = rp; ctx->prm.mu0_lo = lo; ctx->prm.mu0_smin = sm; }
  }

  const size_t B = cfg->B, N = cfg->N, no = h.n_obs > 0 ? h.n_obs : 1;
  auto alloc = [&](double** ptr, size_t n) {
    if (cudaMalloc((void**)ptr, n * sizeof(double)) != cudaSuccess) { ctx->err = "cudaMalloc failed"; return false; }
    return cudaMemsetAsync(*ptr, 0, n * sizeof(double), 0) == cudaSuccess;
  };
  bool ok = true;
  ok = ok && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 8 && ok; ++i) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i) ok = cudaEventCreate(&ctx->tev[i]) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&ctx->ddesc, sizeof(BatchDesc)) == cudaSuccess;
  ok = ok && cudaMemcpy(ctx->ddesc, &h, sizeof(BatchDesc), cudaMemcpyHostToDevice) == cudaSuccess;
  BatchPtrs& p = ctx->p;
  memset(&p, 0, sizeof(p));
  ok = ok && alloc(&ctx->d_tf, B) && alloc(&ctx->d_xinit, B * nx) && alloc(&ctx->d_glo, B * nx) && alloc(&ctx->d_ghi, B * nx);
  ok = ok && alloc(&p.Xp, B * N * nx) && alloc(&p.Up, B * N * nu) && alloc(&p.Xn, B * N * nx) && alloc(&p.Un, B * N * nu);
  ok = ok && alloc(&p.omega, B) && alloc(&p.delta, B) && alloc(&ctx->d_omega_in, B) && alloc(&ctx->d_delta_in, B);
  const size_t NPk = g_np((int)N);                                      // knot-minor block layout (BatchPtrs, common.cuh)
  ok = ok && alloc(&p.f, B * nx * NPk) && alloc(&p.A, B * anz_of(cfg->model_id) * NPk) && alloc(&p.g, B * nx * NPk) && alloc(&p.rows, B * 5 * no * NPk);
  ctx->scratch_stride = scratch_doubles_of(cfg->model_id, (int)N, h.n_obs);
  ok = ok && alloc(&ctx->d_scratch, B * ctx->scratch_stride);
  ok = ok && alloc(&ctx->d_info, B * IPM_NINFO) && alloc(&ctx->d_eval, B * EVAL_NOUT) && alloc(&ctx->d_dual, B * nx);
  ok = ok && cudaMalloc((void**)&ctx->d_accept, B) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&ctx->d_active, B) == cudaSuccess && cudaMemset(ctx->d_active, 1, B) == cudaSuccess;
  if (!ok) { if (ctx->err.empty()) ctx->err = std::string("gusto_create: ") + cudaGetErrorString(cudaGetLastError()); return fail(GUSTO_E_ALLOC); }
  p.active = ctx->d_active;
  p.dual = ctx->d_dual;
  p.tf = ctx->d_tf; p.x_init = ctx->d_xinit; p.goal_lo = ctx->d_glo; p.goal_hi = ctx->d_ghi;
  // initial penalties: omega0, Delta0
  {
    double* tmp = new double[2 * B];
    for (size_t i = 0; i < B; ++i) { tmp[i] = cfg->scp_params[SP_OMEGA0]; tmp[B + i] = cfg->scp_params[SP_DELTA0]; }
    const bool copied = cudaMemcpy(p.omega, tmp, B * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
                        cudaMemcpy(p.delta, tmp + B, B * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    delete[] tmp;
    if (!copied) { ctx->err = std::string("gusto_create: initial penalties: ") + cudaGetErrorString(cudaGetLastError()); return fail(GUSTO_E_CUDA); }
  }
  ctx->ipm_smem = ipm_smem_doubles_of(cfg->model_id, (int)N) * (int)sizeof(double);
  cudaError_t e = cudaSuccess;
  {
    int max_optin = 0, nsm = 1;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, cfg->device);
    ctx->nsm = nsm > 0 ? nsm : 1;
    int pm = ctx->ipm_smem > 0 ? max_optin / ctx->ipm_smem : 1;
    pm = pm < 1 ? 1 : (pm > IPM_MAX_PACK ? IPM_MAX_PACK : pm);
    if (const char* ev = getenv("GUSTO_IPM_PACK")) { const int v = atoi(ev); if (v >= 1 && v <= pm) pm = v; }
    ctx->ipm_pack_max = pm;
    if (const char* ev = getenv("GUSTO_IPM_FORCE_PACK")) { const int v = atoi(ev); if (v >= 1) ctx->ipm_pack_force = v > pm ? pm : v; }
  }
  const int smem_attr = ctx->ipm_smem * ctx->ipm_pack_max;
  switch (cfg->model_id) {
    case DUBINS: e = cudaFuncSetAttribute(ipm_kernel<DUBINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr); break;
    case FREEFLYER_SE2: e = cudaFuncSetAttribute(ipm_kernel<FREEFLYER_SE2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr); break;
    case ASTROBEE_SE3: e = cudaFuncSetAttribute(ipm_kernel<ASTROBEE_SE3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr); break;
    default: e = cudaFuncSetAttribute(ipm_kernel<ASTROBEE_SE3_MANIFOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr); break;
  }
  if (e != cudaSuccess) { ctx->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return fail(GUSTO_E_CUDA); }
  if (cudaDeviceSynchronize() != cudaSuccess) { ctx->err = "gusto_create: device sync failed"; return fail(GUSTO_E_CUDA); }
  *out = ctx;
  return GUSTO_OK;
}

int32_t gusto_destroy(gusto_ctx* ctx) {
  if (!ctx) return GUSTO_OK;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  BatchPtrs& p = ctx->p;
  void* ptrs[] = {ctx->ddesc, ctx->d_tf, ctx->d_xinit, ctx->d_glo, ctx->d_ghi, p.Xp, p.Up, p.Xn, p.Un, p.omega, p.delta,
                  ctx->d_omega_in, ctx->d_delta_in, p.f, p.A, p.g, p.rows, ctx->d_scratch, ctx->d_info, ctx->d_eval, ctx->d_accept, ctx->d_active,
                  ctx->d_dual, ctx->d_Xs, ctx->d_Us, ctx->d_Ps, ctx->d_p0, ctx->d_xgoal, ctx->d_shoot, ctx->d_check, ctx->d_interp};
  for (void* q : ptrs) if (q) cudaFree(q);
  void* to_ptrs[] = {ctx->d_to_scratch, ctx->d_to_eval, ctx->d_to_cmp, ctx->d_to_ref[0], ctx->d_to_ref[1], ctx->d_to_ref[2], ctx->d_to_ref[3]};
  for (void* q : to_ptrs) if (q) cudaFree(q);
  scp_release(ctx);
  for (int i = 0; i < 8; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 2; ++i) if (ctx->tev[i]) cudaEventDestroy(ctx->tev[i]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return GUSTO_OK;
}

#define NEED(c) do { if (!ctx) { g_err = "null context"; return GUSTO_E_ARG; } if (!(c)) { ctx->err = "null argument"; return GUSTO_E_ARG; } } while (0)
#define H2D(dst, src, n) CK(cudaMemcpyAsync(dst, src, (n) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream))
#define D2H(dst, src, n) CK(cudaMemcpyAsync(dst, src, (n) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream))

int32_t gusto_set_problems(gusto_ctx* ctx, const double* x_init, const double* goal_lo, const double* goal_hi, const double* tf) {
  NEED(x_init && goal_lo && goal_hi && tf);
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t B = ctx->cfg.B;
  H2D(ctx->d_xinit, x_init, B * ctx->nx); H2D(ctx->d_glo, goal_lo, B * ctx->nx); H2D(ctx->d_ghi, goal_hi, B * ctx->nx);
  H2D(ctx->d_tf, tf, B);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

static int32_t copy_traj(gusto_ctx* ctx, double* dX, double* dU, const double* X, const double* U, bool to_dev, double* hX, double* hU) {
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t nX = (size_t)ctx->cfg.B * ctx->cfg.N * ctx->nx, nU = (size_t)ctx->cfg.B * ctx->cfg.N * ctx->nu;
  if (to_dev) { H2D(dX, X, nX); H2D(dU, U, nU); }
  else { D2H(hX, dX, nX); D2H(hU, dU, nU); }
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}
int32_t gusto_set_trajectory(gusto_ctx* ctx, const double* X, const double* U) { NEED(X && U); return copy_traj(ctx, ctx->p.Xp, ctx->p.Up, X, U, true, nullptr, nullptr); }
int32_t gusto_set_candidate(gusto_ctx* ctx, const double* X, const double* U) { NEED(X && U); return copy_traj(ctx, ctx->p.Xn, ctx->p.Un, X, U, true, nullptr, nullptr); }
int32_t gusto_get_trajectory(gusto_ctx* ctx, double* X, double* U) { NEED(X && U); return copy_traj(ctx, ctx->p.Xp, ctx->p.Up, nullptr, nullptr, false, X, U); }
int32_t gusto_get_candidate(gusto_ctx* ctx, double* X, double* U) { NEED(X && U); return copy_traj(ctx, ctx->p.Xn, ctx->p.Un, nullptr, nullptr, false, X, U); }

int32_t gusto_set_penalties(gusto_ctx* ctx, const double* omega, const double* delta) {
  NEED(omega && delta);
  CK(cudaSetDevice(ctx->cfg.device));
  H2D(ctx->p.omega, omega, ctx->cfg.B); H2D(ctx->p.delta, delta, ctx->cfg.B);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

static int32_t launch_linearize(gusto_ctx* ctx) {
  const int total = ctx->cfg.B * ctx->cfg.N;
  const int grid = (total + LIN_KNOTS_PER_CTA - 1) / LIN_KNOTS_PER_CTA;
  if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  switch (ctx->cfg.model_id) {
    case DUBINS: linearize_kernel<DUBINS><<<grid, LIN_KNOTS_PER_CTA * 32, 0, ctx->stream>>>(ctx->ddesc, ctx->p); break;
    case FREEFLYER_SE2: linearize_kernel<FREEFLYER_SE2><<<grid, LIN_KNOTS_PER_CTA * 32, 0, ctx->stream>>>(ctx->ddesc, ctx->p); break;
    case ASTROBEE_SE3: linearize_kernel<ASTROBEE_SE3><<<grid, LIN_KNOTS_PER_CTA * 32, 0, ctx->stream>>>(ctx->ddesc, ctx->p); break;
    default: linearize_kernel<ASTROBEE_SE3_MANIFOLD><<<grid, LIN_KNOTS_PER_CTA * 32, 0, ctx->stream>>>(ctx->ddesc, ctx->p); break;
  }
  CK(cudaGetLastError());
  if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (!ctx->capturing) ctx->timed[0] = true;
  ctx->launches++;
  return GUSTO_OK;
}
static int32_t launch_solve(gusto_ctx* ctx) {
  // groups per CTA: as many as it takes to hold the batch in ONE wave (B / #SM, rounded up), within the shared-memory /
  // register limit; small batches stay spread over the SMs
  int pack = (ctx->cfg.B + ctx->nsm - 1) / ctx->nsm;
  pack = pack < 1 ? 1 : (pack > ctx->ipm_pack_max ? ctx->ipm_pack_max : pack);
  if (ctx->ipm_pack_force > 0) pack = ctx->ipm_pack_force;
  const int grid = (ctx->cfg.B + pack - 1) / pack, block = IPM_THREADS * pack, smem = ctx->ipm_smem * pack, sd = ctx->ipm_smem / (int)sizeof(double);
  if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  switch (ctx->cfg.model_id) {
    case DUBINS: ipm_kernel<DUBINS><<<grid, block, smem, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->prm, ctx->d_scratch, ctx->scratch_stride, ctx->d_info, sd); break;
    case FREEFLYER_SE2: ipm_kernel<FREEFLYER_SE2><<<grid, block, smem, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->prm, ctx->d_scratch, ctx->scratch_stride, ctx->d_info, sd); break;
    case ASTROBEE_SE3: ipm_kernel<ASTROBEE_SE3><<<grid, block, smem, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->prm, ctx->d_scratch, ctx->scratch_stride, ctx->d_info, sd); break;
    default: ipm_kernel<ASTROBEE_SE3_MANIFOLD><<<grid, block, smem, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->prm, ctx->d_scratch, ctx->scratch_stride, ctx->d_info, sd); break;
  }
  CK(cudaGetLastError());
  if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  if (!ctx->capturing) ctx->timed[1] = true;
  ctx->launches++;
  return GUSTO_OK;
}
static int32_t launch_evaluate(gusto_ctx* ctx) {
  const int grid = ctx->cfg.B;
  if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  switch (ctx->cfg.model_id) {
    case DUBINS: evaluate_kernel<DUBINS><<<grid, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_eval); break;
    case FREEFLYER_SE2: evaluate_kernel<FREEFLYER_SE2><<<grid, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_eval); break;
    case ASTROBEE_SE3: evaluate_kernel<ASTROBEE_SE3><<<grid, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_eval); break;
    default: evaluate_kernel<ASTROBEE_SE3_MANIFOLD><<<grid, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_eval); break;
  }
  CK(cudaGetLastError());
  if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  if (!ctx->capturing) ctx->timed[2] = true;
  ctx->launches++;
  return GUSTO_OK;
}

int32_t gusto_linearize(gusto_ctx* ctx) {
  NEED(true);
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc = launch_linearize(ctx);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_get_blocks(gusto_ctx* ctx, double* f, double* A, double* g, double* rows) {
  NEED(true);
  CK(cudaSetDevice(ctx->cfg.device));
  // test hook: the kernels keep the blocks knot-minor with A on its sparsity pattern; the caller gets the dense knot-major arrays
  const int B = ctx->cfg.B, N = ctx->cfg.N, no = ctx->hdesc.n_obs, anz = anz_of(ctx->cfg.model_id);
  const size_t np = g_np(N);
  std::vector<double> fc((size_t)B * ctx->nx * np), gc(fc.size()), Ac((size_t)B * anz * np), rc((size_t)B * 5 * (no > 0 ? no : 1) * np);
  D2H(fc.data(), ctx->p.f, fc.size()); D2H(gc.data(), ctx->p.g, gc.size()); D2H(Ac.data(), ctx->p.A, Ac.size());
  if (no > 0) D2H(rc.data(), ctx->p.rows, (size_t)B * 5 * no * np);
  CK(cudaStreamSynchronize(ctx->stream));
  const double* rcp = (rows && no > 0) ? rc.data() : nullptr;
  switch (ctx->cfg.model_id) {
    case DUBINS: blocks_unpack<DUBINS>(B, N, no, fc.data(), Ac.data(), gc.data(), rcp, f, A, g, rows); break;
    case FREEFLYER_SE2: blocks_unpack<FREEFLYER_SE2>(B, N, no, fc.data(), Ac.data(), gc.data(), rcp, f, A, g, rows); break;
    case ASTROBEE_SE3: blocks_unpack<ASTROBEE_SE3>(B, N, no, fc.data(), Ac.data(), gc.data(), rcp, f, A, g, rows); break;
    default: blocks_unpack<ASTROBEE_SE3_MANIFOLD>(B, N, no, fc.data(), Ac.data(), gc.data(), rcp, f, A, g, rows); break;
  }
  return GUSTO_OK;
}

int32_t gusto_solve_subproblem(gusto_ctx* ctx, double* info) {
  NEED(true);
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc = launch_solve(ctx);
  if (rc) return rc;
  if (info) D2H(info, ctx->d_info, (size_t)ctx->cfg.B * IPM_NINFO);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_evaluate(gusto_ctx* ctx, double* out) {
  NEED(out);
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc = launch_evaluate(ctx);
  if (rc) return rc;
  D2H(out, ctx->d_eval, (size_t)ctx->cfg.B * EVAL_NOUT);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

static int32_t launch_accept(gusto_ctx* ctx, const uint8_t* acc_dev, const double* om_dev, const double* de_dev) {
  if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  accept_kernel<<<ctx->cfg.B, 128, 0, ctx->stream>>>(ctx->p, ctx->cfg.N, ctx->nx, ctx->nu, acc_dev, om_dev, de_dev);
  CK(cudaGetLastError());
  if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  if (!ctx->capturing) ctx->timed[3] = true;
  ctx->launches++;
  return GUSTO_OK;
}

int32_t gusto_accept(gusto_ctx* ctx, const uint8_t* accept, const double* omega, const double* delta) {
  NEED(accept);
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t B = ctx->cfg.B;
  CK(cudaMemcpyAsync(ctx->d_accept, accept, B, cudaMemcpyHostToDevice, ctx->stream));
  if (omega) H2D(ctx->d_omega_in, omega, B);
  if (delta) H2D(ctx->d_delta_in, delta, B);
  int32_t rc = launch_accept(ctx, ctx->d_accept, omega ? ctx->d_omega_in : nullptr, delta ? ctx->d_delta_in : nullptr);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_set_active(gusto_ctx* ctx, const uint8_t* active) {
  NEED(active);
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(ctx->d_active, active, ctx->cfg.B, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_iterate(gusto_ctx* ctx, double* out, double* info) {
  NEED(out);
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc;
  if ((rc = launch_linearize(ctx))) return rc;
  if ((rc = launch_solve(ctx))) return rc;
  if ((rc = launch_evaluate(ctx))) return rc;
  D2H(out, ctx->d_eval, (size_t)ctx->cfg.B * EVAL_NOUT);
  if (info) D2H(info, ctx->d_info, (size_t)ctx->cfg.B * IPM_NINFO);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

// One outer iteration for a host-language loop that keeps the trajectories on the HOST: every copy is enqueued on the context
// stream around the three kernels and there is ONE synchronisation at the end (the separate calls synchronise six times).
int32_t gusto_iterate_host(gusto_ctx* ctx, const double* X, const double* U, const double* omega, const double* delta, const uint8_t* active,
                           double* out, double* info, double* Xn, double* Un) {
  NEED(out);
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t B = ctx->cfg.B, nX = B * ctx->cfg.N * ctx->nx, nU = B * ctx->cfg.N * ctx->nu;
  if ((X == nullptr) != (U == nullptr) || (Xn == nullptr) != (Un == nullptr)) { ctx->err = "gusto_iterate_host: X/U and Xn/Un come in pairs"; return GUSTO_E_ARG; }
  if (X) { H2D(ctx->p.Xp, X, nX); H2D(ctx->p.Up, U, nU); }
  if (omega) H2D(ctx->p.omega, omega, B);
  if (delta) H2D(ctx->p.delta, delta, B);
  if (active) CK(cudaMemcpyAsync(ctx->d_active, active, B, cudaMemcpyHostToDevice, ctx->stream));
  int32_t rc;
  if ((rc = launch_linearize(ctx)) || (rc = launch_solve(ctx)) || (rc = launch_evaluate(ctx))) return rc;
  D2H(out, ctx->d_eval, B * EVAL_NOUT);
  if (info) D2H(info, ctx->d_info, B * IPM_NINFO);
  if (Xn) { D2H(Xn, ctx->p.Xn, nX); D2H(Un, ctx->p.Un, nU); }
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_iterate_device(gusto_ctx* ctx) {
  NEED(true);
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc;
  if ((rc = launch_linearize(ctx))) return rc;
  if ((rc = launch_solve(ctx))) return rc;
  if ((rc = launch_evaluate(ctx))) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_accept_device(gusto_ctx* ctx, const uint8_t* accept_dev, const double* omega_dev, const double* delta_dev) {
  NEED(accept_dev);
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc = launch_accept(ctx, accept_dev, omega_dev, delta_dev);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

// ------------------------------------------------------------------------------------------------ TrajOpt variant
static bool trajopt_model(int model) { return model == FREEFLYER_SE2 || model == ASTROBEE_SE3; }

int32_t gusto_trajopt_enable(gusto_ctx* ctx) {
  NEED(true);
  if (ctx->to_ready) return GUSTO_OK;
  // SCPParam_TrajOpt exists for these models only among the in-scope ones; astrobeeSE3manifold would need its quaternion rows as
  // hard per-knot equalities (scp_trajopt.jl:199-207), which the Riccati solve does not carry
  if (!trajopt_model(ctx->cfg.model_id)) { ctx->err = "gusto_trajopt_enable: TrajOpt is built for freeflyerSE2 and astrobeeSE3"; return GUSTO_E_ARG; }
  CK(cudaSetDevice(ctx->cfg.device));
  const int N = ctx->cfg.N, no = ctx->hdesc.n_obs;
  const size_t B = ctx->cfg.B, nX = B * N * ctx->nx, nU = B * N * ctx->nu;
  int sd;
  if (ctx->cfg.model_id == FREEFLYER_SE2) { ctx->to_stride = ipm_trajopt::IpmLayout<FREEFLYER_SE2>::scratch_doubles(N, no); sd = ipm_trajopt::IpmLayout<FREEFLYER_SE2>::smem_doubles(N, IPM_THREADS); }
  else { ctx->to_stride = ipm_trajopt::IpmLayout<ASTROBEE_SE3>::scratch_doubles(N, no); sd = ipm_trajopt::IpmLayout<ASTROBEE_SE3>::smem_doubles(N, IPM_THREADS); }
  ctx->to_smem = ((sd + 15) & ~15) * (int)sizeof(double);
  int max_optin = 0;
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->cfg.device);
  int pm = max_optin / ctx->to_smem;
  ctx->to_pack_max = pm < 1 ? 1 : (pm > IPM_MAX_PACK ? IPM_MAX_PACK : pm);
  const int smem_attr = ctx->to_smem * ctx->to_pack_max;
  if (ctx->cfg.model_id == FREEFLYER_SE2) CK(cudaFuncSetAttribute(ipm_trajopt_kernel<FREEFLYER_SE2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
  else CK(cudaFuncSetAttribute(ipm_trajopt_kernel<ASTROBEE_SE3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
  auto alloc = [&](double** ptr, size_t n) {
    if (cudaMalloc((void**)ptr, n * sizeof(double)) != cudaSuccess) return false;
    return cudaMemset(*ptr, 0, n * sizeof(double)) == cudaSuccess;
  };
  const bool ok = alloc(&ctx->d_to_scratch, B * ctx->to_stride) && alloc(&ctx->d_to_eval, B * TRAJOPT_NOUT) && alloc(&ctx->d_to_cmp, B * 5) &&
                  alloc(&ctx->d_to_ref[0], nX) && alloc(&ctx->d_to_ref[1], nU) && alloc(&ctx->d_to_ref[2], nX) && alloc(&ctx->d_to_ref[3], nU);
  if (!ok) {
    double** all[] = {&ctx->d_to_scratch, &ctx->d_to_eval, &ctx->d_to_cmp, &ctx->d_to_ref[0], &ctx->d_to_ref[1], &ctx->d_to_ref[2], &ctx->d_to_ref[3]};
    for (double** q : all) { if (*q) cudaFree(*q); *q = nullptr; }
    ctx->err = "gusto_trajopt_enable: cudaMalloc failed";
    return GUSTO_E_ALLOC;
  }
  ctx->to_ready = true;
  return GUSTO_OK;
}

static int32_t launch_trajopt_solve(gusto_ctx* ctx) {
  int pack = (ctx->cfg.B + ctx->nsm - 1) / ctx->nsm;
  pack = pack < 1 ? 1 : (pack > ctx->to_pack_max ? ctx->to_pack_max : pack);
  const int grid = (ctx->cfg.B + pack - 1) / pack, block = IPM_THREADS * pack, smem = ctx->to_smem * pack, sd = ctx->to_smem / (int)sizeof(double);
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (ctx->cfg.model_id == FREEFLYER_SE2) ipm_trajopt_kernel<FREEFLYER_SE2><<<grid, block, smem, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->prm, ctx->d_to_scratch, ctx->to_stride, ctx->d_info, sd);
  else ipm_trajopt_kernel<ASTROBEE_SE3><<<grid, block, smem, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->prm, ctx->d_to_scratch, ctx->to_stride, ctx->d_info, sd);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed[1] = true;
  ctx->launches++;
  return GUSTO_OK;
}
static int32_t launch_trajopt_evaluate(gusto_ctx* ctx) {
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  if (ctx->cfg.model_id == FREEFLYER_SE2) trajopt_evaluate_kernel<FREEFLYER_SE2><<<ctx->cfg.B, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_to_eval);
  else trajopt_evaluate_kernel<ASTROBEE_SE3><<<ctx->cfg.B, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_to_eval);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev[5], ctx->stream));
  ctx->timed[2] = true;
  ctx->launches++;
  return GUSTO_OK;
}

int32_t gusto_trajopt_iterate(gusto_ctx* ctx, const double* mu, const double* s, const uint8_t* active, double* out, double* info) {
  NEED(out);
  if (!ctx->to_ready) { ctx->err = "gusto_trajopt_iterate: call gusto_trajopt_enable first"; return GUSTO_E_ARG; }
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t B = ctx->cfg.B;
  if (mu) H2D(ctx->p.omega, mu, B);
  if (s) H2D(ctx->p.delta, s, B);
  if (active) CK(cudaMemcpyAsync(ctx->d_active, active, B, cudaMemcpyHostToDevice, ctx->stream));
  int32_t rc;
  if ((rc = launch_linearize(ctx)) || (rc = launch_trajopt_solve(ctx)) || (rc = launch_trajopt_evaluate(ctx))) return rc;
  D2H(out, ctx->d_to_eval, B * TRAJOPT_NOUT);
  if (info) D2H(info, ctx->d_info, B * IPM_NINFO);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_trajopt_mark(gusto_ctx* ctx, int32_t slot, const uint8_t* which) {
  NEED(true);
  if (!ctx->to_ready || slot < 0 || slot > 1) { ctx->err = "gusto_trajopt_mark: enable first; slot is 0 or 1"; return GUSTO_E_ARG; }
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t B = ctx->cfg.B, sx = (size_t)ctx->cfg.N * ctx->nx, su = (size_t)ctx->cfg.N * ctx->nu;
  if (!which) {
    CK(cudaMemcpyAsync(ctx->d_to_ref[2 * slot], ctx->p.Xp, B * sx * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_to_ref[2 * slot + 1], ctx->p.Up, B * su * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    for (size_t b = 0; b < B; ++b) {          // runs of flagged instances, one copy per run
      if (!which[b]) continue;
      size_t e = b;
      while (e + 1 < B && which[e + 1]) ++e;
      CK(cudaMemcpyAsync(ctx->d_to_ref[2 * slot] + b * sx, ctx->p.Xp + b * sx, (e - b + 1) * sx * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      CK(cudaMemcpyAsync(ctx->d_to_ref[2 * slot + 1] + b * su, ctx->p.Up + b * su, (e - b + 1) * su * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      b = e;
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_trajopt_compare(gusto_ctx* ctx, int32_t slot, double* out) {
  NEED(out);
  if (!ctx->to_ready || slot < 0 || slot > 1) { ctx->err = "gusto_trajopt_compare: enable first; slot is 0 or 1"; return GUSTO_E_ARG; }
  CK(cudaSetDevice(ctx->cfg.device));
  if (ctx->cfg.model_id == FREEFLYER_SE2) trajopt_compare_kernel<FREEFLYER_SE2><<<ctx->cfg.B, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_to_ref[2 * slot], ctx->d_to_ref[2 * slot + 1], ctx->d_to_cmp);
  else trajopt_compare_kernel<ASTROBEE_SE3><<<ctx->cfg.B, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_to_ref[2 * slot], ctx->d_to_ref[2 * slot + 1], ctx->d_to_cmp);
  CK(cudaGetLastError());
  ctx->launches++;
  D2H(out, ctx->d_to_cmp, (size_t)ctx->cfg.B * 5);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_last_kernel_ms(gusto_ctx* ctx, float* ms) {
  NEED(ms);
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 4; ++i) {
    ms[i] = 0.f;
    if (ctx->timed[i]) CK(cudaEventElapsedTime(&ms[i], ctx->ev[2 * i], ctx->ev[2 * i + 1]));
  }
  return GUSTO_OK;
}

int32_t gusto_timer_start(gusto_ctx* ctx) {
  NEED(true);
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventRecord(ctx->tev[0], ctx->stream));
  return GUSTO_OK;
}
int32_t gusto_timer_stop(gusto_ctx* ctx, float* ms) {
  NEED(ms);
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaEventRecord(ctx->tev[1], ctx->stream));
  CK(cudaEventSynchronize(ctx->tev[1]));
  CK(cudaEventElapsedTime(ms, ctx->tev[0], ctx->tev[1]));
  return GUSTO_OK;
}

int32_t gusto_check_trajectory(gusto_ctx* ctx, double* out) {
  NEED(out);
  CK(cudaSetDevice(ctx->cfg.device));
  const int grid = ctx->cfg.B;
  if (!ctx->d_check) CK(cudaMalloc((void**)&ctx->d_check, (size_t)grid * CHECK_NOUT * sizeof(double)));
  double* d_out = ctx->d_check;
  switch (ctx->cfg.model_id) {
    case DUBINS: check_kernel<DUBINS><<<grid, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, d_out); break;
    case FREEFLYER_SE2: check_kernel<FREEFLYER_SE2><<<grid, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, d_out); break;
    case ASTROBEE_SE3: check_kernel<ASTROBEE_SE3><<<grid, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, d_out); break;
    default: check_kernel<ASTROBEE_SE3_MANIFOLD><<<grid, EVAL_THREADS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, d_out); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)grid * CHECK_NOUT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { ctx->err = std::string("gusto_check_trajectory: ") + cudaGetErrorString(e); return GUSTO_E_CUDA; }
  ctx->launches++;
  return GUSTO_OK;
}

int32_t gusto_interpolate_trajectory(gusto_ctx* ctx, int32_t nstep, double* Xfull, double* Ufull) {
  NEED(Xfull && Ufull);
  if (nstep < 1) { ctx->err = "gusto_interpolate_trajectory: nstep must be >= 1"; return GUSTO_E_ARG; }
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t B = ctx->cfg.B, nseg = ctx->cfg.N - 1, nf = (size_t)nstep * nseg;
  const size_t nX = B * (nf + 1) * ctx->nx, nU = B * nf * ctx->nu;
  if (ctx->interp_cap < nX + nU) {                     // grown on demand, kept for the next call
    if (ctx->d_interp) { cudaFree(ctx->d_interp); ctx->d_interp = nullptr; ctx->interp_cap = 0; }
    if (cudaMalloc((void**)&ctx->d_interp, (nX + nU) * sizeof(double)) != cudaSuccess) {
      ctx->d_interp = nullptr; ctx->err = "gusto_interpolate_trajectory: cudaMalloc failed"; return GUSTO_E_ALLOC;
    }
    ctx->interp_cap = nX + nU;
  }
  double *dX = ctx->d_interp, *dU = ctx->d_interp + nX;
  const int total = (int)(B * nseg), grid = (total + 127) / 128;
  switch (ctx->cfg.model_id) {
    case DUBINS: interp_kernel<DUBINS><<<grid, 128, 0, ctx->stream>>>(ctx->ddesc, ctx->p, nstep, dX, dU); break;
    case FREEFLYER_SE2: interp_kernel<FREEFLYER_SE2><<<grid, 128, 0, ctx->stream>>>(ctx->ddesc, ctx->p, nstep, dX, dU); break;
    case ASTROBEE_SE3: interp_kernel<ASTROBEE_SE3><<<grid, 128, 0, ctx->stream>>>(ctx->ddesc, ctx->p, nstep, dX, dU); break;
    default: interp_kernel<ASTROBEE_SE3_MANIFOLD><<<grid, 128, 0, ctx->stream>>>(ctx->ddesc, ctx->p, nstep, dX, dU); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(Xfull, dX, nX * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(Ufull, dU, nU * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { ctx->err = std::string("gusto_interpolate_trajectory: ") + cudaGetErrorString(e); return GUSTO_E_CUDA; }
  ctx->launches++;
  return GUSTO_OK;
}

int32_t gusto_get_duals(gusto_ctx* ctx, double* dual) {
  NEED(dual);
  CK(cudaSetDevice(ctx->cfg.device));
  D2H(dual, ctx->d_dual, (size_t)ctx->cfg.B * ctx->nx);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

// buffers of the shooting refinement, all or nothing: a later call must never find d_Xs set and the others missing
static int32_t shoot_alloc(gusto_ctx* ctx) {
  if (ctx->d_Xs) return GUSTO_OK;
  const size_t B = ctx->cfg.B, N = ctx->cfg.N, nx = ctx->nx, nu = ctx->nu;
  const bool ok = cudaMalloc((void**)&ctx->d_Xs, B * N * nx * sizeof(double)) == cudaSuccess &&
                  cudaMalloc((void**)&ctx->d_Us, B * N * nu * sizeof(double)) == cudaSuccess &&
                  cudaMalloc((void**)&ctx->d_Ps, B * N * nx * sizeof(double)) == cudaSuccess &&
                  cudaMalloc((void**)&ctx->d_p0, B * nx * sizeof(double)) == cudaSuccess &&
                  cudaMalloc((void**)&ctx->d_xgoal, B * nx * sizeof(double)) == cudaSuccess &&
                  cudaMalloc((void**)&ctx->d_shoot, B * SHOOT_NOUT * sizeof(double)) == cudaSuccess;
  if (!ok) {
    double** bufs[] = {&ctx->d_Xs, &ctx->d_Us, &ctx->d_Ps, &ctx->d_p0, &ctx->d_xgoal, &ctx->d_shoot};
    for (double** q : bufs) { if (*q) cudaFree(*q); *q = nullptr; }
    ctx->err = "gusto_shoot: cudaMalloc failed";
    return GUSTO_E_ALLOC;
  }
  return GUSTO_OK;
}

int32_t gusto_set_shooting_trajectory(gusto_ctx* ctx, const double* X, const double* U) {
  NEED(X && U);
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc = shoot_alloc(ctx);
  if (rc) return rc;
  const size_t B = ctx->cfg.B, N = ctx->cfg.N;
  H2D(ctx->d_Xs, X, B * N * ctx->nx); H2D(ctx->d_Us, U, B * N * ctx->nu);
  CK(cudaMemsetAsync(ctx->d_Ps, 0, B * N * ctx->nx * sizeof(double), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_shoot(gusto_ctx* ctx, const double* p0, const double* x_goal, int32_t nsub, int32_t max_iter, double ftol, double* out) {
  NEED(out);
  const int model = ctx->cfg.model_id;
  if (model != DUBINS && model != ASTROBEE_SE3_MANIFOLD) {
    ctx->err = "gusto_shoot: the reference defines shooting_ode! only for DubinsCar and AstrobeeSE3Manifold"; return GUSTO_E_ARG;
  }
  if (nsub < 1 || max_iter < 0 || !(ftol > 0.0)) { ctx->err = "gusto_shoot: need nsub >= 1, max_iter >= 0, ftol > 0"; return GUSTO_E_ARG; }
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t B = ctx->cfg.B, N = ctx->cfg.N, nx = ctx->nx, nu = ctx->nu;
  if (!ctx->d_Xs) {    // SS.traj not seeded by gusto_set_shooting_trajectory: it starts as the trajectory held by the context
    int32_t rc = shoot_alloc(ctx);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_Xs, ctx->p.Xp, B * N * nx * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_Us, ctx->p.Up, B * N * nu * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_Ps, 0, B * N * nx * sizeof(double), ctx->stream));
  }
  if (p0) H2D(ctx->d_p0, p0, B * nx);
  else CK(cudaMemcpyAsync(ctx->d_p0, ctx->d_dual, B * nx * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (x_goal) H2D(ctx->d_xgoal, x_goal, B * nx);
  else {               // ShootingProblem (types.jl:219-226): centre of the goal sets
    std::vector<double> lo(B * nx), hi(B * nx);
    D2H(lo.data(), ctx->d_glo, B * nx);
    D2H(hi.data(), ctx->d_ghi, B * nx);
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < B * nx; ++i) lo[i] = 0.5 * (lo[i] + hi[i]);
    H2D(ctx->d_xgoal, lo.data(), B * nx);
    CK(cudaStreamSynchronize(ctx->stream));
  }
  const int grid = (int)((B + SHOOT_WARPS - 1) / SHOOT_WARPS);
  if (model == DUBINS)
    shoot_kernel<DUBINS><<<grid, 32 * SHOOT_WARPS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_p0, ctx->d_xgoal, nsub, max_iter, ftol, ctx->d_Xs, ctx->d_Us, ctx->d_Ps, ctx->d_shoot);
  else
    shoot_kernel<ASTROBEE_SE3_MANIFOLD><<<grid, 32 * SHOOT_WARPS, 0, ctx->stream>>>(ctx->ddesc, ctx->p, ctx->d_p0, ctx->d_xgoal, nsub, max_iter, ftol, ctx->d_Xs, ctx->d_Us, ctx->d_Ps, ctx->d_shoot);
  CK(cudaGetLastError());
  ctx->launches++;
  D2H(out, ctx->d_shoot, B * SHOOT_NOUT);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int32_t gusto_get_shooting_trajectory(gusto_ctx* ctx, double* X, double* U, double* P) {
  NEED(X || U || P);
  if (!ctx->d_Xs) { ctx->err = "gusto_get_shooting_trajectory: gusto_shoot has not been called"; return GUSTO_E_ARG; }
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t B = ctx->cfg.B, N = ctx->cfg.N;
  if (X) D2H(X, ctx->d_Xs, B * N * ctx->nx);
  if (U) D2H(U, ctx->d_Us, B * N * ctx->nu);
  if (P) D2H(P, ctx->d_Ps, B * N * ctx->nx);
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}


}  // extern "C" (reopened below)

// ------------------------------------------------------------------ device-resident outer loop + status all-gather
// NCCL is bound at run time (dlopen) so that a single-GPU host needs no NCCL at all; the few entry points used are declared
// here with their nccl.h signatures (NCCL 2.x ABI: ncclUniqueId is 128 bytes passed by value, ncclUint8 = 1).
namespace {
struct NcclUid { char internal[128]; };
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool nccl_load(std::string& err) {
  if (g_nccl.h) return true;
  const char* names[] = {getenv("GUSTO_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) { err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return false; }
  NcclApi a;
  a.h = h;
  a.GetUniqueId = (int (*)(NcclUid*))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(void**, int, NcclUid, int))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  a.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
  a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather) { err = "NCCL library lacks the expected symbols"; return false; }
  g_nccl = a;
  return true;
}
constexpr int NCCL_UINT8 = 1;
constexpr int SCP_RING = 4;
}  // namespace

static const double* dev_sp(const gusto_ctx* ctx) { return reinterpret_cast<const double*>(reinterpret_cast<const char*>(ctx->ddesc) + offsetof(BatchDesc, sp)); }

static void scp_release(gusto_ctx* ctx) {
  if (ctx->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; }
  if (ctx->step_graph) { cudaGraphExecDestroy(ctx->step_graph); ctx->step_graph = nullptr; }
  ScpState& s = ctx->scp;
  void* ptrs[] = {s.slot, s.iterations, s.conv_prev, s.j_true, s.j_full, s.converged, s.successful, s.done, s.cnt, s.hist,
                  ctx->d_X0, ctx->d_U0, ctx->d_done_all, ctx->d_nunf};
  for (void* q : ptrs) if (q) cudaFree(q);
  if (ctx->h_nunf) cudaFreeHost(ctx->h_nunf);
  for (int i = 0; i < 4; ++i) { if (ctx->sev[i]) cudaEventDestroy(ctx->sev[i]); if (ctx->fev[i]) cudaEventDestroy(ctx->fev[i]); }
  if (ctx->cstream) cudaStreamDestroy(ctx->cstream);
  s = ScpState{};
  ctx->d_X0 = ctx->d_U0 = nullptr; ctx->d_done_all = nullptr; ctx->d_nunf = nullptr; ctx->h_nunf = nullptr; ctx->cstream = nullptr;
  ctx->scp_ready = false;
}

// buffers of the device-resident loop, allocated on first use (or re-allocated when the communicator changes the gather size)
static int32_t scp_prepare(gusto_ctx* ctx) {
  if (ctx->scp_ready) return GUSTO_OK;
  const size_t B = ctx->cfg.B, N = ctx->cfg.N;
  ScpState& s = ctx->scp;
  s.max_hist = GUSTO_SCP_MAX_HIST;
  bool ok = true;
  auto al = [&](void** q, size_t bytes) { ok = ok && cudaMalloc(q, bytes) == cudaSuccess && cudaMemset(*q, 0, bytes) == cudaSuccess; };
  al((void**)&s.slot, sizeof(int)); al((void**)&s.iterations, B * sizeof(int));
  al((void**)&s.conv_prev, B * sizeof(double)); al((void**)&s.j_true, B * sizeof(double)); al((void**)&s.j_full, B * sizeof(double));
  al((void**)&s.converged, B); al((void**)&s.successful, B); al((void**)&s.done, 2 * B);
  al((void**)&s.cnt, (size_t)s.max_hist * CNT_W * sizeof(int));
  al((void**)&s.hist, (size_t)(s.max_hist + 1) * B * HIST_W * sizeof(double));
  al((void**)&ctx->d_X0, B * N * ctx->nx * sizeof(double)); al((void**)&ctx->d_U0, B * N * ctx->nu * sizeof(double));
  al((void**)&ctx->d_done_all, (size_t)ctx->nranks * B);
  al((void**)&ctx->d_nunf, SCP_RING * sizeof(int));
  ok = ok && cudaMallocHost((void**)&ctx->h_nunf, SCP_RING * sizeof(int)) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->cstream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < SCP_RING && ok; ++i)
    ok = cudaEventCreateWithFlags(&ctx->sev[i], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ctx->fev[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { ctx->err = std::string("gusto_scp: allocation failed: ") + cudaGetErrorString(cudaGetLastError()); scp_release(ctx); return GUSTO_E_ALLOC; }
  s.active = ctx->d_active;
  ctx->scp_ready = true;
  return GUSTO_OK;
}

static int32_t launch_scp_update(gusto_ctx* ctx) {
  const int B = ctx->cfg.B;
  scp_update_kernel<<<B, 128, 0, ctx->stream>>>(ctx->p, ctx->scp, ctx->d_eval, ctx->d_info, B, ctx->cfg.N, ctx->nx, ctx->nu, EVAL_NOUT, IPM_NINFO,
                                                dev_sp(ctx));
  CK(cudaGetLastError());
  scp_advance_kernel<<<1, 1, 0, ctx->stream>>>(ctx->scp);
  CK(cudaGetLastError());
  ctx->launches += 2;
  return GUSTO_OK;
}

// one outer iteration on the main stream: K1 -> K3 -> K4 -> update (+ accept) -> advance; a CUDA graph after the first call
static int32_t enqueue_scp_step(gusto_ctx* ctx) {
  static const bool no_graph = getenv("GUSTO_NO_GRAPH") != nullptr;
  int32_t rc = GUSTO_OK;
  if (no_graph) {
    if ((rc = launch_linearize(ctx)) || (rc = launch_solve(ctx)) || (rc = launch_evaluate(ctx)) || (rc = launch_scp_update(ctx))) return rc;
    return GUSTO_OK;
  }
  if (!ctx->step_graph) {
    cudaGraph_t g = nullptr;
    const int64_t l0 = ctx->launches;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    if (!(rc = launch_linearize(ctx)) && !(rc = launch_solve(ctx)) && !(rc = launch_evaluate(ctx))) rc = launch_scp_update(ctx);
    ctx->capturing = false;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    ctx->launches = l0;
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return GUSTO_E_CUDA; }
    e = cudaGraphInstantiate(&ctx->step_graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { ctx->step_graph = nullptr; ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return GUSTO_E_CUDA; }
  }
  CK(cudaGraphLaunch(ctx->step_graph, ctx->stream));
  ctx->launches += 5;
  return GUSTO_OK;
}

// after iteration `it` (its status bytes sit in half it & 1 of scp.done): gather over the ranks, count the unfinished ones,
// bring the count to the host -- all on the side stream, beside the next iteration's kernels
static int32_t enqueue_status(gusto_ctx* ctx, int it) {
  const int r = it % SCP_RING, B = ctx->cfg.B;
  CK(cudaEventRecord(ctx->sev[r], ctx->stream));
  CK(cudaStreamWaitEvent(ctx->cstream, ctx->sev[r], 0));
  const uint8_t* send = ctx->scp.done + (size_t)(it & 1) * B;
  const uint8_t* all = send;
  if (ctx->comm) {
    const int e = g_nccl.AllGather(send, ctx->d_done_all, (size_t)B, NCCL_UINT8, ctx->comm, ctx->cstream);
    if (e != 0) { ctx->err = std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error"); return GUSTO_E_CUDA; }
    all = ctx->d_done_all;
  }
  scp_count_kernel<<<1, 256, 0, ctx->cstream>>>(all, B * ctx->nranks, ctx->d_nunf + r);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(ctx->h_nunf + r, ctx->d_nunf + r, sizeof(int), cudaMemcpyDeviceToHost, ctx->cstream));
  CK(cudaEventRecord(ctx->fev[r], ctx->cstream));
  return GUSTO_OK;
}

extern "C" {

int32_t gusto_comm_unique_id(uint8_t* id) {
  if (!id) { g_err = "gusto_comm_unique_id: null argument"; return GUSTO_E_ARG; }
  if (!nccl_load(g_err)) return GUSTO_E_STATE;
  NcclUid u;
  const int e = g_nccl.GetUniqueId(&u);
  if (e != 0) { g_err = "ncclGetUniqueId failed"; return GUSTO_E_CUDA; }
  memcpy(id, u.internal, 128);
  return GUSTO_OK;
}

int32_t gusto_comm_init(gusto_ctx* ctx, int32_t rank, int32_t nranks, const uint8_t* id) {
  NEED(id);
  if (nranks < 1 || rank < 0 || rank >= nranks) { ctx->err = "gusto_comm_init: bad rank / nranks"; return GUSTO_E_ARG; }
  if (!nccl_load(ctx->err)) return GUSTO_E_STATE;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaStreamSynchronize(ctx->stream));
  scp_release(ctx);                                   // the gather buffer depends on nranks
  NcclUid u;
  memcpy(u.internal, id, 128);
  void* comm = nullptr;
  const int e = g_nccl.CommInitRank(&comm, nranks, u, rank);
  if (e != 0) { ctx->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error"); return GUSTO_E_CUDA; }
  ctx->comm = comm; ctx->rank = rank; ctx->nranks = nranks;
  return scp_prepare(ctx);
}

int32_t gusto_allgather_status(gusto_ctx* ctx, const uint8_t* done_local, uint8_t* done_all, int32_t* n_unfinished) {
  NEED(done_local);
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc = scp_prepare(ctx);
  if (rc) return rc;
  const size_t B = ctx->cfg.B;
  CK(cudaMemcpyAsync(ctx->scp.done, done_local, B, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = enqueue_status(ctx, 0))) return rc;
  if (done_all) {
    const uint8_t* src = ctx->comm ? ctx->d_done_all : ctx->scp.done;
    CK(cudaMemcpyAsync(done_all, src, B * ctx->nranks, cudaMemcpyDeviceToHost, ctx->cstream));
  }
  CK(cudaStreamSynchronize(ctx->cstream));
  if (n_unfinished) *n_unfinished = ctx->h_nunf[0];
  return GUSTO_OK;
}

int32_t gusto_scp_begin(gusto_ctx* ctx, const double* X0, const double* U0, int32_t force) {
  NEED(true);
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc = scp_prepare(ctx);
  if (rc) return rc;
  const size_t B = ctx->cfg.B, nX = B * ctx->cfg.N * ctx->nx, nU = B * ctx->cfg.N * ctx->nu;
  if ((X0 == nullptr) != (U0 == nullptr)) { ctx->err = "gusto_scp_begin: X0 and U0 must both be given or both be NULL"; return GUSTO_E_ARG; }
  if (X0) { H2D(ctx->d_X0, X0, nX); H2D(ctx->d_U0, U0, nU); }
  else if (!ctx->scp_begun) { ctx->err = "gusto_scp_begin: no stored initial trajectory (first call needs X0, U0)"; return GUSTO_E_STATE; }
  // traj = candidate = traj_init; every instance live; J_true[1], rho_vec[2] from K4 on (traj, traj)  (scp_gusto.jl:60-75)
  CK(cudaMemcpyAsync(ctx->p.Xp, ctx->d_X0, nX * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->p.Up, ctx->d_U0, nU * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->p.Xn, ctx->d_X0, nX * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->p.Un, ctx->d_U0, nU * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_active, 1, B, ctx->stream));
  ctx->scp.force = force ? 1 : 0;
  {
    const int n = (int)B > ctx->scp.max_hist * CNT_W ? (int)B : ctx->scp.max_hist * CNT_W;
    // omega0 / Delta0 first: K4 reads them
    scp_begin_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->p, ctx->scp, ctx->d_eval, (int)B, EVAL_NOUT, dev_sp(ctx), 0);
    CK(cudaGetLastError());
    if ((rc = launch_linearize(ctx)) || (rc = launch_evaluate(ctx))) return rc;
    scp_begin_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->p, ctx->scp, ctx->d_eval, (int)B, EVAL_NOUT, dev_sp(ctx), 1);
    CK(cudaGetLastError());
    ctx->launches += 2;
  }
  ctx->scp_begun = true; ctx->scp_iter = 0; ctx->scp_real = 0;
  return GUSTO_OK;
}

int32_t gusto_scp_run(gusto_ctx* ctx, int32_t max_iter, int32_t* batch_iterations, int32_t* n_unfinished) {
  NEED(true);
  if (!ctx->scp_begun) { ctx->err = "gusto_scp_run: call gusto_scp_begin first"; return GUSTO_E_STATE; }
  CK(cudaSetDevice(ctx->cfg.device));
  int32_t rc;
  int ran = 0, last = -1;
  for (int i = 0; i < max_iter; ++i) {
    const int it = ctx->scp_iter;
    if ((rc = enqueue_scp_step(ctx)) || (rc = enqueue_status(ctx, it))) return rc;
    ctx->scp_iter++;
    ran++;
    if (i >= 1) {        // the count of the previous iteration, read while this one runs: 0 => this one was a no-op
      CK(cudaEventSynchronize(ctx->fev[(it - 1) % SCP_RING]));
      last = ctx->h_nunf[(it - 1) % SCP_RING];
      if (last == 0) { ran--; break; }
    }
  }
  if (last != 0 && ran > 0) {
    const int it = ctx->scp_iter - 1;
    CK(cudaEventSynchronize(ctx->fev[it % SCP_RING]));
    last = ctx->h_nunf[it % SCP_RING];
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->scp_real += ran;
  if (batch_iterations) *batch_iterations = ran;
  if (n_unfinished) *n_unfinished = last < 0 ? 0 : last;
  return GUSTO_OK;
}

int32_t gusto_scp_get(gusto_ctx* ctx, int32_t* iterations, uint8_t* converged, uint8_t* successful, double* hist, int32_t n_hist,
                      int32_t* counters) {
  NEED(true);
  if (!ctx->scp_begun) { ctx->err = "gusto_scp_get: call gusto_scp_begin first"; return GUSTO_E_STATE; }
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t B = ctx->cfg.B;
  if (n_hist < 0 || n_hist > ctx->scp.max_hist + 1) { ctx->err = "gusto_scp_get: n_hist out of range"; return GUSTO_E_ARG; }
  if (iterations) CK(cudaMemcpyAsync(iterations, ctx->scp.iterations, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (converged) CK(cudaMemcpyAsync(converged, ctx->scp.converged, B, cudaMemcpyDeviceToHost, ctx->stream));
  if (successful) CK(cudaMemcpyAsync(successful, ctx->scp.successful, B, cudaMemcpyDeviceToHost, ctx->stream));
  if (hist && n_hist > 0) D2H(hist, ctx->scp.hist, (size_t)n_hist * B * HIST_W);
  if (counters && n_hist > 1) CK(cudaMemcpyAsync(counters, ctx->scp.cnt, (size_t)(n_hist - 1) * CNT_W * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GUSTO_OK;
}

int64_t gusto_launch_count(const gusto_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t gusto_device_ptr(gusto_ctx* ctx, int32_t which, void** ptr, int64_t* n) {
  NEED(ptr && n);
  const int64_t B = ctx->cfg.B, N = ctx->cfg.N;
  switch (which) {
    case 0: *ptr = ctx->d_eval; *n = B * EVAL_NOUT; break;
    case 1: *ptr = ctx->d_info; *n = B * IPM_NINFO; break;
    case 2: *ptr = ctx->p.Xp; *n = B * N * ctx->nx; break;
    case 3: *ptr = ctx->p.Up; *n = B * N * ctx->nu; break;
    case 4: *ptr = ctx->p.Xn; *n = B * N * ctx->nx; break;
    case 5: *ptr = ctx->p.Un; *n = B * N * ctx->nu; break;
    case 6: *ptr = ctx->p.omega; *n = B; break;
    case 7: *ptr = ctx->p.delta; *n = B; break;
    case 8: *ptr = ctx->d_active; *n = B; break;   /* bytes, not doubles */
    default: ctx->err = "gusto_device_ptr: bad selector"; return GUSTO_E_ARG;
  }
  return GUSTO_OK;
}

int64_t gusto_stream_handle(gusto_ctx* ctx) { return ctx ? (int64_t)(uintptr_t)ctx->stream : 0; }

}  // extern "C"
