// Shared definitions for the gusto-b200 CUDA kernels.
//
// Every kernel body is written in CTA-SPMD style: work is expressed as strided loops over (knot, row, slot)
// items (G_PAR_FOR) separated by barriers (G_SYNC), communicating only through shared/global memory.  The same
// source therefore compiles in two ways:
//   * nvcc, sm_100a            -> the product kernels (G_TID = threadIdx.x, G_SYNC = __syncthreads()).
//   * g++ -DGUSTO_HOSTSIM      -> a single-thread host simulation (G_TID = 0, G_NTHR = 1) used ONLY by
//                                 tests/ to check kernel logic in the GPU-less build container.  It is not
//                                 linked into libgusto_b200.so and is not a fallback.
#pragma once
#include <cstdint>
#include <cmath>

#ifdef GUSTO_HOSTSIM
#define GDEV inline
#define GDEV_NOINLINE inline
#define GHD inline
#define G_TID 0
#define G_NTHR 1
#define G_LANE 0
#define G_NLANE 1
#define G_SYNC() ((void)0)
#define G_SYNCWARP() ((void)0)
#else
#include <cuda_runtime.h>
#define GDEV __device__ __forceinline__
#define GDEV_NOINLINE __device__ __noinline__
#define GHD __host__ __device__ __forceinline__
#define G_TID ((int)threadIdx.x)
#define G_NTHR ((int)blockDim.x)
#define G_LANE ((int)(threadIdx.x & 31))
#define G_NLANE 32
#define G_SYNC() __syncthreads()
#define G_SYNCWARP() __syncwarp()
#endif

#define G_PAR_FOR(i, n) for (int i = G_TID; i < (n); i += G_NTHR)
// loops executed by the first warp only (callers guard with `if (G_TID < G_WARP)`), separated by G_SYNCWARP()
#ifdef GUSTO_HOSTSIM
#define G_WARP 1
#else
#define G_WARP 32
#endif
#define G_W0_FOR(i, n) for (int i = G_TID; i < (n); i += G_WARP)

// 8-byte asynchronous global->shared copy (LDGSTS): lets the block-tridiagonal sweeps prefetch the next block's factor
// while the current one is applied.  Host simulation: a plain copy.
#ifdef GUSTO_HOSTSIM
inline void g_cp_async8(double* dst, const double* src) { *dst = *src; }
inline void g_cp_async_wait() {}
#else
__device__ __forceinline__ void g_cp_async8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void g_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

namespace gusto {

enum ModelId : int { DUBINS = 0, FREEFLYER_SE2 = 1, ASTROBEE_SE3 = 2, ASTROBEE_SE3_MANIFOLD = 3 };
enum ObsKind : int { OBS_BOX = 0, OBS_SPHERE = 1 };
enum GoalType : int { GOAL_FREE = 0, GOAL_POINT = 1, GOAL_BOX = 2 };

// robot_params[16] slots (see include/gusto_b200.h)
enum : int {
  RP_MASS = 0, RP_JXX, RP_JYY, RP_JZZ, RP_RADIUS, RP_VMAX, RP_AMAX, RP_WMAX, RP_ALMAX, RP_CLEAR,
  RP_DUB_V, RP_DUB_K, RP_DUB_XMAX0, RP_DUB_XMAX1, RP_DUB_XMAX2, RP_DUB_UMAX
};
enum : int { SP_DELTA0 = 0, SP_OMEGA0, SP_OMEGAMAX, SP_EPS, SP_RHO0, SP_RHO1, SP_BSUCC, SP_BFAIL, SP_GFAIL, SP_CONVTHR };

constexpr int MAX_NX = 13;
constexpr int MAX_NU = 6;
constexpr int MAX_OBS = 64;

// Compile-time description of a dynamics model (mirrors the reference's per-model files, see models.cuh).
template <int M> struct Traits;
template <> struct Traits<DUBINS> {
  static constexpr int NX = 3, NU = 1, WS = 0, HAS_TR = 0, NNORM = 0, NLIN = 6, HAS_QUAT = 0, NBALL = 1;
};
template <> struct Traits<FREEFLYER_SE2> {
  static constexpr int NX = 6, NU = 3, WS = 2, HAS_TR = 1, NNORM = 2, NLIN = 0, HAS_QUAT = 0, NBALL = 2;
};
template <> struct Traits<ASTROBEE_SE3> {
  static constexpr int NX = 12, NU = 6, WS = 3, HAS_TR = 1, NNORM = 2, NLIN = 0, HAS_QUAT = 0, NBALL = 2;
};
template <> struct Traits<ASTROBEE_SE3_MANIFOLD> {
  static constexpr int NX = 13, NU = 6, WS = 3, HAS_TR = 0, NNORM = 2, NLIN = 1, HAS_QUAT = 1, NBALL = 2;
};

// Everything a kernel needs that is shared by the whole batch.  Passed by value (fits the 4 KB param space).
struct BatchDesc {
  int model_id, N, B, n_obs;
  double rp[16];                    // robot_params
  double sp[10];                    // scp_params
  int goal_type[MAX_NX];
  int obs_kind[MAX_OBS];
  double obs_a[MAX_OBS][3];         // box: lo          sphere: centre
  double obs_b[MAX_OBS][3];         // box: hi          sphere: (radius, -, -)
};

// Device-resident per-instance data (all knot-major, instance-major: [B][N][n], i.e. a Julia Array (n, N, B)).
struct BatchPtrs {
  const double* tf;        // [B]
  const double* x_init;    // [B][NX]
  const double* goal_lo;   // [B][NX]
  const double* goal_hi;   // [B][NX]
  double* Xp;  double* Up;     // accepted (previous) trajectory   [B][N][NX], [B][N][NU]
  double* Xn;  double* Un;     // candidate trajectory from the convex solve
  double* omega; double* delta;  // [B] current penalty weight / trust-region size
  // linearization blocks written by linearize_kernel and consumed in place by the solve / evaluate kernels
  double* f;       // [B][N][NX]
  double* A;       // [B][N][NX][NX] row-major: A[i][j] = d f_i / d x_j
  double* g;       // [B][N][NX]      f - A Xp - B Up  (affine part of the trapezoid row, halves summed by the consumer)
  const uint8_t* active;  // [B] instances still iterating (solve/evaluate skip the others); may be null
  double* rows;    // [B][N][n_obs][5] (nhat_x, nhat_y, nhat_z, off, dist0): row value = off - nhat.r, active iff dist0 < toggle
};

GHD double sq(double a) { return a * a; }

}  // namespace gusto
