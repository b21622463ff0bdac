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
// loop over the lanes of the executing warp (any warp)
#define G_LANE_FOR(i, n) for (int i = G_LANE; i < (n); i += G_NLANE)

// Asynchronous global->shared copies (LDGSTS): the block-tridiagonal sweeps prefetch the next block's factor while the
// current one is applied.  16-byte form needs 16-byte aligned source and destination.  Host simulation: plain copies.
#ifdef GUSTO_HOSTSIM
inline void g_cp_async8(double* dst, const double* src) { *dst = *src; }
inline void g_cp_async16(double* dst, const double* src) { dst[0] = src[0]; dst[1] = src[1]; }
inline void g_cp_async_commit() {}
inline void g_cp_async_wait() {}
template <int N> inline void g_cp_async_wait_group() {}
#else
__device__ __forceinline__ void g_cp_async8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void g_cp_async16(double* dst, const double* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void g_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void g_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N> __device__ __forceinline__ void g_cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// 16-byte shared/global accesses (LDS.128 / LDG.128) and the fast FP64 reciprocal (MUFU.RCP64H + Newton steps).
#ifdef GUSTO_HOSTSIM
struct g_d2 { double x, y; };
inline g_d2 g_ld2(const double* p) { g_d2 v; v.x = p[0]; v.y = p[1]; return v; }
inline void g_st2(double* p, double a, double b) { p[0] = a; p[1] = b; }
inline double g_rcp(double x) { return 1.0 / x; }
inline double g_rsqrt(double x) { return 1.0 / sqrt(x); }
#else
typedef double2 g_d2;
__device__ __forceinline__ g_d2 g_ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void g_st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
// Reciprocal: MUFU.RCP64H seed (relative error 2^-23) + two Newton steps = ~1 ulp, without the special-case path of
// __drcp_rn (callers pass finite, normal, non-zero values; anything else surfaces as NaN and the solve reports
// IPM_NUMERICAL).  Measured on the headline batch: solve 6.52 -> 5.98 ms (-DGUSTO_IEEE_RCP restores __drcp_rn).
#ifndef GUSTO_IEEE_RCP
__device__ __forceinline__ double g_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
#else
__device__ __forceinline__ double g_rcp(double x) { return __drcp_rn(x); }
#endif
__device__ __forceinline__ double g_rsqrt(double x) { return rsqrt(x); }
#endif

// TMA (cp.async.bulk) 1-D global -> shared copies completing on an mbarrier, and the mbarrier primitives they need.
// Host simulation: the copy is immediate and the barrier calls are no-ops.
#ifdef GUSTO_HOSTSIM
inline void g_mbar_init(unsigned long long*, unsigned) {}
inline void g_mbar_expect_tx(unsigned long long*, unsigned) {}
inline void g_tma_bulk_g2s(double* dst, const double* src, unsigned bytes, unsigned long long*) { for (unsigned i = 0; i < bytes / 8; ++i) dst[i] = src[i]; }
inline void g_mbar_wait(unsigned long long*, unsigned) {}
inline void g_fence_proxy_async() {}
inline void g_fence_proxy_async_global() {}
inline void g_prefetch_l1(const void*) {}
inline void g_prefetch_l2(const void*) {}
#else
__device__ __forceinline__ void g_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void g_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void g_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// generic-proxy GLOBAL writes that a later cp.async.bulk (async proxy) reads back need the unqualified cross-proxy fence
__device__ __forceinline__ void g_fence_proxy_async_global() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void g_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_tma_bulk_g2s(double* dst, const double* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void g_mbar_wait(unsigned long long* bar, unsigned parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity)
                 : "memory");
  }
}
#endif

// Warp-level FP64 tensor-core tile product (DMMA, mma.sync.m8n8k4.f64): one warp produces an 8 x 8 output tile
//   D[i0 + r][q0 + c] = sum_{m < 4 KS} X(i0 + r, m) * Y(q0 + c, m)          ("A B'" form; both operands are read the same way)
// with X(r, m) = X[r*ldx + m] (or X[m*ldx + r] when TA), likewise Y / TB.  Fragments come straight from shared memory, one
// 8-byte load per operand and k-step: lane l reads row (l >> 2), k-offset (l & 3) -- conflict-free when ld mod 16 is 4 or 12 --
// and ends up holding D[i0 + (l >> 2)][q0 + 2 (l & 3) + {0, 1}], handed to the epilogue `epi(row, col, value)`.
// One DMMA issues 256 multiply-adds in ONE instruction slot (8 DFMA slots otherwise): the solve kernel is issue-bound.
//   -DGUSTO_NO_DMMA : same thread -> element mapping with scalar DFMA (the measured comparison, profiles/r02_history.md)
//   host simulation : plain loops over the 64 elements.
#ifdef GUSTO_HOSTSIM
#define G_WARPID 0
#define G_NWARP 1
#else
#define G_WARPID ((int)(G_TID >> 5))
#define G_NWARP ((int)(G_NTHR >> 5))
__device__ __forceinline__ void g_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif
template <bool T> GDEV double g_tile_at(const double* X, int ld, int row, int m) { return T ? X[m * ld + row] : X[row * ld + m]; }
#ifndef GUSTO_HOSTSIM
template <int KS, bool TA, bool TB>
__device__ __forceinline__ void g_tile_acc(double& c0, double& c1, const double* X, int ldx, int i0, const double* Y, int ldy, int q0) {
  const int g = G_LANE >> 2, t = G_LANE & 3;
#ifndef GUSTO_NO_DMMA
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) g_dmma(c0, c1, g_tile_at<TA>(X, ldx, i0 + g, 4 * ks + t), g_tile_at<TB>(Y, ldy, q0 + g, 4 * ks + t));
#else
#pragma unroll
  for (int m = 0; m < 4 * KS; ++m) {
    const double xa = g_tile_at<TA>(X, ldx, i0 + g, m);
    c0 = fma(xa, g_tile_at<TB>(Y, ldy, q0 + 2 * t, m), c0);
    c1 = fma(xa, g_tile_at<TB>(Y, ldy, q0 + 2 * t + 1, m), c1);
  }
#endif
}
#endif
template <int KS, bool TA, bool TB, typename EPI>
GDEV void g_tile_job(const double* X, int ldx, int i0, const double* Y, int ldy, int q0, EPI&& epi) {
#ifdef GUSTO_HOSTSIM
  for (int r = 0; r < 8; ++r)
    for (int c2 = 0; c2 < 8; ++c2) {
      double s = 0.0;
      for (int m = 0; m < 4 * KS; ++m) s += g_tile_at<TA>(X, ldx, i0 + r, m) * g_tile_at<TB>(Y, ldy, q0 + c2, m);
      epi(i0 + r, q0 + c2, s);
    }
#else
  double c0 = 0.0, c1 = 0.0;
  g_tile_acc<KS, TA, TB>(c0, c1, X, ldx, i0, Y, ldy, q0);
  const int g = G_LANE >> 2, t = G_LANE & 3;
  epi(i0 + g, q0 + 2 * t, c0);
  epi(i0 + g, q0 + 2 * t + 1, c1);
#endif
}
// MT x NT tiles at once (rows i0 + 8 mi, columns q0 + 8 ni): the k-steps of the tiles are interleaved, so MT * NT independent
// DMMA chains are in flight (one DMMA has ~150 cycles of latency here) and every fragment is loaded once per k-step.
template <int KS, int MT, int NT, bool TA, bool TB, typename EPI>
GDEV void g_tile_grid(const double* X, int ldx, int i0, const double* Y, int ldy, int q0, EPI&& epi) {
#ifdef GUSTO_HOSTSIM
  for (int mi = 0; mi < MT; ++mi)
    for (int ni = 0; ni < NT; ++ni) g_tile_job<KS, TA, TB>(X, ldx, i0 + 8 * mi, Y, ldy, q0 + 8 * ni, epi);
#else
  const int g = G_LANE >> 2, t = G_LANE & 3;
  double acc[MT][NT][2];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
#ifndef GUSTO_NO_DMMA
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    double a[MT], b[NT];
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) a[mi] = g_tile_at<TA>(X, ldx, i0 + 8 * mi + g, 4 * ks + t);
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) b[ni] = g_tile_at<TB>(Y, ldy, q0 + 8 * ni + g, 4 * ks + t);
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) g_dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
#else
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) g_tile_acc<KS, TA, TB>(acc[mi][ni][0], acc[mi][ni][1], X, ldx, i0 + 8 * mi, Y, ldy, q0 + 8 * ni);
#endif
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      epi(i0 + 8 * mi + g, q0 + 8 * ni + 2 * t, acc[mi][ni][0]);
      epi(i0 + 8 * mi + g, q0 + 8 * ni + 2 * t + 1, acc[mi][ni][1]);
    }
#endif
}
// ... the sum of two such products (second operand pair X2 / Y2 with its own k-length and orientation)
template <int KS, bool TA, bool TB, int KS2, bool TA2, bool TB2, typename EPI>
GDEV void g_tile_job2(const double* X, int ldx, const double* Y, int ldy, const double* X2, int ldx2, const double* Y2, int ldy2,
                      int i0, int q0, EPI&& epi) {
#ifdef GUSTO_HOSTSIM
  for (int r = 0; r < 8; ++r)
    for (int c2 = 0; c2 < 8; ++c2) {
      double s = 0.0;
      for (int m = 0; m < 4 * KS; ++m) s += g_tile_at<TA>(X, ldx, i0 + r, m) * g_tile_at<TB>(Y, ldy, q0 + c2, m);
      for (int m = 0; m < 4 * KS2; ++m) s += g_tile_at<TA2>(X2, ldx2, i0 + r, m) * g_tile_at<TB2>(Y2, ldy2, q0 + c2, m);
      epi(i0 + r, q0 + c2, s);
    }
#else
  double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;               // two independent chains
  g_tile_acc<KS, TA, TB>(c0, c1, X, ldx, i0, Y, ldy, q0);
  g_tile_acc<KS2, TA2, TB2>(d0, d1, X2, ldx2, i0, Y2, ldy2, q0);
  const int g = G_LANE >> 2, t = G_LANE & 3;
  epi(i0 + g, q0 + 2 * t, c0 + d0);
  epi(i0 + g, q0 + 2 * t + 1, c1 + d1);
#endif
}

// Address-space hint: pointers that travel through the per-instance context struct come back as generic pointers;
// telling the compiler they are shared turns LD.E/ST.E (64-bit generic path) back into LDS/STS.
#ifdef GUSTO_HOSTSIM
#define G_ASSUME_SHARED(p) ((void)0)
#else
#define G_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
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
//   NX, NU, WS (workspace dims of the obstacle rows), which soft/hard row families exist (SCPConstraints(SCPP) of the
//   model file), and the STRUCTURE the convex solve exploits:
//   * x-blocks / u-blocks: a partition of the state / control coordinates such that every inequality row except the
//     state trust region has its gradient inside one block (position | velocity | attitude | rate).  The reduced
//     Hessian of a knot is then  dg*I + blockdiag(blocks) + kappa_tr * g g'  (Sherman-Morrison invertible).
//   * a_row/a_col: the static sparsity pattern of A = df/dx (entries that can be non-zero), ANZ entries.
//   * b_row(a): the single state row driven by control a (B = df/du has one entry per column).
//   * DSPLIT: the dynamics decouple into the coordinates [0, DSPLIT) and [DSPLIT, NX) (translation | rotation of the Astrobee
//     models; 0 = no split): A, B and hence F^-1, Ah, Bh, Gam of the Riccati solve are block diagonal in that partition.
#define GUSTO_BLOCKS(NAME, CNT, O0, N0, O1, N1, O2, N2, O3, N3)                                            \
  static constexpr int NAME##_CNT = CNT;                                                                    \
  GHD static constexpr int NAME##_off(int b) { return b == 0 ? O0 : b == 1 ? O1 : b == 2 ? O2 : O3; }      \
  GHD static constexpr int NAME##_n(int b) { return b == 0 ? N0 : b == 1 ? N1 : b == 2 ? N2 : N3; }        \
  GHD static constexpr int NAME##_pk(int b) {   /* offset of block b in the packed-lower storage */        \
    int o = 0;                                                                                              \
    for (int q = 0; q < b; ++q) o += NAME##_n(q) * (NAME##_n(q) + 1) / 2;                                   \
    return o;                                                                                               \
  }                                                                                                         \
  GHD static constexpr int NAME##_of(int i) {   /* block containing coordinate i */                        \
    int b = 0;                                                                                              \
    for (int q = 0; q < CNT; ++q) if (i >= NAME##_off(q)) b = q;                                            \
    return b;                                                                                               \
  }

template <int M> struct Traits;
template <> struct Traits<DUBINS> {
  static constexpr int NX = 3, NU = 1, WS = 0, HAS_TR = 0, NNORM = 0, NLIN = 6, HAS_QUAT = 0, NBALL = 1, DSPLIT = 0;
  GUSTO_BLOCKS(XB, 3, 0, 1, 1, 1, 2, 1, 0, 0)
  GUSTO_BLOCKS(UB, 1, 0, 1, 0, 0, 0, 0, 0, 0)
  static constexpr int ANZ = 2;
  GHD static constexpr int a_row(int e) { return e; }
  GHD static constexpr int a_col(int) { return 2; }
  GHD static constexpr int b_row(int) { return 2; }
};
template <> struct Traits<FREEFLYER_SE2> {
  static constexpr int NX = 6, NU = 3, WS = 2, HAS_TR = 1, NNORM = 2, NLIN = 0, HAS_QUAT = 0, NBALL = 2, DSPLIT = 0;
  GUSTO_BLOCKS(XB, 4, 0, 2, 2, 1, 3, 2, 5, 1)
  GUSTO_BLOCKS(UB, 2, 0, 2, 2, 1, 0, 0, 0, 0)
  static constexpr int ANZ = 3;
  GHD static constexpr int a_row(int e) { return e; }
  GHD static constexpr int a_col(int e) { return 3 + e; }
  GHD static constexpr int b_row(int a) { return 3 + a; }
};
template <> struct Traits<ASTROBEE_SE3> {
  static constexpr int NX = 12, NU = 6, WS = 3, HAS_TR = 1, NNORM = 2, NLIN = 0, HAS_QUAT = 0, NBALL = 2, DSPLIT = 6;
  GUSTO_BLOCKS(XB, 4, 0, 3, 3, 3, 6, 3, 9, 3)
  GUSTO_BLOCKS(UB, 2, 0, 3, 3, 3, 0, 0, 0, 0)
  // rows 0-2: identity on v | rows 6-8: MRP kinematics wrt (p, w) | rows 9-11: gyroscopic wrt w (off-diagonal)
  static constexpr int ANZ = 27;
  GHD static constexpr int a_row(int e) { return e < 3 ? e : e < 21 ? 6 + (e - 3) / 6 : 9 + (e - 21) / 2; }
  GHD static constexpr int a_col(int e) {
    if (e < 3) return 3 + e;
    if (e < 21) return 6 + (e - 3) % 6;
    const int r = (e - 21) / 2, q = (e - 21) % 2;       // row 9+r: the two columns of {9,10,11} other than 9+r
    return 9 + (q + (q >= r ? 1 : 0));
  }
  GHD static constexpr int b_row(int a) { return a < 3 ? 3 + a : 6 + a; }
};
template <> struct Traits<ASTROBEE_SE3_MANIFOLD> {
  static constexpr int NX = 13, NU = 6, WS = 3, HAS_TR = 0, NNORM = 2, NLIN = 1, HAS_QUAT = 1, NBALL = 2, DSPLIT = 6;
  GUSTO_BLOCKS(XB, 4, 0, 3, 3, 3, 6, 4, 10, 3)
  GUSTO_BLOCKS(UB, 2, 0, 3, 3, 3, 0, 0, 0, 0)
  // rows 0-2: identity on v | rows 6-9: quaternion kinematics wrt (q, w) minus the zero diagonal | rows 10-12: gyroscopic
  static constexpr int ANZ = 33;
  GHD static constexpr int a_row(int e) { return e < 3 ? e : e < 27 ? 6 + (e - 3) / 6 : 10 + (e - 27) / 2; }
  GHD static constexpr int a_col(int e) {
    if (e < 3) return 3 + e;
    if (e < 27) { const int r = (e - 3) / 6, q = (e - 3) % 6; return 6 + (q + (q >= r ? 1 : 0)); }
    const int r = (e - 27) / 2, q = (e - 27) % 2;
    return 10 + (q + (q >= r ? 1 : 0));
  }
  GHD static constexpr int b_row(int a) { return a < 3 ? 3 + a : 7 + a; }
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
  // linearization blocks written by linearize_kernel and consumed in place by the solve / evaluate kernels, knot-minor per
  // instance (NP = g_np(N) doubles per field row; the consumers run one thread per knot and read every field coalesced):
  double* f;       // [B][NX][NP]
  double* A;       // [B][ANZ][NP]    A = d f / d x on its static sparsity pattern (Traits<M>::a_row / a_col), e.g. 27 of 144 entries
  double* g;       // [B][NX][NP]     f - A Xp - B Up  (affine part of the trapezoid row, halves summed by the consumer)
  const uint8_t* active;  // [B] instances still iterating (solve/evaluate skip the others); may be null
  double* rows;    // [B][5][n_obs][NP] (nhat_x, nhat_y, nhat_z, off, dist0): row value = off - nhat.r, active iff dist0 < toggle
  double* dual;    // [B][NX] multiplier of the init rows X[:,1] = x_init of the last solve (= -JuMP.dual, get_dual_jump); may be null
};

GHD double sq(double a) { return a * a; }
// knots per field row of the knot-minor block / scratch layouts: N rounded up to 4 (every row starts on a 32-byte sector)
GHD constexpr int g_np(int N) { return (N + 3) & ~3; }

}  // namespace gusto
