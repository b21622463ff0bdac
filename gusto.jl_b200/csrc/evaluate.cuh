// K4: per-instance evaluation scalars of one SCP iteration (one CTA per instance).
//
// Reference: convergence_metric traj_opt.jl:74-85; trust_region_satisfied_gusto scp_gusto.jl:34-44;
// convex_ineq_satisfied_gusto_jump scp_gusto.jl:316-343; trust_region_ratio_gusto astrobee_se3.jl:383-417
// (manifold :610-642, freeflyer freeflyer_se2.jl:392-427, dubins dubins_car.jl:229-241); cost_true
// astrobee_se3_manifold.jl:73-100; penalized objective scp_gusto.jl:253-314.  Quirks kept: rho's
// linear model omits B(U-Up) and sums over ALL obstacles (q2); freeflyer loops over both hulls of the compound
// robot at the body translation (q11); the soft rows are checked on the convexified rows (q8).
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "sdf.cuh"

namespace gusto {

constexpr int EVAL_NOUT = 8;   // conv, tr_ok, ineq_ok, rho, J_true, J_full, max_k |dX_k|^2, max soft row value

// CTA-wide reductions.  GPU: warp shuffles, then one shared-memory slot per warp (deterministic order; not inlined --
// they are called from dozens of places and the kernel is instruction-cache bound).  Host simulation: identity.
#ifdef GUSTO_HOSTSIM
inline double block_sum(double v, double* red) { (void)red; return v; }
inline double block_max(double v, double* red) { (void)red; return (v == v) ? v : 1e300; }   // NaN-propagating
#else
__device__ __noinline__ double block_sum(double v, double* red) {
  G_ASSUME_SHARED(red);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < nw; ++i) s += red[i];
  __syncthreads();
  return s;
}
__device__ __noinline__ double block_max(double v, double* red) {   // NaN-propagating: a NaN anywhere yields +huge
  G_ASSUME_SHARED(red);
  v = (v == v) ? v : 1e300;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double s = red[0];
  for (int i = 1; i < nw; ++i) s = red[i] > s ? red[i] : s;
  __syncthreads();
  return s;
}
#endif

// X,U: candidate trajectory of instance b; Xp: previous.  red: G_NTHR doubles of shared memory.
template <int M>
GDEV void evaluate_instance(const BatchDesc& d, const BatchPtrs& p, int b, const double* X, const double* U,
                            double* out, double* red) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU;
  const int N = d.N;
  const double h = p.tf[b] / (N - 1);
  const double omega = p.omega[b], Delta = p.delta[b];
  const double eps = d.sp[SP_EPS], cl = d.rp[RP_CLEAR], R = d.rp[RP_RADIUS];
  const double toggle = Delta / 8.0 + cl;                       // scp_gusto.jl:76,156
  const double* Xp = p.Xp + (size_t)b * N * NX;
  const size_t np = g_np(N), fs = (size_t)d.n_obs * np;
  const double* F = p.f + (size_t)b * NX * np;
  const double* A = p.A + (size_t)b * T::ANZ * np;                 // sparsity pattern of A, knot-minor
  const double* rows = p.rows + (size_t)b * 5 * fs;

  double num = 0, den = 0, Jt = 0, Jpen = 0, mdx2 = 0, mnx2 = 0, msoft = -1e300, meq = 0;
  G_PAR_FOR(k, N) {
    const double* x = X + k * NX;
    const double* u = U + k * NU;
    const double* xp = Xp + k * NX;
    double dx[NX], dx2 = 0, nx2 = 0, uu = 0;
    for (int i = 0; i < NX; ++i) { dx[i] = x[i] - xp[i]; dx2 += dx[i] * dx[i]; nx2 += x[i] * x[i]; }
    for (int i = 0; i < NU; ++i) uu += u[i] * u[i];
    mdx2 = dx2 > mdx2 ? dx2 : mdx2;
    mnx2 = nx2 > mnx2 ? nx2 : mnx2;
    Jt += ((k == 0 || k == N - 1) ? 0.5 * h : h) * uu;
    if (T::HAS_TR) { const double v = omega * dx2 - Delta; Jpen += v > 0 ? v : 0; }
    if (k < N - 1) {
      double fn[NX], e2 = 0, l2 = 0;
      dyn_f<M>(x, u, d.rp, fn);
      double lin[NX];
      for (int i = 0; i < NX; ++i) lin[i] = F[i * np + k];
#pragma unroll
      for (int e = 0; e < T::ANZ; ++e) lin[T::a_row(e)] += A[e * np + k] * dx[T::a_col(e)];
      for (int i = 0; i < NX; ++i) { e2 += sq(fn[i] - lin[i]); l2 += lin[i] * lin[i]; }
      num += sqrt(e2);
      den += sqrt(l2);
    }
    for (int j = 0; j < T::NNORM; ++j) {
      int i0, i1; double lim;
      norm_row<M>(j, d.rp, &i0, &i1, &lim);
      double v = -lim * lim;
      for (int i = i0; i < i1; ++i) v += x[i] * x[i];
      msoft = v > msoft ? v : msoft;
      Jpen += omega * v > 0 ? omega * v : 0;
    }
    for (int j = 0; j < T::NLIN; ++j) {
      int i; double sign, bound;
      lin_row<M>(j, d.rp, &i, &sign, &bound);
      const double v = sign * x[i] - bound;
      msoft = v > msoft ? v : msoft;
      Jpen += omega * v > 0 ? omega * v : 0;
    }
    if (T::HAS_QUAT) {
      const double* qp = xp + 6;
      const double* q = x + 6;
      const double nq = sqrt(qp[0] * qp[0] + qp[1] * qp[1] + qp[2] * qp[2] + qp[3] * qp[3]);
      double e = nq - 1.0;
      for (int i = 0; i < 4; ++i) e += qp[i] * (q[i] - qp[i]) / nq;
      const double ae = fabs(e);
      meq = ae > meq ? ae : meq;
      Jpen += omega * e - eps > 0 ? omega * e - eps : 0;
    }
  }
  if (T::WS > 0) {
    constexpr int WS = T::WS > 0 ? T::WS : 1;
    G_PAR_FOR(it, N * d.n_obs) {
      const int i = it / N, k = it - i * N;                         // knot fastest: coalesced row reads
      double r[3], r0[3];
      workspace_location<WS>(X + k * NX, r);                        // (the instance's trajectories are L1-resident by now)
      workspace_location<WS>(Xp + k * NX, r0);
      const double* row = rows + (size_t)i * np + k;
      const double linr = row[3 * fs] - (row[0] * r[0] + row[fs] * r[1] + row[2 * fs] * r[2]);
      double d1, n1[3];
      signed_distance<WS>(r, d.obs_kind[i], d.obs_a[i], d.obs_b[i], R, &d1, n1);
      num += fabs((cl - d1) - linr);
      den += fabs(linr);
      if (M == FREEFLYER_SE2) {       // second hull of the compound robot, offset xb = (0, 0.15, 0) (robot/freeflyer.jl:48)
        double ra[3] = {r[0], r[1] + 0.15, r[2]}, ra0[3] = {r0[0], r0[1] + 0.15, r0[2]}, d0, n0[3];
        signed_distance<WS>(ra0, d.obs_kind[i], d.obs_a[i], d.obs_b[i], R, &d0, n0);
        const double lina = cl - (d0 + n0[0] * (r[0] - r0[0]) + n0[1] * (r[1] - r0[1]) + n0[2] * (r[2] - r0[2]));
        signed_distance<WS>(ra, d.obs_kind[i], d.obs_a[i], d.obs_b[i], R, &d1, n1);
        num += fabs((cl - d1) - lina);
        den += fabs(lina);
      }
      if (row[4 * fs] < toggle) {      // convexified row is live (astrobee_se3.jl:293)
        msoft = linr > msoft ? linr : msoft;
        Jpen += omega * linr > 0 ? omega * linr : 0;
      }
    }
  }
  num = block_sum(num, red);
  den = block_sum(den, red);
  Jt = block_sum(Jt, red);
  Jpen = block_sum(Jpen, red);
  mdx2 = block_max(mdx2, red);
  mnx2 = block_max(mnx2, red);
  msoft = block_max(msoft, red);
  meq = block_max(meq, red);
  if (G_TID == 0) {
    out[0] = sqrt(mdx2) / sqrt(mnx2);
    out[1] = (mdx2 - Delta <= 0) ? 1.0 : 0.0;
    out[2] = (msoft < eps && meq < eps) ? 1.0 : 0.0;
    out[3] = num / den;
    out[4] = Jt;
    out[5] = Jt + Jpen;
    out[6] = mdx2;
    out[7] = msoft;
  }
}

}  // namespace gusto
