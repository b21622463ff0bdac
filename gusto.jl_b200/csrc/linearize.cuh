// K1+K2: per-knot dynamics linearization and convexified obstacle rows, one warp per knot.
//
// Replaces update_model_params! (dynamics/astrobee_se3.jl:130-138: update_f!, update_A! at every knot), the
// affine part of dynamics_constraints (:151-165) and ncsi_obstacle_avoidance_constraints_convexified
// (:282-305, one BulletCollision.distance query per (knot, obstacle)).  Emits, per knot, the blocks the convex
// solve and the evaluation kernel consume in place:
//   f[NX], A on its sparsity pattern [ANZ], g[NX] = f - A x - B u   (trapezoid row k: h/2 (g_{k-1} + g_k) is its constant)
//   rows[5][n_obs] = (nhat, off = clearance - dist0 + nhat.r0, dist0)
// all knot-minor per instance (BatchPtrs, common.cuh): the 8 warps of a CTA hold 8 consecutive knots, so every field row
// receives 64 contiguous bytes per CTA.
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "sdf.cuh"

namespace gusto {

// ws: per-warp scratch of NX*NX + NX doubles (shared memory on the GPU).  bv[NU]: the non-zero entry of each column of B.
template <int M> GDEV void dyn_B_columns(const double* rp, double* bv) {
  using T = Traits<M>;
  double Bm[T::NX * T::NU];
  for (int i = 0; i < T::NX * T::NU; ++i) Bm[i] = 0.0;
  dyn_B<M>(rp, Bm);
  for (int a = 0; a < T::NU; ++a) bv[a] = Bm[T::b_row(a) * T::NU + a];
}
// One knot in three parts, so that the kernel can run the serial part of all the knots of a CTA on the lanes of ONE warp:
//   linearize_fA    (one thread)   f and the dense A into the knot's scratch ws[NX*NX + NX] (zero-filled by the caller)
//   linearize_emit  (one warp)     A on its sparsity pattern, f, g = f - A x - B u
//   linearize_rows  (one warp)     obstacle rows, one lane per collision component (needs the state only)
// Output addressing: field e of A at oA[e * es], of f / g at of[i * es] / og[i * es], field q of obstacle row i at
// orows[(q * n_obs + i) * es].  The kernel stages the 8 knots of a CTA in shared memory (es = 8) and writes every field row of the
// knot-minor global layout as 64 contiguous bytes; the host simulation writes the global layout directly (es = NP).
// pat[ANZ]: a_row(e) * NX + a_col(e) (the kernel keeps it in shared memory; evaluating the pattern functions per entry and knot
// was 17 % of the kernel's instructions).
template <int M> GDEV void linearize_fA(const BatchDesc& d, const double* x, const double* u, double* ws) {
  using T = Traits<M>;
  dyn_f<M>(x, u, d.rp, ws + T::NX * T::NX);
  dyn_A<M>(x, d.rp, ws);
}
template <int M>
GDEV void linearize_emit(const double* x, const double* u, const double* ws, const double* bv, const unsigned char* pat,
                         double* oA, double* of, double* og, size_t es) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU;
  const double* sA = ws;             // [NX*NX]
  const double* sf = ws + NX * NX;   // [NX]
  const int lane = G_LANE;
  for (int e = lane; e < T::ANZ; e += G_NLANE) oA[e * es] = sA[pat[e]];
  for (int i = lane; i < NX; i += G_NLANE) {
    double acc = sf[i];
    for (int j = 0; j < NX; ++j) acc -= sA[i * NX + j] * x[j];
    // B is constant with one entry per column (B[b_row(a)][a] = bv[a], Traits<M>::b_row): no matrix is rebuilt
#pragma unroll
    for (int a = 0; a < NU; ++a) if (T::b_row(a) == i) acc -= bv[a] * u[a];
    of[i * es] = sf[i];
    og[i * es] = acc;
  }
}
template <int M> GDEV void linearize_rows(const BatchDesc& d, const double* x, double* orows, size_t es) {
  using T = Traits<M>;
  if (T::WS > 0) {
    const int lane = G_LANE;
    double r0[3];
    workspace_location<T::WS>(x, r0);
    const size_t fs = (size_t)d.n_obs * es;                                  // field stride
    for (int i = lane; i < d.n_obs; i += G_NLANE) {
      double dist, nh[3];
      signed_distance<(T::WS > 0 ? T::WS : 1)>(r0, d.obs_kind[i], d.obs_a[i], d.obs_b[i], d.rp[RP_RADIUS], &dist, nh);
      double* o = orows + i * es;
      o[0] = nh[0];
      o[fs] = nh[1];
      o[2 * fs] = nh[2];
      o[3 * fs] = d.rp[RP_CLEAR] - dist + nh[0] * r0[0] + nh[1] * r0[1] + nh[2] * r0[2];
      o[4 * fs] = dist;
    }
  }
}

// the knot-minor global layout written directly, one knot after the other (host simulation; BatchPtrs, common.cuh)
template <int M>
GDEV void linearize_knot_global(const BatchDesc& d, const BatchPtrs& p, int b, int k, const double* x, const double* u, double* ws, const double* bv) {
  using T = Traits<M>;
  const size_t np = g_np(d.N);
  unsigned char pat[T::ANZ];
  for (int e = 0; e < T::ANZ; ++e) pat[e] = (unsigned char)(T::a_row(e) * T::NX + T::a_col(e));
  for (int i = 0; i < T::NX * T::NX + T::NX; ++i) ws[i] = 0.0;
  linearize_fA<M>(d, x, u, ws);
  linearize_emit<M>(x, u, ws, bv, pat, p.A + (size_t)b * T::ANZ * np + k, p.f + (size_t)b * T::NX * np + k, p.g + (size_t)b * T::NX * np + k, np);
  linearize_rows<M>(d, x, p.rows + (size_t)b * 5 * d.n_obs * np + k, np);
}

}  // namespace gusto
