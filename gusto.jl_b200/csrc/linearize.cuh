// K1+K2: per-knot dynamics linearization and convexified obstacle rows, one warp per knot.
//
// Replaces update_model_params! (dynamics/astrobee_se3.jl:130-138: update_f!, update_A! at every knot), the
// affine part of dynamics_constraints (:151-165) and ncsi_obstacle_avoidance_constraints_convexified
// (:282-305, one BulletCollision.distance query per (knot, obstacle)).  Emits, per knot, the blocks the convex
// solve and the evaluation kernel consume in place:
//   f[NX], A[NX][NX], g[NX] = f - A x - B u            (trapezoid row k: h/2 (g_{k-1} + g_k) is its constant)
//   rows[n_obs][5] = (nhat, off = clearance - dist0 + nhat.r0, dist0)
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
template <int M>
GDEV void linearize_knot(const BatchDesc& d, const BatchPtrs& p, int b, int k, const double* x, const double* u,
                         double* ws, const double* bv) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU;
  double* sA = ws;             // [NX*NX]
  double* sf = ws + NX * NX;   // [NX]
  const int lane = G_LANE;
  const size_t gk = (size_t)b * d.N + k;

  for (int i = lane; i < NX * NX; i += G_NLANE) sA[i] = 0.0;
  G_SYNCWARP();
  if (lane == 0) {
    dyn_f<M>(x, u, d.rp, sf);
    dyn_A<M>(x, d.rp, sA);
  }
  G_SYNCWARP();
  // coalesced stores of A and f
  double* gA = p.A + gk * (NX * NX);
  for (int i = lane; i < NX * NX; i += G_NLANE) gA[i] = sA[i];
  double* gf = p.f + gk * NX;
  double* gg = p.g + gk * NX;
  for (int i = lane; i < NX; i += G_NLANE) {
    double acc = sf[i];
    for (int j = 0; j < NX; ++j) acc -= sA[i * NX + j] * x[j];
    // B is constant with one entry per column (B[b_row(a)][a] = bv[a], Traits<M>::b_row): no matrix is rebuilt
#pragma unroll
    for (int a = 0; a < NU; ++a) if (T::b_row(a) == i) acc -= bv[a] * u[a];
    gf[i] = sf[i];
    gg[i] = acc;
  }
  // obstacle rows: one lane per collision component
  if (T::WS > 0) {
    double r0[3];
    workspace_location<T::WS>(x, r0);
    double* grow = p.rows + gk * (size_t)d.n_obs * 5;
    for (int i = lane; i < d.n_obs; i += G_NLANE) {
      double dist, nh[3];
      signed_distance<(T::WS > 0 ? T::WS : 1)>(r0, d.obs_kind[i], d.obs_a[i], d.obs_b[i], d.rp[RP_RADIUS], &dist, nh);
      grow[i * 5 + 0] = nh[0];
      grow[i * 5 + 1] = nh[1];
      grow[i * 5 + 2] = nh[2];
      grow[i * 5 + 3] = d.rp[RP_CLEAR] - dist + nh[0] * r0[0] + nh[1] * r0[1] + nh[2] * r0[2];
      grow[i * 5 + 4] = dist;
    }
  }
  G_SYNCWARP();
}

}  // namespace gusto
