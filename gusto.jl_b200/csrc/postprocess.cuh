// K5 / K6: trajectory post-processing, one CTA per instance.
//
// Reference (astrobee_se3.jl; the other models carry the same functions):
//   dynamics_constraint_satisfaction :529-540   J = sum_{k=1}^{N-1} | (X_{k+1} - X_k)/dt - f(X_k, U_k) |_1
//   verify_collision_free            :542-562   first (env_idx, k) -- obstacles outer, knots inner -- with dist < 0
//   interpolate_traj                 :495-527   RK4 upsampling, Nstep = ceil(dt/dt_min) sub-steps per knot interval
//                                               under the zero-order-hold control U_k; every knot is re-seeded exactly
//                                               (stale repmat / Matrix(u,n) calls of the reference restated)
// plus the quantity the parity reports need: the NONLINEAR trapezoid defect  X_{k+1} - X_k - h/2 (f_k + f_{k+1})
// (dynamics_constraints :151-165 evaluated at the trajectory itself).
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "sdf.cuh"
#include "evaluate.cuh"

namespace gusto {

constexpr int CHECK_NOUT = 8;   // J_dyn, max |trapezoid defect|, collision_free, k_first, obstacle_first, dist_first, min dist, max |u-ball| ratio

template <int M>
GDEV void check_instance(const BatchDesc& d, const BatchPtrs& p, int b, const double* X, const double* U, double* out, double* red) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU;
  const int N = d.N;
  const double h = p.tf[b] / (N - 1);
  double J = 0.0, dmax = 0.0, ball = 0.0;
  G_PAR_FOR(k, N - 1) {
    double f0[NX], f1[NX];
    dyn_f<M>(X + k * NX, U + k * NU, d.rp, f0);
    dyn_f<M>(X + (k + 1) * NX, U + (k + 1) * NU, d.rp, f1);
    for (int i = 0; i < NX; ++i) {
      const double dx = X[(k + 1) * NX + i] - X[k * NX + i];
      J += fabs(dx / h - f0[i]);
      const double t = fabs(dx - 0.5 * h * (f0[i] + f1[i]));
      dmax = t > dmax ? t : dmax;
    }
    for (int j = 0; j < T::NBALL; ++j) {          // hard control balls, k = 1..N-1 (quirk q3): |scale .* u| / rad
      int i0, i1; double scale[3], rad;
      ctrl_ball<M>(j, d.rp, &i0, &i1, scale, &rad);
      double s = 0.0;
      for (int i = i0; i < i1; ++i) s += sq(scale[i - i0] * U[k * NU + i]);
      const double r = sqrt(s) / rad;
      ball = r > ball ? r : ball;
    }
  }
  double key = 1e300, dmin = 1e300;
  if (T::WS > 0) {
    constexpr int WS = T::WS > 0 ? T::WS : 1;
    G_PAR_FOR(it, N * d.n_obs) {
      const int i = it / N, k = it - i * N;        // obstacles outer, knots inner: the reference's loop order
      double r[3], dist, nh[3];
      workspace_location<WS>(X + k * NX, r);
      signed_distance<WS>(r, d.obs_kind[i], d.obs_a[i], d.obs_b[i], d.rp[RP_RADIUS], &dist, nh);
      if (dist < 0.0 && (double)it < key) key = (double)it;
      dmin = dist < dmin ? dist : dmin;
    }
  }
  J = block_sum(J, red);
  dmax = block_max(dmax, red);
  ball = block_max(ball, red);
  key = -block_max(-key, red);
  dmin = -block_max(-dmin, red);
  if (G_TID == 0) {
    out[0] = J; out[1] = dmax; out[7] = ball;
    out[6] = (T::WS > 0 && d.n_obs > 0) ? dmin : 0.0;
    if (key < 1e299) {
      constexpr int WS = T::WS > 0 ? T::WS : 1;
      const int it = (int)key, i = it / N, k = it - i * N;
      double r[3], dist, nh[3];
      workspace_location<WS>(X + k * NX, r);
      signed_distance<WS>(r, d.obs_kind[i], d.obs_a[i], d.obs_b[i], d.rp[RP_RADIUS], &dist, nh);
      out[2] = 0.0; out[3] = (double)k; out[4] = (double)i; out[5] = dist;
    } else {
      out[2] = 1.0; out[3] = -1.0; out[4] = -1.0; out[5] = 0.0;
    }
  }
}

// One knot interval k of instance b: nstep RK4 sub-steps from X_k under the held control U_k.
//   Xfull [nstep*(N-1)+1][NX], Ufull [nstep*(N-1)][NU]
template <int M>
GDEV void interpolate_interval(const BatchDesc& d, const BatchPtrs& p, int b, int k, int nstep, const double* X, const double* U,
                               double* Xfull, double* Ufull) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU;
  const int N = d.N;
  const double dt = p.tf[b] / (N - 1) / nstep;
  double x[NX], u[NU];
  for (int i = 0; i < NX; ++i) x[i] = X[k * NX + i];
  for (int i = 0; i < NU; ++i) u[i] = U[k * NU + i];
  const int i0 = nstep * k;
  for (int s = 0; s < nstep; ++s) {
    for (int i = 0; i < NX; ++i) Xfull[(size_t)(i0 + s) * NX + i] = x[i];
    for (int i = 0; i < NU; ++i) Ufull[(size_t)(i0 + s) * NU + i] = u[i];
    double k1[NX], k2[NX], k3[NX], k4[NX], xt[NX];
    dyn_f<M>(x, u, d.rp, k1);
    for (int i = 0; i < NX; ++i) xt[i] = x[i] + 0.5 * dt * k1[i];
    dyn_f<M>(xt, u, d.rp, k2);
    for (int i = 0; i < NX; ++i) xt[i] = x[i] + 0.5 * dt * k2[i];
    dyn_f<M>(xt, u, d.rp, k3);
    for (int i = 0; i < NX; ++i) xt[i] = x[i] + dt * k3[i];
    dyn_f<M>(xt, u, d.rp, k4);
    for (int i = 0; i < NX; ++i) x[i] += dt / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
  }
  if (k == N - 2)                                   // Xfull[:, end] = X[:, end]
    for (int i = 0; i < NX; ++i) Xfull[(size_t)(nstep * (N - 1)) * NX + i] = X[(N - 1) * NX + i];
}

}  // namespace gusto
