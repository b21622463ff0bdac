// Per-model continuous dynamics f(x,u), Jacobian A = df/dx and constant B = df/du, restated as device code.
// Reference (file:line under /root/reference/src):
//   astrobeeSE3          dynamics/astrobee_se3.jl:173-241, utils/quat_functions.jl:253-256 (mrp_derivative)
//   astrobeeSE3manifold  dynamics/astrobee_se3_manifold.jl:224-304
//   freeflyerSE2         dynamics/freeflyer_se2.jl:182-206
//   dubins               dynamics/dubins_car.jl:154-181
// Matrices are row-major: A[i*NX + j] = d f_i / d x_j.
#pragma once
#include "common.cuh"

namespace gusto {

template <int M> GDEV void dyn_f(const double* x, const double* u, const double* rp, double* f);
template <int M> GDEV void dyn_A(const double* x, const double* rp, double* A);   // A must be zero-filled by caller
template <int M> GDEV void dyn_B(const double* rp, double* Bm);                   // Bm must be zero-filled by caller

GDEV void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// ------------------------------------------------------------------------------------------ astrobeeSE3 (MRP)
template <> GDEV void dyn_f<ASTROBEE_SE3>(const double* x, const double* u, const double* rp, double* f) {
  const double *v = x + 3, *p = x + 6, *w = x + 9;
  f[0] = v[0]; f[1] = v[1]; f[2] = v[2];
  const double im = 1.0 / rp[RP_MASS];
  f[3] = u[0] * im; f[4] = u[1] * im; f[5] = u[2] * im;
  const double pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  const double wp = w[0] * p[0] + w[1] * p[1] + w[2] * p[2];
  double wxp[3];
  cross3(w, p, wxp);
  for (int i = 0; i < 3; ++i) f[6 + i] = 0.25 * ((1.0 - pp) * w[i] - 2.0 * wxp[i] + 2.0 * wp * p[i]);
  const double Jw[3] = {rp[RP_JXX] * w[0], rp[RP_JYY] * w[1], rp[RP_JZZ] * w[2]};
  double wxJw[3];
  cross3(w, Jw, wxJw);
  f[9] = (u[3] - wxJw[0]) / rp[RP_JXX];
  f[10] = (u[4] - wxJw[1]) / rp[RP_JYY];
  f[11] = (u[5] - wxJw[2]) / rp[RP_JZZ];
}

template <> GDEV void dyn_A<ASTROBEE_SE3>(const double* x, const double* rp, double* A) {
  constexpr int n = 12;
  const double Jxx = rp[RP_JXX], Jyy = rp[RP_JYY], Jzz = rp[RP_JZZ];
  const double px = x[6], py = x[7], pz = x[8], wx = x[9], wy = x[10], wz = x[11];
  A[0 * n + 3] = 1.0; A[1 * n + 4] = 1.0; A[2 * n + 5] = 1.0;
  const double d = (px * wx) / 2 + (py * wy) / 2 + (pz * wz) / 2;
  A[6 * n + 6] = d;
  A[6 * n + 7] = wz / 2 + (px * wy) / 2 - (py * wx) / 2;
  A[6 * n + 8] = (px * wz) / 2 - wy / 2 - (pz * wx) / 2;
  A[6 * n + 9] = px * px / 4 - py * py / 4 - pz * pz / 4 + 0.25;
  A[6 * n + 10] = (px * py) / 2 - pz / 2;
  A[6 * n + 11] = py / 2 + (px * pz) / 2;
  A[7 * n + 6] = (py * wx) / 2 - (px * wy) / 2 - wz / 2;
  A[7 * n + 7] = d;
  A[7 * n + 8] = wx / 2 + (py * wz) / 2 - (pz * wy) / 2;
  A[7 * n + 9] = pz / 2 + (px * py) / 2;
  A[7 * n + 10] = -px * px / 4 + py * py / 4 - pz * pz / 4 + 0.25;
  A[7 * n + 11] = (py * pz) / 2 - px / 2;
  A[8 * n + 6] = wy / 2 - (px * wz) / 2 + (pz * wx) / 2;
  A[8 * n + 7] = (pz * wy) / 2 - (py * wz) / 2 - wx / 2;
  A[8 * n + 8] = d;
  A[8 * n + 9] = (px * pz) / 2 - py / 2;
  A[8 * n + 10] = px / 2 + (py * pz) / 2;
  A[8 * n + 11] = -px * px / 4 - py * py / 4 + pz * pz / 4 + 0.25;
  A[9 * n + 10] = (Jyy - Jzz) * wz / Jxx;
  A[9 * n + 11] = (Jyy - Jzz) * wy / Jxx;
  A[10 * n + 9] = -(Jxx - Jzz) * wz / Jyy;
  A[10 * n + 11] = -(Jxx - Jzz) * wx / Jyy;
  A[11 * n + 9] = (Jxx - Jyy) * wy / Jzz;
  A[11 * n + 10] = (Jxx - Jyy) * wx / Jzz;
}

template <> GDEV void dyn_B<ASTROBEE_SE3>(const double* rp, double* Bm) {
  constexpr int m = 6;
  const double im = 1.0 / rp[RP_MASS];
  Bm[3 * m + 0] = im; Bm[4 * m + 1] = im; Bm[5 * m + 2] = im;
  Bm[9 * m + 3] = 1.0 / rp[RP_JXX]; Bm[10 * m + 4] = 1.0 / rp[RP_JYY]; Bm[11 * m + 5] = 1.0 / rp[RP_JZZ];
}

// ----------------------------------------------------------------------------- astrobeeSE3manifold (quaternion)
template <> GDEV void dyn_f<ASTROBEE_SE3_MANIFOLD>(const double* x, const double* u, const double* rp, double* f) {
  const double qw = x[6], qx = x[7], qy = x[8], qz = x[9];
  const double* w = x + 10;
  const double wx = w[0], wy = w[1], wz = w[2];
  f[0] = x[3]; f[1] = x[4]; f[2] = x[5];
  const double im = 1.0 / rp[RP_MASS];
  f[3] = u[0] * im; f[4] = u[1] * im; f[5] = u[2] * im;
  f[6] = 0.5 * (-wx * qx - wy * qy - wz * qz);
  f[7] = 0.5 * (wx * qw - wz * qy + wy * qz);
  f[8] = 0.5 * (wy * qw + wz * qx - wx * qz);
  f[9] = 0.5 * (wz * qw - wy * qx + wx * qy);
  const double Jw[3] = {rp[RP_JXX] * wx, rp[RP_JYY] * wy, rp[RP_JZZ] * wz};
  double wxJw[3];
  cross3(w, Jw, wxJw);
  f[10] = (u[3] - wxJw[0]) / rp[RP_JXX];
  f[11] = (u[4] - wxJw[1]) / rp[RP_JYY];
  f[12] = (u[5] - wxJw[2]) / rp[RP_JZZ];
}

template <> GDEV void dyn_A<ASTROBEE_SE3_MANIFOLD>(const double* x, const double* rp, double* A) {
  constexpr int n = 13;
  const double Jxx = rp[RP_JXX], Jyy = rp[RP_JYY], Jzz = rp[RP_JZZ];
  const double qw = x[6], qx = x[7], qy = x[8], qz = x[9], wx = x[10], wy = x[11], wz = x[12];
  A[0 * n + 3] = 1.0; A[1 * n + 4] = 1.0; A[2 * n + 5] = 1.0;
  A[6 * n + 7] = -wx / 2; A[6 * n + 8] = -wy / 2; A[6 * n + 9] = -wz / 2;
  A[6 * n + 10] = -qx / 2; A[6 * n + 11] = -qy / 2; A[6 * n + 12] = -qz / 2;
  A[7 * n + 6] = wx / 2; A[7 * n + 8] = -wz / 2; A[7 * n + 9] = wy / 2;
  A[7 * n + 10] = qw / 2; A[7 * n + 11] = qz / 2; A[7 * n + 12] = -qy / 2;
  A[8 * n + 6] = wy / 2; A[8 * n + 7] = wz / 2; A[8 * n + 9] = -wx / 2;
  A[8 * n + 10] = -qz / 2; A[8 * n + 11] = qw / 2; A[8 * n + 12] = qx / 2;
  A[9 * n + 6] = wz / 2; A[9 * n + 7] = -wy / 2; A[9 * n + 8] = wx / 2;
  A[9 * n + 10] = qy / 2; A[9 * n + 11] = -qx / 2; A[9 * n + 12] = qw / 2;
  A[10 * n + 11] = (Jyy - Jzz) * wz / Jxx;
  A[10 * n + 12] = (Jyy - Jzz) * wy / Jxx;
  A[11 * n + 10] = -(Jxx - Jzz) * wz / Jyy;
  A[11 * n + 12] = -(Jxx - Jzz) * wx / Jyy;
  A[12 * n + 10] = (Jxx - Jyy) * wy / Jzz;
  A[12 * n + 11] = (Jxx - Jyy) * wx / Jzz;
}

template <> GDEV void dyn_B<ASTROBEE_SE3_MANIFOLD>(const double* rp, double* Bm) {
  constexpr int m = 6;
  const double im = 1.0 / rp[RP_MASS];
  Bm[3 * m + 0] = im; Bm[4 * m + 1] = im; Bm[5 * m + 2] = im;
  Bm[10 * m + 3] = 1.0 / rp[RP_JXX]; Bm[11 * m + 4] = 1.0 / rp[RP_JYY]; Bm[12 * m + 5] = 1.0 / rp[RP_JZZ];
}

// -------------------------------------------------------------------------------------------- freeflyerSE2
template <> GDEV void dyn_f<FREEFLYER_SE2>(const double* x, const double* u, const double* rp, double* f) {
  f[0] = x[3]; f[1] = x[4]; f[2] = x[5];
  f[3] = u[0] / rp[RP_MASS];
  f[4] = u[1] / rp[RP_MASS];
  f[5] = u[2] / rp[RP_JXX];
}
template <> GDEV void dyn_A<FREEFLYER_SE2>(const double*, const double*, double* A) {
  A[0 * 6 + 3] = 1.0; A[1 * 6 + 4] = 1.0; A[2 * 6 + 5] = 1.0;
}
template <> GDEV void dyn_B<FREEFLYER_SE2>(const double* rp, double* Bm) {
  Bm[3 * 3 + 0] = 1.0 / rp[RP_MASS]; Bm[4 * 3 + 1] = 1.0 / rp[RP_MASS]; Bm[5 * 3 + 2] = 1.0 / rp[RP_JXX];
}

// -------------------------------------------------------------------------------------------------- dubins
template <> GDEV void dyn_f<DUBINS>(const double* x, const double* u, const double* rp, double* f) {
  f[0] = rp[RP_DUB_V] * cos(x[2]);
  f[1] = rp[RP_DUB_V] * sin(x[2]);
  f[2] = rp[RP_DUB_K] * u[0];
}
template <> GDEV void dyn_A<DUBINS>(const double* x, const double* rp, double* A) {
  A[0 * 3 + 2] = -rp[RP_DUB_V] * sin(x[2]);
  A[1 * 3 + 2] = rp[RP_DUB_V] * cos(x[2]);
}
template <> GDEV void dyn_B<DUBINS>(const double* rp, double* Bm) { Bm[2] = rp[RP_DUB_K]; }

// ------------------------------------------------------------------------------ per-model constraint tables
// Index ranges are constexpr so that the convex solve can place every row in its Hessian block at compile time.
// Soft quadratic state rows |x[i0:i1)|^2 - lim^2 (csi_translational_velocity_bound / csi_angular_velocity_bound:
// astrobee_se3.jl:244-252, astrobee_se3_manifold.jl:321-329, freeflyer_se2.jl:225-233).
template <int M> GHD constexpr int norm_i0(int j) {
  return M == ASTROBEE_SE3 ? (j == 0 ? 3 : 9) : M == ASTROBEE_SE3_MANIFOLD ? (j == 0 ? 3 : 10) : M == FREEFLYER_SE2 ? (j == 0 ? 3 : 5) : 0;
}
template <int M> GHD constexpr int norm_i1(int j) {
  return M == ASTROBEE_SE3 || M == ASTROBEE_SE3_MANIFOLD ? norm_i0<M>(j) + 3 : M == FREEFLYER_SE2 ? (j == 0 ? 5 : 6) : 0;
}
template <int M> GDEV void norm_row(int j, const double* rp, int* i0, int* i1, double* lim) {
  *i0 = norm_i0<M>(j); *i1 = norm_i1<M>(j);
  *lim = j == 0 ? rp[RP_VMAX] : rp[RP_WMAX];
}
// Soft linear state rows sign*x[i] - bound (csi_orientation_sign astrobee_se3_manifold.jl:316-319;
// csi_max/min_bound_constraints dynamics.jl:56-64 for dubins).
template <int M> GHD constexpr int lin_i(int j) { return M == ASTROBEE_SE3_MANIFOLD ? 6 : M == DUBINS ? j % 3 : 0; }
template <int M> GDEV void lin_row(int j, const double* rp, int* i, double* sign, double* bound) {
  *i = lin_i<M>(j);
  if (M == ASTROBEE_SE3_MANIFOLD) { *sign = -1.0; *bound = 0.0; }
  else if (M == DUBINS) { *sign = j < 3 ? 1.0 : -1.0; *bound = rp[RP_DUB_XMAX0 + (j % 3)]; }
  else { *sign = 0.0; *bound = 0.0; }
}
// Hard control balls |scale .* u[i0:i1)|^2 <= rad^2 for k = 1..N-1 (cci_translational_accel_bound /
// cci_angular_accel_bound: astrobee_se3.jl:255-263, freeflyer_se2.jl:236-245; cci_max/min_bound dynamics.jl:73-81).
template <int M> GHD constexpr int ball_i0(int j) {
  return M == ASTROBEE_SE3 || M == ASTROBEE_SE3_MANIFOLD ? 3 * j : M == FREEFLYER_SE2 ? (j == 0 ? 0 : 2) : 0;
}
template <int M> GHD constexpr int ball_i1(int j) {
  return M == ASTROBEE_SE3 || M == ASTROBEE_SE3_MANIFOLD ? 3 * j + 3 : M == FREEFLYER_SE2 ? (j == 0 ? 2 : 3) : 1;
}
template <int M> GDEV void ctrl_ball(int j, const double* rp, int* i0, int* i1, double* scale, double* rad) {
  *i0 = ball_i0<M>(j); *i1 = ball_i1<M>(j);
  if (M == ASTROBEE_SE3 || M == ASTROBEE_SE3_MANIFOLD) {
    if (j == 0) { scale[0] = scale[1] = scale[2] = 1.0 / rp[RP_MASS]; *rad = rp[RP_AMAX]; }
    else { scale[0] = 1.0 / rp[RP_JXX]; scale[1] = 1.0 / rp[RP_JYY]; scale[2] = 1.0 / rp[RP_JZZ]; *rad = rp[RP_ALMAX]; }
  } else if (M == FREEFLYER_SE2) {
    if (j == 0) { scale[0] = scale[1] = 1.0 / rp[RP_MASS]; *rad = rp[RP_AMAX]; }
    else { scale[0] = 1.0 / rp[RP_JXX]; *rad = rp[RP_ALMAX]; }
  } else {
    scale[0] = 1.0; *rad = rp[RP_DUB_UMAX];
  }
}

}  // namespace gusto
