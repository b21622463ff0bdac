// Closed-form signed distance of the robot's bounding sphere (radius R, centre r, never rotated) to one
// collision component, replacing BulletCollision.distance(env, rb_idx, r, env_idx) -> (dist, xbody, xobs)
// (39 call sites in the reference, e.g. dynamics/astrobee_se3.jl:291,403,409; contract in SURVEY.md App. E).
// Returns dist and the outward normal nhat the callers build from (xbody, xobs) (astrobee_se3.jl:296-298).
//   box, centre outside : c = clamp(r, lo, hi); d = |r-c|; nhat = (r-c)/d; dist = d - R
//   box, centre inside  : q = max(lo-r, r-hi) <= 0; j = argmax q (first on ties); nhat = +-e_j; dist = q_j - R
//   sphere (c, rho)     : d = |r-c|; nhat = (r-c)/d; dist = d - rho - R
// WS = 3: sphere vs box/sphere in space (Astrobee3D).  WS = 2: circle vs rectangle in the table plane
// (Freeflyer body cylinder, robot/freeflyer.jl:53-57), nhat_z = 0.
#pragma once
#include "common.cuh"

namespace gusto {

template <int WS>
GDEV void signed_distance(const double* r, int kind, const double* a, const double* b, double R, double* dist,
                          double* nhat) {
  nhat[0] = nhat[1] = nhat[2] = 0.0;
  if (kind == OBS_BOX) {
    double q[3], diff[3];
    bool outside = false;
    double d2 = 0.0;
    int j = 0;
    for (int i = 0; i < WS; ++i) {
      const double ql = a[i] - r[i], qh = r[i] - b[i];
      q[i] = ql > qh ? ql : qh;
      if (q[i] > 0.0) outside = true;
      const double c = r[i] < a[i] ? a[i] : (r[i] > b[i] ? b[i] : r[i]);
      diff[i] = r[i] - c;
      d2 += diff[i] * diff[i];
      if (q[i] > q[j]) j = i;
    }
    if (outside) {
      // 1 / sqrt(d2) once (g_rsqrt, ~1 ulp) instead of an IEEE sqrt and WS IEEE divisions: K1 is issue-bound and this query is
      // most of its instructions; outside => d2 > 0
      const double inv = g_rsqrt(d2), d = d2 * inv;
      for (int i = 0; i < WS; ++i) nhat[i] = diff[i] * inv;
      *dist = d - R;
    } else {
      nhat[j] = (r[j] - b[j] >= a[j] - r[j]) ? 1.0 : -1.0;
      *dist = q[j] - R;
    }
  } else {
    double d2 = 0.0, diff[3];
    for (int i = 0; i < WS; ++i) { diff[i] = r[i] - a[i]; d2 += diff[i] * diff[i]; }
    const double inv = g_rsqrt(d2), d = d2 > 0.0 ? d2 * inv : 0.0;
    for (int i = 0; i < WS; ++i) nhat[i] = diff[i] * inv;      // (centre on centre: NaN normal, as 0 / 0 gave)
    *dist = d - b[0] - R;
  }
}

// get_workspace_location: astrobee_se3.jl:319-321 (X[1:3,k]); freeflyer_se2.jl:334-336 ([X[1:2,k]; 0]).
template <int WS> GDEV void workspace_location(const double* x, double* r) {
  r[0] = WS >= 1 ? x[0] : 0.0;
  r[1] = WS >= 2 ? x[1] : 0.0;
  r[2] = WS >= 3 ? x[2] : 0.0;
}

}  // namespace gusto
