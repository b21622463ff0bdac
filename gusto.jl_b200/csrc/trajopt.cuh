// Evaluation scalars of the TrajOpt SCP variant, solve_trajopt_jump! (/root/reference/src/scp/scp_trajopt.jl:33-157), one CTA per
// instance.  The convex subproblem itself is the second compilation of ipm.cuh (GUSTO_IPM_ALG = 1).
//
//   trajopt_evaluate_instance   candidate (X, U) against the accepted (Xp, Up) the blocks were linearised at:
//       evaluate_xtol = convergence_metric (traj_opt.jl:74-85, scp_trajopt.jl:281-283), cost_true of both (for evaluate_ftol,
//       :285-287) and trust_region_ratio_trajopt (astrobee_se3.jl:419-459, freeflyer_se2.jl:429-469) with its quirks kept:
//       the finite-difference term (X[:,k] - X[:,k]) / dt is zero, so phi_old = |f_p[k]|_1 and phi_new = |f(X_k, U_k)|_1;
//       phi_hat is the l1 norm of the linearised trapezoid row; the obstacle model is linearised at the NEW point and sums
//       over ALL obstacles (both hulls of the freeflyer's compound body, each at the body translation).
//   trajopt_ctol_instance       (X, U) against a reference trajectory (Xr, Ur): evaluate_ctol (:288-312) as the oracle restates it
//       (oracle/gusto_oracle/trajopt.py: the literal code indexes a Dict with 1 and cannot run): sum over the constraint classes
//       of max |c(traj) - c(ref)| and of max |c(traj)|; classes: every norm row, every linear row, obstacle signed distances,
//       goal box, nonlinear trapezoid defect (norm per knot).
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "sdf.cuh"
#include "evaluate.cuh"

namespace gusto {

constexpr int TRAJOPT_NOUT = 8;   // xtol, rho, J_true(U), J_true(Up), max_k |dX_k|^2, rho numerator, rho denominator, sum |linearised defect|_1

template <int M>
GDEV void trajopt_evaluate_instance(const BatchDesc& d, const BatchPtrs& p, int b, const double* X, const double* U, double* out, double* red) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU;
  const int N = d.N;
  const double h = p.tf[b] / (N - 1), hh = 0.5 * h;
  const double cl = d.rp[RP_CLEAR], R = d.rp[RP_RADIUS];
  const double* Xp = p.Xp + (size_t)b * N * NX;
  const double* Up = p.Up + (size_t)b * N * NU;
  const size_t np = g_np(N), fs = (size_t)d.n_obs * np;
  const double* F = p.f + (size_t)b * NX * np;
  const double* A = p.A + (size_t)b * T::ANZ * np;
  const double* G = p.g + (size_t)b * NX * np;
  const double* rows = p.rows + (size_t)b * 5 * fs;
  double Bm[NX * NU];
  for (int i = 0; i < NX * NU; ++i) Bm[i] = 0.0;
  dyn_B<M>(d.rp, Bm);

  double num = 0, den = 0, Jn = 0, Jp = 0, mdx2 = 0, mnx2 = 0, dsum = 0;
  G_PAR_FOR(k, N) {
    const double* x = X + k * NX;
    const double* u = U + k * NU;
    double dx2 = 0, nx2 = 0, uu = 0, up2 = 0;
    for (int i = 0; i < NX; ++i) { const double dxi = x[i] - Xp[k * NX + i]; dx2 += dxi * dxi; nx2 += x[i] * x[i]; }
    for (int i = 0; i < NU; ++i) { uu += u[i] * u[i]; up2 += Up[k * NU + i] * Up[k * NU + i]; }
    mdx2 = dx2 > mdx2 ? dx2 : mdx2;
    mnx2 = nx2 > mnx2 ? nx2 : mnx2;
    const double wk = (k == 0 || k == N - 1) ? 0.5 * h : h;
    Jn += wk * uu; Jp += wk * up2;
    if (k < N - 1) {
      double fn[NX], fl0[NX], fl1[NX], po = 0, pnew = 0, ph = 0;
      dyn_f<M>(x, u, d.rp, fn);
      const double* x1 = x + NX;
      const double* u1 = u + NU;
      for (int i = 0; i < NX; ++i) { fl0[i] = G[i * np + k]; fl1[i] = G[i * np + k + 1]; }
#pragma unroll
      for (int e = 0; e < T::ANZ; ++e) {
        fl0[T::a_row(e)] += A[e * np + k] * x[T::a_col(e)];
        fl1[T::a_row(e)] += A[e * np + k + 1] * x1[T::a_col(e)];
      }
      for (int i = 0; i < NX; ++i)
        for (int a = 0; a < NU; ++a) { fl0[i] += Bm[i * NU + a] * u[a]; fl1[i] += Bm[i * NU + a] * u1[a]; }
      for (int i = 0; i < NX; ++i) {
        po += fabs(F[i * np + k]);
        pnew += fabs(fn[i]);
        ph += fabs(x[i] - x1[i] + hh * (fl0[i] + fl1[i]));
      }
      num += po - pnew;
      den += po - ph;
      dsum += ph;
    }
  }
  if (T::WS > 0) {
    constexpr int WS = T::WS > 0 ? T::WS : 1;
    G_PAR_FOR(it, N * d.n_obs) {
      const int i = it / N, k = it - i * N;
      double r[3], r0[3], d0, d1, n1[3], n0[3];
      workspace_location<WS>(X + k * NX, r);
      workspace_location<WS>(Xp + k * NX, r0);
      d0 = rows[4 * fs + (size_t)i * np + k];                       // dist0 of the linearize kernel
      signed_distance<WS>(r, d.obs_kind[i], d.obs_a[i], d.obs_b[i], R, &d1, n1);
      double dr = 0.0;
      for (int a = 0; a < 3; ++a) dr += n1[a] * (r[a] - r0[a]);
      num += d1 - d0;                                               // (cl - d0) - (cl - d1)
      den += d1 - d0 + dr;                                          // (cl - d0) - (cl - (d1 + nhat.(r - r0)))
      if (M == FREEFLYER_SE2) {                                     // second hull of the compound robot (robot/freeflyer.jl:48)
        double ra[3] = {r[0], r[1] + 0.15, r[2]}, ra0[3] = {r0[0], r0[1] + 0.15, r0[2]};
        signed_distance<WS>(ra0, d.obs_kind[i], d.obs_a[i], d.obs_b[i], R, &d0, n0);
        signed_distance<WS>(ra, d.obs_kind[i], d.obs_a[i], d.obs_b[i], R, &d1, n1);
        dr = 0.0;
        for (int a = 0; a < 3; ++a) dr += n1[a] * (r[a] - r0[a]);
        num += d1 - d0;
        den += d1 - d0 + dr;
      }
      (void)cl;
    }
  }
  num = block_sum(num, red); den = block_sum(den, red); Jn = block_sum(Jn, red); Jp = block_sum(Jp, red); dsum = block_sum(dsum, red);
  mdx2 = block_max(mdx2, red); mnx2 = block_max(mnx2, red);
  if (G_TID == 0) {
    out[0] = sqrt(mdx2) / sqrt(mnx2);
    out[1] = num / den;
    out[2] = Jn; out[3] = Jp; out[4] = mdx2; out[5] = num; out[6] = den; out[7] = dsum;
  }
}

// out[0] = sum over classes of max |c(X,U) - c(Xr,Ur)|, out[1] = sum over classes of max |c(X,U)|  (evaluate_ctol = out[0] / out[1])
template <int M>
GDEV void trajopt_ctol_instance(const BatchDesc& d, const BatchPtrs& p, int b, const double* X, const double* U, const double* Xr, const double* Ur,
                                double* out, double* red) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU, NCL = T::NNORM + T::NLIN + 3;
  const int N = d.N;
  const double hh = 0.5 * p.tf[b] / (N - 1);
  const double cl = d.rp[RP_CLEAR], R = d.rp[RP_RADIUS];
  double mn[NCL], md[NCL];
  for (int c2 = 0; c2 < NCL; ++c2) { mn[c2] = 0.0; md[c2] = 0.0; }
  auto upd = [&](int cls, double a, double r) { const double dv = fabs(a - r), av = fabs(a); mn[cls] = dv > mn[cls] ? dv : mn[cls]; md[cls] = av > md[cls] ? av : md[cls]; };
  G_PAR_FOR(k, N) {
    const double* x = X + k * NX;
    const double* xr = Xr + k * NX;
    for (int j = 0; j < T::NNORM; ++j) {
      int i0, i1; double lim;
      norm_row<M>(j, d.rp, &i0, &i1, &lim);
      double a = -lim * lim, r = -lim * lim;
      for (int i = i0; i < i1; ++i) { a += x[i] * x[i]; r += xr[i] * xr[i]; }
      upd(j, a, r);
    }
    for (int j = 0; j < T::NLIN; ++j) {
      int i; double sign, bound;
      lin_row<M>(j, d.rp, &i, &sign, &bound);
      upd(T::NNORM + j, sign * x[i] - bound, sign * xr[i] - bound);
    }
    if (k < N - 1) {     // nonlinear trapezoid defect, norm of the difference / of the value
      double f0[NX], f1[NX], g0[NX], g1[NX], e2 = 0, a2 = 0;
      dyn_f<M>(x, U + k * NU, d.rp, f0); dyn_f<M>(x + NX, U + (k + 1) * NU, d.rp, f1);
      dyn_f<M>(xr, Ur + k * NU, d.rp, g0); dyn_f<M>(xr + NX, Ur + (k + 1) * NU, d.rp, g1);
      for (int i = 0; i < NX; ++i) {
        const double a = x[i] - x[NX + i] + hh * (f0[i] + f1[i]), r = xr[i] - xr[NX + i] + hh * (g0[i] + g1[i]);
        e2 += (a - r) * (a - r); a2 += a * a;
      }
      const int cls = T::NNORM + T::NLIN + 2;
      e2 = sqrt(e2); a2 = sqrt(a2);
      mn[cls] = e2 > mn[cls] ? e2 : mn[cls]; md[cls] = a2 > md[cls] ? a2 : md[cls];
    }
    if (k == N - 1) {
      for (int i = 0; i < NX; ++i) if (d.goal_type[i] == GOAL_BOX) {
        const double lo = p.goal_lo[(size_t)b * NX + i], hi = p.goal_hi[(size_t)b * NX + i];
        upd(T::NNORM + T::NLIN + 1, x[i] - hi, xr[i] - hi);
        upd(T::NNORM + T::NLIN + 1, lo - x[i], lo - xr[i]);
      }
    }
  }
  if (T::WS > 0) {
    constexpr int WS = T::WS > 0 ? T::WS : 1;
    G_PAR_FOR(it, N * d.n_obs) {
      const int i = it / N, k = it - i * N;
      double r[3], r0[3], d0, d1, nn[3];
      workspace_location<WS>(X + k * NX, r);
      workspace_location<WS>(Xr + k * NX, r0);
      signed_distance<WS>(r, d.obs_kind[i], d.obs_a[i], d.obs_b[i], R, &d1, nn);
      signed_distance<WS>(r0, d.obs_kind[i], d.obs_a[i], d.obs_b[i], R, &d0, nn);
      upd(T::NNORM + T::NLIN, cl - d1, cl - d0);
    }
  }
  double sn = 0.0, sd = 0.0;
  for (int c2 = 0; c2 < NCL; ++c2) { sn += block_max(mn[c2], red); sd += block_max(md[c2], red); }
  if (G_TID == 0) { out[0] = sn; out[1] = sd; }
}

}  // namespace gusto
