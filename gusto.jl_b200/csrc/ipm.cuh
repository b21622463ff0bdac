// K3: batched convex-subproblem solve, one CTA per problem instance.
//
// Solves the GuSTO penalized QCQP assembled by add_constraints_gusto_jump! / add_objective_gusto_jump!
// (/root/reference/src/scp/scp_gusto.jl:192-314; exact form in SURVEY.md App. A) directly from the blocks the
// linearize kernel left in HBM (A_k, g_k, obstacle rows) -- no model object is ever built.  It stands in for
// JuMP.optimize! (:104), i.e. for the barrier methods of Gurobi / Ipopt.
//
// Method: infeasible-start primal-dual interior point with Mehrotra predictor-corrector (same algorithm as
// oracle/gusto_oracle/ipm.py, so iterates can be compared one-to-one), specialised to the problem structure:
//   * every inequality touches one knot only; its slack t and both multipliers are eliminated analytically,
//     leaving a block-diagonal reduced Hessian  H = blkdiag(Hx_k [NXxNX], Hu_k [NUxNU]);
//   * the equality rows (init, trapezoid dynamics, point goal) are block-bidiagonal, so the Schur complement
//     S = A (H + dp I)^-1 A' is block-tridiagonal with N+1 blocks of NX x NX and is factorised by a block
//     Cholesky sweep; Newton directions are then recovered with `nref` steps of iterative refinement against
//     the unregularised KKT matrix (H alone is only positive SEMI-definite: the cost has no state term).
// A first-order splitting (ADMM, prototyped in tools/admm_proto.py) was rejected: GuSTO's accept test compares
// soft rows against eps = 1e-6 (astrobee_se3.jl:31, scp_gusto.jl:318-327), and with a trapezoid double
// integrator over 70 s ADMM needs >2000 iterations for 1e-6 residuals while this method needs 8-25 for 1e-8.
//
// All arithmetic is FP64 (the reference is Float64 throughout).
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "evaluate.cuh"   // block_sum / block_max
#ifdef GUSTO_HOSTSIM
#include <cstdio>
#include <cstdlib>
#endif

namespace gusto {

#ifdef GUSTO_HOSTSIM
GDEV long long g_clock() { return 0; }
#else
GDEV long long g_clock() { return clock64(); }
#endif

struct IpmParams {
  int max_iter;      // Newton iterations cap
  int nref;          // refinement steps on the corrector solve
  double tol;        // max(|r_dual|/(1+omega), |r_eq|, |r_ineq|, mu) <= tol
  double delta_p;    // primal regularisation added to Hx, Hu in the factorisation only
  double delta_d;    // relative regularisation of the Schur diagonal
};

enum : int { IPM_OPTIMAL = 0, IPM_ITERATION_LIMIT = 1, IPM_NUMERICAL = 2 };
constexpr int SLOT_W = 6;   // s, lam, t, lamb, pa (ds*dlam of the predictor), pb (dt*dlamb of the predictor)
constexpr int IPM_NINFO = 8;  // status, iterations, residual, mu, objective, cycles: assemble+slots, factorize, kkt solves

template <int M> struct IpmLayout {
  using T = Traits<M>;
  static constexpr int NX = T::NX, NU = T::NU, NV = NX + NU;
  static constexpr int S_TR = 0;
  static constexpr int S_NORM = 1;
  static constexpr int S_LIN = S_NORM + T::NNORM;
  static constexpr int S_QUAT = S_LIN + T::NLIN;          // hinge, then hard row
  static constexpr int S_BALL = S_QUAT + 2 * T::HAS_QUAT;
  static constexpr int S_BOX = S_BALL + T::NBALL;         // 2*NX goal-box rows (upper, lower per coordinate)
  static constexpr int S_OBS = S_BOX + 2 * NX;
  GHD static int nslots(int n_obs) { return S_OBS + n_obs; }
  // doubles of global scratch per instance
  GHD static size_t scratch_doubles(int N, int n_obs) {
    const size_t nz = (size_t)N * NV, ne = (size_t)(N + 1) * NX;
    return 6 * nz + 5 * ne + (size_t)N * nslots(n_obs) * SLOT_W + (size_t)N * NX * NX * 6 + (size_t)N * NU * NU * 2 +
           (size_t)N * NU * NX + (size_t)(N + 1) * NX * NX * 2;
  }
  GHD static int smem_doubles(int N, int nthr) { return (N + 1) * NX + 4 * NX * NX + 16 + nthr + 16; }
};

template <int M> struct IpmCtx {
  using L = IpmLayout<M>;
  static constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const BatchDesc* d;
  int N, n_obs, S, b;
  double h, omega, Delta, toggle, eps;
  const double *Xp, *Up, *A, *g, *rows, *x_init, *goal_lo, *goal_hi;
  double Bm[NX * NU];     // constant B
  // global scratch
  double *z, *nu, *r, *dz, *t1, *res, *e, *rnu, *dnu, *resnu, *enu, *bS, *slot;
  double *Hx, *Ci, *CA, *W, *Hu, *Cu, *GT, *Ld, *Lo;
  // shared
  double *sy, *sD, *sLo, *sLi, *sB4, *stmp, *red;
};

// ------------------------------------------------------------------------------------------------- slots
struct SlotEval {
  bool valid, is_u, has_t;
  int i0, i1;             // variable range the gradient lives on
  double c0;              // constraint value without the -t term
  double gv[MAX_NX];      // gradient on [i0, i1)
  double hq[MAX_NX];      // diagonal of the constraint Hessian on [i0, i1)
};

template <int M>
GDEV void slot_eval(const IpmCtx<M>& c, int k, int s, const double* x, const double* u, SlotEval& o) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX;
  const double* rp = c.d->rp;
  o.valid = false; o.is_u = false; o.has_t = true; o.i0 = 0; o.i1 = 0; o.c0 = 0.0;
  if (s == L::S_TR) {
    if (!T::HAS_TR) return;
    // stri_state_trust_region (astrobee_se3.jl:308-311) in slack-scaled form: |x - xp|^2 - Delta/omega - t <= 0
    o.valid = true; o.i0 = 0; o.i1 = NX;
    double v = -c.Delta / c.omega;
    for (int i = 0; i < NX; ++i) { const double dxi = x[i] - c.Xp[k * NX + i]; o.gv[i] = 2.0 * dxi; o.hq[i] = 2.0; v += dxi * dxi; }
    o.c0 = v;
  } else if (s < L::S_LIN) {
    int i0, i1; double lim;
    norm_row<M>(s - L::S_NORM, rp, &i0, &i1, &lim);
    o.valid = true; o.i0 = i0; o.i1 = i1;
    double v = -lim * lim;
    for (int i = i0; i < i1; ++i) { o.gv[i - i0] = 2.0 * x[i]; o.hq[i - i0] = 2.0; v += x[i] * x[i]; }
    o.c0 = v;
  } else if (s < L::S_QUAT) {
    int i; double sign, bound;
    lin_row<M>(s - L::S_LIN, rp, &i, &sign, &bound);
    o.valid = true; o.i0 = i; o.i1 = i + 1; o.gv[0] = sign; o.hq[0] = 0.0; o.c0 = sign * x[i] - bound;
  } else if (s < L::S_BALL) {
    // cse_quaternion_norm (astrobee_se3_manifold.jl:308-313): e = a.q - 1, a = qp/|qp|.
    //   hinge  e - eps/omega - t <= 0 (t >= 0)       and the hard row  -e - eps/omega <= 0   
    const double* qp = c.Xp + k * NX + 6;
    const double nq = sqrt(qp[0] * qp[0] + qp[1] * qp[1] + qp[2] * qp[2] + qp[3] * qp[3]);
    double ev = -1.0;
    for (int i = 0; i < 4; ++i) ev += qp[i] / nq * x[6 + i];
    const bool hinge = (s == L::S_QUAT);
    o.valid = true; o.i0 = 6; o.i1 = 10; o.has_t = hinge;
    for (int i = 0; i < 4; ++i) { o.gv[i] = (hinge ? 1.0 : -1.0) * qp[i] / nq; o.hq[i] = 0.0; }
    o.c0 = (hinge ? ev : -ev) - c.eps / c.omega;
  } else if (s < L::S_BOX) {
    if (k >= c.N - 1) return;       // control bounds cover k = 1..N-1 only (astrobee_se3.jl:370-371, quirk q3)
    int i0, i1; double scale[3], rad;
    ctrl_ball<M>(s - L::S_BALL, rp, &i0, &i1, scale, &rad);
    o.valid = true; o.is_u = true; o.has_t = false; o.i0 = i0; o.i1 = i1;
    double v = -rad * rad;
    for (int i = i0; i < i1; ++i) {
      const double s2 = scale[i - i0] * scale[i - i0];
      o.gv[i - i0] = 2.0 * s2 * u[i]; o.hq[i - i0] = 2.0 * s2; v += s2 * u[i] * u[i];
    }
    o.c0 = v;
  } else if (s < L::S_OBS) {
    // csbci_goal_constraints (dynamics.jl:37-42): X[i,N] - ub <= 0, lb - X[i,N] <= 0, hard
    const int j = s - L::S_BOX, i = j >> 1;
    if (k != c.N - 1 || c.d->goal_type[i] != GOAL_BOX) return;
    o.valid = true; o.has_t = false; o.i0 = i; o.i1 = i + 1; o.hq[0] = 0.0;
    if ((j & 1) == 0) { o.gv[0] = 1.0; o.c0 = x[i] - c.goal_hi[i]; }
    else { o.gv[0] = -1.0; o.c0 = c.goal_lo[i] - x[i]; }
  } else {
    // ncsi_obstacle_avoidance_constraints_convexified (astrobee_se3.jl:282-305): off - nhat.r - t <= 0 if dist0 < toggle
    const int i = s - L::S_OBS;
    const double* row = c.rows + ((size_t)k * c.n_obs + i) * 5;
    if (!(row[4] < c.toggle)) return;
    o.valid = true; o.i0 = 0; o.i1 = T::WS;
    double v = row[3];
    for (int a = 0; a < T::WS; ++a) { o.gv[a] = -row[a]; o.hq[a] = 0.0; v -= row[a] * x[a]; }
    o.c0 = v;
  }
}

// ------------------------------------------------------------------------------ small dense helpers (row-major)
template <int n> GDEV bool chol_lower(double* H) {          // in place, lower triangle; upper left untouched
  bool ok = true;
  for (int j = 0; j < n; ++j) {
    double djj = H[j * n + j];
    for (int m = 0; m < j; ++m) djj -= H[j * n + m] * H[j * n + m];
    if (!(djj > 0.0)) { ok = false; djj = 1e-300; }
    const double l = sqrt(djj);
    H[j * n + j] = l;
    for (int i = j + 1; i < n; ++i) {
      double v = H[i * n + j];
      for (int m = 0; m < j; ++m) v -= H[i * n + m] * H[j * n + m];
      H[i * n + j] = v / l;
    }
  }
  return ok;
}
// column `col` of the inverse of lower-triangular C, written into out[:, col] (full column, zeros above col)
template <int n> GDEV void tri_inv_col(const double* C, int col, double* out) {
  double y[n];
  for (int i = 0; i < n; ++i) {
    if (i < col) { y[i] = 0.0; continue; }
    double v = (i == col) ? 1.0 : 0.0;
    for (int m = col; m < i; ++m) v -= C[i * n + m] * y[m];
    y[i] = v / C[i * n + i];
  }
  for (int i = 0; i < n; ++i) out[i * n + col] = y[i];
}

// ------------------------------------------------------------------------------------- structured operators
// Row j of the equality system (j = 0..N), applied to a primal vector v (layout [k][NV]):
//   j = 0      : x_0
//   1..N-1     : (I + h/2 A_{j-1}) x_{j-1} + G u_{j-1} - (I - h/2 A_j) x_j + G u_j          (G = h/2 B)
//   j = N      : M x_{N-1}      (M = diag(goal_type == POINT))
template <int M> GDEV double apply_A_entry(const IpmCtx<M>& c, const double* v, int j, int i) {
  constexpr int NX = IpmCtx<M>::NX, NU = IpmCtx<M>::NU, NV = IpmCtx<M>::NV;
  const double hh = 0.5 * c.h;
  if (j == 0) return v[i];
  if (j == c.N) return c.d->goal_type[i] == GOAL_POINT ? v[(c.N - 1) * NV + i] : 0.0;
  const double* vp = v + (j - 1) * NV;
  const double* vc = v + j * NV;
  const double* Ap = c.A + (size_t)(j - 1) * NX * NX + i * NX;
  const double* Ac = c.A + (size_t)j * NX * NX + i * NX;
  double acc = vp[i] - vc[i];
  for (int m = 0; m < NX; ++m) acc += hh * (Ap[m] * vp[m] + Ac[m] * vc[m]);
  for (int m = 0; m < NU; ++m) acc += hh * c.Bm[i * NU + m] * (vp[NX + m] + vc[NX + m]);
  return acc;
}
// Entry (k, i) of A' nu (i < NX: state part, i >= NX: control part)
template <int M> GDEV double apply_AT_entry(const IpmCtx<M>& c, const double* nu, int k, int i) {
  constexpr int NX = IpmCtx<M>::NX, NU = IpmCtx<M>::NU;
  const double hh = 0.5 * c.h;
  const int N = c.N;
  const double* nk = nu + k * NX;          // row k       (knot k as "current")
  const double* nn = nu + (k + 1) * NX;    // row k + 1   (knot k as "previous")
  const double* Ak = c.A + (size_t)k * NX * NX;
  if (i < NX) {
    double acc;
    if (k == 0) acc = nk[i]; else acc = -nk[i];
    if (k == N - 1) acc += c.d->goal_type[i] == GOAL_POINT ? nn[i] : 0.0; else acc += nn[i];
    double s = 0.0;
    for (int m = 0; m < NX; ++m) {
      double w = 0.0;
      if (k > 0) w += nk[m];
      if (k < N - 1) w += nn[m];
      s += Ak[m * NX + i] * w;
    }
    return acc + hh * s;
  } else {
    const int a = i - NX;
    double s = 0.0;
    for (int m = 0; m < NX; ++m) {
      double w = 0.0;
      if (k > 0) w += nk[m];
      if (k < N - 1) w += nn[m];
      s += c.Bm[m * NU + a] * w;
    }
    return hh * s;
  }
}

// out = (H + dp I)^-1 in   per knot, using the inverse Cholesky factors Ci, Cu  (H^-1 = Ci' Ci)
template <int M> GDEV_NOINLINE void apply_Hinv(const IpmCtx<M>& c, const double* in, double* out) {
  constexpr int NX = IpmCtx<M>::NX, NU = IpmCtx<M>::NU, NV = IpmCtx<M>::NV;
  G_PAR_FOR(k, c.N) {
    const double* Ci = c.Ci + (size_t)k * NX * NX;
    const double* Cu = c.Cu + (size_t)k * NU * NU;
    double t[NX], tu[NU > 0 ? NU : 1];
    for (int i = 0; i < NX; ++i) { double a = 0; for (int m = 0; m <= i; ++m) a += Ci[i * NX + m] * in[k * NV + m]; t[i] = a; }
    for (int i = 0; i < NX; ++i) { double a = 0; for (int m = i; m < NX; ++m) a += Ci[m * NX + i] * t[m]; out[k * NV + i] = a; }
    for (int i = 0; i < NU; ++i) { double a = 0; for (int m = 0; m <= i; ++m) a += Cu[i * NU + m] * in[k * NV + NX + m]; tu[i] = a; }
    for (int i = 0; i < NU; ++i) { double a = 0; for (int m = i; m < NU; ++m) a += Cu[m * NU + i] * tu[m]; out[k * NV + NX + i] = a; }
  }
}

// Block-tridiagonal solve  S y = b  in shared memory (sy holds b on entry, the solution on exit).
// The sweep is a chain of 2(N+1) dependent NX x NX mat-vecs.  It runs on the first warp only (warp-level barriers);
// the factor blocks of the NEXT row are prefetched from HBM/L2 into a shared-memory double buffer with cp.async
// while the current block is applied, so the chain never waits on a global load.
template <int M> GDEV_NOINLINE void schur_solve(const IpmCtx<M>& c) {
  constexpr int NX = IpmCtx<M>::NX, NN = NX * NX;
  const int N = c.N;
  double* y = c.sy;
  double* tmp = c.stmp;
  double* bLd[2] = {c.sD, c.sLo};
  double* bLo[2] = {c.sLi, c.sB4};
  if (G_TID < G_WARP) {
    // ---- forward: y_j = Ld_j (b_j - Lo_j y_{j-1})
    G_W0_FOR(it, NN) { g_cp_async8(&bLd[0][it], c.Ld + it); }
    g_cp_async_wait();
    G_SYNCWARP();
    for (int j = 0; j <= N; ++j) {
      const int cur = j & 1, nxt = cur ^ 1;
      if (j < N) {
        const double* Ldn = c.Ld + (size_t)(j + 1) * NN;
        const double* Lon = c.Lo + (size_t)(j + 1) * NN;
        G_W0_FOR(it, NN) { g_cp_async8(&bLd[nxt][it], Ldn + it); g_cp_async8(&bLo[nxt][it], Lon + it); }
      }
      G_W0_FOR(i, NX) {
        double a = y[j * NX + i];
        if (j > 0) for (int m = 0; m < NX; ++m) a -= bLo[cur][i * NX + m] * y[(j - 1) * NX + m];
        tmp[i] = a;
      }
      G_SYNCWARP();
      G_W0_FOR(i, NX) {
        double a = 0;
        for (int m = 0; m <= i; ++m) a += bLd[cur][i * NX + m] * tmp[m];
        y[j * NX + i] = a;
      }
      g_cp_async_wait();
      G_SYNCWARP();
    }
    // ---- backward: nu_j = Ld_j' (y_j - Lo_{j+1}' nu_{j+1}).  After the forward sweep buffer (N & 1) holds Ld_N, Lo_N.
    for (int j = N; j >= 0; --j) {
      const int cur = j & 1, nxt = cur ^ 1;     // bLd[cur] = Ld_j ; bLo[nxt] = Lo_{j+1} (staged while row j+1 was applied)
      if (j > 0) {
        const double* Ldn = c.Ld + (size_t)(j - 1) * NN;
        G_W0_FOR(it, NN) { g_cp_async8(&bLd[nxt][it], Ldn + it); }
      }
      G_W0_FOR(i, NX) {
        double a = y[j * NX + i];
        if (j < N) for (int m = 0; m < NX; ++m) a -= bLo[nxt][m * NX + i] * y[(j + 1) * NX + m];
        tmp[i] = a;
      }
      G_SYNCWARP();
      // Lo_{j+1} is consumed: its buffer can take Lo_j (needed at step j-1 as "Lo_{(j-1)+1}")
      if (j > 0) {
        const double* Lon = c.Lo + (size_t)j * NN;
        G_W0_FOR(it, NN) { g_cp_async8(&bLo[cur][it], Lon + it); }
      }
      G_W0_FOR(i, NX) {
        double a = 0;
        for (int m = i; m < NX; ++m) a += bLd[cur][m * NX + i] * tmp[m];
        y[j * NX + i] = a;
      }
      g_cp_async_wait();
      G_SYNCWARP();
    }
  }
  G_SYNC();
}

// [dz; dnu] = Ktilde^-1 [r; rnu]  with Ktilde = [[H + dp I, A'], [A, -(dd) ]] via the Schur complement.
template <int M> GDEV_NOINLINE void kkt_solve(const IpmCtx<M>& c, const double* r, const double* rnu, double* dz, double* dnu) {
  constexpr int NX = IpmCtx<M>::NX, NV = IpmCtx<M>::NV;
  const int N = c.N;
  apply_Hinv<M>(c, r, c.t1);
  G_SYNC();
  G_PAR_FOR(it, (N + 1) * NX) { const int j = it / NX, i = it - j * NX; c.sy[it] = apply_A_entry<M>(c, c.t1, j, i) - rnu[it]; }
  G_SYNC();
  schur_solve<M>(c);
  G_PAR_FOR(it, (N + 1) * NX) dnu[it] = c.sy[it];
  G_SYNC();
  G_PAR_FOR(it, N * NV) { const int k = it / NV, i = it - k * NV; c.t1[it] = r[it] - apply_AT_entry<M>(c, dnu, k, i); }
  G_SYNC();
  apply_Hinv<M>(c, c.t1, dz);
  G_SYNC();
}

// Refinement against the exact KKT matrix [[H, A'], [A, 0]].
template <int M> GDEV void kkt_solve_refined(const IpmCtx<M>& c, const double* r, const double* rnu, double* dz, double* dnu, int nref) {
  constexpr int NX = IpmCtx<M>::NX, NU = IpmCtx<M>::NU, NV = IpmCtx<M>::NV;
  const int N = c.N;
  kkt_solve<M>(c, r, rnu, dz, dnu);
  for (int it_ref = 0; it_ref < nref; ++it_ref) {
    G_PAR_FOR(it, N * NV) {
      const int k = it / NV, i = it - k * NV;
      double a = r[it] - apply_AT_entry<M>(c, dnu, k, i);
      if (i < NX) { const double* H = c.Hx + (size_t)k * NX * NX + i * NX; for (int m = 0; m < NX; ++m) a -= H[m] * dz[k * NV + m]; }
      else { const double* H = c.Hu + (size_t)k * NU * NU + (i - NX) * NU; for (int m = 0; m < NU; ++m) a -= H[m] * dz[k * NV + NX + m]; }
      c.res[it] = a;
    }
    G_PAR_FOR(it, (N + 1) * NX) { const int j = it / NX, i = it - j * NX; c.resnu[it] = rnu[it] - apply_A_entry<M>(c, dz, j, i); }
    G_SYNC();
    kkt_solve<M>(c, c.res, c.resnu, c.e, c.enu);
    G_PAR_FOR(it, N * NV) dz[it] += c.e[it];
    G_PAR_FOR(it, (N + 1) * NX) dnu[it] += c.enu[it];
    G_SYNC();
  }
}

// ------------------------------------------------------------------------------------------ factorisation
template <int M> GDEV_NOINLINE bool factorize(const IpmCtx<M>& c, const IpmParams& prm) {
  constexpr int NX = IpmCtx<M>::NX, NU = IpmCtx<M>::NU;
  const int N = c.N;
  const double hh = 0.5 * c.h;
  double bad = 0.0;
  // (1) per knot, on thread-local copies: Cholesky of Hx + dp I and its inverse Ci; inverse Cholesky Cu of Hu + dp I;
  //     GT = Theta G'
  G_PAR_FOR(k, N) {
    double C[NX * NX], Cinv[NX * NX];
    const double* Hx = c.Hx + (size_t)k * NX * NX;
    for (int i = 0; i < NX * NX; ++i) C[i] = Hx[i];
    for (int i = 0; i < NX; ++i) C[i * NX + i] += prm.delta_p;
    if (!chol_lower<NX>(C)) bad = 1.0;
    for (int j = 0; j < NX; ++j) tri_inv_col<NX>(C, j, Cinv);
    double* gCi = c.Ci + (size_t)k * NX * NX;
    for (int i = 0; i < NX * NX; ++i) gCi[i] = Cinv[i];
    double Hu[NU * NU], Cu[NU * NU];
    for (int i = 0; i < NU * NU; ++i) { Hu[i] = c.Hu[(size_t)k * NU * NU + i]; Cu[i] = 0.0; }
    for (int i = 0; i < NU; ++i) Hu[i * NU + i] += prm.delta_p;
    if (!chol_lower<NU>(Hu)) bad = 1.0;
    for (int j = 0; j < NU; ++j) tri_inv_col<NU>(Hu, j, Cu);
    double* gCu = c.Cu + (size_t)k * NU * NU;
    for (int i = 0; i < NU * NU; ++i) gCu[i] = Cu[i];
    // Theta = Cu' Cu ; GT[a][i] = sum_b Theta[a][b] * G[i][b]
    double* GT = c.GT + (size_t)k * NU * NX;
    for (int a = 0; a < NU; ++a)
      for (int i = 0; i < NX; ++i) {
        double s = 0;
        for (int b2 = 0; b2 < NU; ++b2) {
          double th = 0;
          for (int m = (a > b2 ? a : b2); m < NU; ++m) th += Cu[m * NU + a] * Cu[m * NU + b2];
          s += th * hh * c.Bm[i * NU + b2];
        }
        GT[a * NX + i] = s;
      }
  }
  G_SYNC();
  // (3) CA = h/2 * Ci * A'   (CA[m][j] = h/2 sum_q Ci[m][q] A[j][q])
  G_PAR_FOR(it, N * NX) {
    const int k = it / NX, j = it - k * NX;
    const double* Ci = c.Ci + (size_t)k * NX * NX;
    const double* Aj = c.A + (size_t)k * NX * NX + j * NX;
    double* CA = c.CA + (size_t)k * NX * NX;
    for (int m = 0; m < NX; ++m) { double s = 0; for (int q = 0; q <= m; ++q) s += Ci[m * NX + q] * Aj[q]; CA[m * NX + j] = hh * s; }
  }
  G_SYNC();
  // (4) W_LL, W_RR, W_RL column by column:  YL = Ci*Lx', YR = Ci*Rx',  W_XY = YX' YY + [flags] G Theta G'
  G_PAR_FOR(it, N * NX) {
    const int k = it / NX, j = it - k * NX;
    const double* Ci = c.Ci + (size_t)k * NX * NX;
    const double* CA = c.CA + (size_t)k * NX * NX;
    const double* GT = c.GT + (size_t)k * NU * NX;
    double* W = c.W + (size_t)k * 3 * NX * NX;
    const bool first = (k == 0), last = (k == N - 1);
    const double mj = (c.d->goal_type[j] == GOAL_POINT) ? 1.0 : 0.0;
    double ylj[NX], yrj[NX];
    for (int m = 0; m < NX; ++m) {
      ylj[m] = first ? Ci[m * NX + j] : (CA[m * NX + j] - Ci[m * NX + j]);
      yrj[m] = last ? Ci[m * NX + j] * mj : (CA[m * NX + j] + Ci[m * NX + j]);
    }
    for (int i = 0; i < NX; ++i) {
      const double mi = (c.d->goal_type[i] == GOAL_POINT) ? 1.0 : 0.0;
      double ll = 0, rr = 0, rl = 0, xi = 0;
      for (int m = 0; m < NX; ++m) {
        const double yli = first ? Ci[m * NX + i] : (CA[m * NX + i] - Ci[m * NX + i]);
        const double yri = last ? Ci[m * NX + i] * mi : (CA[m * NX + i] + Ci[m * NX + i]);
        ll += yli * ylj[m]; rr += yri * yrj[m]; rl += yri * ylj[m];
      }
      for (int a = 0; a < NU; ++a) xi += hh * c.Bm[i * NU + a] * GT[a * NX + j];
      W[0 * NX * NX + i * NX + j] = ll + (first ? 0.0 : xi);
      W[1 * NX * NX + i * NX + j] = rr + (last ? 0.0 : xi);
      W[2 * NX * NX + i * NX + j] = rl + ((first || last) ? 0.0 : xi);
    }
  }
  G_SYNC();
  // (5) block-tridiagonal Cholesky sweep over rows j = 0..N (sequential in j, parallel inside a block)
  double* D = c.sD; double* Lo = c.sLo; double* Li = c.sLi;
  for (int j = 0; j <= N; ++j) {
    G_PAR_FOR(it, NX * NX) {
      const int i = it / NX, q = it - i * NX;
      double v = 0.0;
      if (j < N) v += c.W[(size_t)j * 3 * NX * NX + it];
      if (j > 0) v += c.W[(size_t)(j - 1) * 3 * NX * NX + NX * NX + it];
      if (j == N && i == q && c.d->goal_type[i] != GOAL_POINT) v += 1.0;
      if (i == q) v += prm.delta_d * v + 1e-300;
      if (j > 0) for (int m = 0; m < NX; ++m) v -= Lo[i * NX + m] * Lo[q * NX + m];
      D[it] = v;
    }
    G_SYNC();
    for (int q = 0; q < NX; ++q) {         // right-looking Cholesky, column q
      if (G_TID == 0) { double p = D[q * NX + q]; if (!(p > 0.0)) { bad = 1.0; p = 1e-300; } D[q * NX + q] = sqrt(p); }
      G_SYNC();
      G_PAR_FOR(i, NX) if (i > q) D[i * NX + q] /= D[q * NX + q];
      G_SYNC();
      G_PAR_FOR(it, NX * NX) { const int i = it / NX, m = it - i * NX; if (m > q && i >= m) D[i * NX + m] -= D[i * NX + q] * D[m * NX + q]; }
      G_SYNC();
    }
    G_PAR_FOR(it, NX * NX) Li[it] = 0.0;
    G_SYNC();
    G_PAR_FOR(col, NX) tri_inv_col<NX>(D, col, Li);
    G_SYNC();
    double* gLd = c.Ld + (size_t)j * NX * NX;
    G_PAR_FOR(it, NX * NX) gLd[it] = Li[it];
    if (j < N) {                          // Lo_{j+1} = W_RL_j * L_jj^-T   (Lo[i][q] = sum_m W_RL[i][m] Li[q][m])
      const double* WRL = c.W + (size_t)j * 3 * NX * NX + 2 * NX * NX;
      double* gLo = c.Lo + (size_t)(j + 1) * NX * NX;
      G_SYNC();
      G_PAR_FOR(it, NX * NX) {
        const int i = it / NX, q = it - i * NX;
        double s = 0; for (int m = 0; m <= q; ++m) s += WRL[i * NX + m] * Li[q * NX + m];
        D[it] = s;                      // D is free again; stage through it so that Lo is not overwritten while read
      }
      G_SYNC();
      G_PAR_FOR(it, NX * NX) { Lo[it] = D[it]; gLo[it] = D[it]; }
    }
    G_SYNC();
  }
  bad = block_max(bad, c.red);
  return bad == 0.0;
}

// ------------------------------------------------------------------------------- assembly of H, rhs, residuals
struct Resid { double rz, rp, rc, mu, npair; };

// phase 0: build Hx, Hu and the predictor rhs (sigma*mu = 0, no second-order term); also residual norms.
// phase 1: corrector rhs with centering target `smu` and the stored predictor products.
template <int M> GDEV_NOINLINE void assemble(const IpmCtx<M>& c, int phase, double smu, Resid* out) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const int N = c.N;
  double rzmax = 0, rcmax = 0, musum = 0, npair = 0;
  G_PAR_FOR(k, N) {
    const double* x = c.z + k * NV;
    const double* u = x + NX;
    double Hx[NX * NX], Hu[NU * NU], rx[NX], ru[NU], gz[NV];
    if (phase == 0) { for (int i = 0; i < NX * NX; ++i) Hx[i] = 0; for (int i = 0; i < NU * NU; ++i) Hu[i] = 0; }
    // gradient of the Lagrangian without inequality terms: cost 2 w_k u  + A' nu
    const double wk = (k == 0 || k == N - 1) ? 0.5 * c.h : c.h;
    for (int i = 0; i < NV; ++i) gz[i] = apply_AT_entry<M>(c, c.nu, k, i);
    for (int i = 0; i < NU; ++i) { gz[NX + i] += 2.0 * wk * u[i]; if (phase == 0) Hu[i * NU + i] = 2.0 * wk; }
    for (int i = 0; i < NX; ++i) rx[i] = 0;
    for (int i = 0; i < NU; ++i) ru[i] = 0;
    for (int s = 0; s < c.S; ++s) {
      SlotEval o;
      slot_eval<M>(c, k, s, x, u, o);
      if (!o.valid) continue;
      double* st = c.slot + ((size_t)k * c.S + s) * SLOT_W;
      const double sa = st[0], la = st[1];
      double kap, bt;
      if (o.has_t) {
        const double t = st[2], lb = st[3];
        const double rca = o.c0 - t + sa, rt = c.omega - la - lb;
        const double wa = la / sa, wb = lb / t;
        const double rsa = sa * la - smu + (phase ? st[4] : 0.0), rsb = t * lb - smu + (phase ? st[5] : 0.0);
        const double ba = (la * rca - rsa) / sa, bb = -rsb / t;
        kap = wa * wb / (wa + wb);
        bt = ba - wa * (ba + bb - rt) / (wa + wb);
        if (phase == 0) {
          rcmax = fabs(rca) > rcmax ? fabs(rca) : rcmax;
          rzmax = fabs(rt) > rzmax ? fabs(rt) : rzmax;
          musum += sa * la + t * lb; npair += 2;
        }
      } else {
        const double rc = o.c0 + sa;
        const double rs = sa * la - smu + (phase ? st[4] : 0.0);
        kap = la / sa;
        bt = (la * rc - rs) / sa;
        if (phase == 0) { rcmax = fabs(rc) > rcmax ? fabs(rc) : rcmax; musum += sa * la; npair += 1; }
      }
      const int n = o.i1 - o.i0;
      double* gdst = o.is_u ? (gz + NX) : gz;
      double* rdst = o.is_u ? ru : rx;
      for (int a = 0; a < n; ++a) { gdst[o.i0 + a] += la * o.gv[a]; rdst[o.i0 + a] -= o.gv[a] * bt; }
      if (phase == 0) {
        double* H = o.is_u ? Hu : Hx;
        const int ld = o.is_u ? NU : NX;
        for (int a = 0; a < n; ++a) {
          H[(o.i0 + a) * ld + o.i0 + a] += la * o.hq[a];
          for (int b2 = 0; b2 < n; ++b2) H[(o.i0 + a) * ld + o.i0 + b2] += kap * o.gv[a] * o.gv[b2];
        }
      }
    }
    for (int i = 0; i < NX; ++i) { c.r[k * NV + i] = rx[i] - gz[i]; if (phase == 0) rzmax = fabs(gz[i]) > rzmax ? fabs(gz[i]) : rzmax; }
    for (int i = 0; i < NU; ++i) { c.r[k * NV + NX + i] = ru[i] - gz[NX + i]; if (phase == 0) rzmax = fabs(gz[NX + i]) > rzmax ? fabs(gz[NX + i]) : rzmax; }
    if (phase == 0) {
      double* gHx = c.Hx + (size_t)k * NX * NX; for (int i = 0; i < NX * NX; ++i) gHx[i] = Hx[i];
      double* gHu = c.Hu + (size_t)k * NU * NU; for (int i = 0; i < NU * NU; ++i) gHu[i] = Hu[i];
    }
  }
  if (phase == 0) {
    // equality residual r_p = A z - b  ;  rnu = -r_p
    double rpmax = 0;
    G_PAR_FOR(it, (N + 1) * NX) {
      const int j = it / NX, i = it - j * NX;
      double v = apply_A_entry<M>(c, c.z, j, i);
      if (j == 0) v -= c.x_init[i];
      else if (j == N) v -= (c.d->goal_type[i] == GOAL_POINT) ? c.goal_lo[i] : 0.0;
      else v += 0.5 * c.h * (c.g[(j - 1) * NX + i] + c.g[j * NX + i]);
      c.rnu[it] = -v;
      rpmax = fabs(v) > rpmax ? fabs(v) : rpmax;
    }
    out->rz = block_max(rzmax, c.red);
    out->rp = block_max(rpmax, c.red);
    out->rc = block_max(rcmax, c.red);
    const double ms = block_sum(musum, c.red), np = block_sum(npair, c.red);
    out->npair = np;
    out->mu = np > 0 ? ms / np : 0.0;
    if (!(out->mu == out->mu)) out->mu = 1e300;
  } else {
    G_SYNC();
  }
}

// Per-slot step from dz; mode 0: predictor (returns max steps and stores nothing), mode 1: store predictor products,
// mode 2: apply step (alpha_p, alpha_d).  Returns through amax[0..1] the largest primal/dual step to the boundary.
template <int M> GDEV_NOINLINE void slot_steps(const IpmCtx<M>& c, int phase, double smu, int mode, double ap, double ad, double* amax,
                                      double* mu_aff) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NV = L::NV;
  const int N = c.N;
  double amp = 1e300, amd = 1e300, musum = 0;
  G_PAR_FOR(it, N * c.S) {
    const int k = it / c.S, s = it - k * c.S;
    const double* x = c.z + k * NV;
    const double* u = x + NX;
    SlotEval o;
    slot_eval<M>(c, k, s, x, u, o);
    if (!o.valid) continue;
    double* st = c.slot + (size_t)it * SLOT_W;
    const double* dv = c.dz + k * NV + (o.is_u ? NX : 0) + o.i0;
    double gdz = 0;
    for (int a = 0; a < o.i1 - o.i0; ++a) gdz += o.gv[a] * dv[a];
    const double sa = st[0], la = st[1];
    if (o.has_t) {
      const double t = st[2], lb = st[3];
      const double rca = o.c0 - t + sa, rt = c.omega - la - lb;
      const double wa = la / sa, wb = lb / t;
      const double rsa = sa * la - smu + (phase ? st[4] : 0.0), rsb = t * lb - smu + (phase ? st[5] : 0.0);
      const double ba = (la * rca - rsa) / sa, bb = -rsb / t;
      const double dt = (wa * gdz + ba + bb - rt) / (wa + wb);
      const double dla = wa * (gdz - dt) + ba, dlb = -wb * dt + bb, ds = -rca - (gdz - dt);
      if (mode == 2) {
        st[0] = sa + ap * ds; st[2] = t + ap * dt; st[1] = la + ad * dla; st[3] = lb + ad * dlb;
      } else {
        if (ds < 0) { const double a = -sa / ds; amp = a < amp ? a : amp; }
        if (dt < 0) { const double a = -t / dt; amp = a < amp ? a : amp; }
        if (dla < 0) { const double a = -la / dla; amd = a < amd ? a : amd; }
        if (dlb < 0) { const double a = -lb / dlb; amd = a < amd ? a : amd; }
        if (mode == 1) { st[4] = ds * dla; st[5] = dt * dlb; musum += (sa + ap * ds) * (la + ap * dla) + (t + ap * dt) * (lb + ap * dlb); }
      }
    } else {
      const double rc = o.c0 + sa;
      const double rs = sa * la - smu + (phase ? st[4] : 0.0);
      const double wa = la / sa, ba = (la * rc - rs) / sa;
      const double dla = wa * gdz + ba, ds = -rc - gdz;
      if (mode == 2) { st[0] = sa + ap * ds; st[1] = la + ad * dla; }
      else {
        if (ds < 0) { const double a = -sa / ds; amp = a < amp ? a : amp; }
        if (dla < 0) { const double a = -la / dla; amd = a < amd ? a : amd; }
        if (mode == 1) { st[4] = ds * dla; st[5] = 0.0; musum += (sa + ap * ds) * (la + ap * dla); }
      }
    }
  }
  if (mode != 2) {
    amax[0] = -block_max(-amp, c.red);
    amax[1] = -block_max(-amd, c.red);
    if (mode == 1) *mu_aff = block_sum(musum, c.red);
  } else {
    G_SYNC();
  }
}

// Keep every complementarity pair above 1e-4 * mu (wide neighbourhood of the central path), as the oracle does.
template <int M> GDEV_NOINLINE void recenter(const IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NV = L::NV;
  const int N = c.N;
  double musum = 0, npair = 0;
  G_PAR_FOR(it, N * c.S) {
    const int k = it / c.S, s = it - k * c.S;
    SlotEval o;
    slot_eval<M>(c, k, s, c.z + k * NV, c.z + k * NV + NX, o);
    if (!o.valid) continue;
    const double* st = c.slot + (size_t)it * SLOT_W;
    musum += st[0] * st[1]; npair += 1;
    if (o.has_t) { musum += st[2] * st[3]; npair += 1; }
  }
  const double ms = block_sum(musum, c.red), np = block_sum(npair, c.red);
  const double floor_ = np > 0 ? 1e-4 * ms / np : 0.0;
  G_PAR_FOR(it, N * c.S) {
    const int k = it / c.S, s = it - k * c.S;
    SlotEval o;
    slot_eval<M>(c, k, s, c.z + k * NV, c.z + k * NV + NX, o);
    if (!o.valid) continue;
    double* st = c.slot + (size_t)it * SLOT_W;
    if (st[0] * st[1] < floor_) st[1] = floor_ / st[0];
    if (o.has_t && st[2] * st[3] < floor_) st[3] = floor_ / st[2];
  }
  G_SYNC();
}

// ------------------------------------------------------------------------------------------------ driver
// scratch: IpmLayout<M>::scratch_doubles() doubles of global memory owned by this instance.
// smem:    IpmLayout<M>::smem_doubles() doubles of shared memory.
// On exit Xn/Un of the instance hold the solution and info[IPM_NINFO] = {status, iters, res, mu, obj,...}.
template <int M>
GDEV void ipm_solve_instance(const BatchDesc& d, const BatchPtrs& p, const IpmParams& prm, int b, double* scratch,
                             double* smem, double* info) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const int N = d.N;
  IpmCtx<M> c;
  c.d = &d; c.N = N; c.n_obs = (Traits<M>::WS > 0) ? d.n_obs : 0; c.S = L::nslots(c.n_obs); c.b = b;
  c.h = p.tf[b] / (N - 1); c.omega = p.omega[b]; c.Delta = p.delta[b];
  c.toggle = c.Delta / 8.0 + d.rp[RP_CLEAR]; c.eps = d.sp[SP_EPS];
  c.Xp = p.Xp + (size_t)b * N * NX; c.Up = p.Up + (size_t)b * N * NU;
  c.A = p.A + (size_t)b * N * NX * NX; c.g = p.g + (size_t)b * N * NX;
  c.rows = p.rows + (size_t)b * N * d.n_obs * 5;
  c.x_init = p.x_init + (size_t)b * NX; c.goal_lo = p.goal_lo + (size_t)b * NX; c.goal_hi = p.goal_hi + (size_t)b * NX;
  for (int i = 0; i < NX * NU; ++i) c.Bm[i] = 0.0;
  dyn_B<M>(d.rp, c.Bm);
  const size_t nz = (size_t)N * NV, ne = (size_t)(N + 1) * NX;
  double* q = scratch;
  c.z = q; q += nz; c.r = q; q += nz; c.dz = q; q += nz; c.t1 = q; q += nz; c.res = q; q += nz; c.e = q; q += nz;
  c.nu = q; q += ne; c.rnu = q; q += ne; c.dnu = q; q += ne; c.resnu = q; q += ne; c.enu = q; q += ne;
  c.slot = q; q += (size_t)N * c.S * SLOT_W;
  c.Hx = q; q += (size_t)N * NX * NX; c.Ci = q; q += (size_t)N * NX * NX; c.CA = q; q += (size_t)N * NX * NX;
  c.W = q; q += (size_t)N * NX * NX * 3;
  c.Hu = q; q += (size_t)N * NU * NU; c.Cu = q; q += (size_t)N * NU * NU; c.GT = q; q += (size_t)N * NU * NX;
  c.Ld = q; q += (size_t)(N + 1) * NX * NX; c.Lo = q; q += (size_t)(N + 1) * NX * NX;
  c.sy = smem; c.sD = c.sy + (N + 1) * NX; c.sLo = c.sD + NX * NX; c.sLi = c.sLo + NX * NX; c.sB4 = c.sLi + NX * NX;
  c.stmp = c.sB4 + NX * NX; c.red = c.stmp + 16;

  // ---- start point: X, U <- previous trajectory (set_start_value, scp_gusto.jl:100-102); slacks one unit inside
  G_PAR_FOR(it, N * NV) { const int k = it / NV, i = it - k * NV; c.z[it] = i < NX ? c.Xp[k * NX + i] : c.Up[k * NU + i - NX]; }
  G_PAR_FOR(it, (N + 1) * NX) c.nu[it] = 0.0;
  G_SYNC();
  G_PAR_FOR(it, N * c.S) {
    const int k = it / c.S, s = it - k * c.S;
    SlotEval o;
    slot_eval<M>(c, k, s, c.z + k * NV, c.z + k * NV + NX, o);
    double* st = c.slot + (size_t)it * SLOT_W;
    for (int i = 0; i < SLOT_W; ++i) st[i] = 0.0;
    if (!o.valid) continue;
    if (o.has_t) {
      const double t = (o.c0 > 0 ? o.c0 : 0.0) + 1.0;
      const double sa = t - o.c0;
      st[0] = sa > 1e-2 ? sa : 1e-2; st[1] = 0.5 * c.omega; st[2] = t; st[3] = 0.5 * c.omega;
    } else {
      st[0] = -o.c0 > 1e-2 ? -o.c0 : 1e-2; st[1] = 1e-2;
    }
  }
  G_SYNC();

  int status = IPM_ITERATION_LIMIT, it_done = 0;
  long long cyc_asm = 0, cyc_fac = 0, cyc_sol = 0, cyc_slot = 0, tc0;
  double res = 1e300, mu = 0;
  const double scd = 1.0 + c.omega;
  for (int iter = 1; iter <= prm.max_iter; ++iter) {
    it_done = iter;
    Resid R;
    tc0 = g_clock();
    assemble<M>(c, 0, 0.0, &R);
    cyc_asm += g_clock() - tc0;
    mu = R.mu;
    res = R.rz / scd;
    res = R.rp > res ? R.rp : res; res = R.rc > res ? R.rc : res; res = mu > res ? mu : res;
#ifdef GUSTO_HOSTSIM
    if (getenv("GUSTO_HOSTSIM_VERBOSE")) printf("  ipm %3d rd=%.2e rp=%.2e rc=%.2e mu=%.2e\n", iter, R.rz, R.rp, R.rc, mu);
#endif
    if (res <= prm.tol) { status = IPM_OPTIMAL; break; }
    if (!(res == res) || res > 1e200) { status = IPM_NUMERICAL; break; }
    tc0 = g_clock();
    const bool fac_ok = factorize<M>(c, prm);
    cyc_fac += g_clock() - tc0;
    if (!fac_ok) {
#ifdef GUSTO_HOSTSIM
      if (getenv("GUSTO_HOSTSIM_VERBOSE")) printf("  factorize: non-positive pivot\n");
#endif
    }
    // predictor
    tc0 = g_clock();
    kkt_solve_refined<M>(c, c.r, c.rnu, c.dz, c.dnu, 0);
    cyc_sol += g_clock() - tc0;
    double am[2], mu_aff = 0;
    tc0 = g_clock();
    slot_steps<M>(c, 0, 0.0, 0, 0, 0, am, &mu_aff);
    double a_aff = am[0] < am[1] ? am[0] : am[1];
    a_aff = a_aff < 1.0 ? a_aff : 1.0;
    slot_steps<M>(c, 0, 0.0, 1, a_aff, a_aff, am, &mu_aff);
    cyc_slot += g_clock() - tc0;
    mu_aff = R.npair > 0 ? mu_aff / R.npair : 0.0;
    double sigma = mu > 0 ? (mu_aff / mu) : 0.0;
    sigma = sigma * sigma * sigma;
    double smu = sigma * mu;
    smu = smu > 0.1 * prm.tol ? smu : 0.1 * prm.tol;
    // corrector
    tc0 = g_clock();
    assemble<M>(c, 1, smu, &R);
    cyc_asm += g_clock() - tc0;
    tc0 = g_clock();
    kkt_solve_refined<M>(c, c.r, c.rnu, c.dz, c.dnu, prm.nref);
    cyc_sol += g_clock() - tc0;
    tc0 = g_clock();
    slot_steps<M>(c, 1, smu, 0, 0, 0, am, &mu_aff);
    double tau = 0.995;
    if (mu < 1.0) { tau = 1.0 - mu; tau = tau > 0.995 ? tau : 0.995; tau = tau < 0.999999 ? tau : 0.999999; }
    double ap = tau * am[0], ad = tau * am[1];
    ap = ap < 1.0 ? ap : 1.0; ad = ad < 1.0 ? ad : 1.0;
    slot_steps<M>(c, 1, smu, 2, ap, ad, am, &mu_aff);
    G_PAR_FOR(it, N * NV) c.z[it] += ap * c.dz[it];
    G_PAR_FOR(it, (N + 1) * NX) c.nu[it] += ad * c.dnu[it];
    G_SYNC();
    recenter<M>(c);
    cyc_slot += g_clock() - tc0;
  }
  if (status == IPM_ITERATION_LIMIT && res <= 1e3 * prm.tol) status = IPM_OPTIMAL;
  {   // a NaN/Inf anywhere in the iterate is a numerical failure, never an answer
    double badz = 0.0;
    G_PAR_FOR(it, N * NV) { const double v = c.z[it]; if (!(v == v) || fabs(v) > 1e100) badz = 1.0; }
    if (block_max(badz, c.red) > 0.0) status = IPM_NUMERICAL;
  }
  // ---- write the candidate trajectory and the objective (cost + omega * sum t)
  double obj = 0;
  G_PAR_FOR(k, N) {
    const double wk = (k == 0 || k == N - 1) ? 0.5 * c.h : c.h;
    for (int i = 0; i < NX; ++i) p.Xn[((size_t)b * N + k) * NX + i] = c.z[k * NV + i];
    for (int i = 0; i < NU; ++i) { const double uv = c.z[k * NV + NX + i]; p.Un[((size_t)b * N + k) * NU + i] = uv; obj += wk * uv * uv; }
  }
  G_PAR_FOR(it, N * c.S) {
    const int k = it / c.S, s = it - k * c.S;
    SlotEval o;
    slot_eval<M>(c, k, s, c.z + k * NV, c.z + k * NV + NX, o);
    if (o.valid && o.has_t) obj += c.omega * c.slot[(size_t)it * SLOT_W + 2];
  }
  obj = block_sum(obj, c.red);
  if (G_TID == 0) {
    info[0] = (double)status; info[1] = (double)it_done; info[2] = res; info[3] = mu; info[4] = obj;
    info[5] = (double)(cyc_asm + cyc_slot); info[6] = (double)cyc_fac; info[7] = (double)cyc_sol;   // SM cycles per phase
  }
}

}  // namespace gusto
