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
//     leaving a block-diagonal reduced Hessian  H = blkdiag(Hx_k, Hu_k);
//   * inside a knot every row except the state trust region lives in ONE coordinate block (position | velocity |
//     attitude | rate, Traits<M>::XB_*), so  Hx_k = blockdiag(<=4x4 blocks) + kappa_tr g g'  and its inverse is
//     blockdiag(P_b) - coef w w'  (Sherman-Morrison); nothing larger than 4x4 is ever factorised per knot;
//   * A_k = df/dx is used through its static sparsity pattern (Traits<M>::a_row/a_col, 27 of 144 entries for SE3);
//   * the equality rows (init, trapezoid dynamics, point goal) are block-bidiagonal, so the Schur complement
//     S = Aeq (H + dp I)^-1 Aeq' is block-tridiagonal with N+1 blocks of NX x NX.  It is factorised as a block
//     L D L' with explicit D_j^-1 (Gauss-Jordan in shared memory on one warp while the other warp prefetches the
//     next block row), so each solve is two chains of N dependent NX x NX mat-vecs streamed through a TMA
//     (cp.async.bulk + mbarrier) ring plus one fully parallel D^-1 pass;
//   * Newton directions are recovered with `nref` steps of iterative refinement against the unregularised KKT
//     matrix (H alone is only positive SEMI-definite: the cost has no state term).
// A first-order splitting (ADMM, prototyped in tools/admm_proto.py) was rejected: GuSTO's accept test compares
// soft rows against eps = 1e-6 (astrobee_se3.jl:31, scp_gusto.jl:318-327), and with a trapezoid double
// integrator over 70 s ADMM needs >2000 iterations for 1e-6 residuals while this method needs 8-25 for 1e-8.
//
// Memory: the iterate z, the direction dz and the Schur right-hand side live in shared memory; multipliers, slack
// records (compacted to the obstacle rows inside the toggle distance), per-knot block inverses and the
// block-tridiagonal factor live in a per-instance global scratch that stays L2-resident.
// All arithmetic is FP64 (the reference is Float64 throughout).
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "evaluate.cuh"   // block_sum / block_max
#ifdef GUSTO_HOSTSIM
#include <cstdio>
#include <cstdlib>
#endif

// ---- instance groups.  One instance is solved by a GROUP of GUSTO_IPM_GROUP threads (2 warps).  A CTA packs several groups
// (one per instance, capi.cu chooses how many): inside this file the SPMD macros address the group, not the CTA --
// G_TID / G_NTHR are the thread's index in / the size of its group, G_SYNC is the group's own named barrier
// (bar.sync id, 64).  Why: the kernel is instruction-fetch bound (stall_no_instruction is its largest stall, ~280 KB of
// SASS cycled through every Newton iteration).  Seven instances as seven groups of ONE 448-thread CTA per SM run 14 %
// faster than as seven 64-thread CTAs (measured, 7.67 -> 6.60 ms on the headline batch): the group size is a compile-time
// constant and the groups of a CTA start together.  An explicit CTA-wide re-alignment barrier per Newton iteration
// (G_CTA_RESYNC, -DGUSTO_IPM_RESYNC) was measured too: 7.11 ms with one per iteration, 7.54 ms with four -- the waiting
// costs more than the shared fetches save, so it is off.
#ifndef GUSTO_IPM_GROUP
#define GUSTO_IPM_GROUP 64
#endif
#ifndef GUSTO_HOSTSIM
#pragma push_macro("G_TID")
#pragma push_macro("G_NTHR")
#pragma push_macro("G_SYNC")
#undef G_TID
#undef G_NTHR
#undef G_SYNC
#define G_TID ((int)(threadIdx.x & (GUSTO_IPM_GROUP - 1)))
#define G_NTHR GUSTO_IPM_GROUP
#define G_SYNC() gusto::g_group_sync()
#ifdef GUSTO_IPM_RESYNC
#define G_CTA_RESYNC() __syncthreads()      /* exited groups no longer count */
#else
#define G_CTA_RESYNC() ((void)0)
#endif
#define block_sum ipm_group_sum
#define block_max ipm_group_max
#else
#define G_CTA_RESYNC() ((void)0)
#endif

namespace gusto {

#ifdef GUSTO_HOSTSIM
GDEV long long g_clock() { return 0; }
#else
GDEV long long g_clock() { return clock64(); }
__device__ __forceinline__ void g_group_sync() {
  asm volatile("bar.sync %0, %1;" ::"r"((int)(threadIdx.x / GUSTO_IPM_GROUP) + 1), "n"(GUSTO_IPM_GROUP) : "memory");
}
// group-wide reductions (see block_sum / block_max in evaluate.cuh); red: one slot per warp of the group
__device__ __noinline__ double ipm_group_sum(double v, double* red) {
  G_ASSUME_SHARED(red);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  constexpr int nw = GUSTO_IPM_GROUP / 32;
  if (G_LANE == 0) red[G_TID >> 5] = v;
  g_group_sync();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < nw; ++i) s += red[i];
  g_group_sync();
  return s;
}
__device__ __noinline__ double ipm_group_max(double v, double* red) {   // NaN-propagating
  G_ASSUME_SHARED(red);
  v = (v == v) ? v : 1e300;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
  constexpr int nw = GUSTO_IPM_GROUP / 32;
  if (G_LANE == 0) red[G_TID >> 5] = v;
  g_group_sync();
  double s = red[0];
#pragma unroll
  for (int i = 1; i < nw; ++i) s = red[i] > s ? red[i] : s;
  g_group_sync();
  return s;
}
#endif

#if defined(GUSTO_PROF_MODE) && GUSTO_PROF_MODE == 3
#define GUSTO_PROF_SOLVE 0
#else
#define GUSTO_PROF_SOLVE 1
#endif

struct IpmParams {
  int max_iter;      // Newton iterations cap
  int nref;          // refinement steps on the corrector solve
  double tol;        // max(|r_dual|/(1+omega), |r_eq|, |r_ineq|, mu) <= tol
  double delta_p;    // primal regularisation added to Hx, Hu in the factorisation only
  double delta_d;    // relative regularisation of the Schur diagonal
};

enum : int { IPM_OPTIMAL = 0, IPM_ITERATION_LIMIT = 1, IPM_NUMERICAL = 2 };
constexpr int SLOT_W = 6;     // s, lam, t, lamb, pa (ds*dlam of the predictor), pb (dt*dlamb of the predictor)
constexpr int OROW_W = 5;     // compacted obstacle row: nhat[3], off, knot
constexpr int IPM_NINFO = 8;  // status, iterations, residual, mu, objective, cycles: assemble+slots, factorize, kkt solves
#ifndef GUSTO_RING_STAGES
#define GUSTO_RING_STAGES 8
#endif
constexpr int RING_STAGES = GUSTO_RING_STAGES;

GHD constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

template <int M> struct IpmLayout {
  using T = Traits<M>;
  static constexpr int NX = T::NX, NU = T::NU, NV = NX + NU, NN = NX * NX, ANZ = T::ANZ;
  static constexpr int GLD = (NX + 1) & ~1;               // row stride of a factor tile in global memory (16-byte rows)
  static constexpr int GT = NX * GLD;
  static constexpr int LDT = 2 * (((NX + 1) / 2) | 1);     // row stride of a shared-memory tile: even, LDT/2 odd -> LDS.128 conflict-free
  static constexpr int TILE = NX * LDT;
  static constexpr int CG = NX <= 12 ? 3 : 4;              // output columns per thread task in the tile products
  static constexpr int NG = (NX + CG - 1) / CG;
  static constexpr int NTASK = NX * NG;                    // tasks of a full product
  GHD static constexpr int nlt() { int n = 0; for (int i = 0; i < NX; ++i) for (int g = 0; g < NG; ++g) if (g * CG <= i) ++n; return n; }
  static constexpr int NLT = nlt();                        // tasks touching the lower triangle
  // special (non-obstacle) slots of a knot
  static constexpr int S_TR = 0;
  static constexpr int S_NORM = T::HAS_TR;
  static constexpr int S_LIN = S_NORM + T::NNORM;
  static constexpr int S_QUAT = S_LIN + T::NLIN;           // hinge, then hard row
  static constexpr int S_BALL = S_QUAT + 2 * T::HAS_QUAT;
  static constexpr int SP = S_BALL + T::NBALL;
  static constexpr int NBOX = 2 * NX;                      // goal-box rows (upper, lower per coordinate), knot N-1 only
  static constexpr int XPK = T::XB_pk(T::XB_CNT), UPK = T::UB_pk(T::UB_CNT);
  // per-knot record: Hb[XPK] P[XPK] w[NX] Hub[UPK] Th[UPK] kap coef
  static constexpr int KD_HB = 0, KD_P = XPK, KD_W = 2 * XPK, KD_HU = KD_W + NX, KD_TH = KD_HU + UPK,
                       KD_KAP = KD_TH + UPK, KD_COEF = KD_KAP + 1, KDW = KD_COEF + 1;
  GHD static size_t rnd(size_t v) { return (v + 1) & ~(size_t)1; }   // keep every array 16-byte aligned
  // doubles of global scratch per instance
  GHD static size_t scratch_doubles(int N, int n_obs) {
    const size_t nz = rnd((size_t)N * NV), ne = rnd((size_t)(N + 1) * NX), no = T::WS > 0 ? n_obs : 0;
    return 3 * nz + 4 * ne + rnd((size_t)N * ANZ) + rnd((size_t)N * SP * SLOT_W) + (size_t)NBOX * SLOT_W +
           rnd((size_t)N * no * SLOT_W) + rnd((size_t)N * no * OROW_W) + rnd((size_t)N * KDW) + (size_t)(N + 1) * 2 * GT;
  }
  GHD static int work_doubles(int N) {      // dz | sy | ring, also the 8 factorisation tiles
    const int ne = (int)rnd((size_t)(N + 1) * NX);
    const int ring = RING_STAGES * GT > ne ? RING_STAGES * GT : ne;      // ring stages are plain copies of the global tiles
    const int w = (int)rnd((size_t)N * NV) + ne + ring, f = FAC_TILES * TILE + 2 * KS + 2 * NX + 4;
    return (w > f ? w : f) + 2;
  }
  // byte tables: (row, col) of the packed lower triangle, (row, group) of the lower-triangular product tasks, then the
  // model tables used with run-time indices: a_row[ANZ], a_col[ANZ], blk_of[NX], ctrl_of[NX]
  static constexpr int TAB_MODEL = NX * (NX + 1) + 2 * NLT;
  static constexpr int TAB_DOUBLES = (TAB_MODEL + 2 * ANZ + 2 * NX + 7) / 8;
  // factor sweep: 10 sweep tiles + Phi, Ah, Y, Z, RR of the Schur-block producer; staged knot record
  static constexpr int FAC_TILES = 15;
  static constexpr int KS_P = 0, KS_W = XPK, KS_TH = XPK + NX, KS_COEF = KS_TH + UPK, KS_A = KS_COEF + 1, KS = (KS_A + ANZ + 1) & ~1;
  GHD static int seg_doubles(int N) { return (N + 4) / 2 + 2; }
  static constexpr int CTX_DOUBLES = 80;                   // the per-instance context struct (IpmCtx) lives in shared memory too
  GHD static int smem_doubles(int N, int nthr) { return CTX_DOUBLES + (int)rnd((size_t)N * NV) + work_doubles(N) + nthr + 16 + seg_doubles(N) + TAB_DOUBLES; }
};

template <int M> struct IpmCtx {
  using L = IpmLayout<M>;
  static constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const BatchDesc* d;
  const double* rp;
  int N, n_obs, b, nact, pmask, bmask;
  double h, hh, omega, Delta, toggle, eps, dd;
  mutable double dp;       // primal regularisation (raised x100 after a failed factorisation)
  double dow, eow;         // Delta / omega, eps / omega
  mutable double floor_;   // pending central-path floor of the complementarity pairs (see pair_floor)
  const double *Xp, *Up, *A, *g, *rows, *x_init, *goal_lo, *goal_hi;
  double bv[NU];          // B has one entry per column: B[b_row(a)][a] = bv[a]
  // global scratch
  double *nu, *dnu, *r, *rnu, *t1, *res, *resnu, *Ac, *sslot, *bslot, *ost, *orow, *kd, *fac;
  // shared
  double *z, *dz, *sy, *ring, *red;
  int* seg;
  mutable unsigned long long ring_bar[RING_STAGES];   // one mbarrier per ring stage (TMA completion)
  mutable unsigned ring_o;                             // tiles fetched so far (stage = o % S, wait parity = (o / S) & 1)
  mutable long long prof[5];      // thread-0 cycle counters: schur rows, factor sweep, forward chain, middle pass, backward chain
  unsigned char* tab;     // [2][NX(NX+1)/2]: row / column of packed-lower entry t; then [2][NLT]: row / group of lower task
};

// shared-memory members, with the address space made known to the compiler
template <int M> GDEV double* sh_z(const IpmCtx<M>& c) { double* p = c.z; G_ASSUME_SHARED(p); return p; }
template <int M> GDEV double* sh_dz(const IpmCtx<M>& c) { double* p = c.dz; G_ASSUME_SHARED(p); return p; }
template <int M> GDEV double* sh_sy(const IpmCtx<M>& c) { double* p = c.sy; G_ASSUME_SHARED(p); return p; }
template <int M> GDEV int* sh_seg(const IpmCtx<M>& c) { return c.seg; }   // (an address-space assumption on this one miscompiles with nvcc 12.9)

// ------------------------------------------------------------------------------------------- slot algebra
// One inequality  c0(z) [- t] + s = 0, s >= 0 (multiplier lam) [, t >= 0 (multiplier lamb), cost omega*t].
// After eliminating (s, lam [, t, lamb]) the row contributes  kap * gv gv' + lam * hess  to H and  -gv * bt  to the rhs.
struct Pair { double rc, rt, wa, wb, ba, bb, iw, kap, bt, la; };

GDEV void pair_eval(const double* st, bool has_t, double c0, double omega, double smu, int phase, Pair& q) {
  const double sa = st[0], la = st[1];
  const double isa = g_rcp(sa);
  q.la = la;
  q.wa = la * isa;
  if (has_t) {
    const double t = st[2], lb = st[3];
    const double it = g_rcp(t);
    q.rc = c0 - t + sa; q.rt = omega - la - lb;
    q.wb = lb * it;
    const double rsa = sa * la - smu + (phase ? st[4] : 0.0), rsb = t * lb - smu + (phase ? st[5] : 0.0);
    q.ba = (la * q.rc - rsa) * isa; q.bb = -rsb * it;
    q.iw = g_rcp(q.wa + q.wb);
    q.kap = q.wa * q.wb * q.iw;
    q.bt = q.ba - q.wa * (q.ba + q.bb - q.rt) * q.iw;
  } else {
    q.rc = c0 + sa; q.rt = 0.0; q.wb = 0.0; q.bb = 0.0; q.iw = 0.0;
    const double rs = sa * la - smu + (phase ? st[4] : 0.0);
    q.ba = (la * q.rc - rs) * isa;
    q.kap = q.wa;
    q.bt = q.ba;
  }
}

struct Stat { double rz, rc, mus, np; };   // running max |dual residual|, max |row residual|, sum s*lam, #pairs
GDEV void pair_stat(const double* st, bool has_t, const Pair& q, Stat& S) {
  S.rc = fabs(q.rc) > S.rc ? fabs(q.rc) : S.rc;
  if (!(q.rc == q.rc)) S.rc = 1e300;
  S.mus += st[0] * st[1]; S.np += 1.0;
  if (has_t) { S.rz = fabs(q.rt) > S.rz ? fabs(q.rt) : S.rz; S.mus += st[2] * st[3]; S.np += 1.0; }
}

// Step of one row given gdz = gv . dz.
//   mode 0: largest steps to the boundary;
//   mode 1: same + store the predictor products ds*dlam and the coefficients of  sum (s + a ds)(lam + a dlam) = c0 + a c1 + a^2 c2
//           (Mehrotra's mu_aff for any step a, so the predictor needs ONE pass over the rows);
//   mode 2: apply (ap, ad) and accumulate the new complementarity sum / pair count (for the central-path floor).
struct StepAcc { double amp, amd, c0, c1, c2, np; };
GDEV void pair_step(double* st, bool has_t, const Pair& q, double gdz, int mode, double ap, double ad, StepAcc& a) {
  const double sa = st[0], la = st[1];
  if (has_t) {
    const double t = st[2], lb = st[3];
    const double dt = (q.wa * gdz + q.ba + q.bb - q.rt) * q.iw;
    const double dla = q.wa * (gdz - dt) + q.ba, dlb = -q.wb * dt + q.bb, ds = -q.rc - (gdz - dt);
    if (mode == 2) {
      const double s1 = sa + ap * ds, t1 = t + ap * dt, l1 = la + ad * dla, b1 = lb + ad * dlb;
      st[0] = s1; st[2] = t1; st[1] = l1; st[3] = b1;
      a.c0 += s1 * l1 + t1 * b1; a.np += 2.0;
    } else {
      if (ds < 0) { const double v = -sa * g_rcp(ds); a.amp = v < a.amp ? v : a.amp; }
      if (dt < 0) { const double v = -t * g_rcp(dt); a.amp = v < a.amp ? v : a.amp; }
      if (dla < 0) { const double v = -la * g_rcp(dla); a.amd = v < a.amd ? v : a.amd; }
      if (dlb < 0) { const double v = -lb * g_rcp(dlb); a.amd = v < a.amd ? v : a.amd; }
      if (mode == 1) {
        st[4] = ds * dla; st[5] = dt * dlb;
        a.c0 += sa * la + t * lb; a.c1 += sa * dla + la * ds + t * dlb + lb * dt; a.c2 += ds * dla + dt * dlb;
      }
    }
  } else {
    const double dla = q.wa * gdz + q.ba, ds = -q.rc - gdz;
    if (mode == 2) {
      const double s1 = sa + ap * ds, l1 = la + ad * dla;
      st[0] = s1; st[1] = l1;
      a.c0 += s1 * l1; a.np += 1.0;
    } else {
      if (ds < 0) { const double v = -sa * g_rcp(ds); a.amp = v < a.amp ? v : a.amp; }
      if (dla < 0) { const double v = -la * g_rcp(dla); a.amd = v < a.amd ? v : a.amd; }
      if (mode == 1) { st[4] = ds * dla; st[5] = 0.0; a.c0 += sa * la; a.c1 += sa * dla + la * ds; a.c2 += ds * dla; }
    }
  }
}
// Keep a complementarity pair above floor = 1e-4 * mu (wide neighbourhood of the central path, as the oracle does).
// Applied lazily by the first reader of the row in the next Newton iteration (assemble, phase 0).
GDEV void pair_floor(double* st, bool has_t, double floor_) {
  if (st[0] * st[1] < floor_) st[1] = floor_ * g_rcp(st[0]);
  if (has_t && st[2] * st[3] < floor_) st[3] = floor_ * g_rcp(st[2]);
}

GDEV void slot_init(double* st, bool valid, bool has_t, double c0, double omega, double t_in = 1.0, double lam_split = 0.5) {
  for (int i = 0; i < SLOT_W; ++i) st[i] = 0.0;
  if (!valid) return;
  if (has_t) {
    // interior start of the penalty slack: t_in inside (the oracle uses one unit).  0.25 saves 0.85 Newton iterations of 8.85 on
    // astrobeeSE3 but costs restarts on freeflyerSE2 at omega = 25 (measured), so it is a per-model setting
    // (slack_start<M>()), like the refinement count; the optimum reached is the same to the solver tolerance.
    const double t = (c0 > 0 ? c0 : 0.0) + t_in;
    const double sa = t - c0;
    st[0] = sa > 1e-2 ? sa : 1e-2; st[1] = lam_split * omega; st[2] = t; st[3] = omega - st[1];
  } else {
    // hard rows: the oracle's s = max(-c, 1e-2), except that a strictly feasible row keeps its exact slack -- a BoxGoal of
    // width 2e-4 (astrobeeSE3manifold notebook) would otherwise start 100x outside its own width on both sides
    st[0] = -c0 > 1e-2 ? -c0 : (-c0 > 1e-6 ? -c0 : 1e-2); st[1] = 1e-2;
  }
}

// ------------------------------------------------------------------------------------------- special slots
// The rows of a knot other than trust region / obstacles / goal box: each lives on <= 4 coordinates [i0, i0+n) of x
// (or u for the control balls), inside one Hessian block.
struct SpecEval {
  bool valid, is_u, has_t;
  int i0, n;
  double c0;
  double gv[4], hq[4];     // gradient and diagonal of the constraint Hessian on [i0, i0+n)
};

template <int M> GHD constexpr bool spec_is_u(int s) { return s >= IpmLayout<M>::S_BALL; }
template <int M> GHD constexpr int spec_i0(int s) {
  using L = IpmLayout<M>;
  return s < L::S_LIN ? norm_i0<M>(s - L::S_NORM) : s < L::S_QUAT ? lin_i<M>(s - L::S_LIN) : s < L::S_BALL ? 6 : ball_i0<M>(s - L::S_BALL);
}
template <int M> GHD constexpr int spec_block(int s) {      // Hessian block (x-blocks, or u-blocks for the balls)
  return spec_is_u<M>(s) ? Traits<M>::UB_of(spec_i0<M>(s)) : Traits<M>::XB_of(spec_i0<M>(s));
}

// s in [S_NORM, SP); x, u: the knot's state/control; xp: previous state of the knot.
template <int M>
GDEV void spec_eval(const IpmCtx<M>& c, int k, int s, const double* x, const double* u, SpecEval& o) {
  using L = IpmLayout<M>;
  const double* rp = c.rp;
  o.valid = true; o.is_u = false; o.has_t = true; o.i0 = 0; o.n = 0; o.c0 = 0.0;
  for (int a = 0; a < 4; ++a) { o.gv[a] = 0.0; o.hq[a] = 0.0; }
  if (s < L::S_LIN) {
    // csi_translational_velocity_bound / csi_angular_velocity_bound: |x[i0:i1)|^2 - lim^2 - t <= 0
    int i0, i1; double lim;
    norm_row<M>(s - L::S_NORM, rp, &i0, &i1, &lim);
    o.i0 = i0; o.n = i1 - i0;
    double v = -lim * lim;
    for (int a = 0; a < 4; ++a) if (a < i1 - i0) { o.gv[a] = 2.0 * x[i0 + a]; o.hq[a] = 2.0; v += x[i0 + a] * x[i0 + a]; }
    o.c0 = v;
  } else if (s < L::S_QUAT) {
    int i; double sign, bound;
    lin_row<M>(s - L::S_LIN, rp, &i, &sign, &bound);
    o.i0 = i; o.n = 1; o.gv[0] = sign; o.c0 = sign * x[i] - bound;
  } else if (s < L::S_BALL) {
    // cse_quaternion_norm (astrobee_se3_manifold.jl:308-313): e = a.q - 1, a = qp/|qp|.
    //   hinge  e - eps/omega - t <= 0 (t >= 0)       and the hard row  -e - eps/omega <= 0   
    const double* qp = c.Xp + k * L::NX + 6;
    const double nq = sqrt(qp[0] * qp[0] + qp[1] * qp[1] + qp[2] * qp[2] + qp[3] * qp[3]);
    double ev = -1.0;
    for (int i = 0; i < 4; ++i) ev += qp[i] / nq * x[6 + i];
    const bool hinge = (s == L::S_QUAT);
    o.i0 = 6; o.n = 4; o.has_t = hinge;
    for (int i = 0; i < 4; ++i) o.gv[i] = (hinge ? 1.0 : -1.0) * qp[i] / nq;
    o.c0 = (hinge ? ev : -ev) - c.eow;
  } else {
    // control balls cover k = 1..N-1 only (astrobee_se3.jl:370-371, quirk q3)
    o.is_u = true; o.has_t = false;
    int i0, i1; double scale[3], rad;
    ctrl_ball<M>(s - L::S_BALL, rp, &i0, &i1, scale, &rad);
    o.i0 = i0; o.n = i1 - i0;
    if (k >= c.N - 1) { o.valid = false; return; }
    double v = -rad * rad;
    for (int a = 0; a < 3; ++a) if (a < i1 - i0) {
      const double s2 = scale[a] * scale[a];
      o.gv[a] = 2.0 * s2 * u[i0 + a]; o.hq[a] = 2.0 * s2; v += s2 * u[i0 + a] * u[i0 + a];
    }
    o.c0 = v;
  }
}

// stri_state_trust_region (astrobee_se3.jl:308-311) in slack-scaled form: |x - xp|^2 - Delta/omega - t <= 0
template <int M> GDEV double tr_c0(const IpmCtx<M>& c, int k, const double* x) {
  constexpr int NX = IpmCtx<M>::NX;
  double v = -c.dow;
  for (int i = 0; i < NX; ++i) { const double dxi = x[i] - c.Xp[k * NX + i]; v += dxi * dxi; }
  return v;
}
// csbci_goal_constraints (dynamics.jl:37-42): X[i,N] - ub <= 0 (j even), lb - X[i,N] <= 0 (j odd); hard
template <int M> GDEV double box_c0(const IpmCtx<M>& c, int j, const double* x) {
  const int i = j >> 1;
  return (j & 1) == 0 ? x[i] - c.goal_hi[i] : c.goal_lo[i] - x[i];
}

// ------------------------------------------------------------------------------------- small SPD blocks
// P = (H + dp I)^-1 for a packed-lower SPD block of order n <= 4 (Cholesky, triangular inverse, Li' Li).
template <int n> GDEV bool spd_inv_packed(const double* H, double dp, double* P) {
  double Lc[n * (n + 1) / 2], Li[n * (n + 1) / 2];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < n; ++j) {
    double dj = H[tri(j, j)] + dp;
#pragma unroll
    for (int m = 0; m < j; ++m) dj -= Lc[tri(j, m)] * Lc[tri(j, m)];
    if (!(dj > 0.0)) { ok = false; dj = 1e-300; }
    const double il = g_rsqrt(dj);
    Lc[tri(j, j)] = il;                       // the diagonal holds 1 / l_jj
#pragma unroll
    for (int i = j + 1; i < n; ++i) {
      double v = H[tri(i, j)];
#pragma unroll
      for (int m = 0; m < j; ++m) v -= Lc[tri(i, m)] * Lc[tri(j, m)];
      Lc[tri(i, j)] = v * il;
    }
  }
#pragma unroll
  for (int j = 0; j < n; ++j) {
    Li[tri(j, j)] = Lc[tri(j, j)];
#pragma unroll
    for (int i = j + 1; i < n; ++i) {
      double s = 0.0;
#pragma unroll
      for (int m = j; m < i; ++m) s += Lc[tri(i, m)] * Li[tri(m, j)];
      Li[tri(i, j)] = -s * Lc[tri(i, i)];
    }
  }
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = 0.0;
#pragma unroll
      for (int m = i; m < n; ++m) s += Li[tri(m, i)] * Li[tri(m, j)];
      P[tri(i, j)] = s;
    }
  return ok;
}

// out[0:n) = Sym(packed) * v[0:n)
template <int n> GDEV void sym_mv(const double* Pk, const double* v, double* out) {
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < n; ++j) s += Pk[tri(i, j)] * v[j];
    out[i] = s;
  }
}

// ------------------------------------------------------------------------------------- structured operators
// Equality system, row j = 0..N, applied to a primal vector v (layout [k][NV]):
//   j = 0      : x_0
//   1..N-1     : (I + h/2 A_{j-1}) x_{j-1} + G u_{j-1} - (I - h/2 A_j) x_j + G u_j          (G = h/2 B)
//   j = N      : M x_{N-1}      (M = diag(goal_type == POINT))
template <int M> GDEV void aeq_row(const IpmCtx<M>& c, const double* v, int j, double* out) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU, NV = NX + NU, ANZ = T::ANZ;
  const int N = c.N;
  if (j == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) out[i] = v[i];
  } else if (j == N) {
#pragma unroll
    for (int i = 0; i < NX; ++i) out[i] = ((c.pmask >> i) & 1) ? v[(N - 1) * NV + i] : 0.0;
  } else {
    const double* vp = v + (j - 1) * NV;
    const double* vc = v + j * NV;
    const double* Ap = c.Ac + (size_t)(j - 1) * ANZ;
    const double* Ak = c.Ac + (size_t)j * ANZ;
    double acc[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) acc[i] = 0.0;
#pragma unroll
    for (int e = 0; e < ANZ; ++e) acc[T::a_row(e)] += Ap[e] * vp[T::a_col(e)] + Ak[e] * vc[T::a_col(e)];
#pragma unroll
    for (int a = 0; a < NU; ++a) acc[T::b_row(a)] += c.bv[a] * (vp[NX + a] + vc[NX + a]);
#pragma unroll
    for (int i = 0; i < NX; ++i) out[i] = vp[i] - vc[i] + c.hh * acc[i];
  }
}
// (Aeq' nu) at knot k: out[0:NX) state part, out[NX:NV) control part
template <int M> GDEV void aeqT_knot(const IpmCtx<M>& c, const double* nu, int k, double* out) {
  using T = Traits<M>;
  constexpr int NX = T::NX, NU = T::NU, ANZ = T::ANZ;
  const int N = c.N;
  const double* nk = nu + k * NX;          // row k       (knot k as "current")
  const double* nn = nu + (k + 1) * NX;    // row k + 1   (knot k as "previous")
  const double* Ak = c.Ac + (size_t)k * ANZ;
  double wv[NX], acc[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    const double a = nk[i], b2 = nn[i];
    wv[i] = (k > 0 ? a : 0.0) + (k < N - 1 ? b2 : 0.0);
    out[i] = (k == 0 ? a : -a) + (k == N - 1 ? (((c.pmask >> i) & 1) ? b2 : 0.0) : b2);
    acc[i] = 0.0;
  }
#pragma unroll
  for (int e = 0; e < ANZ; ++e) acc[T::a_col(e)] += Ak[e] * wv[T::a_row(e)];
#pragma unroll
  for (int i = 0; i < NX; ++i) out[i] += c.hh * acc[i];
#pragma unroll
  for (int a = 0; a < NU; ++a) out[NX + a] = c.hh * c.bv[a] * wv[T::b_row(a)];
}

// out = (H_k + dp I)^-1 in  for knot k:  state part blockdiag(P_b) - coef w w', control part blockdiag(Th_b)
template <int M, int B0 = 0> GDEV void phi_x_blocks(const double* P, const double* in, double* out) {
  using T = Traits<M>;
  if constexpr (B0 < T::XB_CNT) {
    sym_mv<T::XB_n(B0)>(P + T::XB_pk(B0), in + T::XB_off(B0), out + T::XB_off(B0));
    phi_x_blocks<M, B0 + 1>(P, in, out);
  }
}
template <int M, int B0 = 0> GDEV void phi_u_blocks(const double* P, const double* in, double* out) {
  using T = Traits<M>;
  if constexpr (B0 < T::UB_CNT) {
    sym_mv<T::UB_n(B0)>(P + T::UB_pk(B0), in + T::UB_off(B0), out + T::UB_off(B0));
    phi_u_blocks<M, B0 + 1>(P, in, out);
  }
}
template <int M> GDEV void apply_phi(const IpmCtx<M>& c, int k, const double* in, double* out) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX;
  const double* kd = c.kd + (size_t)k * L::KDW;
  phi_x_blocks<M>(kd + L::KD_P, in, out);
  if (T::HAS_TR) {
    const double* w = kd + L::KD_W;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) s += w[i] * in[i];
    s *= kd[L::KD_COEF];
#pragma unroll
    for (int i = 0; i < NX; ++i) out[i] -= s * w[i];
  }
  phi_u_blocks<M>(kd + L::KD_TH, in + NX, out + NX);
}
// out = H_k in  (unregularised)
template <int M> GDEV void apply_H(const IpmCtx<M>& c, int k, const double* in, double* out) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NV = L::NV;
  const double* kd = c.kd + (size_t)k * L::KDW;
  phi_x_blocks<M>(kd + L::KD_HB, in, out);
  if (T::HAS_TR) {
    double gvec[NX], s = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) { gvec[i] = 2.0 * (sh_z<M>(c)[k * NV + i] - c.Xp[k * NX + i]); s += gvec[i] * in[i]; }
    s *= kd[L::KD_KAP];
#pragma unroll
    for (int i = 0; i < NX; ++i) out[i] += s * gvec[i];
  }
  phi_u_blocks<M>(kd + L::KD_HU, in + NX, out + NX);
}

// ------------------------------------------------------------------------- assembly of H, rhs, residuals
struct Resid { double rz, rp, rc, mu, npair; };

// accumulate one row into a block:  gz += lam gv,  rr -= gv bt,  Hb += lam hess + kap gv gv'   (block-local indices)
template <int n> GDEV void block_add(double* Hb, double* gzb, double* rrb, int i0, int m, const double* gv, const double* hq,
                                     const Pair& q, int phase) {
#pragma unroll
  for (int a = 0; a < 4; ++a) if (a < m && i0 + a < n) {
    gzb[i0 + a] += q.la * gv[a];
    rrb[i0 + a] -= gv[a] * q.bt;
    if (phase == 0) {
      Hb[tri(i0 + a, i0 + a)] += q.la * hq[a];
#pragma unroll
      for (int b2 = 0; b2 <= a; ++b2) Hb[tri(i0 + a, i0 + b2)] += q.kap * gv[a] * gv[b2];
    }
  }
}

template <int M> struct KnotAcc {       // what the state blocks of a knot share
  double la_tr, bt_tr, kap_tr, gw;
  Stat st;
};

template <int M, int B0>
GDEV void assemble_xblock(const IpmCtx<M>& c, int k, int phase, double smu, const double* x, double* gz, KnotAcc<M>& ka, bool& ok) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NV = L::NV;
  if constexpr (B0 < T::XB_CNT) {
    constexpr int n = T::XB_n(B0), off = T::XB_off(B0), npk = n * (n + 1) / 2;
    double Hb[npk], rr[n], gtr[n];
#pragma unroll
    for (int i = 0; i < npk; ++i) Hb[i] = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) { rr[i] = 0.0; gtr[i] = 0.0; }
    if (T::HAS_TR) {
#pragma unroll
      for (int i = 0; i < n; ++i) {
        gtr[i] = 2.0 * (x[off + i] - c.Xp[k * NX + off + i]);
        gz[off + i] += ka.la_tr * gtr[i];
        rr[i] -= gtr[i] * ka.bt_tr;
        Hb[tri(i, i)] = 2.0 * ka.la_tr;
      }
    }
    // special rows living in this block
#pragma unroll
    for (int s = L::S_NORM; s < L::S_BALL; ++s) {
      if (spec_block<M>(s) != B0) continue;
      SpecEval o;
      spec_eval<M>(c, k, s, x, x + NX, o);
      double* st = c.sslot + ((size_t)k * L::SP + s) * SLOT_W;
      Pair q;
      if (phase == 0) pair_floor(st, o.has_t, c.floor_);
      pair_eval(st, o.has_t, o.c0, c.omega, smu, phase, q);
      if (phase == 0) pair_stat(st, o.has_t, q, ka.st);
      block_add<n>(Hb, gz + off, rr, o.i0 - off, o.n, o.gv, o.hq, q, phase);
    }
    // convexified obstacle rows (compacted): off - nhat.r - t <= 0
    if (T::WS > 0 && B0 == T::XB_of(0)) {
      constexpr int WS = T::WS > 0 ? T::WS : 1;
      const int s0 = sh_seg<M>(c)[k], s1 = sh_seg<M>(c)[k + 1];
      const double* __restrict__ orow = c.orow;
      double* __restrict__ ost = c.ost;
      for (int p = s0; p < s1; ++p) {
        if (p + 2 < s1) { g_prefetch_l1(orow + (size_t)(p + 2) * OROW_W); g_prefetch_l1(ost + (size_t)(p + 2) * SLOT_W); }
        const double* row = orow + (size_t)p * OROW_W;
        double* st = ost + (size_t)p * SLOT_W;
        double gv[4] = {0, 0, 0, 0}, hq[4] = {0, 0, 0, 0};
        double v = row[3];
#pragma unroll
        for (int a = 0; a < WS; ++a) { gv[a] = -row[a]; v -= row[a] * x[a]; }
        Pair q;
        if (phase == 0) pair_floor(st, true, c.floor_);
        pair_eval(st, true, v, c.omega, smu, phase, q);
        if (phase == 0) pair_stat(st, true, q, ka.st);
        block_add<n>(Hb, gz + off, rr, 0, WS, gv, hq, q, phase);
      }
    }
    // goal box rows (last knot only)
    if (k == c.N - 1 && c.bmask != 0) {       // rare path: sides kept rolled (the kernel is instruction-cache bound)
#pragma unroll
      for (int i = 0; i < n; ++i) {
        if (!((c.bmask >> (off + i)) & 1)) continue;
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
          const int j = 2 * (off + i) + side;
          double* st = c.bslot + (size_t)j * SLOT_W;
          double gv[4] = {side == 0 ? 1.0 : -1.0, 0, 0, 0}, hq[4] = {0, 0, 0, 0};
          Pair q;
          if (phase == 0) pair_floor(st, false, c.floor_);
          pair_eval(st, false, box_c0<M>(c, j, x), c.omega, smu, phase, q);
          if (phase == 0) pair_stat(st, false, q, ka.st);
          block_add<n>(Hb, gz + off, rr, i, 1, gv, hq, q, phase);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < n; ++i) {
      c.r[k * NV + off + i] = rr[i] - gz[off + i];
      if (phase == 0) { const double a = fabs(gz[off + i]); ka.st.rz = a > ka.st.rz ? a : ka.st.rz; if (!(a == a)) ka.st.rz = 1e300; }
    }
    if (phase == 0) {
      double* kd = c.kd + (size_t)k * L::KDW;
      double P[npk];
      if (!spd_inv_packed<n>(Hb, c.dp, P)) ok = false;
#pragma unroll
      for (int i = 0; i < npk; ++i) { kd[L::KD_HB + T::XB_pk(B0) + i] = Hb[i]; kd[L::KD_P + T::XB_pk(B0) + i] = P[i]; }
      if (T::HAS_TR) {
        double w[n];
        sym_mv<n>(P, gtr, w);
#pragma unroll
        for (int i = 0; i < n; ++i) { kd[L::KD_W + off + i] = w[i]; ka.gw += w[i] * gtr[i]; }
      }
    }
    assemble_xblock<M, B0 + 1>(c, k, phase, smu, x, gz, ka, ok);
  }
}

template <int M, int B0>
GDEV void assemble_ublock(const IpmCtx<M>& c, int k, int phase, double smu, const double* x, double wk, double* gz, KnotAcc<M>& ka, bool& ok) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NV = L::NV;
  if constexpr (B0 < T::UB_CNT) {
    constexpr int n = T::UB_n(B0), off = T::UB_off(B0), npk = n * (n + 1) / 2;
    double Hb[npk], rr[n];
#pragma unroll
    for (int i = 0; i < npk; ++i) Hb[i] = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) { rr[i] = 0.0; Hb[tri(i, i)] = 2.0 * wk; }
#pragma unroll
    for (int s = L::S_BALL; s < L::SP; ++s) {
      if (spec_block<M>(s) != B0) continue;
      SpecEval o;
      spec_eval<M>(c, k, s, x, x + NX, o);
      if (!o.valid) continue;
      double* st = c.sslot + ((size_t)k * L::SP + s) * SLOT_W;
      Pair q;
      if (phase == 0) pair_floor(st, false, c.floor_);
      pair_eval(st, false, o.c0, c.omega, smu, phase, q);
      if (phase == 0) pair_stat(st, false, q, ka.st);
      block_add<n>(Hb, gz + NX + off, rr, o.i0 - off, o.n, o.gv, o.hq, q, phase);
    }
#pragma unroll
    for (int i = 0; i < n; ++i) {
      c.r[k * NV + NX + off + i] = rr[i] - gz[NX + off + i];
      if (phase == 0) { const double a = fabs(gz[NX + off + i]); ka.st.rz = a > ka.st.rz ? a : ka.st.rz; if (!(a == a)) ka.st.rz = 1e300; }
    }
    if (phase == 0) {
      double* kd = c.kd + (size_t)k * L::KDW;
      double P[npk];
      if (!spd_inv_packed<n>(Hb, c.dp, P)) ok = false;
#pragma unroll
      for (int i = 0; i < npk; ++i) { kd[L::KD_HU + T::UB_pk(B0) + i] = Hb[i]; kd[L::KD_TH + T::UB_pk(B0) + i] = P[i]; }
    }
    assemble_ublock<M, B0 + 1>(c, k, phase, smu, x, wk, gz, ka, ok);
  }
}

// phase 0: build the per-knot Hessian blocks, their inverses and the predictor rhs (sigma*mu = 0, no second-order
//          term); also the residual norms.  Returns false through *ok_out when a block is not positive definite.
// phase 1: corrector rhs with centering target `smu` and the stored predictor products.
template <int M> GDEV_NOINLINE void assemble(const IpmCtx<M>& c, int phase, double smu, Resid* out, bool* ok_out) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV, ANZ = L::ANZ;
  const int N = c.N;
  Stat S; S.rz = 0; S.rc = 0; S.mus = 0; S.np = 0;
  bool ok = true;
  G_PAR_FOR(k, N) {
    const double* x = sh_z<M>(c) + k * NV;
    const double* u = x + NX;
    double gz[NV];
    aeqT_knot<M>(c, c.nu, k, gz);
    const double wk = (k == 0 || k == N - 1) ? 0.5 * c.h : c.h;
#pragma unroll
    for (int i = 0; i < NU; ++i) gz[NX + i] += 2.0 * wk * u[i];
    KnotAcc<M> ka;
    ka.la_tr = 0; ka.bt_tr = 0; ka.kap_tr = 0; ka.gw = 0; ka.st = S;
    if (T::HAS_TR) {
      double* st = c.sslot + ((size_t)k * L::SP + L::S_TR) * SLOT_W;
      Pair q;
      if (phase == 0) pair_floor(st, true, c.floor_);
      pair_eval(st, true, tr_c0<M>(c, k, x), c.omega, smu, phase, q);
      if (phase == 0) pair_stat(st, true, q, ka.st);
      ka.la_tr = q.la; ka.bt_tr = q.bt; ka.kap_tr = q.kap;
    }
    assemble_xblock<M, 0>(c, k, phase, smu, x, gz, ka, ok);
    assemble_ublock<M, 0>(c, k, phase, smu, x, wk, gz, ka, ok);
    S = ka.st;
    if (phase == 0) {
      double* kd = c.kd + (size_t)k * L::KDW;
      kd[L::KD_KAP] = ka.kap_tr;
      kd[L::KD_COEF] = T::HAS_TR ? ka.kap_tr * g_rcp(1.0 + ka.kap_tr * ka.gw) : 0.0;
    }
    {   // first pass of the KKT solve that follows: t1 = (H + dp I)^-1 r  (kkt_solve)
      double rv[NV], tv[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) rv[i] = c.r[k * NV + i];
      apply_phi<M>(c, k, rv, tv);
#pragma unroll
      for (int i = 0; i < NV; ++i) c.t1[k * NV + i] = tv[i];
    }
  }
  if (phase == 0) {
    // equality residual r_p = Aeq z - b  ;  rnu = -r_p
    double rpmax = 0;
    G_PAR_FOR(j, N + 1) {
      double v[NX];
      aeq_row<M>(c, sh_z<M>(c), j, v);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double t = v[i];
        if (j == 0) t -= c.x_init[i];
        else if (j == N) t -= ((c.pmask >> i) & 1) ? c.goal_lo[i] : 0.0;
        else t += c.hh * (c.g[(j - 1) * NX + i] + c.g[j * NX + i]);
        c.rnu[j * NX + i] = -t;
        const double a = fabs(t);
        rpmax = a > rpmax ? a : rpmax;
        if (!(a == a)) rpmax = 1e300;
      }
    }
    out->rz = block_max(S.rz, c.red);
    out->rp = block_max(rpmax, c.red);
    out->rc = block_max(S.rc, c.red);
    const double ms = block_sum(S.mus, c.red), np = block_sum(S.np, c.red);
    out->npair = np;
    out->mu = np > 0 ? ms / np : 0.0;
    if (!(out->mu == out->mu)) out->mu = 1e300;
    *ok_out = block_max(ok ? 0.0 : 1.0, c.red) == 0.0;
  } else {
    G_SYNC();
  }
}

// ------------------------------------------------------------------------------------------ Schur complement
// S = Aeq (H + dp I)^-1 Aeq' is never stored: its block rows are produced one per sweep step, in shared memory, by the
// threads that do not run the elimination.  Knot k appears in row k as "current" (coefficient Lk = aL A_k + bL I;
// row 0 is x_0 itself) and in row k+1 as "previous" (Rk = aR A_k + diag(dR); row N is the masked goal row):
//   S_kk += Lk Phi Lk' + Xi,   S_{k+1,k} = Rk Phi Lk' + Xi,   S_{k+1,k+1} += Rk Phi Rk' + Xi,   Xi = G Theta_k G',
// with Phi = blockdiag(P_b) - coef w w'.  With Y = (h/2 A) Phi and Z = Y (h/2 A)' (two small tile products) all three
// are element-wise combinations of Z, Y, Y' and Phi.
template <int M> GHD constexpr int ctrl_of_row(int I) { int r = -1; for (int a = 0; a < Traits<M>::NU; ++a) if (Traits<M>::b_row(a) == I) r = a; return r; }

struct SchurTiles { double *Phi, *Ah, *Y, *Z, *RR, *Xi; };

// Threads [t0, t0 + nt) build the dense Phi of the staged knot record `ks`, and refresh Ah (the pattern of A is static,
// so the tile is zeroed once by the caller).  Barrier-free: followed by the caller's barrier.
template <int M> GDEV void schur_build(const IpmCtx<M>& c, const double* ks, const SchurTiles& t, int t0, int nt) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, LDT = L::LDT;
  const unsigned char* mt = c.tab + L::TAB_MODEL;           // a_row | a_col | blk_of | ctrl_of
  G_ASSUME_SHARED(mt);
  const double coef = ks[L::KS_COEF];
  for (int it = G_TID - t0; it < NX * NX; it += nt) {
    const int i = it / NX, q = it - i * NX;
    const int bi = mt[2 * L::ANZ + i];
    double v = 0.0;
    if (bi == mt[2 * L::ANZ + q]) { const int off = T::XB_off(bi); v = ks[L::KS_P + T::XB_pk(bi) + tri(i - off, q - off)]; }
    if (T::HAS_TR) v -= coef * ks[L::KS_W + i] * ks[L::KS_W + q];
    t.Phi[i * LDT + q] = v;
  }
  for (int e = G_TID - t0; e < L::ANZ; e += nt) t.Ah[mt[e] * LDT + mt[L::ANZ + e]] = c.hh * ks[L::KS_A + e];
  // control coupling Xi = G Theta G' (G = h/2 B): one entry per pair of controls of the same block (static pattern too)
  for (int pr = G_TID - t0; pr < L::NU * L::NU; pr += nt) {
    const int a = pr / L::NU, b2 = pr - a * L::NU;
    const int ub = T::UB_of(a);
    if (ub != T::UB_of(b2)) continue;
    const int uo = T::UB_off(ub);
    t.Xi[T::b_row(a) * LDT + T::b_row(b2)] = c.hh * c.hh * c.bv[a] * c.bv[b2] * ks[L::KS_TH + T::UB_pk(ub) + tri(a - uo, b2 - uo)];
  }
}

// out(i, q) = sum_m X[i][m] * Yt[q][m]  for all (i, q), by threads [t0, t0 + nt)
template <int M> GDEV void abt_task(const double* X, const double* Y, int i, int q0, double* acc);
template <int M> GDEV void schur_product(const double* X, const double* Yt, double* out, int t0, int nt) {
  using L = IpmLayout<M>;
  for (int task = G_TID - t0; task < L::NTASK; task += nt) {
    const int i = task / L::NG, q0 = (task - i * L::NG) * L::CG;
    double acc[L::CG];
    abt_task<M>(X, Yt, i, q0, acc);
#pragma unroll
    for (int c2 = 0; c2 < L::CG; ++c2) if (q0 + c2 < L::NX) out[i * L::LDT + q0 + c2] = acc[c2];
  }
}

// Element-wise emission for knot k (k < N):  Sdd (= S_kk) = RR + Lk Phi Lk' [+ Xi] (+ regularisation),
// Sod (= S_{k+1,k}) = Rk Phi Lk' [+ Xi],  RR = Rk Phi Rk' [+ Xi]  (carried to S_{k+1,k+1}).
template <int M> GDEV void schur_emit(const IpmCtx<M>& c, const SchurTiles& t, int k, double* Sdd, double* Sod,
                                      int t0, int nt) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, LDT = L::LDT;
  const int N = c.N;
  // Y and Z already carry the h/2 factors, so the A-coefficients of Lk and Rk are 0 or 1 here
  const double aL = k == 0 ? 0.0 : 1.0, bL = k == 0 ? 1.0 : -1.0;
  const bool last = (k == N - 1);
  const double aR = last ? 0.0 : 1.0;
  const int pmask = c.pmask;
  for (int it = G_TID - t0; it < NX * NX; it += nt) {
    const int i = it / NX, q = it - i * NX;
    const double z = t.Z[i * LDT + q], y = t.Y[i * LDT + q], yt = t.Y[q * LDT + i], ph = t.Phi[i * LDT + q];
    const double dRi = last ? (double)((pmask >> i) & 1) : 1.0, dRq = last ? (double)((pmask >> q) & 1) : 1.0;
    const double xi = t.Xi[i * LDT + q];
    double dd = t.RR[i * LDT + q] + aL * aL * z + aL * bL * (y + yt) + bL * bL * ph + (k >= 1 ? xi : 0.0);
    if (i == q) dd += c.dd * dd + 1e-300;
    Sdd[i * LDT + q] = dd;
    Sod[i * LDT + q] = aR * aL * z + aR * bL * y + aL * dRi * yt + bL * dRi * ph + ((k >= 1 && k <= N - 2) ? xi : 0.0);
    t.RR[i * LDT + q] = aR * aR * z + aR * (y * dRq + dRi * yt) + dRi * ph * dRq + (k <= N - 2 ? xi : 0.0);
  }
}
// Last block row: S_NN = RR (+ identity on the free goal coordinates so that the block stays non-singular)
template <int M> GDEV void schur_emit_last(const IpmCtx<M>& c, const SchurTiles& t, double* Sdd, int t0, int nt) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, LDT = L::LDT;
  for (int it = G_TID - t0; it < NX * NX; it += nt) {
    const int i = it / NX, q = it - i * NX;
    double dd = t.RR[i * LDT + q];
    if (i == q) { if (!((c.pmask >> i) & 1)) dd += 1.0; dd += c.dd * dd + 1e-300; }
    Sdd[i * LDT + q] = dd;
  }
}
// Stage the record of knot k (P, w, Theta, coef, A entries) from global memory into the idle half of the double-buffered
// staging area with 8-byte asynchronous copies (LDGSTS): nothing passes through registers (a register-staged version
// was spilled to local memory by ptxas and serialised three DRAM latencies on the producer warp), and the producer only
// waits for the copies (schur_stage_wait) after its arithmetic, right before the step's barrier.
template <int M> GDEV const double* schur_stage_ptr(const double* kd, const double* Ak, int t) {
  using L = IpmLayout<M>;
  if (t < L::KS_TH) return kd + L::KD_P + t;                      // P | w are contiguous in the record
  if (t < L::KS_COEF) return kd + L::KD_TH + t - L::KS_TH;
  if (t == L::KS_COEF) return kd + L::KD_COEF;
  return Ak + (t - L::KS_A);
}
template <int M> GDEV double schur_stage_src(const double* kd, const double* Ak, int t) {
  using L = IpmLayout<M>;
  return t < L::KS_A + L::ANZ ? *schur_stage_ptr<M>(kd, Ak, t) : 0.0;
}
template <int M> GDEV void schur_stage_async(const IpmCtx<M>& c, int k, double* ks, int t0, int nt) {
  using L = IpmLayout<M>;
  const double* kd = c.kd + (size_t)k * L::KDW;
  const double* Ak = c.Ac + (size_t)k * L::ANZ;
  for (int t = G_TID - t0; t < L::KS_A + L::ANZ; t += nt) g_cp_async8(ks + t, schur_stage_ptr<M>(kd, Ak, t));   // the padding stays 0
}
GDEV void schur_stage_wait() { g_cp_async_wait(); }

// acc[c] = sum_m X[i][m] * Y[q0 + c][m]  over one shared-memory tile row pair (rows are LDT apart, MLEN = GLD terms)
template <int M> GDEV void abt_task(const double* X, const double* Y, int i, int q0, double* acc) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, LDT = L::LDT, CG = L::CG;
#pragma unroll
  for (int c2 = 0; c2 < CG; ++c2) acc[c2] = 0.0;
#pragma unroll
  for (int m = 0; m < L::GLD; m += 2) {
    const g_d2 xv = g_ld2(X + i * LDT + m);
#pragma unroll
    for (int c2 = 0; c2 < CG; ++c2) {
      const int q = q0 + c2 < NX ? q0 + c2 : NX - 1;      // clamped: the caller drops columns >= NX
      const g_d2 yv = g_ld2(Y + q * LDT + m);
      acc[c2] += xv.x * yv.x;
      acc[c2] += xv.y * yv.y;
    }
  }
}

// Factorisation of the block-tridiagonal S.  Block Cholesky recurrence (numerically the stable form: every update is a
// symmetric  D_j = S_jj - Lo_j Lo_j'  with Lo_j = S_{j,j-1} L_{j-1}^-T), but what is STORED is the block L D L' form the
// solves want:  slot (j,0) S_jj -> D_j^-1 = Li_j' Li_j,  slot (j,1) S_{j,j-1} -> V_j = Lo_j Li_{j-1} (= S_{j,j-1} D_{j-1}^-1),
// so that a solve is two chains of single mat-vecs plus one parallel D^-1 pass.  Li_j = L_j^-1 comes out of one
// Gaussian elimination of [D_j | I] in shared memory (NX dependent pivot steps on warp 0 -- the critical path of the
// whole kernel) while the other warp stages block row j+1.  All tile products are row-by-row dot products over
// zero-padded tiles (LDS.128, a 1 x CG register tile per thread).  Returns false on a non-positive pivot.
template <int M> GDEV_NOINLINE bool factorize(const IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NT = NX * (NX + 1) / 2, LDT = L::LDT, TILE = L::TILE, GLD = L::GLD, GT = L::GT;
  constexpr int CG = L::CG, NG = L::NG, NTASK = L::NTASK, NLT = L::NLT;
  const int N = c.N;
  long long tc0 = g_clock();
  // dz | sy | ring are dead here: 15 tiles + pivots + two staged knot records
  double* tile = sh_dz<M>(c);
  double* W = tile;                          // D_j, eliminated in place (lower triangle)
  double* Wr = tile + TILE;                  // I -> unit-lower elimination history (L~^-1)
  double* Li = tile + 2 * TILE;              // L_j^-1 (lower, explicit zeros) and its transpose
  double* LiT = tile + 3 * TILE;
  // ping-pong tiles are addressed arithmetically (an indexed pointer array would live in local memory)
#define GUSTO_SDD(w) (tile + 4 * TILE)                 /* one tile: emitted and consumed within a step */
#define GUSTO_SOD(w) (tile + (6 + (w)) * TILE)
#define GUSTO_LO(w) (tile + (8 + (w)) * TILE)
  SchurTiles st;
  st.Phi = tile + 10 * TILE; st.Ah = tile + 11 * TILE; st.Y = tile + 12 * TILE; st.Z = tile + 13 * TILE; st.RR = tile + 14 * TILE;
  st.Xi = tile + 5 * TILE;
  double* ipv = tile + L::FAC_TILES * TILE;  // 1 / pivot, then 1 / sqrt(pivot)
  double* ksb = ipv + ((NX + 1) & ~1);       // two staged knot records
  const unsigned char* const tab = c.tab;
  G_ASSUME_SHARED(tab);
  const unsigned char* const tab2 = tab + 2 * NT;
  double* const fac = c.fac;
  // producer threads: the second warp (or the only one)
  const int pf0 = G_NTHR > G_WARP ? G_WARP : 0;
  const int npf = G_NTHR > G_WARP ? G_WARP : G_NTHR;
  const bool producer = G_TID >= pf0 && G_TID < pf0 + npf;
  double bad = 0.0;
  G_PAR_FOR(it, L::FAC_TILES * TILE) tile[it] = 0.0;
  // knot 0 (all threads): S_00 -> W, S_10 -> Sod[1], RR <- R0 Phi R0'; stage knot 1
  for (int t = G_TID; t < 2 * L::KS; t += G_NTHR) {
    const int k = t / L::KS, tt = t - k * L::KS;
    if (k < N) ksb[t] = schur_stage_src<M>(c.kd + (size_t)k * L::KDW, c.Ac + (size_t)k * L::ANZ, tt);
  }
  G_SYNC();
  schur_build<M>(c, ksb, st, 0, G_NTHR);
  G_SYNC();
  schur_product<M>(st.Ah, st.Phi, st.Y, 0, G_NTHR);
  G_SYNC();
  schur_product<M>(st.Y, st.Ah, st.Z, 0, G_NTHR);
  G_SYNC();
  schur_emit<M>(c, st, 0, W, GUSTO_SOD(1), 0, G_NTHR);
  G_SYNC();
#if !(defined(GUSTO_PROF_MODE) && GUSTO_PROF_MODE == 3)
  if (G_TID == 0) c.prof[0] += g_clock() - tc0;
#endif
  tc0 = g_clock();
  // my entries of the packed lower triangle during the elimination (warp 0)
  int ei[3], ee[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int t = G_TID + r * G_WARP;
    ei[r] = t < NT ? tab[t] : 0; ee[r] = t < NT ? tab[NT + t] : 0;
  }
  for (int j = 0; j <= N; ++j) {
    const int cur = j & 1, nxt = cur ^ 1;
    // (E) warp 0: eliminate column q from the rows below it, in W (columns > q) and in Wr (columns <= q)
#if defined(GUSTO_PROF_MODE) && GUSTO_PROF_MODE == 3
    long long tp = g_clock();
#define GUSTO_PROF_TICK(slot) do { if (G_TID == 0) { const long long tn = g_clock(); c.prof[slot] += tn - tp; tp = tn; } } while (0)
#else
#define GUSTO_PROF_TICK(slot) ((void)0)
#endif
    if (G_TID < G_WARP) {
#ifdef GUSTO_HOSTSIM
      for (int q = 0; q < NX; ++q) {
        double piv = W[q * LDT + q];
        if (!(piv > 0.0)) { bad = 1.0; piv = 1e-300; }
        const double ip = g_rcp(piv);
        ipv[q] = ip;
        for (int t = 0; t < NT; ++t) {
          const int i = tab[t], e = tab[NT + t];
          if (i <= q) continue;
          const double mult = W[i * LDT + q] * ip;
          if (e < q) Wr[i * LDT + e] -= mult * Wr[q * LDT + e];
          else if (e == q) Wr[i * LDT + e] = -mult;            // Wr starts as the identity, kept implicit
          else W[i * LDT + e] -= mult * W[e * LDT + q];
        }
      }
#else
      // Branch-free: every lane owns <= 3 entries (i, e) of the packed lower triangle; entry (i, e) lives in W while
      // e > q and in Wr afterwards.  All operands of a step are loaded before any arithmetic, and the reciprocal of
      // the NEXT pivot is started from pre-update values so that it overlaps the rank-1 update (it is the longest
      // dependent operation of the step).
      double piv = W[0];
      if (!(piv > 0.0)) { bad = 1.0; piv = 1e-300; }
      double ip = g_rcp(piv);
      for (int q = 0; q < NX; ++q) {
        const int qn = q + 1 < NX ? q + 1 : q;
        const double wq1 = W[qn * LDT + q], d1 = W[qn * LDT + qn];
        double av[3], sv[3], ov[3];
        int to[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int i = ei[r], e = ee[r];
          const bool right = e <= q;
          to[r] = (right ? TILE : 0) + i * LDT + e;
          av[r] = W[i * LDT + q];
          sv[r] = tile[right ? TILE + q * LDT + e : e * LDT + q];
          ov[r] = tile[to[r]];
          if (e == q) { sv[r] = 1.0; ov[r] = 0.0; }            // Wr starts as the identity, kept implicit
        }
        G_SYNCWARP();                                // the look-ahead read W[q+1][q+1] before its owner updates it
        double pn = fma(-(wq1 * ip), wq1, d1);
        if (q + 1 < NX && !(pn > 0.0)) { bad = 1.0; pn = 1e-300; }
        const double ipn = g_rcp(pn);
        if (G_LANE == 0) ipv[q] = ip;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const double nv = fma(-(av[r] * ip), sv[r], ov[r]);
          if (ei[r] > q) tile[to[r]] = nv;
        }
        ip = ipn;
        G_SYNCWARP();
      }
#endif
      G_W0_FOR(q, NX) { const double v = ipv[q]; ipv[q] = v * g_rsqrt(v); }     // sqrt(1/pivot) without the IEEE sqrt sequence
      G_SYNCWARP();
      // (C1) Li = diag(rs) * Wr  (lower triangular) and its transpose -- still on the eliminating warp, which would
      // otherwise wait for the producer
      G_W0_FOR(it, NX * NX) {
        const int i = it / NX, m = it - i * NX;
        const double v = m < i ? Wr[i * LDT + m] * ipv[i] : (m == i ? ipv[i] : 0.0);
        Li[i * LDT + m] = v;
        LiT[m * LDT + i] = v;
      }
    }
    GUSTO_PROF_TICK(0);
    // producer: block row j+1 of S from knot j+1 (and the record of knot j+2 staged for the next step)
    if (producer) {
      const int kn = j + 1;
      if (kn <= N - 1) {
        const double* ks = ksb + (kn & 1) * L::KS;
        if (kn + 1 <= N - 1) schur_stage_async<M>(c, kn + 1, ksb + ((kn + 1) & 1) * L::KS, pf0, npf);
        schur_build<M>(c, ks, st, pf0, npf);
        G_SYNCWARP();
        schur_product<M>(st.Ah, st.Phi, st.Y, pf0, npf);
        G_SYNCWARP();
        schur_product<M>(st.Y, st.Ah, st.Z, pf0, npf);
        schur_stage_wait();
      }
    }
    G_SYNC();
    GUSTO_PROF_TICK(2);
    // (C2) Lo_{j+1} = S_{j+1,j} Li'
    if (j < N) {
      G_PAR_FOR(task, NTASK) {
        const int i = task / NG, q0 = (task - i * NG) * CG;
        double acc[CG];
        abt_task<M>(GUSTO_SOD(nxt), Li, i, q0, acc);
#pragma unroll
        for (int c2 = 0; c2 < CG; ++c2) if (q0 + c2 < NX) GUSTO_LO(nxt)[i * LDT + q0 + c2] = acc[c2];
      }
      // ... and the producer's element-wise emission of block row j+1 (S_{j+1,j+1} -> Sdd, S_{j+2,j+1} -> Sod, carry RR)
      if (j + 1 <= N - 1) schur_emit<M>(c, st, j + 1, GUSTO_SDD(nxt), GUSTO_SOD(cur), 0, G_NTHR);
      else schur_emit_last<M>(c, st, GUSTO_SDD(nxt), 0, G_NTHR);
    }
    G_SYNC();
    GUSTO_PROF_TICK(3);
    // (X) V_{j+1} = Lo_{j+1} Li -> global;  D_j^-1 = Li' Li -> global;  D_{j+1} = S_{j+1,j+1} - Lo_{j+1} Lo_{j+1}' -> W
    {
      double* gD = fac + (size_t)(2 * j) * GT;
      double* gV = fac + (size_t)(2 * (j + 1) + 1) * GT;
      const int nt = j < N ? NTASK + 2 * NLT : NLT;
      G_PAR_FOR(task, nt) {
        double acc[CG];
        if (task < NLT) {                                   // D_j^-1, lower tasks mirrored
          const int i = tab2[task], q0 = tab2[NLT + task] * CG;
          abt_task<M>(LiT, LiT, i, q0, acc);
#pragma unroll
          for (int c2 = 0; c2 < CG; ++c2) if (q0 + c2 <= i) { gD[i * GLD + q0 + c2] = acc[c2]; gD[(q0 + c2) * GLD + i] = acc[c2]; }
        } else if (task < 2 * NLT) {                        // D_{j+1}, lower tasks (the strict upper triangle of W is never read)
          const int i = tab2[task - NLT], q0 = tab2[task] * CG;
          abt_task<M>(GUSTO_LO(nxt), GUSTO_LO(nxt), i, q0, acc);
#pragma unroll
          for (int c2 = 0; c2 < CG; ++c2) if (q0 + c2 < NX) W[i * LDT + q0 + c2] = GUSTO_SDD(nxt)[i * LDT + q0 + c2] - acc[c2];
        } else {
          const int t2 = task - 2 * NLT;
          const int i = t2 / NG, q0 = (t2 - i * NG) * CG;
          abt_task<M>(GUSTO_LO(nxt), LiT, i, q0, acc);
#pragma unroll
          for (int c2 = 0; c2 < CG; ++c2) if (q0 + c2 < NX) gV[i * GLD + q0 + c2] = acc[c2];
        }
      }
    }
    G_SYNC();
    GUSTO_PROF_TICK(4);
  }
#if !(defined(GUSTO_PROF_MODE) && GUSTO_PROF_MODE == 3)
  if (G_TID == 0) c.prof[1] += g_clock() - tc0;
#endif
#undef GUSTO_SDD
#undef GUSTO_SOD
#undef GUSTO_LO
  bad = block_max(bad, c.red);
  return bad == 0.0;
}

// ------------------------------------------------------------------------------------------------ KKT solves
// TMA ring over the V tiles of the factor: one elected lane issues ONE bulk copy (cp.async.bulk, 1-D, GT*8 bytes) per
// tile into stage o % S and the copy completes on that stage's mbarrier; the consumers spin on the barrier's phase
// parity (o / S) & 1.  `o` counts tiles over the whole kernel (fetch order == consume order), so the barriers are
// initialised once.
template <int M> GDEV void ring_fetch(const IpmCtx<M>& c, const double* fac, double* ring, int N, int j, unsigned o) {
  using L = IpmLayout<M>;
  if (j >= 1 && j <= N && G_LANE == 0) {
    unsigned long long* bar = &c.ring_bar[o % RING_STAGES];
    g_mbar_expect_tx(bar, (unsigned)(L::GT * sizeof(double)));
    g_tma_bulk_g2s(ring + (o % RING_STAGES) * L::GT, fac + (size_t)(2 * j + 1) * L::GT, (unsigned)(L::GT * sizeof(double)), bar);
  }
}

// Block-tridiagonal solve  S nu = b  (sy holds b on entry, nu on exit):
//   forward  w_j = b_j - V_j w_{j-1};   middle  v_j = D_j^-1 w_j (all threads);   backward  nu_j = v_j - V_{j+1}' nu_{j+1}.
template <int M> GDEV_NOINLINE void schur_solve(const IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, LDT = L::LDT, GLD = L::GLD, S = RING_STAGES;
  const int N = c.N;
  double* const y = c.sy;
  double* const ring = c.ring;
  const double* const fac = c.fac;
  G_ASSUME_SHARED(y);
  G_ASSUME_SHARED(ring);
  unsigned o = c.ring_o;                          // next tile to consume; of = next tile to fetch
  long long tc0 = g_clock();
  if (G_TID < G_WARP) {
    unsigned of = o;
    g_fence_proxy_async();                        // the ring region was last written by ordinary stores (sweep tiles)
    for (int jj = 1; jj < S; ++jj) if (jj <= N) ring_fetch<M>(c, fac, ring, N, jj, of++);
    for (int j = 1; j <= N; ++j) {
      g_mbar_wait(&c.ring_bar[o % S], (o / S) & 1);               // tile j has landed
      G_SYNCWARP();                               // everyone is done with tile j-1: its stage can be refilled
      if (j + S - 1 <= N) ring_fetch<M>(c, fac, ring, N, j + S - 1, of++);
      const double* R = ring + (o % S) * L::GT;
      G_W0_FOR(i, NX) {
        double a0 = y[j * NX + i], a1 = 0.0;
#pragma unroll
        for (int m = 0; m < GLD; m += 2) {
          const g_d2 rv = g_ld2(R + i * GLD + m);
          a0 -= rv.x * y[(j - 1) * NX + m];
          if (m + 1 < NX) a1 -= rv.y * y[(j - 1) * NX + m + 1];
        }
        y[j * NX + i] = a0 + a1;
      }
      ++o;
    }
    G_SYNCWARP();
  }
  G_SYNC();
  if (G_TID == 0 && GUSTO_PROF_SOLVE) c.prof[2] += g_clock() - tc0;
  tc0 = g_clock();
  // middle: v = D^-1 w, staged through the (idle) ring so that no row is overwritten while still being read.
  // Two rows per thread and round, all loads issued before the arithmetic (the factor streams from L2 / HBM).
  double* v = ring;
  const int ne = (N + 1) * NX;
  for (int it0 = G_TID; it0 < ne; it0 += 2 * G_NTHR) {
    const int it1 = it0 + G_NTHR;
    const bool two = it1 < ne;
    const int j0 = it0 / NX, i0 = it0 - j0 * NX;
    const int j1 = two ? it1 / NX : j0, i1 = two ? it1 - j1 * NX : i0;
    const double* D0 = fac + (size_t)(2 * j0) * L::GT + i0 * GLD;
    const double* D1 = fac + (size_t)(2 * j1) * L::GT + i1 * GLD;
    g_d2 d0[GLD / 2], d1[GLD / 2];
#pragma unroll
    for (int m = 0; m < GLD / 2; ++m) { d0[m] = g_ld2(D0 + 2 * m); d1[m] = g_ld2(D1 + 2 * m); }
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
    for (int m = 0; m < GLD / 2; ++m) {
      a0 += d0[m].x * y[j0 * NX + 2 * m]; b0 += d1[m].x * y[j1 * NX + 2 * m];
      if (2 * m + 1 < NX) { a1 += d0[m].y * y[j0 * NX + 2 * m + 1]; b1 += d1[m].y * y[j1 * NX + 2 * m + 1]; }
    }
    v[it0] = a0 + a1;
    if (two) v[it1] = b0 + b1;
  }
  G_SYNC();
  G_PAR_FOR(it, ne) y[it] = v[it];
  G_SYNC();
  if (G_TID == 0 && GUSTO_PROF_SOLVE) c.prof[3] += g_clock() - tc0;
  tc0 = g_clock();
  if (G_TID < G_WARP) {
    // backward: tiles N, N-1, ..., 1 ; tile t is used at step j = t - 1
    unsigned of = o;
    g_fence_proxy_async();                        // the middle pass wrote the ring region with ordinary stores
    for (int jj = 0; jj < S - 1; ++jj) if (N - jj >= 1) ring_fetch<M>(c, fac, ring, N, N - jj, of++);
    for (int j = N - 1; j >= 0; --j) {
      g_mbar_wait(&c.ring_bar[o % S], (o / S) & 1);
      G_SYNCWARP();
      if (j + 1 - (S - 1) >= 1) ring_fetch<M>(c, fac, ring, N, j + 1 - (S - 1), of++);
      const double* R = ring + (o % S) * L::GT;
      G_W0_FOR(i, NX) {
        double a0 = y[j * NX + i], a1 = 0.0;
#pragma unroll
        for (int m = 0; m + 1 < NX; m += 2) { a0 -= R[m * GLD + i] * y[(j + 1) * NX + m]; a1 -= R[(m + 1) * GLD + i] * y[(j + 1) * NX + m + 1]; }
        if (NX & 1) a0 -= R[(NX - 1) * GLD + i] * y[(j + 1) * NX + NX - 1];
        y[j * NX + i] = a0 + a1;
      }
      ++o;
    }
    G_SYNCWARP();
    if (G_TID == 0) c.ring_o = o;
  }
  G_SYNC();
  if (G_TID == 0 && GUSTO_PROF_SOLVE) c.prof[4] += g_clock() - tc0;
}

// [dz; dnu] (+)= Ktilde^-1 [rin; rnuin]  with Ktilde = [[H + dp I, Aeq'], [Aeq, -dd]] via the Schur complement.
// One Schur-complement solve of  Ktilde [d; dnu] = [rin; rnu - Aeq dz]  with Ktilde = [[H + dp I, Aeq'], [Aeq, -dd]],
// added to (or installed in) dz / dnu.  On entry c.t1 holds  Phi rin (+ dz when accumulating)  -- written by assemble()
// for the first solve of a direction and by the previous call for a refinement solve -- so the Schur right-hand side
// is simply  Aeq t1 - rnu.  When `more` refinement follows, the same per-knot pass that recovers d also leaves the
// next residual  res = rin - Aeq' dnu - H d  (unregularised H, i.e. the exact KKT matrix) in c.res and its Phi-image
// plus the updated dz in c.t1: a refinement step costs two passes and two chains, nothing else.
template <int M> GDEV_NOINLINE void kkt_solve(const IpmCtx<M>& c, const double* rin, bool accumulate, bool more) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NV = L::NV;
  const int N = c.N;
  G_PAR_FOR(j, N + 1) {
    double v[NX];
    aeq_row<M>(c, c.t1, j, v);
#pragma unroll
    for (int i = 0; i < NX; ++i) sh_sy<M>(c)[j * NX + i] = v[i] - c.rnu[j * NX + i];
  }
  G_SYNC();
  schur_solve<M>(c);
  G_PAR_FOR(k, N) {
    double t[NV], d[NV];
    aeqT_knot<M>(c, sh_sy<M>(c), k, t);
#pragma unroll
    for (int i = 0; i < NV; ++i) t[i] = rin[k * NV + i] - t[i];
    apply_phi<M>(c, k, t, d);
    double dzn[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { dzn[i] = accumulate ? sh_dz<M>(c)[k * NV + i] + d[i] : d[i]; sh_dz<M>(c)[k * NV + i] = dzn[i]; }
    if (more) {
      double hd[NV], pr[NV];
      apply_H<M>(c, k, d, hd);
#pragma unroll
      for (int i = 0; i < NV; ++i) { t[i] -= hd[i]; c.res[k * NV + i] = t[i]; }
      apply_phi<M>(c, k, t, pr);
#pragma unroll
      for (int i = 0; i < NV; ++i) c.t1[k * NV + i] = pr[i] + dzn[i];
    }
  }
  {
    double* __restrict__ dnu = c.dnu;
    const double* sy = sh_sy<M>(c);
    const int ne = (N + 1) * NX;
    if (accumulate) {
#pragma unroll 4
      for (int it = G_TID; it < ne; it += G_NTHR) dnu[it] += sy[it];
    } else {
#pragma unroll 4
      for (int it = G_TID; it < ne; it += G_NTHR) dnu[it] = sy[it];
    }
  }
  G_SYNC();
}

// Solve, then `nref` refinement steps against the exact KKT matrix [[H, Aeq'], [Aeq, 0]].
template <int M> GDEV_NOINLINE void kkt_solve_refined(const IpmCtx<M>& c, int nref) {
  kkt_solve<M>(c, c.r, false, nref > 0);
  for (int it_ref = 1; it_ref <= nref; ++it_ref) kkt_solve<M>(c, c.res, true, it_ref < nref);
#if defined(GUSTO_HOSTSIM) && defined(GUSTO_DEBUG_KKT)
  {   // true residuals of the direction against the exact KKT matrix
    using L = IpmLayout<M>;
    constexpr int NX = L::NX, NV = L::NV;
    double rp = 0, rd = 0, np_ = 0, nd = 0;
    for (int k = 0; k < c.N; ++k) {
      double at[NV], hd[NV], dk[NV];
      aeqT_knot<M>(c, c.dnu, k, at);
      for (int i = 0; i < NV; ++i) dk[i] = c.dz[k * NV + i];
      apply_H<M>(c, k, dk, hd);
      for (int i = 0; i < NV; ++i) { rp = fmax(rp, fabs(c.r[k * NV + i] - hd[i] - at[i])); np_ = fmax(np_, fabs(c.r[k * NV + i])); }
    }
    for (int j = 0; j <= c.N; ++j) {
      double v[NX];
      aeq_row<M>(c, c.dz, j, v);
      for (int i = 0; i < NX; ++i) { rd = fmax(rd, fabs(c.rnu[j * NX + i] - v[i])); nd = fmax(nd, fabs(c.rnu[j * NX + i])); }
    }
    printf("    kkt(nref=%d): |r - H dz - A'dnu| = %.2e (|r| %.2e)   |rnu - A dz| = %.2e (|rnu| %.2e)\n", nref, rp, np_, rd, nd);
  }
#endif
}

// --------------------------------------------------------------------------------------------- slot passes
// Flat pass over every live row.  FN(st, has_t, c0, gdz) is called once per row; `want_gdz` says whether the
// directional derivative gv.dz is needed.
template <int M, typename FN> GDEV void for_each_row(const IpmCtx<M>& c, bool want_gdz, FN&& fn) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NV = L::NV;
  const int N = c.N;
  G_PAR_FOR(it, N * L::SP) {
    if (it + G_NTHR < N * L::SP) g_prefetch_l1(c.sslot + (size_t)(it + G_NTHR) * SLOT_W);
    const int k = it / L::SP, s = it - k * L::SP;
    const double* x = sh_z<M>(c) + k * NV;
    double* st = c.sslot + (size_t)it * SLOT_W;
    if (T::HAS_TR && s == L::S_TR) {
      double v = -c.dow, gdz = 0.0;
      for (int i = 0; i < NX; ++i) { const double dxi = x[i] - c.Xp[k * NX + i]; v += dxi * dxi; gdz += 2.0 * dxi * sh_dz<M>(c)[k * NV + i]; }
      fn(st, true, v, gdz);
    } else {
      SpecEval o;
      spec_eval<M>(c, k, s, x, x + NX, o);
      if (!o.valid) continue;
      double gdz = 0.0;
      if (want_gdz) { const double* dv = sh_dz<M>(c) + k * NV + (o.is_u ? NX : 0) + o.i0; for (int a = 0; a < 4; ++a) if (a < o.n) gdz += o.gv[a] * dv[a]; }
      fn(st, o.has_t, o.c0, gdz);
    }
  }
  if (c.bmask != 0) {
    G_PAR_FOR(j, L::NBOX) {
      if (!((c.bmask >> (j >> 1)) & 1)) continue;
      const double* x = sh_z<M>(c) + (N - 1) * NV;
      const double gdz = ((j & 1) == 0 ? 1.0 : -1.0) * sh_dz<M>(c)[(N - 1) * NV + (j >> 1)];
      fn(c.bslot + (size_t)j * SLOT_W, false, box_c0<M>(c, j, x), gdz);
    }
  }
  if (T::WS > 0) {
    constexpr int WS = T::WS > 0 ? T::WS : 1;
    const double* __restrict__ orow = c.orow;
    double* __restrict__ ost = c.ost;
    const double* zs = sh_z<M>(c);
    const double* dzs = sh_dz<M>(c);
    const int nact = c.nact;
    for (int p = G_TID; p < nact; p += G_NTHR) {
      if (p + G_NTHR < nact) { g_prefetch_l1(orow + (size_t)(p + G_NTHR) * OROW_W); g_prefetch_l1(ost + (size_t)(p + G_NTHR) * SLOT_W); }
      const double* row = orow + (size_t)p * OROW_W;
      const int k = (int)row[4];
      const double* x = zs + k * NV;
      const double* dv = dzs + k * NV;
      double v = row[3], gdz = 0.0;
#pragma unroll
      for (int a = 0; a < WS; ++a) { v -= row[a] * x[a]; gdz -= row[a] * dv[a]; }
      fn(ost + (size_t)p * SLOT_W, true, v, gdz);
    }
  }
}

// One pass over the rows (see pair_step).  out[0..1]: largest primal / dual step to the boundary (modes 0, 1);
// out[2..4]: c0, c1, c2 (mode 1);  mode 2: installs the central-path floor for the next iteration.
template <int M> GDEV_NOINLINE void slot_steps(const IpmCtx<M>& c, int phase, double smu, int mode, double ap, double ad, double* out) {
  StepAcc acc; acc.amp = 1e300; acc.amd = 1e300; acc.c0 = 0.0; acc.c1 = 0.0; acc.c2 = 0.0; acc.np = 0.0;
  const double omega = c.omega;
  for_each_row<M>(c, true, [&](double* st, bool has_t, double c0, double gdz) {
    Pair q;
    pair_eval(st, has_t, c0, omega, smu, phase, q);
    pair_step(st, has_t, q, gdz, mode, ap, ad, acc);
  });
  if (mode != 2) {
    out[0] = -block_max(-acc.amp, c.red);
    out[1] = -block_max(-acc.amd, c.red);
    if (mode == 1) { out[2] = block_sum(acc.c0, c.red); out[3] = block_sum(acc.c1, c.red); out[4] = block_sum(acc.c2, c.red); }
  } else {
    const double ms = block_sum(acc.c0, c.red), np = block_sum(acc.np, c.red);
    if (G_TID == 0) c.floor_ = np > 0 ? 1e-4 * ms / np : 0.0;
    G_SYNC();
  }
}

// --------------------------------------------------------------------------------------------------- setup
#ifndef GUSTO_SLACK_START_SE3
#define GUSTO_SLACK_START_SE3 0.25
#endif
template <int M> GHD constexpr double slack_start() { return M == ASTROBEE_SE3 ? GUSTO_SLACK_START_SE3 : 1.0; }
// ... and how the penalty weight omega = lam + lam_t (dual feasibility of t) is split at the start: most soft rows end
// inactive (lam -> 0, lam_t -> omega), so starting lam at 0.1 omega instead of 0.5 omega saves another 0.8 Newton
// iterations on astrobeeSE3 (8.01 -> 7.23, solve 5.87 -> 5.34 ms; 0.02 would give 6.86 but starts badly centred)
#ifndef GUSTO_LAM_SPLIT_SE3
#define GUSTO_LAM_SPLIT_SE3 0.1
#endif
template <int M> GHD constexpr double slack_lam_split() { return M == ASTROBEE_SE3 ? GUSTO_LAM_SPLIT_SE3 : 0.5; }
template <int M> GDEV_NOINLINE void setup(IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV, ANZ = L::ANZ;
  const int N = c.N;
  // start point: X, U <- previous trajectory (set_start_value, scp_gusto.jl:100-102); multipliers 0
  G_PAR_FOR(it, N * NV) { const int k = it / NV, i = it - k * NV; sh_z<M>(c)[it] = i < NX ? c.Xp[k * NX + i] : c.Up[k * NU + i - NX]; }
  G_PAR_FOR(it, (N + 1) * NX) c.nu[it] = 0.0;
  // A_k on its sparsity pattern
  G_PAR_FOR(it, N * ANZ) { const int k = it / ANZ, e = it - k * ANZ; c.Ac[it] = c.A[(size_t)k * NX * NX + T::a_row(e) * NX + T::a_col(e)]; }
  // obstacle rows inside the toggle distance, compacted knot by knot (astrobee_se3.jl:293)
  if (T::WS > 0) {
    G_PAR_FOR(k, N) {
      int cnt = 0;
      for (int i = 0; i < c.n_obs; ++i) cnt += (c.rows[((size_t)k * c.n_obs + i) * 5 + 4] < c.toggle) ? 1 : 0;
      sh_seg<M>(c)[k + 1] = cnt;
    }
    G_SYNC();
    if (G_TID == 0) { sh_seg<M>(c)[0] = 0; for (int k = 0; k < N; ++k) sh_seg<M>(c)[k + 1] += sh_seg<M>(c)[k]; c.nact = sh_seg<M>(c)[N]; }
    G_SYNC();
  } else {
    G_PAR_FOR(k, N + 1) sh_seg<M>(c)[k] = 0;
    if (G_TID == 0) c.nact = 0;
    G_SYNC();
  }
  if (T::WS > 0) {
    G_PAR_FOR(k, N) {
      int p = sh_seg<M>(c)[k];
      const double* x = sh_z<M>(c) + k * NV;
      for (int i = 0; i < c.n_obs; ++i) {
        const double* row = c.rows + ((size_t)k * c.n_obs + i) * 5;
        if (!(row[4] < c.toggle)) continue;
        double* o = c.orow + (size_t)p * OROW_W;
        double v = row[3];
        for (int a = 0; a < 3; ++a) { o[a] = row[a]; if (a < T::WS) v -= row[a] * x[a]; }
        o[3] = row[3]; o[4] = (double)k;
        slot_init(c.ost + (size_t)p * SLOT_W, true, true, v, c.omega, slack_start<M>(), slack_lam_split<M>());
        ++p;
      }
    }
  }
  // special rows: slacks one unit inside
  G_PAR_FOR(it, N * L::SP) {
    const int k = it / L::SP, s = it - k * L::SP;
    const double* x = sh_z<M>(c) + k * NV;
    double* st = c.sslot + (size_t)it * SLOT_W;
    if (T::HAS_TR && s == L::S_TR) slot_init(st, true, true, tr_c0<M>(c, k, x), c.omega, slack_start<M>(), slack_lam_split<M>());
    else { SpecEval o; spec_eval<M>(c, k, s, x, x + NX, o); slot_init(st, o.valid, o.has_t, o.c0, c.omega, slack_start<M>(), slack_lam_split<M>()); }
  }
  G_PAR_FOR(j, L::NBOX) {
    const bool valid = (c.bmask >> (j >> 1)) & 1;
    slot_init(c.bslot + (size_t)j * SLOT_W, valid, false, valid ? box_c0<M>(c, j, sh_z<M>(c) + (N - 1) * NV) : 0.0, c.omega);
  }
  G_SYNC();
}

// ------------------------------------------------------------------------------------------------ driver
// scratch: IpmLayout<M>::scratch_doubles() doubles of global memory owned by this instance (16-byte aligned).
// smem:    IpmLayout<M>::smem_doubles() doubles of shared memory.
// On exit Xn/Un of the instance hold the solution and info[IPM_NINFO] = {status, iters, res, mu, obj,...}.
template <int M>
GDEV void ipm_solve_instance(const BatchDesc& d, const BatchPtrs& p, const IpmParams& prm, int b, double* scratch,
                             double* smem, double* info) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV, NN = L::NN;
  const int N = d.N;
  // The context is ONE struct per instance in shared memory (filled by thread 0): phase functions read it with
  // shared loads instead of per-thread local-memory copies.
  static_assert(sizeof(IpmCtx<M>) <= (size_t)L::CTX_DOUBLES * sizeof(double), "IpmCtx does not fit its shared-memory slot");
  IpmCtx<M>& c = *reinterpret_cast<IpmCtx<M>*>(smem);
  smem += L::CTX_DOUBLES;
  if (G_TID == 0) {
    c.d = &d; c.rp = d.rp; c.N = N; c.n_obs = (T::WS > 0) ? d.n_obs : 0; c.b = b; c.nact = 0;
    c.h = p.tf[b] / (N - 1); c.hh = 0.5 * c.h; c.omega = p.omega[b]; c.Delta = p.delta[b];
    c.toggle = c.Delta / 8.0 + d.rp[RP_CLEAR]; c.eps = d.sp[SP_EPS]; c.dp = prm.delta_p; c.dd = prm.delta_d;
    c.dow = c.Delta / c.omega; c.eow = c.eps / c.omega;
    c.pmask = 0; c.bmask = 0;
    for (int i = 0; i < NX; ++i) { if (d.goal_type[i] == GOAL_POINT) c.pmask |= 1 << i; if (d.goal_type[i] == GOAL_BOX) c.bmask |= 1 << i; }
    c.Xp = p.Xp + (size_t)b * N * NX; c.Up = p.Up + (size_t)b * N * NU;
    c.A = p.A + (size_t)b * N * NX * NX; c.g = p.g + (size_t)b * N * NX;
    c.rows = p.rows + (size_t)b * N * d.n_obs * 5;
    c.x_init = p.x_init + (size_t)b * NX; c.goal_lo = p.goal_lo + (size_t)b * NX; c.goal_hi = p.goal_hi + (size_t)b * NX;
    {
      double Bm[NX * NU];
      for (int i = 0; i < NX * NU; ++i) Bm[i] = 0.0;
      dyn_B<M>(d.rp, Bm);
      for (int a = 0; a < NU; ++a) c.bv[a] = Bm[T::b_row(a) * NU + a];
    }
    const size_t nz = L::rnd((size_t)N * NV), ne = L::rnd((size_t)(N + 1) * NX), no = c.n_obs;
    double* q = scratch;
    c.r = q; q += nz; c.t1 = q; q += nz; c.res = q; q += nz;
    c.nu = q; q += ne; c.dnu = q; q += ne; c.rnu = q; q += ne; c.resnu = q; q += ne;
    c.Ac = q; q += L::rnd((size_t)N * L::ANZ);
    c.sslot = q; q += L::rnd((size_t)N * L::SP * SLOT_W);
    c.bslot = q; q += (size_t)L::NBOX * SLOT_W;
    c.ost = q; q += L::rnd((size_t)N * no * SLOT_W);
    c.orow = q; q += L::rnd((size_t)N * no * OROW_W);
    c.kd = q; q += L::rnd((size_t)N * L::KDW);
    c.fac = q; q += (size_t)(N + 1) * 2 * L::GT;
    c.z = smem; c.dz = c.z + L::rnd((size_t)N * NV); c.sy = c.dz + L::rnd((size_t)N * NV); c.ring = c.sy + L::rnd((size_t)(N + 1) * NX);
    c.red = c.dz + L::work_doubles(N);
    c.seg = reinterpret_cast<int*>(c.red + G_NTHR + 16);
    c.tab = reinterpret_cast<unsigned char*>(c.red + G_NTHR + 16 + L::seg_doubles(N));
    for (int i = 0; i < 5; ++i) c.prof[i] = 0;
    for (int i = 0; i < RING_STAGES; ++i) g_mbar_init(&c.ring_bar[i], 1);
    c.ring_o = 0;
    c.floor_ = 0.0;
    unsigned char* t2 = c.tab + NX * (NX + 1);
    int n = 0;
    for (int i = 0; i < NX; ++i) for (int g = 0; g < L::NG; ++g) if (g * L::CG <= i) { t2[n] = (unsigned char)i; t2[L::NLT + n] = (unsigned char)g; ++n; }
    unsigned char* mt = c.tab + L::TAB_MODEL;                 // a_row | a_col | blk_of | ctrl_of (0xff: no control drives the row)
    for (int e = 0; e < L::ANZ; ++e) { mt[e] = (unsigned char)T::a_row(e); mt[L::ANZ + e] = (unsigned char)T::a_col(e); }
    for (int i = 0; i < NX; ++i) { mt[2 * L::ANZ + i] = (unsigned char)T::XB_of(i); mt[2 * L::ANZ + NX + i] = (unsigned char)ctrl_of_row<M>(i); }
  }
  G_SYNC();
  G_PAR_FOR(t, NX * (NX + 1) / 2) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    c.tab[t] = (unsigned char)i; c.tab[NX * (NX + 1) / 2 + t] = (unsigned char)(t - i * (i + 1) / 2);
  }

  // The solve is attempted with the configured primal regularisation; an attempt that breaks down (non-positive
  // pivot, NaN) or stalls (no progress of the residual for 12 Newton iterations) is restarted from the same start
  // point with delta_p x10, at most twice.  The oracle factorises the full KKT matrix (LU) and only regularises on a
  // failed factorisation (ipm.py); the Schur-complement form used here needs (H + delta_p I)^-1, and on problems
  // whose H is singular in many directions (no trust region: astrobeeSE3manifold) the right delta_p is
  // instance-dependent.
  int status = IPM_ITERATION_LIMIT, it_done = 0;
  long long cyc_asm = 0, cyc_fac = 0, cyc_sol = 0, cyc_slot = 0, tc0;
  double res = 1e300, mu = 0;
  const double scd = 1.0 + c.omega;
  for (int attempt = 0; attempt < 3; ++attempt) {
  if (attempt > 0) {
    G_SYNC();
    if (G_TID == 0) { c.dp *= 10.0; c.floor_ = 0.0; }
    G_SYNC();
#ifdef GUSTO_HOSTSIM
    if (getenv("GUSTO_HOSTSIM_VERBOSE")) printf("  restart with delta_p = %.1e\n", c.dp);
#endif
  }
  setup<M>(c);
  status = IPM_ITERATION_LIMIT;
  double best = 1e300;
  int best_it = 0;
  for (int iter = 1; iter <= prm.max_iter; ++iter) {
    G_CTA_RESYNC();
    ++it_done;
    Resid R;
    bool blocks_ok = true;
    tc0 = g_clock();
    assemble<M>(c, 0, 0.0, &R, &blocks_ok);
    cyc_asm += g_clock() - tc0;
    mu = R.mu;
    res = R.rz / scd;
    res = R.rp > res ? R.rp : res; res = R.rc > res ? R.rc : res; res = mu > res ? mu : res;
#ifdef GUSTO_HOSTSIM
    if (getenv("GUSTO_HOSTSIM_VERBOSE")) printf("  ipm %3d rd=%.2e rp=%.2e rc=%.2e mu=%.2e\n", iter, R.rz, R.rp, R.rc, mu);
#endif
    if (res <= prm.tol) { status = IPM_OPTIMAL; break; }
    if (!(res == res) || res > 1e200) { status = IPM_NUMERICAL; break; }
    if (res < 0.5 * best) { best = res; best_it = iter; }
    else if (iter - best_it >= 12) break;                       // stalled
    tc0 = g_clock();
    const bool fac_ok = factorize<M>(c);
    cyc_fac += g_clock() - tc0;
    if (!(fac_ok && blocks_ok)) { status = IPM_NUMERICAL; break; }   // non-positive pivot
    // predictor
    tc0 = g_clock();
    kkt_solve_refined<M>(c, 0);
    cyc_sol += g_clock() - tc0;
    double am[5];
    tc0 = g_clock();
    slot_steps<M>(c, 0, 0.0, 1, 0, 0, am);
    double a_aff = am[0] < am[1] ? am[0] : am[1];
    a_aff = a_aff < 1.0 ? a_aff : 1.0;
    cyc_slot += g_clock() - tc0;
    double mu_aff = am[2] + a_aff * (am[3] + a_aff * am[4]);
    mu_aff = R.npair > 0 ? mu_aff / R.npair : 0.0;
    double sigma = mu > 0 ? (mu_aff / mu) : 0.0;
    sigma = sigma * sigma * sigma;
    double smu = sigma * mu;
    smu = smu > 0.1 * prm.tol ? smu : 0.1 * prm.tol;
    // corrector
    tc0 = g_clock();
    assemble<M>(c, 1, smu, &R, &blocks_ok);
    cyc_asm += g_clock() - tc0;
    tc0 = g_clock();
    // far from the solution the unrefined direction is accurate enough (its KKT residual is ~1e-7 of the rhs): refine
    // only once the complementarity gap is small
    kkt_solve_refined<M>(c, (T::HAS_TR && mu > 1e-5) ? 0 : prm.nref);
    cyc_sol += g_clock() - tc0;
    tc0 = g_clock();
    slot_steps<M>(c, 1, smu, 0, 0, 0, am);
    double tau = 0.995;
    if (mu < 1.0) { tau = 1.0 - mu; tau = tau > 0.995 ? tau : 0.995; tau = tau < 0.999999 ? tau : 0.999999; }
    double ap = tau * am[0], ad = tau * am[1];
    ap = ap < 1.0 ? ap : 1.0; ad = ad < 1.0 ? ad : 1.0;
    slot_steps<M>(c, 1, smu, 2, ap, ad, am);
    G_PAR_FOR(it, N * NV) sh_z<M>(c)[it] += ap * sh_dz<M>(c)[it];
    {
      double* __restrict__ nu = c.nu;
      const double* __restrict__ dnu = c.dnu;
      const int ne = (N + 1) * NX;
#pragma unroll 4
      for (int it = G_TID; it < ne; it += G_NTHR) nu[it] += ad * dnu[it];
    }
    G_SYNC();
    cyc_slot += g_clock() - tc0;
  }
  if (status == IPM_ITERATION_LIMIT && res <= 1e3 * prm.tol) status = IPM_OPTIMAL;
  {   // a NaN/Inf anywhere in the iterate is a numerical failure, never an answer
    double badz = 0.0;
    G_PAR_FOR(it, N * NV) { const double v = sh_z<M>(c)[it]; if (!(v == v) || fabs(v) > 1e100) badz = 1.0; }
    if (block_max(badz, c.red) > 0.0) status = IPM_NUMERICAL;
  }
  if (status == IPM_OPTIMAL) break;
  }
  // ---- write the candidate trajectory and the objective (cost + omega * sum t)
  double obj = 0;
  G_PAR_FOR(k, N) {
    const double wk = (k == 0 || k == N - 1) ? 0.5 * c.h : c.h;
    for (int i = 0; i < NX; ++i) p.Xn[((size_t)b * N + k) * NX + i] = sh_z<M>(c)[k * NV + i];
    for (int i = 0; i < NU; ++i) { const double uv = sh_z<M>(c)[k * NV + NX + i]; p.Un[((size_t)b * N + k) * NU + i] = uv; obj += wk * uv * uv; }
  }
  {
    const double omega = c.omega;
    for_each_row<M>(c, false, [&](double* st, bool has_t, double, double) { if (has_t) obj += omega * st[2]; });
  }
  // SCPS.dual (scp_gusto.jl:116, get_dual_jump): row 0 of Aeq is  x_0 = x_init  and the Lagrangian is f + nu'(Aeq z - b)
  if (p.dual) G_PAR_FOR(i, NX) p.dual[(size_t)b * NX + i] = c.nu[i];
  obj = block_sum(obj, c.red);
  if (G_TID == 0) {
    info[0] = (double)status; info[1] = (double)it_done; info[2] = res; info[3] = mu; info[4] = obj;
#if !defined(GUSTO_PROF_MODE) || GUSTO_PROF_MODE == 0
    info[5] = (double)(cyc_asm + cyc_slot); info[6] = (double)cyc_fac; info[7] = (double)cyc_sol;   // SM cycles per phase
#elif GUSTO_PROF_MODE == 1      // developer builds: finer split of the factorisation
    info[5] = (double)c.prof[0]; info[6] = (double)c.prof[1]; info[7] = (double)cyc_asm;
#elif GUSTO_PROF_MODE == 2      // ... and of the solves
    info[5] = (double)c.prof[2]; info[6] = (double)c.prof[3]; info[7] = (double)c.prof[4];
#else                           // ... and of one sweep step: elimination | wait for the stager + C1 | C2 (packed: X + loop barrier in info[3])
    info[5] = (double)c.prof[0]; info[6] = (double)(c.prof[1] + c.prof[2]); info[7] = (double)c.prof[3]; info[3] = (double)c.prof[4];
#endif
  }
}

}  // namespace gusto

#ifndef GUSTO_HOSTSIM
#undef block_sum
#undef block_max
#pragma pop_macro("G_SYNC")
#pragma pop_macro("G_NTHR")
#pragma pop_macro("G_TID")
#endif
