// K3: batched convex-subproblem solve, one group of threads per problem instance.
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
//     attitude | rate, Traits<M>::XB_*), so  Hx_k = blockdiag(<=4x4 blocks) + kappa_tr g g';
//   * A_k = df/dx is used through its static sparsity pattern (Traits<M>::a_row/a_col, 27 of 144 entries for SE3);
//   * the Newton system  [[H, Aeq'], [Aeq, 0]] [dz; dnu] = [r; rnu]  is solved by a PRIMAL RICCATI RECURSION over the
//     knots (round 2; round 1 factorised the Schur complement Aeq (H + dp I)^-1 Aeq', which needs a regularised inverse
//     of Hx -- Hx is singular wherever no soft row is active, on astrobeeSE3manifold in most directions).  The
//     trapezoid row  E_j x_{j-1} + G u_{j-1} - F_j x_j + G u_j = rho_j  (E_j = I + h/2 A_{j-1}, F_j = I - h/2 A_j,
//     G = h/2 B) is implicit in x_j and couples u_{j-1} AND u_j; with the shifted state  s_j = x_j - Gam_j u_j,
//     Gam_j = F_j^-1 G (Gam_0 = 0) it becomes  s_{j+1} = Ah_j s_j + Bh_j u_j + ch_j  with  Ah_j = F_{j+1}^-1 E_{j+1},
//     Bh_j = Ah_j Gam_j + Gam_{j+1},  ch_j = -F_{j+1}^-1 rho_{j+1},  and the stage cost picks up the cross term
//     S = Gam'Hx, R = Hu + Gam'Hx Gam.  Ah, Bh, Gam are computed once per solve (setup_dynamics); per Newton iteration
//       Lam_k = R_k + Bh_k'P_{k+1}Bh_k (n_u x n_u, >= Hu_k > 0: the ONLY matrix ever factorised, Cholesky in registers),
//       M_k = S_k + Bh_k'P_{k+1}Ah_k,  K_k = Lam_k^-1 M_k,  P_k = Hx_k + Ah_k'P_{k+1}Ah_k - M_k'K_k,
//     4 group barriers per knot; the backward vector pass of the predictor is fused into the same sweep.  A solve with a
//     new right-hand side is a backward and a forward chain of ONE n_x x n_x mat-vec per knot with the closed-loop
//     tiles  Acl_k = Ah_k - Bh_k K_k  prefetched into registers three knots ahead, plus parallel per-knot passes.  No primal or dynamics regularisation and no iterative refinement; the PointGoal rows  M x_{N-1} = goal
//     are a quadratic penalty  w_N |.|^2 / 2  = dual regularisation 1/w_N of THOSE rows only (w_N = 1e8 + 1e4 omega;
//     the row error after a step is dnu_N / w_N and contracts by ~1e-6 per Newton iteration);
//   * equality multipliers of the original rows are recovered from the Riccati costates:  dnu_j = F_j^-T (P_j s_j - p_j).
// tools/riccati_proto.py is the NumPy prototype of this arithmetic inside the oracle's IPM (same Newton counts on all
// four models, omega = 1 .. 1e10).
// A first-order splitting (ADMM, prototyped in tools/admm_proto.py) was rejected: GuSTO's accept test compares
// soft rows against eps = 1e-6 (astrobee_se3.jl:31, scp_gusto.jl:318-327), and with a trapezoid double
// integrator over 70 s ADMM needs >2000 iterations for 1e-6 residuals while this method needs 8-25 for 1e-8.
//
// Memory: the iterate z, the direction dz and the costate chain live in shared memory; multipliers, slack records
// (compacted to the obstacle rows inside the toggle distance), per-knot Hessian blocks, the per-solve dynamics records
// and the per-iteration Riccati factors (Acl, K, chol(Lam), P) live in a per-instance global scratch.
// All arithmetic is FP64 (the reference is Float64 throughout).
//
// This header is included once per ALGORITHM (no include guard): GUSTO_IPM_ALG = 0 (default) is the GuSTO subproblem described
// above and lives in the inline namespace gusto::ipm_gusto; GUSTO_IPM_ALG = 1 is the TrajOpt subproblem of solve_trajopt_jump!
// (/root/reference/src/scp/scp_trajopt.jl:159-279) in gusto::ipm_trajopt -- same kernel skeleton, compiled
// separately so that the GuSTO kernel's code does not change by a single instruction.  What differs under TrajOpt:
//   * the state trust region is a HARD row  |x_k - xp_k|^2 - s <= 0  (:165-173; `delta` carries s, `omega` carries mu);
//   * the control balls are mu-penalised hinge rows like every other inequality (:222-233); toggle distance = clearance + 1 (:65);
//   * the dynamics rows are l1-PENALISED (:257-275):  d_j(z) - p_j + n_j = 0,  p, n >= 0,  cost mu 1'(p + n).  Eliminating
//     (p, n) and their multipliers leaves the Newton system  [[H, Aeq'], [Aeq, -D]]  with a positive diagonal
//     D_j = p/lam_p + n/lam_n on the dynamics rows: in the shifted-state recursion row j becomes  s_j = Ah s_{j-1} + Bh u_{j-1} +
//     ch_{j-1} + w_j  with a "process noise" w_j = -F_j^-1 D_j dnu_j of cost  1/2 w' W_j^-1 w,  W_j = F_j^-1 D_j F_j^-T.  Minimising
//     over w_j first replaces the cost-to-go of knot j by  P~ = P - P V (I + V'P V)^-1 V'P,  p~ = N'p,  N = I - V (I + V'PV)^-1 V'P
//     (V = F^-1 D^1/2; one n_x x n_x Cholesky per knot: the noise phase of riccati_factor) and the realised state is
//     s_j = N_j s^_j + W_j p~_j.  The chains run with  G_k = N_{k+1} Acl_k  (forward on the realised states, backward on the
//     pre-noise costates  w_k = p_k - P_k ch_{k-1}), so chain_forward / chain_backward are shared with GuSTO unchanged.
#include "common.cuh"
#include "models.cuh"
#include "evaluate.cuh"   // block_sum / block_max
#ifndef GUSTO_IPM_ALG
#define GUSTO_IPM_ALG 0
#endif
#if GUSTO_IPM_ALG == 0
#define GUSTO_IPM_NS_OPEN inline namespace ipm_gusto {
#else
#define GUSTO_IPM_NS_OPEN namespace ipm_trajopt {
#endif
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

#ifndef GUSTO_IPM_COMMON_DEFINED
#define GUSTO_IPM_COMMON_DEFINED
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
#define GUSTO_PROF_CHAINS 0
#else
#define GUSTO_PROF_CHAINS 1
#endif

struct IpmParams {
  int max_iter;      // Newton iterations cap
  double tol;        // max(|r_dual|/(1+omega), |r_eq|, |r_ineq|, mu) <= tol
  double wn_base;    // terminal (PointGoal) penalty  w_N = wn_base + wn_omega * omega
  double wn_omega;
  // centred start (slot_init): mu0 = max(mu0_a omega, mu0_b omega vmax), vmax = the largest violated soft-row value at the start point;
  // mu0 > mu0_cap (a poor start point) or mu0_a <= 0: the tuned default start.  ipm_default_start() fills the defaults.
  double mu0_a, mu0_b, mu0_cap;
  // ... scaled down with the start point's primal infeasibility rp0 (a nearly converged SCP iteration starts next to its optimum):
  // mu0 *= min(1, rp0 / mu0_rp), not below mu0_lo omega; only with an unshrunk trust region and every live obstacle row mu0_smin clear
  double mu0_rp, mu0_lo, mu0_smin;
};
#ifndef GUSTO_MU0_A
#define GUSTO_MU0_A 5e-5
#endif
#ifndef GUSTO_MU0_B
#define GUSTO_MU0_B 1.0
#endif
#ifndef GUSTO_MU0_CAP
#define GUSTO_MU0_CAP 1e-3
#endif
#ifndef GUSTO_MU0_RP
#define GUSTO_MU0_RP 0.1
#endif
#ifndef GUSTO_MU0_LO
#define GUSTO_MU0_LO 1e-9
#endif
#ifndef GUSTO_MU0_SMIN
#define GUSTO_MU0_SMIN 5e-2
#endif
inline void ipm_default_start(IpmParams& prm) {
  prm.mu0_a = GUSTO_MU0_A; prm.mu0_b = GUSTO_MU0_B; prm.mu0_cap = GUSTO_MU0_CAP;
  prm.mu0_rp = GUSTO_MU0_RP; prm.mu0_lo = GUSTO_MU0_LO; prm.mu0_smin = GUSTO_MU0_SMIN;
}

// IPM_ALMOST_OPTIMAL: stalled within 1e3*tol of the tolerance AND far below the SCP's own soft-row threshold eps -- the
// MOI.ALMOST_LOCALLY_SOLVED the reference accepts next to OPTIMAL (scp_gusto.jl:107); the host records it as such.
enum : int { IPM_OPTIMAL = 0, IPM_ITERATION_LIMIT = 1, IPM_NUMERICAL = 2, IPM_ALMOST_OPTIMAL = 3 };
constexpr int SLOT_W = 6;     // s, lam, t, lamb, pa (ds*dlam of the predictor), pb (dt*dlamb of the predictor)
constexpr int OROW_W = 5;     // compacted obstacle row: nhat[3], off, knot
constexpr int IPM_NINFO = 8;  // status, iterations, residual, mu, objective, cycles: assemble+slots, factorize, kkt solves

#ifndef GUSTO_CHAIN_D
#define GUSTO_CHAIN_D 4
#endif
#ifndef GUSTO_RD_FACTOR
#define GUSTO_RD_FACTOR 0.01
#endif
constexpr int CHAIN_STAGES = GUSTO_CHAIN_D;

GHD constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
#endif   // GUSTO_IPM_COMMON_DEFINED

GUSTO_IPM_NS_OPEN
constexpr bool kTO = GUSTO_IPM_ALG == 1;     // TrajOpt subproblem (see the file header)

template <int M> struct IpmLayout {
  using T = Traits<M>;
  static constexpr int NX = T::NX, NU = T::NU, NV = NX + NU, NN = NX * NX, ANZ = T::ANZ;
  // Tiles of the Riccati sweep are operands of 8x8x4 FP64 tensor-core products (g_tile_job, common.cuh): contraction lengths
  // are padded to a multiple of 4 (KP, KU), row counts to a multiple of 8 (NXP, NUP, RXS), and the row stride is the smallest
  // value >= the padded length with stride mod 16 in {4, 12} (conflict-free fragment loads).
  static constexpr int KP = (NX + 3) & ~3, KU = (NU + 3) & ~3;
  GHD static constexpr int ld_of(int kp) { int ld = kp; while ((ld & 15) != 4 && (ld & 15) != 12) ++ld; return ld; }
  static constexpr int LDT = ld_of(KP);                    // 12 (SE3), 20 (manifold), 12 (freeflyer), 4 (dubins)
  static constexpr int LDU = ld_of(KU);                    // row stride of the n_x x n_u tiles (K', Y', M', S')
  static constexpr int NXP = (NX + 7) & ~7, NUP = 8, RXS = (NX + NU + 1 + 7) & ~7;
  static constexpr int TILE = NXP * LDT;
  static constexpr int GT = NX * LDT;                      // closed-loop tile in global memory (one bulk copy per chain step)
  static constexpr int NTX = NX * (NX + 1) / 2, NTU = NU * (NU + 1) / 2;
  // decoupled dynamics blocks (Traits<M>::DSPLIT): state range [dlo(i), dhi(i)) of the block holding coordinate i
  GHD static constexpr int dlo(int i) { return (T::DSPLIT > 0 && i >= T::DSPLIT) ? T::DSPLIT : 0; }
  GHD static constexpr int dhi(int i) { return (T::DSPLIT > 0 && i < T::DSPLIT) ? T::DSPLIT : NX; }
  GHD static constexpr bool dsame(int i, int j) { return dlo(i) == dlo(j); }
  GHD static constexpr bool dcheck() { for (int e = 0; e < ANZ; ++e) if (!dsame(T::a_row(e), T::a_col(e))) return false; return true; }
  static constexpr int DMAX = T::DSPLIT > 0 ? (T::DSPLIT > NX - T::DSPLIT ? T::DSPLIT : NX - T::DSPLIT) : NX;
  // special (non-obstacle) slots of a knot
  static constexpr int S_TR = 0;
  static constexpr int S_NORM = T::HAS_TR;
  static constexpr int S_LIN = S_NORM + T::NNORM;
  static constexpr int S_QUAT = S_LIN + T::NLIN;           // hinge, then hard row
  static constexpr int S_BALL = S_QUAT + 2 * T::HAS_QUAT;
  static constexpr int SP = S_BALL + T::NBALL;
  static constexpr int NBOX = 2 * NX;                      // goal-box rows (upper, lower per coordinate), knot N-1 only
  static constexpr int XPK = T::XB_pk(T::XB_CNT), UPK = T::UB_pk(T::UB_CNT);
  // per-knot Hessian record: Hb[XPK] (packed x-blocks) | Hub[UPK] (packed u-blocks) | w[NX] = sqrt(kap_tr) * 2 (x - xp)
  static constexpr int KD_HB = 0, KD_HU = XPK, KD_W = XPK + UPK, KDW = (KD_W + NX + 1) & ~1;
  // per-knot dynamics record, constant over a solve (setup_dynamics): rows of Ah' (NX), Bh' (NU), Gam' (NU), each LDT long
  static constexpr int CR_AT = 0, CR_BT = NX, CR_GT = NX + NU, CRR = NX + 2 * NU, CRW = CRR * LDT;
  // per-knot Riccati factor of one Newton iteration: K' [NX][LDU], packed P
  static constexpr int KTW = NX * LDU, LPW = (NTU + 1) & ~1, PKW = (NTX + 1) & ~1;   // lp: packed lower L^-1 of Lam = L L'
  GHD static size_t rnd(size_t v) { return (v + 1) & ~(size_t)1; }   // keep every array 16-byte aligned
  // Global scratch is knot-minor ("field-major"): element e of knot k of a per-knot record lives at  e * NP + k  (NP = N rounded up
  // to 4, so every field row starts on a 32-byte sector), and likewise  e * NE + j  for the N + 1 equality rows and  f * PP + p  for
  // the compacted obstacle rows.  The per-knot passes run one THREAD per knot: with this layout a warp's load of one field is 32
  // consecutive doubles (8 sectors) instead of 32 separate sectors -- those passes were bound by L1 wavefronts / exposed global
  // latency (ncu round 2: ric_forward 11 % of the kernel's warp time on 2 % of its instructions).  The knot-serial sweep gathers
  // its inputs with 8-byte asynchronous copies one knot ahead and scatters its factor records with plain stores; only the
  // closed-loop tiles Acl_k (whole-tile copies of the chains) and the dynamics records of the sweep stay knot-major.
  GHD static int np_of(int N) { return (N + 3) & ~3; }
  GHD static int ne_of(int N) { return (N + 1 + 3) & ~3; }
  GHD static int pp_of(int N, int n_obs) { const int v = N * (T::WS > 0 ? n_obs : 0); return v > 0 ? ((v + 3) & ~3) : 4; }
  static constexpr int GSW = NU * NX;                       // Gam' / Bh' entries per knot in the field-major copies
  // TrajOpt: record of one l1-penalised dynamics row entry (row j, coordinate i), field-major [field][i][j]:
  //   p, lam_p, n, lam_n, pa = dp dlam_p, pb = dn dlam_n (predictor products)
  static constexpr int DSLOT_W = 6;
  // doubles of global scratch per instance
  GHD static size_t scratch_doubles(int N, int n_obs) {
    const size_t np = np_of(N), ne = ne_of(N), pp = pp_of(N, n_obs), nn = (size_t)N;
    return 3 * np * NV /* r ra rb */ + 3 * ne * NX /* nu dnu rnu */ + ne * NX /* gsum */ + np * NX /* xps */ + np * SP * SLOT_W + (size_t)NBOX * SLOT_W +
           pp * SLOT_W + pp * OROW_W + np * KDW + np * NN /* Fi */ + nn * CRW + 2 * np * GSW /* gs bs */ + nn * GT /* Acl */ + np * NX * NU /* K */ +
           np * NTU /* lp */ + np * NTX /* P */ + 2 * np * NX /* psi ch */ + np * NU /* kap */ +
           (kTO ? ne * NX * (DSLOT_W + 1) /* dslot dd */ + np * NTX /* P~ */ + np * NN /* N */ : 0);
  }
  // shared-memory tiles of the Riccati sweep (doubles); they alias the direction dz
  static constexpr int XSR = RXS + NUP;                     // staged dynamics record: Ah' | Bh' | ch | 0.. | Gam' (row RXS) | 0..
  static constexpr int F_XS = 0;                            // two of them
  static constexpr int F_PA = F_XS + 2 * XSR * LDT;         // P_{k+1} and the W / P_k under construction, alternating
  static constexpr int F_PB = F_PA + TILE;
  static constexpr int F_Z1 = F_PB + TILE;                  // (P Ah)' rows | (P Bh)' rows | (P ch)' ; later Y' = (L^-1 [M | rt])'
  static constexpr int F_QD = F_Z1 + RXS * LDT;             // dense Hx
  static constexpr int F_SM = F_QD + TILE;                  // S' = Hx Gam, then M' (row NX: rt), then K' (row NX: kap)
  static constexpr int F_LM = F_SM + NXP * LDU;             // Lam
  static constexpr int F_LI = F_LM + NUP * LDU;             // L^-1 (lower triangular)
  static constexpr int F_HU = F_LI + NUP * LDU;             // dense Hu
  static constexpr int SBW = (KDW + NV + NX + 1) & ~1;      // staged inputs of a knot: Hessian record | rhs | ch
  static constexpr int F_SB = F_HU + NUP * LDU;             // two of them
  static constexpr int F_VEC = F_SB + 2 * SBW;              // q[2][KP] | ru[2][KU] | pit[KP]
  static constexpr int V_Q = 0, V_RU = 2 * KP, V_PIT = V_RU + 2 * KU, VECW = V_PIT + KP;
  // TrajOpt noise phase.  Its tiles live in regions of the sweep that are dead while it runs (it is the first thing of a knot
  // step), so that the group's slice of shared memory stays small enough for 7 groups per SM like GuSTO's:
  //   F_j^-1                   in S' | Lam      (written in phases A / B)
  //   W, then P~               in the first rows of the OTHER staged dynamics record (re-staged for knot k - 1 right after)
  //   [I + W P | I]            in Z1             (written in phase A)
  //   N = (I + W P)^-1         in the tile P_k is built in (phase C); the chain tile G_k = N_{k+1} Acl_k at the end of the step
  //                            re-reads N from global scratch (c.nm)
  //   Acl_k (phase D)          in the tile of P_{k+1} / M' (dead after phase C2)
  // Nothing else relies on those regions: only rows / columns that a later phase rewrites completely, or whose products are
  // masked or meet exact zeros of the other operand, are touched (L^-1, whose rows beyond n_u must STAY zero, is not).  Own
  // storage: the vectors D_j, p~, w.
  static constexpr int LDN = NX | 1;                        // odd row stride: column walks are conflict-free
  static constexpr int NTILE = NX * LDN;
  static constexpr int F_NV = F_SM, F_NT = F_Z1, F_NVEC = F_VEC + VECW;
  static_assert(!kTO || (NTILE <= (NXP + NUP) * LDU && NTILE <= (NX + NU) * LDT && 2 * NX * NX <= RXS * LDT && NTILE <= TILE),
                "noise-phase tiles do not fit the regions they alias");
  static constexpr int FAC_DOUBLES = kTO ? F_NVEC + 3 * KP + 2 : F_VEC + VECW;
  static constexpr int GJ_ROWS = (64 / NX) * NX;            // setup_dynamics: (knot, row) pairs of one Gauss-Jordan batch
  static constexpr int GJ_DOUBLES = GJ_ROWS * 2 * NX;
  static_assert(NXP * LDU <= RXS * LDT && GJ_DOUBLES <= FAC_DOUBLES, "aliased tiles do not fit");
  GHD static int work_doubles(int N) {      // dz + the chain ring, also the tiles of the sweep
    const int w = (int)rnd((size_t)N * NV) + CHAIN_STAGES * GT;
    return (w > FAC_DOUBLES ? w : FAC_DOUBLES) + 2;
  }
  // byte tables with run-time indices: a_row[ANZ], a_col[ANZ], blk_of[NX]
  static constexpr int TAB_DOUBLES = (2 * ANZ + NX + 7) / 8;
  GHD static int seg_doubles(int N) { return (N + 4) / 2 + 2; }
  static constexpr int CTX_DOUBLES = kTO ? 88 : 80;                   // the per-instance context struct (IpmCtx) lives in shared memory too
  GHD static int smem_doubles(int N, int nthr) {
    return CTX_DOUBLES + (int)rnd((size_t)N * NV) + work_doubles(N) + (int)rnd((size_t)(N + 1) * NX) + nthr + 16 + seg_doubles(N) + TAB_DOUBLES;
  }
};

template <int M> struct IpmCtx {
  using L = IpmLayout<M>;
  static constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const BatchDesc* d;
  const double* rp;
  int N, n_obs, b, nact, pmask, bmask, reset;   // reset: slack reset on (slack_reset_on<M>() and no BoxGoal rows)
  int NP, NE, PP;         // field strides of the knot-minor scratch arrays (IpmLayout::np_of / ne_of / pp_of)
  double h, hh, omega, Delta, toggle, eps, wN;
  double dow, eow;         // Delta / omega, eps / omega
  double mu0_a, mu0_b, mu0_cap, mu0_rp, mu0_lo, mu0_smin;   // centred start (IpmParams)
  mutable double floor_;   // pending central-path floor of the complementarity pairs (see pair_floor)
  const double *Xp, *Up, *Ac, *g, *rows, *x_init, *goal_lo, *goal_hi;   // Ac, g, rows: the linearize kernel's blocks, knot-minor, read in place
  double bv[NU];          // B has one entry per column: B[b_row(a)][a] = bv[a]
  // global scratch
  double *nu, *dnu, *r, *rnu, *sslot, *bslot, *ost, *orow, *kd;
  double *fi, *cr, *acl, *kt, *lp, *pk, *psi, *ch, *kap;
  double *xps, *gsum, *gs, *bs;   // Xp field-major | h/2 (g_{j-1} + g_j) per equality row | Gam', Bh' field-major
  double *ra, *rb;                // corrector right-hand side  r_corr = r + smu ra + rb  (predictor_pass)
  double *dslot, *dd, *pkt, *nm;  // TrajOpt: l1 records of the dynamics rows | D_j | P~_k | N_k = (I + W_k P_k)^-1, field-major
  // shared
  double *z, *dz, *vp, *red;
  int* seg;
  mutable long long prof[4];      // thread-0 cycle counters: Riccati sweep, chains, per-knot passes, (spare)
  unsigned char* tab;     // a_row[ANZ] | a_col[ANZ] | blk_of[NX]
};

// shared-memory members, with the address space made known to the compiler
template <int M> GDEV double* sh_z(const IpmCtx<M>& c) { double* p = c.z; G_ASSUME_SHARED(p); return p; }
template <int M> GDEV double* sh_dz(const IpmCtx<M>& c) { double* p = c.dz; G_ASSUME_SHARED(p); return p; }
template <int M> GDEV double* sh_vp(const IpmCtx<M>& c) { double* p = c.vp; G_ASSUME_SHARED(p); return p; }
template <int M> GDEV int* sh_seg(const IpmCtx<M>& c) { return c.seg; }   // (an address-space assumption on this one miscompiles with nvcc 12.9)

// ------------------------------------------------------------------------------------------- slot algebra
// One inequality  c0(z) [- t] + s = 0, s >= 0 (multiplier lam) [, t >= 0 (multiplier lamb), cost omega*t].
// After eliminating (s, lam [, t, lamb]) the row contributes  kap * gv gv' + lam * hess  to H and  -gv * bt  to the rhs.
struct Pair { double rc, rt, wa, wb, ba, bb, iw, kap, bt, la, isa, itv; };

// A slot record is addressed as st[f * ss], f = 0 .. SLOT_W-1: the records of the special rows and of the compacted obstacle
// rows are stored field-major ([field][knot] / [field][row]) so that a pass with one thread per knot / per row reads and writes
// them coalesced; the goal-box records (a handful) are contiguous (ss = 1).
GDEV void pair_eval(const double* st, size_t ss, bool has_t, double c0, double omega, double smu, int phase, Pair& q) {
  const double sa = st[0], la = st[ss];
  const double isa = g_rcp(sa);
  q.la = la;
  q.wa = la * isa;
  q.isa = isa; q.itv = 0.0;
  if (has_t) {
    const double t = st[2 * ss], lb = st[3 * ss];
    const double it = g_rcp(t);
    q.itv = it;
    q.rc = c0 - t + sa; q.rt = omega - la - lb;
    q.wb = lb * it;
    const double rsa = sa * la - smu + (phase ? st[4 * ss] : 0.0), rsb = t * lb - smu + (phase ? st[5 * ss] : 0.0);
    q.ba = (la * q.rc - rsa) * isa; q.bb = -rsb * it;
    q.iw = g_rcp(q.wa + q.wb);
    q.kap = q.wa * q.wb * q.iw;
    q.bt = q.ba - q.wa * (q.ba + q.bb - q.rt) * q.iw;
  } else {
    q.rc = c0 + sa; q.rt = 0.0; q.wb = 0.0; q.bb = 0.0; q.iw = 0.0;
    const double rs = sa * la - smu + (phase ? st[4 * ss] : 0.0);
    q.ba = (la * q.rc - rs) * isa;
    q.kap = q.wa;
    q.bt = q.ba;
  }
}

struct Stat { double rz, rc, mus, np; };   // running max |dual residual|, max |row residual|, sum s*lam, #pairs
GDEV void pair_stat(const double* st, size_t ss, bool has_t, const Pair& q, Stat& S) {
  S.rc = fabs(q.rc) > S.rc ? fabs(q.rc) : S.rc;
  if (!(q.rc == q.rc)) S.rc = 1e300;
  S.mus += st[0] * st[ss]; S.np += 1.0;
  if (has_t) { S.rz = fabs(q.rt) > S.rz ? fabs(q.rt) : S.rz; S.mus += st[2 * ss] * st[3 * ss]; S.np += 1.0; }
}

// Step of one row given gdz = gv . dz.
//   mode 0: largest steps to the boundary;
//   mode 1: same + store the predictor products ds*dlam and the coefficients of  sum (s + a ds)(lam + a dlam) = c0 + a c1 + a^2 c2
//           (Mehrotra's mu_aff for any step a, so the predictor needs ONE pass over the rows);
//   mode 2: apply (ap, ad) and accumulate the new complementarity sum / pair count (for the central-path floor).
struct StepAcc { double amp, amd, c0, c1, c2, np; };
// quad = dz' (hess c / 2) dz of a quadratic row: the slack is updated linearly, s += ap ds, while the row moves by ap gdz + ap^2 quad,
// so after a FULL step the row residual c + s [- t] is exactly ap^2 quad -- on the headline batch that second-order remainder
// (1e-4 on the inactive trust-region rows) is what the last one or two Newton iterations of every solve were spent on, with the
// dual residual and the complementarity already converged.  Mode 2 therefore moves the slack to where the row puts it (slack
// reset, as nonlinear interior-point codes do), unless that would take more than 90 % of it (a nearly active row keeps the
// linear update and its fraction-to-boundary guarantee).  -DGUSTO_NO_SLACK_RESET restores the linear update.
GDEV double slack_reset(double s1, double ap, double quad) {
#ifndef GUSTO_NO_SLACK_RESET
  const double s1q = s1 - ap * ap * quad;
  return (quad > 0.0 && s1q >= 0.1 * s1) ? s1q : s1;
#else
  (void)ap; (void)quad;
  return s1;
#endif
}
GDEV void pair_step(double* st, size_t ss, bool has_t, const Pair& q, double gdz, int mode, double ap, double ad, StepAcc& a, double quad = 0.0) {
  const double sa = st[0], la = st[ss];
  if (has_t) {
    const double t = st[2 * ss], lb = st[3 * ss];
    const double dt = (q.wa * gdz + q.ba + q.bb - q.rt) * q.iw;
    const double dla = q.wa * (gdz - dt) + q.ba, dlb = -q.wb * dt + q.bb, ds = -q.rc - (gdz - dt);
    if (mode == 2) {
      const double s0 = sa + ap * ds, s1 = slack_reset(s0, ap, quad), t1 = t + ap * dt, b1 = lb + ad * dlb;
      const double l1 = (la + ad * dla) * (s1 != s0 ? s0 * g_rcp(s1) : 1.0);      // the pair keeps its product: the iterate stays as centred as it was
      st[0] = s1; st[2 * ss] = t1; st[ss] = l1; st[3 * ss] = b1;
      a.c0 += s1 * l1 + t1 * b1; a.np += 2.0;
    } else {
      if (ds < 0) { const double v = -sa * g_rcp(ds); a.amp = v < a.amp ? v : a.amp; }
      if (dt < 0) { const double v = -t * g_rcp(dt); a.amp = v < a.amp ? v : a.amp; }
      if (dla < 0) { const double v = -la * g_rcp(dla); a.amd = v < a.amd ? v : a.amd; }
      if (dlb < 0) { const double v = -lb * g_rcp(dlb); a.amd = v < a.amd ? v : a.amd; }
      if (mode == 1) {
        st[4 * ss] = ds * dla; st[5 * ss] = dt * dlb;
        a.c0 += sa * la + t * lb; a.c1 += sa * dla + la * ds + t * dlb + lb * dt; a.c2 += ds * dla + dt * dlb;
      }
    }
  } else {
    const double dla = q.wa * gdz + q.ba, ds = -q.rc - gdz;
    if (mode == 2) {
      const double s0 = sa + ap * ds, s1 = slack_reset(s0, ap, quad), l1 = (la + ad * dla) * (s1 != s0 ? s0 * g_rcp(s1) : 1.0);
      st[0] = s1; st[ss] = l1;
      a.c0 += s1 * l1; a.np += 1.0;
    } else {
      if (ds < 0) { const double v = -sa * g_rcp(ds); a.amp = v < a.amp ? v : a.amp; }
      if (dla < 0) { const double v = -la * g_rcp(dla); a.amd = v < a.amd ? v : a.amd; }
      if (mode == 1) { st[4 * ss] = ds * dla; st[5 * ss] = 0.0; a.c0 += sa * la; a.c1 += sa * dla + la * ds; a.c2 += ds * dla; }
    }
  }
}
// Keep a complementarity pair above floor = 1e-4 * mu (wide neighbourhood of the central path, as the oracle does).
// Applied lazily by the first reader of the row in the next Newton iteration (assemble, phase 0).
GDEV void pair_floor(double* st, size_t ss, bool has_t, double floor_) {
  if (st[0] * st[ss] < floor_) st[ss] = floor_ * g_rcp(st[0]);
  if (has_t && st[2 * ss] * st[3 * ss] < floor_) st[3 * ss] = floor_ * g_rcp(st[2 * ss]);
}

GDEV void slot_init(double* st, size_t ss, bool valid, bool has_t, double c0, double omega, double t_in = 1.0, double lam_split = 0.5) {
  for (int i = 0; i < SLOT_W; ++i) st[i * ss] = 0.0;
  if (!valid) return;
  if (t_in < 0.0) {
    // Centred start (round 2, last session): both pairs of the row ON the central path at mu0 = -t_in with zero dual residual of t,
    //   lam = mu0 / s,  lamb = mu0 / t,  lam + lamb = omega,  s = t - c0   =>   omega t^2 + (omega a - 2 mu0) t - mu0 a = 0,  a = -c0.
    // The tuned start below leaves e.g. the (far inactive) trust-region pairs at s lam = 0.2 and the solve needs two to three Newton
    // iterations just to bring mu down; from a centred point at mu0 = 5e-5 omega a first-iteration solve of the headline batch takes 4
    // Newton iterations instead of 5.7, later SCP iterations 3 instead of 4.  It needs a start point that violates no soft row (a
    // violated hinge row has t >= c0 and wants mu ~ omega c0): setup() measures that and falls back to the tuned start otherwise
    // (hard tier, first iterations: 30-50 Newton iterations from a centred point at 1e-4, 7-10 from the tuned one).
    const double mu0 = -t_in;
    if (has_t) {
      const double a = -c0, bq = omega * a - 2.0 * mu0;
      double t = (-bq + sqrt(bq * bq + 4.0 * omega * mu0 * a)) / (2.0 * omega);
      if (!(t > 0.0) || !(t + a > 0.0)) t = (c0 > 0 ? c0 : 0.0) + mu0 / omega + 1e-12;
      const double sa = t + a;
      st[0] = sa; st[ss] = mu0 / sa; st[2 * ss] = t; st[3 * ss] = mu0 / t;
    } else {
      const double sa = -c0 > 1e-6 ? -c0 : 1e-2;
      st[0] = sa; st[ss] = mu0 / sa;
    }
    return;
  }
  if (has_t) {
    // interior start of the penalty slack: t_in inside (the oracle uses one unit).  0.25 saves 0.85 Newton iterations of 8.85 on
    // astrobeeSE3 but costs restarts on freeflyerSE2 at omega = 25 (measured), so it is a per-model setting
    // (slack_start<M>()), like the refinement count; the optimum reached is the same to the solver tolerance.
    const double t = (c0 > 0 ? c0 : 0.0) + t_in;
    const double sa = t - c0;
    st[0] = sa > 1e-2 ? sa : 1e-2; st[ss] = lam_split * omega; st[2 * ss] = t; st[3 * ss] = omega - lam_split * omega;
  } else {
    // hard rows: the oracle's s = max(-c, 1e-2), except that a strictly feasible row keeps its exact slack -- a BoxGoal of
    // width 2e-4 (astrobeeSE3manifold notebook) would otherwise start 100x outside its own width on both sides
    st[0] = -c0 > 1e-2 ? -c0 : (-c0 > 1e-6 ? -c0 : 1e-2); st[ss] = 1e-2;
  }
}

// ------------------------------------------------------------------------- TrajOpt: l1-penalised dynamics rows
// Entry i of row j:  d(z) - p + n = 0,  p, n >= 0 (multipliers lp, ln),  cost mu (p + n);  nu is the row's multiplier.
//   stationarity   rp = mu - nu - lp,  rn = mu + nu - ln;      complementarity  p lp = n ln = smu
//   Newton         dlp = -dnu + rp,  dp = (-(p lp - smu + pa) - p dlp) / lp;   dln = dnu + rn,  dn = (-(n ln - smu + pb) - n dln) / ln
//   so             dp - dn = D dnu + beta,   D = p/lp + n/ln,   and the row reads   Aeq dz - D dnu = -(d - p + n) + beta.
// st addresses the record with field stride fs (IpmLayout::DSLOT_W fields).
struct L1Pair { double p, lp, n, ln, rp, rn, ilp, iln; };
GDEV void l1_eval(const double* st, size_t fs, double mu, double nu, L1Pair& q) {
  q.p = st[0]; q.lp = st[fs]; q.n = st[2 * fs]; q.ln = st[3 * fs];
  q.rp = mu - nu - q.lp; q.rn = mu + nu - q.ln;
  q.ilp = g_rcp(q.lp); q.iln = g_rcp(q.ln);
}
// diff = (d - p + n) + (Aeq dz)_row is what dp - dn must equal for the linearised row to hold.  Of the two slacks, the one with the
// smaller p / lam ratio is taken from its complementarity equation (exact in dnu); the other follows from the row: along a relaxed
// defect direction p / lam_p reaches 1e9 and would turn the rounding error of dnu into an O(1e-2) error of dp.
GDEV void l1_dir(const double* st, size_t fs, const L1Pair& q, double dnu, double diff, double smu, int phase, double& dp, double& dlp, double& dn, double& dln) {
  dlp = -dnu + q.rp; dln = dnu + q.rn;
  if (q.p * q.ilp <= q.n * q.iln) {
    dp = (-(q.p * q.lp - smu + (phase ? st[4 * fs] : 0.0)) - q.p * dlp) * q.ilp;
    dn = dp - diff;
  } else {
    dn = (-(q.n * q.ln - smu + (phase ? st[5 * fs] : 0.0)) - q.n * dln) * q.iln;
    dp = dn + diff;
  }
}
// modes as pair_step
GDEV void l1_step(double* st, size_t fs, const L1Pair& q, double dnu, double diff, double smu, int phase, int mode, double ap, double ad, StepAcc& a) {
  double dp, dlp, dn, dln;
  l1_dir(st, fs, q, dnu, diff, smu, phase, dp, dlp, dn, dln);
  if (mode == 2) {
    const double p1 = q.p + ap * dp, n1 = q.n + ap * dn, lp1 = q.lp + ad * dlp, ln1 = q.ln + ad * dln;
    st[0] = p1; st[fs] = lp1; st[2 * fs] = n1; st[3 * fs] = ln1;
    a.c0 += p1 * lp1 + n1 * ln1; a.np += 2.0;
  } else {
    if (dp < 0) { const double v = -q.p * g_rcp(dp); a.amp = v < a.amp ? v : a.amp; }
    if (dn < 0) { const double v = -q.n * g_rcp(dn); a.amp = v < a.amp ? v : a.amp; }
    if (dlp < 0) { const double v = -q.lp * g_rcp(dlp); a.amd = v < a.amd ? v : a.amd; }
    if (dln < 0) { const double v = -q.ln * g_rcp(dln); a.amd = v < a.amd ? v : a.amd; }
    if (mode == 1) {
      st[4 * fs] = dp * dlp; st[5 * fs] = dn * dln;
      a.c0 += q.p * q.lp + q.n * q.ln; a.c1 += q.p * dlp + q.lp * dp + q.n * dln + q.ln * dn; a.c2 += dp * dlp + dn * dln;
    }
  }
}
// start: p - n = d (zero row residual), both t0 inside, multipliers at mu (zero dual residual with nu = 0)
GDEV void l1_init(double* st, size_t fs, double d, double mu) {
  const double t0 = 1e-2;
  st[0] = (d > 0 ? d : 0.0) + t0; st[fs] = mu; st[2 * fs] = (d < 0 ? -d : 0.0) + t0; st[3 * fs] = mu; st[4 * fs] = 0.0; st[5 * fs] = 0.0;
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
    double qp[4];
    for (int i = 0; i < 4; ++i) qp[i] = c.xps[(size_t)(6 + i) * c.NP + k];
    const double nq = sqrt(qp[0] * qp[0] + qp[1] * qp[1] + qp[2] * qp[2] + qp[3] * qp[3]);
    double ev = -1.0;
    for (int i = 0; i < 4; ++i) ev += qp[i] / nq * x[6 + i];
    const bool hinge = (s == L::S_QUAT);
    o.i0 = 6; o.n = 4; o.has_t = hinge;
    for (int i = 0; i < 4; ++i) o.gv[i] = (hinge ? 1.0 : -1.0) * qp[i] / nq;
    o.c0 = (hinge ? ev : -ev) - c.eow;
  } else {
    // control balls cover k = 1..N-1 only (astrobee_se3.jl:370-371, quirk q3)
    o.is_u = true; o.has_t = kTO;                              // TrajOpt penalises them with mu (scp_trajopt.jl:222-233)
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
// (TrajOpt: the hard row |x - xp|^2 - s <= 0, scp_trajopt.jl:165-173; c.dow carries s)
template <int M> GDEV double tr_c0(const IpmCtx<M>& c, int k, const double* x) {
  constexpr int NX = IpmCtx<M>::NX;
  double v = -c.dow;
  for (int i = 0; i < NX; ++i) { const double dxi = x[i] - c.xps[(size_t)i * c.NP + k]; v += dxi * dxi; }
  return v;
}
// csbci_goal_constraints (dynamics.jl:37-42): X[i,N] - ub <= 0 (j even), lb - X[i,N] <= 0 (j odd); hard
template <int M> GDEV double box_c0(const IpmCtx<M>& c, int j, const double* x) {
  const int i = j >> 1;
  return (j & 1) == 0 ? x[i] - c.goal_hi[i] : c.goal_lo[i] - x[i];
}


// ------------------------------------------------------------------------------------- small dense helpers
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
// In-register Cholesky of a packed-lower SPD matrix of order n <= 6; the diagonal of Lc holds 1 / l_jj.
template <int n> GDEV bool chol_packed(const double* H, double* Lc) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < n; ++j) {
    double dj = H[tri(j, j)];
#pragma unroll
    for (int m = 0; m < j; ++m) dj -= Lc[tri(j, m)] * Lc[tri(j, m)];
    if (!(dj > 0.0)) { ok = false; dj = 1e-300; }
    const double il = g_rsqrt(dj);
    Lc[tri(j, j)] = il;
#pragma unroll
    for (int i = j + 1; i < n; ++i) {
      double v = H[tri(i, j)];
#pragma unroll
      for (int m = 0; m < j; ++m) v -= Lc[tri(i, m)] * Lc[tri(j, m)];
      Lc[tri(i, j)] = v * il;
    }
  }
  return ok;
}
// out = Lv in  /  out = Lv' in  for a packed lower-triangular Lv (the stored L^-1 of a Lam_k)
template <int n> GDEV void tri_lower_mv(const double* Lv, const double* in, double* out) {
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
#pragma unroll
    for (int m = 0; m <= i; ++m) s += Lv[tri(i, m)] * in[m];
    out[i] = s;
  }
}
template <int n> GDEV void tri_lower_tmv(const double* Lv, const double* in, double* out) {
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
#pragma unroll
    for (int m = i; m < n; ++m) s += Lv[tri(m, i)] * in[m];
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
    const double* Ak = c.Ac + j;                 // A_j on its pattern, field-major; A_{j-1} one to the left
    const size_t np = c.NP;
    double acc[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) acc[i] = 0.0;
#pragma unroll
    for (int e = 0; e < ANZ; ++e) acc[T::a_row(e)] += Ak[e * np - 1] * vp[T::a_col(e)] + Ak[e * np] * vc[T::a_col(e)];
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
  const double* nk = nu + k;               // row k (knot k as "current"), row k + 1 (knot k as "previous"): field-major [NX][NE]
  const double* Ak = c.Ac + k;
  const size_t np = c.NP, ne = c.NE;
  double wv[NX], acc[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    const double a = nk[i * ne], b2 = nk[i * ne + 1];
    wv[i] = (k > 0 ? a : 0.0) + (k < N - 1 ? b2 : 0.0);
    out[i] = (k == 0 ? a : -a) + (k == N - 1 ? (((c.pmask >> i) & 1) ? b2 : 0.0) : b2);
    acc[i] = 0.0;
  }
#pragma unroll
  for (int e = 0; e < ANZ; ++e) acc[T::a_col(e)] += Ak[e * np] * wv[T::a_row(e)];
#pragma unroll
  for (int i = 0; i < NX; ++i) out[i] += c.hh * acc[i];
#pragma unroll
  for (int a = 0; a < NU; ++a) out[NX + a] = c.hh * c.bv[a] * wv[T::b_row(a)];
}

// out = blockdiag(packed blocks) * in  over the x-blocks / u-blocks
template <int M, int B0 = 0> GDEV void xblocks_mv(const double* P, const double* in, double* out) {
  using T = Traits<M>;
  if constexpr (B0 < T::XB_CNT) {
    sym_mv<T::XB_n(B0)>(P + T::XB_pk(B0), in + T::XB_off(B0), out + T::XB_off(B0));
    xblocks_mv<M, B0 + 1>(P, in, out);
  }
}
template <int M, int B0 = 0> GDEV void ublocks_mv(const double* P, const double* in, double* out) {
  using T = Traits<M>;
  if constexpr (B0 < T::UB_CNT) {
    sym_mv<T::UB_n(B0)>(P + T::UB_pk(B0), in + T::UB_off(B0), out + T::UB_off(B0));
    ublocks_mv<M, B0 + 1>(P, in, out);
  }
}
// out[0:NX) = Hx_k in   (Hx_k = blockdiag(Hb) + w w',  w = sqrt(kap_tr) 2 (x_k - xp_k))
template <int M> GDEV void apply_Hx(const IpmCtx<M>& c, int k, const double* in, double* out) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX;
  double kd[L::KDW];
  for (int i = 0; i < L::KDW; ++i) kd[i] = c.kd[(size_t)i * c.NP + k];
  xblocks_mv<M>(kd + L::KD_HB, in, out);
  if (T::HAS_TR) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) s += kd[L::KD_W + i] * in[i];
#pragma unroll
    for (int i = 0; i < NX; ++i) out[i] += s * kd[L::KD_W + i];
  }
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
  double la_tr, bt_tr, kap_tr;
  Stat st;
};

template <int M, int B0>
GDEV void assemble_xblock(const IpmCtx<M>& c, int k, int phase, double smu, const double* x, double* gz, KnotAcc<M>& ka) {
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
        gtr[i] = 2.0 * (x[off + i] - c.xps[(size_t)(off + i) * c.NP + k]);
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
      const size_t ss = c.NP;
      double* st = c.sslot + (size_t)s * SLOT_W * ss + k;
      Pair q;
      if (phase == 0) pair_floor(st, ss, o.has_t, c.floor_);
      pair_eval(st, ss, o.has_t, o.c0, c.omega, smu, phase, q);
      if (phase == 0) pair_stat(st, ss, o.has_t, q, ka.st);
      block_add<n>(Hb, gz + off, rr, o.i0 - off, o.n, o.gv, o.hq, q, phase);
    }
    // convexified obstacle rows (compacted): off - nhat.r - t <= 0
    if (T::WS > 0 && B0 == T::XB_of(0)) {
      constexpr int WS = T::WS > 0 ? T::WS : 1;
      const int s0 = sh_seg<M>(c)[k], s1 = sh_seg<M>(c)[k + 1];
      const double* __restrict__ orow = c.orow;
      double* __restrict__ ost = c.ost;
      const size_t pp = c.PP;
      for (int p = s0; p < s1; ++p) {
        const double* row = orow + p;
        double* st = ost + p;
        double gv[4] = {0, 0, 0, 0}, hq[4] = {0, 0, 0, 0};
        double v = row[3 * pp];
#pragma unroll
        for (int a = 0; a < WS; ++a) { const double ra = row[a * pp]; gv[a] = -ra; v -= ra * x[a]; }
        Pair q;
        if (phase == 0) pair_floor(st, pp, true, c.floor_);
        pair_eval(st, pp, true, v, c.omega, smu, phase, q);
        if (phase == 0) pair_stat(st, pp, true, q, ka.st);
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
          if (phase == 0) pair_floor(st, 1, false, c.floor_);
          pair_eval(st, 1, false, box_c0<M>(c, j, x), c.omega, smu, phase, q);
          if (phase == 0) pair_stat(st, 1, false, q, ka.st);
          block_add<n>(Hb, gz + off, rr, i, 1, gv, hq, q, phase);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < n; ++i) {
      c.r[(size_t)(off + i) * c.NP + k] = rr[i] - gz[off + i];
      if (phase == 0) { const double a = fabs(gz[off + i]); ka.st.rz = a > ka.st.rz ? a : ka.st.rz; if (!(a == a)) ka.st.rz = 1e300; }
    }
    if (phase == 0) {
      double* kd = c.kd + k;
#pragma unroll
      for (int i = 0; i < npk; ++i) kd[(size_t)(L::KD_HB + T::XB_pk(B0) + i) * c.NP] = Hb[i];
    }
    assemble_xblock<M, B0 + 1>(c, k, phase, smu, x, gz, ka);
  }
}

template <int M, int B0>
GDEV void assemble_ublock(const IpmCtx<M>& c, int k, int phase, double smu, const double* x, double wk, double* gz, KnotAcc<M>& ka) {
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
      const size_t ss = c.NP;
      double* st = c.sslot + (size_t)s * SLOT_W * ss + k;
      Pair q;
      if (phase == 0) pair_floor(st, ss, o.has_t, c.floor_);
      pair_eval(st, ss, o.has_t, o.c0, c.omega, smu, phase, q);
      if (phase == 0) pair_stat(st, ss, o.has_t, q, ka.st);
      block_add<n>(Hb, gz + NX + off, rr, o.i0 - off, o.n, o.gv, o.hq, q, phase);
    }
#pragma unroll
    for (int i = 0; i < n; ++i) {
      c.r[(size_t)(NX + off + i) * c.NP + k] = rr[i] - gz[NX + off + i];
      if (phase == 0) { const double a = fabs(gz[NX + off + i]); ka.st.rz = a > ka.st.rz ? a : ka.st.rz; if (!(a == a)) ka.st.rz = 1e300; }
    }
    if (phase == 0) {
      double* kd = c.kd + k;
#pragma unroll
      for (int i = 0; i < npk; ++i) kd[(size_t)(L::KD_HU + T::UB_pk(B0) + i) * c.NP] = Hb[i];
    }
    assemble_ublock<M, B0 + 1>(c, k, phase, smu, x, wk, gz, ka);
  }
}

// phase 0: build the per-knot Hessian blocks and the predictor rhs (sigma*mu = 0, no second-order term); also the
//          residual norms and the equality residual rnu = b - Aeq z.
// phase 1: corrector rhs with centering target `smu` and the stored predictor products.
template <int M> GDEV_NOINLINE void assemble(const IpmCtx<M>& c, int phase, double smu, Resid* out) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const int N = c.N;
  Stat S; S.rz = 0; S.rc = 0; S.mus = 0; S.np = 0;
  G_PAR_FOR(k, N) {
    const double* x = sh_z<M>(c) + k * NV;
    const double* u = x + NX;
    double gz[NV];
    aeqT_knot<M>(c, c.nu, k, gz);
    const double wk = (k == 0 || k == N - 1) ? 0.5 * c.h : c.h;
#pragma unroll
    for (int i = 0; i < NU; ++i) gz[NX + i] += 2.0 * wk * u[i];
    KnotAcc<M> ka;
    ka.la_tr = 0; ka.bt_tr = 0; ka.kap_tr = 0; ka.st = S;
    if (T::HAS_TR) {
      const size_t ss = c.NP;
      double* st = c.sslot + (size_t)L::S_TR * SLOT_W * ss + k;
      Pair q;
      if (phase == 0) pair_floor(st, ss, !kTO, c.floor_);
      pair_eval(st, ss, !kTO, tr_c0<M>(c, k, x), c.omega, smu, phase, q);
      if (phase == 0) pair_stat(st, ss, !kTO, q, ka.st);
      ka.la_tr = q.la; ka.bt_tr = q.bt; ka.kap_tr = q.kap;
    }
    assemble_xblock<M, 0>(c, k, phase, smu, x, gz, ka);
    assemble_ublock<M, 0>(c, k, phase, smu, x, wk, gz, ka);
    S = ka.st;
    if (phase == 0 && T::HAS_TR) {
      double* kd = c.kd + k;
      const double sk = sqrt(ka.kap_tr);
#pragma unroll
      for (int i = 0; i < NX; ++i) kd[(size_t)(L::KD_W + i) * c.NP] = sk * (2.0 * (x[i] - c.xps[(size_t)i * c.NP + k]));
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
        double t = v[i] + c.gsum[(size_t)i * c.NE + j];           // gsum: -x_init | h/2 (g_{j-1} + g_j) | -goal (setup)
        if (kTO && j >= 1 && j < N) {
          // l1-penalised dynamics row: residual d - p + n, reduced right-hand side rho' = -(d - p + n) + beta (predictor: smu = 0)
          const size_t fs = (size_t)NX * c.NE;
          double* st = c.dslot + (size_t)i * c.NE + j;
          if (st[0] * st[fs] < c.floor_) st[fs] = c.floor_ * g_rcp(st[0]);
          if (st[2 * fs] * st[3 * fs] < c.floor_) st[3 * fs] = c.floor_ * g_rcp(st[2 * fs]);
          L1Pair q;
          l1_eval(st, fs, c.omega, c.nu[(size_t)i * c.NE + j], q);
          const double re = t - q.p + q.n;
          const double beta = -q.p - q.p * q.rp * q.ilp + q.n + q.n * q.rn * q.iln;
          c.dd[(size_t)i * c.NE + j] = q.p * q.ilp + q.n * q.iln;
          const double ad = fabs(q.rp) > fabs(q.rn) ? fabs(q.rp) : fabs(q.rn);
          S.rz = ad > S.rz ? ad : S.rz;
          if (!(ad == ad)) S.rz = 1e300;
          S.mus += q.p * q.lp + q.n * q.ln; S.np += 2.0;
          const double a = fabs(re);
          rpmax = a > rpmax ? a : rpmax;
          if (!(a == a)) rpmax = 1e300;
          t = re - beta;                                          // = -rho'
        } else {
          const double a = fabs(t);
          rpmax = a > rpmax ? a : rpmax;
          if (!(a == a)) rpmax = 1e300;
        }
        v[i] = t;
        c.rnu[(size_t)i * c.NE + j] = -t;
      }
      // ch_{j-1} = -F_j^-1 rho_j  (rho_j = rnu_j = -t): the affine term of the shifted-state recursion, by the thread that holds rho_j
      if (j >= 1 && j < N) {
        const double* fi = c.fi + j;
        const size_t np = c.NP;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double s2 = 0.0;
#pragma unroll
          for (int m = 0; m < NX; ++m) if (L::dsame(i, m)) s2 += fi[(size_t)(i * NX + m) * np] * v[m];
          c.ch[(size_t)i * np + j - 1] = s2;
        }
      }
    }
    out->rz = block_max(S.rz, c.red);
    out->rp = block_max(rpmax, c.red);
    out->rc = block_max(S.rc, c.red);
    const double ms = block_sum(S.mus, c.red), np = block_sum(S.np, c.red);
    out->npair = np;
    out->mu = np > 0 ? ms / np : 0.0;
    if (!(out->mu == out->mu)) out->mu = 1e300;
  } else {
    G_SYNC();
  }
}

// --------------------------------------------------------------------------------- per-solve dynamics records
// Fi_k = F_k^-1 = (I - h/2 A_k)^-1  and, per knot, the rows of  Ah_k' | Bh_k' | Gam_k'  (see the file header).  Constant
// over the Newton iterations of a solve: only the Hessian blocks change.  Gauss-Jordan on [F | I] with one thread per
// (knot, row) pair, 64 / NX knots at a time, rows in shared memory (the pivot row is read by the other rows of its knot);
// F is a small perturbation of the identity (h/2 |A_ii| << 1), so there is no pivoting (a breakdown surfaces as NaN ->
// IPM_NUMERICAL).
// (round 2, late) One THREAD per knot: the decoupled blocks of F_k are inverted in registers by an unrolled in-place Gauss-Jordan
// (no pivoting: F is a small perturbation of the identity), so the whole pass has no barrier -- the cooperative version (one
// thread per (knot, row), pivot rows through shared memory, 2 barriers per pivot, 64 / NX knots at a time) took 580 k cycles per
// solve on the headline batch, 6.5 % of the kernel.  -DGUSTO_SETUP_COOP restores it.
template <int M, int LO, int HI> GDEV void setup_dynamics_block(const IpmCtx<M>& c, int k, double* stg) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, ANZ = L::ANZ, LDT = L::LDT, n = HI - LO;
  const int N = c.N;
  const size_t np = c.NP;
  const double hh = c.hh;
  double a[n][n];
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j < n; ++j) a[i][j] = i == j ? 1.0 : 0.0;
#pragma unroll
  for (int e = 0; e < ANZ; ++e)
    if (T::a_row(e) >= LO && T::a_row(e) < HI) a[T::a_row(e) - LO][T::a_col(e) - LO] -= hh * c.Ac[(size_t)e * np + k];
  // in-place Gauss-Jordan inverse
#pragma unroll
  for (int p = 0; p < n; ++p) {
    const double ip = 1.0 / a[p][p];
    a[p][p] = 1.0;
#pragma unroll
    for (int j = 0; j < n; ++j) a[p][j] *= ip;
#pragma unroll
    for (int r = 0; r < n; ++r) {
      if (r == p) continue;
      const double f = a[r][p];
      a[r][p] = 0.0;
#pragma unroll
      for (int j = 0; j < n; ++j) a[r][j] -= f * a[p][j];
    }
  }
  // Fi_k (rows of this block; entries outside the block stay at the zero the scratch was allocated with)
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j < n; ++j) c.fi[(size_t)((LO + i) * NX + LO + j) * np + k] = a[i][j];
  // Gam_k' (column b_row(a) of Fi_k scaled; zero outside the block of the driven coordinate, and for k = 0)
  // (the knot-major record rows go through the shared-memory stage `stg` = [Ah_{k-1}' rows | Gam_k' rows] and are written to
  // global memory by the whole group, coalesced: one thread per knot storing its own record costs 32 sectors per store instruction)
#pragma unroll
  for (int q = 0; q < NU; ++q) {
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double gv = 0.0;
      if (T::b_row(q) >= LO && T::b_row(q) < HI && k > 0) gv = hh * c.bv[q] * a[i][T::b_row(q) - LO];
      stg[(NX + q) * LDT + LO + i] = gv;
      c.gs[(size_t)(q * NX + LO + i) * np + k] = gv;
    }
  }
  // Ah_{k-1}' = (Fi_k E_k)',  E_k = I + h/2 A_{k-1}: column LO + i of every row j of the record of knot k - 1
  if (k >= 1) {
    double ah[n][n];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < n; ++j) ah[i][j] = a[i][j];
#pragma unroll
    for (int e = 0; e < ANZ; ++e)
      if (T::a_row(e) >= LO && T::a_row(e) < HI) {
        const double av = hh * c.Ac[(size_t)e * np + k - 1];
#pragma unroll
        for (int i = 0; i < n; ++i) ah[i][T::a_col(e) - LO] += av * a[i][T::a_row(e) - LO];
      }
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
      for (int i = 0; i < n; ++i) stg[j * LDT + LO + i] = (j >= LO && j < HI) ? ah[i][j - LO] : 0.0;
  }
  (void)N;
}
#ifndef GUSTO_SETUP_COOP
template <int M> GDEV_NOINLINE void setup_dynamics(const IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, LDT = L::LDT;
  static_assert(L::dcheck(), "Traits<M>::DSPLIT does not decouple the pattern of A");
  const int N = c.N;
  constexpr int RW = (NX + NU) * LDT, W = RW | 1;          // staged rows of one knot; odd slot stride: conflict-free
  double* const stage = sh_dz<M>(c);                       // the work region and the costate chain behind it are idle here
  const int avail = L::work_doubles(N) + (int)L::rnd((size_t)(N + 1) * NX);
  constexpr int GW = NU * LDT;                                // Gam' rows of the previous round's last knot, kept behind the slots
  const int KC = (avail - GW) / W > 0 ? (avail - GW) / W : 1;
  double* const prevG = stage + (size_t)KC * W;
  static_assert(L::CR_AT == 0 && L::CR_BT == NX && L::CR_GT == NX + NU, "record layout");
  for (int k0 = 0; k0 < N; k0 += KC) {
    const int nk = N - k0 < KC ? N - k0 : KC;
    G_PAR_FOR(kk, nk) {
      const int k = k0 + kk;
      double* stg = stage + (size_t)kk * W;
      if constexpr (T::DSPLIT > 0) {
        setup_dynamics_block<M, 0, T::DSPLIT>(c, k, stg);
        setup_dynamics_block<M, T::DSPLIT, NX>(c, k, stg);
      } else {
        setup_dynamics_block<M, 0, NX>(c, k, stg);
      }
    }
    G_SYNC();
    // Bh_{k-1} = Ah_{k-1} Gam_{k-1} + Gam_k from the staged rows (Gam_{k-1}: the previous slot, or the field-major copy the previous
    // round wrote); knot fastest, so the field-major store is coalesced.  (As a separate pass over the records in global memory
    // this was 3 % of the kernel's stall samples, all exposed L2 latency.)
    G_PAR_FOR(it, nk * NU * NX) {
      const int e = it / nk, kk = it - e * nk, k = k0 + kk, a = e / NX, i = e - a * NX;
      if (k < 1) continue;
      const double* stg = stage + (size_t)kk * W;
      double v = stg[(NX + a) * LDT + i];
      for (int m = L::dlo(i); m < L::dhi(i); ++m) {
        const double gm = kk >= 1 ? stg[(NX + a) * LDT + m - W] : prevG[a * LDT + m];
        v += stg[m * LDT + i] * gm;
      }
      c.bs[(size_t)e * c.NP + k - 1] = v;
      c.cr[(size_t)(k - 1) * L::CRW + (L::CR_BT + a) * LDT + i] = v;
    }
    G_PAR_FOR(it, nk * RW) {
      const int kk = it / RW, e = it - kk * RW, k = k0 + kk;
      if (e % LDT >= NX) continue;                          // padding columns of the tile rows stay at their allocation-time zero
      const double v = stage[(size_t)kk * W + e];
      if (e < NX * LDT) { if (k >= 1) c.cr[(size_t)(k - 1) * L::CRW + e] = v; }
      else c.cr[(size_t)k * L::CRW + L::CR_GT * LDT + (e - NX * LDT)] = v;
    }
    G_SYNC();
    if (k0 + KC < N) {
      G_PAR_FOR(t, GW) prevG[t] = stage[(size_t)(nk - 1) * W + NX * LDT + t];
      G_SYNC();                                                    // before the next round's compute overwrites that slot
    }
  }
  G_PAR_FOR(e, RW) c.cr[(size_t)(N - 1) * L::CRW + e] = 0.0;     // no dynamics after the last knot: Ah', Bh' rows of knot N - 1
  G_PAR_FOR(r, NU * NX) c.bs[(size_t)r * c.NP + N - 1] = 0.0;
  G_SYNC();
}
#else
template <int M> GDEV_NOINLINE void setup_dynamics(const IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NN = L::NN, ANZ = L::ANZ, LDT = L::LDT, KB = 64 / NX, W2 = 2 * NX;
  static_assert(L::dcheck(), "Traits<M>::DSPLIT does not decouple the pattern of A");
  const int N = c.N;
  const double hh = c.hh;
  double* const rows = sh_dz<M>(c);            // [KB * NX][2 NX]: row i of [F | W] of knot kb
  for (int k0 = 0; k0 < N; k0 += KB) {
    G_PAR_FOR(t, KB * NX) {
      const int kb = t / NX, i = t - kb * NX, k = k0 + kb;
      double* row = rows + t * W2;
      for (int j = 0; j < W2; ++j) row[j] = 0.0;
      row[i] = 1.0; row[NX + i] = 1.0;
      if (k < N) for (int e = 0; e < ANZ; ++e) if (T::a_row(e) == i) row[T::a_col(e)] -= hh * c.Ac[(size_t)e * c.NP + k];
    }
    G_SYNC();
    // the decoupled blocks are eliminated side by side: step t uses pivot dlo + t of every block
    for (int tp = 0; tp < L::DMAX; ++tp) {
      G_PAR_FOR(t, KB * NX) {
        const int kb = t / NX, i = t - kb * NX;
        const int lo = L::dlo(i), hi = L::dhi(i), p = lo + tp;
        if (i != p) continue;
        double* row = rows + t * W2;
        const double ip = 1.0 / row[p];
        for (int j = lo; j < hi; ++j) { row[j] *= ip; row[NX + j] *= ip; }
      }
      G_SYNC();
      G_PAR_FOR(t, KB * NX) {
        const int kb = t / NX, i = t - kb * NX;
        const int lo = L::dlo(i), hi = L::dhi(i), p = lo + tp;
        if (p >= hi || i == p) continue;
        double* row = rows + t * W2;
        const double* piv = rows + (kb * NX + p) * W2;
        const double f = row[p];
        if (f != 0.0) for (int j = lo; j < hi; ++j) { row[j] -= f * piv[j]; row[NX + j] -= f * piv[NX + j]; }
      }
      G_SYNC();
    }
    // row i of Fi_k: the inverse itself, column i of Gam_k' and of Ah_{k-1}' = (Fi_k E_k)',  E_k = I + h/2 A_{k-1}
    G_PAR_FOR(t, KB * NX) {
      const int kb = t / NX, i = t - kb * NX, k = k0 + kb;
      if (k >= N) continue;
      const double* w = rows + t * W2 + NX;
      double* fo = c.fi + (size_t)(i * NX) * c.NP + k;
      for (int j = 0; j < NX; ++j) fo[(size_t)j * c.NP] = w[j];
      double* cr = c.cr + (size_t)k * L::CRW;
      for (int a = 0; a < NU; ++a) {
        const double gv = k == 0 ? 0.0 : hh * c.bv[a] * w[T::b_row(a)];
        cr[(L::CR_GT + a) * LDT + i] = gv;
        c.gs[(size_t)(a * NX + i) * c.NP + k] = gv;
      }
      if (k >= 1) {
        double ah[NX];
        for (int j = 0; j < NX; ++j) ah[j] = w[j];
        for (int e = 0; e < ANZ; ++e) ah[T::a_col(e)] += hh * c.Ac[(size_t)e * c.NP + k - 1] * w[T::a_row(e)];
        double* crp = c.cr + (size_t)(k - 1) * L::CRW;
        for (int j = 0; j < NX; ++j) crp[(L::CR_AT + j) * LDT + i] = ah[j];
      }
      if (k == N - 1) for (int j = 0; j < NX + NU; ++j) cr[j * LDT + i] = 0.0;      // no dynamics after the last knot
    }
    G_SYNC();
  }
  // Bh_k = Ah_k Gam_k + Gam_{k+1}
  G_PAR_FOR(it, (N - 1) * NU * NX) {
    const int k = it / (NU * NX), r = it - k * (NU * NX), a = r / NX, i = r - a * NX;
    const double* cr = c.cr + (size_t)k * L::CRW;
    const double* gk = cr + (L::CR_GT + a) * LDT;
    double v = c.cr[(size_t)(k + 1) * L::CRW + (L::CR_GT + a) * LDT + i];
    for (int m = L::dlo(i); m < L::dhi(i); ++m) v += cr[(L::CR_AT + m) * LDT + i] * gk[m];
    c.cr[(size_t)k * L::CRW + (L::CR_BT + a) * LDT + i] = v;
    c.bs[(size_t)(a * NX + i) * c.NP + k] = v;
  }
  G_PAR_FOR(r, NU * NX) c.bs[(size_t)r * c.NP + N - 1] = 0.0;      // no dynamics after the last knot
  G_SYNC();
}
#endif

// ------------------------------------------------------------------------------------------- Riccati sweep
// Inputs of knot k that change per Newton iteration (Hessian record, right-hand side, ch_k) are gathered into a staging
// buffer with asynchronous 8-byte copies one knot ahead; the dynamics record goes straight into its tile rows.
template <int M> GDEV void ric_stage_async(const IpmCtx<M>& c, int k, double* tile, int buf) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV, LDT = L::LDT;
  double* XS = tile + L::F_XS + buf * L::XSR * LDT;
  const double* cr = c.cr + (size_t)k * L::CRW;
  G_PAR_FOR(t, (NX + NU) * LDT / 2) g_cp_async16(XS + 2 * t, cr + 2 * t);
  G_PAR_FOR(t, NU * LDT / 2) g_cp_async16(XS + L::RXS * LDT + 2 * t, cr + L::CR_GT * LDT + 2 * t);
  double* SB = tile + L::F_SB + buf * L::SBW;
  const size_t np = c.NP;
  G_PAR_FOR(t, L::KDW) g_cp_async8(SB + t, c.kd + t * np + k);
  G_PAR_FOR(t, NV) g_cp_async8(SB + L::KDW + t, c.r + t * np + k);
  if (k < c.N - 1) G_PAR_FOR(t, NX) g_cp_async8(SB + L::KDW + NV + t, c.ch + t * np + k);
}
// Dense stage inputs of knot k from its staging buffer, by the lanes of ONE warp (no barrier): lane i builds row i of Hx_k
// (+ w_N on the PointGoal coordinates of the last knot) and of Hu_k, q = rx (+ w_N rho_N), ru, and ch_k into its row of the
// dynamics tile.
template <int M, int B0 = 0> GDEV void ric_hx_row(const double* Hb, int i, double* rowv) {
  using T = Traits<M>;
  if constexpr (B0 < T::XB_CNT) {
    constexpr int off = T::XB_off(B0), n = T::XB_n(B0);
    if (i >= off && i < off + n) {
#pragma unroll
      for (int q = 0; q < n; ++q) rowv[off + q] = Hb[T::XB_pk(B0) + tri(i - off, q)];
    }
    ric_hx_row<M, B0 + 1>(Hb, i, rowv);
  }
}
template <int M, int B0 = 0> GDEV void ric_hu_row(const double* Hb, int a, double* rowv) {
  using T = Traits<M>;
  if constexpr (B0 < T::UB_CNT) {
    constexpr int off = T::UB_off(B0), n = T::UB_n(B0);
    if (a >= off && a < off + n) {
#pragma unroll
      for (int q = 0; q < n; ++q) rowv[off + q] = Hb[T::UB_pk(B0) + tri(a - off, q)];
    }
    ric_hu_row<M, B0 + 1>(Hb, a, rowv);
  }
}
template <int M> GDEV void ric_stage_build(const IpmCtx<M>& c, int k, double* tile, int buf) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV, LDT = L::LDT, LDU = L::LDU;
  const int N = c.N;
  double* Qd = tile + L::F_QD;
  double* Hud = tile + L::F_HU;
  double* vec = tile + L::F_VEC;
  double* XS = tile + L::F_XS + buf * L::XSR * LDT;
  const double* SB = tile + L::F_SB + buf * L::SBW;
  const bool last = (k == N - 1);
  G_LANE_FOR(i, NX) {
    double rowv[NX];
#pragma unroll
    for (int q = 0; q < NX; ++q) rowv[q] = 0.0;
    ric_hx_row<M>(SB + L::KD_HB, i, rowv);
    const bool pin = last && ((c.pmask >> i) & 1);
    const double wi = T::HAS_TR ? SB[L::KD_W + i] : 0.0;
#pragma unroll
    for (int q = 0; q < NX; ++q) {
      double v = rowv[q];
      if (T::HAS_TR) v += wi * SB[L::KD_W + q];
      if (pin && q == i) v += c.wN;
      Qd[i * LDT + q] = v;
    }
    double qv = SB[L::KDW + i];
    if (pin) qv += c.wN * c.rnu[(size_t)i * c.NE + N];
    vec[L::V_Q + buf * L::KP + i] = qv;
    XS[(NX + NU) * LDT + i] = last ? 0.0 : SB[L::KDW + NV + i];
  }
  G_LANE_FOR(a, NU) {
    double rowu[NU];
#pragma unroll
    for (int q = 0; q < NU; ++q) rowu[q] = 0.0;
    ric_hu_row<M>(SB + L::KD_HU, a, rowu);
#pragma unroll
    for (int q = 0; q < NU; ++q) Hud[a * LDU + q] = rowu[q];
    vec[L::V_RU + buf * L::KU + a] = SB[L::KDW + NX + a];
  }
}

// Backward sweep k = N-1 .. 0 (see the file header).  Every dense product is a grid of 8x8x4 FP64 tensor-core tiles
// (g_tile_grid, common.cuh) issued by one warp; the two warps of the group split the products of a phase.  Five phases per knot:
//   A   Z1 = [Ah'; Bh'; ch'] P   (= (P Ah)', (P Bh)', (P ch)'),   S' = Hx Gam
//   B   Lam = Hu + Gam'S' + Bh'(P Bh),  rt = ru + Gam'q + Bh'pit,  pit = p_{k+1} - P ch  |  M' = S' + (P Ah)'Bh
//   C   warp 0: Lam = L L' in registers, Li = L^-1  |  warp 1: W = Hx + Ah'(P Ah), dense inputs of knot k-1
//   C2  Y' = [M'; rt'] Li'  (= (L^-1 [M | rt])')  |  BL = Bh Li'
//   D   P_k = W - Y'Y (in place),  K' = Y' Li (row NX: kap)  |  Acl = Ah - BL Y,  p_k = q + Ah'pit - Y' yr
// Leaves, per knot, Acl_k, K_k', L_k^-1, P_k, psi_k = P_{k+1} ch_k in global scratch, and for the right-hand side held in
// c.r / c.rnu (the predictor's): kap_k in c.kap and the costate chain in vp:  vp[0] = p_0,  vp[k+1] = pit_k.  Returns false
// on a non-positive pivot of a Lam_k.
template <int M> GDEV_NOINLINE bool riccati_factor(const IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NU = L::NU, LDT = L::LDT, LDU = L::LDU, NTU = L::NTU, KS = L::KP / 4, KSU = L::KU / 4;
  constexpr int MT_X = L::NXP / 8, MT_Z = L::RXS / 8, MZ0 = (MT_Z + 1) / 2, MZ1 = MT_Z - MZ0;
  const int N = c.N;
  double* const tile = sh_dz<M>(c);
  double* const vp = sh_vp<M>(c);
  double* const Z1 = tile + L::F_Z1;
  double* const Qd = tile + L::F_QD;
  double* const Sm = tile + L::F_SM;
  double* const Lm = tile + L::F_LM;
  double* const Li = tile + L::F_LI;
  double* const Hud = tile + L::F_HU;
  double* const Yt = Z1;                                        // (Z1 is dead after phase B)
  double* const vec = tile + L::F_VEC;
  double* const pit = vec + L::V_PIT;
  const bool w0 = G_NWARP == 1 || G_WARPID == 0, w1 = G_NWARP == 1 || G_WARPID == 1;
  double bad = 0.0;
  const long long tc0 = g_clock();
  G_PAR_FOR(it, L::FAC_DOUBLES) tile[it] = 0.0;
  G_PAR_FOR(i, NX) vp[N * NX + i] = 0.0;
  G_SYNC();
  ric_stage_async<M>(c, N - 1, tile, (N - 1) & 1);
  g_cp_async_wait();
  G_SYNC();
  if (w0) ric_stage_build<M>(c, N - 1, tile, (N - 1) & 1);
  G_SYNC();
#if defined(GUSTO_PROF_MODE) && GUSTO_PROF_MODE == 3
  long long tp = g_clock();
#define GUSTO_PROF_TICK(slot) do { if (G_TID == 0) { const long long tn = g_clock(); c.prof[slot] += tn - tp; tp = tn; } } while (0)
#else
#define GUSTO_PROF_TICK(slot) ((void)0)
#endif
  // TrajOpt: F_j^-1 (its non-zero blocks) and D_j of the NEXT knot step's noise phase, issued once their shared-memory regions
  // are dead at the end of a step: the round trip to global scratch hides behind the chain-tile product and the barriers between
  auto noise_prefetch = [&](int j) {
    double* const Vt = tile + L::F_NV;
    double* const tv = tile + L::F_NVEC;
    const size_t np = c.NP, ne = c.NE;
    G_PAR_FOR(it, NX * NX + NX) {
      if (it < NX * NX) {
        const int i = it / NX, m = it - i * NX;
        if (L::dsame(i, m)) g_cp_async8(Vt + i * L::LDN + m, c.fi + (size_t)(i * NX + m) * np + j);
        else Vt[i * L::LDN + m] = 0.0;
      } else g_cp_async8(tv + (it - NX * NX), c.dd + (size_t)(it - NX * NX) * ne + j);
    }
  };
  for (int k = N - 1; k >= 0; --k) {
    const int cur = k & 1, nxt = cur ^ 1;
    const double* XS = tile + L::F_XS + cur * L::XSR * LDT;        // rows: Ah' (NX) | Bh' (NU) | ch | 0.. | Gam' at row RXS
    const double* Gt = XS + L::RXS * LDT;
    const double* Bt = XS + NX * LDT;
    double* const Pp = tile + (((N - 1 - k) & 1) ? L::F_PB : L::F_PA);   // P_{k+1}
    double* const Wn = tile + (((N - 1 - k) & 1) ? L::F_PA : L::F_PB);   // W, then P_k
    double* const Mt = Pp;                                                // M' (row NX: rt): P_{k+1} is dead after phase A
    double* const BL = Sm;                                                // S' is dead after phase B
    const double* qv = vec + L::V_Q + cur * L::KP;
    const double* ruv = vec + L::V_RU + cur * L::KU;
    const double* pn = vp + (k + 1) * NX;
    if (kTO && k < N - 1) {
      // ---- TrajOpt noise phase (file header): the l1-penalised dynamics row j = k + 1 lets the realised state deviate from the
      //      prediction at a quadratic price; minimising over that deviation first turns (P_j, p_j) into (P~_j, p~_j).
      constexpr int LDN = L::LDN;
      const int j = k + 1;
      const size_t np = c.NP, ne = c.NE;
      double* const Vt = tile + L::F_NV;
      double* const Zt = tile + L::F_XS + nxt * L::XSR * LDT;
      double* const Aug = tile + L::F_NT;                       // [NX][2 NX]
      double* const Nt = Wn;
      double* const tv = tile + L::F_NVEC;
      double* const ptil = tv + L::KP;
      double* const wv = ptil + L::KP;
      // Fi_j and D_j are on their way to shared memory since the end of the previous knot step (noise_prefetch); W = Fi D Fi'
      g_cp_async_wait();
      G_SYNC();
      G_PAR_FOR(it, NX * NX) {
        const int i = it / NX, m = it - i * NX;
        double a = 0.0;
        for (int l = 0; l < NX; ++l) a += Vt[i * LDN + l] * tv[l] * Vt[m * LDN + l];
        Zt[i * LDN + m] = a;                                    // W
      }
      G_SYNC();
      // [I + W P | I]  and  w_j = p_j - P_j ch_k  (the pre-noise costate the chains carry)
      constexpr int LD2 = 2 * NX;
      G_PAR_FOR(it, NX * NX + NX) {
        if (it < NX * NX) {
          const int i = it / NX, q = it - i * NX;
          double a = i == q ? 1.0 : 0.0;
          for (int m = 0; m < NX; ++m) a += Zt[i * LDN + m] * Pp[m * LDT + q];
          Aug[i * LD2 + q] = a;
          Aug[i * LD2 + NX + q] = i == q ? 1.0 : 0.0;
        } else {
          const int i = it - NX * NX;
          double a = pn[i];
          for (int q = 0; q < NX; ++q) a -= Pp[i * LDT + q] * XS[(NX + NU) * LDT + q];
          wv[i] = a;
        }
      }
      G_SYNC();
      // N = (I + W P)^-1 by Gauss-Jordan with row pivoting.  (The textbook forms  N = I - V T^-1 V'P,  P~ = P - P V T^-1 V'P
      // cancel catastrophically once a defect row is relaxed -- D = p / lam_p ~ 1e6 -- and P~ ~ W^-1 << P.)  No row is moved:
      // every thread repeats the pivot search of a column among the rows not used yet (broadcast reads, the same result in every
      // thread, kept as a nibble list in a register), and the whole group updates the live (row, column) pairs of the step --
      // columns left of the pivot column are dead and the pivot column is never read again -- so a column costs ONE barrier.
      static_assert(NX <= 16, "pivot rows are kept as nibbles of one 64-bit register");
      unsigned long long perm = 0;
      unsigned used = 0;
      for (int cc = 0; cc < NX; ++cc) {
        int pr = -1;
        double best = -1.0;
#pragma unroll
        for (int r = 0; r < NX; ++r) {
          const double v = fabs(Aug[r * LD2 + cc]);
          if (!((used >> r) & 1u) && v > best) { best = v; pr = r; }
        }
        if (!(best > 1e-300)) bad = 1.0;
        used |= 1u << pr;
        perm |= (unsigned long long)pr << (4 * cc);
        const double rp = g_rcp(Aug[pr * LD2 + cc]);
        constexpr int RG = GUSTO_IPM_GROUP / LD2 > 0 ? GUSTO_IPM_GROUP / LD2 : 1;   // a thread owns one column and every RG-th row
        G_PAR_FOR(t, RG * LD2) {
          const int g = t / LD2, q = t - g * LD2;
          if (q > cc) {                                         // columns cc + 1 .. 2 NX - 1 are live
            const double pq = Aug[pr * LD2 + q];
            for (int r = g; r < NX; r += RG)
              if (r != pr) Aug[r * LD2 + q] -= Aug[r * LD2 + cc] * rp * pq;
          }
        }
        G_SYNC();
      }
      G_PAR_FOR(it, NX * NX) {
        const int i = it / NX, q = it - i * NX;
        const int pr = (int)((perm >> (4 * i)) & 15u);
        const double a = Aug[pr * LD2 + NX + q] * g_rcp(Aug[pr * LD2 + i]);
        Nt[i * LDN + q] = a;
        c.nm[(size_t)it * np + j] = a;
      }
      G_SYNC();
      G_PAR_FOR(it, NX * NX + NX) {
        if (it < NX * NX) {                                     // P~ = P N, symmetrised (a product: no cancellation)
          const int i = it / NX, q = it - i * NX;
          double a = 0.0;
          for (int m = 0; m < NX; ++m) a += Pp[i * LDT + m] * Nt[m * LDN + q] + Pp[q * LDT + m] * Nt[m * LDN + i];
          Zt[i * LDN + q] = 0.5 * a;
        } else {                                                // p~ = N'p
          const int i = it - NX * NX;
          double a = 0.0;
          for (int m = 0; m < NX; ++m) a += Nt[m * LDN + i] * pn[m];
          ptil[i] = a;
        }
      }
      G_SYNC();
      G_PAR_FOR(it, NX * NX) {
        const int i = it / NX, q = it - i * NX;
        const double a = Zt[i * LDN + q];
        Pp[i * LDT + q] = a;
        if (q <= i) c.pkt[(size_t)tri(i, q) * np + j] = a;
      }
      G_SYNC();
      pn = ptil;
    }
    // ---- phase A
    if (k >= 1) ric_stage_async<M>(c, k - 1, tile, nxt);
    auto z1_store = [&](int r, int q, double v) { if (r <= NX + NU && q < NX) Z1[r * LDT + q] = v; };
    if (w0) g_tile_grid<KS, MZ0, MT_X, false, false>(XS, LDT, 0, Pp, LDT, 0, z1_store);
    if (w1) {
      if constexpr (MZ1 > 0) g_tile_grid<KS, (MZ1 > 0 ? MZ1 : 1), MT_X, false, false>(XS, LDT, 8 * MZ0, Pp, LDT, 0, z1_store);
      g_tile_grid<KS, MT_X, 1, false, false>(Qd, LDT, 0, Gt, LDT, 0, [&](int r, int a, double v) {
        if (r < NX && a < NU) Sm[r * LDU + a] = v;
      });
    }
    G_SYNC();
    GUSTO_PROF_TICK(0);
    // ---- phase B
    if (w0) {
      g_tile_job2<KS, false, true, KS, false, false>(Gt, LDT, Sm, LDU, Bt, LDT, Z1 + NX * LDT, LDT, 0, 0, [&](int a, int b2, double v) {
        if (a < NU && b2 < NU) Lm[a * LDU + b2] = v + Hud[a * LDU + b2];
      });
      G_LANE_FOR(i, NX + NU) {
        if (i < NX) {
          const double ps = Z1[(NX + NU) * LDT + i];             // psi_k = P_{k+1} ch_k  (Z1 is recycled after phase C)
          pit[i] = pn[i] - ps;
          c.psi[(size_t)i * c.NP + k] = ps;
        } else {
          const int a = i - NX;
          double v = ruv[a];
          for (int m = 0; m < NX; ++m) v += Gt[a * LDT + m] * qv[m] + Bt[a * LDT + m] * (pn[m] - Z1[(NX + NU) * LDT + m]);
          Mt[NX * LDU + a] = v;
        }
      }
    }
    if (w1) {
      g_tile_grid<KS, MT_X, 1, false, false>(Z1, LDT, 0, Bt, LDT, 0, [&](int r, int a, double v) {
        if (r < NX && a < L::KU) Mt[r * LDU + a] = a < NU ? v + Sm[r * LDU + a] : 0.0;
      });
    }
    g_cp_async_wait();
    G_SYNC();
    GUSTO_PROF_TICK(1);
    // ---- phase C
    if (w0) {
      double Hl[NTU], Lc[NTU], Lv[NTU];
#pragma unroll
      for (int a = 0; a < NU; ++a)
#pragma unroll
        for (int b2 = 0; b2 <= a; ++b2) Hl[tri(a, b2)] = Lm[a * LDU + b2];
      if (!chol_packed<NU>(Hl, Lc)) bad = 1.0;
#pragma unroll
      for (int j = 0; j < NU; ++j) {
        Lv[tri(j, j)] = Lc[tri(j, j)];
#pragma unroll
        for (int i = j + 1; i < NU; ++i) {
          double sacc = 0.0;
#pragma unroll
          for (int m = j; m < i; ++m) sacc += Lc[tri(i, m)] * Lv[tri(m, j)];
          Lv[tri(i, j)] = -sacc * Lc[tri(i, i)];
        }
      }
      G_LANE_FOR(a, NU) {
#pragma unroll
        for (int b2 = 0; b2 < NU; ++b2) {
          double v = 0.0;
#pragma unroll
          for (int a2 = 0; a2 < NU; ++a2) if (a2 == a && b2 <= a2) v = Lv[tri(a2, b2)];
          Li[a * LDU + b2] = v;
        }
      }
      if (G_LANE == 0) {
        double* lp = c.lp + k;
#pragma unroll
        for (int t = 0; t < NTU; ++t) lp[(size_t)t * c.NP] = Lv[t];
      }
    }
    if (w1) {
      // W does not depend on Lam: it runs beside the (serial) Cholesky of warp 0, then the dense inputs of knot k - 1
      g_tile_grid<KS, MT_X, MT_X, false, false>(XS, LDT, 0, Z1, LDT, 0, [&](int r, int q, double v) {
        if (r < NX && q < NX) Wn[r * LDT + q] = v + Qd[r * LDT + q];
      });
      G_SYNCWARP();
      if (k >= 1) ric_stage_build<M>(c, k - 1, tile, nxt);
    }
    G_SYNC();
    GUSTO_PROF_TICK(2);
    // ---- phase C2
    if (w0) g_tile_grid<KSU, MT_X, 1, false, false>(Mt, LDU, 0, Li, LDU, 0, [&](int r, int a, double v) {
      if (r <= NX && a < L::KU) Yt[r * LDU + a] = v;            // columns NU .. KU-1 are exact zeros (rows of Li beyond NU are)
    });
    if (w1) g_tile_grid<KSU, MT_X, 1, true, false>(Bt, LDT, 0, Li, LDU, 0, [&](int r, int a, double v) {
      if (r < NX && a < NU) BL[r * LDU + a] = v;
    });
    G_SYNC();
    // ---- phase D
    if (w0) {
      g_tile_grid<KSU, MT_X, MT_X, false, false>(Yt, LDU, 0, Yt, LDU, 0, [&](int r, int q, double v) {
        if (r < NX && q < NX) {
          const double pv = Wn[r * LDT + q] - v;
          Wn[r * LDT + q] = pv;
          if (q <= r) c.pk[(size_t)tri(r, q) * c.NP + k] = pv;
        }
      });
      g_tile_grid<KSU, MT_X, 1, false, true>(Yt, LDU, 0, Li, LDU, 0, [&](int r, int a, double v) {
        if (a < NU) {
          if (r < NX) c.kt[(size_t)(r * NU + a) * c.NP + k] = v;
          else if (r == NX) c.kap[(size_t)a * c.NP + k] = v;
        }
      });
    }
    if (w1) {
      g_tile_grid<KSU, MT_X, MT_X, false, false>(BL, LDU, 0, Yt, LDU, 0, [&](int r, int q, double v) {
        if (r < NX && q < NX) {
          if (kTO && k < N - 1) Mt[r * L::LDN + q] = XS[q * LDT + r] - v;                  // multiplied by N_{k+1} below (M' is dead)
          else c.acl[(size_t)k * L::GT + r * LDT + q] = XS[q * LDT + r] - v;
        }
      });
      G_LANE_FOR(i, NX) {
        double v = qv[i];
        for (int m = 0; m < NX; ++m) v += XS[i * LDT + m] * pit[m];
#pragma unroll
        for (int a = 0; a < NU; ++a) v -= Yt[i * LDU + a] * Yt[NX * LDU + a];
        vp[k * NX + i] = v;
        if (kTO) { if (k < N - 1) vp[(k + 1) * NX + i] = tile[L::F_NVEC + 2 * L::KP + i]; }
        else if (k < N - 1) vp[(k + 1) * NX + i] = pit[i];
      }
    }
    G_SYNC();
    if (kTO && k >= 1) noise_prefetch(k);
    if (kTO && k < N - 1) {                                       // chain tile G_k = N_{k+1} Acl_k
      const double* Ng = c.nm + (k + 1);                           // N_{k+1}, written in this step's noise phase
      const double* At = Mt;
      const size_t np = c.NP;
      G_PAR_FOR(it, NX * NX) {
        const int r = it / NX, q = it - r * NX;
        double a = 0.0;
#pragma unroll
        for (int m = 0; m < NX; ++m) a += Ng[(size_t)(r * NX + m) * np] * At[m * L::LDN + q];
        c.acl[(size_t)k * L::GT + r * LDT + q] = a;
      }
      G_SYNC();
    }
    GUSTO_PROF_TICK(3);
  }
  bad = block_max(bad, c.red);
#if !(defined(GUSTO_PROF_MODE) && GUSTO_PROF_MODE == 3)
  if (G_TID == 0) c.prof[0] += g_clock() - tc0;
#endif
  return bad == 0.0;
}

// ------------------------------------------------------------------------------------------------ chains
// One n_x x n_x mat-vec per knot with the closed-loop tiles Acl_k, on warp 0.  The tiles stream from global memory (L2 / HBM:
// with ~220 KB of the SM's shared memory in use L1 is almost gone) into a ring of CHAIN_D stages in the idle part of the work
// region with 16-byte asynchronous copies (LDGSTS, commit groups): plain loads prefetched into registers were measured to
// stall on the scoreboard (~1100 cycles per step, ncu long-scoreboard on the first multiply-add), a TMA ring pays ~90 cycles of
// mbarrier try_wait per step.  The running vector never makes a shared-memory round trip: lane i keeps its entry in a register
// and every lane collects the vector with warp shuffles; the result stores are off the dependent path.
#ifndef GUSTO_CHAIN_D
#define GUSTO_CHAIN_D 4
#endif
template <int M> GDEV void chain_fetch(const IpmCtx<M>& c, double* ring, int slot, int k) {
  using L = IpmLayout<M>;
  const double* src = c.acl + (size_t)k * L::GT;
  double* dst = ring + slot * L::GT;
  G_LANE_FOR(t, L::GT / 2) g_cp_async16(dst + 2 * t, src + 2 * t);
}
// forward:  s_{k+1} = Acl_k s_k + d_k   (s_k in the x-slots of dz; slot k+1 holds d_k on entry, slot 0 holds s_0)
template <int M> GDEV_NOINLINE void chain_forward(const IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NV = L::NV, LDT = L::LDT, HP = (NX + 1) / 2, D = GUSTO_CHAIN_D;
  const int nt = c.N - 1;
  double* const y = c.dz;
  G_ASSUME_SHARED(y);
  const long long tc0 = g_clock();
  if (G_TID < G_WARP) {
#ifdef GUSTO_HOSTSIM
    for (int k = 0; k < nt; ++k)
      for (int i = 0; i < NX; ++i) {
        double a = y[(k + 1) * NV + i];
        for (int m = 0; m < NX; ++m) a += c.acl[(size_t)k * L::GT + i * LDT + m] * y[k * NV + m];
        y[(k + 1) * NV + i] = a;
      }
#else
    double* const ring = y + L::rnd((size_t)c.N * NV);          // idle part of the work region
    const int i = G_LANE;
    const bool act = i < NX;
    const int ir = act ? i : 0;
#pragma unroll
    for (int s2 = 0; s2 < D - 1; ++s2) { if (s2 < nt) chain_fetch<M>(c, ring, s2, s2); g_cp_async_commit(); }
    double sv = y[ir];
#ifdef GUSTO_CHAIN_SKIP
    if (false)
#endif
    for (int k = 0; k < nt; ++k) {
      g_cp_async_wait_group<D - 2>();                            // tile k has landed (this lane's part of it)
      G_SYNCWARP();                                              // ... all lanes' parts; everyone is done with tile k - 1
      if (k + D - 1 < nt) chain_fetch<M>(c, ring, (k + D - 1) % D, k + D - 1);
      g_cp_async_commit();
      const double* R = ring + (k % D) * L::GT + ir * LDT;
      double a0 = act ? y[(k + 1) * NV + i] : 0.0, a1 = 0.0;     // d_k[i]: independent of the chain (idle lanes read nothing: racecheck-clean)
#pragma unroll
      for (int m = 0; m < HP; ++m) {
        const g_d2 rv = g_ld2(R + 2 * m);
        a0 = fma(rv.x, __shfl_sync(0xffffffffu, sv, 2 * m), a0);
        if (2 * m + 1 < NX) a1 = fma(rv.y, __shfl_sync(0xffffffffu, sv, 2 * m + 1), a1);
      }
      sv = a0 + a1;
      if (act) y[(k + 1) * NV + i] = sv;
    }
    g_cp_async_wait();
    G_SYNCWARP();
#endif
  }
  G_SYNC();
  if (G_TID == 0 && GUSTO_PROF_CHAINS) c.prof[1] += g_clock() - tc0;
}
// backward:  vp[k] += Acl_k' vp[k+1]  for k = N-2 .. 0
template <int M> GDEV_NOINLINE void chain_backward(const IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NV = L::NV, LDT = L::LDT, D = GUSTO_CHAIN_D;
  const int nt = c.N - 1;
  double* const y = c.vp;
  G_ASSUME_SHARED(y);
  const long long tc0 = g_clock();
  if (G_TID < G_WARP) {
#ifdef GUSTO_HOSTSIM
    for (int k = nt - 1; k >= 0; --k)
      for (int i = 0; i < NX; ++i) {
        double a = y[k * NX + i];
        for (int m = 0; m < NX; ++m) a += c.acl[(size_t)k * L::GT + m * LDT + i] * y[(k + 1) * NX + m];
        y[k * NX + i] = a;
      }
#else
    double* const ring = c.dz + L::rnd((size_t)c.N * NV);
    G_ASSUME_SHARED(ring);
    const int i = G_LANE;
    const bool act = i < NX;
    const int ir = act ? i : 0;
#pragma unroll
    for (int s2 = 0; s2 < D - 1; ++s2) { if (nt - 1 - s2 >= 0) chain_fetch<M>(c, ring, s2, nt - 1 - s2); g_cp_async_commit(); }
    double sv = y[nt * NX + ir];
#ifdef GUSTO_CHAIN_SKIP
    if (false)
#endif
    for (int k = nt - 1, st = 0; k >= 0; --k, ++st) {
      g_cp_async_wait_group<D - 2>();
      G_SYNCWARP();
      if (k - (D - 1) >= 0) chain_fetch<M>(c, ring, (st + D - 1) % D, k - (D - 1));
      g_cp_async_commit();
      const double* R = ring + (st % D) * L::GT + ir;
      double a0 = act ? y[k * NX + i] : 0.0, a1 = 0.0;
#pragma unroll
      for (int m = 0; m + 1 < NX; m += 2) {
        a0 = fma(R[m * LDT], __shfl_sync(0xffffffffu, sv, m), a0);
        a1 = fma(R[(m + 1) * LDT], __shfl_sync(0xffffffffu, sv, m + 1), a1);
      }
      if (NX & 1) a0 = fma(R[(NX - 1) * LDT], __shfl_sync(0xffffffffu, sv, NX - 1), a0);
      sv = a0 + a1;
      if (act) y[k * NX + i] = sv;
    }
    g_cp_async_wait();
    G_SYNCWARP();
#endif
  }
  G_SYNC();
  if (G_TID == 0 && GUSTO_PROF_CHAINS) c.prof[1] += g_clock() - tc0;
}

// ----------------------------------------------------------------------------------------- per-knot passes
// The s-form control gradient  rs = ru + Gam' rx  of knot k for the right-hand side in c.r / c.rnu.
template <int M> GDEV void ric_rhs_knot(const IpmCtx<M>& c, int k, double* rx, double* rs) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NU = L::NU;
  const size_t np = c.NP;
  const double* r = c.r + k;
#pragma unroll
  for (int i = 0; i < NX; ++i) rx[i] = r[i * np];
  if (k == c.N - 1) {
#pragma unroll
    for (int i = 0; i < NX; ++i) if ((c.pmask >> i) & 1) rx[i] += c.wN * c.rnu[(size_t)i * c.NE + c.N];
  }
  const double* gs = c.gs + k;
#pragma unroll
  for (int a = 0; a < NU; ++a) {
    double v = r[(NX + a) * np];
#pragma unroll
    for (int i = 0; i < NX; ++i) if (L::dsame(i, Traits<M>::b_row(a))) v += gs[(a * NX + i) * np] * rx[i];
    rs[a] = v;
  }
}
// TrajOpt: out = W_j t,  W_j = F_j^-1 D_j F_j^-T  (the "noise covariance" of the l1-penalised dynamics row j)
template <int M> GDEV void noise_W(const IpmCtx<M>& c, int j, const double* t, double* out) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX;
  const size_t np = c.NP, ne = c.NE;
  const double* fi = c.fi + j;
  double y[NX];
#pragma unroll
  for (int m = 0; m < NX; ++m) {
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) if (L::dsame(i, m)) a += fi[(size_t)(i * NX + m) * np] * t[i];
    y[m] = a * c.dd[(size_t)m * ne + j];
  }
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double a = 0.0;
#pragma unroll
    for (int m = 0; m < NX; ++m) if (L::dsame(i, m)) a += fi[(size_t)(i * NX + m) * np] * y[m];
    out[i] = a;
  }
}
// out = Sym(packed, field-major at knot k) * v
template <int M> GDEV void packed_mv(const double* pk_k, size_t np, const double* v, double* out) {
  constexpr int NX = IpmLayout<M>::NX;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < NX; ++q) a += pk_k[(size_t)tri(i, q) * np] * v[q];
    out[i] = a;
  }
}

// Pre-chain part of ric_forward for knot k given kap_k (also called by ric_backward, which has kap_k in registers: one pass, one
// barrier and one global round trip of kap less per corrector solve).
template <int M> GDEV void ric_forward_pre(const IpmCtx<M>& c, int k, const double* kap) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const int N = c.N;
  const size_t np = c.NP, ne = c.NE;
  double* const dz = sh_dz<M>(c);
  const double* const vp = sh_vp<M>(c);
  (void)vp; (void)ne; (void)N;
#pragma unroll
  for (int a = 0; a < NU; ++a) dz[k * NV + NX + a] = kap[a];
  if (k < N - 1) {
    const double* bs = c.bs + k;
    double dk[NX], bk[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v = 0.0;
#pragma unroll
      for (int a = 0; a < NU; ++a) if (L::dsame(i, T::b_row(a))) v += bs[(a * NX + i) * np] * kap[a];
      bk[i] = v;
      dk[i] = v + c.ch[i * np + k];
    }
    if (kTO) {   // realised state: s_{k+1} = N_{k+1} (Acl_k s_k + d_k + W_{k+1} p_{k+1}),  p_{k+1} = w_{k+1} + P_{k+1} ch_k  (products only)
      double pj[NX], wt[NX], chv[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) chv[i] = c.ch[i * np + k];
      packed_mv<M>(c.pk + k + 1, np, chv, pj);
#pragma unroll
      for (int i = 0; i < NX; ++i) pj[i] += vp[(k + 1) * NX + i];
      noise_W<M>(c, k + 1, pj, wt);
#pragma unroll
      for (int i = 0; i < NX; ++i) wt[i] += dk[i];
      const double* nm = c.nm + k + 1;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double a = 0.0;
#pragma unroll
        for (int m = 0; m < NX; ++m) a += nm[(size_t)(i * NX + m) * np] * wt[m];
        dk[i] = a;
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) dz[(k + 1) * NV + i] = dk[i];
  }
  if (k == 0) {
#pragma unroll
    for (int i = 0; i < NX; ++i) dz[i] = c.rnu[i * ne];
  }
}
// Costates for a new right-hand side (the corrector's): bb_k = rx - K'rs - psi_{k-1}, the backward chain, then
// kap_k = Lam_k^-1 (rs + Bh_k' pit_k).
template <int M> GDEV_NOINLINE void ric_backward(const IpmCtx<M>& c, bool fuse_forward_pre = false) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NU = L::NU, NTU = L::NTU;
  const int N = c.N;
  const size_t np = c.NP;
  double* const vp = sh_vp<M>(c);
  long long tc0 = g_clock();
  G_PAR_FOR(k, N) {
    double rx[NX], rs[NU];
    ric_rhs_knot<M>(c, k, rx, rs);
    const double* kt = c.kt + k;
    double psi[NX];
    if (kTO) {     // the corrector changed ch (predictor_pass): psi_{k-1} = P_k ch_{k-1} with the PRE-noise P_k (chain variable w)
#pragma unroll
      for (int i = 0; i < NX; ++i) psi[i] = 0.0;
      if (k >= 1) {
        double chv[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) chv[i] = c.ch[i * np + k - 1];
        packed_mv<M>(c.pk + k, np, chv, psi);
      }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v = rx[i] - (kTO ? psi[i] : (k >= 1 ? c.psi[i * np + k - 1] : 0.0));
#pragma unroll
      for (int a = 0; a < NU; ++a) v -= kt[(i * NU + a) * np] * rs[a];
      vp[k * NX + i] = v;
    }
  }
  G_SYNC();
  if (G_TID == 0 && GUSTO_PROF_CHAINS) c.prof[2] += g_clock() - tc0;
  chain_backward<M>(c);
  tc0 = g_clock();
  G_PAR_FOR(k, N) {
    double rx[NX], rs[NU], Lc[NTU], x[NU];
    ric_rhs_knot<M>(c, k, rx, rs);
    if (k < N - 1) {
      double pit[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) pit[i] = vp[(k + 1) * NX + i];
      if (kTO) {   // pit_k = N_{k+1}' w_{k+1}  (a product: N is tiny along a relaxed defect direction)
        double ww[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) ww[i] = pit[i];
        const double* nm = c.nm + k + 1;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double a = 0.0;
#pragma unroll
          for (int m = 0; m < NX; ++m) a += nm[(size_t)(m * NX + i) * np] * ww[m];
          pit[i] = a;
        }
      }
      const double* bs = c.bs + k;
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        double v = rs[a];
#pragma unroll
        for (int i = 0; i < NX; ++i) if (L::dsame(i, Traits<M>::b_row(a))) v += bs[(a * NX + i) * np] * pit[i];
        rs[a] = v;
      }
    }
    const double* lp = c.lp + k;                           // packed lower L^-1:  kap = L^-T (L^-1 rs)
#pragma unroll
    for (int t = 0; t < NTU; ++t) Lc[t] = lp[t * np];
    tri_lower_mv<NU>(Lc, rs, x);
#pragma unroll
    for (int a = 0; a < NU; ++a) rs[a] = x[a];
    tri_lower_tmv<NU>(Lc, rs, x);
#pragma unroll
    for (int a = 0; a < NU; ++a) if (!fuse_forward_pre) c.kap[a * np + k] = x[a];
    if (fuse_forward_pre) ric_forward_pre<M>(c, k, x);
  }
  G_SYNC();
  if (G_TID == 0 && GUSTO_PROF_CHAINS) c.prof[2] += g_clock() - tc0;
}
// d_k = Bh_k kap_k + ch_k into the x-slot k+1 of dz, s_0 = rho_0 into slot 0, kap_k into the u-slot k; forward chain; then
// per knot  u = kap - K s,  x = s + Gam u  and (want_nu) the equality multipliers  dnu_j = F_j^-T (P_j (s_j - ch_{j-1}) - pit_{j-1}).
template <int M> GDEV_NOINLINE void ric_forward(const IpmCtx<M>& c, bool want_nu, bool pre_done = false) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const int N = c.N;
  const size_t np = c.NP, ne = c.NE;
  double* const dz = sh_dz<M>(c);
  const double* const vp = sh_vp<M>(c);
  long long tc0 = g_clock();
  if (!pre_done) G_PAR_FOR(k, N) {
    double kap[NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) kap[a] = c.kap[a * np + k];
    ric_forward_pre<M>(c, k, kap);
  }
  G_SYNC();
  if (G_TID == 0 && GUSTO_PROF_CHAINS) c.prof[2] += g_clock() - tc0;
  chain_forward<M>(c);
  tc0 = g_clock();
  G_PAR_FOR(k, N) {
    double s[NX], u[NU], x[NX];
    const double* kt = c.kt + k;
    const double* gs = c.gs + k;
#pragma unroll
    for (int i = 0; i < NX; ++i) s[i] = dz[k * NV + i];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      double v = dz[k * NV + NX + a];
#pragma unroll
      for (int i = 0; i < NX; ++i) v -= kt[(i * NU + a) * np] * s[i];
      u[a] = v;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double v = s[i];
#pragma unroll
      for (int a = 0; a < NU; ++a) if (L::dsame(i, T::b_row(a))) v += gs[(a * NX + i) * np] * u[a];
      x[i] = v;
    }
    if (want_nu && k >= 1) {
      const double* pk = c.pk + k;
      const double* fi = c.fi + k;
      double w[NX], lam[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) w[i] = s[i] - c.ch[i * np + k - 1];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double v = -vp[k * NX + i];
#pragma unroll
        for (int j = 0; j < NX; ++j) v += pk[tri(i, j) * np] * w[j];
        lam[i] = v;
      }
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) if (L::dsame(i, j)) v += fi[(i * NX + j) * np] * lam[i];
        c.dnu[j * ne + k] = v;
      }
    }
    if (want_nu && k == N - 1) {
#pragma unroll
      for (int i = 0; i < NX; ++i) c.dnu[i * ne + N] = ((c.pmask >> i) & 1) ? c.wN * (x[i] - c.rnu[i * ne + N]) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) dz[k * NV + i] = x[i];
#pragma unroll
    for (int a = 0; a < NU; ++a) dz[k * NV + NX + a] = u[a];
  }
  G_SYNC();
  if (want_nu) {
    // row 0 (x_0 = x_init): stationarity in x_0,  dnu_0 = rx_0 - Hx_0 dx_0 - E_1' dnu_1,  E_1 = I + h/2 A_0
    if (G_TID == 0) {
      // (unrolled: static indices keep e1 in registers and let the loads go out together; rolled, with e1 in local memory, this
      // one-thread block was 1.7 % of the kernel's stall samples)
      double dx0[NX], hx[NX], e1[NX], d1[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) { dx0[i] = dz[i]; d1[i] = c.dnu[i * ne + 1]; e1[i] = d1[i]; }
      apply_Hx<M>(c, 0, dx0, hx);
#pragma unroll
      for (int e = 0; e < L::ANZ; ++e) e1[T::a_col(e)] += c.hh * c.Ac[e * np] * d1[T::a_row(e)];
#pragma unroll
      for (int i = 0; i < NX; ++i) c.dnu[i * ne] = c.r[i * np] - hx[i] - e1[i];
    }
    G_SYNC();
  }
  if (G_TID == 0 && GUSTO_PROF_CHAINS) c.prof[2] += g_clock() - tc0;
}
// TrajOpt: step of the l1 records of the dynamics rows, one thread per row j = 1 .. N-1 (modes as pair_step)
template <int M> GDEV void l1_rows_step(const IpmCtx<M>& c, double smu, int phase, int mode, double ap, double ad, StepAcc& acc) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX;
  const int N = c.N;
  const size_t ne = c.NE, fs = (size_t)NX * ne;
  G_PAR_FOR(j1, N - 1) {
    const int j = j1 + 1;
    double v[NX], adz[NX];
    aeq_row<M>(c, sh_z<M>(c), j, v);
    aeq_row<M>(c, sh_dz<M>(c), j, adz);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double* st = c.dslot + (size_t)i * ne + j;
      L1Pair q;
      l1_eval(st, fs, c.omega, c.nu[(size_t)i * ne + j], q);
      const double diff = v[i] + c.gsum[(size_t)i * ne + j] - q.p + q.n + adz[i];
      l1_step(st, fs, q, c.dnu[(size_t)i * ne + j], diff, smu, phase, mode, ap, ad, acc);
    }
  }
}

// --------------------------------------------------------------------------------------------- slot passes
// Flat pass over every live row.  FN(st, has_t, c0, gdz) is called once per row; `want_gdz` says whether the
// directional derivative gv.dz is needed.
// Slack reset (pair_step): on where the parity tolerances of the test-suite hold with it -- astrobeeSE3 with PointGoal rows
// (c.reset; the headline configurations).  Elsewhere the minimiser is flat in some directions (attitude inside the quaternion
// dead-band of astrobeeSE3manifold, position inside a BoxGoal, l1 rows of TrajOpt) or the SCP path is close to a decision boundary
// (freeflyerSE2's omega-escalation instances): there the end point depends on how the complementarity products are distributed, and
// with the reset the kernel lands 2e-5 (U) / 5e-4 (X) from the oracle's end point instead of 1e-6 -- inside the solver tolerance,
// outside the tests' -- so those keep the linear slack update.
template <int M> GHD constexpr bool slack_reset_on() { return !kTO && M == ASTROBEE_SE3; }
template <int M, typename FN> GDEV void for_each_row(const IpmCtx<M>& c, bool want_gdz, FN&& fn) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NV = L::NV;
  const int N = c.N;
  const size_t np = c.NP;
  G_PAR_FOR(it, N * L::SP) {
    const int s = it / N, k = it - s * N;                    // knot fastest: the field rows of a slot are read coalesced
    const double* x = sh_z<M>(c) + k * NV;
    double* st = c.sslot + (size_t)s * SLOT_W * np + k;
    if (T::HAS_TR && s == L::S_TR) {
      double v = -c.dow, gdz = 0.0, quad = 0.0;
      for (int i = 0; i < NX; ++i) {
        const double dxi = x[i] - c.xps[i * np + k], dzi = sh_dz<M>(c)[k * NV + i];
        v += dxi * dxi; gdz += 2.0 * dxi * dzi; quad += dzi * dzi;
      }
      fn(st, np, !kTO, v, gdz, c.reset ? quad : 0.0);
    } else {
      SpecEval o;
      spec_eval<M>(c, k, s, x, x + NX, o);
      if (!o.valid) continue;
      double gdz = 0.0, quad = 0.0;
      if (want_gdz) { const double* dv = sh_dz<M>(c) + k * NV + (o.is_u ? NX : 0) + o.i0; for (int a = 0; a < 4; ++a) if (a < o.n) { gdz += o.gv[a] * dv[a]; quad += 0.5 * o.hq[a] * dv[a] * dv[a]; } }
      fn(st, np, o.has_t, o.c0, gdz, c.reset ? quad : 0.0);
    }
  }
  if (c.bmask != 0) {
    G_PAR_FOR(j, L::NBOX) {
      if (!((c.bmask >> (j >> 1)) & 1)) continue;
      const double* x = sh_z<M>(c) + (N - 1) * NV;
      const double gdz = ((j & 1) == 0 ? 1.0 : -1.0) * sh_dz<M>(c)[(N - 1) * NV + (j >> 1)];
      fn(c.bslot + (size_t)j * SLOT_W, (size_t)1, false, box_c0<M>(c, j, x), gdz, 0.0);
    }
  }
  if (T::WS > 0) {
    constexpr int WS = T::WS > 0 ? T::WS : 1;
    const double* __restrict__ orow = c.orow;
    double* __restrict__ ost = c.ost;
    const double* zs = sh_z<M>(c);
    const double* dzs = sh_dz<M>(c);
    const int nact = c.nact;
    const size_t pp = c.PP;
    for (int p = G_TID; p < nact; p += G_NTHR) {
      const double* row = orow + p;
      const int k = (int)row[4 * pp];
      const double* x = zs + k * NV;
      const double* dv = dzs + k * NV;
      double v = row[3 * pp], gdz = 0.0;
#pragma unroll
      for (int a = 0; a < WS; ++a) { const double ra = row[a * pp]; v -= ra * x[a]; gdz -= ra * dv[a]; }
      fn(ost + p, pp, true, v, gdz, 0.0);
    }
  }
}

// One pass over the rows (see pair_step).  out[0..1]: largest primal / dual step to the boundary (modes 0, 1);
// out[2..4]: c0, c1, c2 (mode 1);  mode 2: installs the central-path floor for the next iteration.
template <int M> GDEV_NOINLINE void slot_steps(const IpmCtx<M>& c, int phase, double smu, int mode, double ap, double ad, double* out) {
  StepAcc acc; acc.amp = 1e300; acc.amd = 1e300; acc.c0 = 0.0; acc.c1 = 0.0; acc.c2 = 0.0; acc.np = 0.0;
  const double omega = c.omega;
  for_each_row<M>(c, true, [&](double* st, size_t ss, bool has_t, double c0, double gdz, double quad) {
    Pair q;
    pair_eval(st, ss, has_t, c0, omega, smu, phase, q);
    pair_step(st, ss, has_t, q, gdz, mode, ap, ad, acc, quad);
  });
  if (kTO) l1_rows_step<M>(c, smu, phase, mode, ap, ad, acc);
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

// Predictor pass (round 2): one pass over the rows, ONE THREAD PER KNOT, that does what the flat affine-step pass and the second
// assembly pass did together.  For every row of its knot the thread evaluates the affine step (step limits, the products
// ds dlam / dt dlamb kept in the slot record, the coefficients of Mehrotra's mu_aff) and accumulates how the row's right-hand
// side term  -gv bt  changes from the predictor to the corrector: bt is affine in the centering target and in the two products,
//   dbt = smu X + Y,   X = isa (1 - wa iw) - wa iw it,   Y = -pa isa (1 - wa iw) + pb wa iw it      (hard row: X = isa, Y = -pa isa)
// so  r_corr = r_aff + smu ra + rb  with  ra = -sum gv X,  rb = -sum gv Y  per knot -- known before sigma is.  After the group
// reductions every thread forms sigma and the centering target and the right-hand side is updated in place.  Returns smu.
template <int M> GDEV void pred_row(double* st, size_t ss, bool has_t, double c0, double gdz, double omega, StepAcc& acc, double& X, double& Y) {
  Pair q;
  pair_eval(st, ss, has_t, c0, omega, 0.0, 0, q);
  pair_step(st, ss, has_t, q, gdz, 1, 0.0, 0.0, acc);
  const double pa = st[4 * ss];
  if (has_t) {
    const double pb = st[5 * ss], wi = q.wa * q.iw, om = 1.0 - wi;
    X = q.isa * om - wi * q.itv;
    Y = -pa * q.isa * om + pb * wi * q.itv;
  } else {
    X = q.isa;
    Y = -pa * q.isa;
  }
}
template <int M> GDEV_NOINLINE double predictor_pass(const IpmCtx<M>& c, double npair, double mu, double tol, double* a_aff_out) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NV = L::NV;
  const int N = c.N;
  const size_t np = c.NP, pp = c.PP;
  const double omega = c.omega;
  StepAcc acc; acc.amp = 1e300; acc.amd = 1e300; acc.c0 = 0.0; acc.c1 = 0.0; acc.c2 = 0.0; acc.np = 0.0;
  G_PAR_FOR(k, N) {
    const double* x = sh_z<M>(c) + k * NV;
    const double* dv = sh_dz<M>(c) + k * NV;
    double ra[NV], rb[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { ra[i] = 0.0; rb[i] = 0.0; }
    double X, Y;
    if (T::HAS_TR) {
      double gv[NX], v = -c.dow, gdz = 0.0;
#pragma unroll
      for (int i = 0; i < NX; ++i) { const double dxi = x[i] - c.xps[i * np + k]; gv[i] = 2.0 * dxi; v += dxi * dxi; gdz += gv[i] * dv[i]; }
      pred_row<M>(c.sslot + (size_t)L::S_TR * SLOT_W * np + k, np, !kTO, v, gdz, omega, acc, X, Y);
#pragma unroll
      for (int i = 0; i < NX; ++i) { ra[i] -= gv[i] * X; rb[i] -= gv[i] * Y; }
    }
#pragma unroll
    for (int s = L::S_NORM; s < L::SP; ++s) {
      SpecEval o;
      spec_eval<M>(c, k, s, x, x + NX, o);
      if (!o.valid) continue;
      const int base = (spec_is_u<M>(s) ? NX : 0) + spec_i0<M>(s);
      double gdz = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) if (a < o.n) gdz += o.gv[a] * dv[base + a];
      pred_row<M>(c.sslot + (size_t)s * SLOT_W * np + k, np, o.has_t, o.c0, gdz, omega, acc, X, Y);
#pragma unroll
      for (int a = 0; a < 4; ++a) if (a < o.n) { ra[base + a] -= o.gv[a] * X; rb[base + a] -= o.gv[a] * Y; }
    }
    if (T::WS > 0) {
      constexpr int WS = T::WS > 0 ? T::WS : 1;
      const int s0 = sh_seg<M>(c)[k], s1 = sh_seg<M>(c)[k + 1];
      for (int p = s0; p < s1; ++p) {
        const double* row = c.orow + p;
        double g3[WS], v = row[3 * pp], gdz = 0.0;
#pragma unroll
        for (int a = 0; a < WS; ++a) { g3[a] = -row[a * pp]; v += g3[a] * x[a]; gdz += g3[a] * dv[a]; }
        pred_row<M>(c.ost + p, pp, true, v, gdz, omega, acc, X, Y);
#pragma unroll
        for (int a = 0; a < WS; ++a) { ra[a] -= g3[a] * X; rb[a] -= g3[a] * Y; }
      }
    }
    if (k == N - 1 && c.bmask != 0) {
      for (int j = 0; j < L::NBOX; ++j) {
        const int i = j >> 1;
        if (!((c.bmask >> i) & 1)) continue;
        const double sg = (j & 1) == 0 ? 1.0 : -1.0;
        pred_row<M>(c.bslot + (size_t)j * SLOT_W, 1, false, box_c0<M>(c, j, x), sg * dv[i], omega, acc, X, Y);
        for (int q2 = 0; q2 < NX; ++q2) if (q2 == i) { ra[q2] -= sg * X; rb[q2] -= sg * Y; }
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) { c.ra[i * np + k] = ra[i]; c.rb[i * np + k] = rb[i]; }
  }
  if (kTO) l1_rows_step<M>(c, 0.0, 0, 1, 0.0, 0.0, acc);      // affine step of the l1-penalised dynamics rows (dz, dnu of the predictor solve)
  const double amp = -block_max(-acc.amp, c.red), amd = -block_max(-acc.amd, c.red);
  const double s0 = block_sum(acc.c0, c.red), s1 = block_sum(acc.c1, c.red), s2 = block_sum(acc.c2, c.red);
  double a_aff = amp < amd ? amp : amd;
  a_aff = a_aff < 1.0 ? a_aff : 1.0;
  double mu_aff = s0 + a_aff * (s1 + a_aff * s2);
  mu_aff = npair > 0 ? mu_aff / npair : 0.0;
  double sigma = mu > 0 ? (mu_aff / mu) : 0.0;
  sigma = sigma * sigma * sigma;
  double smu = sigma * mu;
  smu = smu > 0.1 * tol ? smu : 0.1 * tol;
  G_PAR_FOR(it, N * NV) {
    const int i = it / N, k = it - i * N;
    const size_t o = i * np + k;
    c.r[o] += smu * c.ra[o] + c.rb[o];
  }
  if (kTO) {     // corrector right-hand side of the dynamics rows: beta changes by (smu - pa)/lp - (smu - pb)/ln, and ch with it
    const size_t ne = c.NE, fs = (size_t)NX * ne;
    G_PAR_FOR(j1, N - 1) {
      const int j = j1 + 1;
      double v[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const double* st = c.dslot + (size_t)i * ne + j;
        const double rho = c.rnu[(size_t)i * ne + j] + (smu - st[4 * fs]) * g_rcp(st[fs]) - (smu - st[5 * fs]) * g_rcp(st[3 * fs]);
        c.rnu[(size_t)i * ne + j] = rho;
        v[i] = -rho;
      }
      const double* fi = c.fi + j;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        double s2 = 0.0;
#pragma unroll
        for (int m = 0; m < NX; ++m) if (L::dsame(i, m)) s2 += fi[(size_t)(i * NX + m) * np] * v[m];
        c.ch[(size_t)i * np + j - 1] = s2;
      }
    }
  }
  G_SYNC();
  *a_aff_out = a_aff;
  return smu;
}

// TrajOpt: (re)start of the l1 records of the dynamics rows at the current primal iterate
template <int M> GDEV void l1_init_rows(const IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX;
  const int N = c.N;
  const size_t ne = c.NE, fs = (size_t)NX * ne;
  G_PAR_FOR(j1, N - 1) {
    const int j = j1 + 1;
    double v[NX];
    aeq_row<M>(c, sh_z<M>(c), j, v);
#pragma unroll
    for (int i = 0; i < NX; ++i) l1_init(c.dslot + (size_t)i * ne + j, fs, v[i] + c.gsum[(size_t)i * ne + j], c.omega);
  }
}

// --------------------------------------------------------------------------------------------------- setup
#ifndef GUSTO_SLACK_START_SE3
#define GUSTO_SLACK_START_SE3 0.0005
#endif
#ifndef GUSTO_SLACK_START_FF
#define GUSTO_SLACK_START_FF 1.0
#endif
// astrobeeSE3manifold (40 instances, real SCP solves of 6.9 iterations): (1, 0.5) 13.06 Newton iterations per solve, (0.01, 0.02)
// 8.47, (0.001, 0.05) 8.09, (0.001, 0.02) 7.77 (chosen), (0.0001, 0.02) 7.60 with more ALMOST_OPTIMAL endings; dubins (32
// instances): 6.03 -> 4.27.  freeflyerSE2 keeps the oracle's start: (0.1, 0.1) gives 9.98 -> 8.19 but the stalled solves of its
// omega-escalation instances end ALMOST_OPTIMAL more often, and its L3 comparison is the one closest to a decision boundary.
#ifndef GUSTO_SLACK_START_MAN
#define GUSTO_SLACK_START_MAN 0.001
#endif
#ifndef GUSTO_SLACK_START_DUB
#define GUSTO_SLACK_START_DUB 0.001
#endif
template <int M> GHD constexpr double slack_start() {
  return M == ASTROBEE_SE3 ? GUSTO_SLACK_START_SE3 : M == FREEFLYER_SE2 ? GUSTO_SLACK_START_FF : M == ASTROBEE_SE3_MANIFOLD ? GUSTO_SLACK_START_MAN : GUSTO_SLACK_START_DUB;
}
// ... and how the penalty weight omega = lam + lam_t (dual feasibility of t) is split at the start: most soft rows end
// inactive (lam -> 0, lam_t -> omega), so starting lam at 0.1 omega instead of 0.5 omega saves another 0.8 Newton
// iterations on astrobeeSE3 (8.01 -> 7.23, solve 5.87 -> 5.34 ms in round 1).  Round 2 (with the centrality safeguard in the
// driver): scanned on the CPU build of this source over real SCP solves (tools/newton_probe.py, 40 instances each of the easy and
// the hard tier; same optimum, same SCP decisions everywhere) -- (t_in, split) = (0.25, 0.1): 6.80 / 8.60 Newton iterations per
// solve (easy / hard), (0.1, 0.03): 6.23 / 8.18, (0.03, 0.05): 5.89 / 7.98, (0.02, 0.02): 5.65 / 8.16, (0.01, 0.01): 5.35 / 8.38,
// (0.005, 0.005): 5.06 / 8.66, (0.005, 0.02): 5.45 / 8.00, (0.001, 0.03): 5.28 / 7.83, (0.0005, 0.02): 5.08 / 7.91 (chosen: a small
// split costs iterations on the hard tier, a small t_in helps both; the slack itself never starts below 1e-2).
#ifndef GUSTO_LAM_SPLIT_SE3
#define GUSTO_LAM_SPLIT_SE3 0.02
#endif
#ifndef GUSTO_LAM_SPLIT_FF
#define GUSTO_LAM_SPLIT_FF 0.5
#endif
#ifndef GUSTO_LAM_SPLIT_MAN
#define GUSTO_LAM_SPLIT_MAN 0.02
#endif
#ifndef GUSTO_LAM_SPLIT_DUB
#define GUSTO_LAM_SPLIT_DUB 0.02
#endif
template <int M> GHD constexpr double slack_lam_split() {
  return M == ASTROBEE_SE3 ? GUSTO_LAM_SPLIT_SE3 : M == FREEFLYER_SE2 ? GUSTO_LAM_SPLIT_FF : M == ASTROBEE_SE3_MANIFOLD ? GUSTO_LAM_SPLIT_MAN : GUSTO_LAM_SPLIT_DUB;
}
template <int M> GDEV_NOINLINE void setup(IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV, ANZ = L::ANZ;
  const int N = c.N;
  double tin_ = slack_start<M>();
  // start point: X, U <- previous trajectory (set_start_value, scp_gusto.jl:100-102); multipliers 0
  const size_t np = c.NP, ne = c.NE, pp = c.PP;
  G_PAR_FOR(it, N * NV) { const int k = it / NV, i = it - k * NV; sh_z<M>(c)[it] = i < NX ? c.Xp[k * NX + i] : c.Up[k * NU + i - NX]; }
  G_PAR_FOR(it, NX * (int)ne) c.nu[it] = 0.0;
  // field-major copies of what the per-knot passes read every Newton iteration besides the linearize kernel's blocks (A_k on its
  // sparsity pattern is read in place): Xp, and the constant part of the equality residual  -x_init | h/2 (g_{j-1} + g_j) | -goal
  G_PAR_FOR(it, N * NX) { const int i = it / N, k = it - i * N; c.xps[i * np + k] = c.Xp[k * NX + i]; }
  G_PAR_FOR(it, (N + 1) * NX) {
    const int i = it / (N + 1), j = it - i * (N + 1);
    double v;
    if (j == 0) v = -c.x_init[i];
    else if (j == N) v = ((c.pmask >> i) & 1) ? -c.goal_lo[i] : 0.0;
    else v = c.hh * (c.g[i * np + j - 1] + c.g[i * np + j]);
    c.gsum[i * ne + j] = v;
  }
  // obstacle rows inside the toggle distance, compacted knot by knot (astrobee_se3.jl:293)
  if (T::WS > 0) {
    G_PAR_FOR(k, N) {
      int cnt = 0;
      const double* dist = c.rows + 4 * (size_t)c.n_obs * np + k;                    // field 4 (dist0) of [5][n_obs][NP]
      for (int i = 0; i < c.n_obs; ++i) cnt += (dist[i * np] < c.toggle) ? 1 : 0;
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
  // Compaction (thread per knot): the distances of a knot's obstacles are loaded first (independent loads, one round trip) into a
  // bit mask, then the live rows are copied; their start values go through the record's first field to the flat pass below that
  // initialises the records once the start point has been measured.  (One loop with the test in front of each row's loads was a
  // chain of n_obs dependent round trips per thread: 2 % of the kernel's stall samples, three times over.)
  double vloc = 0.0, sloc = 1e300;
  if (T::WS > 0) {
    G_PAR_FOR(k, N) {
      int p = sh_seg<M>(c)[k];
      const double* x = sh_z<M>(c) + k * NV;
      const size_t fs = (size_t)c.n_obs * np;
      unsigned long long live = 0ull;
      for (int i = 0; i < c.n_obs; ++i) live |= (unsigned long long)(c.rows[4 * fs + (size_t)i * np + k] < c.toggle ? 1 : 0) << i;
      for (int i = 0; i < c.n_obs; ++i) {
        if (!((live >> i) & 1ull)) continue;
        const double* row = c.rows + (size_t)i * np + k;
        double* o = c.orow + p;
        const double off = row[3 * fs];
        double v = off;
        for (int a = 0; a < 3; ++a) { const double ra = row[a * fs]; o[a * pp] = ra; if (a < T::WS) v -= ra * x[a]; }
        o[3 * pp] = off; o[4 * pp] = (double)k;
        c.ost[p] = v;
        vloc = v > vloc ? v : vloc;
        if (-v < sloc) sloc = -v;
        ++p;
      }
    }
  }
  // Centred start: on for astrobeeSE3 (GuSTO subproblem) while the penalty weight has not been escalated.  Measured on the CPU build
  // over real SCP solves as the sum over the launches of the slowest instance's Newton count (what a launch costs): astrobeeSE3
  // 15 -> 11, its hard tier 167 -> 167 (with the omega gate; instances past an escalation sit on their hinge rows and a start that
  // close to them stalls: 30-50 iterations), dubins 198 -> 194, freeflyerSE2 158 -> 160, astrobeeSE3manifold 100 -> 146, and the
  // TrajOpt subproblem grows stragglers (25 iterations): those keep the tuned start.
  if (!kTO && M == ASTROBEE_SE3 && c.mu0_a > 0.0 && c.omega <= c.d->sp[SP_OMEGA0]) {
    // centred start when the start point violates no soft row (see slot_init): vmax over the live obstacle rows and the special rows
    G_PAR_FOR(it, N * L::SP) {
      const int s2 = it / N, k = it - s2 * N;
      if (T::HAS_TR && s2 == L::S_TR) continue;                  // x = xp at the start: never violated
      const double* x = sh_z<M>(c) + k * NV;
      SpecEval o; spec_eval<M>(c, k, s2, x, x + NX, o);
      if (o.valid && o.c0 > vloc) vloc = o.c0;
    }
    const double vmax = block_max(vloc, c.red);
    double mu0 = c.mu0_a * c.omega;
    // A nearly converged SCP iteration starts next to its optimum: there the solve can start on the central path at (almost) the
    // final mu and needs ONE Newton step.  The start point's primal infeasibility rp0 (0.1 for a straight line, 3e-4 / 1e-7 in the
    // second / third iteration of the headline problems) measures that, provided nothing else will make the iterate move: the trust
    // region has not been shrunk by a rejection and no live obstacle row is within mu0_smin (5 cm) of its hinge.  astrobeeSE3 hard
    // tier: instances about to touch an obstacle stall for 20-38 iterations from such a start; sum over the 17 launches of a B = 1024
    // solve of the slowest instance's Newton count (GPU): 241 without the scaling, 267 / 243 / 241 with the gate at 2 / 5 / 7 cm, while
    // the headline batch (margins 5-20 cm) keeps its full gain up to 5 cm (bench 388 k -> 453 k instance-iterations/s; 415 k at 7 cm).
    // astrobeeSE3: Newton iterations 4 / 3.2 / 3 -> 4 / 2.6 / 2 over the three SCP iterations of a solve.
    const double smin = -block_max(-sloc, c.red);
    if (c.mu0_rp > 0.0 && c.Delta >= c.d->sp[SP_DELTA0] && smin >= c.mu0_smin) {
      double rpl = 0.0;
      G_PAR_FOR(j, N + 1) {
        double v[NX];
        aeq_row<M>(c, sh_z<M>(c), j, v);
#pragma unroll
        for (int i = 0; i < NX; ++i) { const double a2 = fabs(v[i] + c.gsum[(size_t)i * c.NE + j]); rpl = a2 > rpl ? a2 : rpl; }
      }
      const double rp0 = block_max(rpl, c.red);
      const double f = rp0 * g_rcp(c.mu0_rp);
      if (f < 1.0) mu0 *= f;
      if (mu0 < c.mu0_lo * c.omega) mu0 = c.mu0_lo * c.omega;
    }
    if (c.mu0_b * c.omega * vmax > mu0) mu0 = c.mu0_b * c.omega * vmax;
    if (mu0 <= c.mu0_cap) tin_ = -mu0;
  } else {
    G_SYNC();                                                      // the compacted rows are read by other threads below
  }
  const double tin = tin_;
  if (T::WS > 0) {
    const int nact = c.nact;
    for (int p = G_TID; p < nact; p += G_NTHR) slot_init(c.ost + p, pp, true, true, c.ost[p], c.omega, tin, slack_lam_split<M>());
  }
  // special rows: slacks one unit inside
  G_PAR_FOR(it, N * L::SP) {
    const int s = it / N, k = it - s * N;
    const double* x = sh_z<M>(c) + k * NV;
    double* st = c.sslot + (size_t)s * SLOT_W * np + k;
    if (T::HAS_TR && s == L::S_TR) slot_init(st, np, true, !kTO, -c.dow, c.omega, tin, slack_lam_split<M>());   // x = xp at the start
    else { SpecEval o; spec_eval<M>(c, k, s, x, x + NX, o); slot_init(st, np, o.valid, o.has_t, o.c0, c.omega, tin, slack_lam_split<M>()); }
  }
  G_PAR_FOR(j, L::NBOX) {
    const bool valid = (c.bmask >> (j >> 1)) & 1;
    slot_init(c.bslot + (size_t)j * SLOT_W, 1, valid, false, valid ? box_c0<M>(c, j, sh_z<M>(c) + (N - 1) * NV) : 0.0, c.omega);
  }
  G_SYNC();
  if (kTO) { l1_init_rows<M>(c); G_SYNC(); }
}

// Restart of a solve that cycles: keep the primal iterate z and the equality multipliers, put every slack / multiplier pair back
// to the well-centred cold start of the oracle (slacks one unit inside, penalty weight split evenly) evaluated AT z, and drop the
// pending centrality floor.  The obstacle-row compaction (a function of Xp only) is unchanged.
template <int M> GDEV_NOINLINE void restart_slots(IpmCtx<M>& c) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NV = L::NV;
  const int N = c.N;
  const size_t np = c.NP, pp = c.PP;
  G_PAR_FOR(it, N * L::SP) {
    const int s = it / N, k = it - s * N;
    const double* x = sh_z<M>(c) + k * NV;
    double* st = c.sslot + (size_t)s * SLOT_W * np + k;
    if (T::HAS_TR && s == L::S_TR) slot_init(st, np, true, !kTO, tr_c0<M>(c, k, x), c.omega);
    else { SpecEval o; spec_eval<M>(c, k, s, x, x + NX, o); slot_init(st, np, o.valid, o.has_t, o.c0, c.omega); }
  }
  G_PAR_FOR(j, L::NBOX) {
    const bool valid = (c.bmask >> (j >> 1)) & 1;
    slot_init(c.bslot + (size_t)j * SLOT_W, 1, valid, false, valid ? box_c0<M>(c, j, sh_z<M>(c) + (N - 1) * NV) : 0.0, c.omega);
  }
  if (T::WS > 0) {
    constexpr int WS = T::WS > 0 ? T::WS : 1;
    for (int p = G_TID; p < c.nact; p += G_NTHR) {
      const double* row = c.orow + p;
      const double* x = sh_z<M>(c) + (int)row[4 * pp] * NV;
      double v = row[3 * pp];
#pragma unroll
      for (int a = 0; a < WS; ++a) v -= row[a * pp] * x[a];
      slot_init(c.ost + p, pp, true, true, v, c.omega);
    }
  }
  if (kTO) l1_init_rows<M>(c);
  if (G_TID == 0) c.floor_ = 0.0;
  G_SYNC();
}

#ifdef GUSTO_HOSTSIM
// Host-simulation self check (GUSTO_HOSTSIM_CHECK=1): residual of the Newton system the Riccati solve claims to have solved,
//   H dz + Aeq' dnu = r   and   Aeq dz - D dnu = rnu   (D = 0 outside TrajOpt; the PointGoal rows carry 1/w_N).
template <int M> inline void kkt_check(const IpmCtx<M>& c, const char* tag) {
  using L = IpmLayout<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const int N = c.N;
  const double* dz = c.dz;
  double e1 = 0, e2 = 0, n1 = 0, n2 = 0;
  for (int k = 0; k < N; ++k) {
    double at[NV], hx[NX], hu[NU], kd[L::KDW];
    aeqT_knot<M>(c, c.dnu, k, at);
    apply_Hx<M>(c, k, dz + k * NV, hx);
    for (int i = 0; i < L::KDW; ++i) kd[i] = c.kd[(size_t)i * c.NP + k];
    ublocks_mv<M>(kd + L::KD_HU, dz + k * NV + NX, hu);
    for (int i = 0; i < NV; ++i) {
      const double lhs = (i < NX ? hx[i] : hu[i - NX]) + at[i], rhs = c.r[(size_t)i * c.NP + k];
      e1 = fmax(e1, fabs(lhs - rhs)); n1 = fmax(n1, fabs(rhs));
      if (getenv("GUSTO_HOSTSIM_CHECK")[0] == '2' && fabs(lhs - rhs) > 1e-7) printf("      stat k=%d i=%d lhs %.6e rhs %.6e\n", k, i, lhs, rhs);
    }
  }
  for (int j = 0; j < N; ++j) {
    double v[NX];
    aeq_row<M>(c, dz, j, v);
    for (int i = 0; i < NX; ++i) {
      const double D = (kTO && j >= 1) ? c.dd[(size_t)i * c.NE + j] : 0.0;
      const double lhs = v[i] - D * c.dnu[(size_t)i * c.NE + j], rhs = c.rnu[(size_t)i * c.NE + j];
      e2 = fmax(e2, fabs(lhs - rhs)); n2 = fmax(n2, fabs(rhs));
      if (getenv("GUSTO_HOSTSIM_CHECK")[0] == '2' && fabs(lhs - rhs) > 1e-7) printf("      row j=%d i=%d lhs %.6e rhs %.6e\n", j, i, lhs, rhs);
    }
  }
  printf("    kkt[%s] stationarity %.3e (rhs %.3e)  rows %.3e (rhs %.3e)\n", tag, e1, n1, e2, n2);
}
#endif

// ------------------------------------------------------------------------------------------------ driver
// scratch: IpmLayout<M>::scratch_doubles() doubles of global memory owned by this instance (16-byte aligned, zero-filled
//          once at allocation: the padding columns of the tile records are never written).
// smem:    IpmLayout<M>::smem_doubles() doubles of shared memory.
// On exit Xn/Un of the instance hold the solution and info[IPM_NINFO] = {status, iters, res, mu, obj,...}.
template <int M>
GDEV void ipm_solve_instance(const BatchDesc& d, const BatchPtrs& p, const IpmParams& prm, int b, double* scratch,
                             double* smem, double* info) {
  using L = IpmLayout<M>;
  using T = Traits<M>;
  constexpr int NX = L::NX, NU = L::NU, NV = L::NV;
  const int N = d.N;
  // The context is ONE struct per instance in shared memory (filled by thread 0): phase functions read it with
  // shared loads instead of per-thread local-memory copies.
  static_assert(sizeof(IpmCtx<M>) <= (size_t)L::CTX_DOUBLES * sizeof(double), "IpmCtx does not fit its shared-memory slot");
  IpmCtx<M>& c = *reinterpret_cast<IpmCtx<M>*>(smem);
  smem += L::CTX_DOUBLES;
  if (G_TID == 0) {
    c.d = &d; c.rp = d.rp; c.N = N; c.n_obs = (T::WS > 0) ? d.n_obs : 0; c.b = b; c.nact = 0;
    c.h = p.tf[b] / (N - 1); c.hh = 0.5 * c.h; c.omega = p.omega[b]; c.Delta = p.delta[b];
    c.toggle = kTO ? d.rp[RP_CLEAR] + 1.0 : c.Delta / 8.0 + d.rp[RP_CLEAR]; c.eps = d.sp[SP_EPS];   // scp_gusto.jl:76 | scp_trajopt.jl:65
    // TrajOpt: the noise phase forms P~ = P - X'X at the last knot, where P carries w_N: a smaller penalty keeps that difference
    // accurate (the row error after a step is dnu_N / w_N and vanishes with the step, as a proximal multiplier update does)
    c.mu0_a = prm.mu0_a; c.mu0_b = prm.mu0_b; c.mu0_cap = prm.mu0_cap; c.mu0_rp = prm.mu0_rp; c.mu0_lo = prm.mu0_lo; c.mu0_smin = prm.mu0_smin;
    c.wN = kTO ? 1e-3 * (prm.wn_base + prm.wn_omega * c.omega) : prm.wn_base + prm.wn_omega * c.omega;
    c.dow = kTO ? c.Delta : c.Delta / c.omega; c.eow = c.eps / c.omega;
    c.pmask = 0; c.bmask = 0;
    for (int i = 0; i < NX; ++i) { if (d.goal_type[i] == GOAL_POINT) c.pmask |= 1 << i; if (d.goal_type[i] == GOAL_BOX) c.bmask |= 1 << i; }
    c.reset = (slack_reset_on<M>() && c.bmask == 0) ? 1 : 0;
    c.Xp = p.Xp + (size_t)b * N * NX; c.Up = p.Up + (size_t)b * N * NU;
    c.Ac = p.A + (size_t)b * L::ANZ * g_np(N); c.g = p.g + (size_t)b * NX * g_np(N);
    c.rows = p.rows + (size_t)b * 5 * d.n_obs * g_np(N);
    c.x_init = p.x_init + (size_t)b * NX; c.goal_lo = p.goal_lo + (size_t)b * NX; c.goal_hi = p.goal_hi + (size_t)b * NX;
    {
      double Bm[NX * NU];
      for (int i = 0; i < NX * NU; ++i) Bm[i] = 0.0;
      dyn_B<M>(d.rp, Bm);
      for (int a = 0; a < NU; ++a) c.bv[a] = Bm[T::b_row(a) * NU + a];
    }
    c.NP = L::np_of(N); c.NE = L::ne_of(N); c.PP = L::pp_of(N, c.n_obs);
    const size_t np = c.NP, ne = c.NE, pp = c.PP, nn = (size_t)N;
    double* q = scratch;                                         // same order and sizes as IpmLayout::scratch_doubles
    c.r = q; q += np * NV; c.ra = q; q += np * NV; c.rb = q; q += np * NV;
    c.nu = q; q += ne * NX; c.dnu = q; q += ne * NX; c.rnu = q; q += ne * NX;
    c.gsum = q; q += ne * NX;
    c.xps = q; q += np * NX;
    c.sslot = q; q += np * L::SP * SLOT_W;
    c.bslot = q; q += (size_t)L::NBOX * SLOT_W;
    c.ost = q; q += pp * SLOT_W;
    c.orow = q; q += pp * OROW_W;
    c.kd = q; q += np * L::KDW;
    c.fi = q; q += np * L::NN;
    c.cr = q; q += nn * L::CRW;
    c.gs = q; q += np * L::GSW; c.bs = q; q += np * L::GSW;
    c.acl = q; q += nn * L::GT;
    c.kt = q; q += np * NX * NU;
    c.lp = q; q += np * L::NTU;
    c.pk = q; q += np * L::NTX;
    c.psi = q; q += np * NX;
    c.ch = q; q += np * NX;
    c.kap = q; q += np * NU;
    if (kTO) {
      c.dslot = q; q += ne * NX * L::DSLOT_W; c.dd = q; q += ne * NX;
      c.pkt = q; q += np * L::NTX; c.nm = q; q += np * L::NN;
    } else { c.dslot = nullptr; c.dd = nullptr; c.pkt = nullptr; c.nm = nullptr; }
    c.z = smem; c.dz = c.z + L::rnd((size_t)N * NV);
    c.vp = c.dz + L::work_doubles(N);
    c.red = c.vp + L::rnd((size_t)(N + 1) * NX);
    c.seg = reinterpret_cast<int*>(c.red + G_NTHR + 16);
    c.tab = reinterpret_cast<unsigned char*>(c.red + G_NTHR + 16 + L::seg_doubles(N));
    for (int i = 0; i < 4; ++i) c.prof[i] = 0;
    c.floor_ = 0.0;
    unsigned char* mt = c.tab;                                   // a_row | a_col | blk_of
    for (int e = 0; e < L::ANZ; ++e) { mt[e] = (unsigned char)T::a_row(e); mt[L::ANZ + e] = (unsigned char)T::a_col(e); }
    for (int i = 0; i < NX; ++i) mt[2 * L::ANZ + i] = (unsigned char)T::XB_of(i);
  }
  G_SYNC();

  int status = IPM_ITERATION_LIMIT, it_done = 0;
  long long cyc_asm = 0, cyc_fac = 0, cyc_sol = 0, cyc_slot = 0, tc0;
  double res = 1e300, mu = 0;
  const double scd = 1.0 + c.omega;
  setup<M>(c);
  tc0 = g_clock();
  setup_dynamics<M>(c);
  const long long cyc_setup = g_clock() - tc0;
  double best = 1e300;
  int stall = 0;
  bool polish = false;
  bool broke = false;
  int restarts = 0;
  for (int iter = 1; iter <= prm.max_iter; ++iter) {
    G_CTA_RESYNC();
    ++it_done;
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
    // With the slack reset (pair_step) the row residuals no longer lag behind, and the first iterate with res <= tol can carry a dual
    // residual just under the tolerance -- 1e-9 relative, which the weakly curved state directions (cost on U only) turn into 1e-4 in
    // X.  The solves used to end far below it (the lagging row residual forced one or two more quadratically convergent steps), and
    // the parity tolerances are calibrated on that; likewise the complementarity, whose average mu times the ~1e3 pairs is the error
    // of the objective, used to end at its floor 0.1 tol.  So: accept at once when the dual residual is 100x under the tolerance and
    // mu is at its floor, else take ONE more Newton step and accept then.
    if (res <= prm.tol) {
      if (!c.reset || (R.rz / scd <= GUSTO_RD_FACTOR * prm.tol && mu <= 0.15 * prm.tol) || polish) { status = IPM_OPTIMAL; break; }
      polish = true;
    } else {
      polish = false;
    }
    if (!(res == res) || res > 1e200) { status = IPM_NUMERICAL; break; }
    // The best iterate is kept in Xn / Un (as the oracle keeps it, ipm.py): a solve whose dual residual wanders once the
    // complementarity sits at its floor is stopped after 8 iterations without improvement and answers with that iterate.
    if (res < best) {
      best = res; stall = 0;
      G_PAR_FOR(k, N) {
        for (int i = 0; i < NX; ++i) p.Xn[((size_t)b * N + k) * NX + i] = sh_z<M>(c)[k * NV + i];
        for (int i = 0; i < NU; ++i) p.Un[((size_t)b * N + k) * NU + i] = sh_z<M>(c)[k * NV + NX + i];
      }
    } else if (++stall >= 8 && best <= 1e3 * prm.tol) break;
    // predictor: factorisation sweep with the backward vector pass fused, then the forward pass
    tc0 = g_clock();
    const bool fac_ok = riccati_factor<M>(c);
    cyc_fac += g_clock() - tc0;
    if (!fac_ok) { broke = true; break; }                       // non-positive pivot of a Lam_k: see below
    tc0 = g_clock();
    ric_forward<M>(c, kTO);                                       // TrajOpt: the affine step of the l1 rows needs dnu
    cyc_sol += g_clock() - tc0;
#ifdef GUSTO_HOSTSIM
    if (getenv("GUSTO_HOSTSIM_CHECK")) { if (!kTO) ric_forward<M>(c, true); kkt_check<M>(c, "pred"); }
#endif
    double am[5];
    tc0 = g_clock();
    double a_aff = 1.0;
    const double smu = predictor_pass<M>(c, R.npair, mu, prm.tol, &a_aff);     // affine step, sigma, corrector right-hand side
    cyc_slot += g_clock() - tc0;
    tc0 = g_clock();
    ric_backward<M>(c, true);
    ric_forward<M>(c, true, true);
    cyc_sol += g_clock() - tc0;
#ifdef GUSTO_HOSTSIM
    if (getenv("GUSTO_HOSTSIM_CHECK")) kkt_check<M>(c, "corr");
#endif
    tc0 = g_clock();
    slot_steps<M>(c, 1, smu, 0, 0, 0, am);
    double tau = 0.995;
    if (mu < 1.0) { tau = 1.0 - mu; tau = tau > 0.995 ? tau : 0.995; tau = tau < 0.999999 ? tau : 0.999999; }
    double ap = tau * am[0], ad = tau * am[1];
    ap = ap < 1.0 ? ap : 1.0; ad = ad < 1.0 ? ad : 1.0;
#ifdef GUSTO_HOSTSIM
    if (getenv("GUSTO_HOSTSIM_VERBOSE")) printf("        a_aff=%.3g smu=%.3g ap=%.3g ad=%.3g\n", a_aff, smu, ap, ad);
#endif
    slot_steps<M>(c, 1, smu, 2, ap, ad, am);
    // Safeguard against cycling: plain Mehrotra steps can jam when a few complementarity pairs sit far below the central path
    // (astrobeeSE3manifold instance 347 of the B = 1024 batch: BoxGoal rows 2e-4 apart, steps alternate between a blocked
    // dual and a blocked primal step, the residual wanders at 2e-3 for 50 iterations).  After two iterations without
    // improvement and a short step, the pairs are pulled back to 1e-2 mu instead of 1e-4 mu for the next iteration; solves
    // that improve every iteration never get here.
    if (stall >= 2 && (ap < ad ? ap : ad) < 0.5) { if (G_TID == 0) c.floor_ *= 100.0; G_SYNC(); }
    // ... and when that does not help either (astrobeeSE3manifold instance 151, fifth SCP iteration: a 4-cycle at mu = 5e-9 with the
    // dual residual at 1e-2), the pairs are re-centred from scratch around the current primal iterate, at most twice per solve.
    if (stall >= 12 && restarts < 2) { ++restarts; stall = 0; restart_slots<M>(c); }
    G_PAR_FOR(it, N * NV) sh_z<M>(c)[it] += ap * sh_dz<M>(c)[it];
    {
      double* __restrict__ nu = c.nu;
      const double* __restrict__ dnu = c.dnu;
      const int ne = c.NE * NX;                                  // (padding entries are zero and stay zero)
#pragma unroll 4
      for (int it = G_TID; it < ne; it += G_NTHR) nu[it] += ad * dnu[it];
    }
    G_SYNC();
    cyc_slot += g_clock() - tc0;
  }
  // A stalled solve is accepted only when its best residual is both within 1e3*tol and far below the SCP's own soft-row
  // threshold eps (GuSTO's accept test compares raw row values with eps, scp_gusto.jl:318-327), and is labelled as such.
  // The same holds for a factorisation that breaks down (a Lam_k loses positive definiteness in floating point) once the iterate
  // is that close: at omega = 1e6 and mu = 1e-9 the active soft rows put lam / s ~ 1e15 into Hx, and P_k = W - Y'Y cancels to
  // O(1) errors (freeflyerSE2 instance 10 of the L3 batch, Newton iteration 19, residual 1.6e-7) -- the attainable precision is
  // reached.  A breakdown anywhere else is a NUMERICAL failure (info[3] = -1 tells it apart from a NaN).
  const bool use_best = status == IPM_ITERATION_LIMIT && best <= 1e3 * prm.tol && best <= 0.1 * c.eps;
  if (broke && !use_best) { status = IPM_NUMERICAL; mu = -1.0; }
  if (use_best) {
    status = IPM_ALMOST_OPTIMAL;
    res = best;
    G_SYNC();
    G_PAR_FOR(k, N) {
      for (int i = 0; i < NX; ++i) sh_z<M>(c)[k * NV + i] = p.Xn[((size_t)b * N + k) * NX + i];
      for (int i = 0; i < NU; ++i) sh_z<M>(c)[k * NV + NX + i] = p.Un[((size_t)b * N + k) * NU + i];
    }
    G_SYNC();
  }
  {   // a NaN/Inf anywhere in the iterate is a numerical failure, never an answer
    double badz = 0.0;
    G_PAR_FOR(it, N * NV) { const double v = sh_z<M>(c)[it]; if (!(v == v) || fabs(v) > 1e100) badz = 1.0; }
    if (block_max(badz, c.red) > 0.0) status = IPM_NUMERICAL;
  }
  // ---- write the candidate trajectory and the objective (cost + omega * sum t; for a kept best iterate the slacks are
  //      taken at their optimal values t = max(c0, 0))
  double obj = 0;
  G_PAR_FOR(k, N) {
    const double wk = (k == 0 || k == N - 1) ? 0.5 * c.h : c.h;
    for (int i = 0; i < NX; ++i) p.Xn[((size_t)b * N + k) * NX + i] = sh_z<M>(c)[k * NV + i];
    for (int i = 0; i < NU; ++i) { const double uv = sh_z<M>(c)[k * NV + NX + i]; p.Un[((size_t)b * N + k) * NU + i] = uv; obj += wk * uv * uv; }
  }
  {
    const double omega = c.omega;
    for_each_row<M>(c, false, [&](double* st, size_t ss, bool has_t, double c0, double, double) { if (has_t) obj += omega * (use_best ? (c0 > 0.0 ? c0 : 0.0) : st[2 * ss]); });
    if (kTO) {   // + mu |d_j(z)|_1 over the dynamics rows (the l1 slacks at their optimal values)
      G_PAR_FOR(j1, N - 1) {
        double v[NX];
        aeq_row<M>(c, sh_z<M>(c), j1 + 1, v);
        for (int i = 0; i < NX; ++i) obj += omega * fabs(v[i] + c.gsum[(size_t)i * c.NE + j1 + 1]);
      }
    }
  }
  // SCPS.dual (scp_gusto.jl:116, get_dual_jump): row 0 of Aeq is  x_0 = x_init  and the Lagrangian is f + nu'(Aeq z - b)
  if (p.dual) G_PAR_FOR(i, NX) p.dual[(size_t)b * NX + i] = c.nu[(size_t)i * c.NE];
  obj = block_sum(obj, c.red);
  if (G_TID == 0) {
    info[0] = (double)status; info[1] = (double)it_done; info[2] = res; info[3] = mu; info[4] = obj;
#if !defined(GUSTO_PROF_MODE) || GUSTO_PROF_MODE == 0
    info[5] = (double)(cyc_asm + cyc_slot); info[6] = (double)cyc_fac; info[7] = (double)cyc_sol;   // SM cycles per phase
#elif GUSTO_PROF_MODE == 1      // developer builds: sweep | chains | per-knot passes
    info[5] = (double)c.prof[0]; info[6] = (double)c.prof[1]; info[7] = (double)c.prof[2];
#elif GUSTO_PROF_MODE == 2      // ... setup of the dynamics records | assembly | row passes
    info[5] = (double)cyc_setup; info[6] = (double)cyc_asm; info[7] = (double)cyc_slot;
#else                           // ... phases of the sweep: A | B | C (info[3] = D)
    info[5] = (double)c.prof[0]; info[6] = (double)c.prof[1]; info[7] = (double)c.prof[2]; info[3] = (double)c.prof[3];
#endif
    (void)cyc_setup;
  }
}

}  // namespace ipm_gusto / ipm_trajopt
}  // namespace gusto
#undef GUSTO_IPM_NS_OPEN
#undef GUSTO_PROF_TICK

#ifndef GUSTO_HOSTSIM
#undef block_sum
#undef block_max
#pragma pop_macro("G_SYNC")
#pragma pop_macro("G_NTHR")
#pragma pop_macro("G_TID")
#endif
