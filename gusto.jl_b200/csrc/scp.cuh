// Device-resident GuSTO outer step: accept / reject, Delta / omega schedule and convergence test of
// /root/reference/src/scp/scp_gusto.jl:119-174, one CTA per instance, fused with the trajectory copy of an accepted
// candidate (copy!(SCPS.traj, new_traj), :147).  It is the same decision table as the host-language versions
// (julia/GuSTOB200.jl, host.py::gusto_update) -- those remain the reference for a host that wants the loop in its own
// language; this kernel lets a whole solve_gusto_jump! run without a host round trip per iteration (gusto_scp_run).
// The decision itself (scp_update_instance) is plain host/device code so that the test-only host simulation can check it
// against host.py::gusto_update without a GPU.
#pragma once
#include "common.cuh"

namespace gusto {

// per-iteration history record, one per instance (SCPSolution vectors types.jl:150-173 + SCPParam_GuSTO vectors scp_gusto.jl:15-19)
enum : int { H_JTRUE = 0, H_JFULL, H_SCP_STATUS, H_SOLVER_STATUS, H_ACCEPT, H_CONV, H_DELTA, H_OMEGA, H_RHO, H_TR_OK, H_INEQ_OK, H_NEWTON, HIST_W };
enum : int { SCP_NA = 0, SCP_OK, SCP_INACCURATE, SCP_VIOLATES, SCP_TR_VIOLATED, SCP_SOLVER_FAILED, SCP_INACTIVE };
// per-iteration counters of this rank: instances that ran, whose solve was usable, accepted, still active afterwards
enum : int { C_RAN = 0, C_SOLVED, C_ACCEPTED, C_ACTIVE_AFTER, CNT_W };

struct ScpState {
  int* slot;               // [1] index of the outer iteration being processed (history slot = *slot + 1)
  int* iterations;         // [B] SCPS.iterations
  double* conv_prev;       // [B] convergence_measure[end]
  double* j_true;          // [B] J_true[end]
  double* j_full;          // [B]
  uint8_t* active;         // [B] still iterating (the same bytes as BatchPtrs::active, writable here)
  uint8_t* converged;      // [B]
  uint8_t* successful;     // [B]
  uint8_t* done;           // [2][B] 1 = finished; written by every instance every iteration into half (*slot & 1):
                           //        the send buffer of the status all-gather
  int* cnt;                // [max_hist][CNT_W]
  double* hist;            // [max_hist + 1][B][HIST_W]
  int max_hist;
  int force;               // the reference's force flag (scp_gusto.jl:55,173): never stop on convergence
};

// One instance, one outer iteration.  Reads the evaluation scalars e[] (K4) and the solver record inf[] (K3), updates the
// per-instance state, writes the history record rec[HIST_W] and returns flags: bit 0 accept, bit 1 done, bit 2 ran, bit 3 solved.
GHD int scp_update_instance(const double* e, const double* inf, const double* sp, bool force, bool was_active,
                                   double* delta, double* omega, int* iterations, double* conv_prev, double* j_true, double* j_full,
                                   uint8_t* converged, uint8_t* successful, double* rec) {
  const double Delta = *delta, w = *omega;
  if (!was_active) {
    rec[H_JTRUE] = *j_true; rec[H_JFULL] = *j_full; rec[H_SCP_STATUS] = SCP_INACTIVE; rec[H_SOLVER_STATUS] = -1;
    rec[H_ACCEPT] = 0; rec[H_CONV] = *conv_prev; rec[H_DELTA] = Delta; rec[H_OMEGA] = w; rec[H_RHO] = 0;
    rec[H_TR_OK] = 0; rec[H_INEQ_OK] = 0; rec[H_NEWTON] = 0;
    return 2;
  }
  const int sstat = (int)inf[0];
  const bool ok = sstat == 0 || sstat == 3;                       // OPTIMAL / ALMOST_OPTIMAL continue (scp_gusto.jl:107)
  const double D0 = sp[SP_DELTA0], w_max = sp[SP_OMEGAMAX], rho0 = sp[SP_RHO0], rho1 = sp[SP_RHO1];
  const double conv = e[0], rho = e[3];
  const bool tr_ok = e[1] > 0.5, ineq_ok = e[2] > 0.5;
  double Dn = Delta, wn = w;
  int status = SCP_SOLVER_FAILED, accept = 0;
  bool done = !ok, conv_now = false, succ_now = false;
  if (ok) {
    *iterations += 1;
    if (tr_ok) {
      if (rho > rho1) { status = SCP_INACCURATE; Dn = sp[SP_BFAIL] * Delta; }                       // :125-128
      else {
        accept = 1;                                                                                 // :129-139
        if (rho < rho0) { const double g = sp[SP_BSUCC] * Delta; Dn = g < D0 ? g : D0; }
        if (!ineq_ok) { status = SCP_VIOLATES; wn = sp[SP_GFAIL] * w; } else status = SCP_OK;
      }
    } else { status = SCP_TR_VIOLATED; wn = sp[SP_GFAIL] * w; }                                     // :140-145
    const bool w_exceeded = wn > w_max;                                                             // :163-166
    conv_now = accept && !w_exceeded && *iterations > 2 && (conv + *conv_prev <= sp[SP_CONVTHR]);   // :167-174
    succ_now = conv_now && ineq_ok;
    done = w_exceeded || (conv_now && !force);
    *conv_prev = conv;
    *j_full = inf[4];
    if (accept) *j_true = e[4];
  }
  *delta = Dn; *omega = wn;
  if (conv_now) *converged = 1;
  if (succ_now) *successful = 1;
  rec[H_JTRUE] = *j_true; rec[H_JFULL] = *j_full; rec[H_SCP_STATUS] = status; rec[H_SOLVER_STATUS] = sstat;
  rec[H_ACCEPT] = accept; rec[H_CONV] = *conv_prev; rec[H_DELTA] = Dn; rec[H_OMEGA] = wn;
  rec[H_RHO] = (ok && tr_ok) ? rho : nan(""); rec[H_TR_OK] = tr_ok; rec[H_INEQ_OK] = ineq_ok; rec[H_NEWTON] = inf[1];
  return accept | (done ? 2 : 0) | 4 | (ok ? 8 : 0);
}

#ifndef GUSTO_HOSTSIM
// solve_gusto_jump! :60-75 in two phases around the initial K1 + K4: phase 0 installs Delta0 / omega0 and makes every instance
// live (K4 reads them); phase 1 resets the per-instance state and writes history slot 0: iterations = 0,
// J_true[1] = cost_true(traj_init), rho_vec[2] = ratio(traj_init, traj_init)  (ev0 = K4 on candidate == trajectory).
__global__ void scp_begin_kernel(BatchPtrs p, ScpState s, const double* __restrict__ ev0, int B, int nev, const double* __restrict__ sp, int phase) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (phase == 0) {
    if (b < B) { p.delta[b] = sp[SP_DELTA0]; p.omega[b] = sp[SP_OMEGA0]; s.active[b] = 1; }
    return;
  }
  if (b == 0) *s.slot = 0;
  if (b < s.max_hist * CNT_W) s.cnt[b] = 0;
  if (b >= B) return;
  const double j0 = ev0[(size_t)b * nev + 4];
  s.iterations[b] = 0; s.conv_prev[b] = 0.0; s.j_true[b] = j0; s.j_full[b] = j0;
  s.converged[b] = 0; s.successful[b] = 0; s.done[b] = 0; s.done[B + b] = 0;
  double* rec = s.hist + (size_t)b * HIST_W;
  rec[H_JTRUE] = j0; rec[H_JFULL] = j0; rec[H_SCP_STATUS] = SCP_NA; rec[H_SOLVER_STATUS] = -1; rec[H_ACCEPT] = 1; rec[H_CONV] = 0;
  rec[H_DELTA] = sp[SP_DELTA0]; rec[H_OMEGA] = sp[SP_OMEGA0]; rec[H_RHO] = ev0[(size_t)b * nev + 3]; rec[H_TR_OK] = 0; rec[H_INEQ_OK] = 0;
  rec[H_NEWTON] = 0;
}

__global__ void scp_update_kernel(BatchPtrs p, ScpState s, const double* __restrict__ ev, const double* __restrict__ info,
                                  int B, int N, int nx, int nu, int nev, int ninfo, const double* __restrict__ sp) {
  const int b = blockIdx.x;
  __shared__ int sh_accept;
  if (threadIdx.x == 0) {
    const int it = *s.slot;
    const int h = it + 1 <= s.max_hist ? it + 1 : s.max_hist;
    const int ci = it < s.max_hist ? it : s.max_hist - 1;
    double* rec = s.hist + ((size_t)h * B + b) * HIST_W;
    const bool was_active = s.active[b] != 0;
    const int fl = scp_update_instance(ev + (size_t)b * nev, info + (size_t)b * ninfo, sp, s.force != 0, was_active, p.delta + b, p.omega + b,
                                       s.iterations + b, s.conv_prev + b, s.j_true + b, s.j_full + b, s.converged + b, s.successful + b, rec);
    if (fl & 2) s.active[b] = 0;
    s.done[(size_t)(it & 1) * B + b] = (fl & 2) ? 1 : 0;
    if (fl & 4) {
      atomicAdd(s.cnt + ci * CNT_W + C_RAN, 1);
      if (fl & 8) atomicAdd(s.cnt + ci * CNT_W + C_SOLVED, 1);
      if (fl & 1) atomicAdd(s.cnt + ci * CNT_W + C_ACCEPTED, 1);
      if (!(fl & 2)) atomicAdd(s.cnt + ci * CNT_W + C_ACTIVE_AFTER, 1);
    }
    sh_accept = fl & 1;
  }
  __syncthreads();
  if (sh_accept) {
    const size_t ox = (size_t)b * N * nx, ou = (size_t)b * N * nu;
    for (int i = threadIdx.x; i < N * nx; i += blockDim.x) p.Xp[ox + i] = p.Xn[ox + i];
    for (int i = threadIdx.x; i < N * nu; i += blockDim.x) p.Up[ou + i] = p.Un[ou + i];
  }
}
__global__ void scp_advance_kernel(ScpState s) { *s.slot += 1; }
// number of unfinished instances over the gathered status bytes of all ranks
__global__ void scp_count_kernel(const uint8_t* __restrict__ done_all, int n, int* out) {
  __shared__ int acc;
  if (threadIdx.x == 0) acc = 0;
  __syncthreads();
  int v = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += done_all[i] ? 0 : 1;
  atomicAdd(&acc, v);
  __syncthreads();
  if (threadIdx.x == 0) *out = acc;
}
#endif

}  // namespace gusto
