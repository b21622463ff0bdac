// K7: indirect shooting refinement, one WARP per instance.
//
// Reference:
//   solve!(SS, SP)                 shooting.jl:4-49      Newton on the initial costate:  F(p0) = x_goal - x(tf; x_init, p0) = 0
//   parameterized_shooting_eval!   shooting.jl:51-66
//   shooting_ode!, get_control     dynamics/dubins_car.jl:259-280, dynamics/astrobee_se3_manifold.jl:831-895
//                                  (defined for DubinsCar and AstrobeeSE3Manifold only; the obstacle / bound terms of the
//                                  manifold ODE are commented out in the reference, :880-885)
//   p0 = SCPS.dual                 types.jl:219-226, get_dual_jump (dubins_car.jl:254-257): minus the JuMP dual of the
//                                  init constraints = the multiplier nu of row 0 of the IPM (ipm.cuh writes it to p.dual)
// The reference delegates the arithmetic to DifferentialEquations.jl (adaptive Tsit5) and NLsolve.jl (trust region,
// finite-difference Jacobian, ftol 1e-3, 100 iterations); neither is vendored.  Here: classical RK4 with `nsub` equal
// sub-steps per knot interval, and Levenberg-Marquardt (dx(tf)/dp0 is singular for the quaternion model) on the EXACT
// Jacobian of the discrete flow: lane j < n_x integrates the trajectory in forward-mode dual numbers seeded with e_j, so
// one Jacobian costs ONE integration time.  (A finite-difference Jacobian with NLsolve's absolute step cbrt(eps) is
// useless on the astrobee model: dx(tf)/dp0 ~ 1e5 over 70 s while the attitude costates are ~1e-6.)  The (<= 13 x 13)
// normal equations are solved by lane 0 from shared memory.  oracle/gusto_oracle/shooting.py is the same algorithm in
// NumPy with a complex-step Jacobian.
#pragma once
#include "common.cuh"
#include "models.cuh"

namespace gusto {

constexpr int SHOOT_NOUT = 8;      // status (0 Optimal, 1 Diverged), LM iterations, |F|_inf, J_true, convergence measure, lambda, 0, 0
constexpr int SHOOT_MAX_TRY = 12;  // damping escalations per LM iteration

// forward-mode dual number (value, tangent)
struct Dual {
  double v, d;
  GHD Dual() : v(0.0), d(0.0) {}
  GHD Dual(double a) : v(a), d(0.0) {}
  GHD Dual(double a, double b) : v(a), d(b) {}
};
GHD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
GHD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
GHD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
GHD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.v * b.d + a.d * b.v); }
GHD Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
GHD Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
GHD Dual operator/(Dual a, double b) { return Dual(a.v / b, a.d / b); }
GHD Dual g_sin(Dual a) { return Dual(::sin(a.v), ::cos(a.v) * a.d); }
GHD Dual g_cos(Dual a) { return Dual(::cos(a.v), -::sin(a.v) * a.d); }
GHD double g_sin(double a) { return ::sin(a); }
GHD double g_cos(double a) { return ::cos(a); }

template <int M> struct ShootModel { static constexpr bool DEFINED = (M == DUBINS || M == ASTROBEE_SE3_MANIFOLD); };

template <int M, typename S> GDEV void shoot_control(const double* rp, const S* pc, S* u) {
  if constexpr (M == DUBINS) {
    u[0] = 0.5 * rp[RP_DUB_K] * pc[2];
  } else if constexpr (M == ASTROBEE_SE3_MANIFOLD) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { u[i] = pc[3 + i] / (2.0 * rp[RP_MASS]); u[3 + i] = pc[10 + i] / (2.0 * rp[RP_JXX + i]); }
  }
}

// d/dt [x; p]
template <int M, typename S> GDEV void shoot_ode(const double* rp, const S* y, S* d) {
  using T = Traits<M>;
  constexpr int n = T::NX;
  const S* x = y;
  const S* pc = y + n;
  S u[T::NU > 0 ? T::NU : 1];
  shoot_control<M, S>(rp, pc, u);
#pragma unroll
  for (int i = 0; i < 2 * n; ++i) d[i] = S(0.0);
  if constexpr (M == DUBINS) {
    const double v = rp[RP_DUB_V], k = rp[RP_DUB_K];
    const S s = g_sin(x[2]), c = g_cos(x[2]);
    d[0] = v * c; d[1] = v * s; d[2] = k * u[0];
    d[5] = pc[0] * v * s - pc[1] * v * c;
  } else if constexpr (M == ASTROBEE_SE3_MANIFOLD) {
    const double mass = rp[RP_MASS];
    const S qw = x[6], qx = x[7], qy = x[8], qz = x[9], wx = x[10], wy = x[11], wz = x[12];
    const S pqw = pc[6], pqx = pc[7], pqy = pc[8], pqz = pc[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) { d[i] = x[3 + i]; d[3 + i] = u[i] / mass; }
    d[6] = 0.5 * (-wx * qx - wy * qy - wz * qz);
    d[7] = 0.5 * (wx * qw - wz * qy + wy * qz);
    d[8] = 0.5 * (wy * qw + wz * qx - wx * qz);
    d[9] = 0.5 * (wz * qw - wy * qx + wx * qy);
    const double Jx = rp[RP_JXX], Jy = rp[RP_JYY], Jz = rp[RP_JZZ];
    const S jwx = Jx * wx, jwy = Jy * wy, jwz = Jz * wz;
    d[10] = (u[3] - (wy * jwz - wz * jwy)) / Jx;
    d[11] = (u[4] - (wz * jwx - wx * jwz)) / Jy;
    d[12] = (u[5] - (wx * jwy - wy * jwx)) / Jz;
#pragma unroll
    for (int i = 0; i < 3; ++i) d[n + 3 + i] = -pc[i];
    d[n + 6] = -0.5 * (pqx * wx + pqy * wy + pqz * wz);
    d[n + 7] = -0.5 * (-pqw * wx + pqy * wz - pqz * wy);
    d[n + 8] = -0.5 * (-pqw * wy - pqx * wz + pqz * wx);
    d[n + 9] = -0.5 * (-pqw * wz + pqx * wy - pqy * wx);
    d[n + 10] = -0.5 * (-pqw * qx + pqx * qw - pqy * qz + pqz * qy);
    d[n + 11] = -0.5 * (-pqw * qy + pqx * qz + pqy * qw - pqz * qx);
    d[n + 12] = -0.5 * (-pqw * qz - pqx * qy + pqy * qx + pqz * qw);
  }
}

// one classical RK4 step, in place (accumulating form: y, acc, stage point, slope)
template <int M, typename S> GDEV void shoot_rk4(const double* rp, S* y, double h) {
  constexpr int n2 = 2 * Traits<M>::NX;
  S k[n2], acc[n2], yt[n2];
  shoot_ode<M, S>(rp, y, k);
#pragma unroll
  for (int i = 0; i < n2; ++i) { acc[i] = y[i] + (h / 6.0) * k[i]; yt[i] = y[i] + (0.5 * h) * k[i]; }
  shoot_ode<M, S>(rp, yt, k);
#pragma unroll
  for (int i = 0; i < n2; ++i) { acc[i] = acc[i] + (h / 3.0) * k[i]; yt[i] = y[i] + (0.5 * h) * k[i]; }
  shoot_ode<M, S>(rp, yt, k);
#pragma unroll
  for (int i = 0; i < n2; ++i) { acc[i] = acc[i] + (h / 3.0) * k[i]; yt[i] = y[i] + h * k[i]; }
  shoot_ode<M, S>(rp, yt, k);
#pragma unroll
  for (int i = 0; i < n2; ++i) y[i] = acc[i] + (h / 6.0) * k[i];
}

// x(tf) from (x_init, p0): (N-1) * nsub steps
template <int M> GDEV void shoot_final_state(const double* rp, const double* x_init, const double* p0, double h, int nsteps, double* xf) {
  constexpr int n = Traits<M>::NX;
  double y[2 * n];
#pragma unroll
  for (int i = 0; i < n; ++i) { y[i] = x_init[i]; y[n + i] = p0[i]; }
  for (int s = 0; s < nsteps; ++s) shoot_rk4<M, double>(rp, y, h);
#pragma unroll
  for (int i = 0; i < n; ++i) xf[i] = y[i];
}
// column j of d x(tf) / d p0 (forward-mode tangent seeded with e_j)
template <int M> GDEV void shoot_final_tangent(const double* rp, const double* x_init, const double* p0, int j, double h, int nsteps, double* col) {
  constexpr int n = Traits<M>::NX;
  Dual y[2 * n];
#pragma unroll
  for (int i = 0; i < n; ++i) { y[i] = Dual(x_init[i]); y[n + i] = Dual(p0[i], i == j ? 1.0 : 0.0); }
  for (int s = 0; s < nsteps; ++s) shoot_rk4<M, Dual>(rp, y, h);
#pragma unroll
  for (int i = 0; i < n; ++i) col[i] = y[i].d;
}

// shared-memory workspace of one instance (doubles)
template <int M> struct ShootLayout {
  static constexpr int n = Traits<M>::NX;
  static constexpr int JX = 0;                        // [n][n]   d x_i(tf) / d p_j
  static constexpr int JTJ = JX + n * n;              // [n][n]
  static constexpr int AM = JTJ + n * n;              // [n][n]   damped normal matrix, eliminated in place
  static constexpr int JTF = AM + n * n;              // [n]
  static constexpr int RHS = JTF + n;                 // [n]
  static constexpr int DV = RHS + n;                  // [n]      step
  static constexpr int PC = DV + n;                   // [n]      current costate
  static constexpr int PT = PC + n;                   // [n]      trial costate
  static constexpr int FV = PT + n;                   // [n]      F(p)
  static constexpr int FT = FV + n;                   // [n]      F(trial)
  static constexpr int TOTAL = FT + n + 2;
};

// Xs / Us / Ps: the shooting trajectory of this instance ([N][n_x], [N][n_u], [N][n_x]); on entry Xs holds the previous
// shooting trajectory (SS.traj) against which the convergence measure is taken; it is only overwritten on success.
template <int M>
GDEV void shoot_instance(const BatchDesc& d, const BatchPtrs& p, int b, const double* p0_in, const double* x_goal, int nsub,
                         int max_iter, double ftol, double* sm, double* Xs, double* Us, double* Ps, double* out) {
  using T = Traits<M>;
  using SL = ShootLayout<M>;
  constexpr int n = T::NX, NU = T::NU;
  const int N = d.N;
  const double* rp = d.rp;
  const double* x_init = p.x_init + (size_t)b * n;
  const double h = p.tf[b] / (N - 1) / nsub;
  const int nsteps = (N - 1) * nsub;
  double* Jx = sm + SL::JX; double* JtJ = sm + SL::JTJ; double* Am = sm + SL::AM;
  double* JtF = sm + SL::JTF; double* rhs = sm + SL::RHS; double* dv = sm + SL::DV; double* pc = sm + SL::PC;
  double* pt = sm + SL::PT; double* Fv = sm + SL::FV; double* Ft = sm + SL::FT;
  for (int i = G_LANE; i < n; i += G_NLANE) pc[i] = p0_in[i];
  G_SYNCWARP();
  if (G_LANE == 0) {
    double xf[n];
    shoot_final_state<M>(rp, x_init, pc, h, nsteps, xf);
    for (int i = 0; i < n; ++i) Fv[i] = x_goal[i] - xf[i];
  }
  G_SYNCWARP();
  double lam = 1e-3, fn = 0.0;
  int it = 0, status = 1;
  while (true) {
    fn = 0.0;
    bool finite = true;
    for (int i = 0; i < n; ++i) { const double a = fabs(Fv[i]); fn = a > fn ? a : fn; if (!(a == a) || a > 1e300) finite = false; }
    if (!finite) { fn = 1e300; break; }
    if (fn <= ftol) { status = 0; break; }
    if (it >= max_iter) break;
    ++it;
    // exact Jacobian of the discrete flow: lane j integrates the tangent seeded with e_j
    for (int j = G_LANE; j < n; j += G_NLANE) {
      double pv[n], col[n];
#pragma unroll
      for (int i = 0; i < n; ++i) pv[i] = pc[i];
      shoot_final_tangent<M>(rp, x_init, pv, j, h, nsteps, col);
#pragma unroll
      for (int i = 0; i < n; ++i) Jx[i * n + j] = col[i];
    }
    G_SYNCWARP();
    for (int e = G_LANE; e < n * n; e += G_NLANE) {
      const int a = e / n, c2 = e - a * n;
      double s = 0.0;
      for (int i = 0; i < n; ++i) s += Jx[i * n + a] * Jx[i * n + c2];
      JtJ[e] = s;
    }
    for (int a = G_LANE; a < n; a += G_NLANE) {
      double s = 0.0;
      for (int i = 0; i < n; ++i) s += Jx[i * n + a] * Fv[i];
      JtF[a] = s;
    }
    double f2 = 0.0;
    for (int i = 0; i < n; ++i) f2 += Fv[i] * Fv[i];
    G_SYNCWARP();
    bool accepted = false;
    for (int tr = 0; tr < SHOOT_MAX_TRY; ++tr) {
      if (G_LANE == 0) {
        for (int a = 0; a < n; ++a) {
          for (int c2 = 0; c2 < n; ++c2) Am[a * n + c2] = JtJ[a * n + c2];
          Am[a * n + a] += lam * JtJ[a * n + a] + 1e-14;
          rhs[a] = JtF[a];
        }
        for (int q = 0; q < n; ++q) {                       // SPD: no pivoting
          const double ip = 1.0 / Am[q * n + q];
          for (int i = q + 1; i < n; ++i) {
            const double ml = Am[i * n + q] * ip;
            for (int c2 = q + 1; c2 < n; ++c2) Am[i * n + c2] -= ml * Am[q * n + c2];
            rhs[i] -= ml * rhs[q];
          }
        }
        for (int i = n - 1; i >= 0; --i) {
          double s = rhs[i];
          for (int c2 = i + 1; c2 < n; ++c2) s -= Am[i * n + c2] * dv[c2];
          dv[i] = s / Am[i * n + i];
        }
        double xf[n];
        for (int i = 0; i < n; ++i) pt[i] = pc[i] + dv[i];
        shoot_final_state<M>(rp, x_init, pt, h, nsteps, xf);
        double f2t = 0.0;
        for (int i = 0; i < n; ++i) { Ft[i] = x_goal[i] - xf[i]; f2t += Ft[i] * Ft[i]; }
        Ft[n] = f2t;
      }
      G_SYNCWARP();
      const double f2t = Ft[n];
      const bool ok = (f2t == f2t) && f2t < 1e300 && f2t < f2;
      G_SYNCWARP();                                          // everyone has read f2t before lane 0 may overwrite it
      if (ok) {
        for (int i = G_LANE; i < n; i += G_NLANE) { pc[i] = pt[i]; Fv[i] = Ft[i]; }
        lam = lam * 0.1 > 1e-12 ? lam * 0.1 : 1e-12;
        accepted = true;
        G_SYNCWARP();
        break;
      }
      lam *= 10.0;
    }
    if (!accepted) break;
  }
  // recover the trajectory (shooting.jl:26-43): only a converged attempt replaces SS.traj
  double J = nan(""), conv = nan("");
  if (status == 0) {
    J = 0.0;
    if (G_LANE == 0) {
      double y[2 * n], u[NU > 0 ? NU : 1];
      double num = 0.0, den = 0.0;
      for (int i = 0; i < n; ++i) { y[i] = x_init[i]; y[n + i] = pc[i]; }
      const double dt = p.tf[b] / (N - 1);
      for (int k = 0; k < N; ++k) {
        if (k > 0) for (int s = 0; s < nsub; ++s) shoot_rk4<M, double>(rp, y, h);
        shoot_control<M, double>(rp, y + n, u);
        double dn = 0.0, xn = 0.0, uu = 0.0;
        for (int i = 0; i < n; ++i) { dn += sq(y[i] - Xs[k * n + i]); xn += sq(y[i]); }
        for (int i = 0; i < NU; ++i) uu += sq(u[i]);
        dn = sqrt(dn); xn = sqrt(xn);
        num = dn > num ? dn : num; den = xn > den ? xn : den;
        J += ((k == 0 || k == N - 1) ? 0.5 * dt : dt) * uu;          // cost_true: sum_k dt/2 (|u_{k-1}|^2 + |u_k|^2)
        for (int i = 0; i < n; ++i) { Xs[k * n + i] = y[i]; Ps[k * n + i] = y[n + i]; }
        for (int i = 0; i < NU; ++i) Us[k * NU + i] = u[i];
      }
      conv = num / den;                                               // convergence_metric, traj_opt.jl:74-85
    }
  }
  if (G_LANE == 0) {
    out[0] = (double)status; out[1] = (double)it; out[2] = fn; out[3] = J; out[4] = conv; out[5] = lam; out[6] = 0.0; out[7] = 0.0;
  }
}

}  // namespace gusto
