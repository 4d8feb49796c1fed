// Device-side building blocks shared by all engines: the diagonal linear operator L (TAD.jl:502-566), the
// FourierFlows exponential filter (FF utils.jl makefilter), the dealias box mask (FF domains.jl), and the
// time-stepper stage combines (FF timesteppers.jl) written as per-element register functions so that every
// engine can fuse them into whatever kernel holds the element.
#pragma once
#include "ptf_internal.h"

namespace ptf {

__device__ __forceinline__ double ipow_d(double x, int n) {
  double r = 1.0;  // x^0 == 1, also 0^0 (Julia semantics)
  for (int i = 0; i < n; ++i) r *= x;
  return r;
}

// L = -kappa*kr^2 - eta*l^2 - iota*m^2 - kappa_h*Krsq^n   (same operation order as TAD.jl:515,524)
__device__ __forceinline__ double lin_op(const AxisTables& ax, double kx, double ky, double kz) {
  double kx2 = kx * kx;
  double L = -ax.kappa * kx2;
  double Ksq = kx2;
  if (ax.ndim >= 2) {
    double ky2 = ky * ky;
    L = L - ax.eta * ky2;
    Ksq = Ksq + ky2;
  }
  if (ax.ndim >= 3) {
    double kz2 = kz * kz;
    L = L - ax.iota * kz2;
    Ksq = Ksq + kz2;
  }
  L = L - ax.kappa_h * ipow_d(Ksq, ax.n_kappa_h);
  return L;
}

// filter = 1 for K < innerK, exp(-decay*(K-innerK)^order) otherwise;  K = sqrt(sum (k_a d_a/pi)^2)
__device__ __forceinline__ double filter_val(const AxisTables& ax, double kx, double ky, double kz) {
  double a = kx * ax.fx;
  double Ksq = a * a;
  if (ax.ndim >= 2) {
    double b = ky * ax.fy;
    Ksq += b * b;
  }
  if (ax.ndim >= 3) {
    double c = kz * ax.fz;
    Ksq += c * c;
  }
  double K = sqrt(Ksq);
  if (K < ax.f_inner) return 1.0;
  return exp(-ax.f_decay * pow(K - ax.f_inner, ax.f_order));
}

__device__ __forceinline__ bool dealiased_out(const AxisTables& ax, int64_t ix, int64_t iy, int64_t iz) {
  if (!ax.dealias) return false;
  if (ix >= ax.ax_lo) return true;
  if (ax.ndim >= 2 && iy >= ax.ay_lo && iy < ax.ay_hi) return true;
  if (ax.ndim >= 3 && iz >= ax.az_lo && iz < ax.az_hi) return true;
  return false;
}

__device__ __forceinline__ double2 cmul_r(double2 a, double r) { return make_double2(a.x * r, a.y * r); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cdiv_r(double2 a, double r) { return make_double2(a.x / r, a.y / r); }

// ---------------------------------------------------------------------------------------------------
// Stage combines.  `Nh` is the freshly transformed nonlinear term of this stage, `ss` the stage state it was
// evaluated at.  Each function returns what must be stored; which arrays they live in is the engine's business.
// ---------------------------------------------------------------------------------------------------
enum CombineMode : int {
  // RK4 (FF RK4substeps!/RK4update!):  k = Nh + L*ss
  CM_RK4_S1 = 0,  // acc = k/6           ; next = s0 + dt/2*k      (ss == s0)
  CM_RK4_S2 = 1,  // acc += k/3          ; next = s0 + dt/2*k
  CM_RK4_S3 = 2,  // acc += k/3          ; next = s0 + dt*k
  CM_RK4_S4 = 3,  // s0 = [filter*](s0 + dt*(acc + k/6))
  // ETDRK4 (FF ETDRK4substeps!/ETDRK4update!)
  CM_ETD_S1 = 4,  // N1 = Nh ; s1 = E2*s0 + zeta*N1
  CM_ETD_S2 = 5,  // acc = Nh(N2) ; s2 = E2*s0 + zeta*N2
  CM_ETD_S3 = 6,  // acc += Nh(N3) ; s2 = E2*s1 + zeta*(2*N3 - N1)
  CM_ETD_S4 = 7,  // s0 = [filter*](E*s0 + alpha*N1 + 2*beta*acc + Gamma*N4)
  // ForwardEuler
  CM_EULER = 8,  // s0 = [filter*](s0 + dt*k)
  // LSRK54 (stage i): S2 = A_i*S2 + dt*k ; s0 += B_i*S2 ; filter after stage 5
  CM_LSRK = 9,
  // AB3
  CM_AB3_EULER = 10,  // start-up: s0 += dt*k ; rotate history
  CM_AB3 = 11         // s0 += dt*(23/12 k - 16/12 k1 + 5/12 k2)
};

struct CombineArgs {
  int mode;
  int filtered;  // apply filter on the final stage
  double dt;
  double lsrk_a, lsrk_b;
  int lsrk_last;
};

}  // namespace ptf
