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
  CM_AB3 = 11,        // s0 += dt*(23/12 k - 16/12 k1 + 5/12 k2)
  // not a stepper: s0 = Nh (set_c! through the fused 3-D engine's own forward transforms)
  CM_STORE = 12
};

struct CombineArgs {
  int mode;
  int filtered;  // apply filter on the final stage
  double dt;
  double lsrk_a, lsrk_b;
  int lsrk_last;
};

struct CombinePtrs {
  double2* s0;   // sol
  double2* s1;   // stage state / sol_1
  double2* s2;   // ETDRK4 sol_2
  double2* acc;  // RK4 accumulator / ETD N2+N3 / LSRK S2 / AB3 RHS_{-1}
  double2* n1;   // ETD N1 / AB3 RHS_{-2}
  const double *E, *E2, *zeta, *alpha, *beta, *gamma;  // ETDRK4 coefficient arrays (one member's worth)
};

// One element of the stage combine.  i = element index in the state arrays, ci = index in the (batch-shared) ETD
// coefficient arrays, Nh = transformed nonlinear term.  Performs all loads/stores of the FF stepper for this element
// and returns the state the NEXT calcN is evaluated at (sol_1 / sol_2 / the new sol).
enum { CMASK_RK4 = 1, CMASK_ETD = 2, CMASK_OTHER = 4, CMASK_ALL = 7 };
template <int MASK = CMASK_ALL>
__device__ __forceinline__ double2 combine_at(const CombinePtrs& P, const CombineArgs& A, const AxisTables& ax,
                                              size_t i, size_t ci, double kx, double ky, double kz, double2 Nh) {
  const double dt = A.dt;
  double f = 1.0;
  double2 next = make_double2(0.0, 0.0);
  switch (A.mode) {
    case CM_RK4_S1: if (MASK & CMASK_RK4) {
      double L = lin_op(ax, kx, ky, kz);
      double2 s0 = P.s0[i];
      double2 k = cadd(Nh, cmul_r(s0, L));
      P.acc[i] = cmul_r(k, 1.0 / 6.0);
      next = cadd(s0, cmul_r(k, dt / 2));
      P.s1[i] = next;
    } break;
    case CM_RK4_S2:
    case CM_RK4_S3: if (MASK & CMASK_RK4) {
      double L = lin_op(ax, kx, ky, kz);
      double2 ss = P.s1[i];
      double2 k = cadd(Nh, cmul_r(ss, L));
      P.acc[i] = cadd(P.acc[i], cmul_r(k, 1.0 / 3.0));
      double h = (A.mode == CM_RK4_S2) ? dt / 2 : dt;
      next = cadd(P.s0[i], cmul_r(k, h));
      P.s1[i] = next;
    } break;
    case CM_RK4_S4: if (MASK & CMASK_RK4) {
      double L = lin_op(ax, kx, ky, kz);
      double2 ss = P.s1[i];
      double2 k = cadd(Nh, cmul_r(ss, L));
      double2 sum = cadd(P.acc[i], cmul_r(k, 1.0 / 6.0));
      double2 r = cadd(P.s0[i], cmul_r(sum, dt));
      if (A.filtered) f = filter_val(ax, kx, ky, kz);
      next = cmul_r(r, f);
      P.s0[i] = next;
    } break;
    case CM_ETD_S1: if (MASK & CMASK_ETD) {
      double2 s0 = P.s0[i];
      P.n1[i] = Nh;
      next = cadd(cmul_r(s0, P.E2[ci]), cmul_r(Nh, P.zeta[ci]));
      P.s1[i] = next;
    } break;
    case CM_ETD_S2: if (MASK & CMASK_ETD) {
      P.acc[i] = Nh;
      next = cadd(cmul_r(P.s0[i], P.E2[ci]), cmul_r(Nh, P.zeta[ci]));
      P.s2[i] = next;
    } break;
    case CM_ETD_S3: if (MASK & CMASK_ETD) {
      P.acc[i] = cadd(P.acc[i], Nh);
      double2 n1 = P.n1[i];
      double2 t = make_double2(2 * Nh.x - n1.x, 2 * Nh.y - n1.y);
      next = cadd(cmul_r(P.s1[i], P.E2[ci]), cmul_r(t, P.zeta[ci]));
      P.s2[i] = next;
    } break;
    case CM_ETD_S4: if (MASK & CMASK_ETD) {
      double2 r = cmul_r(P.s0[i], P.E[ci]);
      r = cadd(r, cmul_r(P.n1[i], P.alpha[ci]));
      r = cadd(r, cmul_r(P.acc[i], 2 * P.beta[ci]));
      r = cadd(r, cmul_r(Nh, P.gamma[ci]));
      if (A.filtered) f = filter_val(ax, kx, ky, kz);
      next = cmul_r(r, f);
      P.s0[i] = next;
    } break;
    case CM_EULER: if (MASK & CMASK_OTHER) {
      double L = lin_op(ax, kx, ky, kz);
      double2 s0 = P.s0[i];
      double2 k = cadd(Nh, cmul_r(s0, L));
      double2 r = cadd(s0, cmul_r(k, dt));
      if (A.filtered) f = filter_val(ax, kx, ky, kz);
      next = cmul_r(r, f);
      P.s0[i] = next;
    } break;
    case CM_LSRK: if (MASK & CMASK_OTHER) {
      double L = lin_op(ax, kx, ky, kz);
      double2 s0 = P.s0[i];
      double2 k = cadd(Nh, cmul_r(s0, L));
      double2 S2 = cadd(cmul_r(P.acc[i], A.lsrk_a), cmul_r(k, dt));
      P.acc[i] = S2;
      double2 r = cadd(s0, cmul_r(S2, A.lsrk_b));
      if (A.lsrk_last && A.filtered) {
        f = filter_val(ax, kx, ky, kz);
        r = cmul_r(r, f);
      }
      next = r;
      P.s0[i] = next;
    } break;
    case CM_AB3_EULER:
    case CM_AB3: if (MASK & CMASK_OTHER) {
      double L = lin_op(ax, kx, ky, kz);
      double2 s0 = P.s0[i];
      double2 k = cadd(Nh, cmul_r(s0, L));
      double2 k1 = P.acc[i], k2 = P.n1[i];
      double2 inc = k;
      if (A.mode == CM_AB3) {
        inc.x = 23.0 / 12.0 * k.x - 16.0 / 12.0 * k1.x + 5.0 / 12.0 * k2.x;
        inc.y = 23.0 / 12.0 * k.y - 16.0 / 12.0 * k1.y + 5.0 / 12.0 * k2.y;
      }
      double2 r = cadd(s0, cmul_r(inc, dt));
      if (A.filtered) f = filter_val(ax, kx, ky, kz);
      next = cmul_r(r, f);
      P.s0[i] = next;
      P.n1[i] = k1;
      P.acc[i] = k;
    } break;
    case CM_STORE: if (MASK & CMASK_OTHER) {
      next = Nh;
      P.s0[i] = Nh;
    } break;
  }
  return next;
}

// which state array the NEXT calcN reads after a combine of this mode (host side)
inline int next_state_slot(int mode) {  // 0 = s0, 1 = s1, 2 = s2
  switch (mode) {
    case CM_RK4_S1: case CM_RK4_S2: case CM_RK4_S3: case CM_ETD_S1: return 1;
    case CM_ETD_S2: case CM_ETD_S3: return 2;
    default: return 0;
  }
}

// ---------------------------------------------------------------------------------------------------
// ETDRK4 coefficients (FF getetdcoeffs / getexpLs): 32-point contour mean around dt*L, evaluated on device.
// tlayout = 0: arrays indexed [kz][ky][kx] (canonical); 1: [kx][ky] (the fused 2-D engine's transposed layout);
// 2: [kx][ky_local][kz] (the fused 3-D engine's state layout; ny = local rows, yoff = global index of the first).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cx_mul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cx_div(double2 a, double2 b) {
  double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
__device__ __forceinline__ double2 cx_exp(double2 z) {
  double e = exp(z.x), s, c;
  sincos(z.y, &s, &c);
  return make_double2(e * c, e * s);
}

static __global__ void __launch_bounds__(256) k_etd_coeffs(double* E, double* E2, double* zeta, double* alpha,
                                                           double* beta, double* gamma, AxisTables ax, int64_t nkr,
                                                           int64_t ny, int64_t nz, double dt, int tlayout,
                                                           int64_t yoff = 0) {
  int64_t n = nkr * ny * nz;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t ix, iy, iz;
    if (tlayout == 2) {
      iz = i % nz;
      iy = yoff + (i / nz) % ny;
      ix = i / (nz * ny);
    } else if (tlayout) {
      iy = i % ny;
      ix = i / ny;
      iz = 0;
    } else {
      ix = i % nkr;
      iy = (i / nkr) % ny;
      iz = i / (nkr * ny);
    }
    double L = lin_op(ax, ax.kx[ix], ax.ky[iy], ax.kz[iz]);
    double Ldt = dt * L;
    double sz = 0, sa = 0, sb = 0, sg = 0;
    const int ncirc = 32;
    for (int j = 0; j < ncirc; ++j) {
      double s, c;
      sincospi(2.0 * (j + 0.5) / ncirc, &s, &c);
      double2 z = make_double2(Ldt + c, s);
      double2 ez = cx_exp(z);
      double2 ez2 = cx_exp(make_double2(z.x / 2, z.y / 2));
      double2 z2 = cx_mul(z, z);
      double2 z3 = cx_mul(z2, z);
      sz += cx_div(make_double2(ez2.x - 1.0, ez2.y), z).x;                                   // (e^{z/2}-1)/z
      double2 t = cx_mul(ez, make_double2(4.0 - 3.0 * z.x + z2.x, -3.0 * z.y + z2.y));
      sa += cx_div(make_double2(-4.0 - z.x + t.x, -z.y + t.y), z3).x;                        // alpha
      t = cx_mul(ez, make_double2(-2.0 + z.x, z.y));
      sb += cx_div(make_double2(2.0 + z.x + t.x, z.y + t.y), z3).x;                          // beta
      t = cx_mul(ez, make_double2(4.0 - z.x, -z.y));
      sg += cx_div(make_double2(-4.0 - 3.0 * z.x - z2.x + t.x, -3.0 * z.y - z2.y + t.y), z3).x;  // Gamma
    }
    E[i] = exp(Ldt);
    E2[i] = exp(Ldt / 2);
    zeta[i] = dt * (sz / ncirc);
    alpha[i] = dt * (sa / ncirc);
    beta[i] = dt * (sb / ncirc);
    gamma[i] = dt * (sg / ncirc);
  }
}

}  // namespace ptf
