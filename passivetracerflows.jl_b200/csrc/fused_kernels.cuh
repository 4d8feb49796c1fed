#pragma once
// Kernels of the fused 2-D engine (included by engine_fused.cu and by fused_inst.cu, which is compiled once per
// transform length so that the sizes build in parallel).
//
// Fused 2-D engine: the whole RK4/ETDRK4/... stage as TWO hand-written kernels, no cuFFT on the step path.
//
//   k_fused_y (one CTA sub-group per kr column, all y/l in registers+smem):
//       gather P^x(y,kr) -> forward FFT_y -> N^(kr,l)                      (second half of rfft,  TAD.jl:766)
//       stage combine in registers: addlinearterm! + substepsol!/update!   (FF timesteppers.jl), L/filter on the fly
//       next stage state s' -> A = IFFT_y(s'/N), B = IFFT_y(i*l*s'/N)      (first half of the 2 irffts, TAD.jl:757-761)
//   k_fused_x (one CTA sub-group per pair of y rows):
//       gather A,B(kr,y) -> Z = i*kr*A + i*B (two-for-one Hermitian packing) -> inverse FFT_x -> gx + i*gy
//       p = -u*gx - v*gy                                                    (TAD.jl:764)
//       rows y,y+1 packed as p_y + i*p_{y+1} -> forward FFT_x -> split -> P^x(y,kr)   (first half of rfft)
//
// Data layout in HBM: spectral state (sol, sol_1, acc, ...) and A,B are stored TRANSPOSED, [b][kr][l|y]
// (the column kernel's natural order); P^x is [b][y][kr] (the row kernel's natural order).  Every kernel WRITES
// contiguously and READS the other kernel's layout with 16-byte strided gathers: measured on B200
// (profiles/r01_microbench_strided_bw.txt) strided reads keep 75-95 % of copy bandwidth, strided partial-sector
// writes only 25-47 %.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "fft_core.cuh"
#include "ptf_pointwise.cuh"
#include "ptf_velocity.cuh"
#include "tmem_park.cuh"

namespace ptf {

namespace {

using fft::Cfg;
using fft::out_slot;
using fft::pad_idx;
using fft::Twiddles;

struct YArgs {
  const double2* Px;          // [b][y][kr]     nonlinear term, x-transformed (HAS_IN)
  double2* A;                 // [b][kr][y]     IFFT_y(s')           (HAS_OUT)
  double2* Bf;                // [b][kr][y]     IFFT_y(i*l*s')       (HAS_OUT)
  const double2* next_state;  // [b][kr][l]     array holding s' after the combine (s0 for the prologue)
  CombinePtrs P;
  CombineArgs C;
  AxisTables ax;
  Twiddles tw;
  int nkr;                    // number of columns per member (2-D: nx/2+1;  3-D: (nx/2+1) * nyl)
  double inv_n;               // 1/(nx*ny[*nz]): normalisation of ldiv!(., rfftplan, .)
  const double2* pf_c[4];     // complex state columns the combine will read: L2-prefetched at CTA start
  const double* pf_r[4];      // real coefficient columns (ETDRK4), same
  int pf_ahead;               // > 0: also prefetch the P^x gather of the column `pf_ahead` CTAs ahead
  int ablate;                 // experiment bitmask (timing only): 1 = contiguous instead of gathered P^x
  int stagger;                // first-wave de-phasing: odd CTAs of the first wave start `stagger` cycles late
  int first_wave;             // number of CTAs resident at launch (2 per SM)
  // ---- 3-D (template parameter D3): this kernel is the z-column kernel, one column per (kr, local ky row) ----
  int nkx = 0;                // nx/2+1
  int nyl = 1, yoff = 0;      // local ky rows of this rank's spectral slab and the global index of the first one
  int nzl = 0, zsh = 0;       // planes per rank (nz / P, a power of two) and log2 of it
  int cid0 = 0, cid_end = 0;  // this launch covers columns [cid0, cid_end): a kr chunk of the pipelined slab exchange
  int grid_cap = 0;           // host side only: > 0 = launch at most this many (persistent, grid-striding) CTAs
  // Where the blocks of the slab exchange live, per rank (P <= 16).  NCCL / single GPU: slices of this rank's own
  // send / receive buffers.  P2P mode: pointers into the PEERS' memory (CUDA IPC over NVLink) — the kernel gathers
  // rank r's P^xy block straight from r's send buffer and stores A, C straight into rank p's receive buffers, so the
  // all-to-all is fused into this kernel's loads and stores.
  const double2* Psrc[16];    // [r]: block (r -> this rank) of P^xy, [kr][zl/8][ll][zl%8]
  double2* Adst[16];          // [p]: block (this rank -> p) of A, [kr][ll][zl]
  double2* Cdst[16];          // [p]: same for C
};

enum { FAM_RK4 = 0, FAM_ETD = 1, FAM_OTHER = 2 };

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// FourierFlows filter, out of line: sqrt/exp/pow would otherwise be inlined once per register element
// (a, b, c = k_axis * d_axis / pi per axis; c = 0 in 2-D)
__device__ __noinline__ double filter_slow(double a, double b, double c, double f_inner, double f_decay,
                                           double f_order) {
  double Ksq = a * a;
  Ksq += b * b;
  Ksq += c * c;
  if (Ksq < f_inner * f_inner) return 1.0;   // (K < f_inner for the non-negative K: no square root for the untouched modes)
  double K = sqrt(Ksq);
  if (K < f_inner) return 1.0;
  const double d = K - f_inner;
  if (f_order == 4.0) {   // FourierFlows' default order: two multiplies instead of the general pow()
    const double d2 = d * d;
    return exp(-f_decay * (d2 * d2));
  }
  return exp(-f_decay * pow(d, f_order));
}

// One column of the column kernel: 2-D = (kr; transform axis y), 3-D = (kr, ky row lp; transform axis z).
// ka = wavenumber along the transform axis.
template <bool D3>
__device__ __forceinline__ double col_lin(const AxisTables& ax, double kx, double kyp, double ka) {
  return D3 ? lin_op(ax, kx, kyp, ka) : lin_op(ax, kx, ka, 0.0);
}
template <bool D3>
__device__ __forceinline__ bool col_dealiased(const AxisTables& ax, int kr, int lp, int l) {
  return D3 ? dealiased_out(ax, kr, lp, l) : dealiased_out(ax, kr, l, 0);
}
struct ColId {
  int kr, lp;        // kr index; 3-D: GLOBAL ky index of this column (0 in 2-D)
  double kx, kyp;    // their wavenumbers
  double cax;        // generating constant of the transform axis' wavenumbers (fft::wavenumber_full)
  int nyq;           // its Nyquist sign switch
};

// ---- RK4 family (FF RK4substeps!/RK4update!).  N^ (t_nh) and the new stage state s' (t_w) live in TMEM, so a
// ---- batch of NB elements can have all 3*NB of its 16-byte state loads in flight at once (2 memory round trips
// ---- per column instead of 4) and nothing accumulates in registers.
// FILT: the final stage applies the FourierFlows filter (compiled out otherwise: `fl` is a run-time indexed array in
// local memory, and an unfiltered RK4 step must not pay 16 local loads per thread for it)
template <int NY, int MODE, bool DM, bool FILT, bool D3>
__device__ __forceinline__ void rk4_stage(const YArgs& a, uint32_t t_nh, uint32_t t_w, size_t col, int t,
                                          const ColId& c, bool active, const double (&fl)[16]) {
#ifndef PTF_RK4_NB
#define PTF_RK4_NB 4
#endif
  constexpr int T = Cfg<NY>::T, NB = PTF_RK4_NB;
  const double dt = a.C.dt;
  const double kx = c.kx;
#pragma unroll
  for (int e0 = 0; e0 < 16; e0 += NB) {
    double2 ss[NB], s0[NB], ac[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const size_t i = col + t + T * (e0 + j);
      s0[j] = __ldcg(a.P.s0 + i);
      // opt-in dealias!(sol) at the top of calcN: sol is read as if it had been masked in place at stage 1
      if (DM && col_dealiased<D3>(a.ax, c.kr, c.lp, t + T * (e0 + j))) s0[j] = make_double2(0.0, 0.0);
      if (MODE != CM_RK4_S1) {
        ss[j] = __ldcg(a.P.s1 + i);
        ac[j] = __ldcg(a.P.acc + i);
      }
    }
#pragma unroll
    for (int j0 = 0; j0 < NB; j0 += 4) {
      double2 nh[4];
      tmem::ldn<4>(t_nh + 4 * (e0 + j0), nh);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = j0 + jj, e = e0 + j, l = t + T * e;
        const size_t i = col + l;
        const double L = col_lin<D3>(a.ax, kx, c.kyp, fft::wavenumber_full<NY>(t, e, c.cax, c.nyq));
        const double2 Nh = nh[jj];
        const bool dm = DM && col_dealiased<D3>(a.ax, c.kr, c.lp, l);
        double2 next;
        if (MODE == CM_RK4_S1) {
          double2 k = cadd(Nh, cmul_r(s0[j], L));
          next = cadd(s0[j], cmul_r(k, dt / 2));
          if (dm) next = make_double2(0.0, 0.0);   // the next calcN masks its stage state in place
          if (active) {
            __stcg(a.P.acc + i, cmul_r(k, 1.0 / 6.0));
            __stcg(a.P.s1 + i, next);
          }
        } else if (MODE == CM_RK4_S2 || MODE == CM_RK4_S3) {
          double2 k = cadd(Nh, cmul_r(ss[j], L));
          next = cadd(s0[j], cmul_r(k, MODE == CM_RK4_S2 ? dt / 2 : dt));
          if (dm) next = make_double2(0.0, 0.0);
          if (active) {
            __stcg(a.P.acc + i, cadd(ac[j], cmul_r(k, 1.0 / 3.0)));
            __stcg(a.P.s1 + i, next);
          }
        } else {  // CM_RK4_S4
          double2 k = cadd(Nh, cmul_r(ss[j], L));
          double2 sum = cadd(ac[j], cmul_r(k, 1.0 / 6.0));
          next = cadd(s0[j], cmul_r(sum, dt));
          if (FILT) next = cmul_r(next, fl[e]);
          if (active) __stcg(a.P.s0 + i, next);  // stored unmasked (as the reference leaves sol after the update) ...
          if (dm) next = make_double2(0.0, 0.0); // ... but the next step's first calcN sees it masked
        }
        tmem::st1(t_w + 4 * e, make_double2(next.x * a.inv_n, next.y * a.inv_n));
      }
    }
  }
  tmem::wait_st();
}

// ---- ETDRK4 family (FF ETDRK4substeps!/ETDRK4update!)
template <int NY, int MODE, bool DM, bool FILT, bool D3>
__device__ __forceinline__ void etd_stage(const YArgs& a, const double2 (&v)[16], double2 (&w)[16], size_t col,
                                          size_t ccol, int t, const ColId& c, bool active, const double (&fl)[16]) {
  constexpr int T = Cfg<NY>::T, NB = 4;
#pragma unroll
  for (int e0 = 0; e0 < 16; e0 += NB) {
    double2 sa[NB], n1[NB], ac[NB];
    double c0[NB], c1[NB], c2[NB], c3[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const size_t i = col + t + T * (e0 + j), ci = ccol + t + T * (e0 + j);
      const bool dmj = DM && col_dealiased<D3>(a.ax, c.kr, c.lp, t + T * (e0 + j));
      if (MODE == CM_ETD_S1 || MODE == CM_ETD_S2) {
        sa[j] = __ldcg(a.P.s0 + i);
        if (dmj) sa[j] = make_double2(0.0, 0.0);
        c0[j] = __ldcg(a.P.E2 + ci);
        c1[j] = __ldcg(a.P.zeta + ci);
      } else if (MODE == CM_ETD_S3) {
        sa[j] = __ldcg(a.P.s1 + i);
        n1[j] = __ldcg(a.P.n1 + i);
        ac[j] = __ldcg(a.P.acc + i);
        c0[j] = __ldcg(a.P.E2 + ci);
        c1[j] = __ldcg(a.P.zeta + ci);
      } else {
        sa[j] = __ldcg(a.P.s0 + i);
        if (dmj) sa[j] = make_double2(0.0, 0.0);
        n1[j] = __ldcg(a.P.n1 + i);
        ac[j] = __ldcg(a.P.acc + i);
        c0[j] = __ldcg(a.P.E + ci);
        c1[j] = __ldcg(a.P.alpha + ci);
        c2[j] = __ldcg(a.P.beta + ci);
        c3[j] = __ldcg(a.P.gamma + ci);
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int e = e0 + j, l = t + T * e;
      const size_t i = col + l;
      const double2 Nh = v[out_slot<NY>(e)];
      const bool dm = DM && col_dealiased<D3>(a.ax, c.kr, c.lp, l);
      double2 next;
      if (MODE == CM_ETD_S1) {
        next = cadd(cmul_r(sa[j], c0[j]), cmul_r(Nh, c1[j]));
        if (dm) next = make_double2(0.0, 0.0);
        if (active) {
          __stcg(a.P.n1 + i, Nh);
          __stcg(a.P.s1 + i, next);
        }
      } else if (MODE == CM_ETD_S2) {
        next = cadd(cmul_r(sa[j], c0[j]), cmul_r(Nh, c1[j]));
        if (dm) next = make_double2(0.0, 0.0);
        if (active) {
          __stcg(a.P.acc + i, Nh);
          __stcg(a.P.s2 + i, next);
        }
      } else if (MODE == CM_ETD_S3) {
        double2 tt = make_double2(2 * Nh.x - n1[j].x, 2 * Nh.y - n1[j].y);
        next = cadd(cmul_r(sa[j], c0[j]), cmul_r(tt, c1[j]));
        if (dm) next = make_double2(0.0, 0.0);
        if (active) {
          __stcg(a.P.acc + i, cadd(ac[j], Nh));
          __stcg(a.P.s2 + i, next);
        }
      } else {
        double2 r = cmul_r(sa[j], c0[j]);
        r = cadd(r, cmul_r(n1[j], c1[j]));
        r = cadd(r, cmul_r(ac[j], 2 * c2[j]));
        r = cadd(r, cmul_r(Nh, c3[j]));
        if (FILT) r = cmul_r(r, fl[e]);
        next = r;
        if (active) __stcg(a.P.s0 + i, next);
        if (dm) next = make_double2(0.0, 0.0);
      }
      w[e] = make_double2(next.x * a.inv_n, next.y * a.inv_n);
    }
  }
}

// NT = threads per CTA (64, 128 or 256; >= T).  Small problems use small CTAs so that enough CTAs exist to fill the
// 148 SMs; TMEM is allocated 128 columns per warp-quarter, i.e. 128 columns per CTA up to 4 warps, 256 for 8 warps.
// DM = the opt-in dealias!(sol) mask is compiled in (kept out of the default kernels: it costs registers)
// D3 = z-column kernel of the fused 3-D engine (engine_fused3d.cu): column = (kr, local ky row), transform axis z,
//      input P^xy gathered from [r][kr][zl/8][ll][zl%8] (r = z / nzl: the block received from rank r), outputs
//      A = IFFT_z(s'), C = IFFT_z(i*m*s') written as [p][kr][ll][zl] (p = z / nzl: the block sent to rank p).
// One column of the column kernel (all threads of the CTA call it together: the transforms contain CTA barriers).
template <int NY, int FAM, bool HAS_IN, bool HAS_OUT, int NT, bool DM, bool D3>
__device__ __forceinline__ void fused_y_column(const YArgs& a, const int cid_raw, const int grp, const int t,
                                               double2* __restrict__ smem, const uint32_t t_nh, const uint32_t t_w) {
  constexpr int T = Cfg<NY>::T, F = NT / T, PADN = Cfg<NY>::PADN;
  constexpr bool USE_TMEM = HAS_IN && FAM == FAM_RK4;
  (void)F;
  const bool active = cid_raw < (D3 ? a.cid_end : a.nkr);
  const int cid = active ? cid_raw : 0;
  const int b = blockIdx.y;
  double2* sm = smem + grp * PADN;
  const size_t col = ((size_t)b * a.nkr + cid) * NY;  // this column in the transposed state arrays
  const size_t ccol = (size_t)cid * NY;               // ... in the batch-shared coefficient arrays
  ColId c;
  int ll = 0;                                         // 3-D: local ky row
  if (D3) {
    c.kr = cid / a.nyl;
    ll = cid - c.kr * a.nyl;
    c.lp = a.yoff + ll;
    c.kyp = a.ax.ky[c.lp];
    c.cax = a.ax.cz;
  } else {
    c.kr = cid;
    c.lp = 0;
    c.kyp = 0.0;
    c.cax = a.ax.cy;
  }
  c.nyq = a.ax.nyq_sign;
  c.kx = (double)c.kr * a.ax.cx;   // = a.ax.kx[c.kr], without the load
  const double kx = c.kx;
  double2 w[16];
  double fl[16];  // FourierFlows filter of this thread's 16 modes (final stage of Filtered* steppers only)
  if (HAS_IN && FAM != FAM_OTHER && a.C.filtered && (a.C.mode == CM_RK4_S4 || a.C.mode == CM_ETD_S4)) {
#pragma unroll 1
    for (int e = 0; e < 16; ++e) {
      const double ka = fft::wavenumber_full<NY>(t, e, c.cax, c.nyq);
      fl[e] = D3 ? filter_slow(kx * a.ax.fx, c.kyp * a.ax.fy, ka * a.ax.fz, a.ax.f_inner, a.ax.f_decay, a.ax.f_order)
                 : filter_slow(kx * a.ax.fx, ka * a.ax.fy, 0.0, a.ax.f_inner, a.ax.f_decay, a.ax.f_order);
    }
  }

  if (HAS_IN) {
    double2 v[16];
    if (D3) {
      // P^xy block of rank r = z / nzl: [kr][zl/8][ll][zl%8]; 8 consecutive z of one column are one 128-byte line
      const size_t pin = ((size_t)c.kr * (a.nzl >> 3) * a.nyl + ll) * 8;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int z = t + T * e, r = z >> a.zsh, zl = z & (a.nzl - 1);
        v[e] = __ldcg(a.Psrc[r] + pin + (size_t)(zl >> 3) * a.nyl * 8 + (zl & 7));
      }
    } else {
      // P^x is stored blocked, [b][y/8][kr][y%8]: 8 consecutive y of one column are one 128-byte line, so this
      // column read is fully coalesced (the row kernel pays with 32-byte full-sector stores, which do not stall it)
      const double2* P = a.Px + (size_t)b * NY * a.nkr + (size_t)cid * 8;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int y = t + T * e;
        v[e] = __ldcg((a.ablate & 1) ? (a.Px + col + y) : (P + (size_t)(y >> 3) * a.nkr * 8 + (y & 7)));  // L2-only
      }
    }
    // Pull the state columns the combine needs into L2 while the gather + forward FFT run (fire and forget).
    if ((t & 7) == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (a.pf_c[q]) {
#pragma unroll
          for (int e = 0; e < 16; ++e) prefetch_l2(a.pf_c[q] + col + t + T * e);
        }
    }
    if (FAM == FAM_ETD && (t & 15) == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (a.pf_r[q]) {
#pragma unroll
          for (int e = 0; e < 16; ++e) prefetch_l2(a.pf_r[q] + ccol + t + T * e);
        }
    }
    fft::fft_cta<NY, -1, false, NT>(v, sm, t, a.tw, grp);
    if (!D3 && a.pf_ahead > 0) {  // next wave's gather: one 32-byte sector per (y, kr') pair
      const int kr2 = cid_raw + a.pf_ahead * F;
      if (kr2 < a.nkr) {
        const double2* P2 = a.Px + (size_t)b * NY * a.nkr + (size_t)kr2 * 8;
#pragma unroll
        for (int e = 0; e < 16; ++e) prefetch_l2(P2 + (size_t)((t + T * e) >> 3) * a.nkr * 8 + ((t + T * e) & 7));
      }
    }
    if (FAM == FAM_RK4) {
#pragma unroll
      for (int e = 0; e < 16; ++e) tmem::st1(t_nh + 4 * e, v[out_slot<NY>(e)]);
      tmem::wait_st();
      switch (a.C.mode) {
        case CM_RK4_S1: rk4_stage<NY, CM_RK4_S1, DM, false, D3>(a, t_nh, t_w, col, t, c, active, fl); break;
        case CM_RK4_S2: rk4_stage<NY, CM_RK4_S2, DM, false, D3>(a, t_nh, t_w, col, t, c, active, fl); break;
        case CM_RK4_S3: rk4_stage<NY, CM_RK4_S3, DM, false, D3>(a, t_nh, t_w, col, t, c, active, fl); break;
        default:
          if (a.C.filtered) rk4_stage<NY, CM_RK4_S4, DM, true, D3>(a, t_nh, t_w, col, t, c, active, fl);
          else rk4_stage<NY, CM_RK4_S4, DM, false, D3>(a, t_nh, t_w, col, t, c, active, fl);
          break;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double2 r4[4];
        tmem::ldn<4>(t_w + 16 * q, r4);
#pragma unroll
        for (int j = 0; j < 4; ++j) w[4 * q + j] = r4[j];
      }
    } else if (FAM == FAM_ETD) {
      switch (a.C.mode) {
        case CM_ETD_S1: etd_stage<NY, CM_ETD_S1, DM, false, D3>(a, v, w, col, ccol, t, c, active, fl); break;
        case CM_ETD_S2: etd_stage<NY, CM_ETD_S2, DM, false, D3>(a, v, w, col, ccol, t, c, active, fl); break;
        case CM_ETD_S3: etd_stage<NY, CM_ETD_S3, DM, false, D3>(a, v, w, col, ccol, t, c, active, fl); break;
        default:
          if (a.C.filtered) etd_stage<NY, CM_ETD_S4, DM, true, D3>(a, v, w, col, ccol, t, c, active, fl);
          else etd_stage<NY, CM_ETD_S4, DM, false, D3>(a, v, w, col, ccol, t, c, active, fl);
          break;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int l = t + T * e;
        double2 nx = make_double2(0.0, 0.0);
        const bool dm = DM && col_dealiased<D3>(a.ax, c.kr, c.lp, l);
        if (active) {
          // ForwardEuler / LSRK54 / AB3 evaluate calcN at sol itself: dealias!(sol) acts in place before the combine
          if (dm && a.C.mode != CM_STORE) a.P.s0[col + l] = make_double2(0.0, 0.0);
          const double ka = fft::wavenumber_full<NY>(t, e, c.cax, c.nyq);
          nx = combine_at<CMASK_OTHER>(a.P, a.C, a.ax, col + l, ccol + l, kx, D3 ? c.kyp : ka, D3 ? ka : 0.0,
                                       v[out_slot<NY>(e)]);
        }
        if (dm) nx = make_double2(0.0, 0.0);
        w[e] = make_double2(nx.x * a.inv_n, nx.y * a.inv_n);
      }
    }
  } else {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      double2 s = __ldcg(a.next_state + col + t + T * e);
      if (DM && col_dealiased<D3>(a.ax, c.kr, c.lp, t + T * e)) s = make_double2(0.0, 0.0);
      w[e] = make_double2(s.x * a.inv_n, s.y * a.inv_n);
    }
  }
  if (!HAS_OUT) return;

  // where element l of this column goes in A / B: contiguous in 2-D; 3-D: the block of the rank p that owns plane l
  auto out_ptr = [&](double2* base2d, double2* const* dst3d, int l) -> double2* {
    if (D3) return dst3d[l >> a.zsh] + ((size_t)cid << a.zsh) + (l & (a.nzl - 1));
    return base2d + col + l;
  };
  fft::fft_cta<NY, +1, false, NT>(w, sm, t, a.tw, grp);
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e) __stcg(out_ptr(a.A, a.Adst, t + T * e), w[out_slot<NY>(e)]);
  }
  // derivative along the transform axis: i*l*s'  (3-D: i*m*s')
  if (USE_TMEM) {  // s'/N is still parked in TMEM: no second trip to global memory
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double2 r4[4];
      tmem::ldn<4>(t_w + 16 * q, r4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double ky = fft::wavenumber_full<NY>(t, 4 * q + j, c.cax, c.nyq);
        w[4 * q + j] = make_double2(-ky * r4[j].y, ky * r4[j].x);
      }
    }
  } else {         // s' re-read from the array this thread just wrote: L2 hit, no DRAM traffic
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int l = t + T * e;
      double2 s = __ldcg(a.next_state + col + l);
      if (DM && col_dealiased<D3>(a.ax, c.kr, c.lp, l)) s = make_double2(0.0, 0.0);
      double ky = fft::wavenumber_full<NY>(t, e, c.cax, c.nyq) * a.inv_n;
      w[e] = make_double2(-ky * s.y, ky * s.x);
    }
  }
  fft::fft_cta<NY, +1, false, NT>(w, sm, t, a.tw, grp);
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e) __stcg(out_ptr(a.Bf, a.Cdst, t + T * e), w[out_slot<NY>(e)]);
  }
}

// PERSIST (3-D only) = grid-stride over the columns of the launch: the slab pipeline launches a SMALL persistent grid, so
//      that this (then NVLink-bound) kernel leaves most CTA slots of every SM to the HBM-bound kernels running next to
//      it on the main stream.  Kept out of the default kernels: the loop costs them registers (spills).
template <int NY, int FAM, bool HAS_IN, bool HAS_OUT, int NT, bool DM = false, bool D3 = false, bool PERSIST = false>
__global__ void __launch_bounds__(NT, 512 / NT) k_fused_y(YArgs a) {
  constexpr int T = Cfg<NY>::T, F = NT / T, PADN = Cfg<NY>::PADN;
  constexpr int TCOLS = NT > 128 ? 256 : 128;
  static_assert(NT >= T && NT % T == 0, "CTA must hold whole transforms");
  constexpr bool USE_TMEM = HAS_IN && FAM == FAM_RK4;   // N^ and s' parked in TMEM (see rk4_stage)
  extern __shared__ double2 smem[];
  // De-phase the two CTAs resident on each SM: without this every first-wave CTA starts at the same instant and the
  // whole chip alternates between memory phases (FP64 idle) and FFT phases (HBM idle) in lock-step.
  if (a.stagger > 0 && (int)(blockIdx.x + gridDim.x * blockIdx.y) < a.first_wave && (blockIdx.x & 1)) {
    const long long t0 = clock64();
    while (clock64() - t0 < a.stagger) {
    }
  }

  __shared__ uint32_t tslot;
  uint32_t tbase = 0, t_nh = 0, t_w = 0;
  if (USE_TMEM) {
    tbase = tmem::alloc_cta<TCOLS>(&tslot);
    t_nh = tmem::warp_addr(tbase, 128);
    t_w = t_nh + 64;
  }
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  if (D3 && PERSIST) {
    for (int base = blockIdx.x * F + a.cid0; base < a.cid_end; base += gridDim.x * F) {
      fused_y_column<NY, FAM, HAS_IN, HAS_OUT, NT, DM, D3>(a, base + grp, grp, t, smem, t_nh, t_w);
      if (USE_TMEM) asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();   // the transform buffers and the TMEM parking slots are reused by the next column
      if (USE_TMEM) asm volatile("tcgen05.fence::after_thread_sync;");
    }
  } else {
    fused_y_column<NY, FAM, HAS_IN, HAS_OUT, NT, DM, D3>(a, blockIdx.x * F + grp + (D3 ? a.cid0 : 0), grp, t, smem, t_nh,
                                                         t_w);
  }
  if (USE_TMEM) tmem::free_cta<TCOLS>(tbase);
}

struct XArgs {
  const double2* A;   // [b][kr][y]
  const double2* Bf;  // [b][kr][y]
  double2* Px;        // [b][y][kr]
  VelArgs va;
  AxisTables ax;
  Twiddles tw;
  int nkr, ny;
  int pf_vel;    // 1: L2-prefetch this pair's u,v rows at CTA start
  int pf_ahead;  // > 0: L2-prefetch the A,B gather of the row pair `pf_ahead` CTAs ahead
  int ablate;    // experiment bitmask (timing only): 2 = no u,v loads
  int stagger, first_wave;  // see YArgs
  // ---- 3-D (template parameter D3): b = local z plane; third gradient gz = c2r(C) ----
  const double2* Cf = nullptr;  // [zl][kr][y]   IFFT_y IFFT_z (i*m*s')
  int nzg = 1, zoff = 0;        // global nz (stride of separable z tables) and this rank's first plane
};

// ---- pieces of k_fused_x (free functions so that every register-array index is a compile-time constant) ----
// Z = X + iY with X = i*kr*A (-> gx), Y = B (-> gy) for the lower half k = t + T*e < NX/2; the mirrored bin NX-k
// gets conj(X) + i*conj(Y) and is handed to its owner through shared memory.
template <int NX>
__device__ __forceinline__ void x_lower(double2& ve, int e, int t, double2 Av, double2 Bv, double cx,
                                        double2* __restrict__ sm) {
  constexpr int T = Cfg<NX>::T;
  const int k = t + T * e;
  const double kx = fft::wavenumber_half<NX>(t, e, cx);   // = the kr table entry, without the load
  const double Xx = -kx * Av.y, Xy = kx * Av.x;
  if (e == 0 && t == 0) {
    ve = make_double2(Xx, Bv.x);  // c2r ignores the imaginary part of the DC bin
  } else {
    ve = make_double2(Xx - Bv.y, Xy + Bv.x);
    sm[pad_idx(NX - k)] = make_double2(Xx + Bv.y, Bv.x - Xy);
  }
}
// same packing for two plain half-spectra X = C0 (-> gz of row 0), Y = C1 (-> gz of row 1)
template <int NX>
__device__ __forceinline__ void x_lower_plain(double2& ve, int e, int t, double2 Xv, double2 Yv,
                                              double2* __restrict__ sm) {
  constexpr int T = Cfg<NX>::T;
  const int k = t + T * e;
  if (e == 0 && t == 0) {
    ve = make_double2(Xv.x, Yv.x);
  } else {
    ve = make_double2(Xv.x - Yv.y, Xv.y + Yv.x);
    sm[pad_idx(NX - k)] = make_double2(Xv.x + Yv.y, Yv.x - Xv.y);
  }
}
// bin k = NX/2: real part only (c2r semantics)
template <int NX>
__device__ __forceinline__ void x_nyquist(int t, const double2* Ab, const double2* Bb, int ny, int q, double cx,
                                          double2* __restrict__ sm) {
  constexpr int H = NX / 2;
  if (t == 0) {
    double2 Av = __ldcg(Ab + (size_t)H * ny + q);
    double2 Bv = __ldcg(Bb + (size_t)H * ny + q);
    sm[pad_idx(H)] = make_double2(-((double)H * cx) * Av.y, Bv.x);
  }
}
// velocity row -> TMEM (its latency hides behind the inverse transform that follows)
// D3W: the pair (w of row 0, w of row 1) instead of (u, v) of row q
template <int NX, int VMODE, bool D3W = false, bool D3 = false>
__device__ __forceinline__ void x_request_uv(const XArgs& a, size_t voff, int q, int t, uint32_t t_uv) {
  constexpr int T = Cfg<NX>::T;
  constexpr int UB = 8;  // velocity request batch
  if (VMODE >= 3) {
    // direct mode: only pull the rows into L2 now (one request per 32-byte sector); x_product loads them from there
    // after the transform — no TMEM parking, and this warp does not wait here for HBM
    if ((t & 3) == 0) {
      const size_t i0 = voff + (size_t)q * NX + t;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (D3W) {
          prefetch_l2(a.va.arr[2] + i0 + T * e);
          prefetch_l2(a.va.arr[2] + i0 + NX + T * e);
        } else {
          prefetch_l2(a.va.arr[0] + i0 + T * e);
          prefetch_l2(a.va.arr[1] + i0 + T * e);
        }
      }
    }
  } else if (VMODE != 2) {
    // 3-D: one opaque base per field and compile-time offsets T*e, so that no per-element address outlives this call
    // (hoisted out of the row loop they would cost 64 registers)
    const size_t i0 = D3 ? fft::opaque(voff + (size_t)q * NX + t) : 0;
#pragma unroll
    for (int h = 0; h < 16 / UB; ++h) {
      double2 uv[UB];
#pragma unroll
      for (int j = 0; j < UB; ++j) {
        const size_t i = (D3 ? i0 : voff + (size_t)q * NX + t) + T * (UB * h + j);
        if (a.ablate & 2) uv[j] = make_double2(0.5, 0.25);
        else if (D3W) uv[j] = make_double2(tmem::ldg64(a.va.arr[2] + i), tmem::ldg64(a.va.arr[2] + i + NX));
        else uv[j] = make_double2(tmem::ldg64(a.va.arr[0] + i), tmem::ldg64(a.va.arr[1] + i));
      }
#pragma unroll
      for (int j = 0; j < UB; ++j) tmem::st1(t_uv + 4 * (UB * h + j), uv[j]);
    }
  }
  tmem::wait_st();
}
// physical space: v = gx + i*gy at x = t + T*e;  p = -u*gx - v*gy  (TAD.jl:764)
// D3: both rows' partial products are parked in shared memory (ps[Q*NX + x]); the w*gz term follows in x_product_z
template <int NX, int VMODE, int Q, bool D3>
__device__ __forceinline__ void x_product(const XArgs& a, const double2 (&v)[16], double2 (&w)[16],
                                          double* __restrict__ ps, int t, uint32_t t_uv, int b, int pair) {
  constexpr int T = Cfg<NX>::T;
  const int row = 2 * pair + Q;
  constexpr int PB = VMODE >= 3 ? 8 : 4;  // velocity fetch batch
  const size_t vi0 = VMODE >= 3 ? fft::opaque((D3 ? (size_t)b * a.ny * NX : (size_t)b * a.va.member_stride) +
                                              (size_t)row * NX + t)
                                : 0;
#pragma unroll
  for (int c = 0; c < 16 / PB; ++c) {
    double2 uv[PB];
    if (VMODE >= 3) {
#pragma unroll
      for (int j = 0; j < PB; ++j)
        uv[j] = make_double2(tmem::ldg64(a.va.arr[0] + vi0 + T * (PB * c + j)),
                             tmem::ldg64(a.va.arr[1] + vi0 + T * (PB * c + j)));
    } else if (VMODE != 2) tmem::ldn<PB>(t_uv + 4 * PB * c, uv);
    double pprev = 0.0;
    double2 p0[PB / 2];   // VMODE 4 (2-D, row 1): row 0's parked products of this batch, two points per entry
    if (!D3 && VMODE == 4 && Q == 1) tmem::ldn<PB / 2>(t_uv + 2 * PB * c, p0);
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int e = PB * c + j, x = t + T * e;
      const double2 g = v[out_slot<NX>(e)];
      double u, vv;
      if (VMODE == 2) {
        u = sep_eval(a.va.sep[0], x, row, D3 ? a.zoff + b : 0, NX, a.ny, a.nzg, D3 ? 3 : 2);
        vv = sep_eval(a.va.sep[1], x, row, D3 ? a.zoff + b : 0, NX, a.ny, a.nzg, D3 ? 3 : 2);
      } else {
        u = uv[j].x;
        vv = uv[j].y;
        if (!D3 && a.va.ushift) u += a.va.ushift[b * a.ny + row];  // layered flows: u + U(y, layer)  (TAD.jl:795)
      }
      const double p = -u * g.x - vv * g.y;
      if (D3 && VMODE == 3) {                    // 3-D direct mode: parked in the (free) velocity slot of TMEM, two
        if (j & 1) tmem::st1(t_uv + 32 * Q + 2 * (e - 1), make_double2(pprev, p));   // points per 4-column entry
        pprev = p;
      } else if (D3) ps[Q * NX + x] = p;
      else if (VMODE == 4 && Q == 0) {           // 2-D direct mode 4: row 0's product parked in TMEM as well
        if (j & 1) tmem::st1(t_uv + 2 * (e - 1), make_double2(pprev, p));
        pprev = p;
      } else if (VMODE == 4) w[e] = make_double2((j & 1) ? p0[j / 2].y : p0[j / 2].x, p);
      else if (Q == 0) ps[x] = p;                // row 0: parked in shared memory while row 1 runs
      else w[e] = make_double2(ps[x], p);        // row 1: packed with row 0 as p_y + i*p_{y+1} for the forward FFT
    }
  }
  if ((D3 && VMODE == 3) || (!D3 && VMODE == 4 && Q == 0)) tmem::wait_st();
}
// 3-D: v = gz(row 0) + i*gz(row 1);  p_q -= w_q * gz_q  (TAD.jl:781), completed in the shared-memory parking rows
template <int NX, int VMODE>
__device__ __forceinline__ void x_product_z(const XArgs& a, const double2 (&v)[16], double* __restrict__ ps, int t,
                                            uint32_t t_uv, int b, int pair) {
  constexpr int T = Cfg<NX>::T;
  constexpr int PB = VMODE == 3 ? 8 : 4;
  const size_t vi0 = VMODE == 3 ? fft::opaque((size_t)b * a.ny * NX + (size_t)(2 * pair) * NX + t) : 0;
#pragma unroll
  for (int c = 0; c < 16 / PB; ++c) {
    double2 ww[PB];
    if (VMODE == 3) {
#pragma unroll
      for (int j = 0; j < PB; ++j)
        ww[j] = make_double2(tmem::ldg64(a.va.arr[2] + vi0 + T * (PB * c + j)),
                             tmem::ldg64(a.va.arr[2] + vi0 + NX + T * (PB * c + j)));
    } else if (VMODE != 2) tmem::ldn<PB>(t_uv + 4 * PB * c, ww);
    double2 pa[PB / 2], pb[PB / 2];   // direct mode: the parked partial products of rows 0 and 1
    if (VMODE == 3) {
      tmem::ldn<PB / 2>(t_uv + 2 * PB * c, pa);
      tmem::ldn<PB / 2>(t_uv + 32 + 2 * PB * c, pb);
    }
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int e = PB * c + j, x = t + T * e;
      const double2 g = v[out_slot<NX>(e)];
      double w0, w1;
      if (VMODE == 2) {
        w0 = sep_eval(a.va.sep[2], x, 2 * pair, a.zoff + b, NX, a.ny, a.nzg, 3);
        w1 = sep_eval(a.va.sep[2], x, 2 * pair + 1, a.zoff + b, NX, a.ny, a.nzg, 3);
      } else {
        w0 = ww[j].x;
        w1 = ww[j].y;
      }
      if (VMODE == 3) {   // finished products replace the partial ones in place
        if (j & 1) {
          pa[j / 2].y = pa[j / 2].y - w0 * g.x;
          pb[j / 2].y = pb[j / 2].y - w1 * g.y;
        } else {
          pa[j / 2].x = pa[j / 2].x - w0 * g.x;
          pb[j / 2].x = pb[j / 2].x - w1 * g.y;
        }
      } else {
        ps[x] = ps[x] - w0 * g.x;
        ps[NX + x] = ps[NX + x] - w1 * g.y;
      }
    }
    if (VMODE == 3) {
#pragma unroll
      for (int k = 0; k < PB / 2; ++k) {
        tmem::st1(t_uv + 2 * PB * c + 4 * k, pa[k]);
        tmem::st1(t_uv + 32 + 2 * PB * c + 4 * k, pb[k]);
      }
    }
  }
  if (VMODE == 3) tmem::wait_st();
}

// VMODE 0: velocity arrays; 1: arrays + layered shift U(y,b); 2: separable tables (zero HBM bytes);
// 3: arrays, direct (rows prefetched to L2 before the inverse transform, loaded after it; 3-D default);
// 4: 3 + row 0's product parked in TMEM instead of shared memory (2-D default; launched without the parking row)
//
// Memory choreography (what the ablation study in profiles/ asked for): A[k][y], A[k][y+1] are 32 contiguous bytes,
// so ONE 256-bit load per k fetches both rows of the pair (half the L1 tag look-ups of two 16-byte gathers, one
// memory round trip instead of two); row 1's share is parked in TMEM until row 0 is done.  The velocity rows are
// requested before each inverse transform and parked in TMEM as well, so their latency hides behind the FFT.
// D3 = row kernel of the fused 3-D engine: b = local z plane, a third inverse transform carries gz of both rows.
template <int NX, int VMODE, int NT, bool D3 = false>
__global__ void __launch_bounds__(NT, 512 / NT) k_fused_x(XArgs a) {
  constexpr int T = Cfg<NX>::T, F = NT / T, PADN = Cfg<NX>::PADN, H = NX / 2;
  constexpr int TCOLS = NT > 128 ? 256 : 128;
  static_assert(NT >= T && NT % T == 0, "CTA must hold whole transforms");
  constexpr int GB = 4;  // gather batch (k's per batch)
  extern __shared__ double2 smem[];
  // De-phase the two CTAs resident on each SM: without this every first-wave CTA starts at the same instant and the
  // whole chip alternates between memory phases (FP64 idle) and FFT phases (HBM idle) in lock-step.
  if (a.stagger > 0 && (int)(blockIdx.x + gridDim.x * blockIdx.y) < a.first_wave && (blockIdx.x & 1)) {
    const long long t0 = clock64();
    while (clock64() - t0 < a.stagger) {
    }
  }

  __shared__ uint32_t tslot;
  const uint32_t tbase = tmem::alloc_cta<TCOLS>(&tslot);
  const uint32_t t_in = tmem::warp_addr(tbase, 128);  // 64 columns: row-1 A,B   | 64 columns: u,v of the current row
  const uint32_t t_uv = t_in + 64;
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  const int pair = blockIdx.x * F + grp;
  const int b = blockIdx.y;
  const int ny = a.ny;
  double2* sm = smem + grp * PADN;
  // row-0 product (3-D: both rows' partial products), parked while the other transforms run
  double* ps = reinterpret_cast<double*>(smem + F * PADN) + grp * (D3 ? 2 : 1) * NX;
  const double2* Ab = a.A + (size_t)b * a.nkr * ny + 2 * pair;
  const double2* Bb = a.Bf + (size_t)b * a.nkr * ny + 2 * pair;
  const size_t voff = (D3 ? (size_t)b * ny * NX : (size_t)b * a.va.member_stride) + (size_t)(2 * pair) * NX;
  double2 v[16];
  if (D3) {   // the third transform's inputs: pulled into L2 now, read after the first two transforms
    // (strides pass through fft::opaque so that these addresses are recomputed, not kept live across the transforms)
    const double2* Cb = a.Cf + (size_t)b * a.nkr * ny + 2 * pair + (size_t)t * ny;
    const size_t se = fft::opaque((size_t)T * ny);
#pragma unroll
    for (int e = 0; e < 8; ++e) prefetch_l2(Cb + e * se);
  }

  // row 0's inputs, and the parking of row 1's
#pragma unroll
  for (int h = 0; h < 8 / GB; ++h) {  // batches of GB k's: 2*GB 256-bit requests in flight per thread
    double2 A0[GB], A1[GB], B0[GB], B1[GB];
#pragma unroll
    for (int j = 0; j < GB; ++j) {
      const size_t gi = (size_t)(t + T * (GB * h + j)) * ny;
      tmem::ldg256(Ab + gi, A0[j], A1[j]);
      tmem::ldg256(Bb + gi, B0[j], B1[j]);
    }
#pragma unroll
    for (int j = 0; j < GB; ++j) {
      tmem::st1(t_in + 8 * (GB * h + j), A1[j]);
      tmem::st1(t_in + 8 * (GB * h + j) + 4, B1[j]);
    }
#pragma unroll
    for (int j = 0; j < GB; ++j) x_lower<NX>(v[GB * h + j], GB * h + j, t, A0[j], B0[j], a.ax.cx, sm);
  }
  constexpr int NQ = D3 ? 3 : 2;   // 3-D: a third pass of the same loop transforms (gz row 0, gz row 1)
#pragma unroll 1
  for (int q = 0; q < NQ; ++q) {  // deliberately NOT unrolled: one copy of the inverse transform keeps register
                                  // pressure (and the instruction footprint) down
    if (q == 1) {
      fft::group_sync<T, NT>(grp);  // exchange buffer free (row 0's transform readers are done)
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        double2 AB[4];  // (A1, B1) of k-slots 2h, 2h+1
        tmem::ldn<4>(t_in + 16 * h, AB);
        x_lower<NX>(v[2 * h], 2 * h, t, AB[0], AB[1], a.ax.cx, sm);
        x_lower<NX>(v[2 * h + 1], 2 * h + 1, t, AB[2], AB[3], a.ax.cx, sm);
      }
      if constexpr (D3) {        // the parking slot is free again: fetch the (C row 0, C row 1) pairs into it
        const double2* Cb = a.Cf + (size_t)b * a.nkr * ny + 2 * pair + (size_t)t * ny;
        const size_t se = fft::opaque((size_t)T * ny);
#pragma unroll
        for (int h = 0; h < 8 / GB; ++h) {
          double2 C0[GB], C1[GB];
#pragma unroll
          for (int j = 0; j < GB; ++j) tmem::ldg256(Cb + (GB * h + j) * se, C0[j], C1[j]);
#pragma unroll
          for (int j = 0; j < GB; ++j) {
            tmem::st1(t_in + 8 * (GB * h + j), C0[j]);
            tmem::st1(t_in + 8 * (GB * h + j) + 4, C1[j]);
          }
        }
      }
    }
    if constexpr (D3) {
      if (q == 2) {
        fft::group_sync<T, NT>(grp);  // exchange buffer free (row 1's transform readers are done)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          double2 CC[4];         // (C0, C1) of k-slots 2h, 2h+1
          tmem::ldn<4>(t_in + 16 * h, CC);
          x_lower_plain<NX>(v[2 * h], 2 * h, t, CC[0], CC[1], sm);
          x_lower_plain<NX>(v[2 * h + 1], 2 * h + 1, t, CC[2], CC[3], sm);
        }
        if (t == 0) {
          double2 C0, C1;
          tmem::ldg256(a.Cf + (size_t)b * a.nkr * ny + 2 * pair + (size_t)H * ny, C0, C1);
          sm[pad_idx(H)] = make_double2(C0.x, C1.x);
        }
        x_request_uv<NX, VMODE, true, true>(a, voff, 0, t, t_uv);
      } else {
        x_nyquist<NX>(t, Ab, Bb, ny, q, a.ax.cx, sm);
        x_request_uv<NX, VMODE, false, true>(a, voff, q, t, t_uv);
      }
    } else {
      x_nyquist<NX>(t, Ab, Bb, ny, q, a.ax.cx, sm);
      x_request_uv<NX, VMODE>(a, voff, q, t, t_uv);  // includes tcgen05.wait::st for the parked inputs as well
    }
    fft::group_sync<T, NT>(grp);
#pragma unroll
    for (int e = 8; e < 16; ++e) v[e] = sm[pad_idx(t + T * e)];
    fft::fft_cta<NX, +1, false, NT, D3>(v, sm, t, a.tw, grp);
    if constexpr (D3) {
      if (q == 0) x_product<NX, VMODE, 0, true>(a, v, v, ps, t, t_uv, b, pair);
      else if (q == 1) x_product<NX, VMODE, 1, true>(a, v, v, ps, t, t_uv, b, pair);
      else x_product_z<NX, VMODE>(a, v, ps, t, t_uv, b, pair);
    } else {
      if (q == 0) x_product<NX, VMODE, 0, false>(a, v, v, ps, t, t_uv, b, pair);
    }
  }
  double2 w[16];
  if constexpr (D3 && VMODE == 3) {   // both rows' finished products come back from TMEM, two points per entry
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double2 pa[4], pb[4];
      tmem::ldn<4>(t_uv + 16 * h, pa);
      tmem::ldn<4>(t_uv + 32 + 16 * h, pb);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        w[8 * h + 2 * k] = make_double2(pa[k].x, pb[k].x);
        w[8 * h + 2 * k + 1] = make_double2(pa[k].y, pb[k].y);
      }
    }
  } else if constexpr (D3) {     // both rows' finished products come back from shared memory (same thread wrote them)
#pragma unroll
    for (int e = 0; e < 16; ++e) w[e] = make_double2(ps[t + T * e], ps[NX + t + T * e]);
  } else {
    x_product<NX, VMODE, 1, false>(a, v, w, ps, t, t_uv, b, pair);
  }

  // ---------------- forward transform of the row pair packed as p_y + i*p_{y+1} ----------------
  fft::fft_cta<NX, -1, false, NT, D3>(w, sm, t, a.tw, grp);
  fft::group_sync<T, NT>(grp);
#pragma unroll
  for (int e = 8; e < 16; ++e) sm[pad_idx(t + T * e)] = w[out_slot<NX>(e)];
  fft::group_sync<T, NT>(grp);
  // blocked layout [b][y/8][kr][y%8]: rows y0 = 2*pair and y0+1 of one kr are 32 contiguous, 32-byte aligned bytes
  const int y0 = 2 * pair;
  double2* P0 = a.Px + (size_t)b * ny * a.nkr + (size_t)(y0 >> 3) * a.nkr * 8 + (y0 & 7);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = t + T * e;
    double2 W = w[out_slot<NX>(e)];
    double2 Wm = (e == 0 && t == 0) ? W : sm[pad_idx(NX - k)];
    tmem::stg256(P0 + (size_t)k * 8, make_double2(0.5 * (W.x + Wm.x), 0.5 * (W.y - Wm.y)),
                 make_double2(0.5 * (W.y + Wm.y), 0.5 * (Wm.x - W.x)));
  }
  if (t == 0) {
    double2 W = w[out_slot<NX>(8)];  // index 8*T = NX/2
    tmem::stg256(P0 + (size_t)H * 8, make_double2(W.x, 0.0), make_double2(W.y, 0.0));
  }
  tmem::free_cta<TCOLS>(tbase);
}

// out[b][c][r] = scale * in[b][r][c]   (layout changes at the set/get boundary only — not on the step path)
__global__ void __launch_bounds__(256) k_transpose(const double2* __restrict__ in, double2* __restrict__ out, int R,
                                                   int Cc, double scale) {
  __shared__ double2 tile[32][33];
  const int b = blockIdx.z;
  const double2* I = in + (size_t)b * R * Cc;
  double2* O = out + (size_t)b * R * Cc;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[j][threadIdx.x] = I[(size_t)r * Cc + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < Cc) {
      double2 v = tile[threadIdx.x][j];
      O[(size_t)c * R + r] = make_double2(v.x * scale, v.y * scale);
    }
  }
}

__global__ void __launch_bounds__(256) k_replicate_f(double* __restrict__ c, int64_t npts, int64_t B) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += (int64_t)gridDim.x * blockDim.x) {
    double v = c[i];
    for (int64_t b = 1; b < B; ++b) c[b * npts + i] = v;
  }
}

__global__ void __launch_bounds__(256) k_diag_t(const double2* __restrict__ s, int64_t nkr, int64_t ny, int64_t B,
                                                int64_t nx, double* out) {
  __shared__ double ssum[256];
  __shared__ double smax[256];
  int64_t n = nkr * ny * B;
  double acc = 0, mx = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t ix = (i / ny) % nkr;
    double2 v = s[i];
    double a2 = v.x * v.x + v.y * v.y;
    acc += ((ix == 0 || ix == nx / 2) ? 1.0 : 2.0) * a2;
    mx = fmax(mx, a2);
  }
  ssum[threadIdx.x] = acc;
  smax[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      ssum[threadIdx.x] += ssum[threadIdx.x + o];
      smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(&out[0], ssum[0]);
    atomicMax(reinterpret_cast<unsigned long long*>(&out[1]), (unsigned long long)__double_as_longlong(smax[0]));
  }
}

// ---- stand-alone transform test kernel (ptf_selftest_fft): `count` independent length-N transforms ----
template <int N, int DIR>
__global__ void __launch_bounds__(256, 2) k_fft_test(const double2* __restrict__ in, double2* __restrict__ out,
                                                     int count, Twiddles tw) {
  constexpr int T = Cfg<N>::T, F = 256 / T, PADN = Cfg<N>::PADN;
  extern __shared__ double2 smem[];
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  const int id = blockIdx.x * F + grp;
  const bool active = id < count;
  const size_t base = (size_t)(active ? id : 0) * N;
  double2 v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = in[base + t + T * e];
  fft::fft_cta<N, DIR>(v, smem + grp * PADN, t, tw);
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e) out[base + t + T * e] = v[out_slot<N>(e)];
  }
}

bool is_fused_size(int64_t n) {
  return n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096;
}

// twiddle tables of one length, forward sign, rounded from long double
struct TwiddleSet {
  DevBuf<double2> tw2, tw3;
  Twiddles dev() const { return Twiddles{tw2.p, tw3.p}; }
  void build(int N, int64_t* tally) {
    std::vector<double2> h2(4 * 16), h3(4 * 256);
    const long double PI2 = 6.283185307179586476925286766559005768L;
    const int n2 = N < 256 ? N : 256;  // two-pass lengths (64, 128) keep their last-pass twiddles w_N^(2^m k) here
    for (int m = 0; m < 4; ++m)
      for (int k = 0; k < 16; ++k) {
        long double ang = -PI2 * (long double)((k << m) % n2) / (long double)n2;
        h2[m * 16 + k] = make_double2((double)cosl(ang), (double)sinl(ang));
      }
    for (int m = 0; m < 4; ++m)
      for (int k = 0; k < 256; ++k) {
        long double ang = -PI2 * (long double)(((long)k << m) % N) / (long double)N;
        h3[m * 256 + k] = make_double2((double)cosl(ang), (double)sinl(ang));
      }
    tw2.alloc(h2.size(), tally);
    tw3.alloc(h3.size(), tally);
    PTF_CUDA(cudaMemcpy(tw2.p, h2.data(), h2.size() * sizeof(double2), cudaMemcpyHostToDevice));
    PTF_CUDA(cudaMemcpy(tw3.p, h3.data(), h3.size() * sizeof(double2), cudaMemcpyHostToDevice));
  }
};

// experiment knob (PTF_SMEM_PAD): extra dynamic smem to lower occupancy
inline size_t smem_pad_knob() {
  static const size_t v = [] {
    const char* e = std::getenv("PTF_SMEM_PAD");
    return e ? (size_t)std::atoi(e) : (size_t)0;
  }();
  return v;
}
#define g_smem_pad (smem_pad_knob())
template <int N, int NT = 256>
constexpr size_t y_smem() { return (size_t)(NT / Cfg<N>::T) * Cfg<N>::PADN * sizeof(double2); }
template <int N, int NT = 256>
constexpr size_t x_smem() { return y_smem<N, NT>() + (size_t)(NT / Cfg<N>::T) * N * sizeof(double); }

template <class K>
void allow_smem(K kernel, size_t bytes) {
  PTF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// CTA size for a launch of `items` transform groups (measured on B200, profiles/r01_fft_core_experiments.md §7):
// four independent 128-thread CTAs per SM beat two 256-thread CTAs (more independent phase streams), so 128 threads
// is the default whenever a transform fits (N <= 2048); tiny launches drop to 64 threads to create enough CTAs to
// cover the 148 SMs; 4096-point transforms need 256 threads.
template <int N>
int pick_nt(long items, int n_sm) {
  constexpr int T = Cfg<N>::T;
  const char* fe = std::getenv("PTF_NT");  // experiment knob: force the CTA size (clamped to >= T)
  const int forced = fe ? std::atoi(fe) : 0;
  if (forced == 64 || forced == 128 || forced == 256) return forced < T ? T : forced;
  if (T > 128) return 256;
  const long ctas128 = (items + 128 / T - 1) / (128 / T);
  if (ctas128 >= 2L * n_sm || T > 64) return 128;
  return 64;
}

template <int NY, int NT>
void prep_y_nt() {  // opt in to > 48 KB dynamic shared memory (per device, so done at engine construction)
  if constexpr (NT >= Cfg<NY>::T) {
    allow_smem(k_fused_y<NY, FAM_RK4, false, true, NT, false>, y_smem<NY, NT>() + g_smem_pad);
    allow_smem(k_fused_y<NY, FAM_RK4, true, true, NT, false>, y_smem<NY, NT>() + g_smem_pad);
    allow_smem(k_fused_y<NY, FAM_ETD, true, true, NT, false>, y_smem<NY, NT>() + g_smem_pad);
    allow_smem(k_fused_y<NY, FAM_OTHER, true, true, NT, false>, y_smem<NY, NT>() + g_smem_pad);
    allow_smem(k_fused_y<NY, FAM_RK4, false, true, NT, true>, y_smem<NY, NT>() + g_smem_pad);
    allow_smem(k_fused_y<NY, FAM_RK4, true, true, NT, true>, y_smem<NY, NT>() + g_smem_pad);
    allow_smem(k_fused_y<NY, FAM_ETD, true, true, NT, true>, y_smem<NY, NT>() + g_smem_pad);
    allow_smem(k_fused_y<NY, FAM_OTHER, true, true, NT, true>, y_smem<NY, NT>() + g_smem_pad);
  }
}
template <int NY>
void prep_y() {
  prep_y_nt<NY, 256>();
  prep_y_nt<NY, 128>();
  prep_y_nt<NY, 64>();
}
template <int NX, int NT>
void prep_x_nt() {
  if constexpr (NT >= Cfg<NX>::T) {
    allow_smem(k_fused_x<NX, 0, NT>, x_smem<NX, NT>() + g_smem_pad);
    allow_smem(k_fused_x<NX, 2, NT>, x_smem<NX, NT>() + g_smem_pad);
    allow_smem(k_fused_x<NX, 3, NT>, x_smem<NX, NT>() + g_smem_pad);
    allow_smem(k_fused_x<NX, 4, NT>, y_smem<NX, NT>() + g_smem_pad);
  }
}
template <int NX>
void prep_x() {
  prep_x_nt<NX, 256>();
  prep_x_nt<NX, 128>();
  prep_x_nt<NX, 64>();
}

template <int NY, int NT>
void launch_y_nt(bool has_in, int fam, const YArgs& a, int nb, cudaStream_t st) {
  if constexpr (NT >= Cfg<NY>::T) {
    constexpr int F = NT / Cfg<NY>::T;
    dim3 grid((a.nkr + F - 1) / F, nb, 1);
    size_t sm = y_smem<NY, NT>() + g_smem_pad;
    if (a.ax.dealias) {
      if (!has_in) k_fused_y<NY, FAM_RK4, false, true, NT, true><<<grid, NT, sm, st>>>(a);
      else if (fam == FAM_RK4) k_fused_y<NY, FAM_RK4, true, true, NT, true><<<grid, NT, sm, st>>>(a);
      else if (fam == FAM_ETD) k_fused_y<NY, FAM_ETD, true, true, NT, true><<<grid, NT, sm, st>>>(a);
      else k_fused_y<NY, FAM_OTHER, true, true, NT, true><<<grid, NT, sm, st>>>(a);
      return;
    }
    if (!has_in) k_fused_y<NY, FAM_RK4, false, true, NT, false><<<grid, NT, sm, st>>>(a);
    else if (fam == FAM_RK4) k_fused_y<NY, FAM_RK4, true, true, NT, false><<<grid, NT, sm, st>>>(a);
    else if (fam == FAM_ETD) k_fused_y<NY, FAM_ETD, true, true, NT, false><<<grid, NT, sm, st>>>(a);
    else k_fused_y<NY, FAM_OTHER, true, true, NT, false><<<grid, NT, sm, st>>>(a);
  }
}
template <int NY>
void launch_y(bool has_in, int fam, const YArgs& a, int nb, cudaStream_t st, int n_sm) {
  int nt = pick_nt<NY>((long)a.nkr * nb, n_sm);
  if (nt == 256) launch_y_nt<NY, 256>(has_in, fam, a, nb, st);
  else if (nt == 128) launch_y_nt<NY, 128>(has_in, fam, a, nb, st);
  else launch_y_nt<NY, 64>(has_in, fam, a, nb, st);
}

template <int NX, int NT>
void launch_x_nt(int vmode, const XArgs& a, int nb, cudaStream_t st) {
  if constexpr (NT >= Cfg<NX>::T) {
    constexpr int F = NT / Cfg<NX>::T;
    dim3 grid((a.ny / 2) / F, nb, 1);
    size_t sm = x_smem<NX, NT>() + g_smem_pad;
    if (vmode == 2) k_fused_x<NX, 2, NT><<<grid, NT, sm, st>>>(a);
    else if (vmode == 3) k_fused_x<NX, 3, NT><<<grid, NT, sm, st>>>(a);
    else if (vmode == 4) k_fused_x<NX, 4, NT><<<grid, NT, y_smem<NX, NT>() + g_smem_pad, st>>>(a);
    else k_fused_x<NX, 0, NT><<<grid, NT, sm, st>>>(a);
  }
}
template <int NX>
void launch_x(int vmode, const XArgs& a, int nb, cudaStream_t st, int n_sm) {
  int nt = pick_nt<NX>((long)(a.ny / 2) * nb, n_sm);
  if (nt == 256) launch_x_nt<NX, 256>(vmode, a, nb, st);
  else if (nt == 128) launch_x_nt<NX, 128>(vmode, a, nb, st);
  else launch_x_nt<NX, 64>(vmode, a, nb, st);
}


}  // namespace
}  // namespace ptf
