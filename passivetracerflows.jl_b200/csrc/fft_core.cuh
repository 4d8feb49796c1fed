// Hand-written fp64 complex FFT for one CTA sub-group: N = 256*R3 points (R3 in {1,2,4,8,16}), T = N/16 threads,
// 16 elements per thread held in registers, Stockham autosort passes radix 16 -> 16 -> R3 with padded
// shared-memory exchanges between them.  N = 64 and 128 (the reference's default grid size, TAD.jl:144,166) run two
// passes, radix 16 -> N/16, with one exchange.
//
// Thread <-> data contract (both on entry and on exit):  thread t holds element  t + T*e  in logical slot e.
//   * on entry  v[e]            = x[t + T*e]            (natural slots)
//   * on exit   v[out_slot(e)]  = X[t + T*e]            (compile-time permuted slots; no data movement)
// so global loads/stores are coalesced (consecutive threads, consecutive elements) at both ends and every pointwise
// prologue/epilogue fuses on registers.
//
// Shared memory: one padded buffer of N + N/16 double2 per transform (index i -> i + i/16); with this padding every
// STS.128/LDS.128 of the three exchange patterns is bank-conflict free per quarter-warp.
//
// DIR = -1: forward  X[k] = sum x[n] exp(-2*pi*i*n*k/N);  DIR = +1: unnormalised inverse.
#pragma once
#include <cuda_runtime.h>

namespace ptf {
namespace fft {

#define PTF_HD __host__ __device__ __forceinline__

PTF_HD double2 cadd2(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
PTF_HD double2 csub2(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
PTF_HD double2 cmul2(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// multiply by a table twiddle stored for the FORWARD transform; the inverse uses its conjugate
template <int DIR>
PTF_HD double2 twmul(double2 a, double2 w) {
  if (DIR < 0) return make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
  return make_double2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}
// multiply by the constant (c, -s) forward / (c, +s) inverse
template <int DIR>
PTF_HD double2 cmulc(double2 a, double c, double s) {
  if (DIR < 0) return make_double2(a.x * c + a.y * s, a.y * c - a.x * s);
  return make_double2(a.x * c - a.y * s, a.y * c + a.x * s);
}
// multiply by -i (forward) / +i (inverse)
template <int DIR>
PTF_HD double2 mul_mi(double2 a) {
  if (DIR < 0) return make_double2(a.y, -a.x);
  return make_double2(-a.y, a.x);
}

PTF_HD void dft2(double2& a, double2& b) {
  double2 t = a;
  a = cadd2(t, b);
  b = csub2(t, b);
}

// natural order in, natural order out
template <int DIR>
PTF_HD void dft4(double2& x0, double2& x1, double2& x2, double2& x3) {
  double2 a = cadd2(x0, x2), b = csub2(x0, x2), c = cadd2(x1, x3), d = csub2(x1, x3);
  double2 md = mul_mi<DIR>(d);  // -i*d forward, +i*d inverse
  x0 = cadd2(a, c);
  x2 = csub2(a, c);
  x1 = cadd2(b, md);
  x3 = csub2(b, md);
}

constexpr double C_PI8 = 0.92387953251128675613;   // cos(pi/8)
constexpr double S_PI8 = 0.38268343236508977173;   // sin(pi/8)
constexpr double SQH = 0.70710678118654752440;     // sqrt(1/2)

// slot that holds logical output r after dft16 / dft8
PTF_HD constexpr int sl16(int r) { return ((r & 3) << 2) | (r >> 2); }
PTF_HD constexpr int sl8(int r) { return ((r & 3) << 1) | (r >> 2); }

// 16-point DFT in place (4x4).  Input natural; logical output r is left in slot sl16(r).
template <int DIR>
PTF_HD void dft16(double2 (&v)[16]) {
#pragma unroll
  for (int a = 0; a < 4; ++a) dft4<DIR>(v[a], v[a + 4], v[a + 8], v[a + 12]);
  // v[a + 4c] *= w16^(a*c)
  v[1 + 4] = cmulc<DIR>(v[1 + 4], C_PI8, S_PI8);                  // w^1
  v[1 + 8] = cmulc<DIR>(v[1 + 8], SQH, SQH);                      // w^2
  v[1 + 12] = cmulc<DIR>(v[1 + 12], S_PI8, C_PI8);                // w^3
  v[2 + 4] = cmulc<DIR>(v[2 + 4], SQH, SQH);                      // w^2
  v[2 + 8] = mul_mi<DIR>(v[2 + 8]);                               // w^4 = -i
  v[2 + 12] = cmulc<DIR>(v[2 + 12], -SQH, SQH);                   // w^6
  v[3 + 4] = cmulc<DIR>(v[3 + 4], S_PI8, C_PI8);                  // w^3
  v[3 + 8] = cmulc<DIR>(v[3 + 8], -SQH, SQH);                     // w^6
  v[3 + 12] = cmulc<DIR>(v[3 + 12], -C_PI8, -S_PI8);              // w^9
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4<DIR>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// 8-point DFT in place (2x4) on 8 references.  Input natural; logical output r is left in position sl8(r).
template <int DIR>
PTF_HD void dft8(double2& x0, double2& x1, double2& x2, double2& x3, double2& x4, double2& x5, double2& x6,
                 double2& x7) {
  dft4<DIR>(x0, x2, x4, x6);  // a = 0: u[0][c] -> positions 0,2,4,6  (position a + 2c)
  dft4<DIR>(x1, x3, x5, x7);  // a = 1: u[1][c] -> positions 1,3,5,7
  x3 = cmulc<DIR>(x3, SQH, SQH);    // w8^1
  x5 = mul_mi<DIR>(x5);             // w8^2
  x7 = cmulc<DIR>(x7, -SQH, SQH);   // w8^3
  dft2(x0, x1);
  dft2(x2, x3);
  dft2(x4, x5);
  dft2(x6, x7);
}

PTF_HD constexpr int pad_idx(int i) { return i + (i >> 4); }

#ifdef __CUDACC__
// Wavenumber of element l = t + T*e (T = N/16) of an fftfreq-ordered axis of length N, k = (double)j * c with
// j = l (l < N/2) or l - N: bit-identical to the host-built table (one correctly rounded product of an exact integer),
// without the table load.  e is a compile-time constant at every call site, so j = t + const.
template <int N>
__device__ __forceinline__ double wavenumber_full(int t, int e, double c, int nyq_sign) {
  constexpr int T = N / 16;
  const double j = (double)t + (double)(T * e - (e >= 8 ? N : 0));
  double k = j * c;
  if (e == 8 && nyq_sign > 0 && t == 0) k = -k;   // l == N/2 stored positive on request
  return k;
}
// same for the r2c axis (k = l * c, l <= N/2)
template <int N>
__device__ __forceinline__ double wavenumber_half(int t, int e, double c) {
  return ((double)t + (double)((N / 16) * e)) * c;
}

// Value barrier for the optimiser: the result is "a different value" as far as common-subexpression elimination and
// loop-invariant code motion can tell.  Used on gather strides / bases so that 16 precomputed 64-bit addresses are
// not kept live across a transform (32 registers = spills at 128 registers per thread); recomputing them is 1 IMAD each.
__device__ __forceinline__ size_t opaque(size_t x) {
  asm volatile("" : "+l"(x));
  return x;
}
#endif

// Device-resident twiddle tables for one transform length (forward sign; inverse conjugates on use).
struct Twiddles {
  const double2* tw2;  // [4][16]    N >= 256: w_256^(2^m * k);  N < 256: w_N^(2^m * k);  m = 0..3, k = 0..15
  const double2* tw3;  // [4][256]   w_N^(2^m * k), m = 0..3, k = 0..255   (unused for N <= 256)
};

template <int N>
struct Cfg {
  static constexpr int T = N / 16;     // threads per transform
  static constexpr bool TWO_PASS = N < 256;            // radix 16 -> N/16
  static constexpr int R3 = TWO_PASS ? N / 16 : N / 256;  // radix of the LAST pass (1 = no last pass, N = 256)
  static constexpr int S = (R3 > 0) ? 16 / R3 : 16;  // work items per thread in the last pass
  static constexpr int PADN = N + N / 16;
  static_assert(N == 64 || N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096,
                "unsupported FFT length");
};

// logical element e (index t + T*e) -> register slot, after the transform
template <int N>
PTF_HD constexpr int out_slot(int e) {
  return (Cfg<N>::R3 == 1)    ? sl16(e)
         : (Cfg<N>::R3 == 16) ? sl16(e)
         : (Cfg<N>::R3 == 8)  ? ((e % Cfg<N>::S) + sl8(e / Cfg<N>::S) * Cfg<N>::S)
                              : e;
}

// v[r] *= w^r for r = 1..15 given w^1, w^2, w^4, w^8 (forward-sign table values)
template <int DIR>
PTF_HD void twiddle15(double2 (&v)[16], double2 w1, double2 w2, double2 w4, double2 w8) {
  const double2 w3 = cmul2(w1, w2), w5 = cmul2(w4, w1), w6 = cmul2(w4, w2);
  const double2 w7 = cmul2(w4, w3);
  v[1] = twmul<DIR>(v[1], w1);
  v[2] = twmul<DIR>(v[2], w2);
  v[3] = twmul<DIR>(v[3], w3);
  v[4] = twmul<DIR>(v[4], w4);
  v[5] = twmul<DIR>(v[5], w5);
  v[6] = twmul<DIR>(v[6], w6);
  v[7] = twmul<DIR>(v[7], w7);
  v[8] = twmul<DIR>(v[8], w8);
  v[9] = twmul<DIR>(v[9], cmul2(w8, w1));
  v[10] = twmul<DIR>(v[10], cmul2(w8, w2));
  v[11] = twmul<DIR>(v[11], cmul2(w8, w3));
  v[12] = twmul<DIR>(v[12], cmul2(w8, w4));
  v[13] = twmul<DIR>(v[13], cmul2(w8, w5));
  v[14] = twmul<DIR>(v[14], cmul2(w8, w6));
  v[15] = twmul<DIR>(v[15], cmul2(w8, w7));
}

#ifdef __CUDACC__
// Last Stockham pass: radix R3 over sub-transforms of length NS (= N / R3); work item q of thread t is element
// j = t + q*T, uses register slots q + r*S, twiddles w_N^(r*k) with k = j mod NS read from tab[m*NS + k] = w_N^(2^m k).
// The table entries are fetched by last_pass_load BEFORE the exchange barrier that precedes the pass (the data registers
// are in shared memory at that point, so the loads cost no register pressure and their latency hides behind the barrier).
template <int N>
struct LastTw {
  static constexpr int R3 = Cfg<N>::R3, S = Cfg<N>::S;
  static constexpr int NW = R3 == 2 ? 1 : (R3 == 4 ? 2 : (R3 == 8 ? 3 : 4));
  double2 w[S > 0 ? S : 1][NW];
};
template <int N, int NS>
__device__ __forceinline__ void last_pass_load(LastTw<N>& W, int t, const double2* __restrict__ tab) {
  constexpr int T = Cfg<N>::T, S = Cfg<N>::S;
#pragma unroll
  for (int q = 0; q < S; ++q) {
    const int k = (t + q * T) & (NS - 1);
#pragma unroll
    for (int m = 0; m < LastTw<N>::NW; ++m) W.w[q][m] = __ldg(&tab[m * NS + k]);
  }
}
template <int N, int DIR>
__device__ __forceinline__ void last_pass(double2 (&v)[16], const LastTw<N>& W) {
  constexpr int R3 = Cfg<N>::R3, S = Cfg<N>::S;
#pragma unroll
  for (int q = 0; q < S; ++q) {
    const double2 w1 = W.w[q][0];
    if (R3 == 2) {
      v[q + S] = twmul<DIR>(v[q + S], w1);
      dft2(v[q], v[q + S]);
    } else if (R3 == 4) {
      const double2 w2 = W.w[q][LastTw<N>::NW > 1 ? 1 : 0];
      double2 w3 = cmul2(w1, w2);
      v[q + S] = twmul<DIR>(v[q + S], w1);
      v[q + 2 * S] = twmul<DIR>(v[q + 2 * S], w2);
      v[q + 3 * S] = twmul<DIR>(v[q + 3 * S], w3);
      dft4<DIR>(v[q], v[q + S], v[q + 2 * S], v[q + 3 * S]);
    } else if (R3 == 8) {
      const double2 w2 = W.w[q][LastTw<N>::NW > 1 ? 1 : 0];
      const double2 w4 = W.w[q][LastTw<N>::NW > 2 ? 2 : 0];
      double2 w3 = cmul2(w1, w2), w5 = cmul2(w4, w1), w6 = cmul2(w4, w2);
      double2 w7 = cmul2(w4, w3);
      v[q + S] = twmul<DIR>(v[q + S], w1);
      v[q + 2 * S] = twmul<DIR>(v[q + 2 * S], w2);
      v[q + 3 * S] = twmul<DIR>(v[q + 3 * S], w3);
      v[q + 4 * S] = twmul<DIR>(v[q + 4 * S], w4);
      v[q + 5 * S] = twmul<DIR>(v[q + 5 * S], w5);
      v[q + 6 * S] = twmul<DIR>(v[q + 6 * S], w6);
      v[q + 7 * S] = twmul<DIR>(v[q + 7 * S], w7);
      dft8<DIR>(v[q], v[q + S], v[q + 2 * S], v[q + 3 * S], v[q + 4 * S], v[q + 5 * S], v[q + 6 * S], v[q + 7 * S]);
    } else {  // R3 == 16 (S == 1, q == 0)
      const double2 w2 = W.w[q][LastTw<N>::NW > 1 ? 1 : 0];
      const double2 w4 = W.w[q][LastTw<N>::NW > 2 ? 2 : 0];
      const double2 w8 = W.w[q][LastTw<N>::NW > 3 ? 3 : 0];
      twiddle15<DIR>(v, w1, w2, w4, w8);
      dft16<DIR>(v);
    }
  }
}

// Barrier among the T threads of ONE transform group (group `grp` of a CTA of NT threads).  Groups of a CTA are
// independent transforms: with CTA-wide barriers they would run in lock-step (every exchange waits for the slowest
// group's memory phase); named barriers / warp barriers let each group proceed on its own.
template <int T, int NT>
__device__ __forceinline__ void group_sync(int grp) {
  if (NT == 0 || T >= NT) __syncthreads();
  else if (T >= 32) asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(T) : "memory");
  else __syncwarp();   // T < 32 divides the warp: a group never spans two warps
}

// One transform by T cooperating threads of a CTA; all threads of the CTA must call it (CTA-wide barriers, or with
// NT > 0 barriers among the T threads of group `grp` only).
// `sm` points at this transform's padded buffer; it may be reused by the caller after the call returns AND a barrier.
// OPAQUE_TW: hide the twiddle-table pointers from the optimiser for this call.  A kernel that runs several transforms
// of the same length back to back otherwise gets their (identical) twiddle loads merged and kept live across the
// transforms in between: up to 48 registers, i.e. spills at 128 registers per thread.
// EARLY_TW: request the next pass's twiddles before the exchange barrier (hides their latency; measured better in the
// column kernels, worse in the row kernels, whose register budget is tighter).
template <int N, int DIR, bool OPAQUE_TW = false, int NT = 0, bool EARLY_TW = true>
__device__ __forceinline__ void fft_cta(double2 (&v)[16], double2* __restrict__ sm, int t, const Twiddles& tw_in,
                                        int grp = 0) {
  constexpr int T = Cfg<N>::T, R3 = Cfg<N>::R3;
  Twiddles tw = tw_in;
  if (OPAQUE_TW) {
    asm volatile("" : "+l"(tw.tw2));
    asm volatile("" : "+l"(tw.tw3));
  }
  // ---- pass 1: radix 16, Ns = 1, no twiddles ----
  dft16<DIR>(v);
  group_sync<T, NT>(grp);  // buffer free (previous readers done)
#pragma unroll
  for (int r = 0; r < 16; ++r) sm[pad_idx(16 * t + r)] = v[sl16(r)];
  if (Cfg<N>::TWO_PASS) {  // N = 64, 128: last pass radix N/16 over the 16-point sub-transforms
    LastTw<N> W;
    if (EARLY_TW) last_pass_load<N, 16>(W, t, tw.tw2);   // in flight across the barrier
    group_sync<T, NT>(grp);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = sm[pad_idx(t + T * e)];
    if (!EARLY_TW) last_pass_load<N, 16>(W, t, tw.tw2);
    last_pass<N, DIR>(v, W);
    return;
  }
  // ---- pass 2: radix 16, Ns = 16 ----
  const int k2 = t & 15;
  {
    // w_256^(r*k2), r = 1..15, from 4 table entries (r = 1,2,4,8) and at most 3 chained products; requested before the
    // barrier (the data is in shared memory: no register pressure) so that their latency hides behind the exchange
    double2 w1, w2, w4, w8;
    if (EARLY_TW) {
      w1 = __ldg(&tw.tw2[k2]), w2 = __ldg(&tw.tw2[16 + k2]);
      w4 = __ldg(&tw.tw2[32 + k2]), w8 = __ldg(&tw.tw2[48 + k2]);
    }
    group_sync<T, NT>(grp);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = sm[pad_idx(t + T * e)];
    if (!EARLY_TW) {
      w1 = __ldg(&tw.tw2[k2]), w2 = __ldg(&tw.tw2[16 + k2]);
      w4 = __ldg(&tw.tw2[32 + k2]), w8 = __ldg(&tw.tw2[48 + k2]);
    }
    twiddle15<DIR>(v, w1, w2, w4, w8);
  }
  dft16<DIR>(v);
  if (R3 == 1) return;
  group_sync<T, NT>(grp);
  {
    const int base = 16 * (t - k2) + k2;
#pragma unroll
    for (int r = 0; r < 16; ++r) sm[pad_idx(base + 16 * r)] = v[sl16(r)];
  }
  // ---- pass 3: radix R3, Ns = 256 ----
  LastTw<N> W;
  if (EARLY_TW) last_pass_load<N, 256>(W, t, tw.tw3);
  group_sync<T, NT>(grp);
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = sm[pad_idx(t + T * e)];
  if (!EARLY_TW) last_pass_load<N, 256>(W, t, tw.tw3);
  last_pass<N, DIR>(v, W);
}
#endif  // __CUDACC__

}  // namespace fft
}  // namespace ptf
