#pragma once
// y-column kernels of the fused 3-D engine (engine_fused3d.cu).  One stage of a FourierFlows stepper in 3-D
// (calcN!, TAD.jl:771-786 steady / :725-742 time-varying, + the stage combine) is FOUR hand-written kernels:
//
//   k_fused_y<.., D3>  z columns   gather P^xy -> FFT_z -> N^ -> stage combine -> s' -> A = IFFT_z(s'/N), C = IFFT_z(i m s'/N)
//   k_yinv3            y columns   A, C (pairs of z planes) -> A' = IFFT_y(A), B' = IFFT_y(i l A), C' = IFFT_y(C)
//   k_fused_x<.., D3>  x row pairs c2r of i kr A', B', C' -> p = -u gx - v gy - w gz -> r2c -> P^x
//   k_yfwd3            y columns   gather P^x (pairs of z planes) -> FFT_y -> P^xy
//
// so every 3-D transform costs one read + one write per axis and all pointwise work rides in the FFT I/O.
//
// Layouts in HBM (P ranks; nyl = ny/P spectral ky rows and nzl = nz/P physical planes per rank; P = 1 on one GPU):
//   state (sol, sol_1, acc, ...)   [kr][ll][z]                 z contiguous: the z-column kernel's natural order
//   A, C   as written / sent       [p][kr][ll][zl]             p = z / nzl : the block that goes to rank p
//          as received / read      [r][kr][ll][zl]             r = l / nyl : the block that came from rank r
//   A',B',C'                       [zl][kr][y]                 = the 2-D engine's [b][kr][y] with b = plane
//   P^x                            [zl][y/8][kr][y%8]          = the 2-D engine's blocked layout per plane
//   P^xy   as written / sent       [p][kr][zl/8][ll][zl%8]     p = l / nyl
//          as received / read      [r][kr][zl/8][ll][zl%8]     r = z / nzl
// Every kernel writes whole 32-byte sectors (contiguous runs or z-plane pairs) and reads the other kernel's layout
// with 32-byte pair gathers or 128-byte line gathers; the slab exchange between ranks moves contiguous blocks, so
// there is no pack / unpack pass anywhere.
#include "fused_kernels.cuh"

namespace ptf {
namespace {

struct Y3Args {
  const double2* RA;   // [r][kr][ll][zl]   IFFT_z(s'/N)
  const double2* RC;   // [r][kr][ll][zl]   IFFT_z(i m s'/N)
  double2 *YA, *YB, *YC;  // [zl][kr][y]
  const double2* PX;   // [zl][y/8][kr][y%8]
  double2* PXY;        // [p][kr][zl/8][ll][zl%8]
  double cy;           // generating constant of the ky table (fft::wavenumber_full)
  int nyq_sign;
  Twiddles tw;
  int nkx, nyl, nzl;
  // element l = t + T*e of a y column lives in the block of rank e >> esh (nyl is a multiple of T = ny/16 for P <= 16):
  // offset(e) = base(t) + e * s_e + (e >> esh) * s_r, all in double2 units (see y3_strides)
  int esh;
  long long in_se, in_sr, out_se, out_sr;
  int kr0 = 0, nkr_launch = 0;   // this launch covers kr in [kr0, kr0 + nkr_launch): a chunk of the pipelined exchange
};

// ---- inverse y transforms of one pair of z planes (zl, zl+1) of one kr ----
template <int NY, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_yinv3(Y3Args a) {
  constexpr int T = Cfg<NY>::T, F = NT / T, PADN = Cfg<NY>::PADN;
  constexpr int TCOLS = NT > 128 ? 256 : 128;
  static_assert(NT >= T && NT % T == 0, "CTA must hold whole transforms");
  extern __shared__ double2 smem[];
  __shared__ uint32_t tslot;
  const uint32_t tbase = tmem::alloc_cta<TCOLS>(&tslot);
  const uint32_t t_p = tmem::warp_addr(tbase, 128);  // 64 columns: plane zl+1's input | 64 columns: i*l*A of the current plane
  const uint32_t t_b = t_p + 64;
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  const int zp_raw = blockIdx.x * F + grp;
  const bool active = 2 * zp_raw < a.nzl;
  const int zl = active ? 2 * zp_raw : 0;
  const int kr = a.kr0 + blockIdx.y;
  double2* sm = smem + grp * PADN;
  // element l = t + T*e of this (kr, zl) column in the received blocks [r][kr][ll][zl]
  const size_t ibase = ((size_t)kr * a.nyl + t) * a.nzl + zl;
  const size_t plane = (size_t)a.nkx * NY;
  const size_t ocol = (size_t)kr * NY + t;
  double2 v[16];
  {
    const size_t se = fft::opaque((size_t)a.in_se), sr = (size_t)a.in_sr;
#pragma unroll
    for (int e = 0; e < 16; ++e) prefetch_l2(a.RC + ibase + e * se + (size_t)(e >> a.esh) * sr);
  }
  const size_t se0 = fft::opaque((size_t)a.in_se), sr0 = (size_t)a.in_sr;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    double2 p0[4], p1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      tmem::ldg256(a.RA + ibase + (4 * h + j) * se0 + (size_t)((4 * h + j) >> a.esh) * sr0, p0[j], p1[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tmem::st1(t_p + 4 * (4 * h + j), p1[j]);
      v[4 * h + j] = p0[j];
    }
  }
#pragma unroll 1
  for (int it = 0; it < 6; ++it) {   // NOT unrolled: one copy of the transform
    // it: 0 A(zl) | 1 i l A(zl) | 2 A(zl+1) | 3 i l A(zl+1) | 4 C(zl) | 5 C(zl+1)
    if (it == 4) {
      const size_t se = fft::opaque((size_t)a.in_se), sr = (size_t)a.in_sr;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        double2 p0[4], p1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tmem::ldg256(a.RC + ibase + (4 * h + j) * se + (size_t)((4 * h + j) >> a.esh) * sr, p0[j], p1[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tmem::st1(t_p + 4 * (4 * h + j), p1[j]);
          v[4 * h + j] = p0[j];
        }
      }
      tmem::wait_st();
    } else if (it != 0) {
      const uint32_t src = (it == 2 || it == 5) ? t_p : t_b;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double2 r4[4];
        tmem::ldn<4>(src + 16 * q, r4);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[4 * q + j] = r4[j];
      }
    }
    if (it == 0 || it == 2) {   // the y derivative of this plane, parked until A's transform is stored
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const double ky = fft::wavenumber_full<NY>(t, e, a.cy, a.nyq_sign);
        tmem::st1(t_b + 4 * e, make_double2(-ky * v[e].y, ky * v[e].x));
      }
      tmem::wait_st();
    }
    fft::fft_cta<NY, +1, false, NT>(v, sm, t, a.tw, grp);
    double2* dst = (it == 0 || it == 2) ? a.YA : ((it == 1 || it == 3) ? a.YB : a.YC);
    const int zq = zl + ((it == 2 || it == 3 || it == 5) ? 1 : 0);
    if (active) {
      double2* o = dst + (size_t)zq * plane + ocol;
#pragma unroll
      for (int e = 0; e < 16; ++e) __stcg(o + T * e, v[out_slot<NY>(e)]);
    }
  }
  tmem::free_cta<TCOLS>(tbase);
}

// ---- forward y transforms of one pair of z planes of one kr ----
template <int NY, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_yfwd3(Y3Args a) {
  constexpr int T = Cfg<NY>::T, F = NT / T, PADN = Cfg<NY>::PADN;
  constexpr int TCOLS = NT > 128 ? 256 : 128;
  extern __shared__ double2 smem[];
  __shared__ uint32_t tslot;
  const uint32_t tbase = tmem::alloc_cta<TCOLS>(&tslot);
  const uint32_t t_r = tmem::warp_addr(tbase, 128);
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  const int zp_raw = blockIdx.x * F + grp;
  const bool active = 2 * zp_raw < a.nzl;
  const int zl = active ? 2 * zp_raw : 0;
  const int kr = a.kr0 + blockIdx.y;
  double2* sm = smem + grp * PADN;
  // element l = t + T*e of this (kr, zl) column in the blocks to send [p][kr][zl/8][ll][zl%8]
  const size_t obase = (((size_t)kr * (a.nzl >> 3) + (zl >> 3)) * a.nyl + t) * 8 + (zl & 7);
  double2 v[16];
#pragma unroll 1
  for (int q = 0; q < 2; ++q) {
    // P^x of plane zl+q: [y/8][kr][y%8], 8 consecutive y of this kr are one 128-byte line
    // y = t + T*e: one base + a single stride per element (nothing for the compiler to keep live across the transform)
    const double2* P = a.PX + (size_t)(zl + q) * NY * a.nkx + (size_t)kr * 8;
    if (T >= 8) {
      const double2* Pt = P + (size_t)(t >> 3) * a.nkx * 8 + (t & 7);
      const size_t se = fft::opaque((size_t)(T / 8) * a.nkx * 8);   // opaque: not hoisted out of the q loop
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = __ldcg(Pt + e * se);
    } else {   // T == 4 (NY = 64): y >> 3 = e >> 1, y & 7 = t + 4 (e & 1)
      const double2* Pt = P + t;
      const size_t se = fft::opaque((size_t)a.nkx * 8);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = __ldcg(Pt + (e >> 1) * se + 4 * (e & 1));
    }
    fft::fft_cta<NY, -1, false, NT>(v, sm, t, a.tw, grp);
    if (q == 0) {
#pragma unroll
      for (int e = 0; e < 16; ++e) tmem::st1(t_r + 4 * e, v[out_slot<NY>(e)]);
      tmem::wait_st();
    } else {
      const size_t ose = fft::opaque((size_t)a.out_se), osr = (size_t)a.out_sr;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        double2 r[4];
        tmem::ldn<4>(t_r + 16 * qq, r);
        if (active) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = 4 * qq + j;
            const size_t off = obase + e * ose + (size_t)(e >> a.esh) * osr;
            tmem::stg256(a.PXY + off, r[j], v[out_slot<NY>(e)]);
          }
        }
      }
    }
  }
  tmem::free_cta<TCOLS>(tbase);
}

// ---- cross-GPU barrier of the P2P slab exchange (one process per GPU, peers' flag words mapped through CUDA IPC) ----
// Every rank bumps its epoch counter, publishes the epoch in slot `me` of every peer's flag array (release, system
// scope) and waits until all peers have published theirs here (acquire).  Stream order puts it after the kernel whose
// peer stores / loads it fences.  A rank that never arrives (crashed peer) trips the time-out instead of hanging the GPU.
struct XbArgs {
  unsigned* peer[16];   // [p]: rank p's flag array (own array for p == me)
  int me, P;
};
__global__ void __launch_bounds__(32) k_xbarrier(unsigned* __restrict__ local, XbArgs a) {
  __shared__ unsigned epoch_s;
  if (threadIdx.x == 0) epoch_s = ++local[16];   // slot 16: this rank's barrier count (same sequence on every rank)
  __syncthreads();
  const unsigned epoch = epoch_s;
  const int p = threadIdx.x;
  if (p < a.P) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer[p] + a.me), "r"(epoch) : "memory");
    unsigned v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local + p) : "memory");
      if (clock64() - t0 > 40000000000LL) {   // ~20 s: a peer is gone
        printf("libptf_b200: cross-GPU barrier timed out (rank %d waiting for rank %d, epoch %u, saw %u)\n", a.me, p,
               epoch, v);
        __trap();
      }
    } while ((int)(v - epoch) < 0);
    __threadfence_system();
  }
}

// ---- separable flows (PTF_FLOW_SEPARABLE): u_a = sum_m a_m(t) X_m(x) Y_m(y) Z_m(z) written out once per step ----
// The velocity is frozen at clock.t for all stages of a step (TAD.jl:737), so the three fields are evaluated ONCE per
// step (24 B/pt, ~4 % of a step's traffic) and the row kernel reads them like steady arrays; evaluating the sums
// per point inside the row kernel costs 2-4 table loads per point and component and was measured 3x slower.
// One (y, z) row per block iteration: the row's coefficients a_m(t) Y_m(y) Z_m(z) are formed once (shared memory), every
// point then costs one table load and one FMA per term and component.
__global__ void __launch_bounds__(256) k_sep_fill(VelArgs va, double* __restrict__ u, double* __restrict__ v,
                                                  double* __restrict__ w, int nx, int ny, int nzl, int nzg, int zoff) {
  __shared__ double cs[3][MAX_TERMS];
  double* const out[3] = {u, v, w};
  const int64_t rows = (int64_t)ny * nzl;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int y = (int)(row % ny);
    const int z = (int)(row / ny) + zoff;
    __syncthreads();   // previous row's readers are done
    if (threadIdx.x < 3 * MAX_TERMS) {
      const int c = threadIdx.x / MAX_TERMS, m = threadIdx.x % MAX_TERMS;
      const SepFlow& f = va.sep[c];
      cs[c][m] = m < f.nterms ? f.a[m] * f.yt[(int64_t)m * ny + y] * f.zt[(int64_t)m * nzg + z] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const SepFlow& f = va.sep[c];
      double* o = out[c] + row * nx;
      for (int x = threadIdx.x; x < nx; x += blockDim.x) {
        double acc = 0.0;
        for (int m = 0; m < f.nterms; ++m) acc += cs[c][m] * f.xt[(int64_t)m * nx + x];
        __stcg(o + x, acc);
      }
    }
  }
}

// ---- boundary-only layout changes (set_c! / updatevars! / get_sol / set_sol; not on the step path) ----
// x-transformed planes [zl][y][kr] -> blocked P^x [zl][y/8][kr][y%8]
__global__ void __launch_bounds__(256) k_block_px(const double2* __restrict__ in, double2* __restrict__ out, int nkx,
                                                  int ny, int64_t nplanes) {
  const int64_t n = nplanes * ny * nkx;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int kr = (int)(i % nkx);
    const int y = (int)((i / nkx) % ny);
    const int64_t z = i / ((int64_t)nkx * ny);
    out[(z * ny + (int64_t)(y >> 3) * 8) * nkx + (int64_t)kr * 8 + (y & 7)] = in[i];
  }
}
// A' [zl][kr][y] -> [zl][y][kr] for the final c2r along x, dropping what c2r ignores (Im of the kr = 0 / Nyquist bins)
__global__ void __launch_bounds__(256) k_unblock_c2r(const double2* __restrict__ in, double2* __restrict__ out,
                                                     int nkx, int ny, int64_t nplanes) {
  __shared__ double2 tile[32][33];
  const int64_t b = blockIdx.z;
  const double2* I = in + b * nkx * ny;
  double2* O = out + b * nkx * ny;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;   // c: y (contiguous in), r: kr
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < nkx && c < ny) tile[j][threadIdx.x] = I[(size_t)r * ny + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < nkx && c < ny) {
      double2 v = tile[threadIdx.x][j];
      if (r == 0 || r == nkx - 1) v.y = 0.0;
      O[(size_t)c * nkx + r] = v;
    }
  }
}
// state [kr][ll][z] <-> canonical local spectral slab [z][ll][kr]  (DIR 0: state -> canonical, 1: back)
template <int DIR>
__global__ void __launch_bounds__(256) k_state_canon(const double2* __restrict__ in, double2* __restrict__ out,
                                                     int nkx, int nyl, int nz) {
  __shared__ double2 tile[32][33];
  const int ll = blockIdx.z;
  const int z0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  // state index (kr*nyl + ll)*nz + z  (z contiguous); canonical index (z*nyl + ll)*nkx + kr  (kr contiguous)
  if (DIR == 0) {
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int kr = k0 + j, z = z0 + threadIdx.x;
      if (kr < nkx && z < nz) tile[j][threadIdx.x] = in[((size_t)kr * nyl + ll) * nz + z];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int z = z0 + j, kr = k0 + threadIdx.x;
      if (kr < nkx && z < nz) out[((size_t)z * nyl + ll) * nkx + kr] = tile[threadIdx.x][j];
    }
  } else {
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int z = z0 + j, kr = k0 + threadIdx.x;
      if (kr < nkx && z < nz) tile[j][threadIdx.x] = in[((size_t)z * nyl + ll) * nkx + kr];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int kr = k0 + j, z = z0 + threadIdx.x;
      if (kr < nkx && z < nz) out[((size_t)kr * nyl + ll) * nz + z] = tile[threadIdx.x][j];
    }
  }
}

// ---- launchers ----
template <int N, int NT>
constexpr size_t x3_smem() { return y_smem<N, NT>() + (size_t)(NT / Cfg<N>::T) * 2 * N * sizeof(double); }

template <int N, int NT>
void prep3_nt() {
  if constexpr (NT >= Cfg<N>::T) {
    const size_t ys = y_smem<N, NT>() + g_smem_pad;
    allow_smem(k_fused_y<N, FAM_RK4, false, true, NT, false, true>, ys);
    allow_smem(k_fused_y<N, FAM_RK4, true, true, NT, false, true>, ys);
    allow_smem(k_fused_y<N, FAM_ETD, true, true, NT, false, true>, ys);
    allow_smem(k_fused_y<N, FAM_OTHER, true, true, NT, false, true>, ys);
    allow_smem(k_fused_y<N, FAM_RK4, false, true, NT, true, true>, ys);
    allow_smem(k_fused_y<N, FAM_RK4, true, true, NT, true, true>, ys);
    allow_smem(k_fused_y<N, FAM_ETD, true, true, NT, true, true>, ys);
    allow_smem(k_fused_y<N, FAM_OTHER, true, true, NT, true, true>, ys);
    allow_smem(k_fused_y<N, FAM_RK4, true, true, NT, false, true, true>, ys);
    allow_smem(k_fused_y<N, FAM_ETD, true, true, NT, false, true, true>, ys);
    allow_smem(k_fused_y<N, FAM_OTHER, true, true, NT, false, true, true>, ys);
    allow_smem(k_fused_y<N, FAM_RK4, true, true, NT, true, true, true>, ys);
    allow_smem(k_fused_y<N, FAM_ETD, true, true, NT, true, true, true>, ys);
    allow_smem(k_fused_y<N, FAM_OTHER, true, true, NT, true, true, true>, ys);
    allow_smem(k_fused_x<N, 0, NT, true>, x3_smem<N, NT>() + g_smem_pad);
    allow_smem(k_fused_x<N, 3, NT, true>, y_smem<N, NT>() + g_smem_pad);   // direct mode: products parked in TMEM
    allow_smem(k_yinv3<N, NT>, ys);
    allow_smem(k_yfwd3<N, NT>, ys);
  }
}
template <int N>
void prep3() {
  prep3_nt<N, 128>();
  prep3_nt<N, 64>();
}

// 3-D launches always have plenty of CTAs: 128-thread CTAs (4 per SM) unless a transform needs fewer threads and
// the launch is small
template <int N>
int pick_nt3(long items, int n_sm) {
  constexpr int T = Cfg<N>::T;
  const char* fe = std::getenv("PTF_NT3");
  const int forced = fe ? std::atoi(fe) : 0;
  if (forced == 64 || forced == 128) return forced < T ? T : forced;
  const long ctas128 = (items + 128 / T - 1) / (128 / T);
  if (ctas128 >= 2L * n_sm || T > 64) return 128;
  return 64;
}

template <int N, int NT>
void launch_z3_nt(bool has_in, int fam, const YArgs& a, cudaStream_t st) {
  if constexpr (NT >= Cfg<N>::T) {
    constexpr int F = NT / Cfg<N>::T;
    dim3 grid((a.cid_end - a.cid0 + F - 1) / F, 1, 1);
    const size_t sm = y_smem<N, NT>() + g_smem_pad;
    if (has_in && a.grid_cap > 0 && (int)grid.x > a.grid_cap) {   // small persistent grid (slab pipeline, P2P mode)
      grid.x = a.grid_cap;
      if (a.ax.dealias) {
        if (fam == FAM_RK4) k_fused_y<N, FAM_RK4, true, true, NT, true, true, true><<<grid, NT, sm, st>>>(a);
        else if (fam == FAM_ETD) k_fused_y<N, FAM_ETD, true, true, NT, true, true, true><<<grid, NT, sm, st>>>(a);
        else k_fused_y<N, FAM_OTHER, true, true, NT, true, true, true><<<grid, NT, sm, st>>>(a);
      } else {
        if (fam == FAM_RK4) k_fused_y<N, FAM_RK4, true, true, NT, false, true, true><<<grid, NT, sm, st>>>(a);
        else if (fam == FAM_ETD) k_fused_y<N, FAM_ETD, true, true, NT, false, true, true><<<grid, NT, sm, st>>>(a);
        else k_fused_y<N, FAM_OTHER, true, true, NT, false, true, true><<<grid, NT, sm, st>>>(a);
      }
      return;
    }
    if (a.ax.dealias) {
      if (!has_in) k_fused_y<N, FAM_RK4, false, true, NT, true, true><<<grid, NT, sm, st>>>(a);
      else if (fam == FAM_RK4) k_fused_y<N, FAM_RK4, true, true, NT, true, true><<<grid, NT, sm, st>>>(a);
      else if (fam == FAM_ETD) k_fused_y<N, FAM_ETD, true, true, NT, true, true><<<grid, NT, sm, st>>>(a);
      else k_fused_y<N, FAM_OTHER, true, true, NT, true, true><<<grid, NT, sm, st>>>(a);
      return;
    }
    if (!has_in) k_fused_y<N, FAM_RK4, false, true, NT, false, true><<<grid, NT, sm, st>>>(a);
    else if (fam == FAM_RK4) k_fused_y<N, FAM_RK4, true, true, NT, false, true><<<grid, NT, sm, st>>>(a);
    else if (fam == FAM_ETD) k_fused_y<N, FAM_ETD, true, true, NT, false, true><<<grid, NT, sm, st>>>(a);
    else k_fused_y<N, FAM_OTHER, true, true, NT, false, true><<<grid, NT, sm, st>>>(a);
  }
}
template <int N>
void launch_z3(bool has_in, int fam, const YArgs& a, cudaStream_t st, int n_sm) {
  if (pick_nt3<N>(a.cid_end - a.cid0, n_sm) == 128) launch_z3_nt<N, 128>(has_in, fam, a, st);
  else launch_z3_nt<N, 64>(has_in, fam, a, st);
}

template <int N, int NT>
void launch_x3_nt(int vmode, const XArgs& a, int nplanes, cudaStream_t st) {
  if constexpr (NT >= Cfg<N>::T) {
    constexpr int F = NT / Cfg<N>::T;
    dim3 grid((a.ny / 2 + F - 1) / F, nplanes, 1);
    const size_t sm = x3_smem<N, NT>() + g_smem_pad;
    // separable flows are written out once per step (k_sep_fill): the row kernel always reads arrays
    if (vmode == 3) k_fused_x<N, 3, NT, true><<<grid, NT, y_smem<N, NT>() + g_smem_pad, st>>>(a);
    else k_fused_x<N, 0, NT, true><<<grid, NT, sm, st>>>(a);
  }
}
template <int N>
void launch_x3(int vmode, const XArgs& a, int nplanes, cudaStream_t st, int n_sm) {
  if (pick_nt3<N>((long)(a.ny / 2) * nplanes, n_sm) == 128) launch_x3_nt<N, 128>(vmode, a, nplanes, st);
  else launch_x3_nt<N, 64>(vmode, a, nplanes, st);
}

template <int N, int NT>
void launch_y3_nt(bool inverse, const Y3Args& a, cudaStream_t st) {
  if constexpr (NT >= Cfg<N>::T) {
    constexpr int F = NT / Cfg<N>::T;
    dim3 grid((a.nzl / 2 + F - 1) / F, a.nkr_launch, 1);
    const size_t sm = y_smem<N, NT>() + g_smem_pad;
    if (inverse) k_yinv3<N, NT><<<grid, NT, sm, st>>>(a);
    else k_yfwd3<N, NT><<<grid, NT, sm, st>>>(a);
  }
}
template <int N>
void launch_y3(bool inverse, const Y3Args& a, cudaStream_t st, int n_sm) {
  if (pick_nt3<N>((long)(a.nzl / 2) * a.nkr_launch, n_sm) == 128) launch_y3_nt<N, 128>(inverse, a, st);
  else launch_y3_nt<N, 64>(inverse, a, st);
}

}  // namespace
}  // namespace ptf
