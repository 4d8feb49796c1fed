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
#include <cstring>

#include "fft_core.cuh"
#include "ptf_pointwise.cuh"
#include "ptf_velocity.cuh"

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
  int nkr;
  double inv_n;               // 1/(nx*ny): normalisation of ldiv!(., rfftplan, .)
};

template <int NY, bool HAS_IN, bool HAS_OUT>
__global__ void __launch_bounds__(256, 2) k_fused_y(YArgs a) {
  constexpr int T = Cfg<NY>::T, F = 256 / T, PADN = Cfg<NY>::PADN;
  extern __shared__ double2 smem[];
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  const int kr_raw = blockIdx.x * F + grp;
  const bool active = kr_raw < a.nkr;
  const int kr = active ? kr_raw : 0;
  const int b = blockIdx.y;
  double2* sm = smem + grp * PADN;
  const size_t col = ((size_t)b * a.nkr + kr) * NY;  // this column in the transposed state arrays
  const size_t ccol = (size_t)kr * NY;               // ... in the batch-shared coefficient arrays
  const double kx = a.ax.kx[kr];
  double2 w[16];

  if (HAS_IN) {
    double2 v[16];
    const double2* P = a.Px + (size_t)b * NY * a.nkr + kr;
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = P[(size_t)(t + T * e) * a.nkr];
    fft::fft_cta<NY, -1>(v, sm, t, a.tw);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int l = t + T * e;
      double2 nx = make_double2(0.0, 0.0);
      if (active) nx = combine_at(a.P, a.C, a.ax, col + l, ccol + l, kx, a.ax.ky[l], 0.0, v[out_slot<NY>(e)]);
      w[e] = make_double2(nx.x * a.inv_n, nx.y * a.inv_n);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      double2 s = a.next_state[col + t + T * e];
      w[e] = make_double2(s.x * a.inv_n, s.y * a.inv_n);
    }
  }
  if (!HAS_OUT) return;

  fft::fft_cta<NY, +1>(w, sm, t, a.tw);
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e) a.A[col + t + T * e] = w[out_slot<NY>(e)];
  }
  // y-derivative: i*l*s'  (s' re-read from the array this thread just wrote: L2 hit, no DRAM traffic)
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int l = t + T * e;
    double2 s = a.next_state[col + l];
    double ky = a.ax.ky[l] * a.inv_n;
    w[e] = make_double2(-ky * s.y, ky * s.x);
  }
  fft::fft_cta<NY, +1>(w, sm, t, a.tw);
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e) a.Bf[col + t + T * e] = w[out_slot<NY>(e)];
  }
}

struct XArgs {
  const double2* A;   // [b][kr][y]
  const double2* Bf;  // [b][kr][y]
  double2* Px;        // [b][y][kr]
  VelArgs va;
  AxisTables ax;
  Twiddles tw;
  int nkr, ny;
};

// VMODE 0: velocity arrays; 1: arrays + layered shift U(y,b); 2: separable tables (zero HBM bytes)
template <int NX, int VMODE>
__global__ void __launch_bounds__(256, 2) k_fused_x(XArgs a) {
  constexpr int T = Cfg<NX>::T, F = 256 / T, PADN = Cfg<NX>::PADN, H = NX / 2;
  extern __shared__ double2 smem[];
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  const int pair = blockIdx.x * F + grp;
  const int b = blockIdx.y;
  const int ny = a.ny;
  double2* sm = smem + grp * PADN;
  double* ps = reinterpret_cast<double*>(smem + F * PADN) + grp * NX;  // row-0 product, parked while row 1 runs
  const double2* Ab = a.A + (size_t)b * a.nkr * ny;
  const double2* Bb = a.Bf + (size_t)b * a.nkr * ny;
  const size_t voff = (size_t)b * a.va.member_stride;
  double p1[16];

#pragma unroll 1
  for (int q = 0; q < 2; ++q) {
    const int row = 2 * pair + q;
    double2 v[16];
    __syncthreads();  // exchange buffer free (previous transform's readers are done)
    // lower half of the spectrum: k = t + T*e < NX/2.  Z = X + iY with X = i*kr*A (-> gx), Y = B (-> gy);
    // the mirrored bin NX-k gets conj(X) + i*conj(Y) and is handed to its owner through shared memory.
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = t + T * e;
      double2 Av = Ab[(size_t)k * ny + row];
      double2 Bv = Bb[(size_t)k * ny + row];
      const double kx = a.ax.kx[k];
      const double Xx = -kx * Av.y, Xy = kx * Av.x;
      if (e == 0 && t == 0) {
        v[e] = make_double2(Xx, Bv.x);  // c2r ignores the imaginary part of the DC bin
      } else {
        v[e] = make_double2(Xx - Bv.y, Xy + Bv.x);
        sm[pad_idx(NX - k)] = make_double2(Xx + Bv.y, Bv.x - Xy);
      }
    }
    if (t == 0) {  // Nyquist bin k = NX/2: real part only (c2r semantics)
      double2 Av = Ab[(size_t)H * ny + row];
      double2 Bv = Bb[(size_t)H * ny + row];
      const double kx = a.ax.kx[H];
      sm[pad_idx(H)] = make_double2(-kx * Av.y, Bv.x);
    }
    __syncthreads();
#pragma unroll
    for (int e = 8; e < 16; ++e) v[e] = sm[pad_idx(t + T * e)];
    fft::fft_cta<NX, +1>(v, sm, t, a.tw);
    // physical space: v = gx + i*gy at x = t + T*e
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int x = t + T * e;
      double2 g = v[out_slot<NX>(e)];
      double u, vv;
      if (VMODE == 2) {
        u = sep_eval(a.va.sep[0], x, row, 0, NX, ny, 1, 2);
        vv = sep_eval(a.va.sep[1], x, row, 0, NX, ny, 1, 2);
      } else {
        const size_t i = voff + (size_t)row * NX + x;
        u = a.va.arr[0][i];
        vv = a.va.arr[1][i];
        if (VMODE == 1) u += a.va.ushift[b * ny + row];
      }
      double p = -u * g.x - vv * g.y;
      if (q == 0) ps[x] = p;
      else p1[e] = p;
    }
  }
  // forward transform of the row pair packed as p_y + i*p_{y+1}
  double2 v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = make_double2(ps[t + T * e], p1[e]);
  fft::fft_cta<NX, -1>(v, sm, t, a.tw);
  __syncthreads();
#pragma unroll
  for (int e = 8; e < 16; ++e) sm[pad_idx(t + T * e)] = v[out_slot<NX>(e)];
  __syncthreads();
  double2* P0 = a.Px + ((size_t)b * ny + 2 * pair) * a.nkr;
  double2* P1 = P0 + a.nkr;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = t + T * e;
    double2 W = v[out_slot<NX>(e)];
    double2 Wm = (e == 0 && t == 0) ? W : sm[pad_idx(NX - k)];
    P0[k] = make_double2(0.5 * (W.x + Wm.x), 0.5 * (W.y - Wm.y));
    P1[k] = make_double2(0.5 * (W.y + Wm.y), 0.5 * (Wm.x - W.x));
  }
  if (t == 0) {
    double2 W = v[out_slot<NX>(8)];  // index 8*T = NX/2
    P0[H] = make_double2(W.x, 0.0);
    P1[H] = make_double2(W.y, 0.0);
  }
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

bool is_fused_size(int64_t n) { return n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096; }

// twiddle tables of one length, forward sign, rounded from long double
struct TwiddleSet {
  DevBuf<double2> tw2, tw3;
  Twiddles dev() const { return Twiddles{tw2.p, tw3.p}; }
  void build(int N, int64_t* tally) {
    std::vector<double2> h2(15 * 16), h3(4 * 256);
    const long double PI2 = 6.283185307179586476925286766559005768L;
    for (int r = 1; r < 16; ++r)
      for (int k = 0; k < 16; ++k) {
        long double ang = -PI2 * (long double)(r * k) / 256.0L;
        h2[(r - 1) * 16 + k] = make_double2((double)cosl(ang), (double)sinl(ang));
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

template <int N>
constexpr size_t y_smem() { return (size_t)(256 / Cfg<N>::T) * Cfg<N>::PADN * sizeof(double2); }
template <int N>
constexpr size_t x_smem() { return y_smem<N>() + (size_t)(256 / Cfg<N>::T) * N * sizeof(double); }

template <class K>
void allow_smem(K kernel, size_t bytes) {
  PTF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

template <int NY>
void prep_y() {  // opt in to > 48 KB dynamic shared memory (per device, so done at engine construction)
  allow_smem(k_fused_y<NY, false, true>, y_smem<NY>());
  allow_smem(k_fused_y<NY, true, true>, y_smem<NY>());
  allow_smem(k_fused_y<NY, true, false>, y_smem<NY>());
}
template <int NX>
void prep_x() {
  allow_smem(k_fused_x<NX, 0>, x_smem<NX>());
  allow_smem(k_fused_x<NX, 1>, x_smem<NX>());
  allow_smem(k_fused_x<NX, 2>, x_smem<NX>());
}

template <int NY>
void launch_y(bool has_in, bool has_out, const YArgs& a, int nb, cudaStream_t st) {
  constexpr int F = 256 / Cfg<NY>::T;
  dim3 grid((a.nkr + F - 1) / F, nb, 1);
  size_t sm = y_smem<NY>();
  if (!has_in) k_fused_y<NY, false, true><<<grid, 256, sm, st>>>(a);
  else if (has_out) k_fused_y<NY, true, true><<<grid, 256, sm, st>>>(a);
  else k_fused_y<NY, true, false><<<grid, 256, sm, st>>>(a);
}

template <int NX>
void launch_x(int vmode, const XArgs& a, int nb, cudaStream_t st) {
  constexpr int F = 256 / Cfg<NX>::T;
  dim3 grid((a.ny / 2) / F, nb, 1);
  size_t sm = x_smem<NX>();
  if (vmode == 0) k_fused_x<NX, 0><<<grid, 256, sm, st>>>(a);
  else if (vmode == 1) k_fused_x<NX, 1><<<grid, 256, sm, st>>>(a);
  else k_fused_x<NX, 2><<<grid, 256, sm, st>>>(a);
}

#define PTF_DISPATCH_N(n, ...)                                          \
  switch (n) {                                                          \
    case 256: { constexpr int NN = 256; __VA_ARGS__; } break;                  \
    case 512: { constexpr int NN = 512; __VA_ARGS__; } break;                  \
    case 1024: { constexpr int NN = 1024; __VA_ARGS__; } break;                \
    case 2048: { constexpr int NN = 2048; __VA_ARGS__; } break;                \
    case 4096: { constexpr int NN = 4096; __VA_ARGS__; } break;                \
    default: throw Error(PTF_EUNSUPPORTED, "fused engine: unsupported transform length"); \
  }

class FusedEngine final : public Engine {
 public:
  explicit FusedEngine(Context& c) : ctx(c), g(c.g) {
    nx = (int)g.nx;
    ny = (int)g.ny;
    nkr = (int)g.nkr;
    nb = (int)g.B;
    nspec = (size_t)nkr * ny * nb;
    int base = ctx.st.base;
    auto zalloc = [&](DevBuf<double2>& b) {
      b.alloc(nspec, &dev_bytes);
      PTF_CUDA(cudaMemsetAsync(b.p, 0, b.bytes(), ctx.stream));
    };
    zalloc(s0);
    if (base == PTF_STEPPER_RK4 || base == PTF_STEPPER_ETDRK4) zalloc(s1);
    if (base == PTF_STEPPER_ETDRK4) zalloc(s2);
    if (base != PTF_STEPPER_FORWARD_EULER) zalloc(acc);
    if (base == PTF_STEPPER_ETDRK4 || base == PTF_STEPPER_AB3) zalloc(n1);
    zalloc(A);
    zalloc(Bf);
    zalloc(Px);
    if (base == PTF_STEPPER_ETDRK4)
      for (auto* b : {&cE, &cE2, &cZ, &cA, &cB, &cG}) b->alloc((size_t)nkr * ny, &dev_bytes);
    twx.build(nx, &dev_bytes);
    if (ny != nx) twy_own.build(ny, &dev_bytes);
    vs.init(&g, ctx.stream, ctx.d.flow_kind, &dev_bytes);
    PTF_DISPATCH_N(ny, prep_y<NN>());
    PTF_DISPATCH_N(nx, prep_x<NN>());
    // cuFFT only at the set_c!/updatevars! boundary (canonical <-> transposed layout), never on the step path
    long long n[2] = {ny, nx};
    size_t wf = 0, wi = 0;
    PTF_CUFFT(cufftCreate(&plan_fwd));
    PTF_CUFFT(cufftCreate(&plan_inv));
    PTF_CUFFT(cufftSetAutoAllocation(plan_fwd, 0));
    PTF_CUFFT(cufftSetAutoAllocation(plan_inv, 0));
    PTF_CUFFT(cufftMakePlanMany64(plan_fwd, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, nb, &wf));
    PTF_CUFFT(cufftMakePlanMany64(plan_inv, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, nb, &wi));
    work_bytes = wf > wi ? wf : wi;
    // the cuFFT work area aliases P^x (dead whenever a boundary transform runs) when it fits
    if (work_bytes > Px.bytes()) work.alloc(work_bytes, &dev_bytes);
    void* wa = work.p ? (void*)work.p : (void*)Px.p;
    PTF_CUFFT(cufftSetWorkArea(plan_fwd, wa));
    PTF_CUFFT(cufftSetWorkArea(plan_inv, wa));
    PTF_CUFFT(cufftSetStream(plan_fwd, ctx.stream));
    PTF_CUFFT(cufftSetStream(plan_inv, ctx.stream));
    on_dt_changed();
  }

  ~FusedEngine() override {
    drop_graphs();
    if (plan_fwd) cufftDestroy(plan_fwd);
    if (plan_inv) cufftDestroy(plan_inv);
  }

  const char* name() const override { return "fused"; }
  int id() const override { return PTF_ENGINE_FUSED; }
  cudaStream_t stream() const override { return ctx.stream; }

  Twiddles tw_x() const { return twx.dev(); }
  Twiddles tw_y() const { return ny != nx ? twy_own.dev() : twx.dev(); }

  // ---------------- velocities ----------------
  void sync_vel() {
    if (vs.dirty) drop_graphs();
    vs.dirty = false;
  }
  void set_velocity(int comp, const double* host, int64_t count) override { vs.set_array(comp, host, count); sync_vel(); }
  void set_velocity_separable(int comp, int nterms, const double* xt, const double* yt, const double* zt,
                              const double* coeff0) override {
    vs.set_separable(comp, nterms, xt, yt, zt, coeff0);
    sync_vel();
  }
  void set_velocity_coeffs(int comp, int nterms, const double* a) override { vs.set_coeffs(comp, nterms, a); }
  void set_layered_shift(const double* U) override { vs.set_shift(U); sync_vel(); }

  // ---------------- layout changes at the boundary ----------------
  void transpose(const double2* in, double2* out, int R, int Cc, double scale) {
    dim3 grid((Cc + 31) / 32, (R + 31) / 32, nb), block(32, 8, 1);
    k_transpose<<<grid, block, 0, ctx.stream>>>(in, out, R, Cc, scale);
    ++own_launches;
  }

  void set_c(const double* c_host, bool replicate) override {
    int64_t npts = g.npts();
    double* real = reinterpret_cast<double*>(Bf.p);
    if (replicate && nb > 1) {
      PTF_CUDA(cudaMemcpyAsync(real, c_host, npts * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
      k_replicate_f<<<1184, 256, 0, ctx.stream>>>(real, npts, nb);
      ++own_launches;
    } else {
      PTF_CUDA(cudaMemcpyAsync(real, c_host, npts * nb * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    }
    PTF_CUFFT(cufftExecD2Z(plan_fwd, real, reinterpret_cast<cufftDoubleComplex*>(A.p)));  // canonical [b][l][kr]
    ++lib_calls;
    transpose(A.p, s0.p, ny, nkr, 1.0);
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void get_c(double* c_host) override {
    double* real = reinterpret_cast<double*>(Bf.p);
    transpose(s0.p, A.p, nkr, ny, 1.0 / (double)g.npts());
    PTF_CUFFT(cufftExecZ2D(plan_inv, reinterpret_cast<cufftDoubleComplex*>(A.p), real));
    ++lib_calls;
    PTF_CUDA(cudaMemcpyAsync(c_host, real, g.npts() * nb * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void set_sol(const double* s_host) override {
    PTF_CUDA(cudaMemcpyAsync(A.p, s_host, A.bytes(), cudaMemcpyHostToDevice, ctx.stream));
    transpose(A.p, s0.p, ny, nkr, 1.0);
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void get_sol(double* s_host) override {
    transpose(s0.p, A.p, nkr, ny, 1.0);
    PTF_CUDA(cudaMemcpyAsync(s_host, A.p, A.bytes(), cudaMemcpyDeviceToHost, ctx.stream));
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void on_dt_changed() override {
    drop_graphs();
    if (ctx.st.base == PTF_STEPPER_ETDRK4) {
      k_etd_coeffs<<<1184, 256, 0, ctx.stream>>>(cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p, ctx.ax, nkr, ny, 1, ctx.dt, 1);
      ++own_launches;
      PTF_CUDA(cudaGetLastError());
    }
  }

  // ---------------- the two hot kernels ----------------
  YArgs yargs(int mode, double la, double lb, int llast) const {
    YArgs a;
    a.Px = Px.p;
    a.A = A.p;
    a.Bf = Bf.p;
    a.P = CombinePtrs{s0.p, s1.p, s2.p, acc.p, n1.p, cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p};
    a.C = CombineArgs{mode, ctx.st.filtered ? 1 : 0, ctx.dt, la, lb, llast};
    a.ax = ctx.ax;
    a.tw = tw_y();
    a.nkr = nkr;
    a.inv_n = 1.0 / (double)g.npts();
    int slot = mode < 0 ? 0 : next_state_slot(mode);
    a.next_state = slot == 0 ? s0.p : (slot == 1 ? s1.p : s2.p);
    return a;
  }

  void run_y(bool has_in, bool has_out, int mode, double la = 0, double lb = 0, int llast = 0) {
    YArgs a = yargs(mode, la, lb, llast);
    PTF_DISPATCH_N(ny, launch_y<NN>(has_in, has_out, a, nb, ctx.stream));
    ++own_launches;
  }

  void run_x() {
    XArgs a;
    a.A = A.p;
    a.Bf = Bf.p;
    a.Px = Px.p;
    a.va = vs.va;
    a.ax = ctx.ax;
    a.tw = tw_x();
    a.nkr = nkr;
    a.ny = ny;
    int vmode = (vs.va.kind == PTF_FLOW_SEPARABLE) ? 2 : (vs.va.ushift ? 1 : 0);
    if (vmode != 2 && (!vs.va.arr[0] || !vs.va.arr[1]))
      throw Error(PTF_EINVAL, "velocity fields have not been set (ptf_set_velocity / callback)");
    PTF_DISPATCH_N(nx, launch_x<NN>(vmode, a, nb, ctx.stream));
    ++own_launches;
  }

  void stage(int mode, double la = 0, double lb = 0, int llast = 0) {
    run_x();
    run_y(true, true, mode, la, lb, llast);
  }

  void enqueue_step(int variant) {
    static const double LA[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                 -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
    static const double LB[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                                 1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                                 2277821191437.0 / 14882151754819.0};
    switch (ctx.st.base) {
      case PTF_STEPPER_RK4:
        stage(CM_RK4_S1); stage(CM_RK4_S2); stage(CM_RK4_S3); stage(CM_RK4_S4);
        break;
      case PTF_STEPPER_ETDRK4:
        stage(CM_ETD_S1); stage(CM_ETD_S2); stage(CM_ETD_S3); stage(CM_ETD_S4);
        break;
      case PTF_STEPPER_FORWARD_EULER:
        stage(CM_EULER);
        break;
      case PTF_STEPPER_LSRK54:
        for (int i = 0; i < 5; ++i) stage(CM_LSRK, LA[i], LB[i], i == 4);
        break;
      case PTF_STEPPER_AB3:
        stage(variant == 1 ? CM_AB3_EULER : CM_AB3);
        break;
    }
  }

  void step_once(int64_t step_index) override {
    if (!ab_valid) {  // A,B of the current sol are missing (fresh state): one prologue launch
      run_y(false, true, -1);
      ab_valid = true;
    }
    int variant = (ctx.st.base == PTF_STEPPER_AB3 && step_index < 3) ? 1 : 0;
    if (!ctx.d.use_graph) {
      enqueue_step(variant);
      PTF_CUDA(cudaGetLastError());
      return;
    }
    if (!graph_exec[variant]) {
      int64_t o0 = own_launches;
      cudaGraph_t graph = nullptr;
      PTF_CUDA(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal));
      try {
        enqueue_step(variant);
      } catch (...) {
        cudaStreamEndCapture(ctx.stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      PTF_CUDA(cudaStreamEndCapture(ctx.stream, &graph));
      cudaError_t e = cudaGraphInstantiate(&graph_exec[variant], graph, 0);
      cudaGraphDestroy(graph);
      PTF_CUDA(e);
      per_step_own = own_launches - o0;
      own_launches = o0;
    }
    PTF_CUDA(cudaGraphLaunch(graph_exec[variant], ctx.stream));
    own_launches += per_step_own;
  }

  void drop_graphs() {
    for (auto& ge : graph_exec) {
      if (ge) cudaGraphExecDestroy(ge);
      ge = nullptr;
    }
  }

  void diag(double* mean_c, double* var_c, double* max_abs_sol) override {
    DevBuf<double> out;
    out.alloc(2);
    PTF_CUDA(cudaMemsetAsync(out.p, 0, 2 * sizeof(double), ctx.stream));
    k_diag_t<<<1184, 256, 0, ctx.stream>>>(s0.p, nkr, ny, nb, nx, out.p);
    ++own_launches;
    double h[2];
    double2 dc;
    PTF_CUDA(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaMemcpyAsync(&dc, s0.p, sizeof(dc), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    double N = (double)g.npts();
    double m = dc.x / N;
    double msq = h[0] / (N * N) / (double)nb;
    if (mean_c) *mean_c = m;
    if (var_c) *var_c = msq - m * m;
    if (max_abs_sol) *max_abs_sol = std::sqrt(h[1]);
  }

  // Average device time of one hot kernel measured over `reps` REAL steps (events around every launch of that kernel
  // on the step stream).  The solution is backed up and restored, the clock is untouched.
  float time_kernel(const char* kname, int reps) override {
    std::string k(kname ? kname : "");
    bool want_x = (k == "xkernel");
    if (!want_x && k != "ykernel") throw Error(PTF_EINVAL, "fused engine: unknown kernel '" + k + "' (xkernel|ykernel)");
    if (ctx.st.base != PTF_STEPPER_RK4) throw Error(PTF_EUNSUPPORTED, "kernel timing is implemented for RK4 steps");
    DevBuf<double2> backup;
    backup.alloc(nspec);
    PTF_CUDA(cudaMemcpyAsync(backup.p, s0.p, s0.bytes(), cudaMemcpyDeviceToDevice, ctx.stream));
    if (!ab_valid) {
      run_y(false, true, -1);
      ab_valid = true;
    }
    const int modes[4] = {CM_RK4_S1, CM_RK4_S2, CM_RK4_S3, CM_RK4_S4};
    std::vector<cudaEvent_t> ev(2 * 4 * reps);
    for (auto& e : ev) PTF_CUDA(cudaEventCreate(&e));
    auto one_step = [&](int rep, bool record) {
      for (int s = 0; s < 4; ++s) {
        int idx = 2 * (4 * rep + s);
        if (record && want_x) PTF_CUDA(cudaEventRecord(ev[idx], ctx.stream));
        run_x();
        if (record && want_x) PTF_CUDA(cudaEventRecord(ev[idx + 1], ctx.stream));
        if (record && !want_x) PTF_CUDA(cudaEventRecord(ev[idx], ctx.stream));
        run_y(true, true, modes[s]);
        if (record && !want_x) PTF_CUDA(cudaEventRecord(ev[idx + 1], ctx.stream));
      }
    };
    one_step(0, false);  // warm
    for (int r = 0; r < reps; ++r) one_step(r, true);
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    double total = 0;
    for (int i = 0; i < 4 * reps; ++i) {
      float ms = 0;
      PTF_CUDA(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
      total += ms;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    PTF_CUDA(cudaMemcpyAsync(s0.p, backup.p, s0.bytes(), cudaMemcpyDeviceToDevice, ctx.stream));
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    return (float)(total / (4.0 * reps));
  }

 private:
  Context& ctx;
  Geometry& g;
  int nx, ny, nkr, nb;
  size_t nspec;
  DevBuf<double2> s0, s1, s2, acc, n1, A, Bf, Px;
  DevBuf<double> cE, cE2, cZ, cA, cB, cG;
  DevBuf<char> work;
  size_t work_bytes = 0;
  TwiddleSet twx, twy_own;
  VelocityStore vs;
  cufftHandle plan_fwd = 0, plan_inv = 0;
  bool ab_valid = false;
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int64_t per_step_own = 0;
};

}  // namespace

bool fused_engine_supports(const Context& ctx, std::string* why) {
  auto no = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  if (ctx.g.ndim != 2) return no("only 2-D problems (1-D / 3-D run on the cuFFT engine)");
  if (!is_fused_size(ctx.g.nx) || !is_fused_size(ctx.g.ny)) return no("nx and ny must be powers of two in [256, 4096]");
  if (ctx.d.dealias) return no("dealias option is served by the cuFFT engine");
  if (ctx.g.B > 65535) return no("batch too large for one launch");
  return true;
}

std::unique_ptr<Engine> make_fused_engine(Context& ctx) {
  std::string why;
  if (!fused_engine_supports(ctx, &why)) throw Error(PTF_EUNSUPPORTED, "fused engine: " + why);
  return std::unique_ptr<Engine>(new FusedEngine(ctx));
}

// Test hook: run `count` independent length-n complex transforms through the hand-written FFT core.
void selftest_fft(int n, int dir, int count, const double* in_host, double* out_host) {
  PTF_REQUIRE(is_fused_size(n), "selftest_fft: n must be a power of two in [256, 4096]");
  PTF_REQUIRE(dir == 1 || dir == -1, "selftest_fft: dir must be +1 or -1");
  PTF_REQUIRE(count > 0, "selftest_fft: count must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(PTF_ENODEVICE, "no CUDA device visible");
  }
  TwiddleSet tw;
  tw.build(n, nullptr);
  DevBuf<double2> in, out;
  size_t total = (size_t)n * count;
  in.alloc(total);
  out.alloc(total);
  PTF_CUDA(cudaMemcpy(in.p, in_host, total * sizeof(double2), cudaMemcpyHostToDevice));
  PTF_DISPATCH_N(n, {
    constexpr int F = 256 / Cfg<NN>::T;
    size_t sm = y_smem<NN>();
    int blocks = (count + F - 1) / F;
    if (dir < 0) {
      allow_smem(k_fft_test<NN, -1>, sm);
      k_fft_test<NN, -1><<<blocks, 256, sm>>>(in.p, out.p, count, tw.dev());
    } else {
      allow_smem(k_fft_test<NN, +1>, sm);
      k_fft_test<NN, +1><<<blocks, 256, sm>>>(in.p, out.p, count, tw.dev());
    }
  });
  PTF_CUDA(cudaGetLastError());
  PTF_CUDA(cudaDeviceSynchronize());
  PTF_CUDA(cudaMemcpy(out_host, out.p, total * sizeof(double2), cudaMemcpyDeviceToHost));
}

}  // namespace ptf
