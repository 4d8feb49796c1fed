// Generic engine: cuFFT D2Z/Z2D transforms + fused hand-written fp64 pointwise kernels.  Serves every even
// grid size in 1/2/3-D with a batch (layer / ensemble) axis.  Per stage it launches
//   k_deriv    : s -> i*k_a*s/N for all d axes in one pass over s   (TAD.jl:757-758 etc., + 1/N of ldiv!)
//   Z2D x d    : cuFFT                                               (TAD.jl:760-761)
//   k_product  : p = -u*gx - v*gy - w*gz                             (TAD.jl:764)
//   D2Z        : cuFFT                                               (TAD.jl:766)
//   k_combine  : addlinearterm! + substepsol!/update! of the FF steppers in one pass, L/filter in registers
// The whole step is captured in a CUDA graph.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ptf_pointwise.cuh"
#include "ptf_velocity.cuh"
#include "expr_flow.h"

#ifdef PTF_WITH_NCCL
#include <nccl.h>
#endif

namespace ptf {

namespace {

struct SpecShape {
  int64_t nkr, ny, nz, B;
};

// ---------------------------------------------------------------------------------------------------
// spectral derivative:  dh[a] = i * k_a * s * scale   (scale = 1/N folds the inverse-transform normalisation)
// optional dealias!(s) in place first.
// thread layout: blockDim = (TX, TY); each block covers TY spectral rows; threads stride over kx.
// ---------------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(256) k_deriv(double2* __restrict__ s, double2* __restrict__ d0,
                                               double2* __restrict__ d1, double2* __restrict__ d2, AxisTables ax,
                                               SpecShape sh, double scale) {
  int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  int64_t nrows = sh.ny * sh.nz * sh.B;
  if (row >= nrows) return;
  int64_t iy = row % sh.ny;
  int64_t iz = (row / sh.ny) % sh.nz;
  double ky = (ND >= 2) ? ax.ky[iy] * scale : 0.0;
  double kz = (ND >= 3) ? ax.kz[iz] * scale : 0.0;
  int64_t base = row * sh.nkr;
  for (int64_t ix = threadIdx.x; ix < sh.nkr; ix += blockDim.x) {
    double2 v = s[base + ix];
    if (ax.dealias && dealiased_out(ax, ix, iy, iz)) {
      v = make_double2(0.0, 0.0);
      s[base + ix] = v;
    }
    double kx = ax.kx[ix] * scale;
    d0[base + ix] = make_double2(-kx * v.y, kx * v.x);
    if (ND >= 2) d1[base + ix] = make_double2(-ky * v.y, ky * v.x);
    if (ND >= 3) d2[base + ix] = make_double2(-kz * v.y, kz * v.x);
  }
}

// slab path, step 1 of the shared partial inverse transforms: only TWO fields cross the network,
//   d0 = s*scale  and  d2 = i*kz*s*scale   (layout [kz][ky_local][kx]); optional dealias!(s) in place first.
__global__ void __launch_bounds__(256) k_deriv_z(double2* __restrict__ s, double2* __restrict__ d0,
                                                 double2* __restrict__ d2, AxisTables ax, SpecShape sh, double scale) {
  int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  int64_t nrows = sh.ny * sh.nz * sh.B;
  if (row >= nrows) return;
  int64_t iy = row % sh.ny;
  int64_t iz = (row / sh.ny) % sh.nz;
  double kz = ax.kz[iz] * scale;
  int64_t base = row * sh.nkr;
  for (int64_t ix = threadIdx.x; ix < sh.nkr; ix += blockDim.x) {
    double2 v = s[base + ix];
    if (ax.dealias && dealiased_out(ax, ix, iy, iz)) {
      v = make_double2(0.0, 0.0);
      s[base + ix] = v;
    }
    d0[base + ix] = make_double2(v.x * scale, v.y * scale);
    d2[base + ix] = make_double2(-kz * v.y, kz * v.x);
  }
}

// slab path, step 2: after the transpose the field is [z_local][ky (all)][kx]; the x and y derivatives are applied
// here (i*kx in place, i*ky into `dy`), so they never travel.
__global__ void __launch_bounds__(256) k_deriv_xy_planes(double2* __restrict__ f, double2* __restrict__ dy,
                                                         const double* __restrict__ kxt, const double* __restrict__ kyt,
                                                         int64_t nkr, int64_t ny, int64_t nzl) {
  int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= ny * nzl) return;
  const double ky = kyt[row % ny];
  int64_t base = row * nkr;
  for (int64_t ix = threadIdx.x; ix < nkr; ix += blockDim.x) {
    double2 v = f[base + ix];
    double kx = kxt[ix];
    f[base + ix] = make_double2(-kx * v.y, kx * v.x);
    dy[base + ix] = make_double2(-ky * v.y, ky * v.x);
  }
}

// copy with scale (used by get_c: c = irfft(copy(sol)), TAD.jl:816-818)
__global__ void __launch_bounds__(256) k_scale_copy(const double2* __restrict__ in, double2* __restrict__ out,
                                                    int64_t n, double scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double2 v = in[i];
    out[i] = make_double2(v.x * scale, v.y * scale);
  }
}

// replicate member 0 over the batch axis (set_c! for layered problems, TAD.jl:865)
__global__ void __launch_bounds__(256) k_replicate(double* __restrict__ c, int64_t npts, int64_t B) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += (int64_t)gridDim.x * blockDim.x) {
    double v = c[i];
    for (int64_t b = 1; b < B; ++b) c[b * npts + i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// physical-space product  p = -u*gx - v*gy - w*gz  written over gx.   One thread = two x-adjacent points.
// grid.y = member.
// ---------------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(256) k_product(double* __restrict__ g0, const double* __restrict__ g1,
                                                 const double* __restrict__ g2, VelArgs va, int64_t nx, int64_t ny,
                                                 int64_t nz, int64_t nzg, int64_t zoff) {
  // nz = local planes; separable z tables are global, [nterms][nzg], and this rank's first plane is zoff
  int64_t npts = nx * ny * nz;
  int64_t half = npts >> 1;
  int64_t b = blockIdx.y;
  double2* G0 = reinterpret_cast<double2*>(g0 + b * npts);
  const double2* G1 = reinterpret_cast<const double2*>(g1 + b * npts);
  const double2* G2 = reinterpret_cast<const double2*>(g2 + b * npts);
  int64_t voff = b * va.member_stride;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < half; e += (int64_t)gridDim.x * blockDim.x) {
    double2 u, v = make_double2(0, 0), w = make_double2(0, 0);
    if (va.kind == PTF_FLOW_SEPARABLE) {
      int64_t p = e * 2;
      int64_t i = p % nx;
      int64_t j = (p / nx) % ny;
      int64_t k = p / (nx * ny) + zoff;
      u = make_double2(sep_eval(va.sep[0], i, j, k, nx, ny, nzg, ND), sep_eval(va.sep[0], i + 1, j, k, nx, ny, nzg, ND));
      if (ND >= 2)
        v = make_double2(sep_eval(va.sep[1], i, j, k, nx, ny, nzg, ND),
                         sep_eval(va.sep[1], i + 1, j, k, nx, ny, nzg, ND));
      if (ND >= 3)
        w = make_double2(sep_eval(va.sep[2], i, j, k, nx, ny, nzg, ND),
                         sep_eval(va.sep[2], i + 1, j, k, nx, ny, nzg, ND));
    } else {
      u = reinterpret_cast<const double2*>(va.arr[0] + voff)[e];
      if (ND >= 2) v = reinterpret_cast<const double2*>(va.arr[1] + voff)[e];
      if (ND >= 3) w = reinterpret_cast<const double2*>(va.arr[2] + voff)[e];
      if (va.ushift) {  // u + U(y, layer)   (TAD.jl:795)
        int64_t j = ((e * 2) / nx) % ny;
        double U = va.ushift[b * ny + j];
        u.x += U;
        u.y += U;
      }
    }
    double2 a = G0[e];
    double2 p;
    p.x = -u.x * a.x;
    p.y = -u.y * a.y;
    if (ND >= 2) {
      double2 c = G1[e];
      p.x = p.x - v.x * c.x;
      p.y = p.y - v.y * c.y;
    }
    if (ND >= 3) {
      double2 c = G2[e];
      p.x = p.x - w.x * c.x;
      p.y = p.y - w.y * c.y;
    }
    G0[e] = p;
  }
}

// ---------------------------------------------------------------------------------------------------
// stage combine (FF timesteppers.jl), one pass; see CombineMode.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_combine(const double2* __restrict__ Nh, CombinePtrs P, CombineArgs A,
                                                  AxisTables ax, SpecShape sh) {
  int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  int64_t nrows = sh.ny * sh.nz * sh.B;
  if (row >= nrows) return;
  int64_t iy = row % sh.ny;
  int64_t iz = (row / sh.ny) % sh.nz;
  int64_t crow = (iz * sh.ny + iy) * sh.nkr;  // row base in the (batch-shared) coefficient arrays
  double ky = ax.ky[iy], kz = ax.kz[iz];
  int64_t base = row * sh.nkr;
  for (int64_t ix = threadIdx.x; ix < sh.nkr; ix += blockDim.x)
    combine_at(P, A, ax, (size_t)(base + ix), (size_t)(crow + ix), ax.kx[ix], ky, kz, Nh[base + ix]);
}

// diagnostics: sum of w*|s|^2 (Parseval weights), max |s|, over all members
__global__ void __launch_bounds__(256) k_diag(const double2* __restrict__ s, SpecShape sh, int64_t nx,
                                              double* out /*[2]: sumsq, maxabs*/) {
  __shared__ double ssum[256];
  __shared__ double smax[256];
  int64_t n = sh.nkr * sh.ny * sh.nz * sh.B;
  double acc = 0, mx = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t ix = i % sh.nkr;
    double2 v = s[i];
    double a2 = v.x * v.x + v.y * v.y;
    double w = (ix == 0 || ix == nx / 2) ? 1.0 : 2.0;
    acc += w * a2;
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
    // max of non-negative doubles == max of their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(&out[1]), (unsigned long long)__double_as_longlong(smax[0]));
  }
}

// ---- slab transposes (3-D, one process per GPU).  blk = nzl*nyl*nkr complex values go to / come from each peer.
// pack:   T2[r][zl][jl][kx] = T1[zl][r*nyl + jl][kx]       (after the local 2-D r2c, before the all-to-all)
// unpack: T1[zl][s*nyl + jl][kx] = T2[s][zl][jl][kx]       (after the all-to-all, before the local 2-D c2r)
// MODE 0 pack, 1 unpack, 2 unpack fused with the x/y derivatives of the shared partial inverse transform:
//        f[zl][j][kx] = i*kx*v,  dy[zl][j][kx] = i*ky_j*v   with v = T2[s][zl][jl][kx].
// One warp copies RPW consecutive rows of nkr values, RPW loads in flight per thread before the first store.
template <int MODE>
__global__ void __launch_bounds__(256) k_slab_pack(const double2* __restrict__ src, double2* __restrict__ dst,
                                                   double2* __restrict__ dy, const double* __restrict__ kxt,
                                                   const double* __restrict__ kyt, int64_t nkr, int64_t nyl,
                                                   int64_t nzl, int64_t P, int64_t z0, int64_t nzc) {
  constexpr int RPW = 4;
  const int64_t ny = nyl * P;
  const int64_t rows = nzc * ny;  // (zl, j) rows of nkr values, planes z0 .. z0+nzc-1
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row0 = warp * RPW; row0 < rows; row0 += nwarps * RPW) {
    int64_t a[RPW], b[RPW];
    double ky[RPW];
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
      const int64_t row = row0 + q < rows ? row0 + q : rows - 1;
      const int64_t zl = z0 + row / ny, j = row % ny;
      const int64_t r = j / nyl, jl = j % nyl;
      a[q] = (zl * ny + j) * nkr;                // plane layout   [zl][j][kx]
      b[q] = ((r * nzl + zl) * nyl + jl) * nkr;  // peer-block layout [r][zl][jl][kx]
      ky[q] = MODE == 2 ? kyt[j] : 0.0;
    }
    for (int64_t k = lane; k < nkr; k += 32) {
      double2 v[RPW];
#pragma unroll
      for (int q = 0; q < RPW; ++q) v[q] = __ldcg(src + (MODE == 0 ? a[q] : b[q]) + k);
      const double kx = MODE == 2 ? kxt[k] : 0.0;
#pragma unroll
      for (int q = 0; q < RPW; ++q) {
        if (row0 + q >= rows) break;
        if (MODE == 0) {
          __stcg(dst + b[q] + k, v[q]);
        } else if (MODE == 1) {
          __stcg(dst + a[q] + k, v[q]);
        } else {
          __stcg(dst + a[q] + k, make_double2(-kx * v[q].y, kx * v[q].x));
          __stcg(dy + a[q] + k, make_double2(-ky[q] * v[q].y, ky[q] * v[q].x));
        }
      }
    }
  }
}

inline int pow2_at_least(int64_t v, int cap) {
  int p = 1;
  while (p < v && p < cap) p <<= 1;
  return p;
}

// Optional per-phase device timing of the slab path (PTF_SLAB_PROFILE=1; disables graph capture): CUDA events around
// every phase on the step stream, summed and printed when the engine is destroyed.
struct PhaseTimer {
  bool on = false;
  cudaStream_t st = nullptr;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev;
  const char* names[6] = {"fft2d", "pack", "alltoall", "fftz", "pointwise", "other"};
  void begin(int id) {
    if (!on) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    ev.push_back({id, {a, b}});
  }
  void end() {
    if (!on) return;
    cudaEventRecord(ev.back().second.second, st);
  }
  void report(int rank) {
    if (!on || ev.empty()) return;
    cudaStreamSynchronize(st);
    double tot[6] = {0};
    int cnt[6] = {0};
    for (auto& e : ev) {
      float ms = 0;
      cudaEventElapsedTime(&ms, e.second.first, e.second.second);
      tot[e.first] += ms;
      cnt[e.first]++;
      cudaEventDestroy(e.second.first);
      cudaEventDestroy(e.second.second);
    }
    fprintf(stderr, "[slab profile rank %d]", rank);
    for (int i = 0; i < 6; ++i)
      if (cnt[i]) fprintf(stderr, "  %s: %.2f ms in %d calls (%.3f ms each)", names[i], tot[i], cnt[i], tot[i] / cnt[i]);
    fprintf(stderr, "\n");
    ev.clear();
  }
};

class CufftEngine final : public Engine {
 public:
  explicit CufftEngine(Context& c) : ctx(c), g(c.g) {
    nd = g.ndim;
    nspec = g.lspec() * g.B;
    nreal = g.lpts() * g.B;
    axl = ctx.ax;  // tables as this rank sees them: ky starts at its first spectral row
    axl.ky = ctx.ax.ky + g.yoff;
    axl.ay_lo = ctx.ax.ay_lo - g.yoff;
    axl.ay_hi = ctx.ax.ay_hi - g.yoff;
    int base = ctx.st.base;
    sol.alloc(nspec, &dev_bytes);
    PTF_CUDA(cudaMemsetAsync(sol.p, 0, sol.bytes(), ctx.stream));
    if (base == PTF_STEPPER_RK4 || base == PTF_STEPPER_ETDRK4) s1.alloc(nspec, &dev_bytes);
    if (base == PTF_STEPPER_ETDRK4) s2.alloc(nspec, &dev_bytes);
    if (base != PTF_STEPPER_FORWARD_EULER) {
      acc.alloc(nspec, &dev_bytes);
      PTF_CUDA(cudaMemsetAsync(acc.p, 0, acc.bytes(), ctx.stream));
    }
    if (base == PTF_STEPPER_ETDRK4 || base == PTF_STEPPER_AB3) {
      n1.alloc(nspec, &dev_bytes);
      PTF_CUDA(cudaMemsetAsync(n1.p, 0, n1.bytes(), ctx.stream));
    }
    for (int a = 0; a < nd; ++a) {
      dh[a].alloc(nspec, &dev_bytes);
      gr[a].alloc(nreal, &dev_bytes);
    }
    if (base == PTF_STEPPER_ETDRK4) {
      for (auto* b : {&cE, &cE2, &cZ, &cA, &cB, &cG}) b->alloc(g.lspec(), &dev_bytes);
    }
    vs.init(&g, ctx.stream, ctx.d.flow_kind, &dev_bytes);
    pt.st = ctx.stream;
    pt.on = g.slab && std::getenv("PTF_SLAB_PROFILE") != nullptr;
    if (pt.on) ctx.d.use_graph = 0;
    make_plans();
    on_dt_changed();
  }

  ~CufftEngine() override {
    pt.report(g.rank);
    drop_graphs();
    if (plan_fwd) cufftDestroy(plan_fwd);
    if (plan_inv) cufftDestroy(plan_inv);
    if (plan_z) cufftDestroy(plan_z);
    if (plan_fwd_chunk) cufftDestroy(plan_fwd_chunk);
    if (s_comm) cudaStreamDestroy(s_comm);
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }

  const char* name() const override { return "cufft"; }
  int id() const override { return PTF_ENGINE_CUFFT; }
  cudaStream_t stream() const override { return ctx.stream; }

  void make_plans() {
    if (g.slab) {
      make_slab_plans();
      return;
    }
    long long n[3];
    int rank = nd;
    if (nd == 1) { n[0] = g.nx; }
    if (nd == 2) { n[0] = g.ny; n[1] = g.nx; }
    if (nd == 3) { n[0] = g.nz; n[1] = g.ny; n[2] = g.nx; }
    size_t wf = 0, wi = 0;
    PTF_CUFFT(cufftCreate(&plan_fwd));
    PTF_CUFFT(cufftCreate(&plan_inv));
    PTF_CUFFT(cufftSetAutoAllocation(plan_fwd, 0));
    PTF_CUFFT(cufftSetAutoAllocation(plan_inv, 0));
    PTF_CUFFT(cufftMakePlanMany64(plan_fwd, rank, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, g.B, &wf));
    PTF_CUFFT(cufftMakePlanMany64(plan_inv, rank, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, g.B, &wi));
    size_t w = wf > wi ? wf : wi;
    work.alloc(w ? w : 16, &dev_bytes);
    PTF_CUFFT(cufftSetWorkArea(plan_fwd, work.p));
    PTF_CUFFT(cufftSetWorkArea(plan_inv, work.p));
    PTF_CUFFT(cufftSetStream(plan_fwd, ctx.stream));
    PTF_CUFFT(cufftSetStream(plan_inv, ctx.stream));
  }

  // slab decomposition: local batched 2-D (x,y) transforms over the nzl local planes + strided 1-D z transforms
  void make_slab_plans() {
    long long n2[2] = {g.ny, g.nx};
    long long nz1[1] = {g.nz};
    long long stride = g.nyl * g.nkr;
    size_t w[3] = {0, 0, 0};
    PTF_CUFFT(cufftCreate(&plan_fwd));
    PTF_CUFFT(cufftCreate(&plan_inv));
    PTF_CUFFT(cufftCreate(&plan_z));
    for (cufftHandle pl : {plan_fwd, plan_inv, plan_z}) PTF_CUFFT(cufftSetAutoAllocation(pl, 0));
    PTF_CUFFT(cufftMakePlanMany64(plan_fwd, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, g.nzl, &w[0]));
    PTF_CUFFT(cufftMakePlanMany64(plan_inv, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, g.nzl, &w[1]));
    PTF_CUFFT(cufftMakePlanMany64(plan_z, 1, nz1, nz1, stride, 1, nz1, stride, 1, CUFFT_Z2Z, stride, &w[2]));
    size_t wm = std::max(w[0], std::max(w[1], w[2]));
    work.alloc(wm ? wm : 16, &dev_bytes);
    for (cufftHandle pl : {plan_fwd, plan_inv, plan_z}) {
      PTF_CUFFT(cufftSetWorkArea(pl, work.p));
      PTF_CUFFT(cufftSetStream(pl, ctx.stream));
    }
    T1.alloc(nspec, &dev_bytes);
    T2.alloc(nspec, &dev_bytes);
    // pipeline resources
    const char* pe = std::getenv("PTF_SLAB_PIPELINE");
    slab_pipeline = !(pe && std::atoi(pe) == 0) && nd == 3 && !pt.on;
    if (slab_pipeline) {
      n_chunks = 4;
      while (n_chunks > 1 && g.nzl % n_chunks) n_chunks >>= 1;
      if (const char* ce = std::getenv("PTF_SLAB_CHUNKS")) {
        int c = std::atoi(ce);
        if (c >= 1 && c <= 8 && g.nzl % c == 0) n_chunks = c;
      }
      T3.alloc(nspec, &dev_bytes);
      {  // highest priority: the exchange kernels must get SMs while transforms of the other field are running
        int lo = 0, hi = 0;
        PTF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PTF_CUDA(cudaStreamCreateWithPriority(&s_comm, cudaStreamNonBlocking, hi));
      }
      for (auto& e : ev) PTF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      size_t wc = 0;
      PTF_CUFFT(cufftCreate(&plan_fwd_chunk));
      PTF_CUFFT(cufftSetAutoAllocation(plan_fwd_chunk, 0));
      PTF_CUFFT(cufftMakePlanMany64(plan_fwd_chunk, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, g.nzl / n_chunks, &wc));
      if (wc > work.bytes()) {
        work.alloc(wc, &dev_bytes);
        for (cufftHandle pl : {plan_fwd, plan_inv, plan_z}) PTF_CUFFT(cufftSetWorkArea(pl, work.p));
      }
      PTF_CUFFT(cufftSetWorkArea(plan_fwd_chunk, work.p));
      PTF_CUFFT(cufftSetStream(plan_fwd_chunk, ctx.stream));
    }
  }

  // Exchange of planes [z0, z0+nzc) of every peer block (the whole block by default) on stream `st`.
  void all_to_all(const double2* send, double2* recv, cudaStream_t st = nullptr, int64_t z0 = 0, int64_t nzc = -1) {
#ifdef PTF_WITH_NCCL
    if (!st) st = ctx.stream;
    if (nzc < 0) nzc = g.nzl;
    ncclComm_t comm = (ncclComm_t)ctx.nccl_comm;
    const size_t blk = (size_t)g.nzl * g.nyl * g.nkr;  // complex values per peer
    const size_t off = (size_t)z0 * g.nyl * g.nkr, cnt = (size_t)nzc * g.nyl * g.nkr;
    auto ck = [](ncclResult_t r, const char* what) {
      if (r != ncclSuccess) throw Error(PTF_ENCCL, std::string(what) + ": " + ncclGetErrorString(r));
    };
    // own block: plain device copy; the P-1 peer blocks: one grouped send/recv each over NVLink
    PTF_CUDA(cudaMemcpyAsync(recv + (size_t)g.rank * blk + off, send + (size_t)g.rank * blk + off,
                             cnt * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    ck(ncclGroupStart(), "ncclGroupStart");
    for (int r = 0; r < g.P; ++r) {
      if (r == g.rank) continue;
      ck(ncclSend(send + (size_t)r * blk + off, 2 * cnt, ncclDouble, r, comm, st), "ncclSend");
      ck(ncclRecv(recv + (size_t)r * blk + off, 2 * cnt, ncclDouble, r, comm, st), "ncclRecv");
    }
    ck(ncclGroupEnd(), "ncclGroupEnd");
    ++lib_calls;
#else
    (void)send; (void)recv;
    throw Error(PTF_EUNSUPPORTED, "built without NCCL");
#endif
  }

  // forward r2c of this rank's physical slab -> its spectral slab  (unnormalised, FF mul!(., rfftplan, .))
  void fwd(double* real, double2* spec) {
    if (!g.slab) {
      PTF_CUFFT(cufftExecD2Z(plan_fwd, real, reinterpret_cast<cufftDoubleComplex*>(spec)));
      ++lib_calls;
      return;
    }
    pt.begin(0);
    PTF_CUFFT(cufftExecD2Z(plan_fwd, real, reinterpret_cast<cufftDoubleComplex*>(T1.p)));
    pt.end();
    pt.begin(1);
    k_slab_pack<0><<<2368, 256, 0, ctx.stream>>>(T1.p, T2.p, nullptr, nullptr, nullptr, g.nkr, g.nyl, g.nzl, g.P, 0, g.nzl);
    pt.end();
    ++own_launches;
    pt.begin(2);
    all_to_all(T2.p, spec);  // received blocks [s][nzl][nyl][nkr] are exactly [nz][nyl][nkr]
    pt.end();
    pt.begin(3);
    PTF_CUFFT(cufftExecZ2Z(plan_z, reinterpret_cast<cufftDoubleComplex*>(spec),
                           reinterpret_cast<cufftDoubleComplex*>(spec), CUFFT_FORWARD));
    pt.end();
    lib_calls += 2;
  }

  // inverse c2r (input destroyed), x-axis last as in FFTW/cuFFT c2r; the caller folds in 1/N
  void inv(double2* spec, double* real) {
    if (!g.slab) {
      PTF_CUFFT(cufftExecZ2D(plan_inv, reinterpret_cast<cufftDoubleComplex*>(spec), real));
      ++lib_calls;
      return;
    }
    pt.begin(3);
    PTF_CUFFT(cufftExecZ2Z(plan_z, reinterpret_cast<cufftDoubleComplex*>(spec),
                           reinterpret_cast<cufftDoubleComplex*>(spec), CUFFT_INVERSE));
    pt.end();
    pt.begin(2);
    all_to_all(spec, T2.p);  // chunk r of [nz][nyl][nkr] is the z-range of rank r: no packing on the send side
    pt.end();
    pt.begin(1);
    k_slab_pack<1><<<2368, 256, 0, ctx.stream>>>(T2.p, T1.p, nullptr, nullptr, nullptr, g.nkr, g.nyl, g.nzl, g.P, 0, g.nzl);
    pt.end();
    ++own_launches;
    pt.begin(0);
    PTF_CUFFT(cufftExecZ2D(plan_inv, reinterpret_cast<cufftDoubleComplex*>(T1.p), real));
    pt.end();
    lib_calls += 2;
  }

  SpecShape shape() const { return SpecShape{g.nkr, g.nyl, g.nz, g.B}; }

  void spec_launch_dims(dim3& grid, dim3& block) const {
    int tx = pow2_at_least(g.nkr, 256);
    int ty = 256 / tx;
    int64_t rows = g.nyl * g.nz * g.B;
    block = dim3(tx, ty, 1);
    grid = dim3((unsigned)((rows + ty - 1) / ty), 1, 1);
  }

  int flat_blocks(int64_t n) const {
    int64_t b = (n + 255) / 256;
    int64_t cap = 148 * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
  }

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
  void set_velocity_external(int comp, const double* dev, int64_t count) override {
    vs.set_external(comp, dev, count);
    sync_vel();
  }
  void set_velocity_expr(int comp, const char* expr) override {
    ef.set(comp, expr);
    bool all = true;
    for (int a = 0; a < nd; ++a) all = all && !ef.expr[a].empty();
    if (all && ef.stale) {
      PTF_CUDA(cudaStreamSynchronize(ctx.stream));
      drop_graphs();   // before the old module (referenced by captured kernel nodes) is unloaded
      ef.compile(nd);
    }
  }
  void set_flow_time(double t) override { ef.set_time(t, ctx.stream); }

  // ---------------- state ----------------
  void set_c(const double* c_host, bool replicate) override {
    int64_t npts = g.lpts();
    if (replicate && g.B > 1) {
      PTF_CUDA(cudaMemcpyAsync(gr[0].p, c_host, npts * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
      k_replicate<<<flat_blocks(npts), 256, 0, ctx.stream>>>(gr[0].p, npts, g.B);
      ++own_launches;
    } else {
      PTF_CUDA(cudaMemcpyAsync(gr[0].p, c_host, nreal * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    }
    fwd(gr[0].p, sol.p);
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void get_c(double* c_host) override {
    double scale = 1.0 / (double)g.npts();
    k_scale_copy<<<flat_blocks(nspec), 256, 0, ctx.stream>>>(sol.p, dh[0].p, nspec, scale);
    ++own_launches;
    inv(dh[0].p, gr[0].p);
    PTF_CUDA(cudaMemcpyAsync(c_host, gr[0].p, nreal * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void set_sol(const double* s_host) override {
    PTF_CUDA(cudaMemcpyAsync(sol.p, s_host, sol.bytes(), cudaMemcpyHostToDevice, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  void get_sol(double* s_host) override {
    PTF_CUDA(cudaMemcpyAsync(s_host, sol.p, sol.bytes(), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void on_dt_changed() override {
    drop_graphs();
    if (ctx.st.base == PTF_STEPPER_ETDRK4) {
      k_etd_coeffs<<<flat_blocks(g.lspec()), 256, 0, ctx.stream>>>(cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p, axl, g.nkr,
                                                                    g.nyl, g.nz, ctx.dt, 0);
      ++own_launches;
      PTF_CUDA(cudaGetLastError());
    }
  }

  // ---------------- slab path: one stage with shared partial transforms + comm/compute overlap ----------------
  // inverse: only s and i*kz*s are z-transformed and transposed (2 all-to-alls); the x/y derivatives are applied in
  // the plane layout afterwards.  The all-to-alls run on `s_comm` and overlap the other field's transforms.
  // forward: the 2-D r2c + pack run in NCH plane chunks, each chunk's exchange overlapping the next chunk's FFT.
  void fork_comm_after(cudaEvent_t e) {  // s_comm waits for everything enqueued on the main stream so far
    PTF_CUDA(cudaEventRecord(e, ctx.stream));
    PTF_CUDA(cudaStreamWaitEvent(s_comm, e, 0));
  }
  void join_comm(cudaEvent_t e) {        // the main stream waits for everything enqueued on s_comm so far
    PTF_CUDA(cudaEventRecord(e, s_comm));
    PTF_CUDA(cudaStreamWaitEvent(ctx.stream, e, 0));
  }

  void calcN_slab(double2* ss) {
    dim3 grid, block;
    spec_launch_dims(grid, block);
    const double scale = 1.0 / (double)g.npts();
    auto Z = [](double2* p) { return reinterpret_cast<cufftDoubleComplex*>(p); };
    k_deriv_z<<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, dh[2].p, axl, shape(), scale);
    ++own_launches;
    // field 0 (s): z transform, then its exchange starts on the comm stream
    PTF_CUFFT(cufftExecZ2Z(plan_z, Z(dh[0].p), Z(dh[0].p), CUFFT_INVERSE));
    fork_comm_after(ev[0]);
    all_to_all(dh[0].p, T2.p, s_comm);
    PTF_CUDA(cudaEventRecord(ev[1], s_comm));
    // field 2 (i*kz*s): z transform overlaps exchange 0; its exchange then overlaps the plane work of field 0
    PTF_CUFFT(cufftExecZ2Z(plan_z, Z(dh[2].p), Z(dh[2].p), CUFFT_INVERSE));
    fork_comm_after(ev[2]);
    all_to_all(dh[2].p, T3.p, s_comm);
    PTF_CUDA(cudaEventRecord(ev[3], s_comm));
    // field 0 arrived: unpack, x/y derivatives in the plane layout, two batched 2-D c2r
    PTF_CUDA(cudaStreamWaitEvent(ctx.stream, ev[1], 0));
    // unpack fused with the x/y derivatives (applied in the plane layout, so they never travel)
    k_slab_pack<2><<<2368, 256, 0, ctx.stream>>>(T2.p, T1.p, dh[1].p, ctx.ax.kx, ctx.ax.ky, g.nkr, g.nyl, g.nzl, g.P, 0, g.nzl);
    own_launches += 1;
    PTF_CUFFT(cufftExecZ2D(plan_inv, Z(T1.p), gr[0].p));
    PTF_CUFFT(cufftExecZ2D(plan_inv, Z(dh[1].p), gr[1].p));
    // field 2 arrived
    PTF_CUDA(cudaStreamWaitEvent(ctx.stream, ev[3], 0));
    k_slab_pack<1><<<2368, 256, 0, ctx.stream>>>(T3.p, T1.p, nullptr, nullptr, nullptr, g.nkr, g.nyl, g.nzl, g.P, 0, g.nzl);
    ++own_launches;
    PTF_CUFFT(cufftExecZ2D(plan_inv, Z(T1.p), gr[2].p));
    lib_calls += 7;
    // physical-space product
    int64_t half = g.lpts() / 2;
    dim3 pg((unsigned)flat_blocks(half), (unsigned)g.B, 1);
    VelArgs va = vs.va;
    if (va.kind == PTF_FLOW_EXPR)
      ef.launch(ctx.stream, (int)pg.x, (int)g.B, gr[0].p, gr[1].p, gr[2].p, g.nx, g.ny, g.nzl, 0, g.zoff, g);
    else
      k_product<3><<<pg, 256, 0, ctx.stream>>>(gr[0].p, gr[1].p, gr[2].p, va, g.nx, g.ny, g.nzl, g.nz, g.zoff);
    ++own_launches;
    // forward: chunked 2-D r2c + pack, exchanges pipelined on the comm stream, then the z transform
    const int nch = n_chunks;
    const int64_t nzc = g.nzl / nch;
    for (int c = 0; c < nch; ++c) {
      const int64_t z0 = c * nzc;
      PTF_CUFFT(cufftExecD2Z(plan_fwd_chunk, gr[0].p + z0 * g.nx * g.ny, Z(T1.p + z0 * g.ny * g.nkr)));
      k_slab_pack<0><<<2368, 256, 0, ctx.stream>>>(T1.p, T2.p, nullptr, nullptr, nullptr, g.nkr, g.nyl, g.nzl, g.P, z0, nzc);
      ++own_launches;
      fork_comm_after(ev[4 + c]);
      all_to_all(T2.p, dh[0].p, s_comm, z0, nzc);
    }
    join_comm(ev[4 + nch]);
    PTF_CUFFT(cufftExecZ2Z(plan_z, Z(dh[0].p), Z(dh[0].p), CUFFT_FORWARD));
    lib_calls += nch + 1;
  }

  // ---------------- one stage: N-hat(ss) into dh[0] ----------------
  void calcN(double2* ss) {
    if (g.slab && slab_pipeline) {
      calcN_slab(ss);
      return;
    }
    dim3 grid, block;
    spec_launch_dims(grid, block);
    double scale = 1.0 / (double)g.npts();
    SpecShape sh = shape();
    if (nd == 1)
      k_deriv<1><<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, nullptr, nullptr, axl, sh, scale);
    else if (nd == 2)
      k_deriv<2><<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, dh[1].p, nullptr, axl, sh, scale);
    else
      k_deriv<3><<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, dh[1].p, dh[2].p, axl, sh, scale);
    ++own_launches;
    for (int a = 0; a < nd; ++a) inv(dh[a].p, gr[a].p);
    pt.begin(4);
    int64_t half = g.lpts() / 2;
    VelArgs va = vs.va;
    if (va.kind != PTF_FLOW_SEPARABLE && va.kind != PTF_FLOW_EXPR)
      for (int c = 0; c < nd; ++c)
        if (!va.arr[c]) throw Error(PTF_EINVAL, "velocity fields have not been set (ptf_set_velocity / callback)");
    const int64_t zoff = g.slab ? g.zoff : 0;   // separable z tables are global: [nterms][nz], plane zoff + local k
    dim3 pg((unsigned)flat_blocks(half), (unsigned)g.B, 1);
    const double* g1 = nd >= 2 ? gr[1].p : gr[0].p;
    const double* g2 = nd >= 3 ? gr[2].p : gr[0].p;
    if (va.kind == PTF_FLOW_EXPR)
      ef.launch(ctx.stream, (int)pg.x, (int)g.B, gr[0].p, g1, g2, g.nx, g.ny, g.nzl, 0, g.slab ? g.zoff : 0, g);
    else if (nd == 1)
      k_product<1><<<pg, 256, 0, ctx.stream>>>(gr[0].p, g1, g2, va, g.nx, g.ny, g.nzl, g.nz, zoff);
    else if (nd == 2)
      k_product<2><<<pg, 256, 0, ctx.stream>>>(gr[0].p, g1, g2, va, g.nx, g.ny, g.nzl, g.nz, zoff);
    else
      k_product<3><<<pg, 256, 0, ctx.stream>>>(gr[0].p, g1, g2, va, g.nx, g.ny, g.nzl, g.nz, zoff);
    ++own_launches;
    pt.end();
    fwd(gr[0].p, dh[0].p);
  }

  void combine(int mode, double lsrk_a = 0, double lsrk_b = 0, int lsrk_last = 0) {
    dim3 grid, block;
    spec_launch_dims(grid, block);
    CombinePtrs P{sol.p, s1.p, s2.p, acc.p, n1.p, cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p};
    CombineArgs A{mode, ctx.st.filtered ? 1 : 0, ctx.dt, lsrk_a, lsrk_b, lsrk_last};
    k_combine<<<grid, block, 0, ctx.stream>>>(dh[0].p, P, A, axl, shape());
    ++own_launches;
  }

  void enqueue_step(int variant) {
    static const double LA[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                 -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
    static const double LB[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                                 1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                                 2277821191437.0 / 14882151754819.0};
    switch (ctx.st.base) {
      case PTF_STEPPER_RK4:
        calcN(sol.p); combine(CM_RK4_S1);
        calcN(s1.p);  combine(CM_RK4_S2);
        calcN(s1.p);  combine(CM_RK4_S3);
        calcN(s1.p);  combine(CM_RK4_S4);
        break;
      case PTF_STEPPER_ETDRK4:
        calcN(sol.p); combine(CM_ETD_S1);
        calcN(s1.p);  combine(CM_ETD_S2);
        calcN(s2.p);  combine(CM_ETD_S3);
        calcN(s2.p);  combine(CM_ETD_S4);
        break;
      case PTF_STEPPER_FORWARD_EULER:
        calcN(sol.p); combine(CM_EULER);
        break;
      case PTF_STEPPER_LSRK54:
        for (int i = 0; i < 5; ++i) { calcN(sol.p); combine(CM_LSRK, LA[i], LB[i], i == 4); }
        break;
      case PTF_STEPPER_AB3:
        calcN(sol.p); combine(variant == 1 ? CM_AB3_EULER : CM_AB3);
        break;
    }
  }

  void step_once(int64_t step_index) override {
    int variant = (ctx.st.base == PTF_STEPPER_AB3 && step_index < 3) ? 1 : 0;
    if (!ctx.d.use_graph) {
      enqueue_step(variant);
      PTF_CUDA(cudaGetLastError());
      return;
    }
    if (!graph_exec[variant]) {
      int64_t o0 = own_launches, l0 = lib_calls;
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
      per_step_lib = lib_calls - l0;
      own_launches = o0;
      lib_calls = l0;
    }
    PTF_CUDA(cudaGraphLaunch(graph_exec[variant], ctx.stream));
    own_launches += per_step_own;
    lib_calls += per_step_lib;
  }

  void drop_graphs() {
    for (auto& ge : graph_exec) {
      if (ge) cudaGraphExecDestroy(ge);
      ge = nullptr;
    }
  }

  void diag(double* mean_c, double* var_c, double* max_abs_sol) override {
    DevBuf<double> out;
    out.alloc(4);
    PTF_CUDA(cudaMemsetAsync(out.p, 0, 4 * sizeof(double), ctx.stream));
    k_diag<<<flat_blocks(nspec), 256, 0, ctx.stream>>>(sol.p, shape(), g.nx, out.p);
    ++own_launches;
    PTF_CUDA(cudaMemcpyAsync(out.p + 2, sol.p, sizeof(double2), cudaMemcpyDeviceToDevice, ctx.stream));  // DC mode
#ifdef PTF_WITH_NCCL
    if (g.slab) {  // spectral rows are spread over the ranks; the DC mode lives on rank 0
      ncclComm_t comm = (ncclComm_t)ctx.nccl_comm;
      ncclAllReduce(out.p, out.p, 1, ncclDouble, ncclSum, comm, ctx.stream);
      ncclAllReduce(out.p + 1, out.p + 1, 1, ncclDouble, ncclMax, comm, ctx.stream);
      ncclBroadcast(out.p + 2, out.p + 2, 2, ncclDouble, 0, comm, ctx.stream);
    }
#endif
    double h[4];
    PTF_CUDA(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    double N = (double)g.npts();
    // member-averaged second moment via Parseval; mean from member 0's DC mode
    double m = h[2] / N;
    double msq = h[0] / (N * N) / (double)g.B;
    if (mean_c) *mean_c = m;
    if (var_c) *var_c = msq - m * m;
    if (max_abs_sol) *max_abs_sol = std::sqrt(h[1]);
  }

  float time_kernel(const char* kname, int reps) override {
    std::string k(kname ? kname : "");
    cudaEvent_t e0, e1;
    PTF_CUDA(cudaEventCreate(&e0));
    PTF_CUDA(cudaEventCreate(&e1));
    auto run = [&]() {
      if (k == "deriv") {
        dim3 grid, block;
        spec_launch_dims(grid, block);
        double scale = 1.0 / (double)g.npts();
        // runs on the scratch stage buffer so sol is untouched (dealias writes in place)
        double2* ss = s1.p ? s1.p : dh[0].p;
        if (nd == 1) k_deriv<1><<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, nullptr, nullptr, ctx.ax, shape(), scale);
        else if (nd == 2) k_deriv<2><<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, dh[1].p, nullptr, ctx.ax, shape(), scale);
        else k_deriv<3><<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, dh[1].p, dh[2].p, ctx.ax, shape(), scale);
      } else if (k == "z2d") {
        cufftExecZ2D(plan_inv, reinterpret_cast<cufftDoubleComplex*>(dh[0].p), gr[0].p);
      } else if (k == "d2z") {
        cufftExecD2Z(plan_fwd, gr[0].p, reinterpret_cast<cufftDoubleComplex*>(dh[0].p));
      } else {
        throw Error(PTF_EINVAL, "cufft engine: unknown kernel name '" + k + "' (deriv|z2d|d2z)");
      }
    };
    run();
    PTF_CUDA(cudaEventRecord(e0, ctx.stream));
    for (int i = 0; i < reps; ++i) run();
    PTF_CUDA(cudaEventRecord(e1, ctx.stream));
    PTF_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PTF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms / reps;
  }

 private:
  Context& ctx;
  Geometry& g;
  int nd;
  int64_t nspec, nreal;
  DevBuf<double2> sol, s1, s2, acc, n1, dh[3];
  DevBuf<double> gr[3];
  VelocityStore vs;
  ExprFlow ef;
  DevBuf<double> cE, cE2, cZ, cA, cB, cG;
  DevBuf<char> work;
  PhaseTimer pt;
  cufftHandle plan_fwd = 0, plan_inv = 0, plan_z = 0;
  DevBuf<double2> T1, T2, T3;  // slab transposes
  cufftHandle plan_fwd_chunk = 0;
  cudaStream_t s_comm = nullptr;
  cudaEvent_t ev[16] = {nullptr};
  bool slab_pipeline = false;
  int n_chunks = 1;
  AxisTables axl;
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int64_t per_step_own = 0, per_step_lib = 0;
};

}  // namespace

std::unique_ptr<Engine> make_cufft_engine(Context& ctx) { return std::unique_ptr<Engine>(new CufftEngine(ctx)); }

}  // namespace ptf
