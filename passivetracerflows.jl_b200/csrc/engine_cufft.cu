// Generic engine: cuFFT D2Z/Z2D transforms + fused hand-written fp64 pointwise kernels.  Serves every even
// grid size in 1/2/3-D with a batch (layer / ensemble) axis.  Per stage it launches
//   k_deriv    : s -> i*k_a*s/N for all d axes in one pass over s   (TAD.jl:757-758 etc., + 1/N of ldiv!)
//   Z2D x d    : cuFFT                                               (TAD.jl:760-761)
//   k_product  : p = -u*gx - v*gy - w*gz                             (TAD.jl:764)
//   D2Z        : cuFFT                                               (TAD.jl:766)
//   k_combine  : addlinearterm! + substepsol!/update! of the FF steppers in one pass, L/filter in registers
// The whole step is captured in a CUDA graph.
#include <cmath>
#include <cstring>

#include "ptf_pointwise.cuh"
#include "ptf_velocity.cuh"

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
                                                 int64_t nz) {
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
      int64_t k = p / (nx * ny);
      u = make_double2(sep_eval(va.sep[0], i, j, k, nx, ny, nz, ND), sep_eval(va.sep[0], i + 1, j, k, nx, ny, nz, ND));
      if (ND >= 2)
        v = make_double2(sep_eval(va.sep[1], i, j, k, nx, ny, nz, ND),
                         sep_eval(va.sep[1], i + 1, j, k, nx, ny, nz, ND));
      if (ND >= 3)
        w = make_double2(sep_eval(va.sep[2], i, j, k, nx, ny, nz, ND),
                         sep_eval(va.sep[2], i + 1, j, k, nx, ny, nz, ND));
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

inline int pow2_at_least(int64_t v, int cap) {
  int p = 1;
  while (p < v && p < cap) p <<= 1;
  return p;
}

class CufftEngine final : public Engine {
 public:
  explicit CufftEngine(Context& c) : ctx(c), g(c.g) {
    nd = g.ndim;
    nspec = g.nspec() * g.B;
    nreal = g.npts() * g.B;
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
      for (auto* b : {&cE, &cE2, &cZ, &cA, &cB, &cG}) b->alloc(g.nspec(), &dev_bytes);
    }
    vs.init(&g, ctx.stream, ctx.d.flow_kind, &dev_bytes);
    make_plans();
    on_dt_changed();
  }

  ~CufftEngine() override {
    drop_graphs();
    if (plan_fwd) cufftDestroy(plan_fwd);
    if (plan_inv) cufftDestroy(plan_inv);
  }

  const char* name() const override { return "cufft"; }
  int id() const override { return PTF_ENGINE_CUFFT; }
  cudaStream_t stream() const override { return ctx.stream; }

  void make_plans() {
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

  SpecShape shape() const { return SpecShape{g.nkr, g.ny, g.nz, g.B}; }

  void spec_launch_dims(dim3& grid, dim3& block) const {
    int tx = pow2_at_least(g.nkr, 256);
    int ty = 256 / tx;
    int64_t rows = g.ny * g.nz * g.B;
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

  // ---------------- state ----------------
  void set_c(const double* c_host, bool replicate) override {
    int64_t npts = g.npts();
    if (replicate && g.B > 1) {
      PTF_CUDA(cudaMemcpyAsync(gr[0].p, c_host, npts * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
      k_replicate<<<flat_blocks(npts), 256, 0, ctx.stream>>>(gr[0].p, npts, g.B);
      ++own_launches;
    } else {
      PTF_CUDA(cudaMemcpyAsync(gr[0].p, c_host, nreal * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    }
    PTF_CUFFT(cufftExecD2Z(plan_fwd, gr[0].p, reinterpret_cast<cufftDoubleComplex*>(sol.p)));
    ++lib_calls;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void get_c(double* c_host) override {
    double scale = 1.0 / (double)g.npts();
    k_scale_copy<<<flat_blocks(nspec), 256, 0, ctx.stream>>>(sol.p, dh[0].p, nspec, scale);
    ++own_launches;
    PTF_CUFFT(cufftExecZ2D(plan_inv, reinterpret_cast<cufftDoubleComplex*>(dh[0].p), gr[0].p));
    ++lib_calls;
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
      k_etd_coeffs<<<flat_blocks(g.nspec()), 256, 0, ctx.stream>>>(cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p, ctx.ax, g.nkr,
                                                                    g.ny, g.nz, ctx.dt, 0);
      ++own_launches;
      PTF_CUDA(cudaGetLastError());
    }
  }

  // ---------------- one stage: N-hat(ss) into dh[0] ----------------
  void calcN(double2* ss) {
    dim3 grid, block;
    spec_launch_dims(grid, block);
    double scale = 1.0 / (double)g.npts();
    SpecShape sh = shape();
    if (nd == 1)
      k_deriv<1><<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, nullptr, nullptr, ctx.ax, sh, scale);
    else if (nd == 2)
      k_deriv<2><<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, dh[1].p, nullptr, ctx.ax, sh, scale);
    else
      k_deriv<3><<<grid, block, 0, ctx.stream>>>(ss, dh[0].p, dh[1].p, dh[2].p, ctx.ax, sh, scale);
    ++own_launches;
    for (int a = 0; a < nd; ++a) {
      PTF_CUFFT(cufftExecZ2D(plan_inv, reinterpret_cast<cufftDoubleComplex*>(dh[a].p), gr[a].p));
      ++lib_calls;
    }
    int64_t half = g.npts() / 2;
    dim3 pg((unsigned)flat_blocks(half), (unsigned)g.B, 1);
    const double* g1 = nd >= 2 ? gr[1].p : gr[0].p;
    const double* g2 = nd >= 3 ? gr[2].p : gr[0].p;
    if (nd == 1)
      k_product<1><<<pg, 256, 0, ctx.stream>>>(gr[0].p, g1, g2, vs.va, g.nx, g.ny, g.nz);
    else if (nd == 2)
      k_product<2><<<pg, 256, 0, ctx.stream>>>(gr[0].p, g1, g2, vs.va, g.nx, g.ny, g.nz);
    else
      k_product<3><<<pg, 256, 0, ctx.stream>>>(gr[0].p, g1, g2, vs.va, g.nx, g.ny, g.nz);
    ++own_launches;
    PTF_CUFFT(cufftExecD2Z(plan_fwd, gr[0].p, reinterpret_cast<cufftDoubleComplex*>(dh[0].p)));
    ++lib_calls;
  }

  void combine(int mode, double lsrk_a = 0, double lsrk_b = 0, int lsrk_last = 0) {
    dim3 grid, block;
    spec_launch_dims(grid, block);
    CombinePtrs P{sol.p, s1.p, s2.p, acc.p, n1.p, cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p};
    CombineArgs A{mode, ctx.st.filtered ? 1 : 0, ctx.dt, lsrk_a, lsrk_b, lsrk_last};
    k_combine<<<grid, block, 0, ctx.stream>>>(dh[0].p, P, A, ctx.ax, shape());
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
    out.alloc(2);
    PTF_CUDA(cudaMemsetAsync(out.p, 0, 2 * sizeof(double), ctx.stream));
    k_diag<<<flat_blocks(nspec), 256, 0, ctx.stream>>>(sol.p, shape(), g.nx, out.p);
    ++own_launches;
    double h[2];
    double2 dc;
    PTF_CUDA(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaMemcpyAsync(&dc, sol.p, sizeof(dc), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    double N = (double)g.npts();
    // member-averaged second moment via Parseval; mean from member 0's DC mode
    double m = dc.x / N;
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
  DevBuf<double> cE, cE2, cZ, cA, cB, cG;
  DevBuf<char> work;
  cufftHandle plan_fwd = 0, plan_inv = 0;
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int64_t per_step_own = 0, per_step_lib = 0;
};

}  // namespace

std::unique_ptr<Engine> make_cufft_engine(Context& ctx) { return std::unique_ptr<Engine>(new CufftEngine(ctx)); }

}  // namespace ptf
