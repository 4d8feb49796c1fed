// Fused 1-D engine: a whole time step — every stage's spectral derivative, c2r, product, r2c and stage combine — or
// MANY time steps run in ONE launch by ONE CTA per member, with a shared-memory Stockham FFT.  1-D problems are pure
// latency (BASELINE configs[0]: nx = 128, 1 KB of state): the cuFFT engine needs 12 launches per RK4 step, this
// engine needs one launch per stepforward!(prob, nsteps) call when the velocity does not change between steps.
//
// Reference functions covered: calcN! for OneDGrid (TAD.jl:695-706, 744-754), L (TAD.jl:502-509, 529-536), the FF
// steppers (shared combine_at), set_c!/updatevars! (TAD.jl:815-852).
#include <cmath>
#include <cstring>

#include "ptf_pointwise.cuh"
#include "ptf_velocity.cuh"

namespace ptf {

namespace {

struct Args1D {
  CombinePtrs P;
  CombineArgs C;        // dt, filtered (mode / lsrk fields are set per stage inside the kernel)
  AxisTables ax;
  const double* u;      // velocity array (nx per member, or shared) or nullptr for a separable flow
  SepFlow sep;
  int64_t u_stride;     // 0: shared by all members
  int base;             // PTF_STEPPER_*
  int nx, nkr;
  long long step0;      // clock.step of the first step (AB3 start-up)
  int nsteps;
  double inv_n;
  int state_in_smem;    // 1: sol, stage states, accumulators and ETD coefficients live in shared memory for the whole launch
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Stockham autosort radix-2 FFT of length n in shared memory (two ping-pong buffers); tw[m] = exp(-2*pi*i*m/n),
// m < n/2.  dir = -1 forward, +1 unnormalised inverse.  Returns the buffer that holds the result.  All threads of
// the CTA must call it.
__device__ double2* fft_smem(double2* src, double2* dst, int n, int dir, const double2* __restrict__ tw) {
  const int half = n >> 1;
  for (int ns = 1; ns < n; ns <<= 1) {
    const int tstride = half / ns;  // w_{2ns}^k = tw[k * n/(2ns)]
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
      const int k = j & (ns - 1);
      double2 w = tw[k * tstride];
      if (dir > 0) w.y = -w.y;
      const double2 x0 = src[j];
      const double2 x1 = cmul(src[j + half], w);
      const int i0 = ((j - k) << 1) + k;
      dst[i0] = make_double2(x0.x + x1.x, x0.y + x1.y);
      dst[i0 + ns] = make_double2(x0.x - x1.x, x0.y - x1.y);
    }
    __syncthreads();
    double2* t = src;
    src = dst;
    dst = t;
  }
  return src;
}

__device__ __forceinline__ double vel1d(const Args1D& a, int b, int x) {
  if (a.u) return a.u[(size_t)b * a.u_stride + x];
  double u = 0.0;
  for (int m = 0; m < a.sep.nterms; ++m) u += a.sep.a[m] * a.sep.xt[m * a.nx + x];
  return u;
}

// N^ = rfft(-u * irfft(i*kr*ss)) into `nh` (shared), for the stage state `ss` (global, this member's nkr values)
__device__ void calcN1d(const Args1D& a, int b, double2* ss, double2* w0, double2* w1, const double2* tw, double2* nh) {
  const int nx = a.nx, H = nx >> 1;
  for (int k = threadIdx.x; k <= H; k += blockDim.x) {
    double2 s = ss[k];
    if (a.ax.dealias && dealiased_out(a.ax, k, 0, 0)) {  // opt-in dealias!(sol) at the top of calcN
      s = make_double2(0.0, 0.0);
      ss[k] = s;
    }
    const double kx = a.ax.kx[k] * a.inv_n;
    const double2 X = make_double2(-kx * s.y, kx * s.x);  // i*kr*s / nx
    if (k == 0 || k == H) {
      w0[k] = make_double2(X.x, 0.0);                     // c2r keeps only the real part of the DC / Nyquist bins
    } else {
      w0[k] = X;
      w0[nx - k] = make_double2(X.x, -X.y);               // Hermitian extension
    }
  }
  __syncthreads();
  double2* r = fft_smem(w0, w1, nx, +1, tw);
  double2* o = (r == w0) ? w1 : w0;
  for (int x = threadIdx.x; x < nx; x += blockDim.x) o[x] = make_double2(-vel1d(a, b, x) * r[x].x, 0.0);  // TAD.jl:701,750
  __syncthreads();
  double2* f = fft_smem(o, r, nx, -1, tw);
  for (int k = threadIdx.x; k <= H; k += blockDim.x) nh[k] = f[k];
  __syncthreads();
}

__device__ void combine1d(const Args1D& a, int b, const double2* nh, int mode, double la = 0, double lb = 0,
                          int llast = 0) {
  CombineArgs C = a.C;
  C.mode = mode;
  C.lsrk_a = la;
  C.lsrk_b = lb;
  C.lsrk_last = llast;
  for (int k = threadIdx.x; k < a.nkr; k += blockDim.x)
    combine_at(a.P, C, a.ax, (size_t)b * a.nkr + k, (size_t)k, a.ax.kx[k], 0.0, 0.0, nh[k]);
  __syncthreads();  // the next stage reads what other threads of this CTA just wrote
}

__global__ void __launch_bounds__(512) k_step1d(Args1D a) {
  extern __shared__ double2 sm1[];
  const int nx = a.nx;
  double2* w0 = sm1;
  double2* w1 = w0 + nx;
  double2* nh = w1 + nx;            // nkr
  double2* tw = nh + (nx / 2 + 1);  // nx/2
  const int b = blockIdx.x;
  for (int m = threadIdx.x; m < nx / 2; m += blockDim.x) {
    double s, c;
    sincospi(-2.0 * (double)m / (double)nx, &s, &c);
    tw[m] = make_double2(c, s);
  }
  __syncthreads();
  // Shared-memory residency of the whole problem state: a 1-D problem is a few KB, so every load/store of the stage
  // combines becomes a shared-memory access instead of an L2 round trip (the state is written back once, at the end).
  CombinePtrs G = a.P;  // global arrays
  const int nkr = a.nkr;
  if (a.state_in_smem) {
    double2* base = tw + nx / 2;
    double2** cp[5] = {&a.P.s0, &a.P.s1, &a.P.s2, &a.P.acc, &a.P.n1};
    double2* gp[5] = {G.s0, G.s1, G.s2, G.acc, G.n1};
    for (int q = 0; q < 5; ++q) {
      if (!gp[q]) continue;
      for (int k = threadIdx.x; k < nkr; k += blockDim.x) base[k] = gp[q][(size_t)b * nkr + k];
      *cp[q] = base - (size_t)b * nkr;  // combine_at indexes with b*nkr + k
      base += nkr;
    }
    double* rb = reinterpret_cast<double*>(base);
    const double** rp[6] = {&a.P.E, &a.P.E2, &a.P.zeta, &a.P.alpha, &a.P.beta, &a.P.gamma};
    const double* rg[6] = {G.E, G.E2, G.zeta, G.alpha, G.beta, G.gamma};
    for (int q = 0; q < 6; ++q) {
      if (!rg[q]) continue;
      for (int k = threadIdx.x; k < nkr; k += blockDim.x) rb[k] = rg[q][k];
      *rp[q] = rb;
      rb += nkr + (nkr & 1);
    }
    __syncthreads();
  }
  double2* s0 = a.P.s0 + (size_t)b * a.nkr;
  double2* s1 = a.P.s1 ? a.P.s1 + (size_t)b * a.nkr : nullptr;
  double2* s2 = a.P.s2 ? a.P.s2 + (size_t)b * a.nkr : nullptr;
  const double LA[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                        -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
  const double LB[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                        1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                        2277821191437.0 / 14882151754819.0};
  for (int it = 0; it < a.nsteps; ++it) {
    switch (a.base) {
      case PTF_STEPPER_RK4:
        calcN1d(a, b, s0, w0, w1, tw, nh); combine1d(a, b, nh, CM_RK4_S1);
        calcN1d(a, b, s1, w0, w1, tw, nh); combine1d(a, b, nh, CM_RK4_S2);
        calcN1d(a, b, s1, w0, w1, tw, nh); combine1d(a, b, nh, CM_RK4_S3);
        calcN1d(a, b, s1, w0, w1, tw, nh); combine1d(a, b, nh, CM_RK4_S4);
        break;
      case PTF_STEPPER_ETDRK4:
        calcN1d(a, b, s0, w0, w1, tw, nh); combine1d(a, b, nh, CM_ETD_S1);
        calcN1d(a, b, s1, w0, w1, tw, nh); combine1d(a, b, nh, CM_ETD_S2);
        calcN1d(a, b, s2, w0, w1, tw, nh); combine1d(a, b, nh, CM_ETD_S3);
        calcN1d(a, b, s2, w0, w1, tw, nh); combine1d(a, b, nh, CM_ETD_S4);
        break;
      case PTF_STEPPER_FORWARD_EULER:
        calcN1d(a, b, s0, w0, w1, tw, nh); combine1d(a, b, nh, CM_EULER);
        break;
      case PTF_STEPPER_LSRK54:
        for (int i = 0; i < 5; ++i) {
          calcN1d(a, b, s0, w0, w1, tw, nh);
          combine1d(a, b, nh, CM_LSRK, LA[i], LB[i], i == 4);
        }
        break;
      case PTF_STEPPER_AB3:
        calcN1d(a, b, s0, w0, w1, tw, nh);
        combine1d(a, b, nh, (a.step0 + it < 3) ? CM_AB3_EULER : CM_AB3);
        break;
    }
  }
  if (a.state_in_smem) {  // write the persistent state back (stage scratch s1/s2 need not survive the launch)
    double2* sp[3] = {a.P.s0, a.P.acc, a.P.n1};
    double2* gp[3] = {G.s0, G.acc, G.n1};
    for (int q = 0; q < 3; ++q) {
      if (!gp[q]) continue;
      for (int k = threadIdx.x; k < nkr; k += blockDim.x) gp[q][(size_t)b * nkr + k] = sp[q][(size_t)b * nkr + k];
    }
  }
}

// set_c! (mode 0: sol = rfft(c)) and updatevars! (mode 1: c = irfft(sol)) with the same shared-memory transform
__global__ void __launch_bounds__(512) k_transform1d(double2* sol, double* c, int nx, int mode) {
  extern __shared__ double2 sm1[];
  double2* w0 = sm1;
  double2* w1 = w0 + nx;
  double2* tw = w1 + nx;
  const int b = blockIdx.x, H = nx / 2, nkr = H + 1;
  for (int m = threadIdx.x; m < H; m += blockDim.x) {
    double s, cc;
    sincospi(-2.0 * (double)m / (double)nx, &s, &cc);
    tw[m] = make_double2(cc, s);
  }
  if (mode == 0) {
    for (int x = threadIdx.x; x < nx; x += blockDim.x) w0[x] = make_double2(c[(size_t)b * nx + x], 0.0);
    __syncthreads();
    double2* f = fft_smem(w0, w1, nx, -1, tw);
    for (int k = threadIdx.x; k < nkr; k += blockDim.x) sol[(size_t)b * nkr + k] = f[k];
  } else {
    const double inv = 1.0 / (double)nx;
    for (int k = threadIdx.x; k < nkr; k += blockDim.x) {
      double2 s = sol[(size_t)b * nkr + k];
      s.x *= inv;
      s.y *= inv;
      if (k == 0 || k == H) {
        w0[k] = make_double2(s.x, 0.0);
      } else {
        w0[k] = s;
        w0[nx - k] = make_double2(s.x, -s.y);
      }
    }
    __syncthreads();
    double2* r = fft_smem(w0, w1, nx, +1, tw);
    for (int x = threadIdx.x; x < nx; x += blockDim.x) c[(size_t)b * nx + x] = r[x].x;
  }
}

bool pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

class Fused1DEngine final : public Engine {
 public:
  explicit Fused1DEngine(Context& c) : ctx(c), g(c.g) {
    nx = (int)g.nx;
    nkr = (int)g.nkr;
    nb = (int)g.B;
    nspec = (size_t)nkr * nb;
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
    if (base == PTF_STEPPER_ETDRK4)
      for (auto* b : {&cE, &cE2, &cZ, &cA, &cB, &cG}) b->alloc(nkr, &dev_bytes);
    cbuf.alloc((size_t)nx * nb, &dev_bytes);
    vs.init(&g, ctx.stream, ctx.d.flow_kind, &dev_bytes);
    nthreads = nx / 2 < 64 ? 64 : (nx / 2 > 512 ? 512 : nx / 2);
    smem_step = (size_t)(2 * nx + nkr + nx / 2) * sizeof(double2);
    {
      size_t st = (size_t)5 * nkr * sizeof(double2) + (size_t)6 * (nkr + 1) * sizeof(double);
      state_in_smem = (smem_step + st <= 200 * 1024) ? 1 : 0;
      if (state_in_smem) smem_step += st;
    }
    smem_tr = (size_t)(2 * nx + nx / 2) * sizeof(double2);
    PTF_CUDA(cudaFuncSetAttribute(k_step1d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_step));
    PTF_CUDA(cudaFuncSetAttribute(k_transform1d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tr));
    on_dt_changed();
  }

  const char* name() const override { return "fused"; }
  int id() const override { return PTF_ENGINE_FUSED; }
  cudaStream_t stream() const override { return ctx.stream; }

  void set_velocity(int comp, const double* host, int64_t count) override { vs.set_array(comp, host, count); }
  void set_velocity_separable(int comp, int nterms, const double* xt, const double* yt, const double* zt,
                              const double* coeff0) override {
    vs.set_separable(comp, nterms, xt, yt, zt, coeff0);
  }
  void set_velocity_coeffs(int comp, int nterms, const double* a) override { vs.set_coeffs(comp, nterms, a); }
  void set_layered_shift(const double*) override { throw Error(PTF_EINVAL, "layered velocities are 2-D"); }

  void set_c(const double* c_host, bool replicate) override {
    if (replicate && nb > 1) {
      for (int b = 0; b < nb; ++b)
        PTF_CUDA(cudaMemcpyAsync(cbuf.p + (size_t)b * nx, c_host, nx * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    } else {
      PTF_CUDA(cudaMemcpyAsync(cbuf.p, c_host, cbuf.bytes(), cudaMemcpyHostToDevice, ctx.stream));
    }
    k_transform1d<<<nb, nthreads, smem_tr, ctx.stream>>>(s0.p, cbuf.p, nx, 0);
    ++own_launches;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  void get_c(double* c_host) override {
    k_transform1d<<<nb, nthreads, smem_tr, ctx.stream>>>(s0.p, cbuf.p, nx, 1);
    ++own_launches;
    PTF_CUDA(cudaMemcpyAsync(c_host, cbuf.p, cbuf.bytes(), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  void set_sol(const double* s_host) override {
    PTF_CUDA(cudaMemcpyAsync(s0.p, s_host, s0.bytes(), cudaMemcpyHostToDevice, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  void get_sol(double* s_host) override {
    PTF_CUDA(cudaMemcpyAsync(s_host, s0.p, s0.bytes(), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void on_dt_changed() override {
    if (ctx.st.base == PTF_STEPPER_ETDRK4) {
      k_etd_coeffs<<<4, 256, 0, ctx.stream>>>(cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p, ctx.ax, nkr, 1, 1, ctx.dt, 0);
      ++own_launches;
      PTF_CUDA(cudaGetLastError());
    }
  }

  void launch(int64_t step_index, int nsteps) {
    Args1D a;
    a.P = CombinePtrs{s0.p, s1.p, s2.p, acc.p, n1.p, cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p};
    a.C = CombineArgs{0, ctx.st.filtered ? 1 : 0, ctx.dt, 0.0, 0.0, 0};
    a.ax = ctx.ax;
    a.u = (vs.va.kind == PTF_FLOW_SEPARABLE) ? nullptr : vs.va.arr[0];
    a.sep = vs.va.sep[0];
    a.u_stride = vs.va.member_stride;
    if (vs.va.kind != PTF_FLOW_SEPARABLE && !a.u)
      throw Error(PTF_EINVAL, "velocity field has not been set (ptf_set_velocity / callback)");
    a.base = ctx.st.base;
    a.nx = nx;
    a.nkr = nkr;
    a.step0 = step_index;
    a.nsteps = nsteps;
    a.inv_n = 1.0 / (double)nx;
    a.state_in_smem = state_in_smem;
    k_step1d<<<nb, nthreads, smem_step, ctx.stream>>>(a);
    PTF_CUDA(cudaGetLastError());
    ++own_launches;
  }

  void step_once(int64_t step_index) override { launch(step_index, 1); }
  bool step_many(int64_t first_step, int64_t n) override {
    while (n > 0) {  // the whole stepforward!(prob, n) call in one launch (chunks only guard the int range)
      int chunk = n > (1 << 20) ? (1 << 20) : (int)n;
      launch(first_step, chunk);
      first_step += chunk;
      n -= chunk;
    }
    return true;
  }

  void diag(double* mean_c, double* var_c, double* max_abs_sol) override {  // 1-D state is a few KB: host side
    std::vector<double2> h(nspec);
    get_sol(reinterpret_cast<double*>(h.data()));
    double sumsq = 0, mx = 0;
    for (size_t i = 0; i < nspec; ++i) {
      size_t k = i % nkr;
      double a2 = h[i].x * h[i].x + h[i].y * h[i].y;
      sumsq += ((k == 0 || k == (size_t)nx / 2) ? 1.0 : 2.0) * a2;
      mx = a2 > mx ? a2 : mx;
    }
    double N = (double)nx, m = h[0].x / N;
    if (mean_c) *mean_c = m;
    if (var_c) *var_c = sumsq / (N * N) / nb - m * m;
    if (max_abs_sol) *max_abs_sol = std::sqrt(mx);
  }

  float time_kernel(const char*, int) override {
    throw Error(PTF_EUNSUPPORTED, "1-D engine: a step is a single kernel; use ptf_step_timed");
  }

 private:
  Context& ctx;
  Geometry& g;
  int nx, nkr, nb, nthreads, state_in_smem = 0;
  size_t nspec, smem_step, smem_tr;
  DevBuf<double2> s0, s1, s2, acc, n1;
  DevBuf<double> cE, cE2, cZ, cA, cB, cG, cbuf;
  VelocityStore vs;
};

}  // namespace

bool fused1d_engine_supports(const Context& ctx, std::string* why) {
  auto no = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  if (ctx.g.ndim != 1) return no("not a 1-D problem");
  if (!pow2(ctx.g.nx) || ctx.g.nx < 16 || ctx.g.nx > 2048) return no("nx must be a power of two in [16, 2048]");
  if (ctx.g.B > 65535) return no("batch too large");
  return true;
}

std::unique_ptr<Engine> make_fused1d_engine(Context& ctx) { return std::unique_ptr<Engine>(new Fused1DEngine(ctx)); }

}  // namespace ptf
