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

namespace ptf {

namespace {

constexpr int MAX_TERMS = 8;

struct SepFlow {      // u_comp(x,y,z,t) = sum_m a[m] * X[m][i] * Y[m][j] * Z[m][k]
  int nterms = 0;
  const double* xt = nullptr;  // [nterms][nx]
  const double* yt = nullptr;  // [nterms][ny]
  const double* zt = nullptr;  // [nterms][nz]
  const double* a = nullptr;   // [nterms] device-resident coefficients a_m(t_n): refreshed per step without
                               // touching the captured graph
};

struct VelArgs {
  int kind = 0;                                    // PTF_FLOW_*
  const double* arr[3] = {nullptr, nullptr, nullptr};  // array form (STEADY / CALLBACK / LAYERED)
  int64_t member_stride = 0;                       // 0 when one field is shared by all members
  const double* ushift = nullptr;                  // LAYERED: U(y, layer) added to u, may be null
  SepFlow sep[3];
};

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

__device__ __forceinline__ double sep_eval(const SepFlow& f, int64_t i, int64_t j, int64_t k, int64_t nx, int64_t ny,
                                           int64_t nz, int nd) {
  double u = 0.0;
  for (int m = 0; m < f.nterms; ++m) {
    double t = f.a[m] * f.xt[m * nx + i];
    if (nd >= 2) t *= f.yt[m * ny + j];
    if (nd >= 3) t *= f.zt[m * nz + k];
    u += t;
  }
  return u;
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
struct CombinePtrs {
  const double2* Nh;  // transformed nonlinear term of this stage
  double2* s0;        // sol
  double2* s1;        // stage state / sol_1
  double2* s2;        // ETDRK4 sol_2
  double2* acc;       // RK4 accumulator / ETD N2+N3 / LSRK S2 / AB3 RHS_{-1}
  double2* n1;        // ETD N1 / AB3 RHS_{-2}
  const double *E, *E2, *zeta, *alpha, *beta, *gamma;  // ETDRK4 coefficient arrays [nz][ny][nkr]
};

__global__ void __launch_bounds__(256) k_combine(CombinePtrs P, CombineArgs A, AxisTables ax, SpecShape sh) {
  int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  int64_t nrows = sh.ny * sh.nz * sh.B;
  if (row >= nrows) return;
  int64_t iy = row % sh.ny;
  int64_t iz = (row / sh.ny) % sh.nz;
  int64_t crow = (iz * sh.ny + iy) * sh.nkr;  // row base in the (batch-shared) coefficient arrays
  double ky = ax.ky[iy], kz = ax.kz[iz];
  int64_t base = row * sh.nkr;
  const double dt = A.dt;
  for (int64_t ix = threadIdx.x; ix < sh.nkr; ix += blockDim.x) {
    int64_t i = base + ix;
    double kx = ax.kx[ix];
    double2 Nh = P.Nh[i];
    double f = 1.0;
    switch (A.mode) {
      case CM_RK4_S1: {
        double L = lin_op(ax, kx, ky, kz);
        double2 s0 = P.s0[i];
        double2 k = cadd(Nh, cmul_r(s0, L));
        P.acc[i] = cdiv_r(k, 6.0);
        P.s1[i] = cadd(s0, cmul_r(k, dt / 2));
      } break;
      case CM_RK4_S2:
      case CM_RK4_S3: {
        double L = lin_op(ax, kx, ky, kz);
        double2 ss = P.s1[i];
        double2 k = cadd(Nh, cmul_r(ss, L));
        P.acc[i] = cadd(P.acc[i], cdiv_r(k, 3.0));
        double h = (A.mode == CM_RK4_S2) ? dt / 2 : dt;
        P.s1[i] = cadd(P.s0[i], cmul_r(k, h));
      } break;
      case CM_RK4_S4: {
        double L = lin_op(ax, kx, ky, kz);
        double2 ss = P.s1[i];
        double2 k = cadd(Nh, cmul_r(ss, L));
        double2 sum = cadd(P.acc[i], cdiv_r(k, 6.0));
        double2 r = cadd(P.s0[i], cmul_r(sum, dt));
        if (A.filtered) f = filter_val(ax, kx, ky, kz);
        P.s0[i] = cmul_r(r, f);
      } break;
      case CM_ETD_S1: {
        double2 s0 = P.s0[i];
        P.n1[i] = Nh;
        P.s1[i] = cadd(cmul_r(s0, P.E2[crow + ix]), cmul_r(Nh, P.zeta[crow + ix]));
      } break;
      case CM_ETD_S2: {
        P.acc[i] = Nh;
        P.s2[i] = cadd(cmul_r(P.s0[i], P.E2[crow + ix]), cmul_r(Nh, P.zeta[crow + ix]));
      } break;
      case CM_ETD_S3: {
        P.acc[i] = cadd(P.acc[i], Nh);
        double2 n1 = P.n1[i];
        double2 t = make_double2(2 * Nh.x - n1.x, 2 * Nh.y - n1.y);
        P.s2[i] = cadd(cmul_r(P.s1[i], P.E2[crow + ix]), cmul_r(t, P.zeta[crow + ix]));
      } break;
      case CM_ETD_S4: {
        double2 r = cmul_r(P.s0[i], P.E[crow + ix]);
        r = cadd(r, cmul_r(P.n1[i], P.alpha[crow + ix]));
        r = cadd(r, cmul_r(P.acc[i], 2 * P.beta[crow + ix]));
        r = cadd(r, cmul_r(Nh, P.gamma[crow + ix]));
        if (A.filtered) f = filter_val(ax, kx, ky, kz);
        P.s0[i] = cmul_r(r, f);
      } break;
      case CM_EULER: {
        double L = lin_op(ax, kx, ky, kz);
        double2 s0 = P.s0[i];
        double2 k = cadd(Nh, cmul_r(s0, L));
        double2 r = cadd(s0, cmul_r(k, dt));
        if (A.filtered) f = filter_val(ax, kx, ky, kz);
        P.s0[i] = cmul_r(r, f);
      } break;
      case CM_LSRK: {
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
        P.s0[i] = r;
      } break;
      case CM_AB3_EULER:
      case CM_AB3: {
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
        P.s0[i] = cmul_r(r, f);
        P.n1[i] = k1;
        P.acc[i] = k;
      } break;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// ETDRK4 coefficients (FF getetdcoeffs / getexpLs): 32-point contour mean around dt*L, on device.
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

__global__ void __launch_bounds__(256) k_etd_coeffs(double* E, double* E2, double* zeta, double* alpha, double* beta,
                                                    double* gamma, AxisTables ax, SpecShape sh, double dt) {
  int64_t n = sh.nkr * sh.ny * sh.nz;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t ix = i % sh.nkr;
    int64_t iy = (i / sh.nkr) % sh.ny;
    int64_t iz = i / (sh.nkr * sh.ny);
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
      // zeta: (exp(z/2)-1)/z
      sz += cx_div(make_double2(ez2.x - 1.0, ez2.y), z).x;
      // alpha: (-4 - z + exp(z)(4 - 3z + z^2))/z^3
      double2 t = cx_mul(ez, make_double2(4.0 - 3.0 * z.x + z2.x, -3.0 * z.y + z2.y));
      sa += cx_div(make_double2(-4.0 - z.x + t.x, -z.y + t.y), z3).x;
      // beta: (2 + z + exp(z)(-2 + z))/z^3
      t = cx_mul(ez, make_double2(-2.0 + z.x, z.y));
      sb += cx_div(make_double2(2.0 + z.x + t.x, z.y + t.y), z3).x;
      // gamma: (-4 - 3z - z^2 + exp(z)(4 - z))/z^3
      t = cx_mul(ez, make_double2(4.0 - z.x, -z.y));
      sg += cx_div(make_double2(-4.0 - 3.0 * z.x - z2.x + t.x, -3.0 * z.y - z2.y + t.y), z3).x;
    }
    E[i] = exp(Ldt);
    E2[i] = exp(Ldt / 2);
    zeta[i] = dt * (sz / ncirc);
    alpha[i] = dt * (sa / ncirc);
    beta[i] = dt * (sb / ncirc);
    gamma[i] = dt * (sg / ncirc);
  }
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
    va = VelArgs{};
    va.kind = ctx.d.flow_kind;
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
  void ensure_vel(int comp, int64_t count) {
    if (vel[comp].n != (size_t)count) {
      vel[comp].alloc(count, &dev_bytes);
      drop_graphs();
    }
    va.arr[comp] = vel[comp].p;
  }

  void set_velocity(int comp, const double* host, int64_t count) override {
    PTF_REQUIRE(comp >= 0 && comp < nd, "velocity component out of range");
    PTF_REQUIRE(count == g.npts() || count == g.npts() * g.B, "velocity count must be npts or npts*nbatch");
    PTF_REQUIRE(comp == 0 || vel[0].n == 0 || vel[0].n == (size_t)count,
                "all velocity components must have the same extent");
    ensure_vel(comp, count);
    int64_t ms = (count == g.npts() && g.B > 1) ? 0 : g.npts();
    if (va.member_stride != ms) drop_graphs();
    va.member_stride = ms;
    PTF_CUDA(cudaMemcpyAsync(vel[comp].p, host, count * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void set_velocity_separable(int comp, int nterms, const double* xt, const double* yt, const double* zt,
                              const double* coeff0) override {
    PTF_REQUIRE(comp >= 0 && comp < nd, "velocity component out of range");
    PTF_REQUIRE(nterms >= 0 && nterms <= MAX_TERMS, "separable flow supports at most 8 terms per component");
    SepFlow& f = va.sep[comp];
    f.nterms = nterms;
    auto up = [&](DevBuf<double>& b, const double* h, int64_t n) -> const double* {
      if (!h || nterms == 0) return nullptr;
      b.alloc((size_t)nterms * n, &dev_bytes);
      PTF_CUDA(cudaMemcpy(b.p, h, (size_t)nterms * n * sizeof(double), cudaMemcpyHostToDevice));
      return b.p;
    };
    f.xt = up(sepx[comp], xt, g.nx);
    PTF_REQUIRE(nterms == 0 || f.xt, "separable flow needs an x table");
    f.yt = nd >= 2 ? up(sepy[comp], yt, g.ny) : nullptr;
    f.zt = nd >= 3 ? up(sepz[comp], zt, g.nz) : nullptr;
    PTF_REQUIRE(nterms == 0 || nd < 2 || f.yt, "separable flow needs a y table");
    PTF_REQUIRE(nterms == 0 || nd < 3 || f.zt, "separable flow needs a z table");
    if (!sepa.p) {
      sepa.alloc(3 * MAX_TERMS, &dev_bytes);
      PTF_CUDA(cudaMemset(sepa.p, 0, sepa.bytes()));
    }
    f.a = sepa.p + comp * MAX_TERMS;
    double a0[MAX_TERMS];
    for (int m = 0; m < MAX_TERMS; ++m) a0[m] = (m < nterms) ? (coeff0 ? coeff0[m] : 1.0) : 0.0;
    PTF_CUDA(cudaMemcpy(sepa.p + comp * MAX_TERMS, a0, sizeof(a0), cudaMemcpyHostToDevice));
    drop_graphs();
  }

  void set_velocity_coeffs(int comp, int nterms, const double* a) override {
    SepFlow& f = va.sep[comp];
    PTF_REQUIRE(nterms == f.nterms && f.a, "coefficient count does not match the separable flow");
    // pageable-source async copy: staged before the call returns, ordered on the step stream
    PTF_CUDA(cudaMemcpyAsync(sepa.p + comp * MAX_TERMS, a, nterms * sizeof(double), cudaMemcpyHostToDevice,
                             ctx.stream));
  }

  void set_layered_shift(const double* U) override {
    if (!U) {
      if (va.ushift) drop_graphs();
      va.ushift = nullptr;
      return;
    }
    if (ushift.n != (size_t)(g.B * g.ny)) {
      ushift.alloc(g.B * g.ny, &dev_bytes);
      drop_graphs();
    }
    PTF_CUDA(cudaMemcpyAsync(ushift.p, U, ushift.bytes(), cudaMemcpyHostToDevice, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    if (va.ushift != ushift.p) drop_graphs();
    va.ushift = ushift.p;
  }

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
      k_etd_coeffs<<<flat_blocks(g.nspec()), 256, 0, ctx.stream>>>(cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p, ctx.ax, shape(),
                                                                    ctx.dt);
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
      k_product<1><<<pg, 256, 0, ctx.stream>>>(gr[0].p, g1, g2, va, g.nx, g.ny, g.nz);
    else if (nd == 2)
      k_product<2><<<pg, 256, 0, ctx.stream>>>(gr[0].p, g1, g2, va, g.nx, g.ny, g.nz);
    else
      k_product<3><<<pg, 256, 0, ctx.stream>>>(gr[0].p, g1, g2, va, g.nx, g.ny, g.nz);
    ++own_launches;
    PTF_CUFFT(cufftExecD2Z(plan_fwd, gr[0].p, reinterpret_cast<cufftDoubleComplex*>(dh[0].p)));
    ++lib_calls;
  }

  void combine(int mode, double lsrk_a = 0, double lsrk_b = 0, int lsrk_last = 0) {
    dim3 grid, block;
    spec_launch_dims(grid, block);
    CombinePtrs P{dh[0].p, sol.p, s1.p, s2.p, acc.p, n1.p, cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p};
    CombineArgs A{mode, ctx.st.filtered ? 1 : 0, ctx.dt, lsrk_a, lsrk_b, lsrk_last};
    k_combine<<<grid, block, 0, ctx.stream>>>(P, A, ctx.ax, shape());
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
  DevBuf<double> gr[3], vel[3], sepx[3], sepy[3], sepz[3], sepa, ushift;
  DevBuf<double> cE, cE2, cZ, cA, cB, cG;
  DevBuf<char> work;
  VelArgs va;
  cufftHandle plan_fwd = 0, plan_inv = 0;
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int64_t per_step_own = 0, per_step_lib = 0;
};

}  // namespace

std::unique_ptr<Engine> make_cufft_engine(Context& ctx) { return std::unique_ptr<Engine>(new CufftEngine(ctx)); }

}  // namespace ptf
