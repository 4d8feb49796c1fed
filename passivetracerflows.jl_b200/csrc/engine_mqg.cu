// MultiLayerQG flow solver on the device — the flow that drives the MQG-coupled tracer.
//
// Reference call sites (the solver itself lives in the un-vendored dependency GeophysicalFlows 0.16, Project.toml:25):
//   MultiLayerQG.Problem(nlayers, dev; nx, Lx, f₀, H, b, U, μ, β, dt, stepper, aliased_fraction)   examples/turbulent_advection-diffusion.jl:56-58
//   MultiLayerQG.set_q!                                                                            examples/…:69
//   step_until!(MQGprob, tracer_release_time)                                                      TAD.jl:238
//   MultiLayerQG.updatevars!(MQGprob)                                                              TAD.jl:488, examples/…:151
//   stepforward!(params.MQGprob)                                                                   examples/…:150
//   MQGprob.vars.u .+ MQGprob.params.U, MQGprob.vars.v  read by the tracer's calcN!                TAD.jl:795-796
//
// One stage of the flow's calcN! is 4 launches + 2 batched cuFFT calls, everything L2-resident at the example's size:
//   k_mqg_spec<COMBINE,FRONT>  per wavenumber, all layers in registers: [stage combine of the PREVIOUS stage: N̂ assembled
//                              from the three transformed products, bottom drag, FourierFlows stepper update] then
//                              [dealias!(state) in place, ψ̂ = S⁻¹ q̂, û = −i l ψ̂, v̂ = i kr ψ̂, q̂ copy, all pre-scaled by 1/N]
//   cuFFT Z2D, batch 3·nlayers  (u, v, q)
//   k_mqg_products             u += U(y, layer);  P0 = u·Qx + v·Qy,  P1 = u·q,  P2 = v·q   (in place)
//   cuFFT D2Z, batch 3·nlayers
// N̂ = −P̂0 − i kr P̂1 − i l P̂2 (+ μ|k|²ψ̂ in the bottom layer).  GeophysicalFlows transforms u·Qx and v·Qy separately;
// the forward transform is linear, so transforming their sum differs only in rounding (≈1e-16) and saves one transform.
#include <cmath>
#include <cstring>
#include <vector>

#include "ptf_internal.h"
#include "ptf_pointwise.cuh"

namespace ptf {
namespace {

struct MqgSpecArgs {
  double2* state;      // FRONT: the state the next calcN is evaluated at (dealiased in place)   [NL][ny][nkr]
  double2* spec3;      // COMBINE in: P̂0,P̂1,P̂2 ; FRONT out: û, v̂, q̂ (× 1/N)                    [3][NL][ny][nkr]
  double2* psih;       // ψ̂ of the last FRONT (unnormalised)                                    [NL][ny][nkr]
  const double* sinv;  // S⁻¹                                                                   [NL*NL][ny][nkr]
  AxisTables ax;       // kx, ky; L = −ν|k|^{2nν} through (kappa_h, n_kappa_h); filter; dealias ranges
  int64_t nkr, ny;
  double mu, scale;
  int uniform_bg;                    // Qx == 0 and Qy uniform per layer: the background term is applied spectrally
  double qy[PTF_MQG_MAX_LAYERS];     // that uniform Qy_j
  CombinePtrs P;
  CombineArgs A;
};

template <int NL, bool COMBINE, bool FRONT>
__global__ void __launch_bounds__(128) k_mqg_spec(MqgSpecArgs a) {
  const int64_t ix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t iy = blockIdx.y;
  if (ix >= a.nkr) return;
  const int64_t plane = a.nkr * a.ny;
  const int64_t i = iy * a.nkr + ix;
  const double kx = a.ax.kx[ix], ky = a.ax.ky[iy];
  double2 q[NL];
  if (COMBINE) {
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      double2 f0;
      if (a.uniform_bg) {   // rfft(v*Qy_j) = Qy_j * i kr psi-hat_j (v = irfft(i kr psi-hat), Qy_j a constant); u*Qx = 0
        const double2 pj = a.psih[(int64_t)j * plane + i];
        f0 = make_double2(-a.qy[j] * kx * pj.y, a.qy[j] * kx * pj.x);
        if (ix == a.nkr - 1) {
          // x-Nyquist column: the c2r that produced v dropped Im of this column after the y transform, i.e. v carries
          // the y-Hermitian part (W(l) + conj W(-l)) / 2 of W = i kr psi-hat only.  Invisible while dealias! zeroes the
          // column; with aliased_fraction = 0 (FourierFlows: no-op mask, the example's setting) it is 6.5e-8 of the state.
          const double2 pm = a.psih[(int64_t)j * plane + ((a.ny - iy) % a.ny) * a.nkr + ix];
          const double wx = -a.qy[j] * kx * pm.y, wy = a.qy[j] * kx * pm.x;
          f0 = make_double2(0.5 * (f0.x + wx), 0.5 * (f0.y - wy));
        }
      } else {
        f0 = a.spec3[(0 * NL + j) * plane + i];
      }
      const double2 f1 = a.spec3[(1 * NL + j) * plane + i];
      const double2 f2 = a.spec3[(2 * NL + j) * plane + i];
      // N̂ = −P̂0 − i kr P̂1 − i l P̂2
      double2 Nh = make_double2(-f0.x + (kx * f1.y + ky * f2.y), -f0.y - (kx * f1.x + ky * f2.x));
      if (j == NL - 1 && a.mu != 0.0) {   // bottom linear drag: + μ |k|² ψ̂_n
        const double2 pb = a.psih[(int64_t)j * plane + i];
        const double m = a.mu * (kx * kx + ky * ky);
        Nh.x += m * pb.x;
        Nh.y += m * pb.y;
      }
      q[j] = combine_at<CMASK_ALL>(a.P, a.A, a.ax, (size_t)(j * plane + i), (size_t)i, kx, ky, 0.0, Nh);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NL; ++j) q[j] = a.state[(int64_t)j * plane + i];
  }
  if (FRONT) {
    if (dealiased_out(a.ax, ix, iy, 0)) {   // dealias!(sol, grid): in place on the stage state
#pragma unroll
      for (int j = 0; j < NL; ++j) {
        q[j] = make_double2(0.0, 0.0);
        a.state[(int64_t)j * plane + i] = q[j];
      }
    }
    double2 psi[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      double2 s = make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < NL; ++m) {
        const double w = a.sinv[(int64_t)(j * NL + m) * plane + i];
        s.x += w * q[m].x;
        s.y += w * q[m].y;
      }
      psi[j] = s;
    }
    const double sc = a.scale;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      a.psih[(int64_t)j * plane + i] = psi[j];
      a.spec3[(0 * NL + j) * plane + i] = make_double2(ky * psi[j].y * sc, -ky * psi[j].x * sc);   // −i l ψ̂
      a.spec3[(1 * NL + j) * plane + i] = make_double2(-kx * psi[j].y * sc, kx * psi[j].x * sc);   //  i kr ψ̂
      a.spec3[(2 * NL + j) * plane + i] = make_double2(q[j].x * sc, q[j].y * sc);
    }
  }
}

// phys3 = [u, v, q][NL][ny][nx] -> [u·Qx + v·Qy, u·q, v·q] with u += U(y, layer); two points per thread
__global__ void __launch_bounds__(256) k_mqg_products(double* phys3, const double* Qx, const double* Qy, const double* U,
                                                      int64_t nx, int64_t ny, int64_t NL, int uniform_bg) {
  const int64_t T = nx * ny * NL, half = T / 2, hx = nx / 2;
  double2* u2 = reinterpret_cast<double2*>(phys3);
  double2* v2 = reinterpret_cast<double2*>(phys3 + T);
  double2* q2 = reinterpret_cast<double2*>(phys3 + 2 * T);
  const double2* qx2 = reinterpret_cast<const double2*>(Qx);
  const double2* qy2 = reinterpret_cast<const double2*>(Qy);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < half; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / hx;   // = layer*ny + y
    const double Us = U[row];
    double2 u = u2[e], v = v2[e], q = q2[e];
    u.x += Us;
    u.y += Us;
    if (!uniform_bg) {
      const double2 qx = qx2[e], qy = qy2[e];
      u2[e] = make_double2(u.x * qx.x + v.x * qy.x, u.y * qx.y + v.y * qy.y);
    }
    v2[e] = make_double2(u.x * q.x, u.y * q.y);
    q2[e] = make_double2(v.x * q.x, v.y * q.y);
  }
}

__global__ void __launch_bounds__(256) k_mqg_scale_copy(const double2* in, double2* out, int64_t n, double s) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = make_double2(in[i].x * s, in[i].y * s);
}

// set_q!: zero the (0,0) mode of every layer
__global__ void k_mqg_zero_mean(double2* sol, int64_t plane, int NL) {
  int j = threadIdx.x;
  if (j < NL) sol[(int64_t)j * plane] = make_double2(0.0, 0.0);
}

// q̂ = S ψ̂ (pvfromstreamfunction!), S = −|k|² I + F
template <int NL>
__global__ void __launch_bounds__(128) k_mqg_pv_from_psi(const double2* psih, double2* qh, const double* F, AxisTables ax,
                                                         int64_t nkr, int64_t ny) {
  const int64_t ix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t iy = blockIdx.y;
  if (ix >= nkr) return;
  const int64_t plane = nkr * ny, i = iy * nkr + ix;
  const double kx = ax.kx[ix], ky = ax.ky[iy];
  const double k2 = kx * kx + ky * ky;
  double2 p[NL];
#pragma unroll
  for (int j = 0; j < NL; ++j) p[j] = psih[(int64_t)j * plane + i];
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    double2 s = make_double2(-k2 * p[j].x, -k2 * p[j].y);
#pragma unroll
    for (int m = 0; m < NL; ++m) {
      s.x += F[j * NL + m] * p[m].x;
      s.y += F[j * NL + m] * p[m].y;
    }
    qh[(int64_t)j * plane + i] = s;
  }
}

// Gauss-Jordan inverse of a small dense matrix (row-major, n <= MQG_MAX_LAYERS) with partial pivoting
bool invert_small(const double* A, double* inv, int n) {
  double a[PTF_MQG_MAX_LAYERS][2 * PTF_MQG_MAX_LAYERS];
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      a[r][c] = A[r * n + c];
      a[r][n + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    if (a[piv][c] == 0.0) return false;
    if (piv != c)
      for (int k = 0; k < 2 * n; ++k) std::swap(a[piv][k], a[c][k]);
    const double d = 1.0 / a[c][c];
    for (int k = 0; k < 2 * n; ++k) a[c][k] *= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = a[r][c];
      if (f == 0.0) continue;
      for (int k = 0; k < 2 * n; ++k) a[r][k] -= f * a[c][k];
    }
  }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) inv[r * n + c] = a[r][n + c];
  return true;
}

std::vector<double> wavenumbers(int64_t n, double L, bool half) {
  std::vector<double> k(half ? n / 2 + 1 : n);
  const double k0 = 2.0 * M_PI / L;
  for (int64_t i = 0; i < (int64_t)k.size(); ++i) {
    int64_t j = (half || i < n / 2) ? i : i - n;   // rfftfreq / fftfreq (Nyquist negative on the full axis)
    k[i] = (double)j * k0;
  }
  return k;
}

}  // namespace

// -------------------------------------------------------------------------------------------------------------------
class MqgSolver {
 public:
  explicit MqgSolver(const ptf_mqg_desc& desc) : d(desc) {
    NL = d.nlayers;
    nx = d.nx;
    ny = d.ny;
    nkr = nx / 2 + 1;
    plane = nkr * ny;
    pts = nx * ny;
    PTF_REQUIRE(NL >= 1 && NL <= PTF_MQG_MAX_LAYERS, "nlayers must be between 1 and PTF_MQG_MAX_LAYERS (4)");
    PTF_REQUIRE(nx >= 4 && ny >= 4 && nx % 2 == 0 && ny % 2 == 0, "MultiLayerQG grids need even nx, ny >= 4");
    PTF_REQUIRE(d.Lx > 0 && d.Ly > 0, "domain lengths must be positive");
    PTF_REQUIRE(d.dt > 0, "dt must be positive");
    PTF_REQUIRE(d.n_nu >= 1, "hyperviscosity order n_nu must be >= 1");
    PTF_REQUIRE(d.aliased_fraction >= 0 && d.aliased_fraction < 1, "`aliased_fraction` must be in [0, 1)");
    PTF_REQUIRE(d.H != nullptr, "layer depths H are required");
    PTF_REQUIRE(NL == 1 || d.b != nullptr, "layer buoyancies b are required for nlayers >= 2");
    st.base = d.stepper & 15;
    st.filtered = (d.stepper & PTF_STEPPER_FILTERED) != 0;
    PTF_REQUIRE(st.nstages() > 0, "unknown stepper");
    dt = d.dt;
    PTF_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    stream = own_stream;
    setup_tables();
    allocate();
    make_plans();
    setup_background();
    on_dt_changed();
    PTF_CUDA(cudaStreamSynchronize(stream));
  }

  ~MqgSolver() {
    drop_graphs();
    for (cufftHandle p : {plan_fwd, plan_inv, plan_fwd2, plan_inv2})
      if (p) cufftDestroy(p);
    if (own_stream) cudaStreamDestroy(own_stream);
  }

  // ---- setup ----
  void setup_tables() {
    std::vector<double> kx = wavenumbers(nx, d.Lx, true), ky = wavenumbers(ny, d.Ly, false);
    d_kx.alloc(kx.size(), &dev_bytes);
    d_ky.alloc(ky.size(), &dev_bytes);
    d_kz.alloc(1, &dev_bytes);
    PTF_CUDA(cudaMemcpy(d_kx.p, kx.data(), kx.size() * sizeof(double), cudaMemcpyHostToDevice));
    PTF_CUDA(cudaMemcpy(d_ky.p, ky.data(), ky.size() * sizeof(double), cudaMemcpyHostToDevice));
    PTF_CUDA(cudaMemset(d_kz.p, 0, sizeof(double)));
    ax = AxisTables{};
    ax.kx = d_kx.p;
    ax.ky = d_ky.p;
    ax.kz = d_kz.p;
    ax.ndim = 2;
    ax.kappa = 0.0;
    ax.eta = 0.0;
    ax.kappa_h = d.nu;       // L = −ν |k|^{2nν};  L[0,0] = −ν·0 = 0 for nν >= 1
    ax.n_kappa_h = d.n_nu;
    ax.fx = (d.Lx / (double)nx) / M_PI;
    ax.fy = (d.Ly / (double)ny) / M_PI;
    ax.f_inner = 2.0 / 3.0;
    ax.f_order = 4.0;
    ax.f_decay = -std::log(1e-15) / std::pow(1.0 - 2.0 / 3.0, 4.0);
    ax.dealias = 1;
    const double a = d.aliased_fraction;
    if (a > 0) {   // FF getaliasedwavenumbers
      ax.ax_lo = (int64_t)std::floor((1.0 - a) / 2.0 * (double)nx);
      ax.ay_lo = (int64_t)std::floor((1.0 - a) / 2.0 * (double)ny);
      ax.ay_hi = (int64_t)std::ceil((1.0 + a) / 2.0 * (double)ny);
    } else {       // FF dealias!: `grid.aliased_fraction == 0 && return nothing` — the mask is empty (examples/…:57)
      ax.ax_lo = nx / 2 + 1;
      ax.ay_lo = 0;
      ax.ay_hi = 0;
    }
    hkx = kx;
    hky = ky;
  }

  void allocate() {
    const int64_t ns = plane * NL, nr = pts * NL;
    sol.alloc(ns, &dev_bytes);
    PTF_CUDA(cudaMemsetAsync(sol.p, 0, sol.bytes(), stream));
    const int base = st.base;
    if (base == PTF_STEPPER_RK4 || base == PTF_STEPPER_ETDRK4) s1.alloc(ns, &dev_bytes);
    if (base == PTF_STEPPER_ETDRK4) s2.alloc(ns, &dev_bytes);
    if (base != PTF_STEPPER_FORWARD_EULER) {
      acc.alloc(ns, &dev_bytes);
      PTF_CUDA(cudaMemsetAsync(acc.p, 0, acc.bytes(), stream));
    }
    if (base == PTF_STEPPER_ETDRK4 || base == PTF_STEPPER_AB3) {
      n1.alloc(ns, &dev_bytes);
      PTF_CUDA(cudaMemsetAsync(n1.p, 0, n1.bytes(), stream));
    }
    if (base == PTF_STEPPER_ETDRK4)
      for (auto* b : {&cE, &cE2, &cZ, &cA, &cB, &cG}) b->alloc(plane, &dev_bytes);
    spec3.alloc(3 * ns, &dev_bytes);
    psih.alloc(ns, &dev_bytes);
    PTF_CUDA(cudaMemsetAsync(psih.p, 0, psih.bytes(), stream));
    phys3.alloc(3 * nr, &dev_bytes);
    vars3.alloc(3 * nr, &dev_bytes);   // u (perturbation), v, q of the last updatevars!
    PTF_CUDA(cudaMemsetAsync(vars3.p, 0, vars3.bytes(), stream));
    Qx.alloc(nr, &dev_bytes);
    Qy.alloc(nr, &dev_bytes);
    Ush.alloc(NL * ny, &dev_bytes);
    sinv.alloc((int64_t)NL * NL * plane, &dev_bytes);
    dF.alloc(NL * NL, &dev_bytes);
  }

  void make_plans() {
    long long n[2] = {ny, nx};
    size_t wf = 0, wi = 0;
    PTF_CUFFT(cufftCreate(&plan_fwd));
    PTF_CUFFT(cufftCreate(&plan_inv));
    PTF_CUFFT(cufftSetAutoAllocation(plan_fwd, 0));
    PTF_CUFFT(cufftSetAutoAllocation(plan_inv, 0));
    PTF_CUFFT(cufftMakePlanMany64(plan_fwd, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, 3 * NL, &wf));
    PTF_CUFFT(cufftMakePlanMany64(plan_inv, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, 3 * NL, &wi));
    // two-field batches: forward of (u q, v q) when the background term is spectral; inverse of (u, v) for updatevars!
    size_t wf2 = 0, wi2 = 0;
    PTF_CUFFT(cufftCreate(&plan_fwd2));
    PTF_CUFFT(cufftCreate(&plan_inv2));
    PTF_CUFFT(cufftSetAutoAllocation(plan_fwd2, 0));
    PTF_CUFFT(cufftSetAutoAllocation(plan_inv2, 0));
    PTF_CUFFT(cufftMakePlanMany64(plan_fwd2, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, 2 * NL, &wf2));
    PTF_CUFFT(cufftMakePlanMany64(plan_inv2, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, 2 * NL, &wi2));
    size_t w = std::max(std::max(wf, wi), std::max(wf2, wi2));
    work.alloc(w ? w : 16, &dev_bytes);
    for (cufftHandle p : {plan_fwd, plan_inv, plan_fwd2, plan_inv2}) {
      PTF_CUFFT(cufftSetWorkArea(p, work.p));
      PTF_CUFFT(cufftSetStream(p, stream));
    }
  }

  // MultiLayerQG.Params: stretching matrix, S⁻¹, background PV gradients Qx, Qy
  void setup_background() {
    // U(y, layer)
    hU.assign((size_t)NL * ny, 0.0);
    if (d.U) {
      for (int j = 0; j < NL; ++j)
        for (int64_t y = 0; y < ny; ++y) hU[j * ny + y] = d.U_is_profile ? d.U[j * ny + y] : d.U[j];
    }
    PTF_CUDA(cudaMemcpy(Ush.p, hU.data(), hU.size() * sizeof(double), cudaMemcpyHostToDevice));
    // Uyy = real(ifft(−l² fft(U))) along y: direct O(ny²) DFT, exact zero for uniform U
    std::vector<double> Uyy((size_t)NL * ny, 0.0);
    if (d.U && d.U_is_profile) {
      for (int j = 0; j < NL; ++j) {
        std::vector<double> re(ny), im(ny);
        for (int64_t m = 0; m < ny; ++m) {
          double sr = 0, si = 0;
          for (int64_t y = 0; y < ny; ++y) {
            const double ph = -2.0 * M_PI * (double)((m * y) % ny) / (double)ny;
            sr += hU[j * ny + y] * std::cos(ph);
            si += hU[j * ny + y] * std::sin(ph);
          }
          const double l2 = -hky[m] * hky[m];
          re[m] = l2 * sr;
          im[m] = l2 * si;
        }
        for (int64_t y = 0; y < ny; ++y) {
          double s = 0;
          for (int64_t m = 0; m < ny; ++m) {
            const double ph = 2.0 * M_PI * (double)((m * y) % ny) / (double)ny;
            s += re[m] * std::cos(ph) - im[m] * std::sin(ph);
          }
          Uyy[j * ny + y] = s / (double)ny;
        }
      }
    }
    // stretching matrix F (tridiagonal): sub-diagonal Fm, super-diagonal Fp, diagonal −([Fp;0]+[0;Fm])
    hF.assign((size_t)NL * NL, 0.0);
    std::vector<double> Fm(NL > 1 ? NL - 1 : 0), Fp(NL > 1 ? NL - 1 : 0);
    for (int j = 0; j + 1 < NL; ++j) {
      const double gp = d.b[j] - d.b[j + 1];   // reduced gravity at interface j+½
      PTF_REQUIRE(gp != 0.0 && d.H[j] > 0 && d.H[j + 1] > 0, "layer depths must be positive and buoyancies distinct");
      Fm[j] = d.f0 * d.f0 / (gp * d.H[j + 1]);
      Fp[j] = d.f0 * d.f0 / (gp * d.H[j]);
      hF[(j + 1) * NL + j] = Fm[j];
      hF[j * NL + j + 1] = Fp[j];
    }
    for (int j = 0; j < NL; ++j) hF[j * NL + j] = -((j < NL - 1 ? Fp[j] : 0.0) + (j > 0 ? Fm[j - 1] : 0.0));
    PTF_CUDA(cudaMemcpy(dF.p, hF.data(), hF.size() * sizeof(double), cudaMemcpyHostToDevice));
    // S⁻¹ per wavenumber; (0,0) -> 0
    {
      std::vector<double> h((size_t)NL * NL * plane);
      double S[PTF_MQG_MAX_LAYERS * PTF_MQG_MAX_LAYERS], Si[PTF_MQG_MAX_LAYERS * PTF_MQG_MAX_LAYERS];
      for (int64_t iy = 0; iy < ny; ++iy)
        for (int64_t ix = 0; ix < nkr; ++ix) {
          double k2 = hkx[ix] * hkx[ix] + hky[iy] * hky[iy];
          const bool zero = (k2 == 0.0);
          if (zero) k2 = 1.0;
          for (int r = 0; r < NL; ++r)
            for (int c = 0; c < NL; ++c) S[r * NL + c] = hF[r * NL + c] - (r == c ? k2 : 0.0);
          PTF_REQUIRE(invert_small(S, Si, NL), "stretching operator is singular");
          for (int e = 0; e < NL * NL; ++e) h[(size_t)e * plane + iy * nkr + ix] = zero ? 0.0 : Si[e];
        }
      PTF_CUDA(cudaMemcpy(sinv.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    // Qx, Qy in the reference's operation order: β − Uyy, + topography in the bottom layer, − stretching terms
    std::vector<double> etax((size_t)pts, 0.0), etay((size_t)pts, 0.0);
    if (d.eta) topography_gradients(etax, etay);
    std::vector<double> hQx((size_t)NL * pts, 0.0), hQy((size_t)NL * pts);
    for (int64_t p = 0; p < pts; ++p) hQx[(size_t)(NL - 1) * pts + p] += etax[p] + d.topographic_pv_gradient[0];
    for (int j = 0; j < NL; ++j)
      for (int64_t y = 0; y < ny; ++y) {
        const double Uj = hU[j * ny + y];
        double stretch_a = 0.0, stretch_b = 0.0;
        if (NL >= 2) {
          if (j == 0) {
            stretch_a = Fp[0] * (hU[1 * ny + y] - Uj);
          } else if (j == NL - 1) {
            stretch_a = Fm[NL - 2] * (hU[(NL - 2) * ny + y] - Uj);
          } else {
            stretch_a = Fp[j] * (hU[(j + 1) * ny + y] - Uj);
            stretch_b = Fm[j - 1] * (hU[(j - 1) * ny + y] - Uj);
          }
        }
        for (int64_t x = 0; x < nx; ++x) {
          double qy = d.beta - Uyy[j * ny + y];
          if (j == NL - 1) qy += etay[y * nx + x] + d.topographic_pv_gradient[1];
          hQy[(size_t)j * pts + y * nx + x] = (qy - stretch_a) - stretch_b;
        }
      }
    // background term applied spectrally when Qx == 0 and Qy is one constant per layer (no topography, uniform U):
    // rfft(v*Qy_j) = Qy_j * i kr psi-hat_j exactly, which saves one product field and a third of the forward transforms
    uniform_bg = std::getenv("PTF_MQG_NO_SPECTRAL_BG") == nullptr;
    for (int j = 0; j < NL && uniform_bg; ++j) {
      const double q0 = hQy[(size_t)j * pts];
      qy_uniform[j] = q0;
      for (int64_t p = 0; p < pts && uniform_bg; ++p)
        uniform_bg = hQx[(size_t)j * pts + p] == 0.0 && hQy[(size_t)j * pts + p] == q0;
    }
    PTF_CUDA(cudaMemcpy(Qx.p, hQx.data(), hQx.size() * sizeof(double), cudaMemcpyHostToDevice));
    PTF_CUDA(cudaMemcpy(Qy.p, hQy.data(), hQy.size() * sizeof(double), cudaMemcpyHostToDevice));
  }

  // ∂ₓη, ∂ᵧη of the periodic topographic PV through the device transforms (layer slot 0 of the scratch buffers)
  void topography_gradients(std::vector<double>& etax, std::vector<double>& etay) {
    PTF_CUDA(cudaMemsetAsync(phys3.p, 0, phys3.bytes(), stream));
    PTF_CUDA(cudaMemcpyAsync(phys3.p, d.eta, pts * sizeof(double), cudaMemcpyHostToDevice, stream));
    PTF_CUFFT(cufftExecD2Z(plan_fwd, phys3.p, Z(spec3.p)));
    std::vector<double2> eh((size_t)plane), ex((size_t)plane), ey((size_t)plane);
    PTF_CUDA(cudaMemcpyAsync(eh.data(), spec3.p, plane * sizeof(double2), cudaMemcpyDeviceToHost, stream));
    PTF_CUDA(cudaStreamSynchronize(stream));
    const double sc = 1.0 / (double)pts;
    for (int64_t iy = 0; iy < ny; ++iy)
      for (int64_t ix = 0; ix < nkr; ++ix) {
        const double2 e = eh[iy * nkr + ix];
        ex[iy * nkr + ix] = make_double2(-hkx[ix] * e.y * sc, hkx[ix] * e.x * sc);
        ey[iy * nkr + ix] = make_double2(-hky[iy] * e.y * sc, hky[iy] * e.x * sc);
      }
    PTF_CUDA(cudaMemsetAsync(spec3.p, 0, spec3.bytes(), stream));
    PTF_CUDA(cudaMemcpyAsync(spec3.p, ex.data(), plane * sizeof(double2), cudaMemcpyHostToDevice, stream));
    PTF_CUDA(cudaMemcpyAsync(spec3.p + plane, ey.data(), plane * sizeof(double2), cudaMemcpyHostToDevice, stream));
    PTF_CUFFT(cufftExecZ2D(plan_inv, Z(spec3.p), phys3.p));
    PTF_CUDA(cudaMemcpyAsync(etax.data(), phys3.p, pts * sizeof(double), cudaMemcpyDeviceToHost, stream));
    PTF_CUDA(cudaMemcpyAsync(etay.data(), phys3.p + pts, pts * sizeof(double), cudaMemcpyDeviceToHost, stream));
    PTF_CUDA(cudaStreamSynchronize(stream));
    lib_calls += 2;
  }

  void on_dt_changed() {
    drop_graphs();
    if (st.base == PTF_STEPPER_ETDRK4) {
      k_etd_coeffs<<<(unsigned)((plane + 255) / 256), 256, 0, stream>>>(cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p, ax, nkr, ny, 1,
                                                                      dt, 0);
      ++own_launches;
      PTF_CUDA(cudaGetLastError());
    }
  }

  // Adopt another stream (the coupled tracer's): everything is then ordered on one stream, no events needed.
  void set_stream(cudaStream_t s) {
    PTF_CUDA(cudaStreamSynchronize(stream));
    drop_graphs();
    stream = s;
    for (cufftHandle p : {plan_fwd, plan_inv, plan_fwd2, plan_inv2}) PTF_CUFFT(cufftSetStream(p, stream));
  }

  // ---- one stage ----
  static cufftDoubleComplex* Z(double2* p) { return reinterpret_cast<cufftDoubleComplex*>(p); }

  template <bool COMBINE, bool FRONT>
  void launch_spec(const MqgSpecArgs& a) {
    dim3 block(128, 1, 1), grid((unsigned)((nkr + 127) / 128), (unsigned)ny, 1);
    switch (NL) {
      case 1: k_mqg_spec<1, COMBINE, FRONT><<<grid, block, 0, stream>>>(a); break;
      case 2: k_mqg_spec<2, COMBINE, FRONT><<<grid, block, 0, stream>>>(a); break;
      case 3: k_mqg_spec<3, COMBINE, FRONT><<<grid, block, 0, stream>>>(a); break;
      default: k_mqg_spec<4, COMBINE, FRONT><<<grid, block, 0, stream>>>(a); break;
    }
    ++own_launches;
  }

  MqgSpecArgs spec_args(double2* state, int mode, double la = 0, double lb = 0, int llast = 0) {
    MqgSpecArgs a;
    a.state = state;
    a.spec3 = spec3.p;
    a.psih = psih.p;
    a.sinv = sinv.p;
    a.ax = ax;
    a.nkr = nkr;
    a.ny = ny;
    a.mu = d.mu;
    a.scale = 1.0 / (double)pts;
    a.uniform_bg = uniform_bg ? 1 : 0;
    for (int j = 0; j < PTF_MQG_MAX_LAYERS; ++j) a.qy[j] = qy_uniform[j];
    a.P = CombinePtrs{sol.p, s1.p, s2.p, acc.p, n1.p, cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p};
    a.A = CombineArgs{mode, st.filtered ? 1 : 0, dt, la, lb, llast};
    return a;
  }

  void physical_part() {   // û,v̂,q̂ -> u,v,q -> products -> P̂0,P̂1,P̂2
    PTF_CUFFT(cufftExecZ2D(plan_inv, Z(spec3.p), phys3.p));
    const int64_t half = pts * NL / 2;
    int blocks = (int)std::min<int64_t>((half + 255) / 256, 148 * 8);
    k_mqg_products<<<blocks, 256, 0, stream>>>(phys3.p, Qx.p, Qy.p, Ush.p, nx, ny, NL, uniform_bg ? 1 : 0);
    ++own_launches;
    if (uniform_bg)   // only u*q and v*q need a transform
      PTF_CUFFT(cufftExecD2Z(plan_fwd2, phys3.p + (int64_t)NL * pts, Z(spec3.p + (int64_t)NL * plane)));
    else
      PTF_CUFFT(cufftExecD2Z(plan_fwd, phys3.p, Z(spec3.p)));
    lib_calls += 2;
  }

  double2* slot(int s) { return s == 0 ? sol.p : (s == 1 ? s1.p : s2.p); }

  void enqueue_step(int variant) {
    static const double LA[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                 -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
    static const double LB[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                                 1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                                 2277821191437.0 / 14882151754819.0};
    struct Stage { int mode; double a, b; int last; };
    std::vector<Stage> stages;
    switch (st.base) {
      case PTF_STEPPER_RK4:
        stages = {{CM_RK4_S1, 0, 0, 0}, {CM_RK4_S2, 0, 0, 0}, {CM_RK4_S3, 0, 0, 0}, {CM_RK4_S4, 0, 0, 0}};
        break;
      case PTF_STEPPER_ETDRK4:
        stages = {{CM_ETD_S1, 0, 0, 0}, {CM_ETD_S2, 0, 0, 0}, {CM_ETD_S3, 0, 0, 0}, {CM_ETD_S4, 0, 0, 0}};
        break;
      case PTF_STEPPER_FORWARD_EULER: stages = {{CM_EULER, 0, 0, 0}}; break;
      case PTF_STEPPER_LSRK54:
        for (int i = 0; i < 5; ++i) stages.push_back({CM_LSRK, LA[i], LB[i], i == 4});
        break;
      case PTF_STEPPER_AB3: stages = {{variant == 1 ? CM_AB3_EULER : CM_AB3, 0, 0, 0}}; break;
    }
    launch_spec<false, true>(spec_args(sol.p, 0));   // calcN front of stage 1 on sol
    for (size_t s = 0; s < stages.size(); ++s) {
      physical_part();
      const Stage& g = stages[s];
      double2* nxt = slot(next_state_slot(g.mode));
      if (s + 1 < stages.size())
        launch_spec<true, true>(spec_args(nxt, g.mode, g.a, g.b, g.last));    // combine + next stage's front
      else
        launch_spec<true, false>(spec_args(nxt, g.mode, g.a, g.b, g.last));   // final update only
    }
  }

  void step_once() {
    const int variant = (st.base == PTF_STEPPER_AB3 && step < 3) ? 1 : 0;
    if (!d.use_graph) {
      enqueue_step(variant);
      PTF_CUDA(cudaGetLastError());
    } else {
      if (!graph_exec[variant]) {
        int64_t o0 = own_launches, l0 = lib_calls;
        cudaGraph_t graph = nullptr;
        PTF_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        try {
          enqueue_step(variant);
        } catch (...) {
          cudaStreamEndCapture(stream, &graph);
          if (graph) cudaGraphDestroy(graph);
          throw;
        }
        PTF_CUDA(cudaStreamEndCapture(stream, &graph));
        cudaError_t e = cudaGraphInstantiate(&graph_exec[variant], graph, 0);
        cudaGraphDestroy(graph);
        PTF_CUDA(e);
        per_step_own = own_launches - o0;
        per_step_lib = lib_calls - l0;
        own_launches = o0;
        lib_calls = l0;
      }
      PTF_CUDA(cudaGraphLaunch(graph_exec[variant], stream));
      own_launches += per_step_own;
      lib_calls += per_step_lib;
    }
    t += dt;   // FF stepforward!: clock.t += dt; clock.step += 1 after the update
    step += 1;
  }

  void steps(int64_t n) {
    for (int64_t i = 0; i < n; ++i) step_once();
  }

  void step_until(double t_stop) {
    if (st.base == PTF_STEPPER_ETDRK4)
      throw Error(PTF_EUNSUPPORTED, "step_until! requires an explicit stepper (not ETDRK4)");
    PTF_REQUIRE(t_stop > t, "stop time must be greater than the current time");
    const double dt0 = dt, interval = t_stop - t;
    const int64_t n = (int64_t)std::floor(interval / dt0);
    steps(n);
    const double rem = interval - (double)n * dt0;
    if (rem > 0) {
      dt = rem;
      on_dt_changed();
      steps(1);
      dt = dt0;
      on_dt_changed();
    }
    t = t_stop;
  }

  // MultiLayerQG.updatevars!: dealias!(sol) in place; psi-hat = S^-1 q-hat; u, v, q to physical space (device-resident).
  // full = false transforms u and v only (what the coupled tracer reads): used for all but the last iteration of the
  // in-library loops, whose intermediate q nobody can observe.
  void updatevars(bool full = true) {
    const int gi = full ? 1 : 0;
    auto body = [&]() {
      launch_spec<false, true>(spec_args(sol.p, 0));
      PTF_CUFFT(cufftExecZ2D(full ? plan_inv : plan_inv2, Z(spec3.p), vars3.p));
      ++lib_calls;
    };
    if (!d.use_graph) {
      body();
      PTF_CUDA(cudaGetLastError());
      return;
    }
    if (!upd_graph[gi]) {
      int64_t o0 = own_launches, l0 = lib_calls;
      cudaGraph_t graph = nullptr;
      PTF_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      try {
        body();
      } catch (...) {
        cudaStreamEndCapture(stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      PTF_CUDA(cudaStreamEndCapture(stream, &graph));
      cudaError_t e = cudaGraphInstantiate(&upd_graph[gi], graph, 0);
      cudaGraphDestroy(graph);
      PTF_CUDA(e);
      own_launches = o0;
      lib_calls = l0;
    }
    PTF_CUDA(cudaGraphLaunch(upd_graph[gi], stream));
    own_launches += 1;
    lib_calls += 1;
  }

  void drop_graphs() {
    for (auto& ge : graph_exec) {
      if (ge) cudaGraphExecDestroy(ge);
      ge = nullptr;
    }
    for (auto& ge : upd_graph) {
      if (ge) cudaGraphExecDestroy(ge);
      ge = nullptr;
    }
  }

  // ---- set / get ----
  void set_q(const double* q_host) {   // set_q!: q̂ = rfft(q); q̂[0,0,:] = 0; sol = q̂; updatevars!
    PTF_CUDA(cudaMemsetAsync(phys3.p, 0, phys3.bytes(), stream));
    PTF_CUDA(cudaMemcpyAsync(phys3.p, q_host, pts * NL * sizeof(double), cudaMemcpyHostToDevice, stream));
    PTF_CUFFT(cufftExecD2Z(plan_fwd, phys3.p, Z(spec3.p)));
    ++lib_calls;
    PTF_CUDA(cudaMemcpyAsync(sol.p, spec3.p, sol.bytes(), cudaMemcpyDeviceToDevice, stream));
    k_mqg_zero_mean<<<1, 32, 0, stream>>>(sol.p, plane, NL);
    ++own_launches;
    updatevars();
    PTF_CUDA(cudaStreamSynchronize(stream));
  }

  void set_psi(const double* psi_host) {   // set_ψ!: q̂ = S ψ̂ then set_q!
    PTF_CUDA(cudaMemsetAsync(phys3.p, 0, phys3.bytes(), stream));
    PTF_CUDA(cudaMemcpyAsync(phys3.p, psi_host, pts * NL * sizeof(double), cudaMemcpyHostToDevice, stream));
    PTF_CUFFT(cufftExecD2Z(plan_fwd, phys3.p, Z(spec3.p)));
    ++lib_calls;
    dim3 block(128, 1, 1), grid((unsigned)((nkr + 127) / 128), (unsigned)ny, 1);
    switch (NL) {
      case 1: k_mqg_pv_from_psi<1><<<grid, block, 0, stream>>>(spec3.p, sol.p, dF.p, ax, nkr, ny); break;
      case 2: k_mqg_pv_from_psi<2><<<grid, block, 0, stream>>>(spec3.p, sol.p, dF.p, ax, nkr, ny); break;
      case 3: k_mqg_pv_from_psi<3><<<grid, block, 0, stream>>>(spec3.p, sol.p, dF.p, ax, nkr, ny); break;
      default: k_mqg_pv_from_psi<4><<<grid, block, 0, stream>>>(spec3.p, sol.p, dF.p, ax, nkr, ny); break;
    }
    k_mqg_zero_mean<<<1, 32, 0, stream>>>(sol.p, plane, NL);
    own_launches += 2;
    updatevars();
    PTF_CUDA(cudaStreamSynchronize(stream));
  }

  void set_sol(const double* s_host) {
    PTF_CUDA(cudaMemcpyAsync(sol.p, s_host, sol.bytes(), cudaMemcpyHostToDevice, stream));
    PTF_CUDA(cudaStreamSynchronize(stream));
  }
  void get_sol(double* s_host) {
    PTF_CUDA(cudaMemcpyAsync(s_host, sol.p, sol.bytes(), cudaMemcpyDeviceToHost, stream));
    PTF_CUDA(cudaStreamSynchronize(stream));
  }

  // which: 0 = u (perturbation, as MQGprob.vars.u), 1 = v, 2 = q, 3 = ψ — the values of the last updatevars!
  void get_var(int which, double* host) {
    PTF_REQUIRE(which >= 0 && which <= 3, "variable index must be 0 (u), 1 (v), 2 (q) or 3 (psi)");
    const int64_t nr = pts * NL;
    if (which < 3) {
      PTF_CUDA(cudaMemcpyAsync(host, vars3.p + which * nr, nr * sizeof(double), cudaMemcpyDeviceToHost, stream));
    } else {
      const int64_t ns = plane * NL;
      k_mqg_scale_copy<<<(unsigned)std::min<int64_t>((ns + 255) / 256, 148 * 8), 256, 0, stream>>>(psih.p, spec3.p, ns,
                                                                                                   1.0 / (double)pts);
      ++own_launches;
      PTF_CUFFT(cufftExecZ2D(plan_inv, Z(spec3.p), phys3.p));
      ++lib_calls;
      PTF_CUDA(cudaMemcpyAsync(host, phys3.p, nr * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    PTF_CUDA(cudaStreamSynchronize(stream));
  }

  void get_background(double* qx, double* qy) {
    if (qx) PTF_CUDA(cudaMemcpy(qx, Qx.p, Qx.bytes(), cudaMemcpyDeviceToHost));
    if (qy) PTF_CUDA(cudaMemcpy(qy, Qy.p, Qy.bytes(), cudaMemcpyDeviceToHost));
  }

  const double* dev_u() const { return vars3.p; }
  const double* dev_v() const { return vars3.p + pts * NL; }

  ptf_mqg_desc d;
  int NL = 1;
  int64_t nx = 0, ny = 0, nkr = 0, plane = 0, pts = 0;
  StepperSpec st;
  double dt = 0.01, t = 0.0;
  int64_t step = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  int64_t own_launches = 0, lib_calls = 0, dev_bytes = 0;
  std::vector<double> hU, hF, hkx, hky;

 private:
  AxisTables ax;
  DevBuf<double> d_kx, d_ky, d_kz;
  DevBuf<double2> sol, s1, s2, acc, n1, spec3, psih;
  DevBuf<double> phys3, vars3, Qx, Qy, Ush, sinv, dF;
  DevBuf<double> cE, cE2, cZ, cA, cB, cG;
  DevBuf<char> work;
  cufftHandle plan_fwd = 0, plan_inv = 0, plan_fwd2 = 0, plan_inv2 = 0;
  bool uniform_bg = false;
  double qy_uniform[PTF_MQG_MAX_LAYERS] = {0, 0, 0, 0};
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  cudaGraphExec_t upd_graph[2] = {nullptr, nullptr};
  int64_t per_step_own = 0, per_step_lib = 0;
};

}  // namespace ptf

// -------------------------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------------------------
struct ptf_mqg_handle {
  std::unique_ptr<ptf::MqgSolver> solver;
  std::string last_error;
  int device = 0;
  std::vector<double> H, b, U, eta;   // owned copies of the descriptor's arrays
  std::vector<ptf_handle*> tracers;   // coupled tracer problems (their velocity pointers alias this solver's u, v)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_flow = nullptr, ev_tracer = nullptr;   // cross-stream ordering when a coupled tracer runs on another stream
};

namespace {

thread_local std::string g_mqg_create_error;

template <class F>
int32_t mqg_guarded(ptf_mqg_handle* h, F&& f) {
  try {
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) throw ptf::Error(PTF_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    f();
    return PTF_OK;
  } catch (const ptf::Error& e) {
    h->last_error = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    h->last_error = "host allocation failed";
    return PTF_ENOMEM;
  } catch (const std::exception& e) {
    h->last_error = e.what();
    return PTF_ECUDA;
  }
}

}  // namespace

namespace ptf {
// implemented in ptf_api.cu: point a layered tracer problem's velocity at device-resident fields / detach it
void couple_layered_velocity(ptf_handle* tracer, const double* u_dev, const double* v_dev, const double* U_host,
                             int64_t count);
void decouple_layered_velocity(ptf_handle* tracer);
cudaStream_t tracer_stream(ptf_handle* tracer);
int tracer_device(ptf_handle* tracer);
void tracer_geometry(ptf_handle* tracer, int64_t* nx, int64_t* ny, int64_t* nbatch, double* Lx, double* Ly, double* dt,
                     int* flow_kind, int* ndim);
void tracer_step_one(ptf_handle* tracer);
void tracer_set_mqg(ptf_handle* tracer, ptf_mqg_handle* m);
}  // namespace ptf

extern "C" {

int32_t ptf_mqg_desc_init(ptf_mqg_desc* d) {
  if (!d) return PTF_EINVAL;
  std::memset(d, 0, sizeof(*d));
  d->struct_size = (uint32_t)sizeof(ptf_mqg_desc);
  d->nlayers = 2;
  d->nx = 128;
  d->ny = 128;
  d->Lx = 2.0 * M_PI;
  d->Ly = 2.0 * M_PI;
  d->f0 = 1.0;
  d->beta = 0.0;
  d->mu = 0.0;
  d->nu = 0.0;
  d->n_nu = 1;
  d->dt = 0.01;
  d->stepper = PTF_STEPPER_RK4;
  d->aliased_fraction = 1.0 / 3.0;
  d->device = -1;
  d->use_graph = 1;
  return PTF_OK;
}

const char* ptf_mqg_last_error(const ptf_mqg_handle* h) { return h ? h->last_error.c_str() : g_mqg_create_error.c_str(); }

int32_t ptf_mqg_create(const ptf_mqg_desc* d, ptf_mqg_handle** out) {
  if (!out) return PTF_EINVAL;
  *out = nullptr;
  if (!d || d->struct_size != sizeof(ptf_mqg_desc)) {
    g_mqg_create_error = "descriptor is NULL or struct_size does not match this library (use ptf_mqg_desc_init)";
    return PTF_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_mqg_create_error = "no CUDA device: libptf_b200 has no CPU fallback";
    return PTF_ENODEVICE;
  }
  std::unique_ptr<ptf_mqg_handle> h(new ptf_mqg_handle());
  try {
    int dev = d->device;
    if (dev < 0) PTF_CUDA(cudaGetDevice(&dev));
    if (dev >= ndev) throw ptf::Error(PTF_EINVAL, "device ordinal out of range");
    PTF_CUDA(cudaSetDevice(dev));
    h->device = dev;
    ptf_mqg_desc dd = *d;
    dd.device = dev;
    const int NL = dd.nlayers;
    if (NL < 1 || NL > PTF_MQG_MAX_LAYERS) throw ptf::Error(PTF_EINVAL, "nlayers must be between 1 and 4");
    if (dd.ny <= 0 || dd.nx <= 0) throw ptf::Error(PTF_EINVAL, "grid sizes must be positive");
    if (!dd.H) throw ptf::Error(PTF_EINVAL, "layer depths H are required");
    h->H.assign(dd.H, dd.H + NL);
    dd.H = h->H.data();
    if (dd.b) {
      h->b.assign(dd.b, dd.b + NL);
      dd.b = h->b.data();
    }
    if (dd.U) {
      h->U.assign(dd.U, dd.U + (dd.U_is_profile ? (size_t)NL * dd.ny : (size_t)NL));
      dd.U = h->U.data();
    }
    if (dd.eta) {
      h->eta.assign(dd.eta, dd.eta + (size_t)dd.nx * dd.ny);
      dd.eta = h->eta.data();
    }
    h->solver.reset(new ptf::MqgSolver(dd));
    PTF_CUDA(cudaEventCreate(&h->ev0));
    PTF_CUDA(cudaEventCreate(&h->ev1));
    PTF_CUDA(cudaEventCreateWithFlags(&h->ev_flow, cudaEventDisableTiming));
    PTF_CUDA(cudaEventCreateWithFlags(&h->ev_tracer, cudaEventDisableTiming));
  } catch (const ptf::Error& e) {
    g_mqg_create_error = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    g_mqg_create_error = "host allocation failed";
    return PTF_ENOMEM;
  } catch (const std::exception& e) {
    g_mqg_create_error = e.what();
    return PTF_ECUDA;
  }
  *out = h.release();
  return PTF_OK;
}

int32_t ptf_mqg_destroy(ptf_mqg_handle* h) {
  if (!h) return PTF_OK;
  cudaSetDevice(h->device);
  for (ptf_handle* t : h->tracers) {
    ptf::decouple_layered_velocity(t);
    ptf::tracer_set_mqg(t, nullptr);
  }
  if (h->solver) cudaStreamSynchronize(h->solver->stream);
  h->solver.reset();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_flow) cudaEventDestroy(h->ev_flow);
  if (h->ev_tracer) cudaEventDestroy(h->ev_tracer);
  delete h;
  return PTF_OK;
}

int32_t ptf_mqg_set_q(ptf_mqg_handle* h, const double* q_host) {
  if (!h || !q_host) return PTF_EINVAL;
  return mqg_guarded(h, [&]() { h->solver->set_q(q_host); });
}

int32_t ptf_mqg_set_psi(ptf_mqg_handle* h, const double* psi_host) {
  if (!h || !psi_host) return PTF_EINVAL;
  return mqg_guarded(h, [&]() { h->solver->set_psi(psi_host); });
}

int32_t ptf_mqg_set_sol(ptf_mqg_handle* h, const double* sol_host_interleaved) {
  if (!h || !sol_host_interleaved) return PTF_EINVAL;
  return mqg_guarded(h, [&]() { h->solver->set_sol(sol_host_interleaved); });
}

int32_t ptf_mqg_get_sol(ptf_mqg_handle* h, double* sol_host_interleaved) {
  if (!h || !sol_host_interleaved) return PTF_EINVAL;
  return mqg_guarded(h, [&]() { h->solver->get_sol(sol_host_interleaved); });
}

int32_t ptf_mqg_updatevars(ptf_mqg_handle* h) {
  if (!h) return PTF_EINVAL;
  return mqg_guarded(h, [&]() {
    h->solver->updatevars();
    PTF_CUDA(cudaStreamSynchronize(h->solver->stream));
  });
}

int32_t ptf_mqg_get_var(ptf_mqg_handle* h, int32_t which, double* host) {
  if (!h || !host) return PTF_EINVAL;
  return mqg_guarded(h, [&]() { h->solver->get_var(which, host); });
}

int32_t ptf_mqg_get_background(ptf_mqg_handle* h, double* Qx_host, double* Qy_host) {
  if (!h) return PTF_EINVAL;
  return mqg_guarded(h, [&]() { h->solver->get_background(Qx_host, Qy_host); });
}

int32_t ptf_mqg_step(ptf_mqg_handle* h, int64_t nsteps) {
  if (!h || nsteps < 0) return PTF_EINVAL;
  return mqg_guarded(h, [&]() {
    h->solver->steps(nsteps);
    PTF_CUDA(cudaStreamSynchronize(h->solver->stream));
  });
}

int32_t ptf_mqg_step_until(ptf_mqg_handle* h, double t_stop) {
  if (!h) return PTF_EINVAL;
  return mqg_guarded(h, [&]() {
    h->solver->step_until(t_stop);
    PTF_CUDA(cudaStreamSynchronize(h->solver->stream));
  });
}

int32_t ptf_mqg_step_timed(ptf_mqg_handle* h, int64_t nsteps, int32_t with_updatevars, float* device_ms) {
  if (!h || nsteps < 0) return PTF_EINVAL;
  return mqg_guarded(h, [&]() {
    PTF_CUDA(cudaEventRecord(h->ev0, h->solver->stream));
    for (int64_t i = 0; i < nsteps; ++i) {
      h->solver->steps(1);
      if (with_updatevars) h->solver->updatevars(i + 1 == nsteps);
    }
    PTF_CUDA(cudaEventRecord(h->ev1, h->solver->stream));
    PTF_CUDA(cudaEventSynchronize(h->ev1));
    float ms = 0;
    PTF_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (device_ms) *device_ms = ms;
  });
}

int32_t ptf_mqg_get_clock(const ptf_mqg_handle* h, double* t, int64_t* step, double* dt) {
  if (!h) return PTF_EINVAL;
  if (t) *t = h->solver->t;
  if (step) *step = h->solver->step;
  if (dt) *dt = h->solver->dt;
  return PTF_OK;
}

int32_t ptf_mqg_set_dt(ptf_mqg_handle* h, double dt) {
  if (!h || !(dt > 0)) return PTF_EINVAL;
  return mqg_guarded(h, [&]() {
    if (dt != h->solver->dt) {
      PTF_CUDA(cudaStreamSynchronize(h->solver->stream));
      h->solver->dt = dt;
      h->solver->on_dt_changed();
    }
  });
}

int32_t ptf_mqg_launch_count(const ptf_mqg_handle* h, int64_t* own_kernels, int64_t* library_calls) {
  if (!h) return PTF_EINVAL;
  if (own_kernels) *own_kernels = h->solver->own_launches;
  if (library_calls) *library_calls = h->solver->lib_calls;
  return PTF_OK;
}

/* Problem(MQGprob; κ, η, stepper, tracer_release_time) TAD.jl:225-250: the layered tracer reads MQGprob.vars.u/.v and
 * MQGprob.params.U on the device from now on (TAD.jl:795-796); both problems are ordered on the tracer's stream. */
int32_t ptf_mqg_couple(ptf_mqg_handle* m, ptf_handle* tracer) {
  if (!m || !tracer) return PTF_EINVAL;
  return mqg_guarded(m, [&]() {
    ptf::MqgSolver& s = *m->solver;
    int64_t tnx, tny, tnb;
    double tLx, tLy, tdt;
    int kind, ndim;
    ptf::tracer_geometry(tracer, &tnx, &tny, &tnb, &tLx, &tLy, &tdt, &kind, &ndim);
    PTF_REQUIRE(ptf::tracer_device(tracer) == m->device, "tracer and flow must live on the same device");
    PTF_REQUIRE(kind == PTF_FLOW_LAYERED && ndim == 2, "the tracer problem must be a 2-D PTF_FLOW_LAYERED problem");
    PTF_REQUIRE(tnx == s.nx && tny == s.ny && tnb == s.NL, "tracer grid / layer count differs from the flow's");
    PTF_REQUIRE(tLx == s.d.Lx && tLy == s.d.Ly, "tracer domain differs from the flow's");
    s.set_stream(ptf::tracer_stream(tracer));
    s.updatevars();   // ConstDiffTurbulentFlowParams calls MultiLayerQG.updatevars!(MQGprob)   TAD.jl:488
    ptf::couple_layered_velocity(tracer, s.dev_u(), s.dev_v(), s.hU.data(), s.pts * s.NL);
    ptf::tracer_set_mqg(tracer, m);
    bool known = false;
    for (ptf_handle* t : m->tracers) known = known || (t == tracer);
    if (!known) m->tracers.push_back(tracer);
    PTF_CUDA(cudaStreamSynchronize(s.stream));
  });
}

int32_t ptf_mqg_forget_tracer(ptf_mqg_handle* m, ptf_handle* tracer) {
  if (!m) return PTF_EINVAL;
  for (size_t i = 0; i < m->tracers.size(); ++i)
    if (m->tracers[i] == tracer) {
      m->tracers.erase(m->tracers.begin() + i);
      break;
    }
  // the tracer's stream is about to disappear: go back to the solver's own stream
  if (m->solver && tracer && m->solver->stream == ptf::tracer_stream(tracer)) {
    cudaSetDevice(m->device);
    try {
      m->solver->set_stream(m->solver->own_stream);
    } catch (...) {
    }
  }
  return PTF_OK;
}

/* The loop of examples/turbulent_advection-diffusion.jl:149-151 —
 *   stepforward!(ADprob); stepforward!(params.MQGprob); MultiLayerQG.updatevars!(params.MQGprob)
 * — nsteps times, enqueued on one stream without host synchronisation in between. */
int32_t ptf_mqg_step_coupled(ptf_mqg_handle* m, ptf_handle* tracer, int64_t nsteps, float* device_ms) {
  if (!m || !tracer || nsteps < 0) return PTF_EINVAL;
  return mqg_guarded(m, [&]() {
    ptf::MqgSolver& s = *m->solver;
    bool known = false;
    for (ptf_handle* t : m->tracers) known = known || (t == tracer);
    PTF_REQUIRE(known, "ptf_mqg_step_coupled: the tracer is not coupled to this flow (ptf_mqg_couple)");
    // The flow runs on the stream of the tracer that coupled LAST; any other coupled tracer has its own stream and
    // reads the flow's u, v buffers from there: order the two streams explicitly in that case.
    cudaStream_t ts = ptf::tracer_stream(tracer);
    const bool cross = ts != s.stream;
    PTF_CUDA(cudaEventRecord(m->ev0, s.stream));
    for (int64_t i = 0; i < nsteps; ++i) {
      if (cross) {   // the tracer's product kernels must see the u, v of the flow's last updatevars!
        PTF_CUDA(cudaEventRecord(m->ev_flow, s.stream));
        PTF_CUDA(cudaStreamWaitEvent(ts, m->ev_flow, 0));
      }
      ptf::tracer_step_one(tracer);
      if (cross) {   // ... and the flow must not overwrite them (calcN_advection! uses vars.u, vars.v as scratch) before
        PTF_CUDA(cudaEventRecord(m->ev_tracer, ts));   // the tracer step has read them
        PTF_CUDA(cudaStreamWaitEvent(s.stream, m->ev_tracer, 0));
      }
      s.steps(1);
      s.updatevars(i + 1 == nsteps);   // intermediate iterations only need u, v (the tracer's inputs)
    }
    PTF_CUDA(cudaEventRecord(m->ev1, s.stream));
    PTF_CUDA(cudaEventSynchronize(m->ev1));
    float ms = 0;
    PTF_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
    if (device_ms) *device_ms = ms;
  });
}

}  // extern "C"
