// Internal declarations shared by the C-ABI layer and the engines.  Not part of the public boundary.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>

#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ptf_b200.h"

namespace ptf {

struct Error : std::runtime_error {
  int32_t code;
  Error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PTF_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw ::ptf::Error(_e == cudaErrorMemoryAllocation ? PTF_ENOMEM : PTF_ECUDA,                  \
                         std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                             std::to_string(__LINE__) + ")");                                       \
  } while (0)

#define PTF_CUFFT(expr)                                                                             \
  do {                                                                                              \
    cufftResult _r = (expr);                                                                        \
    if (_r != CUFFT_SUCCESS)                                                                        \
      throw ::ptf::Error(_r == CUFFT_ALLOC_FAILED ? PTF_ENOMEM : PTF_ECUFFT,                        \
                         std::string(#expr) + ": cufft error " + std::to_string((int)_r) + " (" +   \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");                      \
  } while (0)

#define PTF_REQUIRE(cond, msg)                                     \
  do {                                                             \
    if (!(cond)) throw ::ptf::Error(PTF_EINVAL, std::string(msg)); \
  } while (0)

// RAII device buffer
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void alloc(size_t count, int64_t* tally = nullptr) {
    release();
    if (count == 0) return;
    PTF_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    n = count;
    if (tally) *tally += (int64_t)(count * sizeof(T));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// Problem geometry + physics, resolved from the descriptor (host side).
struct Geometry {
  int ndim = 2;
  int64_t nx = 1, ny = 1, nz = 1;  // global physical extents (unused axes = 1)
  int64_t nkr = 1;                 // nx/2+1
  int64_t B = 1;                   // local batch (layers / members held by this rank)
  int64_t Bglobal = 1, Boffset = 0;
  double Lx = 0, Ly = 0, Lz = 0;
  int64_t npts() const { return nx * ny * nz; }   // real points per member (global)
  int64_t nspec() const { return nkr * ny * nz; } // complex coefficients per member (global)
  // slab decomposition (3-D, one process per GPU): this rank holds z-planes [zoff, zoff+nzl) of physical fields
  // and ky-rows [yoff, yoff+nyl) of spectral fields.  Single GPU: nzl = nz, nyl = ny.
  bool slab = false;
  int P = 1, rank = 0;
  int64_t nzl = 1, nyl = 1, zoff = 0, yoff = 0;
  // 2-D slab decomposition (engine_slab2d.cu): physical rows [ypoff, ypoff+nyp), spectral columns kr in
  // [koff, koff+kvalid) stored in kc = ceil(nkr/P) padded columns
  bool slab2d = false;
  int64_t nyp = 1, ypoff = 0, kc = 1, koff = 0, kvalid = 1;
  int64_t lpts() const { return slab2d ? nx * nyp : nx * ny * nzl; }       // local real points per member
  int64_t lspec() const { return slab2d ? kvalid * ny : nkr * nyl * nz; }  // local complex coefficients per member
  std::vector<double> kx, ky, kz;                 // wavenumbers (kz/ky size 1 == {0} on unused axes)
};

// Per-step coefficient set handed to combine kernels.
struct StepperSpec {
  int base = PTF_STEPPER_RK4;
  bool filtered = false;
  int nstages() const {
    switch (base) {
      case PTF_STEPPER_FORWARD_EULER: return 1;
      case PTF_STEPPER_RK4: return 4;
      case PTF_STEPPER_ETDRK4: return 4;
      case PTF_STEPPER_LSRK54: return 5;
      case PTF_STEPPER_AB3: return 1;
    }
    return 0;
  }
};

// Device-resident per-axis tables used by every pointwise kernel (no full-size L / filter arrays are stored:
// L and the filter are evaluated in registers from these, 8*(nkr+ny+nz) bytes instead of 4 B/pt each).
struct AxisTables {
  const double* kx = nullptr;  // [nkr]
  const double* ky = nullptr;  // [ny]
  const double* kz = nullptr;  // [nz]
  // the tables' generating constants: k_a[i] = (double)j * c_a with j = i (r2c axis) or fftfreq order (j = i - n for
  // i >= n/2), exactly as the host builds them — the fused kernels recompute their 16 wavenumbers per thread from
  // these instead of loading them (bit-identical; no L1/L2 latency in front of the combine)
  double cx = 0, cy = 0, cz = 0;
  int nyq_sign = -1;           // > 0: the y/z Nyquist wavenumber is stored positive (ptf_desc.nyquist_sign)
  // kappa*kx^2 etc. are formed in-kernel in the reference's operation order.
  double kappa = 0, eta = 0, iota = 0, kappa_h = 0;
  int n_kappa_h = 0;
  // filter
  double fx = 0, fy = 0, fz = 0;  // dx/pi, dy/pi, dz/pi
  double f_inner = 2.0 / 3.0, f_decay = 0, f_order = 4;
  // dealias (0-based half-open index ranges that are zeroed), enabled flag
  int dealias = 0;
  int64_t ax_lo = 0, ay_lo = 0, ay_hi = 0, az_lo = 0, az_hi = 0;
  int ndim = 2;
};

class Engine {
 public:
  virtual ~Engine() {}
  virtual const char* name() const = 0;
  virtual int id() const = 0;
  virtual void set_velocity(int comp, const double* host, int64_t count) = 0;
  virtual void set_velocity_separable(int comp, int nterms, const double* xt, const double* yt, const double* zt,
                                      const double* coeff0) = 0;
  virtual void set_velocity_coeffs(int comp, int nterms, const double* a) = 0;
  virtual void set_layered_shift(const double* U) = 0;  // U(y, layer) added to u; nullptr = none
  virtual void set_velocity_external(int comp, const double* dev, int64_t count) {  // alias a device field
    (void)comp; (void)dev; (void)count;
    throw Error(PTF_EUNSUPPORTED, "this engine cannot read device-resident velocity fields");
  }
  virtual void set_velocity_expr(int comp, const char* expr) {  // PTF_FLOW_EXPR
    (void)comp; (void)expr;
    throw Error(PTF_EUNSUPPORTED, "this engine does not support expression flows (PTF_FLOW_EXPR runs on the cuFFT pipelines)");
  }
  virtual void set_flow_time(double t) { (void)t; }  // PTF_FLOW_EXPR: clock.t of the step about to run
  virtual void set_c(const double* c_host, bool replicate) = 0;
  virtual void get_c(double* c_host) = 0;
  virtual void set_sol(const double* s_host) = 0;
  virtual void get_sol(double* s_host) = 0;
  virtual void step_once(int64_t step_index) = 0;  // enqueue one full time step on the stream
  // enqueue n steps at once when the engine can (velocity unchanged between them); false = not supported
  virtual bool step_many(int64_t first_step, int64_t n) {
    (void)first_step;
    (void)n;
    return false;
  }
  virtual void on_dt_changed() = 0;
  virtual void diag(double* mean_c, double* var_c, double* max_abs_sol) = 0;
  virtual float time_kernel(const char* name, int reps) = 0;
  virtual cudaStream_t stream() const = 0;
  int64_t own_launches = 0, lib_calls = 0, dev_bytes = 0;
};

struct Context {
  ptf_desc d;
  Geometry g;
  StepperSpec st;
  AxisTables ax;  // device pointers filled by the engine's owner
  DevBuf<double> d_kx, d_ky, d_kz;
  int device = 0;
  double dt = 0.01, t = 0.0;
  int64_t step = 0;
  cudaStream_t stream = nullptr;
  int64_t table_bytes = 0;
  void* nccl_comm = nullptr;  // ncclComm_t when decomposition == PTF_DECOMP_SLAB
};

std::unique_ptr<Engine> make_cufft_engine(Context& ctx);
std::unique_ptr<Engine> make_fused_engine(Context& ctx);  // throws PTF_EUNSUPPORTED when the grid does not qualify
bool fused_engine_supports(const Context& ctx, std::string* why);
std::unique_ptr<Engine> make_fused3d_engine(Context& ctx);  // fused 3-D engine (also slab-decomposed)
bool fused3d_engine_supports(const Context& ctx, std::string* why);
void selftest_fft(int n, int dir, int count, const double* in_host, double* out_host);
std::unique_ptr<Engine> make_slab2d_engine(Context& ctx);
std::unique_ptr<Engine> make_fused1d_engine(Context& ctx);
bool fused1d_engine_supports(const Context& ctx, std::string* why);

}  // namespace ptf

struct ptf_mqg_handle;
struct ptf_handle {
  ptf::Context ctx;
  ptf_mqg_handle* mqg = nullptr;  // coupled MultiLayerQG flow (ptf_mqg_couple), not owned
  std::unique_ptr<ptf::Engine> engine;
  std::string last_error;
  ptf_velocity_fn vel_fn = nullptr;
  void* vel_user = nullptr;
  ptf_coeff_fn coeff_fn = nullptr;
  void* coeff_user = nullptr;
  int sep_nterms[3] = {0, 0, 0};
  double* pinned_vel[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
