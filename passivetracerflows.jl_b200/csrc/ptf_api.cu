// C-ABI layer of libptf_b200.so (see include/ptf_b200.h for the reference interfaces each entry point replaces).
// Host-side setup restates FourierFlows' grid construction (FF domains.jl: x0 = -L/2, rfftfreq / fftfreq
// wavenumbers with a negative y/z Nyquist) — it is part of the product, computed here, not in any oracle.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>

#include "ptf_internal.h"

#ifdef PTF_WITH_NCCL
#include <nccl.h>
#endif

using namespace ptf;

namespace {

thread_local std::string g_create_error;

#ifdef PTF_WITH_NCCL
// One NCCL communicator per (ncclUniqueId, rank) in the process, shared by every handle created with that id
// (a unique id can seed ncclCommInitRank only once).
struct CommEntry {
  ncclComm_t comm;
  int refs;
};
std::mutex g_comm_mu;
std::map<std::string, CommEntry> g_comms;

ncclComm_t acquire_comm(const uint8_t id_bytes[128], int nranks, int rank) {
  std::string key(reinterpret_cast<const char*>(id_bytes), 128);
  key += ":" + std::to_string(nranks) + ":" + std::to_string(rank);
  std::lock_guard<std::mutex> lk(g_comm_mu);
  auto it = g_comms.find(key);
  if (it != g_comms.end()) {
    it->second.refs++;
    return it->second.comm;
  }
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, 128);
  ncclComm_t comm;
  ncclResult_t r = ncclCommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) throw Error(PTF_ENCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
  g_comms[key] = CommEntry{comm, 1};
  return comm;
}

void release_comm(ncclComm_t comm) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  for (auto it = g_comms.begin(); it != g_comms.end(); ++it)
    if (it->second.comm == comm) {
      // communicators stay alive for the life of the process once created: a later handle with the same id reuses
      // them (destroying and re-initialising with a spent id is not possible)
      if (it->second.refs > 0) it->second.refs--;
      return;
    }
}
#endif

const char* status_name(int32_t s) {
  switch (s) {
    case PTF_OK: return "PTF_OK";
    case PTF_EINVAL: return "PTF_EINVAL: invalid argument";
    case PTF_ECUDA: return "PTF_ECUDA: CUDA runtime error";
    case PTF_ECUFFT: return "PTF_ECUFFT: cuFFT error";
    case PTF_ENCCL: return "PTF_ENCCL: NCCL error";
    case PTF_ENOMEM: return "PTF_ENOMEM: out of memory";
    case PTF_EUNSUPPORTED: return "PTF_EUNSUPPORTED: not supported by this build / configuration";
    case PTF_ENODEVICE: return "PTF_ENODEVICE: no CUDA device (this library has no CPU fallback)";
  }
  return "unknown status";
}

template <class F>
int32_t guarded(ptf_handle* h, F&& f) {
  try {
    if (h) PTF_CUDA(cudaSetDevice(h->ctx.device));
    f();
    return PTF_OK;
  } catch (const Error& e) {
    if (h) h->last_error = e.what();
    else g_create_error = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    if (h) h->last_error = "host allocation failed";
    else g_create_error = "host allocation failed";
    return PTF_ENOMEM;
  } catch (const std::exception& e) {
    if (h) h->last_error = e.what();
    else g_create_error = e.what();
    return PTF_EINVAL;
  } catch (...) {
    if (h) h->last_error = "unknown exception";
    else g_create_error = "unknown exception";
    return PTF_EINVAL;
  }
}

// FF domains.jl: kr = rfftfreq(nx, 2pi/Lx*nx), l = fftfreq(ny, 2pi/Ly*ny)
std::vector<double> rfft_wavenumbers(int64_t n, double L) {
  std::vector<double> k(n / 2 + 1);
  double fs_over_n = (2.0 * M_PI / L * (double)n) / (double)n;
  for (int64_t i = 0; i <= n / 2; ++i) k[i] = (double)i * fs_over_n;
  return k;
}
std::vector<double> fft_wavenumbers(int64_t n, double L, int nyquist_sign) {
  std::vector<double> k(n);
  double fs_over_n = (2.0 * M_PI / L * (double)n) / (double)n;
  for (int64_t i = 0; i < n; ++i) {
    int64_t j = (i < n / 2) ? i : i - n;
    k[i] = (double)j * fs_over_n;
  }
  if (nyquist_sign > 0 && n >= 2) k[n / 2] = -k[n / 2];
  return k;
}

void upload(DevBuf<double>& b, const std::vector<double>& v, int64_t* tally) {
  b.alloc(v.size(), tally);
  PTF_CUDA(cudaMemcpy(b.p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
}

void build_context(ptf_handle* h, const ptf_desc* d) {
  Context& c = h->ctx;
  PTF_REQUIRE(d != nullptr, "descriptor is NULL");
  PTF_REQUIRE(d->struct_size == sizeof(ptf_desc), "ptf_desc.struct_size mismatch (use ptf_desc_init)");
  c.d = *d;
  PTF_REQUIRE(d->ndim >= 1 && d->ndim <= 3, "ndim must be 1, 2 or 3");
  for (int a = 0; a < d->ndim; ++a) {
    PTF_REQUIRE(d->n[a] >= 2 && d->n[a] % 2 == 0, "grid sizes must be even and >= 2");
    PTF_REQUIRE(d->L[a] > 0, "domain lengths must be positive");
  }
  PTF_REQUIRE(d->nbatch >= 1, "nbatch must be >= 1");
  PTF_REQUIRE(std::isfinite(d->dt) && d->dt > 0, "dt must be positive");
  PTF_REQUIRE(d->n_kappa_h >= 0, "n_kappa_h must be >= 0");
  PTF_REQUIRE(d->nranks >= 1 && d->rank >= 0 && d->rank < d->nranks, "bad rank / nranks");
  int base = d->stepper & ~PTF_STEPPER_FILTERED;
  PTF_REQUIRE(base >= PTF_STEPPER_FORWARD_EULER && base <= PTF_STEPPER_AB3, "unknown stepper");
  PTF_REQUIRE(d->flow_kind >= PTF_FLOW_STEADY && d->flow_kind <= PTF_FLOW_EXPR, "unknown flow_kind");
  c.st.base = base;
  c.st.filtered = (d->stepper & PTF_STEPPER_FILTERED) != 0;

  Geometry& g = c.g;
  g.ndim = d->ndim;
  g.nx = d->n[0];
  g.ny = d->ndim >= 2 ? d->n[1] : 1;
  g.nz = d->ndim >= 3 ? d->n[2] : 1;
  g.nkr = g.nx / 2 + 1;
  g.Lx = d->L[0];
  g.Ly = d->ndim >= 2 ? d->L[1] : 1.0;
  g.Lz = d->ndim >= 3 ? d->L[2] : 1.0;
  g.Bglobal = d->nbatch;
  g.B = d->nbatch;
  g.Boffset = 0;
  if (d->nranks > 1) {
    if (d->decomposition == PTF_DECOMP_BATCH) {
      PTF_REQUIRE(d->nbatch % d->nranks == 0, "nbatch must be divisible by nranks for PTF_DECOMP_BATCH");
      g.B = d->nbatch / d->nranks;
      g.Boffset = g.B * d->rank;
    } else if (d->decomposition == PTF_DECOMP_SLAB && d->ndim == 2) {
      PTF_REQUIRE(d->nbatch == 1, "slab decomposition needs nbatch == 1");
      PTF_REQUIRE(g.ny % d->nranks == 0, "ny must be divisible by nranks");
      PTF_REQUIRE(d->flow_kind != PTF_FLOW_LAYERED, "layered flows are not slab-decomposed (use PTF_DECOMP_BATCH)");
      g.slab2d = true;
      g.P = d->nranks;
      g.rank = d->rank;
    } else if (d->decomposition == PTF_DECOMP_SLAB) {
      PTF_REQUIRE(d->ndim == 3, "slab decomposition is implemented for 2-D and 3-D problems");
      PTF_REQUIRE(d->nbatch == 1, "slab decomposition needs nbatch == 1");
      PTF_REQUIRE(g.ny % d->nranks == 0 && g.nz % d->nranks == 0, "ny and nz must be divisible by nranks");
      g.slab = true;
      g.P = d->nranks;
      g.rank = d->rank;
    }  // PTF_DECOMP_NONE: independent replicas
  }
  // single-process test hook: run the 2-D slab engine with P = 1 (every kernel but the NCCL exchange)
  if (!g.slab2d && d->nranks == 1 && d->ndim == 2 && d->nbatch == 1 && d->decomposition == PTF_DECOMP_SLAB &&
      d->flow_kind != PTF_FLOW_LAYERED) {
    g.slab2d = true;
    g.P = 1;
    g.rank = 0;
  }
  if (g.slab2d) {
    g.nyp = g.ny / g.P;
    g.ypoff = g.nyp * g.rank;
    g.kc = (g.nkr + g.P - 1) / g.P;
    g.koff = g.kc * g.rank;
    g.kvalid = std::max<int64_t>(0, std::min<int64_t>(g.kc, g.nkr - g.koff));
  }
  g.nzl = g.slab ? g.nz / g.P : g.nz;
  g.nyl = g.slab ? g.ny / g.P : g.ny;
  g.zoff = g.slab ? g.nzl * g.rank : 0;
  g.yoff = g.slab ? g.nyl * g.rank : 0;
  g.kx = rfft_wavenumbers(g.nx, g.Lx);
  g.ky = d->ndim >= 2 ? fft_wavenumbers(g.ny, g.Ly, d->nyquist_sign) : std::vector<double>{0.0};
  g.kz = d->ndim >= 3 ? fft_wavenumbers(g.nz, g.Lz, d->nyquist_sign) : std::vector<double>{0.0};

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(PTF_ENODEVICE, "no CUDA device visible: libptf_b200 has no CPU fallback");
  }
  int dev = d->device;
  if (dev < 0) PTF_CUDA(cudaGetDevice(&dev));
  PTF_REQUIRE(dev < ndev, "device ordinal out of range");
  c.device = dev;
  PTF_CUDA(cudaSetDevice(dev));
  PTF_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  c.dt = d->dt;
  c.t = 0.0;
  c.step = 0;

  if (g.slab || (g.slab2d && g.P > 1)) {
#ifdef PTF_WITH_NCCL
    c.nccl_comm = acquire_comm(d->nccl_id, d->nranks, d->rank);
#else
    throw Error(PTF_EUNSUPPORTED, "built without NCCL");
#endif
  }
  upload(c.d_kx, g.kx, &c.table_bytes);
  upload(c.d_ky, g.ky, &c.table_bytes);
  upload(c.d_kz, g.kz, &c.table_bytes);
  AxisTables& ax = c.ax;
  ax.kx = c.d_kx.p;
  ax.ky = c.d_ky.p;
  ax.kz = c.d_kz.p;
  ax.cx = (2.0 * M_PI / g.Lx * (double)g.nx) / (double)g.nx;   // the same expressions as rfft_/fft_wavenumbers
  ax.cy = (2.0 * M_PI / g.Ly * (double)g.ny) / (double)g.ny;
  ax.cz = (2.0 * M_PI / g.Lz * (double)g.nz) / (double)g.nz;
  ax.nyq_sign = d->nyquist_sign;
  ax.kappa = d->kappa[0];
  ax.eta = d->kappa[1];
  ax.iota = d->kappa[2];
  ax.kappa_h = d->kappa_h;
  ax.n_kappa_h = d->n_kappa_h;
  ax.ndim = d->ndim;
  ax.fx = (g.Lx / (double)g.nx) / M_PI;
  ax.fy = (g.Ly / (double)g.ny) / M_PI;
  ax.fz = (g.Lz / (double)g.nz) / M_PI;
  ax.f_inner = d->filter_inner_k;
  ax.f_order = d->filter_order;
  ax.f_decay = -std::log(d->filter_tol) / std::pow(d->filter_outer_k - d->filter_inner_k, d->filter_order);
  ax.dealias = d->dealias ? 1 : 0;
  {
    // FF getaliasedwavenumbers: 1-based iL = floor((1-a)/2*n)+1, iR = ceil((1+a)/2*n); r2c axis: iL..nkr
    double a = d->aliased_fraction;
    auto lo = [&](int64_t n) { return (int64_t)std::floor((1.0 - a) / 2.0 * (double)n); };       // 0-based iL-1
    auto hi = [&](int64_t n) { return (int64_t)std::ceil((1.0 + a) / 2.0 * (double)n); };        // exclusive
    ax.ax_lo = lo(g.nx);
    ax.ay_lo = lo(g.ny);
    ax.ay_hi = hi(g.ny);
    ax.az_lo = lo(g.nz);
    ax.az_hi = hi(g.nz);
    if (a <= 0.0) {  // FF dealias!: `grid.aliased_fraction == 0 && return nothing` — nothing is zeroed (empty ranges)
      ax.ax_lo = g.nkr;
      ax.ay_lo = ax.ay_hi = 0;
      ax.az_lo = ax.az_hi = 0;
    }
  }
}

void run_velocity_providers(ptf_handle* h) {
  Context& c = h->ctx;
  int nd = c.g.ndim;
  if (c.d.flow_kind == PTF_FLOW_CALLBACK) {
    if (!h->vel_fn) throw Error(PTF_EINVAL, "PTF_FLOW_CALLBACK problem stepped without ptf_set_velocity_callback");
    int64_t count = c.g.lpts() * (c.d.velocity_per_batch ? c.g.B : 1);
    for (int a = 0; a < nd; ++a)
      if (!h->pinned_vel[a]) PTF_CUDA(cudaMallocHost((void**)&h->pinned_vel[a], count * sizeof(double)));
    // evaluated at clock.t — the time at the START of the step, for all stages (TAD.jl:701,718,737)
    h->vel_fn(h->vel_user, c.t, h->pinned_vel[0], nd >= 2 ? h->pinned_vel[1] : nullptr,
              nd >= 3 ? h->pinned_vel[2] : nullptr);
    for (int a = 0; a < nd; ++a) h->engine->set_velocity(a, h->pinned_vel[a], count);
  } else if (c.d.flow_kind == PTF_FLOW_EXPR) {
    h->engine->set_flow_time(c.t);   // evaluated at clock.t for all stages of the step (TAD.jl:701,718,737)
  } else if (c.d.flow_kind == PTF_FLOW_SEPARABLE && h->coeff_fn) {
    for (int a = 0; a < nd; ++a) {
      int nt = h->sep_nterms[a];
      if (nt <= 0) continue;
      double coef[16];
      h->coeff_fn(h->coeff_user, c.t, a, nt, coef);
      h->engine->set_velocity_coeffs(a, nt, coef);
    }
  }
}

void do_steps(ptf_handle* h, int64_t nsteps) {
  Context& c = h->ctx;
  // no per-step host input (steady arrays, or a separable flow without a coefficient callback): the engine may run
  // the whole call in one launch
  const bool per_step_input = c.d.flow_kind == PTF_FLOW_CALLBACK || c.d.flow_kind == PTF_FLOW_EXPR ||
                              (c.d.flow_kind == PTF_FLOW_SEPARABLE && h->coeff_fn != nullptr);
  if (!per_step_input && nsteps > 0 && h->engine->step_many(c.step, nsteps)) {
    for (int64_t i = 0; i < nsteps; ++i) c.t += c.dt;  // same rounding as FF's clock.t += dt per step
    c.step += nsteps;
    return;
  }
  for (int64_t i = 0; i < nsteps; ++i) {
    run_velocity_providers(h);
    h->engine->step_once(c.step);
    c.t += c.dt;  // FF stepforward!: clock.t += dt; clock.step += 1 after the update
    c.step += 1;
  }
}

}  // namespace

// ---- hooks used by the MultiLayerQG coupling (engine_mqg.cu) ----
namespace ptf {
void couple_layered_velocity(ptf_handle* h, const double* u_dev, const double* v_dev, const double* U_host,
                             int64_t count) {
  PTF_CUDA(cudaStreamSynchronize(h->ctx.stream));
  h->engine->set_velocity_external(0, u_dev, count);
  h->engine->set_velocity_external(1, v_dev, count);
  h->engine->set_layered_shift(U_host);   // MQGprob.params.U, added to u in the product (TAD.jl:795)
}
void decouple_layered_velocity(ptf_handle* h) {
  cudaSetDevice(h->ctx.device);
  cudaStreamSynchronize(h->ctx.stream);
  try {
    h->engine->set_velocity_external(0, nullptr, 0);
    h->engine->set_velocity_external(1, nullptr, 0);
  } catch (...) {
  }
}
cudaStream_t tracer_stream(ptf_handle* h) { return h->ctx.stream; }
int tracer_device(ptf_handle* h) { return h->ctx.device; }
void tracer_geometry(ptf_handle* h, int64_t* nx, int64_t* ny, int64_t* nbatch, double* Lx, double* Ly, double* dt,
                     int* flow_kind, int* ndim) {
  const Geometry& g = h->ctx.g;
  *nx = g.nx;
  *ny = g.ny;
  *nbatch = g.B;
  *Lx = g.Lx;
  *Ly = g.Ly;
  *dt = h->ctx.dt;
  *flow_kind = h->ctx.d.flow_kind;
  *ndim = g.ndim;
}
void tracer_step_one(ptf_handle* h) { do_steps(h, 1); }
void tracer_set_mqg(ptf_handle* h, ptf_mqg_handle* m) { h->mqg = m; }
}  // namespace ptf

extern "C" {

int32_t ptf_version(int32_t* major, int32_t* minor) {
  if (major) *major = PTF_VERSION_MAJOR;
  if (minor) *minor = PTF_VERSION_MINOR;
  return PTF_OK;
}

int32_t ptf_device_count(int32_t* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  if (count) *count = n;
  return PTF_OK;
}

const char* ptf_error_string(int32_t status) { return status_name(status); }

const char* ptf_last_error(const ptf_handle* h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

int32_t ptf_nccl_unique_id(uint8_t id[128]) {
#ifdef PTF_WITH_NCCL
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId u;
  if (ncclGetUniqueId(&u) != ncclSuccess) {
    g_create_error = "ncclGetUniqueId failed";
    return PTF_ENCCL;
  }
  std::memcpy(id, &u, 128);
  return PTF_OK;
#else
  (void)id;
  g_create_error = "built without NCCL";
  return PTF_EUNSUPPORTED;
#endif
}

int32_t ptf_desc_init(ptf_desc* d) {
  if (!d) return PTF_EINVAL;
  std::memset(d, 0, sizeof(*d));
  d->struct_size = sizeof(ptf_desc);
  d->ndim = 2;                      // reference defaults: TAD.jl:143-203
  for (int a = 0; a < 3; ++a) {
    d->n[a] = 128;
    d->L[a] = 2.0 * M_PI;
    d->kappa[a] = 0.1;
  }
  d->nbatch = 1;
  d->stepper = PTF_STEPPER_RK4;
  d->kappa_h = 0.0;
  d->n_kappa_h = 0;
  d->dealias = 0;
  d->aliased_fraction = 1.0 / 3.0;
  d->dt = 0.01;
  d->nyquist_sign = -1;
  d->flow_kind = PTF_FLOW_STEADY;
  d->velocity_per_batch = 0;
  d->engine = PTF_ENGINE_AUTO;
  d->device = -1;
  d->decomposition = PTF_DECOMP_NONE;
  d->nranks = 1;
  d->rank = 0;
  d->filter_order = 4.0;
  d->filter_inner_k = 2.0 / 3.0;
  d->filter_outer_k = 1.0;
  d->filter_tol = 1e-15;
  d->use_graph = 1;
  return PTF_OK;
}

int32_t ptf_create(const ptf_desc* d, ptf_handle** out) {
  if (!out) return PTF_EINVAL;
  *out = nullptr;
  ptf_handle* h = nullptr;
  int32_t rc = guarded(nullptr, [&]() {
    h = new ptf_handle();
    build_context(h, d);
    std::string why;
    int want = d->engine;
    const bool one_d = h->ctx.g.ndim == 1;
    const bool expr_flow = d->flow_kind == PTF_FLOW_EXPR;
    if (expr_flow && one_d && want == PTF_ENGINE_FUSED)
      throw Error(PTF_EUNSUPPORTED, "fused 1-D engine: expression flows (PTF_FLOW_EXPR) run on the cuFFT pipeline");
    if (expr_flow && one_d && want == PTF_ENGINE_AUTO) want = PTF_ENGINE_CUFFT;
    if (h->ctx.g.slab2d) {
      if (want == PTF_ENGINE_FUSED) throw Error(PTF_EUNSUPPORTED, "fused engine: 2-D slab decomposition runs on the cuFFT pipeline");
      h->engine = make_slab2d_engine(h->ctx);
    } else if (want == PTF_ENGINE_FUSED) {
      if (one_d) {
        if (!fused1d_engine_supports(h->ctx, &why)) throw Error(PTF_EUNSUPPORTED, "fused engine: " + why);
        h->engine = make_fused1d_engine(h->ctx);
      } else if (h->ctx.g.ndim == 3) {
        if (!fused3d_engine_supports(h->ctx, &why)) throw Error(PTF_EUNSUPPORTED, "fused engine: " + why);
        h->engine = make_fused3d_engine(h->ctx);
      } else {
        if (!fused_engine_supports(h->ctx, &why)) throw Error(PTF_EUNSUPPORTED, "fused engine: " + why);
        h->engine = make_fused_engine(h->ctx);
      }
    } else if (want == PTF_ENGINE_AUTO && one_d && fused1d_engine_supports(h->ctx, &why)) {
      h->engine = make_fused1d_engine(h->ctx);
    } else if (want == PTF_ENGINE_AUTO && !h->ctx.g.slab && fused_engine_supports(h->ctx, &why)) {
      h->engine = make_fused_engine(h->ctx);
    } else if (want == PTF_ENGINE_AUTO && fused3d_engine_supports(h->ctx, &why)) {
      h->engine = make_fused3d_engine(h->ctx);
    } else {
      h->engine = make_cufft_engine(h->ctx);
    }
    PTF_CUDA(cudaEventCreate(&h->ev0));
    PTF_CUDA(cudaEventCreate(&h->ev1));
    PTF_CUDA(cudaStreamSynchronize(h->ctx.stream));
  });
  if (rc != PTF_OK) {
    if (h) ptf_destroy(h);
    return rc;
  }
  *out = h;
  return PTF_OK;
}

int32_t ptf_destroy(ptf_handle* h) {
  if (!h) return PTF_OK;
  cudaSetDevice(h->ctx.device);
  if (h->ctx.stream) cudaStreamSynchronize(h->ctx.stream);
  if (h->mqg) ptf_mqg_forget_tracer(h->mqg, h);   // the coupled flow keeps running on this (soon destroyed) stream: see
                                                  // ptf_mqg_forget_tracer, which moves it back to its own stream
  h->engine.reset();
  for (auto& p : h->pinned_vel)
    if (p) cudaFreeHost(p);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ctx.stream) cudaStreamDestroy(h->ctx.stream);
#ifdef PTF_WITH_NCCL
  if (h->ctx.nccl_comm) release_comm((ncclComm_t)h->ctx.nccl_comm);
#endif
  delete h;
  return PTF_OK;
}

int32_t ptf_local_shape(const ptf_handle* h, int64_t phys_n[4], int64_t spec_n[4], int64_t phys_offset[4],
                        int64_t spec_offset[4]) {
  if (!h) return PTF_EINVAL;
  const Geometry& g = h->ctx.g;
  if (phys_n) { phys_n[0] = g.nx; phys_n[1] = g.ny; phys_n[2] = g.nzl; phys_n[3] = g.B; }
  if (spec_n) { spec_n[0] = g.nkr; spec_n[1] = g.nyl; spec_n[2] = g.nz; spec_n[3] = g.B; }
  if (phys_offset) { phys_offset[0] = phys_offset[1] = 0; phys_offset[2] = g.zoff; phys_offset[3] = g.Boffset; }
  if (spec_offset) { spec_offset[0] = 0; spec_offset[1] = g.yoff; spec_offset[2] = 0; spec_offset[3] = g.Boffset; }
  if (g.slab2d) {  // physical rows [ypoff, ypoff+nyp); spectral columns kr in [koff, koff+kvalid), all ky
    if (phys_n) phys_n[1] = g.nyp;
    if (phys_offset) phys_offset[1] = g.ypoff;
    if (spec_n) { spec_n[0] = g.kvalid; spec_n[1] = g.ny; }
    if (spec_offset) { spec_offset[0] = g.koff; spec_offset[1] = 0; }
  }
  return PTF_OK;
}

int32_t ptf_set_velocity(ptf_handle* h, int32_t comp, const double* host, int64_t count) {
  if (!h || !host) return PTF_EINVAL;
  return guarded(h, [&]() { h->engine->set_velocity(comp, host, count); });
}

int32_t ptf_set_velocity_callback(ptf_handle* h, ptf_velocity_fn fn, void* user) {
  if (!h) return PTF_EINVAL;
  h->vel_fn = fn;
  h->vel_user = user;
  return PTF_OK;
}

int32_t ptf_set_velocity_separable(ptf_handle* h, int32_t comp, int32_t nterms, const double* xtab,
                                   const double* ytab, const double* ztab, const double* coeff0) {
  if (!h) return PTF_EINVAL;
  return guarded(h, [&]() {
    PTF_REQUIRE(h->ctx.d.flow_kind == PTF_FLOW_SEPARABLE, "problem was not created with PTF_FLOW_SEPARABLE");
    PTF_REQUIRE(comp >= 0 && comp < h->ctx.g.ndim, "velocity component out of range");
    h->engine->set_velocity_separable(comp, nterms, xtab, ytab, ztab, coeff0);
    h->sep_nterms[comp] = nterms;
  });
}

int32_t ptf_set_velocity_expr(ptf_handle* h, int32_t comp, const char* expr) {
  if (!h || !expr) return PTF_EINVAL;
  return guarded(h, [&]() {
    PTF_REQUIRE(h->ctx.d.flow_kind == PTF_FLOW_EXPR, "problem was not created with PTF_FLOW_EXPR");
    PTF_REQUIRE(comp >= 0 && comp < h->ctx.g.ndim, "velocity component out of range");
    h->engine->set_velocity_expr(comp, expr);
  });
}

int32_t ptf_set_coeff_callback(ptf_handle* h, ptf_coeff_fn fn, void* user) {
  if (!h) return PTF_EINVAL;
  h->coeff_fn = fn;
  h->coeff_user = user;
  return PTF_OK;
}

int32_t ptf_set_layered_velocity(ptf_handle* h, const double* u, const double* v, const double* U) {
  if (!h || !u || !v) return PTF_EINVAL;
  return guarded(h, [&]() {
    const Geometry& g = h->ctx.g;
    PTF_REQUIRE(g.ndim == 2, "layered velocities are 2-D per layer");
    int64_t count = g.lpts() * g.B;
    h->engine->set_velocity(0, u, count);
    h->engine->set_velocity(1, v, count);
    h->engine->set_layered_shift(U);
  });
}

int32_t ptf_set_c(ptf_handle* h, const double* c_host, int32_t replicate_over_batch) {
  if (!h || !c_host) return PTF_EINVAL;
  return guarded(h, [&]() { h->engine->set_c(c_host, replicate_over_batch != 0); });
}

int32_t ptf_get_c(ptf_handle* h, double* c_host) {
  if (!h || !c_host) return PTF_EINVAL;
  return guarded(h, [&]() { h->engine->get_c(c_host); });
}

int32_t ptf_set_sol(ptf_handle* h, const double* s) {
  if (!h || !s) return PTF_EINVAL;
  return guarded(h, [&]() { h->engine->set_sol(s); });
}

int32_t ptf_get_sol(ptf_handle* h, double* s) {
  if (!h || !s) return PTF_EINVAL;
  return guarded(h, [&]() { h->engine->get_sol(s); });
}

int32_t ptf_get_clock(const ptf_handle* h, double* t, int64_t* step, double* dt) {
  if (!h) return PTF_EINVAL;
  if (t) *t = h->ctx.t;
  if (step) *step = h->ctx.step;
  if (dt) *dt = h->ctx.dt;
  return PTF_OK;
}

int32_t ptf_set_clock(ptf_handle* h, double t, int64_t step) {
  if (!h) return PTF_EINVAL;
  h->ctx.t = t;
  h->ctx.step = step;
  return PTF_OK;
}

int32_t ptf_set_dt(ptf_handle* h, double dt) {
  if (!h) return PTF_EINVAL;
  return guarded(h, [&]() {
    PTF_REQUIRE(std::isfinite(dt) && dt > 0, "dt must be positive");
    if (dt != h->ctx.dt) {
      h->ctx.dt = dt;
      h->engine->on_dt_changed();
    }
  });
}

int32_t ptf_step(ptf_handle* h, int64_t nsteps) {
  if (!h || nsteps < 0) return PTF_EINVAL;
  return guarded(h, [&]() {
    do_steps(h, nsteps);
    PTF_CUDA(cudaStreamSynchronize(h->ctx.stream));  // stepforward! is synchronous in the reference
  });
}

int32_t ptf_step_until(ptf_handle* h, double t_stop) {
  if (!h) return PTF_EINVAL;
  return guarded(h, [&]() {
    Context& c = h->ctx;
    // FF step_until!: only for steppers whose coefficients do not depend on dt
    if (c.st.base == PTF_STEPPER_ETDRK4)
      throw Error(PTF_EUNSUPPORTED, "step_until! requires an explicit stepper (not ETDRK4)");
    PTF_REQUIRE(t_stop > c.t, "stop time must be greater than the current time");
    double dt = c.dt;
    double interval = t_stop - c.t;
    int64_t n = (int64_t)std::floor(interval / dt);
    do_steps(h, n);
    double rem = interval - (double)n * dt;
    if (rem > 0) {
      c.dt = rem;
      h->engine->on_dt_changed();
      do_steps(h, 1);
      c.dt = dt;
      h->engine->on_dt_changed();
    }
    c.t = t_stop;
    PTF_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int32_t ptf_step_timed(ptf_handle* h, int64_t nsteps, float* device_ms) {
  if (!h || nsteps < 0) return PTF_EINVAL;
  return guarded(h, [&]() {
    PTF_CUDA(cudaEventRecord(h->ev0, h->ctx.stream));
    do_steps(h, nsteps);
    PTF_CUDA(cudaEventRecord(h->ev1, h->ctx.stream));
    PTF_CUDA(cudaEventSynchronize(h->ev1));
    float ms = 0;
    PTF_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (device_ms) *device_ms = ms;
  });
}

int32_t ptf_sync(ptf_handle* h) {
  if (!h) return PTF_EINVAL;
  return guarded(h, [&]() { PTF_CUDA(cudaStreamSynchronize(h->ctx.stream)); });
}

int32_t ptf_engine(const ptf_handle* h, int32_t* engine) {
  if (!h || !engine) return PTF_EINVAL;
  *engine = h->engine->id();
  return PTF_OK;
}

int32_t ptf_launch_count(const ptf_handle* h, int64_t* own_kernels, int64_t* library_calls) {
  if (!h) return PTF_EINVAL;
  if (own_kernels) *own_kernels = h->engine->own_launches;
  if (library_calls) *library_calls = h->engine->lib_calls;
  return PTF_OK;
}

int32_t ptf_kernel_timed(ptf_handle* h, const char* name, int32_t reps, float* avg_ms) {
  if (!h || reps <= 0) return PTF_EINVAL;
  return guarded(h, [&]() {
    float ms = h->engine->time_kernel(name, reps);
    if (avg_ms) *avg_ms = ms;
  });
}

int32_t ptf_device_bytes(const ptf_handle* h, int64_t* bytes) {
  if (!h || !bytes) return PTF_EINVAL;
  *bytes = h->engine->dev_bytes + h->ctx.table_bytes;
  return PTF_OK;
}

int32_t ptf_diag(ptf_handle* h, double* mean_c, double* variance_c, double* max_abs_sol) {
  if (!h) return PTF_EINVAL;
  return guarded(h, [&]() { h->engine->diag(mean_c, variance_c, max_abs_sol); });
}

int32_t ptf_selftest_fft(int32_t n, int32_t dir, int32_t count, const double* in_host, double* out_host) {
  if (!in_host || !out_host) return PTF_EINVAL;
  return guarded(nullptr, [&]() { ptf::selftest_fft(n, dir, count, in_host, out_host); });
}

}  // extern "C"
