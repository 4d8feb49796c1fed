// Velocity providers shared by the engines (how u, v, w reach the physical-space product of calcN!, TAD.jl:701-799).
#pragma once
#include "ptf_internal.h"

namespace ptf {

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


// Host-side owner of the velocity data of one problem (arrays, separable tables, layered shift).
// `dirty` is raised whenever a kernel ARGUMENT changed (pointer / stride / term count), i.e. captured graphs are stale.
struct VelocityStore {
  Geometry* g = nullptr;
  cudaStream_t stream = nullptr;
  int64_t* tally = nullptr;
  VelArgs va;
  bool dirty = false;
  DevBuf<double> vel[3], sepx[3], sepy[3], sepz[3], sepa, ushift;

  void init(Geometry* geo, cudaStream_t s, int kind, int64_t* bytes) {
    g = geo;
    stream = s;
    tally = bytes;
    va = VelArgs{};
    va.kind = kind;
  }

  void set_array(int comp, const double* host, int64_t count) {
    PTF_REQUIRE(comp >= 0 && comp < g->ndim, "velocity component out of range");
    PTF_REQUIRE(count == g->lpts() || count == g->lpts() * g->B,
                "velocity count must be the local point count (times nbatch for per-member fields)");
    PTF_REQUIRE(comp == 0 || vel[0].n == 0 || vel[0].n == (size_t)count,
                "all velocity components must have the same extent");
    if (vel[comp].n != (size_t)count) {
      vel[comp].alloc(count, tally);
      dirty = true;
    }
    va.arr[comp] = vel[comp].p;
    int64_t ms = (count == g->lpts() && g->B > 1) ? 0 : g->lpts();
    if (va.member_stride != ms) dirty = true;
    va.member_stride = ms;
    PTF_CUDA(cudaMemcpyAsync(vel[comp].p, host, count * sizeof(double), cudaMemcpyHostToDevice, stream));
    PTF_CUDA(cudaStreamSynchronize(stream));
  }

  // Alias a device-resident field owned by someone else (the coupled MultiLayerQG solver's vars.u / vars.v, TAD.jl:795-796).
  // nullptr detaches.  Everything that reads it must be ordered on the same stream as its producer.
  void set_external(int comp, const double* dev, int64_t count) {
    PTF_REQUIRE(comp >= 0 && comp < g->ndim, "velocity component out of range");
    PTF_REQUIRE(dev == nullptr || count == g->lpts() * g->B, "external velocity must hold one field per member");
    vel[comp].release();
    if (va.arr[comp] != dev) dirty = true;
    va.arr[comp] = dev;
    if (va.member_stride != g->lpts()) dirty = true;
    va.member_stride = g->lpts();
  }

  void set_separable(int comp, int nterms, const double* xt, const double* yt, const double* zt,
                     const double* coeff0) {
    const int nd = g->ndim;
    PTF_REQUIRE(comp >= 0 && comp < nd, "velocity component out of range");
    PTF_REQUIRE(nterms >= 0 && nterms <= MAX_TERMS, "separable flow supports at most 8 terms per component");
    SepFlow& f = va.sep[comp];
    f.nterms = nterms;
    auto up = [&](DevBuf<double>& b, const double* h, int64_t n) -> const double* {
      if (!h || nterms == 0) return nullptr;
      b.alloc((size_t)nterms * n, tally);
      PTF_CUDA(cudaMemcpy(b.p, h, (size_t)nterms * n * sizeof(double), cudaMemcpyHostToDevice));
      return b.p;
    };
    f.xt = up(sepx[comp], xt, g->nx);
    PTF_REQUIRE(nterms == 0 || f.xt, "separable flow needs an x table");
    f.yt = nd >= 2 ? up(sepy[comp], yt, g->ny) : nullptr;
    f.zt = nd >= 3 ? up(sepz[comp], zt, g->nz) : nullptr;
    PTF_REQUIRE(nterms == 0 || nd < 2 || f.yt, "separable flow needs a y table");
    PTF_REQUIRE(nterms == 0 || nd < 3 || f.zt, "separable flow needs a z table");
    if (!sepa.p) {
      sepa.alloc(3 * MAX_TERMS, tally);
      PTF_CUDA(cudaMemset(sepa.p, 0, sepa.bytes()));
    }
    f.a = sepa.p + comp * MAX_TERMS;
    double a0[MAX_TERMS];
    for (int m = 0; m < MAX_TERMS; ++m) a0[m] = (m < nterms) ? (coeff0 ? coeff0[m] : 1.0) : 0.0;
    PTF_CUDA(cudaMemcpy(sepa.p + comp * MAX_TERMS, a0, sizeof(a0), cudaMemcpyHostToDevice));
    dirty = true;
  }

  void set_coeffs(int comp, int nterms, const double* a) {
    SepFlow& f = va.sep[comp];
    PTF_REQUIRE(nterms == f.nterms && f.a, "coefficient count does not match the separable flow");
    // pageable-source async copy: staged before the call returns, ordered on the step stream
    PTF_CUDA(cudaMemcpyAsync(sepa.p + comp * MAX_TERMS, a, nterms * sizeof(double), cudaMemcpyHostToDevice, stream));
  }

  void set_shift(const double* U) {
    if (!U) {
      if (va.ushift) dirty = true;
      va.ushift = nullptr;
      return;
    }
    if (ushift.n != (size_t)(g->B * g->ny)) {
      ushift.alloc(g->B * g->ny, tally);
      dirty = true;
    }
    PTF_CUDA(cudaMemcpyAsync(ushift.p, U, ushift.bytes(), cudaMemcpyHostToDevice, stream));
    PTF_CUDA(cudaStreamSynchronize(stream));
    if (va.ushift != ushift.p) dirty = true;
    va.ushift = ushift.p;
  }
};

}  // namespace ptf
