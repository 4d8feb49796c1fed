// PTF_FLOW_EXPR: time-varying velocities given as CUDA expressions in (x, y, z, t), compiled at run time (NVRTC, straight
// to sm_100a SASS) into the physical-space product kernel of calcN!.
//
// Replaces the closures `u(x, y, z, t)` of ConstDiffTimeVaryingFlowParams (TAD.jl:268-348) that the reference
// broadcasts over `gridpoints(grid)` in every stage (TAD.jl:715-718, 734-737): the expressions are evaluated in registers
// at clock.t (held in a device scalar that is refreshed once per step, so the captured graph never changes), which
// costs zero HBM bytes and no host->device upload — the only practical provider for a 1024^3 grid besides separable
// tables.  Coordinates are formed exactly as FourierFlows / the oracle form them: x_i = -Lx/2 + i*dx (one multiply,
// one add, no FMA contraction).
#include <nvrtc.h>

#include <sstream>

#include "expr_flow.h"

namespace ptf {

namespace {
const char* kTemplate = R"SRC(
#define pi 3.141592653589793238462643383279502884
#ifndef M_PI
#define M_PI pi
#endif
__device__ __forceinline__ double ptf_coord(double c0, long long idx, double d) { return __dadd_rn(c0, __dmul_rn((double)idx, d)); }
__device__ __forceinline__ double ptf_point(double x, double y, double z, double t, double gx, double gy, double gz) {
  double p = -(PTF_EXPR_U) * gx;
#if PTF_ND >= 2
  p = p - (PTF_EXPR_V) * gy;
#endif
#if PTF_ND >= 3
  p = p - (PTF_EXPR_W) * gz;
#endif
  return p;
}
// p = -u*gx - v*gy - w*gz written over gx (TAD.jl:702, 720, 739); one thread = two x-adjacent points; grid.y = member
extern "C" __global__ void __launch_bounds__(256) ptf_product_expr(double* __restrict__ g0, const double* __restrict__ g1,
    const double* __restrict__ g2, const double* __restrict__ tptr, long long nx, long long ny, long long nzl,
    long long joff, long long koff, double x0, double dx, double y0, double dy, double z0, double dz) {
  const double t = *tptr;
  const long long npts = nx * ny * nzl, half = npts >> 1;
  const long long b = blockIdx.y;
  double2* G0 = reinterpret_cast<double2*>(g0 + b * npts);
  const double2* G1 = reinterpret_cast<const double2*>(g1 + b * npts);
  const double2* G2 = reinterpret_cast<const double2*>(g2 + b * npts);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < half; e += (long long)gridDim.x * blockDim.x) {
    const long long p = 2 * e, i = p % nx, j = (p / nx) % ny, k = p / (nx * ny);
    const double y = ptf_coord(y0, joff + j, dy), z = ptf_coord(z0, koff + k, dz);
    const double2 a = G0[e];
    double2 c = make_double2(0.0, 0.0), d = make_double2(0.0, 0.0);
#if PTF_ND >= 2
    c = G1[e];
#endif
#if PTF_ND >= 3
    d = G2[e];
#endif
    G0[e] = make_double2(ptf_point(ptf_coord(x0, i, dx), y, z, t, a.x, c.x, d.x),
                         ptf_point(ptf_coord(x0, i + 1, dx), y, z, t, a.y, c.y, d.y));
  }
}
// The same expressions written out as fields u, v, w at clock.t (fused engines: the velocity is frozen for all stages of
// a step, so it is evaluated once per step and the row kernels read it like steady arrays).
extern "C" __global__ void __launch_bounds__(256) ptf_fill_expr(double* __restrict__ uo, double* __restrict__ vo,
    double* __restrict__ wo, const double* __restrict__ tptr, long long nx, long long ny, long long nzl,
    long long joff, long long koff, double x0, double dx, double y0, double dy, double z0, double dz) {
  const double t = *tptr;
  const long long npts = nx * ny * nzl;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += (long long)gridDim.x * blockDim.x) {
    const long long i = p % nx, j = (p / nx) % ny, k = p / (nx * ny);
    const double x = ptf_coord(x0, i, dx), y = ptf_coord(y0, joff + j, dy), z = ptf_coord(z0, koff + k, dz);
    uo[p] = (PTF_EXPR_U);
#if PTF_ND >= 2
    vo[p] = (PTF_EXPR_V);
#endif
#if PTF_ND >= 3
    wo[p] = (PTF_EXPR_W);
#endif
    (void)x; (void)y; (void)z; (void)t;
  }
}
)SRC";
}  // namespace

ExprFlow::~ExprFlow() {
  if (lib) cudaLibraryUnload(lib);
}

void ExprFlow::set(int comp, const char* e) {
  PTF_REQUIRE(comp >= 0 && comp < 3, "velocity component out of range");
  PTF_REQUIRE(e && *e, "empty velocity expression");
  std::string s(e);
  for (char& ch : s)
    if (ch == '\n' || ch == '\r') ch = ' ';
  PTF_REQUIRE(s.size() < 4096, "velocity expression too long");
  PTF_REQUIRE(s.find('#') == std::string::npos && s.find(';') == std::string::npos && s.find('{') == std::string::npos,
              "velocity expression must be a single C expression in x, y, z, t");
  if (s != expr[comp]) {
    expr[comp] = s;
    stale = true;
  }
}

void ExprFlow::compile(int ndim) {
  for (int a = 0; a < ndim; ++a) PTF_REQUIRE(!expr[a].empty(), "velocity expressions have not been set (ptf_set_velocity_expr)");
  std::ostringstream src;
  src << "#define PTF_ND " << ndim << "\n";
  src << "#define PTF_EXPR_U " << expr[0] << "\n";
  if (ndim >= 2) src << "#define PTF_EXPR_V " << expr[1] << "\n";
  if (ndim >= 3) src << "#define PTF_EXPR_W " << expr[2] << "\n";
  src << kTemplate;
  const std::string code = src.str();
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, code.c_str(), "ptf_expr_flow.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
    throw Error(PTF_ECUDA, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=true", "-lineinfo"};
  nvrtcResult rc = nvrtcCompileProgram(prog, 4, opts);
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nvrtcGetProgramLog(prog, &log[0]);
    nvrtcDestroyProgram(&prog);
    throw Error(PTF_EINVAL, "velocity expression does not compile: " + log);
  }
  size_t n = 0;
  nvrtcGetCUBINSize(prog, &n);
  std::vector<char> cubin(n);
  nvrtcGetCUBIN(prog, cubin.data());
  nvrtcDestroyProgram(&prog);
  if (lib) {
    cudaLibraryUnload(lib);
    lib = nullptr;
  }
  PTF_CUDA(cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  PTF_CUDA(cudaLibraryGetKernel(&kern, lib, "ptf_product_expr"));
  PTF_CUDA(cudaLibraryGetKernel(&kern_fill, lib, "ptf_fill_expr"));
  if (!d_t.p) {
    d_t.alloc(1);
    PTF_CUDA(cudaMemset(d_t.p, 0, sizeof(double)));
  }
  compiled_ndim = ndim;
  stale = false;
}

void ExprFlow::set_time(double t, cudaStream_t st) {
  if (!d_t.p) {
    d_t.alloc(1);
  }
  // pageable-source async copy: staged before the call returns, ordered on the step stream
  PTF_CUDA(cudaMemcpyAsync(d_t.p, &t, sizeof(double), cudaMemcpyHostToDevice, st));
}

void ExprFlow::launch(cudaStream_t st, int blocks, int nbatch, double* g0, const double* g1, const double* g2, int64_t nx,
                      int64_t ny, int64_t nzl, int64_t joff, int64_t koff, const Geometry& g) {
  if (stale || !kern || compiled_ndim != g.ndim) throw Error(PTF_EINVAL, "velocity expressions are not compiled");
  long long a_nx = nx, a_ny = ny, a_nzl = nzl, a_joff = joff, a_koff = koff;
  double x0 = -g.Lx / 2, dx = g.Lx / (double)g.nx;
  double y0 = g.ndim >= 2 ? -g.Ly / 2 : 0.0, dy = g.ndim >= 2 ? g.Ly / (double)g.ny : 0.0;
  double z0 = g.ndim >= 3 ? -g.Lz / 2 : 0.0, dz = g.ndim >= 3 ? g.Lz / (double)g.nz : 0.0;
  const double* tptr = d_t.p;
  void* args[] = {&g0, &g1, &g2, &tptr, &a_nx, &a_ny, &a_nzl, &a_joff, &a_koff, &x0, &dx, &y0, &dy, &z0, &dz};
  PTF_CUDA(cudaLaunchKernel((const void*)kern, dim3((unsigned)blocks, (unsigned)nbatch, 1), dim3(256, 1, 1), args, 0, st));
}

void ExprFlow::fill(cudaStream_t st, double* u, double* v, double* w, int64_t nx, int64_t ny, int64_t nzl, int64_t joff,
                    int64_t koff, const Geometry& g) {
  if (stale || !kern_fill || compiled_ndim != g.ndim) throw Error(PTF_EINVAL, "velocity expressions are not compiled");
  long long a_nx = nx, a_ny = ny, a_nzl = nzl, a_joff = joff, a_koff = koff;
  double x0 = -g.Lx / 2, dx = g.Lx / (double)g.nx;
  double y0 = g.ndim >= 2 ? -g.Ly / 2 : 0.0, dy = g.ndim >= 2 ? g.Ly / (double)g.ny : 0.0;
  double z0 = g.ndim >= 3 ? -g.Lz / 2 : 0.0, dz = g.ndim >= 3 ? g.Lz / (double)g.nz : 0.0;
  const double* tptr = d_t.p;
  void* args[] = {&u, &v, &w, &tptr, &a_nx, &a_ny, &a_nzl, &a_joff, &a_koff, &x0, &dx, &y0, &dy, &z0, &dz};
  PTF_CUDA(cudaLaunchKernel((const void*)kern_fill, dim3(148 * 16, 1, 1), dim3(256, 1, 1), args, 0, st));
}

}  // namespace ptf
