// Slab-decomposed 2-D engine (SURVEY §8e "2-D grids: y-rows sharded, one transpose per transform"): one large 2-D grid
// spread over the P GPUs of a box, one process per GPU.
//
//   physical space:  rank r owns rows y in [r*nyp, (r+1)*nyp), nyp = ny/P, layout [nyp][nx]          (x contiguous)
//   spectral space:  rank r owns columns kr in [r*kc, (r+1)*kc), kc = ceil((nx/2+1)/P), layout [kc][ny] (y contiguous);
//                    nkr = nx/2+1 is odd, so the last rank's range is padded with zero columns
//
//   forward  = batched 1-D r2c along x (local rows) -> row-segment pack -> ALL-TO-ALL -> transpose -> batched 1-D c2c along y
//   inverse  = batched c2c along y -> transpose (send blocks contiguous) -> ALL-TO-ALL -> row assembly -> batched c2r along x
//              (x last, which keeps the c2r semantics of the reference's irfft: SURVEY fact 8)
// All spectral pointwise work (i*kr, i*l, L, filter, dealias, stage combine: the shared combine_at) happens in the
// transposed [kc][ny] layout; `sol` never leaves it except in ptf_get_sol.  The all-to-all (grouped ncclSend/ncclRecv,
// own block by device copy) is the only collective; the second gradient field's y-transform and pack overlap the first
// field's exchange, which runs on a priority stream.  Per RK4 step: 12 exchanges of nkr*ny*16/P^2 bytes per peer.
//
// Replaces, for grids larger than one GPU should hold, the same reference functions as the single-GPU engines:
// calcN! (TAD.jl:708-723, 756-769) + the FourierFlows stage combines + set_c!/updatevars! (TAD.jl:815-872).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#ifdef PTF_WITH_NCCL
#include <nccl.h>
#endif

#include "ptf_internal.h"
#include "ptf_pointwise.cuh"
#include "ptf_velocity.cuh"
#include "expr_flow.h"

namespace ptf {
namespace {

// in [R][C] -> out [C][R], 32x32 tiles through padded shared memory
__global__ void __launch_bounds__(256) k2_transpose(const double2* __restrict__ in, double2* __restrict__ out, int64_t R,
                                                    int64_t C) {
  __shared__ double2 tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = in[r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[c * R + r] = tile[threadIdx.x][i];
  }
}

// rows [nyp][nkr] -> send blocks [dest][nyp][kc]   (zero beyond nkr in the last block)
// only rows [j0, j0+nc) are packed (the forward transform is pipelined in row chunks)
__global__ void __launch_bounds__(256) k2_pack_rows(const double2* __restrict__ rows, double2* __restrict__ blocks,
                                                    int64_t nyp, int64_t nkr, int64_t kc, int P, int64_t j0, int64_t nc) {
  const int64_t n = (int64_t)P * nc * kc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i % kc, jl = j0 + (i / kc) % nc, d = i / (kc * nc);
    const int64_t k = d * kc + c;
    blocks[(d * nyp + jl) * kc + c] = k < nkr ? rows[jl * nkr + k] : make_double2(0.0, 0.0);
  }
}

// received blocks [src][nyp][kc] -> rows [nyp][nkr]
__global__ void __launch_bounds__(256) k2_unpack_rows(const double2* __restrict__ blocks, double2* __restrict__ rows,
                                                      int64_t nyp, int64_t nkr, int64_t kc, int P) {
  const int64_t n = nyp * nkr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i % nkr, jl = i / nkr;
    const int64_t s = k / kc, c = k - s * kc;
    rows[i] = blocks[(s * nyp + jl) * kc + c];
  }
}

// [kc][ny] layout: F0 = i*kr*s/N, F1 = i*l*s/N; opt-in dealias!(s) in place
__global__ void __launch_bounds__(256) k2_deriv(double2* __restrict__ ss, double2* __restrict__ F0,
                                                double2* __restrict__ F1, AxisTables ax, int64_t kc, int64_t ny,
                                                int64_t koff, double scale) {
  const int64_t n = kc * ny;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / ny, j = i - c * ny;
    const double kx = ax.kx[c], ky = ax.ky[j];
    double2 s = ss[i];
    if (dealiased_out(ax, koff + c, j, 0)) {
      s = make_double2(0.0, 0.0);
      ss[i] = s;
    }
    F0[i] = make_double2(-kx * s.y * scale, kx * s.x * scale);
    F1[i] = make_double2(-ky * s.y * scale, ky * s.x * scale);
  }
}

__global__ void __launch_bounds__(256) k2_scale_copy(const double2* __restrict__ in, double2* __restrict__ out, int64_t n,
                                                     double scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = make_double2(in[i].x * scale, in[i].y * scale);
}

// p = -u*gx - v*gy over this rank's rows (TAD.jl:764); separable tables are global in y: offset by ypoff
__global__ void __launch_bounds__(256) k2_product(double* __restrict__ g0, const double* __restrict__ g1, VelArgs va,
                                                  int64_t nx, int64_t nyp, int64_t ny, int64_t ypoff) {
  const int64_t half = nx * nyp / 2;
  double2* G0 = reinterpret_cast<double2*>(g0);
  const double2* G1 = reinterpret_cast<const double2*>(g1);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < half; e += (int64_t)gridDim.x * blockDim.x) {
    double2 u, v;
    if (va.kind == PTF_FLOW_SEPARABLE) {
      const int64_t p = 2 * e, i = p % nx, j = ypoff + p / nx;
      u = make_double2(sep_eval(va.sep[0], i, j, 0, nx, ny, 1, 2), sep_eval(va.sep[0], i + 1, j, 0, nx, ny, 1, 2));
      v = make_double2(sep_eval(va.sep[1], i, j, 0, nx, ny, 1, 2), sep_eval(va.sep[1], i + 1, j, 0, nx, ny, 1, 2));
    } else {
      u = reinterpret_cast<const double2*>(va.arr[0])[e];
      v = reinterpret_cast<const double2*>(va.arr[1])[e];
    }
    const double2 a = G0[e], b = G1[e];
    G0[e] = make_double2(-u.x * a.x - v.x * b.x, -u.y * a.y - v.y * b.y);
  }
}

__global__ void __launch_bounds__(256) k2_combine(const double2* __restrict__ Nh, CombinePtrs P, CombineArgs A,
                                                  AxisTables ax, int64_t kvalid, int64_t ny) {
  const int64_t n = kvalid * ny;   // padded columns are never touched: they stay zero
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / ny, j = i - c * ny;
    combine_at(P, A, ax, (size_t)i, (size_t)i, ax.kx[c], ax.ky[j], 0.0, Nh[i]);
  }
}

// sum of w*|s|^2 (Parseval weights: 1 for kr = 0 and kr = nx/2, else 2) and max |s|^2 over the valid local columns
__global__ void __launch_bounds__(256) k2_diag(const double2* __restrict__ s, int64_t kvalid, int64_t ny, int64_t koff,
                                               int64_t nx, double* out) {
  __shared__ double ssum[256];
  __shared__ double smax[256];
  const int64_t n = kvalid * ny;
  double acc = 0, mx = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = koff + i / ny;
    const double2 v = s[i];
    const double a2 = v.x * v.x + v.y * v.y;
    acc += ((k == 0 || k == nx / 2) ? 1.0 : 2.0) * a2;
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

class Slab2DEngine final : public Engine {
 public:
  explicit Slab2DEngine(Context& c) : ctx(c), g(c.g) {
    P = g.P;
    rank = g.rank;
    nx = g.nx;
    ny = g.ny;
    nkr = g.nkr;
    nyp = g.nyp;
    kc = g.kc;
    koff = g.koff;
    kvalid = g.kvalid;
    nspec = kc * ny;
    nreal = nx * nyp;
    // per-rank tables: kx of the local (padded) columns
    std::vector<double> kxl((size_t)kc, 0.0);
    for (int64_t i = 0; i < kvalid; ++i) kxl[i] = g.kx[koff + i];
    d_kxl.alloc(kc, &dev_bytes);
    PTF_CUDA(cudaMemcpy(d_kxl.p, kxl.data(), kc * sizeof(double), cudaMemcpyHostToDevice));
    axl = ctx.ax;
    axl.kx = d_kxl.p;
    const int base = ctx.st.base;
    auto zalloc = [&](DevBuf<double2>& b, int64_t n) {
      b.alloc(n, &dev_bytes);
      PTF_CUDA(cudaMemsetAsync(b.p, 0, b.bytes(), ctx.stream));
    };
    zalloc(sol, nspec);
    if (base == PTF_STEPPER_RK4 || base == PTF_STEPPER_ETDRK4) zalloc(s1, nspec);
    if (base == PTF_STEPPER_ETDRK4) zalloc(s2, nspec);
    if (base != PTF_STEPPER_FORWARD_EULER) zalloc(acc, nspec);
    if (base == PTF_STEPPER_ETDRK4 || base == PTF_STEPPER_AB3) zalloc(n1, nspec);
    if (base == PTF_STEPPER_ETDRK4)
      for (auto* b : {&cE, &cE2, &cZ, &cA, &cB, &cG}) b->alloc(nspec, &dev_bytes);
    for (int f = 0; f < 2; ++f) {
      zalloc(F[f], nspec);
      zalloc(SB[f], nspec);
      zalloc(RB[f], nspec);
      zalloc(R[f], nyp * nkr);
      G[f].alloc(nreal, &dev_bytes);
    }
    vs.init(&g, ctx.stream, ctx.d.flow_kind, &dev_bytes);
    make_plans();
    int lo = 0, hi = 0;
    PTF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PTF_CUDA(cudaStreamCreateWithPriority(&s_comm, cudaStreamNonBlocking, hi));
    for (auto& e : ev) PTF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    on_dt_changed();
  }

  ~Slab2DEngine() override {
    drop_graphs();
    for (cufftHandle p : {plan_y, plan_xi, plan_xf, plan_yT, plan_Ty, plan_xf_chunk})
      if (p) cufftDestroy(p);
    if (s_comm) cudaStreamDestroy(s_comm);
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }

  const char* name() const override { return "slab2d"; }
  int id() const override { return PTF_ENGINE_CUFFT; }
  cudaStream_t stream() const override { return ctx.stream; }

  void make_plans() {
    long long n_y[1] = {ny}, n_x[1] = {nx};
    size_t w[3] = {0, 0, 0};
    PTF_CUFFT(cufftCreate(&plan_y));
    PTF_CUFFT(cufftCreate(&plan_xi));
    PTF_CUFFT(cufftCreate(&plan_xf));
    for (cufftHandle p : {plan_y, plan_xi, plan_xf}) PTF_CUFFT(cufftSetAutoAllocation(p, 0));
    PTF_CUFFT(cufftMakePlanMany64(plan_y, 1, n_y, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2Z, kc, &w[0]));
    // Optional (PTF_SLAB2D_STRIDED=1): y-transforms that change layout on the fly, [kc][ny] -> [ny][kc] and back, instead
    // of a separate tiled transpose.  Measured at 16384^2 on one B200: 94.1 ms/step vs 90.5 ms/step with the explicit
    // transposes (cuFFT's strided stores cost more than the extra coalesced pass), so it is off by default.
    strided = std::getenv("PTF_SLAB2D_STRIDED") && std::atoi(std::getenv("PTF_SLAB2D_STRIDED")) != 0;
    if (strided) {
      size_t wa = 0, wb = 0;
      PTF_CUFFT(cufftCreate(&plan_yT));
      PTF_CUFFT(cufftCreate(&plan_Ty));
      PTF_CUFFT(cufftSetAutoAllocation(plan_yT, 0));
      PTF_CUFFT(cufftSetAutoAllocation(plan_Ty, 0));
      PTF_CUFFT(cufftMakePlanMany64(plan_yT, 1, n_y, n_y, 1, ny, n_y, kc, 1, CUFFT_Z2Z, kc, &wa));
      PTF_CUFFT(cufftMakePlanMany64(plan_Ty, 1, n_y, n_y, kc, 1, n_y, 1, ny, CUFFT_Z2Z, kc, &wb));
      w[0] = std::max(w[0], std::max(wa, wb));
    }
    PTF_CUFFT(cufftMakePlanMany64(plan_xi, 1, n_x, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, nyp, &w[1]));
    PTF_CUFFT(cufftMakePlanMany64(plan_xf, 1, n_x, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, nyp, &w[2]));
    {  // forward pipeline depth (PTF_SLAB2D_CHUNKS).  Measured at 16384^2 on 2 GPUs: 49.9 ms/step unchunked, 50.9 with
       // 4 chunks, 51.0 with 8 — at P = 2 the exchange is already hidden well enough and the smaller batches cost more
       // than they hide, so the default stays 1 until an 8-GPU measurement says otherwise (parity-tested either way).
      const char* e = std::getenv("PTF_SLAB2D_CHUNKS");
      n_chunks = e ? std::atoi(e) : 1;
      if (n_chunks < 1 || nyp % n_chunks != 0 || nyp / n_chunks < 1) n_chunks = 1;
      if (n_chunks > 1) {
        size_t wc = 0;
        PTF_CUFFT(cufftCreate(&plan_xf_chunk));
        PTF_CUFFT(cufftSetAutoAllocation(plan_xf_chunk, 0));
        PTF_CUFFT(cufftMakePlanMany64(plan_xf_chunk, 1, n_x, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, nyp / n_chunks, &wc));
        w[2] = std::max(w[2], wc);
      }
    }
    size_t wm = std::max(w[0], std::max(w[1], w[2]));
    work.alloc(wm ? wm : 16, &dev_bytes);
    for (cufftHandle p : {plan_y, plan_xi, plan_xf, plan_yT, plan_Ty, plan_xf_chunk}) {
      if (!p) continue;
      PTF_CUFFT(cufftSetWorkArea(p, work.p));
      PTF_CUFFT(cufftSetStream(p, ctx.stream));
    }
  }

  int blocks(int64_t n) const { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 16)); }
  static cufftDoubleComplex* Z(double2* p) { return reinterpret_cast<cufftDoubleComplex*>(p); }

  void transpose(const double2* in, double2* out, int64_t Rr, int64_t Cc) {
    dim3 grid((unsigned)((Cc + 31) / 32), (unsigned)((Rr + 31) / 32), 1), block(32, 8, 1);
    k2_transpose<<<grid, block, 0, ctx.stream>>>(in, out, Rr, Cc);
    ++own_launches;
  }

  // blocks of nyp*kc complex values to / from every peer; own block by device copy
  // rows [j0, j0+nc) of every block only (nc < 0: whole blocks)
  void all_to_all(const double2* send, double2* recv, cudaStream_t st, int64_t j0 = 0, int64_t nc = -1) {
    if (nc < 0) nc = nyp;
    const size_t blk = (size_t)nyp * kc, off = (size_t)j0 * kc, cnt = (size_t)nc * kc;
    PTF_CUDA(cudaMemcpyAsync(recv + (size_t)rank * blk + off, send + (size_t)rank * blk + off, cnt * sizeof(double2),
                             cudaMemcpyDeviceToDevice, st));
    if (P == 1) return;
#ifdef PTF_WITH_NCCL
    ncclComm_t comm = (ncclComm_t)ctx.nccl_comm;
    auto ck = [](ncclResult_t r, const char* what) {
      if (r != ncclSuccess) throw Error(PTF_ENCCL, std::string(what) + ": " + ncclGetErrorString(r));
    };
    ck(ncclGroupStart(), "ncclGroupStart");
    for (int r = 0; r < P; ++r) {
      if (r == rank) continue;
      ck(ncclSend(send + (size_t)r * blk + off, 2 * cnt, ncclDouble, r, comm, st), "ncclSend");
      ck(ncclRecv(recv + (size_t)r * blk + off, 2 * cnt, ncclDouble, r, comm, st), "ncclRecv");
    }
    ck(ncclGroupEnd(), "ncclGroupEnd");
    ++lib_calls;
#else
    throw Error(PTF_EUNSUPPORTED, "built without NCCL");
#endif
  }
  void fork_comm() {   // s_comm waits for everything enqueued on the main stream so far
    PTF_CUDA(cudaEventRecord(ev[evi], ctx.stream));
    PTF_CUDA(cudaStreamWaitEvent(s_comm, ev[evi], 0));
    evi = (evi + 1) % NEV;
  }
  cudaEvent_t mark_comm() {   // event after everything enqueued on s_comm so far
    cudaEvent_t e = ev[evi];
    PTF_CUDA(cudaEventRecord(e, s_comm));
    evi = (evi + 1) % NEV;
    return e;
  }

  // spectral [kc][ny] (this rank's columns) -> physical rows [nyp][nx]; `spec` is destroyed.  The caller folds in 1/N.
  // Split in two so that a second field's y-transform overlaps the first field's exchange.
  cudaEvent_t inv_begin(int f, double2* spec) {
    if (strided) {   // [kc][ny] -> y-transform -> written as [ny][kc]: block d = rows of rank d, contiguous
      PTF_CUFFT(cufftExecZ2Z(plan_yT, Z(spec), Z(SB[f].p), CUFFT_INVERSE));
      ++lib_calls;
    } else {
      PTF_CUFFT(cufftExecZ2Z(plan_y, Z(spec), Z(spec), CUFFT_INVERSE));
      ++lib_calls;
      transpose(spec, SB[f].p, kc, ny);
    }
    fork_comm();
    all_to_all(SB[f].p, RB[f].p, s_comm);
    return mark_comm();
  }
  void inv_end(int f, cudaEvent_t arrived, double* real) {
    PTF_CUDA(cudaStreamWaitEvent(ctx.stream, arrived, 0));
    k2_unpack_rows<<<blocks(nyp * nkr), 256, 0, ctx.stream>>>(RB[f].p, R[f].p, nyp, nkr, kc, P);
    ++own_launches;
    PTF_CUFFT(cufftExecZ2D(plan_xi, Z(R[f].p), real));
    ++lib_calls;
  }
  // physical rows -> spectral [kc][ny], unnormalised
  void fwd(double* real, double2* spec) {
    // row chunks: the exchange of chunk c (priority stream) overlaps the x-transform and pack of chunk c+1
    const int64_t nc = nyp / n_chunks;
    for (int c = 0; c < n_chunks; ++c) {
      const int64_t j0 = c * nc;
      PTF_CUFFT(cufftExecD2Z(n_chunks > 1 ? plan_xf_chunk : plan_xf, real + j0 * nx, Z(R[0].p + j0 * nkr)));
      ++lib_calls;
      k2_pack_rows<<<blocks((int64_t)P * nc * kc), 256, 0, ctx.stream>>>(R[0].p, SB[0].p, nyp, nkr, kc, P, j0, nc);
      ++own_launches;
      fork_comm();
      all_to_all(SB[0].p, RB[0].p, s_comm, j0, nc);
    }
    PTF_CUDA(cudaStreamWaitEvent(ctx.stream, mark_comm(), 0));
    if (strided) {   // [src][nyp][kc] == [ny][kc], read strided, written as [kc][ny]
      PTF_CUFFT(cufftExecZ2Z(plan_Ty, Z(RB[0].p), Z(spec), CUFFT_FORWARD));
    } else {
      transpose(RB[0].p, spec, ny, kc);
      PTF_CUFFT(cufftExecZ2Z(plan_y, Z(spec), Z(spec), CUFFT_FORWARD));
    }
    ++lib_calls;
  }

  // ---------------- velocities ----------------
  void sync_vel() {
    if (vs.dirty) drop_graphs();
    vs.dirty = false;
  }
  void set_velocity(int comp, const double* host, int64_t count) override {
    PTF_REQUIRE(count == nreal, "velocity count must be this rank's nx*ny/P points");
    vs.set_array(comp, host, count);
    sync_vel();
  }
  void set_velocity_separable(int comp, int nterms, const double* xt, const double* yt, const double* zt,
                              const double* coeff0) override {
    vs.set_separable(comp, nterms, xt, yt, zt, coeff0);
    sync_vel();
  }
  void set_velocity_coeffs(int comp, int nterms, const double* a) override { vs.set_coeffs(comp, nterms, a); }
  void set_velocity_expr(int comp, const char* expr) override {
    ef.set(comp, expr);
    if (!ef.expr[0].empty() && !ef.expr[1].empty() && ef.stale) {
      PTF_CUDA(cudaStreamSynchronize(ctx.stream));
      drop_graphs();   // before the old module (referenced by captured kernel nodes) is unloaded
      ef.compile(2);
    }
  }
  void set_flow_time(double t) override { ef.set_time(t, ctx.stream); }
  void set_layered_shift(const double*) override {
    throw Error(PTF_EUNSUPPORTED, "layered flows are not slab-decomposed (shard the layers: PTF_DECOMP_BATCH)");
  }

  // ---------------- state ----------------
  void set_c(const double* c_host, bool) override {
    PTF_CUDA(cudaMemcpyAsync(G[0].p, c_host, nreal * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    fwd(G[0].p, sol.p);
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  void get_c(double* c_host) override {
    k2_scale_copy<<<blocks(nspec), 256, 0, ctx.stream>>>(sol.p, F[0].p, nspec, 1.0 / ((double)nx * (double)ny));
    ++own_launches;
    inv_end(0, inv_begin(0, F[0].p), G[0].p);
    PTF_CUDA(cudaMemcpyAsync(c_host, G[0].p, nreal * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  // boundary layout of the local spectral block: (kvalid, ny) column-major == [ny][kvalid], kr fastest
  void set_sol(const double* s_host) override {
    PTF_CUDA(cudaMemsetAsync(F[0].p, 0, F[0].bytes(), ctx.stream));
    if (kvalid > 0)
      PTF_CUDA(cudaMemcpy2DAsync(F[0].p, kc * sizeof(double2), s_host, kvalid * sizeof(double2), kvalid * sizeof(double2),
                                 ny, cudaMemcpyHostToDevice, ctx.stream));
    transpose(F[0].p, sol.p, ny, kc);
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  void get_sol(double* s_host) override {
    transpose(sol.p, F[0].p, kc, ny);
    if (kvalid > 0)
      PTF_CUDA(cudaMemcpy2DAsync(s_host, kvalid * sizeof(double2), F[0].p, kc * sizeof(double2), kvalid * sizeof(double2),
                                 ny, cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void on_dt_changed() override {
    drop_graphs();
    if (ctx.st.base == PTF_STEPPER_ETDRK4) {
      // tlayout 1 = [kx][ky]; padded columns get harmless values (kx = 0) and are never read
      k_etd_coeffs<<<blocks(nspec), 256, 0, ctx.stream>>>(cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p, axl, kc, ny, 1, ctx.dt, 1);
      ++own_launches;
      PTF_CUDA(cudaGetLastError());
    }
  }

  // ---------------- one stage: N-hat(ss) into F[0] ----------------
  void calcN(double2* ss) {
    const double scale = 1.0 / ((double)nx * (double)ny);
    k2_deriv<<<blocks(nspec), 256, 0, ctx.stream>>>(ss, F[0].p, F[1].p, axl, kc, ny, koff, scale);
    ++own_launches;
    cudaEvent_t a0 = inv_begin(0, F[0].p);   // exchange 0 runs on s_comm ...
    cudaEvent_t a1 = inv_begin(1, F[1].p);   // ... while field 1 is y-transformed and packed
    inv_end(0, a0, G[0].p);
    inv_end(1, a1, G[1].p);
    if (vs.va.kind == PTF_FLOW_EXPR) {
      ef.launch(ctx.stream, blocks(nreal / 2), 1, G[0].p, G[1].p, G[1].p, nx, nyp, 1, g.ypoff, 0, g);
    } else {
      if (vs.va.kind != PTF_FLOW_SEPARABLE && (!vs.va.arr[0] || !vs.va.arr[1]))
        throw Error(PTF_EINVAL, "velocity fields have not been set (ptf_set_velocity / callback)");
      k2_product<<<blocks(nreal / 2), 256, 0, ctx.stream>>>(G[0].p, G[1].p, vs.va, nx, nyp, ny, g.ypoff);
    }
    ++own_launches;
    fwd(G[0].p, F[0].p);
  }

  void combine(int mode, double la = 0, double lb = 0, int llast = 0) {
    if (kvalid == 0) return;
    CombinePtrs Pp{sol.p, s1.p, s2.p, acc.p, n1.p, cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p};
    CombineArgs A{mode, ctx.st.filtered ? 1 : 0, ctx.dt, la, lb, llast};
    k2_combine<<<blocks(kvalid * ny), 256, 0, ctx.stream>>>(F[0].p, Pp, A, axl, kvalid, ny);
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
    const int variant = (ctx.st.base == PTF_STEPPER_AB3 && step_index < 3) ? 1 : 0;
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
    if (kvalid > 0) {
      k2_diag<<<blocks(kvalid * ny), 256, 0, ctx.stream>>>(sol.p, kvalid, ny, koff, nx, out.p);
      ++own_launches;
    }
    PTF_CUDA(cudaMemcpyAsync(out.p + 2, sol.p, sizeof(double2), cudaMemcpyDeviceToDevice, ctx.stream));  // rank 0: DC mode
#ifdef PTF_WITH_NCCL
    if (P > 1) {
      ncclComm_t comm = (ncclComm_t)ctx.nccl_comm;
      ncclAllReduce(out.p, out.p, 1, ncclDouble, ncclSum, comm, ctx.stream);
      ncclAllReduce(out.p + 1, out.p + 1, 1, ncclDouble, ncclMax, comm, ctx.stream);
      ncclBroadcast(out.p + 2, out.p + 2, 2, ncclDouble, 0, comm, ctx.stream);
    }
#endif
    double h[4];
    PTF_CUDA(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    const double N = (double)nx * (double)ny;
    const double m = h[2] / N, msq = h[0] / (N * N);
    if (mean_c) *mean_c = m;
    if (var_c) *var_c = msq - m * m;
    if (max_abs_sol) *max_abs_sol = std::sqrt(h[1]);
  }

  float time_kernel(const char* kname, int) override {
    throw Error(PTF_EINVAL, std::string("slab2d engine: no in-place kernel timer for '") + (kname ? kname : "") + "'");
  }

 private:
  Context& ctx;
  Geometry& g;
  int P = 1, rank = 0;
  int64_t nx = 0, ny = 0, nkr = 0, nyp = 0, kc = 0, koff = 0, kvalid = 0, nspec = 0, nreal = 0;
  DevBuf<double> d_kxl;
  AxisTables axl;
  DevBuf<double2> sol, s1, s2, acc, n1, F[2], SB[2], RB[2], R[2];
  DevBuf<double> G[2];
  DevBuf<double> cE, cE2, cZ, cA, cB, cG;
  DevBuf<char> work;
  VelocityStore vs;
  ExprFlow ef;
  cufftHandle plan_y = 0, plan_xi = 0, plan_xf = 0, plan_yT = 0, plan_Ty = 0, plan_xf_chunk = 0;
  int n_chunks = 1;
  bool strided = false;
  cudaStream_t s_comm = nullptr;
  static constexpr int NEV = 16;
  cudaEvent_t ev[NEV] = {nullptr};
  int evi = 0;
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int64_t per_step_own = 0, per_step_lib = 0;
};

}  // namespace

std::unique_ptr<Engine> make_slab2d_engine(Context& ctx) { return std::unique_ptr<Engine>(new Slab2DEngine(ctx)); }

}  // namespace ptf
