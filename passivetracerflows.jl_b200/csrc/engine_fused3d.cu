// Fused 3-D engine, host side (kernels: fused_kernels.cuh with D3 = true, fused3d_kernels.cuh).  calcN! in 3-D
// (TAD.jl:771-786 steady, :725-742 time-varying) + the FourierFlows stage combine as four hand-written kernels per
// stage, no cuFFT on the step path; one process per GPU with z-slabs (physical) / ky-slabs (spectral) when the
// problem is slab-decomposed, the all-to-all between the y- and z-column kernels moving contiguous blocks that the
// producing kernels already wrote in the receiver's layout (no pack / unpack passes).
#include <cmath>

#include "fused3d_kernels.cuh"
#include "expr_flow.h"

#ifdef PTF_WITH_NCCL
#include <nccl.h>
#endif

namespace ptf {

#define PTF_DECL_INST3(N)                                                                           \
  void fused3_prep_##N();                                                                           \
  void fused3_launch_z_##N(bool has_in, int fam, const void* yargs, cudaStream_t st, int n_sm);     \
  void fused3_launch_x_##N(int vmode, const void* xargs, int nplanes, cudaStream_t st, int n_sm);   \
  void fused3_launch_y_##N(bool inverse, const void* y3args, cudaStream_t st, int n_sm);
PTF_DECL_INST3(64)
PTF_DECL_INST3(128)
PTF_DECL_INST3(256)
PTF_DECL_INST3(512)
PTF_DECL_INST3(1024)

#define PTF_DISPATCH_N3(n, FN, ...)                                                              \
  switch (n) {                                                                                   \
    case 64: FN##_64(__VA_ARGS__); break;                                                        \
    case 128: FN##_128(__VA_ARGS__); break;                                                      \
    case 256: FN##_256(__VA_ARGS__); break;                                                      \
    case 512: FN##_512(__VA_ARGS__); break;                                                      \
    case 1024: FN##_1024(__VA_ARGS__); break;                                                    \
    default: throw Error(PTF_EUNSUPPORTED, "fused 3-D engine: unsupported transform length");    \
  }

namespace {

bool is_fused3_size(int64_t n) { return n == 64 || n == 128 || n == 256 || n == 512 || n == 1024; }
int ilog2(int64_t v) {
  int s = 0;
  while ((int64_t(1) << s) < v) ++s;
  return s;
}

class Fused3DEngine final : public Engine {
 public:
  explicit Fused3DEngine(Context& c) : ctx(c), g(c.g) {
    nx = (int)g.nx;
    ny = (int)g.ny;
    nz = (int)g.nz;
    nkx = (int)g.nkr;
    P = g.slab ? g.P : 1;
    rank = g.slab ? g.rank : 0;
    nyl = (int)g.nyl;
    nzl = (int)g.nzl;
    nspec = (size_t)nkx * nyl * nz;
    blk = (size_t)nkx * nyl * nzl;
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
    // Exchange mode for P > 1: peer-to-peer over NVLink (the z-column kernel gathers P^xy from the peers' send
    // buffers and stores A, C into the peers' receive buffers; CUDA IPC, no collective call on the step path), or
    // NCCL send/recv pipelined over kr chunks when PTF_F3_P2P=0 or the peers cannot be mapped.
    {
      const char* pe = std::getenv("PTF_F3_P2P");
      want_p2p = P > 1 && !(pe && std::atoi(pe) == 0);
    }
    for (auto* b : {&U1, &U3, &U4, &YA, &YB, &YC}) zalloc(*b);
    if (want_p2p) {
      try {
        setup_p2p();
      } catch (const Error& e) {
        fprintf(stderr, "libptf_b200: P2P slab exchange unavailable (%s); using the NCCL pipeline\n", e.what());
        p2p = false;
      }
    }
    zalloc(U2);
    if (P == 1) {
      ZA = U1.p; ZC = U2.p; RA = U1.p; RC = U2.p; PX = U3.p; PXY = U4.p; RP = U4.p;
    } else if (p2p) {   // PXY: local send buffer the peers read; RA, RC: receive buffers the peers write; P^x has its
                        // own buffer: a peer's z kernel may store chunk c of A while chunk c+1 of P^x is still being read
      PXY = U1.p; RA = U3.p; RC = U4.p; PX = U2.p; ZA = nullptr; ZC = nullptr; RP = nullptr;
    } else {
      ZA = U1.p; ZC = U2.p; RA = U3.p; RC = U4.p; PX = U3.p; PXY = U1.p; RP = U4.p;
    }
    if (base == PTF_STEPPER_ETDRK4)
      for (auto* b : {&cE, &cE2, &cZ, &cA, &cB, &cG}) b->alloc(nspec, &dev_bytes);
    twx.build(nx, &dev_bytes);
    if (ny != nx) twy_own.build(ny, &dev_bytes);
    if (nz != nx && nz != ny) twz_own.build(nz, &dev_bytes);
    vs.init(&g, ctx.stream, ctx.d.flow_kind, &dev_bytes);
    PTF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx.device));
    PTF_DISPATCH_N3(nx, fused3_prep);
    if (ny != nx) PTF_DISPATCH_N3(ny, fused3_prep);
    if (nz != nx && nz != ny) PTF_DISPATCH_N3(nz, fused3_prep);
    // cuFFT only at the set_c!/updatevars! boundary: batched 1-D transforms along x of the local planes
    long long n1d[1] = {nx};
    size_t wf = 0, wi = 0;
    PTF_CUFFT(cufftCreate(&plan_fwd));
    PTF_CUFFT(cufftCreate(&plan_inv));
    PTF_CUFFT(cufftMakePlanMany64(plan_fwd, 1, n1d, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, (long long)ny * nzl, &wf));
    PTF_CUFFT(cufftMakePlanMany64(plan_inv, 1, n1d, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, (long long)ny * nzl, &wi));
    PTF_CUFFT(cufftSetStream(plan_fwd, ctx.stream));
    PTF_CUFFT(cufftSetStream(plan_inv, ctx.stream));
    dev_bytes += (int64_t)(wf + wi);
    // pipelined exchange: kr chunks, exchanges on a high-priority stream overlapping the column kernels.
    // PTF_F3_CHUNKS = number of chunks (default 4); on one GPU it switches the chunked launch sequence on as a test
    // hook (every kernel launch and event of the slab pipeline, with the exchange itself a no-op).
    if (std::getenv("PTF_NO_GRAPH")) ctx.d.use_graph = 0;   // experiment knob: eager launches instead of a captured graph
    const char* ce = std::getenv("PTF_F3_CHUNKS");
    // P2P mode runs UNPIPELINED by default: measured on 8 B200 at 1024^3 (profiles/r02_slab3d_scaling.md), the
    // z-column kernel needs the whole machine's CTAs in flight to keep NVLink busy (33.8 ms/step with full-size launches
    // against 35.9 / 39.3 ms with 2 / 1 persistent CTAs per SM next to the y kernels).
    pipelined = (P > 1 && !p2p) || (ce && std::atoi(ce) > 1);
    if (P > 1 && !p2p) ctx.d.use_graph = 0;   // measured: the captured multi-stream NCCL pipeline runs 15 % slower
    if (pipelined) {
      n_chunks = 4;
      if (ce) {
        int c = std::atoi(ce);
        if (c >= 1 && c <= MAXCH) n_chunks = c;
      }
      if (n_chunks > nkx) n_chunks = nkx;
      if (n_chunks == 1 && P > 1 && p2p) pipelined = false;
      if (const char* ze = std::getenv("PTF_F3_ZCTAS")) z_ctas = std::atoi(ze);   // 0 = full-size grid
      int lo = 0, hi = 0;
      PTF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      PTF_CUDA(cudaStreamCreateWithPriority(&s_comm, cudaStreamNonBlocking, hi));
      for (auto& row : ev)
        for (auto& e : row) PTF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    on_dt_changed();
  }

  ~Fused3DEngine() override {
    drop_graphs();
    teardown_p2p();
    if (s_comm) cudaStreamDestroy(s_comm);
    for (auto& row : ev)
      for (auto& e : row)
        if (e) cudaEventDestroy(e);
    if (plan_fwd) cufftDestroy(plan_fwd);
    if (plan_inv) cufftDestroy(plan_inv);
  }

  const char* name() const override { return "fused3d"; }
  int id() const override { return PTF_ENGINE_FUSED; }
  cudaStream_t stream() const override { return ctx.stream; }

  Twiddles tw_x() const { return twx.dev(); }
  Twiddles tw_y() const { return ny != nx ? twy_own.dev() : twx.dev(); }
  Twiddles tw_z() const { return nz == nx ? twx.dev() : (nz == ny ? tw_y() : twz_own.dev()); }

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
    sep_dirty = true;
  }
  void set_velocity_coeffs(int comp, int nterms, const double* a) override {
    vs.set_coeffs(comp, nterms, a);
    sep_dirty = true;
  }
  void set_velocity_expr(int comp, const char* expr) override {
    ef.set(comp, expr);
    bool all = true;
    for (int a = 0; a < 3; ++a) all = all && !ef.expr[a].empty();
    if (all && ef.stale) {
      PTF_CUDA(cudaStreamSynchronize(ctx.stream));
      ef.compile(3);
    }
    sep_dirty = true;
  }
  void set_flow_time(double t) override {
    ef.set_time(t, ctx.stream);
    sep_dirty = true;
  }
  // separable / expression flows: the three fields of the step about to run, evaluated at clock.t once
  void refresh_separable() {
    if ((vs.va.kind != PTF_FLOW_SEPARABLE && vs.va.kind != PTF_FLOW_EXPR) || !sep_dirty) return;
    const size_t nreal = (size_t)nx * ny * nzl;
    for (auto& b : sepv)
      if (b.n != nreal) b.alloc(nreal, &dev_bytes);
    if (vs.va.kind == PTF_FLOW_EXPR) {
      ef.fill(ctx.stream, sepv[0].p, sepv[1].p, sepv[2].p, nx, ny, nzl, 0, g.zoff, g);
      ++own_launches;
      sep_dirty = false;
      return;
    }
    k_sep_fill<<<148 * 8, 256, 0, ctx.stream>>>(vs.va, sepv[0].p, sepv[1].p, sepv[2].p, nx, ny, nzl, nz, (int)g.zoff);
    ++own_launches;
    sep_dirty = false;
  }
  void set_layered_shift(const double*) override {
    throw Error(PTF_EUNSUPPORTED, "layered flows are 2-D per layer");
  }

  // ---------------- the exchange between the y- and z-column kernels (the only collective) ----------------
  // kr range [k0, k1) of every peer block (the whole block by default) on stream `st`
  void exchange(const double2* send, double2* recv, int k0 = 0, int k1 = -1, cudaStream_t st = nullptr) {
    if (P == 1 || p2p) return;  // same buffer / the z-column kernel moved the data itself
#ifdef PTF_WITH_NCCL
    if (k1 < 0) k1 = nkx;
    if (!st) st = ctx.stream;
    ncclComm_t comm = (ncclComm_t)ctx.nccl_comm;
    auto ck = [](ncclResult_t r, const char* what) {
      if (r != ncclSuccess) throw Error(PTF_ENCCL, std::string(what) + ": " + ncclGetErrorString(r));
    };
    const size_t off = (size_t)k0 * nyl * nzl, cnt = (size_t)(k1 - k0) * nyl * nzl;
    PTF_CUDA(cudaMemcpyAsync(recv + (size_t)rank * blk + off, send + (size_t)rank * blk + off, cnt * sizeof(double2),
                             cudaMemcpyDeviceToDevice, st));
    ck(ncclGroupStart(), "ncclGroupStart");
    for (int r = 0; r < P; ++r) {
      if (r == rank) continue;
      ck(ncclSend(send + (size_t)r * blk + off, 2 * cnt, ncclDouble, r, comm, st), "ncclSend");
      ck(ncclRecv(recv + (size_t)r * blk + off, 2 * cnt, ncclDouble, r, comm, st), "ncclRecv");
    }
    ck(ncclGroupEnd(), "ncclGroupEnd");
    ++lib_calls;
#else
    (void)send; (void)recv; (void)k0; (void)k1; (void)st;
    throw Error(PTF_EUNSUPPORTED, "built without NCCL");
#endif
  }
  int chunk_lo(int c) const { return (int)((long)nkx * c / n_chunks); }

  // ---------------- kernels ----------------
  YArgs zargs(int mode, double la, double lb, int llast) const {
    YArgs a;
    a.Px = RP;
    a.A = ZA;
    a.Bf = ZC;
    a.P = CombinePtrs{s0.p, s1.p, s2.p, acc.p, n1.p, cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p};
    a.C = CombineArgs{mode, ctx.st.filtered ? 1 : 0, ctx.dt, la, lb, llast};
    a.ax = ctx.ax;
    a.tw = tw_z();
    a.nkr = nkx * nyl;
    a.inv_n = 1.0 / (double)g.npts();
    int slot = mode < 0 ? 0 : next_state_slot(mode);
    a.next_state = slot == 0 ? s0.p : (slot == 1 ? s1.p : s2.p);
    for (int q = 0; q < 4; ++q) {
      a.pf_c[q] = nullptr;
      a.pf_r[q] = nullptr;
    }
    a.pf_ahead = 0;
    a.ablate = 0;
    a.stagger = 0;
    a.first_wave = 0;
    a.nkx = nkx;
    a.nyl = nyl;
    a.yoff = (int)g.yoff;
    a.nzl = nzl;
    a.zsh = ilog2(nzl);
    a.cid0 = 0;
    a.cid_end = nkx * nyl;
    for (int r = 0; r < 16; ++r) {
      a.Psrc[r] = nullptr;
      a.Adst[r] = nullptr;
      a.Cdst[r] = nullptr;
    }
    for (int r = 0; r < P; ++r) {
      if (p2p) {   // block (r -> me) of r's send buffer; block (me -> r) of r's receive buffers
        a.Psrc[r] = pxy_peer[r] + (size_t)rank * blk;
        a.Adst[r] = ra_peer[r] + (size_t)rank * blk;
        a.Cdst[r] = rc_peer[r] + (size_t)rank * blk;
      } else {
        a.Psrc[r] = RP + (size_t)r * blk;
        a.Adst[r] = ZA + (size_t)r * blk;
        a.Cdst[r] = ZC + (size_t)r * blk;
      }
    }
    return a;
  }

  // ---------------- P2P exchange: peers' buffers mapped with CUDA IPC, cross-GPU barrier kernel ----------------
  void setup_p2p() {
#ifdef PTF_WITH_NCCL
    ncclComm_t comm = (ncclComm_t)ctx.nccl_comm;
    flags.alloc(32, &dev_bytes);
    PTF_CUDA(cudaMemsetAsync(flags.p, 0, flags.bytes(), ctx.stream));
    struct Handles { cudaIpcMemHandle_t h[4]; };
    static_assert(sizeof(Handles) == 256, "IPC handle size");
    Handles mine;
    PTF_CUDA(cudaIpcGetMemHandle(&mine.h[0], U1.p));
    PTF_CUDA(cudaIpcGetMemHandle(&mine.h[1], U3.p));
    PTF_CUDA(cudaIpcGetMemHandle(&mine.h[2], U4.p));
    PTF_CUDA(cudaIpcGetMemHandle(&mine.h[3], flags.p));
    DevBuf<unsigned char> hb;
    hb.alloc((size_t)P * sizeof(Handles));
    PTF_CUDA(cudaMemcpyAsync(hb.p + (size_t)rank * sizeof(Handles), &mine, sizeof(Handles), cudaMemcpyHostToDevice,
                             ctx.stream));
    ncclResult_t r = ncclAllGather(hb.p + (size_t)rank * sizeof(Handles), hb.p, sizeof(Handles), ncclUint8, comm,
                                   ctx.stream);
    if (r != ncclSuccess) throw Error(PTF_ENCCL, std::string("ncclAllGather(IPC handles): ") + ncclGetErrorString(r));
    std::vector<Handles> all(P);
    PTF_CUDA(cudaMemcpyAsync(all.data(), hb.p, (size_t)P * sizeof(Handles), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    int ok = 1;
    for (int q = 0; q < P && ok; ++q) {
      if (q == rank) {
        pxy_peer[q] = U1.p; ra_peer[q] = U3.p; rc_peer[q] = U4.p; flag_peer[q] = flags.p;
        continue;
      }
      void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
      for (int i = 0; i < 4; ++i) {
        cudaError_t e = cudaIpcOpenMemHandle(&ptr[i], all[q].h[i], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          cudaGetLastError();
          ok = 0;
          break;
        }
        opened.push_back(ptr[i]);
      }
      if (!ok) break;
      pxy_peer[q] = (const double2*)ptr[0];
      ra_peer[q] = (double2*)ptr[1];
      rc_peer[q] = (double2*)ptr[2];
      flag_peer[q] = (unsigned*)ptr[3];
    }
    // all ranks must agree (a rank that failed to map a peer would otherwise wait for barriers nobody joins)
    DevBuf<int> agree;
    agree.alloc(1);
    PTF_CUDA(cudaMemcpyAsync(agree.p, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
    ncclAllReduce(agree.p, agree.p, 1, ncclInt, ncclMin, comm, ctx.stream);
    PTF_CUDA(cudaMemcpyAsync(&ok, agree.p, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    if (!ok) {
      for (void* q : opened) cudaIpcCloseMemHandle(q);
      opened.clear();
      throw Error(PTF_EUNSUPPORTED, "cudaIpcOpenMemHandle failed on at least one rank");
    }
    p2p = true;
#else
    throw Error(PTF_EUNSUPPORTED, "built without NCCL");
#endif
  }
  void teardown_p2p() {
    if (!p2p) return;
#ifdef PTF_WITH_NCCL
    cudaStreamSynchronize(ctx.stream);
    for (void* q : opened) cudaIpcCloseMemHandle(q);
    opened.clear();
    // nobody frees a buffer a peer may still have mapped: one tiny collective as a host-level barrier
    ncclComm_t comm = (ncclComm_t)ctx.nccl_comm;
    DevBuf<int> t;
    t.alloc(1);
    cudaMemsetAsync(t.p, 0, sizeof(int), ctx.stream);
    ncclAllReduce(t.p, t.p, 1, ncclInt, ncclSum, comm, ctx.stream);
    cudaStreamSynchronize(ctx.stream);
#endif
    p2p = false;
  }
  void xbarrier(cudaStream_t st = nullptr) {
    if (!p2p) return;
    if (!st) st = ctx.stream;
    XbArgs a;
    for (int q = 0; q < 16; ++q) a.peer[q] = q < P ? flag_peer[q] : nullptr;
    a.me = rank;
    a.P = P;
    k_xbarrier<<<1, 32, 0, st>>>(flags.p, a);
    ++own_launches;
  }

  void run_z(bool has_in, int mode, double la = 0, double lb = 0, int llast = 0, int fam_override = -1,
             bool unmasked = false, int k0 = 0, int k1 = -1, cudaStream_t st = nullptr, int grid_cap = 0) {
    if (!st) st = ctx.stream;
    YArgs a = zargs(mode, la, lb, llast);
    if (k1 >= 0) {
      a.cid0 = k0 * nyl;
      a.cid_end = k1 * nyl;
    }
    a.grid_cap = grid_cap;
    if (unmasked) a.ax.dealias = 0;   // updatevars! transforms sol as it is stored (TAD.jl:815-821)
    int fam = (ctx.st.base == PTF_STEPPER_RK4) ? FAM_RK4 : (ctx.st.base == PTF_STEPPER_ETDRK4 ? FAM_ETD : FAM_OTHER);
    if (fam_override >= 0) fam = fam_override;
    // P2P mode: this kernel reads the peers' P^xy and writes the peers' A, C.  Barrier before: every rank has finished
    // writing its P^xy and reading its previous A, C.  Barrier after: all A, C have landed, all P^xy reads are done.
    xbarrier(st);
    PTF_DISPATCH_N3(nz, fused3_launch_z, has_in, fam, &a, st, n_sm);
    ++own_launches;
    xbarrier(st);
  }

  Y3Args y3args() const {
    Y3Args a;
    a.RA = RA;
    a.RC = RC;
    a.YA = YA.p;
    a.YB = YB.p;
    a.YC = YC.p;
    a.PX = PX;
    a.PXY = PXY;
    a.cy = ctx.ax.cy;
    a.nyq_sign = ctx.ax.nyq_sign;
    a.tw = tw_y();
    a.nkx = nkx;
    a.nyl = nyl;
    a.nzl = nzl;
    // l = t + T*e, T = ny/16: rank index = e >> esh with 2^esh = 16/P elements per thread and rank
    const long long T = ny / 16;
    a.esh = ilog2(16 / P);
    a.in_se = T * nzl;                                   // [r][kr][ll][zl]
    a.in_sr = (long long)blk - (long long)nyl * nzl;
    a.out_se = T * 8;                                    // [p][kr][zl/8][ll][zl%8]
    a.out_sr = (long long)blk - (long long)nyl * 8;
    a.kr0 = 0;
    a.nkr_launch = nkx;
    return a;
  }
  void run_yinv(int k0 = 0, int k1 = -1) {
    Y3Args a = y3args();
    if (k1 >= 0) {
      a.kr0 = k0;
      a.nkr_launch = k1 - k0;
    }
    PTF_DISPATCH_N3(ny, fused3_launch_y, true, &a, ctx.stream, n_sm);
    ++own_launches;
  }
  void run_yfwd(int k0 = 0, int k1 = -1) {
    Y3Args a = y3args();
    if (k1 >= 0) {
      a.kr0 = k0;
      a.nkr_launch = k1 - k0;
    }
    PTF_DISPATCH_N3(ny, fused3_launch_y, false, &a, ctx.stream, n_sm);
    ++own_launches;
  }

  void run_x() {
    XArgs a;
    a.A = YA.p;
    a.Bf = YB.p;
    a.Cf = YC.p;
    a.Px = PX;
    a.va = vs.va;
    a.ax = ctx.ax;
    a.tw = tw_x();
    a.nkr = nkx;
    a.ny = ny;
    a.pf_vel = 0;
    a.pf_ahead = 0;
    a.ablate = 0;
    a.stagger = 0;
    a.first_wave = 0;
    a.nzg = nz;
    a.zoff = (int)g.zoff;
    // direct mode (default): velocity rows prefetched to L2 and loaded after the transforms, partial products parked in
    // TMEM -> 4 instead of 3 CTAs per SM (measured: 256^3 row kernel 0.252 -> 0.185 ms, 1024^3 19.8 -> 17.2 ms)
    const char* xd = std::getenv("PTF_X_DIRECT");
    int vmode = (xd && std::atoi(xd) == 0) ? 0 : 3;
    if (vs.va.kind == PTF_FLOW_SEPARABLE || vs.va.kind == PTF_FLOW_EXPR) {
      if (!sepv[0].p) throw Error(PTF_EINVAL, "separable tables / velocity expressions have not been set");
      for (int c = 0; c < 3; ++c) a.va.arr[c] = sepv[c].p;
    } else if (!vs.va.arr[0] || !vs.va.arr[1] || !vs.va.arr[2]) {
      throw Error(PTF_EINVAL, "velocity fields have not been set (ptf_set_velocity / callback)");
    }
    PTF_DISPATCH_N3(nx, fused3_launch_x, vmode, &a, nzl, ctx.stream, n_sm);
    ++own_launches;
  }

  // A, C of the current sol (fresh state): one z-column launch without input + the exchange
  void prologue() {
    run_z(false, -1);
    exchange(ZA, RA);
    exchange(ZC, RC);
    ac_valid = true;
  }

  // One stage.  Single GPU: yinv -> x -> yfwd -> z.  Slab-decomposed: the same with the two exchanges pipelined over
  // kr chunks on the high-priority stream, so that chunk c travels while chunk c+1 is being computed:
  //   main:  x | yfwd(0) yfwd(1) ... | z(0) z(1) ...            | (next stage) yinv(0) yinv(1) ... x
  //   comm:      P^xy(0)  P^xy(1) ...  |  A,C(0)  A,C(1) ...
  // The stage therefore starts with the y-inverse kernels of the chunks as they arrive.
  void stage(int mode, double la = 0, double lb = 0, int llast = 0) {
    if (!pipelined) {
      run_yinv();
      run_x();
      run_yfwd();
      exchange(PXY, RP);
      run_z(true, mode, la, lb, llast);
      exchange(ZA, RA);
      exchange(ZC, RC);
      return;
    }
    const int nc = n_chunks;
    for (int c = 0; c < nc; ++c) {   // A, C chunks of the previous stage (or of the prologue) as they arrive
      if (ac_in_flight) PTF_CUDA(cudaStreamWaitEvent(ctx.stream, ev[3][c], 0));
      run_yinv(chunk_lo(c), chunk_lo(c + 1));
    }
    run_x();
    if (p2p) {
      // P2P mode: the z-column kernel of chunk c (NVLink-bound: it reads the peers' P^xy and writes the peers' A, C)
      // runs on the priority stream next to the HBM-bound y kernels of the other chunks on the main stream.
      for (int c = 0; c < nc; ++c) {
        const int k0 = chunk_lo(c), k1 = chunk_lo(c + 1);
        run_yfwd(k0, k1);
        PTF_CUDA(cudaEventRecord(ev[0][c], ctx.stream));
        PTF_CUDA(cudaStreamWaitEvent(s_comm, ev[0][c], 0));
        // barrier, kernel, barrier on the priority stream; a small persistent grid (z_ctas CTAs per SM)
        run_z(true, mode, la, lb, llast, -1, false, k0, k1, s_comm, z_ctas * n_sm);
        PTF_CUDA(cudaEventRecord(ev[3][c], s_comm));
      }
      ac_in_flight = true;
      return;
    }
    for (int c = 0; c < nc; ++c) {
      const int k0 = chunk_lo(c), k1 = chunk_lo(c + 1);
      run_yfwd(k0, k1);
      PTF_CUDA(cudaEventRecord(ev[0][c], ctx.stream));
      PTF_CUDA(cudaStreamWaitEvent(s_comm, ev[0][c], 0));
      exchange(PXY, RP, k0, k1, s_comm);
      PTF_CUDA(cudaEventRecord(ev[1][c], s_comm));
    }
    for (int c = 0; c < nc; ++c) {
      const int k0 = chunk_lo(c), k1 = chunk_lo(c + 1);
      PTF_CUDA(cudaStreamWaitEvent(ctx.stream, ev[1][c], 0));
      // (single-GPU test hook: PTF_F3_ZCTAS also caps this launch, exercising the persistent grid-stride path)
      run_z(true, mode, la, lb, llast, -1, false, k0, k1, nullptr,
            (P == 1 && std::getenv("PTF_F3_ZCTAS")) ? z_ctas * n_sm : 0);
      PTF_CUDA(cudaEventRecord(ev[2][c], ctx.stream));
      PTF_CUDA(cudaStreamWaitEvent(s_comm, ev[2][c], 0));
      exchange(ZA, RA, k0, k1, s_comm);
      exchange(ZC, RC, k0, k1, s_comm);
      PTF_CUDA(cudaEventRecord(ev[3][c], s_comm));
    }
    ac_in_flight = true;
  }
  // the main stream waits for the A, C exchanges still in flight on the comm stream (end of a step)
  void join_exchanges() {
    if (!ac_in_flight) return;
    for (int c = 0; c < n_chunks; ++c) PTF_CUDA(cudaStreamWaitEvent(ctx.stream, ev[3][c], 0));
    ac_in_flight = false;
  }

  // ---------------- boundary ----------------
  int flat_blocks(int64_t n) const {
    int64_t b = (n + 255) / 256;
    int64_t cap = 148 * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
  }

  void set_c(const double* c_host, bool replicate) override {
    (void)replicate;  // nbatch == 1
    double* real = reinterpret_cast<double*>(YB.p);
    const int64_t nreal = (int64_t)nx * ny * nzl;
    PTF_CUDA(cudaMemcpyAsync(real, c_host, nreal * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    PTF_CUFFT(cufftExecD2Z(plan_fwd, real, reinterpret_cast<cufftDoubleComplex*>(YC.p)));   // [zl][y][kr]
    ++lib_calls;
    k_block_px<<<flat_blocks((int64_t)nkx * ny * nzl), 256, 0, ctx.stream>>>(YC.p, PX, nkx, ny, nzl);
    ++own_launches;
    run_yfwd();
    exchange(PXY, RP);
    run_z(true, CM_STORE, 0, 0, 0, FAM_OTHER);   // sol = FFT_z(...) ; leaves A, C of the new sol behind
    exchange(ZA, RA);
    exchange(ZC, RC);
    ac_valid = true;
    PTF_CUDA(cudaGetLastError());
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void get_c(double* c_host) override {
    if (ctx.ax.dealias) {  // A of the UNMASKED sol (the scratch then no longer holds what the next calcN needs)
      run_z(false, -1, 0, 0, 0, -1, true);
      exchange(ZA, RA);
      ac_valid = false;
    } else {
      prologue();          // A = IFFT_z(sol / N) in the plane owners' layout
    }
    run_yinv();            // A' = IFFT_y(A)
    dim3 grid((ny + 31) / 32, (nkx + 31) / 32, nzl), block(32, 8, 1);
    k_unblock_c2r<<<grid, block, 0, ctx.stream>>>(YA.p, YC.p, nkx, ny, nzl);
    ++own_launches;
    double* real = reinterpret_cast<double*>(YB.p);
    PTF_CUFFT(cufftExecZ2D(plan_inv, reinterpret_cast<cufftDoubleComplex*>(YC.p), real));
    ++lib_calls;
    PTF_CUDA(cudaMemcpyAsync(c_host, real, (size_t)nx * ny * nzl * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void set_sol(const double* s_host) override {
    PTF_CUDA(cudaMemcpyAsync(YA.p, s_host, nspec * sizeof(double2), cudaMemcpyHostToDevice, ctx.stream));
    dim3 grid((nz + 31) / 32, (nkx + 31) / 32, nyl), block(32, 8, 1);
    k_state_canon<1><<<grid, block, 0, ctx.stream>>>(YA.p, s0.p, nkx, nyl, nz);
    ++own_launches;
    ac_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void get_sol(double* s_host) override {
    dim3 grid((nz + 31) / 32, (nkx + 31) / 32, nyl), block(32, 8, 1);
    k_state_canon<0><<<grid, block, 0, ctx.stream>>>(s0.p, YA.p, nkx, nyl, nz);
    ++own_launches;
    PTF_CUDA(cudaMemcpyAsync(s_host, YA.p, nspec * sizeof(double2), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void on_dt_changed() override {
    drop_graphs();
    if (ctx.st.base == PTF_STEPPER_ETDRK4) {
      k_etd_coeffs<<<1184, 256, 0, ctx.stream>>>(cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p, ctx.ax, nkx, nyl, nz, ctx.dt, 2,
                                                 g.yoff);
      ++own_launches;
      PTF_CUDA(cudaGetLastError());
    }
  }

  void enqueue_step(int variant) {
    static const double LA[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                 -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
    static const double LB[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                                 1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                                 2277821191437.0 / 14882151754819.0};
    switch (ctx.st.base) {
      case PTF_STEPPER_RK4:
        stage(CM_RK4_S1); stage(CM_RK4_S2); stage(CM_RK4_S3); stage(CM_RK4_S4);
        break;
      case PTF_STEPPER_ETDRK4:
        stage(CM_ETD_S1); stage(CM_ETD_S2); stage(CM_ETD_S3); stage(CM_ETD_S4);
        break;
      case PTF_STEPPER_FORWARD_EULER:
        stage(CM_EULER);
        break;
      case PTF_STEPPER_LSRK54:
        for (int i = 0; i < 5; ++i) stage(CM_LSRK, LA[i], LB[i], i == 4);
        break;
      case PTF_STEPPER_AB3:
        stage(variant == 1 ? CM_AB3_EULER : CM_AB3);
        break;
    }
    join_exchanges();   // a step (and a captured graph) ends with every exchange joined back into the step stream
  }

  void step_once(int64_t step_index) override {
    refresh_separable();
    if (!ac_valid) prologue();
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
    k_diag_t<<<1184, 256, 0, ctx.stream>>>(s0.p, nkx, (int64_t)nyl * nz, 1, nx, out.p);
    ++own_launches;
    PTF_CUDA(cudaMemcpyAsync(out.p + 2, s0.p, sizeof(double2), cudaMemcpyDeviceToDevice, ctx.stream));  // DC mode
#ifdef PTF_WITH_NCCL
    if (P > 1) {  // spectral rows are spread over the ranks; the DC mode lives on rank 0
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
    double m = h[2] / N;
    double msq = h[0] / (N * N);
    if (mean_c) *mean_c = m;
    if (var_c) *var_c = msq - m * m;
    if (max_abs_sol) *max_abs_sol = std::sqrt(h[1]);
  }

  // Average device time of one of the four stage kernels over `reps` REAL RK4 steps (events around every launch of
  // that kernel on the step stream; unpipelined sequence).  "exchange" = one field's all-to-all (three per stage),
  // run alone on the stream, i.e. the time the collective costs when nothing hides it.
  // The solution is backed up and restored, the clock is untouched.
  float time_kernel(const char* kname, int reps) override {
    std::string k(kname ? kname : "");
    int which = k == "zkernel" ? 0 : k == "yinv" ? 1 : k == "xkernel" ? 2 : k == "yfwd" ? 3 : k == "exchange" ? 4 : -1;
    if (which < 0)
      throw Error(PTF_EINVAL, "fused 3-D engine: unknown kernel '" + k + "' (zkernel|yinv|xkernel|yfwd|exchange)");
    if (which == 4 && P == 1) throw Error(PTF_EUNSUPPORTED, "no exchange on a single GPU");
    if (which == 4 && p2p)
      throw Error(PTF_EUNSUPPORTED, "P2P mode: the exchange is fused into the z-column kernel's loads and stores");
    if (ctx.st.base != PTF_STEPPER_RK4) throw Error(PTF_EUNSUPPORTED, "kernel timing is implemented for RK4 steps");
    DevBuf<double2> backup;
    backup.alloc(nspec);
    PTF_CUDA(cudaMemcpyAsync(backup.p, s0.p, s0.bytes(), cudaMemcpyDeviceToDevice, ctx.stream));
    refresh_separable();
    if (!ac_valid) prologue();
    const int modes[4] = {CM_RK4_S1, CM_RK4_S2, CM_RK4_S3, CM_RK4_S4};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
    bool record = false;
    auto timed = [&](int id, auto&& fn) {
      const bool on = record && which == id;
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (on) {
        PTF_CUDA(cudaEventCreate(&e0));
        PTF_CUDA(cudaEventCreate(&e1));
        PTF_CUDA(cudaEventRecord(e0, ctx.stream));
      }
      fn();
      if (on) {
        PTF_CUDA(cudaEventRecord(e1, ctx.stream));
        evs.push_back({e0, e1});
      }
    };
    auto one_step = [&]() {
      for (int s = 0; s < 4; ++s) {
        timed(1, [&] { run_yinv(); });
        timed(2, [&] { run_x(); });
        timed(3, [&] { run_yfwd(); });
        timed(4, [&] { exchange(PXY, RP); });
        timed(0, [&] { run_z(true, modes[s]); });
        timed(4, [&] { exchange(ZA, RA); });
        timed(4, [&] { exchange(ZC, RC); });
      }
    };
    one_step();  // warm
    record = true;
    for (int r = 0; r < reps; ++r) one_step();
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    double total = 0;
    for (auto& e : evs) {
      float ms = 0;
      PTF_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
      total += ms;
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
    PTF_CUDA(cudaMemcpyAsync(s0.p, backup.p, s0.bytes(), cudaMemcpyDeviceToDevice, ctx.stream));
    ac_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    return (float)(total / (double)evs.size());
  }

 private:
  Context& ctx;
  Geometry& g;
  int nx, ny, nz, nkx, nyl, nzl, P, rank;
  size_t nspec, blk;
  DevBuf<double2> s0, s1, s2, acc, n1, U1, U2, U3, U4, YA, YB, YC;
  double2 *ZA = nullptr, *ZC = nullptr, *RA = nullptr, *RC = nullptr, *PX = nullptr, *PXY = nullptr, *RP = nullptr;
  DevBuf<double> cE, cE2, cZ, cA, cB, cG;
  TwiddleSet twx, twy_own, twz_own;
  VelocityStore vs;
  DevBuf<double> sepv[3];   // separable / expression flows: u, v, w of the current step (local planes)
  ExprFlow ef;              // PTF_FLOW_EXPR: run-time compiled expressions (fill kernel)
  bool sep_dirty = true;
  cufftHandle plan_fwd = 0, plan_inv = 0;
  bool ac_valid = false;
  // P2P slab exchange (CUDA IPC over NVLink)
  bool want_p2p = false, p2p = false;
  DevBuf<unsigned> flags;            // [0..15] peers' epochs, [16] own barrier count
  std::vector<void*> opened;
  const double2* pxy_peer[16] = {};
  double2* ra_peer[16] = {};
  double2* rc_peer[16] = {};
  unsigned* flag_peer[16] = {};
  // pipelined slab exchange (NCCL mode)
  static constexpr int MAXCH = 16;
  int n_chunks = 1, z_ctas = 1;
  bool pipelined = false, ac_in_flight = false;
  cudaStream_t s_comm = nullptr;
  cudaEvent_t ev[4][MAXCH] = {};
  int n_sm = 148;
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int64_t per_step_own = 0, per_step_lib = 0;
};

}  // namespace

bool fused3d_engine_supports(const Context& ctx, std::string* why) {
  auto no = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  const Geometry& g = ctx.g;
  if (g.ndim != 3) return no("not a 3-D problem");
  if (!is_fused3_size(g.nx) || !is_fused3_size(g.ny) || !is_fused3_size(g.nz))
    return no("nx, ny and nz must be powers of two in [64, 1024]");
  if (g.B != 1) return no("3-D ensembles run on the cuFFT engine");
  if (ctx.d.flow_kind == PTF_FLOW_LAYERED) return no("layered flows are 2-D per layer");
  const int P = g.slab ? g.P : 1;
  if ((P & (P - 1)) || P > 16) return no("the number of ranks must be a power of two, at most 16");
  if (g.nzl < 8 || g.nyl < 1) return no("each rank needs at least 8 z planes");
  return true;
}

std::unique_ptr<Engine> make_fused3d_engine(Context& ctx) {
  std::string why;
  if (!fused3d_engine_supports(ctx, &why)) throw Error(PTF_EUNSUPPORTED, "fused 3-D engine: " + why);
  return std::unique_ptr<Engine>(new Fused3DEngine(ctx));
}

}  // namespace ptf
