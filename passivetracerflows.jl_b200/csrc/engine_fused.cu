// Fused 2-D engine, host side: buffers, layouts at the boundary, graph capture, stage sequencing.  The kernels live
// in fused_kernels.cuh and are instantiated per transform length in fused_inst.cu (one object file per length).
#include "fused_kernels.cuh"
#include "expr_flow.h"

namespace ptf {

// per-length entry points (fused_inst.cu)
#define PTF_DECL_INST(N)                                                                                   \
  void fused_prep_##N();                                                                                   \
  void fused_launch_y_##N(bool has_in, int fam, const void* yargs, int nb, cudaStream_t st, int n_sm);     \
  void fused_launch_x_##N(int vmode, const void* xargs, int nb, cudaStream_t st, int n_sm);                \
  void fused_selftest_##N(int dir, int count, const double2* in, double2* out, const void* tw);
PTF_DECL_INST(64)
PTF_DECL_INST(128)
PTF_DECL_INST(256)
PTF_DECL_INST(512)
PTF_DECL_INST(1024)
PTF_DECL_INST(2048)
PTF_DECL_INST(4096)

#define PTF_DISPATCH_N(n, FN, ...)                                                               \
  switch (n) {                                                                                   \
    case 64: FN##_64(__VA_ARGS__); break;                                                        \
    case 128: FN##_128(__VA_ARGS__); break;                                                      \
    case 256: FN##_256(__VA_ARGS__); break;                                                      \
    case 512: FN##_512(__VA_ARGS__); break;                                                      \
    case 1024: FN##_1024(__VA_ARGS__); break;                                                    \
    case 2048: FN##_2048(__VA_ARGS__); break;                                                    \
    case 4096: FN##_4096(__VA_ARGS__); break;                                                    \
    default: throw Error(PTF_EUNSUPPORTED, "fused engine: unsupported transform length");        \
  }

namespace {

class FusedEngine final : public Engine {
 public:
  explicit FusedEngine(Context& c) : ctx(c), g(c.g) {
    nx = (int)g.nx;
    ny = (int)g.ny;
    nkr = (int)g.nkr;
    nb = (int)g.B;
    nspec = (size_t)nkr * ny * nb;
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
    zalloc(A);
    zalloc(Bf);
    zalloc(Px);
    if (base == PTF_STEPPER_ETDRK4)
      for (auto* b : {&cE, &cE2, &cZ, &cA, &cB, &cG}) b->alloc((size_t)nkr * ny, &dev_bytes);
    twx.build(nx, &dev_bytes);
    if (ny != nx) twy_own.build(ny, &dev_bytes);
    vs.init(&g, ctx.stream, ctx.d.flow_kind, &dev_bytes);
    {  // tuning knobs (defaults chosen from B200 measurements, profiles/); env overrides for experiments
      auto env_int = [](const char* name, int dflt) {
        const char* v = std::getenv(name);
        return v ? std::atoi(v) : dflt;
      };
      tune_pf_state = env_int("PTF_PF_STATE", 0);
      tune_pf_vel = env_int("PTF_PF_VEL", 0);
      tune_pf_ahead_y = env_int("PTF_PF_AHEAD_Y", 0);
      tune_pf_ahead_x = env_int("PTF_PF_AHEAD_X", 0);
      tune_ablate_x = env_int("PTF_ABLATE_X", 0);
      // measured on B200 at 4096^2 (profiles/r01_fft_core_experiments.md): -2 % step time; only applied to launches
      // of >= 4 waves (see run_x / yargs), where an 8-10 us head start is negligible
      tune_stagger_x = env_int("PTF_STAGGER_X", 20000);
      tune_stagger_y = env_int("PTF_STAGGER_Y", 25000);
      PTF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx.device));
      tune_ablate_y = env_int("PTF_ABLATE_Y", 0);
      // 1: velocities loaded after the transforms (row kernel 0.229 -> 0.206 ms at 4096^2); 2 (default): row 0's product
      // parked in TMEM too, no shared-memory parking row, no spills (0.206 -> 0.192 ms; profiles/r02_vmode4.txt)
      tune_x_direct = env_int("PTF_X_DIRECT", 2);
    }
    PTF_DISPATCH_N(ny, fused_prep);
    if (nx != ny) PTF_DISPATCH_N(nx, fused_prep);
    // cuFFT only at the set_c!/updatevars! boundary (canonical <-> transposed layout), never on the step path
    long long n[2] = {ny, nx};
    size_t wf = 0, wi = 0;
    PTF_CUFFT(cufftCreate(&plan_fwd));
    PTF_CUFFT(cufftCreate(&plan_inv));
    PTF_CUFFT(cufftSetAutoAllocation(plan_fwd, 0));
    PTF_CUFFT(cufftSetAutoAllocation(plan_inv, 0));
    PTF_CUFFT(cufftMakePlanMany64(plan_fwd, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, nb, &wf));
    PTF_CUFFT(cufftMakePlanMany64(plan_inv, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, nb, &wi));
    work_bytes = wf > wi ? wf : wi;
    // the cuFFT work area aliases P^x (dead whenever a boundary transform runs) when it fits
    if (work_bytes > Px.bytes()) work.alloc(work_bytes, &dev_bytes);
    void* wa = work.p ? (void*)work.p : (void*)Px.p;
    PTF_CUFFT(cufftSetWorkArea(plan_fwd, wa));
    PTF_CUFFT(cufftSetWorkArea(plan_inv, wa));
    PTF_CUFFT(cufftSetStream(plan_fwd, ctx.stream));
    PTF_CUFFT(cufftSetStream(plan_inv, ctx.stream));
    on_dt_changed();
  }

  ~FusedEngine() override {
    drop_graphs();
    if (plan_fwd) cufftDestroy(plan_fwd);
    if (plan_inv) cufftDestroy(plan_inv);
  }

  const char* name() const override { return "fused"; }
  int id() const override { return PTF_ENGINE_FUSED; }
  cudaStream_t stream() const override { return ctx.stream; }

  Twiddles tw_x() const { return twx.dev(); }
  Twiddles tw_y() const { return ny != nx ? twy_own.dev() : twx.dev(); }

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
  void set_velocity_external(int comp, const double* dev, int64_t count) override {
    vs.set_external(comp, dev, count);
    sync_vel();
  }
  // PTF_FLOW_EXPR: the expressions are written out as fields u, v once per step (the velocity is frozen at clock.t for
  // all stages, TAD.jl:718) by a run-time compiled fill kernel; the row kernel reads them like steady arrays
  void set_velocity_expr(int comp, const char* expr) override {
    ef.set(comp, expr);
    bool all = !ef.expr[0].empty() && !ef.expr[1].empty();
    if (all && ef.stale) {
      PTF_CUDA(cudaStreamSynchronize(ctx.stream));
      ef.compile(2);
    }
    expr_dirty = true;
  }
  void set_flow_time(double t) override {
    ef.set_time(t, ctx.stream);
    expr_dirty = true;
  }
  void refresh_expr() {
    if (ctx.d.flow_kind != PTF_FLOW_EXPR || !expr_dirty) return;
    const size_t nreal = (size_t)nx * ny;
    bool fresh = false;
    for (int c = 0; c < 2; ++c)
      if (exprv[c].n != nreal) {
        exprv[c].alloc(nreal, &dev_bytes);
        fresh = true;
      }
    if (fresh) {   // one field shared by all members
      vs.va.arr[0] = exprv[0].p;
      vs.va.arr[1] = exprv[1].p;
      vs.va.member_stride = 0;
      drop_graphs();
    }
    ef.fill(ctx.stream, exprv[0].p, exprv[1].p, nullptr, nx, ny, 1, 0, 0, g);
    ++own_launches;
    expr_dirty = false;
  }

  // ---------------- layout changes at the boundary ----------------
  void transpose(const double2* in, double2* out, int R, int Cc, double scale) {
    dim3 grid((Cc + 31) / 32, (R + 31) / 32, nb), block(32, 8, 1);
    k_transpose<<<grid, block, 0, ctx.stream>>>(in, out, R, Cc, scale);
    ++own_launches;
  }

  void set_c(const double* c_host, bool replicate) override {
    int64_t npts = g.npts();
    double* real = reinterpret_cast<double*>(Bf.p);
    if (replicate && nb > 1) {
      PTF_CUDA(cudaMemcpyAsync(real, c_host, npts * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
      k_replicate_f<<<1184, 256, 0, ctx.stream>>>(real, npts, nb);
      ++own_launches;
    } else {
      PTF_CUDA(cudaMemcpyAsync(real, c_host, npts * nb * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    }
    PTF_CUFFT(cufftExecD2Z(plan_fwd, real, reinterpret_cast<cufftDoubleComplex*>(A.p)));  // canonical [b][l][kr]
    ++lib_calls;
    transpose(A.p, s0.p, ny, nkr, 1.0);
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void get_c(double* c_host) override {
    double* real = reinterpret_cast<double*>(Bf.p);
    transpose(s0.p, A.p, nkr, ny, 1.0 / (double)g.npts());
    PTF_CUFFT(cufftExecZ2D(plan_inv, reinterpret_cast<cufftDoubleComplex*>(A.p), real));
    ++lib_calls;
    PTF_CUDA(cudaMemcpyAsync(c_host, real, g.npts() * nb * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void set_sol(const double* s_host) override {
    PTF_CUDA(cudaMemcpyAsync(A.p, s_host, A.bytes(), cudaMemcpyHostToDevice, ctx.stream));
    transpose(A.p, s0.p, ny, nkr, 1.0);
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void get_sol(double* s_host) override {
    transpose(s0.p, A.p, nkr, ny, 1.0);
    PTF_CUDA(cudaMemcpyAsync(s_host, A.p, A.bytes(), cudaMemcpyDeviceToHost, ctx.stream));
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void on_dt_changed() override {
    drop_graphs();
    if (ctx.st.base == PTF_STEPPER_ETDRK4) {
      k_etd_coeffs<<<1184, 256, 0, ctx.stream>>>(cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p, ctx.ax, nkr, ny, 1, ctx.dt, 1);
      ++own_launches;
      PTF_CUDA(cudaGetLastError());
    }
  }

  // ---------------- the two hot kernels ----------------
  YArgs yargs(int mode, double la, double lb, int llast) const {
    YArgs a{};
    a.Px = Px.p;
    a.A = A.p;
    a.Bf = Bf.p;
    a.P = CombinePtrs{s0.p, s1.p, s2.p, acc.p, n1.p, cE.p, cE2.p, cZ.p, cA.p, cB.p, cG.p};
    a.C = CombineArgs{mode, ctx.st.filtered ? 1 : 0, ctx.dt, la, lb, llast};
    a.ax = ctx.ax;
    a.tw = tw_y();
    a.nkr = nkr;
    a.inv_n = 1.0 / (double)g.npts();
    int slot = mode < 0 ? 0 : next_state_slot(mode);
    a.next_state = slot == 0 ? s0.p : (slot == 1 ? s1.p : s2.p);
    for (int q = 0; q < 4; ++q) {
      a.pf_c[q] = nullptr;
      a.pf_r[q] = nullptr;
    }
    if (tune_pf_state) {
      switch (mode) {
        case CM_RK4_S1: a.pf_c[0] = s0.p; break;
        case CM_RK4_S2: case CM_RK4_S3: case CM_RK4_S4:
          a.pf_c[0] = s0.p; a.pf_c[1] = s1.p; a.pf_c[2] = acc.p; break;
        case CM_ETD_S1: case CM_ETD_S2: a.pf_c[0] = s0.p; a.pf_r[0] = cE2.p; a.pf_r[1] = cZ.p; break;
        case CM_ETD_S3:
          a.pf_c[0] = s1.p; a.pf_c[1] = n1.p; a.pf_c[2] = acc.p; a.pf_r[0] = cE2.p; a.pf_r[1] = cZ.p; break;
        case CM_ETD_S4:
          a.pf_c[0] = s0.p; a.pf_c[1] = n1.p; a.pf_c[2] = acc.p;
          a.pf_r[0] = cE.p; a.pf_r[1] = cA.p; a.pf_r[2] = cB.p; a.pf_r[3] = cG.p; break;
        default: break;
      }
    }
    a.pf_ahead = tune_pf_ahead_y;
    a.ablate = tune_ablate_y;
    a.first_wave = 2 * n_sm;
    {
      const int fy = 256 / (ny / 16);
      const long ctas = (long)((nkr + fy - 1) / fy) * nb;
      a.stagger = ctas >= 4L * a.first_wave ? tune_stagger_y : 0;
    }
    return a;
  }

  void run_y(bool has_in, bool has_out, int mode, double la = 0, double lb = 0, int llast = 0) {
    (void)has_out;
    YArgs a = yargs(mode, la, lb, llast);
    int fam = (ctx.st.base == PTF_STEPPER_RK4) ? FAM_RK4 : (ctx.st.base == PTF_STEPPER_ETDRK4 ? FAM_ETD : FAM_OTHER);
    PTF_DISPATCH_N(ny, fused_launch_y, has_in, fam, &a, nb, ctx.stream, n_sm);
    ++own_launches;
  }

  void run_x() {
    XArgs a;
    a.A = A.p;
    a.Bf = Bf.p;
    a.Px = Px.p;
    a.va = vs.va;
    a.ax = ctx.ax;
    a.tw = tw_x();
    a.nkr = nkr;
    a.ny = ny;
    a.pf_vel = tune_pf_vel;
    a.pf_ahead = tune_pf_ahead_x;
    a.ablate = tune_ablate_x;
    a.first_wave = 2 * n_sm;
    {
      const int fx = 256 / (nx / 16);
      const long ctas = (long)((ny / 2) / fx) * nb;
      a.stagger = ctas >= 4L * a.first_wave ? tune_stagger_x : 0;
    }
    int vmode = (vs.va.kind == PTF_FLOW_SEPARABLE) ? 2 : (vs.va.ushift ? 1 : 0);
    if (vmode != 2 && tune_x_direct) vmode = tune_x_direct == 2 ? 4 : 3;
    if (vmode != 2 && (!vs.va.arr[0] || !vs.va.arr[1]))
      throw Error(PTF_EINVAL, "velocity fields have not been set (ptf_set_velocity / callback)");
    PTF_DISPATCH_N(nx, fused_launch_x, vmode, &a, nb, ctx.stream, n_sm);
    ++own_launches;
  }

  void stage(int mode, double la = 0, double lb = 0, int llast = 0) {
    run_x();
    run_y(true, true, mode, la, lb, llast);
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
  }

  void step_once(int64_t step_index) override {
    refresh_expr();
    if (!ab_valid) {  // A,B of the current sol are missing (fresh state): one prologue launch
      run_y(false, true, -1);
      ab_valid = true;
    }
    int variant = (ctx.st.base == PTF_STEPPER_AB3 && step_index < 3) ? 1 : 0;
    if (!ctx.d.use_graph) {
      enqueue_step(variant);
      PTF_CUDA(cudaGetLastError());
      return;
    }
    if (!graph_exec[variant]) {
      int64_t o0 = own_launches;
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
      own_launches = o0;
    }
    PTF_CUDA(cudaGraphLaunch(graph_exec[variant], ctx.stream));
    own_launches += per_step_own;
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
    k_diag_t<<<1184, 256, 0, ctx.stream>>>(s0.p, nkr, ny, nb, nx, out.p);
    ++own_launches;
    double h[2];
    double2 dc;
    PTF_CUDA(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaMemcpyAsync(&dc, s0.p, sizeof(dc), cudaMemcpyDeviceToHost, ctx.stream));
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    double N = (double)g.npts();
    double m = dc.x / N;
    double msq = h[0] / (N * N) / (double)nb;
    if (mean_c) *mean_c = m;
    if (var_c) *var_c = msq - m * m;
    if (max_abs_sol) *max_abs_sol = std::sqrt(h[1]);
  }

  // Average device time of one hot kernel measured over `reps` REAL steps (events around every launch of that kernel
  // on the step stream).  The solution is backed up and restored, the clock is untouched.
  float time_kernel(const char* kname, int reps) override {
    std::string k(kname ? kname : "");
    bool want_x = (k == "xkernel");
    if (!want_x && k != "ykernel") throw Error(PTF_EINVAL, "fused engine: unknown kernel '" + k + "' (xkernel|ykernel)");
    if (ctx.st.base != PTF_STEPPER_RK4) throw Error(PTF_EUNSUPPORTED, "kernel timing is implemented for RK4 steps");
    DevBuf<double2> backup;
    backup.alloc(nspec);
    PTF_CUDA(cudaMemcpyAsync(backup.p, s0.p, s0.bytes(), cudaMemcpyDeviceToDevice, ctx.stream));
    refresh_expr();
    if (!ab_valid) {
      run_y(false, true, -1);
      ab_valid = true;
    }
    const int modes[4] = {CM_RK4_S1, CM_RK4_S2, CM_RK4_S3, CM_RK4_S4};
    std::vector<cudaEvent_t> ev(2 * 4 * reps);
    for (auto& e : ev) PTF_CUDA(cudaEventCreate(&e));
    auto one_step = [&](int rep, bool record) {
      for (int s = 0; s < 4; ++s) {
        int idx = 2 * (4 * rep + s);
        if (record && want_x) PTF_CUDA(cudaEventRecord(ev[idx], ctx.stream));
        run_x();
        if (record && want_x) PTF_CUDA(cudaEventRecord(ev[idx + 1], ctx.stream));
        if (record && !want_x) PTF_CUDA(cudaEventRecord(ev[idx], ctx.stream));
        run_y(true, true, modes[s]);
        if (record && !want_x) PTF_CUDA(cudaEventRecord(ev[idx + 1], ctx.stream));
      }
    };
    one_step(0, false);  // warm
    for (int r = 0; r < reps; ++r) one_step(r, true);
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    double total = 0;
    for (int i = 0; i < 4 * reps; ++i) {
      float ms = 0;
      PTF_CUDA(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
      total += ms;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    PTF_CUDA(cudaMemcpyAsync(s0.p, backup.p, s0.bytes(), cudaMemcpyDeviceToDevice, ctx.stream));
    ab_valid = false;
    PTF_CUDA(cudaStreamSynchronize(ctx.stream));
    return (float)(total / (4.0 * reps));
  }

 private:
  Context& ctx;
  Geometry& g;
  int nx, ny, nkr, nb;
  size_t nspec;
  DevBuf<double2> s0, s1, s2, acc, n1, A, Bf, Px;
  DevBuf<double> cE, cE2, cZ, cA, cB, cG;
  DevBuf<char> work;
  size_t work_bytes = 0;
  TwiddleSet twx, twy_own;
  VelocityStore vs;
  ExprFlow ef;                 // PTF_FLOW_EXPR
  DevBuf<double> exprv[2];
  bool expr_dirty = true;
  cufftHandle plan_fwd = 0, plan_inv = 0;
  bool ab_valid = false;
  int tune_ablate_x = 0, tune_ablate_y = 0, tune_stagger_x = 0, tune_stagger_y = 0, n_sm = 148;
  int tune_pf_state = 0, tune_pf_vel = 0, tune_pf_ahead_y = 0, tune_pf_ahead_x = 0, tune_x_direct = 0;
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int64_t per_step_own = 0;
};

}  // namespace

bool fused_engine_supports(const Context& ctx, std::string* why) {
  auto no = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  if (ctx.g.ndim != 2) return no("only 2-D problems (1-D / 3-D run on the cuFFT engine)");
  if (!is_fused_size(ctx.g.nx) || !is_fused_size(ctx.g.ny)) return no("nx and ny must be powers of two in [64, 4096]");
  if (ctx.g.B > 65535) return no("batch too large for one launch");
  return true;
}

std::unique_ptr<Engine> make_fused_engine(Context& ctx) {
  std::string why;
  if (!fused_engine_supports(ctx, &why)) throw Error(PTF_EUNSUPPORTED, "fused engine: " + why);
  return std::unique_ptr<Engine>(new FusedEngine(ctx));
}

// Test hook: run `count` independent length-n complex transforms through the hand-written FFT core
// (dir = -1 / +1), or the TMEM pairing test (dir = 2: transforms along axis 0 of a row-major [n][count] matrix).
void selftest_fft(int n, int dir, int count, const double* in_host, double* out_host) {
  PTF_REQUIRE(is_fused_size(n), "selftest_fft: n must be a power of two in [64, 4096]");
  PTF_REQUIRE(dir == 1 || dir == -1 || dir == 2, "selftest_fft: dir must be +1, -1 or 2 (pair test)");
  PTF_REQUIRE(count > 0, "selftest_fft: count must be positive");
  PTF_REQUIRE(dir != 2 || count % 2 == 0, "pair test needs an even count");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(PTF_ENODEVICE, "no CUDA device visible");
  }
  TwiddleSet tw;
  tw.build(n, nullptr);
  Twiddles twd = tw.dev();
  DevBuf<double2> in, out;
  size_t total = (size_t)n * count;
  in.alloc(total);
  out.alloc(total);
  PTF_CUDA(cudaMemcpy(in.p, in_host, total * sizeof(double2), cudaMemcpyHostToDevice));
  PTF_DISPATCH_N(n, fused_selftest, dir, count, in.p, out.p, &twd);
  PTF_CUDA(cudaGetLastError());
  PTF_CUDA(cudaDeviceSynchronize());
  PTF_CUDA(cudaMemcpy(out_host, out.p, total * sizeof(double2), cudaMemcpyDeviceToHost));
}

}  // namespace ptf
