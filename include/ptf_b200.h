/* ptf_b200.h — C ABI of libptf_b200.so, the B200-native drop-in for the hot path of
 * PassiveTracerFlows.jl's TracerAdvectionDiffusion module.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes, returns an int32 status
 * (0 == PTF_OK) and never lets a C++ exception cross the boundary.  The library owns all
 * device memory, cuFFT plans, streams, graphs and the NCCL communicator; the caller owns the
 * host buffers, which only need to stay valid for the duration of the call.
 *
 * Reference interfaces replaced (paths relative to the reference checkout; TAD.jl =
 * src/traceradvectiondiffusion.jl; "FF" = FourierFlows.jl 0.10.5, an un-vendored dependency
 * pinned by Project.toml:24, whose pieces on this path are re-implemented here):
 *
 *   ptf_create             TracerAdvectionDiffusion.Problem(dev, flow; nx,Lx,...,κ,η,ι,dt,stepper)
 *                          TAD.jl:143-216, Problem(MQGprob; ...) TAD.jl:225-250, and the low-level
 *                          route ConstDiffSteadyFlowParams(κ,η,κh,nκh,u,v,grid)+Equation+FF.Problem
 *                          (test/test_traceradvectiondiffusion.jl:420-423); Equation -> L TAD.jl:502-566;
 *                          Vars TAD.jl:643-683; FF TimeStepper constructors; FF makefilter.
 *   ptf_set_velocity       steady u.(x,y) arrays of ConstDiffSteadyFlowParams TAD.jl:426-452
 *   ptf_set_velocity_callback   the u(x,y,t) closures of ConstDiffTimeVaryingFlowParams TAD.jl:268-348,
 *                          evaluated at clock.t once per step (TAD.jl:701,718,737)
 *   ptf_set_velocity_separable  same closures, for flows of the form sum_m a_m(t) X_m(x) Y_m(y) Z_m(z)
 *   ptf_set_velocity_expr  same closures, written as CUDA expressions and compiled at run time (NVRTC)
 *   ptf_set_layered_velocity    MQGprob.vars.u .+ MQGprob.params.U, MQGprob.vars.v  TAD.jl:795-796
 *   ptf_set_c              set_c!(prob, c)        TAD.jl:844-872
 *   ptf_get_c              updatevars!(prob) + read of prob.vars.c   TAD.jl:815-837
 *   ptf_get_sol/ptf_set_sol     prob.sol (read by FF Output / diagnostics, examples/onedim_gaussiandiffusion.jl:68-72)
 *   ptf_step               stepforward!(prob, nsteps)  (FF timesteppers.jl) which calls calcN! TAD.jl:695-804
 *   ptf_step_until         step_until!(prob, t)   (FF; used at TAD.jl:238)
 *   ptf_get_clock/ptf_set_clock/ptf_set_dt   prob.clock.{t,step,dt}
 *
 * Array layout at the boundary is Julia's: dense column-major (nx, ny, nz, nbatch) float64 for physical
 * fields and (nx/2+1, ny, nz, nbatch) interleaved complex128 for spectral ones — x fastest.  This is what
 * `pointer(A)` of the reference's own arrays gives, no transposes are needed (TAD.jl:646-665,678-679).
 *
 * Threading: a handle is not re-entrant (one in-flight call per handle).  Different handles may be driven from
 * different OS threads.  Every entry point selects the handle's CUDA device itself.
 */
#ifndef PTF_B200_H
#define PTF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTF_VERSION_MAJOR 0
#define PTF_VERSION_MINOR 1

/* status codes */
enum {
  PTF_OK = 0,
  PTF_EINVAL = 1,       /* bad argument / descriptor (maps to Julia ArgumentError, cf. TAD.jl:234) */
  PTF_ECUDA = 2,        /* CUDA runtime error */
  PTF_ECUFFT = 3,       /* cuFFT error */
  PTF_ENCCL = 4,        /* NCCL error */
  PTF_ENOMEM = 5,       /* device or host allocation failed */
  PTF_EUNSUPPORTED = 6, /* valid request this build cannot serve (e.g. fused engine for a non power-of-two grid) */
  PTF_ENODEVICE = 7     /* no CUDA device: there is NO CPU fallback */
};

/* FourierFlows stepper names (string kwarg `stepper`, TAD.jl:148,173,201,228) */
enum {
  PTF_STEPPER_FORWARD_EULER = 0,
  PTF_STEPPER_RK4 = 1,
  PTF_STEPPER_ETDRK4 = 2,
  PTF_STEPPER_LSRK54 = 3,
  PTF_STEPPER_AB3 = 4,
  PTF_STEPPER_FILTERED = 16 /* OR-ed flag: FilteredRK4 = PTF_STEPPER_RK4 | PTF_STEPPER_FILTERED, ... */
};

/* how the advecting velocity is supplied */
enum {
  PTF_FLOW_STEADY = 0,    /* arrays uploaded once with ptf_set_velocity (AbstractSteadyFlowParams) */
  PTF_FLOW_CALLBACK = 1,  /* host callback fills u,v,w at clock.t once per step (AbstractTimeVaryingFlowParams) */
  PTF_FLOW_SEPARABLE = 2, /* sum_m a_m(t) X_m(x) Y_m(y) Z_m(z); coefficient callback per step; zero HBM bytes */
  PTF_FLOW_LAYERED = 3,   /* per-layer u,v (+U(y,layer)) re-supplied by the caller between steps (MQG coupling) */
  PTF_FLOW_EXPR = 4       /* u,v,w given as CUDA expressions in (x,y,z,t), compiled at run time into the product kernel and
                             evaluated in registers at clock.t: zero HBM bytes, no per-step upload (cuFFT pipelines only) */
};

/* which implementation executes the step */
enum {
  PTF_ENGINE_AUTO = 0,  /* fused hand-written FFT pipeline when the grid qualifies, else the cuFFT pipeline */
  PTF_ENGINE_CUFFT = 1, /* cuFFT D2Z/Z2D + fused pointwise kernels: any even grid size */
  PTF_ENGINE_FUSED = 2  /* hand-written shared-memory FFTs with all pointwise work fused into their I/O: 2-D grids
                           64..4096 per axis, 3-D grids 64..1024 per axis (also slab-decomposed), 1-D <= 2048 */
};

/* how a multi-rank job is partitioned (one process per GPU; rank/nranks below) */
enum {
  PTF_DECOMP_NONE = 0,  /* single GPU, or independent replicas */
  PTF_DECOMP_BATCH = 1, /* ensemble members / layers sharded over ranks, no communication */
  PTF_DECOMP_SLAB = 2   /* last physical axis sharded; one all-to-all per transform: peer-to-peer over NVLink fused into
                           the z-column kernel (fused 3-D engine, CUDA IPC) or NCCL send/recv (cuFFT / 2-D slab engines) */
};

typedef struct ptf_handle ptf_handle;

/* Velocity callback: fill u[,v[,w]] (each n[0]*n[1]*n[2](*nbatch if per_batch) doubles, x fastest; unused
 * components are NULL) with the flow at time t.  Called on the calling thread, once per time step. */
typedef void (*ptf_velocity_fn)(void* user, double t, double* u, double* v, double* w);
/* Separable-flow coefficient callback: fill a[0..nterms) for component `comp` at time t. */
typedef void (*ptf_coeff_fn)(void* user, double t, int32_t comp, int32_t nterms, double* a);

typedef struct ptf_desc {
  uint32_t struct_size;      /* = sizeof(ptf_desc); set by ptf_desc_init */
  int32_t ndim;              /* 1, 2 or 3 */
  int64_t n[3];              /* nx, ny, nz (even) */
  double L[3];               /* Lx, Ly, Lz */
  int32_t nbatch;            /* layers or ensemble members (>= 1); transforms never cross this axis */
  int32_t stepper;           /* PTF_STEPPER_* [| PTF_STEPPER_FILTERED] */
  double kappa[3];           /* κ, η, ι */
  double kappa_h;            /* isotropic hyperdiffusivity κh */
  int32_t n_kappa_h;         /* its order nκh */
  int32_t dealias;           /* 0 = off (reference behaviour); 1 = dealias!(sol) at the top of every calcN */
  double aliased_fraction;   /* FF default 1/3 */
  double dt;
  int32_t nyquist_sign;      /* sign of the y/z Nyquist wavenumber: -1 = FF/fftfreq convention */
  int32_t flow_kind;         /* PTF_FLOW_* */
  int32_t velocity_per_batch;/* 0: one velocity field shared by all members; 1: (…, nbatch) fields */
  int32_t engine;            /* PTF_ENGINE_* */
  int32_t device;            /* CUDA device ordinal; -1 = current device */
  int32_t decomposition;     /* PTF_DECOMP_* */
  int32_t nranks;            /* processes in the job (1 = single GPU) */
  int32_t rank;              /* this process */
  uint8_t nccl_id[128];      /* ncclUniqueId from ptf_nccl_unique_id on rank 0, broadcast by the host side */
  double filter_order;       /* FF makefilter: 4 */
  double filter_inner_k;     /* 2/3 */
  double filter_outer_k;     /* 1 */
  double filter_tol;         /* 1e-15 */
  int32_t use_graph;         /* 1 = capture each step in a CUDA graph (default) */
  int32_t reserved[7];
} ptf_desc;

/* library */
int32_t ptf_version(int32_t* major, int32_t* minor);
int32_t ptf_device_count(int32_t* count);
const char* ptf_error_string(int32_t status);
const char* ptf_last_error(const ptf_handle* h); /* h == NULL: last error of a failed ptf_create on this thread */
int32_t ptf_nccl_unique_id(uint8_t id[128]);

/* construction */
int32_t ptf_desc_init(ptf_desc* d); /* reference defaults: ndim 2, n 128, L 2π, κ=η=ι 0.1, dt 0.01, RK4 */
int32_t ptf_create(const ptf_desc* d, ptf_handle** out);
int32_t ptf_destroy(ptf_handle* h);

/* local (this rank's) extents: physical and spectral element counts and the slab/batch offsets */
int32_t ptf_local_shape(const ptf_handle* h, int64_t phys_n[4], int64_t spec_n[4], int64_t phys_offset[4],
                        int64_t spec_offset[4]);

/* velocities */
int32_t ptf_set_velocity(ptf_handle* h, int32_t comp, const double* host, int64_t count);
int32_t ptf_set_velocity_callback(ptf_handle* h, ptf_velocity_fn fn, void* user);
int32_t ptf_set_velocity_separable(ptf_handle* h, int32_t comp, int32_t nterms, const double* xtab,
                                   const double* ytab, const double* ztab, const double* coeff0);
int32_t ptf_set_coeff_callback(ptf_handle* h, ptf_coeff_fn fn, void* user);
/* PTF_FLOW_EXPR: component `comp` as a C expression in the doubles x, y, z, t (CUDA device math: sin, cos, exp, ...; `pi`),
 * e.g. "(sin(z) + cos(y)) * (1 + 0.5*sin(t))".  The same closures the reference keeps in ConstDiffTimeVaryingFlowParams
 * (TAD.jl:268-348) and evaluates at clock.t on gridpoints(grid) (TAD.jl:715-718).  Compiled when the last component is
 * set; a compile error returns PTF_EINVAL with the compiler log in ptf_last_error. */
int32_t ptf_set_velocity_expr(ptf_handle* h, int32_t comp, const char* expr);
int32_t ptf_set_layered_velocity(ptf_handle* h, const double* u, const double* v, const double* U);

/* state */
int32_t ptf_set_c(ptf_handle* h, const double* c_host, int32_t replicate_over_batch);
int32_t ptf_get_c(ptf_handle* h, double* c_host);
int32_t ptf_set_sol(ptf_handle* h, const double* sol_host_interleaved);
int32_t ptf_get_sol(ptf_handle* h, double* sol_host_interleaved);
int32_t ptf_get_clock(const ptf_handle* h, double* t, int64_t* step, double* dt);
int32_t ptf_set_clock(ptf_handle* h, double t, int64_t step);
int32_t ptf_set_dt(ptf_handle* h, double dt);

/* stepping */
int32_t ptf_step(ptf_handle* h, int64_t nsteps);
int32_t ptf_step_until(ptf_handle* h, double t_stop);
int32_t ptf_step_timed(ptf_handle* h, int64_t nsteps, float* device_ms); /* CUDA-event time on the step stream */
int32_t ptf_sync(ptf_handle* h);

/* introspection for benchmarks / tests */
int32_t ptf_engine(const ptf_handle* h, int32_t* engine);              /* which engine was selected */
int32_t ptf_launch_count(const ptf_handle* h, int64_t* own_kernels, int64_t* library_calls);
int32_t ptf_kernel_timed(ptf_handle* h, const char* name, int32_t reps, float* avg_ms); /* time one hot kernel in place */
int32_t ptf_device_bytes(const ptf_handle* h, int64_t* bytes);

/* device-side diagnostics (callers' side of the path: FF Diagnostic-style scalars without a full read-back) */
int32_t ptf_diag(ptf_handle* h, double* mean_c, double* variance_c, double* max_abs_sol);

/* test hook: `count` independent length-n (power of two, 256..4096) complex transforms through the hand-written
 * shared-memory FFT core; dir = -1 forward, +1 unnormalised inverse; interleaved complex128 host buffers */
int32_t ptf_selftest_fft(int32_t n, int32_t dir, int32_t count, const double* in_host, double* out_host);

/* ------------------------------------------------------------------------------------------------------------------
 * MultiLayerQG flow solver: the flow that advects the tracer of Problem(MQGprob; ...) TAD.jl:225-250.
 * The solver itself is GeophysicalFlows.MultiLayerQG (un-vendored, pinned 0.16 by Project.toml:25); these entry points
 * replace the calls the reference makes into it:
 *   ptf_mqg_create        MultiLayerQG.Problem(nlayers, dev; nx, Lx, f₀, H, b, U, μ, β, dt, stepper, aliased_fraction)
 *                         examples/turbulent_advection-diffusion.jl:56-58
 *   ptf_mqg_set_q         MultiLayerQG.set_q!(MQGprob, q₀)                      examples/…:69
 *   ptf_mqg_step          stepforward!(params.MQGprob)                          examples/…:150
 *   ptf_mqg_step_until    step_until!(MQGprob, tracer_release_time)             TAD.jl:238
 *   ptf_mqg_updatevars    MultiLayerQG.updatevars!(MQGprob)                     TAD.jl:488, examples/…:151
 *   ptf_mqg_get_var       MQGprob.vars.{u,v,q,ψ} (u, v read at TAD.jl:795-796; ψ at examples/…:110-118)
 *   ptf_mqg_couple        ConstDiffTurbulentFlowParams(κ, η, tracer_release_time, MQGprob)   TAD.jl:485-491 — the
 *                         layered tracer problem reads the flow's u, v (+U) on the device from then on
 *   ptf_mqg_step_coupled  the example's loop body (examples/…:149-151), nsteps times, without host round trips
 * Arrays: (nx, ny, nlayers) column-major float64 / (nx/2+1, ny, nlayers) complex128, x fastest.
 * ------------------------------------------------------------------------------------------------------------------ */
#define PTF_MQG_MAX_LAYERS 4

typedef struct ptf_mqg_handle ptf_mqg_handle;

typedef struct ptf_mqg_desc {
  uint32_t struct_size;   /* = sizeof(ptf_mqg_desc); set by ptf_mqg_desc_init */
  int32_t nlayers;        /* 1 .. PTF_MQG_MAX_LAYERS */
  int64_t nx, ny;         /* even */
  double Lx, Ly;
  double f0, beta;        /* f₀, β */
  const double* H;        /* [nlayers] rest depths */
  const double* b;        /* [nlayers] Boussinesq buoyancies (nlayers >= 2) */
  const double* U;        /* imposed zonal flow: [nlayers], or [nlayers][ny] when U_is_profile; NULL = 0 */
  int32_t U_is_profile;
  int32_t n_nu;           /* hyperviscosity order nν (>= 1) */
  const double* eta;      /* [ny][nx] periodic topographic PV f₀h/H_n, or NULL */
  double topographic_pv_gradient[2];
  double mu, nu;          /* bottom drag μ, (hyper)viscosity ν */
  double dt;
  int32_t stepper;        /* PTF_STEPPER_* [| PTF_STEPPER_FILTERED] */
  int32_t device;         /* CUDA ordinal; -1 = current */
  double aliased_fraction;/* FF default 1/3; the reference's example uses 0 */
  int32_t use_graph;
  int32_t reserved[7];
} ptf_mqg_desc;

int32_t ptf_mqg_desc_init(ptf_mqg_desc* d); /* GeophysicalFlows defaults: 2 layers, 128², 2π, f₀ 1, β 0, RK4, dt 0.01 */
int32_t ptf_mqg_create(const ptf_mqg_desc* d, ptf_mqg_handle** out);
int32_t ptf_mqg_destroy(ptf_mqg_handle* h);
const char* ptf_mqg_last_error(const ptf_mqg_handle* h);
int32_t ptf_mqg_set_q(ptf_mqg_handle* h, const double* q_host);
int32_t ptf_mqg_set_psi(ptf_mqg_handle* h, const double* psi_host);
int32_t ptf_mqg_set_sol(ptf_mqg_handle* h, const double* sol_host_interleaved);
int32_t ptf_mqg_get_sol(ptf_mqg_handle* h, double* sol_host_interleaved);
int32_t ptf_mqg_updatevars(ptf_mqg_handle* h);
int32_t ptf_mqg_get_var(ptf_mqg_handle* h, int32_t which /* 0 u, 1 v, 2 q, 3 psi */, double* host);
int32_t ptf_mqg_get_background(ptf_mqg_handle* h, double* Qx_host, double* Qy_host);
int32_t ptf_mqg_step(ptf_mqg_handle* h, int64_t nsteps);
int32_t ptf_mqg_step_until(ptf_mqg_handle* h, double t_stop);
int32_t ptf_mqg_step_timed(ptf_mqg_handle* h, int64_t nsteps, int32_t with_updatevars, float* device_ms);
int32_t ptf_mqg_get_clock(const ptf_mqg_handle* h, double* t, int64_t* step, double* dt);
int32_t ptf_mqg_set_dt(ptf_mqg_handle* h, double dt);
int32_t ptf_mqg_launch_count(const ptf_mqg_handle* h, int64_t* own_kernels, int64_t* library_calls);
int32_t ptf_mqg_couple(ptf_mqg_handle* flow, ptf_handle* tracer);
int32_t ptf_mqg_forget_tracer(ptf_mqg_handle* flow, ptf_handle* tracer);
int32_t ptf_mqg_step_coupled(ptf_mqg_handle* flow, ptf_handle* tracer, int64_t nsteps, float* device_ms);

#ifdef __cplusplus
}
#endif
#endif /* PTF_B200_H */
