# PTFB200.jl — Julia-side glue for libptf_b200.so (the B200-native TracerAdvectionDiffusion hot path).
#
# Drop-in usage (same names as PassiveTracerFlows.TracerAdvectionDiffusion, new device type `B200`):
#
#     using PTFB200
#     prob = PTFB200.Problem(B200(), TwoDAdvectingFlow(; u, v, steadyflow=true); nx, Lx, κ, dt, stepper="RK4")
#     PTFB200.set_c!(prob, c0); stepforward!(prob, 25); PTFB200.updatevars!(prob); prob.vars.c
#
# UNTESTED HERE: this image has no Julia.  Every ccall below binds one symbol of include/ptf_b200.h; the Python
# ctypes binding (passivetracerflows.jl_b200/_capi.py) is the tested twin of this file.
module PTFB200

export B200, Problem, set_c!, updatevars!, stepforward!, step_until!

const LIB = get(ENV, "PTF_B200_LIB", "libptf_b200.so")

struct B200                      # <: FourierFlows.Device in the real package (device seam, test/runtests.jl:12)
  device::Int32
end
B200() = B200(-1)

# mirror of `struct ptf_desc` (include/ptf_b200.h) — field order and types must match
mutable struct PtfDesc
  struct_size::UInt32; ndim::Int32
  n::NTuple{3,Int64}; L::NTuple{3,Float64}
  nbatch::Int32; stepper::Int32
  kappa::NTuple{3,Float64}; kappa_h::Float64
  n_kappa_h::Int32; dealias::Int32
  aliased_fraction::Float64; dt::Float64
  nyquist_sign::Int32; flow_kind::Int32; velocity_per_batch::Int32; engine::Int32
  device::Int32; decomposition::Int32; nranks::Int32; rank::Int32
  nccl_id::NTuple{128,UInt8}
  filter_order::Float64; filter_inner_k::Float64; filter_outer_k::Float64; filter_tol::Float64
  use_graph::Int32; reserved::NTuple{7,Int32}
  PtfDesc() = new()
end

const STEPPERS = Dict("ForwardEuler"=>0, "RK4"=>1, "ETDRK4"=>2, "LSRK54"=>3, "AB3"=>4)
stepper_id(s) = startswith(s, "Filtered") ? (STEPPERS[s[9:end]] | 16) : STEPPERS[s]

function check(status::Int32, h=C_NULL)
  status == 0 && return
  msg = unsafe_string(ccall((:ptf_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
  status == 1 ? throw(ArgumentError(msg)) : error("libptf_b200 [status $status]: $msg")   # cf. TAD.jl:234
end

mutable struct Clock; dt::Float64; t::Float64; step::Int; end
mutable struct Vars; c::Array{Float64}; ch::Array{ComplexF64}; end

mutable struct Prob
  h::Ptr{Cvoid}; sol::Array{ComplexF64}; clock::Clock; vars::Vars
  n::NTuple{3,Int}; L::NTuple{3,Float64}; ndim::Int
  velocity                                   # keeps the @cfunction closure alive
  nbatch::Int                                # layers of an MQG-coupled problem (1 otherwise)
end

gridpoints1(n, L) = range(-L/2, step=L/n, length=n)

"Problem(dev::B200, flow; nx, Lx, ny, Ly, nz, Lz, κ, η, ι, dt, stepper) — TAD.jl:143-216"
function Problem(dev::B200, flow; nx=128, Lx=2π, ny=nx, Ly=Lx, nz=nx, Lz=Lx, κ=0.1, η=κ, ι=κ, dt=0.01,
                 stepper="RK4", κh=0.0, nκh=0, dealias=false)
  fns  = hasproperty(flow, :w) ? (flow.u, flow.v, flow.w) : hasproperty(flow, :v) ? (flow.u, flow.v) : (flow.u,)
  ndim = length(fns)
  d = PtfDesc()
  check(ccall((:ptf_desc_init, LIB), Int32, (Ref{PtfDesc},), d))
  d.ndim = ndim; d.n = (nx, ny, nz); d.L = (Lx, Ly, Lz); d.kappa = (κ, η, ι); d.kappa_h = κh; d.n_kappa_h = nκh
  d.dt = dt; d.stepper = stepper_id(stepper); d.dealias = dealias; d.device = dev.device
  d.flow_kind = flow.steadyflow ? 0 : 1       # PTF_FLOW_STEADY / PTF_FLOW_CALLBACK
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ptf_create, LIB), Int32, (Ref{PtfDesc}, Ref{Ptr{Cvoid}}), d, h))
  dims  = (nx, ny, nz)[1:ndim]
  sdims = (nx ÷ 2 + 1, dims[2:end]...)
  x = gridpoints1(nx, Lx); y = gridpoints1(ny, Ly); z = gridpoints1(nz, Lz)
  pts(i...) = ndim == 1 ? (x[i[1]],) : ndim == 2 ? (x[i[1]], y[i[2]]) : (x[i[1]], y[i[2]], z[i[3]])
  keep = nothing
  if flow.steadyflow                         # u.(x, y) evaluated once on gridpoints (TAD.jl:426-452)
    for (a, f) in enumerate(fns)
      U = [Float64(f(pts(Tuple(I)...)...)) for I in CartesianIndices(dims)]
      GC.@preserve U check(ccall((:ptf_set_velocity, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64),
                                 h[], a - 1, U, length(U)), h[])
    end
  else                                       # evaluated at clock.t once per step (TAD.jl:701,718,737)
    npts = prod(dims)
    function cb(user::Ptr{Cvoid}, t::Float64, pu::Ptr{Float64}, pv::Ptr{Float64}, pw::Ptr{Float64})::Cvoid
      for (f, p) in zip(fns, (pu, pv, pw))
        out = unsafe_wrap(Array, p, npts)
        for (j, I) in enumerate(CartesianIndices(dims)); out[j] = f(pts(Tuple(I)...)..., t); end
      end
      nothing
    end
    keep = @cfunction($cb, Cvoid, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}))
    check(ccall((:ptf_set_velocity_callback, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), h[], keep, C_NULL), h[])
  end
  p = Prob(h[], zeros(ComplexF64, sdims), Clock(dt, 0.0, 0), Vars(zeros(dims), zeros(ComplexF64, sdims)),
           (nx, ny, nz), (Lx, Ly, Lz), ndim, keep, 1)
  finalizer(q -> ccall((:ptf_destroy, LIB), Int32, (Ptr{Cvoid},), q.h), p)
  return p
end

"set_c!(prob, c) — TAD.jl:844-872; a 2-D c given to a layered problem is repeated over the layers (TAD.jl:865)"
function set_c!(p::Prob, c::Array{Float64})
  replicate = (p.nbatch > 1 && ndims(c) == p.ndim) ? 1 : 0
  GC.@preserve c check(ccall((:ptf_set_c, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32), p.h, c, replicate), p.h)
  updatevars!(p)
end

"updatevars!(prob) — TAD.jl:815-837: refreshes the host mirrors prob.vars.c and prob.sol"
function updatevars!(p::Prob)
  check(ccall((:ptf_get_c, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), p.h, p.vars.c), p.h)
  check(ccall((:ptf_get_sol, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}), p.h, p.sol), p.h)
  p.vars.ch = p.sol
  nothing
end

function sync_clock!(p::Prob)
  t = Ref(0.0); s = Ref(Int64(0)); dt = Ref(0.0)
  ccall((:ptf_get_clock, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Int64}, Ref{Float64}), p.h, t, s, dt)
  p.clock.t, p.clock.step, p.clock.dt = t[], s[], dt[]
end

"stepforward!(prob[, nsteps]) — FourierFlows timesteppers.jl"
function stepforward!(p::Prob, nsteps::Integer=1)
  check(ccall((:ptf_step, LIB), Int32, (Ptr{Cvoid}, Int64), p.h, nsteps), p.h)
  sync_clock!(p)
end

"step_until!(prob, stop_time) — FourierFlows (used at TAD.jl:238)"
function step_until!(p::Prob, stop_time)
  check(ccall((:ptf_step_until, LIB), Int32, (Ptr{Cvoid}, Float64), p.h, stop_time), p.h)
  sync_clock!(p)
end

"""
Time-varying flow given as CUDA expressions in x, y, z, t (PTF_FLOW_EXPR = 4 in `d.flow_kind` at creation): the closures of
ConstDiffTimeVaryingFlowParams (TAD.jl:268-348) evaluated at clock.t in registers, e.g.
`set_velocity_expr!(prob, 0, "(sin(z) + cos(y)) * (1 + 0.5*sin(t))")`.  A syntax error throws ArgumentError with the log.
"""
set_velocity_expr!(p::Prob, comp::Integer, expr::AbstractString) =
  check(ccall((:ptf_set_velocity_expr, LIB), Int32, (Ptr{Cvoid}, Int32, Cstring), p.h, comp, expr), p.h)

# ------------------------------------------------------------------------------------------------------------
# MultiLayerQG flow on the device (ptf_mqg_* in include/ptf_b200.h): the calls the reference makes into
# GeophysicalFlows.MultiLayerQG — examples/turbulent_advection-diffusion.jl:56-69,149-151; TAD.jl:225-250,488,795-796
# ------------------------------------------------------------------------------------------------------------
mutable struct PtfMqgDesc
  struct_size::UInt32; nlayers::Int32
  nx::Int64; ny::Int64
  Lx::Float64; Ly::Float64; f0::Float64; beta::Float64
  H::Ptr{Float64}; b::Ptr{Float64}; U::Ptr{Float64}
  U_is_profile::Int32; n_nu::Int32
  eta::Ptr{Float64}
  topographic_pv_gradient::NTuple{2,Float64}
  mu::Float64; nu::Float64; dt::Float64
  stepper::Int32; device::Int32
  aliased_fraction::Float64
  use_graph::Int32; reserved::NTuple{7,Int32}
  PtfMqgDesc() = new()
end

mutable struct MQGProb
  h::Ptr{Cvoid}; nlayers::Int; nx::Int; ny::Int; clock::Clock; U::Vector{Float64}
end

checkmqg(status::Int32, h=C_NULL) = status == 0 ? nothing :
  (msg = unsafe_string(ccall((:ptf_mqg_last_error, LIB), Cstring, (Ptr{Cvoid},), h));
   status == 1 ? throw(ArgumentError(msg)) : error("libptf_b200 status $status: $msg"))

"MultiLayerQG.Problem(nlayers, B200(); nx, Lx, f₀, H, b, U, μ, β, dt, stepper, aliased_fraction) — examples/…:56-58"
function MultiLayerQGProblem(nlayers::Int, dev::B200; nx=128, ny=nx, Lx=2π, Ly=Lx, f₀=1.0, β=0.0, U=zeros(nlayers),
                             H=fill(1/nlayers, nlayers), b=-(1 .+ (0:nlayers-1)/nlayers), μ=0.0, ν=0.0, nν=1,
                             dt=0.01, stepper="RK4", aliased_fraction=1/3)
  d = PtfMqgDesc(); ccall((:ptf_mqg_desc_init, LIB), Int32, (Ref{PtfMqgDesc},), d)
  Hv, bv, Uv = Float64.(collect(H)), Float64.(collect(b)), Float64.(collect(U))
  d.nlayers, d.nx, d.ny, d.Lx, d.Ly, d.f0, d.beta = nlayers, nx, ny, Lx, Ly, f₀, β
  d.mu, d.nu, d.n_nu, d.dt, d.aliased_fraction, d.device = μ, ν, nν, dt, aliased_fraction, dev.device
  filt = startswith(stepper, "Filtered")
  d.stepper = STEPPERS[filt ? stepper[9:end] : stepper] | (filt ? 16 : 0)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve Hv bv Uv begin
    d.H, d.b, d.U = pointer(Hv), pointer(bv), pointer(Uv)       # copied by the library before ptf_mqg_create returns
    checkmqg(ccall((:ptf_mqg_create, LIB), Int32, (Ref{PtfMqgDesc}, Ref{Ptr{Cvoid}}), d, h))
  end
  p = MQGProb(h[], nlayers, nx, ny, Clock(dt, 0.0, 0), Uv)
  finalizer(q -> ccall((:ptf_mqg_destroy, LIB), Int32, (Ptr{Cvoid},), q.h), p)
  return p
end

"MultiLayerQG.set_q!(MQGprob, q) — examples/…:69"
set_q!(p::MQGProb, q::Array{Float64,3}) =
  GC.@preserve q checkmqg(ccall((:ptf_mqg_set_q, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), p.h, q), p.h)

"MultiLayerQG.updatevars!(MQGprob) — TAD.jl:488, examples/…:151 (device side; read fields with mqg_var)"
updatevars!(p::MQGProb) = checkmqg(ccall((:ptf_mqg_updatevars, LIB), Int32, (Ptr{Cvoid},), p.h), p.h)

"MQGprob.vars.u / .v / .q / .ψ as of the last updatevars! (which = 0, 1, 2, 3)"
function mqg_var(p::MQGProb, which::Integer)
  out = Array{Float64}(undef, p.nx, p.ny, p.nlayers)
  checkmqg(ccall((:ptf_mqg_get_var, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), p.h, which, out), p.h)
  return out
end

function sync_clock!(p::MQGProb)
  t = Ref(0.0); s = Ref(Int64(0)); dt = Ref(0.0)
  ccall((:ptf_mqg_get_clock, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Int64}, Ref{Float64}), p.h, t, s, dt)
  p.clock.t, p.clock.step, p.clock.dt = t[], s[], dt[]
end

"stepforward!(MQGprob[, nsteps]) — examples/…:150"
function stepforward!(p::MQGProb, nsteps::Integer=1)
  checkmqg(ccall((:ptf_mqg_step, LIB), Int32, (Ptr{Cvoid}, Int64), p.h, nsteps), p.h); sync_clock!(p)
end

"step_until!(MQGprob, t) — TAD.jl:238"
function step_until!(p::MQGProb, t)
  checkmqg(ccall((:ptf_mqg_step_until, LIB), Int32, (Ptr{Cvoid}, Float64), p.h, t), p.h); sync_clock!(p)
end

"""
Problem(MQGprob; κ, η, stepper, tracer_release_time) — TAD.jl:225-250: the layered tracer problem on the flow's grid and
dt (flow_kind = PTF_FLOW_LAYERED, nbatch = nlayers, one velocity field per layer), coupled to the device-resident flow:
its calcN! reads MQGprob.vars.u .+ params.U and vars.v (TAD.jl:795-796) directly from the flow solver's device buffers.
"""
function Problem(flow::MQGProb; κ=0.1, η=κ, stepper="FilteredRK4", tracer_release_time=0, Lx=2π, Ly=Lx, device=-1)
  d = PtfDesc()
  check(ccall((:ptf_desc_init, LIB), Int32, (Ref{PtfDesc},), d))
  d.ndim = 2; d.n = (flow.nx, flow.ny, 1); d.L = (Lx, Ly, 1.0); d.kappa = (κ, η, κ)
  d.dt = flow.clock.dt; d.stepper = stepper_id(stepper); d.device = device
  d.flow_kind = 3                                # PTF_FLOW_LAYERED
  d.nbatch = flow.nlayers; d.velocity_per_batch = 1
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ptf_create, LIB), Int32, (Ref{PtfDesc}, Ref{Ptr{Cvoid}}), d, h))
  dims, sdims = (flow.nx, flow.ny, flow.nlayers), (flow.nx ÷ 2 + 1, flow.ny, flow.nlayers)
  p = Prob(h[], zeros(ComplexF64, sdims), Clock(flow.clock.dt, 0.0, 0), Vars(zeros(dims), zeros(ComplexF64, sdims)),
           (flow.nx, flow.ny, 1), (Lx, Ly, 1.0), 2, nothing, flow.nlayers)
  finalizer(q -> ccall((:ptf_destroy, LIB), Int32, (Ptr{Cvoid},), q.h), p)
  couple!(p, flow; tracer_release_time=tracer_release_time)
  return p
end

"couples an existing layered tracer problem to the flow (what Problem(MQGprob; …) does after creating it)"
function couple!(tracer::Prob, flow::MQGProb; tracer_release_time=0)
  tracer_release_time < 0 && throw(ArgumentError("tracer_release_time must be non-negative!"))   # TAD.jl:234
  tracer_release_time > 0 && step_until!(flow, tracer_release_time)                                # TAD.jl:236-239
  checkmqg(ccall((:ptf_mqg_couple, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), flow.h, tracer.h), flow.h)
end

"the loop body of examples/…:149-151, nsteps times, without host round trips; returns device milliseconds"
function step_coupled!(tracer::Prob, flow::MQGProb, nsteps::Integer=1)
  ms = Ref(Float32(0))
  checkmqg(ccall((:ptf_mqg_step_coupled, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Float32}),
                 flow.h, tracer.h, nsteps, ms), flow.h)
  sync_clock!(tracer); sync_clock!(flow)
  return ms[]
end

end # module
