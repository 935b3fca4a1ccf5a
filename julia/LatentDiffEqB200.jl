# LatentDiffEqB200.jl -- Julia glue a LatentDiffEq.jl maintainer would add to route `diffeq_layer` to libldeq.so.
#
# WRITTEN TO SPEC, NOT EXECUTABLE IN THE BUILD IMAGE (no Julia there).  It binds include/ldeq.h with `ccall` and
# adds more specific `diffeq_layer` methods for CuArray inputs plus `ChainRulesCore.rrule`s, leaving the
# LatentDiffEqModel / GOKU / LatentODE constructors, the layer API and the diffeq struct fields untouched
# (reference src/models/GOKU.jl:98-130, src/models/LatentODE.jl:61-78, src/models/LatentDiffEqModel.jl:101-113).
module LatentDiffEqB200

using CUDA, ChainRulesCore, Flux, Functors
import LatentDiffEq
import LatentDiffEq: diffeq_layer, transform_after_diffeq, apply_pattern_extractor, Encoder, Decoder, GOKU, LatentODE

const libldeq = get(ENV, "LDEQ_LIB", "libldeq.so")

# mirror of `struct ldeq_opts` (include/ldeq.h); field order and types must match
Base.@kwdef mutable struct Opts
    abstol::Cdouble = 1e-6
    reltol::Cdouble = 1e-3
    adaptive::Int32 = 1
    controller_pow::Int32 = 0
    dt::Cdouble = 0.0
    dtmax::Cdouble = 0.0
    dtmin::Cdouble = 0.0
    maxiters::Int64 = 1_000_000
    gamma::Cdouble = 0.9
    qmin::Cdouble = 0.2
    qmax::Cdouble = 10.0
    beta1::Cdouble = 7 / 50
    beta2::Cdouble = 2 / 25
    qoldinit::Cdouble = 1e-4
    qsteady_min::Cdouble = 1.0
    qsteady_max::Cdouble = 1.0
    tape_steps::Int32 = 0
    norm_mode::Int32 = 0
    mlp_math::Int32 = 0
    sensealg::Int32 = 1   # 1 = LDEQ_SENSE_FORWARD_DUAL, the reference's dual-number re-solves (default); 0 = discrete adjoint
    solver::Int32 = 0     # ldeq_solver: 0 = Tsit5 (the reference's structs), 1 = DP5, 2 = BS3, 3 = RK4 (fixed step only)
    reserved_::Int32 = 0
end

# explicit opt-in sensealg of this glue (not a SciMLSensitivity type): the discrete adjoint of the taped accepted steps
struct DiscreteAdjoint end

# the diffeq struct's `sensealg` and `solver` fields (pendulum.jl:11,58; read at GOKU.jl:106-107) -> ldeq_opts codes.
# Anything libldeq does not implement is an error, never silently replaced.
function sensealg_code(s)
    s isa DiscreteAdjoint && return Int32(0)
    nameof(typeof(s)) === :ForwardDiffSensitivity && return Int32(1)   # SciMLSensitivity / DiffEqSensitivity, whichever is loaded
    nameof(typeof(s)) === :InterpolatingAdjoint && return Int32(2)     # LatentODE path only (ldeq_mlp_*): NeuralODE's default
    error("sensealg $(typeof(s)) not supported by libldeq: use ForwardDiffSensitivity() (the reference's) or LatentDiffEqB200.DiscreteAdjoint()")
end
const SOLVER_CODES = Dict(:Tsit5 => Int32(0), :DP5 => Int32(1), :BS3 => Int32(2), :RK4 => Int32(3))
solver_code(s) = get(SOLVER_CODES, nameof(typeof(s))) do
    error("solver $(typeof(s)) not supported by libldeq: Tsit5(), DP5(), BS3() and RK4() (fixed step) are built")
end

# the `kwargs` field of the diffeq struct is splatted into `solve` by the reference (GOKU.jl:108,121)
function Opts(kwargs::Union{NamedTuple,Base.Pairs,Dict})
    o = Opts()
    for (k, v) in pairs(kwargs)
        k === :saveat && continue
        hasproperty(o, k) || error("solver option $k is not supported by libldeq")
        setproperty!(o, k, convert(fieldtype(Opts, k), v))
    end
    return o
end

# everything the reference reads from a GOKU diffeq struct (GOKU.jl:105-108): kwargs, sensealg, solver.  The solver comes
# first: the PI-controller exponents OrdinaryDiffEq defaults to depend on the algorithm (ldeq_opts_default_solver), and
# the struct's kwargs then override them like they override OrdinaryDiffEq's defaults in `solve`.
function Opts(diffeq)
    base = Ref(Opts())
    code = hasproperty(diffeq, :solver) ? solver_code(diffeq.solver) : Int32(0)
    ccall((:ldeq_opts_default_solver, libldeq), Cint, (Ptr{Opts}, Cint), base, code) == 0 || error("ldeq_opts_default_solver($code)")
    o = base[]
    for (k, v) in pairs(diffeq.kwargs)
        k === :saveat && continue
        hasproperty(o, k) || error("solver option $k is not supported by libldeq")
        setproperty!(o, k, convert(fieldtype(Opts, k), v))
    end
    hasproperty(diffeq, :sensealg) && (o.sensealg = sensealg_code(diffeq.sensealg))
    return o
end

check(h, rc) = rc == 0 || error("libldeq error $rc: " * unsafe_string(ccall((:ldeq_last_error, libldeq), Cstring, (Ptr{Cvoid},), h)))

const HANDLES = Dict{Int,Ptr{Cvoid}}()
function handle()
    dev = CUDA.deviceid(CUDA.device())
    get!(HANDLES, dev) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ldeq_create, libldeq), Cint, (Ptr{Ptr{Cvoid}}, Cint), h, dev)
        rc == 0 || error("ldeq_create failed ($rc): no usable CUDA device")
        h[]
    end
end

dtype_code(::Type{Float32}) = Cint(0)
dtype_code(::Type{Float64}) = Cint(1)

# Right-hand side of a diffeq struct.  Built-ins: define `rhs_kind(::MyPendulum) = 0` (Pendulum) or `1` (Pendulum_friction).
# Any other `f!`: define `rhs_source(::MyDiffEq)` returning CUDA C for
#     template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t)
# (compiled once per handle by NVRTC through ldeq_rhs_from_source; z_dim / p_dim come from `prob.u0` / `prob.p`,
# exactly what GOKU.jl:207-208 reads).
rhs_kind(diffeq) = nothing
rhs_source(diffeq) = nothing

# one ldeq_rhs per (handle, diffeq type): created on first use, released by `release!` (or process exit)
const RHS_CACHE = Dict{Tuple{Ptr{Cvoid},DataType},Ptr{Cvoid}}()
function rhs_object(h, diffeq)
    get!(RHS_CACHE, (h, typeof(diffeq))) do
        r = Ref{Ptr{Cvoid}}(C_NULL)
        if rhs_kind(diffeq) !== nothing
            check(h, ccall((:ldeq_rhs_builtin, libldeq), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), h, rhs_kind(diffeq), r))
        elseif rhs_source(diffeq) !== nothing
            check(h, ccall((:ldeq_rhs_from_source, libldeq), Cint, (Ptr{Cvoid}, Cstring, Cint, Cint, Ptr{Ptr{Cvoid}}),
                           h, rhs_source(diffeq), length(diffeq.prob.u0), length(diffeq.prob.p), r))
        else
            error("define LatentDiffEqB200.rhs_kind(::$(typeof(diffeq))) or LatentDiffEqB200.rhs_source(::$(typeof(diffeq)))")
        end
        r[]
    end
end

# free every cached right-hand side and handle (call before unloading the library / at the end of a script)
function release!()
    for ((h, _), r) in RHS_CACHE
        ccall((:ldeq_rhs_free, libldeq), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}), h, r)
    end
    empty!(RHS_CACHE)
    for (_, h) in HANDLES
        ccall((:ldeq_destroy, libldeq), Cvoid, (Ptr{Cvoid},), h)
    end
    empty!(HANDLES)
end
atexit(release!)

mutable struct Tape
    ptr::Ptr{Cvoid}
    mlp::Bool
end

# ---- GOKU: B independent solves (replaces GOKU.jl:101-128 for CuArray inputs) -------------------------------
function solve_fwd(diffeq, ẑ₀::CuMatrix{T}, θ̂::CuMatrix{T}, t; tape::Bool) where {T}
    h = handle()
    z, B = size(ẑ₀)
    tt = collect(Float64, t)                           # the reference's t is a Float64 range
    ẑ = CUDA.zeros(T, z, B, length(tt))                # (z, B, T): what permutedims(.,[1,3,2]) yields, GOKU.jl:125
    opts = Ref(Opts(diffeq))                           # kwargs + sensealg + solver
    tp = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ldeq_solve_fwd, libldeq), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Ref{Opts}, CuPtr{Cvoid},
                CuPtr{Int32}, CuPtr{Int32}, CuPtr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
               h, rhs_object(h, diffeq), dtype_code(T), ẑ₀, θ̂, tt, B, length(tt), opts, ẑ,
               CU_NULL, CU_NULL, CU_NULL, tape ? tp : C_NULL, CUDA.stream().handle)
    check(h, rc)
    return ẑ, Tape(tp[], false)
end

function diffeq_layer(decoder::Decoder{M}, l̂::Tuple{<:CuMatrix,<:CuMatrix}, t) where {M<:GOKU}
    ẑ, _ = solve_fwd(decoder.diffeq, l̂[1], l̂[2], t; tape = false)
    return transform_after_diffeq(ẑ, decoder.diffeq)
end

function ChainRulesCore.rrule(::typeof(diffeq_layer), decoder::Decoder{M}, l̂::Tuple{<:CuMatrix,<:CuMatrix}, t) where {M<:GOKU}
    ẑ₀, θ̂ = l̂
    ẑ, tape = solve_fwd(decoder.diffeq, ẑ₀, θ̂, t; tape = true)
    function pullback(Δ)
        h = handle()
        dẑ₀, dθ̂ = similar(ẑ₀), similar(θ̂)
        check(h, ccall((:ldeq_solve_bwd, libldeq), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
                       h, tape.ptr, CuArray(unthunk(Δ)), dẑ₀, dθ̂, CUDA.stream().handle))
        ccall((:ldeq_tape_free, libldeq), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), h, tape.ptr, CUDA.stream().handle)
        return NoTangent(), NoTangent(), Tangent{typeof(l̂)}(dẑ₀, dθ̂), NoTangent()
    end
    # NB: a non-identity transform_after_diffeq must be composed by the caller's AD (it is plain Julia code)
    return transform_after_diffeq(ẑ, decoder.diffeq), pullback
end

# ---- LatentODE: one solve on the (D,B) matrix state (replaces LatentODE.jl:70-72 for CuArray inputs) ---------
function mlp_dims(dudt::Flux.Chain)
    d = Int32[size(dudt[1].weight, 2)]
    for l in dudt
        push!(d, size(l.weight, 1))
    end
    d
end

function diffeq_layer(decoder::Decoder{LatentODE}, ẑ₀::CuMatrix{T}, t) where {T}
    ẑ, _ = mlp_fwd(decoder.diffeq, ẑ₀, t; tape = false)
    return transform_after_diffeq(ẑ, decoder.diffeq)
end

function mlp_fwd(diffeq, ẑ₀::CuMatrix{T}, t; tape::Bool) where {T}
    h = handle()
    u0 = diffeq.augment_dim == 0 ? ẑ₀ : vcat(ẑ₀, CUDA.zeros(T, diffeq.augment_dim, size(ẑ₀, 2)))   # AugmentedNDELayer
    D, B = size(u0)
    p, _ = Flux.destructure(diffeq.dudt)              # per layer vec(W) column-major, then b (DiffEqFlux order)
    dims = mlp_dims(diffeq.dudt)
    tt = collect(Float64, t)
    ẑ = CUDA.zeros(T, D, B, length(tt))
    opts = Ref(Opts(diffeq))                          # kwargs + solver (a NODE struct has no sensealg field, nODE.jl:3-12)
    # NeuralODE's default sensitivity algorithm is InterpolatingAdjoint (DiffEqFlux 1.52): LDEQ_SENSE_INTERPOLATING_ADJOINT
    # whenever the solve runs in the reference's configuration (batch-global norm, exact arithmetic); the per-trajectory
    # norm and the bf16x3 tensor-core path are performance deviations whose reverse pass is the discrete adjoint
    if !hasproperty(diffeq, :sensealg)
        opts[].sensealg = (opts[].norm_mode == 0 && opts[].mlp_math == 0) ? Int32(2) : Int32(0)
    end
    tp = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ldeq_mlp_solve_fwd, libldeq), Cint,
               (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cdouble}, Cint, Cint, Ref{Opts}, CuPtr{Cvoid},
                CuPtr{Int32}, CuPtr{Int32}, CuPtr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
               h, dtype_code(T), u0, CuArray(p), dims, length(dims) - 1, tt, B, length(tt), opts, ẑ,
               CU_NULL, CU_NULL, CU_NULL, tape ? tp : C_NULL, CUDA.stream().handle)
    check(h, rc)
    return ẑ, Tape(tp[], true)
end

# reverse pass of the LatentODE method (replaces DiffEqFlux's InterpolatingAdjoint pullback, LatentODE.jl:70-72 under Zygote;
# the tape remembers the sensealg of the forward call: the continuous adjoint kernel or the discrete adjoint of the taped steps):
# ldeq_mlp_solve_bwd -> (dẑ₀ (D,B), dparams_flat in Flux.destructure order), the latter restructured into a tangent of `dudt`
function ChainRulesCore.rrule(::typeof(diffeq_layer), decoder::Decoder{LatentODE}, ẑ₀::CuMatrix{T}, t) where {T}
    diffeq = decoder.diffeq
    ẑ, tape = mlp_fwd(diffeq, ẑ₀, t; tape = true)
    p, re = Flux.destructure(diffeq.dudt)
    function pullback(Δ)
        h = handle()
        Δc = CuArray{T}(unthunk(Δ))                       # (D+aug, B, T), same layout as ẑ
        du0 = CUDA.zeros(T, size(ẑ, 1), size(ẑ, 2))
        dp = CUDA.zeros(T, length(p))
        check(h, ccall((:ldeq_mlp_solve_bwd, libldeq), Cint,
                       (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
                       h, tape.ptr, Δc, du0, dp, CUDA.stream().handle))
        ccall((:ldeq_mlp_tape_free, libldeq), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), h, tape.ptr, CUDA.stream().handle)
        dẑ₀ = diffeq.augment_dim == 0 ? du0 : du0[1:size(ẑ₀, 1), :]   # the zero-padded rows carry no input gradient
        ddudt = re(Array(dp))                              # a Chain whose arrays are the gradients (structural tangent)
        ddecoder = Tangent{typeof(decoder)}(; diffeq = Tangent{typeof(diffeq)}(; dudt = ddudt))
        return NoTangent(), ddecoder, dẑ₀, NoTangent()
    end
    return transform_after_diffeq(ẑ, diffeq), pullback
end

# ---- the host-array path the reference actually takes (GOKU.jl:102-103: `cpu(...)`, then `solve` on the CPU) ---------
# Forward: ldeq_solve_fwd_host; pullback: ldeq_solve_bwd_host (host cotangent in, host gradients out).  The library cuts
# the batch into column slabs and overlaps the copies with the kernels on its own streams.
function solve_fwd_host(diffeq, ẑ₀::Matrix{T}, θ̂::Matrix{T}, t; tape::Bool) where {T<:Union{Float32,Float64}}
    h = handle()
    z, B = size(ẑ₀)
    tt = collect(Float64, t)
    ẑ = Array{T}(undef, z, B, length(tt))
    tp = Ref{Ptr{Cvoid}}(C_NULL)
    check(h, ccall((:ldeq_solve_fwd_host, libldeq), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Ref{Opts}, Ptr{Cvoid},
                    Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
                   h, rhs_object(h, diffeq), dtype_code(T), ẑ₀, θ̂, tt, B, length(tt), Ref(Opts(diffeq)), ẑ,
                   C_NULL, C_NULL, C_NULL, tape ? tp : C_NULL, CUDA.stream().handle))
    return ẑ, Tape(tp[], false)
end

function diffeq_layer(decoder::Decoder{M}, l̂::Tuple{Matrix{T},Matrix{T}}, t) where {M<:GOKU,T<:Union{Float32,Float64}}
    ẑ, _ = solve_fwd_host(decoder.diffeq, l̂[1], l̂[2], t; tape = false)
    return transform_after_diffeq(ẑ, decoder.diffeq)
end

function ChainRulesCore.rrule(::typeof(diffeq_layer), decoder::Decoder{M}, l̂::Tuple{Matrix{T},Matrix{T}}, t) where {M<:GOKU,T<:Union{Float32,Float64}}
    ẑ₀, θ̂ = l̂
    ẑ, tape = solve_fwd_host(decoder.diffeq, ẑ₀, θ̂, t; tape = true)
    function pullback(Δ)
        h = handle()
        Δh = convert(Array{T,3}, unthunk(Δ))
        dẑ₀, dθ̂ = similar(ẑ₀), similar(θ̂)
        check(h, ccall((:ldeq_solve_bwd_host, libldeq), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                       h, tape.ptr, Δh, dẑ₀, dθ̂, CUDA.stream().handle))
        ccall((:ldeq_tape_free, libldeq), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), h, tape.ptr, CUDA.stream().handle)
        return NoTangent(), NoTangent(), Tangent{typeof(l̂)}(dẑ₀, dθ̂), NoTangent()
    end
    return transform_after_diffeq(ẑ, decoder.diffeq), pullback
end

# ---- the steps either side of the solve that libldeq also provides (GOKU.jl:155-173, model_train.jl:225-238, :138) ------
# z̃ = μ + ε ⊙ exp(logσ²/2) with ε drawn on the device (the reference draws it on the host and uploads it, GOKU.jl:169-170)
function sample_device(μ::CuMatrix{Float32}, logσ²::CuMatrix{Float32}; seed::Integer = rand(UInt64), offset::Integer = 0)
    h = handle()
    z̃, ε = similar(μ), similar(μ)
    check(h, ccall((:ldeq_sample, libldeq), Cint,
                   (Ptr{Cvoid}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, Int64, UInt64, UInt64, Ptr{Cvoid}),
                   h, μ, logσ², z̃, ε, length(μ), seed, offset, CUDA.stream().handle))
    return z̃, ε
end

# ---- apply_pattern_extractor (GOKU.jl:30-49): the three recurrent stacks as persistent kernels ----------------------------
# fe_out is the (F,B,T) CuArray the feature extractor returns.  Flux.destructure of a Chain(RNN, RNN) / Chain(LSTM, LSTM)
# is exactly the flat layout ldeq_pattern_extractor_* read (per layer Wi column-major, Wh, b, state0 [h0, c0]); `re` of the
# flat gradients rebuilds the parameter cotangents.  Built for rnn_output_dim = 16 and rnn_input_dim in {16, 32, 64}.
function pattern_extractor_fwd(fe_out::CuArray{Float32,3}, p_rnn::CuVector{Float32}, p_f::CuVector{Float32}, p_b::CuVector{Float32}; tape::Bool)
    h = handle()
    F, B, T = size(fe_out)
    z0o, θo = CUDA.zeros(Float32, 16, B), CUDA.zeros(Float32, 32, B)
    tp = Ref{Ptr{Cvoid}}(C_NULL)
    check(h, ccall((:ldeq_pattern_extractor_fwd, libldeq), Cint,
                   (Ptr{Cvoid}, CuPtr{Cfloat}, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat},
                    Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
                   h, fe_out, B, T, F, 16, p_rnn, p_f, p_b, z0o, θo, tape ? tp : C_NULL, CUDA.stream().handle))
    return z0o, θo, tp[]
end

function apply_pattern_extractor(encoder::Encoder{M}, fe_out::CuArray{Float32,3}) where {M<:GOKU}
    pe_z₀, pe_f, pe_b = encoder.pattern_extractor
    z0o, θo, _ = pattern_extractor_fwd(fe_out, Flux.destructure(pe_z₀)[1], Flux.destructure(pe_f)[1], Flux.destructure(pe_b)[1]; tape = false)
    return z0o, θo
end

function ChainRulesCore.rrule(::typeof(apply_pattern_extractor), encoder::Encoder{M}, fe_out::CuArray{Float32,3}) where {M<:GOKU}
    h = handle()
    pe_z₀, pe_f, pe_b = encoder.pattern_extractor
    (p_rnn, re_rnn), (p_f, re_f), (p_b, re_b) = Flux.destructure(pe_z₀), Flux.destructure(pe_f), Flux.destructure(pe_b)
    z0o, θo, tp = pattern_extractor_fwd(fe_out, p_rnn, p_f, p_b; tape = true)
    function pullback((Δz0, Δθ))
        dx, g_rnn, g_f, g_b = similar(fe_out), similar(p_rnn), similar(p_f), similar(p_b)
        check(h, ccall((:ldeq_pattern_extractor_bwd, libldeq), Cint,
                       (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat},
                        CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cvoid}),
                       h, tp, fe_out, p_rnn, p_f, p_b, CuArray{Float32}(unthunk(Δz0)), CuArray{Float32}(unthunk(Δθ)), dx, g_rnn, g_f, g_b,
                       CUDA.stream().handle))
        ccall((:ldeq_pe_tape_free, libldeq), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), h, tp, CUDA.stream().handle)
        # parameter cotangents as structural tangents of the three Chains (Functors.fmapstructure of the rebuilt gradient models)
        d_pe = map((re, g) -> Functors.fmapstructure(identity, re(g)), (re_rnn, re_f, re_b), (g_rnn, g_f, g_b))
        return NoTangent(), Tangent{typeof(encoder)}(; pattern_extractor = d_pe), dx
    end
    return (z0o, θo), pullback
end

# loss_batch's reduction + its gradient in one pass: x, x̂ (P,B,T); μs / logσ²s tuples of (d,B) heads
function elbo_fwd_bwd(x::CuArray{Float32,3}, x̂::CuArray{Float32,3}, μs::Tuple, logσ²s::Tuple, β::Real; grad_scale::Real = 1)
    h = handle()
    P, B, T = size(x)
    nh = length(μs)
    loss = CUDA.zeros(Float32, 3)
    dx̂ = similar(x̂)
    dμs, dlvs = map(similar, μs), map(similar, logσ²s)
    ptrs(xs) = Ptr{Cvoid}[reinterpret(Ptr{Cvoid}, pointer(a)) for a in xs]
    check(h, ccall((:ldeq_elbo_fwd_bwd, libldeq), Cint,
                   (Ptr{Cvoid}, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Cint, Cfloat, Cint, Cint, Cint,
                    Cfloat, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
                   h, x, x̂, ptrs(μs), ptrs(logσ²s), Int32[size(m, 1) for m in μs], nh, β, B, T, P, grad_scale, loss, dx̂,
                   ptrs(dμs), ptrs(dlvs), CUDA.stream().handle))
    return loss, dx̂, dμs, dlvs
end

# the same with the reconstructor's sigmoid output activation folded into the kernel: `a` are the pre-activations of its last
# Dense layer (GOKU.jl:265-268); returns the gradient with respect to `a`
function elbo_logits_fwd_bwd(x::CuArray{Float32,3}, a::CuArray{Float32,3}, μs::Tuple, logσ²s::Tuple, β::Real; grad_scale::Real = 1)
    h = handle()
    P, B, T = size(x)
    nh = length(μs)
    loss = CUDA.zeros(Float32, 3)
    da = similar(a)
    dμs, dlvs = map(similar, μs), map(similar, logσ²s)
    ptrs(xs) = Ptr{Cvoid}[reinterpret(Ptr{Cvoid}, pointer(v)) for v in xs]
    check(h, ccall((:ldeq_elbo_logits_fwd_bwd, libldeq), Cint,
                   (Ptr{Cvoid}, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Cint, Cfloat, Cint, Cint, Cint,
                    Cfloat, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
                   h, x, a, ptrs(μs), ptrs(logσ²s), Int32[size(m, 1) for m in μs], nh, β, B, T, P, grad_scale, loss, da,
                   ptrs(dμs), ptrs(dlvs), CUDA.stream().handle))
    return loss, da, dμs, dlvs
end

# Flux ADAMW(η, (β₁, β₂), decay) on the flat parameter vector of Flux.destructure(model); `step` is 1-based
adamw_step!(p::CuVector{Float32}, g::CuVector{Float32}, m::CuVector{Float32}, v::CuVector{Float32}, step::Integer;
            η = 1e-3, β = (0.9, 0.999), ϵ = 1e-8, decay = 1f-3, grad_scale = 1f0) =
    check(handle(), ccall((:ldeq_adamw_step, libldeq), Cint,
                          (Ptr{Cvoid}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, Int64, Cdouble, Cdouble, Cdouble, Cdouble,
                           Cfloat, Int64, Cfloat, Ptr{Cvoid}),
                          handle(), p, g, m, v, length(p), η, β[1], β[2], ϵ, decay, step, grad_scale, CUDA.stream().handle))


# ---- data-parallel training: one Julia process per GPU, flat gradient summed with NCCL through the C ABI ---------------
# (no reference counterpart: examples/pendulum_friction-less/model_train.jl:186-208 is a single-process loop)
function comm_init!(h, rank::Integer, nranks::Integer, idfile::AbstractString)
    id = Vector{UInt8}(undef, 128)
    if rank == 0
        check(h, ccall((:ldeq_comm_unique_id, libldeq), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h, id))
        write(idfile * ".tmp", id); mv(idfile * ".tmp", idfile; force = true)
    else
        while !isfile(idfile); sleep(0.05); end
        id = read(idfile)
    end
    check(h, ccall((:ldeq_comm_init, libldeq), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), h, id, rank, nranks))
end

# g = the flat gradient of Flux.destructure(model) order, a CuVector{Float32}; summed in place over the ranks
allreduce_grads!(h, g::CuVector{Float32}) =
    check(h, ccall((:ldeq_allreduce_grads, libldeq), Cint, (Ptr{Cvoid}, CuPtr{Cfloat}, Int64, Ptr{Cvoid}),
                   h, g, length(g), CUDA.stream().handle))

end # module
