# LatentDiffEqB200.jl -- Julia glue a LatentDiffEq.jl maintainer would add to route `diffeq_layer` to libldeq.so.
#
# WRITTEN TO SPEC, NOT EXECUTABLE IN THE BUILD IMAGE (no Julia there).  It binds include/ldeq.h with `ccall` and
# adds more specific `diffeq_layer` methods for CuArray inputs plus `ChainRulesCore.rrule`s, leaving the
# LatentDiffEqModel / GOKU / LatentODE constructors, the layer API and the diffeq struct fields untouched
# (reference src/models/GOKU.jl:98-130, src/models/LatentODE.jl:61-78, src/models/LatentDiffEqModel.jl:101-113).
module LatentDiffEqB200

using CUDA, ChainRulesCore, Flux
import LatentDiffEq
import LatentDiffEq: diffeq_layer, transform_after_diffeq, Decoder, GOKU, LatentODE

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
    sensealg::Int32 = 0   # 0 discrete adjoint of the primal steps, 1 the reference's dual-number re-solves (LDEQ_SENSE_FORWARD_DUAL)
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

# which built-in right-hand side a diffeq struct stands for; user structs add a method (or use ldeq_rhs_from_source)
rhs_kind(diffeq) = error("define LatentDiffEqB200.rhs_kind(::$(typeof(diffeq))) (0 = Pendulum, 1 = Pendulum_friction)")

function rhs_object(h, diffeq)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(h, ccall((:ldeq_rhs_builtin, libldeq), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), h, rhs_kind(diffeq), r))
    r[]
end

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
    opts = Ref(Opts(diffeq.kwargs))
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
    opts = Ref(Opts(diffeq.kwargs))
    tp = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ldeq_mlp_solve_fwd, libldeq), Cint,
               (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cdouble}, Cint, Cint, Ref{Opts}, CuPtr{Cvoid},
                CuPtr{Int32}, CuPtr{Int32}, CuPtr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
               h, dtype_code(T), u0, CuArray(p), dims, length(dims) - 1, tt, B, length(tt), opts, ẑ,
               CU_NULL, CU_NULL, CU_NULL, tape ? tp : C_NULL, CUDA.stream().handle)
    check(h, rc)
    return ẑ, Tape(tp[], true)
end

# reverse pass of the LatentODE method (replaces DiffEqFlux's InterpolatingAdjoint pullback, LatentODE.jl:70-72 under Zygote):
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

# ---- the host-array methods of the reference (GOKU.jl:102-103: `cpu(...)`) can use the *_host entry points ------
function diffeq_layer(decoder::Decoder{M}, l̂::Tuple{Matrix{Float32},Matrix{Float32}}, t) where {M<:GOKU}
    h = handle()
    ẑ₀, θ̂ = l̂
    z, B = size(ẑ₀)
    tt = collect(Float64, t)
    ẑ = Array{Float32}(undef, z, B, length(tt))
    check(h, ccall((:ldeq_solve_fwd_host, libldeq), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Ref{Opts}, Ptr{Cvoid},
                    Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
                   h, rhs_object(h, decoder.diffeq), Cint(0), ẑ₀, θ̂, tt, B, length(tt), Ref(Opts(decoder.diffeq.kwargs)), ẑ,
                   C_NULL, C_NULL, C_NULL, C_NULL, CUDA.stream().handle))
    return transform_after_diffeq(ẑ, decoder.diffeq)
end


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
