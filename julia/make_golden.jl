# make_golden.jl -- run ONCE under real Julia (the reference's own environment) to pin the oracle against the reference:
#
#     julia --project=/path/to/LatentDiffEq.jl julia/make_golden.jl [repo_root]
#
# Reads tests/golden/julia_inputs.bson (the seeded inputs of SURVEY.md 8(d), written by tests/golden/make_julia_inputs.py)
# and writes tests/golden/julia_golden.bson with what the REFERENCE computes on them:
#   * C1  GOKU friction-less pendulum, Float32 state / Float64 time: `diffeq_layer(decoder, (z0, theta), t)` of
#         src/models/GOKU.jl:98-130 (adaptive defaults, and fixed step dt = 0.05), its Zygote pullback (ForwardDiffSensitivity,
#         pendulum.jl:11), and per-trajectory destats (naccept / nreject) from a direct `solve`;
#   * C3  pendulum with friction in Float64 and Float32 on identical inputs;
#   * Base.sin / Base.cos(::Float32) on probe arguments (pins oracle/ldeq_oracle.cpp::jl_sinf, csrc/ldeq_julia_trig.cuh);
#   * DiffEqBase.fastpow on probe arguments (pins the oracle's fastpow restatement, SURVEY.md A.3).
# tests/test_golden.py::test_*_against_julia_golden consumes the file when it exists (and is skipped until then).
#
# NOT EXECUTABLE IN THE BUILD IMAGE (no Julia there); written against the reference's API as read from its sources.
using LatentDiffEq, OrdinaryDiffEq, Flux, Zygote, BSON, DiffEqBase
try
    @eval using SciMLSensitivity
catch
    @eval using DiffEqSensitivity      # the example environments pin the old package name (model_train.jl:16)
end

root = length(ARGS) >= 1 ? ARGS[1] : dirname(@__DIR__)
include(joinpath(root, "..", "reference", "examples", "pendulum_friction-less", "pendulum.jl"))  # Pendulum, Pendulum_friction; adjust the path

inp = BSON.load(joinpath(root, "tests", "golden", "julia_inputs.bson"))

function decoder_for(diffeq)
    enc, dec = default_layers(GOKU_basic(), 784, diffeq; device = cpu)
    LatentDiffEqModel(GOKU_basic(), enc, dec).decoder
end

function run_case(c, diffeq_ctor, T)
    z0, th, t, d = T.(c[:z0]), T.(c[:theta]), range(0.0, step = 0.05, length = length(c[:t])), T.(c[:dtraj])
    out = Dict{Symbol,Any}()
    for (name, kw) in ((:adaptive, NamedTuple()), (:fixed, (adaptive = false, dt = 0.05)))
        diffeq = diffeq_ctor(; kw...)                        # kwargs are splatted into solve (pendulum.jl:11,43; GOKU.jl:108)
        dec = decoder_for(diffeq)
        ẑ, back = Zygote.pullback(l̂ -> LatentDiffEq.diffeq_layer(dec, l̂, t), (z0, th))
        (dz0, dth), = back(d)
        # destats straight from the integrator, trajectory by trajectory (what EnsembleThreads runs, GOKU.jl:111-121)
        na, nr = Int[], Int[]
        for b in 1:size(z0, 2)
            prob = remake(diffeq.prob; u0 = z0[:, b], p = th[:, b], tspan = (t[1], t[end]))
            sol = solve(prob, diffeq.solver; saveat = t, diffeq.kwargs...)
            push!(na, sol.destats.naccept); push!(nr, sol.destats.nreject)
        end
        out[name] = Dict(:traj => ẑ, :dz0 => dz0, :dtheta => dth, :naccept => na, :nreject => nr)
    end
    out
end

golden = Dict{Symbol,Any}()
golden[:c1_f32] = run_case(inp[:c1], Pendulum, Float32)
golden[:c3_f64] = run_case(inp[:c3], Pendulum_friction, Float64)
golden[:c3_f32] = run_case(inp[:c3], Pendulum_friction, Float32)
x = inp[:trig_x]
golden[:trig] = Dict(:sin => sin.(x), :cos => cos.(x))
golden[:fastpow] = Dict(:b1 => [DiffEqBase.fastpow(v, inp[:fastpow_y][1]) for v in inp[:fastpow_x]],
                        :b2 => [DiffEqBase.fastpow(v, inp[:fastpow_y][2]) for v in inp[:fastpow_x]])
golden[:versions] = Dict(:julia => string(VERSION))
BSON.bson(joinpath(root, "tests", "golden", "julia_golden.bson"), golden)
println("wrote tests/golden/julia_golden.bson")
