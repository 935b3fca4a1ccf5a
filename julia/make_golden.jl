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
#   * C3 with `solver = DP5()` / `BS3()` (adaptive and fixed step dt = 0.08) and `RK4()` (fixed step), Float64: pins the oracle's
#         restatement of the other solvers (SURVEY.md 8(f)4), same outputs as above;
#   * the pattern extractor (GOKU.jl:30-49): `Chain(RNN(32,16,relu), RNN(16,16,relu))`, two `Chain(LSTM(32,16), LSTM(16,16))`
#         and LatentODE's `Chain(RNN(32,32,relu), RNN(32,32,relu))` rebuilt from the fixture's `Flux.destructure` vectors; final
#         states and the Zygote gradients with respect to the frames and the parameter vectors;
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

# `solver`: the diffeq struct's field, a keyword of its constructor (pendulum.jl:11,58)
function run_case(c, diffeq_ctor, T; solver = nothing, modes = ((:adaptive, NamedTuple()), (:fixed, (adaptive = false, dt = 0.05))))
    z0, th, t, d = T.(c[:z0]), T.(c[:theta]), range(0.0, step = 0.05, length = length(c[:t])), T.(c[:dtraj])
    out = Dict{Symbol,Any}()
    for (name, kw) in modes
        diffeq = solver === nothing ? diffeq_ctor(; kw...) : diffeq_ctor(; solver = solver, kw...)   # kwargs are splatted into solve (pendulum.jl:11,43; GOKU.jl:108)
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
# other solvers of the diffeq struct (SURVEY.md 8(f)4): C3's right-hand side in Float64
golden[:c3_f64_dp5] = run_case(inp[:c3], Pendulum_friction, Float64; solver = DP5(), modes = ((:adaptive, NamedTuple()), (:fixed, (adaptive = false, dt = 0.08))))
golden[:c3_f64_bs3] = run_case(inp[:c3], Pendulum_friction, Float64; solver = BS3(), modes = ((:adaptive, NamedTuple()), (:fixed, (adaptive = false, dt = 0.08))))
golden[:c3_f64_rk4] = run_case(inp[:c3], Pendulum_friction, Float64; solver = RK4(), modes = ((:fixed, (adaptive = false, dt = 0.08)),))

# the pattern extractor (GOKU.jl:30-49, LatentODE.jl:20-34) on the fixture's frames and Flux.destructure vectors
function run_pattern_extractor(pe)
    x = Float32.(pe[:x])                                     # (F, B, T)
    frames = Flux.unstack(x, 3)
    _, re_rnn = Flux.destructure(Chain(RNN(32, 16, relu), RNN(16, 16, relu)))
    _, re_lstm = Flux.destructure(Chain(LSTM(32, 16), LSTM(16, 16)))
    _, re_rnn32 = Flux.destructure(Chain(RNN(32, 32, relu), RNN(32, 32, relu)))
    final(re, p, fr) = (m = re(p); Flux.reset!(m); [m(f) for f in fr][end])
    f16(x, pr, pf, pb) = (fr = Flux.unstack(x, 3); (final(re_rnn, pr, reverse(fr)), vcat(final(re_lstm, pf, fr), final(re_lstm, pb, reverse(fr)))))
    (z0o, θo), back = Zygote.pullback(f16, x, Float32.(pe[:rnn]), Float32.(pe[:lstm_f]), Float32.(pe[:lstm_b]))
    dx, drnn, dlf, dlb = back((Float32.(pe[:dz0]), Float32.(pe[:dtheta])))
    f32(x, pr) = final(re_rnn32, pr, reverse(Flux.unstack(x, 3)))
    z32, back32 = Zygote.pullback(f32, x, Float32.(pe[:rnn32]))
    dx32, drnn32 = back32(Float32.(pe[:dz0_32]))
    Dict(:z0_out => z0o, :theta_out => θo, :dx => dx, :d_rnn => drnn, :d_lstm_f => dlf, :d_lstm_b => dlb,
         :z0_out_32 => z32, :dx_32 => dx32, :d_rnn32 => drnn32)
end
haskey(inp, :pe) && (golden[:pe] = run_pattern_extractor(inp[:pe]))

x = inp[:trig_x]
golden[:trig] = Dict(:sin => sin.(x), :cos => cos.(x))
golden[:fastpow] = Dict(:b1 => [DiffEqBase.fastpow(v, inp[:fastpow_y][1]) for v in inp[:fastpow_x]],
                        :b2 => [DiffEqBase.fastpow(v, inp[:fastpow_y][2]) for v in inp[:fastpow_x]])
golden[:versions] = Dict(:julia => string(VERSION))
BSON.bson(joinpath(root, "tests", "golden", "julia_golden.bson"), golden)
println("wrote tests/golden/julia_golden.bson")
