# GradusB200Ext.jl -- reference-side binding for libgradus_b200 (the `ccall` shim a Gradus.jl maintainer adds).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  Every entry point used below is
# exercised through the very same C ABI from Python (gradus.jl_b200/_cabi.py, tests/).  Struct layouts mirror
# include/gradus_b200.h field for field.
#
# It replaces ext/GradusDiffEqGPUExt/GradusDiffEqGPUExt.jl:10-31 by adding a method of
#     Gradus.ensemble_solve_tracing_problem(ensemble, problem, config; ...)      (src/tracing/tracing.jl:113-196)
# for the new ensemble type `EnsembleB200`, next to `EnsembleEndpointThreads` (src/Gradus.jl:412).
module GradusB200Ext

using Gradus
using Gradus: TracingConfiguration, GeodesicPoint, StatusCodes, KerrMetric, JohannsenPsaltisMetric, ThinDisc, ShakuraSunyaev,
    DatumPlane, PolarChart, PolarPlane, GeometricGrid, LinearGrid, InverseGrid
using StaticArrays
import SciMLBase

const libgradus_b200 = get(ENV, "GRADUS_B200_LIB", "libgradus_b200")

"""
    EnsembleB200(devices = [0])

Integrate every ray of an ensemble on the listed CUDA devices (B200, sm_100a) with libgradus_b200.
"""
struct EnsembleB200
    devices::Vector{Int}
end
EnsembleB200() = EnsembleB200([0])
Gradus.restrict_ensemble(::Gradus.AbstractMetric, e::EnsembleB200) = e

# ---- POD mirrors of include/gradus_b200.h ---------------------------------------------------------------------
struct CProblem
    metric_kind::Int32; geometry_kind::Int32; callback_kind::Int32; pow_mode::Int32
    metric_params::NTuple{8,Float64}; observer::NTuple{4,Float64}; geometry_params::NTuple{4,Float64}
    gtol::Float64; chart_inner::Float64; chart_outer::Float64; callback_delta::Float64
    lambda_min::Float64; lambda_max::Float64; abstol::Float64; reltol::Float64
    dtmax::Float64; mu::Float64; maxiters::Int64
end
struct CIC
    kind::Int32; grid_kind::Int32; width::Int64; height::Int64
    lo0::Float64; hi0::Float64; lo1::Float64; hi1::Float64
    x::NTuple{4,Ptr{Float64}}; v::NTuple{4,Ptr{Float64}}; n::Int64
end
struct CRange
    first::Int64; count::Int64; stride::Int64; block::Int64
end
struct CEndpoints
    status::Ptr{Int32}; lambda_max::Ptr{Float64}
    x::NTuple{4,Ptr{Float64}}; v::NTuple{4,Ptr{Float64}}; x_init::NTuple{4,Ptr{Float64}}; v_init::NTuple{4,Ptr{Float64}}
    naccept::Ptr{Int32}; nreject::Ptr{Int32}; flags::Ptr{Int32}
end

_mp8(v...) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 8)
_metric(m::KerrMetric) = (Int32(0), _mp8(m.M, m.a))
_metric(m::JohannsenPsaltisMetric) = (Int32(1), _mp8(m.M, m.a, m.ϵ3))
_metric(m::JohannsenMetric) = (Int32(2), _mp8(m.M, m.a, m.α13, m.α22, m.α52, m.ϵ3))
_metric(m::BumblebeeMetric) = (Int32(3), _mp8(m.M, m.a, m.l))
# slot 4 (metric_params[3]) carries the charge of the test particle, q for photons and q/μ otherwise
# (geodesic_ode_problem(::KerrNewmanMetric), src/metrics/kerr-newman-ad.jl:74-78); the caller passes trace.q, trace.μ
_metric(m::KerrNewmanMetric; q = 0.0, μ = 0.0) = (Int32(4), _mp8(m.M, m.a, m.Q, isapprox(μ, 0.0) ? q : q / μ))
_metric(m) = throw(ArgumentError("EnsembleB200 has no closed-form right-hand side for $(typeof(m)); there is no CPU fallback"))
_geometry(::Nothing) = (Int32(0), (0.0, 0.0, 0.0, 0.0))
_geometry(d::ThinDisc) = (Int32(1), (d.inner_radius, d.outer_radius, 0.0, 0.0))
_geometry(d::ShakuraSunyaev) = (Int32(2), (d.Ṁ_Ṁedd, d.inv_η, d.inner_radius, 0.0))
_geometry(d::DatumPlane) = (Int32(3), (d.height, 0.0, 0.0, 0.0))
_geometry(d) = throw(ArgumentError("geometry $(typeof(d)) is outside the EnsembleB200 scope"))

# The shim recognises `domain_upper_hemisphere(δ)` (src/tracing/callbacks.jl:31-39) by the closure type of its condition
# and reads δ from the captured variable; anything else cannot run on the device.
function _callback(cb)
    isnothing(cb) && return (Int32(0), 0.0)
    if cb isa SciMLBase.DiscreteCallback && nameof(typeof(cb.condition)) === Symbol("#_domain_upper_hemisphere_check")
        return (Int32(1), Float64(cb.condition.δ))
    end
    throw(ArgumentError("only `domain_upper_hemisphere` user callbacks can run on the device"))
end

"""Render-grid initial conditions as data (what `_render_velocity_function`, src/rendering/rendering.jl:140-163, closes over)."""
struct RenderGridVelocity{T}
    image_width::Int; image_height::Int; αlims::Tuple{T,T}; βlims::Tuple{T,T}
end

function _ic(config::TracingConfiguration, problem, keep)
    v = config.velocity
    np = (C_NULL, C_NULL, C_NULL, C_NULL) .|> p -> convert(Ptr{Float64}, p)
    if v isa RenderGridVelocity
        return CIC(0, 0, v.image_width, v.image_height, v.αlims[1], v.αlims[2], v.βlims[1], v.βlims[2], np, np, v.image_width * v.image_height)
    elseif v isa PolarPlane
        gk = v.grid isa GeometricGrid ? 1 : v.grid isa InverseGrid ? 2 : v.grid isa LinearGrid ? 0 :
             throw(ArgumentError("grid $(typeof(v.grid)) is outside the EnsembleB200 scope"))
        return CIC(1, gk, v.Nr, v.Nθ, v.r_min, v.r_max, v.θ_min, v.θ_max, np, np, v.Nr * v.Nθ)
    else
        # generic path: evaluate prob_func on host threads into SoA (corona ensembles, lamp-post.jl:89-100)
        n = config.trajectories
        xs = [Vector{Float64}(undef, n) for _ = 1:4]; vs = [Vector{Float64}(undef, n) for _ = 1:4]
        Threads.@threads for i = 1:n
            u0 = problem.prob_func(problem.prob, i, 0).u0
            for k = 1:4
                xs[k][i] = u0[k]; vs[k][i] = u0[4+k]
            end
        end
        append!(keep, xs); append!(keep, vs)
        return CIC(2, 0, 0, 0, 0.0, 0.0, 0.0, 0.0, Tuple(pointer.(xs)), Tuple(pointer.(vs)), n)
    end
end

function _check(rc, ctx)
    rc == 0 && return
    msg = unsafe_string(ccall((:gb200_last_error, libgradus_b200), Cstring, (Ptr{Cvoid},), ctx))
    rc == -1 || rc == -4 ? throw(ArgumentError(msg)) : error("libgradus_b200: $msg")
end

function Gradus.ensemble_solve_tracing_problem(
    ensemble::EnsembleB200,
    problem::SciMLBase.EnsembleProblem,
    config::TracingConfiguration{T};
    progress_bar = nothing,
    save_on = false,
    solver_opts...,
) where {T}
    save_on && error("Cannot use `EnsembleB200` with `save_on`")                      # tracing.jl:159-161
    isempty(solver_opts) || throw(ArgumentError("unrecognised solver options $(keys(solver_opts))"))  # KeywordArgError
    config.solver isa Gradus.Tsit5 || throw(ArgumentError("EnsembleB200 integrates with Tsit5 only"))
    isnothing(progress_bar) || @warn "Progress bar not supported with EnsembleB200"
    mk, mp = _metric(config.metric)
    gk, gp = _geometry(config.geometry)
    ck, cδ = _callback(config.callback)
    chart = config.chart::PolarChart
    keep = Any[]
    p = CProblem(mk, gk, ck, 0, mp, Tuple(Float64.(config.position)), gp, 1e-2, chart.inner_radius, chart.outer_radius, cδ,
                 config.λ_domain[1], config.λ_domain[2], config.abstol, config.reltol, 0.0, 0.0, 0)
    ic = _ic(config, problem, keep)
    n = ic.n
    status = Vector{Int32}(undef, n); λ = Vector{Float64}(undef, n)
    x = [Vector{Float64}(undef, n) for _ = 1:4]; v = [Vector{Float64}(undef, n) for _ = 1:4]
    x0 = [Vector{Float64}(undef, n) for _ = 1:4]; v0 = [Vector{Float64}(undef, n) for _ = 1:4]
    ndev = length(ensemble.devices)
    GC.@preserve keep status λ x v x0 v0 begin
        Threads.@threads for d = 1:ndev                 # contiguous ray blocks, one context per GPU
            first = (d - 1) * n ÷ ndev; count = d * n ÷ ndev - first
            ctx = Ref{Ptr{Cvoid}}()
            _check(ccall((:gb200_init, libgradus_b200), Cint, (Cint, Ref{Ptr{Cvoid}}), ensemble.devices[d], ctx), C_NULL)
            off(a) = pointer(a, first + 1)
            out = CEndpoints(off(status), off(λ), Tuple(off.(x)), Tuple(off.(v)), Tuple(off.(x0)), Tuple(off.(v0)), C_NULL, C_NULL, C_NULL)
            rc = ccall((:gb200_trace, libgradus_b200), Cint, (Ptr{Cvoid}, Ref{CProblem}, Ref{CIC}, Ref{CRange}, Ref{CEndpoints}),
                       ctx[], p, ic, CRange(first, count, 1, 1), out)
            _check(rc, ctx[])
            ccall((:gb200_destroy, libgradus_b200), Cvoid, (Ptr{Cvoid},), ctx[])
        end
    end
    # return contract of tracing.jl:179-189: a concretely typed Vector{GeodesicPoint{T,Nothing}} in ray order
    map(1:n) do i
        GeodesicPoint(StatusCodes.T(status[i]), T(config.λ_domain[1]), λ[i],
                      SVector{4,T}(x0[1][i], x0[2][i], x0[3][i], x0[4][i]), SVector{4,T}(x[1][i], x[2][i], x[3][i], x[4][i]),
                      SVector{4,T}(v0[1][i], v0[2][i], v0[3][i], v0[4][i]), SVector{4,T}(v[1][i], v[2][i], v[3][i], v[4][i]), nothing)
    end
end

# device-side initial conditions for the two structured generators: the closures of the reference become data
Gradus._render_velocity_function(m::Union{KerrMetric{T},JohannsenPsaltisMetric{T}}, position, w, h, αlims, βlims) where {T} =
    RenderGridVelocity{T}(w, h, T.(αlims), T.(βlims))
# (when the ensemble is not EnsembleB200 the maintainer keeps the stock closure: dispatch on the ensemble in render_configuration)

export EnsembleB200
end # module
