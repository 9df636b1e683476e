#!/usr/bin/env julia
# tests/golden/make_golden.jl -- REFERENCE-DERIVED golden traces.
#
#     julia tests/golden/make_golden.jl /path/to/DZOptimization.jl  [tests/golden/golden_ref.json]
#     python -m pytest tests/test_golden_ref.py            # turns the oracle's parity from "unpinned" to pinned
#
# Why this file exists.  The hot path this repository re-implements (BFGSOptimizer / GradientDescentOptimizer `step!`) is
# COMMENTED-OUT code in the reference (legacy/DZOptimization.jl is one block comment, the BFGS part a second, nested one)
# and calls helpers that are defined nowhere in the reference tree (SURVEY.md section 0.3).  The build/test image of
# this repository has no Julia, so the committed fixtures (golden.json, golden_next.json) come from a Python
# restatement.  This script is the recipe a maintainer WITH Julia (>= 1.10) runs once to pin the oracle against the
# reference's own source text:
#
#   1. it reads legacy/DZOptimization.jl from the reference checkout and strips exactly the four comment markers
#      (`#=` on lines 1 and 698, `=#` on lines 1144 and 1147) -- every other character is the reference's;
#   2. it supplies the missing helpers per SURVEY.md section 8.0 (marked [GLUE] below: LineSearchFunctor,
#      quadratic_line_search, scalar_mul!, add!, negate!/3, mul!, norm, NULL_CONSTRAINT, extended Rosenbrock);
#   3. it replays the reference's `step!` (and its own `run_and_test!` invariants, :998-1049) on the inputs of
#      golden.json and writes every float as a C99 hex literal in golden.json's schema.
#
# What comes out is reference code wherever the reference has code: `step!(::BFGSOptimizer)` :891-994,
# `update_inverse_hessian!` :864-889, the constructor :762-810, `identity_matrix!` :712-720, `step!(::GradientDescentOptimizer)`
# :393-449 with its constructor :330-374, `find_three_point_bracket` :49-172, `QuadraticLineSearch` :181-216, the BLAS-1
# kernels of legacy/Kernels.jl, `rosenbrock_*` and `riesz_*` of legacy/ExampleFunctions.jl and legacy/PCG.jl.  The GD
# traces contain NO glue except NULL_CONSTRAINT and rsqrt; the BFGS traces contain the [GLUE] line-search driver, which
# re-uses the reference's bracket / interpolation logic line for line but takes the first step t1 as an argument.
#
# Dependencies: none beyond Julia's stdlib.  legacy/Kernels.jl and legacy/ExampleFunctions.jl `using MultiFloats, SIMD`
# only for their MultiFloat SIMD methods, which this path never calls; if those packages are not installed the two tiny
# stand-ins below satisfy the `using` lines ([GLUE] rsqrt(x::Float64) = inv(sqrt(x)), as in SURVEY.md 8c).
#
# This script has not been executed in this repository's environment (no Julia there); it is written against the
# reference text cited above and is meant to be fixed, not trusted, if a Julia version rejects a line.

length(ARGS) >= 1 || error("usage: julia make_golden.jl /path/to/DZOptimization.jl [out.json]")
const REF = ARGS[1]
const OUT = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "golden_ref.json")
const LEGACY = joinpath(REF, "legacy")

# ------------------------------------------------------------------ 1. the reference text, comment markers removed
function uncommented_legacy_source()
    lines = readlines(joinpath(LEGACY, "DZOptimization.jl"))
    for (ln, want) in ((1, "#="), (698, "#="), (1144, "=#"), (1147, "=#"))
        strip(lines[ln]) == want ||
            error("legacy/DZOptimization.jl:$ln is `$(lines[ln])`, expected `$want` -- different reference version?")
    end
    keep = String[]
    for (ln, text) in enumerate(lines)
        ln in (1, 698, 1144, 1147) && continue
        # lines 3-6: include()s and `using .Kernels` -- done by hand below so that the stand-in packages are in scope
        (3 <= ln <= 6) && continue
        push!(keep, text)
    end
    return join(keep, "\n")
end

nested(src) = replace(replace(src, "using MultiFloats:" => "using ..MultiFloats:"), "using SIMD:" => "using ..SIMD:")

module RefHost

# ---- stand-ins for the two packages whose SIMD MultiFloat methods this path never calls  [GLUE]
module MultiFloats
export MultiFloat, MultiFloatVec, rsqrt, mfvgather
struct MultiFloat{T,N} <: AbstractFloat end
struct MultiFloatVec{M,T,N} end
@inline rsqrt(x::AbstractFloat) = inv(sqrt(x))            # [GLUE] SURVEY.md 8c: 1.0 / sqrt(x)
mfvgather(args...) = error("MultiFloat SIMD path is outside the golden traces")
end
module SIMD
export Vec
struct Vec{M,T}
    data::NTuple{M,T}
end
end

using LinearAlgebra: LinearAlgebra                           # NOT imported into scope: mul! / norm below are [GLUE]

end # module RefHost (re-opened below with Core.eval / include_string)

Base.include_string(RefHost, nested(read(joinpath(LEGACY, "Kernels.jl"), String)), "legacy/Kernels.jl")
Base.include_string(RefHost, nested(read(joinpath(LEGACY, "ExampleFunctions.jl"), String)), "legacy/ExampleFunctions.jl")
Base.include_string(RefHost, read(joinpath(LEGACY, "PCG.jl"), String), "legacy/PCG.jl")

# ------------------------------------------------------------------ 2. [GLUE] the helpers the reference never defines
Base.include_string(RefHost, raw"""
using .Kernels: dot, norm2, inv_norm, negate!, scale!, delta!, axpy!          # legacy/DZOptimization.jl:6
import .Kernels: negate!

const NULL_CONSTRAINT = x -> true                                              # [GLUE] used at :384, :555, :759

# [GLUE] norm = sqrt(sequential sum of squares)  (call sites :921, :928; cf. :424 `sqrt(norm2(...))`)
norm(x::AbstractArray{T}) where {T} = sqrt(norm2(x, length(x)))

# [GLUE] mul!(y, A, x): row i = sum_j A[i,j]*x[j], j ascending, starting from zero  (call sites :875, :958-960)
function mul!(y::AbstractVector{T}, A::AbstractMatrix{T}, x::AbstractVector{T}) where {T}
    m, n = size(A)
    @inbounds for i = 1:m
        acc = zero(T)
        for j = 1:n
            acc += A[i, j] * x[j]
        end
        y[i] = acc
    end
    return y
end

# [GLUE] BLAS-1 helpers, semantics fixed by their use sites
scalar_mul!(x::AbstractArray{T}, a::T) where {T} = (for i in eachindex(x); @inbounds x[i] = x[i] * a; end; x)      # :874
negate!(dst::AbstractArray{T}, src::AbstractArray{T}, n::Int) where {T} =                                            # :943-944
    (for i = 1:n; @inbounds dst[i] = -src[i]; end; dst)
add!(dst::AbstractArray{T}, a::T, x::AbstractArray{T}, n::Int) where {T} =                                           # :945, :973
    (for i = 1:n; @inbounds dst[i] = dst[i] + a * x[i]; end; dst)
add!(dst::AbstractArray{T}, x::AbstractArray{T}, n::Int) where {T} =                                                 # :949-950
    (for i = 1:n; @inbounds dst[i] = dst[i] + x[i]; end; dst)

# [GLUE] LineSearchFunctor (:749-750, :786-791): the five constructor arguments in the order of :786-791
struct LineSearchFunctor{F,C,T,N}
    objective_function::F
    constraint_function!::C
    current_point::Array{T,N}
    new_point::Array{T,N}          # the optimizer's _scratch_space, shared by both functors
    step_direction::Array{T,N}
end
# probe(t): w = x - t*dir (sign from :945, :973); +Inf when infeasible (as LineSearchEvaluator :36-44)
function (lsf::LineSearchFunctor{F,C,T,N})(t::T) where {F,C,T,N}
    x, w, d = lsf.current_point, lsf.new_point, lsf.step_direction
    @inbounds for i in eachindex(x)
        w[i] = x[i] + (-t) * d[i]
    end
    return lsf.constraint_function!(w) ? lsf.objective_function(w) : typemax(T)
end

# [GLUE] quadratic_line_search(functor, f0, t1) -> (t*, f*): find_three_point_bracket (:49-172) with first step t1
# instead of 1 and unlimited doublings, followed by QuadraticLineSearch (:191-216).  Branch for branch the reference's.
function quadratic_line_search(lsf::LineSearchFunctor{F,C,T,N}, f0::T, t1::T) where {F,C,T,N}
    _zero = zero(T)
    x, w, d = lsf.current_point, lsf.new_point, lsf.step_direction
    bracket = (_zero, f0, _zero, f0)
    while true                                                        # `break` = the early returns of :64-123
        (!isfinite(f0) || !isfinite(t1) || iszero(t1)) && break       # :64-66 + [GLUE] guard for a zero direction norm
        all(iszero, d) && break                                       # :71-85
        step_size = t1
        step_is_small = false
        moved() = begin
            changed = false
            @inbounds for i in eachindex(x)
                new = x[i] + (-step_size) * d[i]
                changed |= (x[i] != new)
                w[i] = new
            end
            changed
        end
        point_changed = moved()
        doublings = 0
        while !point_changed                                          # :89-101
            step_size += step_size
            step_is_small = true
            point_changed = moved()
            doublings += 1
            doublings >= 4096 && break                                # [GLUE] DZO_LINESEARCH_CAP (never reached)
        end
        point_changed || break
        is_feasible = lsf.constraint_function!(w)                     # :104
        if step_is_small && (!is_feasible || x == w)                  # :107-123
            break
        end
        f1 = is_feasible ? lsf.objective_function(w) : typemax(T)     # :126
        if f1 <= f0                                                   # :130
            reference_point = copy(w)                                 # :136
            while true                                                # :143-156 (max_increases = 0: unlimited)
                double_step_size = step_size + step_size
                f2 = lsf(double_step_size)
                if !isfinite(f2) || (f2 > f1) || (w == reference_point)
                    bracket = (step_size, f1, double_step_size, f2)
                    break
                end
                step_size = double_step_size
                f1 = f2
                copy!(reference_point, w)
            end
        else                                                          # :157-171
            _half = inv(one(T) + one(T))
            while true
                half_step_size = _half * step_size
                f2 = lsf(half_step_size)
                if f2 <= f0
                    bracket = (half_step_size, f2, step_size, f1)
                    break
                end
                step_size = half_step_size
                f1 = f2
            end
        end
        break
    end
    (x1, f1, x2, f2) = bracket
    xb, fb = _zero, f0                                                # :196-214 verbatim
    if f1 < fb
        xb, fb = x1, f1
    end
    if f2 < fb
        xb, fb = x2, f2
    end
    delta_1 = f0 - f1
    delta_2 = f2 - f1
    sum_deltas = delta_1 + delta_2
    if (delta_1 >= _zero) && (delta_2 >= _zero) && (sum_deltas > _zero)
        twice_delta_1 = delta_1 + delta_1
        delta_ratio = (twice_delta_1 + sum_deltas) / (sum_deltas + sum_deltas)
        xq = delta_ratio * x1
        fq = lsf(xq)
        if fq < fb
            xb, fb = xq, fq
        end
    end
    return (xb, fb)
end

# [GLUE] extended Rosenbrock: sum over consecutive pairs, k ascending, of the reference's n = 2 function
# (legacy/ExampleFunctions.jl:10-24); reduces to it at n = 2
function ext_rosenbrock_function(v::Vector{T}) where {T}
    acc = zero(T)
    for k = 1:2:length(v)
        acc += ExampleFunctions.rosenbrock_function(T[v[k], v[k+1]])
    end
    return acc
end
function ext_rosenbrock_gradient!(g::Vector{T}, v::Vector{T}) where {T}
    g2 = Vector{T}(undef, 2)
    for k = 1:2:length(v)
        ExampleFunctions.rosenbrock_gradient!(g2, T[v[k], v[k+1]])
        g[k], g[k+1] = g2[1], g2[2]
    end
    return g
end
""", "glue.jl")

# ------------------------------------------------------------------ the reference itself
Base.include_string(RefHost, uncommented_legacy_source(), "legacy/DZOptimization.jl (comment markers removed)")

# ------------------------------------------------------------------ 3. replay and dump
hexf(x::Float64) = begin                                               # C99 / Python float.hex() spelling
    x == 0 && return signbit(x) ? "-0x0.0p+0" : "0x0.0p+0"
    isnan(x) && return "nan"
    isinf(x) && return x > 0 ? "inf" : "-inf"
    bits = reinterpret(UInt64, abs(x))
    e = Int((bits >> 52) & 0x7ff)
    m = bits & 0x000fffffffffffff
    lead, ex = e == 0 ? (0, -1022) : (1, e - 1023)
    string(signbit(x) ? "-" : "", "0x", lead, ".", string(m, base=16, pad=13), "p", ex >= 0 ? "+" : "", ex)
end
hexv(v) = [hexf(Float64(a)) for a in v]
pcg(count, seed) = RefHost.PCG.random_fill!(zeros(count), seed)

function bfgs_trace(x0::Vector{Float64}, step::Float64, iters::Int; keep_points=true)
    f, g! = length(x0) == 2 ? (RefHost.ExampleFunctions.rosenbrock_function, RefHost.ExampleFunctions.rosenbrock_gradient!) :
            (RefHost.ext_rosenbrock_function, RefHost.ext_rosenbrock_gradient!)
    opt = RefHost.BFGSOptimizer(f, g!, copy(x0), step)
    rows = Any[]
    for _ = 1:iters
        RefHost.step!(opt)
        row = Dict{String,Any}("f" => hexf(opt.current_objective_value[]), "L" => hexf(opt.last_step_length[]),
            "type" => Int(opt.last_step_type[]), "iter" => opt.iteration_count[], "term" => opt.has_terminated[])
        keep_points && (row["x"] = hexv(opt.current_point))
        push!(rows, row)
        opt.has_terminated[] && break
    end
    Dict{String,Any}("x0" => hexv(x0), "step" => step, "tree" => false, "rows" => rows,
        "final_x" => hexv(opt.current_point), "final_d" => hexv(opt.next_step_direction),
        "final_H_row0" => hexv(opt.approximate_inverse_hessian[1, :]))
end

function gd_trace(f, g!, c!, x0::Array{Float64}, step::Float64, iters::Int, max_increases::Int)
    opt = RefHost.GradientDescentOptimizer(c!, f, g!, RefHost.QuadraticLineSearch(max_increases), copy(x0), step)
    rows = Any[]
    for _ = 1:iters
        RefHost.step!(opt)
        push!(rows, Dict{String,Any}("f" => hexf(opt.current_objective_value[]), "L" => hexf(opt.last_step_length[]),
            "iter" => opt.iteration_count[], "term" => opt.has_terminated[]))
        opt.has_terminated[] && break
    end
    Dict{String,Any}("x0" => hexv(vec(x0)), "step" => step, "tree" => false, "max_increases" => max_increases,
        "rows" => rows, "final_x" => hexv(vec(opt.current_point)), "final_d" => hexv(vec(opt.next_step_direction)))
end

# the reference's own invariants (:998-1049) on the README problem: must not throw
RefHost.run_and_test!(RefHost.BFGSOptimizer(RefHost.ExampleFunctions.rosenbrock_function,
    RefHost.ExampleFunctions.rosenbrock_gradient!, pcg(2, 0), 1.0))

g = Dict{String,Any}()
g["made_by"] = "tests/golden/make_golden.jl on Julia $(VERSION) from $(REF)"
g["pcg"] = Dict(string(seed) => hexv(pcg(8, seed)) for seed in (0, 1, 2024, 2^63 + 5))
# config 1: README Rosenbrock n=2 from "rand(2)" (PCG seeds), step 1.0, run to has_converged  -- same inputs as golden.json
g["c1_rosenbrock_n2"] = [bfgs_trace(pcg(2, s), 1.0, 500) for s in 0:5]
# config 2 (three problems of the batch): n=16, x0 = 4u-2, seed 2024
u = pcg(48, 2024)
g["c2_rosenbrock_n16"] = [bfgs_trace([4.0 * a - 2.0 for a in u[16p+1:16p+16]], 1.0, 40) for p in 0:2]
# GradientDescentOptimizer, sequential order: Riesz energy of 20 free points in the plane (golden.json's gd_riesz_free_N20_seq)
function sphere_points(N, dim, seed)
    uu = pcg(N * dim, seed)
    P = Matrix{Float64}(undef, dim, N)
    for j = 1:N
        p = [2.0 * uu[(j-1)*dim+k] - 1.0 for k = 1:dim]
        s = sqrt(sum(c * c for c in p))          # k ascending
        P[:, j] = p ./ s
    end
    P
end
g["gd_riesz_free_N20_seq"] = gd_trace(RefHost.ExampleFunctions.riesz_energy, RefHost.ExampleFunctions.riesz_gradient!,
    RefHost.NULL_CONSTRAINT, sphere_points(20, 2, 5), 1e-2, 10, 3)
u = pcg(16, 6)
g["gd_rosenbrock_n16_seq"] = gd_trace(RefHost.ext_rosenbrock_function, RefHost.ext_rosenbrock_gradient!,
    RefHost.NULL_CONSTRAINT, [4.0 * a - 2.0 for a in u], 1e-2, 20, 0)

# minimal JSON writer (stdlib only)
js(x::AbstractString) = "\"" * escape_string(x) * "\""
js(x::Bool) = x ? "true" : "false"
js(x::Integer) = string(x)
js(x::AbstractFloat) = repr(Float64(x))
js(x::AbstractVector) = "[" * join((js(a) for a in x), ", ") * "]"
js(x::AbstractDict) = "{" * join((js(string(k)) * ": " * js(v) for (k, v) in sort(collect(x), by=first)), ",\n") * "}"
open(OUT, "w") do io
    write(io, js(g))
end
println("wrote ", OUT, ": ", join(sort(collect(keys(g))), ", "))
