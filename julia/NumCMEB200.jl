# NumCMEB200.jl -- Julia-side glue that keeps NumCME.jl's API for the FSP right-hand-side path while every
# kernel runs in libncme (hand-written CUDA for sm_100a, C ABI in include/ncme.h).
#
# NOT RUNNABLE IN THE BUILD ENVIRONMENT (no Julia there); written against include/ncme.h and reviewed by eye
# against the reference signatures quoted next to each method.  The Python package numcme.jl_b200/ is the executed
# mirror of exactly this file (same call sequence per method) and is what the parity tests drive.  What IS checked
# mechanically (tests/test_julia_glue.py): every ccall against the C ABI (symbol, arity, scalar/pointer category of
# each argument), that NO method defined here has a signature type-equal to a method of the reference (nothing of
# NumCME is overwritten), and that the solve entry points, the Broadcast surface and the multi-GPU plumbing exist.
#
# Usage: `include("NumCMEB200.jl"); using .NumCMEB200` next to `using NumCME`.  The reference's CPU types and methods
# keep working untouched; the B200 path is selected by TYPE:
#   * spaces / matrices:   StateSpaceSparseB200, FspMatrixSparseB200, ForwardSensFspMatrixSparseB200 (same constructor
#                          arguments, same generic functions: expand!, deleteat!, matvec!, matvecadd!, matvec, *, size, ...)
#   * solves:              wrap the reference's algorithm object:  solve(model, p0, tspan, OnB200(alg); kwargs...)
#                          alg = AdaptiveFspSparse(...), AdaptiveForwardSensFspSparse(...), or an ODE method / nothing
#                          (fixed space).  `ode_method = nothing` -> native device-resident BDF/GMRES of libncme;
#                          a DifferentialEquations.jl algorithm -> the reference's own loop with the FSP vector as a
#                          DeviceVector (Broadcast surface below: linear combinations, the residual/WRMS forms).
module NumCMEB200

using NumCME
using StaticArrays: MVector
import DifferentialEquations as DE
using DifferentialEquations.DiffEqBase: AbstractODEAlgorithm
import NumCME: expand!, deleteat!, get_state_count, get_sink_count, get_states, get_statedict,
    get_state_connectivity, get_sink_connectivity, get_stoich_matrix, matvec!, matvecadd!, matvec,
    get_rowcount, get_colcount, get_parameters, get_propensities, init!, adapt!, solve, get_propensity_gradients,
    get_gradient_sparsity_patterns, get_parameter_count, get_probability, get_sensitivity
import Base: size, *
import Base.Broadcast: Broadcasted, BroadcastStyle
import LinearAlgebra
import LinearAlgebra: mul!

const libncme = get(ENV, "NCME_LIB", joinpath(@__DIR__, "..", "numcme.jl_b200", "lib", "libncme.so"))

# ------------------------------------------------------------------------------------------------ errors
struct NcmeError <: Exception
    code::Cint
    msg::String
end
function check(code::Cint)
    code == 0 && return nothing
    msg = unsafe_string(ccall((:ncme_last_error, libncme), Cstring, ()))
    code == -1 && throw(ArgumentError(msg))          # NCME_ERR_ARG  <-> ArgumentError / DimensionMismatch
    code == -3 && throw(OutOfMemoryError())
    throw(NcmeError(code, msg))
end

# ------------------------------------------------------------------------------------------------ context
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ncme_ctx_create, libncme), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        finalizer(c -> ccall((:ncme_ctx_destroy, libncme), Cint, (Ptr{Cvoid},), c.h), new(r[]))
    end
end
const DEFAULT_CTX = Ref{Union{Nothing,Context}}(nothing)
default_ctx() = (DEFAULT_CTX[] === nothing && (DEFAULT_CTX[] = Context(0)); DEFAULT_CTX[])

# ------------------------------------------------------------------------------------------------ multi-GPU plumbing
# One Julia process per GPU (Distributed.jl / MPI.jl launch them); rank 0 creates the NCCL id and the host language
# broadcasts its 128 bytes, exactly like numcme.jl_b200/parallel.py does through torch.distributed.
mutable struct Comm
    ctx::Context
    h::Ptr{Cvoid}
    rank::Int
    nranks::Int
end
function unique_id()
    buf = zeros(UInt8, 128)
    check(ccall((:ncme_comm_unique_id, libncme), Cint, (Ptr{UInt8},), buf))
    buf
end
function Comm(ctx::Context, rank::Integer, nranks::Integer, uid::Vector{UInt8})
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ncme_comm_create, libncme), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}, Ref{Ptr{Cvoid}}), ctx.h, rank, nranks, uid, r))
    finalizer(c -> ccall((:ncme_comm_destroy, libncme), Cint, (Ptr{Cvoid},), c.h), Comm(ctx, r[], rank, nranks))
end
_commh(c::Union{Nothing,Comm}) = c === nothing ? C_NULL : c.h
# row cuts of ncme_matrix_create_sharded: contiguous blocks, boundaries on multiples of 64 rows
function shard_bounds(n::Integer, nranks::Integer)
    cuts = Int64[0]
    for r in 1:nranks-1
        push!(cuts, min(n, cld((n * r) ÷ nranks, 64) * 64))
    end
    push!(cuts, n)
    cuts
end

# ------------------------------------------------------------------------------------------------ device vector
# The FSP vector stays in HBM.  AbstractVector surface needed by the reference's own code:
#   u[end-R+1:end], u[1:end-R]  (fspsolve.jl:146,172-173)  -> getindex(::UnitRange) downloads a slice
#   similar / copy / zero / length / size / fill! / copyto!, reductions (sum, dot, norm, any(isnan, .)) and the
#   Broadcast surface further down (K7).  Element-wise scalar indexing is an ERROR unless allowscalar(true): an
#   integrator that silently walks the vector element by element would do one PCIe round trip per entry.
mutable struct DeviceVector <: AbstractVector{Float64}
    ctx::Context
    ptr::Ptr{Cvoid}
    n::Int
    owned::Bool
end
function DeviceVector(ctx::Context, n::Integer)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ncme_dmalloc, libncme), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx.h, 8 * max(n, 1), r))
    v = DeviceVector(ctx, r[], n, true)
    finalizer(x -> x.owned && ccall((:ncme_dfree, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.h, x.ptr), v)
end
function DeviceVector(ctx::Context, a::Vector{Float64})
    v = DeviceVector(ctx, length(a))
    check(ccall((:ncme_h2d, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Csize_t), ctx.h, v.ptr, a, 8 * length(a)))
    v
end
const ALLOW_SCALAR = Ref(false)
allowscalar(flag::Bool) = (ALLOW_SCALAR[] = flag)
Base.size(v::DeviceVector) = (v.n,)
Base.length(v::DeviceVector) = v.n
Base.similar(v::DeviceVector) = DeviceVector(v.ctx, v.n)
Base.similar(v::DeviceVector, ::Type{Float64}) = DeviceVector(v.ctx, v.n)
Base.similar(v::DeviceVector, ::Type{Float64}, dims::Tuple{Int}) = DeviceVector(v.ctx, dims[1])
Base.zero(v::DeviceVector) = fill!(similar(v), 0.0)
Base.vec(v::DeviceVector) = v
Base.view(v::DeviceVector, r::UnitRange{<:Integer}) = DeviceVector(v.ctx, v.ptr + 8 * (first(r) - 1), length(r), false)
function Base.getindex(v::DeviceVector, r::UnitRange{<:Integer})        # slice download (host Vector)
    out = Vector{Float64}(undef, length(r))
    isempty(r) && return out
    check(ccall((:ncme_d2h, libncme), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Csize_t), v.ctx.h, out,
        v.ptr + 8 * (first(r) - 1), 8 * length(r)))
    out
end
function Base.getindex(v::DeviceVector, i::Integer)
    ALLOW_SCALAR[] || error("scalar indexing of a DeviceVector (one PCIe round trip per element); use ranges, the fused " *
                            "vector operations, or NumCMEB200.allowscalar(true) for debugging")
    v[i:i][1]
end
Base.setindex!(v::DeviceVector, x, i::Integer) = error("scalar setindex! on a DeviceVector is not supported; upload a slice with copyto!(view(v, r), host)")
function Base.copyto!(y::DeviceVector, x::Vector{Float64})              # upload
    length(x) == y.n || throw(DimensionMismatch("copyto!: lengths differ"))
    check(ccall((:ncme_h2d, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Csize_t), y.ctx.h, y.ptr, x, 8 * length(x)))
    y
end
Base.Array(v::DeviceVector) = v[1:v.n]
Base.fill!(v::DeviceVector, a::Real) = (check(ccall((:ncme_vec_fill, libncme), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Cvoid}), v.ctx.h, v.n, a, v.ptr)); v)
function Base.copyto!(y::DeviceVector, x::DeviceVector)
    x.n == y.n || throw(DimensionMismatch("copyto!: lengths differ"))
    check(ccall((:ncme_vec_copy, libncme), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}), y.ctx.h, y.n, x.ptr, y.ptr)); y
end
Base.copy(x::DeviceVector) = copyto!(similar(x), x)
function Base.sum(v::DeviceVector)
    r = Ref{Float64}(0)
    check(ccall((:ncme_vec_sum, libncme), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ref{Float64}), v.ctx.h, v.n, v.ptr, r)); r[]
end
function LinearAlgebra.dot(x::DeviceVector, y::DeviceVector)
    r = Ref{Float64}(0)
    check(ccall((:ncme_vec_dot, libncme), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}), x.ctx.h, x.n, x.ptr, y.ptr, r)); r[]
end
LinearAlgebra.norm(x::DeviceVector) = sqrt(LinearAlgebra.dot(x, x))
Base.sum(::typeof(abs2), x::DeviceVector) = LinearAlgebra.dot(x, x)
Base.mapreduce(::typeof(abs2), ::typeof(+), x::DeviceVector) = LinearAlgebra.dot(x, x)
function hasnonfinite(x::DeviceVector)
    r = Ref{Cint}(0)
    check(ccall((:ncme_vec_any_nonfinite, libncme), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ref{Cint}), x.ctx.h, x.n, x.ptr, r)); r[] != 0
end
Base.any(::typeof(isnan), x::DeviceVector) = hasnonfinite(x)
scale!(x::DeviceVector, a::Real) = (check(ccall((:ncme_vec_scale, libncme), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Cvoid}), x.ctx.h, x.n, a, x.ptr)); x)
LinearAlgebra.rmul!(x::DeviceVector, a::Real) = scale!(x, a)
axpy!(a::Real, x::DeviceVector, y::DeviceVector) = (check(ccall((:ncme_vec_axpy, libncme), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Cvoid}, Ptr{Cvoid}), y.ctx.h, y.n, a, x.ptr, y.ptr)); y)
LinearAlgebra.axpy!(a::Real, x::DeviceVector, y::DeviceVector) = axpy!(a, x, y)
function lincomb!(out::DeviceVector, coefs::Vector{Float64}, xs::Vector{DeviceVector})       # out = sum_k c_k x_k, k <= 8
    ps = Ptr{Cvoid}[x.ptr for x in xs]
    check(ccall((:ncme_vec_lincomb, libncme), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{Float64}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
        out.ctx.h, out.n, length(coefs), coefs, ps, out.ptr)); out
end
function wrms(x::DeviceVector, u0::DeviceVector, u1::DeviceVector, atol::Real, rtol::Real)    # the integrator's error norm
    r = Ref{Float64}(0)
    check(ccall((:ncme_vec_wrms, libncme), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Ref{Float64}),
        x.ctx.h, x.n, x.ptr, u0.ptr, u1.ptr, atol, rtol, r)); r[]
end
function residuals!(out::DeviceVector, x::DeviceVector, u0::DeviceVector, u1::DeviceVector, atol::Real, rtol::Real)
    check(ccall((:ncme_vec_residuals, libncme), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Ptr{Cvoid}),
        out.ctx.h, out.n, x.ptr, u0.ptr, u1.ptr, atol, rtol, out.ptr)); out
end
shift!(x::DeviceVector, a::Real) = (check(ccall((:ncme_vec_shift, libncme), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Cvoid}), x.ctx.h, x.n, a, x.ptr)); x)

# ---- Broadcast surface (SURVEY.md H3).  Call sites: DifferentialEquations.jl algorithms update `u` with fused
# broadcasts (`@.. u = uprev + dt*(a21*k1 + ...)`, always LINEAR in the vectors) and form the error estimate with
# `calculate_residuals!` (x / (atol + rtol*max(|u0|,|u1|))); the reference's loop hands them `u` (fspsolve.jl:158-161,
# rstepadapters.jl:101-102).  Without CUDA.jl there is no broadcast code generation, so a Broadcasted tree over
# DeviceVectors is pattern-matched: linear combinations lower to ncme_vec_lincomb (+ ncme_vec_shift for a constant
# term), the residual form to ncme_vec_residuals; anything else is a loud error, never a silent element-wise loop.
struct DeviceStyle <: Broadcast.AbstractArrayStyle{1} end
DeviceStyle(::Val{N}) where {N} = DeviceStyle()
BroadcastStyle(::Type{DeviceVector}) = DeviceStyle()

struct LinForm                      # c0 + sum_k coefs[k] * vecs[k]
    c0::Float64
    coefs::Vector{Float64}
    vecs::Vector{DeviceVector}
end
_isconst(l::LinForm) = isempty(l.vecs)
_scale(l::LinForm, a::Float64) = LinForm(a * l.c0, a .* l.coefs, l.vecs)
_add(a::LinForm, b::LinForm) = LinForm(a.c0 + b.c0, vcat(a.coefs, b.coefs), vcat(a.vecs, b.vecs))
_lin(x::Number) = LinForm(Float64(x), Float64[], DeviceVector[])
_lin(x::Base.RefValue) = _lin(x[])
_lin(x::Tuple{<:Number}) = _lin(x[1])
_lin(v::DeviceVector) = LinForm(0.0, [1.0], [v])
_lin(x) = nothing                   # host arrays and everything else: not part of the device surface
function _lin(bc::Broadcasted)
    f = bc.f
    ls = map(_lin, bc.args)
    any(l -> l === nothing, ls) && return nothing
    if f === (+)
        return reduce(_add, ls)
    elseif f === (-)
        length(ls) == 1 && return _scale(ls[1], -1.0)
        length(ls) == 2 && return _add(ls[1], _scale(ls[2], -1.0))
    elseif f === (*)
        nonconst = [l for l in ls if !_isconst(l)]
        length(nonconst) > 1 && return nothing            # product of two vectors: not linear
        a = prod(Float64[l.c0 for l in ls if _isconst(l)]; init = 1.0)
        return isempty(nonconst) ? LinForm(a, Float64[], DeviceVector[]) : _scale(nonconst[1], a)
    elseif f === (/)
        (length(ls) == 2 && _isconst(ls[2])) && return _scale(ls[1], 1.0 / ls[2].c0)
    elseif f === muladd
        length(ls) == 3 || return nothing
        (_isconst(ls[1]) || _isconst(ls[2])) || return nothing
        prodform = _isconst(ls[1]) ? _scale(ls[2], ls[1].c0) : _scale(ls[1], ls[2].c0)
        return _add(prodform, ls[3])
    elseif f === identity || f === float || f === Float64
        length(ls) == 1 && return ls[1]
    end
    nothing
end
_firstdev(x) = nothing
_firstdev(v::DeviceVector) = v
function _firstdev(bc::Broadcasted)
    for a in bc.args
        d = _firstdev(a)
        d === nothing || return d
    end
    nothing
end
function Base.similar(bc::Broadcasted{DeviceStyle}, ::Type{T}) where {T}
    T === Float64 || error("DeviceVector broadcast: only Float64 results are supported (got $T)")
    d = _firstdev(bc)
    DeviceVector(d.ctx, d.n)
end
function Base.copyto!(dest::DeviceVector, bc::Broadcasted{DeviceStyle})
    # (1) the residual form of OrdinaryDiffEq: calculate_residuals(x, u0, u1, atol, rtol, internalnorm, t)
    if nameof(bc.f) === :calculate_residuals && length(bc.args) >= 5 && bc.args[1] isa DeviceVector &&
       bc.args[2] isa DeviceVector && bc.args[3] isa DeviceVector && bc.args[4] isa Number && bc.args[5] isa Number
        return residuals!(dest, bc.args[1], bc.args[2], bc.args[3], bc.args[4], bc.args[5])
    end
    # (2) linear combinations
    l = _lin(bc)
    l === nothing && error("DeviceVector broadcast: `$(bc.f)` over device vectors is not a linear combination or the " *
                           "residual form; supported: +, -, scalar*vector, vector/scalar, muladd, calculate_residuals. " *
                           "Use ode_method = nothing (native device integrator) or add the fused kernel to libncme.")
    all(v -> v.n == dest.n, l.vecs) || throw(DimensionMismatch("DeviceVector broadcast: lengths differ"))
    if _isconst(l)
        return fill!(dest, l.c0)
    end
    # ncme_vec_lincomb takes up to 8 terms and its output may alias any input: chunk, accumulating through dest
    coefs, vecs = copy(l.coefs), copy(l.vecs)
    first_chunk = true
    while !isempty(coefs)
        k = min(length(coefs), first_chunk ? 8 : 7)
        cs, vs = coefs[1:k], vecs[1:k]
        first_chunk || (pushfirst!(cs, 1.0); pushfirst!(vs, dest))
        lincomb!(dest, cs, vs)
        coefs, vecs = coefs[k+1:end], vecs[k+1:end]
        first_chunk = false
    end
    l.c0 == 0.0 || shift!(dest, l.c0)
    dest
end
Base.copyto!(dest::DeviceVector, bc::Broadcasted{<:Broadcast.AbstractArrayStyle{0}}) = fill!(dest, bc[CartesianIndex()])
# error-norm and instability hooks handed to DE.init for DiffEq algorithms (the defaults loop over elements)
_internalnorm(u::DeviceVector, t) = sqrt(LinearAlgebra.dot(u, u) / max(length(u), 1))
_internalnorm(u::Number, t) = abs(u)
_unstable_check(dt, u::DeviceVector, p, t) = hasnonfinite(u)
_unstable_check(dt, u, p, t) = any(isnan, u)

# ------------------------------------------------------------------------------------------------ StateSpaceSparse
# reference: src/statespace/sparse/sparsestatespace.jl:22-40 (struct), :103-144 (ctors)
mutable struct StateSpaceSparseB200{NS,NR} <: NumCME.AbstractStateSpaceSparse{NS,NR,Int64,UInt32}
    ctx::Context
    h::Ptr{Cvoid}
    stoich_matrix::Matrix{Int64}
end
function StateSpaceSparseB200(stoich::Matrix{<:Integer}, initstates::Vector; ctx::Context = default_ctx())
    S = Matrix{Int64}(stoich)                      # column-major == the ABI's reaction-major layout
    ns, nr = size(S)
    flat = Int64[x for st in initstates for x in st]
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ncme_space_create, libncme), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Int64, Ptr{Int64}, Ref{Ptr{Cvoid}}),
        ctx.h, ns, nr, S, length(initstates), flat, r))
    sp = StateSpaceSparseB200{ns,nr}(ctx, r[], S)
    finalizer(x -> ccall((:ncme_space_destroy, libncme), Cint, (Ptr{Cvoid},), x.h), sp)
end
StateSpaceSparseB200(stoich::Matrix{<:Integer}, x0::Vector{<:Integer}; kw...) = StateSpaceSparseB200(stoich, [x0]; kw...)

get_stoich_matrix(sp::StateSpaceSparseB200) = sp.stoich_matrix
function get_state_count(sp::StateSpaceSparseB200)
    r = Ref{Int64}(0); check(ccall((:ncme_space_state_count, libncme), Cint, (Ptr{Cvoid}, Ref{Int64}), sp.h, r)); Int(r[])
end
get_sink_count(sp::StateSpaceSparseB200{NS,NR}) where {NS,NR} = UInt32(NR)
function get_states(sp::StateSpaceSparseB200{NS,NR}, first::Integer = 0, count::Integer = -1) where {NS,NR}   # sparsestatespace.jl:69
    n = count < 0 ? get_state_count(sp) - first : count
    out = Vector{MVector{NS,Int64}}(undef, n)      # contiguous n*NS Int64: exactly the ABI's state-major layout
    n > 0 && check(ccall((:ncme_space_download_states, libncme), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}), sp.h, first, n, out))
    out
end
function _connectivity(sp::StateSpaceSparseB200{NS,NR}) where {NS,NR}
    n = get_state_count(sp)
    sc = Vector{MVector{NR,UInt32}}(undef, n); kc = Vector{MVector{NR,UInt32}}(undef, n)
    n > 0 && check(ccall((:ncme_space_download_connectivity, libncme), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}), sp.h, 0, n, sc, kc))
    sc, kc
end
get_state_connectivity(sp::StateSpaceSparseB200) = _connectivity(sp)[1]
get_sink_connectivity(sp::StateSpaceSparseB200) = _connectivity(sp)[2]
get_statedict(sp::StateSpaceSparseB200) = Dict(x => UInt32(i) for (i, x) in enumerate(get_states(sp)))   # host materialisation
function Base.getproperty(sp::StateSpaceSparseB200, f::Symbol)          # `space.states` etc. as in the reference struct
    f === :states && return get_states(sp)
    f === :state2idx && return get_statedict(sp)
    f === :state_connectivity && return get_state_connectivity(sp)
    f === :sink_connectivity && return get_sink_connectivity(sp)
    f === :sink_count && return get_sink_count(sp)
    getfield(sp, f)
end

# expand!(statespace, expansionlevel; onlyreactions = [])        sparsestatespace.jl:153
function expand!(sp::StateSpaceSparseB200, expansionlevel::Integer; onlyreactions = [])
    only = Int32[r for r in onlyreactions]
    check(ccall((:ncme_space_expand, libncme), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}), sp.h, expansionlevel, length(only), only))
    nothing
end
# deleteat!(statespace, ids)                                      sparsestatespace.jl:276
function deleteat!(sp::StateSpaceSparseB200, ids::Vector{T}) where {T<:Integer}
    v = Int64[i for i in ids]
    check(ccall((:ncme_space_delete, libncme), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}), sp.h, length(v), v))
    nothing
end

# sum(p, dims) for a device-resident probability vector over a device space        fspvector.jl:66-99
# (reduced states in the order of their first occurrence, like the reference; values accumulated with fp64 atomics)
function Base.sum(p::DeviceVector, sp::StateSpaceSparseB200{NS,NR}, dims::AbstractVector{<:Integer}) where {NS,NR}
    d = Int32[x for x in unique(dims)]
    nkeep = NS - length(d)
    nred = Ref{Int64}(0)
    check(ccall((:ncme_space_marginal, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, Int64, Ref{Int64}, Ptr{Int64}, Ptr{Float64}),
                sp.h, p.ptr, length(d), d, 0, nred, C_NULL, C_NULL))                       # size query
    states = Vector{MVector{nkeep,Int64}}(undef, nred[]); vals = zeros(Float64, nred[])
    nred[] > 0 && check(ccall((:ncme_space_marginal, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, Int64, Ref{Int64}, Ptr{Cvoid}, Ptr{Float64}),
                sp.h, p.ptr, length(d), d, nred[], nred, states, vals))
    FspVectorSparse(states, vals)
end

# ------------------------------------------------------------------------------------------------ FspMatrixSparse
# SURVEY H1: a JointTimeVaryingPropensity f(t,x,p) that is numerically c(t) g(x) on the current states is handed to the
# library as separable (g = f(t_ref, .), c(t) = f(t, x*, p) / g(x*)): one host call per right-hand side instead of n
# calls + an upload per distinct t (`_update_sparsematrix!`, fspsparsematrix.jl:154-166).  Same test as the executed
# Python mirror (numcme.jl_b200/fspmatrix.py: detect_rank1); the product form is re-checked on sentinel states at
# every t actually used; a mismatch raises SeparabilityError, which solve() answers by repeating the segment on the
# exact joint path.
struct SeparabilityError <: Exception
    reaction::Int
    t::Float64
end
const _PROBE_TIMES = (0.0, 0.7310585786300049, 19.098300562505255, 738.90560989306495, 5459.8150033144236, 28813.3)
function detect_rank1(f, states, θ; rtol = 1e-12)
    g = nothing
    for t in _PROBE_TIMES
        v = Float64[f(t, x, θ) for x in states]
        all(isfinite, v) || return nothing
        if g === nothing
            any(!iszero, v) && (g = v)
            continue
        end
        supp = g .!= 0.0
        any(!iszero, v[.!supp]) && return nothing
        ratio = v[supp] ./ g[supp]
        (maximum(abs, ratio) > 0 && maximum(abs, ratio .- ratio[1]) > rtol * max(abs(ratio[1]), 1e-300)) && return nothing
    end
    g === nothing && return nothing
    nz = sortperm(abs.(g), rev = true)[1:count(!iszero, g)]
    sent = unique([nz[1]; [nz[k] for k in (length(nz) ÷ 3, 2 * length(nz) ÷ 3, length(nz)) if 1 < k <= length(nz)]])
    return g, sent
end

# reference: src/fspmatrix/sparse/fspsparsematrix.jl:9-27 (struct), :47-108 (ctor)
mutable struct FspMatrixSparseB200{NS,NR} <: NumCME.AbstractFspMatrix
    ctx::Context
    h::Ptr{Cvoid}
    comm::Union{Nothing,Comm}
    parameters::Vector{Any}
    states::Vector{MVector{NS,Int64}}
    rowcount::Int64
    colcount::Int64
    propensities::Vector{<:Propensity}
    kinds::Vector{Int32}
    t_cache::Float64
    coef::Vector{Float64}
    tfactors::Dict{Int,Any}      # reaction => t -> c_r(t) for every reaction the library treats as separable
end
# comm !== nothing: this rank's row block only (K8).  The host evaluates the state factors
# of its own rows + predecessor window only (ncme_matrix_shard_window / ncme_matrix_create_window): evaluation and
# upload shrink with the number of ranks.
function FspMatrixSparseB200(space::StateSpaceSparseB200{NS,NR}, props::Vector{<:Propensity}; parameters = [],
                             detect_separable::Bool = true, comm::Union{Nothing,Comm} = nothing) where {NS,NR}
    states = get_states(space)                      # the host copy the reference keeps (`deepcopy(space.states)`, :97)
    n = length(states)
    kinds = Int32[!istimevarying(a) ? 0 : (istimeseparable(a) ? 1 : 2) for a in props]
    windowed = comm !== nothing && comm.nranks > 1 && n > 0
    lo, hi = 0, n
    if windowed
        w = zeros(Int64, 4)
        check(ccall((:ncme_matrix_shard_window, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}), space.h, comm.h, w))
        lo, hi = Int(w[3]), Int(w[4])
    end
    nw = hi - lo
    G = zeros(Float64, max(nw, 1), NR)              # column r = state factor of reaction r over the window: the ABI's reaction-major layout
    tfactors = Dict{Int,Any}()
    for (r, a) in enumerate(props)                  # host evaluation of the opaque closures, once per (state, reaction) (:129)
        kinds[r] == 0 && (for i in 1:nw; G[i, r] = a.f(states[lo+i], parameters); end)
        kinds[r] == 1 && (for i in 1:nw; G[i, r] = a.statefactor(states[lo+i], parameters); end; tfactors[r] = t -> a.tfactor(t, parameters))
        if kinds[r] == 2 && detect_separable
            found = detect_rank1(a.f, states, parameters)
            if found !== nothing
                g, sent = found
                G[1:nw, r] .= g[lo+1:hi]; kinds[r] = 1
                tfactors[r] = function (t)
                    c = a.f(t, states[sent[1]], parameters) / g[sent[1]]
                    for k in sent[2:end]
                        ck = a.f(t, states[k], parameters) / g[k]
                        abs(ck - c) > 1e-9 * max(abs(c), abs(ck), 1e-300) && throw(SeparabilityError(r, t))
                    end
                    c
                end
            end
        end
    end
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    if windowed
        check(ccall((:ncme_matrix_create_window, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Int64, Int64, Ref{Ptr{Cvoid}}),
                    space.h, comm.h, kinds, G, lo, hi, ref))
    else
        check(ccall((:ncme_matrix_create_sharded, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                    space.h, _commh(comm), kinds, G, ref))
    end
    A = FspMatrixSparseB200{NS,NR}(space.ctx, ref[], comm, Vector{Any}(parameters), states, n + NR, n + NR, props, kinds, -Inf, ones(NR), tfactors)
    finalizer(x -> ccall((:ncme_matrix_destroy, libncme), Cint, (Ptr{Cvoid},), x.h), A)
end
# Rebuild after an adapt! (fspsolve.jl:176 rebuilds from scratch): only the states appended since `previous` was built
# (always the tail of the state list) are evaluated on the host, the factors of the surviving states are carried over
# on the device (SURVEY.md H8).  Falls back to the full constructor when `previous` is not the matrix the space was
# last assembled into or a joint propensity found to be c(t) g(x) stops being so on the new states.
function FspMatrixSparseB200(space::StateSpaceSparseB200{NS,NR}, previous::FspMatrixSparseB200{NS,NR}; detect_separable::Bool = true) where {NS,NR}
    props, parameters, comm = previous.propensities, previous.parameters, previous.comm
    full() = FspMatrixSparseB200(space, props; parameters, detect_separable, comm)
    nk = Ref{Int64}(0); nn = Ref{Int64}(0)
    check(ccall((:ncme_space_new_count, libncme), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), space.h, nk, nn))
    nkept, nnew = Int(nk[]), Int(nn[])
    any(previous.kinds[r] == 1 && !istimeseparable(props[r]) for r in 1:NR) && return full()   # rank-1 joint reactions: re-detect
    (nkept == 0 || (comm !== nothing && comm.nranks > 1)) && return full()                    # windowed builds cannot seed an incremental one
    states = get_states(space)
    n = length(states)
    G = zeros(Float64, max(nnew, 1), NR)
    for (r, a) in enumerate(props), i in 1:nnew
        previous.kinds[r] == 0 && (G[i, r] = a.f(states[nkept+i], parameters))
        previous.kinds[r] == 1 && (G[i, r] = a.statefactor(states[nkept+i], parameters))
    end
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    code = ccall((:ncme_matrix_create_incremental, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                 space.h, C_NULL, previous.h, previous.kinds, G, ref)
    code == 0 || return full()
    A = FspMatrixSparseB200{NS,NR}(space.ctx, ref[], nothing, previous.parameters, states, n + NR, n + NR, props, copy(previous.kinds), -Inf, ones(NR), previous.tfactors)
    finalizer(x -> ccall((:ncme_matrix_destroy, libncme), Cint, (Ptr{Cvoid},), x.h), A)
end
get_parameters(A::FspMatrixSparseB200) = A.parameters
get_states(A::FspMatrixSparseB200) = A.states
get_rowcount(A::FspMatrixSparseB200) = A.rowcount
get_colcount(A::FspMatrixSparseB200) = A.colcount
get_propensities(A::FspMatrixSparseB200) = A.propensities
size(A::FspMatrixSparseB200) = (A.rowcount, A.colcount)
function size(A::FspMatrixSparseB200, dim::Integer)                     # fspsparsematrix.jl:181-186
    !(1 <= dim <= 2) && throw(ArgumentError("Second argument must be either 1 or 2."))
    dim == 1 ? A.rowcount : A.colcount
end
# {row_lo, row_hi, halo_lo, halo_hi, n_global, interior_begin, interior_end, nranks} of this rank's shard
function shard_info(A::FspMatrixSparseB200)
    info = zeros(Int64, 8)
    check(ccall((:ncme_matrix_shard_info, libncme), Cint, (Ptr{Cvoid}, Ptr{Int64}), A.h, info))
    info
end

# time-dependent pieces: separable factors are one host scalar per reaction per call (:204); joint reactions are
# re-evaluated on the host when t changes (:206-212) and uploaded
function _prepare!(A::FspMatrixSparseB200, t::Real)
    θ = A.parameters
    for (r, tf) in A.tfactors                       # separable reactions and joint ones found to be rank-1 (kinds[r] == 1)
        A.coef[r] = tf(t)
    end
    if t != A.t_cache
        A.t_cache = t
        lo, hi = 0, length(A.states)
        if A.comm !== nothing                      # row-sharded: this rank's rows + predecessor window only
            i = shard_info(A)
            lo, hi = Int(i[1] - i[3]), Int(i[2] + i[4])
        end
        for (r, a) in enumerate(A.propensities)
            if A.kinds[r] == 2
                vals = Float64[a.f(t, A.states[k], θ) for k in lo+1:hi]
                check(ccall((:ncme_matrix_set_joint_values, libncme), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), A.h, r, vals))
            end
        end
    end
    A.coef
end

# matvec!(out, t, A, v)   fspsparsematrix.jl:196   /   matvecadd!(out, t, A, v)   :226
# Sharded matrices: vectors are this rank's slice [rows | R sinks]; inputs need halo margins (ShardedVector below).
function _apply!(out, t, A::FspMatrixSparseB200, v, beta::Float64)
    N = A.comm === nothing ? A.rowcount : (i = shard_info(A); i[2] - i[1] + (A.rowcount - i[5]))
    (length(out) == N && length(v) == N) || throw(DimensionMismatch("matvec!: vector lengths must equal $N"))
    coef = _prepare!(A, t)
    if out isa DeviceVector && v isa DeviceVector            # device-resident: one kernel launch, asynchronous
        check(ccall((:ncme_matvec, libncme), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}, Float64), A.h, coef, v.ptr, out.ptr, beta))
    else                                                      # host Vector{Float64} / contiguous views: H2D + kernel + D2H
        check(ccall((:ncme_matvec_host, libncme), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64), A.h, coef, v, out, beta))
    end
    nothing
end
matvec!(out, t, A::FspMatrixSparseB200, v) = _apply!(out, t, A, v, 0.0)
matvecadd!(out, t, A::FspMatrixSparseB200, v) = _apply!(out, t, A, v, 1.0)
matvec(t, A::FspMatrixSparseB200, v) = (w = similar(v); matvec!(w, t, A, v); w)
*(A::FspMatrixSparseB200, v::Vector{Float64}) = matvec(0.0, A, v)
*(A::FspMatrixSparseB200, v::DeviceVector) = matvec(0.0, A, v)
# addition (the reference imports mul! but defines no method, fspsparsematrix.jl:1): mul! at the cached time
mul!(y, A::FspMatrixSparseB200, x) = (matvec!(y, isfinite(A.t_cache) ? A.t_cache : 0.0, A, x); y)

# Local slice [rows row_lo..row_hi | R sink entries] of an FSP vector with the halo margins a sharded matvec input needs
struct ShardedVector
    buf::DeviceVector
    v::DeviceVector               # the local slice (what matvec! / ncme_solve_segment take)
    lo::Int; hi::Int; nglobal::Int
end
function ShardedVector(A::FspMatrixSparseB200{NS,NR}) where {NS,NR}
    i = shard_info(A)
    nloc = Int(i[2] - i[1])
    buf = fill!(DeviceVector(A.ctx, Int(i[3]) + nloc + NR + Int(i[4])), 0.0)
    ShardedVector(buf, view(buf, Int(i[3])+1:Int(i[3])+nloc+NR), Int(i[1]), Int(i[2]), Int(i[5]))
end
function load!(s::ShardedVector, pfull::DeviceVector, sinks::Vector{Float64})
    nloc = s.hi - s.lo
    nloc > 0 && copyto!(view(s.v, 1:nloc), view(pfull, s.lo+1:s.hi))
    copyto!(view(s.v, nloc+1:nloc+length(sinks)), sinks)
    s
end
function gather(s::ShardedVector, comm::Union{Nothing,Comm})           # p (all states) on every rank
    full = DeviceVector(s.v.ctx, s.nglobal)
    if comm !== nothing && comm.nranks > 1
        cuts = shard_bounds(s.nglobal, comm.nranks)
        counts = Int64[cuts[r+1] - cuts[r] for r in 1:comm.nranks]; displs = Int64[cuts[r] for r in 1:comm.nranks]
        check(ccall((:ncme_comm_allgatherv, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}),
                    comm.h, s.v.ptr, full.ptr, counts, displs))
    elseif s.nglobal > 0
        copyto!(full, view(s.v, 1:s.nglobal))
    end
    full
end

# ------------------------------------------------------------------------------------------------ ForwardSensFspMatrixSparse
# reference: src/forwardsensfspmatrix/forwardsensfspmatrixsparse/sensfspmatrixsparse.jl:9-22 (struct), :31-95 (ctor),
# :97-142 (matvec!).  Vector layout [p; s_1; ...; s_P], each block n + R long.  A(t) is read once for all P + 1 blocks
# and every dA/dtheta entry is streamed once (ncme_sens_matvec, one fused launch).
mutable struct ForwardSensFspMatrixSparseB200{NS,NR} <: NumCME.ForwardSensFspMatrix
    fspmatrix::FspMatrixSparseB200{NS,NR}
    h::Ptr{Cvoid}
    propensity_gradients::Vector{<:PropensityGradient}
    entries::Vector{Tuple{Int,Int}}          # (reaction, parameter) pairs of the gradient sparsity pattern, parameter-major
    dcoef::Vector{Float64}
end
function ForwardSensFspMatrixSparseB200(model::CmeModelWithSensitivity, space::StateSpaceSparseB200{NS,NR}) where {NS,NR}
    θ = get_parameters(model)
    # the derivative entries follow the user's classification of every reaction: no separability detection here
    A = FspMatrixSparseB200(space, get_propensities(model); parameters = θ, detect_separable = false)
    grads = get_propensity_gradients(model)
    pattern = get_gradient_sparsity_patterns(model)
    P = get_parameter_count(model)
    ents = [(r, ip) for ip in 1:P for r in 1:NR if pattern[r, ip]]       # the reference's order (nzrange over CSC columns)
    n = length(A.states)
    dvals = zeros(Float64, max(n, 1), max(length(ents), 1))               # entry-major nentries x n for the ABI
    for (e, (r, ip)) in enumerate(ents)
        g = grads[r]
        A.kinds[r] == 0 && (for i in 1:n; dvals[i, e] = g.pardiffs[ip](A.states[i], θ); end)
        A.kinds[r] == 1 && (for i in 1:n; dvals[i, e] = g.statefactor_pardiffs[ip](A.states[i], θ); end)
    end
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ncme_sensmatrix_create, libncme), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                A.h, P, length(ents), Int32[r for (r, _) in ents], Int32[ip for (_, ip) in ents], dvals, ref))
    SA = ForwardSensFspMatrixSparseB200{NS,NR}(A, ref[], grads, ents, zeros(max(length(ents), 1)))
    finalizer(x -> ccall((:ncme_sensmatrix_destroy, libncme), Cint, (Ptr{Cvoid},), x.h), SA)
end
# Rebuild after an adapt! (forwardsenscmesparse.jl:140 rebuilds from scratch inside the loop): the plain matrix is rebuilt
# incrementally from `previous.fspmatrix`, the parameter derivatives are evaluated on the appended states only and the
# survivors' rows are carried over on the device (ncme_sensmatrix_create_incremental).  Falls back to the full
# constructor whenever the plain matrix did (too few kept states, a space assembled elsewhere in between).
function ForwardSensFspMatrixSparseB200(model::CmeModelWithSensitivity, space::StateSpaceSparseB200{NS,NR},
                                        previous::ForwardSensFspMatrixSparseB200{NS,NR}) where {NS,NR}
    full() = ForwardSensFspMatrixSparseB200(model, space)
    grads = get_propensity_gradients(model)
    grads === previous.propensity_gradients || return full()
    nk = Ref{Int64}(0); nn = Ref{Int64}(0)
    check(ccall((:ncme_space_new_count, libncme), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), space.h, nk, nn))
    nkept, nnew = Int(nk[]), Int(nn[])
    nkept == 0 && return full()
    A = FspMatrixSparseB200(space, previous.fspmatrix; detect_separable = false)
    θ = get_parameters(model); P = get_parameter_count(model); ents = previous.entries
    dvals = zeros(Float64, max(nnew, 1), max(length(ents), 1))            # entry-major nentries x n_new for the ABI
    for (e, (r, ip)) in enumerate(ents), i in 1:nnew
        A.kinds[r] == 0 && (dvals[i, e] = grads[r].pardiffs[ip](A.states[nkept+i], θ))
        A.kinds[r] == 1 && (dvals[i, e] = grads[r].statefactor_pardiffs[ip](A.states[nkept+i], θ))
    end
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    code = ccall((:ncme_sensmatrix_create_incremental, libncme), Cint,
                 (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                 A.h, previous.h, P, length(ents), Int32[r for (r, _) in ents], Int32[ip for (_, ip) in ents], dvals, ref)
    code == 0 || return full()               # `A` was built from scratch (not from previous.fspmatrix): start over
    SA = ForwardSensFspMatrixSparseB200{NS,NR}(A, ref[], grads, ents, zeros(max(length(ents), 1)))
    finalizer(x -> ccall((:ncme_sensmatrix_destroy, libncme), Cint, (Ptr{Cvoid},), x.h), SA)
end
get_propensity_gradients(SA::ForwardSensFspMatrixSparseB200) = SA.propensity_gradients

# time factors of A and of the derivative entries at t (sensfspmatrixsparse.jl:124-139)
function _prepare!(SA::ForwardSensFspMatrixSparseB200, t::Real)
    A = SA.fspmatrix
    θ = A.parameters
    coef = _prepare!(A, t)
    for (e, (r, ip)) in enumerate(SA.entries)
        if A.kinds[r] == 1                                               # d tfactor / d theta_ip   (:124-132)
            SA.dcoef[e] = SA.propensity_gradients[r].tfactor_pardiffs[ip](t, θ)
        elseif A.kinds[r] == 2                                           # joint: d f / d theta_ip over all states (:134-139, see Q6)
            vals = Float64[SA.propensity_gradients[r].pardiffs[ip](t, x, θ) for x in A.states]
            check(ccall((:ncme_sensmatrix_set_joint_values, libncme), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), SA.h, e - 1, vals))
        end
    end
    coef
end
# matvec!(out, t, SA, vs)                                         sensfspmatrixsparse.jl:97
function matvec!(out::DeviceVector, t::Real, SA::ForwardSensFspMatrixSparseB200, vs::DeviceVector)
    A = SA.fspmatrix
    P = length(A.parameters)
    (length(out) == (P + 1) * A.rowcount && length(vs) == (P + 1) * A.rowcount) ||
        throw(DimensionMismatch("matvec!: expected vectors of length $((P + 1) * A.rowcount)"))
    coef = _prepare!(SA, t)
    check(ccall((:ncme_sens_matvec, libncme), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}),
                SA.h, coef, SA.dcoef, vs.ptr, out.ptr))
    nothing
end
function matvec!(out::Vector{Float64}, t::Real, SA::ForwardSensFspMatrixSparseB200, vs::Vector{Float64})   # host vectors: staged
    ctx = SA.fspmatrix.ctx
    dv = DeviceVector(ctx, vs); dw = DeviceVector(ctx, length(out))
    matvec!(dw, t, SA, dv)
    copyto!(out, Array(dw))
    nothing
end

# ------------------------------------------------------------------------------------------------ adapters
# init!(space, adapter, p, t, fsptol)            rstepadapters.jl:23 / :74
function init!(space::StateSpaceSparseB200, adapter::Union{RStepAdapter,SelectiveRStepAdapter}, p::DeviceVector, t, fsptol)
    expand!(space, adapter.initial_step_count)
    _grow(p, get_state_count(space))
end
function _grow(p::DeviceVector, n::Integer)                              # append!(p, zeros(...))
    n == p.n && return p
    q = fill!(DeviceVector(p.ctx, n), 0.0)
    p.n > 0 && copyto!(view(q, 1:p.n), p)
    q
end
# prune rule of adapt! (sortperm + cumsum + threshold + deleteat!, rstepadapters.jl:41-45 / :93-96) on the device;
# returns the compaction of every vector in `vecs` (p first: it carries the mass the rule looks at)
function _prune!(space::StateSpaceSparseB200, vecs::Vector{DeviceVector}, thr::Float64, strict::Bool)
    dc = Ref{Int64}(0)
    check(ccall((:ncme_space_prune_by_mass, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Cint, Ref{Int64}),
        space.h, vecs[1].ptr, thr, strict, dc))
    dc[] > 0 || return vecs
    map(vecs) do v
        q = DeviceVector(v.ctx, get_state_count(space))
        check(ccall((:ncme_space_compact_vector, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), space.h, v.ptr, q.ptr))
        q
    end
end
# adapt!(space, adapter, p, sinks, t, tend, fsptol; integrator)      rstepadapters.jl:35 / :86
function adapt!(space::StateSpaceSparseB200, adapter::Union{RStepAdapter,SelectiveRStepAdapter}, p::DeviceVector,
    sinks::Vector{Float64}, t, tend, fsptol; dsinks::Union{Nothing,Vector{Float64}} = nothing)
    strict = adapter isa SelectiveRStepAdapter
    adapter.dropstates && p.n > 0 && (p = _prune!(space, DeviceVector[p], 1.0 - t * fsptol / tend, strict)[1])
    strict ? expand!(space, adapter.max_step_count; onlyreactions = findall(dsinks .> 0)) : expand!(space, adapter.max_step_count)
    _grow(p, get_state_count(space))
end
# sensitivity adapters: init!(space, adapter, p, S, t, fsptol) / adapt!(space, adapter, p, S, sinks, dsinks, t, tend, fsptol)
# fsspaceadapterssparse.jl:21-30 / :37-60   (vecs = [p, S_1..S_P], all compacted / grown alike)
function init!(space::StateSpaceSparseB200, adapter::ForwardSensRStepAdapter, vecs::Vector{DeviceVector}, t, fsptol)
    expand!(space, adapter.initial_step_count)
    DeviceVector[_grow(v, get_state_count(space)) for v in vecs]
end
function adapt!(space::StateSpaceSparseB200, adapter::ForwardSensRStepAdapter, vecs::Vector{DeviceVector}, t, tend, fsptol)
    adapter.dropstates && vecs[1].n > 0 && (vecs = _prune!(space, vecs, 1.0 - t * fsptol / tend, false))
    expand!(space, adapter.max_step_count)
    DeviceVector[_grow(v, get_state_count(space)) for v in vecs]
end

# ------------------------------------------------------------------------------------------------ solve
# The B200 path is selected by wrapping the reference's algorithm object: no method of the reference is overwritten
# (its signatures, fspsolve.jl:10-14 / :105-112 and forwardsenscmesparse.jl:99-108, stay the only ones for their types).
struct OnB200{T}
    alg::T                       # AdaptiveFspSparse | AdaptiveForwardSensFspSparse | AbstractODEAlgorithm | Nothing
    ctx::Context
    comm::Union{Nothing,Comm}
end
OnB200(alg; ctx::Context = default_ctx(), comm::Union{Nothing,Comm} = nothing) = OnB200{typeof(alg)}(alg, ctx, comm)

struct SolveOpts
    rtol::Float64; atol::Float64; event_slope::Float64; check_event::Cint; save_every_step::Cint
    nsave::Cint; save_t::Ptr{Float64}; h_init::Float64; max_steps::Int64; method::Cint
end
mutable struct SolveStats
    t_final::Float64; h_last::Float64; event_hit::Cint; nsaved::Cint
    steps::Int64; rejected::Int64; rhs_evals::Int64; launches::Int64
    SolveStats() = new(0, 0, 0, 0, 0, 0, 0, 0)
end
# Callbacks run inside C frames: no exception may cross them.  A failing callback stores its exception in the box,
# asks the library to stop (ncme_request_abort; the segment returns NCME_ERR_ABORTED) and solve rethrows it.
mutable struct CallbackBox
    prepare::Any                 # t -> (coef, dcoef)  time factors of the matrix (and of the derivative entries)
    ncoef::Int
    ndcoef::Int
    len::Int                     # length of the vector handed to the save callback
    saved::Vector{Tuple{Float64,Vector{Float64}}}
    err::Any
end
function _coef_cb(t::Float64, coef::Ptr{Float64}, user::Ptr{Cvoid})::Cvoid
    box = unsafe_pointer_to_objref(user)::CallbackBox
    try
        c, dc = box.prepare(t)
        for r in 1:box.ncoef; unsafe_store!(coef, c[r], r); end
        for e in 1:box.ndcoef; unsafe_store!(coef, dc[e], box.ncoef + e); end
    catch err
        box.err = err
        ccall((:ncme_request_abort, libncme), Cvoid, ())
    end
    nothing
end
function _save_cb(t::Float64, u::Ptr{Float64}, user::Ptr{Cvoid})::Cvoid
    box = unsafe_pointer_to_objref(user)::CallbackBox
    try
        push!(box.saved, (t, copy(unsafe_wrap(Array, u, box.len))))
    catch err
        box.err = err
        ccall((:ncme_request_abort, libncme), Cvoid, ())
    end
    nothing
end
_saveat(saveat, tspan) = saveat isa Number ? collect(Float64, tspan[1]:saveat:tspan[2]) : collect(Float64, saveat)
# one call of ncme_solve_segment / ncme_sens_solve_segment (native BDF/GMRES; the fused step kernel below 2e6 rows)
function _segment!(entry::Symbol, h::Ptr{Cvoid}, box::CallbackBox, u::DeviceVector, t0, t1, sv::Vector{Float64}, odertol, odeatol,
                   event_slope::Union{Nothing,Float64})
    ccoef = @cfunction(_coef_cb, Cvoid, (Float64, Ptr{Float64}, Ptr{Cvoid}))
    csave = @cfunction(_save_cb, Cvoid, (Float64, Ptr{Float64}, Ptr{Cvoid}))
    stats = SolveStats()
    opts = SolveOpts(odertol, odeatol, event_slope === nothing ? 0.0 : event_slope, event_slope === nothing ? 0 : 1,
                     isempty(sv) ? 1 : 0, length(sv), pointer(sv), 0.0, 0, 1)
    box.err = nothing
    code = GC.@preserve box sv begin
        if entry === :plain
            ccall((:ncme_solve_segment, libncme), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Ptr{Cvoid}, Ref{SolveOpts}, Ref{SolveStats}),
                  h, ccoef, csave, pointer_from_objref(box), t0, t1, u.ptr, opts, stats)
        else
            ccall((:ncme_sens_solve_segment, libncme), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Ptr{Cvoid}, Ref{SolveOpts}, Ref{SolveStats}),
                  h, ccoef, csave, pointer_from_objref(box), t0, t1, u.ptr, opts, stats)
        end
    end
    box.err === nothing || throw(box.err)       # the exception of a user closure (or SeparabilityError) raised in a callback
    check(code)
    stats
end
_plainbox(A::FspMatrixSparseB200{NS,NR}, len) where {NS,NR} =
    CallbackBox(t -> (_prepare!(A, t), Float64[]), NR, 0, len, Tuple{Float64,Vector{Float64}}[], nothing)
_newoutput(::Val{NS}) where {NS} = FspOutputSparse{NS,Int64,Float64}(t = Float64[], p = FspVectorSparse{NS,Int64,Float64}[], sinks = Vector{Float64}[])
# values of p0 placed on the states of `space` (duplicates / negative states are dropped by the space, :221)
function _place(space::StateSpaceSparseB200, states0, vals0)
    d = get_statedict(space)
    a = zeros(Float64, get_state_count(space))
    for (x, v) in zip(states0, vals0)
        i = get(d, x, UInt32(0)); i > 0 && (a[i] = v)
    end
    a
end

# ---- fixed state space: solve(model, p0, tspan, ode_method; saveat, fsptol, odeatol, odertol)      fspsolve.jl:10-41
function solve(model::CmeModel, p0::FspVectorSparse{NS,IntT,RealT}, tspan::Union{Vector,Tuple},
    on::OnB200{<:Union{Nothing,AbstractODEAlgorithm}}; saveat = [], fsptol::AbstractFloat = 1.0e-6,
    odeatol::AbstractFloat = 1.0e-6, odertol::AbstractFloat = 1.0e-4, detect_separable::Bool = true) where {NS,IntT<:Integer,RealT<:AbstractFloat}
    ctx, comm = on.ctx, on.comm
    space = StateSpaceSparseB200(model.stoich_matrix, p0.states; ctx)
    R = Int(get_sink_count(space)); n = get_state_count(space)
    A = FspMatrixSparseB200(space, model.propensities; parameters = model.parameters, detect_separable, comm)
    out = _newoutput(Val(NS))
    pfull = DeviceVector(ctx, _place(space, p0.states, p0.values))
    if on.alg !== nothing                                   # a DifferentialEquations.jl algorithm on the device vector
        comm === nothing || throw(ArgumentError("DifferentialEquations.jl algorithms run on one GPU; use ode_method = nothing with comm"))
        u0 = fill!(DeviceVector(ctx, n + R), 0.0); copyto!(view(u0, 1:n), pfull)
        prob = DE.ODEProblem((du, u, θ, t) -> (matvec!(du, t, A, u); nothing), u0, tspan, model.parameters)
        sol = DE.solve(prob, on.alg; abstol = odeatol, reltol = odertol, saveat = saveat, internalnorm = _internalnorm, unstable_check = _unstable_check)
        for (t, u) in zip(sol.t, sol.u)
            uu = Array(u)
            push!(out.t, t); push!(out.p, FspVectorSparse(A.states, uu[1:n])); push!(out.sinks, uu[n+1:end])
        end
        return out
    end
    sv = _saveat(saveat, tspan)
    while true
        s = ShardedVector(A); load!(s, pfull, zeros(R))
        nloc = s.hi - s.lo
        box = _plainbox(A, nloc + R)
        try
            _segment!(:plain, A.h, box, s.v, Float64(tspan[1]), Float64(tspan[2]), sv, odertol, odeatol, nothing)
        catch err                                            # a detected c(t) g(x) form broke down: exact joint path
            (err isa SeparabilityError && detect_separable) || rethrow()
            detect_separable = false
            A = FspMatrixSparseB200(space, model.propensities; parameters = model.parameters, detect_separable = false, comm)
            continue
        end
        for (t, uu) in box.saved                             # sharded: every rank holds its own rows of each slice
            push!(out.t, t); push!(out.p, FspVectorSparse(A.states[s.lo+1:s.hi], uu[1:nloc])); push!(out.sinks, uu[nloc+1:end])
        end
        return out
    end
end

# ---- adaptive: solve(model, p0, tspan, AdaptiveFspSparse; saveat, fsptol, odeatol, odertol, verbose)   fspsolve.jl:105-197
function solve(model::CmeModel, p0::FspVectorSparse{NS,IntT,RealT}, tspan::Tuple{AbstractFloat,AbstractFloat},
    on::OnB200{AdaptiveFspSparse}; saveat = [], fsptol::AbstractFloat = 1.0e-6, odeatol::AbstractFloat = 1.0e-6,
    odertol::AbstractFloat = 1.0e-4, verbose::Bool = false, detect_separable::Bool = true) where {NS,IntT<:Integer,RealT<:AbstractFloat}
    ctx, comm, alg = on.ctx, on.comm, on.alg
    tstart, tend = min(tspan...), max(tspan...)
    sv = _saveat(saveat, tspan)
    adapter = alg.space_adapter
    space = StateSpaceSparseB200(model.stoich_matrix, p0.states; ctx)
    R = Int(get_sink_count(space))
    p = init!(space, adapter, DeviceVector(ctx, _place(space, p0.states, p0.values)), tstart, fsptol)   # all states, replicated
    sinks = zeros(R); tnow = tstart
    A = FspMatrixSparseB200(space, model.propensities; parameters = model.parameters, detect_separable, comm)
    out = _newoutput(Val(NS))
    while tnow < tend
        n = get_state_count(space)
        if alg.ode_method !== nothing
            # a DifferentialEquations.jl algorithm: the reference's own segment (fspsolve.jl:138-161) with `u` in HBM
            comm === nothing || throw(ArgumentError("DifferentialEquations.jl algorithms run on one GPU; use ode_method = nothing with comm"))
            u0 = DeviceVector(ctx, n + R); copyto!(view(u0, 1:n), p); copyto!(view(u0, n+1:n+R), sinks)
            cond(u, t, integrator) = sum(u[n+1:n+R]) - fsptol * t / tend
            cb = DE.ContinuousCallback(cond, integ -> DE.terminate!(integ); save_positions = (false, false), interp_points = 100, abstol = eps())
            prob = DE.ODEProblem((du, u, θ, t) -> (matvec!(du, t, A, u); nothing), u0, (tnow, tend), model.parameters)
            integ = DE.init(prob, alg.ode_method; abstol = odeatol, reltol = odertol, callback = cb, saveat = sv,
                            internalnorm = _internalnorm, unstable_check = _unstable_check)
            DE.step!(integ, tend - tnow, true)
            for (t, u) in zip(integ.sol.t, integ.sol.u)
                uu = Array(u)
                push!(out.t, t); push!(out.p, FspVectorSparse(A.states, uu[1:n])); push!(out.sinks, uu[n+1:end])
            end
            tnow = integ.t
            ufin = integ.u; hit = tnow < tend
            sinks = ufin[n+1:n+R]
            pfin = copy(view(ufin, 1:n))
            dsinks = nothing
            if hit && adapter isa SelectiveRStepAdapter
                du = similar(ufin); matvec!(du, tnow, A, ufin); dsinks = du[n+1:n+R]
            end
        else
            s = ShardedVector(A); load!(s, p, sinks)
            nloc = s.hi - s.lo
            box = _plainbox(A, nloc + R)
            stats = try
                _segment!(:plain, A.h, box, s.v, tnow, tend, sv, odertol, odeatol, fsptol / tend)
            catch err
                # a joint propensity detected as c(t) g(x) broke the product form at a time the integrator used: discard
                # this segment (p, sinks still hold its initial state) and repeat it on the exact joint path
                (err isa SeparabilityError && detect_separable) || rethrow()
                detect_separable = false
                A = FspMatrixSparseB200(space, model.propensities; parameters = model.parameters, detect_separable = false, comm)
                continue
            end
            for (t, uu) in box.saved
                push!(out.t, t); push!(out.p, FspVectorSparse(A.states[s.lo+1:s.hi], uu[1:nloc])); push!(out.sinks, uu[nloc+1:end])
            end
            tnow = stats.t_final
            hit = stats.event_hit != 0 && tnow < tend
            sinks = s.v[nloc+1:nloc+R]
            dsinks = nothing
            if hit && adapter isa SelectiveRStepAdapter                  # get_du!(du, integrator) (rstepadapters.jl:100-103)
                du = DeviceVector(ctx, nloc + R); matvec!(du, tnow, A, s.v); dsinks = du[nloc+1:nloc+R]
            end
            pfin = gather(s, comm)
        end
        if hit
            p = adapt!(space, adapter, pfin, sinks, tnow, tend, fsptol; dsinks)
            A = FspMatrixSparseB200(space, A; detect_separable)           # incremental rebuild (fspsolve.jl:176)
            sum(sinks) >= tnow * fsptol / tend && (sinks .-= eps())      # fspsolve.jl:179-181
            verbose && println("t = $(round(tnow, digits=2)). Update state space. New size: $(get_state_count(space)).")
        else
            push!(out.t, tnow); push!(out.p, FspVectorSparse(A.states, Array(pfin))); push!(out.sinks, sinks)
            tnow = tend
        end
    end
    out
end

# ---- forward sensitivity: solve(model::CmeModelWithSensitivity, ic, tspan, AdaptiveForwardSensFspSparse; ...)
# forwardsenscmesparse.jl:99-215 with the block vector U = [p; s_1; ...; s_P] in HBM and K2 as right-hand side
function solve(model::CmeModelWithSensitivity, ic::ForwardSensFspInitialConditionSparse{NS,IntT,RealT},
    tspan::Tuple{AbstractFloat,AbstractFloat}, on::OnB200{AdaptiveForwardSensFspSparse}; saveat = [],
    fsptol::AbstractFloat = 1.0e-6, odeatol::AbstractFloat = 1.0e-10, odertol::AbstractFloat = 1.0e-4,
    verbose::Bool = false) where {NS,IntT<:Integer,RealT<:AbstractFloat}
    ctx, alg = on.ctx, on.alg
    on.comm === nothing || throw(ArgumentError("the forward-sensitivity solve runs on one GPU"))
    alg.ode_method === nothing || throw(ArgumentError("OnB200 forward-sensitivity solve: use ode_method = nothing (native device integrator)"))
    tstart, tend = min(tspan...), max(tspan...)
    P = get_parameter_count(model)
    length(get_sensitivity(ic)) ≠ P && throw(ArgumentError("Initial condition does not match CME model. Initial condition must contain `np` sensitivity vectors where `np` is the number of CME model parameters."))
    adapter = alg.space_adapter
    sv = _saveat(saveat, tspan)
    space = StateSpaceSparseB200(get_stoich_matrix(model), get_states(ic); ctx)
    R = Int(get_sink_count(space))
    vecs = DeviceVector[DeviceVector(ctx, _place(space, get_states(ic), v)) for v in [[get_probability(ic)]; get_sensitivity(ic)]]
    vecs = init!(space, adapter, vecs, tstart, fsptol)
    sinks = zeros(R); dsinks = [zeros(R) for _ in 1:P]
    out = NumCME.ForwardSensFspOutputSparse{NS,Int64,Float64}(t = Float64[], p = FspVectorSparse{NS,Int64,Float64}[],   # sensoutputsparse.jl:20-26
        sinks = AbstractVector{Float64}[], S = Vector{FspVectorSparse{NS,Int64,Float64}}[], dsinks = Vector{AbstractVector{Float64}}[])
    function pushslice!(t, states, uu, n)
        N = n + R
        push!(out.t, t); push!(out.p, FspVectorSparse(states, uu[1:n])); push!(out.sinks, uu[n+1:N])
        push!(out.S, [FspVectorSparse(states, uu[ip*N+1:ip*N+n]) for ip in 1:P])
        push!(out.dsinks, AbstractVector{Float64}[uu[ip*N+n+1:(ip+1)*N] for ip in 1:P])
    end
    tnow = tstart
    SA = nothing
    while tnow < tend
        SA = SA === nothing ? ForwardSensFspMatrixSparseB200(model, space) : ForwardSensFspMatrixSparseB200(model, space, SA)
        n = get_state_count(space); N = n + R
        U = fill!(DeviceVector(ctx, N * (P + 1)), 0.0)
        for (b, v) in enumerate(vecs); copyto!(view(U, (b-1)*N+1:(b-1)*N+n), v); end
        copyto!(view(U, n+1:N), sinks)
        for ip in 1:P; copyto!(view(U, ip*N+n+1:(ip+1)*N), dsinks[ip]); end
        box = CallbackBox(t -> (_prepare!(SA, t), SA.dcoef), R, length(SA.entries), N * (P + 1), Tuple{Float64,Vector{Float64}}[], nothing)
        stats = _segment!(:sens, SA.h, box, U, tnow, tend, sv, odertol, odeatol, fsptol / tend)
        states = SA.fspmatrix.states
        for (t, uu) in box.saved; pushslice!(t, states, uu, n); end
        tnow = stats.t_final
        if stats.event_hit != 0 && tnow < tend
            vecs = DeviceVector[copy(view(U, (b-1)*N+1:(b-1)*N+n)) for b in 1:P+1]
            sinks = U[n+1:N]; dsinks = [U[ip*N+n+1:(ip+1)*N] for ip in 1:P]
            vecs = adapt!(space, adapter, vecs, tnow, tend, fsptol)
            sum(sinks) >= tnow * fsptol / tend && (sinks .-= eps())      # forwardsenscmesparse.jl:187-189
            verbose && println("At t = $(round(tnow, digits=2)): update sate space. New size: $(get_state_count(space)) states.")
        else
            pushslice!(tnow, states, Array(U), n)
            tnow = tend
        end
    end
    out
end

export Context, Comm, DeviceVector, ShardedVector, StateSpaceSparseB200, FspMatrixSparseB200, ForwardSensFspMatrixSparseB200,
    OnB200, lincomb!, wrms, residuals!, axpy!, scale!, shift!, hasnonfinite, allowscalar, unique_id, shard_bounds, shard_info

end # module
