# NumCMEB200.jl -- Julia-side glue that keeps NumCME.jl's API for the FSP right-hand-side path while every
# kernel runs in libncme (hand-written CUDA for sm_100a, C ABI in include/ncme.h).
#
# NOT RUNNABLE IN THE BUILD ENVIRONMENT (no Julia there); written against include/ncme.h and reviewed by eye
# against the reference signatures quoted next to each method.  The Python package numcme.jl_b200/ is the executed
# mirror of exactly this file (same call sequence per method) and is what the parity tests drive.
#
# Usage inside NumCME.jl: `include("NumCMEB200.jl"); using .NumCMEB200` after the reference's own includes; the
# methods below add dispatch on the handle-backed types, the reference's CPU types keep working untouched.
module NumCMEB200

using NumCME
using StaticArrays: MVector
import NumCME: expand!, deleteat!, get_state_count, get_sink_count, get_states, get_statedict,
    get_state_connectivity, get_sink_connectivity, get_stoich_matrix, matvec!, matvecadd!, matvec,
    get_rowcount, get_colcount, get_parameters, get_propensities, init!, adapt!, solve, get_propensity_gradients,
    get_gradient_sparsity_patterns, get_parameter_count
import Base: size, *
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

# ------------------------------------------------------------------------------------------------ device vector
# The FSP vector stays in HBM.  AbstractVector surface needed by the reference's own code:
#   u[end-R+1:end], u[1:end-R]  (fspsolve.jl:146,172-173)  -> getindex(::UnitRange) downloads a slice
#   similar / copy / length / size / fill!  and the fused ops below (K7) for integrators that run on the host.
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
Base.size(v::DeviceVector) = (v.n,)
Base.length(v::DeviceVector) = v.n
Base.similar(v::DeviceVector) = DeviceVector(v.ctx, v.n)
Base.view(v::DeviceVector, r::UnitRange{<:Integer}) = DeviceVector(v.ctx, v.ptr + 8 * (first(r) - 1), length(r), false)
function Base.getindex(v::DeviceVector, r::UnitRange{<:Integer})        # slice download (host Vector)
    out = Vector{Float64}(undef, length(r))
    check(ccall((:ncme_d2h, libncme), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Csize_t), v.ctx.h, out,
        v.ptr + 8 * (first(r) - 1), 8 * length(r)))
    out
end
Base.getindex(v::DeviceVector, i::Integer) = v[i:i][1]
Base.Array(v::DeviceVector) = v[1:v.n]
Base.fill!(v::DeviceVector, a::Real) = (check(ccall((:ncme_vec_fill, libncme), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Cvoid}), v.ctx.h, v.n, a, v.ptr)); v)
Base.copyto!(y::DeviceVector, x::DeviceVector) = (check(ccall((:ncme_vec_copy, libncme), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}), y.ctx.h, y.n, x.ptr, y.ptr)); y)
Base.copy(x::DeviceVector) = copyto!(similar(x), x)
function Base.sum(v::DeviceVector)
    r = Ref{Float64}(0)
    check(ccall((:ncme_vec_sum, libncme), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ref{Float64}), v.ctx.h, v.n, v.ptr, r)); r[]
end
axpy!(a::Real, x::DeviceVector, y::DeviceVector) = (check(ccall((:ncme_vec_axpy, libncme), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Cvoid}, Ptr{Cvoid}), y.ctx.h, y.n, a, x.ptr, y.ptr)); y)
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
# Generic broadcast over DeviceVector is deliberately NOT defined: without CUDA.jl there is no broadcast code
# generation; an integrator must call lincomb!/axpy!/wrms (or use ode_method = nothing, below).

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
function get_states(sp::StateSpaceSparseB200{NS,NR}) where {NS,NR}        # sparsestatespace.jl:69
    n = get_state_count(sp)
    out = Vector{MVector{NS,Int64}}(undef, n)      # contiguous n*NS Int64: exactly the ABI's state-major layout
    n > 0 && check(ccall((:ncme_space_download_states, libncme), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}), sp.h, 0, n, out))
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
# every t actually used.
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
function FspMatrixSparseB200(space::StateSpaceSparseB200{NS,NR}, props::Vector{<:Propensity}; parameters = [],
                             detect_separable::Bool = true) where {NS,NR}
    states = get_states(space)                      # the host copy the reference keeps (`deepcopy(space.states)`, :97)
    n = length(states)
    kinds = Int32[!istimevarying(a) ? 0 : (istimeseparable(a) ? 1 : 2) for a in props]
    G = zeros(Float64, n, NR)                       # column r = state factor of reaction r: the ABI's reaction-major n x nr
    tfactors = Dict{Int,Any}()
    for (r, a) in enumerate(props)                  # host evaluation of the opaque closures, once per (state, reaction) (:129)
        kinds[r] == 0 && (for i in 1:n; G[i, r] = a.f(states[i], parameters); end)
        kinds[r] == 1 && (for i in 1:n; G[i, r] = a.statefactor(states[i], parameters); end; tfactors[r] = t -> a.tfactor(t, parameters))
        if kinds[r] == 2 && detect_separable
            found = detect_rank1(a.f, states, parameters)
            if found !== nothing
                g, sent = found
                G[:, r] .= g; kinds[r] = 1
                tfactors[r] = function (t)
                    c = a.f(t, states[sent[1]], parameters) / g[sent[1]]
                    for k in sent[2:end]
                        ck = a.f(t, states[k], parameters) / g[k]
                        abs(ck - c) > 1e-9 * max(abs(c), abs(ck), 1e-300) &&
                            error("propensity $r was classified as c(t) g(x) but is not separable at t = $t; pass detect_separable = false")
                    end
                    c
                end
            end
        end
    end
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ncme_matrix_create, libncme), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ref{Ptr{Cvoid}}), space.h, kinds, G, ref))
    A = FspMatrixSparseB200{NS,NR}(space.ctx, ref[], Vector{Any}(parameters), states, n + NR, n + NR, props, kinds, -Inf, ones(NR), tfactors)
    finalizer(x -> ccall((:ncme_matrix_destroy, libncme), Cint, (Ptr{Cvoid},), x.h), A)
end
# Rebuild after an adapt! (fspsolve.jl:176 rebuilds from scratch): only the states appended since `previous` was built
# (always the tail of the state list) are evaluated on the host, the factors of the surviving states are carried over
# on the device (SURVEY.md H8).  Falls back to the full constructor when `previous` is not the matrix the space was
# last assembled into or a joint propensity found to be c(t) g(x) stops being so on the new states.
function FspMatrixSparseB200(space::StateSpaceSparseB200{NS,NR}, previous::FspMatrixSparseB200{NS,NR}) where {NS,NR}
    props, parameters = previous.propensities, previous.parameters
    nk = Ref{Int64}(0); nn = Ref{Int64}(0)
    check(ccall((:ncme_space_new_count, libncme), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), space.h, nk, nn))
    states = get_states(space)
    n, nkept, nnew = length(states), Int(nk[]), Int(nn[])
    any(previous.kinds[r] == 1 && !istimeseparable(props[r]) for r in 1:NR) &&         # rank-1 joint reactions: re-detect
        return FspMatrixSparseB200(space, props; parameters)
    nkept == 0 && return FspMatrixSparseB200(space, props; parameters)
    G = zeros(Float64, nnew, NR)
    for (r, a) in enumerate(props), i in 1:nnew
        previous.kinds[r] == 0 && (G[i, r] = a.f(states[nkept+i], parameters))
        previous.kinds[r] == 1 && (G[i, r] = a.statefactor(states[nkept+i], parameters))
    end
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    code = ccall((:ncme_matrix_create_incremental, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                 space.h, C_NULL, previous.h, previous.kinds, G, ref)
    code == 0 || return FspMatrixSparseB200(space, props; parameters)
    A = FspMatrixSparseB200{NS,NR}(space.ctx, ref[], previous.parameters, states, n + NR, n + NR, props, copy(previous.kinds), -Inf, ones(NR), previous.tfactors)
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

# time-dependent pieces: separable factors are one host scalar per reaction per call (:204); joint reactions are
# re-evaluated on the host when t changes (:206-212) and uploaded
function _prepare!(A::FspMatrixSparseB200, t::Real)
    θ = A.parameters
    for (r, tf) in A.tfactors                       # separable reactions and joint ones found to be rank-1 (kinds[r] == 1)
        A.coef[r] = tf(t)
    end
    if t != A.t_cache
        A.t_cache = t
        for (r, a) in enumerate(A.propensities)
            if A.kinds[r] == 2
                vals = Float64[a.f(t, x, θ) for x in A.states]
                check(ccall((:ncme_matrix_set_joint_values, libncme), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), A.h, r, vals))
            end
        end
    end
    A.coef
end

# matvec!(out, t, A, v)   fspsparsematrix.jl:196   /   matvecadd!(out, t, A, v)   :226
function _apply!(out, t, A::FspMatrixSparseB200, v, beta::Float64)
    (length(out) == A.rowcount && length(v) == A.colcount) || throw(DimensionMismatch("matvec!: vector lengths must equal $(A.rowcount)"))
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
# addition (the reference imports mul! but defines no method, fspsparsematrix.jl:1): mul! at the cached time
mul!(y, A::FspMatrixSparseB200, x) = (matvec!(y, isfinite(A.t_cache) ? A.t_cache : 0.0, A, x); y)

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
    dvals = zeros(Float64, n, max(length(ents), 1))                       # entry-major nentries x n for the ABI
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
get_propensity_gradients(SA::ForwardSensFspMatrixSparseB200) = SA.propensity_gradients

# matvec!(out, t, SA, vs)                                         sensfspmatrixsparse.jl:97
function matvec!(out::DeviceVector, t::Real, SA::ForwardSensFspMatrixSparseB200, vs::DeviceVector)
    A = SA.fspmatrix
    θ = A.parameters
    P = length(θ)
    (length(out) == (P + 1) * A.rowcount && length(vs) == (P + 1) * A.rowcount) ||
        throw(DimensionMismatch("matvec!: expected vectors of length $((P + 1) * A.rowcount)"))
    coef = _prepare!(A, t)
    for (e, (r, ip)) in enumerate(SA.entries)
        if A.kinds[r] == 1                                               # d tfactor / d theta_ip   (:124-132)
            SA.dcoef[e] = SA.propensity_gradients[r].tfactor_pardiffs[ip](t, θ)
        elseif A.kinds[r] == 2                                           # joint: d f / d theta_ip over all states (:134-139, see Q6)
            vals = Float64[SA.propensity_gradients[r].pardiffs[ip](t, x, θ) for x in A.states]
            check(ccall((:ncme_sensmatrix_set_joint_values, libncme), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), SA.h, e - 1, vals))
        end
    end
    check(ccall((:ncme_sens_matvec, libncme), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}),
                SA.h, coef, SA.dcoef, vs.ptr, out.ptr))
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
# adapt!(space, adapter, p, sinks, t, tend, fsptol; integrator)      rstepadapters.jl:35 / :86
function adapt!(space::StateSpaceSparseB200, adapter::Union{RStepAdapter,SelectiveRStepAdapter}, p::DeviceVector,
    sinks::Vector{Float64}, t, tend, fsptol; dsinks::Union{Nothing,Vector{Float64}} = nothing)
    strict = adapter isa SelectiveRStepAdapter
    if adapter.dropstates
        dc = Ref{Int64}(0)                                               # sortperm + cumsum + threshold + deleteat! on the device
        check(ccall((:ncme_space_prune_by_mass, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Cint, Ref{Int64}),
            space.h, p.ptr, 1.0 - t * fsptol / tend, strict, dc))
        if dc[] > 0
            q = DeviceVector(p.ctx, get_state_count(space))
            check(ccall((:ncme_space_compact_vector, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), space.h, p.ptr, q.ptr))
            p = q
        end
    end
    strict ? expand!(space, adapter.max_step_count; onlyreactions = findall(dsinks .> 0)) : expand!(space, adapter.max_step_count)
    _grow(p, get_state_count(space))
end

# ------------------------------------------------------------------------------------------------ solve
# solve(model, p0, tspan, AdaptiveFspSparse(ode_method = nothing, space_adapter); saveat, fsptol, odeatol, odertol, verbose)
# fspsolve.jl:105-197 with the third-party integrator replaced by ncme_solve_segment (u never leaves HBM).
struct SolveOpts
    rtol::Float64; atol::Float64; event_slope::Float64; check_event::Cint; save_every_step::Cint
    nsave::Cint; save_t::Ptr{Float64}; h_init::Float64; max_steps::Int64; method::Cint
end
mutable struct SolveStats
    t_final::Float64; h_last::Float64; event_hit::Cint; nsaved::Cint
    steps::Int64; rejected::Int64; rhs_evals::Int64; launches::Int64
    SolveStats() = new(0, 0, 0, 0, 0, 0, 0, 0)
end
function _coef_cb(t::Float64, coef::Ptr{Float64}, user::Ptr{Cvoid})::Cvoid
    A = unsafe_pointer_to_objref(user)[1]::FspMatrixSparseB200
    c = _prepare!(A, t)
    for r in eachindex(c); unsafe_store!(coef, c[r], r); end
    nothing
end
function _save_cb(t::Float64, u::Ptr{Float64}, user::Ptr{Cvoid})::Cvoid
    A, sink = unsafe_pointer_to_objref(user)
    push!(sink, (t, copy(unsafe_wrap(Array, u, A.rowcount))))
    nothing
end
function solve(model::CmeModel, p0::FspVectorSparse{NS,IntT,RealT}, tspan::Tuple{AbstractFloat,AbstractFloat},
    alg::AdaptiveFspSparse; saveat = [], fsptol = 1.0e-6, odeatol = 1.0e-6, odertol = 1.0e-4, verbose = false,
    ctx::Context = default_ctx()) where {NS,IntT,RealT}
    alg.ode_method === nothing || return invoke(solve, Tuple{CmeModel,FspVectorSparse,Tuple,AdaptiveFspSparse}, model, p0, tspan, alg;
        saveat, fsptol, odeatol, odertol, verbose)              # a DifferentialEquations.jl algorithm: the reference's own loop
    tstart, tend = min(tspan...), max(tspan...)
    sv = saveat isa Number ? collect(Float64, tspan[1]:saveat:tspan[2]) : collect(Float64, saveat)
    adapter = alg.space_adapter
    space = StateSpaceSparseB200(model.stoich_matrix, p0.states; ctx)
    R = Int(get_sink_count(space))
    p = init!(space, adapter, DeviceVector(ctx, copy(p0.values)), tstart, fsptol)
    sinks = zeros(R); tnow = tstart
    A = FspMatrixSparseB200(space, model.propensities; parameters = model.parameters)
    out = FspOutputSparse{NS,Int64,Float64}(t = Float64[], p = FspVectorSparse{NS,Int64,Float64}[], sinks = Vector{Float64}[])
    ccoef = @cfunction(_coef_cb, Cvoid, (Float64, Ptr{Float64}, Ptr{Cvoid}))
    csave = @cfunction(_save_cb, Cvoid, (Float64, Ptr{Float64}, Ptr{Cvoid}))
    while tnow < tend
        n = get_state_count(space)
        u = DeviceVector(ctx, n + R); copyto!(view(u, 1:n), p)
        check(ccall((:ncme_h2d, libncme), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Csize_t), ctx.h, u.ptr + 8n, sinks, 8R))
        saved = Tuple{Float64,Vector{Float64}}[]; box = Ref((A, saved)); stats = SolveStats()
        opts = SolveOpts(odertol, odeatol, fsptol / tend, 1, isempty(sv) ? 1 : 0, length(sv), pointer(sv), 0.0, 0, 1)   # method 1: native BDF/GMRES (fused step kernel below 2e6 rows)
        GC.@preserve box sv check(ccall((:ncme_solve_segment, libncme), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Ptr{Cvoid}, Ref{SolveOpts}, Ref{SolveStats}),
            A.h, ccoef, csave, pointer_from_objref(box), tnow, tend, u.ptr, opts, stats))
        for (t, uu) in saved
            push!(out.t, t); push!(out.p, FspVectorSparse(A.states, uu[1:n])); push!(out.sinks, uu[n+1:end])
        end
        tnow = stats.t_final
        sinks = u[n+1:n+R]
        if stats.event_hit != 0 && tnow < tend
            dsinks = nothing
            if adapter isa SelectiveRStepAdapter                         # get_du!(du, integrator) (rstepadapters.jl:100-103)
                du = similar(u); matvec!(du, tnow, A, u); dsinks = du[n+1:n+R]
            end
            p = adapt!(space, adapter, copy(view(u, 1:n)), sinks, tnow, tend, fsptol; dsinks)
            A = FspMatrixSparseB200(space, A)                             # incremental rebuild (fspsolve.jl:176)
            sum(sinks) >= tnow * fsptol / tend && (sinks .-= eps())      # fspsolve.jl:179-181
            verbose && println("t = $(round(tnow, digits=2)). Update state space. New size: $(get_state_count(space)).")
        else
            push!(out.t, tnow); push!(out.p, FspVectorSparse(A.states, u[1:n])); push!(out.sinks, sinks)
        end
    end
    out
end

export Context, DeviceVector, StateSpaceSparseB200, FspMatrixSparseB200, ForwardSensFspMatrixSparseB200, lincomb!, wrms, axpy!

end # module
