# ISOKANNB200.jl -- the reference-side binding of libisokann_b200.so (include/isokann_b200.h, ABI version 3).
#
# Load it next to ISOKANN.jl (`include("ISOKANNB200.jl"); using .ISOKANNB200`) with ENV["ISOKANN_B200_LIB"]
# pointing at the built library.  It adds *methods* to the reference's own generic functions, dispatching on a
# `B200Model` wrapper, and a more specific method of `ISOKANN.gpu(::Iso)`, so that
#
#     iso = Iso(data; gpu=true); run!(iso, n)
#
# is unchanged (`Iso(...)` ends in `gpu(iso)`, src/iso.jl:41, which lands here):
#
#   hook                                   reference definition            replaced by
#   gpu(iso) / cpu(iso)                    src/iso.jl:256-257              B200Model + upload/download of params and
#                                                                          optimiser state
#   model(x)                               Flux.Chain call                 isokann_forward
#   isotarget(t, model, xs, ys)            src/isotarget.jl:34,100,152     isokann_target  (target stays resident)
#   train_batch!(model, xs, ys, opt, mb)   src/iso.jl:179-194              isokann_train_epoch
#   chis / chicoords / dchidx / addcoords! src/iso.jl:203,211,238          isokann_chis / _forward / _chi_vjp / _append_data
#   validationloss / rates                 src/iso.jl:160-168,339-351      isokann_validationloss / _rates
#   residual_ritz / residual_subspace      src/isotarget.jl:787-821        isokann_residual_ritz / _residual_subspace
#
# STATUS: written against the reference sources and Optimisers.jl 0.4 / Flux 0.16 state-tree layout, but NOT
# EXECUTED in the build environment (there is no Julia there).  The Python ctypes mirror (isokann.jl_b200/lib.py,
# engine.py) and the C program tests/abi_c/roundtrip.c bind the identical symbols and are what the tests drive.

module ISOKANNB200

using ISOKANN, Flux, Optimisers, Random
import ISOKANN: isotarget, train_batch!, chis, chicoords, Iso, SimulationData,
                TransformShiftscale, TransformISA, TransformPseudoInv

const LIB = get(ENV, "ISOKANN_B200_LIB", "libisokann_b200.so")
const ENABLED = get(ENV, "ISOKANN_B200", "1") != "0"

# mirror of isokann_config (include/isokann_b200.h)
struct Config
    n_layers::Int32
    widths::NTuple{9,Int32}
    layernorm::Int32
    ln_eps::Float32
    activation::Int32
    last_activation::Int32
    optimiser::Int32
    eta::Float32; lambda::Float32; beta1::Float32; beta2::Float32; eps::Float32; rho::Float32
    featurizer::Int32
    n_atoms::Int32
    n_index::Int32
    index::Ptr{Int32}
    device::Int32
    gemm_mode::Int32
    chunk::Int64
end

struct TargetOpts
    permute::Int32; whitening::Int32; normalize::Int32; direct::Int32; eigenvecs::Int32
end

const DOMAIN_MESSAGES = Dict(
    1 => "Could not compute the shift-scale. chi function is constant",
    2 => "The ISOKANN model collapsed under training. Try reducing the learning rate or increasing regularization",
    3 => "Could not compute the simplex transformation. The subspace might be singular/collapsed",
    4 => "Could not compute the pseudoinverse. The subspace might be singular/collapsed")

mutable struct B200Model
    handle::Ptr{Cvoid}
    chain::Flux.Chain          # host mirror: keeps model.layers / inputdim / outputdim / show working
    rule                       # the Optimisers rule (OptimiserChain(WeightDecay, Adam|Nesterov))
    nparams::Int
    N::Int
    target_stamp::Int          # bumped whenever the resident target changes
    function B200Model(h, chain, rule, P)
        m = new(h, chain, rule, P, 0, 0)
        finalizer(x -> ccall((:isokann_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), m)
    end
end

# what host code pokes at (src/iso.jl:261, src/models.jl:26-31)
Base.getproperty(m::B200Model, s::Symbol) = s === :layers ? getfield(m, :chain).layers : getfield(m, s)
ISOKANN.inputdim(m::B200Model) = ISOKANN.inputdim(m.chain)
ISOKANN.outputdim(m::B200Model) = ISOKANN.outputdim(m.chain)
ISOKANN.iscuda(::B200Model) = false     # callers hand over and receive host arrays (src/iso.jl:211, src/models.jl:35)

function check(m::B200Model, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:isokann_last_error, LIB), Cstring, (Ptr{Cvoid},), m.handle))
    1 <= rc <= 4 && throw(DomainError(rc, isempty(msg) ? DOMAIN_MESSAGES[Int(rc)] : msg))
    error("libisokann_b200 status $rc: $msg")
end

actid(f) = f === identity ? 0 : f in (Flux.sigmoid, Flux.sigmoid_fast) ? 1 : f in (tanh, Flux.tanh_fast) ? 2 : f === Flux.relu ? 3 :
           error("unsupported activation $f")

# the trainable arrays in the order of the flat parameter vector: [LN.scale, LN.bias,] W1, b1, W2, b2, ...
# (explicit walk over the Chain instead of Flux.trainables so that the optimiser tree below uses the same order)
function param_arrays(chain::Flux.Chain)
    out = AbstractArray[]
    for l in chain.layers
        if l isa Flux.LayerNorm
            push!(out, l.diag.scale, l.diag.bias)
        elseif l isa Flux.Dense
            push!(out, l.weight, l.bias)
        else
            error("unsupported layer $(typeof(l)): the library runs [LayerNorm,] Dense... chains")
        end
    end
    out
end

# the Optimisers.Leaf of every trainable array of a Flux.setup state tree, same order as param_arrays
function state_leaves(tree, chain::Flux.Chain)
    out = Optimisers.Leaf[]
    for (l, t) in zip(chain.layers, tree.layers)
        if l isa Flux.LayerNorm
            push!(out, t.diag.scale, t.diag.bias)
        else
            push!(out, t.weight, t.bias)
        end
    end
    out
end

flatparams(chain) = reduce(vcat, [vec(Float32.(p)) for p in param_arrays(chain)])

# Iso(data; opt) replaces the rule by Flux.setup(opt, model) (src/iso.jl:27): recover the rule from any leaf, as
# optimizerstring does (src/models.jl:23)
rule_of(opt::Optimisers.AbstractRule, chain) = opt
rule_of(tree, chain) = first(state_leaves(tree, chain)).rule

featspec(::typeof(identity)) = (0, 0, Int32[])
featspec(::ISOKANN.OpenMM.FeaturesCoords) = (0, 0, Int32[])
featspec(::ISOKANN.OpenMM.FeaturesAll) = (1, 0, Int32[])
featspec(f::ISOKANN.OpenMM.FeaturesAtoms) = (2, length(f.atominds), Int32.(f.atominds))
featspec(f::ISOKANN.OpenMM.FeaturesPairs) = (3, length(f.pairs), Int32.(collect(Iterators.flatten(f.pairs))))

"""B200Model(chain, rule, data): Iso(data; model, opt) (src/iso.jl:17-43) on the library"""
function B200Model(chain::Flux.Chain, rule::Optimisers.AbstractRule, data::SimulationData; device=0, gemm=0)
    layers = collect(chain.layers)
    ln = layers[1] isa Flux.LayerNorm
    dense = ln ? layers[2:end] : layers
    widths = Int32[size(dense[1].weight, 2); [size(l.weight, 1) for l in dense]]
    wt = ntuple(i -> i <= length(widths) ? widths[i] : Int32(0), 9)
    wd, inner = rule.opts                                   # OptimiserChain(WeightDecay(λ), Adam|Nesterov)
    isadam = inner isa Optimisers.Adam
    kind, nidx, idx = featspec(data.featurizer)
    D = size(data.coords[1], 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve idx begin
        cfg = Ref(Config(length(dense), wt, ln, ln ? Float32(layers[1].ϵ) : 1f-5, actid(dense[1].σ), actid(dense[end].σ),
            isadam ? 1 : 0, Float32(inner.eta), Float32(wd.lambda), isadam ? Float32(inner.beta[1]) : 0.9f0,
            isadam ? Float32(inner.beta[2]) : 0.999f0, isadam ? Float32(inner.epsilon) : 1f-8,
            isadam ? 0.9f0 : Float32(inner.rho), kind, kind == 0 ? 0 : D ÷ 3, nidx, pointer(idx), device, gemm, 0))
        rc = ccall((:isokann_create, LIB), Int32, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h)
    end
    rc == 0 || error("isokann_create failed ($rc): " * unsafe_string(ccall((:isokann_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    P = Int(ccall((:isokann_num_params, LIB), Int64, (Ptr{Cvoid},), h[]))
    m = B200Model(h[], chain, rule, P)
    upload!(m)
    setdata!(m, data)
    return m
end

function upload!(m::B200Model)
    flat = flatparams(m.chain)
    check(m, ccall((:isokann_upload_params, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), m.handle, flat, length(flat)))
end

"""pull the parameters back into the Flux.Chain (cpu(iso), src/iso.jl:257)"""
function download!(m::B200Model)
    flat = Vector{Float32}(undef, m.nparams)
    check(m, ccall((:isokann_download_params, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), m.handle, flat, m.nparams))
    o = 0
    for p in param_arrays(m.chain)
        copyto!(p, reshape(view(flat, o+1:o+length(p)), size(p))); o += length(p)
    end
    m.chain
end

isadam(m::B200Model) = m.rule.opts[2] isa Optimisers.Adam

"""optimiser state of a Flux.setup tree -> library.  Leaf state of OptimiserChain(WeightDecay, Adam) is
(nothing, (mt, vt, (β1^t, β2^t))); of OptimiserChain(WeightDecay, Nesterov) it is (nothing, velocity)."""
function upload_opt_state!(m::B200Model, tree)
    tree isa Optimisers.AbstractRule && return            # fresh rule: the library starts from zero state too
    leaves = state_leaves(tree, m.chain)
    if isadam(m)
        mt = reduce(vcat, [vec(Float32.(l.state[2][1])) for l in leaves])
        vt = reduce(vcat, [vec(Float32.(l.state[2][2])) for l in leaves])
        bt = Float32[leaves[1].state[2][3]...]
        check(m, ccall((:isokann_upload_opt_state, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Int64),
            m.handle, mt, vt, bt, m.nparams))
    else
        vel = reduce(vcat, [vec(Float32.(l.state[2])) for l in leaves])
        check(m, ccall((:isokann_upload_opt_state, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Int64),
            m.handle, vel, C_NULL, C_NULL, m.nparams))
    end
end

"""library -> a fresh Flux.setup(rule, chain) tree carrying the device-side optimiser state"""
function download_opt_state(m::B200Model)
    tree = Flux.setup(m.rule, m.chain)
    mt = Vector{Float32}(undef, m.nparams); vt = similar(mt); bt = zeros(Float32, 2)
    check(m, ccall((:isokann_download_opt_state, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Int64),
        m.handle, mt, vt, bt, m.nparams))
    o = 0
    for (p, leaf) in zip(param_arrays(m.chain), state_leaves(tree, m.chain))
        r = o+1:o+length(p); o += length(p)
        if isadam(m)
            leaf.state = (nothing, (reshape(mt[r], size(p)), reshape(vt[r], size(p)), (bt[1], bt[2])))
        else
            leaf.state = (nothing, reshape(mt[r], size(p)))
        end
    end
    tree
end

"""SimulationData upload (src/simulation.jl:110-114): coordinates, not cached features.  Float64 coordinates go
through the Float64 entry point (differences are formed on the device after the upload)."""
function setdata!(m::B200Model, data::SimulationData)
    xs, ys = data.coords
    D, K, N = size(ys)
    if eltype(xs) == Float64
        check(m, ccall((:isokann_set_data_f64, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64),
            m.handle, Array(xs), Array(ys), D, K, N))
    else
        check(m, ccall((:isokann_set_data, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int64, Int64, Int64),
            m.handle, Array{Float32}(xs), Array{Float32}(ys), D, K, N))
    end
    m.N = N
end

# model(x): chis / chicoords / user code (src/iso.jl:203,211; src/isotarget.jl:18)
function (m::B200Model)(x::AbstractArray{<:Real}; is_features=true)
    x = Array{Float32}(x)
    rows = size(x, 1); M = length(x) ÷ rows
    d = ISOKANN.outputdim(m.chain)
    out = Array{Float32}(undef, d, size(x)[2:end]...)
    check(m, ccall((:isokann_forward, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64, Int64, Int32, Ptr{Float32}),
        m.handle, x, rows, M, is_features, out))
    out
end

chis(iso::Iso{<:B200Model}) = (out = Array{Float32}(undef, ISOKANN.outputdim(iso.model.chain), iso.model.N);
    check(iso.model, ccall((:isokann_chis, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}), iso.model.handle, out)); out)
chicoords(iso::Iso{<:B200Model}, xs) = iso.model(xs; is_features=false)

# expectation(model, ys) on the resident Koopman samples (src/isotarget.jl:18)
function koopman_resident(m::B200Model)
    out = Array{Float32}(undef, ISOKANN.outputdim(m.chain), m.N)
    check(m, ccall((:isokann_koopman, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}), m.handle, out)); out
end
ISOKANN.koopman(iso::Iso{<:B200Model}) = koopman_resident(iso.model)

transform_id(::TransformShiftscale) = (0, TargetOpts(1, 0, 1, 1, 1))
transform_id(t::TransformISA) = (1, TargetOpts(t.permute, t.whitening, 1, 1, 1))
transform_id(t::TransformPseudoInv) = (2, TargetOpts(t.permute, 0, t.normalize, t.direct, t.eigenvecs))

"""The d x N target of the current iteration, resident on the device.  It behaves like a Matrix{Float32} (loggers
may index it: the first access downloads it once), but train_batch! recognises it and skips the upload, so the
target never round-trips through the host inside run!."""
mutable struct ResidentTarget <: AbstractMatrix{Float32}
    m::B200Model
    dims::Tuple{Int,Int}
    stamp::Int
    host::Union{Nothing,Matrix{Float32}}
end
Base.size(t::ResidentTarget) = t.dims
function hostcopy(t::ResidentTarget)
    if t.host === nothing
        t.stamp == t.m.target_stamp || error("this target is no longer resident (a newer one replaced it)")
        out = Matrix{Float32}(undef, t.dims...)
        check(t.m, ccall((:isokann_download_target, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}), t.m.handle, out))
        t.host = out
    end
    t.host
end
Base.getindex(t::ResidentTarget, i::Int, j::Int) = hostcopy(t)[i, j]
Base.Array(t::ResidentTarget) = copy(hostcopy(t))

# isotarget(target, model, xs, ys) (src/isotarget.jl:12,34,100,152): xs/ys are already resident
function isotarget(t::Union{TransformShiftscale,TransformISA,TransformPseudoInv}, m::B200Model, xs, ys)
    id, opts = transform_id(t)
    check(m, ccall((:isokann_target, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{TargetOpts}, Ptr{Float32}),
        m.handle, id, Ref(opts), C_NULL))
    m.target_stamp += 1
    ResidentTarget(m, (ISOKANN.outputdim(m.chain), m.N), m.target_stamp, nothing)
end

function run_epoch(m::B200Model, N, minibatch, shuffle, partial)
    # the one randperm the DataLoader would draw this epoch (MLUtils shuffleobs), same RNG, same call
    perm = shuffle ? Int64.(randperm(Random.default_rng(), N)) : collect(Int64, 1:N)
    loss = Ref{Float64}(0)
    check(m, ccall((:isokann_train_epoch, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int32, Ref{Float64}),
        m.handle, perm, minibatch, partial, loss))
    loss[]
end

# train_batch!(model, xs, target, opt, minibatch) (src/iso.jl:179-194).  `xs::AbstractMatrix` makes these methods
# strictly more specific than the reference's (model, xs::AbstractMatrix, ys::AbstractMatrix, opt, minibatch).
function train_batch!(m::B200Model, xs::AbstractMatrix, target::ResidentTarget, opt, minibatch; shuffle=true, partial=false)
    target.stamp == m.target_stamp || return train_batch!(m, xs, Array(target), opt, minibatch; shuffle, partial)
    run_epoch(m, size(target, 2), minibatch, shuffle, partial)
end
function train_batch!(m::B200Model, xs::AbstractMatrix, target::AbstractMatrix, opt, minibatch; shuffle=true, partial=false)
    t = Array{Float32}(target)      # user-defined targets (scripts/251126_carsten/main.jl:132) come from the host
    check(m, ccall((:isokann_set_target, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64, Int64), m.handle, t, size(t, 1), size(t, 2)))
    m.target_stamp += 1
    run_epoch(m, size(t, 2), minibatch, shuffle, partial)
end

# dchidx(iso, x) (src/utils/minimumpath.jl:3-7): the Zygote pullback through chicoords is one library call
function ISOKANN.dchidx(iso::Iso{<:B200Model}, x::AbstractVecOrMat)
    m = iso.model
    xf = Array{Float32}(x); M = length(xf) ÷ size(xf, 1)
    out = similar(xf)
    check(m, ccall((:isokann_chi_vjp, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int64, Int64, Int32, Ptr{Float32}, Ptr{Float32}),
        m.handle, xf, size(xf, 1), M, 0, C_NULL, out))
    out
end

# validationloss(iso, valdata) (src/iso.jl:160-168) in one call, nothing but the scalar comes back
function ISOKANN.validationloss(iso::Iso{<:B200Model}, valdata::SimulationData)
    vx, vy = Array{Float32}(valdata.coords[1]), Array{Float32}(valdata.coords[2])
    D, K, Nv = size(vy)
    out = Ref{Float64}(0)
    check(iso.model, ccall((:isokann_validationloss, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int64, Int64, Int64, Ref{Float64}),
        iso.model.handle, vx, vy, D, K, Nv, out))
    out[]
end

# rates(iso) (src/iso.jl:339-343): log(Kchi / chi) / lagtime; chi and Kchi stay on the device
function ISOKANN.rates(iso::Iso{<:B200Model})
    n = max(ISOKANN.outputdim(iso.model.chain), 2)
    q = Matrix{Float64}(undef, n, n)
    check(iso.model, ccall((:isokann_rates, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}), iso.model.handle, q, C_NULL))
    q ./ ISOKANN.lagtime(iso.data.sim)
end

# residual_subspace(iso) (src/isotarget.jl:805-821) on the resident data; `res=false` returns only relres (d numbers)
function ISOKANN.residual_subspace(iso::Iso{<:B200Model}; V_norms=false, res=true)
    m = iso.model
    d = ISOKANN.outputdim(m.chain)
    relres = Vector{Float64}(undef, d)
    R = res ? Matrix{Float64}(undef, m.N, d) : nothing
    check(m, ccall((:isokann_residual_subspace, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}),
        m.handle, V_norms, relres, res ? R : C_NULL))
    (; res=R, relres)
end

# residual_ritz(iso) (src/isotarget.jl:787-802); vecs are given in the basis Q whose R has a positive diagonal
# (Q = V / R with R = cholesky(V'V).U), which is what the returned named tuple's Q would be -- form it on the host if
# needed; `residues=false` returns only the O(d^2) quantities
function ISOKANN.residual_ritz(iso::Iso{<:B200Model}; residues=true)
    m = iso.model
    d = ISOKANN.outputdim(m.chain)
    vals = Vector{ComplexF64}(undef, d); vecs = Matrix{ComplexF64}(undef, d, d); relres = Vector{Float64}(undef, d)
    R = residues ? Matrix{ComplexF64}(undef, m.N, d) : nothing
    check(m, ccall((:isokann_residual_ritz, LIB), Int32,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{ComplexF64}),
        m.handle, vals, vecs, relres, residues ? R : C_NULL))
    if all(iszero ∘ imag, vals)      # a real spectrum gives real results in the reference
        return (; residues=residues ? real.(R) : nothing, relres, vals=real.(vals), vecs=real.(vecs))
    end
    (; residues=R, relres, vals, vecs)
end

# addcoords!(iso, coords) (src/iso.jl:238): propagate on the host as before, upload only the new block
function ISOKANN.addcoords!(iso::Iso{<:B200Model}, coords::AbstractMatrix)
    new = SimulationData(iso.data.sim, coords, ISOKANN.nk(iso.data), featurizer=iso.data.featurizer)
    xs, ys = Array{Float32}(new.coords[1]), Array{Float32}(new.coords[2])
    check(iso.model, ccall((:isokann_append_data, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int64, Int64, Int64),
        iso.model.handle, xs, ys, size(ys, 1), size(ys, 2), size(ys, 3)))
    iso.model.N += size(xs, 2)
    iso.data = ISOKANN.mergedata(iso.data, new)
    nothing
end

# iso.data = iso.data[end-cutoff+1:end] (run_kde!, src/iso.jl:288-290): drop the oldest points on the device
function cutoff!(iso::Iso{<:B200Model}, cutoff::Integer)
    length(iso.data) > cutoff || return nothing
    check(iso.model, ccall((:isokann_keep_last, LIB), Int32, (Ptr{Cvoid}, Int64), iso.model.handle, cutoff))
    iso.data = iso.data[end-cutoff+1:end]
    iso.model.N = cutoff
    nothing
end

# model(propfeatures(data)) for resample_kde / chistratcoords (src/simulation.jl:199-207,227-228): (d, K, N)
function propchis(iso::Iso{<:B200Model})
    m = iso.model
    out = Array{Float32}(undef, ISOKANN.outputdim(m.chain), ISOKANN.nk(iso.data), m.N)
    check(m, ccall((:isokann_chis_prop, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}), m.handle, out)); out
end

"""Upload that overlaps the next Koopman pass: ys, then xs, stream in behind the call.  The library page-locks both
arrays itself (cached per array); keep them alive and unmodified, and call release_host_buffers! before they may be
garbage collected."""
function setdata_async!(m::B200Model, xs::Matrix{Float32}, ys::Array{Float32,3}; offset=0, nlocal=size(ys, 3))
    D, K, _ = size(ys)
    check(m, ccall((:isokann_set_data_async, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int64, Int64, Int64, Int64, Int64),
        m.handle, xs, ys, D, K, size(xs, 2), offset, nlocal))
    m.N = size(xs, 2)
end
release_host_buffers!(m::B200Model) = check(m, ccall((:isokann_release_host_buffers, LIB), Int32, (Ptr{Cvoid},), m.handle))

# One Julia process per GPU: rank 0 creates the id, the host broadcasts the 128 bytes (MPI / Distributed), every
# rank joins before its first setdata!; ys then holds only this rank's contiguous slice of the start points.
unique_id() = (id = Vector{UInt8}(undef, 128); ccall((:isokann_comm_get_unique_id, LIB), Int32, (Ptr{UInt8},), id); id)
comm_init!(m::B200Model, world::Integer, rank::Integer, id::Vector{UInt8}) =
    check(m, ccall((:isokann_comm_init, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), m.handle, world, rank, id))

"""Julia's `randperm(Xoshiro(s0,s1,s2,s3), N)` replayed by the library (Xoshiro256++ + randperm!/ltm52 as written in
Random); returns the permutation and the advanced state"""
function lib_randperm(state::NTuple{4,UInt64}, N::Integer)
    st = UInt64[state...]
    out = Vector{Int64}(undef, N)
    ccall((:isokann_randperm, LIB), Int32, (Ptr{UInt64}, Int64, Ptr{Int64}), st, N, out)
    out, (st[1], st[2], st[3], st[4])
end

"""run!(iso, n, epochs) without host round trips (src/iso.jl:72-94)"""
function run_fused!(iso::Iso{<:B200Model}, n=1, epochs=1)
    m = iso.model
    id, opts = transform_id(iso.target)
    perms = reduce(hcat, [Int64.(randperm(Random.default_rng(), m.N)) for _ in 1:n*epochs])
    losses = Vector{Float64}(undef, n * epochs)
    check(m, ccall((:isokann_iterate, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ref{TargetOpts}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Float64}),
        m.handle, id, Ref(opts), n, epochs, iso.minibatch, perms, losses))
    m.target_stamp += 1
    append!(iso.losses, losses)
    iso
end

"""b200(iso): moves model, optimiser state and data behind the library, like gpu(iso) (src/iso.jl:256)"""
function b200(iso::Iso{<:Flux.Chain}; kw...)
    chain = Flux.cpu(iso.model)
    rule = rule_of(iso.opt, chain)
    m = B200Model(chain, rule, iso.data; kw...)
    upload_opt_state!(m, iso.opt)
    # iso.opt keeps the rule: run! only calls Optimisers.setup on it when it is an AbstractRule (src/iso.jl:74), and
    # train_batch!(::B200Model, ...) ignores it -- the state lives on the device
    Iso(m, iso.opt, iso.data, iso.target, iso.losses, iso.loggers, iso.minibatch)
end
b200(iso::Iso{<:B200Model}; kw...) = iso

# Iso(data; gpu=true) ends in ISOKANN.gpu(iso) (src/iso.jl:20,41): this method is more specific than gpu(::Iso), so
# the stock constructor lands on the library (set ENV["ISOKANN_B200"]="0" to keep Flux.gpu)
if ENABLED
    ISOKANN.gpu(iso::Iso{<:Flux.Chain}) = b200(iso)
    ISOKANN.gpu(iso::Iso{<:B200Model}) = iso
end

"""cpu(iso) (src/iso.jl:257): a plain Flux.Chain + Optimisers state tree again, so save (src/iso.jl:405-408, JLD2)
and every host-side analysis work on it"""
function ISOKANN.cpu(iso::Iso{<:B200Model})
    m = iso.model
    chain = deepcopy(download!(m))
    Iso(chain, download_opt_state(m), iso.data, iso.target, iso.losses, iso.loggers, iso.minibatch)
end

end # module
