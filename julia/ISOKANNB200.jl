# ISOKANNB200.jl -- the reference-side binding of libisokann_b200.so.
#
# Drop this file into ISOKANN.jl (`include("ISOKANNB200.jl")` after src/isotarget.jl) and set
# ENV["ISOKANN_B200_LIB"] to the built library.  It adds *methods* to the reference's own generic
# functions, dispatching on a `B200Model` wrapper, so `iso = Iso(data); run!(iso, n)` is unchanged:
#
#   hook                                   reference definition            replaced by
#   featurizer(coords)                     src/utils/features.jl:22-35     isokann_featurize
#   model(x)                               Flux.Chain call                 isokann_forward
#   isotarget(t, model, xs, ys)            src/isotarget.jl:34,100,152     isokann_target
#   train_batch!(model, xs, ys, opt, mb)   src/iso.jl:179-194              isokann_train_epoch
#   gpu(iso) / cpu(iso)                    src/iso.jl:256-257              upload/download params + opt state
#
# NOT EXECUTED in the build environment (no Julia there); the Python ctypes mirror
# (isokann.jl_b200/lib.py, engine.py) binds the identical symbols and is what the tests drive.

module ISOKANNB200

using ISOKANN, Flux, Optimisers, Random
import ISOKANN: isotarget, train_batch!, expectation, chis, chicoords, features, propfeatures,
                TransformShiftscale, TransformISA, TransformPseudoInv, SimulationData, Iso

const LIB = get(ENV, "ISOKANN_B200_LIB", "libisokann_b200.so")
const MAX_LAYERS = 8

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
    nparams::Int
    N::Int
    function B200Model(h, chain, P)
        m = new(h, chain, P, 0)
        finalizer(x -> ccall((:isokann_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), m)
    end
end

function check(m::B200Model, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:isokann_last_error, LIB), Cstring, (Ptr{Cvoid},), m.handle))
    1 <= rc <= 4 && throw(DomainError(rc, isempty(msg) ? DOMAIN_MESSAGES[Int(rc)] : msg))
    error("libisokann_b200 status $rc: $msg")
end

actid(f) = f === identity ? 0 : f in (Flux.sigmoid, Flux.sigmoid_fast) ? 1 : f in (tanh, Flux.tanh_fast) ? 2 : f === Flux.relu ? 3 :
           error("unsupported activation $f")

# flat parameter vector in Functors order: [LN.scale, LN.bias,] W1 (column-major), b1, ...
flatparams(chain) = reduce(vcat, vec.(Flux.trainables(chain)))

featspec(::typeof(identity)) = (0, 0, Int32[])
featspec(::ISOKANN.OpenMM.FeaturesCoords) = (0, 0, Int32[])
featspec(::ISOKANN.OpenMM.FeaturesAll) = (1, 0, Int32[])
featspec(f::ISOKANN.OpenMM.FeaturesAtoms) = (2, length(f.atominds), Int32.(f.atominds))
featspec(f::ISOKANN.OpenMM.FeaturesPairs) = (3, length(f.pairs), Int32.(collect(Iterators.flatten(f.pairs))))

"""B200Model(chain, rule, data): Iso(data; model, opt) (src/iso.jl:17-43) on the library"""
function B200Model(chain::Flux.Chain, rule, data::SimulationData; device=0, gemm=0)
    layers = collect(chain.layers)
    ln = layers[1] isa Flux.LayerNorm
    dense = ln ? layers[2:end] : layers
    widths = Int32[size(dense[1].weight, 2); [size(l.weight, 1) for l in dense]]
    wt = ntuple(i -> i <= length(widths) ? widths[i] : Int32(0), 9)
    wd, inner = rule.opts                                   # OptimiserChain(WeightDecay(λ), Adam|Nesterov)
    isadam = inner isa Optimisers.Adam
    kind, nidx, idx = featspec(data.featurizer)
    D = size(data.coords[1], 1)
    cfg = Ref(Config(length(dense), wt, ln, ln ? Float32(layers[1].ϵ) : 1f-5, actid(dense[1].σ), actid(dense[end].σ),
        isadam ? 1 : 0, Float32(inner.eta), Float32(wd.lambda), isadam ? Float32(inner.beta[1]) : 0.9f0,
        isadam ? Float32(inner.beta[2]) : 0.999f0, isadam ? Float32(inner.epsilon) : 1f-8,
        isadam ? 0.9f0 : Float32(inner.rho), kind, kind == 0 ? 0 : D ÷ 3, nidx, pointer(idx), device, gemm, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve idx begin
        rc = ccall((:isokann_create, LIB), Int32, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h)
    end
    rc == 0 || error("isokann_create failed ($rc): " * unsafe_string(ccall((:isokann_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    P = Int(ccall((:isokann_num_params, LIB), Int64, (Ptr{Cvoid},), h[]))
    m = B200Model(h[], chain, P)
    upload!(m)
    setdata!(m, data)
    return m
end

function upload!(m::B200Model)
    flat = Float32.(flatparams(m.chain))
    check(m, ccall((:isokann_upload_params, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), m.handle, flat, length(flat)))
end

"""cpu(iso): pull the parameters back into the Flux.Chain (src/iso.jl:257)"""
function download!(m::B200Model)
    flat = Vector{Float32}(undef, m.nparams)
    check(m, ccall((:isokann_download_params, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), m.handle, flat, m.nparams))
    o = 0
    for p in Flux.trainables(m.chain)
        copyto!(p, reshape(view(flat, o+1:o+length(p)), size(p))); o += length(p)
    end
    m.chain
end

"""SimulationData upload (src/simulation.jl:110-114): coordinates, not cached features"""
function setdata!(m::B200Model, data::SimulationData)
    xs, ys = Float32.(data.coords[1]), Float32.(data.coords[2])
    D, K, N = size(ys)
    check(m, ccall((:isokann_set_data, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int64, Int64, Int64),
        m.handle, xs, ys, D, K, N))
    m.N = N
end

# model(x): chis / chicoords / user code (src/iso.jl:203,211; src/isotarget.jl:18)
function (m::B200Model)(x::AbstractArray{<:Real}; is_features=true)
    x = Float32.(x)
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

transform_id(::TransformShiftscale) = (0, TargetOpts(1, 0, 1, 1, 1))
transform_id(t::TransformISA) = (1, TargetOpts(t.permute, t.whitening, 1, 1, 1))
transform_id(t::TransformPseudoInv) = (2, TargetOpts(t.permute, 0, t.normalize, t.direct, t.eigenvecs))

# isotarget(target, model, xs, ys) (src/isotarget.jl:12,34,100,152): xs/ys are already resident
function isotarget(t::Union{TransformShiftscale,TransformISA,TransformPseudoInv}, m::B200Model, xs, ys)
    id, opts = transform_id(t)
    out = Array{Float32}(undef, ISOKANN.outputdim(m.chain), m.N)
    check(m, ccall((:isokann_target, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{TargetOpts}, Ptr{Float32}),
        m.handle, id, Ref(opts), out))
    out
end

# train_batch!(model, xs, target, opt, minibatch) (src/iso.jl:179-194)
function train_batch!(m::B200Model, xs, target::AbstractMatrix, opt, minibatch; shuffle=true, partial=false)
    t = Float32.(target)
    check(m, ccall((:isokann_set_target, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64, Int64), m.handle, t, size(t, 1), size(t, 2)))
    N = size(t, 2)
    # the one randperm the DataLoader would draw this epoch (MLUtils shuffleobs), same RNG, same call
    perm = shuffle ? Int64.(randperm(Random.default_rng(), N)) : collect(Int64, 1:N)
    loss = Ref{Float64}(0)
    check(m, ccall((:isokann_train_epoch, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int32, Ref{Float64}),
        m.handle, perm, minibatch, partial, loss))
    loss[]
end

# dchidx(iso, x) (src/utils/minimumpath.jl:3-7): the Zygote pullback through chicoords is one library call
function ISOKANN.dchidx(iso::Iso{<:B200Model}, x::AbstractVecOrMat)
    m = iso.model
    xf = Float32.(x); M = length(xf) ÷ size(xf, 1)
    out = similar(xf)
    check(m, ccall((:isokann_chi_vjp, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int64, Int64, Int32, Ptr{Float32}, Ptr{Float32}),
        m.handle, xf, size(xf, 1), M, 0, C_NULL, out))
    out
end

# addcoords!(iso, coords) (src/iso.jl:238): propagate on the host as before, upload only the new block
function ISOKANN.addcoords!(iso::Iso{<:B200Model}, coords::AbstractMatrix)
    new = SimulationData(iso.data.sim, coords, ISOKANN.nk(iso.data), featurizer=iso.data.featurizer)
    xs, ys = Float32.(new.coords[1]), Float32.(new.coords[2])
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

"""Upload that overlaps the next Koopman pass: ys, then xs, stream in behind the call (keep both arrays alive and
unmodified until results computed from them have come back; page-lock them for a truly asynchronous copy)"""
function setdata_async!(m::B200Model, xs::Matrix{Float32}, ys::Array{Float32,3}; offset=0, nlocal=size(ys, 3))
    D, K, _ = size(ys)
    check(m, ccall((:isokann_set_data_async, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int64, Int64, Int64, Int64, Int64),
        m.handle, xs, ys, D, K, size(xs, 2), offset, nlocal))
    m.N = size(xs, 2)
end

# One Julia process per GPU: rank 0 creates the id, the host broadcasts the 128 bytes (MPI / Distributed), every
# rank joins before its first setdata!; ys then holds only this rank's contiguous slice of the start points.
unique_id() = (id = Vector{UInt8}(undef, 128); ccall((:isokann_comm_get_unique_id, LIB), Int32, (Ptr{UInt8},), id); id)
comm_init!(m::B200Model, world::Integer, rank::Integer, id::Vector{UInt8}) =
    check(m, ccall((:isokann_comm_init, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), m.handle, world, rank, id))

"""run!(iso, n, epochs) without host round trips (src/iso.jl:72-94)"""
function run_fused!(iso::Iso{<:B200Model}, n=1, epochs=1)
    m = iso.model
    id, opts = transform_id(iso.target)
    perms = reduce(hcat, [Int64.(randperm(Random.default_rng(), m.N)) for _ in 1:n*epochs])
    losses = Vector{Float64}(undef, n * epochs)
    check(m, ccall((:isokann_iterate, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ref{TargetOpts}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Float64}),
        m.handle, id, Ref(opts), n, epochs, iso.minibatch, perms, losses))
    append!(iso.losses, losses)
    iso
end

"""b200(iso): like gpu(iso) (src/iso.jl:256) -- moves model, optimiser state and data behind the library"""
b200(iso::Iso; kw...) = Iso(B200Model(iso.model, iso.opt isa Optimisers.AbstractRule ? iso.opt : iso.optrule, iso.data; kw...),
    iso.opt, iso.data, iso.target, iso.losses, iso.loggers, iso.minibatch)

end # module
