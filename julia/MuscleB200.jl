# MuscleB200.jl — the reference-side binding of libmuscle_b200.so (include/muscle_b200.h).
#
# UNEXECUTED IN THIS REPOSITORY'S BUILD IMAGE: Julia is not installed there (no network), so this file is
# source-only. The Python/ctypes host in `muscle.jl_b200/` binds the very same entry points and is what
# the tests and benchmarks run; this is the stub a Muscle.jl maintainer would add (INTEGRATION.md).
#
# It plugs into Muscle's existing dispatch exactly like the cuTENSOR extension does:
#   * a new backend type next to src/Backend.jl:6-14,
#   * `Muscle.Domain(::Type{<:B200Array}) = DomainB200()`            (pattern: ext/MuscleCUDAExt.jl:7),
#   * `choose_backend_rule(binary_einsum, ::DomainB200, ::DomainB200)` (pattern: binary_einsum.jl:21),
#   * the 4-argument backend methods invoked at src/Operations/binary_einsum.jl:50 and :68.
module MuscleB200

using Muscle
using Muscle: Tensor, Index, inds, Backend, Domain
using Libdl

const libmuscle_b200 = Ref{String}(get(ENV, "MUSCLE_B200_LIB", "libmuscle_b200.so"))

# ---- status → Julia exceptions (include/muscle_b200.h: mb200_status_t) -------------------------------
const MB200_OK = Cint(0)
last_error() = unsafe_string(ccall((:mb200_last_error_string, libmuscle_b200[]), Cstring, ()))
function check(status::Cint)
    status == MB200_OK && return nothing
    msg = last_error()
    if status == 1 || status == 2          # INVALID_ARGUMENT / NOT_SUPPORTED
        throw(ArgumentError(msg))          # what @argcheck raises, binary_einsum.jl:82-83
    elseif status == 3                     # DIMENSION_MISMATCH
        throw(DimensionMismatch(msg))      # src/Tensor.jl:23
    else
        throw(ErrorException("libmuscle_b200 status $status: $msg"))
    end
end

# ---- handle (one per device per task-thread; the library locks internally) ---------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mb200_create, libmuscle_b200[]), Cint, (Ref{Ptr{Cvoid}}, Cint), r, device))
        h = new(r[])
        finalizer(h -> ccall((:mb200_destroy, libmuscle_b200[]), Cint, (Ptr{Cvoid},), h.ptr), h)
        return h
    end
end
const DEFAULT_HANDLE = Ref{Union{Nothing,Handle}}(nothing)
handle() = something(DEFAULT_HANDLE[], (DEFAULT_HANDLE[] = Handle(0)))

# ---- device array: dense, column-major, complex interleaved — a Julia Array that lives in HBM ---------
mutable struct B200Array{T,N} <: AbstractArray{T,N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
    function B200Array{T,N}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N}
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mb200_malloc, libmuscle_b200[]), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Csize_t),
                    handle().ptr, r, max(1, prod(dims)) * sizeof(T)))
        a = new{T,N}(r[], dims)
        finalizer(a -> ccall((:mb200_free, libmuscle_b200[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), handle().ptr, a.ptr), a)
        return a
    end
end
Base.size(a::B200Array) = a.dims
Base.similar(a::B200Array, ::Type{T}, dims::Dims{N}) where {T,N} = B200Array{T,N}(undef, dims)
Base.unsafe_convert(::Type{Ptr{T}}, a::B200Array{T}) where {T} = Ptr{T}(a.ptr)
Base.getindex(::B200Array, I...) = error("scalar indexing of a B200Array is not supported; use Array(a)")

function B200Array(x::Array{T,N}) where {T,N}
    a = B200Array{T,N}(undef, size(x))
    GC.@preserve x check(ccall((:mb200_memcpy_h2d, libmuscle_b200[]), Cint,
                               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), handle().ptr, a.ptr, pointer(x), sizeof(x)))
    check(ccall((:mb200_stream_sync, libmuscle_b200[]), Cint, (Ptr{Cvoid},), handle().ptr))
    return a
end
function Base.Array(a::B200Array{T,N}) where {T,N}
    x = Array{T,N}(undef, size(a))
    GC.@preserve x check(ccall((:mb200_memcpy_d2h, libmuscle_b200[]), Cint,
                               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), handle().ptr, pointer(x), a.ptr, sizeof(x)))
    check(ccall((:mb200_stream_sync, libmuscle_b200[]), Cint, (Ptr{Cvoid},), handle().ptr))
    return x
end

# ---- Backend / Domain plumbing -------------------------------------------------------------------------
struct BackendB200 <: Muscle.Backend end
struct DomainB200 <: Muscle.Domain end
Muscle.Domain(::Type{<:B200Array}) = DomainB200()
Muscle.choose_backend_rule(::typeof(Muscle.binary_einsum), ::DomainB200, ::DomainB200) = BackendB200()
Muscle.choose_backend_rule(::typeof(Muscle.binary_einsum!), ::DomainB200, ::DomainB200, ::DomainB200) = BackendB200()
# hybrid host / device operands (pattern of binary_einsum.jl:23-24 for Reactant): the host operand is uploaded by binary_einsum!
Muscle.choose_backend_rule(::typeof(Muscle.binary_einsum), ::DomainB200, ::Muscle.DomainHost) = BackendB200()
Muscle.choose_backend_rule(::typeof(Muscle.binary_einsum), ::Muscle.DomainHost, ::DomainB200) = BackendB200()

dtype_enum(::Type{Float32}) = Cint(0)
dtype_enum(::Type{Float64}) = Cint(1)
dtype_enum(::Type{ComplexF32}) = Cint(2)
dtype_enum(::Type{ComplexF64}) = Cint(3)
dtype_enum(::Type{T}) where {T} = throw(ArgumentError("eltype $T is not supported by BackendB200"))

# Index → int mode ids, the reference's own convention (ext/MuscleCUDAExt.jl:24-27)
function modes(inds_c, inds_a, inds_b)
    indmap = Dict{Index,Int32}()
    for ind in Iterators.flatten((inds_a, inds_b))
        get!(indmap, ind, Int32(length(indmap)))
    end
    all(i -> haskey(indmap, i), inds_c) || throw(ArgumentError("an index of the output is found in neither operand"))
    return Int32[indmap[i] for i in inds_c], Int32[indmap[i] for i in inds_a], Int32[indmap[i] for i in inds_b]
end

# binary_einsum(::BackendB200, inds_c, a, b) — the method called at src/Operations/binary_einsum.jl:50
function Muscle.binary_einsum(::BackendB200, inds_c, a::Tensor, b::Tensor)
    T = Base.promote_eltype(a, b)                                   # ext/MuscleCUDAExt.jl:16
    dims = map(inds_c) do i
        i ∈ inds(a) ? size(a, i) : i ∈ inds(b) ? size(b, i) : throw(ArgumentError("index $i not found in a nor b"))
    end
    # the result lives where the device operand lives (hybrid host / device operands: on the device)
    proto = parent(a) isa B200Array ? parent(a) : (parent(b) isa B200Array ? parent(b) : parent(a))
    c = Tensor(similar(proto, T, Tuple(dims)), collect(Index, inds_c))
    Muscle.binary_einsum!(BackendB200(), c, a, b)
    return c
end

# binary_einsum!(::BackendB200, c, a, b) — the method called at src/Operations/binary_einsum.jl:68
function Muscle.binary_einsum!(::BackendB200, c::Tensor, a::Tensor, b::Tensor)
    mc, ma, mb = modes(inds(c), inds(a), inds(b))
    ea, eb = Int64[size(a)...], Int64[size(b)...]
    pa, pb, pc = parent(a), parent(b), parent(c)
    # dispatch on ALL THREE parents: the device entry takes device pointers only, the host entry host pointers only
    # (this shim drives one device per process - handle() - so device arrays always share a GPU)
    ondev = (pa isa B200Array, pb isa B200Array, pc isa B200Array)
    if all(ondev)
        entry = :device
    elseif !any(ondev)
        entry = :host
    else
        # hybrid operands (cf. binary_einsum.jl:23-24): upload the host ones next to the device ones, then recurse
        pc isa B200Array || throw(ArgumentError("binary_einsum!: c on the host needs host operands"))
        a2 = pa isa B200Array ? a : Tensor(B200Array(pa), inds(a))
        b2 = pb isa B200Array ? b : Tensor(B200Array(pb), inds(b))
        return Muscle.binary_einsum!(BackendB200(), c, a2, b2)
    end
    GC.@preserve pa pb pc mc ma mb ea eb begin
        if entry === :device
            check(ccall((:mb200_binary_einsum, libmuscle_b200[]), Cint,
                (Ptr{Cvoid},
                 Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int64},
                 Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int64}, Ptr{Int64},
                 Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
                handle().ptr,
                pointer_of(pc), dtype_enum(eltype(c)), length(mc), mc, C_NULL,
                pointer_of(pa), dtype_enum(eltype(a)), length(ma), ma, ea, C_NULL,
                pointer_of(pb), dtype_enum(eltype(b)), length(mb), mb, eb, C_NULL))
            # stream-ordered; the shim synchronises before handing the result back to Julia code
            check(ccall((:mb200_stream_sync, libmuscle_b200[]), Cint, (Ptr{Cvoid},), handle().ptr))
        else
            # host Arrays (with_backend(BackendB200()) on DomainHost operands): the library stages through HBM
            check(ccall((:mb200_binary_einsum_host, libmuscle_b200[]), Cint,
                (Ptr{Cvoid},
                 Ptr{Cvoid}, Cint, Cint, Ptr{Int32},
                 Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int64},
                 Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int64}),
                handle().ptr,
                pointer_of(pc), dtype_enum(eltype(c)), length(mc), mc,
                pointer_of(pa), dtype_enum(eltype(a)), length(ma), ma, ea,
                pointer_of(pb), dtype_enum(eltype(b)), length(mb), mb, eb))
        end
    end
    return c
end

# Arithmetic of Float32 / ComplexF32 contractions on this thread's handle (mb200_compute_type_t): :default = tensor-core split
# scheme (TF32 + BF16 cross terms, rel. Frobenius error <= 1e-5), :fp32 = strict FP32 FMAs (what cuTENSOR's COMPUTE_32F and
# BackendBase give), :tf32x3 = the classic three-pass split.
function set_compute_type!(kind::Symbol)
    v = kind === :default ? 0 : kind === :fp32 ? 1 : kind === :tf32x3 ? 2 : throw(ArgumentError("unknown compute type $kind"))
    check(ccall((:mb200_set_compute_type, libmuscle_b200[]), Cint, (Ptr{Cvoid}, Cint), handle().ptr, v))
end

pointer_of(x::B200Array) = x.ptr
pointer_of(x::Array) = Ptr{Cvoid}(pointer(x))

sync() = check(ccall((:mb200_stream_sync, libmuscle_b200[]), Cint, (Ptr{Cvoid},), handle().ptr))

# ---- the einsum family next to the hot path (device arrays) -------------------------------------------------
Muscle.choose_backend_rule(::typeof(Muscle.unary_einsum), ::DomainB200) = BackendB200()              # unary_einsum.jl:20-21
Muscle.choose_backend_rule(::typeof(Muscle.unary_einsum!), ::DomainB200, ::DomainB200) = BackendB200()
Muscle.choose_backend_rule(::typeof(Muscle.hadamard), ::DomainB200, ::DomainB200) = BackendB200()    # hadamard.jl:4-5
Muscle.choose_backend_rule(::typeof(Muscle.hadamard!), ::DomainB200, ::DomainB200, ::DomainB200) = BackendB200()

function modes1(inds_x, inds_y)
    indmap = Dict{Index,Int32}()
    for ind in inds_x
        get!(indmap, ind, Int32(length(indmap)))
    end
    all(i -> haskey(indmap, i), inds_y) || throw(ArgumentError("Output indices must be a subset of input indices"))
    return Int32[indmap[i] for i in inds_x], Int32[indmap[i] for i in inds_y]
end

# unary_einsum(::BackendB200, inds_y, x) / unary_einsum!(::BackendB200, y, x) — replaces ext/MuscleOMEinsumExt.jl:25-38
function Muscle.unary_einsum(::BackendB200, inds_y, x::Tensor)
    y = Tensor(similar(parent(x), Tuple(size(x, i) for i in inds_y)), collect(Index, inds_y))
    return Muscle.unary_einsum!(BackendB200(), y, x)
end
function Muscle.unary_einsum!(::BackendB200, y::Tensor, x::Tensor)
    mx, my = modes1(inds(x), inds(y))
    ex = Int64[size(x)...]
    px, py = parent(x), parent(y)
    GC.@preserve px py mx my ex begin
        check(ccall((:mb200_unary_einsum, libmuscle_b200[]), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int64},
             Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
            handle().ptr, pointer_of(py), dtype_enum(eltype(y)), length(my), my, C_NULL,
            pointer_of(px), dtype_enum(eltype(x)), length(mx), mx, ex, C_NULL))
        sync()
    end
    return y
end

# hadamard(::BackendB200, a, b) / hadamard!(::BackendB200, c, a, b) — replaces src/Operations/hadamard.jl:42-77
function Muscle.hadamard(::BackendB200, a::Tensor, b::Tensor)
    ndims(a) >= ndims(b) || return Muscle.hadamard(BackendB200(), b, a)
    c = Tensor(similar(parent(a), Base.promote_eltype(a, b), size(a)), inds(a))
    return Muscle.hadamard!(BackendB200(), c, a, b)
end
function Muscle.hadamard!(::BackendB200, c::Tensor, a::Tensor, b::Tensor)
    ndims(a) >= ndims(b) || return Muscle.hadamard!(BackendB200(), c, b, a)
    inds(c) == inds(a) || throw(ArgumentError("inds(c) == inds(a) must hold"))
    ma, mb = modes1(inds(a), inds(b))
    ea, eb = Int64[size(a)...], Int64[size(b)...]
    pa, pb, pc = parent(a), parent(b), parent(c)
    GC.@preserve pa pb pc ma mb ea eb begin
        check(ccall((:mb200_hadamard, libmuscle_b200[]), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint,
             Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int64},
             Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int64}),
            handle().ptr, pointer_of(pc), dtype_enum(eltype(c)),
            pointer_of(pa), dtype_enum(eltype(a)), length(ma), ma, ea,
            pointer_of(pb), dtype_enum(eltype(b)), length(mb), mb, eb))
        sync()
    end
    return c
end

# ---- tensor_svd_thin(::BackendB200, A; inds_u, inds_v, ind_s) — replaces src/Operations/tensor_svd.jl:100-124 -----------
Muscle.choose_backend_rule(::typeof(Muscle.tensor_svd_thin), ::DomainB200) = BackendB200()
Muscle.choose_backend_rule(::typeof(Muscle.simple_update), ::DomainB200, ::DomainB200, ::DomainB200) = BackendB200()
function Muscle.tensor_svd_thin(::BackendB200, A::Tensor; inds_u=(), inds_v=(), ind_s=Index(gensym(:vind)), kwargs...)
    inds_u, inds_v = Muscle.factorinds(inds(A), inds_u, inds_v)
    ind_s ∉ inds(A) || throw(ArgumentError("ind_s must not be an index of A"))
    left, right = map(i -> size(A, i), inds_u), map(i -> size(A, i), inds_v)
    Amat = permutedims(A, [inds_u..., inds_v...])        # Base.permutedims(::B200Array, perm) = mb200_permute
    T, k = eltype(A), min(prod(left), prod(right))
    U = B200Array{T,length(left) + 1}(undef, (left..., k))
    S = B200Array{real(T),1}(undef, (k,))
    Vt = B200Array{T,length(right) + 1}(undef, (right..., k))
    pA = parent(Amat)
    GC.@preserve pA U S Vt begin
        check(ccall((:mb200_svd_thin, libmuscle_b200[]), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64, Cdouble, Cint),
            handle().ptr, U.ptr, S.ptr, Vt.ptr, pointer_of(pA), dtype_enum(T), prod(left), prod(right), 0.0, 0))
        sync()
    end
    return Tensor(U, [inds_u; ind_s]), Tensor(S, [ind_s]), Tensor(Vt, [inds_v; ind_s])
end
# tensor_qr_thin(::BackendB200, A; inds_q, inds_r, ind_virtual) — replaces src/Operations/tensor_qr.jl:57-79
Muscle.choose_backend_rule(::typeof(Muscle.tensor_qr_thin), ::DomainB200) = BackendB200()
function Muscle.tensor_qr_thin(::BackendB200, A::Tensor; inds_q=(), inds_r=(), ind_virtual=Index(gensym(:qr)), kwargs...)
    ind_virtual ∉ inds(A) || throw(ArgumentError("new virtual bond name ($ind_virtual) cannot be already be present"))
    inds_q, inds_r = Muscle.factorinds(inds(A), inds_q, inds_r)
    left, right = map(i -> size(A, i), inds_q), map(i -> size(A, i), inds_r)
    Amat = permutedims(A, [inds_q..., inds_r...])
    T, k = eltype(A), min(prod(left), prod(right))
    Q = B200Array{T,length(left) + 1}(undef, (left..., k))
    R = B200Array{T,length(right) + 1}(undef, (k, right...))
    pA = parent(Amat)
    GC.@preserve pA Q R begin
        check(ccall((:mb200_qr_thin, libmuscle_b200[]), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64),
            handle().ptr, Q.ptr, R.ptr, pointer_of(pA), dtype_enum(T), prod(left), prod(right)))
        sync()
    end
    return Tensor(Q, [inds_q..., ind_virtual]), Tensor(R, [ind_virtual, inds_r...])
end
# `simple_update(::Backend, ...)` (src/Operations/simple_update.jl:35-82) is written against binary_einsum,
# tensor_svd_thin and hadamard! only, so with the methods above it runs on B200Arrays unchanged.

function Base.permutedims(a::B200Array{T,N}, perm) where {T,N}
    out = B200Array{T,N}(undef, ntuple(d -> a.dims[perm[d]], N))
    ext, p0 = Int64[a.dims...], Int32[p - 1 for p in perm]
    GC.@preserve a out ext p0 begin
        check(ccall((:mb200_permute, libmuscle_b200[]), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Int32}, UInt32),
            handle().ptr, out.ptr, a.ptr, dtype_enum(T), N, ext, p0, 0))
        sync()
    end
    return out
end

# ---- Dagger bridge (SURVEY 8f row 4): `task_binary_einsum` (ext/MuscleDaggerExt/binary_einsum.jl:60-62) calls
# `binary_einsum(a, b; ...)` per chunk; chunks that are B200Arrays select BackendB200 through the Domain rule above,
# so Dagger drives one BackendB200 handle per worker process / GPU without further glue.

end # module
