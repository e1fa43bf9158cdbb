# SPDX-License-Identifier: MIT
#
# MarkovModelsB200.jl — the reference-side binding of libmarkov_b200.so: the methods a
# MarkovModels.jl maintainer adds so that `compile` / `batch` / `αrecursion` / `βrecursion` /
# `pdfposteriors` / `bestpath` run in the B200 library instead of the CUDA.jl kernels of
# src/linalg.jl.  NOT RUNNABLE IN THIS IMAGE (no julia binary, SURVEY.md G4): it documents the
# ccall layer; the same entry points are exercised through ctypes by tests/ and bench.py.
#
# Memory facts used (SURVEY.md A.4): Matrix{LogSemiring{Float32}} is bit-identical to
# Matrix{Float32}; the CPU FSM stores T̂ as SparseMatrixCSC{K,Int64} (colptr/rowval/nzval,
# 1-based) and α̂ as SparseVector{K,Int64} — passed as they are with index_base = 1.
module MarkovModelsB200

using CUDA, SparseArrays, Semirings
import MarkovModels: FSM, nstates

const LIB = "libmarkov_b200"

semiring_code(::Type{<:LogSemiring}) = Cint(0)
semiring_code(::Type{<:TropicalSemiring}) = Cint(1)
semiring_code(::Type{<:ProbSemiring}) = Cint(2)      # native at the operator level; graphs run as their LogSemiring image
dtype_code(::Type{Float32}) = Cint(0)
dtype_code(::Type{Float64}) = Cint(1)
payload(::Type{<:Semiring{T}}) where T = T   # val(x)::T

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:mk_last_error, LIB), Cstring, ()))
    rc == 22 ? throw(DimensionMismatch(msg)) : error("libmarkov_b200 error $rc: $msg")
end

# CompiledFSM{K} (src/inference.jl:3-12) -> an opaque mk_graph handle on the current device
mutable struct B200CompiledFSM{K}
    handle::Ptr{Cvoid}
    nstates_hat::Int
    npdf_hat::Int
end

# compile(fsm, Ĉ): Ĉ has one 1̄ per row (examples/prepare-lfmmi-graphs.jl:15-23)
function compile(fsm::FSM{K}, Ĉ::AbstractSparseMatrix{K}) where K
    T = payload(K)
    state2pdf = Cint[findnz(Ĉ[s, :])[1][1] for s in 1:size(Ĉ, 1)]
    T̂, α̂ = fsm.T̂, fsm.α̂
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mk_graph_create, LIB), Cint,
                (Ref{Ptr{Cvoid}}, Cint, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid},
                 Int64, Ptr{Int64}, Ptr{Cvoid}, Ptr{Cint}, Int64, Cint, Cint),
                h, semiring_code(K), dtype_code(T), size(T̂, 1), nnz(T̂),
                T̂.colptr, T̂.rowval, reinterpret(T, T̂.nzval),
                nnz(α̂), α̂.nzind, reinterpret(T, α̂.nzval), state2pdf, size(Ĉ, 2), 1, -1))
    c = B200CompiledFSM{K}(h[], size(T̂, 1), size(Ĉ, 2))
    finalizer(x -> ccall((:mk_graph_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), c)
end

# batch(fsm1, fsms...) (src/inference.jl:28-36): a descriptor, no blockdiag
mutable struct B200Batch{K}
    handle::Ptr{Cvoid}
    fsms::Vector{B200CompiledFSM{K}}   # keeps the graphs alive
end
function batch(fsm1::B200CompiledFSM{K}, fsms::B200CompiledFSM{K}...) where K
    all = [fsm1, fsms...]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mk_batch_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Int64),
                h, [f.handle for f in all], length(all)))
    b = B200Batch{K}(h[], all)
    finalizer(x -> ccall((:mk_batch_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), b)
end

# pdfposteriors(cfsm, V̂s) (src/inference.jl:164-180): V̂s are the expanded D̂ x N̂ CuMatrix{K};
# vcat'ed exactly like :146 so that element (b, d, n) sits at b*D̂ + d + n*B*D̂.
function pdfposteriors(b::B200Batch{K}, V̂s::Vector{<:CuMatrix{K}}) where K
    T = payload(K)
    V̂ = vcat(V̂s...)
    B, D̂, N̂ = length(V̂s), size(V̂s[1], 1), size(V̂s[1], 2)
    post = CUDA.zeros(T, B, D̂ - 1, N̂ - 1)
    ttl = CUDA.zeros(T, B)
    check(ccall((:mk_pdfposteriors, LIB), Cint,
                (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Cint, Ptr{Cint},
                 CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
                b.handle, pointer(V̂), D̂, 1, B * D̂, D̂, N̂, 1, C_NULL,
                pointer(post), pointer(ttl), CUDA.stream().handle))
    post, ttl
end

# The same call with the statistics a data-parallel step exchanges (the accumulation the caller of
# examples/test_cuda.jl:140-152 does on the host): stats = [Σ logZ, #frames, occupancy[1:D]] (Float64, device)
function pdfposteriors_stats(b::B200Batch{K}, V̂s::Vector{<:CuMatrix{K}}) where K
    T = payload(K)
    V̂ = vcat(V̂s...)
    B, D̂, N̂ = length(V̂s), size(V̂s[1], 1), size(V̂s[1], 2)
    post = CUDA.zeros(T, B, D̂ - 1, N̂ - 1)
    ttl = CUDA.zeros(T, B)
    stats = CUDA.zeros(Float64, D̂ + 1)
    check(ccall((:mk_pdfposteriors_stats, LIB), Cint,
                (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Cint, Ptr{Cint},
                 CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Float64}, Ptr{Cvoid}),
                b.handle, pointer(V̂), D̂, 1, B * D̂, D̂, N̂, 1, C_NULL,
                pointer(post), pointer(ttl), pointer(stats), CUDA.stream().handle))
    post, ttl, stats
end

# ---- multi-GPU: the batch shards by utterance; ONE sum all-reduce of `stats` per step -------------------
# One process per GPU (e.g. under MPI.jl): rank 0 creates the NCCL id, everybody gets it, everybody joins.
mutable struct B200Comm
    handle::Ptr{Cvoid}
end
function unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:mk_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    id
end
function B200Comm(nranks::Integer, rank::Integer, id::Vector{UInt8}; device::Integer = -1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mk_comm_init_rank, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{UInt8}, Cint), h, nranks, rank, id, device))
    finalizer(x -> ccall((:mk_comm_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), B200Comm(h[]))
end
# one process driving n GPUs: one communicator per device 0..n-1
function B200Comm(n_gpus::Integer)
    hs = Vector{Ptr{Cvoid}}(undef, n_gpus)
    check(ccall((:mk_comm_init, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), hs, n_gpus))
    [finalizer(x -> ccall((:mk_comm_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), B200Comm(h)) for h in hs]
end
function allreduce!(c::B200Comm, stats::CuVector{Float64})
    check(ccall((:mk_allreduce_stats, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Int64, Ptr{Cvoid}),
                c.handle, pointer(stats), length(stats), CUDA.stream().handle))
    stats
end

# ---- CPU arrays in, CPU arrays out (the reference's CPU FSM path), optionally overlapped --------------------
# pdfposteriors on host matrices (un-expanded D x N per utterance, stacked (B, D, N) b-fastest in `lhs`); with
# wait = false the call returns after enqueueing its copies and kernels: keep `lhs` / the outputs alive and pinned
# (CUDA.pin) until wait(b).  Two batches built from the same compiled graph overlap each other's copies.
function pdfposteriors_host!(post::Array{T,3}, ttl::Vector{T}, b::B200Batch{K}, lhs::Array{T,3},
                             seqlengths::Vector{<:Integer}; wait::Bool = true) where {K,T}
    B, D, N = size(lhs)
    fn = wait ? :mk_pdfposteriors_host : :mk_pdfposteriors_host_begin
    check(ccall((fn, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Cint, Ptr{Cint}, Ptr{Cvoid}, Ptr{Cvoid}),
                b.handle, lhs, 1, B, B * D, D, N, 0, Cint.(seqlengths), post, ttl))
    post, ttl
end
Base.wait(b::B200Batch) = check(ccall((:mk_batch_wait, LIB), Cint, (Ptr{Cvoid},), b.handle))

# αrecursion / βrecursion (src/inference.jl:62-74, 99-110): (ΣŜ) x N̂ CuMatrix{K}
for (fn, sym) in ((:αrecursion, :mk_alpha), (:βrecursion, :mk_beta))
    @eval function $fn(b::B200Batch{K}, V̂s::Vector{<:CuMatrix{K}}) where K
        T = payload(K)
        V̂ = vcat(V̂s...)
        B, D̂, N̂ = length(V̂s), size(V̂s[1], 1), size(V̂s[1], 2)
        total = sum(f.nstates_hat for f in b.fsms)
        out = CuArray{K}(undef, total, N̂)
        check(ccall(($(QuoteNode(sym)), LIB), Cint,
                    (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Cint, Ptr{Cint},
                     CuPtr{Cvoid}, Ptr{Cvoid}),
                    b.handle, pointer(V̂), D̂, 1, B * D̂, D̂, N̂, 1, C_NULL,
                    pointer(out), CUDA.stream().handle))
        out
    end
end

# bestpath(cfsm, lhs) (historical signature, examples/demo.ipynb cell 23): un-expanded D x N
# likelihoods for every utterance, (B, D, N) CuArray, plus sequence lengths
function bestpath(b::B200Batch{K}, lhs::CuArray{T,3}, seqlengths::Vector{<:Integer}) where {K<:TropicalSemiring,T}
    B, D, N = size(lhs)
    path = CUDA.zeros(Cint, N, B)   # column-major (N, B) == the ABI's [B][T]
    score = CUDA.zeros(T, B)
    check(ccall((:mk_bestpath, LIB), Cint,
                (Ptr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Cint, Ptr{Cint},
                 CuPtr{Cint}, CuPtr{Cvoid}, Ptr{Cvoid}),
                b.handle, pointer(lhs), 1, B, B * D, D, N, 0, Cint.(seqlengths),
                pointer(path), pointer(score), CUDA.stream().handle))
    permutedims(path), score
end

# ---- operator level: the methods of src/linalg.jl themselves -----------------------------------
# These REPLACE the bodies of src/linalg.jl:163-184 (mul! SpMV), :240-262 (mul! SpMM) and :299-320
# (_copyto! sparse-vector broadcast): same signatures, same @boundscheck behaviour (the library
# returns 22 -> DimensionMismatch), same "empty matrix launches nothing" rule.  CuSparseMatrixCSR
# stores Cint 1-based rowPtr / colVal — passed as they are with index_base = 1.
import LinearAlgebra
using CUDA.CUSPARSE: CuSparseMatrixCSR, CuSparseVector

function LinearAlgebra.mul!(c::CuVector{K}, A::CuSparseMatrixCSR{K}, b::CuVector{K}) where K<:Semiring
    check(ccall((:mk_spmv, LIB), Cint,
                (Cint, Cint, Int64, Int64, Int64, CuPtr{Cint}, CuPtr{Cint}, CuPtr{Cvoid}, Cint,
                 CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
                semiring_code(K), dtype_code(payload(K)), size(A, 1), size(A, 2), length(A.nzVal),
                pointer(A.rowPtr), pointer(A.colVal), pointer(A.nzVal), 1,
                pointer(b), length(b), pointer(c), length(c), CUDA.stream().handle))
    c
end

function LinearAlgebra.mul!(C::CuMatrix{K}, A::CuSparseMatrixCSR{K}, B::CuMatrix{K},
                            α::Number, β::Number) where K<:Semiring
    (β == 0 || β == 1) || (LinearAlgebra.rmul!(C, β); β = true)     # src/linalg.jl:246-248
    check(ccall((:mk_spmm, LIB), Cint,
                (Cint, Cint, Int64, Int64, Int64, CuPtr{Cint}, CuPtr{Cint}, CuPtr{Cvoid}, Cint,
                 CuPtr{Cvoid}, Int64, Int64, Int64, CuPtr{Cvoid}, Int64, Int64, Int64, Cint, Ptr{Cvoid}),
                semiring_code(K), dtype_code(payload(K)), size(A, 1), size(A, 2), length(A.nzVal),
                pointer(A.rowPtr), pointer(A.colVal), pointer(A.nzVal), 1,
                pointer(B), size(B, 1), size(B, 2), stride(B, 2),
                pointer(C), size(C, 1), size(C, 2), stride(C, 2), β == 1 ? 1 : 0, CUDA.stream().handle))
    C
end

# _copyto!(f, dest, x::CuSparseVector, y::CuVector) — src/linalg.jl:299-320; f is * or /
function _copyto!(f::Union{typeof(*), typeof(/)}, dest::CuArray{K}, x::CuSparseVector{K}, y::CuVector{K}) where K<:Semiring
    nzInd, nzVal = SparseArrays.nonzeroinds(x), SparseArrays.nonzeros(x)
    check(ccall((:mk_spvec_bcast, LIB), Cint,
                (Cint, Cint, Cint, Int64, Int64, CuPtr{Cint}, CuPtr{Cvoid}, Cint, CuPtr{Cvoid}, Int64,
                 CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
                semiring_code(K), dtype_code(payload(K)), f === (*) ? 0 : 1, length(x), length(nzVal),
                pointer(nzInd), pointer(nzVal), 1, pointer(y), length(y), pointer(dest), length(dest),
                CUDA.stream().handle))
    dest
end

# ---- graph preparation: REPLACE the bodies of src/linalg.jl:12-157 ---------------------------------------------
using CUDA.CUSPARSE: CuSparseMatrixCSC

function _transpose_arrays(::Type{K}, ptr, idx, val, n_ptr, n_idx) where K
    out_ptr, out_idx, out_val = CuVector{Cint}(undef, n_idx + 1), similar(idx), similar(val)
    check(ccall((:mk_sparse_transpose, LIB), Cint,
                (Cint, Int64, Int64, Int64, CuPtr{Cint}, CuPtr{Cint}, CuPtr{Cvoid}, Cint, CuPtr{Cint}, CuPtr{Cint},
                 CuPtr{Cvoid}, Ptr{Cvoid}),
                dtype_code(payload(K)), n_ptr, n_idx, length(val), pointer(ptr), pointer(idx), pointer(val), 1,
                pointer(out_ptr), pointer(out_idx), pointer(out_val), CUDA.stream().handle))
    out_ptr, out_idx, out_val
end
# CuSparseMatrixCSR(::CuSparseMatrixCSC) and back (src/linalg.jl:12-49)
function CUDA.CUSPARSE.CuSparseMatrixCSR(M::CuSparseMatrixCSC{K}) where K <: Semiring
    CuSparseMatrixCSR{K}(_transpose_arrays(K, M.colPtr, M.rowVal, M.nzVal, size(M, 2), size(M, 1))..., M.dims)
end
function CUDA.CUSPARSE.CuSparseMatrixCSC(M::CuSparseMatrixCSR{K}) where K <: Semiring
    CuSparseMatrixCSC{K}(_transpose_arrays(K, M.rowPtr, M.colVal, M.nzVal, size(M, 1), size(M, 2))..., M.dims)
end
# copy(M') / copy(transpose(M)) (src/linalg.jl:55-67)
function Base.copy(Mᵀ::Union{LinearAlgebra.Adjoint{K, <:CuSparseMatrixCSR}, LinearAlgebra.Transpose{K, <:CuSparseMatrixCSR}}) where K <: Semiring
    M = parent(Mᵀ)
    CuSparseMatrixCSR{K}(_transpose_arrays(K, M.rowPtr, M.colVal, M.nzVal, size(M, 1), size(M, 2))..., reverse(M.dims))
end
function Base.copy(Mᵀ::Union{LinearAlgebra.Adjoint{K, <:CuSparseMatrixCSC}, LinearAlgebra.Transpose{K, <:CuSparseMatrixCSC}}) where K <: Semiring
    M = parent(Mᵀ)
    CuSparseMatrixCSC{K}(_transpose_arrays(K, M.colPtr, M.rowVal, M.nzVal, size(M, 2), size(M, 1))..., reverse(M.dims))
end

# blockdiag (src/linalg.jl:73-131): one launch for all blocks instead of three copies per block
function _blockdiag(::Type{K}, ptrs, idxs, vals, dim_ptr, dim_idx) where K
    nnzs = Int64[length(v) for v in vals]
    out_ptr = CuVector{Cint}(undef, sum(dim_ptr) + 1)
    out_idx = CuVector{Cint}(undef, sum(nnzs))
    out_val = CuVector{K}(undef, sum(nnzs))
    GC.@preserve ptrs idxs vals begin
        check(ccall((:mk_blockdiag, LIB), Cint,
                    (Cint, Int64, Ptr{CuPtr{Cint}}, Ptr{CuPtr{Cint}}, Ptr{CuPtr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
                     Cint, CuPtr{Cint}, CuPtr{Cint}, CuPtr{Cvoid}, Ptr{Cvoid}),
                    dtype_code(payload(K)), length(ptrs), pointer.(ptrs), pointer.(idxs), pointer.(vals),
                    Int64.(dim_ptr), Int64.(dim_idx), nnzs, 1, pointer(out_ptr), pointer(out_idx), pointer(out_val),
                    CUDA.stream().handle))
    end
    out_ptr, out_idx, out_val
end
function SparseArrays.blockdiag(X::CuSparseMatrixCSR{K}...) where K <: Semiring
    m, n = sum(size(x, 1) for x in X), sum(size(x, 2) for x in X)
    CuSparseMatrixCSR{K}(_blockdiag(K, [x.rowPtr for x in X], [x.colVal for x in X], [x.nzVal for x in X],
                                    [size(x, 1) for x in X], [size(x, 2) for x in X])..., (m, n))
end
function SparseArrays.blockdiag(X::CuSparseMatrixCSC{K}...) where K <: Semiring
    m, n = sum(size(x, 1) for x in X), sum(size(x, 2) for x in X)
    CuSparseMatrixCSC{K}(_blockdiag(K, [x.colPtr for x in X], [x.rowVal for x in X], [x.nzVal for x in X],
                                    [size(x, 2) for x in X], [size(x, 1) for x in X])..., (m, n))
end

# vcat(::CuSparseVector...) (src/linalg.jl:137-157)
function Base.vcat(X::CuSparseVector{K}...) where K <: Semiring
    nnzs = Int64[length(SparseArrays.nonzeros(x)) for x in X]
    iPtr, nzVal = CuVector{Cint}(undef, sum(nnzs)), CuVector{K}(undef, sum(nnzs))
    inds, vals = [SparseArrays.nonzeroinds(x) for x in X], [SparseArrays.nonzeros(x) for x in X]
    GC.@preserve inds vals begin
        check(ccall((:mk_vcat_spvec, LIB), Cint,
                    (Cint, Int64, Ptr{CuPtr{Cint}}, Ptr{CuPtr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, CuPtr{Cint}, CuPtr{Cvoid}, Ptr{Cvoid}),
                    dtype_code(payload(K)), length(X), pointer.(inds), pointer.(vals), Int64[length(x) for x in X], nnzs,
                    pointer(iPtr), pointer(nzVal), CUDA.stream().handle))
    end
    CuSparseVector{K}(iPtr, nzVal, sum(length(x) for x in X))
end

end # module
