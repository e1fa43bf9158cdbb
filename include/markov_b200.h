/* SPDX-License-Identifier: MIT
 *
 * markov_b200.h — C ABI of libmarkov_b200.so: the B200-native (sm_100a) drop-in for
 * MarkovModels.jl's batched semiring inference path.
 *
 * The reference has no FFI: its "operator API" is Julia method dispatch from
 * src/inference.jl onto the GPU methods of src/linalg.jl (mul!, blockdiag, vcat,
 * broadcast!).  This boundary sits one level higher so that the per-frame host
 * loop (src/inference.jl:69-72,105-108) moves into the kernels.  Each entry point
 * names the reference interface it replaces.  A Julia shim binds these with
 * ccall; tests/bench bind them with ctypes.
 *
 * Conventions
 *   - every call returns an int status (MK_OK == 0); mk_last_error() gives the
 *     message of the calling thread's last failure.  No exceptions cross the ABI.
 *   - DimensionMismatch (src/linalg.jl:166-167,242-244) maps to MK_EINVAL.
 *   - handles are immutable after creation except for their private workspace:
 *     one in-flight call per mk_batch at a time; distinct batches may run
 *     concurrently on distinct streams.
 *   - "device" pointers are CUDA device pointers on the graph's device; the call
 *     is asynchronous on `stream` (a cudaStream_t passed as void*), like CUDA.jl
 *     launches on the task-local stream.  `_host` variants take host pointers,
 *     do the H2D/D2H copies themselves and return after synchronising.
 *   - payload floats only: Matrix{LogSemiring{Float32}} is bit-identical to
 *     Matrix{Float32}, so buffers are passed as float* / double*.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails with MK_ECUDA.
 */
#ifndef MARKOV_B200_H
#define MARKOV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MK_ABI_VERSION 2

enum mk_status {
    MK_OK = 0,
    MK_EINVAL = 22,   /* dimension mismatch / bad argument  (DimensionMismatch) */
    MK_ENOMEM = 12,   /* host or device allocation failed */
    MK_ENOTSUP = 95,  /* valid request the library does not implement */
    MK_ECUDA = 1000   /* CUDA runtime error; text in mk_last_error() */
};

/* Semirings.jl types the path is instantiated with (src/MarkovModels.jl:12). */
/* MK_PROB (ProbSemiring: ⊕ = +, ⊗ = *, 0̄ = 0, 1̄ = 1): native at the operator level; a graph created with it is stored
 * and run as its LogSemiring image (log.(weights)), emissions (probabilities) enter through log, α / β / totals leave
 * through exp — same posteriors, the range of the log semiring inside. */
enum mk_semiring { MK_LOG = 0, MK_TROPICAL = 1, MK_PROB = 2 };
enum mk_dtype { MK_F32 = 0, MK_F64 = 1 };

typedef struct mk_graph mk_graph; /* one compiled FSM resident on one GPU */
typedef struct mk_batch mk_batch; /* ragged batch of graphs (virtual rawunion) */

int mk_abi_version(void);
const char* mk_last_error(void);

/* Number of CUDA devices visible (0 when none / no driver). */
int mk_device_count(void);

/* FSM + compile + adapt(CuArray, ·)   — src/fsm.jl:7-28,42-48, src/inference.jl:3-26.
 *
 * The graph is the reference's extended matrix T̂ = [T ω; 0̄ 1̄] (phony final state
 * last, src/fsm.jl:22-27) exactly as the CPU FSM stores it, a
 * SparseMatrixCSC{K,Int64}: colptr (n_states_hat+1), rowval (nnz_hat), nzval
 * (nnz_hat) — column = destination state, rowval = source state.  α̂ is the
 * SparseVector nzind/nzval pair.  index_base is 1 for Julia arrays, 0 for C.
 * state2pdf[s] (same base) replaces Ĉ/Ĉᵀ: Ĉ has exactly one 1̄ per row
 * (examples/prepare-lfmmi-graphs.jl:15-23); the phony state must map to the phony
 * pdf n_pdf_hat (1-based) / n_pdf_hat-1 (0-based).  Inputs are copied; the caller
 * keeps ownership.  The library builds both orientations (T̂ and T̂ᵀ, the
 * copy(T̂') of src/inference.jl:12) on the device `device` (-1 = current). */
int mk_graph_create(mk_graph** out, int semiring, int dtype, int64_t n_states_hat,
                    int64_t nnz_hat, const int64_t* colptr, const int64_t* rowval,
                    const void* nzval, int64_t n_init, const int64_t* init_idx,
                    const void* init_w, const int32_t* state2pdf, int64_t n_pdf_hat,
                    int index_base, int device);
int mk_graph_destroy(mk_graph* g);
int mk_graph_info(const mk_graph* g, int64_t* n_states_hat, int64_t* nnz_hat,
                  int64_t* n_pdf_hat, int* semiring, int* dtype);

/* rawunion / batch   — src/fsmops.jl:28-36, src/inference.jl:28-36 (+ GPU blockdiag/vcat,
 * src/linalg.jl:73-157).  No block-diagonal matrix is materialised: the batch is a
 * descriptor; identical handles are stored once (the replicated denominator).  State
 * numbering of the virtual union is concatenation order: utterance b owns rows
 * [off_b, off_b + Ŝ_b).  All graphs must share semiring, dtype, device, n_pdf_hat. */
int mk_batch_create(mk_batch** out, mk_graph* const* graphs, int64_t B);
int mk_batch_destroy(mk_batch* b);
int mk_batch_info(const mk_batch* b, int64_t* B, int64_t* total_states_hat);

/* Emission argument shared by the calls below.
 * `ll` points at payload floats; element (b, d, n) (0-based utterance, pdf, frame)
 * lives at ll[b*stride_b + d*stride_d + n*stride_n] (strides in elements).
 *   expanded == 0: ll holds the un-padded D x T likelihoods; `expand`
 *       (src/inference.jl:54-60) is applied inside using seqlens (host int32[B] or
 *       NULL = all T): D̂ = D+1 must equal the graphs' n_pdf_hat, N̂ = T+1.
 *   expanded == 1: ll already holds the D̂ x N̂ matrices V̂ the reference's callers
 *       pass (examples/test_cuda.jl:124-128); then D == n_pdf_hat, T == N̂, seqlens
 *       must be NULL.
 */

/* αrecursion / βrecursion   — src/inference.jl:62-74, :99-110.
 * out: (ΣŜ_b) x N̂ column-major payload array on the device (state fastest). */
int mk_alpha(mk_batch* b, const void* ll, int64_t stride_b, int64_t stride_d, int64_t stride_n,
             int64_t D, int64_t T, int expanded, const int32_t* seqlens, void* out_A,
             void* stream);
int mk_beta(mk_batch* b, const void* ll, int64_t stride_b, int64_t stride_d, int64_t stride_n,
            int64_t D, int64_t T, int expanded, const int32_t* seqlens, void* out_B,
            void* stream);

/* pdfposteriors   — src/inference.jl:145-161 (and pdfposteriors2 :164-180).
 * out_post: the reference's (B, D, N) column-major array (b fastest), exp-domain,
 * real pdfs and real frames only (Ẑ[:, 1:end-1, 1:end-1], :160).  out_logz: B
 * totals (the reference's `ttl`, minimum over frames of the per-frame sums, :159).
 * Unreachable final state: posteriors 0, logz -Inf (the pdfposteriors3 convention,
 * src/inference.jl:198-200). */
int mk_pdfposteriors(mk_batch* b, const void* ll, int64_t stride_b, int64_t stride_d,
                     int64_t stride_n, int64_t D, int64_t T, int expanded,
                     const int32_t* seqlens, void* out_post, void* out_logz, void* stream);

/* pdfposteriors + the statistics a data-parallel training step exchanges (the accumulation the
 * reference's caller does on the host, examples/test_cuda.jl:140-152).  out_stats: DEVICE float64
 * [D + 2], overwritten with {Σ_b logZ_b, Σ_b frames of b, occupancy[d] = Σ_{b,n} out_post[b,d,n]};
 * the occupancy falls out of the normalisation pass over out_post, no extra pass.  NULL = not
 * wanted (then identical to mk_pdfposteriors). */
int mk_pdfposteriors_stats(mk_batch* b, const void* ll, int64_t stride_b, int64_t stride_d,
                           int64_t stride_n, int64_t D, int64_t T, int expanded,
                           const int32_t* seqlens, void* out_post, void* out_logz,
                           double* out_stats, void* stream);

/* bestpath   — absent from the 0.10.0 sources (src/MarkovModels.jl:56-57 are commented
 * exports); historical signature test/test_algorithms.jl:279-281.  Tropical graphs only,
 * expanded must be 0.  out_path: int32 [B][T] (utterance-major), 1-based state ids local
 * to each utterance's graph for frames < seqlens[b], 0 after; ties resolve to the
 * smallest predecessor index.  out_score: B path scores (-Inf and an all-zero path
 * when the final state is unreachable). */
int mk_bestpath(mk_batch* b, const void* ll, int64_t stride_b, int64_t stride_d,
                int64_t stride_n, int64_t D, int64_t T, int expanded, const int32_t* seqlens,
                int32_t* out_path, void* out_score, void* stream);

/* Host-buffer variants: what a caller holding CPU arrays (the reference's CPU
 * FSM path) calls.  ll/out_* are HOST pointers (pinned gives full PCIe rate);
 * copies in, computes, copies out, synchronises. */
int mk_pdfposteriors_host(mk_batch* b, const void* ll, int64_t stride_b, int64_t stride_d,
                          int64_t stride_n, int64_t D, int64_t T, int expanded,
                          const int32_t* seqlens, void* out_post, void* out_logz);
/* The same call split in two, for callers that keep several batches in flight (double buffering: one
 * mk_batch per buffer — they share the graph — so that the next batch's host-to-device copy and forward sweep
 * overlap this batch's device-to-host copy; PCIe is full duplex).  _begin enqueues the copies and kernels on
 * the batch's own streams and returns; the host buffers must stay valid and untouched (PINNED memory, or the
 * copies are not asynchronous) until mk_batch_wait(b) has returned. */
int mk_pdfposteriors_host_begin(mk_batch* b, const void* ll, int64_t stride_b, int64_t stride_d,
                                int64_t stride_n, int64_t D, int64_t T, int expanded,
                                const int32_t* seqlens, void* out_post, void* out_logz);
int mk_batch_wait(mk_batch* b);
/* Two batches in flight on one GPU (two streams, or two _begin calls): with overlap enabled on BOTH, the shared-graph
 * sweeps run with half the threads per CTA, so that the two batches' cooperative kernels are co-resident on every SM
 * and each fills the other's grid-barrier waits.  A lone batch is faster with overlap off (the default). */
int mk_batch_set_overlap(mk_batch* b, int enable);
int mk_bestpath_host(mk_batch* b, const void* ll, int64_t stride_b, int64_t stride_d,
                     int64_t stride_n, int64_t D, int64_t T, int expanded,
                     const int32_t* seqlens, int32_t* out_path, void* out_score);

/* LF-MMI gradient — the step after the two pdfposteriors calls of the reference's training
 * loop (examples/test_cuda.jl:118-152: permutedims :120, numerator / denominator posteriors
 * :140-143, their difference :152):
 *   grad[b, n, d] = scale * (den_post[b, d, n] - num_post[b, d, n])   for n < seqlens[b], else 0
 * num_post / den_post: (B, D, N) b-fastest device arrays as written by mk_pdfposteriors; grad:
 * the network's (B, T, D) device array, element strides given; seqlens_dev: DEVICE int32[B] or
 * NULL.  With scale = 1 this is d(-Σ_b (logZ_num - logZ_den)) / d(loglikes). */
int mk_lfmmi_grad(int dtype, const void* num_post, const void* den_post, int64_t B, int64_t D,
                  int64_t N, const int32_t* seqlens_dev, double scale, void* grad,
                  int64_t stride_b, int64_t stride_n, int64_t stride_d, void* stream);

/* ---- Operator level: the GPU methods of src/linalg.jl on caller-owned DEVICE arrays --------------------
 * A = CuSparseMatrixCSR{K}: rowptr (n_rows+1), colval (nnz), nzval (nnz), Cint indices, index_base 1 as CUDA.jl
 * stores them (0 accepted).  K ∈ {MK_LOG, MK_TROPICAL, MK_PROB} × {MK_F32, MK_F64} — the types the reference's
 * enabled tests instantiate (test/test_linalg.jl:34-54, 88-108).  Asynchronous on `stream`, current device.
 *
 * mk_spmv — LinearAlgebra.mul!(c::CuVector{K}, A::CuSparseMatrixCSR{K}, b::CuVector{K})
 *   (src/linalg.jl:163-184, kernel :213-233): c[i] = ⊕_k A[i,k] ⊗ b[k].  size(A,2) != len_b or
 *   size(A,1) != len_c -> MK_EINVAL (DimensionMismatch, :166-167); nnz == 0 -> no launch, c untouched (:169). */
int mk_spmv(int semiring, int dtype, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* rowptr,
            const int32_t* colval, const void* nzval, int index_base, const void* b, int64_t len_b,
            void* c, int64_t len_c, void* stream);
/* mk_spmm — LinearAlgebra.mul!(C::CuMatrix{K}, A::CuSparseMatrixCSR{K}, B::CuMatrix{K}, α, β)
 *   (src/linalg.jl:240-262, kernel :268-280): C = (β ⊗ C) ⊕ A ⊗ B with β ∈ {0, 1} as the reference's callers
 *   use it: accumulate = 0 overwrites C (β = 0: `fill!(C, 0̄)` then ⊕=), accumulate = 1 keeps it (β = 1).
 *   α is ignored, as in the reference.  B, C column-major with leading dimensions ldb, ldc (elements).
 *   Dimension errors (:242-244) -> MK_EINVAL. */
int mk_spmm(int semiring, int dtype, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* rowptr,
            const int32_t* colval, const void* nzval, int index_base, const void* B, int64_t rows_b,
            int64_t cols_b, int64_t ldb, void* C, int64_t rows_c, int64_t cols_c, int64_t ldc,
            int accumulate, void* stream);
/* mk_spvec_bcast — broadcast of a CuSparseVector{K} with a dense CuVector{K} (src/linalg.jl:287-338):
 *   dest .= 0̄;  dest[nzind[k]] = op == 0 ? nzval[k] ⊗ y[nzind[k]]   (elmul!, :292)
 *                                        : nzval[k] ⊘ y[nzind[k]]   (eldiv!, :294). */
int mk_spvec_bcast(int semiring, int dtype, int op, int64_t n, int64_t nnz, const int32_t* nzind,
                   const void* nzval, int index_base, const void* y, int64_t len_y, void* dest,
                   int64_t len_dest, void* stream);

/* ---- Graph preparation at the operator level (src/linalg.jl:12-157), DEVICE arrays, asynchronous on `stream` unless
 * noted.  The payload is opaque (dtype gives the element size): these operators move semiring values, never combine
 * them, so they serve Log / Tropical / Prob alike (test/test_linalg.jl:1-32, 56-86 instantiate all three).
 *
 * mk_blockdiag — SparseArrays.blockdiag(X::CuSparseMatrixCSC{K}...) (src/linalg.jl:73-100) and the CSR method
 *   (:102-131): block i is (ptr[i] [dim_ptr[i] + 1], idx[i], val[i]) with dim_ptr = columns for CSC / rows for CSR
 *   and dim_idx the other dimension; out_ptr has Σ dim_ptr + 1 entries, out_idx / out_val Σ nnz.  ptr / idx / val /
 *   dim_* / nnz are HOST arrays (of device pointers / sizes).  One launch for all blocks (the reference: three copies
 *   per block).
 * mk_vcat_spvec — Base.vcat(X::CuSparseVector{K}...) (src/linalg.jl:137-157): nzind shifted by the lengths so far.
 * mk_sparse_transpose — the CSC <-> CSR conversions and copy(transpose(M)) / copy(M') (src/linalg.jl:12-67), which
 *   the reference delegates to CUSPARSE csr2csc: the arrays (ptr [n_ptr + 1], idx, val) of a matrix compressed along
 *   its first-named dimension become the arrays compressed along the other one (out_ptr [n_idx + 1]), indices
 *   ascending inside every output segment.  CSR(A) -> CSC(A), CSC(A) -> CSR(A), and — reading the result with the
 *   roles swapped — CSR(A) -> CSR(Aᵀ). */
int mk_blockdiag(int dtype, int64_t n_mats, const int32_t* const* ptr, const int32_t* const* idx,
                 const void* const* val, const int64_t* dim_ptr, const int64_t* dim_idx, const int64_t* nnz,
                 int index_base, int32_t* out_ptr, int32_t* out_idx, void* out_val, void* stream);
int mk_vcat_spvec(int dtype, int64_t n_vecs, const int32_t* const* nzind, const void* const* nzval,
                  const int64_t* len, const int64_t* nnz, int32_t* out_ind, void* out_val, void* stream);
int mk_sparse_transpose(int dtype, int64_t n_ptr, int64_t n_idx, int64_t nnz, const int32_t* ptr,
                        const int32_t* idx, const void* val, int index_base, int32_t* out_ptr,
                        int32_t* out_idx, void* out_val, void* stream);

/* ---- Multi-GPU: the path shards by utterance (block-diagonal batch, src/fsmops.jl:28-36); the only
 * exchange is ONE sum all-reduce per step of the mk_pdfposteriors_stats vector.  NCCL is
 * bound at run time (dlopen of MK_NCCL_LIB / libnccl.so.2: the copy a CUDA.jl or PyTorch process already
 * holds); without it these return MK_ENOTSUP.
 *   one process per GPU:   rank 0 calls mk_comm_unique_id and hands the 128 bytes to the other ranks (file,
 *                          MPI, a socket ...); every rank then calls mk_comm_init_rank (device -1 = current).
 *   one process, n GPUs:   mk_comm_init fills out[0..n_gpus) with one communicator per device 0..n_gpus-1;
 *                          mk_allreduce_stats_all enqueues the n all-reduces as one NCCL group.
 * mk_allreduce_stats: in-place float64 sum over the ranks, asynchronous on `stream` of the communicator's
 * device. */
typedef struct mk_comm mk_comm;
int mk_comm_unique_id(void* id128);
int mk_comm_init_rank(mk_comm** out, int n_ranks, int rank, const void* id128, int device);
int mk_comm_init(mk_comm** out, int n_gpus);
int mk_allreduce_stats(mk_comm* c, double* stats, int64_t count, void* stream);
int mk_allreduce_stats_all(mk_comm* const* comms, double* const* stats, int n_gpus, int64_t count,
                           void* const* streams);
int mk_comm_destroy(mk_comm* c);

/* Instrumentation: number of kernels this library launched on the calling thread since
 * the last reset (bench.py's gpu_launches), and workspace bytes held by a batch. */
int64_t mk_launch_count(int reset);
int64_t mk_batch_workspace_bytes(const mk_batch* b);
/* When enabled, CUDA events are recorded (on the call's stream) around every launch of the
 * dominant kernel — the shared-graph forward-backward kernel — into a ring of 64 pairs;
 * mk_batch_kernel_ms waits for and returns the durations of the most recent launches (up to
 * `cap`, oldest first).  Enabling resets the ring. */
int mk_batch_profile(mk_batch* b, int enable);
/* The MUFU (exp2) rate of `device` (-1 = current), measured with a micro-kernel: the peak the SFU half of the
 * roofline (roofline time = max(t_HBM, t_SFU)) is quoted against.  Synchronous, ~10 ms. */
int mk_measure_sfu_peak(int device, double* ops_per_second);
int mk_batch_kernel_ms(mk_batch* b, float* ms, int cap, int* n);

#ifdef __cplusplus
}
#endif
#endif /* MARKOV_B200_H */
