// SPDX-License-Identifier: MIT
//
// linalg.cuh — the reference's operator level (src/linalg.jl) as stand-alone sm_100a kernels:
//   spmv_kernel   c = A ⊗ b          mul!(c, A::CuSparseMatrixCSR{K}, b::CuVector{K})       src/linalg.jl:163-233
//   spmm_kernel   C = [C ⊕] A ⊗ B    mul!(C, A::CuSparseMatrixCSR{K}, B::CuMatrix{K}, α, β)  src/linalg.jl:240-280
//   spvec kernels dest = f.(x_sparse, y_dense), f ∈ {⊗, ⊘}                                  src/linalg.jl:287-338
// for K ∈ {LogSemiring, TropicalSemiring, ProbSemiring} × {Float32, Float64}: the operator coverage of the
// reference's enabled tests (test/test_linalg.jl:34-54, 88-108).  The inference entry points do NOT go through
// these (their recursions are fused, kernels.cuh); these serve callers of `mul!` itself and the
// emission-free algorithms of src/algorithms.jl (totalsum / totalcumsum).
//
// All three are HBM-bound streaming kernels: rowptr/colval/nzval are read once, coalesced; b is gathered
// (L2-resident for the graph sizes of the path).  ⊕ of the Log semiring is a single-pass running
// (max, scaled sum) pair per lane — one exp per arc, one log per row — combined across the lanes of the row's
// group with shuffles; the result has the semiring's full range (no exp of an un-shifted value anywhere).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace mk {

enum { LSR_LOG = 0, LSR_TROPICAL = 1, LSR_PROB = 2 };

template <typename T> __device__ __forceinline__ T lin_neg_inf();
template <> __device__ __forceinline__ float lin_neg_inf<float>() { return -INFINITY; }
template <> __device__ __forceinline__ double lin_neg_inf<double>() { return -(double)INFINITY; }
// Float32: ex2.approx / lg2.approx (2 ulp) — the arguments of exp are differences to the running maximum (<= 0), the
// argument of log is a sum in [1, nnz]: the result keeps ~1e-6 relative accuracy, far inside the 1e-4 bar.
__device__ __forceinline__ float lin_exp(float x) { return __expf(x); }
__device__ __forceinline__ double lin_exp(double x) { return exp(x); }
__device__ __forceinline__ float lin_log(float x) { return __logf(x); }
__device__ __forceinline__ double lin_log(double x) { return log(x); }

// Running ⊕ of one lane.  Log: value = m + log(s) with m the running maximum (s = 0 ⇔ nothing seen).
template <typename T, int SR> struct Acc;
template <typename T> struct Acc<T, LSR_LOG> {
    T m, s;
    __device__ __forceinline__ Acc() : m(lin_neg_inf<T>()), s(T(0)) {}
    __device__ __forceinline__ void add_prod(T w, T x) {  // ⊕= w ⊗ x
        const T v = w + x;
        if (!(v > lin_neg_inf<T>())) return;  // 0̄ (or NaN): nothing to add
        if (v <= m) s += lin_exp(v - m);
        else { s = s * lin_exp(m - v) + T(1); m = v; }  // first term: s = 0·exp(-Inf) + 1
    }
    __device__ __forceinline__ void merge(T om, T os) {
        if (!(os > T(0))) return;
        if (om <= m) s += os * lin_exp(om - m);
        else { s = s * lin_exp(m - om) + os; m = om; }
    }
    template <int LANES> __device__ __forceinline__ void reduce() {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
            const T om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
            merge(om, os);
        }
    }
    __device__ __forceinline__ T value() const { return s > T(0) ? m + lin_log(s) : lin_neg_inf<T>(); }
    __device__ __forceinline__ void add_value(T v) { add_prod(v, T(0)); }
    __device__ __forceinline__ void spill(T& a, T& b_) const { a = m; b_ = s; }
};
template <typename T> struct Acc<T, LSR_TROPICAL> {
    T m;
    __device__ __forceinline__ Acc() : m(lin_neg_inf<T>()) {}
    __device__ __forceinline__ void add_prod(T w, T x) { const T v = w + x; m = v > m ? v : m; }
    template <int LANES> __device__ __forceinline__ void reduce() {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
            const T om = __shfl_xor_sync(0xffffffffu, m, o);
            m = om > m ? om : m;
        }
    }
    __device__ __forceinline__ T value() const { return m; }
    __device__ __forceinline__ void add_value(T v) { m = v > m ? v : m; }
    __device__ __forceinline__ void spill(T& a, T& b_) const { a = m; b_ = T(0); }
    __device__ __forceinline__ void merge(T om, T) { m = om > m ? om : m; }
};
template <typename T> struct Acc<T, LSR_PROB> {
    T s;
    __device__ __forceinline__ Acc() : s(T(0)) {}
    __device__ __forceinline__ void add_prod(T w, T x) { s = fma(w, x, s); }
    template <int LANES> __device__ __forceinline__ void reduce() {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    __device__ __forceinline__ T value() const { return s; }
    __device__ __forceinline__ void add_value(T v) { s += v; }
    __device__ __forceinline__ void spill(T& a, T& b_) const { a = s; b_ = T(0); }
    __device__ __forceinline__ void merge(T os, T) { s += os; }
};

// c[r] = ⊕_k nzval[k] ⊗ b[colval[k]] — LANES lanes per row (the reference spends a whole warp per row,
// src/linalg.jl:213-233; rows of the path's graphs hold ~17 arcs, so the host picks 2/4/32 from nnz / rows).
// A row is taken in chunks of LANES x R arcs: every lane first requests its R (colval, nzval) pairs and the R
// gathers from b — all in flight together — and only then folds them.  Log semiring: the chunk's maximum is made
// group-uniform with LANES-wide shuffles, every lane adds exp(v - M) of its own arcs to a lane-local sum (one exp per
// arc, no data-dependent branches), a later chunk rescales that sum by exp(M_old - M_new); one shuffle-sum at the end.
template <typename T, int SR, int LANES, int R>
__global__ void spmv_kernel(long long n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
                            const T* __restrict__ nzval, int base, const T* __restrict__ b, T* __restrict__ c,
                            int long_row, int* __restrict__ worklist, int worklist_cap) {
    constexpr int U = 1;  // rows per lane group in flight (2 measured slower: 1.02 vs 0.65 ms at cfg 3)
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = int(gid % LANES);
    const long long rows_per_pass = (long long)gridDim.x * blockDim.x / LANES;
    // every lane of a warp runs the same number of passes: the shuffles below are warp-wide
    const long long passes = (n_rows + rows_per_pass * U - 1) / (rows_per_pass * U);
    long long r0 = gid / LANES;
    const T zero = SR == LSR_PROB ? T(0) : lin_neg_inf<T>();
    // (the row pointers of the next pass are requested before this pass's arcs: one round trip less per pass)
    int nbeg[U], nend[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * rows_per_pass;
        nbeg[u] = nend[u] = 0;
        if (r < n_rows) { nbeg[u] = rowptr[r] - base; nend[u] = rowptr[r + 1] - base; }
    }
    for (long long it = 0; it < passes; ++it, r0 += rows_per_pass * U) {
        int beg[U], end[U], len = 0;
        bool skip[U] = {};
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * rows_per_pass;
            beg[u] = nbeg[u]; end[u] = nend[u];
            const long long rn = r + rows_per_pass * U;
            nbeg[u] = nend[u] = 0;
            if (rn < n_rows) { nbeg[u] = rowptr[rn] - base; nend[u] = rowptr[rn + 1] - base; }
            // A row far longer than the rest (the phony final state of an FSM collects every final weight: 9 300 arcs
            // against a mean of 17 at cfg 3) would keep one lane group busy for hundreds of chunks while the grid
            // drains: it goes to the work list of spmv_long_kernel (a whole CTA per row) instead.
            const bool is_long = worklist != nullptr && end[u] - beg[u] > long_row;
            int slot = 0;
            if (is_long && lane == 0) slot = atomicAdd(worklist, 1);
            slot = __shfl_sync(0xffffffffu, slot, 0, LANES);  // (every lane of the warp takes part; is_long is group-uniform)
            if (is_long && slot < worklist_cap) {
                if (lane == 0) worklist[1 + slot] = int(r);
                beg[u] = end[u] = 0;
                skip[u] = true;
            }
            len = max(len, end[u] - beg[u]);
        }
        // chunks of this warp: the longest row among the warp's groups decides (shuffles are warp-wide)
#pragma unroll
        for (int o = 16; o >= LANES; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
        T M[U], S[U], A[U];  // Log: running maximum (group-uniform) and lane-local scaled sum; else lane-local accumulator
#pragma unroll
        for (int u = 0; u < U; ++u) { M[u] = lin_neg_inf<T>(); S[u] = T(0); A[u] = zero; }
        for (int k0 = 0; k0 < len; k0 += LANES * R) {
            T v[U][R];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int k = beg[u] + k0 + q * LANES + lane;
                    v[u][q] = zero;
                    if (k < end[u]) {
                        const T w = nzval[k];
                        const T x = b[colval[k] - base];
                        v[u][q] = SR == LSR_PROB ? w * x : w + x;
                    }
                }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (SR == LSR_LOG) {
                    T m = v[u][0];
#pragma unroll
                    for (int q = 1; q < R; ++q) m = v[u][q] > m ? v[u][q] : m;
#pragma unroll
                    for (int o = LANES / 2; o > 0; o >>= 1) { const T om = __shfl_xor_sync(0xffffffffu, m, o); m = om > m ? om : m; }
                    const T Mn = m > M[u] ? m : M[u];
                    if (Mn > lin_neg_inf<T>()) {  // (everything seen so far is 0̄ otherwise)
                        T sum = S[u] * lin_exp(M[u] - Mn);   // first chunk: S = 0
#pragma unroll
                        for (int q = 0; q < R; ++q) sum += lin_exp(v[u][q] - Mn);  // exp(-Inf) = 0 for pads and 0̄ entries
                        S[u] = sum;
                        M[u] = Mn;
                    }
                } else if (SR == LSR_TROPICAL) {
#pragma unroll
                    for (int q = 0; q < R; ++q) A[u] = v[u][q] > A[u] ? v[u][q] : A[u];
                } else {
#pragma unroll
                    for (int q = 0; q < R; ++q) A[u] += v[u][q];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            T out;
            if (SR == LSR_LOG) {
                T sum = S[u];
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                out = sum > T(0) ? M[u] + lin_log(sum) : lin_neg_inf<T>();
            } else {
                T acc = A[u];
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) {
                    const T oa = __shfl_xor_sync(0xffffffffu, acc, o);
                    acc = SR == LSR_TROPICAL ? (oa > acc ? oa : acc) : acc + oa;
                }
                out = acc;
            }
            const long long r = r0 + u * rows_per_pass;
            if (lane == 0 && r < n_rows && !skip[u]) c[r] = out;
        }
    }
}

// The long rows of spmv_kernel's work list ({count, rows...}): one CTA of 256 threads per row, grid-stride over the list.
template <typename T, int SR>
__global__ void spmv_long_kernel(const int* __restrict__ worklist, int worklist_cap, const int32_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ colval, const T* __restrict__ nzval, int base,
                                 const T* __restrict__ b, T* __restrict__ c) {
    __shared__ T sm[32], ss[32];
    const int count = min(worklist[0], worklist_cap);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = blockIdx.x; i < count; i += gridDim.x) {
        const int r = worklist[1 + i];
        const int beg = rowptr[r] - base, end = rowptr[r + 1] - base;
        Acc<T, SR> acc;
        // (8 arcs and their gathers in flight per thread: one arc at a time measured 44 us for the 128 final-state rows of
        // cfg 3 — 9 300 arcs each, 36 dependent round trips per thread)
        constexpr int Q = 8;
        for (int k0 = beg + threadIdx.x; k0 < end; k0 += blockDim.x * Q) {
            int col[Q];
            T w[Q], x[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int k = k0 + q * blockDim.x;
                if (k < end) { col[q] = colval[k] - base; w[q] = nzval[k]; }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q)
                if (k0 + q * blockDim.x < end) x[q] = b[col[q]];
#pragma unroll
            for (int q = 0; q < Q; ++q)
                if (k0 + q * blockDim.x < end) acc.add_prod(w[q], x[q]);
        }
        acc.template reduce<32>();
        if (lane == 0) acc.spill(sm[warp], ss[warp]);
        __syncthreads();
        if (threadIdx.x == 0) {
            Acc<T, SR> tot;
            for (int w = 0; w < int(blockDim.x >> 5); ++w) tot.merge(sm[w], ss[w]);
            c[r] = tot.value();
        }
        __syncthreads();
    }
}

// C[i, j] = [C[i, j] ⊕] ⊕_k nzval[k] ⊗ B[colval[k], j], column-major C (ldc) and B (ldb).
// One thread per row i and chunk of CJ columns, i fastest: the writes to C, the row pointers and (for the short rows
// of Ĉ / T̂) the arcs of neighbouring rows are contiguous across a warp, every arc is read once per CJ columns and its
// CJ gathers from B are independent loads in flight together — the reference's kernel (:268-280) strides a thread over
// rows with the column loop inside and a global read-modify-write per arc.
// grid = (ceil(m / 256), min(ceil(n_cols_b / CJ), 65535)).
template <typename T, int SR, int CJ>
__global__ void spmm_kernel(long long n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
                            const T* __restrict__ nzval, int base, const T* __restrict__ B, long long ldb,
                            T* __restrict__ C, long long ldc, long long n_cols_b, int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int beg = rowptr[i] - base, end = rowptr[i + 1] - base;
    for (long long j0 = (long long)blockIdx.y * CJ; j0 < n_cols_b; j0 += (long long)gridDim.y * CJ) {
        const T* Bj = B + j0 * ldb;
        if (end - beg == 1 && !accumulate && j0 + CJ <= n_cols_b) {
            // one arc per row — the state-to-pdf maps Ĉ of the path (exactly one 1̄ per row,
            // examples/prepare-lfmmi-graphs.jl:15-23): the ⊕ over a single term is the term, no exp / log
            const T w = nzval[beg];
            const T* src = Bj + (colval[beg] - base);
            T x[CJ];
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) x[jj] = __ldg(src + jj * ldb);
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                T v;
                if (SR == LSR_PROB) v = w * x[jj];
                else { v = w + x[jj]; v = v > lin_neg_inf<T>() ? v : lin_neg_inf<T>(); }
                C[(j0 + jj) * ldc + i] = v;
            }
            continue;
        }
        Acc<T, SR> acc[CJ];
        if (j0 + CJ <= n_cols_b) {
            for (int k = beg; k < end; ++k) {
                const T w = nzval[k];
                const T* src = Bj + (colval[k] - base);
                T x[CJ];
#pragma unroll
                for (int jj = 0; jj < CJ; ++jj) x[jj] = __ldg(src + jj * ldb);
#pragma unroll
                for (int jj = 0; jj < CJ; ++jj) acc[jj].add_prod(w, x[jj]);
            }
        } else {
            for (int k = beg; k < end; ++k) {
                const T w = nzval[k];
                const T* src = Bj + (colval[k] - base);
#pragma unroll
                for (int jj = 0; jj < CJ; ++jj)
                    if (j0 + jj < n_cols_b) acc[jj].add_prod(w, __ldg(src + jj * ldb));
            }
        }
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj) {
            if (j0 + jj >= n_cols_b) break;
            T* dst = C + (j0 + jj) * ldc + i;
            if (accumulate) acc[jj].add_value(*dst);
#ifdef MK_SPMM_STCS
            __stcs(dst, acc[jj].value());
#else
            *dst = acc[jj].value();
#endif
        }
    }
}

// ---- SpMM through a shared-memory window of B ---------------------------------------------------------------------
// spmm_kernel's gathers fetch 4 bytes of every 32-byte sector they touch: at cfg 3 (Ĉ·V̂: 3.84 M rows x 151 columns) that is
// 18.5 GB through the L2 -> L1 path for 2.6 GB of algorithmic traffic, and the 1.1 ms it takes is exactly that path's bandwidth
// (148 SMs x 64 B/clk).  The matrices of the path are block-diagonal (one block per utterance, src/fsmops.jl:28-36), so the
// columns a block of consecutive rows touches form a narrow window: spmm_window_kernel finds it per block of `rb_rows` rows, and
// spmm_staged_kernel copies that window of CJ columns of B into shared memory — coalesced, every sector used in full — as
// [column of A][CJ] so that an arc's CJ operands are one 16-byte shared-memory load.  A block whose window does not fit
// (a matrix without that structure) reads B from global memory exactly like spmm_kernel; no host decision, no synchronisation.
__global__ void spmm_window_kernel(long long n_rows, int rb_rows, const int32_t* __restrict__ rowptr,
                                   const int32_t* __restrict__ colval, int base, int* __restrict__ win) {
    __shared__ int slo[32], shi[32];
    const long long r0 = (long long)blockIdx.x * rb_rows;
    const long long r1 = r0 + rb_rows < n_rows ? r0 + rb_rows : n_rows;
    const int beg = rowptr[r0] - base, end = rowptr[r1] - base;
    // win[3b + 2]: every row of the block holds exactly one arc (Ĉ: examples/prepare-lfmmi-graphs.jl:15-23) — the row loop of
    // spmm_staged_kernel then needs no row pointers at all
    int one = 1;
    for (long long r = r0 + threadIdx.x; r < r1; r += blockDim.x) one &= (rowptr[r + 1] - rowptr[r] == 1) ? 1 : 0;
    one = __syncthreads_and(one);
    int lo = 0x7fffffff, hi = -1;
    for (int k = beg + threadIdx.x; k < end; k += blockDim.x) {
        const int c = colval[k] - base;
        lo = min(lo, c);
        hi = max(hi, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < int(blockDim.x >> 5); ++w) { lo = min(lo, slo[w]); hi = max(hi, shi[w]); }
        win[3 * blockIdx.x] = lo;
        win[3 * blockIdx.x + 1] = hi;
        win[3 * blockIdx.x + 2] = one;
    }
}

// CJ operands of one arc: from the staged window (16-byte loads) or straight from B.
template <typename T, int CJ, bool STAGED>
__device__ __forceinline__ void spmm_operands(T (&x)[CJ], const T* __restrict__ sB, const T* __restrict__ Bj, long long ldb,
                                              int col, int lo, int ncj, T zero) {
    if (STAGED) {
        constexpr int PER = 16 / int(sizeof(T));
        const uint4* src = reinterpret_cast<const uint4*>(sB + size_t(col - lo) * CJ);
#pragma unroll
        for (int q = 0; q < CJ / PER; ++q) {
            const uint4 v = src[q];
            T t[PER];
            memcpy(t, &v, 16);
#pragma unroll
            for (int e = 0; e < PER; ++e) x[q * PER + e] = t[e];
        }
    } else {
        const T* src = Bj + col;
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj) x[jj] = jj < ncj ? __ldg(src + jj * ldb) : zero;
    }
}

// One row with any number of arcs (or β = 1).  C0: this row block's part of the first column of the chunk; l: row in block.
// (Out of line: its accumulators would otherwise share the register budget of the one-arc path.)
template <typename T, int SR, int CJ, bool STAGED>
__device__ __noinline__ void spmm_one_row(int l, int beg, int end, const int32_t* __restrict__ colval,
                                          const T* __restrict__ nzval, int base, const T* __restrict__ sB, int lo,
                                          const T* __restrict__ Bj, long long ldb, T* __restrict__ C0, long long ldc, int ncj,
                                          int accumulate) {
    const T zero = SR == LSR_PROB ? T(0) : lin_neg_inf<T>();
    Acc<T, SR> acc[CJ];
    T x[CJ];
    for (int k = beg; k < end; ++k) {
        const T w = nzval[k];
        spmm_operands<T, CJ, STAGED>(x, sB, Bj, ldb, colval[k] - base, lo, ncj, zero);
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj) acc[jj].add_prod(w, x[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) {
        if (jj >= ncj) break;
        T* dst = C0 + jj * ldc + l;
        if (accumulate) acc[jj].add_value(*dst);
        __stcs(dst, acc[jj].value());
    }
}

// U rows of one thread (l0, l0 + step, ...; all inside the block).  A row is a chain of dependent loads (row pointer -> arc ->
// operand -> store): the U row pointers, then the U arcs, are requested together; the operands come from shared memory and are
// consumed row by row (holding all U x CJ of them would not fit 64 registers — measured: spills through L1, 1.13 ms at cfg 3).
template <typename T, int SR, int CJ, int U, bool STAGED, bool FULL>
__device__ __forceinline__ void spmm_row_batch(int l0, int step, const int32_t* __restrict__ rp,
                                               const int32_t* __restrict__ colval, const T* __restrict__ nzval, int base,
                                               const T* __restrict__ sB, int lo, const T* __restrict__ Bj, long long ldb,
                                               T* const (&Cr)[CJ], long long ldc, int ncj, int accumulate) {
    const T zero = SR == LSR_PROB ? T(0) : lin_neg_inf<T>();
    int beg[U], end[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { beg[u] = rp[l0 + u * step] - base; end[u] = rp[l0 + u * step + 1] - base; }
    bool single = !accumulate;
#pragma unroll
    for (int u = 0; u < U; ++u) single = single && end[u] - beg[u] == 1;
    if (single) {  // one arc per row (Ĉ): the ⊕ over a single term is the term
        int col[U];
        T w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { col[u] = colval[beg[u]] - base; w[u] = nzval[beg[u]]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            T x[CJ];
            spmm_operands<T, CJ, STAGED>(x, sB, Bj, ldb, col[u], lo, ncj, zero);
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                if (!FULL && jj >= ncj) break;
                T v;
                if (SR == LSR_PROB) v = w[u] * x[jj];
                else { v = w[u] + x[jj]; v = v > lin_neg_inf<T>() ? v : lin_neg_inf<T>(); }
                __stcs(Cr[jj] + (l0 + u * step), v);
            }
        }
        return;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        spmm_one_row<T, SR, CJ, STAGED>(l0 + u * step, beg[u], end[u], colval, nzval, base, sB, lo, Bj, ldb, Cr[0], ldc, ncj,
                                        accumulate);
}

// Row pointers of U rows of one thread; `single` = every one of them holds exactly one arc.
template <int U>
__device__ __forceinline__ void spmm_row_ptrs(const int32_t* __restrict__ rp, int l0, int step, int base, int (&beg)[U],
                                              bool& single) {
    int len[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int b0 = rp[l0 + u * step], b1 = rp[l0 + u * step + 1];
        beg[u] = b0 - base;
        len[u] = b1 - b0;
    }
    single = true;
#pragma unroll
    for (int u = 0; u < U; ++u) single = single && len[u] == 1;
}

// The rows of one block: 32-bit offsets into the block; C is written with streaming stores so that the arcs and B stay in L2 for
// the CTAs of the same row block that work on the other column chunks.  Software-pipelined by one batch: the row pointers of the
// next U rows are requested while the arcs of this batch are in flight (one L2 round trip per batch instead of two).
template <typename T, int SR, int CJ, int U, bool STAGED, bool FULL>
__device__ __forceinline__ void spmm_rows(int nrow, const int32_t* __restrict__ rp, const int32_t* __restrict__ colval,
                                          const T* __restrict__ nzval, int base, const T* __restrict__ sB, int lo,
                                          const T* __restrict__ Bj, long long ldb, T* const (&Cr)[CJ], long long ldc,
                                          int ncj, int accumulate) {
    const T zero = SR == LSR_PROB ? T(0) : lin_neg_inf<T>();
    const int step = int(blockDim.x);
    int l0 = int(threadIdx.x);
    if (!accumulate) {
        // one arc per row (Ĉ): the ⊕ over a single term is the term.  (No call in this loop: the out-of-line general row would
        // force everything that lives across it into local memory.)
        bool go = l0 + (U - 1) * step < nrow;
        int beg[U];
        if (go) spmm_row_ptrs<U>(rp, l0, step, base, beg, go);
        while (go) {
            int col[U];
            T w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { col[u] = colval[beg[u]] - base; w[u] = nzval[beg[u]]; }
            const int ln = l0 + step * U;
            go = ln + (U - 1) * step < nrow;
            if (go) spmm_row_ptrs<U>(rp, ln, step, base, beg, go);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                T x[CJ];
                spmm_operands<T, CJ, STAGED>(x, sB, Bj, ldb, col[u], lo, ncj, zero);
#pragma unroll
                for (int jj = 0; jj < CJ; ++jj) {
                    if (!FULL && jj >= ncj) break;
                    T v;
                    if (SR == LSR_PROB) v = w[u] * x[jj];
                    else { v = w[u] + x[jj]; v = v > lin_neg_inf<T>() ? v : lin_neg_inf<T>(); }
                    __stcs(Cr[jj] + (l0 + u * step), v);
                }
            }
            l0 = ln;
        }
    }
    // whatever is left — from the first batch with a row of another length on, or everything when β = 1
    for (; l0 + (U - 1) * step < nrow; l0 += step * U)
        spmm_row_batch<T, SR, CJ, U, STAGED, FULL>(l0, step, rp, colval, nzval, base, sB, lo, Bj, ldb, Cr, ldc, ncj, accumulate);
    for (; l0 < nrow; l0 += step)
        spmm_row_batch<T, SR, CJ, 1, STAGED, FULL>(l0, step, rp, colval, nzval, base, sB, lo, Bj, ldb, Cr, ldc, ncj, accumulate);
}

// Float32, a block of one-arc rows, four full columns, β = 0, everything 16-byte aligned: a thread takes FOUR ADJACENT rows — their
// arcs are four consecutive entries of colval / nzval (one 16-byte load each, no row pointers), their operands four 16-byte
// shared-memory loads, and every column gets ONE 16-byte streaming store (a warp writes 512 contiguous bytes per column).
// ~4 instructions per output element against ~9 of the row-per-thread loop.  The next quad's arcs are requested one trip ahead.
template <int SR>
__device__ __forceinline__ void spmm_rows_quad(int nrow, int a0, const int32_t* __restrict__ colval,
                                               const float* __restrict__ nzval, int base, const float* __restrict__ sB, int lo,
                                               float* const (&Cr)[4]) {
    const int nq = nrow >> 2, step = int(blockDim.x);
    const int4* c4 = reinterpret_cast<const int4*>(colval + a0);
    const float4* w4 = reinterpret_cast<const float4*>(nzval + a0);
    int q = int(threadIdx.x);
    int4 col = make_int4(0, 0, 0, 0);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) { col = c4[q]; w = w4[q]; }
    while (q < nq) {
        const int qn = q + step;
        int4 coln = col;
        float4 wn = w;
        if (qn < nq) { coln = c4[qn]; wn = w4[qn]; }
        const int cc[4] = {col.x, col.y, col.z, col.w};
        const float ww[4] = {w.x, w.y, w.z, w.w};
        float x[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float4 t = *reinterpret_cast<const float4*>(sB + size_t(cc[r] - base - lo) * 4);
            x[r][0] = t.x; x[r][1] = t.y; x[r][2] = t.z; x[r][3] = t.w;
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            float v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (SR == LSR_PROB) v[r] = ww[r] * x[r][jj];
                else { v[r] = ww[r] + x[r][jj]; v[r] = v[r] > lin_neg_inf<float>() ? v[r] : lin_neg_inf<float>(); }
            }
            __stcs(reinterpret_cast<float4*>(Cr[jj]) + q, make_float4(v[0], v[1], v[2], v[3]));
        }
        q = qn;
        col = coln;
        w = wn;
    }
}

// grid = (min(ceil(n_cols_b / CJ), 65535), row blocks): the CTAs of one row block — one per column chunk — run side by side and
// share its arcs and its window of B through L2.  Dynamic shared memory = max_window * CJ * sizeof(T).
template <typename T, int SR, int CJ, int THREADS, int U>
__global__ void __launch_bounds__(THREADS, 2)
spmm_staged_kernel(long long n_rows, int rb_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
                   const T* __restrict__ nzval, int base, const T* __restrict__ B, long long ldb, T* __restrict__ C,
                   long long ldc, long long n_cols_b, int accumulate, const int* __restrict__ win, int max_window) {
    extern __shared__ __align__(16) unsigned char spmm_smem[];
    T* sB = reinterpret_cast<T*>(spmm_smem);
    const long long r0 = (long long)blockIdx.y * rb_rows;
    const int nrow = int(r0 + rb_rows < n_rows ? rb_rows : n_rows - r0);
    const int lo = win[3 * blockIdx.y], hi = win[3 * blockIdx.y + 1];
    const bool one_arc_rows = win[3 * blockIdx.y + 2] != 0;
    const int W = hi >= lo ? hi - lo + 1 : 0;
    const bool staged = W <= max_window;  // (block-uniform)
    const T zero = SR == LSR_PROB ? T(0) : lin_neg_inf<T>();
    const int32_t* rp = rowptr + r0;
    for (long long j0 = (long long)blockIdx.x * CJ; j0 < n_cols_b; j0 += (long long)gridDim.x * CJ) {
        const int ncj = int(n_cols_b - j0 < CJ ? n_cols_b - j0 : CJ);
        const T* Bj = B + j0 * ldb;
        T* Cr[CJ];
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj) Cr[jj] = C + (j0 + (jj < ncj ? jj : 0)) * ldc + r0;
        if (staged) {
            __syncthreads();  // the readers of the previous window are done
            constexpr int PER = 16 / int(sizeof(T));
            const T* Bw = Bj + lo;
            // (four window positions = 16 loads in flight per thread: a window is 6-12 positions per thread, and taken one
            // or two at a time the copy was a chain of as many L2 / HBM round trips — 22 % of the stall samples)
            constexpr int SU = sizeof(T) == 4 ? 4 : 2;
            for (int c0 = threadIdx.x; c0 < W; c0 += THREADS * SU) {
                T x[SU][CJ];
#pragma unroll
                for (int u = 0; u < SU; ++u) {
                    const int c = c0 + u * THREADS;
#pragma unroll
                    for (int jj = 0; jj < CJ; ++jj) x[u][jj] = (c < W && jj < ncj) ? __ldg(Bw + jj * ldb + c) : zero;
                }
#pragma unroll
                for (int u = 0; u < SU; ++u) {
                    const int c = c0 + u * THREADS;
                    if (c < W) {
                        uint4* dst = reinterpret_cast<uint4*>(sB + size_t(c) * CJ);
#pragma unroll
                        for (int q = 0; q < CJ / PER; ++q) {
                            uint4 v;
                            memcpy(&v, &x[u][q * PER], 16);
                            dst[q] = v;
                        }
                    }
                }
            }
            __syncthreads();
            if constexpr (sizeof(T) == 4 && CJ == 4) {
                const int a0 = one_arc_rows ? rp[0] - base : 0;
                const bool quad = one_arc_rows && ncj == CJ && !accumulate && (ldc & 3) == 0 &&
                                  ((reinterpret_cast<uintptr_t>(Cr[0]) | reinterpret_cast<uintptr_t>(colval + a0) |
                                    reinterpret_cast<uintptr_t>(nzval + a0)) & 15) == 0;
                if (quad) {  // (block-uniform)
                    spmm_rows_quad<SR>(nrow, a0, colval, nzval, base, sB, lo, Cr);
                    // the last nrow % 4 rows of the matrix
                    const int l = (nrow & ~3) + int(threadIdx.x);
                    if (l < nrow)
                        spmm_row_batch<T, SR, CJ, 1, true, true>(l, THREADS, rp, colval, nzval, base, sB, lo, Bj, ldb, Cr, ldc, ncj, accumulate);
                    continue;
                }
            }
            if (ncj == CJ) spmm_rows<T, SR, CJ, U, true, true>(nrow, rp, colval, nzval, base, sB, lo, Bj, ldb, Cr, ldc, ncj, accumulate);
            else spmm_rows<T, SR, CJ, U, true, false>(nrow, rp, colval, nzval, base, sB, lo, Bj, ldb, Cr, ldc, ncj, accumulate);
        } else {
            spmm_rows<T, SR, CJ, U, false, false>(nrow, rp, colval, nzval, base, sB, lo, Bj, ldb, Cr, ldc, ncj, accumulate);
        }
    }
}

// dest .= 0̄ ;  dest[nzind[k]] = f(nzval[k], y[nzind[k]])   (src/linalg.jl:299-338)
template <typename T> __global__ void fill_kernel(T* dest, long long n, T v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dest[i] = v;
}
template <typename T, int SR, int OP /* 0: ⊗, 1: ⊘ */>
__global__ void spvec_bcast_kernel(long long nnz, const int32_t* __restrict__ nzind, const T* __restrict__ nzval,
                                   int base, const T* __restrict__ y, T* __restrict__ dest) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
        const long long i = nzind[k] - base;
        const T a = nzval[k], v = y[i];
        T r;
        if (SR == LSR_PROB) r = OP == 0 ? a * v : a / v;
        else r = OP == 0 ? a + v : a - v;
        dest[i] = r;
    }
}

}  // namespace mk
