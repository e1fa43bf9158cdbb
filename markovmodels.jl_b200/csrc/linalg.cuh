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
// src/linalg.jl:213-233; rows of the path's graphs hold ~17 arcs, so the host picks 4/8/32 from nnz / rows).
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
    __shared__ T sm[8], ss[8];
    const int count = min(worklist[0], worklist_cap);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = blockIdx.x; i < count; i += gridDim.x) {
        const int r = worklist[1 + i];
        const int beg = rowptr[r] - base, end = rowptr[r + 1] - base;
        Acc<T, SR> acc;
        for (int k = beg + threadIdx.x; k < end; k += blockDim.x) acc.add_prod(nzval[k], b[colval[k] - base]);
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
