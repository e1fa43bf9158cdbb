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
};

// c[r] = ⊕_k nzval[k] ⊗ b[colval[k]] — LANES lanes per row (the reference spends a whole warp per row,
// src/linalg.jl:213-233; rows of the path's graphs hold ~17 arcs, so the host picks 4/8/32 from nnz / rows).
template <typename T, int SR, int LANES>
__global__ void spmv_kernel(long long n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
                            const T* __restrict__ nzval, int base, const T* __restrict__ b, T* __restrict__ c) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = int(gid % LANES);
    const long long rows_per_pass = (long long)gridDim.x * blockDim.x / LANES;
    // every lane of a warp runs the same number of passes: the shuffles below are warp-wide
    const long long passes = (n_rows + rows_per_pass - 1) / rows_per_pass;
    long long r = gid / LANES;
    for (long long it = 0; it < passes; ++it, r += rows_per_pass) {
        Acc<T, SR> acc;
        if (r < n_rows) {
            const int beg = rowptr[r] - base, end = rowptr[r + 1] - base;
            int k = beg + lane;
            for (; k + LANES < end; k += 2 * LANES) {  // two arcs per trip: four loads and two gathers in flight
                const int c0 = colval[k], c1 = colval[k + LANES];
                const T w0 = nzval[k], w1 = nzval[k + LANES];
                const T x0 = b[c0 - base], x1 = b[c1 - base];
                acc.add_prod(w0, x0);
                acc.add_prod(w1, x1);
            }
            if (k < end) acc.add_prod(nzval[k], b[colval[k] - base]);
        }
        acc.template reduce<LANES>();
        if (lane == 0 && r < n_rows) c[r] = acc.value();
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
        Acc<T, SR> acc[CJ];
        const T* Bj = B + j0 * ldb;
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
            *dst = acc[jj].value();
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
