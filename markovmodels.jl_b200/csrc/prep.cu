// SPDX-License-Identifier: MIT
//
// prep.cu — the graph-preparation operators of src/linalg.jl on caller-owned DEVICE arrays:
//   blockdiag(::CuSparseMatrixCSC...) / blockdiag(::CuSparseMatrixCSR...)      src/linalg.jl:73-131
//   vcat(::CuSparseVector...)                                                  src/linalg.jl:137-157
//   CuSparseMatrixCSR(::CuSparseMatrixCSC), CuSparseMatrixCSC(::CuSparseMatrixCSR), copy(Mᵀ)   src/linalg.jl:12-67
// The fused inference path never needs them (a batch is a descriptor there); they exist for callers of `mul!` itself,
// who prepare T̂ᵀ, Ĉ, Ĉᵀ and the block-diagonal batch the way the reference does.  The reference issues 3 small copies
// with a scalar offset per block (3·B launches for a batch); here one launch concatenates every block, offsets applied
// on the fly.  The payload is opaque here (4- or 8-byte elements): these operators move semiring values, they never
// combine them.
#include "../../include/markov_b200.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cstdint>
#include <cstring>
#include <vector>

int mk_set_error(int code, const char* fmt, ...);  // markov_b200.cu: mk_last_error's message
void mk_note_launches(int n);                      // markov_b200.cu: mk_launch_count's counter

namespace {

#define PCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return mk_set_error(MK_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// one piece of a concatenation: `count` elements from `src` to `dst`; int32 pieces get `add` added
struct Piece {
    const void* src;
    void* dst;
    long long count;
    int add;
    int is_index;  // 1: int32 with offset; 0: payload of esize bytes
};

// grid (chunks, n_pieces): a piece is swept by its row of blocks
__global__ void concat_kernel(const Piece* pieces, int esize) {
    const Piece p = pieces[blockIdx.y];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.count; i += stride) {
        if (p.is_index) {
            static_cast<int*>(p.dst)[i] = static_cast<const int*>(p.src)[i] + p.add;
        } else if (esize == 4) {
            static_cast<uint32_t*>(p.dst)[i] = static_cast<const uint32_t*>(p.src)[i];
        } else {
            static_cast<uint64_t*>(p.dst)[i] = static_cast<const uint64_t*>(p.src)[i];
        }
    }
}

int run_pieces(const std::vector<Piece>& pieces, int esize, cudaStream_t st) {
    if (pieces.empty()) return MK_OK;
    Piece* d = nullptr;
    PCK(cudaMallocAsync(&d, pieces.size() * sizeof(Piece), st));
    // (pageable source: the runtime stages it before returning, the vector may die afterwards)
    PCK(cudaMemcpyAsync(d, pieces.data(), pieces.size() * sizeof(Piece), cudaMemcpyHostToDevice, st));
    long long longest = 1;
    for (const Piece& p : pieces) longest = p.count > longest ? p.count : longest;
    const int chunks = int(std::min<long long>((longest + 255) / 256, 1024));
    for (size_t p0 = 0; p0 < pieces.size(); p0 += 65535) {
        const unsigned np = unsigned(std::min<size_t>(65535, pieces.size() - p0));
        concat_kernel<<<dim3(chunks, np), 256, 0, st>>>(d + p0, esize);
        PCK(cudaGetLastError());
        mk_note_launches(1);
    }
    PCK(cudaFreeAsync(d, st));
    return MK_OK;
}

// ---- CSR(A) -> CSR(Aᵀ)  (= CSC(A) read as CSR of the transpose) ------------------------------------------------
// row_of[k] for every stored element k
__global__ void expand_rows_kernel(const int* ptr, int base, long long n_rows, int* row_of) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x)
        for (int k = ptr[r] - base; k < ptr[r + 1] - base; ++k) row_of[k] = int(r);
}
__global__ void iota_kernel(int* x, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = int(i);
}
// after the stable sort by column: element j of the transpose is source element perm[j]
__global__ void gather_kernel(const int* perm, const int* row_of, const void* val, int esize, int base, long long nnz,
                              int* out_idx, void* out_val) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (long long)gridDim.x * blockDim.x) {
        const int k = perm[j];
        out_idx[j] = row_of[k] + base;
        if (esize == 4) static_cast<uint32_t*>(out_val)[j] = static_cast<const uint32_t*>(val)[k];
        else static_cast<uint64_t*>(out_val)[j] = static_cast<const uint64_t*>(val)[k];
    }
}
// out_ptr[c] = first position whose (sorted) column is >= c + base   (c = 0 .. n_cols)
__global__ void ptr_kernel(const int* sorted_cols, long long nnz, long long n_cols, int base, int* out_ptr) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c <= n_cols; c += (long long)gridDim.x * blockDim.x) {
        long long lo = 0, hi = nnz;
        const int key = int(c) + base;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (sorted_cols[mid] < key) lo = mid + 1;
            else hi = mid;
        }
        out_ptr[c] = int(lo) + base;
    }
}

int blocks_for(long long n) { return int(std::min<long long>(std::max<long long>((n + 255) / 256, 1), 148 * 16)); }

}  // namespace

extern "C" {

int mk_blockdiag(int dtype, int64_t n_mats, const int32_t* const* ptr, const int32_t* const* idx, const void* const* val,
                 const int64_t* dim_ptr, const int64_t* dim_idx, const int64_t* nnz, int index_base, int32_t* out_ptr,
                 int32_t* out_idx, void* out_val, void* stream) {
    if (dtype != MK_F32 && dtype != MK_F64) return mk_set_error(MK_EINVAL, "bad dtype %d", dtype);
    if (n_mats < 0 || (n_mats > 0 && (!ptr || !idx || !val || !dim_ptr || !dim_idx || !nnz)) || !out_ptr)
        return mk_set_error(MK_EINVAL, "null argument");
    if (index_base != 0 && index_base != 1) return mk_set_error(MK_EINVAL, "index_base must be 0 or 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int esize = dtype == MK_F32 ? 4 : 8;
    static const int kBase[2] = {0, 1};
    if (n_mats == 0) {  // an empty pointer array: one entry
        PCK(cudaMemcpyAsync(out_ptr, &kBase[index_base], sizeof(int), cudaMemcpyHostToDevice, st));
        return MK_OK;
    }
    std::vector<Piece> pieces;
    long long nnz_sofar = 0, p_sofar = 0, i_sofar = 0;
    for (int64_t i = 0; i < n_mats; ++i) {
        if (dim_ptr[i] < 0 || dim_idx[i] < 0 || nnz[i] < 0 || !ptr[i]) return mk_set_error(MK_EINVAL, "bad block %lld", (long long)i);
        if (nnz_sofar + nnz[i] > 0x7fffffffLL - 1 || i_sofar + dim_idx[i] > 0x7fffffffLL - 1)
            return mk_set_error(MK_ENOTSUP, "the block-diagonal matrix needs more than Cint indices");
        if (nnz[i] > 0 && (!idx[i] || !val[i] || !out_idx || !out_val)) return mk_set_error(MK_EINVAL, "null index/value array");
        // the block's pointer array without its last entry (the next block's first entry takes that place); the very
        // last entry of the result — colPtr[n+1] = nnz + 1, src/linalg.jl:97,128 — is the last block's last entry
        const bool last = i + 1 == n_mats;
        pieces.push_back({ptr[i], out_ptr + p_sofar, dim_ptr[i] + (last ? 1 : 0), int(nnz_sofar), 1});
        if (nnz[i] > 0) {
            pieces.push_back({idx[i], out_idx + nnz_sofar, nnz[i], int(i_sofar), 1});
            pieces.push_back({val[i], static_cast<char*>(out_val) + size_t(nnz_sofar) * esize, nnz[i], 0, 0});
        }
        nnz_sofar += nnz[i]; p_sofar += dim_ptr[i]; i_sofar += dim_idx[i];
    }
    return run_pieces(pieces, esize, st);
}

int mk_vcat_spvec(int dtype, int64_t n_vecs, const int32_t* const* nzind, const void* const* nzval, const int64_t* len,
                  const int64_t* nnz, int32_t* out_ind, void* out_val, void* stream) {
    if (dtype != MK_F32 && dtype != MK_F64) return mk_set_error(MK_EINVAL, "bad dtype %d", dtype);
    if (n_vecs < 0 || (n_vecs > 0 && (!nzind || !nzval || !len || !nnz))) return mk_set_error(MK_EINVAL, "null argument");
    const int esize = dtype == MK_F32 ? 4 : 8;
    std::vector<Piece> pieces;
    long long n_sofar = 0, nnz_sofar = 0;
    for (int64_t i = 0; i < n_vecs; ++i) {
        if (len[i] < 0 || nnz[i] < 0 || nnz[i] > len[i]) return mk_set_error(MK_EINVAL, "bad vector %lld", (long long)i);
        if (n_sofar + len[i] > 0x7fffffffLL - 1) return mk_set_error(MK_ENOTSUP, "the concatenation needs more than Cint indices");
        if (nnz[i] > 0) {
            if (!nzind[i] || !nzval[i] || !out_ind || !out_val) return mk_set_error(MK_EINVAL, "null index/value array");
            pieces.push_back({nzind[i], out_ind + nnz_sofar, nnz[i], int(n_sofar), 1});
            pieces.push_back({nzval[i], static_cast<char*>(out_val) + size_t(nnz_sofar) * esize, nnz[i], 0, 0});
        }
        n_sofar += len[i]; nnz_sofar += nnz[i];
    }
    return run_pieces(pieces, esize, static_cast<cudaStream_t>(stream));
}

int mk_sparse_transpose(int dtype, int64_t n_ptr, int64_t n_idx, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                        const void* val, int index_base, int32_t* out_ptr, int32_t* out_idx, void* out_val, void* stream) {
    if (dtype != MK_F32 && dtype != MK_F64) return mk_set_error(MK_EINVAL, "bad dtype %d", dtype);
    if (n_ptr < 0 || n_idx < 0 || nnz < 0 || !ptr || !out_ptr) return mk_set_error(MK_EINVAL, "bad arguments");
    if (index_base != 0 && index_base != 1) return mk_set_error(MK_EINVAL, "index_base must be 0 or 1");
    if (nnz > 0 && (!idx || !val || !out_idx || !out_val)) return mk_set_error(MK_EINVAL, "null index/value array");
    if (nnz > 0x7fffffffLL - 1 || n_ptr > 0x7fffffffLL - 1 || n_idx > 0x7fffffffLL - 1)
        return mk_set_error(MK_ENOTSUP, "sizes beyond Cint");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int esize = dtype == MK_F32 ? 4 : 8;
    if (nnz == 0) {
        ptr_kernel<<<blocks_for(n_idx + 1), 256, 0, st>>>(nullptr, 0, n_idx, index_base, out_ptr);
        PCK(cudaGetLastError());
        mk_note_launches(1);
        return MK_OK;
    }
    // stable radix sort of the stored elements by their index (column of a CSR matrix): inside a column the original
    // order — ascending rows — survives, which is what CUSPARSE's csr2csc delivers to the reference
    int *row_of = nullptr, *perm_in = nullptr, *perm_out = nullptr, *keys_out = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    int bits = 1;
    while ((1LL << bits) < n_idx + index_base + 1 && bits < 32) ++bits;
    PCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, idx, keys_out, perm_in, perm_out, int(nnz), 0, bits, st));
    PCK(cudaMallocAsync(&row_of, size_t(nnz) * 4 * sizeof(int) + tmp_bytes + 256, st));
    perm_in = row_of + nnz; perm_out = perm_in + nnz; keys_out = perm_out + nnz;
    tmp = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(keys_out + nnz) + 255) & ~uintptr_t(255));
    expand_rows_kernel<<<blocks_for(n_ptr), 256, 0, st>>>(ptr, index_base, n_ptr, row_of);
    iota_kernel<<<blocks_for(nnz), 256, 0, st>>>(perm_in, nnz);
    PCK(cudaGetLastError());
    PCK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, idx, keys_out, perm_in, perm_out, int(nnz), 0, bits, st));
    gather_kernel<<<blocks_for(nnz), 256, 0, st>>>(perm_out, row_of, val, esize, index_base, nnz, out_idx, out_val);
    ptr_kernel<<<blocks_for(n_idx + 1), 256, 0, st>>>(keys_out, nnz, n_idx, index_base, out_ptr);
    PCK(cudaGetLastError());
    mk_note_launches(4);  // (+ the library sort's own kernels)
    PCK(cudaFreeAsync(row_of, st));
    return MK_OK;
}

}  // extern "C"
