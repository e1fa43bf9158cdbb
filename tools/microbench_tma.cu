// SPDX-License-Identifier: MIT
// microbench_tma.cu — which way of gathering 512-byte state-vector rows out of L2 feeds an SM best?
//   (a) register-destination LDG.128 gathers, 8 rows in flight per warp (the round-1 shared_fb_kernel);
//   (b) cp.async.bulk (TMA 1-D) into a per-warp shared-memory ring with mbarrier completion and a lookahead of
//       M-1 passes, consumed with LDS.128 + FFMA (the round-2 design);
// plus the MUFU ex2 rate (the SFU half of the roofline) and the grid-barrier cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_tma tools/microbench_tma.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// (b) every warp: passes of ROWS rows; lanes 0..ROWS-1 each issue one 512 B bulk copy; M slots per warp
template <int ROWS>
__global__ void __launch_bounds__(1024, 1) bulk_gather_kernel(const float4* __restrict__ vec, const int* __restrict__ idx,
                                                              int n_per_warp, int M, float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t slot_bytes = size_t(ROWS) * 512;
    unsigned char* ring = smem + size_t(warp) * M * slot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(nw) * M * slot_bytes) + warp * 8;
    if (lane == 0)
        for (int s = 0; s < M; ++s) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int* my = idx + size_t(gw) * n_per_warp;
    const int npass = n_per_warp / ROWS;
    auto issue = [&](int j, int slot) {
        const uint32_t bar = smem_u32(bars + slot);
        if (lane == 0) mbar_expect_tx(bar, ROWS * 512);
        __syncwarp();
        if (lane < ROWS) {
            const int r = __ldg(my + j * ROWS + lane);
            bulk_g2s(smem_u32(ring + size_t(slot) * slot_bytes + lane * 512), vec + size_t(r) * 32, 512, bar);
        }
    };
    int islot = 0;
    for (int j = 0; j < M - 1 && j < npass; ++j) { issue(j, islot); islot = islot + 1 == M ? 0 : islot + 1; }
    float acc[4] = {0, 0, 0, 0};
    int cslot = 0;
    uint32_t parity = 0;
    for (int j = 0; j < npass; ++j) {
        if (j + M - 1 < npass) { issue(j + M - 1, islot); islot = islot + 1 == M ? 0 : islot + 1; }
        const uint32_t bar = smem_u32(bars + cslot);
        while (!mbar_try_wait(bar, parity)) {}
        const float4* rows = reinterpret_cast<const float4*>(ring + size_t(cslot) * slot_bytes) + lane;
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
            const float4 v = rows[k * 32];
            acc[0] = fmaf(v.x, 0.5f, acc[0]); acc[1] = fmaf(v.y, 0.5f, acc[1]);
            acc[2] = fmaf(v.z, 0.5f, acc[2]); acc[3] = fmaf(v.w, 0.5f, acc[3]);
        }
        __syncwarp();
        if (++cslot == M) { cslot = 0; parity ^= 1; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

// (a) register-destination gathers, 8 in flight per warp
__global__ void ldg_gather_kernel(const float4* __restrict__ vec, const int* __restrict__ idx, int n_per_warp, float* out) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int* my = idx + size_t(gw) * n_per_warp;
    float acc = 0;
    for (int i = 0; i < n_per_warp; i += 8) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcg(vec + size_t(__ldg(my + i + k)) * 32 + lane);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// (c) cp.async (LDGSTS.128, per-lane addresses: one warp instruction copies a 512 B row) into a per-warp ring,
//     commit_group / wait_group per pass; every lane reads back exactly the 16 bytes it copied (no barrier at all)
template <int ROWS, int M>
__global__ void __launch_bounds__(1024, 1) ldgsts_gather_kernel(const float4* __restrict__ vec, const int* __restrict__ idx,
                                                                int n_per_warp, float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t slot_bytes = size_t(ROWS) * 512;
    const uint32_t ring = smem_u32(smem + size_t(warp) * M * slot_bytes) + lane * 16;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int* my = idx + size_t(gw) * n_per_warp;
    const int npass = n_per_warp / ROWS;
    auto issue = [&](int j, int slot) {
        if (j < npass) {
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
                const int r = __ldg(my + j * ROWS + k);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + uint32_t(slot * slot_bytes) + k * 512),
                             "l"(vec + size_t(r) * 32 + lane) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int islot = 0;
    for (int j = 0; j < M - 1; ++j) { issue(j, islot); islot = islot + 1 == M ? 0 : islot + 1; }
    float acc[4] = {0, 0, 0, 0};
    int cslot = 0;
    for (int j = 0; j < npass; ++j) {
        issue(j + M - 1, islot); islot = islot + 1 == M ? 0 : islot + 1;
        asm volatile("cp.async.wait_group %0;" ::"n"(M - 1) : "memory");
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(ring + uint32_t(cslot * slot_bytes) + k * 512) : "memory");
            acc[0] = fmaf(v.x, 0.5f, acc[0]); acc[1] = fmaf(v.y, 0.5f, acc[1]);
            acc[2] = fmaf(v.z, 0.5f, acc[2]); acc[3] = fmaf(v.w, 0.5f, acc[3]);
        }
        if (++cslot == M) cslot = 0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

// (d) register-destination gathers, software-pipelined: pass j+1 (8 rows) is requested before pass j is folded
__global__ void ldg2_gather_kernel(const float4* __restrict__ vec, const int* __restrict__ idx, int n_per_warp, float* out) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int* my = idx + size_t(gw) * n_per_warp;
    float acc = 0;
    float4 v[8], w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(vec + size_t(__ldg(my + k)) * 32 + lane);
    for (int i = 8; i < n_per_warp; i += 16) {
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __ldcg(vec + size_t(__ldg(my + i + k)) * 32 + lane);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
        if (i + 8 < n_per_warp) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldcg(vec + size_t(__ldg(my + i + 8 + k)) * 32 + lane);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += w[k].x + w[k].y + w[k].z + w[k].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void ex2_kernel(float* out, int iters) {
    float a[8];
    for (int k = 0; k < 8; ++k) a[k] = -0.001f * (threadIdx.x + k);
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
    float s = 0;
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
    }
    __syncthreads();
}
__global__ void sync_kernel(unsigned* ctr, int iters) {
    unsigned target = 0;
    for (int i = 0; i < iters; ++i) grid_sync(ctr, target);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount, clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max clock %d MHz\n", prop.name, sms, clk / 1000);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    float* out;
    CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024));

    // 1. MUFU ex2 (the SFU peak the roofline's SFU term is measured against)
    for (int rep = 0; rep < 3; ++rep) {
        int iters = 4096, blocks = sms * 2, threads = 1024;
        CK(cudaEventRecord(e0));
        ex2_kernel<<<blocks, threads>>>(out, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double ops = double(blocks) * threads * iters * 8;
        printf("ex2: %.3f ms, %.3f Tops/s (%.2f per clk per SM at %d MHz)\n", ms, ops / ms / 1e9,
               ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    }
    // 2. gathers of 512 B rows out of a 30k-row (15 MB, L2-resident) vector
    const int rows = 30001;
    size_t bytes = size_t(rows) * 512;
    float4* vec;
    CK(cudaMalloc(&vec, bytes));
    CK(cudaMemset(vec, 0, bytes));
    const int n_per_warp = 4096;
    std::vector<int> h(size_t(sms) * 32 * n_per_warp);
    unsigned s = 12345;
    for (auto& x : h) { s = s * 1664525u + 1013904223u; x = (s >> 8) % rows; }
    int* idx;
    CK(cudaMalloc(&idx, h.size() * 4));
    CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    for (int threads : {256, 384, 512, 768, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            ldg_gather_kernel<<<sms, threads>>>(vec, idx, n_per_warp, out);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        double gb = double(sms) * (threads / 32) * n_per_warp * 512 / 1e9;
        printf("ldg gather : %4d thr/SM (8 rows in flight per warp): %.3f ms, %.0f GB/s\n", threads, ms, gb / (ms * 1e-3));
    }
    for (int threads : {256, 384, 512, 768}) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            ldg2_gather_kernel<<<sms, threads>>>(vec, idx, n_per_warp, out);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        double gb = double(sms) * (threads / 32) * n_per_warp * 512 / 1e9;
        printf("ldg2 gather: %4d thr/SM (16 rows in flight per warp): %.3f ms, %.0f GB/s\n", threads, ms, gb / (ms * 1e-3));
    }
    {
        auto run_c = [&](auto kern, int rowsp, int M, int threads) {
            const int nw = threads / 32;
            size_t smem = size_t(nw) * M * rowsp * 512;
            if (smem > 200 * 1024) return;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                kern<<<sms, threads, smem>>>(vec, idx, n_per_warp, out);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            double gb = double(sms) * nw * n_per_warp * 512 / 1e9;
            printf("ldgsts gather: %4d thr/SM, %2d rows/pass, %d slots (%3zu KB smem): %.3f ms, %.0f GB/s\n", threads, rowsp, M,
                   smem / 1024, ms, gb / (ms * 1e-3));
        };
        for (int threads : {256, 384, 512, 768}) {
            run_c(ldgsts_gather_kernel<8, 2>, 8, 2, threads);
            run_c(ldgsts_gather_kernel<8, 3>, 8, 3, threads);
            run_c(ldgsts_gather_kernel<8, 4>, 8, 4, threads);
            run_c(ldgsts_gather_kernel<4, 4>, 4, 4, threads);
            run_c(ldgsts_gather_kernel<4, 6>, 4, 6, threads);
        }
    }
    CK(cudaFuncSetAttribute(bulk_gather_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(bulk_gather_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(bulk_gather_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int rowsp : {4, 8, 16})
        for (int threads : {256, 384, 512, 768}) {
            for (int M : {2, 3, 4, 6}) {
                const int nw = threads / 32;
                size_t smem = size_t(nw) * M * rowsp * 512 + size_t(nw) * 64;
                if (smem > 200 * 1024) continue;
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaEventRecord(e0));
                    if (rowsp == 4) bulk_gather_kernel<4><<<sms, threads, smem>>>(vec, idx, n_per_warp, M, out);
                    else if (rowsp == 8) bulk_gather_kernel<8><<<sms, threads, smem>>>(vec, idx, n_per_warp, M, out);
                    else bulk_gather_kernel<16><<<sms, threads, smem>>>(vec, idx, n_per_warp, M, out);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
                }
                double gb = double(sms) * nw * n_per_warp * 512 / 1e9;
                printf("bulk gather: %4d thr/SM, %2d rows/pass, %d slots (%3zu KB smem): %.3f ms, %.0f GB/s\n", threads, rowsp, M,
                       smem / 1024, ms, gb / (ms * 1e-3));
            }
        }
    // 3. grid barrier
    unsigned* ctr;
    CK(cudaMalloc(&ctr, 4));
    for (int threads : {256, 384, 512, 1024}) {
        CK(cudaMemset(ctr, 0, 4));
        int iters = 2000;
        void* args[] = {&ctr, &iters};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void*)sync_kernel, dim3(sms), dim3(threads), args, 0, 0));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("grid barrier: %d CTAs x %d threads: %.2f us per barrier\n", sms, threads, ms * 1e3 / iters);
    }
    return 0;
}
