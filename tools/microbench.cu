// SPDX-License-Identifier: MIT
// microbench.cu — measures the three on-chip ceilings the shared-graph kernel is designed
// around: MUFU ex2 rate, L2 gather bandwidth for 512 B segments, grid-barrier cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

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

// every warp gathers `n` 512-byte rows (row = idx[i]) with one 16 B load per lane, 8 in flight
__global__ void gather_kernel(const float4* __restrict__ vec, const int* __restrict__ idx, int n_per_warp,
                              float* out) {
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

__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
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

    // 1. MUFU ex2
    for (int rep = 0; rep < 3; ++rep) {
        int iters = 4096, blocks = sms * 2, threads = 1024;
        CK(cudaEventRecord(e0));
        ex2_kernel<<<blocks, threads>>>(out, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double ops = double(blocks) * threads * iters * 8;
        printf("ex2: %.3f ms, %.2f Tops/s (%.1f per clk per SM at %d MHz)\n", ms, ops / ms / 1e9,
               ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    }
    // 2. L2 gather of 512 B rows
    for (int rows : {15000, 30001, 60000, 240000}) {
        size_t bytes = size_t(rows) * 512;
        float4* vec;
        CK(cudaMalloc(&vec, bytes));
        CK(cudaMemset(vec, 0, bytes));
        for (int threads : {512, 1024}) {
            int warps = sms * threads / 32, n_per_warp = 2048;
            std::vector<int> h(size_t(warps) * n_per_warp);
            unsigned s = 12345;
            for (auto& x : h) { s = s * 1664525u + 1013904223u; x = (s >> 8) % rows; }
            int* idx;
            CK(cudaMalloc(&idx, h.size() * 4));
            CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                gather_kernel<<<sms, threads>>>(vec, idx, n_per_warp, out);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            double gb = double(warps) * n_per_warp * 512 / 1e9;
            printf("gather: %6d rows (%.1f MB), %4d thr/SM: %.3f ms, %.0f GB/s\n", rows, bytes / 1e6, threads, ms, gb / (ms * 1e-3));
            CK(cudaFree(idx));
        }
        CK(cudaFree(vec));
    }
    // 3. grid barrier
    unsigned* ctr;
    CK(cudaMalloc(&ctr, 4));
    for (int threads : {256, 512, 1024}) {
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
