// SPDX-License-Identifier: MIT
// How fast can the CTAs of a thread-block cluster gather 512-byte rows out of each other's shared memory?  The question behind a
// cluster / DSMEM variant of the shared-graph kernel (state vector of the frame in flight spread over the shared memory of a
// cluster instead of L2): its per-arc operation is exactly this gather.  Compared with the same gather from L2-resident global
// memory (what the kernel does today; tools/microbench_tma.cu measures that one in more variants).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_dsmem tools/microbench_dsmem.cu && tools/microbench_dsmem
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

constexpr int kThreads = 512, kRowFloats = 128, kRows = 192;  // 192 rows x 512 B = 96 KB of shared memory per CTA
constexpr int kIters = 2048, kUnroll = 8;

__device__ __forceinline__ unsigned mix(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// mode 0: rows of the CTA's own shared memory; 1: rows of a random CTA of the cluster (ld.shared::cluster); 2: rows of an
// L2-resident global array (15 MB, like the kernel's state vector)
template <int MODE>
__global__ void __launch_bounds__(kThreads) gather_kernel(const float* __restrict__ gsrc, int grows, float* out, int cs) {
    extern __shared__ __align__(16) float rows[];
    cg::cluster_group cluster = cg::this_cluster();
    for (int i = threadIdx.x; i < kRows * kRowFloats; i += kThreads) rows[i] = float(i & 7);
    cluster.sync();
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < kIters; it += kUnroll) {
        float4 v[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            const unsigned h = mix(unsigned(warp) * 9781u + unsigned(it + k));
            if (MODE == 2) {
                v[k] = __ldcg(reinterpret_cast<const float4*>(gsrc + size_t(h % unsigned(grows)) * kRowFloats) + lane);
            } else {
                const float* base = rows;
                if (MODE == 1) base = cluster.map_shared_rank(rows, (h >> 20) % unsigned(cs));
                v[k] = *(reinterpret_cast<const float4*>(base + size_t(h % unsigned(kRows)) * kRowFloats) + lane);
            }
        }
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
    }
    cluster.sync();  // (no CTA may exit while a neighbour still reads its shared memory)
    if (acc.x + acc.y + acc.z + acc.w == -1.f) out[0] = acc.x;
}

template <int MODE> static float run(int cs, const float* gsrc, int grows, float* out, int grid) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kRows * kRowFloats * sizeof(float);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cfg.dynamicSmemBytes));
    if (cs > 8) cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchKernelEx(&cfg, gather_kernel<MODE>, gsrc, grows, out, cs);
        cudaEventRecord(e1);
        if (err != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) {
            printf("  launch failed (cluster %d): %s\n", cs, cudaGetErrorString(cudaGetLastError()));
            return -1.f;
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grows = 30001;  // 15.4 MB: the state vector of cfg 3 (30 001 states x 128 utterances x 4 B)
    float *g, *out;
    cudaMalloc(&g, size_t(grows) * kRowFloats * sizeof(float)); cudaMemset(g, 0, size_t(grows) * kRowFloats * sizeof(float));
    cudaMalloc(&out, 64);
    printf("%d SMs; one CTA of %d threads per SM, every warp gathers %d rows of 512 B, %d in flight per lane\n", sms, kThreads, kIters, kUnroll);
    for (int cs : {1, 2, 4, 8, 16}) {
        const int grid = sms / cs * cs;
        const double bytes = double(grid) * (kThreads / 32) * kIters * 512.0;
        const float t0 = run<0>(cs, g, grows, out, grid), t1 = cs > 1 ? run<1>(cs, g, grows, out, grid) : -1.f,
                    t2 = run<2>(cs, g, grows, out, grid);
        printf("cluster %2d (%3d CTAs): own shared memory %7.2f TB/s   cluster shared memory %7.2f TB/s   L2-resident global %7.2f TB/s\n",
               cs, grid, t0 > 0 ? bytes / t0 * 1e-9 : 0.0, t1 > 0 ? bytes / t1 * 1e-9 : 0.0, t2 > 0 ? bytes / t2 * 1e-9 : 0.0);
    }
    return 0;
}
