// SPDX-License-Identifier: MIT
//
// kernels.cuh — sm_100a device code of libmarkov_b200.so.
//
// Two kernel families implement the reference's time recursions
// (src/inference.jl:62-74 αrecursion, :99-110 βrecursion, :145-161 pdfposteriors)
// with the host-side frame loop moved on to the device:
//
//  * shared_fb_kernel  — a group of U utterances that share ONE graph (the replicated
//    LF-MMI denominator).  Persistent cooperative grid, one CTA per SM.  State vectors
//    live in global memory as [state][utterance] so that one arc (src -> dst, w) is
//    loaded once and applied to 128 utterances: lane l owns utterances 4l..4l+3 and
//    fetches them with one 16-byte load, a warp therefore gathers one contiguous 512 B
//    segment per arc (L2-resident: two frames of state are 2*Ŝ*U*4 B ≈ 31 MB ≪ 126 MB).
//    No cross-lane reduction is needed; the ⊕ over a row's arcs runs per lane.  Rows are
//    split over warps by arc count.  One grid barrier per frame (the vectors are NOT resident in shared memory or
//    cluster DSMEM: 2 x 15 MB per frame pair against 227 KB per SM / 3.6 MB per 16-CTA cluster).
//    Replaces K1 (SpMV per frame), K2 (Ĉ·V̂ gather, Ĉᵀ scatter-reduce), K3 (α̂ ⊙ e₁), K4
//    (per-frame broadcasts, γ, exp, sums) of SURVEY.md §2.2.
//
//  * small_fb_kernel — one CTA per utterance for graphs that fit shared memory (LF-MMI
//    numerators: a few hundred states each, all distinct).  α/β vectors ping-pong in
//    shared memory, arcs come through L1, one __syncthreads per frame.
//
// Semiring arithmetic (Semirings.jl).  Shared-graph kernel, Log: every stored vector exists as a normalised
// log2 row and as its linear copy 2^(v + H); a row's ⊕ is Σ lin[neighbour] · W with linear arc weights — one FFMA (half
// a packed fma.rn.f32x2) per arc and utterance, MUFU work per state only (lg2 of the sum, ex2 for the linear copy and for
// γ) — and a row whose linear sum underflows is redone exactly from the log2 copies (two-pass max + Σ ex2).  Per-utterance
// kernel and the exact path: max + log(Σ exp(x - max)).  Tropical ⊕ = max.  ⊗ = +.  -Inf is the semiring zero and never
// produces NaN (the max of an all--Inf row is replaced by 0 before subtracting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mk {

enum { SR_LOG = 0, SR_TROP = 1 };

#ifndef MK_THREADS
#define MK_THREADS 512
#endif
constexpr int kSharedThreads = MK_THREADS;  // 16 warps / CTA, 1 CTA / SM (128 registers per thread: one pass of gathers must not spill)
constexpr int kSharedWarps = kSharedThreads / 32;
constexpr int kChunk = 8;       // arcs cached in registers per ⊕ chunk
constexpr int kTileUtts = 128;  // utterances covered by one warp pass (32 lanes x 4)

template <typename T> struct Arc {
    int idx;  // source state (in-arcs, forward) or destination state (out-arcs, backward)
    T w;
};
static_assert(sizeof(Arc<float>) == 8, "arc f32 = 8 B");
static_assert(sizeof(Arc<double>) == 16, "arc f64 = 16 B");

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return __int_as_float(0xff800000); }
template <> __device__ __forceinline__ double neg_inf<double>() {
    return __longlong_as_double(0xfff0000000000000LL);
}

// ---- scalar math ---------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float exp_(float x) { return ex2_approx(x * 1.4426950408889634f); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float log_(float x) { return lg2_approx(x) * 0.6931471805599453f; }
__device__ __forceinline__ double log_(double x) { return log(x); }
// log2-domain variants: the shared-graph kernel stores its Log-semiring vectors in log2 units so that
// ⊕ needs no multiply around ex2 / lg2
__device__ __forceinline__ float ex2_(float x) { return ex2_approx(x); }
__device__ __forceinline__ double ex2_(double x) { return exp2(x); }
__device__ __forceinline__ float lg2_(float x) { return lg2_approx(x); }
__device__ __forceinline__ double lg2_(double x) { return log2(x); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }

// ---- arc and 4-wide vector access --------------------------------------------------------------
__device__ __forceinline__ Arc<float> ld_arc(const Arc<float>* p) {
    int2 t = __ldg(reinterpret_cast<const int2*>(p));
    Arc<float> a; a.idx = t.x; a.w = __int_as_float(t.y); return a;
}
__device__ __forceinline__ Arc<double> ld_arc(const Arc<double>* p) {
    int4 t = __ldg(reinterpret_cast<const int4*>(p));
    Arc<double> a; a.idx = t.x; a.w = __hiloint2double(t.w, t.z); return a;
}

// generic-address variants: the shared-graph kernel keeps its arcs in shared memory when they fit
__device__ __forceinline__ Arc<float> ld_arc_g(const Arc<float>* p) {
    int2 t = *reinterpret_cast<const int2*>(p);
    Arc<float> a; a.idx = t.x; a.w = __int_as_float(t.y); return a;
}
__device__ __forceinline__ Arc<double> ld_arc_g(const Arc<double>* p) {
    int4 t = *reinterpret_cast<const int4*>(p);
    Arc<double> a; a.idx = t.x; a.w = __hiloint2double(t.w, t.z); return a;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <typename T> struct V4 { T v[4]; };

// L2-coherent (L1-bypassing) accesses: the state vectors are rewritten by other SMs each frame.
__device__ __forceinline__ V4<float> ld4_cg(const float* p) {
    float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    V4<float> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
__device__ __forceinline__ V4<double> ld4_cg(const double* p) {
    double2 a = __ldcg(reinterpret_cast<const double2*>(p));
    double2 b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    V4<double> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; return r;
}
// streaming (read-once) loads: the α store on the backward sweep, emissions
__device__ __forceinline__ V4<float> ld4_cs(const float* p) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    V4<float> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
__device__ __forceinline__ V4<double> ld4_cs(const double* p) {
    double2 a = __ldcs(reinterpret_cast<const double2*>(p));
    double2 b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
    V4<double> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; return r;
}
// L1-allocating loads for rows that were prefetched into L1 (the α store on the backward sweep:
// written earlier in this launch by other SMs with L1-bypassing stores, never cached before)
__device__ __forceinline__ V4<float> ld4_ca(const float* p) {
    V4<float> r;
    asm volatile("ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ V4<double> ld4_ca(const double* p) {
    V4<double> r;
    asm volatile("ld.global.ca.v2.f64 {%0, %1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p) : "memory");
    asm volatile("ld.global.ca.v2.f64 {%0, %1}, [%2];" : "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p + 2) : "memory");
    return r;
}
#ifdef MK_ABLATE
__device__ int g_no_stores;
#endif
__device__ __forceinline__ void st4_cg(float* p, const V4<float>& x) {
#ifdef MK_ABLATE
    if (g_no_stores) return;
#endif
    __stcg(reinterpret_cast<float4*>(p), make_float4(x.v[0], x.v[1], x.v[2], x.v[3]));
}
__device__ __forceinline__ void st4_cg(double* p, const V4<double>& x) {
    __stcg(reinterpret_cast<double2*>(p), make_double2(x.v[0], x.v[1]));
    __stcg(reinterpret_cast<double2*>(p) + 1, make_double2(x.v[2], x.v[3]));
}

// Stores of the backward ping-pong vectors (b ⊗ e' and its linear copies): rewritten two frames later and gathered in
// between, they should never leave L2 — but the 2.3 GB α stream of the same sweep pushes their dirty lines out
// (2.5 GB of write-backs per call against 0.23 GB of posteriors).  L2::evict_last keeps more of them resident, and
// L2::evict_first on the forward sweep's α-store writes (below) leaves less of that stream in L2 when the backward
// sweep starts: DRAM traffic per call 7.93 -> 6.64 GB, backward write-backs 2.57 -> 1.38 GB, kernel pair -0.5 %
// (tools/ab_variants.py on one box: 7.468 / 7.441 / 7.433 ms for MK_L2_HINTS = 0 / 1 / 2; evict_first on the emission
// and α reads as well — tried as level 3 — brought the reads UP by 0.2 GB and was dropped).  The policy operands are
// the encodings `createpolicy.fractional.L2::evict_{last,first}.b64 p, 1.0` produces (as in CUTLASS' CacheHintSm90).
#ifndef MK_L2_HINTS
#define MK_L2_HINTS 2
#endif
#if MK_L2_HINTS >= 1
__device__ __forceinline__ void st4_keep(float* p, const V4<float>& x) {
    asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(x.v[0]), "f"(x.v[1]),
                 "f"(x.v[2]), "f"(x.v[3]), "l"(0x14F0000000000000ull)  // createpolicy.fractional.L2::evict_last, 1.0
                 : "memory");
}
__device__ __forceinline__ void st4_keep(double* p, const V4<double>& x) { st4_cg(p, x); }
#else
template <typename T> __device__ __forceinline__ void st4_keep(T* p, const V4<T>& x) { st4_cg(p, x); }
#endif
// The α store on the forward sweep: written once, read back milliseconds later — first in line for eviction.
#if MK_L2_HINTS >= 2
__device__ __forceinline__ void st4_stream(float* p, const V4<float>& x) {
    asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(x.v[0]), "f"(x.v[1]),
                 "f"(x.v[2]), "f"(x.v[3]), "l"(0x12F0000000000000ull)  // createpolicy.fractional.L2::evict_first, 1.0
                 : "memory");
}
__device__ __forceinline__ void st4_stream(double* p, const V4<double>& x) { st4_cg(p, x); }
#else
template <typename T> __device__ __forceinline__ void st4_stream(T* p, const V4<T>& x) { st4_cg(p, x); }
#endif

// posterior accumulation into the (B, D, N) output: Log -> add, Tropical -> max
__device__ __forceinline__ void red_add4(float* p, const V4<float>& x) {
#ifdef MK_ABLATE
    if (g_no_stores) return;
#endif
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x.v[0]),
                 "f"(x.v[1]), "f"(x.v[2]), "f"(x.v[3])
                 : "memory");
}
__device__ __forceinline__ void red_add4(double* p, const V4<double>& x) {
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(p + j, x.v[j]);
}
// values are >= 0 so the IEEE bit patterns order like signed integers
__device__ __forceinline__ void red_max1(float* p, float x) {
    atomicMax(reinterpret_cast<int*>(p), __float_as_int(x));
}
__device__ __forceinline__ void red_max1(double* p, double x) {
    atomicMax(reinterpret_cast<long long*>(p), __double_as_longlong(x));
}
template <int SR, typename T> __device__ __forceinline__ void red1(T* p, T x) {
    if (SR == SR_LOG) atomicAdd(p, x);
    else red_max1(p, x);
}
template <int SR, typename T> __device__ __forceinline__ T lin_add(T a, T b) {
    return SR == SR_LOG ? a + b : max_(a, b);
}

// CTA barrier that does not require the warps to arrive converged (barrier.sync without .aligned): used after the
// tile loops, where the lanes beyond a partial utterance tile skip the work of the live lanes.
__device__ __forceinline__ void cta_sync_unaligned() {
    asm volatile("barrier.sync 0;" ::: "memory");
    // ... and the lanes leave it together: the aligned barriers that follow (__syncthreads in the scalar phase and in
    // grid_sync) need converged warps.  (A plain __syncwarp() here is elided by the compiler, which takes the
    // reconvergence point after the divergent block for granted; compute-sanitizer --tool synccheck showed the lanes
    // of a partial tile arriving at the next __syncthreads one by one.)
    asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
}

// ---- grid barrier ------------------------------------------------------------------------------
// Monotonic counter; the kernel is launched cooperatively (all CTAs co-resident).  Writers'
// stores are ordered by bar.sync + fence + the atomic; the waiting thread uses an acquire load
// and every consumer then reads other CTAs' data with L1-bypassing loads only (ld4_cg).
#ifdef MK_ABLATE
#define MK_ABL(p, bit) ((p).ablate & (bit))
#else
#define MK_ABL(p, bit) false
#endif
#ifdef MK_PROFILE_BARRIER
__shared__ long long t_last;  // thread 0: when the CTA left the last grid barrier
__device__ unsigned long long g_prof[148 * 4];
__device__ unsigned long long g_redo;  // exact-fallback events  // per CTA: cycles before arriving, cycles waiting, ...
#endif
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& target) {
#ifdef MK_PROFILE_BARRIER
    long long t0 = clock64();
#endif
    __syncthreads();
#ifdef MK_PROFILE_BARRIER
    long long t1 = clock64();
#endif
    if (threadIdx.x == 0) {
        target += gridDim.x;
        // release: the CTA's writes (ordered before this thread by bar.sync) become visible before the arrival
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
#ifdef MK_PROFILE_BARRIER
    if (threadIdx.x == 0) {
        long long t2 = clock64();
        if (target > gridDim.x) {  // skip the first barrier
            g_prof[blockIdx.x * 4 + 0] += (unsigned long long)(t0 - t_last);  // thread 0: work since last barrier
            // (slot 1: scalar phase, barrier exit -> chunk loop start, added by the kernel)
            g_prof[blockIdx.x * 4 + 2] += (unsigned long long)(t2 - t1);      // CTA waiting for the grid
        }
        t_last = t2;
    }
#endif
}

// ---- per-lane ⊕ over one row's arcs for 4 utterances (exact two-pass form, log2 units) -----------
// One chunk of CNT arcs: gather CNT x 16 B, then fold into the running (m, s) pair (Log) or
// into m (Tropical).
template <typename T, int SR, int CNT>
__device__ __forceinline__ void chunk_fold(const Arc<T>* __restrict__ arcs, const T* vec, int U4,
                                           int uoff, bool first, T (&m)[4], T (&s)[4]) {
    T x[CNT][4];
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        Arc<T> a = ld_arc(arcs + k);
        V4<T> v = ld4_cg(vec + size_t(a.idx) * U4 + uoff);
#pragma unroll
        for (int j = 0; j < 4; ++j) x[k][j] = v.v[j] + a.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        T mc = x[0][j];
#pragma unroll
        for (int k = 1; k < CNT; ++k) mc = max_(mc, x[k][j]);
        if (SR == SR_TROP) {
            m[j] = max_(m[j], mc);
        } else {
            T mn = first ? mc : max_(m[j], mc);
            T ms = (mn == neg_inf<T>()) ? T(0) : mn;
            T acc = first ? T(0) : s[j] * ex2_(m[j] - ms);
#pragma unroll
            for (int k = 0; k < CNT; ++k) acc += ex2_(x[k][j] - ms);
            s[j] = acc;
            m[j] = mn;
        }
    }
}

// vec: [state][U4] payload; uoff: this lane's utterance offset (multiple of 4).
template <typename T, int SR>
__device__ __forceinline__ V4<T> row_reduce(const Arc<T>* __restrict__ arcs, int beg, int end,
                                            const T* vec, int U4, int uoff) {
    T m[4], s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { m[j] = neg_inf<T>(); s[j] = T(0); }
    int a = beg;
    bool first = true;
    for (; a + kChunk <= end; a += kChunk) {
        chunk_fold<T, SR, kChunk>(arcs + a, vec, U4, uoff, first, m, s);
        first = false;
    }
    switch (end - a) {  // warp-uniform
        case 1: chunk_fold<T, SR, 1>(arcs + a, vec, U4, uoff, first, m, s); break;
        case 2: chunk_fold<T, SR, 2>(arcs + a, vec, U4, uoff, first, m, s); break;
        case 3: chunk_fold<T, SR, 3>(arcs + a, vec, U4, uoff, first, m, s); break;
        case 4: chunk_fold<T, SR, 4>(arcs + a, vec, U4, uoff, first, m, s); break;
        case 5: chunk_fold<T, SR, 5>(arcs + a, vec, U4, uoff, first, m, s); break;
        case 6: chunk_fold<T, SR, 6>(arcs + a, vec, U4, uoff, first, m, s); break;
        case 7: chunk_fold<T, SR, 7>(arcs + a, vec, U4, uoff, first, m, s); break;
        default: break;
    }
    V4<T> out;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (SR == SR_TROP) out.v[j] = m[j];
        else out.v[j] = (s[j] > T(0)) ? m[j] + lg2_(s[j]) : neg_inf<T>();
    }
    return out;
}

// The same ⊕ for forward rows of the Log semiring whose in-arcs may name a merged run (index Ŝ + g): the run's
// members are its consecutive rows, each reached with the arc's weight — exactly the un-merged graph.  The
// α store keeps no log2 row for the virtual sources q_g (only their linear copies exist), so the rare exact
// path expands them.  Plain two-pass loops: this code runs for a handful of rows per call.
template <typename T>
__device__ __forceinline__ void row_reduce_runs(const Arc<T>* __restrict__ arcs, int beg, int end, const T* vec, int U4, int uoff,
                                             const int2* __restrict__ runs, int S, T* out4) {
    T m[4], s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { m[j] = neg_inf<T>(); s[j] = T(0); }
    for (int pass = 0; pass < 2; ++pass) {
        for (int a = beg; a < end; ++a) {
            const Arc<T> arc = ld_arc(arcs + a);
            int r0 = arc.idx, r1 = arc.idx + 1;
            if (arc.idx >= S) {
                const int2 run = __ldg(runs + (arc.idx - S));
                r0 = run.x;
                r1 = run.x + run.y;
            }
            for (int r = r0; r < r1; ++r) {
                const V4<T> v = ld4_cg(vec + size_t(r) * U4 + uoff);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const T x = v.v[j] + arc.w;
                    if (pass == 0) m[j] = max_(m[j], x);
                    else if (m[j] != neg_inf<T>()) s[j] += ex2_(x - m[j]);
                }
            }
        }
        if (m[0] == neg_inf<T>() && m[1] == neg_inf<T>() && m[2] == neg_inf<T>() && m[3] == neg_inf<T>()) break;  // a dead row
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) out4[j] = (s[j] > T(0)) ? m[j] + lg2_(s[j]) : neg_inf<T>();
}

// ---- ordered integer keys: float max through integer atomicMax (works for negatives, -Inf) ----
constexpr int kKeyMin = int(0x80000000);
__device__ __forceinline__ int fkey(float x) {
    int b = __float_as_int(x);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
// the per-frame normaliser: the (float-rounded) maximum if it is finite, else 0
template <int SR, typename T> __device__ __forceinline__ T shift_from_key(int k) {
    if (SR != SR_LOG || k == kKeyMin) return T(0);
    float f = fkey_inv(k);
    return (f > -3.0e38f && f < 3.0e38f) ? T(f) : T(0);
}
template <typename T> __device__ __forceinline__ bool all_zero_bar(const V4<T>& e) {  // all four are 0̄
    return e.v[0] == neg_inf<T>() && e.v[1] == neg_inf<T>() && e.v[2] == neg_inf<T>() && e.v[3] == neg_inf<T>();
}

// ---- work plan of one direction ---------------------------------------------------------------
// Built by the host (markov_b200.cu, build_plan).  Items are rows (or segments of long forward
// rows) in row order.  Their arcs are re-laid out item by item into PADDED arrays: every item
// starts on a multiple of four arcs (pads: weight 0 / 0̄, never loaded), so that the streaming loop
// fetches four neighbour offsets and four weights with one 16-byte shared-memory load each.
template <typename T> struct DirPlan {
    const int4* items;           // {row, pdf, slot or -1, run flags (see FwdFin / BwdFin)}
    const int2* item_pa;         // {first padded arc (multiple of 4), number of arcs}
    const int2* item_arcs;       // {beg, end} in the un-padded arc array (exact fallback only)
    const int4* chunks;          // {parc_begin, parc_end, item_begin, item_end}
    const int* cta_chunks;       // [grid + 1]: CTA c pulls chunks [cta_chunks[c], cta_chunks[c+1])
    const int* pidx;             // padded arcs: neighbour state
    const T* pw;                 // padded arcs: Log: linear weight 2^(w - R - H); Tropical: w
    const Arc<T>* arcs;          // un-padded arcs (w - R, kernel units) for the exact fallback
    T R;                         // bound on the ⊕ exponents (0 for Tropical)
    T H;                         // Log: the linear copy of a stored value v is 2^(v + H) (<= 2^headroom)
    const int2* runs;            // forward plans: merged runs {first row, number of rows}; arc index Ŝ + g = run g
    int n_states;                // Ŝ
};

// Single-pass ⊕ of an item's arcs, resolved to log2 Σ 2^(v + w) for 4 utterances.  acc is the
// linear sum Σ lin·W (Log) or the running maximum (Tropical).  A sum that underflowed for an
// utterance whose emission is alive is redone with the exact two-pass row_reduce over the log2
// copies (which also recognises a genuinely dead row).  `live`: bit j set = utterance j needs its value.
template <typename T> __device__ __forceinline__ T tiny_sum() { return sizeof(T) == 4 ? T(1e-30) : T(1e-280); }  // (the bound sits at 2^100 / 2^900, see build_graph)
template <typename T> __device__ __forceinline__ T min4(const V4<T>& x) {
    return fmin(fmin(x.v[0], x.v[1]), fmin(x.v[2], x.v[3]));
}
template <typename T, int SR>
__device__ __forceinline__ void redo_row(const DirPlan<T>* pl, int item, const T* vec, int U4, int uoff, const T* acc4,
                                      unsigned live, T* val4) {
#ifdef MK_PROFILE_BARRIER
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_redo, 1ull);
#endif
    const T tiny = tiny_sum<T>();
    bool need = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) need |= (acc4[j] < tiny) && ((live >> j) & 1u);
    if (!need) return;
    const int2 ar = __ldg(pl->item_arcs + item);
    if (ar.y <= ar.x) return;
    V4<T> r;
    if (SR == SR_LOG && pl->runs) {
        T o4[4];
        row_reduce_runs<T>(pl->arcs, ar.x, ar.y, vec, U4, uoff, pl->runs, pl->n_states, o4);
#pragma unroll
        for (int j = 0; j < 4; ++j) r.v[j] = o4[j];
    } else {
        r = row_reduce<T, SR>(pl->arcs, ar.x, ar.y, vec, U4, uoff);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if ((acc4[j] < tiny) && ((live >> j) & 1u)) val4[j] = r.v[j] + pl->R;
}
template <typename T> __device__ __forceinline__ unsigned live_mask(const V4<T>& e) {
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) m |= (e.v[j] != neg_inf<T>()) ? (1u << j) : 0u;
    return m;
}
// [U4] ints in the CTA's shared-memory scalar block (after s_key, see shared_fb_kernel): exactness flags of the frame
__device__ __forceinline__ int* exact_flags(int U4, size_t tsize) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    return reinterpret_cast<int*>(smem_raw + size_t(U4) * (2 * sizeof(double) + 3 * tsize + sizeof(int)));
}
// The rare part of resolve_sum (inlined: as out-of-line calls, entered by the few live lanes of a partial utterance tile, it left
// compute-sanitizer's synccheck reporting divergent barriers — lanes return from such a call one by one — and cost 0.5 %): which
// utterances really need the exact path?  Not those whose emission is
// 0̄, and not those whose all-zero sum is exact by construction: the per-frame, per-utterance flags in shared memory
// (exact_flags) say whether the gather source of this frame lives on the seed states only — bit5: the frame after α̂
// (forward) / the frame before an utterance's phony frames (backward); bit6: the frames after an utterance's first
// phony frame (forward) — and the same bits of item.w say that this item has no arc from such a state.  In a ragged
// batch the finished utterances would otherwise send the 146 segments of the phony final state's row through the
// serial exact path in every frame (measured: 5x per frame).
template <typename T, int SR>
__device__ __forceinline__ V4<T> resolve_slow(V4<T> acc, V4<T> val, unsigned live, int item_w, const DirPlan<T>* pl, int item,
                                           const T* vec, int U4, int uoff) {
    const int4 xf = *reinterpret_cast<const int4*>(exact_flags(U4, sizeof(T)) + uoff);
    const int xw[4] = {xf.x, xf.y, xf.z, xf.w};
    unsigned need = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (acc.v[j] < tiny_sum<T>() && ((live >> j) & 1u) && !((item_w & xw[j] & 96) && acc.v[j] == T(0))) need |= 1u << j;
    if (need) {
        T a4[4], v4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { a4[j] = acc.v[j]; v4[j] = val.v[j]; }
        redo_row<T, SR>(pl, item, vec, U4, uoff, a4, need, v4);
#pragma unroll
        for (int j = 0; j < 4; ++j) val.v[j] = v4[j];
    }
    return val;
}
template <typename T, int SR>
__device__ __forceinline__ V4<T> resolve_sum(const V4<T>& acc, const V4<T>& e, bool need_all, bool dead, int item_w,
                                             const DirPlan<T>& pl, int item, const T* vec, int U4, int uoff) {
    if (SR == SR_TROP) return acc;
    V4<T> val;
#pragma unroll
    for (int j = 0; j < 4; ++j) val.v[j] = lg2_(acc.v[j]) + pl.R;  // acc == 0 -> -Inf
    if (min4(acc) < tiny_sum<T>() && !dead) {
        val = resolve_slow<T, SR>(acc, val, need_all ? 15u : live_mask(e), item_w, &pl, item, vec, U4, uoff);
    }
    return val;
}

// ---- arc source: the CTA's shared-memory cache (SA) or the global padded arrays ---------------------
// The cache holds, per padded arc, the row offset pre-multiplied (idx * U4/4, in units of one
// lane's 4 utterances) and the weight; per item its two records.

// Float32 weights sit in shared memory as PAIRS {w, w} (8 bytes per arc): the operand layout of the packed
// fma.rn.f32x2 of the Log fold below
template <typename T> struct CacheW { static constexpr int bytes = sizeof(T) == 4 ? 8 : int(sizeof(T)); };
__device__ __forceinline__ void lds_w4(unsigned a, float (&w)[4]) {
    float d0, d1, d2, d3;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w[0]), "=f"(d0), "=f"(w[1]), "=f"(d1) : "r"(a));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w[2]), "=f"(d2), "=f"(w[3]), "=f"(d3) : "r"(a + 16u));
}
__device__ __forceinline__ void lds_w4_pairs(unsigned a, unsigned long long (&w2)[4]) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w2[0]), "=l"(w2[1]) : "r"(a));
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w2[2]), "=l"(w2[3]) : "r"(a + 16u));
}
template <typename T, bool SA> struct ArcSrc {
    const int* gidx; const T* gw;  // global
    unsigned soff, sw;             // shared addresses of (virtual) padded arc 0
    const int4* gitems; const int2* gpa; const int4* gchunks;
    unsigned sitems, spa;          // shared addresses of (virtual) item 0
    unsigned schunks;              // shared address of (virtual) chunk 0
    __device__ __forceinline__ int4 chunk(int c) const {
        if (SA) {
            int4 r;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(schunks + unsigned(c) * 16u));
            return r;
        }
        return __ldg(gchunks + c);
    }
    int U4q;                       // U4 / 4
    int zpdf;                      // backward sweep: item.z carries the pdf index (no slots there)
    // item record with pre-multiplied offsets: {row * U4/4, pdf * U4/4, slot (forward) or pdf (backward), flags}
    __device__ __forceinline__ static int4 cook(int4 r, int U4q, int zpdf) {
        if (zpdf) r.z = r.y;
        r.x *= U4q;
        r.y *= U4q;
        return r;
    }
    __device__ __forceinline__ int4 item(int i) const {
        if (SA) {
            int4 r;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(sitems + unsigned(i) * 16u));
            return r;
        }
        return cook(__ldg(gitems + i), U4q, zpdf);
    }
    __device__ __forceinline__ int2 item_pa(int i) const {
        if (SA) {
            int2 r;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(spa + unsigned(i) * 8u));
            return r;
        }
        return __ldg(gpa + i);
    }
    // row offsets of the quad's four arcs, in units of one lane's 4 utterances
    __device__ __forceinline__ void offsets(int aq, unsigned (&off)[4]) const {
        if (SA) {
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(off[0]), "=r"(off[1]), "=r"(off[2]), "=r"(off[3]) : "r"(soff + unsigned(aq) * 4u));
        } else {
            const int4 ix = __ldg(reinterpret_cast<const int4*>(gidx + aq));
            off[0] = unsigned(ix.x) * U4q; off[1] = unsigned(ix.y) * U4q;
            off[2] = unsigned(ix.z) * U4q; off[3] = unsigned(ix.w) * U4q;
        }
    }
    __device__ __forceinline__ void weights(int aq, T (&w)[4]) const;
    // Float32: the quad's four weights as {w, w} pairs (operands of fma.rn.f32x2)
    __device__ __forceinline__ void weights2(int aq, unsigned long long (&w2)[4]) const {
        if (SA) {
            lds_w4_pairs(sw + unsigned(aq) * 8u, w2);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float w = __ldg(reinterpret_cast<const float*>(gw) + aq + k);
                asm("mov.b64 %0, {%1, %1};" : "=l"(w2[k]) : "f"(w));
            }
        }
    }
};
__device__ __forceinline__ void lds_w4(unsigned a, double (&w)[4]) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(w[0]), "=d"(w[1]) : "r"(a));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(w[2]), "=d"(w[3]) : "r"(a + 16u));
}
template <typename T, bool SA>
__device__ __forceinline__ void ArcSrc<T, SA>::weights(int aq, T (&w)[4]) const {
    if (SA) {
        lds_w4(sw + unsigned(aq) * unsigned(CacheW<T>::bytes), w);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = __ldg(gw + aq + k);
    }
}

// ---- streaming loop ----------------------------------------------------------------------------------
// One warp streams the items of a chunk for 128 utterances (lane l: utterances 4l..4l+3).  Per pass
// it gathers up to 4 * kPassQuads neighbour rows straight into registers (one 16-byte L2 load per
// arc and lane: a warp reads one contiguous 512 B row) and then folds them with the arc weights:
//   Log      acc += lin[neighbour] * W      (linear copies x linear weights: one FFMA per arc and utterance)
//   Tropical acc  = max(acc, v[neighbour] + w)
// Nothing passes through shared memory except the arc records themselves; the latency of a pass is
// hidden by the other warps of the SM, each of which has its own pass in flight.
#ifndef MK_PASS_QUADS
#define MK_PASS_QUADS 2
#endif
template <typename T> struct PassOf { static constexpr int quads = sizeof(T) == 4 ? MK_PASS_QUADS : (MK_PASS_QUADS + 1) / 2; };

// Float32 Log fold with Blackwell's packed FP32 pipe: the lane's four utterances are two f32x2 pairs, an arc costs two
// fma.rn.f32x2 instead of four FFMA (sm_100+; bit-identical: each half is an IEEE fma.rn).  Gathers land as two 64-bit
// registers, the accumulators stay packed until the item is finalised.
#ifndef MK_FFMA2
#define MK_FFMA2 1
#endif
__device__ __forceinline__ void ld2x2_cg(const float* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("ld.global.cg.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p));
}
__device__ __forceinline__ void fma2(unsigned long long& acc, unsigned long long v, unsigned long long w) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(v), "l"(w));
}
template <bool SA, class Fin>
__device__ __forceinline__ void stream_items_f32x2(const ArcSrc<float, SA>& src, const int i0, const int i1,
                                                   const float* vec_lane, Fin& fin) {
    constexpr int kPassQuads = PassOf<float>::quads;
    for (int item = i0; item < i1; ++item) {
        fin.prefetch(src, item);
        if (fin.is_passive()) {
            fin.passive();
            continue;
        }
        const int2 pa = src.item_pa(item);
        unsigned long long acc01 = 0ull, acc23 = 0ull;
        int a = pa.x, rem = pa.y;
        while (rem > 0) {
            unsigned long long v01[kPassQuads * 4], v23[kPassQuads * 4];
#pragma unroll
            for (int q = 0; q < kPassQuads; ++q) {
                const int left = rem - q * 4;
                if (left > 0) {  // warp-uniform
                    unsigned off[4];
                    src.offsets(a + q * 4, off);
                    if (left >= 4) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) ld2x2_cg(vec_lane + size_t(off[k]) * 4, v01[q * 4 + k], v23[q * 4 + k]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k < left) ld2x2_cg(vec_lane + size_t(off[k]) * 4, v01[q * 4 + k], v23[q * 4 + k]);
                            else v01[q * 4 + k] = v23[q * 4 + k] = 0ull;  // (pad weights are 0)
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kPassQuads; ++q) {
                if (rem - q * 4 > 0) {
                    unsigned long long w2[4];
                    src.weights2(a + q * 4, w2);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        fma2(acc01, v01[q * 4 + k], w2[k]);
                        fma2(acc23, v23[q * 4 + k], w2[k]);
                    }
                }
            }
            a += kPassQuads * 4;
            rem -= kPassQuads * 4;
        }
        V4<float> acc;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.v[0]), "=f"(acc.v[1]) : "l"(acc01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.v[2]), "=f"(acc.v[3]) : "l"(acc23));
        fin(item, acc);
    }
}

template <typename T, int SR, bool SA, class Fin>
__device__ __forceinline__ void stream_items(const ArcSrc<T, SA>& src, const int i0, const int i1,
                                             const T* vec_lane, Fin& fin) {
    if constexpr (MK_FFMA2 && sizeof(T) == 4 && SR == SR_LOG) {
#ifndef MK_ABLATE
        stream_items_f32x2<SA>(src, i0, i1, vec_lane, fin);
        return;
#endif
    }
    constexpr int kPassQuads = PassOf<T>::quads;
    for (int item = i0; item < i1; ++item) {
        fin.prefetch(src, item);  // item record, emission (and α) rows: in flight together with the gathers
        if (fin.is_passive()) {   // rows of a merged run that reuse the run's ⊕ (backward)
            fin.passive();
            continue;
        }
        const int2 pa = src.item_pa(item);
        V4<T> acc;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc.v[j] = SR == SR_LOG ? T(0) : neg_inf<T>();
        int a = pa.x, rem = pa.y;
        while (rem > 0) {
            V4<T> v[kPassQuads * 4];
#pragma unroll
            for (int q = 0; q < kPassQuads; ++q) {
                const int left = rem - q * 4;
                if (left > 0) {  // warp-uniform
                    unsigned off[4];
                    src.offsets(a + q * 4, off);
                    if (MK_ABL(fin.p, 1)) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[q * 4 + k].v[j] = T(1);
                    } else if (left >= 4) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) v[q * 4 + k] = ld4_cg(vec_lane + size_t(off[k]) * 4);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k < left) {
                                v[q * 4 + k] = ld4_cg(vec_lane + size_t(off[k]) * 4);
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[q * 4 + k].v[j] = T(0);  // (pad weights are 0 / 0̄)
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kPassQuads; ++q) {
                if (rem - q * 4 > 0) {
                    T w[4];
                    src.weights(a + q * 4, w);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (SR == SR_LOG) acc.v[j] = fma_(v[q * 4 + k].v[j], w[k], acc.v[j]);
                            else acc.v[j] = max_(acc.v[j], v[q * 4 + k].v[j] + w[k]);
                        }
                }
            }
            a += kPassQuads * 4;
            rem -= kPassQuads * 4;
        }
        if (!MK_ABL(fin.p, 2)) fin(item, acc);
    }
}

// ================================================================================================
// Shared-graph kernel
// ================================================================================================
// Long forward rows (in-degree above the split threshold, e.g. the phony final state whose in-arcs
// are all the final weights) are cut into segment items; a segment writes its partial ⊕ to a
// scratch slot and every CTA combines the slots of the long rows at the start of the next frame
// (redundantly, identical bits) so no second grid barrier is needed.
//
// Normalisation (Log semiring; all stored quantities in log2 units):
//   a_n[s] = log2 Σ_i 2^(a_{n-1}[i] + w_is) + (e_n[s] - emax_n) - shift_n,   shift_n = max_s a_{n-1}[s],
//   α_n = (a_n + Ca_n) ln 2,   Ca_n = Σ_{k<=n} (shift_k + emax_k)   (float64),
// and likewise b_n, Cb_n for β.  Hence a_n <= log2(max column ⊕-sum of T̂) and the ⊕ of a row is
// evaluated in ONE pass against the compile-time bound R (the arcs hold w - R, so every exponent
// is <= 0): one ex2 per arc, no running maximum, no rescaling.  If a row's sum underflows although
// the utterance is alive, the row is redone with the exact two-pass row_reduce.  Stored values stay
// O(10) whatever the sequence length, so Float32 keeps ~1e-6 absolute accuracy in the log domain
// where the un-normalised recursion loses ulp(|α|) per ⊕.
template <typename T> struct SharedParams {
    int S;       // Ŝ states incl. phony final (last)
    int Sq;      // rows of the forward vector: Ŝ + merged-run sources q_g
    int Dh;      // D̂ pdfs incl. phony (last)
    int N1;      // N̂ frames incl. phony (last)
    int U4;      // utterances in the group, padded to a multiple of 4
    int ntiles;  // ceil(U4 / 128)
    DirPlan<T> fwd, bwd;                          // T̂ᵀ rows (by destination) / T̂ rows (by source)
    int cache_f, cache_b;                         // padded arcs per CTA held in shared memory (SA kernels)
    int cache_items_f, cache_items_b;             // item records per CTA held in shared memory
    int n_long; const int4* fwd_long;             // {row, pseudo_beg, pseudo_end, pdf}
    const Arc<T>* fwd_long_arcs;                  // pseudo arcs {slot, 1̄} of the long rows
    int n_slots; T* part;                         // [2][n_slots][U4] segment partials
    const T* init_dense;                          // α̂ as a dense vector [S] (kernel units)
    const T* E;      // expanded, transposed emissions (kernel units, NOT normalised): [N1][Dh][U4]
    const T* emax;   // [N1][U4] per-frame emission maxima (kernel units), subtracted together with the shift
    T* alpha;        // [N1][Ŝ][U4] normalised a_n (log2 units / tropical values)
    T* bt;           // [2][S][U4]    b_{n+1} ⊗ e'_{n+1} ping-pong
    T* flin;         // [2][Sq][U4]   the forward gather source (ping-pong): Log linear copies 2^(a_n + H_f), Tropical a_n itself;
                     //               rows Ŝ.. are the merged runs' virtual sources q_g
    T* blin;         // [2][S][U4]    Log: linear copies 2^(b_{n+1} ⊗ e'_{n+1} + H_b), the gather source
    T* beta_out;     // optional [N1][S][U4]  normalised b_n
    int* gkey;       // [2][N1][U4]   per-frame maxima (ordered keys): forward, backward
    double* Coff;    // [2][N1][U4]   Ca_n, Cb_n
    // posterior output, the reference's (B, D, N) b-fastest array
    T* post; int B; int D; int Tn;
    const int* utt_b;  // [U4] global utterance index per group lane (-1 = padding)
    int post_vec4;     // 1: the 4 utterances of every lane are b0..b0+3, 16 B aligned
    int post_ld;       // > 0: `post` is a staging array [Tn][D][post_ld] in LANE order (column = group lane u): any utterance
                       // order keeps the 16-byte reductions; normalize_permuted_kernel moves it to the caller's (B, D, N) array
    T* zsum;           // [N1][B] per-frame normalisers (linear, relative to lz)
    T* lz;             // [B] forward total log-likelihood (natural log)
    double* lz2;       // [U4] the same in kernel units, handed from the forward to the backward launch
    unsigned* barrier;
    // calibration (mk_graph_create): per-CTA cycles between leaving a grid barrier and arriving at the next, summed over
    // the launch — [2][grid], forward then backward — or null
    unsigned long long* cta_cycles;
    int do_fwd, do_bwd, do_post;
    // frame segment of this launch: the forward sweep runs frames [n_lo, n_hi) upwards, the backward sweep the same
    // range downwards.  A sweep cut into several launches (host-buffer pipeline: copies overlap the kernels) carries
    // its running per-utterance scalars through `carry_C` ([U4] float64) and `carry_shift` ([U4]).
    int n_lo, n_hi;
    double* carry_C; T* carry_shift;
    int ablate;        // debug builds (MK_ABLATE): 1 no gathers, 2 no finalise, 4 no chunk work, 8 no emission/α loads, 16 no stores,
                       // 64 no exact path (every all-zero sum taken at face value)
    int bwd_dead_ok;   // the library applied `expand`: co-unreachable rows have β = 0̄ before the last frame
    // Ragged batches (the intent of the reference's PartialVector drafts, src/inference.jl:76-90,112-127): frames an
    // utterance tile needs, [ntiles] device ints in [2, N1] or null (= N1 everywhere).  Past its sequence length an
    // utterance only carries the phony final state along (e_n alive on the phony pdf only, 1̄ self-loop), so a tile
    // whose longest utterance has L frames stops after frame L (0-based) in the forward sweep and starts there — with
    // the reference's B[:,end] = 1̄ — in the backward sweep: exactly the values the full sweep would produce.
    const int* tile_n1;
    const int* seqlens;  // device [B] sequence lengths, or null (= Tn for every utterance); meaningful when bwd_dead_ok
};
template <typename T> __device__ __forceinline__ int tile_limit(const SharedParams<T>& p, int tile) {
    return p.tile_n1 ? __ldg(p.tile_n1 + tile) : p.N1;
}

template <typename T> __device__ __forceinline__ V4<T> ld4_nc(const T* p);
template <> __device__ __forceinline__ V4<float> ld4_nc<float>(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    V4<float> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <> __device__ __forceinline__ V4<double> ld4_nc<double>(const double* p) {
    double2 a = __ldg(reinterpret_cast<const double2*>(p));
    double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    V4<double> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; return r;
}

// linear copy of a stored (log2) row: 2^(v + H); 0̄ -> 0, values below 2^-126 flush to 0 (the rows that
// then sum to less than `tiny` are redone exactly from the log2 copies, see resolve_sum)
template <typename T> __device__ __forceinline__ void st_lin(T* dst, const V4<T>& val, T H) {
    V4<T> l;
#pragma unroll
    for (int j = 0; j < 4; ++j) l.v[j] = ex2_(val.v[j] + H);
    st4_cg(dst, l);
}

// combine the segment partials of the long rows into a_m (every CTA, identical result)
template <typename T, int SR>
__device__ __forceinline__ void fwd_combine(const SharedParams<T>& p, int m, const T* s_shift, int* s_key) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int U4 = p.U4;
    const size_t frame = size_t(p.Sq) * U4;                          // linear copies: Ŝ + runs rows
    const size_t frame_a = size_t(p.S) * U4;  // α store
    const T* Em = p.E + size_t(m) * p.Dh * U4;
    const T* part = p.part + size_t(m & 1) * p.n_slots * U4;
    for (int k = warp; k < p.n_long; k += int(blockDim.x >> 5)) {
        const int4 lr = __ldg(p.fwd_long + k);
        const int r = lr.x;
        for (int tile = 0; tile < p.ntiles; ++tile) {
            const int uoff = tile * kTileUtts + lane * 4;
            if (uoff >= U4 || m >= tile_limit(p, tile)) continue;
            V4<T> e = ld4_nc<T>(Em + size_t(lr.w) * U4 + uoff);
            V4<T> val;
            if (m == 0) {
                T a0 = __ldg(p.init_dense + r);
#pragma unroll
                for (int j = 0; j < 4; ++j) val.v[j] = a0;
            } else if (all_zero_bar(e)) {
#pragma unroll
                for (int j = 0; j < 4; ++j) val.v[j] = neg_inf<T>();
            } else {
                val = row_reduce<T, SR>(p.fwd_long_arcs, lr.y, lr.z, part, U4, uoff);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // (e - emax first: the two are close for the pdfs that matter, the difference is exact)
                val.v[j] = val.v[j] + (e.v[j] - __ldg(p.emax + size_t(m) * U4 + uoff + j)) - s_shift[uoff + j];
                atomicMax(&s_key[uoff + j], fkey(float(val.v[j])));
            }
            st4_cg(p.alpha + size_t(m) * frame_a + size_t(r) * U4 + uoff, val);
            if (SR == SR_LOG) st_lin(p.flin + size_t(m & 1) * frame + size_t(r) * U4 + uoff, val, p.fwd.H);
            else st4_cg(p.flin + size_t(m & 1) * frame + size_t(r) * U4 + uoff, val);
        }
    }
}

// Finalisers.  Item records arrive with pre-multiplied offsets (ArcSrc::item): it.x = row * U4/4,
// it.y = pdf * U4/4 — in units of one lane's four utterances, so that every row address is one
// IMAD.WIDE away from a lane base pointer.
template <typename T, int SR> struct FwdFin {
    const SharedParams<T>& p;
    const T* prev;      // previous frame's log2 vector (exact fallback)
    T* cur_l;           // lane bases (+ uoff): this frame's log2 vector,
    T* lin_l;           //   its linear copies (Log; gathered by the next frame),
    T* part_l;          //   the segment partials,
    const T* En_l;      //   this frame's emissions
    int uoff;
    T c[4];             // -shift_n of the lane's utterances
    T cm[4];            // -emax_n: e'_n = e_n - emax_n (formed first: exact for the pdfs near the maximum)
    T mx[4];            // running maxima of this tile's a values, flushed once per tile
    int4 it;            // the item being streamed and its emissions, requested when the item starts
    V4<T> e;
    V4<T> qacc;         // merged run: Σ of the members' linear copies (Log) / their maximum (Tropical)
    __device__ __forceinline__ FwdFin(const SharedParams<T>& p_, const T* prev_, T* cur, T* lin, T* part, const T* En,
                                      int uoff_, const T* s_shift, const T* emax_n)
        : p(p_), prev(prev_), cur_l(cur + uoff_), lin_l(lin + uoff_), part_l(part + uoff_), En_l(En + uoff_), uoff(uoff_) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            c[j] = -s_shift[uoff_ + j]; cm[j] = -__ldg(emax_n + uoff_ + j);
            mx[j] = neg_inf<T>(); qacc.v[j] = T(0);
        }
    }
    template <class Src> __device__ __forceinline__ void prefetch(const Src& src, int item) {
        it = src.item(item);
        if (MK_ABL(p, 8)) { for (int j = 0; j < 4; ++j) e.v[j] = T(-1); return; }
        e = ld4_nc<T>(En_l + size_t(unsigned(it.y)) * 4);
    }
    __device__ __forceinline__ bool is_passive() const { return false; }
    __device__ __forceinline__ void passive() {}
    // Row merging: item.w bit0 = member of a run, bit1 = first, bit2 = last, bits 8.. = run index g.  The
    // run's virtual source q_g = ⊕_members a (row Ŝ + g of this frame's vector) feeds the members' common
    // successors in the next frame.  lin: the member's linear copy (Log) / its value (Tropical).
    __device__ __forceinline__ void emit_q(const V4<T>& lin) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (it.w & 2) qacc.v[j] = lin.v[j];
            else qacc.v[j] = SR == SR_LOG ? qacc.v[j] + lin.v[j] : max_(qacc.v[j], lin.v[j]);
        }
        if (it.w & 4) {
            const unsigned qoff = unsigned(p.S + (it.w >> 8)) * unsigned(p.U4 >> 2);
            st4_cg(lin_l + size_t(qoff) * 4, qacc);  // (no row for q_g in the α store: the exact fallback expands runs)
        }
    }
    // val: the row's normalised a_n (log2 / tropical)
    __device__ __forceinline__ void store(V4<T>& val) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mx[j] = max_(mx[j], val.v[j]);
        st4_stream(cur_l + size_t(unsigned(it.x)) * 4, val);  // the α store: Ŝ rows per frame, read back much later
        // the gather source of the next frame: the linear copy (Log) / the value itself (Tropical), in the L2-resident
        // ping-pong that also holds the merged runs' virtual sources
        V4<T> lin;
#pragma unroll
        for (int j = 0; j < 4; ++j) lin.v[j] = SR == SR_LOG ? ex2_(val.v[j] + p.fwd.H) : val.v[j];
        // (rows of a merged run are only ever gathered through the run's virtual source)
        if (!(it.w & 1)) st4_cg(lin_l + size_t(unsigned(it.x)) * 4, lin);
        else emit_q(lin);
    }
    __device__ __forceinline__ void operator()(int item, const V4<T>& acc) {
        // item.w bit3: no initial state reaches this row — α = 0̄ in every frame; bit4: the row has no arcs
        V4<T> val = resolve_sum<T, SR>(acc, e, false, (it.w & 24) || MK_ABL(p, 64), it.w, p.fwd, item, prev, p.U4,
                                       uoff);  // T̂ᵀ A[:,n-1] (:70)
        if (it.z >= 0) {  // segment of a long row: partial ⊕ only
            st4_cg(part_l + size_t(it.z) * p.U4, val);
            return;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) val.v[j] += (e.v[j] + cm[j]) + c[j];  // ⊗ e_n (:71), normalised
        store(val);
    }
};

template <typename T, int SR> struct BwdFin {
    const SharedParams<T>& p;
    const T* bt_next;   // b_{n+1} ⊗ e'_{n+1} (log2; exact fallback)
    T* bt_l;            // lane bases (+ uoff): b_n ⊗ e'_n,
    T* lin_l;           //   its linear copies (Log; gathered by the next frame of the sweep),
    const T* En_l;      //   this frame's emissions,
    const T* An_l;      //   this frame's a_n,
    T* beta_l;          //   optional β output,
    T* post_l;          //   posterior rows of this frame (pdf 0)
    int n, uoff;
    bool post_on;       // this frame and these utterances have posterior rows
    T c[4], g[4];       // -shift_n and Ca_n + Cb_n - log Z of the lane's utterances
    T cm[4];            // -emax_n: b_n ⊗ e'_n = b_n + e_n - emax_n
    T mx[4], zs[4];     // running maxima of b ⊗ e and posterior mass, flushed once per chunk
    int4 it;            // the item being streamed (it.z = pdf), its emissions and α, requested when the item starts
    V4<T> e, a;
    V4<T> last;         // b_n of the last item that owned arcs: rows of a merged run share it (item.w bit0)
    __device__ __forceinline__ BwdFin(const SharedParams<T>& p_, const T* bt_next_, T* bt_cur, T* lin, const T* En,
                                      const T* An, int n_, int uoff_, const T* s_shift, const T* s_g)
        : p(p_), bt_next(bt_next_), bt_l(bt_cur + uoff_), lin_l(lin + uoff_), En_l(En + uoff_), An_l(An + uoff_), n(n_),
          uoff(uoff_) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            c[j] = -s_shift[uoff_ + j]; g[j] = s_g[uoff_ + j];
            cm[j] = -__ldg(p.emax + size_t(n_) * p.U4 + uoff_ + j);
            mx[j] = neg_inf<T>(); zs[j] = T(0); last.v[j] = neg_inf<T>();
        }
        beta_l = p.beta_out ? p.beta_out + size_t(n_) * p.S * p.U4 + uoff_ : nullptr;
        post_on = p.do_post && n_ < p.Tn;
        post_l = post_on ? p.post + size_t(n_) * p.D * (p.post_ld > 0 ? p.post_ld : p.B) : nullptr;
    }
    // items without arcs: rows of a merged run reuse the ⊕ just resolved (item.w bit0)
    __device__ __forceinline__ bool is_passive() const { return it.w & 1; }
    __device__ __forceinline__ void passive() { finish(last); }
    template <class Src> __device__ __forceinline__ void prefetch(const Src& src, int item) {
        it = src.item(item);
        if (MK_ABL(p, 8)) { for (int j = 0; j < 4; ++j) { e.v[j] = T(-1); a.v[j] = T(-1); } return; }
        e = ld4_nc<T>(En_l + size_t(unsigned(it.y)) * 4);
        if (p.do_post) a = ld4_cs(An_l + size_t(unsigned(it.x)) * 4);
    }
    __device__ __forceinline__ void finish(V4<T> beta) {
        if (beta_l) st4_cg(beta_l + size_t(unsigned(it.x)) * 4, beta);
        if (p.do_post) {
            // γ = α ⊗ β ⊘ Z, exp, per-pdf ⊕  (:154-160)
            V4<T> pg;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const T x = a.v[j] + beta.v[j] + g[j];
                pg.v[j] = SR == SR_LOG ? ex2_(x) : exp_(x);
                zs[j] = lin_add<SR>(zs[j], pg.v[j]);
            }
            const int pdf = it.z;
            if (post_on && pdf < p.D) {
                T* dst = post_l + size_t(pdf) * (p.post_ld > 0 ? p.post_ld : p.B);
                if (p.post_ld > 0 && SR == SR_LOG) {
                    red_add4(dst + uoff, pg);
                } else if (p.post_vec4 && SR == SR_LOG) {
                    red_add4(dst + p.utt_b[uoff], pg);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int b = p.utt_b[uoff + j];
                        if (b >= 0 && pg.v[j] > T(0)) red1<SR>(dst + b, pg.v[j]);
                    }
                }
            }
        }
        if (n > 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                beta.v[j] += e.v[j] + cm[j];
                mx[j] = max_(mx[j], beta.v[j]);
            }
            st4_keep(bt_l + size_t(unsigned(it.x)) * 4, beta);
            if (SR == SR_LOG) {
                V4<T> lin;
#pragma unroll
                for (int j = 0; j < 4; ++j) lin.v[j] = ex2_(beta.v[j] + p.bwd.H);
                st4_keep(lin_l + size_t(unsigned(it.x)) * 4, lin);
            }
        }
    }
    __device__ __forceinline__ void operator()(int item, const V4<T>& acc) {
        // an explicit β output needs β_n even where e_n = 0̄ kills α_n and b_n ⊗ e_n
        // item.w bit3: the phony final state is unreachable from this row — β = 0̄ (under `expand` emissions)
        // (the owner of a merged run resolves the ⊕ for the rows tied to it as well: their pdfs, hence their emissions,
        // differ from the owner's — its own 0̄ emission must not leave an underflowed sum unresolved for them)
        V4<T> beta = resolve_sum<T, SR>(acc, e, p.beta_out != nullptr || (it.w & 2), ((it.w & 8) && p.bwd_dead_ok) || (it.w & 16) || MK_ABL(p, 64),
                                        it.w, p.bwd, item, bt_next, p.U4, uoff);  // (:106-107)
#pragma unroll
        for (int j = 0; j < 4; ++j) beta.v[j] += c[j];
        last = beta;
        finish(beta);
    }
};

// fill this CTA's shared-memory arc cache for one direction; returns the arc source
template <typename T, bool SA>
__device__ __forceinline__ ArcSrc<T, SA> make_arc_src(const DirPlan<T>& pl, int cap, int cap_items, unsigned char* smem,
                                                      int U4q, int zpdf) {
    ArcSrc<T, SA> src;
    src.gidx = pl.pidx; src.gw = pl.pw; src.U4q = U4q; src.zpdf = zpdf;
    src.gitems = pl.items; src.gpa = pl.item_pa; src.gchunks = pl.chunks;
    src.soff = src.sw = src.sitems = src.spa = src.schunks = 0;
    if (SA && cap + cap_items > 0) {  // (both 0: the other sweep's plan, unused in this launch)
        // the CTA's chunks cover one contiguous range of padded arcs and of items
        int a0 = 0x7fffffff, a1 = 0, i0 = 0x7fffffff, i1 = 0;
        for (int c = pl.cta_chunks[blockIdx.x]; c < pl.cta_chunks[blockIdx.x + 1]; ++c) {
            const int4 ch = pl.chunks[c];
            a0 = min(a0, ch.x);
            a1 = max(a1, ch.y);
            i0 = min(i0, ch.z);
            i1 = max(i1, ch.w);
        }
        if (a0 > a1) a0 = a1 = 0;
        if (i0 > i1) i0 = i1 = 0;
        unsigned* s_off = reinterpret_cast<unsigned*>(smem);
        unsigned char* s_w = smem + size_t(cap) * 4;
        int4* s_items = reinterpret_cast<int4*>(smem + size_t(cap) * (4 + CacheW<T>::bytes));
        int2* s_pa = reinterpret_cast<int2*>(smem + size_t(cap) * (4 + CacheW<T>::bytes) + size_t(cap_items) * 16);
        int4* s_chunks = reinterpret_cast<int4*>(smem + size_t(cap) * (4 + CacheW<T>::bytes) + ((size_t(cap_items) * 24 + 15) & ~size_t(15)));
        const int c0 = pl.cta_chunks[blockIdx.x], c1 = pl.cta_chunks[blockIdx.x + 1];
        for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) s_chunks[c - c0] = pl.chunks[c];
        src.schunks = unsigned(__cvta_generic_to_shared(s_chunks)) - unsigned(c0) * 16u;
        for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
            s_items[i - i0] = ArcSrc<T, SA>::cook(pl.items[i], U4q, zpdf);
            s_pa[i - i0] = pl.item_pa[i];
        }
        for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x) {
            s_off[a - a0] = unsigned(pl.pidx[a]) * unsigned(U4q);
            if (sizeof(T) == 4) {
                const T w = pl.pw[a];
                reinterpret_cast<T*>(s_w)[2 * (a - a0)] = w;
                reinterpret_cast<T*>(s_w)[2 * (a - a0) + 1] = w;
            } else {
                reinterpret_cast<T*>(s_w)[a - a0] = pl.pw[a];
            }
        }
        src.sitems = unsigned(__cvta_generic_to_shared(s_items)) - unsigned(i0) * 16u;
        src.spa = unsigned(__cvta_generic_to_shared(s_pa)) - unsigned(i0) * 8u;
        src.soff = unsigned(__cvta_generic_to_shared(s_off)) - unsigned(a0) * 4u;
        src.sw = unsigned(__cvta_generic_to_shared(s_w)) - unsigned(a0) * unsigned(CacheW<T>::bytes);
    }
    return src;
}
// padded arcs (offset, weight; cap is a multiple of 4), then the CTA's item records
__host__ __device__ inline size_t arc_cache_bytes(int cap, int items, int chunks, size_t tsize) {
    return size_t(cap) * (4 + (tsize == 4 ? 8 : tsize)) + ((size_t(items) * 24 + 15) & ~size_t(15)) + size_t(chunks) * 16;
}
__host__ __device__ inline size_t shared_scalars_bytes(int U4, size_t tsize) {
    return (size_t(U4) * (2 * sizeof(double) + 3 * tsize + 3 * sizeof(int)) + size_t((U4 + 127) / 128) * sizeof(int) + 16 + 15) &
           ~size_t(15);
}

// PHASE 0: forward sweep (αrecursion), leaves log Z in p.lz2.  PHASE 1: backward sweep (βrecursion + γ).
// Two launches: each sweep gets its own register allocation.
template <typename T, int SR, bool SA, int PHASE>
__global__ void __launch_bounds__(kSharedThreads, 1) shared_fb_kernel(const __grid_constant__ SharedParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = p.S, U4 = p.U4;
    double* s_C = reinterpret_cast<double*>(smem_raw);   // [U4] running Ca (forward) / Cb (backward)
    double* s_lz = s_C + U4;                              // [U4] log Z
    T* s_shift = reinterpret_cast<T*>(s_lz + U4);         // [U4] shift of the current frame
    T* s_g = s_shift + U4;                                // [U4] Ca_n + Cb_n - log Z
    T* s_z = s_g + U4;                                    // [U4] per-frame posterior mass
    int* s_key = reinterpret_cast<int*>(s_z + U4);        // [U4] running maxima
    int* s_exact = s_key + U4;                            // [U4] exactness flags of the frame (resolve_sum, exact_flags())
    int* s_next = s_exact + U4;                           // [ntiles] dynamic chunk counters, one per utterance tile
    int* s_len = s_next + (U4 + kTileUtts - 1) / kTileUtts;  // [U4] sequence lengths: read once (the per-frame scalar phase is
                                                          // serial, and utt_b -> seqlens were two dependent L2 round trips in it:
                                                          // the grid barrier's acquire empties L1 every frame)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t frame = size_t(S) * U4;      // β-side vectors: Ŝ rows
    const size_t frame_q = size_t(p.Sq) * U4;  // forward vectors with the merged-run rows: Ŝ + runs
    const size_t frame_a = frame;              // α store: the states only (the gather source is the flin ping-pong)
    unsigned bar_target = 0;

    // This CTA's arcs stay in shared memory for the whole launch: the per-frame fence of the grid
    // barrier invalidates L1, shared memory survives.
    unsigned char* cache = smem_raw + shared_scalars_bytes(U4, sizeof(T));
    const ArcSrc<T, SA> fwd_src = make_arc_src<T, SA>(p.fwd, PHASE == 0 ? p.cache_f : 0, PHASE == 0 ? p.cache_items_f : 0,
                                                      cache, U4 >> 2, 0);
    const ArcSrc<T, SA> bwd_src = make_arc_src<T, SA>(p.bwd, PHASE == 1 ? p.cache_b : 0, PHASE == 1 ? p.cache_items_b : 0,
                                                      cache, U4 >> 2, 1);

    const bool first_segment = PHASE == 0 ? p.n_lo == 0 : p.n_hi == p.N1;
    if (first_segment) {   // this sweep's per-frame maxima
        int* keys = p.gkey + size_t(PHASE) * p.N1 * U4;
        for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < size_t(p.N1) * U4;
             i += size_t(gridDim.x) * blockDim.x)
            keys[i] = kKeyMin;
    }
    for (int u = threadIdx.x; u < U4; u += blockDim.x) {
        s_C[u] = first_segment ? 0.0 : p.carry_C[u];
        s_shift[u] = (first_segment || PHASE == 1) ? T(0) : p.carry_shift[u];
        s_lz[u] = 0.0; s_g[u] = T(0); s_z[u] = T(0); s_key[u] = kKeyMin;
        const int b = p.utt_b[u];
        s_len[u] = (b >= 0 && p.seqlens) ? __ldg(p.seqlens + b) : p.Tn;
    }
    grid_sync(p.barrier, bar_target);
    long long t_mark = 0, t_work = 0;  // (thread 0, calibration launches only)
    if (p.cta_cycles && threadIdx.x == 0) t_mark = clock64();

    // ---------------------------------------------------------------- forward (αrecursion)
    if (PHASE == 0) {
        const int c0 = p.fwd.cta_chunks[blockIdx.x], c1 = p.fwd.cta_chunks[blockIdx.x + 1];
        for (int n = p.n_lo; n < p.n_hi; ++n) {
            if (n >= 1 && p.n_long) {
                fwd_combine<T, SR>(p, n - 1, s_shift, s_key);
                __syncthreads();
            }
            for (int u = threadIdx.x; u < U4; u += blockDim.x) {
                if (n >= tile_limit(p, u / kTileUtts)) { s_key[u] = kKeyMin; continue; }  // Ca, shift stay at the tile's last frame
                {   // is the gather source α_{n-1} confined to the seed states?  (only meaningful under `expand`)
                    const int L = s_len[u];
                    s_exact[u] = (n == 1 ? 32 : 0) | ((p.bwd_dead_ok && n - 1 >= L) ? 64 : 0);
                }
                int key = kKeyMin;
                if (n >= 1) key = __ldcg(p.gkey + size_t(n - 1) * U4 + u);
                const T em = __ldg(p.emax + size_t(n) * U4 + u);  // (requested together with the key: one L2 round trip)
                T sh = T(0);
                if (n >= 1) sh = shift_from_key<SR, T>(max(key, s_key[u]));
                s_shift[u] = sh;
                s_C[u] += double(sh) + double(em);
                s_key[u] = kKeyMin;
                if (blockIdx.x == 0) p.Coff[size_t(n) * U4 + u] = s_C[u];
            }
            for (int t = threadIdx.x; t < p.ntiles; t += blockDim.x) s_next[t] = c0;
            __syncthreads();
            // Utterance tiles in turn; inside a tile the warps pull the CTA's chunks dynamically, largest first.
            // The finaliser (lane pointers, per-utterance scalars, running maxima) is set up once per tile.
#ifdef MK_PROFILE_BARRIER
            const long long t_busy0 = clock64();
            if (threadIdx.x == 0 && bar_target > gridDim.x) g_prof[blockIdx.x * 4 + 1] += (unsigned long long)(t_busy0 - t_last);
#endif
            for (int tile = 0; tile < p.ntiles; ++tile) {
                if (n >= tile_limit(p, tile)) continue;  // ragged batch: this tile's utterances are all finished
                const bool live = tile * kTileUtts + lane * 4 < U4;  // (lanes beyond the batch stay converged for the pulls)
                const int uoff = live ? tile * kTileUtts + lane * 4 : 0;
                FwdFin<T, SR> fin(p, p.alpha + size_t(n > 0 ? n - 1 : 0) * frame_a, p.alpha + size_t(n) * frame_a,
                                  p.flin + size_t(n & 1) * frame_q, p.part + size_t(n & 1) * p.n_slots * U4,
                                  p.E + size_t(n) * p.Dh * U4, uoff, s_shift, p.emax + size_t(n) * U4);
                // gather source: the previous frame's linear copies (Log) / values (Tropical), merged-run sources included
                const T* gsrc = p.flin + size_t((n - 1) & 1) * frame_q + uoff;
                for (;;) {
                    int wk = 0;
                    if (lane == 0) wk = atomicAdd(s_next + tile, 1);
                    wk = __shfl_sync(0xffffffffu, wk, 0);
                    if (wk >= c1) break;
                    if (MK_ABL(p, 4)) continue;
                    const int4 ch = fwd_src.chunk(wk);
                    if (live) {
                        if (n == 0) {
                            for (int i = ch.z; i < ch.w; ++i) {
                                fin.prefetch(fwd_src, i);
                                if (fin.it.z >= 0) continue;
                                const T a0 = __ldg(p.init_dense + unsigned(fin.it.x) / unsigned(U4 >> 2));  // A[:,1] = α̂ ⊗ e₁  (:68)
                                V4<T> val;
#pragma unroll
                                for (int j = 0; j < 4; ++j) val.v[j] = a0 + (fin.e.v[j] + fin.cm[j]);
                                fin.store(val);
                            }
                        } else {
                            stream_items<T, SR, SA>(fwd_src, ch.z, ch.w, gsrc, fin);
                        }
                    }
                    __syncwarp();
                }
                if (live) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) atomicMax(&s_key[uoff + j], fkey(float(fin.mx[j])));
                }
                __syncwarp();
            }
#ifdef MK_PROFILE_BARRIER
            if (lane == 0) atomicAdd(&g_prof[blockIdx.x * 4 + 3], (unsigned long long)(clock64() - t_busy0));
#endif
            cta_sync_unaligned();
            for (int u = threadIdx.x; u < U4; u += blockDim.x) {
                int k = s_key[u];
                if (k != kKeyMin) atomicMax(p.gkey + size_t(n) * U4 + u, k);
                s_key[u] = kKeyMin;
            }
            if (p.cta_cycles && threadIdx.x == 0) t_work += clock64() - t_mark;
            grid_sync(p.barrier, bar_target);
            if (p.cta_cycles && threadIdx.x == 0) t_mark = clock64();
        }
        if (p.cta_cycles && threadIdx.x == 0) p.cta_cycles[blockIdx.x] = (unsigned long long)t_work;
        if (p.n_hi < p.N1) {  // the sweep continues in the next launch
            if (blockIdx.x == 0)
                for (int u = threadIdx.x; u < U4; u += blockDim.x) { p.carry_C[u] = s_C[u]; p.carry_shift[u] = s_shift[u]; }
            return;
        }
        if (p.n_long) {
            fwd_combine<T, SR>(p, p.N1 - 1, s_shift, s_key);
            __syncthreads();
        }
        // log Z = α_{N̂}[phony final] = a + Ca   (kernel units inside, natural log out)
        for (int u = threadIdx.x; u < U4; u += blockDim.x) {
            const T* last = p.alpha + size_t(tile_limit(p, u / kTileUtts) - 1) * frame_a + size_t(S - 1) * U4;
            T a = __ldcg(last + u);
            double z = (a == neg_inf<T>()) ? double(a) : double(a) + s_C[u];
            int b = p.utt_b[u];
            if (blockIdx.x == 0) {
                p.lz2[u] = z;
                if (b >= 0) p.lz[b] = T(SR == SR_LOG ? z * 0.6931471805599453 : z);
            }
        }
        return;
    }

    // ---------------------------------------------------------------- backward (βrecursion + γ)
    for (int u = threadIdx.x; u < U4; u += blockDim.x) {
        s_key[u] = kKeyMin; s_z[u] = T(0);
        s_lz[u] = p.do_post ? p.lz2[u] : 0.0;
    }
    __syncthreads();
    const int c0 = p.bwd.cta_chunks[blockIdx.x], c1 = p.bwd.cta_chunks[blockIdx.x + 1];
    int* gkey_b = p.gkey + size_t(p.N1) * U4;
    double* Cb = p.Coff + size_t(p.N1) * U4;
    for (int n = p.n_hi - 1; n >= p.n_lo; --n) {
        for (int u = threadIdx.x; u < U4; u += blockDim.x) {
            const int lim = tile_limit(p, u / kTileUtts);
            if (n >= lim) continue;  // this tile's backward sweep starts at frame lim - 1
            {   // is the gather source b_{n+1} ⊗ e_{n+1} confined to the phony final state?
                const int L = s_len[u];
                s_exact[u] = (p.bwd_dead_ok && n + 1 >= L) ? 32 : 0;
            }
            // (the three global loads of the phase are requested together: one L2 round trip, not three)
            int key = kKeyMin;
            T em = T(0);
            double ca = 0.0;
            if (n < lim - 1) {
                key = __ldcg(gkey_b + size_t(n + 1) * U4 + u);
                em = __ldg(p.emax + size_t(n + 1) * U4 + u);
            }
            if (p.do_post) ca = __ldcg(p.Coff + size_t(n) * U4 + u);
            T sh = T(0);
            if (n < lim - 1) {
                sh = shift_from_key<SR, T>(key);
                s_C[u] += double(sh) + double(em);
            }
            s_shift[u] = sh;
            if (blockIdx.x == 0 && p.beta_out) Cb[size_t(n) * U4 + u] = s_C[u];
            if (p.do_post) {
                double lz = s_lz[u];
                s_g[u] = (lz == double(neg_inf<T>())) ? T(0) : T(ca + s_C[u] - lz);
            }
        }
        for (int t = threadIdx.x; t < p.ntiles; t += blockDim.x) s_next[t] = c0;
        __syncthreads();
#ifdef MK_PROFILE_BARRIER
        const long long t_busy0 = clock64();
        if (threadIdx.x == 0 && bar_target > gridDim.x) g_prof[blockIdx.x * 4 + 1] += (unsigned long long)(t_busy0 - t_last);
#endif
        for (int tile = 0; tile < p.ntiles; ++tile) {  // (as in the forward sweep)
            const int lim = tile_limit(p, tile);
            if (n >= lim) continue;
            const bool live = tile * kTileUtts + lane * 4 < U4;
            const int uoff = live ? tile * kTileUtts + lane * 4 : 0;
            BwdFin<T, SR> fin(p, p.bt + size_t((n + 1) & 1) * frame, p.bt + size_t(n & 1) * frame,
                              p.blin + size_t(n & 1) * frame, p.E + size_t(n) * p.Dh * U4,
                              p.alpha + size_t(n) * frame_a, n, uoff, s_shift, s_g);
            const T* gsrc = (SR == SR_LOG ? p.blin + size_t((n + 1) & 1) * frame : fin.bt_next) + uoff;
            for (;;) {
                int wk = 0;
                if (lane == 0) wk = atomicAdd(s_next + tile, 1);
                wk = __shfl_sync(0xffffffffu, wk, 0);
                if (wk >= c1) break;
                if (MK_ABL(p, 4)) continue;
                const int4 ch = bwd_src.chunk(wk);
                if (live) {
                    if (n == lim - 1) {
                        for (int i = ch.z; i < ch.w; ++i) {
                            fin.prefetch(bwd_src, i);
                            V4<T> beta;
#pragma unroll
                            for (int j = 0; j < 4; ++j) beta.v[j] = T(0);  // B[:,end] = 1̄  (:104)
                            fin.finish(beta);
                        }
                    } else {
                        stream_items<T, SR, SA>(bwd_src, ch.z, ch.w, gsrc, fin);
                    }
                }
                __syncwarp();
            }
            if (live) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n > 0) atomicMax(&s_key[uoff + j], fkey(float(fin.mx[j])));
                    if (p.do_post && fin.zs[j] > T(0)) {
                        if (SR == SR_LOG) atomicAdd(&s_z[uoff + j], fin.zs[j]);
                        else red_max1(&s_z[uoff + j], fin.zs[j]);
                    }
                }
            }
            __syncwarp();
        }
#ifdef MK_PROFILE_BARRIER
        if (lane == 0) atomicAdd(&g_prof[blockIdx.x * 4 + 3], (unsigned long long)(clock64() - t_busy0));
#endif
        cta_sync_unaligned();
        for (int u = threadIdx.x; u < U4; u += blockDim.x) {
            int k = s_key[u];
            if (k != kKeyMin) atomicMax(gkey_b + size_t(n) * U4 + u, k);
            s_key[u] = kKeyMin;
            if (p.do_post) {
                int b = p.utt_b[u];
                T v = s_z[u];
                if (b >= 0 && v > T(0)) red1<SR>(p.zsum + size_t(n) * p.B + b, v);
                s_z[u] = T(0);
            }
        }
        if (p.cta_cycles && threadIdx.x == 0) t_work += clock64() - t_mark;
        grid_sync(p.barrier, bar_target);
        if (p.cta_cycles && threadIdx.x == 0) t_mark = clock64();
    }
    if (p.cta_cycles && threadIdx.x == 0) p.cta_cycles[gridDim.x + blockIdx.x] = (unsigned long long)t_work;
    if (p.n_lo > 0 && blockIdx.x == 0)  // the sweep continues in the next launch
        for (int u = threadIdx.x; u < U4; u += blockDim.x) p.carry_C[u] = s_C[u];
}

// ================================================================================================
// Emission transpose + expand  (src/inference.jl:54-60 expand, :146-150 vcat / Ĉ·V̂ — the pdf
// gather itself happens in the recursion kernels through state->pdf)
//   E[n][d][u] for the utterances of one shared-graph group.
// ================================================================================================
template <typename T> struct EmisParams {
    const T* ll; long long sb, sd, sn;
    int D, Tn, expanded;       // as passed by the caller
    int Dh, N1;                // D̂, N̂
    const int* seqlens;        // device [B] or null
    const int* utt_b;          // [U4]
    int U4;
    T* E;
    int n0;                    // first frame of this launch (grid z counts from it)
    T scale;                   // log2(e) for the Log semiring (the kernels work in log2 units), 1 for Tropical
    int* emax_key;             // [N1][U4] ordered keys of the per-(frame, utterance) maxima of E, or null
};

template <typename T>
__device__ __forceinline__ T emission(const T* ll, long long sb, long long sd, long long sn, int D,
                                      int expanded, int L, int b, int d, int n) {
    // `expanded`: bit0 = the caller passed D̂ x N̂ matrices; bit1 = ProbSemiring payloads (probabilities): the recursions run
    // in the log semiring, a loaded value enters as its logarithm (0̄ = 0 -> -Inf, 1̄ = 1 -> 0: `expand` is unchanged)
    if (expanded & 2) {
        if (expanded & 1) return log_(ll[b * sb + d * sd + n * sn]);
        if (d < D) return n < L ? log_(ll[b * sb + d * sd + n * sn]) : neg_inf<T>();
        return n < L ? neg_inf<T>() : T(0);
    }
    if (expanded) return ll[b * sb + d * sd + n * sn];
    if (d < D) return n < L ? ll[b * sb + d * sd + n * sn] : neg_inf<T>();
    return n < L ? neg_inf<T>() : T(0);
}

// grid: (ceil(Dh/32), ceil(U4/32), N1), block (32, 8)
template <typename T> __global__ void expand_transpose_kernel(EmisParams<T> p) {
    __shared__ T tile[32][33];
    const int n = p.n0 + blockIdx.z;
    const int d0 = blockIdx.x * 32, u0 = blockIdx.y * 32;
    {   // the thread's four (utterance, pdf) elements: indices, lengths and values each requested together
        const int d = d0 + threadIdx.x;
        int b4[4], L4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int u = u0 + threadIdx.y + 8 * i;
            b4[i] = (u < p.U4 && d < p.Dh) ? p.utt_b[u] : -1;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) L4[i] = (b4[i] >= 0 && p.seqlens) ? p.seqlens[b4[i]] : p.Tn;
        T v4[4];
        unsigned loaded = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            T cst = neg_inf<T>();
            bool need = b4[i] >= 0;
            if (need && !(p.expanded & 1)) {
                if (d < p.D) need = n < L4[i];
                else { need = false; cst = n < L4[i] ? neg_inf<T>() : T(0); }
            }
            v4[i] = need ? __ldg(p.ll + b4[i] * p.sb + d * p.sd + n * p.sn) : cst;
            loaded |= need ? 1u << i : 0u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            T x = v4[i];
            if ((p.expanded & 2) && ((loaded >> i) & 1u)) x = log_(x);  // ProbSemiring payloads
            tile[threadIdx.y + 8 * i][threadIdx.x] = x * p.scale;
        }
    }
    __syncthreads();
    T m = neg_inf<T>();
    for (int k = threadIdx.y; k < 32; k += 8) {
        int d = d0 + k, u = u0 + threadIdx.x;
        if (d < p.Dh && u < p.U4) {
            const T v = tile[threadIdx.x][k];
            p.E[(size_t(n) * p.Dh + d) * p.U4 + u] = v;
            m = max_(m, v);
        }
    }
    // per-(frame, utterance) maximum over the pdfs: the emission part of the per-frame normaliser
    // (the recursion kernels subtract it together with the shift; E itself stays un-normalised)
    if (p.emax_key && u0 + threadIdx.x < p.U4 && m > neg_inf<T>())
        atomicMax(p.emax_key + size_t(n) * p.U4 + u0 + threadIdx.x, fkey(float(m)));
}

// The same through wider tiles: (64 pdfs | 32 for Float64) x 128 utterances per block of 8 warps.  Reading, a warp takes one
// utterance at a time and its lanes consecutive pdfs (128-byte requests when pdfs are contiguous in the caller's array,
// 16 utterances x 2 requests in flight per lane); writing, a warp takes one pdf row and stores the 128 utterances as four
// 128-byte pieces of the row's 512 bytes; both shared-memory phases are conflict-free (row pitch 129).  The per-(frame,
// utterance) maxima are reduced over the block's 8 warps before the atomics.  Step 7.75 -> 7.68 ms on the 128 x 150 x 3000
// call against the 32 x 32 tiles above (MK_NARROW_TRANSPOSE=1 selects those).
//   grid (ceil(Dh / kWideD), ceil(U4 / 128), frames), block 256
template <typename T> struct WideTile { static constexpr int d = sizeof(T) == 4 ? 64 : 32; };
template <typename T> __global__ void __launch_bounds__(256) expand_transpose_wide_kernel(EmisParams<T> p) {
    constexpr int DT = WideTile<T>::d;
    __shared__ T tile[DT][129];
    __shared__ T s_max[8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = p.n0 + blockIdx.z;
    const int d0 = blockIdx.x * DT, u0 = blockIdx.y * 128;
    // The warp's 16 utterances: index and length first (lanes 0-15 fetch one each: two dependent loads once per warp instead of
    // once per utterance), then the emissions in batches of kTrBatch utterances — every load of a batch is requested before any
    // value is used (the straightforward loop had 2-4 loads in flight per lane: the compiler does not hoist loads over
    // emission()'s branches).
#ifndef MK_TR_BATCH
#define MK_TR_BATCH 8
#endif
    constexpr int kTrBatch = MK_TR_BATCH, kPerLane = DT / 32;
    int my_b = -1, my_L = 0;
    if (lane < 16) {
        const int u = u0 + warp + 8 * lane;
        if (u < p.U4) {
            my_b = p.utt_b[u];
            if (my_b >= 0) my_L = p.seqlens ? p.seqlens[my_b] : p.Tn;
        }
    }
#pragma unroll
    for (int k0 = 0; k0 < 16; k0 += kTrBatch) {
        T v[kTrBatch][kPerLane];
        unsigned loaded = 0;
#pragma unroll
        for (int k = 0; k < kTrBatch; ++k) {
            const int b = __shfl_sync(0xffffffffu, my_b, k0 + k), L = __shfl_sync(0xffffffffu, my_L, k0 + k);
#pragma unroll
            for (int j = 0; j < kPerLane; ++j) {
                const int d = d0 + lane + 32 * j;
                // `expand` (src/inference.jl:54-60) without the load where the value is a constant; see emission()
                T cst = neg_inf<T>();
                bool need = b >= 0 && d < p.Dh;
                if (need && !(p.expanded & 1)) {
                    if (d < p.D) need = n < L;
                    else { need = false; cst = n < L ? neg_inf<T>() : T(0); }
                }
                v[k][j] = need ? __ldg(p.ll + b * p.sb + d * p.sd + n * p.sn) : cst;
                loaded |= need ? 1u << (k * kPerLane + j) : 0u;
            }
        }
#pragma unroll
        for (int k = 0; k < kTrBatch; ++k)
#pragma unroll
            for (int j = 0; j < kPerLane; ++j) {
                T x = v[k][j];
                if ((p.expanded & 2) && ((loaded >> (k * kPerLane + j)) & 1u)) x = log_(x);  // ProbSemiring payloads
                tile[lane + 32 * j][warp + 8 * (k0 + k)] = x * p.scale;
            }
    }
    __syncthreads();
    T m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = neg_inf<T>();
    for (int dl = warp; dl < DT; dl += 8) {
        const int d = d0 + dl;
        if (d >= p.Dh) break;
        T* row = p.E + (size_t(n) * p.Dh + d) * p.U4 + u0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ul = lane + 32 * k;
            if (u0 + ul < p.U4) {
                const T v = tile[dl][ul];
                row[ul] = v;
                m[k] = max_(m[k], v);
            }
        }
    }
    if (!p.emax_key) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) s_max[warp][lane + 32 * k] = m[k];
    __syncthreads();
    if (threadIdx.x < 128 && u0 + threadIdx.x < p.U4) {
        T mm = s_max[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < 8; ++w) mm = max_(mm, s_max[w][threadIdx.x]);
        if (mm > neg_inf<T>()) atomicMax(p.emax_key + size_t(n) * p.U4 + u0 + threadIdx.x, fkey(float(mm)));
    }
}

// emax[n][u] from its key (0 when the whole column is 0̄; keys start as 0x80808080 = memset 0x80)
template <typename T> __global__ void emission_max_decode_kernel(const int* key, T* emax, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float m = fkey_inv(key[i]);
    emax[i] = (m > -3.0e38f && m < 3.0e38f) ? T(m) : T(0);
}

// ================================================================================================
// Small-graph kernel: one CTA per utterance
// ================================================================================================
template <typename T> struct UttDesc {
    const int* in_ptr;  const Arc<T>* in_arcs;
    const int* out_ptr; const Arc<T>* out_arcs;
    const int* pdf;
    const T* init_dense;
    int S;        // Ŝ
    int b;        // global utterance index
    long long ws_off;   // offset of this utterance's [N1][S] block in the α workspace
    long long out_off;  // state offset off_b in the virtual union (for user-layout outputs)
    long long c_off;    // offset of this utterance's [N1] block in the Ca workspace
};

template <typename T> struct SmallParams {
    const UttDesc<T>* utts;
    const T* ll; long long sb, sd, sn;
    int D, Tn, expanded, Dh, N1;
    const int* seqlens;
    // α destination.  alpha_user == 0: normalised a_n into the workspace, element (n, s) at
    // alpha + ws_off + n*S + s.  alpha_user == 1: un-normalised α into the caller's (ΣŜ x N̂)
    // array, element at alpha + out_off + n*alpha_sn + s.
    T* alpha; long long alpha_sn; int alpha_user;
    T* beta_out; long long beta_sn;  // user layout only (un-normalised β)
    double* Ca;                      // [Σ N1] forward offsets
    T* post; int B;
    T* zsum; T* lz;
    int do_fwd, do_bwd, do_post;
    // frame segment of this launch: the forward sweep runs frames [n_lo, n_hi) upwards, the backward sweep the same
    // range downwards.  A sweep cut into several launches (host-buffer pipeline: copies overlap the kernels) carries
    // its running per-utterance scalars through `carry_C` ([U4] float64) and `carry_shift` ([U4]).
    int n_lo, n_hi;
    double* carry_C; T* carry_shift;
};

template <typename T, int SR>
__device__ __forceinline__ T small_row(const Arc<T>* __restrict__ arcs, int beg, int end, const T* vec) {
    T m = neg_inf<T>();
    for (int a = beg; a < end; ++a) {
        Arc<T> arc = ld_arc(arcs + a);
        m = max_(m, arc.w + vec[arc.idx]);
    }
    if (SR == SR_TROP || m == neg_inf<T>()) return m;
    T s = T(0);
    for (int a = beg; a < end; ++a) {
        Arc<T> arc = ld_arc(arcs + a);
        s += exp_(arc.w + vec[arc.idx] - m);
    }
    return m + log_(s);
}

// block-wide maximum of per-thread keys through a parity-double-buffered shared array: the
// value written in frame n is read by every thread at the start of frame n+1 (after the frame's
// __syncthreads), so no extra barrier is needed.
__device__ __forceinline__ void publish_key(int* s_keys, int parity, int key) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
    if ((threadIdx.x & 31) == 0) s_keys[parity * 32 + (threadIdx.x >> 5)] = key;
}
__device__ __forceinline__ int collect_key(const int* s_keys, int parity) {
    int k = kKeyMin;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) k = max(k, s_keys[parity * 32 + w]);
    return k;
}

template <typename T, int SR> __global__ void __launch_bounds__(1024, 1) small_fb_kernel(SmallParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const UttDesc<T> u = p.utts[blockIdx.x];
    const int S = u.S;
    T* v0 = reinterpret_cast<T*>(smem_raw);
    T* v1 = v0 + S;
    T* s_red = v1 + S;                               // [2][32] posterior mass per warp (frame parity)
    int* s_keys = reinterpret_cast<int*>(s_red + 64);  // [2][32] maxima per warp (frame parity)
    const int b = u.b;
    const int L = p.seqlens ? p.seqlens[b] : p.Tn;
    T* A = p.alpha + (p.alpha_user ? u.out_off : u.ws_off);
    const long long a_sn = p.alpha_user ? p.alpha_sn : S;
    double* Ca = p.Ca + u.c_off;
    double C = 0.0;  // every thread tracks the same running offset
    // A thread's first state (usually its only one: numerator graphs have a few hundred states) keeps its pdf and arc
    // range in registers, and its emission — one dependent DRAM read per frame, the longest latency of the frame — is
    // requested one frame ahead so that it is in flight while the current frame is computed.
    const int s0 = threadIdx.x;
    const bool has0 = s0 < S;
    const int pdf0 = has0 ? u.pdf[s0] : 0;
    const int ib0 = has0 ? u.in_ptr[s0] : 0, ie0 = has0 ? u.in_ptr[s0 + 1] : 0;
    const int ob0 = has0 ? u.out_ptr[s0] : 0, oe0 = has0 ? u.out_ptr[s0 + 1] : 0;

    if (p.do_fwd) {
        T e_ahead = has0 ? emission<T>(p.ll, p.sb, p.sd, p.sn, p.D, p.expanded, L, b, pdf0, 0) : neg_inf<T>();
        for (int n = 0; n < p.N1; ++n) {
            const T* prev = (n & 1) ? v0 : v1;
            T* cur = (n & 1) ? v1 : v0;
            const T e0 = e_ahead;
            if (has0 && n + 1 < p.N1) e_ahead = emission<T>(p.ll, p.sb, p.sd, p.sn, p.D, p.expanded, L, b, pdf0, n + 1);
            const T sh = n > 0 ? shift_from_key<SR, T>(collect_key(s_keys, (n - 1) & 1)) : T(0);
            C += double(sh);
            if (threadIdx.x == 0) Ca[n] = C;
            int key = kKeyMin;
            for (int s = threadIdx.x; s < S; s += blockDim.x) {
                const bool first = s == s0;
                T e = first ? e0 : emission<T>(p.ll, p.sb, p.sd, p.sn, p.D, p.expanded, L, b, u.pdf[s], n);
                T acc = (n == 0) ? u.init_dense[s]
                        : (e == neg_inf<T>()) ? e
                        : small_row<T, SR>(u.in_arcs, first ? ib0 : u.in_ptr[s], first ? ie0 : u.in_ptr[s + 1], prev);
                acc = acc + e - sh;
                cur[s] = acc;
                key = max(key, fkey(float(acc)));
                A[size_t(n) * a_sn + s] = p.alpha_user ? ((p.expanded & 2) ? T(exp(double(acc) + C)) : T(double(acc) + C)) : acc;
            }
            publish_key(s_keys, n & 1, key);
            __syncthreads();
        }
    }
    // log Z = a_{N̂}[phony final] + Ca   (all threads; also orders the α store for the sweep below)
    double lz = 0.0;
    if (p.do_fwd) {
        const T a = ((p.N1 - 1) & 1 ? v1 : v0)[S - 1];
        lz = (a == neg_inf<T>()) ? double(a) : double(a) + C;
        if (threadIdx.x == 0) p.lz[b] = T(lz);
    }
    if (!p.do_bwd) return;
    __syncthreads();

    T* Bo = p.beta_out ? p.beta_out + u.out_off : nullptr;
    C = 0.0;  // now Cb
    T e_ahead = has0 ? emission<T>(p.ll, p.sb, p.sd, p.sn, p.D, p.expanded, L, b, pdf0, p.N1 - 1) : neg_inf<T>();
    T a_ahead = (has0 && p.do_post) ? A[size_t(p.N1 - 1) * a_sn + s0] : T(0);
    for (int n = p.N1 - 1; n >= 0; --n) {
        const T* nxt = (n & 1) ? v0 : v1;  // b_{n+1} ⊗ e_{n+1}
        T* cur = (n & 1) ? v1 : v0;
        const T e0 = e_ahead, a0 = a_ahead;
        if (has0 && n > 0) {  // the next frame of the sweep: emission and α row in flight during this one
            e_ahead = emission<T>(p.ll, p.sb, p.sd, p.sn, p.D, p.expanded, L, b, pdf0, n - 1);
            if (p.do_post) a_ahead = A[size_t(n - 1) * a_sn + s0];
        }
        const T sh = n < p.N1 - 1 ? shift_from_key<SR, T>(collect_key(s_keys, (n + 1) & 1)) : T(0);
        C += double(sh);
        T g = T(0);
        if (p.do_post && lz != double(neg_inf<T>())) g = T(Ca[n] + C - lz);
        T zs = T(0);
        int key = kKeyMin;
        for (int i = threadIdx.x; i < S; i += blockDim.x) {
            const bool first = i == s0;
            const int pdf = first ? pdf0 : u.pdf[i];
            T e = first ? e0 : emission<T>(p.ll, p.sb, p.sd, p.sn, p.D, p.expanded, L, b, pdf, n);
            const bool dead = !Bo && e == neg_inf<T>();
            T beta = (n == p.N1 - 1) ? T(0)
                     : dead ? neg_inf<T>()
                            : small_row<T, SR>(u.out_arcs, first ? ob0 : u.out_ptr[i], first ? oe0 : u.out_ptr[i + 1], nxt) - sh;
            if (Bo) Bo[size_t(n) * p.beta_sn + i] = (p.expanded & 2) ? T(exp(double(beta) + C)) : T(double(beta) + C);
            if (p.do_post && !dead) {
                T pg = exp_((first ? a0 : A[size_t(n) * a_sn + i]) + beta + g);
                zs = lin_add<SR>(zs, pg);
                if (n < p.Tn && pdf < p.D && pg > T(0)) red1<SR>(p.post + (size_t(n) * p.D + pdf) * p.B + b, pg);
            }
            beta += e;
            cur[i] = beta;
            key = max(key, fkey(float(beta)));
        }
        publish_key(s_keys, n & 1, key);
        if (p.do_post) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) zs = lin_add<SR>(zs, __shfl_xor_sync(0xffffffffu, zs, o));
            if ((threadIdx.x & 31) == 0) s_red[(n & 1) * 32 + (threadIdx.x >> 5)] = zs;
        }
        __syncthreads();
        if (p.do_post && threadIdx.x == 0) {
            T z = T(0);
            for (int w = 0; w < (blockDim.x + 31) / 32; ++w) z = lin_add<SR>(z, s_red[(n & 1) * 32 + w]);
            p.zsum[size_t(n) * p.B + b] = z;
        }
    }
}

// ================================================================================================
// Normalisation: Ẑ ./ sums, ttl = minimum(sums)   (src/inference.jl:157-160)
// ================================================================================================
// post[n][d][b] /= zsum[n][b]  (Ẑ ./ sums, :158) and, when `occ` is given, the pdf occupancy of the data-parallel
// step statistics: occ[d] += Σ_{n,b} post[n][d][b] (float64).  One warp per pdf d and run of kNormFrames frames:
// the lanes sweep the utterance-fastest rows (512 B at B = 128) with 16-byte accesses when B % 4 == 0, keep a partial
// sum over the run, and issue ONE float64 atomic per (d, run) — 20 per pdf at T = 150.
//   grid (ceil(D / warps per block), ceil(Tn / kNormFrames)), block 256
constexpr int kNormFrames = 8;
template <typename T>
__global__ void normalize_post_kernel(T* post, const T* zsum, int B, int D, int Tn, double* occ) {
    const int lane = threadIdx.x & 31, d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= D) return;
    const int n0 = blockIdx.y * kNormFrames, n1 = min(Tn, n0 + kNormFrames);
    T sum = T(0);
    const bool vec = sizeof(T) == 4 && (B & 3) == 0 && (reinterpret_cast<uintptr_t>(post) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(zsum) & 15) == 0;
    for (int n = n0; n < n1; ++n) {
        T* row = post + (size_t(n) * D + d) * B;
        const T* z = zsum + size_t(n) * B;
        if (vec) {
            for (int b = lane * 4; b < B; b += 128) {
                const float4 zz = *reinterpret_cast<const float4*>(z + b);
                float4 v = *reinterpret_cast<float4*>(row + b);
                v.x = zz.x > 0.f ? v.x / zz.x : 0.f; v.y = zz.y > 0.f ? v.y / zz.y : 0.f;
                v.z = zz.z > 0.f ? v.z / zz.z : 0.f; v.w = zz.w > 0.f ? v.w / zz.w : 0.f;
                *reinterpret_cast<float4*>(row + b) = v;
                sum += T((v.x + v.y) + (v.z + v.w));
            }
        } else {
            for (int b = lane; b < B; b += 32) {
                const T zz = z[b];
                const T v = zz > T(0) ? row[b] / zz : T(0);
                row[b] = v;
                sum += v;
            }
        }
    }
    if (!occ) return;
    double s = double(sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && s != 0.0) atomicAdd(occ + d, s);
}
// The same from a LANE-ordered staging array (SharedParams::post_ld): post[n][d][b] = stage[n][d][lane_of[b]] / zsum[n][b].
// Used when the library sorted a ragged group's utterances by length (any permutation of the lanes).
template <typename T>
__global__ void normalize_permuted_kernel(const T* stage, int ld, const int* lane_of, T* post, const T* zsum, int B, int D,
                                          int Tn, double* occ) {
    const int lane = threadIdx.x & 31, d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= D) return;
    const int n0 = blockIdx.y * kNormFrames, n1 = min(Tn, n0 + kNormFrames);
    T sum = T(0);
    for (int n = n0; n < n1; ++n) {
        const T* src = stage + (size_t(n) * D + d) * ld;
        T* row = post + (size_t(n) * D + d) * B;
        const T* z = zsum + size_t(n) * B;
        for (int b = lane; b < B; b += 32) {
            const int u = lane_of[b];
            const T zz = z[b];
            const T v = (u >= 0 && zz > T(0)) ? src[u] / zz : T(0);
            row[b] = v;
            sum += v;
        }
    }
    if (!occ) return;
    double s = double(sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && s != 0.0) atomicAdd(occ + d, s);
}
// logz[b] = lz[b] + log(min_n zsum[n][b])
// (nlimit: frames that were evaluated for utterance b — its tile's limit in a ragged batch — or null = N1)
// stats (optional, float64): stats[0] += Σ_b logz[b], stats[1] += Σ_b frames of b (seqlens[b], or Tn)
template <typename T>
__global__ void total_kernel(const T* zsum, const T* lz, T* logz, int B, int N1, const int* nlimit, double* stats,
                             const int* seqlens, int Tn, int prob) {
    // one warp per utterance: the lanes share the frames (a lone thread walked 151 dependent L2 loads: 36 us)
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const T l = lz[b];
    T out;
    if (l == neg_inf<T>()) {
        out = l;
    } else {
        const int nl = nlimit ? nlimit[b] : N1;
        T mn = zsum[b];  // (frame 0 always exists)
        for (int n = lane; n < nl; n += 32) mn = fmin(mn, zsum[size_t(n) * B + b]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        out = mn > T(0) ? l + T(log(double(mn))) : neg_inf<T>();
    }
    if (lane != 0) return;
    logz[b] = prob ? T(exp(double(out))) : out;  // (ProbSemiring: the total as a probability; the statistics keep the logarithm)
    if (stats) {
        atomicAdd(stats, double(out));
        atomicAdd(stats + 1, double(seqlens ? seqlens[b] : Tn));
    }
}

// ================================================================================================
// LF-MMI gradient (caller pattern examples/test_cuda.jl:118-152: permutedims :120, the two
// pdfposteriors calls :140-143, their difference :152):
//   grad[b][n][d] = scale * (den[n][d][b] - num[n][d][b])        for n < len[b], else 0
// num / den are (B, D, N) b-fastest posterior arrays = [N][D][B]; grad is the network's (B, T, D) array
// with element strides (gsb, gsn, gsd).  One 32 x 32 (d, b) tile per block: transposes through shared
// memory so that both sides are coalesced.   grid (ceil(D/32), ceil(B/32), N), block (32, 8)
// ================================================================================================
template <typename T>
__global__ void lfmmi_grad_kernel(const T* num, const T* den, int B, int D, int N, const int* seqlens, T scale, T* grad,
                                  long long gsb, long long gsn, long long gsd) {
    __shared__ T tile[32][33];
    const int n = blockIdx.z, d0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += 8) {
        const int d = d0 + k, b = b0 + threadIdx.x;
        T v = T(0);
        if (d < D && b < B) {
            const size_t i = (size_t(n) * D + d) * B + b;
            v = scale * (den[i] - num[i]);
        }
        tile[k][threadIdx.x] = v;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += 8) {
        const int b = b0 + k, d = d0 + threadIdx.x;
        if (b < B && d < D) {
            const bool live = !seqlens || n < seqlens[b];
            grad[b * gsb + n * gsn + d * gsd] = live ? tile[threadIdx.x][k] : T(0);
        }
    }
}

// ================================================================================================
// Layout conversion for the αrecursion / βrecursion entry points:
//   src [N1][S][U4] (shared-graph layout, normalised) + C[n][u] -> dst[(off_b + s) + total*n]
//   (reference layout, un-normalised)
// grid (ceil(S/32), ceil(U4/32), N1), block (32, 8)
// ================================================================================================
template <typename T>
__global__ void unpack_states_kernel(const T* src, int S, int S_src /* rows per source frame */, int U4, const int* utt_b,
                                     const long long* utt_off, const double* C /* [N1][U4] */, double unit, T* dst,
                                     long long total, int n0, int prob) {
    __shared__ T tile[32][33];
    const int n = n0 + blockIdx.z, s0 = blockIdx.x * 32, u0 = blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += 8) {
        int s = s0 + k, u = u0 + threadIdx.x;
        if (s < S && u < U4) tile[k][threadIdx.x] = src[(size_t(n) * S_src + s) * U4 + u];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += 8) {
        int u = u0 + k, s = s0 + threadIdx.x;
        if (s < S && u < U4 && utt_b[u] >= 0)
        {
            const double v = (double(tile[threadIdx.x][k]) + C[size_t(n) * U4 + u]) * unit;
            dst[utt_off[u] + s + total * n] = prob ? T(exp(v)) : T(v);
        }
    }
}

// ================================================================================================
// Viterbi back-trace (bestpath).  One warp per utterance walks N̂-1 steps
// back from the phony final state; at each step the lanes scan the in-arcs of the current
// state and pick the first maximum in ascending predecessor order.
//   α(n, s) of utterance k lives at alpha + base[k] + n*sn[k] + s*ss[k].
// ================================================================================================
template <typename T> struct TraceDesc {
    const int* in_ptr; const Arc<T>* in_arcs;
    long long base, sn, ss;
    int S, b;
};
template <typename T>
__global__ void backtrace_kernel(const TraceDesc<T>* descs, int nutts, const T* alpha, int N1, int Tn,
                                 const int* seqlens, int* path, T* score) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= nutts) return;
    const TraceDesc<T> d = descs[k];
    const int L = seqlens ? seqlens[d.b] : Tn;
    int* out = path + size_t(d.b) * Tn;
    for (int t = lane; t < Tn; t += 32) out[t] = 0;
    __syncwarp();
    const T* A = alpha + d.base;
    T sc = A[size_t(N1 - 1) * d.sn + size_t(d.S - 1) * d.ss];
    if (lane == 0) score[d.b] = sc;
    if (sc == neg_inf<T>()) return;
    int cur = d.S - 1;
    for (int n = N1 - 1; n >= 1; --n) {
        const T* prev = A + size_t(n - 1) * d.sn;
        const int beg = d.in_ptr[cur], end = d.in_ptr[cur + 1];
        T best = neg_inf<T>();
        int arg = 0x7fffffff;
        for (int a = beg + lane; a < end; a += 32) {  // ascending within a lane
            Arc<T> arc = ld_arc(d.in_arcs + a);
            T v = arc.w + prev[size_t(arc.idx) * d.ss];
            if (v > best) { best = v; arg = arc.idx; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            T ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        cur = arg;
        if (lane == 0 && n - 1 < L) out[n - 1] = cur + 1;
    }
}


// ================================================================================================
// MUFU (SFU) peak: the SFU half of the roofline (roofline time = max(t_HBM, t_SFU)) is quoted against
// the ex2 rate MEASURED on the box, not a data-sheet number.  Eight independent ex2 chains per thread.
// ================================================================================================
__global__ void sfu_peak_kernel(float* out, int iters) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = -0.001f * float(threadIdx.x + k);
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace mk
