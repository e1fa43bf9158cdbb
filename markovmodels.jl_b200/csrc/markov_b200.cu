// SPDX-License-Identifier: MIT
//
// markov_b200.cu — host side of libmarkov_b200.so: graph compilation (FSM + compile + adapt,
// src/fsm.jl:7-48, src/inference.jl:3-26), the ragged batch descriptor (rawunion / batch,
// src/fsmops.jl:28-36, src/inference.jl:28-36), kernel dispatch, and the C ABI declared in
// include/markov_b200.h.  No CPU compute path exists here: every entry point either runs the
// CUDA kernels of kernels.cuh or fails.
#include "../../include/markov_b200.h"
#include "kernels.cuh"
#include "linalg.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <memory>
#include <atomic>
#include <mutex>
#include <functional>
#include <limits>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace mk;

// ------------------------------------------------------------------------------------------------
// errors, instrumentation
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local int64_t g_launches = 0;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
// the same for the other translation units of the library (prep.cu)
int mk_set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return fail(code, "%s", buf);
}
void mk_note_launches(int n) { g_launches += n; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(MK_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                 \
    } while (0)
#define TRY(call)                   \
    do {                            \
        int rc_ = (call);           \
        if (rc_ != MK_OK) return rc_; \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return MK_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 16 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            return fail(MK_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return MK_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <typename T> static int upload(const std::vector<T>& h, void** d) {
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    CK(cudaMalloc(d, bytes));
    if (!h.empty()) CK(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return MK_OK;
}

// ------------------------------------------------------------------------------------------------
// mk_graph
// ------------------------------------------------------------------------------------------------
// device-side work plan of one direction (kernels.cuh DirPlan)
struct DirDev {
    int4 *items = nullptr, *chunks = nullptr;
    int2 *item_arcs = nullptr, *item_pa = nullptr;
    int *cta_chunks = nullptr, *pidx = nullptr;
    void *pw = nullptr, *arcs = nullptr;  // arcs: un-padded Arc<T> holding w - R in kernel units
    int cache_cap = 0;  // largest per-CTA range of padded arcs (multiple of 4)
    int cache_items = 0;  // largest per-CTA number of items
    int cache_chunks = 0; // largest per-CTA number of chunks
    double R = 0;       // bound on the ⊕ exponents, kernel units
    double H = 0;       // Log: linear copy of a stored value v = 2^(v + H), kernel units
    void release() {
        cudaFree(items); cudaFree(chunks); cudaFree(item_arcs); cudaFree(cta_chunks); cudaFree(pidx);
        cudaFree(pw); cudaFree(arcs); cudaFree(item_pa);
    }
};

struct mk_graph {
    int semiring = 0, dtype = 0, device = 0, n_sms = 0;
    bool prob = false;  // a ProbSemiring graph: stored and run as its LogSemiring image (log.(weights)); inputs / outputs converted
    bool calibrated = false;                 // the plans' CTA split follows measured per-CTA speeds
    unsigned long long* d_cta_cycles = nullptr;  // set while the calibration call runs: [2][n_sms]
    int64_t S = 0, nnz = 0, Dh = 0;
    int max_in_deg = 0, max_out_deg = 0;
    // true weights, both orientations: per-utterance kernel, back-trace
    int *d_in_ptr = nullptr, *d_out_ptr = nullptr, *d_pdf = nullptr;
    void *d_in_arcs = nullptr, *d_out_arcs = nullptr, *d_init_dense = nullptr;
    // shared-graph kernel plans
    DirDev fwd, bwd;
    int4* d_fwd_long = nullptr;
    void *d_fwd_long_arcs = nullptr, *d_init_dense_s = nullptr;
    int n_long = 0, n_slots = 0;
    int n_runs = 0;  // merged runs: the forward vector has Ŝ + n_runs rows
    int2* d_runs = nullptr;  // {first row, number of rows} of every run (exact fallback of the forward sweep)
    size_t bytes = 0;
    ~mk_graph() {
        cudaFree(d_in_ptr); cudaFree(d_out_ptr); cudaFree(d_pdf);
        cudaFree(d_in_arcs); cudaFree(d_out_arcs); cudaFree(d_init_dense);
        fwd.release(); bwd.release();
        cudaFree(d_fwd_long); cudaFree(d_fwd_long_arcs); cudaFree(d_init_dense_s); cudaFree(d_runs);
    }
};

// Work plan of one direction.  Rows with more than kLongRow arcs (forward only — typically the
// phony final state: one in-arc per final state) are cut into segment items that spread over all
// CTAs; each segment writes a partial ⊕ into a scratch slot and every CTA combines the slots at
// the start of the next frame (kernels.cuh, fwd_combine).  Items are cut into per-CTA ranges
// balanced by arcs, and inside a CTA into chunks of ~kChunkArcs arcs that the warps pull
// dynamically, largest first.  The arcs are re-laid out chunk by chunk into padded arrays (see
// kernels.cuh DirPlan): every chunk spans a multiple of four arcs, every item owns >= 1 arc.
constexpr int kItemDead = 8;    // item.w bit3 (both sweeps): statically dead row, see kernels.cuh
constexpr int kItemEmpty = 16;  // item.w bit4 (both sweeps): the row has no arcs
constexpr int kItemNoSeed = 32;  // item.w bit5: forward — no in-arc from an initial state; backward — no arc into the phony final state
constexpr int kItemNoFinalPred = 64;  // item.w bit6 (forward): no in-arc from the phony final state (every row but the phony state's self-loop segment)
constexpr double kItemCost = 12.0;
constexpr int kLongRow = 128;
constexpr int kMinSegment = 64;
constexpr int kMaxSlotsPerRow = 160;
// Target cost (arcs + kItemCost per item) of a dynamically scheduled chunk: about six chunks per warp of
// the CTA, between 24 and 44 (measured optimum for the 30k-state denominator with register-destination
// gathers: 40-48; small chunks keep the tail of a frame short).  MK_CHUNK_ARCS overrides, for tuning.
static double chunk_cost_target(double cta_total) {
    const char* e = getenv("MK_CHUNK_ARCS");
    if (e && atoi(e) >= 4) return atoi(e);
    return std::min(44.0, std::max(24.0, cta_total / (6.0 * kSharedWarps)));
}

template <typename T> struct DirHost {
    std::vector<int4> items, chunks;
    std::vector<int2> item_arcs, item_pa;
    std::vector<int> cta_chunks, pidx;
    std::vector<T> pw;
    int cache_cap = 0, cache_items = 0, cache_chunks = 0;
};

// gflags[r] (kernels.cuh, item.w): run bookkeeping of row merging.  `tied[r]` = row r must stay in the
// same chunk as row r-1 (a non-first member of a run).  Backward (`reuse` rows): a tied row owns no
// arcs, it reuses the ⊕ of the run's first row.
template <typename T>
static void build_plan(const std::vector<int>& ptr, const std::vector<Arc<T>>& arcs, const std::vector<int>& pdf,
                       int S, int n_ctas, bool split, const std::vector<int>& gflags, const std::vector<char>& tied,
                       bool reuse, bool linear, double H, DirHost<T>& d, std::vector<int4>& long_rows,
                       std::vector<Arc<T>>& long_arcs, int& n_slots, const std::vector<double>* share = nullptr) {
    // Log semiring: the padded (streamed) arcs hold LINEAR weights 2^(w - R - H) that multiply the linear
    // copies 2^(v + H) of the vector; pads are 0.  Tropical: w itself, pads 0̄.
    const T ninf = linear ? T(0) : -std::numeric_limits<T>::infinity();
    n_slots = 0;
    for (int r = 0; r < S; ++r) {
        const int beg = ptr[r], end = ptr[r + 1], deg = end - beg;
        if (reuse && tied[r]) {
            // β of this row = β of the run's first row: no arcs
            d.items.push_back(make_int4(r, pdf[r], -1, gflags[r]));
            d.item_arcs.push_back(make_int2(beg, beg));
            continue;
        }
        if (!split || deg <= kLongRow) {
            d.items.push_back(make_int4(r, pdf[r], -1, gflags[r]));
            d.item_arcs.push_back(make_int2(beg, end));
            continue;
        }
        const int seg = std::max(kMinSegment, (deg + kMaxSlotsPerRow - 1) / kMaxSlotsPerRow);
        const int pseudo_beg = int(long_arcs.size());
        for (int a = beg; a < end; a += seg) {
            // (segments of a long row: bit6 per segment — the phony final state's in-arcs are all the final weights
            // plus its own self-loop, and only the segment holding the self-loop ever sees a live source once an
            // utterance is past its last frame)
            bool from_final = false;
            for (int k = a; k < std::min(end, a + seg); ++k) from_final = from_final || arcs[k].idx == S - 1;
            d.items.push_back(make_int4(r, pdf[r], n_slots, from_final ? 0 : kItemNoFinalPred));
            d.item_arcs.push_back(make_int2(a, std::min(end, a + seg)));
            Arc<T> pa;
            std::memset(&pa, 0, sizeof pa);
            pa.idx = n_slots++;
            pa.w = T(0);  // 1̄
            long_arcs.push_back(pa);
        }
        long_rows.push_back(make_int4(r, pseudo_beg, int(long_arcs.size()), pdf[r]));
    }
    const int n_items = int(d.items.size());
    // cost of an item in arc units: its arcs plus the finalise (measured ≈ 12 arcs' worth of warp time)
    auto cost = [&](int i) { return double(d.item_arcs[i].y - d.item_arcs[i].x) + kItemCost; };
    double total = 0;
    for (int i = 0; i < n_items; ++i) total += cost(i);
    std::vector<int> cta_items(n_ctas + 1, n_items);
    cta_items[0] = 0;
    double acc = 0;
    int i = 0;
    // `share` (calibration, mk_graph_create): CTA k's part of the cost — the measured relative speed of its SM on this
    // graph — instead of 1 / n_ctas
    std::vector<double> cum(n_ctas + 1, 0.0);
    for (int k = 0; k < n_ctas; ++k) cum[k + 1] = cum[k] + (share ? (*share)[k] : 1.0 / n_ctas);
    for (int k = 1; k < n_ctas; ++k) {
        const double target = total * cum[k] / cum[n_ctas];
        while (i < n_items) {
            double c = cost(i);
            if (acc + 0.5 * c > target) break;
            acc += c;
            ++i;
        }
        while (i < n_items && tied[d.items[i].x] && d.items[i].z < 0) ++i;  // never cut a run
        cta_items[k] = i;
    }
    d.cta_chunks.assign(n_ctas + 1, 0);
    d.item_pa.assign(n_items, make_int2(0, 0));
    auto emit = [&](int idx, T w) { d.pidx.push_back(idx); d.pw.push_back(w); };
    for (int k = 0; k < n_ctas; ++k) {
        const size_t first = d.chunks.size();
        const int cta_arc0 = int(d.pidx.size());
        int cb = cta_items[k], pb = int(d.pidx.size());
        // guided self-scheduling: chunks shrink from ~2x the target to 1/4 of it along the CTA's work, so
        // that the last chunks pulled (they are sorted largest first) are small and the frame ends evenly
        double cta_total = 0, done = 0, n_cost = 0;
        for (int it = cta_items[k]; it < cta_items[k + 1]; ++it) cta_total += cost(it);
        const double chunk_target = chunk_cost_target(cta_total);
        for (int it = cta_items[k]; it < cta_items[k + 1]; ++it) {
            const int2 ar = d.item_arcs[it];
            const bool owns_arcs = !(reuse && tied[d.items[it].x]);
            d.item_pa[it] = make_int2(int(d.pidx.size()), owns_arcs ? ar.y - ar.x : 0);
            if (owns_arcs) {
                for (int a = ar.x; a < ar.y; ++a)
                    emit(arcs[a].idx, linear ? T(std::exp2(double(arcs[a].w) - H)) : arcs[a].w);
                while (d.pidx.size() % 4) emit(0, ninf);  // every item starts on a quad
            }
            n_cost += cost(it);
            const bool next_tied = it + 1 < cta_items[k + 1] && tied[d.items[it + 1].x] && d.items[it + 1].z < 0;
            const double frac = cta_total > 0 ? done / cta_total : 1.0;
            const double target = std::max(2.0 * kItemCost, chunk_target * (2.0 - 1.75 * frac));
            if ((n_cost >= target && !next_tied) || it + 1 == cta_items[k + 1]) {
                done += n_cost;
                n_cost = 0;
                d.chunks.push_back(make_int4(pb, int(d.pidx.size()), cb, it + 1));
                cb = it + 1;
                pb = int(d.pidx.size());
            }
        }
        std::stable_sort(d.chunks.begin() + first, d.chunks.end(),
                         [](const int4& x, const int4& y) { return x.y - x.x > y.y - y.x; });
        d.cta_chunks[k + 1] = int(d.chunks.size());
        d.cache_chunks = std::max(d.cache_chunks, int(d.chunks.size() - first));
        d.cache_cap = std::max(d.cache_cap, int(d.pidx.size()) - cta_arc0);
        d.cache_items = std::max(d.cache_items, cta_items[k + 1] - cta_items[k]);
    }
    for (int q = 0; q < 4; ++q) emit(0, ninf);  // the global (uncached) path reads whole quads
}

template <typename T> static int upload_plan(const DirHost<T>& hst, const std::vector<Arc<T>>& arcs, DirDev& dev) {
    TRY(upload(hst.items, (void**)&dev.items));
    TRY(upload(hst.item_arcs, (void**)&dev.item_arcs));
    TRY(upload(hst.chunks, (void**)&dev.chunks));
    TRY(upload(hst.cta_chunks, (void**)&dev.cta_chunks));
    TRY(upload(hst.pidx, (void**)&dev.pidx));
    TRY(upload(hst.pw, &dev.pw));
    TRY(upload(hst.item_pa, (void**)&dev.item_pa));
    TRY(upload(arcs, &dev.arcs));
    dev.cache_cap = hst.cache_cap;
    dev.cache_items = hst.cache_items;
    dev.cache_chunks = hst.cache_chunks;
    return MK_OK;
}

// log of the largest ⊕-sum of a row's weights (Log semiring bound; rows given by ptr)
template <typename T> static double max_row_logsum(const std::vector<int>& ptr, const std::vector<Arc<T>>& arcs, int S) {
    double best = -std::numeric_limits<double>::infinity();
    for (int r = 0; r < S; ++r) {
        double m = -std::numeric_limits<double>::infinity();
        for (int a = ptr[r]; a < ptr[r + 1]; ++a) m = std::max(m, double(arcs[a].w));
        if (!(m > -std::numeric_limits<double>::infinity())) continue;
        double sum = 0;
        for (int a = ptr[r]; a < ptr[r + 1]; ++a) sum += std::exp(double(arcs[a].w) - m);
        best = std::max(best, m + std::log(sum));
    }
    return best;
}

// Runs a short synthetic pdfposteriors on the freshly built graph and returns the work cycles every CTA spent per sweep
// (defined after the dispatch code below).
static int calibrate_graph(mk_graph* g, std::vector<double>& fwd_cycles, std::vector<double>& bwd_cycles);

template <typename T>
static int build_graph(mk_graph* g, const int64_t* colptr, const int64_t* rowval, const void* nzval_,
                       int64_t n_init, const int64_t* init_idx, const void* init_w_,
                       const int32_t* state2pdf, int base) {
    const T* nzval = static_cast<const T*>(nzval_);
    const T* init_w = static_cast<const T*>(init_w_);
    const int S = int(g->S);
    const int64_t nnz = g->nnz;
    const T ninf = -std::numeric_limits<T>::infinity();

    std::vector<int> in_ptr(S + 1), out_ptr(S + 1, 0);
    std::vector<Arc<T>> in_arcs(nnz), out_arcs(nnz);
    for (int j = 0; j <= S; ++j) {
        int64_t v = colptr[j] - base;
        if (v < 0 || v > nnz || (j > 0 && v < in_ptr[j - 1]))
            return fail(MK_EINVAL, "colptr is not a valid CSC pointer array at column %d", j);
        in_ptr[j] = int(v);
    }
    if (in_ptr[0] != 0 || in_ptr[S] != nnz) return fail(MK_EINVAL, "colptr does not span nnz");
    for (int64_t a = 0; a < nnz; ++a) {
        int64_t s = rowval[a] - base;
        if (s < 0 || s >= S) return fail(MK_EINVAL, "rowval[%lld] out of range", (long long)a);
        std::memset(&in_arcs[a], 0, sizeof(Arc<T>));
        in_arcs[a].idx = int(s);
        in_arcs[a].w = g->prob ? T(std::log(double(nzval[a]))) : nzval[a];
    }
    // Julia's CSC keeps row indices ascending inside a column; enforce it (defines the tie rule)
    for (int j = 0; j < S; ++j) {
        auto b = in_arcs.begin() + in_ptr[j], e = in_arcs.begin() + in_ptr[j + 1];
        if (!std::is_sorted(b, e, [](const Arc<T>& x, const Arc<T>& y) { return x.idx < y.idx; }))
            std::stable_sort(b, e, [](const Arc<T>& x, const Arc<T>& y) { return x.idx < y.idx; });
        g->max_in_deg = std::max(g->max_in_deg, in_ptr[j + 1] - in_ptr[j]);
    }
    // T̂ rows (by source): the transpose, destinations ascending inside a row
    for (int64_t a = 0; a < nnz; ++a) out_ptr[in_arcs[a].idx + 1]++;
    for (int i = 0; i < S; ++i) {
        g->max_out_deg = std::max(g->max_out_deg, out_ptr[i + 1]);
        out_ptr[i + 1] += out_ptr[i];
    }
    {
        std::vector<int> fill(out_ptr.begin(), out_ptr.end() - 1);
        for (int j = 0; j < S; ++j)
            for (int a = in_ptr[j]; a < in_ptr[j + 1]; ++a) {
                int pos = fill[in_arcs[a].idx]++;
                std::memset(&out_arcs[pos], 0, sizeof(Arc<T>));
                out_arcs[pos].idx = j;
                out_arcs[pos].w = in_arcs[a].w;
            }
    }
    std::vector<int> pdf(S);
    for (int s = 0; s < S; ++s) {
        int64_t d = int64_t(state2pdf[s]) - base;
        if (d < 0 || d >= g->Dh) return fail(MK_EINVAL, "state2pdf[%d] out of range", s);
        pdf[s] = int(d);
    }
    if (pdf[S - 1] != g->Dh - 1)
        return fail(MK_EINVAL, "the phony final state must map to the phony pdf (n_pdf_hat)");
    for (int s = 0; s + 1 < S; ++s)
        if (pdf[s] == g->Dh - 1)
            return fail(MK_EINVAL, "state %d maps to the phony pdf; only the phony final state may", s);
    std::vector<T> init(S, ninf);
    for (int64_t k = 0; k < n_init; ++k) {
        int64_t s = init_idx[k] - base;
        if (s < 0 || s >= S) return fail(MK_EINVAL, "init_idx[%lld] out of range", (long long)k);
        init[s] = g->prob ? T(std::log(double(init_w[k]))) : init_w[k];
    }
    // Row merging: runs of adjacent states with bit-identical out-arc lists (the A/B state pairs of
    // the chain topology: same successors, same weights).  Backward: β is computed once per run and
    // reused by the other members.  Forward: the members' mass reaches their common successors through
    // ONE virtual source q_g = ⊕_{i in run} a[i] (row Ŝ + g of the forward vector), produced by the warp
    // that finalises the run — so every such successor row loses |run| - 1 in-arcs.  MK_NO_MERGE=1 disables.
    std::vector<int> grp(S, -1);
    std::vector<char> tied(S, 0), run_last(S, 0);
    int n_runs = 0, max_run = 1;
    if (!(getenv("MK_NO_MERGE") && atoi(getenv("MK_NO_MERGE")))) {
        auto same_row = [&](int x, int y) {
            const int n = out_ptr[x + 1] - out_ptr[x];
            if (n == 0 || n != out_ptr[y + 1] - out_ptr[y]) return false;
            return std::memcmp(&out_arcs[out_ptr[x]], &out_arcs[out_ptr[y]], size_t(n) * sizeof(Arc<T>)) == 0;
        };
        for (int i = 0; i + 1 < S;) {
            int j = i + 1;
            while (j < S && j - i < 8 && same_row(i, j)) ++j;
            bool ok = j - i >= 2;
            for (int k = i; ok && k < j; ++k) ok = in_ptr[k + 1] - in_ptr[k] <= kLongRow;  // segments cannot join a run
            if (!ok) { ++i; continue; }
            for (int k = i; k < j; ++k) { grp[k] = n_runs; tied[k] = k > i; }
            run_last[j - 1] = 1;
            max_run = std::max(max_run, j - i);
            ++n_runs;
            i = j;
        }
    }
    g->n_runs = n_runs;
    std::vector<int2> runs(n_runs);
    for (int k = 0; k < S; ++k)
        if (grp[k] >= 0) {
            if (!tied[k]) runs[grp[k]] = make_int2(k, 1);
            else runs[grp[k]].y += 1;
        }
    // forward in-arcs with the runs' members replaced by their virtual source
    std::vector<int> in_ptr_m(S + 1, 0);
    std::vector<Arc<T>> in_arcs_m;
    in_arcs_m.reserve(nnz);
    for (int j = 0; j < S; ++j) {
        for (int a = in_ptr[j]; a < in_ptr[j + 1]; ++a) {
            const int i = in_arcs[a].idx;
            if (grp[i] < 0) in_arcs_m.push_back(in_arcs[a]);
            else if (!tied[i]) { Arc<T> m = in_arcs[a]; m.idx = S + grp[i]; in_arcs_m.push_back(m); }
        }
        in_ptr_m[j + 1] = int(in_arcs_m.size());
    }
    // Static trimming: a state that no initial state reaches has α = 0̄ in every frame; a state from which
    // the phony final state cannot be reached has β = 0̄ in every frame before the last (when `expand`
    // builds the emissions: the phony frame kills every real state).  Such rows are flagged DEAD: their
    // all-zero ⊕ is taken at face value instead of going through the exact-fallback path every frame.
    std::vector<char> fwd_live(S, 0), bwd_live(S, 0);
    {
        std::vector<int> stack;
        for (int s = 0; s < S; ++s) if (init[s] > ninf) { fwd_live[s] = 1; stack.push_back(s); }
        while (!stack.empty()) {
            int s = stack.back(); stack.pop_back();
            for (int a = out_ptr[s]; a < out_ptr[s + 1]; ++a) {
                int t = out_arcs[a].idx;
                if (!fwd_live[t] && out_arcs[a].w > ninf) { fwd_live[t] = 1; stack.push_back(t); }
            }
        }
        bwd_live[S - 1] = 1; stack.push_back(S - 1);
        while (!stack.empty()) {
            int s = stack.back(); stack.pop_back();
            for (int a = in_ptr[s]; a < in_ptr[s + 1]; ++a) {
                int t = in_arcs[a].idx;
                if (!bwd_live[t] && in_arcs[a].w > ninf) { bwd_live[t] = 1; stack.push_back(t); }
            }
        }
    }
    // item flags (kernels.cuh): forward  bit0 = run member, bit1 = first, bit2 = last, bits 8.. = run index;
    //                           backward bit0 = reuse the previous item's ⊕, bit1 = owner of a run (its ⊕ is reused);
    //                           both: bit3 = dead row, bit4 = no arcs
    std::vector<int> gf_fwd(S, 0), gf_bwd(S, 0);
    for (int s = 0; s < S; ++s)
        if (grp[s] >= 0) {
            gf_fwd[s] = 1 | (tied[s] ? 0 : 2) | (run_last[s] ? 4 : 0) | (grp[s] << 8);
            gf_bwd[s] = tied[s] ? 1 : 2;  // bit1: this row's ⊕ is shared by the rows tied to it
        }
    for (int s = 0; s < S; ++s) {
        if (!fwd_live[s]) gf_fwd[s] |= kItemDead;
        if (!bwd_live[s]) gf_bwd[s] |= kItemDead;
        // a row without arcs sums to 0̄ whatever the emissions are
        if (in_ptr_m[s + 1] == in_ptr_m[s]) gf_fwd[s] |= kItemEmpty;
        if (out_ptr[s + 1] == out_ptr[s]) gf_bwd[s] |= kItemEmpty;
        // rows whose ⊕ is exactly 0̄ in the frame after α̂ / the frame before the phony frames (kernels.cuh, exact0)
        bool seeded = false;
        for (int a = in_ptr[s]; a < in_ptr[s + 1] && !seeded; ++a) seeded = init[in_arcs[a].idx] > ninf && in_arcs[a].w > ninf;
        if (!seeded) gf_fwd[s] |= kItemNoSeed;
        bool from_final = false;
        for (int a = in_ptr[s]; a < in_ptr[s + 1] && !from_final; ++a) from_final = in_arcs[a].idx == S - 1 && in_arcs[a].w > ninf;
        if (!from_final) gf_fwd[s] |= kItemNoFinalPred;
        bool to_final = false;
        for (int a = out_ptr[s]; a < out_ptr[s + 1] && !to_final; ++a) to_final = out_arcs[a].idx == S - 1 && out_arcs[a].w > ninf;
        if (!to_final) gf_bwd[s] |= kItemNoSeed;
    }

    // Bounds for the single-pass ⊕ (kernels.cuh): stored a_n <= max(log max column-sum, max α̂),
    // stored b_n ⊗ e' <= max(log max row-sum, 0); every exponent v + (w - R) is then <= 0.
    double R_f = 0, R_b = 0, vf = 0, vb = 0;
    if (g->semiring == MK_LOG) {
        const double ninf_d = -std::numeric_limits<double>::infinity();
        double wmax = ninf_d, li = ninf_d;
        for (int64_t a = 0; a < nnz; ++a) wmax = std::max(wmax, double(in_arcs[a].w));
        for (int64_t k = 0; k < n_init; ++k) li = std::max(li, double(init_w[k]));
        if (!(li > ninf_d && li < std::numeric_limits<double>::infinity())) li = 0;
        if (wmax > ninf_d && wmax < std::numeric_limits<double>::infinity()) {
            vf = std::max(max_row_logsum<T>(in_ptr, in_arcs, S), li) + std::log(double(max_run));  // q_g <= max + log|run|
            vb = std::max(max_row_logsum<T>(out_ptr, out_arcs, S), 0.0);
            R_f = vf + wmax;
            R_b = vb + wmax;
        } else {
            vf = li + std::log(double(max_run));  // (no usable arc: only frame 0 holds finite values)
        }
    }
    // the shared-graph kernel works in log2 units for the Log semiring.  The bound is placed at 2^kHeadroom
    // rather than at 1 so that the linear sums use the whole float range: terms may be as small as
    // 2^-100 (the underflow threshold of resolve_sum) and a row of up to 2^20 arcs still cannot overflow.
    const double unit = g->semiring == MK_LOG ? 1.4426950408889634 : 1.0;
    const double headroom = g->semiring == MK_LOG ? (sizeof(T) == 4 ? 100.0 : 900.0) : 0.0;  // log2 units
    R_f -= headroom / unit;
    R_b -= headroom / unit;
    std::vector<Arc<T>> in_s(in_arcs_m), out_s(out_arcs);
    for (auto& a : in_s) a.w = T((double(a.w) - R_f) * unit);
    for (auto& a : out_s) a.w = T((double(a.w) - R_b) * unit);
    std::vector<T> init_s(init);
    for (auto& x : init_s) x = T(double(x) * unit);
    g->fwd.R = R_f * unit;
    g->bwd.R = R_b * unit;
    // linear copies: stored values are <= vf (vb), so 2^(v + H) <= 2^headroom and the linear weights are
    // 2^((w - wmax) unit) <= 1
    const bool linear = g->semiring == MK_LOG;
    g->fwd.H = linear ? headroom - vf * unit : 0.0;
    g->bwd.H = linear ? headroom - vb * unit : 0.0;

    DirHost<T> fwd, bwd;
    std::vector<int4> fwd_long, no_long;
    std::vector<Arc<T>> fwd_long_arcs, no_arcs;
    int no_slots = 0;
    build_plan<T>(in_ptr_m, in_s, pdf, S, g->n_sms, true, gf_fwd, tied, false, linear, g->fwd.H, fwd, fwd_long,
                  fwd_long_arcs, g->n_slots);
    build_plan<T>(out_ptr, out_s, pdf, S, g->n_sms, false, gf_bwd, tied, true, linear, g->bwd.H, bwd, no_long,
                  no_arcs, no_slots);
    g->n_long = int(fwd_long.size());

    TRY(upload(in_ptr, (void**)&g->d_in_ptr));
    TRY(upload(out_ptr, (void**)&g->d_out_ptr));
    TRY(upload(in_arcs, &g->d_in_arcs));
    TRY(upload(out_arcs, &g->d_out_arcs));
    TRY(upload(pdf, (void**)&g->d_pdf));
    TRY(upload(init, &g->d_init_dense));
    TRY(upload(init_s, &g->d_init_dense_s));
    TRY(upload_plan<T>(fwd, in_s, g->fwd));
    TRY(upload_plan<T>(bwd, out_s, g->bwd));
    TRY(upload(runs, (void**)&g->d_runs));
    TRY(upload(fwd_long, (void**)&g->d_fwd_long));
    TRY(upload(fwd_long_arcs, &g->d_fwd_long_arcs));
    g->bytes = 2 * (S + 1) * sizeof(int) + 4 * nnz * sizeof(Arc<T>) + S * (sizeof(int) + 2 * sizeof(T)) +
               (fwd.pidx.size() + bwd.pidx.size()) * (sizeof(int) + sizeof(T));
    // Calibration: the cost model (arcs + 12 per item) and the SMs are not uniform — on cfg 3 the CTAs on one block of SM
    // ids take 8 % longer per frame than the mean, and every frame waits for the slowest.  One short synthetic call on the
    // new graph measures each CTA's work cycles per sweep; the plans are then rebuilt with every CTA's share of the cost
    // proportional to its measured speed (clipped to +-25 %), three times over (MK_CALIBRATE=n: n rounds, 0 disables); graphs that
    // never take the shared-graph kernel (fewer than 2 048 states) are not calibrated.
    const char* cal = getenv("MK_CALIBRATE");
    if (!(cal && cal[0] == '0') && S >= 2048) {
        const int rounds = (cal && atoi(cal) >= 1) ? std::min(atoi(cal), 12) : 3;
        std::vector<double> sf(g->n_sms, 1.0), sb(g->n_sms, 1.0);  // shares so far (relative)
        for (int round = 0; round < rounds; ++round) {
            std::vector<double> tf, tb;
            if (calibrate_graph(g, tf, tb) != MK_OK || int(tf.size()) != g->n_sms || int(tb.size()) != g->n_sms) {
                cudaGetLastError();  // (a failed calibration leaves the current plan in place)
                break;
            }
            // a CTA that took t with share s runs at speed s / t: the next share is proportional to it
            auto update = [&](std::vector<double>& sh, const std::vector<double>& t) {
                double mean_v = 0;
                std::vector<double> v(t.size(), 0.0);
                for (size_t k = 0; k < t.size(); ++k) { v[k] = t[k] > 0 ? sh[k] / t[k] : 0.0; mean_v += v[k]; }
                mean_v /= double(t.size());
                if (!(mean_v > 0)) return;
                for (size_t k = 0; k < t.size(); ++k) sh[k] = v[k] > 0 ? std::min(1.25, std::max(0.75, v[k] / mean_v)) : 1.0;
            };
            update(sf, tf);
            update(sb, tb);
            DirHost<T> fwd2, bwd2;
            std::vector<int4> fwd_long2, no_long2;
            std::vector<Arc<T>> fwd_long_arcs2, no_arcs2;
            int n_slots2 = 0, no_slots2 = 0;
            build_plan<T>(in_ptr_m, in_s, pdf, S, g->n_sms, true, gf_fwd, tied, false, linear, g->fwd.H, fwd2, fwd_long2,
                          fwd_long_arcs2, n_slots2, &sf);
            build_plan<T>(out_ptr, out_s, pdf, S, g->n_sms, false, gf_bwd, tied, true, linear, g->bwd.H, bwd2, no_long2,
                          no_arcs2, no_slots2, &sb);
            if (n_slots2 != g->n_slots || fwd_long2.size() != fwd_long.size()) break;  // (the items do not depend on the split)
            CK(cudaDeviceSynchronize());
            const double Rf = g->fwd.R, Hf = g->fwd.H, Rb = g->bwd.R, Hb = g->bwd.H;
            g->fwd.release(); g->bwd.release();
            g->fwd = DirDev(); g->bwd = DirDev();
            g->fwd.R = Rf; g->fwd.H = Hf; g->bwd.R = Rb; g->bwd.H = Hb;
            TRY(upload_plan<T>(fwd2, in_s, g->fwd));
            TRY(upload_plan<T>(bwd2, out_s, g->bwd));
            g->calibrated = true;
        }
    }
    return MK_OK;
}

// ------------------------------------------------------------------------------------------------
// mk_batch
// ------------------------------------------------------------------------------------------------
struct Group {  // utterances sharing one graph, run by shared_fb_kernel
    mk_graph* g = nullptr;
    std::vector<int> utts;
    int U4 = 0;
    bool vec4 = false;
    int* d_utt_b = nullptr;
    long long* d_utt_off = nullptr;
    DevBuf E, emax, emax_key, alpha, bt, flin, blin, part, gkey, coff, lz2, carry, tile_n1, trace;
    std::vector<unsigned char> h_trace;  // back-trace descriptors (bestpath), uploaded once
    size_t trace_tsize = 0;
    DevBuf post_stage, lane_of;          // ragged groups sorted by length: lane-ordered posteriors, utterance -> lane
    std::vector<int> h_lane_of;
    bool staged = false;                 // the current call scatters into post_stage
    std::vector<int> h_tile_n1;  // ragged batches: frames each utterance tile needs (staging for tile_n1)
    std::vector<int> h_order;    // group lane u -> utterance b of the current call (utts, or its length-sorted quads)
    std::vector<int> h_utt_b;    // staging for d_utt_b when the order changes
    bool permuted = false;       // d_utt_b currently holds a sorted order
};

struct mk_batch {
    int64_t B = 0, total = 0;
    int semiring = 0, dtype = 0, device = 0, n_sms = 0;
    bool prob = false;
    int64_t Dh = 0;
    std::vector<mk_graph*> graphs;
    std::vector<int64_t> off;
    std::vector<Group> groups;
    std::vector<int> small;  // utterances run by small_fb_kernel
    int small_smax = 0;
    int64_t small_cached_n1 = -1;
    DevBuf small_descs, small_alpha, small_ca, zsum, lz, seqlens, barrier, trace, h_ll, h_post, h_logz, h_path, zlimit;
    std::vector<int> h_zlimit;  // ragged batches: frames evaluated per utterance (staging for zlimit)
    std::vector<unsigned char> h_trace;  // back-trace descriptors of the per-utterance kernel's utterances (bestpath)
    int64_t trace_n1 = -1;
    size_t trace_tsize = 0;
    bool ragged_cut = false;    // the current call stops at least one utterance tile early
    bool host_pending = false;  // a mk_pdfposteriors_host_begin call has not been waited for
    // Threads per CTA of the shared-graph kernel: kSharedThreads (one CTA per SM owns the SM), or half of it so that the
    // sweeps of TWO batches in flight share every SM (mk_batch_set_overlap): what one waits for at its grid barrier,
    // the other computes.
    int shared_threads = kSharedThreads;
    cudaStream_t own_stream = nullptr, copy_stream = nullptr;
    // One in-flight call per batch: the workspaces are shared by every entry point.  Each call records `ev_last` on
    // its stream when it has enqueued its work; a following call on ANOTHER stream (the *_host entry points run on
    // own_stream, device calls on the caller's) waits for it before touching the workspaces.
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
    static constexpr int kMaxSegments = 16;
    cudaEvent_t ev_h2d[kMaxSegments] = {}, ev_done[kMaxSegments] = {};
    size_t max_smem_optin = 0;
    // optional timing of the dominant kernel (bench.py roofline): events around the last
    // shared_fb_kernel launch, on the stream it was launched on
    static constexpr int kProfRing = 64;
    bool profile = false;
    int prof_n = 0;  // launches recorded since profiling was (re)enabled
    cudaEvent_t ev0[kProfRing] = {}, ev1[kProfRing] = {};
    ~mk_batch() {
        for (auto& gr : groups) {
            cudaFree(gr.d_utt_b); cudaFree(gr.d_utt_off);
            gr.E.release(); gr.alpha.release(); gr.bt.release(); gr.flin.release(); gr.blin.release();
            gr.part.release(); gr.gkey.release(); gr.coff.release(); gr.emax.release(); gr.emax_key.release(); gr.lz2.release(); gr.carry.release(); gr.tile_n1.release(); gr.trace.release(); gr.post_stage.release(); gr.lane_of.release();
        }
        DevBuf* all[] = {&small_descs, &small_alpha, &small_ca, &zsum, &lz, &seqlens, &barrier, &trace,
                         &h_ll, &h_post, &h_logz, &h_path, &zlimit};
        for (DevBuf* d : all) d->release();
        if (own_stream) cudaStreamDestroy(own_stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (ev_last) cudaEventDestroy(ev_last);
        for (int i = 0; i < kMaxSegments; ++i) {
            if (ev_h2d[i]) cudaEventDestroy(ev_h2d[i]);
            if (ev_done[i]) cudaEventDestroy(ev_done[i]);
        }
        for (int i = 0; i < kProfRing; ++i) {
            if (ev0[i]) cudaEventDestroy(ev0[i]);
            if (ev1[i]) cudaEventDestroy(ev1[i]);
        }
    }
    size_t ws_bytes() const {
        size_t t = small_descs.cap + small_alpha.cap + small_ca.cap + zsum.cap + lz.cap + seqlens.cap + barrier.cap +
                   trace.cap + h_ll.cap + h_post.cap + h_logz.cap + h_path.cap;
        for (auto& gr : groups) t += gr.E.cap + gr.alpha.cap + gr.bt.cap + gr.flin.cap + gr.blin.cap + gr.part.cap + gr.gkey.cap + gr.coff.cap + gr.emax.cap;
        return t;
    }
};

static size_t tsize(int dtype) { return dtype == MK_F32 ? 4 : 8; }
static size_t small_smem_bytes(int S, int dtype) { return (2 * size_t(S) + 64) * tsize(dtype) + 64 * sizeof(int); }

// ------------------------------------------------------------------------------------------------
// run
// ------------------------------------------------------------------------------------------------
enum Mode { MODE_ALPHA, MODE_BETA, MODE_POST, MODE_BEST };

// A call cut into frame segments (mk_pdfposteriors_host): the emissions of segment k arrive while the
// forward sweep works on segment k-1, the posteriors of segment k leave while the backward sweep works
// on segment k-1.  `before_fwd(k)` runs before the emission transform of segment k is enqueued (waits for
// the copy), `after_bwd(k)` after the posteriors of segment k have been normalised (starts the copy).
struct Segments {
    std::vector<int> f;  // frame boundaries: f[0] = 0 < f[1] < ... < f[K] = N̂
    std::function<int(int)> before_fwd, after_bwd;
};

static bool ragged_cut_enabled() {
    const char* e = getenv("MK_RAGGED_CUT");
    return !(e && e[0] == '0');
}

static bool ragged_sort_enabled() {
    const char* e = getenv("MK_RAGGED_SORT");
    return !(e && e[0] == '0');
}

struct CallArgs {
    const void* ll; int64_t sb, sd, sn, D, T; int expanded; const int32_t* seqlens;
    void* out0;  // A / B / post / path
    void* out1;  // logz / score
    double* stats = nullptr;  // MODE_POST, optional: device float64 [D + 2] step statistics {Σ logZ, #frames, occupancy[D]}
    cudaStream_t stream;
    const Segments* seg = nullptr;  // host pipeline (single shared-graph group only)
};

template <typename T, int SR>
static int launch_shared(mk_batch* bt, Group& gr, Mode mode, const CallArgs& c, int Dh, int N1,
                         int Dout, int Tout, const int* d_seqlens, const Segments* seg = nullptr) {
    mk_graph* g = gr.g;
    const int S = int(g->S), U4 = gr.U4;
    const int Sq = S + g->n_runs;  // rows of the forward vector: states + merged-run sources
    const size_t frame = size_t(S) * U4 * sizeof(T), frame_q = size_t(Sq) * U4 * sizeof(T);
    if (size_t(Sq) * U4 >= (size_t(1) << 31)) return fail(MK_ENOTSUP, "Ŝ*U exceeds 2^31 in one group");
    TRY(gr.E.ensure(size_t(N1) * Dh * U4 * sizeof(T)));
    // α store: the states only (the merged runs' sources q_g live in the forward gather ping-pong)
    const size_t frame_a = frame;
    TRY(gr.alpha.ensure(size_t(N1) * frame_a));
    if (mode == MODE_POST || mode == MODE_BETA) TRY(gr.bt.ensure(2 * frame));
    if (mode != MODE_BETA) TRY(gr.flin.ensure(2 * frame_q));  // forward gather source of the two frames in flight
    if (SR == SR_LOG && (mode == MODE_POST || mode == MODE_BETA)) TRY(gr.blin.ensure(2 * frame));  // backward: linear copies
    TRY(gr.part.ensure(2 * size_t(std::max(g->n_slots, 1)) * U4 * sizeof(T)));
    TRY(gr.gkey.ensure(2 * size_t(N1) * U4 * sizeof(int)));
    TRY(gr.emax.ensure(size_t(N1) * U4 * sizeof(T)));
    if (SR == SR_LOG) TRY(gr.emax_key.ensure(size_t(N1) * U4 * sizeof(int)));
    TRY(gr.lz2.ensure(size_t(U4) * sizeof(double)));
    TRY(gr.coff.ensure(2 * size_t(N1) * U4 * sizeof(double)));
    TRY(gr.carry.ensure(size_t(U4) * (sizeof(double) + sizeof(T))));

    // Utterance order of this call.  Ragged posteriors over several utterance tiles: the group's lane quads (4 consecutive
    // utterances: the unit of the 16-byte posterior reductions) are sorted by length, longest first, so that the per-tile
    // frame limits below bite whatever order the caller's batch has.  Everything downstream goes through d_utt_b (the
    // emission transform, lz / zsum / the posterior scatter), so results land at the caller's utterance indices.
    const bool want_cut = mode == MODE_POST && c.seqlens && !c.expanded && ragged_cut_enabled();
    gr.h_order = gr.utts;
    bool permute = false;
    gr.staged = false;
    // Single utterances sorted by length, longest first, when the whole batch is this one group: the posteriors are then
    // scattered into a lane-ordered staging array (16-byte reductions whatever the order) and the normalisation pass
    // writes them to the caller's utterance order (MK_RAGGED_SORT=quads keeps the quad sort below, =0 no sort at all).
    const char* sort_env = getenv("MK_RAGGED_SORT");
    const bool by_utt = want_cut && U4 > kTileUtts && ragged_sort_enabled() && bt->groups.size() == 1 && bt->small.empty() &&
                        !(sort_env && sort_env[0] == 'q');
    if (by_utt) {
        std::vector<int> q(gr.utts.size());
        for (size_t i = 0; i < q.size(); ++i) q[i] = int(i);
        std::stable_sort(q.begin(), q.end(), [&](int a, int b) { return c.seqlens[gr.utts[a]] > c.seqlens[gr.utts[b]]; });
        for (size_t i = 0; i < q.size() && !permute; ++i) permute = q[i] != int(i);
        if (permute) {
            for (size_t i = 0; i < q.size(); ++i) gr.h_order[i] = gr.utts[q[i]];
            gr.staged = true;
        }
    } else if (want_cut && U4 > kTileUtts && gr.utts.size() % 4 == 0 && ragged_sort_enabled()) {
        const int nq = int(gr.utts.size() / 4);
        std::vector<int> q(nq), key(nq, 0);
        for (int i = 0; i < nq; ++i) {
            q[i] = i;
            for (int j = 0; j < 4; ++j) key[i] = std::max(key[i], int(c.seqlens[gr.utts[4 * i + j]]));
        }
        std::stable_sort(q.begin(), q.end(), [&](int a, int b) { return key[a] > key[b]; });
        for (int i = 0; i < nq && !permute; ++i) permute = q[i] != i;
        if (permute)
            for (int i = 0; i < nq; ++i)
                for (int j = 0; j < 4; ++j) gr.h_order[4 * i + j] = gr.utts[4 * q[i] + j];
    }
    if (permute || gr.permuted) {
        gr.h_utt_b.assign(U4, -1);
        std::copy(gr.h_order.begin(), gr.h_order.end(), gr.h_utt_b.begin());
        CK(cudaMemcpyAsync(gr.d_utt_b, gr.h_utt_b.data(), size_t(U4) * sizeof(int), cudaMemcpyHostToDevice, c.stream));
        gr.permuted = permute;
    }

    if (gr.staged) {
        gr.h_lane_of.assign(size_t(bt->B), -1);
        for (size_t u = 0; u < gr.h_order.size(); ++u) gr.h_lane_of[gr.h_order[u]] = int(u);
        TRY(gr.lane_of.ensure(size_t(bt->B) * sizeof(int)));
        CK(cudaMemcpyAsync(gr.lane_of.p, gr.h_lane_of.data(), size_t(bt->B) * sizeof(int), cudaMemcpyHostToDevice, c.stream));
        const size_t stage_bytes = size_t(Tout) * Dout * U4 * sizeof(T);
        TRY(gr.post_stage.ensure(stage_bytes));
        CK(cudaMemsetAsync(gr.post_stage.p, 0, stage_bytes, c.stream));
    }
    // frame segments of this call (one, unless the host pipeline cut it)
    std::vector<int> fb = seg ? seg->f : std::vector<int>{0, N1};
    const int K = int(fb.size()) - 1;

    EmisParams<T> ep;
    ep.ll = static_cast<const T*>(c.ll); ep.sb = c.sb; ep.sd = c.sd; ep.sn = c.sn;
    ep.D = int(c.D); ep.Tn = int(c.T); ep.expanded = c.expanded | (bt->prob ? 2 : 0); ep.Dh = Dh; ep.N1 = N1;
    ep.seqlens = d_seqlens; ep.utt_b = gr.d_utt_b; ep.U4 = U4;
    ep.scale = SR == SR_LOG ? T(1.4426950408889634) : T(1);
    // expand + transpose + per-frame emission maxima of the frames [n0, n1)
    auto emissions = [&](int n0, int n1) -> int {
        const int nf = n1 - n0;
        EmisParams<T> e = ep;
        e.n0 = n0;
        e.E = static_cast<T*>(gr.E.p);  // (the kernel indexes both arrays by the absolute frame n0 + blockIdx.z)
        e.emax_key = SR == SR_LOG ? static_cast<int*>(gr.emax_key.p) : nullptr;
        int* keys = static_cast<int*>(gr.emax_key.p) + size_t(n0) * U4;
        // keys start below every finite value (0x80808080 decodes to -3.4e38)
        if (SR == SR_LOG) CK(cudaMemsetAsync(keys, 0x80, size_t(nf) * U4 * sizeof(int), c.stream));
        for (int z0 = 0; z0 < nf; z0 += 65535) {  // (gridDim.z is limited to 65535 frames per launch)
            e.n0 = n0 + z0;
            // (32 x 32 tiles when two batches share the SMs: the small blocks slip in beside the other batch's sweep, the
            // wide ones wait for it — 7.15 against 7.65 ms per batch; MK_NARROW_TRANSPOSE forces them, for A/B timing)
            if (getenv("MK_NARROW_TRANSPOSE") || bt->shared_threads != kSharedThreads) {
                dim3 eg((Dh + 31) / 32, (U4 + 31) / 32, std::min(nf - z0, 65535)), eb(32, 8);
                expand_transpose_kernel<T><<<eg, eb, 0, c.stream>>>(e);
            } else {
                constexpr int DT = WideTile<T>::d;
                dim3 eg((Dh + DT - 1) / DT, (U4 + 127) / 128, std::min(nf - z0, 65535));
                expand_transpose_wide_kernel<T><<<eg, 256, 0, c.stream>>>(e);
            }
            CK(cudaGetLastError());
            ++g_launches;
        }
        T* emax = static_cast<T*>(gr.emax.p) + size_t(n0) * U4;
        if (SR == SR_LOG) {
            const int count = nf * U4;
            emission_max_decode_kernel<T><<<(count + 255) / 256, 256, 0, c.stream>>>(keys, emax, count);
            CK(cudaGetLastError());
            ++g_launches;
        } else {
            CK(cudaMemsetAsync(emax, 0, size_t(nf) * U4 * sizeof(T), c.stream));
        }
        return MK_OK;
    };
    if (!seg) TRY(emissions(0, N1));

    SharedParams<T> p;
    p.S = S; p.Sq = Sq; p.Dh = Dh; p.N1 = N1; p.U4 = U4; p.ntiles = (U4 + kTileUtts - 1) / kTileUtts;
    auto plan = [](const DirDev& d) {
        DirPlan<T> q;
        q.items = d.items; q.item_arcs = d.item_arcs; q.chunks = d.chunks; q.cta_chunks = d.cta_chunks;
        q.pidx = d.pidx; q.pw = static_cast<const T*>(d.pw); q.item_pa = d.item_pa;
        q.arcs = static_cast<const Arc<T>*>(d.arcs); q.R = T(d.R); q.H = T(d.H);
        q.runs = nullptr; q.n_states = 0;
        return q;
    };
    p.fwd = plan(g->fwd); p.bwd = plan(g->bwd);
    p.fwd.runs = g->n_runs ? g->d_runs : nullptr; p.fwd.n_states = S;
    p.n_long = g->n_long; p.fwd_long = g->d_fwd_long;
    p.fwd_long_arcs = static_cast<const Arc<T>*>(g->d_fwd_long_arcs);
    p.n_slots = g->n_slots; p.part = static_cast<T*>(gr.part.p);
    p.init_dense = static_cast<const T*>(g->d_init_dense_s);
    p.emax = static_cast<const T*>(gr.emax.p);
    p.lz2 = static_cast<double*>(gr.lz2.p);
    p.gkey = static_cast<int*>(gr.gkey.p); p.Coff = static_cast<double*>(gr.coff.p);
    p.E = static_cast<const T*>(gr.E.p);
    p.alpha = static_cast<T*>(gr.alpha.p);
    p.bt = static_cast<T*>(gr.bt.p);
    p.flin = static_cast<T*>(gr.flin.p); p.blin = static_cast<T*>(gr.blin.p);
    p.beta_out = nullptr;
    p.post = nullptr; p.B = int(bt->B); p.D = Dout; p.Tn = Tout;
    p.utt_b = gr.d_utt_b; p.post_vec4 = 0; p.post_ld = 0;
    p.zsum = static_cast<T*>(bt->zsum.p); p.lz = static_cast<T*>(bt->lz.p);
    p.barrier = static_cast<unsigned*>(bt->barrier.p);
    p.cta_cycles = g->d_cta_cycles;
    p.carry_C = static_cast<double*>(gr.carry.p);
    p.carry_shift = reinterpret_cast<T*>(static_cast<double*>(gr.carry.p) + U4);
    p.n_lo = 0; p.n_hi = N1;
    p.do_fwd = p.do_bwd = p.do_post = 0;
    p.bwd_dead_ok = c.expanded ? 0 : 1;
    // Ragged batch: an utterance tile (128 consecutive utterances of the group) whose longest
    // sequence has L < T frames runs frames 0..L only.  The lane quads were sorted by length above
    // (MK_RAGGED_SORT=0 keeps the caller's order); any order is correct.  MK_RAGGED_CUT=0 disables the limits (the tests
    // compare both).
    p.tile_n1 = nullptr;
    p.seqlens = d_seqlens;
    if (want_cut) {
        gr.h_tile_n1.assign(p.ntiles, 2);
        bool cut = false;
        for (int t = 0; t < p.ntiles; ++t) {
            int lim = 2;
            const size_t k1 = std::min(gr.h_order.size(), size_t(t + 1) * kTileUtts);
            for (size_t k = size_t(t) * kTileUtts; k < k1; ++k) lim = std::max(lim, c.seqlens[gr.h_order[k]] + 1);
            gr.h_tile_n1[t] = lim = std::min(lim, N1);
            cut = cut || lim < N1;
            for (size_t k = size_t(t) * kTileUtts; k < k1; ++k) bt->h_zlimit[gr.h_order[k]] = lim;
        }
        if (cut) {
            TRY(gr.tile_n1.ensure(p.ntiles * sizeof(int)));
            CK(cudaMemcpyAsync(gr.tile_n1.p, gr.h_tile_n1.data(), p.ntiles * sizeof(int), cudaMemcpyHostToDevice, c.stream));
            p.tile_n1 = static_cast<const int*>(gr.tile_n1.p);
            bt->ragged_cut = true;
        }
    }
    p.ablate = getenv("MK_ABLATE") ? atoi(getenv("MK_ABLATE")) : 0;
#ifdef MK_ABLATE
    { int ns = (p.ablate & 16) ? 1 : 0; CK(cudaMemcpyToSymbol(g_no_stores, &ns, sizeof ns)); }
#endif
    switch (mode) {
        case MODE_ALPHA: case MODE_BEST: p.do_fwd = 1; break;
        case MODE_BETA: p.do_bwd = 1; p.beta_out = static_cast<T*>(gr.alpha.p); break;
        case MODE_POST:
            p.do_fwd = p.do_bwd = p.do_post = 1;
            p.post = static_cast<T*>(c.out0);
            p.post_vec4 = (gr.vec4 && bt->B % 4 == 0 && (reinterpret_cast<uintptr_t>(c.out0) & 15) == 0) ? 1 : 0;
            if (gr.staged) { p.post = static_cast<T*>(gr.post_stage.p); p.post_ld = U4; }
            break;
    }
    void* args[] = {&p};
    const size_t scal = shared_scalars_bytes(U4, sizeof(T));
    if (scal + 16 * 1024 > bt->max_smem_optin)
        return fail(MK_ENOTSUP, "%d utterances share one graph: the per-utterance scalars (%zu bytes of shared memory) "
                                "leave no room for the kernel's working set; split the batch into several mk_batch objects",
                    U4, scal);
    const int slot = bt->prof_n % mk_batch::kProfRing;
    if (bt->profile) CK(cudaEventRecord(bt->ev0[slot], c.stream));
    for (int phase = 0; phase < 2; ++phase) {
        if (phase == 0 ? !p.do_fwd : !p.do_bwd) continue;
        const DirDev& dd = phase == 0 ? g->fwd : g->bwd;
        // shared-memory arc cache of this sweep when it fits next to the scalars
        size_t smem = scal;
        const size_t need = arc_cache_bytes(dd.cache_cap, dd.cache_items, dd.cache_chunks, sizeof(T));
        const bool sa = smem + need <= bt->max_smem_optin;
        p.cache_f = p.cache_b = p.cache_items_f = p.cache_items_b = 0;
        if (sa) {
            smem += need;
            if (phase == 0) { p.cache_f = dd.cache_cap; p.cache_items_f = dd.cache_items; }
            else { p.cache_b = dd.cache_cap; p.cache_items_b = dd.cache_items; }
        }
        void (*kern)(SharedParams<T>) =
            phase == 0 ? (sa ? shared_fb_kernel<T, SR, true, 0> : shared_fb_kernel<T, SR, false, 0>)
                       : (sa ? shared_fb_kernel<T, SR, true, 1> : shared_fb_kernel<T, SR, false, 1>);
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        for (int kk = 0; kk < K; ++kk) {  // forward: segments upwards; backward: downwards
            const int k = phase == 0 ? kk : K - 1 - kk;
            p.n_lo = fb[k]; p.n_hi = fb[k + 1];
            if (seg && phase == 0) {
                if (seg->before_fwd) TRY(seg->before_fwd(k));
                TRY(emissions(p.n_lo, p.n_hi));
            }
            CK(cudaMemsetAsync(bt->barrier.p, 0, sizeof(unsigned), c.stream));
            CK(cudaLaunchCooperativeKernel((void*)kern, dim3(g->n_sms), dim3(bt->shared_threads), args, smem, c.stream));
            ++g_launches;
            if (seg && phase == 1 && mode == MODE_POST) {
                // Ẑ ./ sums for the real frames of this segment, then hand them to the caller
                const int t0 = fb[k], t1 = std::min(fb[k + 1], Tout);
                if (t1 > t0) {
                    dim3 ng((Dout + 7) / 8, (t1 - t0 + kNormFrames - 1) / kNormFrames);
                    if (gr.staged)
                        normalize_permuted_kernel<T><<<ng, 256, 0, c.stream>>>(
                            static_cast<const T*>(gr.post_stage.p) + size_t(t0) * Dout * U4, U4,
                            static_cast<const int*>(gr.lane_of.p), static_cast<T*>(c.out0) + size_t(t0) * Dout * bt->B,
                            static_cast<const T*>(bt->zsum.p) + size_t(t0) * bt->B, int(bt->B), Dout, t1 - t0,
                            c.stats ? c.stats + 2 : nullptr);
                    else
                        normalize_post_kernel<T><<<ng, 256, 0, c.stream>>>(
                            static_cast<T*>(c.out0) + size_t(t0) * Dout * bt->B,
                            static_cast<const T*>(bt->zsum.p) + size_t(t0) * bt->B, int(bt->B), Dout, t1 - t0,
                            c.stats ? c.stats + 2 : nullptr);
                    CK(cudaGetLastError());
                    ++g_launches;
                }
                if (seg->after_bwd) TRY(seg->after_bwd(k));
            }
        }
    }
    if (bt->profile) { CK(cudaEventRecord(bt->ev1[slot], c.stream)); ++bt->prof_n; }

    if (mode == MODE_ALPHA || mode == MODE_BETA) {
        const double* C = static_cast<const double*>(gr.coff.p) + (mode == MODE_BETA ? size_t(N1) * U4 : 0);
        for (int z0 = 0; z0 < N1; z0 += 65535) {  // (gridDim.z is limited to 65535 frames per launch)
            dim3 ug((S + 31) / 32, (U4 + 31) / 32, std::min(N1 - z0, 65535)), ub(32, 8);
            unpack_states_kernel<T><<<ug, ub, 0, c.stream>>>(static_cast<const T*>(gr.alpha.p), S,
                                                            S, U4, gr.d_utt_b,
                                                            gr.d_utt_off, C, SR == SR_LOG ? 0.6931471805599453 : 1.0,
                                                            static_cast<T*>(c.out0), bt->total, z0, bt->prob ? 1 : 0);
            CK(cudaGetLastError());
            ++g_launches;
        }
    }
    return MK_OK;
}

template <typename T, int SR>
static int launch_small(mk_batch* bt, Mode mode, const CallArgs& c, int Dh, int N1, int Dout, int Tout,
                        const int* d_seqlens) {
    const int n = int(bt->small.size());
    if (n == 0) return MK_OK;
    // descriptors (ws_off depends on N̂)
    if (bt->small_cached_n1 != N1) {
        std::vector<UttDesc<T>> descs(n);
        long long off = 0;
        for (int k = 0; k < n; ++k) {
            int b = bt->small[k];
            mk_graph* g = bt->graphs[b];
            UttDesc<T>& d = descs[k];
            d.in_ptr = g->d_in_ptr; d.in_arcs = static_cast<const Arc<T>*>(g->d_in_arcs);
            d.out_ptr = g->d_out_ptr; d.out_arcs = static_cast<const Arc<T>*>(g->d_out_arcs);
            d.pdf = g->d_pdf; d.init_dense = static_cast<const T*>(g->d_init_dense);
            d.S = int(g->S); d.b = b; d.ws_off = off; d.out_off = bt->off[b]; d.c_off = (long long)k * N1;
            off += (long long)N1 * g->S;
        }
        TRY(bt->small_descs.ensure(n * sizeof(UttDesc<T>)));
        TRY(bt->small_alpha.ensure(size_t(off) * sizeof(T)));
        TRY(bt->small_ca.ensure(size_t(n) * N1 * sizeof(double)));
        CK(cudaMemcpyAsync(bt->small_descs.p, descs.data(), n * sizeof(UttDesc<T>), cudaMemcpyHostToDevice,
                           c.stream));
        CK(cudaStreamSynchronize(c.stream));  // descs is a stack-lifetime host buffer
        bt->small_cached_n1 = N1;
    }
    SmallParams<T> p;
    p.utts = static_cast<const UttDesc<T>*>(bt->small_descs.p);
    p.ll = static_cast<const T*>(c.ll); p.sb = c.sb; p.sd = c.sd; p.sn = c.sn;
    p.D = Dout; p.Tn = Tout; p.expanded = c.expanded | (bt->prob ? 2 : 0); p.Dh = Dh; p.N1 = N1;
    p.seqlens = d_seqlens;
    p.alpha = static_cast<T*>(bt->small_alpha.p); p.alpha_sn = 0; p.alpha_user = 0;
    p.beta_out = nullptr; p.beta_sn = 0;
    p.Ca = static_cast<double*>(bt->small_ca.p);
    p.post = nullptr; p.B = int(bt->B);
    p.zsum = static_cast<T*>(bt->zsum.p); p.lz = static_cast<T*>(bt->lz.p);
    p.do_fwd = p.do_bwd = p.do_post = 0;
    switch (mode) {
        case MODE_ALPHA: p.do_fwd = 1; p.alpha = static_cast<T*>(c.out0); p.alpha_sn = bt->total; p.alpha_user = 1; break;
        case MODE_BEST: p.do_fwd = 1; break;
        case MODE_BETA: p.do_bwd = 1; p.beta_out = static_cast<T*>(c.out0); p.beta_sn = bt->total; break;
        case MODE_POST: p.do_fwd = p.do_bwd = p.do_post = 1; p.post = static_cast<T*>(c.out0); break;
    }
    const int smax = bt->small_smax;
    int threads = std::min(1024, std::max(64, ((smax + 31) / 32) * 32));
    size_t smem = small_smem_bytes(smax, bt->dtype);
    auto kern = small_fb_kernel<T, SR>;
    if (smem > 48 * 1024)
        CK(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<n, threads, smem, c.stream>>>(p);
    CK(cudaGetLastError());
    ++g_launches;
    return MK_OK;
}

template <typename T, int SR> static int run(mk_batch* bt, Mode mode, const CallArgs& c) {
    DeviceGuard guard(bt->device);
    if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", bt->device);
    const int B = int(bt->B);
    if (c.D <= 0 || c.T <= 0) return fail(MK_EINVAL, "D and T must be positive");
    const int Dh = c.expanded ? int(c.D) : int(c.D) + 1;
    const int N1 = c.expanded ? int(c.T) : int(c.T) + 1;
    const int Dout = Dh - 1, Tout = N1 - 1;
    if (Dh != bt->Dh)
        return fail(MK_EINVAL, "DimensionMismatch: emissions have %d pdfs (incl. phony), graphs expect %lld",
                    Dh, (long long)bt->Dh);
    if (c.expanded && c.seqlens) return fail(MK_EINVAL, "seqlens must be NULL with expanded emissions");
    if (N1 < 2) return fail(MK_EINVAL, "need at least one real frame");
    if (!c.ll || !c.out0) return fail(MK_EINVAL, "null buffer");
    const int* d_seqlens = nullptr;
    if (c.seqlens) {
        for (int b = 0; b < B; ++b)
            if (c.seqlens[b] < 0 || c.seqlens[b] > c.T)
                return fail(MK_EINVAL, "seqlens[%d] = %d outside [0, T]", b, c.seqlens[b]);
        TRY(bt->seqlens.ensure(B * sizeof(int)));
        CK(cudaMemcpyAsync(bt->seqlens.p, c.seqlens, B * sizeof(int), cudaMemcpyHostToDevice, c.stream));
        d_seqlens = static_cast<const int*>(bt->seqlens.p);
    }
    TRY(bt->barrier.ensure(256));
    TRY(bt->lz.ensure(B * sizeof(T)));
    TRY(bt->zsum.ensure(size_t(N1) * B * sizeof(T)));
    if (mode == MODE_POST) {
        if (!c.out1) return fail(MK_EINVAL, "null logz buffer");
        CK(cudaMemsetAsync(c.out0, 0, size_t(Tout) * Dout * B * sizeof(T), c.stream));
        CK(cudaMemsetAsync(bt->zsum.p, 0, size_t(N1) * B * sizeof(T), c.stream));
        if (c.stats) CK(cudaMemsetAsync(c.stats, 0, size_t(Dout + 2) * sizeof(double), c.stream));
    }
    const Segments* seg = (c.seg && bt->groups.size() == 1 && bt->small.empty()) ? c.seg : nullptr;
    bt->ragged_cut = false;
    bt->h_zlimit.assign(B, N1);
    for (auto& gr : bt->groups) TRY((launch_shared<T, SR>(bt, gr, mode, c, Dh, N1, Dout, Tout, d_seqlens, seg)));
    TRY((launch_small<T, SR>(bt, mode, c, Dh, N1, Dout, Tout, d_seqlens)));

    if (mode == MODE_POST) {
        if (!seg) {  // (a segmented call normalises segment by segment, inside launch_shared)
            dim3 ng((Dout + 7) / 8, (Tout + kNormFrames - 1) / kNormFrames);
            if (bt->groups.size() == 1 && bt->groups[0].staged) {
                Group& gr = bt->groups[0];
                normalize_permuted_kernel<T><<<ng, 256, 0, c.stream>>>(
                    static_cast<const T*>(gr.post_stage.p), gr.U4, static_cast<const int*>(gr.lane_of.p),
                    static_cast<T*>(c.out0), static_cast<const T*>(bt->zsum.p), B, Dout, Tout, c.stats ? c.stats + 2 : nullptr);
            } else {
                normalize_post_kernel<T><<<ng, 256, 0, c.stream>>>(static_cast<T*>(c.out0),
                                                                 static_cast<const T*>(bt->zsum.p), B, Dout, Tout,
                                                                 c.stats ? c.stats + 2 : nullptr);
            }
            CK(cudaGetLastError());
        }
        const int* zlimit = nullptr;
        if (bt->ragged_cut) {  // frames past a tile's limit were never evaluated: keep them out of minimum(sums)
            TRY(bt->zlimit.ensure(B * sizeof(int)));
            CK(cudaMemcpyAsync(bt->zlimit.p, bt->h_zlimit.data(), B * sizeof(int), cudaMemcpyHostToDevice, c.stream));
            zlimit = static_cast<const int*>(bt->zlimit.p);
        }
        total_kernel<T><<<(B * 32 + 127) / 128, 128, 0, c.stream>>>(static_cast<const T*>(bt->zsum.p),
                                                             static_cast<const T*>(bt->lz.p),
                                                             static_cast<T*>(c.out1), B, N1, zlimit, c.stats,
                                                             c.expanded ? nullptr : d_seqlens, Tout, bt->prob ? 1 : 0);
        CK(cudaGetLastError());
        g_launches += 2;
    }
    if (mode == MODE_BEST) {
        if (!c.out1) return fail(MK_EINVAL, "null score buffer");
        // Back-trace descriptors: static per batch (graph pointers, strides of the α workspaces), so they are built and
        // uploaded once — per group at the first bestpath call, for the per-utterance kernel again when N̂ changes — from
        // host vectors the batch owns; the call enqueues kernels only and never synchronises the host.
        for (auto& gr : bt->groups) {
            const int n = int(gr.utts.size());
            if (gr.trace_tsize != sizeof(T)) {
                std::vector<TraceDesc<T>> descs(n);
                for (int k = 0; k < n; ++k) {
                    TraceDesc<T>& d = descs[k];
                    d.in_ptr = gr.g->d_in_ptr; d.in_arcs = static_cast<const Arc<T>*>(gr.g->d_in_arcs);
                    d.base = (long long)k; d.sn = (long long)gr.g->S * gr.U4; d.ss = gr.U4;
                    d.S = int(gr.g->S); d.b = gr.utts[k];
                }
                gr.h_trace.resize(descs.size() * sizeof(TraceDesc<T>));
                std::memcpy(gr.h_trace.data(), descs.data(), gr.h_trace.size());
                TRY(gr.trace.ensure(gr.h_trace.size()));
                CK(cudaMemcpyAsync(gr.trace.p, gr.h_trace.data(), gr.h_trace.size(), cudaMemcpyHostToDevice, c.stream));
                gr.trace_tsize = sizeof(T);
            }
            backtrace_kernel<T><<<(n * 32 + 127) / 128, 128, 0, c.stream>>>(
                static_cast<const TraceDesc<T>*>(gr.trace.p), n, static_cast<const T*>(gr.alpha.p), N1, Tout,
                d_seqlens, static_cast<int*>(c.out0), static_cast<T*>(c.out1));
            CK(cudaGetLastError());
            ++g_launches;
        }
        if (!bt->small.empty()) {
            const int n = int(bt->small.size());
            if (bt->trace_n1 != N1 || bt->trace_tsize != sizeof(T)) {
                std::vector<TraceDesc<T>> descs;
                descs.reserve(n);
                long long off = 0;
                for (int b : bt->small) {
                    mk_graph* g = bt->graphs[b];
                    TraceDesc<T> d;
                    d.in_ptr = g->d_in_ptr; d.in_arcs = static_cast<const Arc<T>*>(g->d_in_arcs);
                    d.base = off; d.sn = g->S; d.ss = 1; d.S = int(g->S); d.b = b;
                    off += (long long)N1 * g->S;
                    descs.push_back(d);
                }
                // (a previous call may still read the old table: the upload is ordered after it on the call's stream,
                // and dispatch() orders calls on different streams)
                bt->h_trace.resize(descs.size() * sizeof(TraceDesc<T>));
                std::memcpy(bt->h_trace.data(), descs.data(), bt->h_trace.size());
                TRY(bt->trace.ensure(bt->h_trace.size()));
                CK(cudaMemcpyAsync(bt->trace.p, bt->h_trace.data(), bt->h_trace.size(), cudaMemcpyHostToDevice, c.stream));
                bt->trace_n1 = N1; bt->trace_tsize = sizeof(T);
            }
            backtrace_kernel<T><<<(n * 32 + 127) / 128, 128, 0, c.stream>>>(
                static_cast<const TraceDesc<T>*>(bt->trace.p), n, static_cast<const T*>(bt->small_alpha.p), N1,
                Tout, d_seqlens, static_cast<int*>(c.out0), static_cast<T*>(c.out1));
            CK(cudaGetLastError());
            ++g_launches;
        }
    }
    return MK_OK;
}

static int dispatch(mk_batch* bt, Mode mode, const CallArgs& c) {
    if (!bt) return fail(MK_EINVAL, "null batch");
    if (mode == MODE_BEST && bt->semiring != MK_TROPICAL)
        return fail(MK_EINVAL, "bestpath needs TropicalSemiring graphs");
    if (mode == MODE_BEST && c.expanded) return fail(MK_EINVAL, "bestpath takes un-expanded emissions");
    {   // order this call after the previous one on the same batch if that one ran on another stream
        DeviceGuard guard(bt->device);
        if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", bt->device);
        if (!bt->ev_last) CK(cudaEventCreateWithFlags(&bt->ev_last, cudaEventDisableTiming));
        if (bt->has_last && bt->last_stream != c.stream) CK(cudaStreamWaitEvent(c.stream, bt->ev_last, 0));
    }
    int rc;
    if (bt->dtype == MK_F32)
        rc = bt->semiring == MK_LOG ? run<float, SR_LOG>(bt, mode, c) : run<float, SR_TROP>(bt, mode, c);
    else
        rc = bt->semiring == MK_LOG ? run<double, SR_LOG>(bt, mode, c) : run<double, SR_TROP>(bt, mode, c);
    {   // (also after a failed call: whatever it enqueued before failing still uses the workspaces)
        DeviceGuard guard(bt->device);
        if (guard.ok && cudaEventRecord(bt->ev_last, c.stream) == cudaSuccess) {
            bt->last_stream = c.stream;
            bt->has_last = true;
        }
    }
    return rc;
}

static int calibrate_graph(mk_graph* g, std::vector<double>& fwd_cycles, std::vector<double>& bwd_cycles) {
    // 128 utterances x 24 frames of all-zero log-likelihoods (the work of a frame does not depend on the values) through
    // the public entry points; the second call is the measured one (the first fills the caches and sizes the workspaces)
    const int B = 128, T = 24;
    const int64_t D = g->Dh - 1;
    const size_t ts = g->dtype == MK_F32 ? 4 : 8;
    std::vector<mk_graph*> graphs(B, g);
    mk_batch* bt = nullptr;
    int rc = mk_batch_create(&bt, graphs.data(), B);
    if (rc != MK_OK) return rc;
    void *ll = nullptr, *post = nullptr, *logz = nullptr;
    unsigned long long* cyc = nullptr;
    auto cleanup = [&]() {
        g->d_cta_cycles = nullptr;
        cudaFree(ll); cudaFree(post); cudaFree(logz); cudaFree(cyc);
        mk_batch_destroy(bt);
    };
    const size_t n_ll = size_t(B) * T * D;
    if (bt->groups.size() != 1 || cudaMalloc(&ll, n_ll * ts) != cudaSuccess || cudaMalloc(&post, n_ll * ts) != cudaSuccess ||
        cudaMalloc(&logz, B * ts) != cudaSuccess || cudaMalloc(&cyc, 2 * sizeof(unsigned long long) * g->n_sms) != cudaSuccess) {
        cleanup();
        cudaGetLastError();
        return MK_ENOMEM;
    }
    cudaMemset(ll, 0, n_ll * ts);
    cudaMemset(cyc, 0, 2 * sizeof(unsigned long long) * g->n_sms);
    for (int rep = 0; rep < 2 && rc == MK_OK; ++rep) {
        g->d_cta_cycles = rep == 1 ? cyc : nullptr;
        rc = mk_pdfposteriors(bt, ll, T * D, 1, D, D, T, 0, nullptr, post, logz, nullptr);
    }
    if (rc == MK_OK && cudaDeviceSynchronize() != cudaSuccess) rc = MK_ECUDA;
    std::vector<unsigned long long> h(2 * size_t(g->n_sms));
    if (rc == MK_OK && cudaMemcpy(h.data(), cyc, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = MK_ECUDA;
    cleanup();
    if (rc != MK_OK) return rc;
    fwd_cycles.assign(h.begin(), h.begin() + g->n_sms);
    bwd_cycles.assign(h.begin() + g->n_sms, h.end());
    return MK_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// operator level: mul! / sparse-vector broadcast of src/linalg.jl on caller-owned device arrays
// ------------------------------------------------------------------------------------------------
static int check_sr_dtype(int semiring, int dtype) {
    if (semiring != MK_LOG && semiring != MK_TROPICAL && semiring != MK_PROB)
        return fail(MK_EINVAL, "unknown semiring %d", semiring);
    if (dtype != MK_F32 && dtype != MK_F64) return fail(MK_EINVAL, "unknown dtype %d", dtype);
    return MK_OK;
}
static int check_csr(int64_t n_rows, int64_t n_cols, int64_t nnz, const void* rowptr, const void* colval,
                     const void* nzval, int index_base) {
    if (n_rows < 0 || n_cols < 0 || nnz < 0) return fail(MK_EINVAL, "negative matrix dimension");
    if (index_base != 0 && index_base != 1) return fail(MK_EINVAL, "index_base must be 0 or 1");
    if (n_rows >= (int64_t(1) << 31) || n_cols >= (int64_t(1) << 31) || nnz >= (int64_t(1) << 31))
        return fail(MK_ENOTSUP, "CSR dimensions exceed the Cint index type");
    if (!rowptr || (nnz > 0 && (!colval || !nzval))) return fail(MK_EINVAL, "null CSR array");
    return MK_OK;
}
static int device_sms() {  // (cached per device: the attribute query is not free and the operators are launch-bound for small graphs)
    static int cache[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        return 0;
    }
    if (dev >= 0 && dev < 64) cache[dev] = sms;
    return sms;
}

// Scratch for mk_spmv's long-row work list ({count, rows...}): one 256 KB buffer per (device, stream), created on first
// use and kept for the life of the process — stream order makes reuse by consecutive calls on that stream safe, and
// distinct streams never share one.  (The stream-ordered allocator was tried first: with the default pool's release
// threshold of 0 a call occasionally paid a real cudaMalloc, 0.4 -> 5 ms.)
constexpr int kSpmvWorklistCap = 1 << 16;
static int* spmv_worklist(cudaStream_t st, int cap) {
    struct Entry { int dev; cudaStream_t st; int* p; };
    static std::mutex mu;
    static std::vector<Entry> entries;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : entries)
        if (e.dev == dev && e.st == st) return e.p;
    if (entries.size() >= 256) return nullptr;
    int* p = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&p), size_t(cap + 1) * sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    entries.push_back({dev, st, p});
    return p;
}

template <typename T, int SR>
static int spmv_launch(int64_t n_rows, int64_t nnz, const int32_t* rowptr, const int32_t* colval, const void* nzval,
                       int base, const void* b, void* c, int sms, cudaStream_t st) {
    // lanes per row and arcs per lane and chunk from the mean row length: Ĉ-like matrices (1 arc per row) 4 x 2; the
    // path's graphs (~17 arcs per row) 2 x 12 — 16 rows per warp with up to 24 arcs each in flight (measured on the
    // block-diagonal T̂ᵀ of cfg 3: 8 lanes x 4 arcs 0.37 ms, 4 x 8 0.268, 2 x 16 0.261, 2 x 12 0.259, 1 x 32 0.34);
    // dense-ish rows (ω as a 1-row matrix) a warp.
    const double mean = double(nnz) / double(std::max<int64_t>(n_rows, 1));
    const T* v = static_cast<const T*>(nzval); const T* bb = static_cast<const T*>(b); T* cc = static_cast<T*>(c);
    const int lanes = mean <= 6 ? 4 : (mean <= 24 ? 2 : 32);
    const int per_lane = mean <= 6 ? 2 : (mean <= 24 ? 12 : 4);
    const int threads = 256;
    const int64_t want = (n_rows * lanes + threads - 1) / threads;
    const int blocks = int(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(sms) * 32)));  // grid-stride beyond
    // rows far longer than a lane group's chunk go to a work list and get a CTA each (stream-ordered scratch)
    const int long_row = lanes == 32 ? 2048 : 512;
    const int cap = int(std::min<int64_t>(n_rows, kSpmvWorklistCap));
    int* wl = nnz > long_row ? spmv_worklist(st, kSpmvWorklistCap) : nullptr;  // (none: the long rows are done in place)
    if (wl) CK(cudaMemsetAsync(wl, 0, sizeof(int), st));
    if (lanes == 4) spmv_kernel<T, SR, 4, 2><<<blocks, threads, 0, st>>>(n_rows, rowptr, colval, v, base, bb, cc, long_row, wl, cap);
    else if (lanes == 2) spmv_kernel<T, SR, 2, 12><<<blocks, threads, 0, st>>>(n_rows, rowptr, colval, v, base, bb, cc, long_row, wl, cap);
    else spmv_kernel<T, SR, 32, 4><<<blocks, threads, 0, st>>>(n_rows, rowptr, colval, v, base, bb, cc, long_row, wl, cap);
    CK(cudaGetLastError());
    ++g_launches;
    if (wl) {
        spmv_long_kernel<T, SR><<<sms, 512, 0, st>>>(wl, cap, rowptr, colval, v, base, bb, cc);
        CK(cudaGetLastError());
        ++g_launches;
    }
    return MK_OK;
}

template <typename T, int SR>
static int spmm_launch(int64_t n_rows, const int32_t* rowptr, const int32_t* colval, const void* nzval, int base,
                       const void* B, int64_t ldb, void* C, int64_t ldc, int64_t cols, int accumulate, cudaStream_t st) {
#ifndef MK_SPMM_CJ
#define MK_SPMM_CJ 8
#endif
    constexpr int CJ = MK_SPMM_CJ;  // columns per thread (tools/ab_variants.py; Ĉ·V̂ at cfg 3)
    // Large products go through the shared-memory window of B (linalg.cuh, spmm_staged_kernel): blocks of rb_rows rows x 4
    // columns, the column window of every row block found by a pre-pass on the same stream.  MK_SPMM_STAGED=0 keeps the
    // direct kernel; MK_SPMM_RB overrides the rows per block.
    static const int staged_on = [] { const char* e = getenv("MK_SPMM_STAGED"); return e ? atoi(e) : 1; }();
    static const int rb_env = [] { const char* e = getenv("MK_SPMM_RB"); return e ? atoi(e) : 0; }();
    if (staged_on && n_rows >= (1 << 17) && cols >= 4) {
        constexpr int SCJ = 4;
        constexpr int kWindowBytes = 110 * 1024;  // two CTAs per SM
        const int rb_rows = rb_env > 0 ? rb_env : 16384;
        const int64_t nblocks = (n_rows + rb_rows - 1) / rb_rows;
        int* win = 3 * nblocks <= kSpmvWorklistCap ? spmv_worklist(st, kSpmvWorklistCap) : nullptr;  // {lo, hi, one-arc rows} per block
        if (win) {
            spmm_window_kernel<<<unsigned(nblocks), 512, 0, st>>>(n_rows, rb_rows, rowptr, colval, base, win);
            CK(cudaGetLastError());
            dim3 sgrid(unsigned(std::min<int64_t>((cols + SCJ - 1) / SCJ, 65535)), unsigned(nblocks));
            // 512 threads x 8 rows in flight per thread, two CTAs per SM (measured at cfg 3: 512 x 4 0.71 ms, 256 x 16 0.74 ms,
            // 256 x 8 0.79 ms, this 0.67 ms; spmm_kernel 1.11 ms)
            constexpr int THREADS = 512, U = 8;
            CK(cudaFuncSetAttribute(spmm_staged_kernel<T, SR, SCJ, THREADS, U>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kWindowBytes));  // (per device; cheap)
            spmm_staged_kernel<T, SR, SCJ, THREADS, U><<<sgrid, THREADS, kWindowBytes, st>>>(
                n_rows, rb_rows, rowptr, colval, static_cast<const T*>(nzval), base, static_cast<const T*>(B), ldb,
                static_cast<T*>(C), ldc, cols, accumulate, win, kWindowBytes / int(SCJ * sizeof(T)));
            CK(cudaGetLastError());
            g_launches += 2;
            return MK_OK;
        }
    }
    dim3 grid(unsigned((n_rows + 255) / 256), unsigned(std::min<int64_t>((cols + CJ - 1) / CJ, 65535)));
    spmm_kernel<T, SR, CJ><<<grid, 256, 0, st>>>(n_rows, rowptr, colval, static_cast<const T*>(nzval), base,
                                                 static_cast<const T*>(B), ldb, static_cast<T*>(C), ldc, cols, accumulate);
    CK(cudaGetLastError());
    ++g_launches;
    return MK_OK;
}

template <typename T, int SR>
static int spvec_launch(int op, int64_t n, int64_t nnz, const int32_t* nzind, const void* nzval, int base,
                        const void* y, void* dest, int sms, cudaStream_t st) {
    const T zero = SR == LSR_PROB ? T(0) : -std::numeric_limits<T>::infinity();
    if (n > 0) {
        const int blocks = int(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, int64_t(sms) * 16)));
        fill_kernel<T><<<blocks, 256, 0, st>>>(static_cast<T*>(dest), n, zero);
        CK(cudaGetLastError());
        ++g_launches;
    }
    if (nnz > 0) {  // (src/linalg.jl:306: no launch for an empty sparse vector)
        const int blocks = int(std::max<int64_t>(1, std::min<int64_t>((nnz + 255) / 256, int64_t(sms) * 16)));
        if (op == 0)
            spvec_bcast_kernel<T, SR, 0><<<blocks, 256, 0, st>>>(nnz, nzind, static_cast<const T*>(nzval), base,
                                                                static_cast<const T*>(y), static_cast<T*>(dest));
        else
            spvec_bcast_kernel<T, SR, 1><<<blocks, 256, 0, st>>>(nnz, nzind, static_cast<const T*>(nzval), base,
                                                                static_cast<const T*>(y), static_cast<T*>(dest));
        CK(cudaGetLastError());
        ++g_launches;
    }
    return MK_OK;
}

#define MK_SR_DISPATCH(fn, ...)                                                              \
    do {                                                                                     \
        if (dtype == MK_F32) {                                                               \
            if (semiring == MK_LOG) return fn<float, LSR_LOG>(__VA_ARGS__);                  \
            if (semiring == MK_TROPICAL) return fn<float, LSR_TROPICAL>(__VA_ARGS__);        \
            return fn<float, LSR_PROB>(__VA_ARGS__);                                         \
        }                                                                                    \
        if (semiring == MK_LOG) return fn<double, LSR_LOG>(__VA_ARGS__);                     \
        if (semiring == MK_TROPICAL) return fn<double, LSR_TROPICAL>(__VA_ARGS__);           \
        return fn<double, LSR_PROB>(__VA_ARGS__);                                            \
    } while (0)


extern "C" {

int mk_abi_version(void) { return MK_ABI_VERSION; }
const char* mk_last_error(void) { return g_err.c_str(); }
int64_t mk_launch_count(int reset) {
    int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}
int mk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mk_graph_create(mk_graph** out, int semiring, int dtype, int64_t n_states_hat, int64_t nnz_hat,
                    const int64_t* colptr, const int64_t* rowval, const void* nzval, int64_t n_init,
                    const int64_t* init_idx, const void* init_w, const int32_t* state2pdf,
                    int64_t n_pdf_hat, int index_base, int device) {
    if (!out) return fail(MK_EINVAL, "null out");
    *out = nullptr;
    if (semiring != MK_LOG && semiring != MK_TROPICAL && semiring != MK_PROB) return fail(MK_EINVAL, "unknown semiring %d", semiring);
    if (dtype != MK_F32 && dtype != MK_F64) return fail(MK_EINVAL, "unknown dtype %d", dtype);
    if (index_base != 0 && index_base != 1) return fail(MK_EINVAL, "index_base must be 0 or 1");
    if (n_states_hat < 2 || n_states_hat > (int64_t(1) << 30)) return fail(MK_EINVAL, "bad n_states_hat");
    if (nnz_hat < 0 || nnz_hat > (int64_t(1) << 31) - 2) return fail(MK_EINVAL, "bad nnz_hat");
    if (n_pdf_hat < 2 || n_pdf_hat > (int64_t(1) << 30)) return fail(MK_EINVAL, "bad n_pdf_hat");
    if (!colptr || (nnz_hat && (!rowval || !nzval)) || !state2pdf || (n_init && (!init_idx || !init_w)))
        return fail(MK_EINVAL, "null array");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MK_ECUDA, "no CUDA device available (libmarkov_b200 has no CPU fallback)");
    }
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= ndev) return fail(MK_EINVAL, "device %d out of range", device);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", device);
    mk_graph* g = new (std::nothrow) mk_graph;
    if (!g) return fail(MK_ENOMEM, "out of host memory");
    // ProbSemiring: the graph is its LogSemiring image (x -> log x is a semiring isomorphism); emissions enter through
    // log, α / β / totals leave through exp, posteriors are the same numbers
    g->prob = semiring == MK_PROB;
    g->semiring = g->prob ? int(MK_LOG) : semiring; g->dtype = dtype; g->device = device;
    g->S = n_states_hat; g->nnz = nnz_hat; g->Dh = n_pdf_hat;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete g; return fail(MK_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    g->n_sms = prop.multiProcessorCount;
    int rc = dtype == MK_F32
                 ? build_graph<float>(g, colptr, rowval, nzval, n_init, init_idx, init_w, state2pdf, index_base)
                 : build_graph<double>(g, colptr, rowval, nzval, n_init, init_idx, init_w, state2pdf, index_base);
    if (rc != MK_OK) { delete g; return rc; }
    *out = g;
    return MK_OK;
}

int mk_graph_destroy(mk_graph* g) {
    if (!g) return MK_OK;
    DeviceGuard guard(g->device);
    delete g;
    return MK_OK;
}

int mk_graph_info(const mk_graph* g, int64_t* n_states_hat, int64_t* nnz_hat, int64_t* n_pdf_hat,
                  int* semiring, int* dtype) {
    if (!g) return fail(MK_EINVAL, "null graph");
    if (n_states_hat) *n_states_hat = g->S;
    if (nnz_hat) *nnz_hat = g->nnz;
    if (n_pdf_hat) *n_pdf_hat = g->Dh;
    if (semiring) *semiring = g->prob ? int(MK_PROB) : g->semiring;
    if (dtype) *dtype = g->dtype;
    return MK_OK;
}

int mk_batch_create(mk_batch** out, mk_graph* const* graphs, int64_t B) {
    if (!out) return fail(MK_EINVAL, "null out");
    *out = nullptr;
    if (!graphs || B <= 0 || B > (1 << 24)) return fail(MK_EINVAL, "bad batch size");
    for (int64_t b = 0; b < B; ++b)
        if (!graphs[b]) return fail(MK_EINVAL, "graphs[%lld] is null", (long long)b);
    mk_graph* g0 = graphs[0];
    for (int64_t b = 1; b < B; ++b) {
        mk_graph* g = graphs[b];
        if (g->semiring != g0->semiring || g->prob != g0->prob || g->dtype != g0->dtype || g->device != g0->device || g->Dh != g0->Dh)
            return fail(MK_EINVAL, "graphs[%lld] differs in semiring/dtype/device/n_pdf_hat", (long long)b);
    }
    DeviceGuard guard(g0->device);
    if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", g0->device);
    mk_batch* bt = new (std::nothrow) mk_batch;
    if (!bt) return fail(MK_ENOMEM, "out of host memory");
    bt->B = B; bt->semiring = g0->semiring; bt->prob = g0->prob; bt->dtype = g0->dtype; bt->device = g0->device;
    bt->n_sms = g0->n_sms; bt->Dh = g0->Dh;
    bt->graphs.assign(graphs, graphs + B);
    bt->off.resize(B);
    for (int64_t b = 0; b < B; ++b) { bt->off[b] = bt->total; bt->total += graphs[b]->S; }
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, bt->device);
    bt->max_smem_optin = size_t(optin);

    // Route every distinct graph: the replicated large graph goes to the shared-graph kernel,
    // graphs that fit shared memory to the per-utterance kernel.  MK_FORCE_KERNEL=shared|small
    // overrides (tests exercise both on the same inputs).
    const char* force = getenv("MK_FORCE_KERNEL");
    std::map<mk_graph*, std::vector<int>> by_graph;
    std::vector<mk_graph*> order;
    for (int64_t b = 0; b < B; ++b) {
        auto it = by_graph.find(graphs[b]);
        if (it == by_graph.end()) { order.push_back(graphs[b]); by_graph[graphs[b]] = {int(b)}; }
        else it->second.push_back(int(b));
    }
    for (mk_graph* g : order) {
        std::vector<int>& utts = by_graph[g];
        bool fits_small = small_smem_bytes(int(g->S), g->dtype) <= bt->max_smem_optin;
        bool shared = !fits_small || (utts.size() >= 8 && g->S >= 2048);
        if (force && !strcmp(force, "shared")) shared = true;
        if (force && !strcmp(force, "small") && fits_small) shared = false;
        if (!shared) {
            for (int b : utts) { bt->small.push_back(b); bt->small_smax = std::max(bt->small_smax, int(g->S)); }
            continue;
        }
        bt->groups.emplace_back();
        Group& gr = bt->groups.back();
        gr.g = g; gr.utts = utts;
        gr.U4 = int((utts.size() + 3) / 4 * 4);
        std::vector<int> ub(gr.U4, -1);
        std::vector<long long> uo(gr.U4, 0);
        bool vec4 = utts.size() % 4 == 0;
        for (size_t k = 0; k < utts.size(); ++k) {
            ub[k] = utts[k]; uo[k] = bt->off[utts[k]];
            if (k % 4 == 0) vec4 = vec4 && utts[k] % 4 == 0;
            else vec4 = vec4 && utts[k] == utts[k - 1] + 1;
        }
        gr.vec4 = vec4;
        int rc = upload(ub, (void**)&gr.d_utt_b);
        if (rc == MK_OK) rc = upload(uo, (void**)&gr.d_utt_off);
        if (rc != MK_OK) { delete bt; return rc; }
    }
    std::sort(bt->small.begin(), bt->small.end());
    cudaError_t e = cudaStreamCreateWithFlags(&bt->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete bt; return fail(MK_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    *out = bt;
    return MK_OK;
}

static void set_host_pending(mk_batch* b, bool pending);
int mk_batch_destroy(mk_batch* b) {
    if (!b) return MK_OK;
    DeviceGuard guard(b->device);
    set_host_pending(b, false);  // (the destructor drains the batch's streams)
    delete b;
    return MK_OK;
}

int mk_batch_info(const mk_batch* b, int64_t* B, int64_t* total_states_hat) {
    if (!b) return fail(MK_EINVAL, "null batch");
    if (B) *B = b->B;
    if (total_states_hat) *total_states_hat = b->total;
    return MK_OK;
}

int64_t mk_batch_workspace_bytes(const mk_batch* b) { return b ? int64_t(b->ws_bytes()) : 0; }

int mk_lfmmi_grad(int dtype, const void* num_post, const void* den_post, int64_t B, int64_t D, int64_t N,
                  const int32_t* seqlens_dev, double scale, void* grad, int64_t stride_b, int64_t stride_n,
                  int64_t stride_d, void* stream) {
    if (dtype != MK_F32 && dtype != MK_F64) return fail(MK_EINVAL, "unknown dtype %d", dtype);
    if (!num_post || !den_post || !grad) return fail(MK_EINVAL, "null buffer");
    if (B <= 0 || D <= 0 || N <= 0 || N > 65535 || (B + 31) / 32 > 65535)
        return fail(MK_EINVAL, "DimensionMismatch: bad (B, D, N) = (%lld, %lld, %lld)", (long long)B, (long long)D, (long long)N);
    dim3 grid(unsigned((D + 31) / 32), unsigned((B + 31) / 32), unsigned(N)), block(32, 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == MK_F32)
        lfmmi_grad_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(num_post), static_cast<const float*>(den_post),
                                                        int(B), int(D), int(N), seqlens_dev, float(scale),
                                                        static_cast<float*>(grad), stride_b, stride_n, stride_d);
    else
        lfmmi_grad_kernel<double><<<grid, block, 0, st>>>(static_cast<const double*>(num_post), static_cast<const double*>(den_post),
                                                         int(B), int(D), int(N), seqlens_dev, scale,
                                                         static_cast<double*>(grad), stride_b, stride_n, stride_d);
    CK(cudaGetLastError());
    ++g_launches;
    return MK_OK;
}

#ifdef MK_PROFILE_BARRIER
// debug: per-CTA cycle counters of grid_sync (work, CTA wait, grid wait); resets after reading
int mk_debug_barrier_profile(unsigned long long* out /* [148*4] */) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * 148 * 4));
    static unsigned long long zeros[148 * 4];
    CK(cudaMemcpyToSymbol(g_prof, zeros, sizeof zeros));
    CK(cudaMemcpyFromSymbol(out + 148 * 4, g_redo, sizeof(unsigned long long)));
    CK(cudaMemcpyToSymbol(g_redo, zeros, sizeof(unsigned long long)));
    return MK_OK;
}
#endif

int mk_measure_sfu_peak(int device, double* ops_per_second) {
    if (!ops_per_second) return fail(MK_EINVAL, "null result pointer");
    if (device < 0) CK(cudaGetDevice(&device));
    DeviceGuard guard(device);
    if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", device);
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 2, threads = 1024, iters = 4096;
    float* out = nullptr;
    CK(cudaMalloc(&out, sizeof(float) * blocks * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 0.f;
    for (int rep = 0; rep < 4; ++rep) {  // (the first launch warms the clocks up)
        CK(cudaEventRecord(e0, nullptr));
        sfu_peak_kernel<<<blocks, threads>>>(out, iters);
        CK(cudaEventRecord(e1, nullptr));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && (best == 0.f || ms < best)) best = ms;
        ++g_launches;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *ops_per_second = double(blocks) * threads * iters * 8 / (double(best) * 1e-3);
    return MK_OK;
}

int mk_batch_profile(mk_batch* b, int enable) {
    if (!b) return fail(MK_EINVAL, "null batch");
    DeviceGuard guard(b->device);
    if (enable && !b->ev0[0])
        for (int i = 0; i < mk_batch::kProfRing; ++i) {
            CK(cudaEventCreate(&b->ev0[i]));
            CK(cudaEventCreate(&b->ev1[i]));
        }
    b->profile = enable != 0;
    b->prof_n = 0;
    return MK_OK;
}

int mk_batch_kernel_ms(mk_batch* b, float* ms, int cap, int* n) {
    if (!b || !ms || !n) return fail(MK_EINVAL, "null argument");
    DeviceGuard guard(b->device);
    int have = std::min(b->prof_n, int(mk_batch::kProfRing));
    int cnt = std::min(have, cap);
    for (int k = 0; k < cnt; ++k) {  // the most recent `cnt` launches, oldest first
        int slot = (b->prof_n - cnt + k) % mk_batch::kProfRing;
        CK(cudaEventSynchronize(b->ev1[slot]));
        CK(cudaEventElapsedTime(&ms[k], b->ev0[slot], b->ev1[slot]));
    }
    *n = cnt;
    return MK_OK;
}

static CallArgs mkargs(const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D, int64_t T, int expanded,
                       const int32_t* seqlens, void* o0, void* o1, void* stream) {
    CallArgs c;
    c.ll = ll; c.sb = sb; c.sd = sd; c.sn = sn; c.D = D; c.T = T; c.expanded = expanded; c.seqlens = seqlens;
    c.out0 = o0; c.out1 = o1; c.stream = static_cast<cudaStream_t>(stream);
    return c;
}

int mk_alpha(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D, int64_t T,
             int expanded, const int32_t* seqlens, void* out_A, void* stream) {
    return dispatch(b, MODE_ALPHA, mkargs(ll, sb, sd, sn, D, T, expanded, seqlens, out_A, nullptr, stream));
}
int mk_beta(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D, int64_t T,
            int expanded, const int32_t* seqlens, void* out_B, void* stream) {
    return dispatch(b, MODE_BETA, mkargs(ll, sb, sd, sn, D, T, expanded, seqlens, out_B, nullptr, stream));
}
int mk_pdfposteriors(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D, int64_t T,
                     int expanded, const int32_t* seqlens, void* out_post, void* out_logz, void* stream) {
    return dispatch(b, MODE_POST, mkargs(ll, sb, sd, sn, D, T, expanded, seqlens, out_post, out_logz, stream));
}
int mk_pdfposteriors_stats(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D, int64_t T,
                           int expanded, const int32_t* seqlens, void* out_post, void* out_logz, double* out_stats,
                           void* stream) {
    CallArgs c = mkargs(ll, sb, sd, sn, D, T, expanded, seqlens, out_post, out_logz, stream);
    c.stats = out_stats;
    return dispatch(b, MODE_POST, c);
}
int mk_bestpath(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D, int64_t T,
                int expanded, const int32_t* seqlens, int32_t* out_path, void* out_score, void* stream) {
    return dispatch(b, MODE_BEST, mkargs(ll, sb, sd, sn, D, T, expanded, seqlens, out_path, out_score, stream));
}

static int64_t extent(int64_t B, int64_t D, int64_t T, int64_t sb, int64_t sd, int64_t sn) {
    return (B - 1) * sb + (D - 1) * sd + (T - 1) * sn + 1;
}

static int posteriors_host(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D, int64_t T,
                           int expanded, const int32_t* seqlens, void* out_post, void* out_logz, bool wait);

int mk_pdfposteriors_host(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D,
                          int64_t T, int expanded, const int32_t* seqlens, void* out_post, void* out_logz) {
    return posteriors_host(b, ll, sb, sd, sn, D, T, expanded, seqlens, out_post, out_logz, true);
}
int mk_pdfposteriors_host_begin(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D,
                                int64_t T, int expanded, const int32_t* seqlens, void* out_post, void* out_logz) {
    return posteriors_host(b, ll, sb, sd, sn, D, T, expanded, seqlens, out_post, out_logz, false);
}
int mk_batch_set_overlap(mk_batch* b, int enable) {
    if (!b) return fail(MK_EINVAL, "null batch");
    b->shared_threads = enable ? kSharedThreads / 2 : kSharedThreads;
    const char* e = getenv("MK_OVERLAP_THREADS");  // (tuning)
    if (enable && e && atoi(e) >= 64 && atoi(e) <= kSharedThreads && atoi(e) % 32 == 0) b->shared_threads = atoi(e);
    return MK_OK;
}
// Host-buffer calls begun and not yet waited for, per device: a call that starts with none in flight has no neighbour to
// hide its copies behind.
static std::atomic<int> g_host_inflight[64];
static void set_host_pending(mk_batch* b, bool pending) {
    if (b->host_pending == pending) return;
    b->host_pending = pending;
    if (b->device >= 0 && b->device < 64) g_host_inflight[b->device] += pending ? 1 : -1;
}

int mk_batch_wait(mk_batch* b) {
    if (!b) return fail(MK_EINVAL, "null batch");
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", b->device);
    if (b->own_stream) CK(cudaStreamSynchronize(b->own_stream));
    if (b->copy_stream) CK(cudaStreamSynchronize(b->copy_stream));
    set_host_pending(b, false);
    return MK_OK;
}

static int posteriors_host(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D, int64_t T,
                           int expanded, const int32_t* seqlens, void* out_post, void* out_logz, bool wait) {
    if (!b) return fail(MK_EINVAL, "null batch");
    if (!ll || !out_post || !out_logz) return fail(MK_EINVAL, "null buffer");
    if (sb < 0 || sd < 0 || sn < 0 || D <= 0 || T <= 0) return fail(MK_EINVAL, "bad strides/dims");
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", b->device);
    // one call per batch at a time: the staging buffers and workspaces belong to the batch (several batches overlap)
    if (b->host_pending) TRY(mk_batch_wait(b));
    const bool alone = b->device < 0 || b->device >= 64 || g_host_inflight[b->device].load() == 0;
    set_host_pending(b, !wait);
    const size_t ts = tsize(b->dtype);
    const int64_t Dout = expanded ? D - 1 : D, Tout = expanded ? T - 1 : T;
    const size_t in_bytes = size_t(extent(b->B, D, T, sb, sd, sn)) * ts;
    const size_t post_bytes = size_t(b->B) * Dout * Tout * ts;
    TRY(b->h_ll.ensure(in_bytes));
    TRY(b->h_post.ensure(post_bytes));
    TRY(b->h_logz.ensure(b->B * ts));
    cudaStream_t st = b->own_stream;
    // Pipeline: when the whole batch is one shared-graph group and the emissions are [b][t][d] rows, the call is
    // cut into frame segments; the copy stream brings slice k+1 in while the forward sweep runs slice k, and
    // takes the posteriors of segment k out while the backward sweep runs segment k-1.  MK_NO_PIPELINE=1 disables.
    // Segments: a lone blocking call wants many (the un-overlapped first copy in and last copy out shrink: 8 segments
    // 9.78 ms, 12 9.59 ms on the 128 x 150 x 3000 call).  Calls that overlap each other (begin / wait) want none: the
    // neighbour calls hide the copies, segments only add launches and event waits — per step at cfg 3, two batches in
    // flight: 1 segment 8.46 ms, 2 8.07, 4 8.16; three in flight: 1 segment 7.77, 2 8.2, 4 8.5 (tools/e2e_depth_probe.py).
    // The first call of such a sequence (none in flight when it begins) has no neighbour yet: 4 segments.
    int K = int(std::min<int64_t>(wait ? 12 : (alone ? 4 : 1), T / 8));
    if (getenv("MK_SEGMENTS") && atoi(getenv("MK_SEGMENTS")) >= 1) K = int(std::min<int64_t>(std::min<int64_t>(atoi(getenv("MK_SEGMENTS")), int64_t(mk_batch::kMaxSegments)), T / 8));  // (tuning)
    const bool pipe = b->groups.size() == 1 && b->small.empty() && !expanded && sd == 1 && sn == D && sb == T * D &&
                      K >= 2 && !(getenv("MK_NO_PIPELINE") && atoi(getenv("MK_NO_PIPELINE")));
    if (pipe) {
        if (!b->copy_stream) CK(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < K; ++k) {
            if (!b->ev_h2d[k]) CK(cudaEventCreateWithFlags(&b->ev_h2d[k], cudaEventDisableTiming));
            if (!b->ev_done[k]) CK(cudaEventCreateWithFlags(&b->ev_done[k], cudaEventDisableTiming));
        }
        cudaStream_t sc = b->copy_stream;
        Segments seg;
        for (int k = 0; k < K; ++k) seg.f.push_back(int(T * k / K));
        seg.f.push_back(int(T) + 1);  // the last segment also holds the phony frame
        const size_t row = size_t(D) * ts, pitch = size_t(T) * row;
        for (int k = 0; k < K; ++k) {
            const int t0 = seg.f[k], t1 = std::min<int>(seg.f[k + 1], int(T));
            CK(cudaMemcpy2DAsync(static_cast<char*>(b->h_ll.p) + size_t(t0) * row, pitch,
                                 static_cast<const char*>(ll) + size_t(t0) * row, pitch, size_t(t1 - t0) * row,
                                 size_t(b->B), cudaMemcpyHostToDevice, sc));
            CK(cudaEventRecord(b->ev_h2d[k], sc));
        }
        seg.before_fwd = [&](int k) -> int {
            CK(cudaStreamWaitEvent(st, b->ev_h2d[k], 0));
            return MK_OK;
        };
        const size_t frame_bytes = size_t(b->B) * Dout * ts;
        seg.after_bwd = [&](int k) -> int {
            const int t0 = seg.f[k], t1 = std::min<int>(seg.f[k + 1], int(T));
            CK(cudaEventRecord(b->ev_done[k], st));
            CK(cudaStreamWaitEvent(sc, b->ev_done[k], 0));
            CK(cudaMemcpyAsync(static_cast<char*>(out_post) + size_t(t0) * frame_bytes,
                               static_cast<char*>(b->h_post.p) + size_t(t0) * frame_bytes, size_t(t1 - t0) * frame_bytes,
                               cudaMemcpyDeviceToHost, sc));
            return MK_OK;
        };
        CallArgs c = mkargs(b->h_ll.p, sb, sd, sn, D, T, expanded, seqlens, b->h_post.p, b->h_logz.p, st);
        c.seg = &seg;
        int rc = dispatch(b, MODE_POST, c);
        if (rc == MK_OK) {
            cudaError_t e = cudaMemcpyAsync(out_logz, b->h_logz.p, b->B * ts, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) rc = fail(MK_ECUDA, "cudaMemcpyAsync(logz) failed: %s", cudaGetErrorString(e));
        }
        if (wait || rc != MK_OK) {  // (the lambdas only enqueue work during dispatch; a failed call drains what it enqueued)
            cudaStreamSynchronize(st);
            cudaStreamSynchronize(sc);
        }
        if (rc != MK_OK) return rc;
        CK(cudaGetLastError());
        return MK_OK;
    }
    CK(cudaMemcpyAsync(b->h_ll.p, ll, in_bytes, cudaMemcpyHostToDevice, st));
    TRY(dispatch(b, MODE_POST, mkargs(b->h_ll.p, sb, sd, sn, D, T, expanded, seqlens, b->h_post.p, b->h_logz.p, st)));
    CK(cudaMemcpyAsync(out_post, b->h_post.p, post_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_logz, b->h_logz.p, b->B * ts, cudaMemcpyDeviceToHost, st));
    if (wait) CK(cudaStreamSynchronize(st));
    return MK_OK;
}

int mk_bestpath_host(mk_batch* b, const void* ll, int64_t sb, int64_t sd, int64_t sn, int64_t D, int64_t T,
                     int expanded, const int32_t* seqlens, int32_t* out_path, void* out_score) {
    if (!b) return fail(MK_EINVAL, "null batch");
    if (!ll || !out_path || !out_score) return fail(MK_EINVAL, "null buffer");
    if (sb < 0 || sd < 0 || sn < 0 || D <= 0 || T <= 0) return fail(MK_EINVAL, "bad strides/dims");
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", b->device);
    const size_t ts = tsize(b->dtype);
    const size_t in_bytes = size_t(extent(b->B, D, T, sb, sd, sn)) * ts;
    const size_t path_bytes = size_t(b->B) * T * sizeof(int32_t);
    TRY(b->h_ll.ensure(in_bytes));
    TRY(b->h_path.ensure(path_bytes));
    TRY(b->h_logz.ensure(b->B * ts));
    cudaStream_t st = b->own_stream;
    CK(cudaMemcpyAsync(b->h_ll.p, ll, in_bytes, cudaMemcpyHostToDevice, st));
    TRY(dispatch(b, MODE_BEST, mkargs(b->h_ll.p, sb, sd, sn, D, T, expanded, seqlens, b->h_path.p, b->h_logz.p, st)));
    CK(cudaMemcpyAsync(out_path, b->h_path.p, path_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_score, b->h_logz.p, b->B * ts, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return MK_OK;
}

int mk_spmv(int semiring, int dtype, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* rowptr,
            const int32_t* colval, const void* nzval, int index_base, const void* b, int64_t len_b, void* c,
            int64_t len_c, void* stream) {
    TRY(check_sr_dtype(semiring, dtype));
    TRY(check_csr(n_rows, n_cols, nnz, rowptr, colval, nzval, index_base));
    if (n_cols != len_b) return fail(MK_EINVAL, "DimensionMismatch: size(A, 2) = %lld, length(b) = %lld",
                                     (long long)n_cols, (long long)len_b);
    if (n_rows != len_c) return fail(MK_EINVAL, "DimensionMismatch: size(A, 1) = %lld, length(c) = %lld",
                                     (long long)n_rows, (long long)len_c);
    const int sms = device_sms();
    if (!sms) return fail(MK_ECUDA, "no usable CUDA device (libmarkov_b200 has no CPU fallback)");
    if (nnz == 0 || n_rows == 0) return MK_OK;  // src/linalg.jl:169: an empty matrix launches nothing, c is untouched
    if (!b || !c) return fail(MK_EINVAL, "null vector");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MK_SR_DISPATCH(spmv_launch, n_rows, nnz, rowptr, colval, nzval, index_base, b, c, sms, st);
}

int mk_spmm(int semiring, int dtype, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* rowptr,
            const int32_t* colval, const void* nzval, int index_base, const void* B, int64_t rows_b, int64_t cols_b,
            int64_t ldb, void* C, int64_t rows_c, int64_t cols_c, int64_t ldc, int accumulate, void* stream) {
    TRY(check_sr_dtype(semiring, dtype));
    TRY(check_csr(n_rows, n_cols, nnz, rowptr, colval, nzval, index_base));
    if (n_cols != rows_b) return fail(MK_EINVAL, "DimensionMismatch: size(A, 2) = %lld, size(B, 1) = %lld",
                                      (long long)n_cols, (long long)rows_b);
    if (n_rows != rows_c) return fail(MK_EINVAL, "DimensionMismatch: size(A, 1) = %lld, size(C, 1) = %lld",
                                      (long long)n_rows, (long long)rows_c);
    if (cols_b != cols_c) return fail(MK_EINVAL, "DimensionMismatch: size(B, 2) = %lld, size(C, 2) = %lld",
                                      (long long)cols_b, (long long)cols_c);
    if (ldb < rows_b || ldc < rows_c) return fail(MK_EINVAL, "leading dimension smaller than the row count");
    const int sms = device_sms();
    if (!sms) return fail(MK_ECUDA, "no usable CUDA device (libmarkov_b200 has no CPU fallback)");
    if (n_rows == 0 || cols_c == 0) return MK_OK;
    // β = 1 with an empty A changes nothing; β = 0 still clears C (src/linalg.jl:246-249) — the kernel does both
    if (nnz == 0 && accumulate) return MK_OK;
    if (!C || (nnz > 0 && !B)) return fail(MK_EINVAL, "null matrix");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MK_SR_DISPATCH(spmm_launch, n_rows, rowptr, colval, nzval, index_base, B, ldb, C, ldc, cols_c, accumulate ? 1 : 0, st);
}

int mk_spvec_bcast(int semiring, int dtype, int op, int64_t n, int64_t nnz, const int32_t* nzind, const void* nzval,
                   int index_base, const void* y, int64_t len_y, void* dest, int64_t len_dest, void* stream) {
    TRY(check_sr_dtype(semiring, dtype));
    if (op != 0 && op != 1) return fail(MK_EINVAL, "op must be 0 (*) or 1 (/)");
    if (index_base != 0 && index_base != 1) return fail(MK_EINVAL, "index_base must be 0 or 1");
    if (n < 0 || nnz < 0 || nnz > n) return fail(MK_EINVAL, "bad sparse vector sizes");
    if (len_y != n || len_dest != n)
        return fail(MK_EINVAL, "DimensionMismatch: sparse vector of length %lld, y %lld, dest %lld", (long long)n,
                    (long long)len_y, (long long)len_dest);
    if (nnz > 0 && (!nzind || !nzval || !y)) return fail(MK_EINVAL, "null array");
    if (n > 0 && !dest) return fail(MK_EINVAL, "null destination");
    const int sms = device_sms();
    if (!sms) return fail(MK_ECUDA, "no usable CUDA device (libmarkov_b200 has no CPU fallback)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MK_SR_DISPATCH(spvec_launch, op, n, nnz, nzind, nzval, index_base, y, dest, sms, st);
}

// ---- data-parallel exchange: NCCL, bound at run time (no link-time dependency) -------------------------------
// The only cross-GPU step of the path is ONE sum all-reduce of the step statistics.  libnccl is looked
// up with dlopen: MK_NCCL_LIB, then the sonames a CUDA.jl / PyTorch process already has loaded.
namespace {
struct NcclId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("MK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
            api.why = dlerror();
        }
        if (!api.h) return;
        auto sym = [&](const char* n) { void* p = dlsym(api.h, n); if (!p) { api.why = std::string("missing symbol ") + n; } return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.AllReduce || !api.CommDestroy ||
            !api.GroupStart || !api.GroupEnd || !api.GetErrorString) {
            dlclose(api.h);
            api.h = nullptr;
        }
    });
    return &api;
}
int need_nccl(NcclApi** out) {
    NcclApi* a = nccl_api();
    if (!a->h) return fail(MK_ENOTSUP, "libnccl is not available (%s); set MK_NCCL_LIB", a->why.c_str());
    *out = a;
    return MK_OK;
}
#define NCK(api, call)                                                                                   \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != 0) return fail(MK_ECUDA, "%s failed: %s", #call, (api)->GetErrorString(r_));           \
    } while (0)
}  // namespace

struct mk_comm {
    NcclComm comm = nullptr;
    int device = 0, n_ranks = 1, rank = 0;
};

int mk_comm_unique_id(void* id128) {
    if (!id128) return fail(MK_EINVAL, "null id buffer");
    NcclApi* a;
    TRY(need_nccl(&a));
    NcclId id;
    NCK(a, a->GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return MK_OK;
}

int mk_comm_init_rank(mk_comm** out, int n_ranks, int rank, const void* id128, int device) {
    if (!out || !id128) return fail(MK_EINVAL, "null argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(MK_EINVAL, "rank %d outside [0, %d)", rank, n_ranks);
    NcclApi* a;
    TRY(need_nccl(&a));
    if (device < 0) CK(cudaGetDevice(&device));
    DeviceGuard guard(device);
    if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", device);
    NcclId id;
    std::memcpy(&id, id128, sizeof id);
    std::unique_ptr<mk_comm> c(new mk_comm);
    c->device = device; c->n_ranks = n_ranks; c->rank = rank;
    NCK(a, a->CommInitRank(&c->comm, n_ranks, id, rank));
    *out = c.release();
    return MK_OK;
}

int mk_comm_init(mk_comm** out, int n_gpus) {
    if (!out || n_gpus < 1) return fail(MK_EINVAL, "bad arguments");
    NcclApi* a;
    TRY(need_nccl(&a));
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < n_gpus)
        return fail(MK_ECUDA, "%d CUDA devices requested, %d usable", n_gpus, have);
    std::vector<NcclComm> comms(n_gpus);
    std::vector<int> devs(n_gpus);
    for (int i = 0; i < n_gpus; ++i) devs[i] = i;
    NCK(a, a->CommInitAll(comms.data(), n_gpus, devs.data()));
    for (int i = 0; i < n_gpus; ++i) {
        out[i] = new mk_comm;
        out[i]->comm = comms[i]; out[i]->device = i; out[i]->n_ranks = n_gpus; out[i]->rank = i;
    }
    return MK_OK;
}

int mk_allreduce_stats(mk_comm* c, double* stats, int64_t count, void* stream) {
    if (!c || !c->comm) return fail(MK_EINVAL, "null communicator");
    if (!stats || count < 0) return fail(MK_EINVAL, "bad buffer");
    if (count == 0) return MK_OK;
    NcclApi* a;
    TRY(need_nccl(&a));
    DeviceGuard guard(c->device);
    if (!guard.ok) return fail(MK_ECUDA, "cannot select CUDA device %d", c->device);
    NCK(a, a->AllReduce(stats, stats, size_t(count), /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->comm,
                        static_cast<cudaStream_t>(stream)));
    return MK_OK;
}

int mk_allreduce_stats_all(mk_comm* const* comms, double* const* stats, int n_gpus, int64_t count, void* const* streams) {
    if (!comms || !stats || n_gpus < 1) return fail(MK_EINVAL, "bad arguments");
    NcclApi* a;
    TRY(need_nccl(&a));
    NCK(a, a->GroupStart());
    int rc = MK_OK;
    for (int i = 0; i < n_gpus && rc == MK_OK; ++i)
        rc = mk_allreduce_stats(comms[i], stats[i], count, streams ? streams[i] : nullptr);
    NCK(a, a->GroupEnd());
    return rc;
}

int mk_comm_destroy(mk_comm* c) {
    if (!c) return MK_OK;
    NcclApi* a = nccl_api();
    if (a->h && c->comm) {
        DeviceGuard guard(c->device);
        a->CommDestroy(c->comm);
    }
    delete c;
    return MK_OK;
}

}  // extern "C"
