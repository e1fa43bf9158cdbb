// SPDX-License-Identifier: MIT
//
// oracle.cpp — CPU restatement of MarkovModels.jl's batched semiring inference
// path.  TEST INFRASTRUCTURE ONLY: nothing under markovmodels.jl_b200/ may
// include, link or call this file.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker.
//
// PARITY STATUS: "parity unpinned" against a *running* reference — the
// reference is Julia (no julia binary in this image) and its scalar arithmetic
// lives in the un-vendored Semirings.jl 0.5 (Project.toml:11,18).  The oracle is
// pinned only to the literal known-answer vectors the reference tree holds
// (examples/demo.ipynb cell 13, test/test_semirings.jl:3-7,
// test/test_linalg.jl:93-95, test/test_algorithms.jl:218-283) — see
// tests/test_oracle_golden.py.
//
// What is restated (reference file:line):
//   * LogSemiring / TropicalSemiring scalars   — Semirings.jl 0.5 (external);
//     use sites src/inference.jl:68-71,104-107,154-160, src/linalg.jl:207
//   * expand                                   — src/inference.jl:54-60
//   * Ĉ * V̂ (state <- pdf emission gather)     — src/inference.jl:150
//   * αrecursion                               — src/inference.jl:62-74
//   * βrecursion                               — src/inference.jl:99-110
//   * pdfposteriors                            — src/inference.jl:145-161
//   * CPU sparse mul! accumulation order       — Julia stdlib SparseArrays
//     (column-major push; per destination the terms arrive in ascending source
//     order and are folded with pairwise ⊕)
//   * bestpath (absent from the 0.10.0 tree) — αrecursion in the
//     tropical semiring + arg-max back-trace, first maximum in ascending
//     predecessor order (= Julia argmax).
//
// Conventions: all indices 0-based here (wrappers convert).  A graph is the
// extended matrix T̂ (src/fsm.jl:19-28) with the phony final state last, given
// in BOTH orientations:
//   in_*  : CSC of T̂  (column = destination state, entries = source states ascending)
//   out_* : CSR of T̂  (row = source state, entries = destination states ascending)
// Emissions are passed un-expanded, V[b] = D x T column-major (pdf fastest); the
// expand() padding is applied here.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

template <class T> constexpr T NEG_INF() { return -std::numeric_limits<T>::infinity(); }

// ---- Semirings.jl restatement -------------------------------------------------
template <class T> struct LogSR {
    static T zero() { return NEG_INF<T>(); }
    static T one() { return T(0); }
    // logaddexp; must return the other operand when one is -Inf
    // (test/test_semirings.jl:3-7 KAT)
    static T add(T x, T y) {
        if (x == NEG_INF<T>()) return y;
        if (y == NEG_INF<T>()) return x;
        T m = x > y ? x : y;
        T d = x > y ? y - x : x - y;  // -|x-y|
        return m + std::log1p(std::exp(d));
    }
    static T mul(T x, T y) { return x + y; }
    static T div(T x, T y) { return x - y; }
};
template <class T> struct TropSR {
    static T zero() { return NEG_INF<T>(); }
    static T one() { return T(0); }
    static T add(T x, T y) { return x > y ? x : y; }  // max-plus (assumption A-TROP)
    static T mul(T x, T y) { return x + y; }
    static T div(T x, T y) { return x - y; }
};

struct Graph {
    int64_t S;  // Ŝ, number of states incl. phony final
    const int64_t *in_ptr, *in_src;
    const void* in_w;
    const int64_t *out_ptr, *out_dst;
    const void* out_w;
    int64_t n_init;
    const int64_t* init_idx;
    const void* init_w;
    const int32_t* state2pdf;  // 0-based, phony state -> D
    int64_t Dhat;              // D+1
};

// expand (src/inference.jl:54-60) fused with the Ĉ*V̂ gather (:150):
// lhs[s, n] for n in 0..N (N+1 columns), column-major S x (T+1)
template <class T, class SR>
void state_emissions(const Graph& g, const T* V, int64_t D, int64_t Tn, int64_t ldv, int64_t L,
                     std::vector<T>& lhs) {
    const int64_t S = g.S, N1 = Tn + 1;
    lhs.assign(size_t(S) * N1, SR::zero());
    for (int64_t n = 0; n < N1; ++n) {
        T* col = lhs.data() + size_t(n) * S;
        for (int64_t s = 0; s < S; ++s) {
            int32_t d = g.state2pdf[s];
            T v;
            if (d < D) v = (n < L) ? V[size_t(n) * ldv + d] : SR::zero();
            else v = (n < L) ? SR::zero() : SR::one();
            col[s] = v;
        }
    }
}

// αrecursion (src/inference.jl:62-74).  A: S x N1 column-major.
template <class T, class SR>
void alpha_rec(const Graph& g, const T* lhs, int64_t N1, T* A) {
    const int64_t S = g.S;
    const T* w = static_cast<const T*>(g.in_w);
    const T* iw = static_cast<const T*>(g.init_w);
    // A[:,1] = α̂ .* lhs[:,1]   (sparse-vector broadcast: zero elsewhere)  :68
    for (int64_t s = 0; s < S; ++s) A[s] = SR::zero();
    for (int64_t k = 0; k < g.n_init; ++k) {
        int64_t s = g.init_idx[k];
        A[s] = SR::mul(iw[k], lhs[s]);
    }
    for (int64_t n = 1; n < N1; ++n) {
        const T* prev = A + size_t(n - 1) * S;
        T* cur = A + size_t(n) * S;
        const T* e = lhs + size_t(n) * S;
        for (int64_t j = 0; j < S; ++j) {
            // buffer = T̂ᵀ * A[:,n-1]  (:70) — ascending source order, pairwise ⊕
            T acc = SR::zero();
            for (int64_t a = g.in_ptr[j]; a < g.in_ptr[j + 1]; ++a)
                acc = SR::add(acc, SR::mul(w[a], prev[g.in_src[a]]));
            cur[j] = SR::mul(acc, e[j]);  // :71
        }
    }
}

// βrecursion (src/inference.jl:99-110).  Bm: S x N1 column-major.
template <class T, class SR>
void beta_rec(const Graph& g, const T* lhs, int64_t N1, T* Bm) {
    const int64_t S = g.S;
    const T* w = static_cast<const T*>(g.out_w);
    std::vector<T> buf(S);
    for (int64_t s = 0; s < S; ++s) Bm[size_t(N1 - 1) * S + s] = SR::one();  // :104
    for (int64_t n = N1 - 2; n >= 0; --n) {
        const T* nxt = Bm + size_t(n + 1) * S;
        const T* e = lhs + size_t(n + 1) * S;
        for (int64_t s = 0; s < S; ++s) buf[s] = SR::mul(nxt[s], e[s]);  // :106
        T* cur = Bm + size_t(n) * S;
        for (int64_t i = 0; i < S; ++i) {  // :107
            T acc = SR::zero();
            for (int64_t a = g.out_ptr[i]; a < g.out_ptr[i + 1]; ++a)
                acc = SR::add(acc, SR::mul(w[a], buf[g.out_dst[a]]));
            cur[i] = acc;
        }
    }
}

// pdfposteriors for one utterance (src/inference.jl:145-161).
// post: D x T column-major with leading stride given by (sd, sn): post[d*sd + n*sn].
template <class T, class SR>
T pdfpost_one(const Graph& g, const T* V, int64_t D, int64_t Tn, int64_t ldv, int64_t L, T* post,
              int64_t sd, int64_t sn) {
    const int64_t S = g.S, N1 = Tn + 1, Dh = D + 1;
    std::vector<T> lhs, A(size_t(S) * N1), Bm(size_t(S) * N1), P(Dh);
    state_emissions<T, SR>(g, V, D, Tn, ldv, L, lhs);
    alpha_rec<T, SR>(g, lhs.data(), N1, A.data());
    beta_rec<T, SR>(g, lhs.data(), N1, Bm.data());
    T ttl = std::numeric_limits<T>::infinity();
    for (int64_t n = 0; n < N1; ++n) {
        // AB = Ĉᵀ * (A .* B): per-pdf ⊕ in ascending state order (:154-155)
        for (int64_t d = 0; d < Dh; ++d) P[d] = SR::zero();
        for (int64_t s = 0; s < S; ++s) {
            T ab = SR::mul(A[size_t(n) * S + s], Bm[size_t(n) * S + s]);
            int32_t d = g.state2pdf[s];
            P[d] = SR::add(P[d], ab);
        }
        T sum = SR::zero();  // sums = sum(Ẑ, dims=2) (:157)
        for (int64_t d = 0; d < Dh; ++d) sum = SR::add(sum, P[d]);
        if (sum < ttl) ttl = sum;  // minimum(sums) (:159)
        if (n < Tn) {
            for (int64_t d = 0; d < D; ++d) {
                // Ẑ ./ sums then exp∘val (:158,160).  -Inf ⊘ -Inf (unreachable final)
                // is NaN in the reference; we adopt pdfposteriors3's convention
                // (src/inference.jl:198-200): zero.
                T q = (P[d] == NEG_INF<T>()) ? NEG_INF<T>() : SR::div(P[d], sum);
                post[d * sd + n * sn] = std::exp(q);
            }
        }
    }
    return ttl;
}

// bestpath for one utterance (ours).  path: T entries,
// 1-based state ids for frames < L, 0 for frames >= L.  Returns the path score.
template <class T>
T viterbi_one(const Graph& g, const T* V, int64_t D, int64_t Tn, int64_t ldv, int64_t L,
              int32_t* path) {
    typedef TropSR<T> SR;
    const int64_t S = g.S, N1 = Tn + 1;
    std::vector<T> lhs, A(size_t(S) * N1);
    state_emissions<T, SR>(g, V, D, Tn, ldv, L, lhs);
    alpha_rec<T, SR>(g, lhs.data(), N1, A.data());
    const T* w = static_cast<const T*>(g.in_w);
    for (int64_t t = 0; t < Tn; ++t) path[t] = 0;
    T score = A[size_t(N1 - 1) * S + (S - 1)];
    if (score == NEG_INF<T>()) return score;
    int64_t cur = S - 1;  // phony final at frame N1-1
    for (int64_t n = N1 - 1; n >= 1; --n) {
        // ψ = argmax_i T̂[i,cur] ⊗ A[i,n-1]; first maximum in ascending i
        T best = NEG_INF<T>();
        int64_t arg = -1;
        for (int64_t a = g.in_ptr[cur]; a < g.in_ptr[cur + 1]; ++a) {
            T v = w[a] + A[size_t(n - 1) * S + g.in_src[a]];
            if (v > best) { best = v; arg = g.in_src[a]; }
        }
        cur = arg;
        if (n - 1 < L) path[n - 1] = int32_t(cur + 1);
    }
    return score;
}

template <class T, class SR>
int run_alpha_beta(const Graph& g, const void* V, int64_t D, int64_t Tn, int64_t ldv, int64_t L,
                   void* A, void* Bm) {
    std::vector<T> lhs;
    state_emissions<T, SR>(g, static_cast<const T*>(V), D, Tn, ldv, L, lhs);
    if (A) alpha_rec<T, SR>(g, lhs.data(), Tn + 1, static_cast<T*>(A));
    if (Bm) beta_rec<T, SR>(g, lhs.data(), Tn + 1, static_cast<T*>(Bm));
    return 0;
}

}  // namespace

extern "C" {

// dtype: 0 = f32, 1 = f64.  semiring: 0 = Log, 1 = Tropical.
struct orc_graph {
    int64_t S, Dhat, n_init;
    const int64_t *in_ptr, *in_src;
    const void* in_w;
    const int64_t *out_ptr, *out_dst;
    const void* out_w;
    const int64_t* init_idx;
    const void* init_w;
    const int32_t* state2pdf;
};

static Graph to_graph(const orc_graph* og) {
    Graph g;
    g.S = og->S; g.Dhat = og->Dhat; g.n_init = og->n_init;
    g.in_ptr = og->in_ptr; g.in_src = og->in_src; g.in_w = og->in_w;
    g.out_ptr = og->out_ptr; g.out_dst = og->out_dst; g.out_w = og->out_w;
    g.init_idx = og->init_idx; g.init_w = og->init_w; g.state2pdf = og->state2pdf;
    return g;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Scalar ⊕ for the KATs (test/test_semirings.jl:3-7).
double orc_logaddexp_f64(double x, double y) { return LogSR<double>::add(x, y); }
float orc_logaddexp_f32(float x, float y) { return LogSR<float>::add(x, y); }

// Generic semiring CSR SpMV c = A ⊗ b (src/linalg.jl:163-184 contract; CPU order).
// semiring 2 = Prob (plain + and *), for the test_linalg.jl:93-95 KAT.
int orc_spmv_f64(int semiring, int64_t nrows, const int64_t* rowptr, const int64_t* colidx,
                 const double* w, const double* b, double* c) {
    for (int64_t r = 0; r < nrows; ++r) {
        double acc = semiring == 2 ? 0.0 : NEG_INF<double>();
        for (int64_t a = rowptr[r]; a < rowptr[r + 1]; ++a) {
            double x = b[colidx[a]];
            if (semiring == 0) acc = LogSR<double>::add(acc, w[a] + x);
            else if (semiring == 1) acc = TropSR<double>::add(acc, w[a] + x);
            else acc += w[a] * x;
        }
        c[r] = acc;
    }
    return 0;
}

// α / β for one utterance: A, Bm are Ŝ x (T+1) column-major (either may be NULL).
int orc_alpha_beta(int dtype, int semiring, const orc_graph* og, const void* V, int64_t D,
                   int64_t Tn, int64_t ldv, int64_t L, void* A, void* Bm) {
    Graph g = to_graph(og);
    if (g.Dhat != D + 1) return 22;
    if (dtype == 0 && semiring == 0) return run_alpha_beta<float, LogSR<float>>(g, V, D, Tn, ldv, L, A, Bm);
    if (dtype == 0 && semiring == 1) return run_alpha_beta<float, TropSR<float>>(g, V, D, Tn, ldv, L, A, Bm);
    if (dtype == 1 && semiring == 0) return run_alpha_beta<double, LogSR<double>>(g, V, D, Tn, ldv, L, A, Bm);
    if (dtype == 1 && semiring == 1) return run_alpha_beta<double, TropSR<double>>(g, V, D, Tn, ldv, L, A, Bm);
    return 22;
}

// pdfposteriors over a batch of B utterances.  graphs[b] may repeat.  V is
// B x (D x T column-major), utterance stride = ldv*T elements.  post is the
// reference's (B, D, N) column-major array (b fastest).  threads<=0: all cores;
// threads==1: the reference's single-threaded behaviour.
int orc_pdfposteriors(int dtype, int semiring, int64_t B, const orc_graph* const* graphs,
                      const void* V, int64_t D, int64_t Tn, int64_t ldv, const int32_t* seqlens,
                      void* post, void* logz, int threads) {
#ifdef _OPENMP
    int nt = threads <= 0 ? omp_get_max_threads() : threads;
#else
    int nt = 1; (void)threads;
#endif
    (void)nt;
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int64_t b = 0; b < B; ++b) {
        Graph g = to_graph(graphs[b]);
        if (g.Dhat != D + 1) { err = 22; continue; }
        int64_t L = seqlens ? seqlens[b] : Tn;
        size_t voff = size_t(b) * ldv * Tn;
        if (dtype == 0) {
            const float* v = static_cast<const float*>(V) + voff;
            float* p = static_cast<float*>(post) + b;
            float z = semiring == 0 ? pdfpost_one<float, LogSR<float>>(g, v, D, Tn, ldv, L, p, B, B * D)
                                    : pdfpost_one<float, TropSR<float>>(g, v, D, Tn, ldv, L, p, B, B * D);
            static_cast<float*>(logz)[b] = z;
        } else {
            const double* v = static_cast<const double*>(V) + voff;
            double* p = static_cast<double*>(post) + b;
            double z = semiring == 0 ? pdfpost_one<double, LogSR<double>>(g, v, D, Tn, ldv, L, p, B, B * D)
                                     : pdfpost_one<double, TropSR<double>>(g, v, D, Tn, ldv, L, p, B, B * D);
            static_cast<double*>(logz)[b] = z;
        }
    }
    return err;
}

// bestpath over a batch.  path: B x T int32 row-major (utterance-major), score: B.
int orc_bestpath(int dtype, int64_t B, const orc_graph* const* graphs, const void* V, int64_t D,
                 int64_t Tn, int64_t ldv, const int32_t* seqlens, int32_t* path, void* score,
                 int threads) {
#ifdef _OPENMP
    int nt = threads <= 0 ? omp_get_max_threads() : threads;
#else
    int nt = 1; (void)threads;
#endif
    (void)nt;
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int64_t b = 0; b < B; ++b) {
        Graph g = to_graph(graphs[b]);
        if (g.Dhat != D + 1) { err = 22; continue; }
        int64_t L = seqlens ? seqlens[b] : Tn;
        size_t voff = size_t(b) * ldv * Tn;
        if (dtype == 0)
            static_cast<float*>(score)[b] =
                viterbi_one<float>(g, static_cast<const float*>(V) + voff, D, Tn, ldv, L, path + b * Tn);
        else
            static_cast<double*>(score)[b] =
                viterbi_one<double>(g, static_cast<const double*>(V) + voff, D, Tn, ldv, L, path + b * Tn);
    }
    return err;
}

}  // extern "C"
