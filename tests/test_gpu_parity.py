# SPDX-License-Identifier: MIT
"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): log-likelihoods and pdf posteriors within 1e-4 relative
in Float32 and 1e-9 in Float64; best-path state sequences bit-exact."""
import os

import numpy as np
import pytest

from conftest import load_golden_fsm

pytestmark = pytest.mark.gpu

TOL = {np.float32: dict(rtol=1e-4, atol=1e-6), np.float64: dict(rtol=1e-9, atol=1e-12)}
# α/β are log-domain values of magnitude 10..1000: relative 1e-4 (f32) / 1e-9 (f64), with an
# absolute floor for entries near 0
TOL_STATE = {np.float32: dict(rtol=1e-4, atol=1e-4), np.float64: dict(rtol=1e-9, atol=1e-9)}


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch


class forced:
    """MK_FORCE_KERNEL is read when the batch is created."""

    def __init__(self, which):
        self.which = which

    def __enter__(self):
        self.old = os.environ.get("MK_FORCE_KERNEL")
        if self.which:
            os.environ["MK_FORCE_KERNEL"] = self.which
        else:
            os.environ.pop("MK_FORCE_KERNEL", None)

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("MK_FORCE_KERNEL", None)
        else:
            os.environ["MK_FORCE_KERNEL"] = self.old


def gpu_batch(mm, graphs, D, force=None):
    """graphs: list of (fsm, pdfids); identical objects are compiled once."""
    cache = {}
    cs = []
    for f, p in graphs:
        if id(f) not in cache:
            cache[id(f)] = mm.compile(f, mm.statemap(f, D, p))
        cs.append(cache[id(f)])
    with forced(force):
        return mm.batch(*cs)


def orc_graphs(orc, graphs, D):
    cache = {}
    out = []
    for f, p in graphs:
        if id(f) not in cache:
            cache[id(f)] = orc.OracleGraph(f, p, D)
        out.append(cache[id(f)])
    return out


def dev(torch, V_btd):
    """(B, T, D) host array -> the (B, D, T) strided device view a network output gives."""
    return torch.from_numpy(np.ascontiguousarray(V_btd)).cuda().permute(0, 2, 1)


def check_posteriors(mm, orc, graphs, D, V, lens, post, ttl, dtype, ttl_atol=0.0):
    """The parity bar.  Float64: 1e-9 against the oracle.  Float32: the oracle evaluated in Float32
    (the reference's arithmetic) carries its own rounding error — ulp(|α|) per ⊕, |α| growing with
    the frame index — so two correct Float32 evaluations cannot agree to 1e-4 on long sequences.
    The CUDA path must (i) match the exact answer (the Float64 oracle on the same Float32 inputs)
    within 1e-4 relative, and (ii) be no further from the Float32 oracle than twice that
    oracle's own distance from the exact answer."""
    post, ttl = np.asarray(post.cpu() if hasattr(post, "cpu") else post), np.asarray(ttl.cpu() if hasattr(ttl, "cpu") else ttl)
    opost, ottl = orc.pdfposteriors(orc_graphs(orc, graphs, D), V, lens)
    if dtype == np.float64:
        np.testing.assert_allclose(ttl, ottl, rtol=1e-9)
        np.testing.assert_allclose(post, opost, **TOL[dtype])
        return
    K64 = (mm.LogSemiring if graphs[0][0].K.code == 0 else mm.TropicalSemiring)[np.float64]
    cache = {}
    g64 = []
    for f, p in graphs:
        if id(f) not in cache:
            cache[id(f)] = (f.astype(K64), p)
        g64.append(cache[id(f)])
    xpost, xttl = orc.pdfposteriors(orc_graphs(orc, g64, D), V.astype(np.float64), lens)
    np.testing.assert_allclose(ttl, xttl, rtol=1e-4, atol=ttl_atol)
    np.testing.assert_allclose(ttl, ottl, rtol=1e-4, atol=ttl_atol)
    np.testing.assert_allclose(post, xpost, **TOL[dtype])                       # (i)
    ref_err = np.abs(opost - xpost).max()
    assert np.abs(post - opost).max() <= 2 * ref_err + 1e-6, (np.abs(post - opost).max(), ref_err)  # (ii)


def assert_states_close(got, want, dtype):
    got, want = np.asarray(got), np.asarray(want)
    inf_w = np.isneginf(want)
    np.testing.assert_array_equal(np.isneginf(got), inf_w)
    np.testing.assert_allclose(got[~inf_w], want[~inf_w], **TOL_STATE[dtype])


# ---------------------------------------------------------------------------------------------
# known answers straight from the reference tree
# ---------------------------------------------------------------------------------------------
GAMMA_DEMO = np.array([[1, .5, 1 / 6, 0, 0], [0, .5, 2 / 3, .5, 0], [0, 0, 1 / 6, .5, 1]])


@pytest.mark.parametrize("force", ["small", "shared"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_demo_posteriors(torch, mm, dtype, force):
    """examples/demo.ipynb cell 13."""
    K = mm.LogSemiring[dtype]
    g = mm.graphs.hmm3(K)
    b = gpu_batch(mm, [g], 3, force)
    post, ttl = mm.pdfposteriors(b, torch.zeros((1, 3, 5), dtype=torch.float32 if dtype == np.float32 else torch.float64,
                                             device="cuda"))
    np.testing.assert_allclose(post[0].cpu().numpy(), GAMMA_DEMO, atol=2e-6 if dtype == np.float32 else 1e-12)
    assert float(ttl[0]) == pytest.approx(np.log(6 / 32), rel=1e-5 if dtype == np.float32 else 1e-12)


@pytest.mark.parametrize("force", ["small", "shared"])
def test_ragged_batch_kat(torch, mm, orc, force):
    """test/test_algorithms.jl:218-248: lhs = ones(3,7,2), seqlengths = [5,7]."""
    K = mm.LogSemiring[np.float32]
    g = mm.graphs.hmm3(K)
    b = gpu_batch(mm, [g, g], 3, force)
    V = np.ones((2, 7, 3), np.float32)
    post, ttl = mm.pdfposteriors(b, dev(torch, V), seqlengths=[5, 7])
    post, ttl = post.cpu().numpy(), ttl.cpu().numpy()
    opost, ottl = orc.pdfposteriors(orc_graphs(orc, [g, g], 3), V, [5, 7])
    np.testing.assert_allclose(post, opost, **TOL[np.float32])
    np.testing.assert_allclose(ttl, ottl, rtol=1e-4)
    assert np.all(post[0, :, 5:] == 0.0)


@pytest.mark.parametrize("force", ["small", "shared"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bestpath_chain(torch, mm, dtype, force):
    """test/test_algorithms.jl:262-283 => "a b c d"."""
    K = mm.TropicalSemiring[dtype]
    g = mm.graphs.chain(K)
    b = gpu_batch(mm, [g, g], 4, force)
    path, score = mm.bestpath(b, dev(torch, np.ones((2, 6, 4), dtype)), seqlengths=[4, 3])
    np.testing.assert_array_equal(path[0].cpu().numpy(), [1, 2, 3, 4, 0, 0])
    assert float(score[0]) == 4.0
    assert float(score[1]) == -np.inf and np.all(path[1].cpu().numpy() == 0)


@pytest.mark.parametrize("dtype,rel", [(np.float32, 1e-4), (np.float64, 1e-9)])
@pytest.mark.parametrize("force", ["small", "shared"])
def test_den_fsm_wsj_logz(torch, mm, dtype, rel, force):
    """the reference's benchmark graph, lhs = ones(84, N)."""
    K = mm.LogSemiring[dtype]
    g = load_golden_fsm("den_fsm_wsj", K)
    b = gpu_batch(mm, [g] * 4, 84, force)
    for N, want in ((20, 12.578496038935), (100, 92.531024082122)):
        V = np.ones((4, N, 84), dtype)
        post, ttl = mm.pdfposteriors(b, dev(torch, V))
        np.testing.assert_allclose(ttl.cpu().numpy(), want, rtol=rel)
        np.testing.assert_allclose(post.cpu().numpy().sum(axis=1), 1.0, atol=1e-5 if dtype == np.float32 else 1e-10)


def test_num_fsm_wsj_unreachable(torch, mm):
    """Appendix B item 7: final state unreachable in < 166 frames -> posteriors 0, logZ -Inf."""
    K = mm.LogSemiring[np.float32]
    g = load_golden_fsm("num_fsm_wsj", K)
    b = gpu_batch(mm, [g, g], 84)
    post, ttl = mm.pdfposteriors(b, dev(torch, np.zeros((2, 166, 84), np.float32)), seqlengths=[165, 166])
    post, ttl = post.cpu().numpy(), ttl.cpu().numpy()
    assert ttl[0] == -np.inf and np.all(post[0] == 0.0) and not np.isnan(post).any()
    assert np.isfinite(ttl[1])
    np.testing.assert_allclose(post[1].sum(axis=0), 1.0, atol=1e-5)


# ---------------------------------------------------------------------------------------------
# seeded random inputs vs the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_numerator_batch_vs_oracle(torch, mm, orc, dtype):
    """cfg 2 shape (distinct per-utterance graphs, ragged lengths), reduced batch."""
    K = mm.LogSemiring[dtype]
    rng = np.random.default_rng(202)
    B, T, D = 12, 60, 200
    graphs = [mm.graphs.numerator(K, np.random.default_rng(202 + k), D, n_phones=int(rng.integers(8, 20))) for k in range(B)]
    V = (rng.standard_normal((B, T, D)) * 2).astype(dtype)
    lens = rng.integers(T // 2, T + 1, B).astype(np.int32)
    lens[0] = T
    b = gpu_batch(mm, graphs, D)
    post, ttl = mm.pdfposteriors(b, dev(torch, V), seqlengths=lens)
    assert bool(torch.isfinite(ttl).all())
    check_posteriors(mm, orc, graphs, D, V, lens, post, ttl, dtype)
    for k in range(B):
        assert np.all(post[k, :, int(lens[k]):].cpu().numpy() == 0.0)


@pytest.mark.parametrize("dtype,force", [(np.float32, None), (np.float32, "small"), (np.float64, None)])
def test_denominator_vs_oracle(torch, mm, orc, dtype, force):
    """cfg 3 shape, reduced: replicated denominator through the shared-graph kernel (U not a
    multiple of 4, ragged lengths) and, forced, through the per-utterance kernel."""
    K = mm.LogSemiring[dtype]
    rng = np.random.default_rng(303)
    B, T, D = 10, 40, 300
    g = mm.graphs.denominator(K, n_tokens=1500, n_pdf=D, seed=303)
    V = (rng.standard_normal((B, T, D)) * 2).astype(dtype)
    lens = rng.integers(T // 2, T + 1, B).astype(np.int32)
    b = gpu_batch(mm, [g] * B, D, force)
    post, ttl = mm.pdfposteriors(b, dev(torch, V), seqlengths=lens)
    check_posteriors(mm, orc, [g] * B, D, V, lens, post, ttl, dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_linear_copies_underflow_falls_back_exactly(torch, mm, orc, dtype):
    """The shared-graph kernel accumulates linear copies 2^(v + H) of the state vectors and redoes a row
    exactly (log domain) when its linear sum underflows.  A long left-to-right chain with sharply peaked
    emissions drives most rows of most frames far below 2^-126 of the frame maximum; posteriors and
    log-likelihoods must still match the Float64 oracle."""
    K = mm.LogSemiring[dtype]
    rng = np.random.default_rng(77)
    S, T, B = 90, 120, 8
    D = S
    f, pdf = mm.graphs.hmm3(K, S)  # left-to-right, self-loop + forward arc per state
    # every frame strongly prefers one pdf (60 nats above the rest): the forward mass collapses on a
    # narrow band, everything else is ~e^-60 per frame away from it
    V = (rng.standard_normal((B, T, D)) * 2).astype(dtype)
    for b in range(B):
        best = np.minimum(np.arange(T) * S // T + rng.integers(0, 3, T), S - 1)
        V[b, np.arange(T), best] += 60.0
    lens = rng.integers(T - 10, T + 1, B).astype(np.int32)
    b_ = gpu_batch(mm, [(f, pdf)] * B, D, "shared")
    post, ttl = mm.pdfposteriors(b_, dev(torch, V), seqlengths=lens)
    K64 = mm.LogSemiring[np.float64]
    xpost, xttl = orc.pdfposteriors(orc_graphs(orc, [(f.astype(K64), pdf)] * B, D), V.astype(np.float64), lens)
    assert np.all(np.isfinite(xttl))
    np.testing.assert_allclose(ttl.cpu().numpy(), xttl, rtol=1e-4 if dtype == np.float32 else 1e-9)
    # Float32: emissions of magnitude 100 (log2 units) are stored with ulp(100) = 7.6e-6 absolute resolution and
    # that rounding enters every frame; over 120 frames the posteriors keep 3e-4 relative (the Float32 oracle,
    # the reference's arithmetic, loses all significant digits on this input)
    tol = dict(rtol=3e-4, atol=1e-6) if dtype == np.float32 else TOL[dtype]
    np.testing.assert_allclose(post.cpu().numpy(), xpost, **tol)
    A = mm.αrecursion(b_, dev(torch, V), seqlengths=lens).cpu().numpy()
    og = orc_graphs(orc, [(f.astype(K64), pdf)] * B, D)
    for k in range(B):
        oA, _ = orc.alpha_beta(og[k], V[k].astype(np.float64), lens[k], want_beta=False)
        got = A[b_.offsets[k]:b_.offsets[k + 1]]
        if dtype == np.float64:
            assert_states_close(got, oA, dtype)
        else:
            # Float32: a state 4 000 log2 units below the frame maximum is stored with ulp(4000) = 2.4e-4
            # absolute resolution (the reference's un-normalised Float32 α has ulp(|α|) everywhere)
            np.testing.assert_array_equal(np.isneginf(got), np.isneginf(oA))
            fin = ~np.isneginf(oA)
            np.testing.assert_allclose(got[fin], oA[fin], rtol=1e-4, atol=5e-3)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_underflow_fallback_expands_merged_runs(torch, mm, orc, dtype):
    """The forward α store keeps no log2 row for the merged-run sources q_g; the exact fallback reaches a run's
    members through the run table.  Chain-topology denominator (every token is a run of two rows) with one pdf
    per frame 60 nats above the rest: most rows underflow in the linear domain in most frames."""
    K = mm.LogSemiring[dtype]
    rng = np.random.default_rng(91)
    B, T, D = 8, 50, 40
    g = mm.graphs.denominator(K, n_tokens=150, n_pdf=D, seed=17)
    V = (rng.standard_normal((B, T, D)) * 2).astype(dtype)
    for b in range(B):
        V[b, np.arange(T), rng.integers(0, D, T)] += 60.0
    lens = rng.integers(T - 10, T + 1, B).astype(np.int32)
    b_ = gpu_batch(mm, [g] * B, D, "shared")
    post, ttl = mm.pdfposteriors(b_, dev(torch, V), seqlengths=lens)
    K64 = mm.LogSemiring[np.float64]
    g64 = (g[0].astype(K64), g[1])
    xpost, xttl = orc.pdfposteriors(orc_graphs(orc, [g64] * B, D), V.astype(np.float64), lens)
    assert np.all(np.isfinite(xttl))
    np.testing.assert_allclose(ttl.cpu().numpy(), xttl, rtol=1e-4 if dtype == np.float32 else 1e-9)
    tol = dict(rtol=3e-4, atol=1e-6) if dtype == np.float32 else TOL[dtype]
    np.testing.assert_allclose(post.cpu().numpy(), xpost, **tol)
    A = mm.αrecursion(b_, dev(torch, V), seqlengths=lens).cpu().numpy()
    og = orc_graphs(orc, [g64] * B, D)
    for k in range(B):
        oA, _ = orc.alpha_beta(og[k], V[k].astype(np.float64), lens[k], want_beta=False)
        got = A[b_.offsets[k]:b_.offsets[k + 1]]
        np.testing.assert_array_equal(np.isneginf(got), np.isneginf(oA))
        fin = ~np.isneginf(oA)
        np.testing.assert_allclose(got[fin], oA[fin], rtol=1e-4 if dtype == np.float32 else 1e-9,
                                   atol=5e-3 if dtype == np.float32 else 1e-9)


def test_more_than_one_utterance_tile(torch, mm, orc):
    """A warp covers 128 utterances (32 lanes x 4); larger groups are worked tile by tile.  130 utterances:
    one full tile and one with two live lanes... of which half a lane is padding."""
    K = mm.LogSemiring[np.float32]
    rng = np.random.default_rng(130)
    B, T, D = 130, 16, 60
    g = mm.graphs.denominator(K, n_tokens=400, n_pdf=D, seed=13)
    V = (rng.standard_normal((B, T, D)) * 2).astype(np.float32)
    lens = rng.integers(T // 2, T + 1, B).astype(np.int32)
    b = gpu_batch(mm, [g] * B, D, "shared")
    post, ttl = mm.pdfposteriors(b, dev(torch, V), seqlengths=lens)
    check_posteriors(mm, orc, [g] * B, D, V, lens, post, ttl, np.float32)
    Kt = mm.TropicalSemiring[np.float32]
    gt = (g[0].astype(Kt), g[1])
    bt = gpu_batch(mm, [gt] * B, D, "shared")
    path, score = mm.bestpath(bt, dev(torch, V), seqlengths=lens)
    opath, oscore = orc.bestpath(orc_graphs(orc, [gt] * B, D), V, lens)
    np.testing.assert_array_equal(path.cpu().numpy(), opath)
    np.testing.assert_array_equal(score.cpu().numpy(), oscore)


def test_row_merging_is_transparent(torch, mm, orc):
    """The graph compiler merges runs of adjacent states with identical out-arc lists (the A/B pairs of
    the chain topology).  With and without merging (MK_NO_MERGE=1) the results agree with each other
    and with the oracle — posteriors, α, β and the tropical best path (bit-exact)."""
    rng = np.random.default_rng(77)
    B, T, D = 9, 35, 200
    V = (rng.standard_normal((B, T, D)) * 2).astype(np.float32)
    lens = rng.integers(T // 2, T + 1, B).astype(np.int32)
    out = {}
    for merge in (True, False):
        old = os.environ.get("MK_NO_MERGE")
        os.environ["MK_NO_MERGE"] = "0" if merge else "1"
        try:
            K = mm.LogSemiring[np.float32]
            g = mm.graphs.denominator(K, n_tokens=1300, n_pdf=D, seed=11)
            b = gpu_batch(mm, [g] * B, D, "shared")
            post, ttl = mm.pdfposteriors(b, dev(torch, V), seqlengths=lens)
            check_posteriors(mm, orc, [g] * B, D, V, lens, post, ttl, np.float32)
            A = mm.αrecursion(b, dev(torch, V), seqlengths=lens).cpu().numpy()
            Bm = mm.βrecursion(b, dev(torch, V), seqlengths=lens).cpu().numpy()
            Kt = mm.TropicalSemiring[np.float32]
            gt = (g[0].astype(Kt), g[1])
            bt = gpu_batch(mm, [gt] * B, D, "shared")
            path, score = mm.bestpath(bt, dev(torch, V), seqlengths=lens)
            opath, oscore = orc.bestpath(orc_graphs(orc, [gt] * B, D), V, lens)
            np.testing.assert_array_equal(path.cpu().numpy(), opath)
            np.testing.assert_array_equal(score.cpu().numpy(), oscore)
            out[merge] = (post.cpu().numpy(), ttl.cpu().numpy(), A, Bm)
        finally:
            if old is None:
                os.environ.pop("MK_NO_MERGE", None)
            else:
                os.environ["MK_NO_MERGE"] = old
    np.testing.assert_allclose(out[True][1], out[False][1], rtol=1e-6)
    np.testing.assert_allclose(out[True][0], out[False][0], rtol=1e-4, atol=1e-7)
    assert_states_close(out[True][2], out[False][2], np.float32)
    assert_states_close(out[True][3], out[False][3], np.float32)
    og = orc_graphs(orc, [g], D)[0]
    oA, oB = orc.alpha_beta(og, V[0], lens[0])
    assert_states_close(out[True][2][:g[0].nstates_hat], oA, np.float32)
    assert_states_close(out[True][3][:g[0].nstates_hat], oB, np.float32)


def test_mixed_batch_vs_oracle(torch, mm, orc):
    """One batch holding the replicated denominator (shared-graph kernel, interleaved utterance
    indices) and distinct small graphs (per-utterance kernel)."""
    K = mm.LogSemiring[np.float32]
    rng = np.random.default_rng(11)
    T, D = 30, 120
    den = mm.graphs.denominator(K, n_tokens=1100, n_pdf=D, seed=5)
    nums = [mm.graphs.numerator(K, np.random.default_rng(50 + k), D, n_phones=6) for k in range(5)]
    graphs = []
    for k in range(9):
        graphs.append(den)
        if k < 5:
            graphs.append(nums[k])
    B = len(graphs)
    V = (rng.standard_normal((B, T, D)) * 2).astype(np.float32)
    lens = rng.integers(T // 2, T + 1, B).astype(np.int32)
    b = gpu_batch(mm, graphs, D)
    post, ttl = mm.pdfposteriors(b, dev(torch, V), seqlengths=lens)
    check_posteriors(mm, orc, graphs, D, V, lens, post, ttl, np.float32)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("semiring", ["log", "tropical"])
@pytest.mark.parametrize("force", ["small", "shared"])
def test_alpha_beta_recursions(torch, mm, orc, dtype, semiring, force):
    """αrecursion / βrecursion outputs in the reference's (ΣŜ x N̂) layout."""
    K = (mm.LogSemiring if semiring == "log" else mm.TropicalSemiring)[dtype]
    rng = np.random.default_rng(21)
    T, D = 25, 60
    g1 = mm.graphs.denominator(K, n_tokens=150, n_pdf=D, seed=9)
    g2 = mm.graphs.numerator(K, rng, D, n_phones=7)
    graphs = [g1, g2, g1, g1, g2]
    B = len(graphs)
    V = (rng.standard_normal((B, T, D)) * 2).astype(dtype)
    lens = np.array([T, T - 3, T - 10, T, 12], np.int32)
    b = gpu_batch(mm, graphs, D, force)
    A = mm.αrecursion(b, dev(torch, V), seqlengths=lens).cpu().numpy()
    Bm = mm.βrecursion(b, dev(torch, V), seqlengths=lens).cpu().numpy()
    assert A.shape == (b.total_states_hat, T + 1) == Bm.shape
    og = orc_graphs(orc, graphs, D)
    for k in range(B):
        oA, oB = orc.alpha_beta(og[k], V[k], lens[k])
        lo, hi = b.offsets[k], b.offsets[k + 1]
        assert_states_close(A[lo:hi], oA, dtype)
        assert_states_close(Bm[lo:hi], oB, dtype)


def test_tropical_alpha_bit_exact(torch, mm, orc):
    """max / + are exact in floating point: tropical α must match the oracle bit for bit."""
    K = mm.TropicalSemiring[np.float32]
    rng = np.random.default_rng(3)
    T, D = 30, 80
    g = mm.graphs.denominator(K, n_tokens=400, n_pdf=D, seed=4)
    V = (rng.standard_normal((3, T, D)) * 2).astype(np.float32)
    for force in ("small", "shared"):
        b = gpu_batch(mm, [g] * 3, D, force)
        A = mm.αrecursion(b, dev(torch, V)).cpu().numpy()
        for k in range(3):
            oA, _ = orc.alpha_beta(orc_graphs(orc, [g], D)[0], V[k], want_beta=False)
            np.testing.assert_array_equal(A[b.offsets[k]:b.offsets[k + 1]], oA)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("force", ["small", "shared"])
def test_bestpath_vs_oracle(torch, mm, orc, dtype, force):
    K = mm.TropicalSemiring[dtype]
    rng = np.random.default_rng(505)
    T, D = 50, 100
    den = mm.graphs.denominator(K, n_tokens=500, n_pdf=D, seed=6)
    loop = mm.graphs.phone_loop(K, n_phones=11)
    loop = (loop[0], loop[1] % D)
    graphs = [den] * 6 + [loop] * 3
    B = len(graphs)
    V = (rng.standard_normal((B, T, D)) * 2).astype(dtype)
    lens = rng.integers(T // 2, T + 1, B).astype(np.int32)
    b = gpu_batch(mm, graphs, D, force)
    path, score = mm.bestpath(b, dev(torch, V), seqlengths=lens)
    opath, oscore = orc.bestpath(orc_graphs(orc, graphs, D), V, lens)
    np.testing.assert_array_equal(path.cpu().numpy(), opath)
    np.testing.assert_array_equal(score.cpu().numpy(), oscore)
    for k in range(B):
        assert np.all(opath[k, :lens[k]] > 0) and np.all(opath[k, lens[k]:] == 0)


def test_bestpath_calls_queue_without_host_sync(torch, mm, orc):
    """mk_bestpath is asynchronous (include/markov_b200.h): several calls with different inputs and frame counts are
    enqueued back to back on a side stream — nothing waits for the GPU in between — and every one of them is right."""
    K = mm.TropicalSemiring[np.float32]
    rng = np.random.default_rng(77)
    D = 80
    den = mm.graphs.denominator(K, n_tokens=400, n_pdf=D, seed=9)
    loop = mm.graphs.phone_loop(K, n_phones=9)
    loop = (loop[0], loop[1] % D)
    graphs = [den] * 8 + [loop] * 2     # shared-graph group + per-utterance kernel (whose trace table depends on N̂)
    B = len(graphs)
    b = gpu_batch(mm, graphs, D)
    og = orc_graphs(orc, graphs, D)
    cases = [(rng.standard_normal((B, T, D)) * 2).astype(np.float32) for T in (30, 45, 30)]
    st = torch.cuda.Stream()
    outs = []
    with torch.cuda.stream(st):
        devs = [dev(torch, V) for V in cases]
        for Vd in devs:
            outs.append(mm.bestpath(b, Vd))
    st.synchronize()
    for V, (path, score) in zip(cases, outs):
        opath, oscore = orc.bestpath(og, V)
        np.testing.assert_array_equal(path.cpu().numpy(), opath)
        np.testing.assert_array_equal(score.cpu().numpy(), oscore)


def test_bestpath_ties_take_smallest_predecessor(torch, mm, orc):
    """Two exactly tied branches 1->2->4 and 1->3->4: the path goes through state 2."""
    K = mm.TropicalSemiring[np.float32]
    fsm = mm.FSM.from_arrays(K, 4, [0, 0, 1, 2], [1, 2, 3, 3], [0.0] * 4, [0], [0.0], [3], [0.0])
    pdf = np.zeros(4, np.int64)
    V = np.zeros((1, 3, 1), np.float32)
    for force in ("small", "shared"):
        b = gpu_batch(mm, [(fsm, pdf)], 1, force)
        path, score = mm.bestpath(b, dev(torch, V))
        np.testing.assert_array_equal(path.cpu().numpy(), [[1, 2, 4]])
    opath, _ = orc.bestpath(orc_graphs(orc, [(fsm, pdf)], 1), V)
    np.testing.assert_array_equal(opath, [[1, 2, 4]])


def test_expanded_inputs_and_reference_call_shape(torch, mm, orc):
    """The reference's call: pdfposteriors(rawunion(fsms...), V̂s, Ĉs) with expanded D̂ x N̂
    matrices (examples/test_cuda.jl:124-128) gives the same result as the un-expanded call."""
    K = mm.LogSemiring[np.float32]
    rng = np.random.default_rng(8)
    T, D = 20, 40
    fsms = [mm.graphs.numerator(K, rng, D, n_phones=5) for _ in range(3)]
    lens = [20, 14, 17]
    V = (rng.standard_normal((3, T, D)) * 2).astype(np.float32)
    Vhats = [torch.from_numpy(mm.expand(V[k].T, lens[k])).cuda() for k in range(3)]
    Cs = [mm.statemap(f, D, p) for f, p in fsms]
    post, ttl = mm.pdfposteriors(mm.rawunion(*[f for f, _ in fsms]), Vhats, Cs)
    assert tuple(post.shape) == (3, D, T)
    check_posteriors(mm, orc, fsms, D, V, lens, post, ttl, np.float32)


def test_host_buffer_entry_point(torch, mm, orc):
    """numpy in -> numpy out through mk_pdfposteriors_host / mk_bestpath_host."""
    K = mm.LogSemiring[np.float32]
    rng = np.random.default_rng(12)
    B, T, D = 9, 30, 100
    g = mm.graphs.denominator(K, n_tokens=1100, n_pdf=D, seed=2)
    V = (rng.standard_normal((B, T, D)) * 2).astype(np.float32)
    b = gpu_batch(mm, [g] * B, D)
    post, ttl = mm.pdfposteriors(b, V.transpose(0, 2, 1))
    assert isinstance(post, np.ndarray)
    check_posteriors(mm, orc, [g] * B, D, V, None, post, ttl, np.float32)
    Kt = mm.TropicalSemiring[np.float32]
    gt = mm.graphs.denominator(Kt, n_tokens=1100, n_pdf=D, seed=2)
    path, score = mm.bestpath(gpu_batch(mm, [gt] * B, D), V.transpose(0, 2, 1))
    opath, oscore = orc.bestpath(orc_graphs(orc, [gt] * B, D), V)
    np.testing.assert_array_equal(path, opath)
    np.testing.assert_array_equal(score, oscore)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_host_pipeline_matches_single_launch(torch, mm, orc, dtype):
    """mk_pdfposteriors_host cuts long calls into frame segments (copies overlap the sweeps, the running
    per-utterance scalars are carried from launch to launch); MK_NO_PIPELINE=1 runs the same call in one piece.
    Ragged lengths, a segment boundary inside the padding of some utterances."""
    K = mm.LogSemiring[dtype]
    rng = np.random.default_rng(66)
    B, T, D = 8, 70, 120
    g = mm.graphs.denominator(K, n_tokens=700, n_pdf=D, seed=6)
    V = (rng.standard_normal((B, T, D)) * 2).astype(dtype)
    lens = rng.integers(20, T + 1, B).astype(np.int32)
    lens[0] = T
    b = gpu_batch(mm, [g] * B, D, "shared")
    out = {}
    for flag in ("0", "1"):
        os.environ["MK_NO_PIPELINE"] = flag
        try:
            post, ttl = mm.pdfposteriors(b, V.transpose(0, 2, 1), seqlengths=lens)
            out[flag] = (np.array(post), np.array(ttl))
        finally:
            os.environ.pop("MK_NO_PIPELINE", None)
    rt = 1e-5 if dtype == np.float32 else 1e-12
    np.testing.assert_allclose(out["0"][1], out["1"][1], rtol=rt)
    np.testing.assert_allclose(out["0"][0], out["1"][0], rtol=rt, atol=1e-9)
    check_posteriors(mm, orc, [g] * B, D, V, lens, out["0"][0], out["0"][1], dtype)
    dpost, dttl = mm.pdfposteriors(b, dev(torch, V), seqlengths=lens)  # device path, one launch pair
    np.testing.assert_allclose(out["0"][1], dttl.cpu().numpy(), rtol=rt)
    np.testing.assert_allclose(out["0"][0], dpost.cpu().numpy(), rtol=rt, atol=1e-9)


def test_host_calls_overlap_across_batches(torch, mm, orc):
    """mk_pdfposteriors_host_begin / mk_batch_wait: two batch objects (one graph) keep a call each in flight, on pinned
    host buffers; same results as the blocking call, and a second begin on a busy batch waits for the first."""
    K = mm.LogSemiring[np.float32]
    rng = np.random.default_rng(91)
    B, T, D = 8, 64, 100
    g = mm.graphs.denominator(K, n_tokens=600, n_pdf=D, seed=8)
    bs = [gpu_batch(mm, [g] * B, D, "shared") for _ in range(2)]
    Vs = [torch.from_numpy((rng.standard_normal((B, T, D)) * 2).astype(np.float32)).pin_memory() for _ in range(3)]
    outs = [(torch.empty((T, D, B)).pin_memory(), torch.empty((B,)).pin_memory()) for _ in range(3)]
    with pytest.raises(TypeError):
        mm.pdfposteriors(bs[0], Vs[0].numpy().transpose(0, 2, 1), wait=False)   # needs out=
    for k in range(3):   # k = 2 reuses batch 0 while its first call may still run: the library waits for it
        mm.pdfposteriors(bs[k & 1], Vs[k].numpy().transpose(0, 2, 1), out=(outs[k][0].numpy(), outs[k][1].numpy()), wait=False)
    bs[0].wait(); bs[1].wait()
    for k in range(3):
        post, ttl = mm.pdfposteriors(bs[0], Vs[k].numpy().transpose(0, 2, 1))   # blocking call, library-owned outputs
        # (the per-pdf sums are float atomics: their order, hence the last bits, differ from run to run)
        np.testing.assert_allclose(outs[k][1].numpy(), np.asarray(ttl), rtol=1e-6)
        np.testing.assert_allclose(outs[k][0].numpy().transpose(2, 1, 0), np.asarray(post), rtol=1e-5, atol=1e-9)
    check_posteriors(mm, orc, [g] * B, D, Vs[2].numpy(), np.full(B, T, np.int32), outs[2][0].numpy().transpose(2, 1, 0),
                     outs[2][1].numpy(), np.float32)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("force", ["small", "shared"])
def test_prob_semiring_graphs_run_as_their_log_image(torch, mm, orc, dtype, force):
    """FSM{ProbSemiring} through the fused recursions (the reference instantiates Prob for the same mul!, test/test_linalg.jl:89):
    probabilities in (weights, emissions), probabilities out (α, β, totals); posteriors equal the LogSemiring run on the
    logarithms, which the oracle checks."""
    Kp, Kl = mm.ProbSemiring[dtype], mm.LogSemiring[dtype]
    rng = np.random.default_rng(12)
    B, T, D = 5, 12, 30
    gl = mm.graphs.denominator(Kl, n_tokens=120, n_pdf=D, seed=3)
    fl, pdf = gl
    fp = mm.FSM(Kp, fl.nstates_hat, fl.init_idx, np.exp(fl.init_w.astype(np.float64)).astype(dtype), fl.colptr, fl.rowval,
                np.exp(fl.nzval.astype(np.float64)).astype(dtype))
    logV = (rng.standard_normal((B, T, D)) * 1.5).astype(dtype)
    lens = rng.integers(T // 2, T + 1, B).astype(np.int32)
    bl = gpu_batch(mm, [gl] * B, D, force)
    bp = gpu_batch(mm, [(fp, pdf)] * B, D, force)
    post_l, ttl_l = mm.pdfposteriors(bl, dev(torch, logV), seqlengths=lens)
    post_p, ttl_p = mm.pdfposteriors(bp, dev(torch, np.exp(logV)), seqlengths=lens)
    rt = 2e-4 if dtype == np.float32 else 1e-9
    np.testing.assert_allclose(post_p.cpu().numpy(), post_l.cpu().numpy(), rtol=rt, atol=1e-7)
    np.testing.assert_allclose(ttl_p.cpu().numpy(), np.exp(ttl_l.cpu().numpy().astype(np.float64)), rtol=rt)
    check_posteriors(mm, orc, [gl] * B, D, logV, lens, post_p, np.log(ttl_p.cpu().numpy().astype(np.float64)).astype(dtype), dtype)
    A_l = mm.αrecursion(bl, dev(torch, logV), seqlengths=lens).cpu().numpy().astype(np.float64)
    A_p = mm.αrecursion(bp, dev(torch, np.exp(logV)), seqlengths=lens).cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(A_p, np.exp(A_l), rtol=rt, atol=0)
    B_l = mm.βrecursion(bl, dev(torch, logV), seqlengths=lens).cpu().numpy().astype(np.float64)
    B_p = mm.βrecursion(bp, dev(torch, np.exp(logV)), seqlengths=lens).cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(B_p, np.exp(B_l), rtol=rt, atol=0)


def test_two_batches_share_the_sms(torch, mm, orc):
    """mk_batch_set_overlap: the shared-graph sweeps of two batches run with half the threads per CTA, co-resident on every
    SM, on two streams; same posteriors as the default launch."""
    K = mm.LogSemiring[np.float32]
    rng = np.random.default_rng(92)
    B, T, D = 12, 40, 90
    g = mm.graphs.denominator(K, n_tokens=900, n_pdf=D, seed=10)
    bs = [gpu_batch(mm, [g] * B, D, "shared") for _ in range(2)]
    Vs = [(rng.standard_normal((B, T, D)) * 2).astype(np.float32) for _ in range(2)]
    lens = [rng.integers(T // 2, T + 1, B).astype(np.int32) for _ in range(2)]
    for b in bs:
        b.set_overlap(True)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for k in range(2):
        with torch.cuda.stream(streams[k]):
            outs.append(mm.pdfposteriors(bs[k], dev(torch, Vs[k]), seqlengths=lens[k]))
    torch.cuda.synchronize()
    for k in range(2):
        check_posteriors(mm, orc, [g] * B, D, Vs[k], lens[k], outs[k][0], outs[k][1], np.float32)
    bs[0].set_overlap(False)
    post, ttl = mm.pdfposteriors(bs[0], dev(torch, Vs[0]), seqlengths=lens[0])
    np.testing.assert_allclose(ttl.cpu().numpy(), outs[0][1].cpu().numpy(), rtol=1e-6)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("order", ["sorted", "shuffled"])
def test_ragged_tiles_stop_early(torch, mm, orc, dtype, order):
    """SURVEY.md §8f rank 4 (the intent of the PartialVector drafts, src/inference.jl:76-90,112-127): in a ragged batch an
    utterance tile (128 utterances) whose longest sequence has L < T frames runs frames 0..L only — forward sweep stops
    there, backward sweep starts there with B[:,end] = 1̄.  Same results as the full sweeps (MK_RAGGED_CUT=0) and as the
    oracle; device path, host pipeline (frame segments), and lengths down to 1."""
    K = mm.LogSemiring[dtype]
    rng = np.random.default_rng(77)
    B, T, D = 300, 36, 60   # three tiles: 128 + 128 + 44
    g = mm.graphs.denominator(K, n_tokens=400, n_pdf=D, seed=17)
    V = (rng.standard_normal((B, T, D)) * 2).astype(dtype)
    lens = np.sort(rng.integers(1, T + 1, B))[::-1].astype(np.int32)
    lens[0] = T
    lens[-1] = 1
    if order == "shuffled":
        lens = rng.permutation(lens).astype(np.int32)
        lens[200:] = np.minimum(lens[200:], 9)   # the last tile still stops early
    b = gpu_batch(mm, [g] * B, D, "shared")
    out = {}
    for flag in ("1", "0"):
        os.environ["MK_RAGGED_CUT"] = flag
        try:
            post, ttl = mm.pdfposteriors(b, dev(torch, V), seqlengths=lens)
            out[flag] = (post.cpu().numpy(), ttl.cpu().numpy())
        finally:
            os.environ.pop("MK_RAGGED_CUT", None)
    rt = 2e-5 if dtype == np.float32 else 1e-11
    np.testing.assert_allclose(out["1"][1], out["0"][1], rtol=rt)
    np.testing.assert_allclose(out["1"][0], out["0"][0], rtol=rt, atol=1e-9)
    # (one- and two-frame utterances have |log Z| < 0.1: Float32 rounding of the emissions alone is 1e-5 absolute)
    check_posteriors(mm, orc, [g] * B, D, V, lens, out["1"][0], out["1"][1], dtype, ttl_atol=2e-5)
    for n in range(B):  # nothing past an utterance's length, every real frame sums to 1
        assert not out["1"][0][n, :, lens[n]:].any()
        np.testing.assert_allclose(out["1"][0][n, :, :lens[n]].sum(axis=0), 1.0, rtol=1e-4)
    # host-buffer entry point: the same call cut into frame segments
    hpost, httl = mm.pdfposteriors(b, V.transpose(0, 2, 1), seqlengths=lens)
    np.testing.assert_allclose(np.array(httl), out["1"][1], rtol=rt)
    np.testing.assert_allclose(np.array(hpost), out["1"][0], rtol=rt, atol=1e-9)
    # the other entry points are unaffected by a previous cut call on the same batch
    A = mm.αrecursion(b, dev(torch, V), seqlengths=lens).cpu().numpy()
    og = orc_graphs(orc, [g], D)[0]
    for k in (0, 150, B - 1):
        oA = orc.alpha_beta(og, V[k], lens[k], want_beta=False)[0]
        assert_states_close(A[b.offsets[k]:b.offsets[k + 1]], oA, dtype)


def test_ragged_cut_saves_frames(torch, mm):
    """The cut is visible in the kernel time: 4 tiles of which 3 stop after a quarter of the frames."""
    K = mm.LogSemiring[np.float32]
    rng = np.random.default_rng(78)
    B, T, D = 512, 64, 200
    g = mm.graphs.denominator(K, n_tokens=3000, n_pdf=D, seed=18)
    V = torch.from_numpy((rng.standard_normal((B, T, D)) * 2).astype(np.float32)).cuda().permute(0, 2, 1)
    lens = np.full(B, T // 4, np.int32)
    lens[:128] = T
    b = gpu_batch(mm, [g] * B, D, "shared")
    ms = {}
    for flag in ("1", "0"):
        os.environ["MK_RAGGED_CUT"] = flag
        try:
            for _ in range(2):
                mm.pdfposteriors(b, V, seqlengths=lens)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                mm.pdfposteriors(b, V, seqlengths=lens)
            e1.record()
            torch.cuda.synchronize()
            ms[flag] = e0.elapsed_time(e1) / 3
        finally:
            os.environ.pop("MK_RAGGED_CUT", None)
    print(f"ragged cut: {ms['1']:.3f} ms vs {ms['0']:.3f} ms per call")
    assert ms["1"] < 0.8 * ms["0"], ms


def test_dimension_mismatch(torch, mm):
    """@boundscheck ... throw(DimensionMismatch()) (src/linalg.jl:166-167)."""
    K = mm.LogSemiring[np.float32]
    g = mm.graphs.hmm3(K)
    b = gpu_batch(mm, [g, g], 3)
    with pytest.raises(mm.DimensionMismatch):
        mm.pdfposteriors(b, torch.zeros((2, 7, 5), device="cuda"))  # 7 pdfs for a 3-pdf graph
    with pytest.raises(mm.DimensionMismatch):
        mm.pdfposteriors(b, torch.zeros((3, 3, 5), device="cuda"))  # 3 utterances for 2 FSMs
    with pytest.raises(mm.DimensionMismatch):
        mm.pdfposteriors(b, torch.zeros((2, 3, 5), device="cuda"), seqlengths=[5, 6])
    with pytest.raises(mm.DimensionMismatch):
        mm.bestpath(b, torch.zeros((2, 3, 5), device="cuda"))  # Log graphs
    Kd = mm.LogSemiring[np.float64]
    with pytest.raises(mm.DimensionMismatch):
        mm.batch(mm.compile(g[0], mm.statemap(g[0], 3, g[1])),
                 mm.compile(*(lambda f, p: (f, mm.statemap(f, 3, p)))(*mm.graphs.hmm3(Kd))))


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties + a sampled oracle comparison
# ---------------------------------------------------------------------------------------------
def test_cfg3_full_size_properties(torch, mm, orc):
    """cfg 3: S=30 000, ~520k arcs, D=3 000, B=128, T=150, Float32."""
    K = mm.LogSemiring[np.float32]
    B, T, D = 128, 150, 3000
    g = mm.graphs.denominator(K)
    gen = torch.Generator(device="cuda").manual_seed(303)
    V = torch.randn((B, T, D), generator=gen, device="cuda") * 2
    b = gpu_batch(mm, [g] * B, D)
    post, ttl = mm.pdfposteriors(b, V.permute(0, 2, 1))
    torch.cuda.synchronize()
    assert tuple(post.shape) == (B, D, T) and post.stride() == (1, B, B * D)
    assert bool(torch.isfinite(ttl).all()) and bool(torch.isfinite(post).all()) and float(post.min()) >= 0.0
    # every frame's pdf posteriors sum to 1 (the reference normalises per frame, :157-158)
    torch.testing.assert_close(post.sum(dim=1), torch.ones((B, T), device="cuda"), rtol=0, atol=2e-4)
    # linearity in a per-utterance constant: adding c to every log-likelihood adds T*c to logZ
    # and leaves the posteriors unchanged
    post2, ttl2 = mm.pdfposteriors(b, (V + 0.5).permute(0, 2, 1))
    torch.testing.assert_close(ttl2, ttl + 0.5 * T, rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(post2, post, rtol=1e-3, atol=1e-6)
    # sampled oracle comparison: 3 utterances of the 128
    idx = [0, 77, 127]
    Vh = V[idx].cpu().numpy()
    check_posteriors(mm, orc, [g] * len(idx), D, Vh, None, post[idx], ttl[idx], np.float32)


def test_cfg5_bestpath_full_graph(torch, mm, orc):
    """cfg 5 graph (tropical denominator) at a reduced batch: path is a valid arc sequence whose
    score equals the reported one, and matches the oracle on sampled utterances."""
    K = mm.TropicalSemiring[np.float32]
    B, T, D = 32, 100, 3000
    g = mm.graphs.denominator(K)
    fsm, pdfids = g
    rng = np.random.default_rng(505)
    V = (rng.standard_normal((B, T, D)) * 2).astype(np.float32)
    b = gpu_batch(mm, [g] * B, D)
    path, score = mm.bestpath(b, dev(torch, V))
    path, score = path.cpu().numpy(), score.cpu().numpy()
    import scipy.sparse as sp
    src, dst, w = fsm.arcs_hat()
    M = sp.csr_matrix((w.astype(np.float64) + 1e3, (src, dst)), shape=(fsm.nstates_hat,) * 2)  # shift: keep zeros explicit
    for k in (0, 13, 31):
        p = path[k] - 1
        tot = float(fsm.α[p[0]]) + float(V[k, 0, pdfids[p[0]]])
        for n in range(1, T):
            a = M[p[n - 1], p[n]]
            assert a != 0, "path uses a non-existent arc"
            tot += (a - 1e3) + float(V[k, n, pdfids[p[n]]])
        tot += float(fsm.ω[p[-1]])
        assert tot == pytest.approx(float(score[k]), rel=1e-5)
    idx = [0, 13]
    opath, oscore = orc.bestpath(orc_graphs(orc, [g] * 2, D), V[idx])
    np.testing.assert_array_equal(path[idx], opath)
    np.testing.assert_array_equal(score[idx], oscore)
