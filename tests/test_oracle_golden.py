# SPDX-License-Identifier: MIT
"""Pin the oracle to every literal known-answer vector the reference tree holds for this path
(R = emitted by the reference, D = derived from the reference's maths)."""
import numpy as np
import pytest

from conftest import load_golden_fsm

F64 = np.float64


def _graph(orc, fsm, pdfids, numpdf):
    return orc.OracleGraph(fsm, pdfids, numpdf)


# 2. R  test/test_semirings.jl:3-7
def test_logaddexp_kat(orc):
    assert orc.logaddexp(2.0, 3.0) == pytest.approx(3.3132616875182228, abs=1e-15)
    assert orc.logaddexp(10002.0, 10003.0) == pytest.approx(10000 + 3.3132616875182228, abs=1e-11)
    assert orc.logaddexp(-np.inf, 1.5) == 1.5 and orc.logaddexp(1.5, -np.inf) == 1.5
    assert orc.logaddexp(-np.inf, -np.inf) == -np.inf
    assert orc.logaddexp(2.0, 3.0, np.float32) == pytest.approx(3.3132617, rel=1e-6)


# 3. R-inputs / D-values  test/test_linalg.jl:93-95:  A = sparse([1,2,2,3,4],[3,1,2,1,3],[1,2,3,4,5],4,3)
def test_spmv_kat(orc):
    rows = np.array([1, 2, 2, 3, 4]) - 1
    cols = np.array([3, 1, 2, 1, 3]) - 1
    vals = np.array([1.0, 2, 3, 4, 5])
    order = np.lexsort((cols, rows))
    rowptr = np.zeros(5, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    x = np.array([1.0, 2, 3])
    log = orc.spmv(0, rowptr, cols[order], vals[order], x)
    np.testing.assert_allclose(log, [4, 5.126928011042972, 5, 8], rtol=0, atol=1e-14)
    np.testing.assert_array_equal(orc.spmv(1, rowptr, cols[order], vals[order], x), [4, 5, 5, 8])
    np.testing.assert_array_equal(orc.spmv(2, rowptr, cols[order], vals[order], x), [3, 8, 4, 15])
    # SpMM with reshape(1:12, 3, 4): column j of the Log result = [3j+1, 3j+2.1269.., 3j+2, 3j+5]
    Bm = np.arange(1, 13, dtype=F64).reshape(4, 3).T
    for j in range(4):
        col = orc.spmv(0, rowptr, cols[order], vals[order], Bm[:, j])
        np.testing.assert_allclose(col, [3 * j + 4, 3 * j + 5.126928011042972, 3 * j + 5, 3 * j + 8], atol=1e-13)


# 1. R  examples/demo.ipynb cell 13: 3-state L-R HMM, lhs = zeros(3,5)
GAMMA_DEMO = np.array([[1, .5, 1 / 6, 0, 0], [0, .5, 2 / 3, .5, 0], [0, 0, 1 / 6, .5, 1]])


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 2e-6)])
def test_demo_posteriors(orc, mm, dtype, tol):
    K = mm.LogSemiring[dtype]
    fsm, pdfids = mm.graphs.hmm3(K)
    g = _graph(orc, fsm, pdfids, 3)
    post, ttl = orc.pdfposteriors([g], np.zeros((1, 5, 3)))
    np.testing.assert_allclose(post[0], GAMMA_DEMO, atol=tol)
    assert ttl[0] == pytest.approx(np.log(6 / 32), abs=tol * 10)  # D: -1.6739764335716716
    # the per-frame totals are identical for every frame: α_n·β_n sums to logZ
    A, Bm = orc.alpha_beta(g, np.zeros((5, 3)))
    z = np.logaddexp.reduce(A + Bm, axis=0)
    np.testing.assert_allclose(z, np.log(6 / 32), atol=tol * 10)
    # second, independent restatement (dense logsumexp, test/test_algorithms.jl:28-63)
    dpost, dz = orc.dense_forward_backward(fsm, pdfids, np.zeros((3, 5)))
    np.testing.assert_allclose(dpost, GAMMA_DEMO, atol=tol)  # (f32: the weights are rounded)
    assert dz == pytest.approx(np.log(6 / 32), abs=tol * 10)


# 4. R-design  test/test_algorithms.jl:218-248: ragged batch, lhs = ones(3,7,2), seqlengths = [5,7]
def test_ragged_batch(orc, mm):
    K = mm.LogSemiring[np.float32]
    fsm, pdfids = mm.graphs.hmm3(K)
    g = _graph(orc, fsm, pdfids, 3)
    V = np.ones((2, 7, 3), np.float32)
    post, ttl = orc.pdfposteriors([g, g], V, [5, 7])
    p1, t1 = orc.pdfposteriors([g], V[:1, :5])
    p2, t2 = orc.pdfposteriors([g], V[1:, :7])
    np.testing.assert_allclose(post[0, :, :5], p1[0], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(post[1], p2[0], rtol=1e-6, atol=1e-7)
    assert ttl[0] == pytest.approx(t1[0], rel=1e-6) and ttl[1] == pytest.approx(t2[0], rel=1e-6)
    assert np.all(post[0, :, 5:] == 0.0)  # @test all(γ[:,6:7,1] .== zero(T))
    for k, L in enumerate((5, 7)):
        d, dz = orc.dense_forward_backward(fsm, pdfids, np.ones((3, 7)), L)
        np.testing.assert_allclose(post[k, :, :L], d, atol=2e-6)
        assert ttl[k] == pytest.approx(dz, rel=1e-6)


# 5. R  test/test_algorithms.jl:262-283: chain a->b->c->d, lhs = ones(4,4), Tropical => "a b c d"
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bestpath_chain(orc, mm, dtype):
    K = mm.TropicalSemiring[dtype]
    fsm, pdfids = mm.graphs.chain(K)
    g = _graph(orc, fsm, pdfids, 4)
    path, score = orc.bestpath([g], np.ones((1, 4, 4)))
    np.testing.assert_array_equal(path[0], [1, 2, 3, 4])
    assert score[0] == 4.0
    # too short / too long: no path through a 4-chain in 3 or 5 frames
    path, score = orc.bestpath([g], np.ones((1, 5, 4)))
    assert score[0] == -np.inf and np.all(path == 0)
    path, score = orc.bestpath([g, g], np.ones((2, 6, 4)), [4, 3])
    np.testing.assert_array_equal(path[0], [1, 2, 3, 4, 0, 0])
    assert score[1] == -np.inf and np.all(path[1] == 0)


# 6. D  misc/benchmark/den_fsm_wsj.txt with lhs = ones(84, N) (benchmark.jl:31) and zeros
DEN_LOGZ = {("ones", 20): 12.578496038935, ("ones", 100): 92.531024082122, ("zeros", 20): -7.421503961065,
            ("zeros", 100): -7.468975917878}


def test_den_fsm_wsj_fixture(orc, mm):
    K = mm.LogSemiring[np.float64]
    fsm, pdfids = load_golden_fsm("den_fsm_wsj", K)
    assert fsm.nstates == 3032 and fsm.nnz_hat == 50984 + 942 + 1 and pdfids.max() + 1 == 84
    assert fsm.init_idx.size == 38
    g = _graph(orc, fsm, pdfids, 84)
    for (kind, N), want in DEN_LOGZ.items():
        V = np.ones((1, N, 84)) if kind == "ones" else np.zeros((1, N, 84))
        post, ttl = orc.pdfposteriors([g], V)
        assert ttl[0] == pytest.approx(want, abs=2e-9), (kind, N)
        np.testing.assert_allclose(post[0].sum(axis=0), 1.0, atol=1e-9)
    # active (finite-α) states per frame: 38, 577, 1972, 2911, 2994
    A, _ = orc.alpha_beta(g, np.ones((6, 84)), want_beta=False)
    assert list(np.isfinite(A[:-1, :5]).sum(axis=0)) == [38, 577, 1972, 2911, 2994]


@pytest.mark.parametrize("N,want", [(700, 692.168685813936)])
def test_den_fsm_wsj_benchmark_length(orc, mm, N, want):
    K = mm.LogSemiring[np.float64]
    fsm, pdfids = load_golden_fsm("den_fsm_wsj", K)
    g = _graph(orc, fsm, pdfids, 84)
    _, ttl = orc.pdfposteriors([g], np.ones((1, N, 84)))
    assert ttl[0] == pytest.approx(want, abs=5e-9)


# 7. D  num_fsm_wsj.txt: final unreachable in fewer than 166 frames
def test_num_fsm_wsj_unreachable(orc, mm):
    K = mm.LogSemiring[np.float64]
    fsm, pdfids = load_golden_fsm("num_fsm_wsj", K)
    assert fsm.nstates == 454
    g = _graph(orc, fsm, pdfids, 84)
    post, ttl = orc.pdfposteriors([g], np.zeros((1, 165, 84)))
    assert ttl[0] == -np.inf and np.all(post == 0.0)  # pdfposteriors3 convention, src/inference.jl:198-200
    post, ttl = orc.pdfposteriors([g], np.zeros((1, 166, 84)))
    assert np.isfinite(ttl[0])
    np.testing.assert_allclose(post[0].sum(axis=0), 1.0, atol=1e-9)


# C++ oracle vs the dense float64 restatement on a random mid-size graph, Log and Tropical
def test_cpp_vs_dense_random(orc, mm):
    rng = np.random.default_rng(7)
    K = mm.LogSemiring[np.float64]
    fsm, pdfids = mm.graphs.phone_loop(K, n_phones=5)
    D = fsm.nstates
    V = rng.standard_normal((D, 40))
    g = _graph(orc, fsm, pdfids, D)
    post, ttl = orc.pdfposteriors([g], V.T[None])
    dpost, dz = orc.dense_forward_backward(fsm, pdfids, V)
    np.testing.assert_allclose(post[0], dpost, atol=1e-12)
    assert ttl[0] == pytest.approx(dz, abs=1e-10)
    Kt = mm.TropicalSemiring[np.float64]
    fsm_t = mm.FSM(Kt, fsm.nstates_hat, fsm.init_idx, fsm.init_w, fsm.colptr, fsm.rowval, fsm.nzval)
    path, score = orc.bestpath([_graph(orc, fsm_t, pdfids, D)], V.T[None])
    dpath, dscore = orc.dense_viterbi(fsm_t, pdfids, V)
    np.testing.assert_array_equal(path[0], dpath)
    assert score[0] == pytest.approx(dscore, abs=1e-10)
