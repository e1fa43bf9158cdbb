# SPDX-License-Identifier: MIT
"""totalweightsum (src/algorithms.jl:8-16, :32-36) on the device against a dense Float64 restatement —
the quantity the reference's FSM tests compare (test/test_fsms.jl:9-16)."""
import numpy as np
import pytest

from test_gpu_parity import torch  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def _dense_totalcumsum(fsm, n, tropical):
    """totalcumsum(α, T, ω, n) with dense matrices in Float64 (log / tropical payloads)."""
    a, T, w = fsm.α.astype(np.float64), fsm.T.astype(np.float64), fsm.ω.astype(np.float64)

    def oplus(x, axis=None):
        x = np.asarray(x, np.float64)
        if tropical:
            return np.max(x, axis=axis)
        m = np.max(x, axis=axis, keepdims=True)
        m = np.where(np.isfinite(m), m, 0.0)
        return np.squeeze(m, axis=axis) + np.log(np.sum(np.exp(x - m), axis=axis))

    v = a
    terms = [oplus(v + w, axis=0)]
    with np.errstate(divide="ignore"):
        for _ in range(2, n + 1):
            v = oplus(T + v[:, None], axis=0)  # (Tᵀ v)[j] = ⊕_i T[i, j] ⊗ v[i]
            terms.append(oplus(v + w, axis=0))
        return float(oplus(np.array(terms), axis=0))


@pytest.mark.parametrize("semiring", ["log", "tropical"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_totalweightsum_vs_dense(torch, mm, semiring, dtype):
    K = (mm.LogSemiring if semiring == "log" else mm.TropicalSemiring)[dtype]
    rng = np.random.default_rng(4)
    graphs = [mm.graphs.hmm3(K, 5)[0], mm.graphs.phone_loop(K, n_phones=4)[0],
              mm.graphs.numerator(K, rng, 50, n_phones=6)[0]]
    for fsm in graphs:
        for n in (1, 2, fsm.nstates, fsm.nstates + 7):
            with np.errstate(divide="ignore", invalid="ignore"):
                want = _dense_totalcumsum(fsm, n, semiring == "tropical")
            got = mm.totalweightsum(fsm, n)
            if np.isneginf(want):
                assert np.isneginf(got)
            else:
                assert got == pytest.approx(want, rel=1e-4 if dtype == np.float32 else 1e-9, abs=1e-5 if dtype == np.float32 else 1e-10)


def test_totalweightsum_union_is_sum_of_parts(torch, mm):
    """The reference's use: two FSMs are compared through their total weight sums (test/test_fsms.jl:9-16);
    the union of two FSMs sums them."""
    K = mm.LogSemiring[np.float64]
    f1, f2 = mm.graphs.hmm3(K, 4)[0], mm.graphs.phone_loop(K, n_phones=3)[0]
    u = mm.union(f1, f2)
    n = 9
    a, b, c = mm.totalweightsum(f1, n), mm.totalweightsum(f2, n), mm.totalweightsum(u, n)
    assert c == pytest.approx(np.logaddexp(a, b), rel=1e-9)
