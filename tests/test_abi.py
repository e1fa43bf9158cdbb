# SPDX-License-Identifier: MIT
"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/markov_b200.h declares, and the host logic (FSM construction, rawunion, expand, Ĉ
handling) matches the reference's semantics.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(mm):
    hdr = open(os.path.join(ROOT, "include", "markov_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mk_[a-z_0-9]+)\s*\(", hdr))
    assert {"mk_graph_create", "mk_batch_create", "mk_alpha", "mk_beta", "mk_pdfposteriors", "mk_bestpath",
            "mk_pdfposteriors_host", "mk_bestpath_host", "mk_last_error"} <= declared
    lib = mm.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from markov_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    assert lib.mk_abi_version() == 2


def test_no_cpu_fallback(mm):
    """Without a CUDA device compile() must fail loudly, not fall back."""
    if mm.lib().mk_device_count() > 0:
        pytest.skip("a GPU is present")
    K = mm.LogSemiring[np.float32]
    fsm, pdfids = mm.graphs.hmm3(K)
    with pytest.raises(mm.MarkovError) as ei:
        mm.compile(fsm, mm.statemap(fsm, 3, pdfids), device=-1)
    assert ei.value.code == 1000 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "markovmodels.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f


def test_fsm_extension(mm):
    """T̂ = [T ω; 0̄ 1̄], α̂ = [α; 0̄]  (src/fsm.jl:19-28)."""
    K = mm.LogSemiring[np.float64]
    one = K.one
    fsm = mm.FSM.from_pairs(K, [(1, one)], [((1, 1), one), ((1, 2), one), ((2, 2), one), ((2, 1), one)],
                            [(2, one)], [1, 2])  # the FSM of test/test_fsms.jl:25-51
    assert fsm.nstates == 2 and fsm.nstates_hat == 3 and fsm.nnz_hat == 6
    src, dst, w = fsm.arcs_hat()
    assert sorted(zip(src.tolist(), dst.tolist())) == [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2), (2, 2)]
    np.testing.assert_array_equal(fsm.α, [0.0, -np.inf])
    np.testing.assert_array_equal(fsm.ω, [-np.inf, 0.0])
    js = mm.FSM.from_json('{"semiring": "LogSemiring{Float64}", "initstates": [[1, 0.0]], "arcs": [[1,1,0.0],'
                          '[1,2,0.0],[2,2,0.0],[2,1,0.0]], "finalstates": [[2, 0.0]], "labels": [1, 2]}')
    np.testing.assert_array_equal(js.colptr, fsm.colptr)
    np.testing.assert_array_equal(js.rowval, fsm.rowval)
    np.testing.assert_array_equal(js.nzval, fsm.nzval)


def test_duplicate_arcs_are_semiring_summed(mm):
    K = mm.LogSemiring[np.float64]
    fsm = mm.FSM.from_arrays(K, 2, [0, 0], [1, 1], [np.log(0.25), np.log(0.25)], [0], [0.0], [1], [0.0])
    assert fsm.T[0, 1] == pytest.approx(np.log(0.5))
    Kt = mm.TropicalSemiring[np.float64]
    fsm = mm.FSM.from_arrays(Kt, 2, [0, 0], [1, 1], [-1.0, -2.0], [0], [0.0], [1], [0.0])
    assert fsm.T[0, 1] == -1.0


def test_rawunion_keeps_one_phony_final_per_operand(mm):
    """src/fsmops.jl:28-36 vs union :8-17."""
    K = mm.LogSemiring[np.float32]
    a, _ = mm.graphs.hmm3(K)
    b, _ = mm.graphs.chain(K)
    u = mm.rawunion(a, b, a)
    assert u.nstates_hat == a.nstates_hat * 2 + b.nstates_hat
    assert u.nnz_hat == 2 * a.nnz_hat + b.nnz_hat
    assert len(u.parts) == 3 and u.parts[0] is a and u.parts[2] is a
    np.testing.assert_array_equal(u.init_idx, [0, 4, 9])
    # block-diagonal: no arc crosses a block boundary
    src, dst, _ = u.arcs_hat()
    blk = np.searchsorted([4, 9, 13], src, side="right")
    np.testing.assert_array_equal(blk, np.searchsorted([4, 9, 13], dst, side="right"))
    m = mm.union(a, b)
    assert m.nstates == a.nstates + b.nstates and m.nstates_hat == m.nstates + 1


def test_renorm_rows_sum_to_one(mm):
    K = mm.LogSemiring[np.float64]
    fsm, _ = mm.graphs.phone_loop(K, n_phones=4)
    src, dst, w = fsm.arcs_hat()
    tot = np.full(fsm.nstates_hat, -np.inf)
    np.logaddexp.at(tot, src, w)
    np.testing.assert_allclose(tot, 0.0, atol=1e-12)
    den, _ = mm.graphs.denominator(K, n_tokens=200, n_pdf=40)
    src, dst, w = den.arcs_hat()
    tot = np.full(den.nstates_hat, -np.inf)
    np.logaddexp.at(tot, src, w)
    np.testing.assert_allclose(tot, 0.0, atol=1e-9)


def test_expand(mm):
    """src/inference.jl:54-60."""
    V = np.arange(6, dtype=np.float32).reshape(2, 3)
    E = mm.expand(V, 2)
    assert E.shape == (3, 4)
    np.testing.assert_array_equal(E[:2, :2], V[:, :2])
    assert np.all(E[:2, 2:] == -np.inf) and np.all(E[2, :2] == -np.inf) and np.all(E[2, 2:] == 0.0)
    E = mm.expand(V)
    np.testing.assert_array_equal(E[:2, :3], V)
    assert np.all(E[2, :3] == -np.inf) and E[2, 3] == 0.0 and np.all(E[:2, 3] == -np.inf)


def test_statemap(mm):
    """examples/prepare-lfmmi-graphs.jl:15-23: Ĉ[s, pdf(s)] = 1̄, phony state -> phony pdf."""
    K = mm.LogSemiring[np.float32]
    fsm, pdfids = mm.graphs.hmm3(K)
    sm = mm.statemap(fsm, 5, [4, 0, 2])
    np.testing.assert_array_equal(sm.state2pdf, [4, 0, 2, 5])
    dense = sm.dense(K)
    assert dense.shape == (4, 6) and np.all(np.isfinite(dense).sum(axis=1) == 1)
    from markov_b200.inference import _as_statemap
    np.testing.assert_array_equal(_as_statemap(fsm, dense).state2pdf, sm.state2pdf)
    with pytest.raises(IndexError):
        mm.statemap(fsm, 2, [0, 1, 2])


def test_semiring_descriptors(mm):
    K = mm.LogSemiring[np.float32]
    assert K is mm.LogSemiring[np.float32] and K != mm.LogSemiring[np.float64]
    assert K.zero == -np.inf and K.one == 0.0 and repr(K) == "LogSemiring{Float32}"
    assert K.add(2.0, 3.0) == pytest.approx(3.3132617, rel=1e-6)
    assert mm.TropicalSemiring[np.float64].add(2.0, 3.0) == 3.0


def test_comm_exports_validate_arguments(mm):
    """The multi-GPU exports of SURVEY.md 8(b) exist and reject bad arguments without a GPU."""
    import ctypes as C
    from markov_b200 import _lib
    lib = mm.lib()
    assert lib.mk_comm_unique_id(None) == _lib.MK_EINVAL
    h = C.c_void_p()
    ident = (C.c_char * 128)()
    assert lib.mk_comm_init_rank(C.byref(h), 2, 5, ident, -1) == _lib.MK_EINVAL      # rank outside [0, n_ranks)
    assert lib.mk_comm_init_rank(None, 1, 0, ident, -1) == _lib.MK_EINVAL
    assert lib.mk_allreduce_stats(None, None, 4, None) == _lib.MK_EINVAL
    assert lib.mk_comm_destroy(None) == _lib.MK_OK
    assert b"communicator" in lib.mk_last_error() or b"null" in lib.mk_last_error()
