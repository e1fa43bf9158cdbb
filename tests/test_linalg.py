# SPDX-License-Identifier: MIT
"""Operator level (mk_spmv / mk_spmm / mk_spvec_bcast) against the oracle — these read like the reference's
own enabled tests, test/test_linalg.jl: "mul!" (:88-108, the same 4 x 3 matrix, dense 3 x 4 matrix and vector,
LogSemiring / ProbSemiring / TropicalSemiring x Float32 / Float64) and the sparse-vector broadcasts (:34-54),
where the reference compares its GPU kernels with the CPU generic path; here the CPU side is the oracle's
CSR SpMV restatement (oracle.spmv, src/linalg.jl:163-184 contract, CPU accumulation order)."""
import numpy as np
import pytest

from test_gpu_parity import torch  # noqa: F401  (fixture)

SEMIRINGS = ["LogSemiring", "ProbSemiring", "TropicalSemiring"]
DTYPES = [np.float32, np.float64]


def _K(mm, name, dtype):
    return getattr(mm, name)[dtype]


def _rel(dtype):
    return 1e-5 if dtype == np.float32 else 1e-12


def _oracle_mul(orc, K, I, J, V, m, n, B):  # noqa: E741
    """A ⊗ B column by column with the oracle's SpMV (float64 payloads)."""
    order = np.lexsort((J, I))
    I, J, V = np.asarray(I)[order], np.asarray(J)[order], np.asarray(V, np.float64)[order]  # noqa: E741
    rowptr = np.concatenate(([0], np.cumsum(np.bincount(I - 1, minlength=m))))
    B = np.asarray(B, np.float64)
    if B.ndim == 1:
        return orc.spmv(K.code, rowptr, J - 1, V, B)
    return np.stack([orc.spmv(K.code, rowptr, J - 1, V, B[:, j]) for j in range(B.shape[1])], axis=1)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sr", SEMIRINGS)
def test_mul_reference_testset(torch, mm, orc, sr, dtype):
    """test/test_linalg.jl:88-108 verbatim: sm = sparse([1,2,2,3,4], [3,1,2,1,3], K[1,2,3,4,5], 4, 3),
    dm = reshape(K.(1:12), 3, 4), dv = K.(1:3); mul!(similar(dm, 4, 4), sm, dm) and mul!(similar(dv, 4), sm, dv)."""
    K = _K(mm, sr, dtype)
    I, J, V = [1, 2, 2, 3, 4], [3, 1, 2, 1, 3], [1, 2, 3, 4, 5]  # noqa: E741
    dm = np.arange(1, 13, dtype=np.float64).reshape(4, 3).T  # column-major reshape(1:12, 3, 4)
    dv = np.arange(1, 4, dtype=np.float64)
    cu_sm = mm.CuSparseMatrixCSR(K, I, J, V, 4, 3)
    cu_dm = mm.linalg.to_colmajor(K, dm)
    cu_dv = torch.from_numpy(dv.astype(dtype)).cuda()

    cu_r = mm.mul_(mm.linalg.colmajor(K, 4, 4, fill=123.0), cu_sm, cu_dm)  # `similar`: uninitialised, β = 0 clears it
    np.testing.assert_allclose(cu_r.cpu().numpy(), _oracle_mul(orc, K, I, J, V, 4, 3, dm), rtol=_rel(dtype))

    cu_r = mm.mul_(torch.empty(4, dtype=cu_dv.dtype, device="cuda"), cu_sm, cu_dv)
    np.testing.assert_allclose(cu_r.cpu().numpy(), _oracle_mul(orc, K, I, J, V, 4, 3, dv), rtol=_rel(dtype))


@pytest.mark.gpu
def test_mul_known_answers(torch, mm):
    """Literal values for the testset above (hand-derivable): row 1 = A[1,3] ⊗ b[3], row 2 = A[2,1] ⊗ b[1] ⊕ A[2,2] ⊗ b[2]…"""
    I, J, V = [1, 2, 2, 3, 4], [3, 1, 2, 1, 3], [1, 2, 3, 4, 5]  # noqa: E741
    dv = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64, device="cuda")
    want = {"LogSemiring": [4.0, np.logaddexp(3.0, 5.0), 5.0, 8.0], "TropicalSemiring": [4.0, 5.0, 5.0, 8.0],
            "ProbSemiring": [3.0, 8.0, 4.0, 15.0]}
    for sr, w in want.items():
        A = mm.CuSparseMatrixCSR(_K(mm, sr, np.float64), I, J, V, 4, 3)
        got = mm.mul_(torch.empty(4, dtype=torch.float64, device="cuda"), A, dv)
        np.testing.assert_allclose(got.cpu().numpy(), w, rtol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sr", SEMIRINGS)
def test_mul_random_graph_sizes(torch, mm, orc, sr, dtype):
    """The shapes the path uses: T̂ᵀ of a denominator-like graph (≈17 arcs per row -> 8 lanes per row), a Ĉ-like
    matrix (one entry per row -> 4 lanes), a dense-ish short-wide matrix (-> a warp per row), empty rows, 0̄ entries."""
    K = _K(mm, sr, dtype)
    rng = np.random.default_rng(7)
    for m, n, per_row in ((3000, 3000, 17), (5000, 300, 1), (37, 4000, 900), (129, 65, 3)):
        I = np.repeat(np.arange(1, m + 1), per_row)  # noqa: E741
        J = rng.integers(1, n + 1, I.size)
        keep = rng.random(I.size) < 0.9
        keep[I == 2] = False  # an empty row
        I, J = I[keep], J[keep]  # noqa: E741
        if m == 3000:  # two rows far longer than the rest (the phony final state's row of T̂ᵀ): the CTA-per-row path
            for row, cnt in ((6, 2900), (2999, 1500)):
                I = np.concatenate([I, np.full(cnt, row)])  # noqa: E741
                J = np.concatenate([J, np.arange(1, cnt + 1)])
        _, first = np.unique(np.stack([I, J]), axis=1, return_index=True)  # distinct entries (no ⊕ of duplicates)
        I, J = I[first], J[first]  # noqa: E741
        if K.code == 2:
            V = rng.random(I.size)
            b = rng.random(n)
            V[::11] = 0.0
        else:
            V = rng.standard_normal(I.size) * 3
            b = rng.standard_normal(n) * 30  # a wide dynamic range: exp(b) alone would overflow Float32
            V[::11] = -np.inf
            b[::13] = -np.inf
        V, b = V.astype(dtype), b.astype(dtype)
        A = mm.CuSparseMatrixCSR(K, I, J, V, m, n)
        got = mm.mul_(torch.full((m,), 777.0, dtype=torch.from_numpy(b).dtype, device="cuda"), A,
                      torch.from_numpy(b).cuda()).cpu().numpy()
        want = _oracle_mul(orc, K, I, J, V, m, n, b)
        np.testing.assert_allclose(got, want, rtol=1e-4 if dtype == np.float32 else 1e-11, atol=1e-30)
        assert got[1] == K.zero  # the empty row is 0̄, not the previous content
        # matrix form, 5 columns, with a leading dimension larger than the row count, β = 0 and β = 1
        Bm = np.stack([np.roll(b, k) for k in range(5)], axis=1)
        cuB = mm.linalg.to_colmajor(K, Bm)
        big = mm.linalg.colmajor(K, m + 3, 5, fill=float(K.zero))
        C = big[:m, :]
        assert C.stride(1) == m + 3
        mm.mul_(C, A, cuB)
        wantM = _oracle_mul(orc, K, I, J, V, m, n, Bm)
        np.testing.assert_allclose(C.cpu().numpy(), wantM, rtol=1e-4 if dtype == np.float32 else 1e-11, atol=1e-30)
        np.testing.assert_array_equal(big[m:, :].cpu().numpy(), np.full((3, 5), K.zero))  # padding rows untouched
        mm.mul_(C, A, cuB, True, True)  # C ⊕= A ⊗ B  ->  C ⊕ C
        twice = K.add_ufunc(wantM, wantM)
        np.testing.assert_allclose(C.cpu().numpy(), twice, rtol=1e-4 if dtype == np.float32 else 1e-11, atol=1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sr", SEMIRINGS)
def test_mul_vector_block_diagonal_large(torch, mm, orc, sr, dtype):
    """blockdiag(T̂ᵀ ...) ⊗ b at scale (one frame of the reference's αrecursion, src/inference.jl:69-72): matrices of >= 2^16
    graph-like rows go through the kernel that keeps a window of b in shared memory.  14 blocks of 5 000 x 5 000 (a row block
    of the kernel sees one or two of them; in Float64 the window holds only 15 360 entries), ~17 arcs per row, every 97th row
    with arcs all over the matrix (outside any window), one long row per block (the phony final state's), an empty row,
    0̄ entries."""
    K = _K(mm, sr, dtype)
    rng = np.random.default_rng(13)
    blk, nblk = 5000, 14
    m = n = blk * nblk
    per_row = rng.integers(8, 27, m)
    I = np.repeat(np.arange(1, m + 1), per_row)  # noqa: E741
    J = ((I - 1) // blk) * blk + rng.integers(0, blk, I.size) + 1
    far = (I % 97) == 0
    J[far] = rng.integers(1, n + 1, int(far.sum()))
    keep = I != 3  # an empty row
    I, J = I[keep], J[keep]  # noqa: E741
    for k in range(nblk):  # the long rows: 1 200 arcs from the block's first columns
        I = np.concatenate([I, np.full(1200, (k + 1) * blk)])  # noqa: E741
        J = np.concatenate([J, k * blk + np.arange(1, 1201)])
    _, first = np.unique(np.stack([I, J]), axis=1, return_index=True)
    I, J = I[first], J[first]  # noqa: E741
    if K.code == 2:
        V, b = rng.random(I.size), rng.random(n)
        V[::11] = 0.0
    else:
        V, b = rng.standard_normal(I.size) * 3, rng.standard_normal(n) * 30
        V[::11] = -np.inf
        b[::13] = -np.inf
    V, b = V.astype(dtype), b.astype(dtype)
    A = mm.CuSparseMatrixCSR(K, I, J, V, m, n)
    got = mm.mul_(torch.full((m,), 777.0, dtype=torch.from_numpy(b).dtype, device="cuda"), A,
                  torch.from_numpy(b).cuda()).cpu().numpy()
    want = _oracle_mul(orc, K, I, J, V, m, n, b)
    np.testing.assert_allclose(got, want, rtol=1e-4 if dtype == np.float32 else 1e-11, atol=1e-30)
    assert got[2] == K.zero


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sr", SEMIRINGS)
def test_mul_matrix_block_diagonal_large(torch, mm, orc, sr, dtype):
    """Ĉ·V̂ at scale (src/inference.jl:150): products with >= 2^17 rows stage the column window of every block of rows
    in shared memory.  Block-diagonal part (narrow windows: staged; in Float64 some windows are too wide and read B
    directly), a tail of rows with columns all over the matrix (never staged), one-arc and many-arc rows, empty rows,
    0̄ entries, 7 columns (a full chunk of 4 and a partial one), padded leading dimensions, β = 0 and β = 1."""
    K = _K(mm, sr, dtype)
    rng = np.random.default_rng(11)
    blk_rows, blk_cols, nblk, tail = 7000, 900, 20, 12000
    m, n = blk_rows * nblk + tail, blk_cols * nblk
    per_row = rng.integers(0, 4, m)
    per_row[rng.random(m) < 0.6] = 1
    I = np.repeat(np.arange(1, m + 1), per_row)  # noqa: E741
    blk = np.minimum((I - 1) // blk_rows, nblk - 1)
    J = blk * blk_cols + rng.integers(0, blk_cols, I.size) + 1
    in_tail = I > blk_rows * nblk
    J[in_tail] = rng.integers(1, n + 1, int(in_tail.sum()))
    _, first = np.unique(np.stack([I, J]), axis=1, return_index=True)
    I, J = I[first], J[first]  # noqa: E741
    if K.code == 2:
        V, Bm = rng.random(I.size), rng.random((n, 7))
        V[::11] = 0.0
    else:
        V, Bm = rng.standard_normal(I.size) * 3, rng.standard_normal((n, 7)) * 30
        V[::11] = -np.inf
        Bm[::13, :] = -np.inf
    V, Bm = V.astype(dtype), Bm.astype(dtype)
    A = mm.CuSparseMatrixCSR(K, I, J, V, m, n)
    bigB = mm.linalg.colmajor(K, n + 5, 7, fill=float(K.zero))
    bigB[:n, :] = torch.from_numpy(Bm).cuda()
    big = mm.linalg.colmajor(K, m + 3, 7, fill=777.0)
    C = big[:m, :]
    mm.mul_(C, A, bigB[:n, :])
    want = _oracle_mul(orc, K, I, J, V, m, n, Bm)
    tol = dict(rtol=1e-4 if dtype == np.float32 else 1e-11, atol=1e-30)
    np.testing.assert_allclose(C.cpu().numpy(), want, **tol)
    np.testing.assert_array_equal(big[m:, :].cpu().numpy(), np.full((3, 7), 777.0))  # padding rows untouched
    mm.mul_(C, A, bigB[:n, :], True, True)
    np.testing.assert_allclose(C.cpu().numpy(), K.add_ufunc(want, want), **tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sr", SEMIRINGS)
def test_mul_matrix_one_arc_rows_large(torch, mm, sr, dtype):
    """Ĉ·V̂ (src/inference.jl:150; Ĉ has exactly one entry per row, examples/prepare-lfmmi-graphs.jl:15-23) at batch size: the
    Float32 kernel takes four adjacent rows per thread and writes 16 bytes per column.  The ⊕ over one term is the term, so the
    result is exact: C[i, j] = A[i, c_i] ⊗ B[c_i, j].  Row counts divisible by four and not (tail rows, unaligned columns),
    9 columns (two full chunks and a partial one), 0̄ weights and operands."""
    K = _K(mm, sr, dtype)
    rng = np.random.default_rng(19)
    blk_rows, blk_cols = 9000, 700
    for m in (144000, 144003):
        nblk = (m + blk_rows - 1) // blk_rows
        n = nblk * blk_cols
        I = np.arange(1, m + 1)  # noqa: E741
        J = ((I - 1) // blk_rows) * blk_cols + rng.integers(0, blk_cols, m) + 1
        if K.code == 2:
            V, Bm = rng.random(m), rng.random((n, 9))
            V[::17] = 0.0
            want = V[:, None] * Bm[J - 1, :]
        else:
            V, Bm = rng.standard_normal(m) * 3, rng.standard_normal((n, 9)) * 30
            V[::17] = -np.inf
            Bm[::13, :] = -np.inf
            want = V[:, None] + Bm[J - 1, :]
        V, Bm = V.astype(dtype), Bm.astype(dtype)
        want = (V[:, None] * Bm[J - 1, :]) if K.code == 2 else (V[:, None] + Bm[J - 1, :])
        A = mm.CuSparseMatrixCSR(K, I, J, V, m, n)
        C = mm.mul_(mm.linalg.colmajor(K, m, 9, fill=777.0), A, mm.linalg.to_colmajor(K, Bm))
        np.testing.assert_array_equal(C.cpu().numpy(), want.astype(dtype))


@pytest.mark.gpu
def test_mul_zero_based_indices_large(torch, mm):
    """The C ABI takes CUSPARSE-style 0-based arrays as well (`index_base = 0`): the batch-sized kernels (row-block windows of
    mk_spmm, lane groups + long-row list of mk_spmv) must give the same bits as with CUDA.jl's 1-based arrays."""
    from markov_b200 import _lib
    K = mm.LogSemiring[np.float32]
    rng = np.random.default_rng(17)
    blk, nblk = 6000, 24
    m, n = blk * nblk, 500 * nblk
    per_row = rng.integers(1, 4, m)
    I = np.repeat(np.arange(1, m + 1), per_row)  # noqa: E741
    J = ((I - 1) // blk) * 500 + rng.integers(0, 500, I.size) + 1
    _, first = np.unique(np.stack([I, J]), axis=1, return_index=True)
    I, J = I[first], J[first]  # noqa: E741
    V = (rng.standard_normal(I.size) * 3).astype(np.float32)
    A = mm.CuSparseMatrixCSR(K, I, J, V, m, n)
    B = mm.linalg.to_colmajor(K, (rng.standard_normal((n, 6)) * 5).astype(np.float32))
    b = B[:, 0].contiguous()
    C1 = mm.mul_(mm.linalg.colmajor(K, m, 6), A, B)
    c1 = mm.mul_(torch.empty(m, device="cuda"), A, b)
    rp0, cv0 = (A.rowPtr - 1).contiguous(), (A.colVal - 1).contiguous()
    C0 = mm.linalg.colmajor(K, m, 6)
    c0 = torch.empty(m, device="cuda")
    l = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(l.mk_spmm(K.code, K.dtype_code, m, n, A.nnz, rp0.data_ptr(), cv0.data_ptr(), A.nzVal.data_ptr(), 0, B.data_ptr(),
                         n, 6, B.stride(1), C0.data_ptr(), m, 6, C0.stride(1), 0, st))
    _lib.check(l.mk_spmv(K.code, K.dtype_code, m, n, A.nnz, rp0.data_ptr(), cv0.data_ptr(), A.nzVal.data_ptr(), 0, b.data_ptr(),
                         n, c0.data_ptr(), m, st))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(C0.cpu().numpy(), C1.cpu().numpy())
    np.testing.assert_array_equal(c0.cpu().numpy(), c1.cpu().numpy())


@pytest.mark.gpu
def test_mul_log_full_range(torch, mm):
    """⊕ of the Log semiring never exponentiates an un-shifted value: payloads around ±1e4 (exp overflows /
    underflows in both precisions) still give max + log(count)."""
    for dtype in DTYPES:
        K = mm.LogSemiring[dtype]
        for off in (1.0e4, -1.0e4):
            A = mm.CuSparseMatrixCSR(K, [1] * 40, list(range(1, 41)), [off] * 40, 1, 40)
            b = torch.zeros(40, dtype=torch.from_numpy(np.zeros(1, dtype)).dtype, device="cuda")
            got = float(mm.mul_(torch.empty(1, dtype=b.dtype, device="cuda"), A, b)[0])
            assert got == pytest.approx(off + np.log(40.0), rel=1e-6)


@pytest.mark.gpu
def test_mul_dimension_mismatch_and_empty(torch, mm):
    """@boundscheck of src/linalg.jl:166-167, 242-244 -> DimensionMismatch; an empty matrix launches nothing (:169)."""
    K = mm.LogSemiring[np.float32]
    A = mm.CuSparseMatrixCSR(K, [1, 2], [1, 3], [0.5, 0.25], 2, 3)
    f = lambda *s: torch.zeros(*s, dtype=torch.float32, device="cuda")  # noqa: E731
    with pytest.raises(mm.DimensionMismatch):
        mm.mul_(f(2), A, f(4))
    with pytest.raises(mm.DimensionMismatch):
        mm.mul_(f(3), A, f(3))
    with pytest.raises(mm.DimensionMismatch):
        mm.mul_(mm.linalg.colmajor(K, 2, 4), A, mm.linalg.colmajor(K, 3, 5))
    with pytest.raises(mm.DimensionMismatch):
        mm.mul_(mm.linalg.colmajor(K, 2, 4), A, mm.linalg.colmajor(K, 2, 4))
    E = mm.CuSparseMatrixCSR(K, [], [], [], 2, 3)
    c = torch.full((2,), 5.0, dtype=torch.float32, device="cuda")
    n0 = mm.lib().mk_launch_count(1)
    mm.mul_(c, E, f(3))
    assert mm.lib().mk_launch_count(0) == 0 and n0 >= 0
    np.testing.assert_array_equal(c.cpu().numpy(), [5.0, 5.0])
    C = mm.linalg.colmajor(K, 2, 2, fill=5.0)
    mm.mul_(C, E, mm.linalg.colmajor(K, 3, 2, fill=0.0))  # β = 0: fill!(C, 0̄) happens even when A is empty (:246-248)
    np.testing.assert_array_equal(C.cpu().numpy(), np.full((2, 2), -np.inf, np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sr", SEMIRINGS)
def test_sparse_vector_broadcast(torch, mm, sr, dtype):
    """test/test_linalg.jl:34-54: x (sparse) .* y and x ./ y against the dense computation."""
    K = _K(mm, sr, dtype)
    rng = np.random.default_rng(11)
    n = 1000
    idx = np.sort(rng.choice(np.arange(1, n + 1), 137, replace=False))
    xv = (rng.random(137) + 0.1).astype(dtype)
    y = (rng.random(n) + 0.1).astype(dtype)
    x = mm.CuSparseVector(K, idx, xv, n)
    cu_y = torch.from_numpy(y).cuda()
    dense = np.full(n, K.zero, dtype)
    for op, fn in ((K.mul, mm.elmul_), (K.div, mm.eldiv_)):
        want = dense.copy()
        want[idx - 1] = op(xv, y[idx - 1])
        out = torch.full((n,), 9.0, dtype=cu_y.dtype, device="cuda")
        got = fn(out, cu_y, x) if fn is mm.elmul_ else fn(out, x, cu_y)
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-6 if dtype == np.float32 else 1e-15)
    with pytest.raises(mm.DimensionMismatch):
        mm.elmul_(torch.zeros(n + 1, dtype=cu_y.dtype, device="cuda"), cu_y, x)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sr", SEMIRINGS)
def test_totalsum_totalcumsum_vs_dense(torch, mm, sr, dtype):
    """totalsum / totalcumsum (src/algorithms.jl:8-29) as host loops of mul! — against dense Float64 recursions in
    the probability domain (all three semirings agree there: Log = log of Prob; Tropical = max-product)."""
    K = _K(mm, sr, dtype)
    KL = mm.LogSemiring[np.float64]
    rng = np.random.default_rng(5)
    for base in (mm.graphs.hmm3(KL, 5)[0], mm.graphs.phone_loop(KL, n_phones=4)[0],
                 mm.graphs.numerator(KL, rng, 50, n_phones=6)[0]):
        S = base.nstates
        with np.errstate(divide="ignore"):
            a, T, w = np.exp(base.α), np.exp(base.T), np.exp(base.ω)  # probabilities
        src, dst = np.nonzero(T)
        conv = (lambda p: p) if K.code == 2 else (lambda p: np.log(p))
        with np.errstate(divide="ignore"):
            fsm = mm.FSM.from_arrays(K, S, src, dst, conv(T[src, dst]), np.flatnonzero(a), conv(a[a > 0]),
                                     np.flatnonzero(w), conv(w[w > 0]))
        for n in (1, 2, S, S + 5):
            v, terms = a.copy(), []
            for i in range(1, n + 1):
                if i > 1:
                    v = (T * v[:, None]).max(axis=0) if K.code == 1 else T.T @ v
                terms.append((v * w).max() if K.code == 1 else v @ w)
            want_sum = terms[-1]
            want_cum = max(terms) if K.code == 1 else sum(terms)
            got_sum, got_cum = mm.totalsum(fsm, n), mm.totalcumsum(fsm, n)
            if K.code != 2:
                got_sum, got_cum = np.exp(got_sum), np.exp(got_cum)
            rel = 2e-4 if dtype == np.float32 else 1e-9
            assert got_sum == pytest.approx(want_sum, rel=rel, abs=1e-300)
            assert got_cum == pytest.approx(want_cum, rel=rel, abs=1e-300)
            assert mm.totalweightsum(fsm, n) == pytest.approx(mm.totalcumsum(fsm, n), rel=rel, abs=1e-6)


def test_prob_semiring_host_ops(mm):
    """ProbSemiring scalars: ⊕ = +, ⊗ = *, ⊘ = /, 0̄ = 0, 1̄ = 1; graphs in it are rejected by the fused
    recursions with MK_ENOTSUP-style guidance, not silently mis-computed."""
    K = mm.ProbSemiring[np.float64]
    assert (K.zero, K.one) == (0.0, 1.0)
    assert K.add(0.25, 0.5) == 0.75 and K.mul(0.25, 0.5) == 0.125 and K.div(0.25, 0.5) == 0.5
    assert K.sum([0.1, 0.2, 0.3]) == pytest.approx(0.6)
    f = mm.FSM.from_json('{"semiring": "ProbSemiring{Float64}", "initstates": [[1, 1.0]], "arcs": [[1,1,0.5],[1,2,0.5]],'
                         ' "finalstates": [[2, 1.0]], "labels": [1, 2]}')
    assert f.K == K
    np.testing.assert_array_equal(f.T, [[0.5, 0.5], [0.0, 0.0]])
    r = mm.renorm(mm.FSM.from_pairs(K, [(1, 2.0)], [((1, 1), 1.0), ((1, 2), 3.0)], [(2, 5.0)], [1, 2]))
    np.testing.assert_allclose(r.T, [[0.25, 0.75], [0.0, 0.0]])
    np.testing.assert_allclose(r.ω, [0.0, 1.0])
    np.testing.assert_allclose(r.α, [1.0, 0.0])


def test_operator_entry_points_reject_bad_arguments_without_a_gpu(mm):
    """Argument validation happens before any device work, so it is checkable here: unknown semiring / dtype,
    DimensionMismatch, bad index base."""
    l = mm.lib()
    assert l.mk_spmv(7, 0, 1, 1, 0, None, None, None, 1, None, 1, None, 1, None) == 22
    assert l.mk_spmv(0, 5, 1, 1, 0, None, None, None, 1, None, 1, None, 1, None) == 22
    import ctypes
    rp = (ctypes.c_int32 * 3)(1, 1, 1)
    assert l.mk_spmv(0, 0, 2, 3, 0, rp, None, None, 1, None, 4, None, 2, None) == 22  # size(A,2) != length(b)
    assert b"DimensionMismatch" in l.mk_last_error()
    assert l.mk_spmv(0, 0, 2, 3, 0, rp, None, None, 2, None, 3, None, 2, None) == 22  # index_base
    assert l.mk_spmm(2, 1, 2, 3, 0, rp, None, None, 1, None, 3, 4, 3, None, 2, 5, 2, 0, None) == 22  # size(B,2) != size(C,2)
    assert l.mk_spvec_bcast(0, 0, 2, 4, 1, None, None, 1, None, 4, None, 4, None) == 22  # op
    if l.mk_device_count() == 0:
        assert l.mk_spmv(0, 0, 2, 3, 0, rp, None, None, 1, None, 3, None, 2, None) == 1000  # no CPU fallback
        assert b"no CPU fallback" in l.mk_last_error()


def test_csr_constructor_host_logic(mm):
    """``CuSparseMatrixCSR(K, I, J, V, m, n)`` = ``CuSparseMatrixCSR(sparse(I, J, V, m, n))``: rows sorted, columns ascending
    inside a row, 1-based Cint rowPtr / colVal exactly as CUDA.jl stores them, duplicates combined with the semiring's ⊕
    (Julia's `sparse` combines with +, i.e. ⊕ of the element type).  Host-side logic only (CPU tensors)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    m, n, nnz = 40, 23, 300
    I = rng.integers(1, m + 1, nnz)  # noqa: E741
    J = rng.integers(1, n + 1, nnz)
    V = rng.random(nnz) + 0.5
    for fam, comb in ((mm.ProbSemiring, np.add), (mm.LogSemiring, np.logaddexp), (mm.TropicalSemiring, np.maximum)):
        K = fam[np.float64]
        A = mm.CuSparseMatrixCSR(K, I, J, V, m, n, device="cpu")
        rowptr, colval, nzval = A.rowPtr.numpy(), A.colVal.numpy(), A.nzVal.numpy()
        assert rowptr.dtype == np.int32 and colval.dtype == np.int32 and rowptr[0] == 1 and rowptr[-1] == A.nnz + 1
        dense = np.full((m, n), np.nan)
        for i, j, v in zip(I - 1, J - 1, V):
            dense[i, j] = v if np.isnan(dense[i, j]) else comb(dense[i, j], v)
        got = np.full((m, n), np.nan)
        for r in range(m):
            cols = colval[rowptr[r] - 1:rowptr[r + 1] - 1]
            assert (np.diff(cols) > 0).all()  # ascending, no duplicates left
            got[r, cols - 1] = nzval[rowptr[r] - 1:rowptr[r + 1] - 1]
        np.testing.assert_allclose(got, dense, rtol=1e-14, equal_nan=True)
        if fam is mm.ProbSemiring:  # + is what scipy's COO -> CSR does too
            ref = sp.coo_matrix((V, (I - 1, J - 1)), shape=(m, n)).tocsr()
            ref.sort_indices()
            np.testing.assert_array_equal(rowptr - 1, ref.indptr)
            np.testing.assert_array_equal(colval - 1, ref.indices)
            np.testing.assert_allclose(nzval, ref.data, rtol=1e-14)
    E = mm.CuSparseMatrixCSR(mm.LogSemiring[np.float32], [], [], [], 3, 2, device="cpu")
    np.testing.assert_array_equal(E.rowPtr.numpy(), [1, 1, 1, 1])
    with pytest.raises(IndexError):
        mm.CuSparseMatrixCSR(mm.LogSemiring[np.float32], [4], [1], [0.0], 3, 2, device="cpu")


# ---- graph preparation: vcat / blockdiag / CSC <-> CSR / copy(transpose)  (test/test_linalg.jl:1-32, 56-86) ----------
SEMIRINGS = ["LogSemiring", "ProbSemiring", "TropicalSemiring"]


@pytest.mark.gpu
@pytest.mark.parametrize("pK", SEMIRINGS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_vcat_sparse_vectors(mm, pK, dtype):
    """test/test_linalg.jl:1-14: vcat(cu_sv, cu_sv) against the CPU vcat of sparsevec([1, 3], K[0.1, 0.2], 3)."""
    K = getattr(mm, pK)[dtype]
    sv = mm.CuSparseVector(K, [1, 3], [0.1, 0.2], 3)
    r = mm.vcat(sv, sv)
    assert isinstance(r, mm.CuSparseVector) and r.n == 6
    np.testing.assert_array_equal(r.nzInd.cpu().numpy(), [1, 3, 4, 6])
    np.testing.assert_array_equal(r.nzVal.cpu().numpy(), np.asarray([0.1, 0.2, 0.1, 0.2], dtype))
    # ragged: different lengths and an empty vector in the middle
    a = mm.CuSparseVector(K, [2, 5], [1.5, -2.0], 5)
    e = mm.CuSparseVector(K, [], [], 4)
    b = mm.CuSparseVector(K, [1], [7.0], 2)
    r = mm.vcat(a, e, b)
    assert r.n == 11
    np.testing.assert_array_equal(r.nzInd.cpu().numpy(), [2, 5, 10])
    np.testing.assert_array_equal(r.nzVal.cpu().numpy(), np.asarray([1.5, -2.0, 7.0], dtype))


@pytest.mark.gpu
@pytest.mark.parametrize("pK", SEMIRINGS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_blockdiag(mm, pK, dtype):
    """test/test_linalg.jl:16-32: blockdiag of three copies of sparse([1, 2, 2], [3, 2, 3], K[1, 2, 3], 3, 3), CSC and CSR,
    against scipy's block_diag; plus ragged shapes with an all-zero block."""
    import scipy.sparse as sp
    K = getattr(mm, pK)[dtype]
    I, J, V = [1, 2, 2], [3, 2, 3], [1.0, 2.0, 3.0]  # noqa: E741
    csr = mm.CuSparseMatrixCSR(K, I, J, V, 3, 3)
    csc = mm.CuSparseMatrixCSC(K, I, J, V, 3, 3)
    want = sp.block_diag([csr.to_scipy()] * 3).toarray()
    for blk in (csr, csc):
        r = mm.blockdiag(blk, blk, blk)
        assert type(r) is type(blk) and r.shape == (9, 9) and r.nnz == 9
        np.testing.assert_array_equal(r.to_scipy().toarray(), want)
    a = mm.CuSparseMatrixCSR(K, [1, 2, 2, 3, 4], [3, 1, 2, 1, 3], [1.0, 2.0, 3.0, 4.0, 5.0], 4, 3)
    z = mm.CuSparseMatrixCSR(K, [], [], [], 2, 5)
    b = mm.CuSparseMatrixCSR(K, [1], [2], [9.0], 1, 2)
    r = mm.blockdiag(a, z, b)
    assert r.shape == (7, 10)
    np.testing.assert_array_equal(r.to_scipy().toarray(), sp.block_diag([a.to_scipy(), z.to_scipy(), b.to_scipy()]).toarray())
    with pytest.raises(TypeError):
        mm.blockdiag(csr, csc)


@pytest.mark.gpu
@pytest.mark.parametrize("pK", SEMIRINGS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_csc_csr_conversion_and_transpose(mm, pK, dtype):
    """test/test_linalg.jl:56-86: CSC -> CSR -> CSC round trip and copy(transpose) of both storage kinds, on
    sparse([1, 2, 2, 3, 4], [3, 1, 2, 1, 3], K[1, 2, 3, 4, 5], 4, 3)."""
    K = getattr(mm, pK)[dtype]
    I, J, V = [1, 2, 2, 3, 4], [3, 1, 2, 1, 3], [1.0, 2.0, 3.0, 4.0, 5.0]  # noqa: E741
    csc = mm.CuSparseMatrixCSC(K, I, J, V, 4, 3)
    dense = csc.to_scipy().toarray()
    csr = mm.csr_from_csc(csc)
    np.testing.assert_array_equal(csr.to_scipy().toarray(), dense)
    back = mm.CuSparseMatrixCSC(csr)
    np.testing.assert_array_equal(back.to_scipy().toarray(), dense)
    for k in ("colPtr", "rowVal", "nzVal"):   # canonical arrays: the round trip is the identity
        np.testing.assert_array_equal(getattr(back, k).cpu().numpy(), getattr(csc, k).cpu().numpy())
    for M in (csc, csr):
        t = mm.copy_transpose(M)
        assert type(t) is type(M) and t.shape == (3, 4)
        np.testing.assert_array_equal(t.to_scipy().toarray(), dense.T)
    # indices come out ascending inside every segment (what CUSPARSE csr2csc gives the reference)
    t = mm.copy_transpose(csr)
    rp, cv = t.rowPtr.cpu().numpy() - 1, t.colVal.cpu().numpy()
    assert all(np.all(np.diff(cv[rp[r]:rp[r + 1]]) > 0) for r in range(3))


@pytest.mark.gpu
def test_transpose_of_a_graph_sized_matrix_feeds_mul(mm, orc):
    """The reference prepares T̂ᵀ with copy(T̂') (src/inference.jl:12) and multiplies with it: do the same on the synthetic
    denominator graph (empty rows, a 9 000-arc column) and check mul! on the materialised transpose against the oracle."""
    import torch
    K = mm.LogSemiring[np.float32]
    den, _ = mm.graphs.denominator(K, n_tokens=2000, n_pdf=100, seed=4)
    src, dst, w = den.arcs_hat()
    S = den.nstates_hat
    T = mm.CuSparseMatrixCSR(K, src + 1, dst + 1, w, S, S)      # T̂: rows = sources
    Tt = mm.copy_transpose(T)                                    # T̂ᵀ: rows = destinations
    ref = mm.CuSparseMatrixCSR(K, dst + 1, src + 1, w, S, S)
    for k in ("rowPtr", "colVal", "nzVal"):
        np.testing.assert_array_equal(getattr(Tt, k).cpu().numpy(), getattr(ref, k).cpu().numpy())
    rng = np.random.default_rng(1)
    x = rng.standard_normal(S).astype(np.float32)
    got = mm.mul_(torch.empty(S, dtype=torch.float32, device="cuda"), Tt, torch.from_numpy(x).cuda())
    order = np.lexsort((src, dst))
    rowptr = np.concatenate(([0], np.cumsum(np.bincount(dst, minlength=S))))
    want = orc.spmv(K.code, rowptr, src[order], w[order].astype(np.float64), x.astype(np.float64))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-5)


def test_prep_operators_reject_bad_arguments(mm):
    """No GPU needed: argument validation of the graph-preparation exports."""
    import ctypes as C
    from markov_b200 import _lib
    l = mm.lib()
    assert l.mk_vcat_spvec(7, 0, None, None, None, None, None, None, None) == _lib.MK_EINVAL       # dtype
    assert l.mk_vcat_spvec(0, 2, None, None, None, None, None, None, None) == _lib.MK_EINVAL       # null tables
    assert l.mk_sparse_transpose(0, -1, 3, 0, None, None, None, 1, None, None, None, None) == _lib.MK_EINVAL
    one = (C.c_int64 * 1)(3)
    neg = (C.c_int64 * 1)(-1)
    p = (C.c_void_p * 1)(8)
    out = C.c_void_p(8)
    assert l.mk_blockdiag(0, 1, p, p, p, one, one, neg, 1, out, out, out, None) == _lib.MK_EINVAL  # negative nnz
    assert l.mk_blockdiag(0, 1, p, p, p, one, one, one, 2, out, out, out, None) == _lib.MK_EINVAL  # index_base
