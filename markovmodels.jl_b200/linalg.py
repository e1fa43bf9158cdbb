# SPDX-License-Identifier: MIT
"""Operator level — host mirror of the GPU methods of ``/root/reference/src/linalg.jl`` on top of
``mk_spmv`` / ``mk_spmm`` / ``mk_spvec_bcast``:

    CuSparseMatrixCSR(K, I, J, V, m, n)     ``CuSparseMatrixCSR(adapt(CuArray, sparse(I, J, V, m, n)))``
    CuSparseVector(K, I, V, n)              ``adapt(CuArray, sparsevec(I, V, n))``
    mul_(c, A, b)                           ``mul!(c, A, b)``            src/linalg.jl:163-184
    mul_(C, A, B, α, β)                     ``mul!(C, A, B, α, β)``      src/linalg.jl:240-262
    elmul_(out, x, y) / eldiv_(out, x, y)   sparse-vector broadcasts     src/linalg.jl:287-338
    blockdiag(A, B, ...)                    ``blockdiag(X::CuSparseMatrixCSR...)`` / ``CSC...``   src/linalg.jl:73-131
    vcat(x, y, ...)                         ``vcat(X::CuSparseVector...)``                        src/linalg.jl:137-157
    CuSparseMatrixCSC(A) / CuSparseMatrixCSR(A)   the conversions                                 src/linalg.jl:12-49
    copy_transpose(A)                       ``copy(transpose(A))`` / ``copy(A')``                 src/linalg.jl:55-67

(``!`` is not an identifier character in Python, hence the trailing underscore.)  Arrays are torch
CUDA tensors of payload floats; matrices are column-major like Julia's (``colmajor`` allocates one;
any 2-D tensor with ``stride(0) == 1`` is accepted).  Indices are ``Cint`` and 1-based on the
device, exactly as CUDA.jl stores a ``CuSparseMatrixCSR`` — the constructors take Julia's 1-based
``I, J``.  The inference entry points do not go through these operators (their recursions are
fused); they exist for callers of ``mul!`` itself and for ``algorithms.totalsum``.
"""
import numpy as np

from . import _lib
from .semirings import SemiringType


def _torch():
    import torch
    return torch


def _tdtype(K):
    torch = _torch()
    return torch.float32 if K.dtype == np.float32 else torch.float64


def colmajor(K, m, n, fill=None, device="cuda"):
    """An ``m x n`` column-major payload matrix (``similar(dm, m, n)``), optionally filled."""
    torch = _torch()
    buf = torch.empty((n, m), dtype=_tdtype(K), device=device)
    if fill is not None:
        buf.fill_(float(fill))
    return buf.T


def to_colmajor(K, M, device="cuda"):
    """Upload a host ``m x n`` array as a column-major device matrix."""
    torch = _torch()
    M = np.asarray(M, K.dtype)
    return torch.from_numpy(np.ascontiguousarray(M.T)).to(device).T


class CuSparseMatrixCSR:
    """``CuSparseMatrixCSR{K}``: ``rowPtr`` (m+1), ``colVal`` (nnz), ``nzVal`` (nnz), 1-based ``Cint`` indices."""

    def __init__(self, K, I, J, V, m, n, device="cuda"):  # noqa: E741 - Julia's sparse(I, J, V, m, n)
        if not isinstance(K, SemiringType):
            raise TypeError("K must be a semiring type")
        torch = _torch()
        I = np.asarray(I, np.int64)  # noqa: E741
        J = np.asarray(J, np.int64)
        V = np.asarray(V, K.dtype)
        if I.size and (I.min() < 1 or I.max() > m or J.min() < 1 or J.max() > n):
            raise IndexError("sparse(I, J, V, m, n): index outside the matrix")
        order = np.lexsort((J, I))
        I, J, V = I[order], J[order], V[order]  # noqa: E741
        if I.size > 1:  # duplicates combine with ⊕, as `sparse` does with +
            first = np.concatenate(([True], (I[1:] != I[:-1]) | (J[1:] != J[:-1])))
            if not first.all():
                V = K.add_ufunc.reduceat(V, np.flatnonzero(first)).astype(K.dtype)
                I, J = I[first], J[first]  # noqa: E741
        rowptr = 1 + np.concatenate(([0], np.cumsum(np.bincount(I - 1, minlength=m))))
        self.K, self.shape = K, (int(m), int(n))
        self.rowPtr = torch.from_numpy(rowptr.astype(np.int32)).to(device)
        self.colVal = torch.from_numpy(J.astype(np.int32)).to(device)
        self.nzVal = torch.from_numpy(V).to(device)

    @property
    def nnz(self):
        return int(self.nzVal.numel())

    def size(self, d=None):
        return self.shape if d is None else self.shape[d - 1]

    @classmethod
    def from_arrays(cls, K, rowPtr, colVal, nzVal, m, n):
        """Wrap existing device arrays (1-based ``Cint`` indices) without copying."""
        A = cls.__new__(cls)
        A.K, A.shape, A.rowPtr, A.colVal, A.nzVal = K, (int(m), int(n)), rowPtr, colVal, nzVal
        return A

    def to_scipy(self):
        """Host copy as a scipy CSR matrix of payload floats (tests)."""
        import scipy.sparse as sp
        return sp.csr_matrix((self.nzVal.cpu().numpy(), self.colVal.cpu().numpy() - 1, self.rowPtr.cpu().numpy() - 1),
                             shape=self.shape)


class CuSparseMatrixCSC:
    """``CuSparseMatrixCSC{K}``: ``colPtr`` (n+1), ``rowVal`` (nnz), ``nzVal`` (nnz), 1-based ``Cint`` indices.
    ``CuSparseMatrixCSC(A::CuSparseMatrixCSR)`` converts (src/linalg.jl:32-49); ``CuSparseMatrixCSC(K, I, J, V, m, n)``
    is ``adapt(CuArray, sparse(I, J, V, m, n))``."""

    def __init__(self, K, *args, device="cuda"):
        if isinstance(K, CuSparseMatrixCSR):   # conversion
            A = K
            ptr, idx, val = _sparse_transpose(A.K, A.rowPtr, A.colVal, A.nzVal, A.shape[0], A.shape[1])
            self.K, self.shape, self.colPtr, self.rowVal, self.nzVal = A.K, A.shape, ptr, idx, val
            return
        I, J, V, m, n = args  # noqa: E741
        t = CuSparseMatrixCSR(K, J, I, V, n, m, device=device)   # CSC(A) holds the arrays of CSR(Aᵀ)
        self.K, self.shape, self.colPtr, self.rowVal, self.nzVal = K, (int(m), int(n)), t.rowPtr, t.colVal, t.nzVal

    @property
    def nnz(self):
        return int(self.nzVal.numel())

    def size(self, d=None):
        return self.shape if d is None else self.shape[d - 1]

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzVal.cpu().numpy(), self.rowVal.cpu().numpy() - 1, self.colPtr.cpu().numpy() - 1),
                             shape=self.shape)


def _sparse_transpose(K, ptr, idx, val, n_ptr, n_idx):
    torch = _torch()
    out_ptr = torch.empty(n_idx + 1, dtype=torch.int32, device=val.device)
    out_idx = torch.empty_like(idx)
    out_val = torch.empty_like(val)
    _lib.check(_lib.lib().mk_sparse_transpose(K.dtype_code, n_ptr, n_idx, int(val.numel()), ptr.data_ptr(), idx.data_ptr(),
                                              val.data_ptr(), 1, out_ptr.data_ptr(), out_idx.data_ptr(), out_val.data_ptr(),
                                              _stream()))
    return out_ptr, out_idx, out_val


def csr_from_csc(A):
    """``CuSparseMatrixCSR(A::CuSparseMatrixCSC)`` (src/linalg.jl:12-30)."""
    ptr, idx, val = _sparse_transpose(A.K, A.colPtr, A.rowVal, A.nzVal, A.shape[1], A.shape[0])
    return CuSparseMatrixCSR.from_arrays(A.K, ptr, idx, val, *A.shape)


def copy_transpose(A):
    """``copy(transpose(A))`` / ``copy(A')`` (src/linalg.jl:55-67): the materialised transpose, same storage kind.
    As in the reference, the arrays of CSR(A) are read as CSC(Aᵀ) and converted back."""
    m, n = A.shape
    if isinstance(A, CuSparseMatrixCSR):
        ptr, idx, val = _sparse_transpose(A.K, A.rowPtr, A.colVal, A.nzVal, m, n)
        return CuSparseMatrixCSR.from_arrays(A.K, ptr, idx, val, n, m)
    ptr, idx, val = _sparse_transpose(A.K, A.colPtr, A.rowVal, A.nzVal, n, m)
    T = CuSparseMatrixCSC.__new__(CuSparseMatrixCSC)
    T.K, T.shape, T.colPtr, T.rowVal, T.nzVal = A.K, (n, m), ptr, idx, val
    return T


def blockdiag(*Ms):
    """``blockdiag(X::CuSparseMatrixCSR{K}...)`` / ``blockdiag(X::CuSparseMatrixCSC{K}...)`` (src/linalg.jl:73-131)."""
    import ctypes as C
    torch = _torch()
    if not Ms:
        raise ValueError("blockdiag needs at least one matrix")
    csr = isinstance(Ms[0], CuSparseMatrixCSR)
    if any(isinstance(M, CuSparseMatrixCSR) != csr or M.K is not Ms[0].K for M in Ms):
        raise TypeError("blockdiag: all blocks must share storage kind and semiring")
    K, k = Ms[0].K, len(Ms)
    ptrs = [(M.rowPtr if csr else M.colPtr) for M in Ms]
    idxs = [(M.colVal if csr else M.rowVal) for M in Ms]
    dim_p = [M.shape[0 if csr else 1] for M in Ms]
    dim_i = [M.shape[1 if csr else 0] for M in Ms]
    nnz = [M.nnz for M in Ms]
    dev = Ms[0].nzVal.device
    out_ptr = torch.empty(sum(dim_p) + 1, dtype=torch.int32, device=dev)
    out_idx = torch.empty(sum(nnz), dtype=torch.int32, device=dev)
    out_val = torch.empty(sum(nnz), dtype=_tdtype(K), device=dev)
    vp, i64 = C.c_void_p * k, C.c_int64 * k
    _lib.check(_lib.lib().mk_blockdiag(K.dtype_code, k, vp(*[t.data_ptr() for t in ptrs]), vp(*[t.data_ptr() for t in idxs]),
                                       vp(*[M.nzVal.data_ptr() for M in Ms]), i64(*dim_p), i64(*dim_i), i64(*nnz), 1,
                                       out_ptr.data_ptr(), out_idx.data_ptr(), out_val.data_ptr(), _stream()))
    m, n = sum(M.shape[0] for M in Ms), sum(M.shape[1] for M in Ms)
    if csr:
        return CuSparseMatrixCSR.from_arrays(K, out_ptr, out_idx, out_val, m, n)
    R = CuSparseMatrixCSC.__new__(CuSparseMatrixCSC)
    R.K, R.shape, R.colPtr, R.rowVal, R.nzVal = K, (m, n), out_ptr, out_idx, out_val
    return R


def vcat(*xs):
    """``vcat(X::CuSparseVector{K}...)`` (src/linalg.jl:137-157)."""
    import ctypes as C
    torch = _torch()
    if not xs:
        raise ValueError("vcat needs at least one vector")
    K, k = xs[0].K, len(xs)
    if any(x.K is not K for x in xs):
        raise TypeError("vcat: all vectors must share the semiring")
    nnz = [int(x.nzVal.numel()) for x in xs]
    dev = xs[0].nzVal.device
    out_ind = torch.empty(sum(nnz), dtype=torch.int32, device=dev)
    out_val = torch.empty(sum(nnz), dtype=_tdtype(K), device=dev)
    vp, i64 = C.c_void_p * k, C.c_int64 * k
    _lib.check(_lib.lib().mk_vcat_spvec(K.dtype_code, k, vp(*[x.nzInd.data_ptr() for x in xs]),
                                        vp(*[x.nzVal.data_ptr() for x in xs]), i64(*[x.n for x in xs]), i64(*nnz),
                                        out_ind.data_ptr(), out_val.data_ptr(), _stream()))
    r = CuSparseVector.__new__(CuSparseVector)
    r.K, r.n, r.nzInd, r.nzVal = K, sum(x.n for x in xs), out_ind, out_val
    return r


class CuSparseVector:
    """``CuSparseVector{K}``: ``nzInd`` (1-based ``Cint``), ``nzVal``, length ``n``."""

    def __init__(self, K, I, V, n, device="cuda"):  # noqa: E741
        torch = _torch()
        I = np.asarray(I, np.int64)  # noqa: E741
        V = np.asarray(V, K.dtype)
        if I.size and (I.min() < 1 or I.max() > n):
            raise IndexError("sparsevec(I, V, n): index outside the vector")
        order = np.argsort(I, kind="stable")
        self.K, self.n = K, int(n)
        self.nzInd = torch.from_numpy(I[order].astype(np.int32)).to(device)
        self.nzVal = torch.from_numpy(V[order]).to(device)


def _stream():
    return _torch().cuda.current_stream().cuda_stream


def _check_payload(K, *tensors):
    dt = _tdtype(K)
    for t in tensors:
        if t.dtype != dt:
            raise TypeError(f"payload dtype {t.dtype} does not match {K}")
        if not t.is_cuda:
            raise TypeError("operator-level calls take CUDA tensors (libmarkov_b200 has no CPU fallback)")


def mul_(C, A, B, α=True, β=False):
    """``mul!(c, A, b)`` / ``mul!(C, A, B, α, β)`` for ``A::CuSparseMatrixCSR{K}`` (src/linalg.jl:163-184, 240-262).

    Vector form: ``c[i] = ⊕_k A[i,k] ⊗ b[k]``; an empty ``A`` launches nothing and leaves ``c`` as it is (:169).
    Matrix form: ``β == 1`` accumulates into ``C``, ``β == 0`` overwrites (:246-248); other ``β`` (``rmul!``) is
    not something the reference's callers use -> ``MK_ENOTSUP``.  ``α`` is ignored, as in the reference.
    Raises :class:`DimensionMismatch` like the reference's ``@boundscheck`` (:166-167, :242-244).  Returns ``C``."""
    K = A.K
    l = _lib.lib()
    _check_payload(K, C, B)
    m, n = A.shape
    if C.dim() == 1 and B.dim() == 1:
        if not (C.is_contiguous() and B.is_contiguous()):
            raise ValueError("vectors must be contiguous")
        _lib.check(l.mk_spmv(K.code, K.dtype_code, m, n, A.nnz, A.rowPtr.data_ptr(), A.colVal.data_ptr(),
                             A.nzVal.data_ptr(), 1, B.data_ptr(), B.numel(), C.data_ptr(), C.numel(), _stream()))
        return C
    if C.dim() != 2 or B.dim() != 2:
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, "mul!: C and B must both be vectors or both be matrices")
    if β not in (0, 1, False, True):
        raise _lib.MarkovError(_lib.MK_ENOTSUP, "mul!: β must be 0 or 1")
    for t in (C, B):
        if t.shape[0] > 1 and t.stride(0) != 1:
            raise ValueError("matrices must be column-major (stride(0) == 1); see linalg.colmajor")
    ldb = B.stride(1) if B.shape[1] > 1 else max(B.shape[0], 1)
    ldc = C.stride(1) if C.shape[1] > 1 else max(C.shape[0], 1)
    _lib.check(l.mk_spmm(K.code, K.dtype_code, m, n, A.nnz, A.rowPtr.data_ptr(), A.colVal.data_ptr(),
                         A.nzVal.data_ptr(), 1, B.data_ptr(), B.shape[0], B.shape[1], ldb, C.data_ptr(), C.shape[0],
                         C.shape[1], ldc, 1 if β else 0, _stream()))
    return C


def _bcast(op, out, x, y):
    K = x.K
    _check_payload(K, out, y)
    _lib.check(_lib.lib().mk_spvec_bcast(K.code, K.dtype_code, op, x.n, int(x.nzVal.numel()), x.nzInd.data_ptr(),
                                         x.nzVal.data_ptr(), 1, y.data_ptr(), y.numel(), out.data_ptr(), out.numel(),
                                         _stream()))
    return out


def elmul_(out, x, y):
    """``elmul!(out, x::AbstractVector, y::CuSparseVector)`` (src/linalg.jl:292): ``out .= x .* y``."""
    if isinstance(x, CuSparseVector):
        x, y = y, x
    return _bcast(0, out, y, x)


def eldiv_(out, x, y):
    """``eldiv!(out, x::CuSparseVector, y::AbstractVector)`` (src/linalg.jl:294): ``out .= x ./ y``."""
    return _bcast(1, out, x, y)
