# SPDX-License-Identifier: MIT
"""FSM data model — host mirror of ``/root/reference/src/fsm.jl`` and the batching half of
``src/fsmops.jl``.

``FSM{K,L}`` (src/fsm.jl:7-17) holds the *extended* graph: ``α̂ = [α; 0̄]`` and
``T̂ = [T ω; 0̄ᵀ 1̄]`` with a phony final state appended last (src/fsm.jl:19-28), plus labels
``λ``.  The reference stores ``T̂`` as ``SparseMatrixCSC{K,Int64}``; this class keeps the same
three arrays (``colptr``, ``rowval``, ``nzval`` — column = destination state) 0-based, which is
what ``mk_graph_create`` consumes, so a Julia shim can hand its arrays over without reshaping.

Only what the inference path needs is mirrored: construction (arrays, pair lists, JSON),
``.α/.T/.ω`` views, ``nstates``, ``renorm`` (used to build test graphs), ``union`` and
``rawunion``.  The offline graph algebra (cat, compose, determinize, minimize, …) is out of
scope.
"""
import json

import numpy as np

from .semirings import LogSemiring, ProbSemiring, SemiringType, TropicalSemiring


def _coo_to_csc(K, n_rows, n_cols, row, col, val):
    """``sparse(I, J, V, m, n)``: duplicates are combined with the semiring ⊕."""
    row = np.asarray(row, np.int64)
    col = np.asarray(col, np.int64)
    val = np.asarray(val, K.dtype)
    order = np.lexsort((row, col))
    row, col, val = row[order], col[order], val[order]
    if row.size:
        if row.min() < 0 or row.max() >= n_rows or col.min() < 0 or col.max() >= n_cols:
            raise IndexError("arc endpoint outside the state range")
        dup = (row[1:] == row[:-1]) & (col[1:] == col[:-1])
        if dup.any():
            keep = np.concatenate(([True], ~dup))
            starts = np.flatnonzero(keep)
            out = val[starts].copy()
            for k, s in enumerate(starts):
                e = starts[k + 1] if k + 1 < len(starts) else len(val)
                for v in val[s + 1:e]:
                    out[k] = K.add(out[k], v)
            row, col, val = row[keep], col[keep], out
    colptr = np.zeros(n_cols + 1, np.int64)
    np.add.at(colptr, col + 1, 1)
    np.cumsum(colptr, out=colptr)
    return colptr, row.copy(), val.copy()


class FSM:
    """``FSM{K,L}``: extended graph (α̂, T̂, λ) — src/fsm.jl:7-17.

    Attributes (all 0-based):
      K            semiring descriptor
      nstates_hat  Ŝ = S + 1 (phony final state is index Ŝ-1)
      init_idx/init_w   the stored entries of α̂ (SparseVector nzind/nzval)
      colptr/rowval/nzval   T̂ as CSC: column = destination, rowval = source (ascending)
      labels       λ, one per real state
    """

    def __init__(self, K, nstates_hat, init_idx, init_w, colptr, rowval, nzval, labels=None, parts=None):
        if not isinstance(K, SemiringType):
            raise TypeError("K must be a semiring type such as LogSemiring[np.float32]")
        self.K = K
        self.nstates_hat = int(nstates_hat)
        self.init_idx = np.ascontiguousarray(init_idx, np.int64)
        self.init_w = np.ascontiguousarray(init_w, K.dtype)
        self.colptr = np.ascontiguousarray(colptr, np.int64)
        self.rowval = np.ascontiguousarray(rowval, np.int64)
        self.nzval = np.ascontiguousarray(nzval, K.dtype)
        self.labels = list(labels) if labels is not None else list(range(1, self.nstates_hat))
        # rawunion remembers its operands so the batch can be described without materialising
        # the block-diagonal matrix on the device (identical operands are compiled once)
        self.parts = parts
        if self.colptr.shape != (self.nstates_hat + 1,):
            raise ValueError("DimensionMismatch: colptr length")

    # ---- constructors -------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, K, nstates, src, dst, w, init_idx, init_w, final_idx, final_w, labels=None):
        """``FSM(α, T, ω, λ)`` (src/fsm.jl:19-28) from 0-based arrays: builds T̂ and α̂."""
        S = int(nstates)
        src = np.asarray(src, np.int64)
        dst = np.asarray(dst, np.int64)
        final_idx = np.asarray(final_idx, np.int64)
        if final_idx.size and (final_idx.min() < 0 or final_idx.max() >= S):
            raise IndexError("final state outside the state range")
        if src.size and (max(src.max(), dst.max()) >= S or min(src.min(), dst.min()) < 0):
            raise IndexError("arc endpoint outside the state range")
        rows = np.concatenate([src, final_idx, [S]])                       # T | ω | phony self-loop
        cols = np.concatenate([dst, np.full(final_idx.size, S), [S]])
        vals = np.concatenate([np.asarray(w, K.dtype), np.asarray(final_w, K.dtype), [K.one]])
        colptr, rowval, nzval = _coo_to_csc(K, S + 1, S + 1, rows, cols, vals)
        init_idx = np.asarray(init_idx, np.int64)
        order = np.argsort(init_idx, kind="stable")
        return cls(K, S + 1, init_idx[order], np.asarray(init_w, K.dtype)[order], colptr, rowval, nzval, labels)

    @classmethod
    def from_pairs(cls, K, initws, arcs, finalws, labels):
        """The pair-list constructor (src/fsm.jl:50-71): 1-based state ids as in the reference,
        ``initws = [(s, w)]``, ``arcs = [((s, d), w)]``, ``finalws = [(s, w)]``."""
        S = len(labels)
        src = [a[0][0] - 1 for a in arcs]
        dst = [a[0][1] - 1 for a in arcs]
        w = [a[1] for a in arcs]
        return cls.from_arrays(K, S, src, dst, w, [s - 1 for s, _ in initws], [x for _, x in initws],
                               [s - 1 for s, _ in finalws], [x for _, x in finalws], labels)

    @classmethod
    def from_json(cls, s):
        """``FSM(::AbstractString)`` (src/fsm.jl:73-82); keys semiring, initstates, arcs,
        finalstates, labels (example test/test_fsms.jl:42-51)."""
        data = json.loads(s)
        K = _parse_semiring(data["semiring"])
        return cls.from_pairs(K, [(a, K(b)) for a, b in data["initstates"]],
                              [((a, b), K(c)) for a, b, c in data["arcs"]],
                              [(a, K(b)) for a, b in data["finalstates"]], list(data["labels"]))

    def astype(self, K):
        """The same graph with another semiring type / payload precision (``convert`` in the
        reference's tests, test/test_algorithms.jl:277)."""
        parts = None
        if self.parts is not None:  # a rawunion stays batchable: convert every operand once (identical operands stay identical)
            done = {}
            parts = [done.setdefault(id(p), p.astype(K)) for p in self.parts]
        return FSM(K, self.nstates_hat, self.init_idx, self.init_w.astype(K.dtype), self.colptr, self.rowval,
                   self.nzval.astype(K.dtype), self.labels, parts=parts)

    # ---- views (src/fsm.jl:30-40) --------------------------------------------------------------
    @property
    def nstates(self):
        return self.nstates_hat - 1

    @property
    def nnz_hat(self):
        return int(self.nzval.size)

    def arcs_hat(self):
        """(src, dst, w) triplets of T̂, 0-based, column-major order."""
        dst = np.repeat(np.arange(self.nstates_hat, dtype=np.int64), np.diff(self.colptr))
        return self.rowval, dst, self.nzval

    @property
    def α(self):
        v = np.full(self.nstates, self.K.zero, self.K.dtype)
        m = self.init_idx < self.nstates
        v[self.init_idx[m]] = self.init_w[m]
        return v

    @property
    def ω(self):
        S = self.nstates
        v = np.full(S, self.K.zero, self.K.dtype)
        a, b = self.colptr[S], self.colptr[S + 1]
        src = self.rowval[a:b]
        m = src < S
        v[src[m]] = self.nzval[a:b][m]
        return v

    @property
    def T(self):
        """Dense S x S payload matrix (0̄ where no arc) — small graphs / tests only."""
        S = self.nstates
        M = np.full((S, S), self.K.zero, self.K.dtype)
        src, dst, w = self.arcs_hat()
        m = (src < S) & (dst < S)
        M[src[m], dst[m]] = w[m]
        return M

    def __repr__(self):
        return f"FSM{{{self.K}}}(nstates={self.nstates}, nnz(T̂)={self.nnz_hat})"


def _parse_semiring(name):
    fam = {"LogSemiring": LogSemiring, "TropicalSemiring": TropicalSemiring,
           "ProbSemiring": ProbSemiring}.get(name.split("{")[0])
    if fam is None:
        raise ValueError(f"unsupported semiring {name!r} (the path covers Log/Tropical/Prob)")
    return fam[np.float32 if "Float32" in name else np.float64]


def nstates(fsm):
    """``nstates(m::FSM)`` (src/fsm.jl:84)."""
    return fsm.nstates


def renorm(fsm):
    """``renorm`` (src/fsmops.jl:71-80): scale every state's outgoing weights (arcs + final) to
    ⊕-sum to 1̄ and the initial weights likewise."""
    K = fsm.K
    S = fsm.nstates
    src, dst, w = fsm.arcs_hat()
    real = src < S  # drop the phony self-loop
    src, dst, w = src[real], dst[real], w[real]
    tot = np.full(S, K.zero, K.dtype)
    K.add_ufunc.at(tot, src, w)
    w = K.div(w, tot[src])
    fin = dst == S
    a = fsm.init_w
    return FSM.from_arrays(K, S, src[~fin], dst[~fin], w[~fin], fsm.init_idx, K.div(a, K.sum(a)),
                           src[fin], w[fin], fsm.labels)


def rawunion(fsm1, *fsms):
    """``rawunion`` (src/fsmops.jl:28-36): stack the extended storages — ``vcat`` of α̂,
    ``blockdiag`` of T̂ — keeping one phony final state per operand.  The result is "several
    independent FSMs packed in a single structure"; it remembers its operands so the device
    batch never materialises the block-diagonal matrix."""
    parts = []
    for f in (fsm1,) + fsms:
        if f.K != fsm1.K:
            raise TypeError("rawunion: FSMs must share the semiring type")
        parts.extend(f.parts if f.parts is not None else [f])
    off = np.cumsum([0] + [p.nstates_hat for p in parts])
    nnz_off = np.cumsum([0] + [p.nnz_hat for p in parts])
    colptr = np.concatenate([[0]] + [p.colptr[1:] + nnz_off[k] for k, p in enumerate(parts)])
    rowval = np.concatenate([p.rowval + off[k] for k, p in enumerate(parts)])
    nzval = np.concatenate([p.nzval for p in parts])
    init_idx = np.concatenate([p.init_idx + off[k] for k, p in enumerate(parts)])
    init_w = np.concatenate([p.init_w for p in parts])
    labels = [l for p in parts for l in p.labels]
    return FSM(fsm1.K, int(off[-1]), init_idx, init_w, colptr, rowval, nzval, labels, parts=parts)


def union(fsm1, *fsms):
    """``union`` (src/fsmops.jl:8-17): block-diagonal T with ONE shared phony final state.  Not
    valid for batched inference — kept for API parity."""
    K = fsm1.K
    all_f = (fsm1,) + fsms
    off = np.cumsum([0] + [f.nstates for f in all_f])
    S = int(off[-1])
    src, dst, w, fi, fw, ii, iw, labels = [], [], [], [], [], [], [], []
    for k, f in enumerate(all_f):
        if f.K != K:
            raise TypeError("union: FSMs must share the semiring type")
        s, d, x = f.arcs_hat()
        real = s < f.nstates
        s, d, x = s[real], d[real], x[real]
        fin = d == f.nstates
        src.append(s[~fin] + off[k]); dst.append(d[~fin] + off[k]); w.append(x[~fin])
        fi.append(s[fin] + off[k]); fw.append(x[fin])
        ii.append(f.init_idx + off[k]); iw.append(f.init_w)
        labels.extend(f.labels)
    cat = np.concatenate
    return FSM.from_arrays(K, S, cat(src), cat(dst), cat(w), cat(ii), cat(iw), cat(fi), cat(fw), labels)
