# SPDX-License-Identifier: MIT
"""oracle — CPU restatement of the reference's inference path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``markovmodels.jl_b200``) never
does.  PARITY STATUS: "parity unpinned" against a running reference (no Julia here; Semirings.jl
un-vendored) — pinned to the literal known-answer vectors of the reference tree only, see
``oracle.cpp`` and ``tests/test_oracle_golden.py``.

Two independent restatements:
  * ``oracle.cpp`` (C++, f32/f64, OpenMP over utterances) — follows src/inference.jl line by
    line including the CPU sparse ``mul!`` accumulation order; the checker and the CPU baseline;
  * ``dense_forward_backward`` below (numpy float64, dense ``logsumexp``) — the pattern of the
    reference's own stale test oracle, test/test_algorithms.jl:28-63; cross-checks the C++ one.

``oracle/_ref`` (the compiled reference) does not exist: the reference is pure Julia and cannot
be built with gcc/g++.
"""
import ctypes as C
import hashlib
import os
import platform
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "oracle.cpp")
_lib = None


def _cpu_signature():
    """The build is -march=native, so the library is tied to the host CPU: the file name carries
    a hash of the CPU model and its ISA flags, and a box with another CPU compiles its own copy."""
    model, flags = platform.machine(), ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and "@" not in model:
                model += "@" + line.split(":", 1)[1].strip()
            elif line.startswith("flags"):
                flags = " ".join(sorted(line.split(":", 1)[1].split()))
                break
    except OSError:
        pass
    return hashlib.sha1((model + "|" + flags).encode()).hexdigest()[:12]


LIB = os.path.join(_HERE, "_build", f"liboracle_{_cpu_signature()}.so")


CXXFLAGS = ["-O3", "-march=native", "-fopenmp", "-fno-fast-math", "-ffp-contract=off", "-std=c++17"]


def build_flags():
    """The compiler line of the CPU legs, for the bench JSON."""
    return "g++ " + " ".join(CXXFLAGS)


def build(force=False):
    """g++ -O3 -march=native -fopenmp for THIS host's CPU; no -ffast-math: -Inf arithmetic
    must be IEEE (and -O3 -march=native without it keeps the scalar ⊕ order of oracle.cpp)."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["g++"] + CXXFLAGS + ["-shared", "-fPIC", "-o", LIB + ".tmp", SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)  # (atomic: xdist workers / torchrun ranks may build at the same time)
    return LIB


class _OrcGraph(C.Structure):
    _fields_ = [("S", C.c_int64), ("Dhat", C.c_int64), ("n_init", C.c_int64),
                ("in_ptr", C.c_void_p), ("in_src", C.c_void_p), ("in_w", C.c_void_p),
                ("out_ptr", C.c_void_p), ("out_dst", C.c_void_p), ("out_w", C.c_void_p),
                ("init_idx", C.c_void_p), ("init_w", C.c_void_p), ("state2pdf", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB)
        l.orc_num_threads.restype = C.c_int
        l.orc_logaddexp_f64.restype = C.c_double
        l.orc_logaddexp_f64.argtypes = [C.c_double, C.c_double]
        l.orc_logaddexp_f32.restype = C.c_float
        l.orc_logaddexp_f32.argtypes = [C.c_float, C.c_float]
        l.orc_spmv_f64.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 5
        l.orc_alpha_beta.argtypes = [C.c_int, C.c_int, C.POINTER(_OrcGraph), C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        l.orc_pdfposteriors.argtypes = [C.c_int, C.c_int, C.c_int64, C.POINTER(C.POINTER(_OrcGraph)), C.c_void_p,
                                        C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int]
        l.orc_bestpath.argtypes = [C.c_int, C.c_int64, C.POINTER(C.POINTER(_OrcGraph)), C.c_void_p, C.c_int64,
                                   C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = l
    return _lib


def num_threads():
    """Host threads the CPU legs use: the cores this process may run on.  NOT omp_get_max_threads():
    torchrun exports OMP_NUM_THREADS=1, which would silently turn the all-core baseline into a 1-core one."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class OracleGraph:
    """Host arrays of one extended graph in both orientations.  Built from the duck-typed fields
    of a host FSM (K, nstates_hat, colptr, rowval, nzval, init_idx, init_w) plus the 0-based
    state→pdf map of its real states."""

    def __init__(self, fsm, pdfids, numpdf):
        K = fsm.K
        self.K = K
        self.S = S = int(fsm.nstates_hat)
        self.numpdf = int(numpdf)
        self.in_ptr = np.ascontiguousarray(fsm.colptr, np.int64)
        self.in_src = np.ascontiguousarray(fsm.rowval, np.int64)
        self.in_w = np.ascontiguousarray(fsm.nzval, K.dtype)
        dst = np.repeat(np.arange(S, dtype=np.int64), np.diff(self.in_ptr))
        order = np.lexsort((dst, self.in_src))  # by source, destinations ascending
        self.out_dst = np.ascontiguousarray(dst[order])
        self.out_w = np.ascontiguousarray(self.in_w[order])
        self.out_ptr = np.zeros(S + 1, np.int64)
        np.add.at(self.out_ptr, self.in_src + 1, 1)
        np.cumsum(self.out_ptr, out=self.out_ptr)
        self.init_idx = np.ascontiguousarray(fsm.init_idx, np.int64)
        self.init_w = np.ascontiguousarray(fsm.init_w, K.dtype)
        pdfids = np.asarray(pdfids, np.int64)
        assert pdfids.shape == (S - 1,)
        self.state2pdf = np.ascontiguousarray(np.concatenate([pdfids, [numpdf]]), np.int32)
        g = _OrcGraph()
        g.S, g.Dhat, g.n_init = S, numpdf + 1, self.init_idx.size
        for name in ("in_ptr", "in_src", "in_w", "out_ptr", "out_dst", "out_w", "init_idx", "init_w", "state2pdf"):
            setattr(g, name, getattr(self, name).ctypes.data)
        self.c = g


def _prep(graphs, V, seqlengths):
    K = graphs[0].K
    V = np.ascontiguousarray(V, K.dtype)  # (B, T, D): per utterance D x T column-major
    B, T, D = V.shape
    assert len(graphs) == B
    arr = (C.POINTER(_OrcGraph) * B)(*[C.pointer(g.c) for g in graphs])
    sl = None if seqlengths is None else np.ascontiguousarray(seqlengths, np.int32)
    return K, V, B, T, D, arr, sl


def pdfposteriors(graphs, V_btd, seqlengths=None, threads=0):
    """src/inference.jl:145-161 over a batch.  ``V_btd``: (B, T, D) array (frame-major per
    utterance = the reference's D x T column-major matrices).  Returns (post (B, D, T) with b
    fastest in memory, ttl (B,))."""
    K, V, B, T, D, arr, sl = _prep(graphs, V_btd, seqlengths)
    post = np.zeros((T, D, B), K.dtype)
    ttl = np.zeros(B, K.dtype)
    rc = lib().orc_pdfposteriors(K.dtype_code, K.code, B, arr, V.ctypes.data, D, T, D,
                                 None if sl is None else sl.ctypes.data, post.ctypes.data, ttl.ctypes.data,
                                 threads if threads > 0 else num_threads())
    if rc:
        raise ValueError("DimensionMismatch in oracle pdfposteriors")
    return post.transpose(2, 1, 0), ttl


def bestpath(graphs, V_btd, seqlengths=None, threads=0):
    K, V, B, T, D, arr, sl = _prep(graphs, V_btd, seqlengths)
    path = np.zeros((B, T), np.int32)
    score = np.zeros(B, K.dtype)
    rc = lib().orc_bestpath(K.dtype_code, B, arr, V.ctypes.data, D, T, D,
                            None if sl is None else sl.ctypes.data, path.ctypes.data, score.ctypes.data,
                            threads if threads > 0 else num_threads())
    if rc:
        raise ValueError("DimensionMismatch in oracle bestpath")
    return path, score


def alpha_beta(graph, V_td, seqlength=None, want_alpha=True, want_beta=True):
    """αrecursion / βrecursion (src/inference.jl:62-74, 99-110) for one utterance; returns Ŝ x N̂
    arrays (column-major, like the reference)."""
    K = graph.K
    V = np.ascontiguousarray(V_td, K.dtype)
    T, D = V.shape
    L = T if seqlength is None else int(seqlength)
    A = np.zeros((T + 1, graph.S), K.dtype) if want_alpha else None
    Bm = np.zeros((T + 1, graph.S), K.dtype) if want_beta else None
    rc = lib().orc_alpha_beta(K.dtype_code, K.code, C.byref(graph.c), V.ctypes.data, D, T, D, L,
                              A.ctypes.data if want_alpha else None, Bm.ctypes.data if want_beta else None)
    if rc:
        raise ValueError("DimensionMismatch in oracle alpha_beta")
    return (A.T if want_alpha else None), (Bm.T if want_beta else None)


def logaddexp(x, y, dtype=np.float64):
    l = lib()
    return l.orc_logaddexp_f64(x, y) if np.dtype(dtype) == np.float64 else l.orc_logaddexp_f32(x, y)


def spmv(semiring, rowptr, colidx, w, b):
    """Semiring CSR SpMV on float64 payloads (0 Log, 1 Tropical, 2 Prob)."""
    rowptr = np.ascontiguousarray(rowptr, np.int64)
    colidx = np.ascontiguousarray(colidx, np.int64)
    w = np.ascontiguousarray(w, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    c = np.zeros(rowptr.size - 1)
    lib().orc_spmv_f64(semiring, rowptr.size - 1, rowptr.ctypes.data, colidx.ctypes.data, w.ctypes.data,
                       b.ctypes.data, c.ctypes.data)
    return c


# ---------------------------------------------------------------------------------------------
# independent dense float64 restatement (pattern of test/test_algorithms.jl:28-63)
# ---------------------------------------------------------------------------------------------
def _lse(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore"):
        return np.squeeze(m, axis) + np.log(np.sum(np.exp(x - m), axis=axis))


def dense_forward_backward(fsm, pdfids, V_dt, seqlength=None):
    """Dense log-domain forward-backward on the un-extended graph (α, T, ω), float64.
    Returns (γ pdf posteriors D x L, logZ)."""
    S = fsm.nstates
    A = fsm.T.astype(np.float64)
    init = fsm.α.astype(np.float64)
    final = fsm.ω.astype(np.float64)
    V = np.asarray(V_dt, np.float64)
    D, T = V.shape
    L = T if seqlength is None else int(seqlength)
    lhs = V[np.asarray(pdfids), :L]
    la = np.full((S, L), -np.inf)
    la[:, 0] = init + lhs[:, 0]
    for n in range(1, L):
        la[:, n] = lhs[:, n] + _lse(A + la[:, n - 1][:, None], 0)
    lb = np.full((S, L), -np.inf)
    lb[:, -1] = final
    for n in range(L - 2, -1, -1):
        lb[:, n] = _lse(A + (lb[:, n + 1] + lhs[:, n + 1])[None, :], 1)
    lg = la + lb
    logz = _lse(lg, 0)
    with np.errstate(invalid="ignore"):
        gs = np.exp(lg - logz[None, :])
    post = np.zeros((D, L))
    np.add.at(post, np.asarray(pdfids), gs)
    return post, float(np.min(logz))


def dense_viterbi(fsm, pdfids, V_dt, seqlength=None):
    """Dense max-plus Viterbi with first-maximum tie breaking; returns (1-based path, score)."""
    S = fsm.nstates
    A = fsm.T.astype(np.float64)
    V = np.asarray(V_dt, np.float64)
    L = V.shape[1] if seqlength is None else int(seqlength)
    lhs = V[np.asarray(pdfids), :L]
    d = fsm.α.astype(np.float64) + lhs[:, 0]
    psi = np.zeros((S, L), np.int64)
    for n in range(1, L):
        c = A + d[:, None]
        psi[:, n] = np.argmax(c, axis=0)
        d = c.max(axis=0) + lhs[:, n]
    fin = d + fsm.ω.astype(np.float64)
    s = int(np.argmax(fin))
    path = [s]
    for n in range(L - 1, 0, -1):
        s = int(psi[s, n])
        path.append(s)
    return np.array(path[::-1]) + 1, float(fin.max())
