# SPDX-License-Identifier: MIT
"""Inference API — host mirror of ``/root/reference/src/inference.jl`` on top of the C ABI.

Same names and argument meaning as the reference:

    compile(fsm, Ĉ)                 src/inference.jl:11-12   -> CompiledFSM (device resident)
    batch(cfsm1, cfsms...)          src/inference.jl:28-36   -> BatchedFSM
    expand(V, seqlength)            src/inference.jl:54-60
    αrecursion / βrecursion         src/inference.jl:62-74, 99-110
    pdfposteriors(fsm, V̂s, Ĉs)      src/inference.jl:145-161 (and pdfposteriors2 :164-180)
    bestpath                        absent from 0.10.0; historical signature
                                    test/test_algorithms.jl:279-281

The reference selects the device by array type (``adapt(CuArray, ·)``); here ``compile`` puts
the graph on the current CUDA device and the recursions run in ``libmarkov_b200.so``.  torch is
used for device memory and streams only.  Emissions may be passed

  * the reference's way: a list ``V̂s`` of expanded ``D̂ x N̂`` matrices (one per utterance), or
  * un-expanded as one ``(B, D, T)`` array (any strides, e.g. ``net_out.permute(0, 2, 1)``) plus
    ``seqlengths`` — ``expand`` is then applied inside the kernels.

torch CUDA tensors stay on the device (asynchronous on the current stream); numpy arrays go
through the ``*_host`` entry points (H2D + compute + D2H).
"""
import ctypes as C

import numpy as np

from . import _lib
from .fsm import FSM
from .semirings import MK_TROPICAL


def _torch():
    import torch
    return torch


# ---------------------------------------------------------------------------------------------
# Ĉ handling
# ---------------------------------------------------------------------------------------------
class StateMap:
    """Ĉ: the (S+1) x (D+1) state→pdf matrix with exactly one 1̄ per row
    (examples/prepare-lfmmi-graphs.jl:15-23), stored as ``state2pdf`` (0-based; the phony
    state maps to the phony pdf ``numpdf``)."""

    def __init__(self, state2pdf, numpdf):
        self.state2pdf = np.ascontiguousarray(state2pdf, np.int32)
        self.numpdf = int(numpdf)
        if self.state2pdf[-1] != self.numpdf:
            raise ValueError("the phony final state must map to the phony pdf")

    @property
    def n_pdf_hat(self):
        return self.numpdf + 1

    def dense(self, K):
        M = np.full((self.state2pdf.size, self.numpdf + 1), K.zero, K.dtype)
        M[np.arange(self.state2pdf.size), self.state2pdf] = K.one
        return M


def statemap(fsm, numpdf, pdfids=None):
    """``statemap(fsm, numpdf)`` (examples/prepare-lfmmi-graphs.jl:15-23).  ``pdfids`` are the
    0-based pdf ids of the real states (default: the last element of each state's label, which
    the reference uses, minus one)."""
    if pdfids is None:
        pdfids = [(l[-1] if isinstance(l, (tuple, list)) else l) - 1 for l in fsm.labels]
    pdfids = np.asarray(pdfids, np.int64)
    if pdfids.shape != (fsm.nstates,):
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, "one pdf id per state expected")
    if pdfids.size and (pdfids.min() < 0 or pdfids.max() >= numpdf):
        raise IndexError("pdf id outside [0, numpdf)")
    return StateMap(np.concatenate([pdfids, [numpdf]]), numpdf)


def _as_statemap(fsm, Ĉ):
    """Ĉ as a state→pdf map.  A dense Ĉ holds the payloads of ``K``: its stored entries are the ones that differ
    from ``zero(K)`` (-Inf for Log / Tropical, 0 for Prob)."""
    if isinstance(Ĉ, StateMap):
        return Ĉ
    if hasattr(Ĉ, "tocsr"):  # scipy sparse with stored 1̄ entries
        csr = Ĉ.tocsr()
        if not np.all(np.diff(csr.indptr) == 1):
            raise _lib.MarkovError(_lib.MK_ENOTSUP, "Ĉ must have exactly one stored entry per row")
        return StateMap(csr.indices, csr.shape[1] - 1)
    Ĉ = np.asarray(Ĉ)
    if Ĉ.ndim == 1:
        return StateMap(Ĉ, int(Ĉ[-1]))
    fin = Ĉ != fsm.K.zero
    if not np.all(fin.sum(axis=1) == 1):
        raise _lib.MarkovError(_lib.MK_ENOTSUP, "Ĉ must have exactly one 1̄ per row")
    return StateMap(fin.argmax(axis=1), Ĉ.shape[1] - 1)


# ---------------------------------------------------------------------------------------------
# compile / batch
# ---------------------------------------------------------------------------------------------
class CompiledFSM:
    """``CompiledFSM{K}`` (src/inference.jl:3-9): (α̂, T̂, T̂ᵀ, Ĉ, Ĉᵀ) resident on one GPU."""

    def __init__(self, fsm, smap, device=None):
        if fsm.parts is not None:
            raise TypeError("compile the operands of a rawunion separately and batch() them")
        if smap.state2pdf.size != fsm.nstates_hat:
            raise _lib.DimensionMismatch(_lib.MK_EINVAL, "Ĉ has a different number of rows than T̂")
        self.K = fsm.K
        self.nstates_hat = fsm.nstates_hat
        self.n_pdf_hat = smap.n_pdf_hat
        self.fsm, self.smap = fsm, smap
        if device is None:
            torch = _torch()
            device = torch.cuda.current_device() if torch.cuda.is_available() else -1
        h = C.c_void_p()
        l = _lib.lib()
        _lib.check(l.mk_graph_create(
            C.byref(h), fsm.K.code, fsm.K.dtype_code, fsm.nstates_hat, fsm.nnz_hat,
            fsm.colptr.ctypes.data, fsm.rowval.ctypes.data, fsm.nzval.ctypes.data,
            fsm.init_idx.size, fsm.init_idx.ctypes.data, fsm.init_w.ctypes.data,
            smap.state2pdf.ctypes.data, smap.n_pdf_hat, 0, int(device)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().mk_graph_destroy(h)
            except Exception:
                pass
            self._h = None


def compile(fsm, Ĉ, device=None):  # noqa: A001 - the reference's name
    """``compile(fsm, Ĉ)`` (src/inference.jl:11-12) + ``adapt(CuArray, ·)`` (:14-26)."""
    return CompiledFSM(fsm, _as_statemap(fsm, Ĉ), device)


class BatchedFSM:
    """``batch(fsm1, fsms...)`` (src/inference.jl:28-36): the virtual rawunion of compiled FSMs.
    Identical operands are stored once; no block-diagonal matrix is built."""

    def __init__(self, cfsms):
        cfsms = list(cfsms)
        if not cfsms:
            raise ValueError("empty batch")
        self.cfsms = cfsms
        self.K = cfsms[0].K
        self.B = len(cfsms)
        self.n_pdf_hat = cfsms[0].n_pdf_hat
        self.offsets = np.cumsum([0] + [c.nstates_hat for c in cfsms])
        self.total_states_hat = int(self.offsets[-1])
        arr = (C.c_void_p * self.B)(*[c._h.value for c in cfsms])
        h = C.c_void_p()
        _lib.check(_lib.lib().mk_batch_create(C.byref(h), arr, self.B))
        self._h = h

    def set_overlap(self, enable=True):
        """Share every SM with a second batch in flight (``mk_batch_set_overlap``): enable on both batches."""
        _lib.check(_lib.lib().mk_batch_set_overlap(self._h, int(enable)))

    def wait(self):
        """Complete a ``pdfposteriors(..., wait=False)`` call on this batch (``mk_batch_wait``)."""
        _lib.check(_lib.lib().mk_batch_wait(self._h))
        self._keep = None

    def workspace_bytes(self):
        return int(_lib.lib().mk_batch_workspace_bytes(self._h))

    def profile(self, enable=True):
        _lib.check(_lib.lib().mk_batch_profile(self._h, int(enable)))

    def kernel_ms(self, cap=64):
        """Durations (ms) of the most recent shared-graph kernel launches, oldest first."""
        ms = (C.c_float * cap)()
        n = C.c_int()
        _lib.check(_lib.lib().mk_batch_kernel_ms(self._h, ms, cap, C.byref(n)))
        return list(ms[:n.value])

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().mk_batch_destroy(h)
            except Exception:
                pass
            self._h = None


def batch(fsm1, *fsms):
    return BatchedFSM((fsm1,) + fsms)


def _as_batch(x, Ĉs, B):
    """Accept what the reference's entry points accept: a BatchedFSM / CompiledFSM, or a
    (rawunion'd) FSM with its list of Ĉ matrices (src/inference.jl:145)."""
    if isinstance(x, BatchedFSM):
        return x
    if isinstance(x, CompiledFSM):
        return BatchedFSM([x] * B)
    if isinstance(x, FSM):
        parts = x.parts if x.parts is not None else [x]
        if Ĉs is None:
            raise TypeError("pdfposteriors(fsm::FSM, V̂s, Ĉs) needs the Ĉ matrices")
        if not isinstance(Ĉs, (list, tuple)):
            Ĉs = [Ĉs] * len(parts)
        if len(Ĉs) != len(parts):
            raise _lib.DimensionMismatch(_lib.MK_EINVAL, "one Ĉ per FSM of the union expected")
        cache = {}
        out = []
        for p, c in zip(parts, Ĉs):  # the replicated denominator is compiled once
            key = (id(p), id(c))
            if key not in cache:
                cache[key] = compile(p, c)
            out.append(cache[key])
        return BatchedFSM(out)
    raise TypeError(f"cannot run inference on {type(x).__name__}")


# ---------------------------------------------------------------------------------------------
# emissions
# ---------------------------------------------------------------------------------------------
def expand(V, seqlength=None, K=None):
    """``expand`` (src/inference.jl:54-60): D x N payload matrix -> (D+1) x (N+1) with the phony
    pdf row and phony frame column, filled with ``zero(K)`` / ``one(K)`` (default: the Log / Tropical
    payloads -Inf / 0; ``K=ProbSemiring[...]`` gives 0 / 1).  Works on numpy arrays and torch tensors."""
    is_t = type(V).__module__.startswith("torch")
    D, N = V.shape
    L = N if seqlength is None else int(seqlength)
    zero, one = (-float("inf"), 0.0) if K is None else (float(K.zero), float(K.one))
    if is_t:
        torch = _torch()
        out = torch.full((D + 1, N + 1), zero, dtype=V.dtype, device=V.device)
        out[:D, :L] = V[:, :L]
        out[D, L:] = one
        return out
    V = np.asarray(V)
    out = np.full((D + 1, N + 1), zero, V.dtype)
    out[:D, :L] = V[:, :L]
    out[D, L:] = one
    return out


class _Emis:
    pass


def _emissions(V, K, n_pdf_hat):
    """Normalise the emission argument to (pointer, strides, D, T, expanded, on_device)."""
    e = _Emis()
    if isinstance(V, (list, tuple)):  # the reference's V̂s: one matrix per utterance
        if type(V[0]).__module__.startswith("torch"):
            V = _torch().stack(list(V))
        else:
            V = np.stack([np.asarray(v) for v in V])
    if V.ndim != 3:
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, "emissions must be (B, D, T)")
    e.on_device = type(V).__module__.startswith("torch")
    if e.on_device:
        torch = _torch()
        want = torch.float32 if K.dtype == np.float32 else torch.float64
        if V.dtype != want:
            V = V.to(want)
        if not V.is_cuda:  # CPU torch tensor: treat as host array
            V = V.numpy()
            e.on_device = False
    if e.on_device:
        e.ptr, e.strides = V.data_ptr(), tuple(V.stride())
    else:
        V = np.asarray(V, K.dtype)
        if any(s < 0 for s in V.strides):
            V = np.ascontiguousarray(V)
        e.ptr, e.strides = V.ctypes.data, tuple(s // V.itemsize for s in V.strides)
    e.keep = V
    e.B, e.D, e.T = V.shape
    if e.D == n_pdf_hat:
        e.expanded = 1
    elif e.D == n_pdf_hat - 1:
        e.expanded = 0
    else:
        raise _lib.DimensionMismatch(
            _lib.MK_EINVAL, f"emissions have {e.D} pdfs, the graphs expect {n_pdf_hat - 1} (+1 when expanded)")
    return e


def _seqlens(seqlengths, B):
    if seqlengths is None:
        return None
    a = np.ascontiguousarray(seqlengths, np.int32)
    if a.shape != (B,):
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, "one sequence length per utterance expected")
    return a


def _slp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _tdtype(K):
    torch = _torch()
    return torch.float32 if K.dtype == np.float32 else torch.float64


# ---------------------------------------------------------------------------------------------
# recursions
# ---------------------------------------------------------------------------------------------
def _state_recursion(fn_name, x, V, Ĉs, seqlengths):
    torch = _torch()
    B = len(V) if isinstance(V, (list, tuple)) else V.shape[0]
    b = _as_batch(x, Ĉs, B)
    e = _emissions(V, b.K, b.n_pdf_hat)
    if e.B != b.B:
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, f"{e.B} emission matrices for {b.B} FSMs")
    sl = _seqlens(seqlengths, b.B)
    was_host = not e.on_device
    if was_host:
        Vd = torch.as_tensor(e.keep).cuda()
        e = _emissions(Vd, b.K, b.n_pdf_hat)
    N1 = e.T if e.expanded else e.T + 1
    out = torch.empty((N1, b.total_states_hat), dtype=_tdtype(b.K), device="cuda")
    fn = getattr(_lib.lib(), fn_name)
    _lib.check(fn(b._h, e.ptr, *e.strides, e.D, e.T, e.expanded, _slp(sl), out.data_ptr(), _stream()))
    out = out.t()  # (ΣŜ, N̂), column-major like the reference
    return out.cpu().numpy() if was_host else out


def αrecursion(x, V, Ĉs=None, seqlengths=None):
    """``αrecursion`` (src/inference.jl:62-74) for a whole batch: returns A, (ΣŜ_b) x N̂."""
    return _state_recursion("mk_alpha", x, V, Ĉs, seqlengths)


def βrecursion(x, V, Ĉs=None, seqlengths=None):
    """``βrecursion`` (src/inference.jl:99-110) for a whole batch: returns B, (ΣŜ_b) x N̂."""
    return _state_recursion("mk_beta", x, V, Ĉs, seqlengths)


def _check_out(buf, shape, dtype, on_device, name):
    """A caller-supplied output buffer must be exactly what the kernels write: payload dtype, shape, contiguous, on
    the side (device / host) of the emissions.  Anything else would be a silent out-of-bounds write."""
    if on_device:
        if not (type(buf).__module__.startswith("torch") and buf.is_cuda):
            raise TypeError(f"out: {name} must be a CUDA tensor for device emissions")
        ok = buf.is_contiguous()
    else:
        if not isinstance(buf, np.ndarray):
            raise TypeError(f"out: {name} must be a numpy array for host emissions")
        ok = buf.flags.c_contiguous and buf.flags.writeable
    if buf.dtype != dtype:
        raise TypeError(f"out: {name} has dtype {buf.dtype}, the semiring's payload type is {dtype}")
    if tuple(buf.shape) != tuple(shape) or not ok:
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, f"out: {name} must be a contiguous {tuple(shape)} array, got "
                                                     f"{tuple(buf.shape)}{'' if ok else ' (not contiguous)'}")


def pdfposteriors(x, V, Ĉs=None, seqlengths=None, out=None, stats=None, wait=True):
    """``pdfposteriors(fsm, V̂s, Ĉs)`` (src/inference.jl:145-161).

    Returns ``(post, ttl)``: ``post`` is the ``(B, D, N)`` array of pdf posteriors (exp domain,
    utterance index fastest in memory, as the reference lays it out) and ``ttl`` the ``B``
    total log-likelihoods.  ``out=(post_buf, ttl_buf)`` supplies the output storage: contiguous
    ``(N, D, B)`` and ``(B,)`` buffers of the payload dtype (device tensors for device input,
    e.g. pinned numpy arrays for host input).  ``stats`` (device input only): a float64 CUDA tensor
    of ``D + 2`` entries that receives the data-parallel step statistics ``[Σ logZ, #frames,
    occupancy[D]]`` (``mk_pdfposteriors_stats``; see :mod:`sharding`).  ``wait=False`` (host input with
    ``out=`` PINNED buffers only): enqueue the copies and kernels and return; ``x.wait()`` completes the call
    (``mk_pdfposteriors_host_begin`` / ``mk_batch_wait`` — keep three ``batch`` objects in flight: the copies of
    one call then fit beside the sweeps of the other two)."""
    B = len(V) if isinstance(V, (list, tuple)) else V.shape[0]
    b = _as_batch(x, Ĉs, B)
    e = _emissions(V, b.K, b.n_pdf_hat)
    if e.B != b.B:
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, f"{e.B} emission matrices for {b.B} FSMs")
    sl = _seqlens(seqlengths, b.B)
    Do, To = (e.D - 1, e.T - 1) if e.expanded else (e.D, e.T)
    l = _lib.lib()
    if e.on_device:
        torch = _torch()
        if out is not None:
            post, ttl = out
            _check_out(post, (To, Do, b.B), _tdtype(b.K), True, "post")
            _check_out(ttl, (b.B,), _tdtype(b.K), True, "ttl")
        else:
            post = torch.empty((To, Do, b.B), dtype=_tdtype(b.K), device="cuda")
            ttl = torch.empty((b.B,), dtype=_tdtype(b.K), device="cuda")
        sp = None
        if stats is not None:
            if not (stats.is_cuda and stats.dtype == torch.float64 and stats.is_contiguous() and stats.numel() == Do + 2):
                raise _lib.DimensionMismatch(_lib.MK_EINVAL, f"stats must be a contiguous float64 CUDA tensor of {Do + 2} entries")
            sp = stats.data_ptr()
        _lib.check(l.mk_pdfposteriors_stats(b._h, e.ptr, *e.strides, e.D, e.T, e.expanded, _slp(sl),
                                            post.data_ptr(), ttl.data_ptr(), sp, _stream()))
        return post.permute(2, 1, 0), ttl
    if stats is not None:
        raise TypeError("stats= needs device emissions (the host-buffer call returns posteriors to the host)")
    if out is not None:
        post, ttl = out
        _check_out(post, (To, Do, b.B), b.K.dtype, False, "post")
        _check_out(ttl, (b.B,), b.K.dtype, False, "ttl")
    else:
        post = np.empty((To, Do, b.B), b.K.dtype)
        ttl = np.empty((b.B,), b.K.dtype)
    if not wait:
        if out is None:
            raise TypeError("wait=False needs out= (pinned host buffers that outlive the call)")
        b._keep = (e, sl, post, ttl)   # the host arrays stay referenced until wait()
        _lib.check(l.mk_pdfposteriors_host_begin(b._h, e.ptr, *e.strides, e.D, e.T, e.expanded, _slp(sl),
                                                 post.ctypes.data, ttl.ctypes.data))
        return post.transpose(2, 1, 0), ttl
    _lib.check(l.mk_pdfposteriors_host(b._h, e.ptr, *e.strides, e.D, e.T, e.expanded, _slp(sl),
                                       post.ctypes.data, ttl.ctypes.data))
    return post.transpose(2, 1, 0), ttl


def bestpath(x, V, Ĉs=None, seqlengths=None):
    """Viterbi best path (TropicalSemiring graphs).  Returns ``(paths, scores)``: ``paths`` is a
    ``(B, T)`` int32 array of 1-based state ids (0 after each utterance's length)."""
    B = len(V) if isinstance(V, (list, tuple)) else V.shape[0]
    b = _as_batch(x, Ĉs, B)
    if b.K.code != MK_TROPICAL:
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, "bestpath needs TropicalSemiring graphs")
    e = _emissions(V, b.K, b.n_pdf_hat)
    if e.B != b.B:
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, f"{e.B} emission matrices for {b.B} FSMs")
    sl = _seqlens(seqlengths, b.B)
    l = _lib.lib()
    if e.on_device:
        torch = _torch()
        path = torch.empty((b.B, e.T), dtype=torch.int32, device="cuda")
        score = torch.empty((b.B,), dtype=_tdtype(b.K), device="cuda")
        _lib.check(l.mk_bestpath(b._h, e.ptr, *e.strides, e.D, e.T, e.expanded, _slp(sl),
                                 path.data_ptr(), score.data_ptr(), _stream()))
        return path, score
    path = np.empty((b.B, e.T), np.int32)
    score = np.empty((b.B,), b.K.dtype)
    _lib.check(l.mk_bestpath_host(b._h, e.ptr, *e.strides, e.D, e.T, e.expanded, _slp(sl),
                                  path.ctypes.data, score.ctypes.data))
    return path, score


# ASCII aliases
alpha_recursion = αrecursion
beta_recursion = βrecursion
