# SPDX-License-Identifier: MIT
"""``totalweightsum`` on the device — the reference's
``totalweightsum(fsm, n) = totalcumsum(fsm.α, fsm.T, fsm.ω, n)`` (src/algorithms.jl:8-16, :32-36):

    v₁ = α,  vᵢ = Tᵀ vᵢ₋₁,   total = ⊕_{i=1..n} vᵢ · ω

is the emission-free forward recursion.  On the extended graph (src/fsm.jl:19-28) the phony final state
has a 1̄ self-loop and collects ``vᵢ · ω`` at every step, so ``total`` is ``αrecursion``'s value of the
phony final state in column ``n + 1`` when every emission is 1̄ — the same kernels as the hot path, no
new device code.  The reference uses it to compare FSMs (test/test_fsms.jl:9-16).
"""
import numpy as np

from .inference import compile, statemap, αrecursion
from .linalg import CuSparseMatrixCSR, mul_
from .semirings import MK_PROB


def _operands(fsm, device="cuda"):
    """(α, CSR(Tᵀ), ω as a 1 x S CSR matrix) of ``fsm`` on the device — ``fsm.α, fsm.T, fsm.ω`` (src/fsm.jl:30-40)."""
    import torch
    K, S = fsm.K, fsm.nstates
    src, dst, w = fsm.arcs_hat()
    real = (src < S) & (dst < S)
    Tt = CuSparseMatrixCSR(K, dst[real] + 1, src[real] + 1, w[real], S, S, device=device)  # rows of Tᵀ = destinations
    fin = (dst == S) & (src < S)
    om = CuSparseMatrixCSR(K, np.ones(int(fin.sum()), np.int64), src[fin] + 1, w[fin], 1, S, device=device)
    return torch.from_numpy(fsm.α).to(device), Tt, om


def _dot(om, v, K):
    """``dot(v, ω)`` = ⊕_i v_i ⊗ ω_i: ω as a one-row matrix through ``mul!``."""
    import torch
    out = torch.full((1,), float(K.zero), dtype=v.dtype, device=v.device)  # an empty ω launches nothing
    return mul_(out, om, v)


def totalcumsum(fsm, n=None):
    """``totalcumsum(α, T, ω, n)`` (src/algorithms.jl:8-16) of ``fsm``'s (α, T, ω): ``⊕_{i=1..n} (Tᵀ)^{i-1} α · ω`` —
    the reference's host loop of ``T' * v`` products, each one ``mul!`` on the device (``mk_spmv``); any of the
    Log / Tropical / Prob semirings.  Returns the payload value."""
    import torch
    K = fsm.K
    n = fsm.nstates if n is None else int(n)
    if n < 1:
        raise ValueError("n must be >= 1")
    v, Tt, om = _operands(fsm)
    total = _dot(om, v, K)
    nxt = torch.empty_like(v)
    step = torch.empty_like(total)
    for _ in range(2, n + 1):
        nxt.fill_(float(K.zero))  # (a graph without arcs launches nothing: Tᵀ v = 0̄)
        mul_(nxt, Tt, v)
        v, nxt = nxt, v
        step.fill_(float(K.zero))
        mul_(step, om, v)
        total = _oplus(K, total, step)
    return float(total[0])


def totalsum(fsm, n=None):
    """``totalsum(α, T, ω, n)`` (src/algorithms.jl:23-29): ``(Tᵀ)^{n-1} α · ω`` (the non-cumulative variant)."""
    import torch
    K = fsm.K
    n = fsm.nstates if n is None else int(n)
    if n < 1:
        raise ValueError("n must be >= 1")
    v, Tt, om = _operands(fsm)
    nxt = torch.empty_like(v)
    for _ in range(2, n + 1):
        nxt.fill_(float(K.zero))
        mul_(nxt, Tt, v)
        v, nxt = nxt, v
    return float(_dot(om, v, K)[0])


def _oplus(K, x, y):
    """⊕ of two device scalars — again the operator: [x y] · [1̄ 1̄]ᵀ would do, but a 1 x 2 ``mul!`` per step is
    all launch latency; torch's elementwise ops on a single element are plumbing, not the path."""
    import torch
    if K.code == MK_PROB:
        return x + y
    return torch.logaddexp(x, y) if K.code == 0 else torch.maximum(x, y)


def totalweightsum(fsm, n=None):
    """``totalweightsum(fsm, n = nstates(fsm))``: payload value (log / tropical weight) of the ``n``-th partial
    total weight sum of ``fsm``.  Log / Tropical graphs run the emission-free forward recursion on the hot-path
    kernels (one launch for all ``n`` steps); ``ProbSemiring`` graphs take the reference's route, ``totalcumsum``
    through ``mul!``."""
    import torch
    n = fsm.nstates if n is None else int(n)
    if n < 1:
        raise ValueError("n must be >= 1")
    if fsm.K.code == MK_PROB:
        return totalcumsum(fsm, n)
    cfsm = compile(fsm, statemap(fsm, 1, np.zeros(fsm.nstates, np.int64)))  # every state emits pdf 1
    # expanded emissions, all 1̄: D̂ x N̂ zeros with N̂ = n + 1 columns
    dt = torch.float32 if fsm.K.dtype == np.float32 else torch.float64
    V = torch.zeros((1, 2, n + 1), dtype=dt, device="cuda")
    A = αrecursion(cfsm, V)
    return float(A[-1, n])
