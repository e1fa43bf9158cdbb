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


def totalweightsum(fsm, n=None):
    """``totalweightsum(fsm, n = nstates(fsm))``: payload value (log / tropical weight) of the ``n``-th partial
    total weight sum of ``fsm``."""
    import torch
    n = fsm.nstates if n is None else int(n)
    if n < 1:
        raise ValueError("n must be >= 1")
    cfsm = compile(fsm, statemap(fsm, 1, np.zeros(fsm.nstates, np.int64)))  # every state emits pdf 1
    # expanded emissions, all 1̄: D̂ x N̂ zeros with N̂ = n + 1 columns
    dt = torch.float32 if fsm.K.dtype == np.float32 else torch.float64
    V = torch.zeros((1, 2, n + 1), dtype=dt, device="cuda")
    A = αrecursion(cfsm, V)
    return float(A[-1, n])
