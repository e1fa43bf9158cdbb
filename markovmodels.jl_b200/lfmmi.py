# SPDX-License-Identifier: MIT
"""LF-MMI objective on top of ``pdfposteriors`` — the caller either side of the hot path.

The reference ships no loss function; its training loop (examples/test_cuda.jl:118-152) calls
``pdfposteriors`` on the batched numerator graphs and on the replicated denominator graph with the
same network output (``permutedims`` to (B, D, T), :120), and uses the difference of the two
posterior arrays (:152) as the gradient.  ``lfmmi_loss`` is that step as a
``torch.autograd.Function``:

    loss = - Σ_b (logZ_num[b] - logZ_den[b])
    d loss / d loglikes[b, t, d] = γ_den[b, d, t] - γ_num[b, d, t]      (0 beyond an utterance's length)

The subtraction and the (B, D, N) → (B, T, D) layout change run in one kernel of the library
(``mk_lfmmi_grad``); torch only holds the tensors and chains the gradient.
"""
import ctypes as C

import numpy as np

from . import _lib
from .inference import pdfposteriors
from .semirings import MK_F32, MK_F64


def _torch():
    import torch
    return torch


def lfmmi_grad(num_post, den_post, seqlengths=None, scale=1.0, out=None):
    """``scale * (den_post - num_post)`` in the network's ``(B, T, D)`` layout.  ``num_post`` /
    ``den_post``: the ``(B, D, N)`` views returned by :func:`pdfposteriors` (b fastest in memory)."""
    torch = _torch()
    B, D, N = num_post.shape
    if tuple(den_post.shape) != (B, D, N):
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, "numerator and denominator posteriors differ in shape")
    nbase, dbase = num_post.permute(2, 1, 0), den_post.permute(2, 1, 0)
    if not (nbase.is_contiguous() and dbase.is_contiguous()):
        nbase, dbase = nbase.contiguous(), dbase.contiguous()
    grad = out if out is not None else torch.empty((B, N, D), dtype=num_post.dtype, device=num_post.device)
    sl = None
    if seqlengths is not None:
        sl = torch.as_tensor(np.ascontiguousarray(seqlengths, np.int32)).to(num_post.device)
    dtype = MK_F32 if num_post.dtype == torch.float32 else MK_F64
    _lib.check(_lib.lib().mk_lfmmi_grad(dtype, nbase.data_ptr(), dbase.data_ptr(), B, D, N,
                                        sl.data_ptr() if sl is not None else None, float(scale), grad.data_ptr(),
                                        *grad.stride(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return grad


def _make_function():
    torch = _torch()

    class _LFMMI(torch.autograd.Function):
        @staticmethod
        def forward(ctx, loglikes, num, den, seqlengths):
            V = loglikes.detach().permute(0, 2, 1)  # (B, D, T) view, as the reference's permutedims (:120)
            num_post, num_ttl = pdfposteriors(num, V, seqlengths=seqlengths)
            den_post, den_ttl = pdfposteriors(den, V, seqlengths=seqlengths)
            ctx.grad = lfmmi_grad(num_post, den_post, seqlengths)
            ctx.mark_non_differentiable(num_ttl, den_ttl)
            return -(num_ttl - den_ttl).sum(), num_ttl, den_ttl

        @staticmethod
        def backward(ctx, g_loss, _g_num, _g_den):
            return ctx.grad * g_loss, None, None, None

    return _LFMMI


_FN = None


def lfmmi_loss(loglikes, num, den, seqlengths=None):
    """LF-MMI loss of a ``(B, T, D)`` CUDA tensor of log-likelihoods.

    ``num``: BatchedFSM of the utterances' numerator graphs, ``den``: BatchedFSM of B copies of the
    denominator graph (``batch(*[cden] * B)``).  Returns ``(loss, logZ_num, logZ_den)``; ``loss`` is
    differentiable with respect to ``loglikes``."""
    global _FN
    if _FN is None:
        _FN = _make_function()
    if not loglikes.is_cuda or loglikes.dim() != 3:
        raise _lib.DimensionMismatch(_lib.MK_EINVAL, "loglikes must be a (B, T, D) CUDA tensor")
    return _FN.apply(loglikes, num, den, seqlengths)
