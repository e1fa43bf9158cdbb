# SPDX-License-Identifier: MIT
"""markov_b200 — B200-native (sm_100a) drop-in for MarkovModels.jl's batched semiring inference
path (αrecursion / βrecursion / pdfposteriors / bestpath in the Log and Tropical semirings).

The compute lives in ``csrc/libmarkov_b200.so`` (C ABI: ``include/markov_b200.h``); this package
is the host-side mirror of the reference's API names (``/root/reference/src/MarkovModels.jl:14-45``).
Import it as ``markov_b200`` (the directory name ``markovmodels.jl_b200`` is not an identifier;
``markov_b200.py`` at the repository root registers it).
"""
from ._lib import DimensionMismatch, MarkovError, build, lib  # noqa: F401
from .fsm import FSM, nstates, rawunion, renorm, union  # noqa: F401
from .inference import (BatchedFSM, CompiledFSM, StateMap, alpha_recursion, batch, bestpath,  # noqa: F401
                        beta_recursion, compile, expand, pdfposteriors, statemap, αrecursion, βrecursion)
from .semirings import LogSemiring, ProbSemiring, TropicalSemiring  # noqa: F401
from . import algorithms, graphs, lfmmi, linalg, sharding  # noqa: F401
from .algorithms import totalcumsum, totalsum, totalweightsum  # noqa: F401
from .linalg import (CuSparseMatrixCSC, CuSparseMatrixCSR, CuSparseVector, blockdiag, copy_transpose,  # noqa: F401
                     csr_from_csc, eldiv_, elmul_, mul_, vcat)
from .lfmmi import lfmmi_grad, lfmmi_loss  # noqa: F401

__version__ = "0.1.0"
