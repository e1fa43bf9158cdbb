# SPDX-License-Identifier: MIT
"""Semiring type descriptors — host mirror of the Semirings.jl types the path uses.

The reference imports ``LogSemiring{T}`` / ``TropicalSemiring{T}`` from the external
Semirings.jl 0.5 (``/root/reference/src/MarkovModels.jl:12``, ``Project.toml:11,18``).  They
are isbits wrappers of one float, so arrays of them are bit-identical to float arrays
(``src/linalg.jl:15-28`` re-wraps buffers that way); here ``K = LogSemiring[np.float32]`` is a
descriptor object carrying the dtype, the ABI code and host-side scalar ops, and arrays hold
the payload floats.

    ⊕         ⊗      0̄      1̄
    Log       logaddexp  +   -Inf   0
    Tropical  max        +   -Inf   0      (max-plus; SURVEY.md A.1, assumption A-TROP)
    Prob      +          *   0      1      (operator level only: mul! / broadcasts / totalsum,
                                            test/test_linalg.jl:89, src/algorithms.jl:8-36)
"""
import numpy as np

MK_LOG, MK_TROPICAL, MK_PROB = 0, 1, 2
MK_F32, MK_F64 = 0, 1


class SemiringType:
    """A concrete semiring ``K`` (e.g. ``LogSemiring[np.float32]``)."""

    _cache = {}

    def __init__(self, name, code, dtype):
        self.name = name
        self.code = code
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError(f"{name} supports Float32/Float64 payloads, got {dtype}")
        self.dtype_code = MK_F32 if self.dtype == np.float32 else MK_F64
        self.zero = self.dtype.type(0.0 if code == MK_PROB else -np.inf)
        self.one = self.dtype.type(1.0 if code == MK_PROB else 0.0)

    @property
    def add_ufunc(self):
        """The numpy ufunc of ⊕ (``.at`` / ``.reduce`` work on it)."""
        return {MK_LOG: np.logaddexp, MK_TROPICAL: np.maximum, MK_PROB: np.add}[self.code]

    # scalar / elementwise host ops on payload floats (graph construction only)
    def add(self, x, y):
        return self.add_ufunc(x, y).astype(self.dtype)

    def mul(self, x, y):
        x, y = np.asarray(x, self.dtype), np.asarray(y, self.dtype)
        return (x * y if self.code == MK_PROB else x + y).astype(self.dtype)

    def div(self, x, y):
        x, y = np.asarray(x, self.dtype), np.asarray(y, self.dtype)
        return (x / y if self.code == MK_PROB else x - y).astype(self.dtype)

    def sum(self, x):
        x = np.asarray(x, self.dtype)
        if x.size == 0:
            return self.zero
        return self.dtype.type(self.add_ufunc.reduce(x))

    def __call__(self, x):
        """``K(x)``: payload constructor (src/fsm.jl:77 ``K(b)``)."""
        return self.dtype.type(x)

    def __repr__(self):
        return f"{self.name}{{{'Float32' if self.dtype_code == MK_F32 else 'Float64'}}}"

    def __eq__(self, other):
        return isinstance(other, SemiringType) and (self.code, self.dtype) == (other.code, other.dtype)

    def __hash__(self):
        return hash((self.code, self.dtype.str))


class _SemiringFamily:
    def __init__(self, name, code):
        self.name, self.code = name, code

    def __getitem__(self, dtype):
        key = (self.code, np.dtype(dtype).str)
        if key not in SemiringType._cache:
            SemiringType._cache[key] = SemiringType(self.name, self.code, dtype)
        return SemiringType._cache[key]

    __call__ = __getitem__


LogSemiring = _SemiringFamily("LogSemiring", MK_LOG)
TropicalSemiring = _SemiringFamily("TropicalSemiring", MK_TROPICAL)
ProbSemiring = _SemiringFamily("ProbSemiring", MK_PROB)
