# SPDX-License-Identifier: MIT
"""ctypes binding of ``libmarkov_b200.so`` (the C ABI of ``include/markov_b200.h``).

The library is built in-tree (``csrc/libmarkov_b200.so``) by :func:`build`.  There is no CPU
fallback: if the shared object is missing, loading raises; if no CUDA device is usable, every
compute entry point returns ``MK_ECUDA`` and the wrappers raise :class:`MarkovError`.
"""
import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# MARKOV_B200_LIB selects another build of the same ABI (A/B timing of kernel variants, tools/ab_variants.py)
LIB_PATH = os.environ.get("MARKOV_B200_LIB") or os.path.join(CSRC, "libmarkov_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include", "markov_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

MK_OK, MK_EINVAL, MK_ENOMEM, MK_ENOTSUP, MK_ECUDA = 0, 22, 12, 95, 1000


class MarkovError(RuntimeError):
    """A failed libmarkov_b200 call.  ``code == MK_EINVAL`` is the reference's
    ``DimensionMismatch`` (src/linalg.jl:166-167)."""

    def __init__(self, code, msg):
        super().__init__(f"libmarkov_b200 error {code}: {msg}")
        self.code = code


class DimensionMismatch(MarkovError, ValueError):
    pass


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))] + [INCLUDE]
    return any(os.path.getmtime(s) > t for s in srcs)


SOURCES = ["markov_b200.cu", "prep.cu"]


def build(force=False, verbose=False):
    """Compile ``csrc/*.cu`` for sm_100a into ``csrc/libmarkov_b200.so`` (nvcc cross-compiles
    without a GPU): one object per translation unit, compiled side by side, then linked."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    for knob in ("MK_PASS_QUADS", "MK_THREADS", "MK_PROFILE_BARRIER", "MK_ABLATE", "MK_SPMM_CJ", "MK_SPMM_STCS", "MK_L2_HINTS", "MK_TR_BATCH"):  # kernel tuning knobs (defaults in kernels.cuh)
        if os.environ.get(knob):
            flags.insert(0, f"-D{knob}={os.environ[knob]}")
    if verbose:
        flags.insert(0, "-Xptxas=-v")
    tag = os.path.splitext(os.path.basename(LIB_PATH))[0]

    def compile_one(src):
        obj = os.path.join(CSRC, f"{tag}.{os.path.splitext(src)[0]}.o")
        res = subprocess.run([nvcc] + flags + ["-c", "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(len(SOURCES)) as pool:
        done = list(pool.map(compile_one, SOURCES))
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + [o for o, _ in done],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc (link) failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(err for _, err in done))
    return LIB_PATH


_lib = None


def _point_at_nccl():
    """mk_comm_* bind libnccl at run time (dlopen); tell them where PyTorch's bundled copy lives unless the caller did."""
    if os.environ.get("MK_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["MK_NCCL_LIB"] = cand
                return
    except Exception:
        pass

_i64, _i32, _vp = C.c_int64, C.c_int32, C.c_void_p
_EMIS = [_vp, _i64, _i64, _i64, _i64, _i64, C.c_int, C.POINTER(_i32)]  # ll, sb, sd, sn, D, T, expanded, seqlens

SIGNATURES = {
    "mk_abi_version": (C.c_int, []),
    "mk_last_error": (C.c_char_p, []),
    "mk_device_count": (C.c_int, []),
    "mk_graph_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp,
                                  _vp, _i64, C.c_int, C.c_int]),
    "mk_graph_destroy": (C.c_int, [_vp]),
    "mk_graph_info": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(C.c_int),
                                C.POINTER(C.c_int)]),
    "mk_batch_create": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), _i64]),
    "mk_batch_destroy": (C.c_int, [_vp]),
    "mk_batch_info": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "mk_alpha": (C.c_int, [_vp] + _EMIS + [_vp, _vp]),
    "mk_beta": (C.c_int, [_vp] + _EMIS + [_vp, _vp]),
    "mk_pdfposteriors": (C.c_int, [_vp] + _EMIS + [_vp, _vp, _vp]),
    "mk_pdfposteriors_stats": (C.c_int, [_vp] + _EMIS + [_vp, _vp, _vp, _vp]),
    "mk_bestpath": (C.c_int, [_vp] + _EMIS + [_vp, _vp, _vp]),
    "mk_pdfposteriors_host": (C.c_int, [_vp] + _EMIS + [_vp, _vp]),
    "mk_bestpath_host": (C.c_int, [_vp] + _EMIS + [_vp, _vp]),
    "mk_pdfposteriors_host_begin": (C.c_int, [_vp] + _EMIS + [_vp, _vp]),
    "mk_batch_wait": (C.c_int, [_vp]),
    "mk_batch_set_overlap": (C.c_int, [_vp, C.c_int]),
    "mk_lfmmi_grad": (C.c_int, [C.c_int, _vp, _vp, _i64, _i64, _i64, _vp, C.c_double, _vp, _i64, _i64, _i64, _vp]),
    "mk_spmv": (C.c_int, [C.c_int, C.c_int, _i64, _i64, _i64, _vp, _vp, _vp, C.c_int, _vp, _i64, _vp, _i64, _vp]),
    "mk_spmm": (C.c_int, [C.c_int, C.c_int, _i64, _i64, _i64, _vp, _vp, _vp, C.c_int, _vp, _i64, _i64, _i64, _vp, _i64,
                          _i64, _i64, C.c_int, _vp]),
    "mk_spvec_bcast": (C.c_int, [C.c_int, C.c_int, C.c_int, _i64, _i64, _vp, _vp, C.c_int, _vp, _i64, _vp, _i64, _vp]),
    "mk_comm_unique_id": (C.c_int, [_vp]),
    "mk_comm_init_rank": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp, C.c_int]),
    "mk_comm_init": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "mk_allreduce_stats": (C.c_int, [_vp, _vp, _i64, _vp]),
    "mk_allreduce_stats_all": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), C.c_int, _i64, C.POINTER(_vp)]),
    "mk_comm_destroy": (C.c_int, [_vp]),
    "mk_blockdiag": (C.c_int, [C.c_int, _i64, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64),
                               C.POINTER(_i64), C.POINTER(_i64), C.c_int, _vp, _vp, _vp, _vp]),
    "mk_vcat_spvec": (C.c_int, [C.c_int, _i64, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64), _vp, _vp, _vp]),
    "mk_sparse_transpose": (C.c_int, [C.c_int, _i64, _i64, _i64, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "mk_launch_count": (_i64, [C.c_int]),
    "mk_batch_workspace_bytes": (_i64, [_vp]),
    "mk_batch_profile": (C.c_int, [_vp, C.c_int]),
    "mk_measure_sfu_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "mk_batch_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
}


def lib():
    """The loaded library.  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libmarkov_b200 has no CPU fallback)")
        _point_at_nccl()
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        if l.mk_abi_version() != 2:
            raise RuntimeError("libmarkov_b200.so ABI version mismatch; rebuild")
        _lib = l
    return _lib


def check(rc):
    if rc != MK_OK:
        msg = lib().mk_last_error().decode("utf-8", "replace")
        raise (DimensionMismatch if rc == MK_EINVAL else MarkovError)(rc, msg)
