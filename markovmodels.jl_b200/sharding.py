# SPDX-License-Identifier: MIT
"""Multi-GPU data parallelism for the batched inference path.

The reference has no multi-GPU code: its batch is one block-diagonal FSM
(``rawunion``, src/fsmops.jl:28-36) whose blocks — the utterances — are independent.  The path
therefore shards by utterance: one process per GPU (``torch.distributed``), each rank owns a
contiguous slice of the batch and a replica of the shared (denominator) graph, emissions and
posteriors stay on the owning GPU, and the only exchange is ONE all-reduce per step of
``[Σ_b logZ_b, #frames, pdf occupancy[D]]`` (~12 KB; NCCL over NVLink on GPUs, gloo in the CPU
tests).  No collective touches the recursion itself.

On GPUs nothing of this runs in Python or in eager torch: ``pdfposteriors(..., stats=buf)`` makes the library
write the statistics (the occupancy falls out of its normalisation pass), and :class:`Communicator` is the
library's own NCCL binding (``mk_comm_init_rank`` / ``mk_allreduce_stats``), enqueued on the same stream —
one launch, no host synchronisation.  ``local_stats`` / ``allreduce_stats`` are the host-side / gloo forms
of the same two steps.
"""
import ctypes as C

import numpy as np

from . import _lib


class Communicator:
    """One NCCL communicator per rank through the C ABI (``mk_comm_*``): what a Julia host would ``ccall``.
    The 128-byte unique id travels from rank 0 to the others through ``torch.distributed`` here (any channel
    works: a file, MPI, a socket)."""

    def __init__(self, rank, world, device=-1, group=None):
        import torch
        import torch.distributed as dist
        l = _lib.lib()
        ident = (C.c_char * 128)()
        if rank == 0:
            _lib.check(l.mk_comm_unique_id(ident))
        if world > 1:
            box = [bytes(ident.raw) if rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)
            ident = (C.c_char * 128).from_buffer_copy(box[0])
        h = C.c_void_p()
        _lib.check(l.mk_comm_init_rank(C.byref(h), world, rank, ident, device))
        self._h, self.rank, self.world = h, rank, world

    def allreduce_(self, stats):
        """In-place float64 sum over the ranks, asynchronous on the current torch stream."""
        import torch
        if not (stats.is_cuda and stats.dtype == torch.float64 and stats.is_contiguous()):
            raise TypeError("stats must be a contiguous float64 CUDA tensor")
        _lib.check(_lib.lib().mk_allreduce_stats(self._h, stats.data_ptr(), stats.numel(),
                                                 torch.cuda.current_stream().cuda_stream))
        return stats

    def close(self):
        if self._h:
            _lib.lib().mk_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_bounds(n_utts, rank, world):
    """Contiguous, balanced slice [lo, hi) of ``n_utts`` utterances for ``rank`` of ``world``."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    base, rem = divmod(int(n_utts), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_length(seqlengths, world):
    """Length-balanced assignment for ragged batches: utterances sorted by length (longest first)
    are dealt to the ranks in a snake order so that every rank gets the same number of utterances
    (±1) and nearly the same number of frames.  Returns a list of index arrays, one per rank."""
    order = np.argsort(-np.asarray(seqlengths), kind="stable")
    out = [[] for _ in range(world)]
    for k, idx in enumerate(order):
        r = k % (2 * world)
        out[r if r < world else 2 * world - 1 - r].append(int(idx))
    return [np.asarray(sorted(o), np.int64) for o in out]


def local_stats(post, ttl, seqlengths=None):
    """[Σ logZ, #frames, occupancy[D]] of this rank's utterances.  ``post`` is the (B, D, N)
    posterior array (torch tensor or numpy), ``ttl`` the B log-likelihoods."""
    is_t = type(post).__module__.startswith("torch")
    B, D, N = post.shape
    frames = float(B * N if seqlengths is None else int(np.sum(np.asarray(seqlengths))))
    if is_t:
        import torch
        stats = torch.empty(D + 2, dtype=torch.float64, device=post.device)
        stats[0] = ttl.double().sum()
        stats[1] = frames
        # (reduce in the payload dtype, then widen: a float64 reduction would first copy the whole array)
        stats[2:] = post.sum(dim=(0, 2)).double()
        return stats
    stats = np.empty(D + 2, np.float64)
    stats[0] = np.sum(ttl, dtype=np.float64)
    stats[1] = frames
    stats[2:] = post.sum(axis=(0, 2), dtype=np.float64)
    return stats


def allreduce_stats(stats, group=None):
    """Sum the per-rank statistics over the data-parallel group (no-op without a process group).
    Returns a torch tensor (on the device of ``stats`` if it was one)."""
    import torch
    import torch.distributed as dist
    t = stats if isinstance(stats, torch.Tensor) else torch.from_numpy(np.asarray(stats))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
