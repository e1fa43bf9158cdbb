# SPDX-License-Identifier: MIT
"""The N>1 path on CPU: utterance sharding + the one all-reduce of [Σ logZ, #frames, occupancy],
world_size 2 over gloo.  The per-rank compute is played by the oracle (the CUDA path needs a GPU);
what is under test is the host-side partition and the exchange."""
import os
import socket

import numpy as np
import pytest


def test_shard_bounds_cover_and_balance(mm):
    for n, w in ((128, 8), (1024, 8), (10, 4), (3, 8), (0, 2)):
        spans = [mm.sharding.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[k][1] == spans[k + 1][0] for k in range(w - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mm.sharding.shard_bounds(8, 2, 2)


def test_shard_by_length_balances_frames(mm):
    rng = np.random.default_rng(0)
    lens = rng.integers(75, 151, 128)
    parts = mm.sharding.shard_by_length(lens, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(128))
    assert all(len(p) == 16 for p in parts)
    frames = [lens[p].sum() for p in parts]
    assert max(frames) - min(frames) <= 0.02 * np.mean(frames)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import markov_b200 as mm
    import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K = mm.LogSemiring[np.float64]
        fsm, pdfids = mm.graphs.denominator(K, n_tokens=60, n_pdf=24, seed=3)  # replicated graph
        g = oracle.OracleGraph(fsm, pdfids, 24)
        rng = np.random.default_rng(5)  # same stream on every rank: each takes its slice
        B, T, D = 6, 12, 24
        V = rng.standard_normal((B, T, D)) * 2
        lens = rng.integers(6, T + 1, B)
        lo, hi = mm.sharding.shard_bounds(B, rank, world)
        post, ttl = oracle.pdfposteriors([g] * (hi - lo), V[lo:hi], lens[lo:hi])
        stats = mm.sharding.allreduce_stats(mm.sharding.local_stats(post, ttl, lens[lo:hi]))
        np.save(os.path.join(out_dir, f"stats{rank}.npy"), stats.numpy())
        if rank == 0:
            fpost, fttl = oracle.pdfposteriors([g] * B, V, lens)
            np.save(os.path.join(out_dir, "full.npy"), mm.sharding.local_stats(fpost, fttl, lens))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    s0, s1 = np.load(tmp_path / "stats0.npy"), np.load(tmp_path / "stats1.npy")
    full = np.load(tmp_path / "full.npy")
    np.testing.assert_array_equal(s0, s1)  # every rank holds the reduced statistics
    np.testing.assert_allclose(s0, full, rtol=1e-12, atol=1e-12)
    assert s0[1] == full[1] and abs(s0[2:].sum() - full[1]) < 1e-6  # occupancy sums to #frames


@pytest.mark.gpu
def test_library_step_statistics_match_the_host_form():
    """mk_pdfposteriors_stats writes [Σ logZ, #frames, occupancy[D]] itself; same numbers as local_stats on the outputs
    (shared-graph and per-utterance kernels, ragged lengths), and a 1-rank communicator leaves them unchanged."""
    import torch
    import markov_b200 as mm
    K = mm.LogSemiring[np.float32]
    rng = np.random.default_rng(5)
    D, T = 60, 40
    den, den_pdf = mm.graphs.denominator(K, n_tokens=1500, n_pdf=D, seed=2)
    cden = mm.compile(den, mm.statemap(den, D, den_pdf))
    nums = [mm.graphs.numerator(K, rng, D, n_phones=6) for _ in range(3)]
    cn = [mm.compile(f, mm.statemap(f, D, p)) for f, p in nums]
    for graphs in ([cden] * 12, cn, [cden] * 8 + cn):
        B = len(graphs)
        V = torch.from_numpy((rng.standard_normal((B, T, D)) * 2).astype(np.float32)).cuda().permute(0, 2, 1)
        lens = rng.integers(T // 2, T + 1, B).astype(np.int32)
        stats = torch.full((D + 2,), 7.0, dtype=torch.float64, device="cuda")   # (overwritten, not accumulated)
        post, ttl = mm.pdfposteriors(mm.batch(*graphs), V, seqlengths=lens, stats=stats)
        want = mm.sharding.local_stats(post, ttl, lens)
        torch.testing.assert_close(stats, want, rtol=1e-5, atol=1e-4)
        assert float(stats[1]) == float(lens.sum())
        np.testing.assert_allclose(float(stats[2:].sum()), float(lens.sum()), rtol=1e-4)   # posteriors sum to 1 per frame
    comm = mm.sharding.Communicator(0, 1)
    before = stats.clone()
    comm.allreduce_(stats)
    torch.cuda.synchronize()
    torch.testing.assert_close(stats, before, rtol=0, atol=0)
    comm.close()
    with pytest.raises(TypeError):
        mm.pdfposteriors(mm.batch(*cn), V[:3].cpu().numpy(), seqlengths=lens[:3], stats=stats)
