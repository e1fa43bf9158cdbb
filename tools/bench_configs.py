# SPDX-License-Identifier: MIT
"""Secondary measurements over BASELINE.json's other configs (1 GPU, device-resident inputs, CUDA
events, 3 warm-ups).  bench.py stays the contract benchmark (configs[2]); these lines document the
rest of SURVEY.md §8d.  One JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import markov_b200 as mm

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden_fsm  # noqa: E402


def timed(fn, n=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def emit(name, **kw):
    print(json.dumps({"config": name, **kw}), flush=True)


def main():
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg4", "ragged", "cfg5", "real", "linalg"]
    f32 = mm.LogSemiring[np.float32]
    if "cfg1" in which:  # phone loop, Float64, T=500: CPU anchor + GPU parity
        import oracle
        K = mm.LogSemiring[np.float64]
        fsm, pdf = mm.graphs.phone_loop(K, 33)
        D, T = fsm.nstates, 500
        V = np.random.default_rng(101).standard_normal((1, T, D))
        g = oracle.OracleGraph(fsm, pdf, D)
        t0 = time.perf_counter(); opost, ottl = oracle.pdfposteriors([g], V, threads=1); t_cpu = time.perf_counter() - t0
        b = mm.batch(mm.compile(fsm, mm.statemap(fsm, D, pdf)))
        Vd = torch.from_numpy(V).cuda().permute(0, 2, 1)
        post, ttl = mm.pdfposteriors(b, Vd)
        err = float(np.abs(post.cpu().numpy() - opost).max())
        Kt = mm.TropicalSemiring[np.float64]
        ft = fsm.astype(Kt)
        bt = mm.batch(mm.compile(ft, mm.statemap(ft, D, pdf)))
        path, score = mm.bestpath(bt, Vd)
        opath, oscore = oracle.bestpath([oracle.OracleGraph(ft, pdf, D)], V)
        ms = timed(lambda: mm.pdfposteriors(b, Vd))
        emit("cfg1 phone-loop 99 states, T=500, LogSemiring{Float64}", cpu_oracle_frames_per_s=T / t_cpu,
             gpu_frames_per_s=T / (ms * 1e-3), max_abs_posterior_err=err, logz_rel_err=float(abs(ttl[0] - ottl[0]) / abs(ottl[0])),
             bestpath_equal=bool((path.cpu().numpy() == opath).all()), score_equal=bool(score.cpu().numpy()[0] == oscore[0]))
    if "cfg2" in which:  # 128 distinct numerator graphs, T=150, f32
        B, T, D = 128, 150, 3000
        graphs = [mm.graphs.numerator(f32, np.random.default_rng(202 + k), D) for k in range(B)]
        cs = [mm.compile(f, mm.statemap(f, D, p)) for f, p in graphs]
        b = mm.batch(*cs)
        V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(202), device="cuda") * 2).permute(0, 2, 1)
        post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
        ms = timed(lambda: mm.pdfposteriors(b, V, out=(post, ttl)))
        lens = torch.randint(75, 151, (B,), generator=torch.Generator().manual_seed(1)).numpy().astype(np.int32)
        ms_r = timed(lambda: mm.pdfposteriors(b, V, seqlengths=lens, out=(post, ttl)))
        emit("cfg2 128 numerator graphs, T=150, f32", states_mean=float(np.mean([f.nstates for f, _ in graphs])),
             ms=ms, frames_per_s=B * T / (ms * 1e-3), ragged_ms=ms_r, finite=bool(torch.isfinite(ttl).all()))
    if "cfg4" in which:  # denominator, B=1024 on one GPU
        B, T, D = 1024, 150, 3000
        fsm, pdf = mm.graphs.denominator(f32)
        c = mm.compile(fsm, mm.statemap(fsm, D, pdf))
        b = mm.batch(*[c] * B)
        V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(404), device="cuda") * 2).permute(0, 2, 1)
        post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
        ms = timed(lambda: mm.pdfposteriors(b, V, out=(post, ttl)), n=3, warm=2)
        emit("cfg4 denominator B=1024 on 1 GPU, T=150, f32", ms=ms, frames_per_s=B * T / (ms * 1e-3),
             workspace_GB=b.workspace_bytes() / 1e9, mean_logz=float(ttl.mean()))
        del b, post, V
        torch.cuda.empty_cache()
    if "ragged" in which:  # cfg 4's batch with ragged lengths U[75, 150]: per-tile frame limits on / off, sorted / as drawn
        B, T, D = 1024, 150, 3000
        fsm, pdf = mm.graphs.denominator(f32)
        c = mm.compile(fsm, mm.statemap(fsm, D, pdf))
        b = mm.batch(*[c] * B)
        V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(404), device="cuda") * 2).permute(0, 2, 1)
        post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
        drawn = np.random.default_rng(404).integers(75, 151, B).astype(np.int32)
        res = {}
        for name, lens in (("as_drawn", drawn), ("sorted", np.sort(drawn)[::-1].copy())):
            for flag in ("1", "0"):
                os.environ["MK_RAGGED_CUT"] = flag
                ms = timed(lambda: mm.pdfposteriors(b, V, seqlengths=lens, out=(post, ttl)), n=3, warm=2)
                res[f"{name}_cut{flag}_ms"] = ms
                res[f"{name}_cut{flag}_real_frames_per_s"] = float(lens.sum()) / (ms * 1e-3)
        os.environ.pop("MK_RAGGED_CUT", None)
        emit("ragged: denominator B=1024 on 1 GPU, lengths U[75,150], f32 (real frames = sum of lengths)", **res)
        del b, post, V
        torch.cuda.empty_cache()
    if "cfg5" in which:  # Viterbi on the denominator, B=512, T=500
        B, T, D = 512, 500, 3000
        Kt = mm.TropicalSemiring[np.float32]
        fsm, pdf = mm.graphs.denominator(Kt)
        c = mm.compile(fsm, mm.statemap(fsm, D, pdf))
        b = mm.batch(*[c] * B)
        V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(505), device="cuda") * 2).permute(0, 2, 1)
        ms = timed(lambda: mm.bestpath(b, V), n=2, warm=1)
        path, score = mm.bestpath(b, V)
        emit("cfg5 bestpath on the denominator, B=512, T=500, TropicalSemiring{Float32}", ms=ms,
             frames_per_s=B * T / (ms * 1e-3), workspace_GB=b.workspace_bytes() / 1e9,
             all_paths_complete=bool((path > 0).all()), mean_score=float(score.mean()))
        del b, V
        torch.cuda.empty_cache()
    if "real" in which:  # the reference's own benchmark: den_fsm_wsj, B=128, N=700, lhs = ones
        B, T, D = 128, 700, 84
        fsm, pdf = load_golden_fsm("den_fsm_wsj", f32)
        c = mm.compile(fsm, mm.statemap(fsm, D, pdf))
        b = mm.batch(*[c] * B)
        V = torch.ones((B, D, T), device="cuda")
        post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
        ms = timed(lambda: mm.pdfposteriors(b, V, out=(post, ttl)))
        emit("real: misc/benchmark den_fsm_wsj (3032 states), B=128, N=700, lhs=ones, f32", ms=ms,
             frames_per_s=B * T / (ms * 1e-3), logz=float(ttl[0]), logz_expected=692.168685813936,
             reference_gtx1080_frames_per_s=44730, reference_cpu_frames_per_s=263)

    if "linalg" in which:  # the reference's operator level at cfg 3's sizes: K1 (mul! SpMV) and K2 (mul! SpMM)
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                           "MEASURED_PEAKS.json")))["hbm_gbs"]
        B, T, D = 128, 150, 3000
        fsm, pdf = mm.graphs.denominator(f32)
        S = fsm.nstates_hat
        src, dst, w = fsm.arcs_hat()
        # T̂ᵀ of ONE graph as CSR (rows = destinations), then the reference's blockdiag over the batch
        # (src/linalg.jl:73-131) built with torch on the device: 128 x 523k arcs = 67M nnz, 520 MB of (colVal, nzVal)
        one = mm.CuSparseMatrixCSR(f32, dst + 1, src + 1, w, S, S)
        nnz1 = one.nnz
        blk = mm.CuSparseMatrixCSR.__new__(mm.CuSparseMatrixCSR)
        blk.K, blk.shape = f32, (B * S, B * S)
        offs = torch.arange(B, device="cuda", dtype=torch.int32)
        blk.rowPtr = torch.cat([(one.rowPtr[:-1][None, :] + offs[:, None] * nnz1).reshape(-1),
                                torch.tensor([B * nnz1 + 1], device="cuda", dtype=torch.int32)])
        blk.colVal = (one.colVal[None, :] + offs[:, None] * S).reshape(-1).contiguous()
        blk.nzVal = one.nzVal.repeat(B)
        x = torch.randn(B * S, device="cuda")
        y = torch.empty(B * S, device="cuda")
        ms = timed(lambda: mm.mul_(y, blk, x), n=20)
        by = blk.nnz * 8 + (B * S + 1) * 4 + 2 * B * S * 4  # colVal + nzVal, rowPtr, b read once, c written once
        emit("linalg K1: mul!(c, blockdiag(T̂ᵀ x128), b) — one frame of the reference's αrecursion at cfg 3, f32",
             rows=B * S, nnz=blk.nnz, ms=ms, algorithmic_GB=by / 1e9, achieved_GBps=by / (ms * 1e-3) / 1e9,
             hbm_peak_GBps=peak, frac=by / (ms * 1e-3) / 1e9 / peak)
        # one graph, L2-resident: what a frame costs when the graph is NOT replicated per utterance
        x1 = torch.randn(S, device="cuda"); y1 = torch.empty(S, device="cuda")
        ms1 = timed(lambda: mm.mul_(y1, one, x1), n=50)
        emit("linalg K1: mul!(c, T̂ᵀ, b) — one graph (30k rows, 523k arcs, L2-resident)", ms=ms1, arcs_per_s=nnz1 / (ms1 * 1e-3))
        del blk, x, y
        torch.cuda.empty_cache()
        # K2: Ĉ · V̂ (src/inference.jl:150): Ĉ = blockdiag of state->pdf indicator matrices (one 1̄ per row), V̂ = (B·D̂) x N̂
        Dh, N1 = D + 1, T + 1
        pdf_hat = np.concatenate([np.asarray(pdf, np.int64), [D]])
        I = torch.arange(1, B * S + 1, device="cuda", dtype=torch.int32)  # noqa: E741
        Chat = mm.CuSparseMatrixCSR.__new__(mm.CuSparseMatrixCSR)
        Chat.K, Chat.shape = f32, (B * S, B * Dh)
        Chat.rowPtr = torch.arange(1, B * S + 2, device="cuda", dtype=torch.int32)
        Chat.colVal = (torch.from_numpy(pdf_hat.astype(np.int32)).cuda()[None, :] + 1 + offs[:, None] * Dh).reshape(-1).contiguous()
        Chat.nzVal = torch.zeros(B * S, device="cuda")
        Vh = mm.linalg.colmajor(f32, B * Dh, N1)
        Vh.normal_()
        CV = mm.linalg.colmajor(f32, B * S, N1)
        ms2 = timed(lambda: mm.mul_(CV, Chat, Vh), n=5)
        by2 = (B * S * N1 + B * Dh * N1) * 4 + B * S * 12  # Ĉ·V̂ written once, V̂ read once, Ĉ read once
        emit("linalg K2: mul!(ĈV̂, Ĉ, V̂) — the SpMM in front of the reference's recursions at cfg 3, f32",
             rows=B * S, cols=N1, ms=ms2, algorithmic_GB=by2 / 1e9, achieved_GBps=by2 / (ms2 * 1e-3) / 1e9,
             hbm_peak_GBps=peak, frac=by2 / (ms2 * 1e-3) / 1e9 / peak)
        del I


if __name__ == "__main__":
    main()
