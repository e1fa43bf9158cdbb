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
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg4", "cfg5", "real"]
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


if __name__ == "__main__":
    main()
