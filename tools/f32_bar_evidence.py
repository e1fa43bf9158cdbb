# SPDX-License-Identifier: MIT
"""Evidence for the Float32 posterior bar.

north_star asks for pdf posteriors within 1e-4 of "the reference's CPU implementation" in Float32.  That
implementation (restated by oracle.cpp) runs the UN-normalised log-domain recursion in Float32, so it carries its own
rounding error.  This script measures, at cfg 3 (the bench workload, BASELINE.json configs[2]):

    ref_err = max |post(oracle f32) - post(oracle f64)|      the reference arithmetic's distance from the exact answer
    gpu_err = max |post(CUDA f32)   - post(oracle f64)|      ours            (only with a GPU; --no-gpu skips it)
    gpu_ref = max |post(CUDA f32)   - post(oracle f32)|

(the f64 oracle evaluates the same Float32 inputs exactly enough to serve as the truth: its own error is ~1e-13), plus
the same three numbers for the log-likelihoods, relative.  Prints one JSON line; tests/test_gpu_parity.py asserts
gpu_err <= ref_err at the same shape.   python tools/f32_bar_evidence.py [--utts 16] [--no-gpu]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import markov_b200 as mm
import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=16)
ap.add_argument("--frames", type=int, default=150)
ap.add_argument("--no-gpu", action="store_true")
a = ap.parse_args()
B, T, D = 128, a.frames, 3000
K32, K64 = mm.LogSemiring[np.float32], mm.LogSemiring[np.float64]
fsm, pdf = mm.graphs.denominator(K32)
idx = np.linspace(0, B - 1, a.utts).astype(int)
out = {"workload": "cfg 3 (30000 states, 3000 pdfs, T=%d), %d utterances of the 128" % (T, a.utts)}
if not a.no_gpu:
    import torch
    V = torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(303), device="cuda") * 2
    c = mm.compile(fsm, mm.statemap(fsm, D, pdf))
    post, ttl = mm.pdfposteriors(mm.batch(*[c] * B), V.permute(0, 2, 1))
    Vh = V[idx].cpu().numpy()
    gpost, gttl = post[idx].cpu().numpy(), ttl[idx].cpu().numpy()
else:
    Vh = (np.random.default_rng(303).standard_normal((a.utts, T, D)) * 2).astype(np.float32)
t0 = time.time()
p32, z32 = oracle.pdfposteriors([oracle.OracleGraph(fsm, pdf, D)] * a.utts, Vh)
p64, z64 = oracle.pdfposteriors([oracle.OracleGraph(fsm.astype(K64), pdf, D)] * a.utts, Vh.astype(np.float64))
out["oracle_seconds"] = round(time.time() - t0, 1)
big = p64 > 1e-3   # relative errors where the posterior is not negligible
out["ref_err_abs"] = float(np.abs(p32 - p64).max())
out["ref_err_rel_p>1e-3"] = float((np.abs(p32 - p64)[big] / p64[big]).max())
out["ref_logz_rel"] = float(np.abs((z32 - z64) / z64).max())
if not a.no_gpu:
    out["gpu_err_abs"] = float(np.abs(gpost - p64).max())
    out["gpu_err_rel_p>1e-3"] = float((np.abs(gpost - p64)[big] / p64[big]).max())
    out["gpu_vs_ref_abs"] = float(np.abs(gpost - p32).max())
    out["gpu_logz_rel"] = float(np.abs((gttl - z64) / z64).max())
    out["gpu_closer_than_reference_arithmetic"] = bool(out["gpu_err_abs"] <= out["ref_err_abs"])
print(json.dumps(out))
