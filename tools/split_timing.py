# SPDX-License-Identifier: MIT
"""Time the forward-only, backward-only and full calls of the cfg-3 workload (device resident)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import markov_b200 as mm

B, T, D = int(os.environ.get("B", 128)), int(os.environ.get("T", 150)), 3000
K = mm.LogSemiring[np.float32]
fsm, pdf = mm.graphs.denominator(K)
c = mm.compile(fsm, mm.statemap(fsm, D, pdf))
b = mm.batch(*[c] * B)
V = (torch.randn((B, T, D), device="cuda") * 2).permute(0, 2, 1)
lib = mm.lib()
import ctypes as C
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
b.profile(True)
post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
t = timed(lambda: mm.pdfposteriors(b, V, out=(post, ttl))); print("pdfposteriors ms", t, "kernel", b.kernel_ms(3))
if os.environ.get("FULL", "1") == "1":
    t = timed(lambda: mm.αrecursion(b, V)); print("alpha (fwd + unpack) ms", t, "kernel", b.kernel_ms(3))
    t = timed(lambda: mm.βrecursion(b, V)); print("beta (bwd, no posterior) ms", t, "kernel", b.kernel_ms(3))
