# SPDX-License-Identifier: MIT
"""Debug (library built with MK_PROFILE_BARRIER=1): where do CTAs spend a frame — working, waiting for
their own warps, or waiting for the grid?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import markov_b200 as mm
K = mm.LogSemiring[np.float32]
if len(sys.argv) > 1 and sys.argv[1] == "wsj":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from conftest import load_golden_fsm
    B, T, D = 128, 700, 84
    fsm, pdf = load_golden_fsm("den_fsm_wsj", K)
else:
    B, T, D = 128, 150, 3000
    fsm, pdf = mm.graphs.denominator(K)
c = mm.compile(fsm, mm.statemap(fsm, D, pdf)); b = mm.batch(*[c] * B)
V = (torch.randn((B, T, D), device="cuda") * 2).permute(0, 2, 1)
post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
lib = C.CDLL(mm._lib.LIB_PATH)
buf = (C.c_ulonglong * (148 * 4 + 1))()
mm.pdfposteriors(b, V, out=(post, ttl)); lib.mk_debug_barrier_profile(buf)
mm.pdfposteriors(b, V, out=(post, ttl)); lib.mk_debug_barrier_profile(buf)
a = np.array(buf[:148 * 4], np.float64).reshape(148, 4)
print('exact-fallback events per call:', buf[148 * 4], 'of', 2 * (T + 1) * fsm.nstates_hat, 'row evaluations')
nb = 2 * (T + 1)
print("per barrier interval, cycles (mean over CTAs / min / max):")
for k, name in enumerate(["thread0 work since last barrier", "scalar phase (barrier exit -> chunk loop)", "CTA waits for the grid"]):
    x = a[:, k] / nb
    print(f"  {name:34s} {x.mean():9.0f} {x.min():9.0f} {x.max():9.0f}")
order = np.argsort(-a[:, 0])
print('slowest CTAs (work cycles/frame):', [(int(c), int(a[c, 0] / nb)) for c in order[:8]])
print('fastest CTAs:', [(int(c), int(a[c, 0] / nb)) for c in order[-4:]])
tot = a[:, :3].sum(1) / nb
print("  total per frame", tot.mean(), "cycles =", tot.mean() / 1.965e3, "us")
x = a[:, 3] / nb / 16
print(f"  chunk loop, cycles per warp per frame (mean over the CTA's 16 warps)  {x.mean():9.0f} {x.min():9.0f} {x.max():9.0f}")
w = a[:, 0] + a[:, 1]
print("  busy (work + CTA wait) spread: min %.0f max %.0f  (max/mean %.2f)" % ((w / nb).min(), (w / nb).max(), w.max() / w.mean()))
