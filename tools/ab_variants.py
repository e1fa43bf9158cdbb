# SPDX-License-Identifier: MIT
"""A/B timing of library builds on ONE box (chips differ by a few percent, so variants are only comparable inside
one gpurun call): for every .so given, a fresh process loads it through MARKOV_B200_LIB and times cfg 3's
pdfposteriors (device-resident inputs, CUDA events, 3 warm-ups, 10 calls) plus the shared-graph kernel pair."""
import os, subprocess, sys

CHILD = r'''
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(sys.argv[0]))) if False else os.getcwd())
import numpy as np, torch
import markov_b200 as mm
from markov_b200 import _lib
raw = C.CDLL(_lib.LIB_PATH)
_lib.SIGNATURES = {k: v for k, v in _lib.SIGNATURES.items() if hasattr(raw, k)}   # older builds lack newer entry points
K = mm.LogSemiring[np.float32]
B, T, D = 128, 150, 3000
fsm, pdf = mm.graphs.denominator(K)
c = mm.compile(fsm, mm.statemap(fsm, D, pdf)); b = mm.batch(*[c] * B)
V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(303), device="cuda") * 2).permute(0, 2, 1)
post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
for _ in range(3): mm.pdfposteriors(b, V, out=(post, ttl))
torch.cuda.synchronize()
mm.lib().mk_batch_profile(b._h, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): mm.pdfposteriors(b, V, out=(post, ttl))
e1.record(); torch.cuda.synchronize()
ms = (C.c_float * 64)(); n = C.c_int(0)
mm.lib().mk_batch_kernel_ms(b._h, ms, 64, C.byref(n))
k = sorted(ms[:n.value])
print(f"{os.path.basename(_lib.LIB_PATH):12s} step {e0.elapsed_time(e1) / 10:7.3f} ms   kernel pair median {k[len(k) // 2]:7.3f} min {k[0]:7.3f} ms   mean logZ {float(ttl.mean()):.4f}", flush=True)
'''

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rep in range(int(os.environ.get("AB_REPEATS", "2"))):
    for lib in sys.argv[1:]:
        env = dict(os.environ, MARKOV_B200_LIB=os.path.abspath(lib))
        subprocess.run([sys.executable, "-c", CHILD], cwd=root, env=env)
