# SPDX-License-Identifier: MIT
"""cfg 3 (BASELINE.json configs[2]) pdfposteriors, a few calls — the workload for ncu captures of the shared-graph kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import markov_b200 as mm
K = mm.LogSemiring[np.float32]
B, T, D = 128, 150, 3000
fsm, pdf = mm.graphs.denominator(K)
c = mm.compile(fsm, mm.statemap(fsm, D, pdf)); b = mm.batch(*[c] * B)
V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(303), device="cuda") * 2).permute(0, 2, 1)
post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    mm.pdfposteriors(b, V, out=(post, ttl))
torch.cuda.synchronize()
print("mean logZ", float(ttl.mean()))
