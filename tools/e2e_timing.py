import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import markov_b200 as mm
B, T, D = 128, 150, 3000
K = mm.LogSemiring[np.float32]
fsm, pdf = mm.graphs.denominator(K)
c = mm.compile(fsm, mm.statemap(fsm, D, pdf)); b = mm.batch(*[c] * B)
Vh = torch.empty((B, T, D), pin_memory=True); Vh.copy_(torch.randn((B, T, D)) * 2)
post_h = torch.empty((T, D, B), pin_memory=True); ttl_h = torch.empty((B,), pin_memory=True)
Vn, pn, tn = Vh.numpy().transpose(0, 2, 1), post_h.numpy(), ttl_h.numpy()
for _ in range(3): mm.pdfposteriors(b, Vn, out=(pn, tn))
t0 = time.perf_counter()
for _ in range(10): mm.pdfposteriors(b, Vn, out=(pn, tn))
print("segments", os.environ.get("MK_SEGMENTS"), "e2e ms", (time.perf_counter() - t0) * 100)
