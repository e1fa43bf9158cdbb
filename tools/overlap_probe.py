# SPDX-License-Identifier: MIT
"""Two batches in flight on one GPU: device-resident cfg 3 calls on two streams, with and without SM sharing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import markov_b200 as mm
K = mm.LogSemiring[np.float32]
B, T, D = 128, 150, 3000
fsm, pdf = mm.graphs.denominator(K)
c = mm.compile(fsm, mm.statemap(fsm, D, pdf))
NB = int(os.environ.get('NB', '2'))
bs = [mm.batch(*[c] * B) for _ in range(NB)]
V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(303), device="cuda") * 2).permute(0, 2, 1)
outs = [(torch.empty((T, D, B), device="cuda"), torch.empty((B,), device="cuda")) for _ in range(NB)]
streams = [torch.cuda.Stream() for _ in range(NB)]
def run(n, two):
    for k in range(n):
        j = k % NB if two else 0
        with torch.cuda.stream(streams[j]):
            mm.pdfposteriors(bs[j], V, out=outs[j])
for label, two, overlap in (("one stream, 512 threads", False, 0), ("two streams, 512 threads", True, 0), ("two streams, SM sharing", True, 1), ("one stream, 512 threads", False, 0)):
    for b in bs: b.set_overlap(overlap)
    run(2 * NB, two); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(24, two); torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / 24
    print(f"{label:28s} {ms:7.3f} ms per batch  ({B * T / ms / 1e3:.3f} M frames/s)  mean logZ {float(outs[0][1].mean()):.4f} {float(outs[1][1].mean()):.4f}", flush=True)
