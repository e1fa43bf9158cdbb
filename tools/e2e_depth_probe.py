# SPDX-License-Identifier: MIT
"""Host-buffer pdfposteriors calls in flight: how many batch objects (2, 3, 4) and how many frame segments per call
(MK_SEGMENTS) give the best step at cfg 3, 24 steps per trial, 3 trials each.  Usage: python tools/e2e_depth_probe.py [DEPTH ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

import markov_b200 as mm
depths = [int(x) for x in sys.argv[1:]] or [2, 3]
B, T, D = 128, 150, 3000
K = mm.LogSemiring[np.float32]
fsm, pdf = mm.graphs.denominator(K)
c = mm.compile(fsm, mm.statemap(fsm, D, pdf))
bs = [mm.batch(*[c] * B) for _ in range(max(depths))]
Vh = torch.empty((B, T, D), pin_memory=True); Vh.copy_(torch.randn((B, T, D)) * 2)
Vn = Vh.numpy().transpose(0, 2, 1)
outs = []
for _ in range(max(depths)):
    p = torch.empty((T, D, B), pin_memory=True); t = torch.empty((B,), pin_memory=True)
    outs.append((p.numpy(), t.numpy(), p, t))


def run(n, depth):
    chk = 0.0
    for k in range(n):
        j = k % depth
        if k >= depth:
            bs[j].wait(); chk += float(outs[j][1].sum())
        mm.pdfposteriors(bs[j], Vn, out=outs[j][:2], wait=False)
    for k in range(max(0, n - depth), n):
        bs[k % depth].wait(); chk += float(outs[k % depth][1].sum())
    return chk


n = 24
for rep in range(3):
    for depth in depths:
        for seg in (0, 1, 2, 4):  # 0: the library's own choice
            if seg:
                os.environ["MK_SEGMENTS"] = str(seg)
            else:
                os.environ.pop("MK_SEGMENTS", None)
            run(2 * depth, depth)
            torch.cuda.synchronize()
            t0 = time.perf_counter(); run(n, depth); torch.cuda.synchronize()
            print("rep", rep, "depth", depth, "segments", seg, "ms/step %.3f" % ((time.perf_counter() - t0) * 1e3 / n), flush=True)
