# SPDX-License-Identifier: MIT
"""Where does a ragged batch lose time?  cfg 3's graph, B = 256 (two utterance tiles), T = 150, four length
patterns; per-call time of pdfposteriors with the per-tile frame limits off (MK_RAGGED_CUT=0) so that every
pattern runs the same number of frame-tiles.  With a MK_PROFILE_BARRIER=1 build also the exact-fallback count."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import markov_b200 as mm
from bench_configs import timed

K = mm.LogSemiring[np.float32]
B, T, D = 256, 150, 3000
fsm, pdf = mm.graphs.denominator(K)
c = mm.compile(fsm, mm.statemap(fsm, D, pdf)); b = mm.batch(*[c] * B)
V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(1), device="cuda") * 2).permute(0, 2, 1)
post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
lib = C.CDLL(mm._lib.LIB_PATH)
dbg = hasattr(lib, "mk_debug_barrier_profile")
buf = (C.c_ulonglong * (148 * 4 + 1))()
pat = {
    "all 150": np.full(B, 150),
    "all 75": np.full(B, 75),
    "tile0 150, tile1 75": np.concatenate([np.full(128, 150), np.full(128, 75)]),
    "alternating 150/75 (mixed lanes)": np.tile([150, 75], B // 2),
    "U[75,150] as drawn": np.random.default_rng(0).integers(75, 151, B),
    "U[75,150] sorted": np.sort(np.random.default_rng(0).integers(75, 151, B))[::-1].copy(),
}
if len(sys.argv) > 1:  # pattern filter: substrings
    pat = {k: v for k, v in pat.items() if any(a in k for a in sys.argv[1:])}
for cut in (os.environ.get("MK_RAGGED_CUT"),) if os.environ.get("MK_RAGGED_CUT") else ("0", "1"):
    os.environ["MK_RAGGED_CUT"] = cut
    for name, lens in pat.items():
        lens = lens.astype(np.int32)
        ms = timed(lambda: mm.pdfposteriors(b, V, seqlengths=lens, out=(post, ttl)), n=3, warm=2)
        extra = ""
        if dbg:
            lib.mk_debug_barrier_profile(buf)
            mm.pdfposteriors(b, V, seqlengths=lens, out=(post, ttl))
            lib.mk_debug_barrier_profile(buf)
            extra = f"  exact-fallback events {buf[148 * 4]}"
        print(f"cut={cut}  {name:36s} {ms:8.2f} ms{extra}", flush=True)
