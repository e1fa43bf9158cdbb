# SPDX-License-Identifier: MIT
"""Generate the committed golden fixtures from the reference tree (run in the build container,
where /root/reference exists; the GPU box only sees the committed .npz files).

  den_fsm_wsj.npz / num_fsm_wsj.npz — the reference's own benchmark graphs
      (/root/reference/misc/benchmark/{den,num}_fsm_wsj.txt, OpenFst text written by
      misc/benchmark/generatefsm.jl:42-57) converted to arrays: arcs (src, dst, -cost), initial
      and final weights, pdf id per state (0-based).  Data only, no reference source code.

The known answers that go with them are literals in
tests/test_oracle_golden.py.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import markov_b200 as mm  # noqa: E402

REF = "/root/reference/misc/benchmark"
HERE = os.path.dirname(os.path.abspath(__file__))


def convert(name):
    K = mm.LogSemiring[np.float64]
    fsm, pdfids = mm.graphs.load_openfst_text(os.path.join(REF, name + ".txt"), K)
    src, dst, w = fsm.arcs_hat()
    S = fsm.nstates
    real = (src < S) & (dst < S)
    fin = (src < S) & (dst == S)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), nstates=S, src=src[real].astype(np.int32), dst=dst[real].astype(np.int32),
        w=w[real].astype(np.float64), init_idx=fsm.init_idx.astype(np.int32), init_w=fsm.init_w.astype(np.float64),
        final_idx=src[fin].astype(np.int32), final_w=w[fin].astype(np.float64), pdfids=pdfids.astype(np.int32))
    print(name, fsm, "pdfs", pdfids.max() + 1)


if __name__ == "__main__":
    convert("den_fsm_wsj")
    convert("num_fsm_wsj")
