# SPDX-License-Identifier: MIT
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden_fsm(name, K):
    """Rebuild an FSM from a committed tests/golden/*.npz fixture (made by make_golden.py)."""
    import markov_b200 as mm
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    fsm = mm.FSM.from_arrays(K, int(z["nstates"]), z["src"], z["dst"], z["w"].astype(K.dtype), z["init_idx"],
                             z["init_w"].astype(K.dtype), z["final_idx"], z["final_w"].astype(K.dtype))
    return fsm, z["pdfids"].astype(np.int64)


@pytest.fixture(scope="session")
def mm():
    import markov_b200
    return markov_b200


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle
