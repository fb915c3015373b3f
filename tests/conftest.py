import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle_built():
    import oracle_lib as O
    O.build_oracle()
    return O


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = os.path.join(HERE, "golden")
    return {n[:-4]: np.load(os.path.join(g, n), allow_pickle=False) for n in os.listdir(g) if n.endswith(".npz")}
