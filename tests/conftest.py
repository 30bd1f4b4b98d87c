import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ctx():
    from dipper_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def make_msa(n, L, seed=1, regime="tiefree", **kw):
    from dipper_b200 import synth
    codes, info = synth.evolve(n, L, seed=seed, regime=regime, **kw)
    return codes, synth.pack4_np(codes), info
