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
    O.build()
    return O


@pytest.fixture(scope="session")
def engines():
    """One engine per precision mode, shared by the GPU tests."""
    import imageprocessor_b200 as ip
    made = {}

    def get(precision=ip.PRECISION_EXACT, **kw):
        key = (precision, tuple(sorted(kw.items())))
        if key not in made:
            made[key] = ip.Engine(devices=[0], precision=precision, **kw)
        return made[key]

    yield get
    for e in made.values():
        e.close()
