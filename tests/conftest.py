import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    import numpy

    def load(name):
        return numpy.load(os.path.join(GOLDEN, name + '.npz'))
    return load


@pytest.fixture(scope='session', autouse=True)
def built_artefacts():
    """The shared libraries are build products (git-ignored): build them once if a clean checkout has none, so that
    `pytest -m "not gpu"` works without a prior `__graft_entry__.build()` (nvcc cross-compiles without a GPU)."""
    needed = [os.path.join(ROOT, 'nufhe_b200', 'csrc', 'libnufhe_b200.so'),
              os.path.join(ROOT, 'nufhe_b200', 'csrc', 'libnb_host_emul.so'),
              os.path.join(ROOT, 'oracle', 'libnufhe_oracle.so')]
    if not all(os.path.exists(p) for p in needed):
        import __graft_entry__ as g
        g.build()
