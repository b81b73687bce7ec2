import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """libsemb.so is git-ignored: build it (nvcc cross-compiles without a GPU) if a fresh checkout lacks it."""
    lib = os.path.join(ROOT, "spectralelements.jl_b200", "lib", "libsemb.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    return lib


@pytest.fixture(scope="session")
def sem():
    """The product package (ctypes over libsemb.so).  Fails loudly if the library cannot be built / loaded."""
    _ensure_built()
    import spectralelements_jl_b200 as sem
    sem._lib.load()
    return sem


@pytest.fixture(scope="session")
def ctx(sem):
    c = sem.init(0)
    yield c
    sem.finalize()
