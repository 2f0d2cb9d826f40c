import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fme_mod():
    import fme_loader
    return fme_loader.load()


@pytest.fixture(scope="session")
def orc():
    import oracle_bindings
    return oracle_bindings.oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own compiled code, or None when oracle/_ref/libhmref.so is unavailable."""
    import oracle_bindings
    return oracle_bindings.reference()
