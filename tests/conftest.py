import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ptb():
    """The product package; building libptb.so on demand (nvcc cross-compiles without a GPU)."""
    import importlib
    builder = importlib.import_module("distributed-path-tracer_b200.build")
    builder.build()
    return importlib.import_module("distributed-path-tracer_b200")


@pytest.fixture(scope="session")
def procedural(ptb):
    import importlib
    return importlib.import_module("distributed-path-tracer_b200.procedural")


@pytest.fixture(scope="session")
def portlib():
    import portlib as pl
    pl.build()
    return pl


@pytest.fixture(scope="session")
def reflib():
    import reflib as rl
    return rl


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
