import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "causalgpslc.jl_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def kats():
    import json
    return json.load(open(os.path.join(GOLDEN, "reference_kats.json")))


@pytest.fixture(scope="session")
def lib_built():
    """Build the shared library once per session if it is missing (nvcc cross-compiles without a GPU)."""
    from gpslc_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import subprocess
        subprocess.run(["bash", os.path.join(ROOT, "causalgpslc.jl_b200", "build.sh")], check=True)
    return _lib.load()


@pytest.fixture(scope="session")
def ctx(lib_built):
    import gpslc_b200
    return gpslc_b200.Context(0)
