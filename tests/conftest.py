import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # the oracle is test infrastructure: build it on demand (gcc is in both images)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    # the product library must already be built (__graft_entry__.build()); build it if the tree is fresh
    if not os.path.exists(os.path.join(ROOT, "metrovector_b200", "libmvf_b200.so")):
        subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "metrovector_b200", "csrc")], check=True)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def engine():
    """The GPU engine.  No skip: a GPU test on a box without the native path must fail loudly."""
    import metrovector_b200 as mv
    eng = mv.GpuEngine(0)
    yield eng
    eng.shutdown()
