import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native():
    import pslite_b200

    return pslite_b200.native()


@pytest.fixture(scope="session")
def built_native_tree():
    """make the C++ library, apps and unit tests once per session"""
    import subprocess

    subprocess.run(["make", "-j", "8", "all"], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "build")
