import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ctx():
    """One afc_ctx on cuda:0 for the gpu-marked parity tests (fails loudly if the extension is missing)."""
    import agentfield_b200 as afb
    c = afb.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def hostsim():
    """TEST-ONLY CPU build of the kernel logic (tests/hostsim)."""
    import ctypes
    import subprocess
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    return ctypes.CDLL(os.path.join(d, "libafc_hostsim.so"))
