import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port():
    from oracle import oracle as O
    return O.load_port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref/libdgref.so not built (needs /root/reference)")
    return O.load_ref()


def _oracles():
    from oracle import oracle as O
    out = ["port"]
    if O.have_ref():
        out.append("ref")
    return out


@pytest.fixture(scope="session", params=_oracles())
def orc(request):
    """Every known-answer test runs on the C restatement and, when built, on the unmodified reference."""
    from oracle import oracle as O
    return O.load_port() if request.param == "port" else O.load_ref()
