import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kats():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session", params=["ref", "plain"])
def oracle_lib(request):
    """Both oracle builds: the one linked against the reference's own C++ and the stand-alone one."""
    from oracle import pyoracle

    ref, plain = pyoracle.lib_paths()
    path = ref if request.param == "ref" else plain
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} not built")
    return pyoracle.load(prefer_ref=(request.param == "ref"))
