import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    """Build the oracle .so (gcc) and libmcig.so (nvcc) if missing: both are git-ignored artefacts."""
    import orc
    if not os.path.exists(orc.ORACLE_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    from mcintegratorplusplus_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        from mcintegratorplusplus_b200 import build
        build.build()


@pytest.fixture(scope="session")
def oracle():
    _ensure_built()
    import orc
    return orc.oracle()


@pytest.fixture(scope="session")
def golden_runs():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "ref_runs.json")))


@pytest.fixture(scope="session")
def golden_callback():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "ref_callback.json")))


@pytest.fixture(scope="session")
def golden_est():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "ref_estimators.json")))


@pytest.fixture(scope="session")
def mcig():
    _ensure_built()
    import mcintegratorplusplus_b200 as m
    return m


def fromhex(v):
    return [float.fromhex(s) for s in v]
