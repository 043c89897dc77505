import json
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_cases():
    with open(os.path.join(ROOT, "tests", "golden", "cases.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_lib():
    """Builds the oracle's C helper (gcc) once."""
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "libncc_rowdot.so"], check=True, capture_output=True)


@pytest.fixture(scope="session")
def fpm_built():
    from fastest_image_pattern_matching_b200 import build
    return build()


@pytest.fixture(scope="session")
def matcher(fpm_built):
    from fastest_image_pattern_matching_b200 import TemplateMatcher
    m = TemplateMatcher(0)
    yield m
    m.close()
