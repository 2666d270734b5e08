import os
import sys
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fx(tmp_path_factory):
    """Unpacked tests/golden/catt_fixture.tar.gz (the reference's sample input, golden CSV and V/J references)."""
    d = tmp_path_factory.mktemp("catt_fixture")
    with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
        tar.extractall(d)
    return str(d)


@pytest.fixture(scope="session")
def trb(fx):
    from catt_b200.refg import chain_config
    return chain_config("hs", "TRB", "CDR3", os.path.join(fx, "resource"))
