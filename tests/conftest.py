import importlib
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def kfd():
    """Unpacked reference kfd fixture trees (tests/golden/*.tar.gz)."""
    import kfd_fixtures
    yield kfd_fixtures
    kfd_fixtures.cleanup()


@pytest.fixture(scope="session")
def pkg():
    """The product package; its directory name has a hyphen, so import by string."""
    return importlib.import_module("k8s-device-plugin_b200")


@pytest.fixture
def short_dir():
    """A short scratch dir for unix sockets (sun_path is limited to 107 chars)."""
    import shutil
    import tempfile
    d = tempfile.mkdtemp(prefix="b2s_", dir="/tmp")
    yield d
    shutil.rmtree(d, ignore_errors=True)
