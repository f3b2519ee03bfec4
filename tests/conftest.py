import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The scalar C restatement (test infrastructure). Built on demand by oracle/Makefile."""
    import subprocess
    from _util import ORACLE_SO, OpalCLibrary
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return OpalCLibrary(ORACLE_SO)


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference, when oracle/_ref/ was built (needs /root/reference at build time)."""
    from _util import REF_SO, OpalCLibrary
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libopal_ref.so not built")
    return OpalCLibrary(REF_SO)


@pytest.fixture(scope="session")
def product():
    """The CUDA library through its C ABI. No fallback: a missing .so or device is a failure."""
    from opal_b200.handle import OpalB200
    return OpalB200()
