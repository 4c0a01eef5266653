import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import ilf_oracle
    ilf_oracle.build()
    return ilf_oracle


@pytest.fixture(scope="session")
def ilf_lib():
    """The product's CUDA library.  Built in-tree by __graft_entry__.build(); GPU tests fail loudly without it."""
    import vvcsoftware_vtm_b200 as v
    v.load_library()
    return v
