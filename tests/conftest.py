import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The restated CPU matcher (oracle/matcher_oracle.cpp), built on demand."""
    from oracle import bindings
    if not os.path.exists(bindings.ORACLE_MATCHER_SO):
        bindings.build()
    return bindings.MatcherLib("oracle")


@pytest.fixture(scope="session")
def reference_lib():
    """The reference's chargrid.cpp compiled verbatim; skipped where it was never built."""
    from oracle import bindings
    if not bindings.have_reference():
        if os.path.isdir("/root/reference/src/matcher"):
            bindings.build()
        else:
            pytest.skip("oracle/_ref/libref_chargrid.so not present")
    return bindings.MatcherLib("reference")
