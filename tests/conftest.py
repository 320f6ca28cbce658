"""pytest configuration: `gpu` marker, shared fixtures.

-m "not gpu": the oracle against the reference-built library and the committed golden vectors, host logic
              (camera, lattice, loader), C-ABI surface, multi-rank strip logic on gloo.  No CUDA calls.
-m gpu:       parity of the CUDA path against the oracle, always through the C-ABI (libugrt.so).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """oracle/libugoracle.so (plain-C restatement) - built on demand from source."""
    from oracle import harness
    if not harness.available("oracle"):
        harness.build("oracle")
    return harness.load("oracle")


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref/libugref.so: the reference's own kernel sources compiled as host C++.  Built here when
    /root/reference exists; on the GPU box the prebuilt library travels with the snapshot."""
    from oracle import harness
    if not harness.available("ref") and os.path.isdir("/root/reference"):
        harness.build("ref")
    if not harness.available("ref"):
        pytest.skip("oracle/_ref/libugref.so not available (no /root/reference and no prebuilt library)")
    return harness.load("ref")


@pytest.fixture(scope="session")
def ug():
    """The product package; libugrt.so must exist (built by __graft_entry__.build())."""
    import ugrt  # noqa: F401
    import uniformgrid_raytracing_b200 as pkg
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    pkg.lib()
    return pkg


@pytest.fixture(scope="session")
def renderer(ug):
    r = ug.Renderer(0)
    yield r
    r.close()
