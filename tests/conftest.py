import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def rl():
    import pyraylib
    from oracle import bindings as ob
    if not os.path.exists(pyraylib.PRODUCT_LIB) or not os.path.exists(ob.RESTATE_LIB):
        import __graft_entry__
        __graft_entry__.build()
    return pyraylib


@pytest.fixture(scope="session")
def prod(rl):
    p = rl.Product()
    p.lib.Raylib_Initialize()
    yield p
    p.lib.Raylib_Terminate()


@pytest.fixture(scope="session")
def gpu(prod):
    if prod.device_count() <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible (the product has no CPU path)")
    return prod


@pytest.fixture(scope="session")
def ref(rl):
    """The compiled reference (oracle/_ref). Present in the build container and shipped to the GPU box."""
    from oracle import bindings as ob
    if not os.path.exists(ob.REF_LIB) or not os.path.exists(ob.REF_SCENES):
        pytest.skip("oracle/_ref not built (needs /root/reference); golden fixtures cover this host")
    return ob.Reference()


@pytest.fixture(scope="session")
def restate(rl):
    from oracle import bindings as ob
    return ob.Restatement()


def load_golden(cfg):
    return np.load(os.path.join(GOLDEN, "config%d.npz" % cfg))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
