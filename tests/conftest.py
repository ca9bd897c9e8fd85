import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


GOLDEN_NAMES = ["arctic_a0001", "vaiueo2d", "synthetic48k_u7", "synthetic16k_u11"]


@pytest.fixture(scope="session")
def reference_lib():
    """The unmodified reference compiled by oracle/Makefile (test infrastructure only)."""
    from oracle import ref
    p = ref.ref_path()
    if not os.path.exists(p):
        if os.path.isdir("/root/reference/externs/WORLD_v2/src"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/libworld_ref.so"],
                           check=True, capture_output=True)
        else:
            pytest.skip("oracle/_ref/libworld_ref.so not built and /root/reference is absent")
    return ref.load()


@pytest.fixture(scope="session")
def wb():
    """The product library on a GPU box."""
    import hts_train_world_b200 as m
    m.init(0)
    return m
