import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_CASES = [
    "hw4", "reflectance", "dof", "spherelight", "spheres_blur", "checkertexture", "checkertexture_nogloss",
    "texture", "textureog", "window", "staircase", "rectprism", "checkercylinder", "chkpt2_mocap", "boundary_mocap",
]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    """Build (if needed) and return the path of oracle/liboracle.so -- the checker, never the product."""
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("drt_oracle.cpp", "drt_skeleton_oracle.cpp")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(src) for src in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return so


def load_case(name):
    from distraytracer_b200.scene import load_fixture
    return load_fixture(os.path.join(GOLDEN, name + ".npz"))
