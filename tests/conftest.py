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
    # the slab-box prism classes: the reference's own `prismcyl` scene, and scenes that put RectPrism /
    # RectPrismWithCylinder / RectPrismWithHoles in front of its constructors (tests/golden/make_golden.py)
    "prismcyl", "prism_box", "prism_cyl", "prism_cyl_side", "prism_holes", "prism_holes_back",
]
# In the reference a hit on the wall of a prism's hole WRITES the hole's colour into the prism for good (geometry.cpp:1650,
# 2005, 2044): every later ray sees it, so its picture depends on the order the pixels are rendered in.  The oracle
# reproduces that to the bit under DRT_ORACLE_Q19=persist (sequential mode only); by default -- and in the CUDA path -- the
# colour belongs to the hit that set it.
Q19_CASES = {"prism_cyl", "prism_cyl_side", "prism_holes", "prism_holes_back"}


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
