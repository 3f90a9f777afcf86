"""Writes distraytracer_b200/data/: the scenes the BASELINE configurations are built from (scenes.py), as scene-only
copies of the reference builders' exports in tests/golden/ (no golden images), plus the mocap clip.  The package must not
depend on the tests directory."""
import os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distraytracer_b200.scene import load_fixture, save_fixture
G = os.path.join(ROOT, "tests", "golden"); D = os.path.join(ROOT, "distraytracer_b200", "data")
os.makedirs(D, exist_ok=True)
for name in ("checkertexture", "reflectance", "chkpt2_mocap"):
    scene, settings, _ = load_fixture(os.path.join(G, name + ".npz"))
    save_fixture(os.path.join(D, name + "_scene.npz"), scene, settings)
for f in ("mocap_90.asf", "mocap_90_16_frames880_1000.amc", "mocap_bones_880_999.npy"):
    shutil.copyfile(os.path.join(G, f), os.path.join(D, f))
print(sorted(os.listdir(D)))
