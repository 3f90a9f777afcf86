import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import load_case
from distraytracer_b200 import runtime, abi
from distraytracer_b200.scene import Scene
from oracle.harness import Oracle, ORACLE_KEYED, compare
scene, settings, _ = load_case("boundary_mocap")
def run(tag, prims, lights, st):
    sc = Scene(prims, lights, scene.textures)
    want, wab, _, _ = Oracle(sc).render(st, mode=ORACLE_KEYED)
    got, _ = runtime.DeviceScene(sc, 0).render_float(st)
    c = compare(want, got)
    print(f"{tag:30s} within1={c['frac_within_1']:.5f} nbad={c['n_bad']}", flush=True)
    return want, got
L = scene.lights
for sel in ([0], [1], [2], [0, 1], [0, 1, 2]):
    run(f"lights {sel}", scene.prims, [L[i] for i in sel], settings)
s2 = abi.copy_struct(settings); s2.reflect = 0
run("reflect=0", scene.prims, L, s2)
s3 = abi.copy_struct(settings); s3.nogloss = 1
run("nogloss=1", scene.prims, L, s3)
s4 = abi.copy_struct(settings); s4.antialias_samples = 1
w, g = run("aa=1", scene.prims, L, s4)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag2.npz"), want=w, got=g)
