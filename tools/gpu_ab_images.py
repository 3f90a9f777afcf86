"""Are two builds of libdrt.so bit-identical?  Renders fixture scenes and BASELINE configurations (small) with the in-tree library and
with $DRT_LIB_B in a second process and compares the float images exactly.  usage: DRT_LIB_B=variants/libdrt_x.so python tools/gpu_ab_images.py"""
import os, subprocess, sys, pickle
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    from conftest import GOLDEN_CASES, load_case
    from fuzz_cases import mutated_case, random_scene
    from distraytracer_b200 import runtime, scenes
    out = {}
    for c in GOLDEN_CASES:
        sc, s, _ = load_case(c); out[c] = runtime.DeviceScene(sc, 0).render_float(s)[0]
    for seed in range(40):
        c, sc, s = mutated_case(seed); out[f"mut{seed}"] = runtime.DeviceScene(sc, 0).render_float(s)[0]
        c, sc, s = random_scene(seed); out[f"rnd{seed}"] = runtime.DeviceScene(sc, 0).render_float(s)[0]
    sc, s = scenes.config4_frame(30, 320, 180, 16); out["c4"] = runtime.DeviceScene(sc, 0).render_float(s)[0]
    pickle.dump(out, open(sys.argv[2], "wb")); sys.exit(0)
def run(lib, path):
    env = dict(os.environ); 
    if lib: env["DRT_LIB"] = os.path.abspath(lib)
    subprocess.check_call([sys.executable, __file__, "worker", path], env=env)
    return pickle.load(open(path, "rb"))
a = run(None, "/tmp/ab_a.pkl"); b = run(os.environ["DRT_LIB_B"], "/tmp/ab_b.pkl")
diff = [k for k in a if not np.array_equal(a[k], b[k], equal_nan=True)]
print(len(a), "images;", "bit-identical" if not diff else f"DIFFERENT: {diff}")
