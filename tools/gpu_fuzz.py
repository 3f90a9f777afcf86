"""GPU vs oracle on mutated fixture scenes (diagnostic for a GPU box; not collected by pytest yet -- written when this
round's GPU budget was spent, so it has not run; promote it to tests/ once it has).

The CPU twin (tests/test_oracle_fuzz.py) pins the oracle on the compiled reference for the same kind of mutations; this
script closes the loop for the kernels: stored fixture scenes with random motion flags, BRDF models, roughness,
reflective materials, glossy flags and settings, rendered by libdrt.so and by the oracle under the keyed stream.
  python tools/gpu_fuzz.py [n_cases]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from conftest import GOLDEN_CASES, load_case  # noqa: E402
from distraytracer_b200 import runtime, abi  # noqa: E402
from distraytracer_b200.scene import Scene  # noqa: E402
from oracle.harness import Oracle, ORACLE_KEYED, compare  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
for seed in range(n_cases):
    rng = np.random.default_rng(9000 + seed)
    case = GOLDEN_CASES[int(rng.integers(len(GOLDEN_CASES)))]
    scene, settings, _ = load_case(case)
    s = abi.copy_struct(settings)
    s.xRes, s.yRes = int(rng.integers(24, 65)), int(rng.integers(18, 49))
    s.antialias_samples = int(rng.choice([1, 4, 9]))
    s.aperture = float(rng.choice([0.0, 0.2]))
    s.brdf_samples = int(rng.integers(1, 4))
    s.max_depth = int(rng.integers(1, 7))
    s.blur_samples = int(rng.integers(0, 4))
    s.frame_range = int(rng.integers(1, 9))
    if rng.random() < 0.4:
        s.frame_prism, s.frame_blur = 0, int(rng.choice([0, 100000]))
    s.seed = int(rng.integers(1, 1 << 30))
    prims = [abi.copy_struct(p) for p in scene.prims]
    for p in prims:
        if p.flags & abi.FLAG_LIGHT:
            continue
        if rng.random() < 0.3:
            p.flags ^= abi.FLAG_MOTION
        if rng.random() < 0.5:
            p.model = int(rng.choice([abi.MODEL_LAMBERT, abi.MODEL_OREN_NAYAR, abi.MODEL_COOK_TORRANCE]))
            p.roughness = float(np.float32(rng.uniform(0.1, 0.9)))
            p.refr[0], p.refr[1] = 0.958, 6.69
        if rng.random() < 0.3:
            p.material = int(rng.choice([abi.MAT_NONE, abi.MAT_STEEL, abi.MAT_ALUMINUM, abi.MAT_LINOLEUM]))
            if rng.random() < 0.5:
                p.flags ^= abi.FLAG_GLOSSY
    sc = Scene(prims, scene.lights, scene.textures)
    want, _, _, _ = Oracle(sc).render(s, mode=ORACLE_KEYED)
    got, _ = runtime.DeviceScene(sc, 0).render_float(s)
    st = compare(want, got)
    ok = st["frac_within_1"] >= 0.999
    bad += not ok
    print(f"seed {seed:3d} {case:24s} {s.xRes}x{s.yRes} aa={s.antialias_samples} depth={s.max_depth} lobes={s.brdf_samples} "
          f"blur={s.blur_samples} within1={st['frac_within_1']:.5f} max={st['max']} {'ok' if ok else 'MISMATCH'}", flush=True)
print("mismatching cases:", bad)
sys.exit(1 if bad else 0)
