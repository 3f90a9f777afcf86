"""Mutated fixture scenes for the GPU-vs-oracle fuzz (tests/test_gpu_parity.py, tools/gpu_fuzz.py): stored fixture
scenes with random motion flags, BRDF models, roughness, reflective materials, glossy flags and settings.  The CPU twin
(tests/test_oracle_fuzz.py) pins the oracle on the compiled reference for the same kind of mutations."""
import numpy as np

from conftest import GOLDEN_CASES, load_case


def mutated_case(seed):
    """-> (fixture name, Scene, settings); deterministic in `seed`."""
    from distraytracer_b200 import abi
    from distraytracer_b200.scene import Scene
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
    return case, Scene(prims, scene.lights, scene.textures), s
