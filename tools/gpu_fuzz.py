"""GPU vs oracle on mutated fixture scenes, any number of seeds (tests/test_gpu_parity.py runs the first 14).
  python tools/gpu_fuzz.py [n_cases]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fuzz_cases import mutated_case  # noqa: E402
from distraytracer_b200 import runtime  # noqa: E402
from oracle.harness import Oracle, ORACLE_KEYED, compare  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
for seed in range(n_cases):
    case, sc, s = mutated_case(seed)
    want, _, _, _ = Oracle(sc).render(s, mode=ORACLE_KEYED)
    got, _ = runtime.DeviceScene(sc, 0).render_float(s)
    st = compare(want, got)
    ok = st["frac_within_1"] >= 0.999
    bad += not ok
    print(f"seed {seed:3d} {case:24s} {s.xRes}x{s.yRes} aa={s.antialias_samples} depth={s.max_depth} lobes={s.brdf_samples} "
          f"blur={s.blur_samples} within1={st['frac_within_1']:.5f} max={st['max']} {'ok' if ok else 'MISMATCH'}", flush=True)
print("mismatching cases:", bad)
sys.exit(1 if bad else 0)
