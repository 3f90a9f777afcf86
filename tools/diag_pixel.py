"""One pixel of a mutated-scene seed, GPU vs oracle, with both sides' event counters.
usage: python tools/diag_pixel.py <seed> <image row> <image column> [aa]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fuzz_cases import mutated_case
from distraytracer_b200 import runtime, abi
from oracle.harness import Oracle, ORACLE_KEYED

seed, row, col = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
case, sc, s = mutated_case(seed)
if len(sys.argv) > 4: s.antialias_samples = int(sys.argv[4])
s.max_depth = 1; s.blur_samples = 0
tile = abi.Tile(col, s.yRes - 1 - row, 1, 1, 0)
want, _, oc, _ = Oracle(sc).render(s, tile, mode=ORACLE_KEYED)
gc = abi.Counters(); gc.collect = 1
got, _ = runtime.DeviceScene(sc, 0).render_float(s, tile, counters=gc)
f = lambda c: dict(rays=c.rays, shadow=c.shadow_rays, shade=c.shade_evals, nodes=c.node_tests, tests=list(c.prim_tests))
print("oracle", want.ravel().round(2).tolist(), f(oc))
print("gpu   ", got.ravel().round(2).tolist(), f(gc))
