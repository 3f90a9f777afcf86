"""Diagnose one mutated-scene seed: where GPU and oracle differ, and which setting makes the difference go away."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fuzz_cases import mutated_case
from distraytracer_b200 import runtime, abi
from oracle.harness import Oracle, ORACLE_KEYED, compare

seed = int(sys.argv[1])
case, sc, s0 = mutated_case(seed)

def run(tag, mut):
    s = abi.copy_struct(s0); mut(s)
    want, wab, _, _ = Oracle(sc).render(s, mode=ORACLE_KEYED)
    got, _ = runtime.DeviceScene(sc, 0).render_float(s)
    d = np.abs(np.floor(np.nan_to_num(want)) - np.floor(np.nan_to_num(got))).max(axis=-1)
    bad = np.argwhere(d > 1)
    print(f"{tag:28s} differing pixels {len(bad)}: " + " ".join(f"({y},{x}) o={want[y,x].round(2).tolist()} g={got[y,x].round(2).tolist()}" for y, x in bad[:4]), flush=True)
    return bad

run("as is", lambda s: None)
run("blur_samples 0", lambda s: setattr(s, "blur_samples", 0))
run("aperture 0", lambda s: setattr(s, "aperture", 0.0))
run("aa 1", lambda s: setattr(s, "antialias_samples", 1))
run("max_depth 1", lambda s: setattr(s, "max_depth", 1))
run("reflect 0", lambda s: setattr(s, "reflect", 0))
run("frame_range 1", lambda s: setattr(s, "frame_range", 1))
run("frame_blur 100000", lambda s: setattr(s, "frame_blur", 100000))
run("fp32", lambda s: setattr(s, "precision", 1))
