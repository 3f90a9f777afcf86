import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import load_case
from distraytracer_b200 import runtime, abi
from oracle.harness import Oracle, ORACLE_KEYED
case = sys.argv[1]
scene, settings, _ = load_case(case)
for a in sys.argv[2:]:
    k, v = a.split("="); setattr(settings, k, type(getattr(settings, k))(float(v)))
want, wab, _, _ = Oracle(scene).render(settings, mode=ORACLE_KEYED)
got, _ = runtime.DeviceScene(scene, 0).render_float(settings)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"diag_{case}.npz"), want=want, got=got, wab=wab)
