import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distraytracer_b200 import runtime, abi, scenes
from oracle.harness import Oracle, ORACLE_KEYED, compare
scene, s = scenes.config2(480, 270, 16)
want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
dev = runtime.DeviceScene(scene, 0)
for prec in (0, 1):
    s.precision = prec
    got, _ = dev.render_float(s)
    print("config2 480x270x16 precision", prec, compare(want, got), flush=True)
