"""Quick parity/throughput table on a GPU box (diagnostic; the tests are the gate)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import GOLDEN_CASES, load_case
from distraytracer_b200 import runtime, abi
from oracle.harness import Oracle, ORACLE_KEYED, compare

for case in GOLDEN_CASES:
    scene, settings, _ = load_case(case)
    want, wab, _, osec = Oracle(scene).render(settings, mode=ORACLE_KEYED)
    dev = runtime.DeviceScene(scene, 0)
    for prec in (0, 1):
        settings.precision = prec
        cnt = abi.Counters()
        got, _ = dev.render_float(settings, counters=cnt)
        st = compare(want, got)
        print(f"{case:24s} prec={prec} within1={st['frac_within_1']:.5f} max={st['max']:3d} nbad={st['n_bad']:5d} "
              f"kernel={cnt.kernel_ms:8.3f} ms launches={cnt.kernel_launches} oracle={osec*1e3:.0f} ms", flush=True)
