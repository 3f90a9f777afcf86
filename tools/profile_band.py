"""Small, representative slice of the bench workload for ncu (a 1920x48 band through the
doors / glass block at 64 spp): keeps profiler replay time in seconds, not minutes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distraytracer_b200 import runtime, abi, scenes
cfg = os.environ.get("BAND_CONFIG", "c2")        # c2: the bench workload; c5: the 999 698-triangle terrain at 4K (LBVH traversal); c3, c4
scene, st = (scenes.config5() if cfg == "c5" else scenes.config3() if cfg == "c3" else scenes.config4_frame(30) if cfg == "c4"
             else scenes.config2(1920, 1080, 64))
st.precision = int(os.environ.get("DRT_PRECISION", "0"))
dev = runtime.DeviceScene(scene, 0)
y0 = int(os.environ.get("BAND_Y0", "1000" if cfg == "c5" else "470")); h = int(os.environ.get("BAND_H", "16" if cfg == "c5" else "48"))
tile = abi.Tile(0, y0, st.xRes, h, 0)
cnt = abi.Counters()
for i in range(3):
    dev.render_device(st, tile, cnt)
    print(f"band y0={y0} h={h}: kernel {cnt.kernel_ms:.3f} ms, {cnt.kernel_launches} launches", flush=True)
if os.environ.get("BAND_COUNT"):
    c2 = abi.Counters(); c2.collect = 1
    dev.render_device(st, tile, c2)
    pt = list(c2.prim_tests)
    print(f"samples {c2.samples} rays {c2.rays} shadow {c2.shadow_rays} node {c2.node_tests} rect {pt[abi.PRIM_RECTANGLE]} sph {pt[abi.PRIM_SPHERE]} tri {pt[abi.PRIM_TRIANGLE]} shade {c2.shade_evals} kernel {c2.kernel_ms:.1f} ms")
