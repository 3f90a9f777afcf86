"""Device-timed throughput of every BASELINE.json configuration on one GPU (diagnostic table;
bench.py is the contractual line for configs[1])."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from distraytracer_b200 import runtime, abi, scenes

def run(tag, scene, st, reps=2):
    t0 = time.time(); dev = runtime.DeviceScene(scene, 0); tb = time.time() - t0
    cnt = abi.Counters()
    spp = int(np.sqrt(st.antialias_samples)) ** 2
    n = st.xRes * st.yRes * spp
    ms = []
    for _ in range(reps + 1):
        dev.render_device(st, None, cnt); ms.append(cnt.kernel_ms)
    best = min(ms[1:])
    print(json.dumps({"config": tag, "res": [st.xRes, st.yRes], "spp": spp, "samples": n, "ms": best,
                      "Msamples_per_s": n / best / 1e3, "frames_per_s": 1e3 / best, "scene_create_s": tb,
                      "launches": cnt.kernel_launches}), flush=True)
    dev.close()

which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
if "c1" in which: run("C1 checkertexture 640x480 1spp", *scenes.config1())
if "c2" in which: run("C2 1080p 64spp all effects", *scenes.config2())
if "c3" in which: run("C3 Oren-Nayar + cloud background 1080p 256spp", *scenes.config3())
if "c4" in which: run("C4 mocap frame 30 velocity blur 1080p 16spp", *scenes.config4_frame(30))
if "c5" in which: run("C5 999698-triangle terrain 4K 64spp", *scenes.config5())


def mocap_video(n_frames=120, xres=1920, yres=1080, spp=16):
    """C4 end to end on one GPU: ASF/AMC -> drt_skeleton (all frames posed by skeleton_fk) -> per frame
    drt_scene_pose_skeleton + drt_render into a host frame (wall clock, includes H2D of the scene and D2H of the frame).
    Measured with one scene handle, and with two handles on two host threads taking alternate frames (what
    drt_host.h::renderVideo does): while one frame drains and is copied back, the other keeps the SMs busy."""
    import threading
    golden = os.path.join(ROOT, "tests", "golden")
    t0 = time.time()
    skel = runtime.DeviceSkeleton(os.path.join(golden, "mocap_90.asf"), os.path.join(golden, "mocap_90_16_frames880_1000.amc"))
    t_load = time.time() - t0
    scene, st0 = scenes.config4_frame(0, xres, yres, spp)
    first = next(i for i, p in enumerate(scene.prims) if p.type == abi.PRIM_CYLINDER)
    n = xres * yres * (int(np.sqrt(spp)) ** 2)
    for handles in (1, 2):
        devs = [runtime.DeviceScene(scene, 0) for _ in range(handles)]
        outs = [np.empty((yres, xres, 3), dtype=np.uint8) for _ in range(handles)]
        pose_s = [0.0] * handles

        def work(k, frames):
            st = abi.copy_struct(st0)
            for f in frames:
                st.frame = f; st.seed = 1000 + f
                tp = time.time(); devs[k].pose_skeleton(skel, f, first); pose_s[k] += time.time() - tp
                devs[k].render(st, out=outs[k])

        for k in range(handles):
            work(k, range(3))                              # warm-up
        pose_s = [0.0] * handles
        th = [threading.Thread(target=work, args=(k, range(k, n_frames, handles))) for k in range(handles)]
        t0 = time.time()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.time() - t0
        print(json.dumps({"config": f"C4 mocap video {n_frames} frames {xres}x{yres} {spp}spp velocity blur, skeleton path, e2e wall clock",
                          "scene_handles": handles, "frames_per_s": n_frames / dt, "Msamples_per_s": n * n_frames / dt / 1e6,
                          "ms_per_frame": 1e3 * dt / n_frames, "pose_ms_per_frame": 1e3 * sum(pose_s) / n_frames,
                          "skeleton_load_s": t_load, "fk_kernel_ms": skel.fk_ms, "clip_frames": skel.n_frames,
                          "bones": skel.n_cylinders}), flush=True)


if "c4video" in which: mocap_video()
