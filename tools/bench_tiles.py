"""One frame cut across the GPUs of a box (BASELINE configs 3 and 5: "single frame tiled across 8 GPUs").

One host thread per GPU (ctypes releases the GIL inside drt_render) claims blocks of `--rows` image rows from a shared
counter -- sky rows and object rows differ in cost by orders of magnitude -- and renders each block straight into its
place in the pinned host frame.  The scene is replicated; there is no collective.  The same partition is what
drt_host.h::renderFrame does in C++.  Wall clock around the whole frame, after one warm-up frame, for 1 GPU and for all.

  python tools/bench_tiles.py [c2|c3|c5] [--rows 15] [--streams 2] [--gpus N]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from distraytracer_b200 import runtime, abi, scenes, shard  # noqa: E402


def frame_on(devs, st, rows, frame):
    yres = st.yRes
    blocks = [(y0, min(y0 + rows, yres)) for y0 in range(0, yres, rows)]
    nxt = [0]
    lock = threading.Lock()
    errs = []

    def work(dev):
        try:
            while True:
                with lock:
                    k = nxt[0]; nxt[0] += 1
                if k >= len(blocks):
                    return
                y0, y1 = blocks[k]
                out = frame[yres - y1: yres - y0]                       # PPM row order (shard.place_band)
                dev.render(st, abi.Tile(0, y0, st.xRes, y1 - y0, dev.device), out=out)
        except Exception as e:                                          # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(d,)) for d in devs]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    if errs:
        raise errs[0]
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="c5")
    ap.add_argument("--rows", type=int, default=15)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--streams", type=int, default=2, help="scene handles (streams) per GPU: the drain of one block overlaps the ramp of the next")
    args = ap.parse_args()
    scene, st = {"c2": scenes.config2, "c3": scenes.config3, "c5": scenes.config5}[args.config]()
    n = args.gpus or runtime.device_count()
    spp = int(np.sqrt(st.antialias_samples)) ** 2
    samples = st.xRes * st.yRes * spp
    import torch
    frame = torch.empty((st.yRes, st.xRes, 3), dtype=torch.uint8).pin_memory().numpy()
    devs = [runtime.DeviceScene(scene, d) for d in range(n) for _ in range(args.streams)]   # GPU-major
    ref = None
    for use in ([1, n] if n > 1 else [1]):
        frame_on(devs[:use * args.streams], st, args.rows, frame)       # warm-up (scratch allocation, clocks)
        dt = min(frame_on(devs[:use * args.streams], st, args.rows, frame) for _ in range(2))
        if use == 1:
            ref = frame.copy(); t1 = dt
        print(json.dumps({"config": args.config, "res": [st.xRes, st.yRes], "spp": spp, "gpus": use, "streams_per_gpu": args.streams, "rows_per_block": args.rows,
                          "ms_per_frame": 1e3 * dt, "frames_per_s": 1 / dt, "Msamples_per_s": samples / dt / 1e6,
                          "speedup_vs_1gpu": t1 / dt, "same_image_as_1gpu": bool(np.array_equal(ref, frame))}), flush=True)


if __name__ == "__main__":
    main()
