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
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from distraytracer_b200 import runtime, abi, scenes, shard  # noqa: E402


def frame_on(devs, st, rows, frame):
    return shard.render_frame_blocks(devs, st, frame, rows)


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
