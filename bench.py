#!/usr/bin/env python
"""bench.py -- throughput of the distributed ray-tracing hot path.

Workload (BASELINE.json configs[1]): the reference's checkertexture scene at 1920x1080,
64 spp, with depth of field, glossy reflection, rectangle-light soft shadows and glass
refraction (distraytracer_b200.scenes.config2).  One "step" = one full frame
(132.7 M camera samples).  With N GPUs every rank renders its own frames (frames sharded
across ranks, no data-path collective): weak scaling.

  value : whole-job Msamples/s with the scene already resident in HBM, device-timed
          (CUDA events on the launching stream inside libdrt.so, max over ranks).
  e2e   : the same metric through drt_render() with HOST buffers: per step the scene
          primitives are re-uploaded (drt_scene_update_prims, H2D) and the finished
          u8 frame is copied back (D2H), host wall clock around the calls.
  --impl reference : the reference's own CPU renderer (oracle/_ref, the unmodified
          reference sources compiled here) on all host cores, one process per core on
          disjoint row bands of the same frame.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

XRES, YRES, SPP = 1920, 1080, 64
METRIC = "samples_per_s_1080p_64spp"
UNIT = "Msamples/s"

# algorithmic cost table, SURVEY.md 8(d): FP32-equivalent ops per event
COST = {"node": 20, "sphere": 32, "triangle": 52, "rectangle": 36, "cylinder": 70, "shade": 110, "ray": 60}


def workload():
    from distraytracer_b200 import scenes
    scene, settings = scenes.config2(XRES, YRES, SPP)
    return scene, settings


def config_dict(n_gpus):
    return {
        "workload": "configs[1]: checkertexture scene 1920x1080 64spp, DOF aperture 0.2 / focal 10, glossy floor+doors, "
                    "rectangle-light soft shadows, glass block (Fresnel refraction), max_depth 10, brdf_samples 2",
        "samples_per_step": XRES * YRES * SPP,
        "partition": f"frames round-robin over {n_gpus} rank(s), no collective; scene replicated",
        "precision": "reference: f64 vectors, f32 scalar temporaries where the reference narrows, no FMA contraction",
        "l2": "working set is the 2.1 GB per-frame sample buffer (> 126 MB L2); no explicit flush",
    }


# ---------------------------------------------------------------------------
# clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.dev = device_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            # under load = samples at or above the median of the upper half
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------
# CPU reference leg
_WORKER = {}


def _ref_worker(args):
    """One task = one row of the frame on one reference renderer instance (the reference is not
    thread-safe: one process per core, state cached per process)."""
    scene_npz, settings_bytes, y0, x0, x1, use_ref = args
    y1 = y0 + 1
    if "st" not in _WORKER:
        sys.path.insert(0, ROOT)
        from distraytracer_b200.scene import Scene, settings_from_bytes
        import numpy as _np
        scene = Scene.from_npz_dict(_np.load(scene_npz))
        st = settings_from_bytes(settings_bytes)
        _WORKER["st"] = st
        if use_ref:
            from oracle.harness import Ref
            r = Ref()
            r.load(scene)
            r.set_settings(st)
            r.rng(0)                  # the reference's own mt19937 / random_device draws
            _WORKER["ref"] = r
        else:
            from oracle.harness import Oracle
            _WORKER["oracle"] = Oracle(scene)
    st = _WORKER["st"]
    if use_ref:
        _, _, sec = _WORKER["ref"].render_loop(st.frame, y0, y1, reset_policy=0, seed=0, x0=x0, x1=x1)
    else:
        from oracle.harness import ORACLE_KEYED
        from distraytracer_b200 import abi
        _, _, _, sec = _WORKER["oracle"].render(st, abi.Tile(x0, y0, x1 - x0, y1 - y0, 0), mode=ORACLE_KEYED)
    return sec


class CpuReference:
    """Times the reference's CPU implementation on a bounded sample: `rows_per_core` rows per
    host core, bands spread over the frame so sky, floor, doors and glass are all sampled."""

    def __init__(self, scene, settings, rows_per_core=1):
        import multiprocessing as mp
        from oracle.harness import ref_available
        from distraytracer_b200.scene import settings_to_bytes
        self.cores = os.cpu_count() or 1
        self.kind = "reference" if ref_available() else "port"
        self.rows_per_core = rows_per_core
        self.tmp = tempfile.mktemp(suffix=".npz")
        np.savez(self.tmp, **scene.to_npz_dict())
        self.settings = settings
        self.sbytes = settings_to_bytes(settings)
        self.pool = mp.get_context("spawn").Pool(self.cores)
        # stratified sample of the frame: rows evenly spaced over the height, and in each row 8
        # segments of `seg_px` pixels evenly spaced over the width; (row, segment) tasks are handed
        # out dynamically (some pixels cost 1000x others: glass + glossy ray trees)
        self.n_rows = self.cores * rows_per_core
        self.seg_px = 30
        ys = np.linspace(0, settings.yRes - 1, self.n_rows).astype(int)
        xs = np.linspace(0, settings.xRes - self.seg_px, 8).astype(int)
        self.bands = [(int(y), int(x), int(x) + self.seg_px) for y in ys for x in xs]
        self.samples = len(self.bands) * self.seg_px * (int(np.sqrt(settings.antialias_samples)) ** 2)

    def step(self):
        t0 = time.perf_counter()
        list(self.pool.imap_unordered(_ref_worker, [(self.tmp, self.sbytes, y, a, b, self.kind == "reference")
                                                    for y, a, b in self.bands], chunksize=1))
        return time.perf_counter() - t0

    def close(self):
        self.pool.close(); self.pool.join()
        try:
            os.unlink(self.tmp)
        except OSError:
            pass

    def describe(self):
        return (f"stratified sample of the {self.settings.xRes}x{self.settings.yRes} {SPP}spp frame: {self.n_rows} rows x 8 segments "
                f"of {self.seg_px} px = {self.samples} samples per step, {len(self.bands)} tasks over {self.cores} processes "
                f"(one per core); "
                + ("oracle/_ref (unmodified reference sources, its own RNG)" if self.kind == "reference"
                   else "oracle/ C++ restatement (reference tree not built)"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene, settings = workload()
    ref = CpuReference(scene, settings, rows_per_core=args.ref_rows)
    for _ in range(args.warmup):
        ref.step()
    t = 0.0
    for _ in range(args.steps):
        t += ref.step()
    ref.close()
    value = ref.samples * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus),
        "frames_per_s": value * 1e6 / (XRES * YRES * SPP),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": ref.describe()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from distraytracer_b200 import runtime, abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or runtime.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: distraytracer_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scene, settings = workload()
    dev = runtime.DeviceScene(scene, local)
    tile = abi.Tile(0, 0, settings.xRes, settings.yRes, local)
    samples_per_step = XRES * YRES * SPP
    frame_bytes = XRES * YRES * 3
    host_frame = torch.empty((YRES, XRES, 3), dtype=torch.uint8).pin_memory().numpy()
    prims = list(scene.prims)
    h2d_bytes = None

    from distraytracer_b200 import shard
    my_frames = shard.frames_for_rank(rank, world, (args.warmup + args.steps) * world)

    def set_frame(step):
        # frames are sharded round-robin: rank r renders frames r, r+world, ...
        settings.frame = my_frames[step]
        settings.seed = 1000 + settings.frame

    # ---- device-resident leg (value) ------------------------------------------------
    cnt = abi.Counters()
    for w in range(args.warmup):
        set_frame(w)
        dev.render_device(settings, tile, cnt)
    clocks = ClockSampler(local) if rank == 0 else None
    barrier()
    if clocks:
        clocks.start()
    t_wall0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for k in range(args.steps):
        set_frame(args.warmup + k)
        dev.render_device(settings, tile, cnt)          # synchronises its stream; kernel_ms from CUDA events
        dev_ms += cnt.kernel_ms
        launches += cnt.kernel_launches
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clock_info = clocks.stop() if clocks else None

    # ---- end-to-end leg (e2e): host buffers in, host frame out ----------------------
    for w in range(min(args.warmup, 2)):
        dev.update_prims(prims); dev.render(settings, tile, out=host_frame)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        set_frame(args.warmup + k)
        dev.update_prims(prims)                          # H2D: the frame's primitives
        dev.render(settings, tile, out=host_frame)       # kernels + D2H of the u8 frame
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)

    def maxr(x):
        return shard.max_over_ranks(x, dist if world > 1 else None, device="cuda")

    dev_ms, wall_ms, e2e_ms = maxr(dev_ms), maxr(wall_ms), maxr(e2e_ms)

    if rank == 0:
        total = samples_per_step * args.steps * world
        value = total / (dev_ms * 1e-3) / 1e6
        e2e_value = total / (e2e_ms * 1e-3) / 1e6
        # informational: the same frame through the single-precision build of the kernels (settings.precision = FP32;
        # passes the <= 1/255 on >= 99.9 % of pixels bar on every fixture except the degenerate boundary scene)
        fp32_info = None
        try:
            s32 = abi.copy_struct(settings); s32.precision = abi.PRECISION_FP32
            c32 = abi.Counters()
            dev.render_device(s32, tile, c32); dev.render_device(s32, tile, c32)
            fp32_info = {"value": samples_per_step / (c32.kernel_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": c32.kernel_ms,
                         "note": "one GPU, device-timed, not the headline: the headline runs in the reference's precision"}
        except Exception as e:
            fp32_info = {"error": str(e)}
        # one untimed instrumented frame for the roofline accounting
        c2 = abi.Counters(); c2.collect = 1
        events_note = None
        try:
            dev.render_device(settings, tile, c2)
        except Exception as e:      # the timed numbers above stand on their own; report the accounting as missing
            c2 = abi.Counters(); events_note = f"instrumented frame failed: {e}"
        pt = list(c2.prim_tests)
        ops = (c2.node_tests * COST["node"] + pt[abi.PRIM_SPHERE] * COST["sphere"] + pt[abi.PRIM_TRIANGLE] * COST["triangle"]
               + pt[abi.PRIM_RECTANGLE] * COST["rectangle"] + pt[abi.PRIM_CYLINDER] * COST["cylinder"]
               + c2.shade_evals * COST["shade"] + (c2.rays + c2.shadow_rays) * COST["ray"])
        step_s = dev_ms * 1e-3 / args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        sm_mhz = (clock_info or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        props = torch.cuda.get_device_properties(local)
        # algorithmic bytes per frame: 16 B sample record written by render_samples and read once by
        # resolve, plus the 3 B/pixel frame (the scene itself lives in L1/L2)
        alg_bytes = samples_per_step * 32 + frame_bytes
        # dram__bytes_read.sum + dram__bytes_write.sum of the render_wave launch of one step, from the ncu launch list of
        # this same command (profiles/r1_launches_dram_bench_one_launch.csv): 64.44 GB read + 146.65 GB written.
        # The excess over the algorithmic bytes is the CTA ray pools and hit buffers (64 B per pushed ray, 80 B per hit,
        # each written once and read once, streamed with .cs): with 8192-hit passes the 148 CTAs' working set (~350 MB)
        # exceeds the L2.
        wave_launches_per_step = max(1, launches // max(1, args.steps) // 2)     # render_wave + resolve per row chunk
        measured_traffic = (64438689536 + 146654461952) // wave_launches_per_step
        fp64_peak = props.multi_processor_count * 64 * 2 * sm_mhz * 1e6 / 1e12     # TFLOP/s at the observed clock
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(world),
            "frames_per_s": value * 1e6 / samples_per_step,
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_scene_bytes(dev, scene)),
                    "d2h_bytes_per_step": frame_bytes, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": clock_info,
            "roofline": {"bound": "hbm", "achieved": alg_bytes / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": alg_bytes / step_s / 1e9 / hbm_peak, "traffic": measured_traffic,
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                         "launches_per_step": wave_launches_per_step,
                         "note": "kernel render_wave<double>, one launch per step (a frame fits one row chunk of <= 2^28 samples): achieved = algorithmic bytes of the launch / its device time (CUDA events on the library's stream; resolve is 0.5 ms of the 207), traffic = DRAM bytes of the launch (ncu). The path is issue/latency bound in the FP64+ALU pipes, not HBM bound: see roofline_fp; traffic = ray-pool and hit-buffer spill, 1.0 TB/s"},
            "fp32_variant": fp32_info,
            # what actually bounds the kernel: warp-instruction issue.  Instructions per frame from the same ncu launch list
            # as `traffic` (smsp__inst_executed.sum = 96.2 G for render_wave<double>), time measured live.
            "roofline_issue": {"bound": "warp_issue", "achieved": 96205331080 / step_s / 1e9,
                               "peak": props.multi_processor_count * 4 * sm_mhz * 1e6 / 1e9, "unit": "Gwarp-inst/s",
                               "frac": 96205331080 / step_s / (props.multi_processor_count * 4 * sm_mhz * 1e6),
                               "note": "peak = SMs x 4 schedulers x observed SM clock (one warp instruction per scheduler per "
                                       "cycle); 26 of 32 lanes are active per instruction (profiles/r1_ncu_full_band_wave_final.txt)"},
            "roofline_fp": {"bound": "fp64_pipe", "achieved": ops / step_s / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": ops / step_s / 1e12 / fp64_peak,
                            "note": events_note or "algorithmic ops = event counters x SURVEY.md 8(d) cost table; peak = SMs x 64 DFMA x 2 x observed SM clock",
                            "events": {"samples": c2.samples, "rays": c2.rays, "shadow_rays": c2.shadow_rays, "node_tests": c2.node_tests,
                                       "rect_tests": pt[abi.PRIM_RECTANGLE], "sphere_tests": pt[abi.PRIM_SPHERE],
                                       "tri_tests": pt[abi.PRIM_TRIANGLE], "cyl_tests": pt[abi.PRIM_CYLINDER], "shade_evals": c2.shade_evals}},
        }
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(scene, settings, rows_per_core=2 * args.ref_rows)
            sec = ref.step()
            ref.close()
            line["cpu_baseline"] = {"value": ref.samples / sec / 1e6, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                                    "sample": ref.describe(), "seconds": sec}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def h2d_scene_bytes(dev, scene):
    """Bytes drt_scene_update_prims uploads per call (both precisions' geom/prim/node/light tables
    are small; counted from the POD input the host hands over)."""
    import ctypes as C
    from distraytracer_b200 import abi
    return len(scene.prims) * C.sizeof(abi.Prim)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-rows", type=int, default=1, help="rows per host core in one CPU-reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
