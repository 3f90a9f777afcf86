"""Generate the golden fixtures in tests/golden/ from the reference itself.

Run in the build container (needs /root/reference and oracle/_ref/libdrt_ref.so,
built by `make -C oracle ref`):

    python tests/golden/make_golden.py

For every scene the reference can build from in-repo assets this
stores, in <case>.npz:
  * the scene exactly as the reference's builder produced it, flattened to the
    drt.h PODs by oracle/ref_driver.cpp (prims, lights, textures) + the settings;
  * `ref_f32`: the image the UNMODIFIED reference rayColor produced (float, before
    writePPM's truncation) under the deterministic sequential sample stream
    (oracle/drt_rng.h, reset per pixel, seed in settings.seed);
  * `ref_aborted`: pixels where the reference itself terminates (bare `throw;`).
The reference publishes no golden vectors of its own; these are
"outputs of the reference itself run here".
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.harness import Ref  # noqa: E402
from distraytracer_b200.scene import save_fixture  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
W, H, SEED = 160, 120, 7

# (case name, reference builder, frame, antialias_samples, overrides)
CASES = [
    ("hw4", "hw4", 0, 1, {}),
    ("reflectance", "reflectance", 40, 4, {}),                    # DOF (aperture 0.2) + glossy aluminium
    ("dof", "dof", 0, 4, {}),
    ("spherelight", "spherelight", 0, 4, {}),                      # sphere area light
    ("spheres_blur", "spheres", 10, 1, {}),                        # motion flag -> blur re-traces
    ("checkertexture", "checkertexture", 0, 1, {"aperture": 0.0}),  # BASELINE config 1
    ("checkertexture_nogloss", "checkertexture", 0, 1, {"aperture": 0.0, "nogloss": 1}),
    ("texture", "texture", 0, 1, {"aperture": 0.0}),
    ("textureog", "textureog", 0, 1, {"aperture": 0.0}),
    ("window", "window", 0, 1, {"aperture": 0.0}),
    ("staircase", "staircase", 5, 1, {"aperture": 0.0}),
    ("rectprism", "rectprism", 13, 1, {"aperture": 0.0}),
    ("checkercylinder", "checkercylinder", 0, 4, {"aperture": 0.0}),
    ("chkpt2_mocap", "chkpt2", 30, 4, {}),                         # 29 bone cylinders + 2 sphere lights
    ("boundary_mocap", "boundary", 1, 4, {"aperture": 0.0}),       # rect lights + sphere light + glossy
]


def main():
    r = Ref(mocap=True)
    for case, builder, frame, aa, kw in CASES:
        r.reset()
        r.build(builder, frame)
        s = r.settings()
        s.xRes, s.yRes, s.antialias_samples, s.frame, s.seed = W, H, aa, frame, SEED
        for k, v in kw.items():
            setattr(s, k, v)
        r.set_settings(s)
        scene = r.export()
        r.rng(1, SEED, 0)
        img, aborted, sec = r.render_loop(frame, reset_policy=1, seed=SEED)
        save_fixture(os.path.join(OUT, case + ".npz"), scene, s,
                     ref_f32=img, ref_aborted=aborted.astype(np.uint8))
        print(f"{case:24s} prims={len(scene.prims):3d} lights={len(scene.lights)} tex={len(scene.textures)} "
              f"aborted={int(aborted.sum())} ref {sec:.2f}s")

    # mocap bone end points for frames 0..119 (BASELINE config 4), scene.h:637-659
    bones = np.stack([r.mocap_bones(f) for f in range(120)])
    np.save(os.path.join(OUT, "mocap_bones_0_119.npy"), bones.astype(np.float64))
    print("mocap bones", bones.shape)
    # the clip itself for the ASF/AMC ingest tests: the skeleton file and the first 121 frames of the motion
    # (3 header lines + 121 x 30 lines), copied as DATA fixtures
    from oracle.harness import REFERENCE_ROOT
    with open(os.path.join(REFERENCE_ROOT, "90.asf"), "rb") as f:
        asf = f.read()
    with open(os.path.join(REFERENCE_ROOT, "90_16_v3.amc"), "rb") as f:
        amc_lines = f.read().split(b"\n")
    with open(os.path.join(OUT, "mocap_90.asf"), "wb") as f:
        f.write(asf)
    with open(os.path.join(OUT, "mocap_90_16_first121.amc"), "wb") as f:
        f.write(b"\n".join(amc_lines[:3 + 121 * 30]) + b"\n")

    # value-noise known answers (noise.h) through the reference's renderImageCloud
    r.reset()
    s = r.settings()
    s.xRes, s.yRes = 64, 48
    r.set_settings(s)
    cloud = r.render_cloud(3.0)
    np.save(os.path.join(OUT, "cloud_frame3_64x48.npy"), cloud)
    print("cloud", cloud.shape, cloud.mean())


if __name__ == "__main__":
    main()
