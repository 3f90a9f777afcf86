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
MOCAP_START = 880

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
    ("chkpt2_mocap", "chkpt2", 910, 4, {}),                        # 29 bone cylinders + 2 sphere lights; clip frame 910 = frame 30 of the stored window
    ("boundary_mocap", "boundary", 1, 4, {"aperture": 0.0}),       # rect lights + sphere light + glossy
    ("prismcyl", "prismcyl", 7, 1, {"aperture": 0.0}),             # `./render prismcyl 7` (scene.h:3227-3263): RectPrismWithCylinder
]


# ---- the slab-box prism classes -----------------------------------------------------------------------
# Only RectPrismWithCylinder is instantiated by a reference scene builder, and that scene renders black (the hole's cap
# planes coincide with the box faces and intersectCap has no radius test, so every ray through the front face is "in the
# hole").  These scenes put the three classes in front of the reference's own intersect / intersectShadow / getNorm through
# its constructors (drtref_load_scene): a floor and a sphere to catch and cast shadows, two point lights.
def prism(kind, lo, hi, color, holes=()):
    """Axis-aligned prism [lo, hi] with the corner naming of scene.h:3232-3238: A..D the x = lo face, E..H = A..D + back."""
    from distraytracer_b200 import abi, scenes
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    p = scenes.new_prim()
    p.type = kind
    A = np.array([lo[0], lo[1], lo[2]]); B = np.array([lo[0], lo[1], hi[2]]); C = np.array([lo[0], hi[1], hi[2]]); D = np.array([lo[0], hi[1], lo[2]])
    back = np.array([hi[0] - lo[0], 0, 0])
    for dst, v in zip((p.A, p.B, p.C, p.D, p.E, p.F, p.G, p.H), (A, B, C, D, A + back, B + back, C + back, D + back)):
        dst[:] = list(v)
    p.color[:] = list(color)
    p.center[:] = list((lo + hi) / 2)
    p.n_holes = len(holes)
    for k, h in enumerate(holes):
        p.holes[k].type = h["type"]; p.holes[k].c1[:] = h["c1"]; p.holes[k].c2[:] = h.get("c2", [0, 0, 0])
        p.holes[k].radius = float(np.float32(h["radius"])); p.holes[k].color[:] = h["color"]
    return p


def prism_scene(kind):
    from distraytracer_b200 import abi, scenes
    from distraytracer_b200.scene import Scene
    floor = scenes.rectangle((-4, -2.2, -6), (8, -2.2, -6), (8, -2.2, 6), (-4, -2.2, 6), (0.7, 0.7, 0.7), name=abi.NAME_OTHER)
    ball = scenes.sphere((3.0, 0.0, 2.6), 0.8, (0.2, 0.9, 0.3))
    if kind == abi.PRIM_RECTPRISM:
        pr = prism(kind, (0, -2, -2), (1, 1.5, 1.2), (1, 0.2, 0.1))
    elif kind == abi.PRIM_RECTPRISM_CYL:
        # one cylinder whose caps stick out of the box on both sides, one that ends inside it
        pr = prism(kind, (0, -2, -2), (1, 2, 2), (1, 0, 0),
                   holes=[dict(type=abi.PRIM_CYLINDER, c1=[-0.5, 0.3, -0.4], c2=[1.5, 0.3, -0.4], radius=0.9, color=[0, 0, 1]),
                          dict(type=abi.PRIM_CYLINDER, c1=[0.2, -1.2, 1.0], c2=[0.7, -1.0, 1.3], radius=0.4, color=[1, 1, 0])])
    else:
        pr = prism(kind, (0, -2, -2), (1, 2, 2), (1, 0, 0),
                   holes=[dict(type=abi.PRIM_SPHERE, c1=[0.5, 0.6, -0.5], radius=0.8, color=[0, 0, 1]),
                          dict(type=abi.PRIM_CYLINDER, c1=[-0.3, -1.0, 1.0], c2=[1.4, -0.9, 1.1], radius=0.5, color=[1, 1, 0])])
    return Scene([pr, floor, ball], [scenes.point_light((-5, 3, 1), (1, 1, 1)), scenes.point_light((6, 4, -3), (0.6, 0.6, 0.9))], [])


# (case, class, eye): RectPrismWithHoles::getNorm throws on three of the six faces, so its views abandon thousands of pixels
PRISM_CASES = [("prism_box", "PRIM_RECTPRISM", (-3, 2.5, 5)), ("prism_cyl", "PRIM_RECTPRISM_CYL", (-6, 0.5, 1)),
               ("prism_cyl_side", "PRIM_RECTPRISM_CYL", (-3, 2.5, 5)), ("prism_holes", "PRIM_RECTPRISM_HOLES", (-3, 2.5, 5)),
               ("prism_holes_back", "PRIM_RECTPRISM_HOLES", (5, 1.0, 4))]


def main():
    r = Ref(mocap=True)
    if os.environ.get("GOLDEN_ONLY_MOCAP"):
        return mocap_and_cloud(r)
    for case, builder, frame, aa, kw in CASES:
        if os.environ.get("GOLDEN_ONLY_PRISMS") and case != "prismcyl":
            continue
        if os.environ.get("GOLDEN_ONLY") and case not in os.environ["GOLDEN_ONLY"].split(","):
            continue
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

    if os.environ.get("GOLDEN_ONLY"):
        return
    from distraytracer_b200 import abi
    for case, kind, eye in PRISM_CASES:
        scene = prism_scene(getattr(abi, kind))
        r.reset()
        s = r.settings()
        s.xRes, s.yRes, s.antialias_samples, s.frame, s.seed, s.aperture = W, H, 1, 0, SEED, 0.0
        s.eye[:] = list(eye); s.lookingAt[:] = [0.5, 0, 0]
        r.load(scene); r.set_settings(s)
        r.rng(1, SEED, 0)
        img, aborted, sec = r.render_loop(0, reset_policy=1, seed=SEED)
        save_fixture(os.path.join(OUT, case + ".npz"), scene, s, ref_f32=img, ref_aborted=aborted.astype(np.uint8))
        print(f"{case:24s} prims={len(scene.prims):3d} lights={len(scene.lights)} aborted={int(aborted.sum())} mean={img.mean():.2f} ref {sec:.2f}s")
    if os.environ.get("GOLDEN_ONLY_PRISMS"):
        return
    mocap_and_cloud(r)


def mocap_and_cloud(r):

    # mocap bone end points (BASELINE config 4), scene.h:637-659.  The clip holds its first pose for 560 frames; the window
    # MOCAP_START .. MOCAP_START + 120 is the middle of the acrobatic sequence, where every bone moves.
    bones = np.stack([r.mocap_bones(MOCAP_START + f) for f in range(120)])
    np.save(os.path.join(OUT, "mocap_bones_880_999.npy"), bones.astype(np.float64))
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
    # the same 121 frames, renumbered from 1 (3 header lines, then a frame-number line + 29 bone lines per frame)
    window = amc_lines[3 + MOCAP_START * 30: 3 + (MOCAP_START + 121) * 30]
    for k in range(121):
        assert int(window[k * 30]) == MOCAP_START + k + 1
        window[k * 30] = str(k + 1).encode()
    with open(os.path.join(OUT, "mocap_90_16_frames880_1000.amc"), "wb") as f:
        f.write(b"\n".join(amc_lines[:3] + window) + b"\n")

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
