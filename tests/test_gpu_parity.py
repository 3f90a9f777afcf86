"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on
identical sample positions (keyed sample stream), plus the committed golden data.

Bar (BASELINE.json north_star): per-pixel error <= 1/255 on >= 99.9 % of pixels in
deterministic fixed-sample mode.  Pixels where the reference itself aborts are
written as 0 by both sides and take part in the comparison.
"""
import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_CASES, load_case

pytestmark = pytest.mark.gpu

TOL_FRAC = 0.999   # fraction of pixels within 1/255


def _gpu(scene):
    from distraytracer_b200 import runtime
    assert runtime.device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return runtime.DeviceScene(scene, 0)


@pytest.mark.parametrize("precision", [0, 1], ids=["reference", "fp32"])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_cuda_matches_oracle_same_samples(oracle_lib, case, precision):
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    scene, settings, _ = load_case(case)
    settings.precision = precision
    want, want_ab, _, _ = Oracle(scene).render(settings, mode=ORACLE_KEYED)
    got, got_u8 = _gpu(scene).render_float(settings)
    st = compare(want, got)
    # reference precision: the bar.  fp32: same bar, except on the one fixture whose image is
    # decided by rounding residues (area-light panels lying IN the ceiling plane: the shadow
    # rays' plane tests divide ~1e-17 by ~1e-17), where only double arithmetic can follow.
    bar = TOL_FRAC if (precision == 0 or case != "boundary_mocap") else 0.75
    assert st["frac_within_1"] >= bar, (case, st)
    from oracle.harness import quantize
    assert np.array_equal(got_u8, quantize(got)), "u8 output is not writePPM's truncation of the float image"


def test_cuda_cloud_frame_matches_reference_golden(oracle_lib):
    """renderImageCloud (noise.h integer hash must be exact) against the reference's own output."""
    from distraytracer_b200 import abi
    gold = np.load(GOLDEN + "/cloud_frame3_64x48.npy")
    scene, settings, _ = load_case("hw4")
    s = abi.copy_struct(settings)
    s.xRes, s.yRes, s.cloud_only, s.frame = 64, 48, 1, 3
    s.eye[:] = [0.5, 1.5, 1]; s.up[:] = [0, 0, 1]; s.lookingAt[:] = [0.5, -1, 1]
    for precision in (0, 1):
        s.precision = precision
        got = _gpu(scene).render(s)
        d = np.abs(got.astype(int) - gold.astype(int)).max(axis=-1)
        assert (d <= 1).mean() >= TOL_FRAC, (precision, int(d.max()), float((d <= 1).mean()))


def test_cuda_perlin_background_and_tiles(oracle_lib):
    """perlin_cloud background behind geometry + tiles: four quarter tiles == full frame."""
    from distraytracer_b200 import abi
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    scene, settings, _ = load_case("reflectance")
    s = abi.copy_struct(settings)
    s.xRes, s.yRes, s.perlin_cloud = 96, 64, 1
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    dev = _gpu(scene)
    got, full = dev.render_float(s)
    assert compare(want, got)["frac_within_1"] >= TOL_FRAC
    tiles = np.zeros_like(full)
    for (x0, y0) in [(0, 0), (48, 0), (0, 32), (48, 32)]:
        t = dev.render(s, abi.Tile(x0, y0, 48, 32, 0))
        r0 = s.yRes - (y0 + 32)
        tiles[r0:r0 + 32, x0:x0 + 48] = t
    assert np.array_equal(tiles, full)


def test_cuda_full_size_properties(oracle_lib):
    """BASELINE config sizes through size-independent properties: determinism (same seed ->
    identical bytes), seed sensitivity, and a 16-row band of the 1080p/64spp frame equal to the
    same rows of the oracle."""
    from distraytracer_b200 import abi
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    scene, settings, _ = load_case("checkertexture")
    s = abi.copy_struct(settings)
    s.xRes, s.yRes, s.antialias_samples, s.aperture, s.focal_length = 1920, 1080, 64, 0.2, 10.0
    dev = _gpu(scene)
    band = abi.Tile(0, 500, 1920, 16, 0)
    a = dev.render(s, band)
    b = dev.render(s, band)
    assert np.array_equal(a, b)
    s2 = abi.copy_struct(s); s2.seed = s.seed + 1
    assert not np.array_equal(a, dev.render(s2, band))
    sub = abi.Tile(800, 500, 160, 16, 0)
    want, _, _, _ = Oracle(scene).render(s, sub, mode=ORACLE_KEYED)
    got, _ = dev.render_float(s, sub)
    assert compare(want, got)["frac_within_1"] >= TOL_FRAC


def test_scene_update_and_errors(oracle_lib):
    from distraytracer_b200 import abi, runtime
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    from distraytracer_b200.scene import Scene
    scene, settings, _ = load_case("chkpt2_mocap")
    bones = np.load(GOLDEN + "/mocap_bones_880_999.npy")
    dev = _gpu(scene)
    prims = [abi.copy_struct(p) for p in scene.prims]
    k = 0
    for p in prims:                                   # re-pose the bone cylinders to mocap frame 60
        if p.type == abi.PRIM_CYLINDER:
            p.c1[:] = bones[60, k, 0]; p.c2[:] = bones[60, k, 1]
            p.center[:] = (bones[60, k, 0] + bones[60, k, 1]) / 2
            k += 1
    dev.update_prims(prims)
    got, _ = dev.render_float(settings)
    want, _, _, _ = Oracle(Scene(prims, scene.lights, scene.textures)).render(settings, mode=ORACLE_KEYED)
    assert compare(want, got)["frac_within_1"] >= TOL_FRAC
    with pytest.raises(runtime.DrtError):             # count must not change
        dev.update_prims(prims[:-1])
    bad = abi.copy_struct(settings)
    bad.eye[:] = [0, 0, 0]; bad.lookingAt[:] = [0, 1, 0]; bad.up[:] = [0, 1, 0]   # up x gaze == 0 exactly
    with pytest.raises(runtime.DrtError) as e:        # gaze == up, render_final_project.cpp:992-996
        dev.render(bad)
    assert e.value.code == abi.ERR_SCENE


@pytest.mark.parametrize("cfg", ["config1", "config2", "config3", "config4", "config4_two_pose", "config4_long_shutter"])
def test_cuda_matches_oracle_on_baseline_configs(oracle_lib, cfg):
    """The BASELINE.json workloads at reduced resolution / spp (the oracle finishes in seconds):
    C2 exercises DOF + glossy floor + rectangle-light soft shadows + glass refraction (NaN
    Fresnel terms included, written as 0 like the reference's writePPM does), C3 the cloud
    background behind Oren-Nayar spheres, C4 velocity motion blur of the mocap bones."""
    from distraytracer_b200 import scenes
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    if cfg == "config1":
        scene, s = scenes.config1(); s.xRes, s.yRes = 320, 240
    elif cfg == "config2":
        scene, s = scenes.config2(240, 135, 16)
    elif cfg == "config3":
        scene, s = scenes.config3(96, 54, 4)
    elif cfg == "config4":
        scene, s = scenes.config4_frame(37, 160, 90, 4)
    elif cfg == "config4_two_pose":
        # every bone between its poses at frame and frame + 1: end points on their own paths, axis re-derived per time
        # sample (DRT_FLAG_VERTEX_MOTION); the slab filter keeps culling through the swept boxes
        scene, s = scenes.config4_frame(37, 160, 90, 4, two_pose=True)
        s.blur_samples = 3
    else:
        # a shutter longer than the swept boxes cover (frame_range 3 > 1): re-traces fall back to testing every geom
        scene, s = scenes.config4_frame(37, 160, 90, 4, two_pose=True)
        s.frame_range = 3
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (cfg, st)


def test_cuda_mesh_lbvh_matches_oracle(oracle_lib):
    """Triangle mesh through the device-built LBVH (BASELINE config 5 at reduced size: 2 x 23^2 =
    1058 textured Oren-Nayar triangles + a sphere, DOF) against the oracle fed the same triangles
    as individual Triangle primitives."""
    from distraytracer_b200 import scenes
    from distraytracer_b200.scene import Scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    scene, s = scenes.config5(n=24, xres=160, yres=90, spp=4)
    flat = Scene(list(scene.prims) + scenes.mesh_to_prims(scene.mesh), scene.lights, scene.textures)
    want, _, _, _ = Oracle(flat).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, st


def test_cuda_mesh_per_triangle_materials_match_oracle(oracle_lib):
    """The reference's model builders give every Triangle its own roughness from a roughness map (scene.h:372-378).  The
    mesh carries that as a material table + one index per triangle; the picture must be the oracle's, which gets one
    Triangle primitive per face with its own roughness."""
    from distraytracer_b200 import scenes, abi
    from distraytracer_b200.scene import Scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    scene, s = scenes.config5(n=24, xres=160, yres=90, spp=4)
    rng = np.random.default_rng(11)
    nt = len(scene.mesh["indices"])
    fr = (rng.integers(0, 766, size=nt) / np.float32(3 * 255)).astype(np.float32)     # what face_roughness_from_map yields
    values, ids = np.unique(fr, return_inverse=True)
    mats = []
    for v in values:
        m = abi.copy_struct(scene.mesh["material"]); m.roughness = float(v); mats.append(m)
    mesh = dict(scene.mesh, materials=mats, material_ids=ids.astype(np.int32))
    tabled = Scene(scene.prims, scene.lights, scene.textures, mesh=mesh)
    flat = Scene(list(scene.prims) + scenes.mesh_to_prims(mesh), scene.lights, scene.textures)
    want, _, _, _ = Oracle(flat).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(tabled).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, st
    # ... and it is not the picture of the single-material mesh
    one, _ = _gpu(scene).render_float(s)
    assert compare(one, got)["mae"] > 0.05
    # the oracle's own mesh expansion reads the same table
    via_mesh, _, _, _ = Oracle(tabled).render(s, mode=ORACLE_KEYED)
    assert np.array_equal(want, via_mesh)


def test_cuda_obj_ingest_renders_like_the_direct_mesh(oracle_lib):
    """A mesh that went through the OBJ text format and `ingest.mesh_from_obj` (vertex unification, no
    UV rewrite) renders byte-identically to the mesh it was written from."""
    from distraytracer_b200 import scenes, ingest
    from distraytracer_b200.scene import Scene
    scene, s = scenes.config5(n=16, xres=96, yres=54, spp=4)
    obj = ingest.parse_obj(ingest.mesh_to_obj(scene.mesh))
    mesh2 = ingest.mesh_from_obj(obj, scene.mesh["material"], wrap_uv=False, flip_v=False)
    a = _gpu(scene).render(s)
    b = _gpu(Scene(scene.prims, scene.lights, scene.textures, mesh=mesh2)).render(s)
    assert np.array_equal(a, b) and a.std() > 5


def test_cuda_mesh_full_size_matches_oracle_on_crops(oracle_lib):
    """BASELINE config 5 at FULL size -- 999 698 triangles through the device-built LBVH, 3840x2160, DOF, the motion-blurred
    sphere -- against the oracle on crops of the 4K frame at 4 spp.  The reference's generateBVH is O(n^2) per node and
    cannot build this scene (helpers.h:452-453), so the oracle walks its median-split stand-in tree (oracle/drt_oracle.cpp
    FastBuilder: same gather semantics, validated against the reference-order tree in test_oracle_golden.py)."""
    from distraytracer_b200 import scenes, abi
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    scene, s = scenes.config5()
    s.antialias_samples = 4
    oracle = Oracle(scene, builder=1)
    dev = _gpu(scene)
    crops = [abi.Tile(1600, 1060, 640, 20, 0),     # through the sphere: velocity-blur re-traces, its shadow on the terrain
             abi.Tile(1904, 700, 32, 700, 0),      # a column from the near terrain up over the sphere to the far edge
             abi.Tile(40, 300, 96, 54, 0),         # the left edge of the terrain against the empty background
             abi.Tile(2700, 1200, 96, 54, 0), abi.Tile(1000, 600, 96, 54, 0), abi.Tile(2900, 500, 96, 54, 0)]
    total = bad = 0
    for tile in crops:
        want, _, _, _ = oracle.render(s, tile, mode=ORACLE_KEYED)
        got, _ = dev.render_float(s, tile)
        st = compare(want, got)
        assert st["frac_within_1"] >= TOL_FRAC, ((tile.x0, tile.y0), st)
        total += tile.width * tile.height; bad += st["n_bad"]
        assert want.std() > 0.5, "crop shows nothing"
    assert total >= 50000 and bad <= total * (1 - TOL_FRAC)


def test_cuda_mesh_full_size_builds_and_is_deterministic(oracle_lib):
    """999 698 triangles: the LBVH builds on the device, a 4K/64spp band renders, twice the same."""
    from distraytracer_b200 import scenes, abi
    scene, s = scenes.config5()
    dev = _gpu(scene)
    band = abi.Tile(0, 1000, 3840, 8, 0)
    a = dev.render(s, band)
    b = dev.render(s, band)
    assert np.array_equal(a, b) and a.std() > 5


def test_production_spp_mae_against_high_spp_reference(oracle_lib):
    """North-star statistical bar: at production spp (64) the image must be within a mean absolute
    error of 0.5/255 of a high-spp reference render.  The reference render is the oracle in
    DRT_ORACLE_STREAM mode -- the sequential sample stream in the reference's own draw order,
    bit-identical to the compiled reference (tests/test_oracle_golden.py) -- i.e. a different sample
    set drawn a different way (per-pixel lens shuffle included), so this checks that the keyed
    sampler of the CUDA path is unbiased, not just that it repeats the oracle.
    "High spp" is 32 independent 64-spp frames averaged (2048 spp): the reference's own jitter is
    `(i + u)/9` with a hard-coded 9 and int truncation (render_final_project.cpp:1052, quirk Q1), so
    a single frame with n = sqrt(spp) > 9 samples neighbouring PIXELS and is a different image."""
    from distraytracer_b200 import abi
    from oracle.harness import Oracle, ORACLE_STREAM
    # (case, overrides, single 64-spp frame must meet the bar).  On the area-light scenes the
    # reference's OWN 64-spp frame is 1.2-1.3/255 away from its converged mean (Monte Carlo noise,
    # measured with the oracle), so there the bar is asserted on the mean of 16 frames (noise / 4)
    # and the single frame only has to be as close as the reference's own single frame is.
    for case, extra, single in (("reflectance", {}, True), ("boundary_mocap", {"aperture": 0.2}, False),
                                ("spherelight", {"aperture": 0.2}, False)):
        scene, settings, _ = load_case(case)
        s = abi.copy_struct(settings)
        s.xRes, s.yRes, s.antialias_samples = 64, 36, 64
        for k, v in extra.items():
            setattr(s, k, v)
        o = Oracle(scene)
        dev = _gpu(scene)
        ref_frames, our_frames = [], []
        for k in range(16):
            q = abi.copy_struct(s); q.seed = 100 + k
            ref_frames.append(np.clip(o.render(q, mode=ORACLE_STREAM)[0], 0, 255))
            q.seed = 5000 + k
            our_frames.append(np.clip(dev.render_float(q)[0], 0, 255))
        ref = np.mean(ref_frames, axis=0)
        mae_mean = float(np.abs(np.mean(our_frames, axis=0) - ref).mean())
        assert mae_mean <= 0.5, (case, "mean of 16 frames", mae_mean)
        q = abi.copy_struct(s); q.seed = 777
        ref_self = float(np.abs(np.clip(o.render(q, mode=ORACLE_STREAM)[0], 0, 255) - ref).mean())
        ours_single = float(np.abs(our_frames[0] - ref).mean())
        if single:
            assert ours_single <= 0.5, (case, ours_single)
        assert ours_single <= 1.2 * ref_self + 0.05, (case, ours_single, ref_self)


def test_cuda_many_shapes_match_oracle(oracle_lib):
    """337 analytic shapes (more geoms than the shared-memory slab table holds): the CUDA path gathers candidates through
    its own tree over the geoms, the oracle through the reference-order BVH -- same image, ties of t included (the steel
    floor and the spheres' mirror reflections see every shape)."""
    from distraytracer_b200 import scenes
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    scene, s = scenes.many_shapes()
    assert len(scene.prims) > 300
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    dev = _gpu(scene)
    got, _ = dev.render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, st
    assert float(np.abs(want).max()) > 0.2                    # not a black frame


@pytest.mark.parametrize("variant", ["one_pixel", "edge_tile", "depth1", "noreflect", "nogloss", "no_lights", "aa2", "blur_ref_mode",
                                     "brdf5_depth4", "depth32", "dof_aa10", "dof_aa24", "dof_scan_fallback"])
def test_cuda_edge_cases_match_oracle(oracle_lib, variant):
    """Edge cases of the settings surface: degenerate tiles, depth / switch extremes, spp that is not
    a square, a light-less scene, and the reference's own motion-blur mode with moving "rectangle"
    shapes after frame_prism (bumpBVH path, render_final_project.cpp:1095-1210)."""
    from distraytracer_b200 import abi
    from distraytracer_b200.scene import Scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    scene, settings, _ = load_case("boundary_mocap" if variant == "blur_ref_mode" else "checkertexture")
    s = abi.copy_struct(settings)
    s.xRes, s.yRes = 96, 72
    tile = None
    if variant == "one_pixel":
        tile = abi.Tile(47, 30, 1, 1, 0)
    elif variant == "edge_tile":
        tile = abi.Tile(90, 66, 6, 6, 0)          # top-right corner of the frame in loop coordinates
    elif variant == "depth1":
        s.max_depth = 1
    elif variant == "noreflect":
        s.reflect = 0
    elif variant == "nogloss":
        s.nogloss = 1
    elif variant == "no_lights":
        scene = Scene(scene.prims, [], scene.textures)
    elif variant == "aa2":
        s.antialias_samples, s.aperture = 2, 0.2   # n = int(sqrt(2)) = 1 -> 1 spp
    elif variant == "dof_aa10":
        # 10 lens points shuffled, 9 used (sampled_n = int(sqrt(10))^2): the shuffle decides WHICH nine (helpers.h:270-279)
        s.antialias_samples, s.aperture = 10, 0.3
    elif variant == "dof_aa24":
        # 16 spp of 24 lens points: batches of 1024 samples are not pixel aligned (1024 % 16 == 0 but 24-entry permutations)
        s.antialias_samples, s.aperture = 24, 0.3
    elif variant == "dof_scan_fallback":
        # 4400 lens points per pixel do not fit the shared-memory permutation table: per-sample scan (lensIndexScan)
        s.antialias_samples, s.aperture, s.xRes, s.yRes, s.nogloss, s.max_depth = 4400, 0.3, 8, 6, 1, 2
    elif variant == "brdf5_depth4":
        s.brdf_samples, s.max_depth = 5, 4         # widest glossy fan the ray pool is sized for
    elif variant == "depth32":
        s.max_depth = 32                           # deepest recursion the boundary accepts
    elif variant == "blur_ref_mode":
        # frame >= frame_prism: shapes named "rectangle" move in y for the blur re-traces; flag them
        prims = [abi.copy_struct(p) for p in scene.prims]
        for p in prims:
            if p.name == abi.NAME_RECTANGLE:
                p.flags |= abi.FLAG_MOTION
        scene = Scene(prims, scene.lights, scene.textures)
        s.frame, s.frame_prism, s.frame_blur, s.frame_range, s.blur_samples = 1700, 960, 1600, 8, 2
        s.antialias_samples = 4
    want, _, _, _ = Oracle(scene).render(s, tile, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s, tile)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (variant, st)


def test_video_frames_through_scene_update(oracle_lib):
    """BASELINE config 4 mechanics: one resident scene, per frame only the bone cylinders are re-posed
    (drt_scene_update_prims) -- must equal building the scene from scratch for that frame."""
    from distraytracer_b200 import scenes
    scene0, s0 = scenes.config4_frame(10, 160, 90, 4)
    dev = _gpu(scene0)
    for f in (11, 57, 119):
        scene_f, s_f = scenes.config4_frame(f, 160, 90, 4)
        dev.update_prims(scene_f.prims)
        a = dev.render(s_f)
        b = _gpu(scene_f).render(s_f)
        assert np.array_equal(a, b), f


def test_update_lights_equals_a_fresh_scene(oracle_lib):
    """drt_scene_update_lights: moving the light of a resident scene gives the frame a newly created scene gives."""
    from distraytracer_b200 import abi
    from distraytracer_b200.scene import Scene
    scene, settings, _ = load_case("reflectance")
    s = abi.copy_struct(settings); s.xRes, s.yRes = 96, 72
    dev = _gpu(scene)
    a0 = dev.render(s)
    lights = [abi.copy_struct(l) for l in scene.lights]
    lights[0].center[2] += 7.0
    dev.update_lights(lights)
    a1 = dev.render(s)
    b1 = _gpu(Scene(scene.prims, lights, scene.textures)).render(s)
    assert np.array_equal(a1, b1) and not np.array_equal(a0, a1)


@pytest.mark.gpu
def test_out_of_scope_inputs_are_rejected_not_rendered(oracle_lib):
    """What the hot path does not cover fails loudly with DRT_ERR_UNSUPPORTED (never a silent approximation): primitive
    classes outside drt_prim_type, holes of a class the reference has no cap / far-root test for, a textured
    sphere (GeoPrimitive::getUV has no body, geometry.h:36), recursion deeper than 32, more glossy lobes than the ray
    pool is sized for."""
    from distraytracer_b200 import abi, runtime
    from distraytracer_b200.scene import Scene
    scene, settings, _ = load_case("checkertexture")
    prims = [abi.copy_struct(p) for p in scene.prims]
    prims[0].type = abi.PRIM_TYPE_COUNT                     # no such class
    with pytest.raises(runtime.DrtError) as e:
        runtime.DeviceScene(Scene(prims, scene.lights, scene.textures), 0)
    assert e.value.code == abi.ERR_UNSUPPORTED
    prism, _, _ = load_case("prism_cyl")
    prims = [abi.copy_struct(p) for p in prism.prims]
    prims[0].holes[0].type = abi.PRIM_SPHERE                # RectPrismWithCylinder::holes are Cylinders (intersectCap)
    with pytest.raises(runtime.DrtError) as e:
        runtime.DeviceScene(Scene(prims, prism.lights, prism.textures), 0)
    assert e.value.code == abi.ERR_UNSUPPORTED
    prims = [abi.copy_struct(p) for p in scene.prims]
    ball = next(p for p in prims if p.type == abi.PRIM_SPHERE)
    ball.flags |= abi.FLAG_TEXTURE; ball.tex_frame = 0
    with pytest.raises(runtime.DrtError) as e:
        runtime.DeviceScene(Scene(prims, scene.lights, scene.textures), 0)
    assert e.value.code == abi.ERR_UNSUPPORTED
    dev = _gpu(scene)
    for field, value in (("max_depth", 33), ("brdf_samples", 7), ("blur_samples", 65)):
        s = abi.copy_struct(settings)
        s.xRes, s.yRes = 32, 24
        setattr(s, field, value)
        with pytest.raises(runtime.DrtError) as e:
            dev.render(s)
        assert e.value.code == abi.ERR_UNSUPPORTED, field


@pytest.mark.gpu
def test_row_chunking_does_not_change_the_image(oracle_lib):
    """A frame larger than the per-launch sample bound is rendered in row chunks (one render_wave launch each); the
    picture is the same however it is chunked (DRT_CHUNK_LOG2 shrinks the bound so a small frame needs many launches)."""
    import os
    from distraytracer_b200 import abi
    scene, settings, _ = load_case("checkertexture")
    s = abi.copy_struct(settings)
    s.xRes, s.yRes, s.antialias_samples, s.aperture = 96, 72, 16, 0.2
    dev = _gpu(scene)
    one, many = abi.Counters(), abi.Counters()
    a = dev.render(s, counters=one)
    os.environ["DRT_CHUNK_LOG2"] = "12"            # 4096 samples = 2 rows of this frame per launch
    try:
        b = dev.render(s, counters=many)
    finally:
        del os.environ["DRT_CHUNK_LOG2"]
    assert np.array_equal(a, b) and a.std() > 5
    assert one.kernel_launches == 2 and many.kernel_launches == 2 * 36


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(14)) + [43, 151])   # 43, 151: depth of field past a wall whose corners are listed diagonally (GF_SPILL)
def test_cuda_matches_oracle_on_mutated_scenes(oracle_lib, seed):
    """Fuzz: fixture scenes with random motion flags, BRDF models, roughness, reflective materials, glossy flags and
    settings (tests/fuzz_cases.py); the oracle is pinned on the compiled reference for the same kind of mutations by
    tests/test_oracle_fuzz.py.  On B200 all 14 are bit-identical in u8; the bar asserted is the north-star one."""
    from fuzz_cases import mutated_case
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    case, scene, s = mutated_case(seed)
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (case, st)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(16))
def test_cuda_matches_oracle_on_random_scenes(oracle_lib, seed):
    """Scenes no reference builder produced (fuzz_cases.random_scene; the oracle is pinned on the compiled reference for
    the same scenes in tests/test_oracle_fuzz.py): random shapes of every analytic class, rectangles with their corners in
    any order, all three light kinds, random cameras.  tools/gpu_fuzz.py N scenes runs more seeds (600 were bit-identical)."""
    from fuzz_cases import random_scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    _, scene, s = random_scene(seed)
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (seed, st)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_cuda_matches_oracle_on_random_meshes(oracle_lib, seed):
    """Random triangle meshes (bumpy grids with holes, triangle soups; per-triangle material tables; textured or not)
    through the device-built 4-wide LBVH vs the oracle fed the same triangles as Triangle primitives through its
    reference-order walk (fuzz_cases.random_mesh_scene; tools/gpu_fuzz.py N meshes ran 400 seeds bit-identical)."""
    from fuzz_cases import random_mesh_scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    _, scene, flat, s = random_mesh_scene(seed)
    want, _, _, _ = Oracle(flat).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (seed, st)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2, 3, 31, 72, 93, 249])
def test_cuda_matches_oracle_on_random_scenes_in_velocity_blur(oracle_lib, seed):
    """DRT_BLUR_VELOCITY (the library's own blur mode) on random scenes: random velocities, two-pose
    cylinders, rectangles of the diagonal kind (Q20: in a velocity re-trace every primitive is a candidate, there is no
    reference gather to replay -- seeds 31 ... 249 are the ones on which the oracle twin once culled such a rectangle)."""
    from fuzz_cases import random_motion_scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    _, scene, s = random_motion_scene(seed)
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (seed, st)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,seed", [("big", 0), ("big", 1), ("big", 2), ("big", 3), ("sky", 0), ("sky", 1), ("sky", 2), ("sky", 3)])
def test_cuda_matches_oracle_on_big_scenes_and_cloud_backgrounds(oracle_lib, kind, seed):
    """fuzz_cases.random_big_scene (257..700 shapes: the tree over the geoms) and fuzz_cases.sky_case (mutated fixtures in
    front of the value-noise clouds: cloud_corners + per-sample corner offsets); tools/gpu_fuzz.py ran 300 / 200 seeds."""
    from fuzz_cases import random_big_scene, sky_case
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    _, scene, s = (random_big_scene if kind == "big" else sky_case)(seed)
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (kind, seed, st)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_cuda_matches_oracle_on_moving_meshes(oracle_lib, seed):
    """A mesh in motion (DRT_BLUR_VELOCITY: the whole mesh by its material's velocity -- the traversal moves the ray through
    the static tree, the exact test runs on the displaced vertices) against the oracle twin, which moves every triangle
    (fuzz_cases.random_moving_mesh_scene; tools/gpu_fuzz.py ran 300 seeds bit-identical)."""
    from fuzz_cases import random_moving_mesh_scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    _, scene, flat, s = random_moving_mesh_scene(seed)
    want, _, _, _ = Oracle(flat).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (seed, st)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_cuda_matches_oracle_on_random_scenes_with_the_reference_blur(oracle_lib, seed):
    """fuzz_cases.random_refblur_scene: random scenes in the reference's own blur mode (moving "rectangle" shapes, bumped
    leaf boxes: the node-by-node walk of the replayed reference tree); the oracle is pinned on the compiled reference for
    these scenes in tests/test_oracle_fuzz.py; tools/gpu_fuzz.py ran 400 seeds bit-identical."""
    from fuzz_cases import random_refblur_scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    _, scene, s = random_refblur_scene(seed)
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (seed, st)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_cuda_matches_oracle_on_random_glass(oracle_lib, seed):
    """Random scenes with closed blocks of glass triangles: Fresnel split, total internal reflection, rays that start
    inside the glass, deeper recursion (fuzz_cases.random_glass_scene; tools/gpu_fuzz.py ran 400 seeds bit-identical)."""
    from fuzz_cases import random_glass_scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    _, scene, s = random_glass_scene(seed)
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (seed, st)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_cuda_matches_oracle_on_random_prisms(oracle_lib, seed):
    """Random slab-box prisms (RectPrism / RectPrismWithCylinder / RectPrismWithHoles with random holes, a14), seen from
    anywhere, the inside included (fuzz_cases.random_prism_scene; tools/gpu_fuzz.py ran 400 seeds bit-identical; about 9 %
    of the pixels of these scenes are ones where the reference throws, which both sides abandon)."""
    from fuzz_cases import random_prism_scene
    from oracle.harness import Oracle, ORACLE_KEYED, compare
    _, scene, s = random_prism_scene(seed)
    want, _, _, _ = Oracle(scene).render(s, mode=ORACLE_KEYED)
    got, _ = _gpu(scene).render_float(s)
    st = compare(want, got)
    assert st["frac_within_1"] >= TOL_FRAC, (seed, st)


@pytest.mark.parametrize("variant", ["c2", "perlin_aa10", "chunks"])
def test_render_multi_equals_the_single_device_frame(oracle_lib, variant, monkeypatch):
    """drt_render_multi (one frame on several scene handles, units of ~1024 samples claimed from one shared counter, every
    handle resolving its own pixels into the gathering handle's frame) against drt_render: byte-identical, whoever rendered
    which unit.  With one GPU the handles share the device; with more, every GPU takes part (peer access)."""
    from distraytracer_b200 import runtime, scenes, abi
    if variant == "c2":
        scene, s = scenes.config2(240, 135, 16)
    elif variant == "perlin_aa10":
        scene, s = scenes.config3(97, 53, 10)      # 9 spp of 10 lens points: units of 114 pixels, ragged last unit, cloud corners
        s.aperture = 0.15
    else:
        scene, s = scenes.config2(160, 90, 16)
        monkeypatch.setenv("DRT_CHUNK_LOG2", "16")   # 65536 samples per launch: the frame is several row chunks
    ndev = runtime.device_count()
    devs = list(range(ndev)) if ndev > 1 else [0, 0]
    handles = [runtime.DeviceScene(scene, d) for d in devs]
    want = handles[0].render(s)
    got, cnt = runtime.render_multi(handles, s, counters=True)
    assert np.array_equal(want, got)
    assert got.std() > 5 and all(c.kernel_launches >= 2 for c in cnt)
    # a second frame through the same handles (counters and ownership maps are reset per call)
    s.seed += 1
    assert np.array_equal(handles[-1].render(s), runtime.render_multi(handles, s))
    # sub-tile
    tile = abi.Tile(13, 7, 50, 31, 0)
    assert np.array_equal(handles[0].render(s, tile), runtime.render_multi(handles, s, tile=tile))
