"""Mutated fixture scenes for the GPU-vs-oracle fuzz (tests/test_gpu_parity.py, tools/gpu_fuzz.py): stored fixture
scenes with random motion flags, BRDF models, roughness, reflective materials, glossy flags and settings, and random scenes
no builder of the reference ever produced.  The CPU twin (tests/test_oracle_fuzz.py) pins the oracle on the compiled
reference for the same mutations and the same random scenes."""
import numpy as np

from conftest import GOLDEN_CASES, load_case


def mutated_case(seed):
    """-> (fixture name, Scene, settings); deterministic in `seed`."""
    from distraytracer_b200 import abi
    from distraytracer_b200.scene import Scene
    rng = np.random.default_rng(9000 + seed)
    case = GOLDEN_CASES[int(rng.integers(len(GOLDEN_CASES)))]
    scene, settings, _ = load_case(case)
    s = abi.copy_struct(settings)
    s.xRes, s.yRes = int(rng.integers(24, 65)), int(rng.integers(18, 49))
    s.antialias_samples = int(rng.choice([1, 4, 9]))
    s.aperture = float(rng.choice([0.0, 0.2]))
    s.brdf_samples = int(rng.integers(1, 4))
    s.max_depth = int(rng.integers(1, 7))
    s.blur_samples = int(rng.integers(0, 4))
    s.frame_range = int(rng.integers(1, 9))
    if rng.random() < 0.4:
        s.frame_prism, s.frame_blur = 0, int(rng.choice([0, 100000]))
    s.seed = int(rng.integers(1, 1 << 30))
    prims = [abi.copy_struct(p) for p in scene.prims]
    for p in prims:
        if p.flags & abi.FLAG_LIGHT:
            continue
        if rng.random() < 0.3:
            p.flags ^= abi.FLAG_MOTION
        if rng.random() < 0.5:
            p.model = int(rng.choice([abi.MODEL_LAMBERT, abi.MODEL_OREN_NAYAR, abi.MODEL_COOK_TORRANCE]))
            p.roughness = float(np.float32(rng.uniform(0.1, 0.9)))
            p.refr[0], p.refr[1] = 0.958, 6.69
        if rng.random() < 0.3:
            p.material = int(rng.choice([abi.MAT_NONE, abi.MAT_STEEL, abi.MAT_ALUMINUM, abi.MAT_LINOLEUM]))
            if rng.random() < 0.5:
                p.flags ^= abi.FLAG_GLOSSY
    return case, Scene(prims, scene.lights, scene.textures), s


def random_scene(seed):
    """A scene the reference never built: random spheres, cylinders, triangles, rectangles (corners in any order, so some
    are the diagonal kind) and checkerboards around the origin, random materials / models, one to three lights."""
    import numpy as np
    from distraytracer_b200 import abi, scenes
    from distraytracer_b200.scene import Scene
    rng = np.random.default_rng(55000 + seed)
    base, settings, _ = load_case("checkertexture")
    P = lambda s=3.0: rng.uniform(-s, s, 3)
    col = lambda: tuple(rng.uniform(0.2, 1.0, 3))
    mat = lambda: int(rng.choice([abi.MAT_NONE, abi.MAT_NONE, abi.MAT_STEEL, abi.MAT_ALUMINUM]))
    mod = lambda: int(rng.choice([abi.MODEL_LAMBERT, abi.MODEL_OREN_NAYAR, abi.MODEL_COOK_TORRANCE]))
    prims = []
    for _ in range(int(rng.integers(3, 14))):
        kind = int(rng.integers(5))
        if kind == 0:
            p = scenes.sphere(P(), float(rng.uniform(0.2, 1.2)), col(), material=mat(), model=mod())
        elif kind == 1:
            a = P(); p = scenes.cylinder(a, a + rng.normal(0, 1.0, 3), float(rng.uniform(0.1, 0.5)), col(), material=mat(), model=mod())
        elif kind == 2:
            a = P(); p = scenes.triangle(a, a + rng.normal(0, 1.5, 3), a + rng.normal(0, 1.5, 3), col(), material=mat(), model=mod())
        else:
            a = P(4.0); u = rng.normal(0, 2.0, 3); v = np.cross(u, rng.normal(0, 1, 3)); v *= float(rng.uniform(0.5, 3.0)) / max(np.linalg.norm(v), 1e-9)
            corners = [a, a + u, a + u + v, a + v]
            if rng.random() < 0.35:
                corners = [corners[i] for i in rng.permutation(4)]       # any corner order: the reference takes what it is given
            if kind == 3:
                p = scenes.rectangle(*corners, col(), material=mat(), model=mod(), name=int(rng.choice([abi.NAME_RECTANGLE, abi.NAME_OTHER])))
            else:
                p = scenes.checkerboard(*corners, col(), col(), float(rng.uniform(0.2, 1.0)), material=mat(), model=mod())
        p.roughness = float(np.float32(rng.uniform(0.1, 0.9))); p.refr[0], p.refr[1] = 0.958, 6.69
        if rng.random() < 0.3: p.flags |= abi.FLAG_GLOSSY
        prims.append(p)
    lights = []
    for _ in range(int(rng.integers(1, 4))):
        kind = int(rng.integers(3))
        if kind == 0:
            lights.append(scenes.point_light(P(6.0) + np.array([0, 4, 0]), col()))
        elif kind == 1:
            a = P(3.0) + np.array([0, 5, 0]); u = np.array([float(rng.uniform(0.5, 2)), 0, 0]); v = np.array([0, 0, float(rng.uniform(0.5, 2))])
            lp, ll = scenes.rectangle_light(a, a + u, a + u + v, a + v, col(), len(prims)); prims.append(lp); lights.append(ll)
        else:
            lp, ll = scenes.sphere_light(P(4.0) + np.array([0, 5, 0]), float(rng.uniform(0.2, 0.8)), col(), len(prims)); prims.append(lp); lights.append(ll)
    s = abi.copy_struct(settings)
    s.xRes, s.yRes = int(rng.integers(24, 65)), int(rng.integers(18, 49))
    s.antialias_samples = int(rng.choice([1, 4, 9])); s.aperture = float(rng.choice([0.0, 0.2]))
    s.brdf_samples = int(rng.integers(1, 4)); s.max_depth = int(rng.integers(1, 6)); s.blur_samples = 0
    s.seed = int(rng.integers(1, 1 << 30))
    v = rng.normal(0, 1, 3); v /= np.linalg.norm(v)
    eye = v * float(rng.uniform(5, 9)); eye[1] = abs(eye[1]) * 0.6
    s.eye[:] = [float(x) for x in eye]; s.lookingAt[:] = [float(x) for x in rng.uniform(-0.5, 0.5, 3)]; s.up[:] = [0, 1, 0]
    s.focal_length = float(np.linalg.norm(eye))
    return "random_scene", Scene(prims, lights, base.textures), s


def random_mesh_scene(seed):
    """A random triangle mesh (a bumpy grid with random holes, or a soup of random triangles; 40 to ~2000 triangles, with
    texcoords and, half of the time, a per-triangle material table) plus a sphere or two, one or two lights, depth of
    field.  -> (name, Scene with the mesh -- what the CUDA path takes, through its device-built 4-wide LBVH --, the same
    scene with the triangles as Triangle primitives -- what the oracle's reference-order walk takes --, settings)."""
    import numpy as np
    from distraytracer_b200 import abi, scenes
    from distraytracer_b200.scene import Scene
    rng = np.random.default_rng(66000 + seed)
    base, settings, _ = load_case("checkertexture")
    if rng.random() < 0.6:
        n = int(rng.integers(6, 33))
        mesh = scenes.terrain_mesh(n, size=float(rng.uniform(3, 10)), height=float(rng.uniform(0.2, 2.0)))
        keep = rng.random(len(mesh["indices"])) > float(rng.uniform(0.0, 0.3))
        mesh["indices"] = np.ascontiguousarray(mesh["indices"][keep])
    else:
        nt = int(rng.integers(40, 400))
        a = rng.uniform(-4, 4, (nt, 1, 3)); a[:, :, 1] *= 0.3
        verts = (a + rng.normal(0, float(rng.uniform(0.2, 1.0)), (nt, 3, 3))).reshape(-1, 3).astype(np.float32)
        mesh = {"vertices": verts, "indices": np.arange(3 * nt, dtype=np.int32).reshape(nt, 3),
                "texcoords": rng.uniform(0, 1, (3 * nt, 2)).astype(np.float32)}
    textured = rng.random() < 0.5
    mesh["material"] = scenes.mesh_material(tex_frame=2 if textured else -1,
                                            model=int(rng.choice([abi.MODEL_LAMBERT, abi.MODEL_OREN_NAYAR, abi.MODEL_COOK_TORRANCE])),
                                            roughness=float(rng.uniform(0.2, 0.8)))
    mesh["material"].refr[0], mesh["material"].refr[1] = 0.958, 6.69
    if rng.random() < 0.5:
        mats = []
        for v in rng.uniform(0.1, 0.9, 5):
            m = abi.copy_struct(mesh["material"]); m.roughness = float(np.float32(v)); m.color[0] = float(v); mats.append(m)
        mesh["materials"] = mats
        mesh["material_ids"] = rng.integers(0, 5, len(mesh["indices"])).astype(np.int32)
    prims = [scenes.sphere(rng.uniform(-2, 2, 3) + np.array([0, 1.5, 0]), float(rng.uniform(0.3, 0.9)), tuple(rng.uniform(0.2, 1, 3)),
                           material=int(rng.choice([abi.MAT_NONE, abi.MAT_STEEL])))]
    if rng.random() < 0.5:
        prims.append(scenes.rectangle((-6, -1.5, 6), (6, -1.5, 6), (6, -1.5, -6), (-6, -1.5, -6), (0.6, 0.6, 0.6), name=abi.NAME_OTHER))
    lights = [scenes.point_light(rng.uniform(-5, 5, 3) + np.array([0, 7, 0]), (1.0, 1.0, 1.0))]
    if rng.random() < 0.5:
        lp, ll = scenes.sphere_light(rng.uniform(-3, 3, 3) + np.array([0, 6, 0]), 0.5, (0.8, 0.8, 1.0), len(prims))
        prims.append(lp); lights.append(ll)
    s = abi.copy_struct(settings)
    s.xRes, s.yRes = int(rng.integers(24, 65)), int(rng.integers(18, 49))
    s.antialias_samples = int(rng.choice([1, 4])); s.aperture = float(rng.choice([0.0, 0.2]))
    s.brdf_samples = int(rng.integers(1, 3)); s.max_depth = int(rng.integers(1, 5)); s.blur_samples = 0
    s.seed = int(rng.integers(1, 1 << 30))
    v = rng.normal(0, 1, 3); v /= np.linalg.norm(v)
    eye = v * float(rng.uniform(6, 11)); eye[1] = abs(eye[1]) * 0.7 + 1.0
    s.eye[:] = [float(x) for x in eye]; s.lookingAt[:] = [0.0, 0.0, 0.0]; s.up[:] = [0, 1, 0]
    s.focal_length = float(np.linalg.norm(eye))
    with_mesh = Scene(prims, lights, base.textures, mesh=mesh)
    flat = Scene(list(prims) + scenes.mesh_to_prims(mesh), lights, base.textures)
    return "random_mesh", with_mesh, flat, s


def random_motion_scene(seed):
    """random_scene with the library's own blur mode (DRT_BLUR_VELOCITY, SURVEY 8(f)1: the reference has no counterpart, the
    oracle twin is the checker): random velocities on flagged primitives, cylinders partly with two-pose end points."""
    import numpy as np
    from distraytracer_b200 import abi
    from distraytracer_b200.scene import Scene
    case, scene, s = random_scene(seed)
    rng = np.random.default_rng(88000 + seed)
    prims = [abi.copy_struct(p) for p in scene.prims]
    for p in prims:
        if (p.flags & abi.FLAG_LIGHT) or rng.random() < 0.5:
            continue
        p.flags |= abi.FLAG_MOTION
        v = rng.normal(0, 0.6, 3)
        p.velocity[0], p.velocity[1], p.velocity[2] = (float(x) for x in v)
        if p.type == abi.PRIM_CYLINDER and rng.random() < 0.6:
            p.flags |= abi.FLAG_VERTEX_MOTION
            w = rng.normal(0, 0.6, 3)
            p.velocity2[0], p.velocity2[1], p.velocity2[2] = (float(x) for x in w)
    s.blur_mode, s.blur_samples, s.frame_range = abi.BLUR_VELOCITY, int(rng.integers(1, 4)), 1
    return "random_motion", Scene(prims, scene.lights, scene.textures), s


def random_big_scene(seed):
    """More analytic shapes than the CUDA path's shared-memory filter table holds (257 .. 700): small spheres, cylinders and
    rectangles scattered over a floor -- the CUDA path gathers through its 4-wide tree over the geoms."""
    import numpy as np
    from distraytracer_b200 import abi, scenes
    from distraytracer_b200.scene import Scene
    rng = np.random.default_rng(99000 + seed)
    base, settings, _ = load_case("checkertexture")
    prims = [scenes.rectangle((-9, 0, 9), (9, 0, 9), (9, 0, -9), (-9, 0, -9), (0.7, 0.7, 0.7),
                              material=int(rng.choice([abi.MAT_NONE, abi.MAT_STEEL])), name=abi.NAME_OTHER)]
    for _ in range(int(rng.integers(257, 700))):
        c = rng.uniform(-8, 8, 3); c[1] = float(rng.uniform(0.1, 2.5))
        col = tuple(rng.uniform(0.2, 1.0, 3))
        kind = int(rng.integers(3))
        if kind == 0:
            prims.append(scenes.sphere(c, float(rng.uniform(0.08, 0.35)), col, material=int(rng.choice([abi.MAT_NONE, abi.MAT_NONE, abi.MAT_STEEL]))))
        elif kind == 1:
            prims.append(scenes.cylinder(c, c + rng.normal(0, 0.4, 3), float(rng.uniform(0.05, 0.15)), col))
        else:
            u = rng.normal(0, 0.4, 3); v = np.cross(u, rng.normal(0, 1, 3)); v *= 0.4 / max(np.linalg.norm(v), 1e-9)
            prims.append(scenes.rectangle(c, c + u, c + u + v, c + v, col, name=abi.NAME_OTHER))
    lp, ll = scenes.rectangle_light((-1, 6, -1), (1, 6, -1), (1, 6, 1), (-1, 6, 1), (1.0, 1.0, 0.9), len(prims))
    prims.append(lp)
    lights = [ll, scenes.point_light(rng.uniform(-6, 6, 3) + np.array([0, 8, 0]), (0.7, 0.7, 0.7))]
    s = abi.copy_struct(settings)
    s.xRes, s.yRes = int(rng.integers(24, 65)), int(rng.integers(18, 49))
    s.antialias_samples = int(rng.choice([1, 4])); s.aperture = float(rng.choice([0.0, 0.2]))
    s.brdf_samples = int(rng.integers(1, 3)); s.max_depth = int(rng.integers(1, 5)); s.blur_samples = 0
    s.seed = int(rng.integers(1, 1 << 30))
    v = rng.normal(0, 1, 3); v /= np.linalg.norm(v)
    eye = v * float(rng.uniform(7, 14)); eye[1] = abs(eye[1]) * 0.6 + 1.5
    s.eye[:] = [float(x) for x in eye]; s.lookingAt[:] = [0.0, 0.5, 0.0]; s.up[:] = [0, 1, 0]
    s.focal_length = float(np.linalg.norm(eye))
    return "random_big", Scene(prims, lights, base.textures), s


def sky_case(seed):
    """mutated_case with the value-noise cloud background switched on (perlin_cloud: cloud_corners + the per-sample corner
    offsets) at a small size -- the oracle marches 200 noise steps per pixel corner."""
    case, scene, s = mutated_case(seed)
    s.perlin_cloud = 1
    s.xRes, s.yRes = min(s.xRes, 40), min(s.yRes, 28)
    return case + "+sky", scene, s


def random_moving_mesh_scene(seed):
    """random_mesh_scene in velocity-blur mode with the MESH in motion (DRT_BLUR_VELOCITY moves a mesh as a whole by its
    materials' common velocity) and, half of the time, the sphere too."""
    import numpy as np
    from distraytracer_b200 import abi, scenes
    from distraytracer_b200.scene import Scene
    case, with_mesh, flat, s = random_mesh_scene(seed)
    rng = np.random.default_rng(44000 + seed)
    vel = rng.normal(0, 0.5, 3)
    mesh = dict(with_mesh.mesh)
    def move(m):
        m = abi.copy_struct(m); m.flags |= abi.FLAG_MOTION
        m.velocity[0], m.velocity[1], m.velocity[2] = (float(x) for x in vel)
        return m
    mesh["material"] = move(mesh["material"])
    if mesh.get("materials"):
        mesh["materials"] = [move(m) for m in mesh["materials"]]
    prims = [abi.copy_struct(p) for p in with_mesh.prims]
    if rng.random() < 0.5:
        prims[0].flags |= abi.FLAG_MOTION
        w = rng.normal(0, 0.5, 3); prims[0].velocity[0], prims[0].velocity[1], prims[0].velocity[2] = (float(x) for x in w)
    s.blur_mode, s.blur_samples, s.frame_range = abi.BLUR_VELOCITY, int(rng.integers(1, 4)), 1
    return ("random_moving_mesh", Scene(prims, with_mesh.lights, with_mesh.textures, mesh=mesh),
            Scene(list(prims) + scenes.mesh_to_prims(mesh), with_mesh.lights, with_mesh.textures), s)


def random_prism_scene(seed):
    """The slab-box prism classes (RectPrism / RectPrismWithCylinder / RectPrismWithHoles) with random
    extents and random holes (spheres and cylinders, inside, across and outside the box), over a floor, next to a sphere,
    under random lights, seen from anywhere -- including from inside the prism."""
    import numpy as np
    from distraytracer_b200 import abi, scenes
    from distraytracer_b200.scene import Scene
    rng = np.random.default_rng(33000 + seed)
    base, settings, _ = load_case("checkertexture")
    prims = [scenes.rectangle((-6, -2.2, -6), (8, -2.2, -6), (8, -2.2, 6), (-6, -2.2, 6), (0.7, 0.7, 0.7), name=abi.NAME_OTHER),
             scenes.sphere((3.0, 0.0, 2.6), 0.8, (0.2, 0.9, 0.3), material=int(rng.choice([abi.MAT_NONE, abi.MAT_STEEL])))]
    for _ in range(int(rng.integers(1, 3))):
        kind = int(rng.choice([abi.PRIM_RECTPRISM, abi.PRIM_RECTPRISM_CYL, abi.PRIM_RECTPRISM_HOLES]))
        lo = rng.uniform(-3, 1, 3); hi = lo + rng.uniform(0.6, 3.5, 3)
        p = scenes.new_prim(); p.type = kind
        A = np.array([lo[0], lo[1], lo[2]]); B = np.array([lo[0], lo[1], hi[2]]); C = np.array([lo[0], hi[1], hi[2]]); D = np.array([lo[0], hi[1], lo[2]])
        back = np.array([hi[0] - lo[0], 0, 0])
        for dst, v in zip((p.A, p.B, p.C, p.D, p.E, p.F, p.G, p.H), (A, B, C, D, A + back, B + back, C + back, D + back)):
            dst[:] = list(v)
        p.color[:] = list(rng.uniform(0.2, 1.0, 3)); p.center[:] = list((lo + hi) / 2)
        p.material = int(rng.choice([abi.MAT_NONE, abi.MAT_NONE, abi.MAT_STEEL]))
        p.model = int(rng.choice([abi.MODEL_LAMBERT, abi.MODEL_OREN_NAYAR])); p.roughness = float(np.float32(rng.uniform(0.2, 0.8)))
        if kind != abi.PRIM_RECTPRISM:
            p.n_holes = int(rng.integers(1, 4))
            for k in range(p.n_holes):
                sph = kind == abi.PRIM_RECTPRISM_HOLES and rng.random() < 0.5
                c1 = rng.uniform(lo - 0.5, hi + 0.5)
                p.holes[k].type = abi.PRIM_SPHERE if sph else abi.PRIM_CYLINDER
                p.holes[k].c1[:] = list(c1); p.holes[k].c2[:] = list(c1 + rng.normal(0, 1.2, 3))
                p.holes[k].radius = float(np.float32(rng.uniform(0.2, 1.0))); p.holes[k].color[:] = list(rng.uniform(0, 1, 3))
        prims.append(p)
    lights = [scenes.point_light(rng.uniform(-6, 6, 3) + np.array([0, 5, 0]), (1, 1, 1))]
    if rng.random() < 0.5:
        lights.append(scenes.point_light(rng.uniform(-6, 6, 3) + np.array([0, 4, 0]), (0.6, 0.6, 0.9)))
    s = abi.copy_struct(settings)
    s.xRes, s.yRes = int(rng.integers(24, 65)), int(rng.integers(18, 49))
    s.antialias_samples = int(rng.choice([1, 4])); s.aperture = float(rng.choice([0.0, 0.2]))
    s.brdf_samples = int(rng.integers(1, 3)); s.max_depth = int(rng.integers(1, 5)); s.blur_samples = 0
    s.seed = int(rng.integers(1, 1 << 30))
    v = rng.normal(0, 1, 3); v /= np.linalg.norm(v)
    eye = v * float(rng.uniform(0.5, 9)); eye[1] = abs(eye[1]) * 0.7
    s.eye[:] = [float(x) for x in eye]; s.lookingAt[:] = [float(x) for x in rng.uniform(-1, 1, 3)]; s.up[:] = [0, 1, 0]
    s.focal_length = float(max(np.linalg.norm(eye), 1.0))
    return "random_prism", Scene(prims, lights, base.textures), s


def random_glass_scene(seed):
    """random_scene plus one or two closed blocks of glass triangles (the reference's refraction path is only reachable on
    `mesh` triangles): Fresnel split, total internal reflection, rays starting inside the glass."""
    import numpy as np
    from distraytracer_b200 import abi, scenes
    from distraytracer_b200.scene import Scene
    case, scene, s = random_scene(seed)
    rng = np.random.default_rng(22000 + seed)
    prims = [abi.copy_struct(p) for p in scene.prims]
    lights = [abi.copy_struct(l) for l in scene.lights]
    extra = []
    for _ in range(int(rng.integers(1, 3))):
        lo = rng.uniform(-2.5, 1.5, 3); hi = lo + rng.uniform(0.5, 2.0, 3)
        extra.extend(scenes.glass_block(lo, hi))
    # area lights point at their shapes by index: the blocks go in front, the indices move up
    for l in lights:
        if l.prim_index >= 0: l.prim_index += len(extra)
    s.max_depth = int(rng.integers(2, 8))
    return "random_glass", Scene(extra + prims, lights, scene.textures), s


def random_refblur_scene(seed):
    """random_scene in the REFERENCE's blur mode after frame_prism: shapes named "rectangle" move in y during the blur
    re-traces of samples that hit a primitive flagged `motion` (render_final_project.cpp:1095-1210), leaf boxes are widened
    but interior ones are not (bumpBVH, Q14) -- on random geometry, rectangles of the diagonal kind included."""
    import numpy as np
    from distraytracer_b200 import abi
    from distraytracer_b200.scene import Scene
    case, scene, s = random_scene(seed)
    rng = np.random.default_rng(11000 + seed)
    prims = [abi.copy_struct(p) for p in scene.prims]
    for p in prims:
        if not (p.flags & abi.FLAG_LIGHT) and rng.random() < 0.5:
            p.flags |= abi.FLAG_MOTION
        # the reference's Rectangle constructor names every rectangle and checkerboard "rectangle" (geometry.cpp:618,636):
        # those are the shapes its blur moves.  The library takes the name as data; the reference cannot build the others.
        if p.type in (abi.PRIM_RECTANGLE, abi.PRIM_CHECKERBOARD, abi.PRIM_CHECKERBOARD_HOLE) and p.name == abi.NAME_OTHER:
            p.name = abi.NAME_RECTANGLE
    s.blur_mode = abi.BLUR_REFERENCE
    s.blur_samples, s.frame_range = int(rng.integers(1, 4)), int(rng.integers(1, 9))
    s.frame = int(rng.integers(0, 2000)); s.frame_prism = 0; s.frame_blur = int(rng.choice([0, 100000]))
    return "random_refblur", Scene(prims, scene.lights, scene.textures), s
