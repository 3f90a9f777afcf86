"""Scene builders for the BASELINE.json configurations, producing the POD scene of
include/drt.h.  They play the role of the reference's build*() functions in scene.h
for the synthetic benchmark configurations.  The exported reference scenes they start
from are package data (distraytracer_b200/data/).
"""
import math
import os

import numpy as np

from . import abi
from .scene import Scene, load_fixture

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def data_path(name):
    """A file of the package's own data directory (written by tools/make_package_data.py): the reference builders'
    exported scenes the configurations start from, and the mocap clip."""
    return os.path.join(DATA, name)


def _v(dst, src):
    dst[0], dst[1], dst[2] = float(src[0]), float(src[1]), float(src[2])


def new_prim():
    p = abi.Prim()
    p.tex_frame = -1
    return p


def sphere(center, radius, color, material=abi.MAT_NONE, motion=False, model=abi.MODEL_LAMBERT):
    """Sphere(c, r, col, material, in_motion, shader), geometry.cpp:94-104."""
    p = new_prim()
    p.type, p.material, p.model = abi.PRIM_SPHERE, material, model
    p.flags = abi.FLAG_MOTION if motion else 0
    _v(p.center, center); _v(p.color, color)
    p.radius = float(np.float32(radius))
    return p


def cylinder(c1, c2, radius, color, material=abi.MAT_NONE, motion=False, model=abi.MODEL_LAMBERT):
    """Cylinder(v1, v2, r, col, ...), geometry.cpp:227-240."""
    p = new_prim()
    p.type, p.material, p.model = abi.PRIM_CYLINDER, material, model
    p.flags = abi.FLAG_MOTION if motion else 0
    _v(p.c1, c1); _v(p.c2, c2); _v(p.color, color)
    _v(p.center, (np.asarray(c1, float) + np.asarray(c2, float)) / 2)
    p.radius = float(np.float32(radius))
    return p


def rectangle(a, b, c, d, color, material=abi.MAT_NONE, motion=False, tex_frame=-1, model=abi.MODEL_LAMBERT,
              name=abi.NAME_RECTANGLE):
    """Rectangle(a,b,c,d,col,material,in_motion,texframe,shader), geometry.cpp:621-638."""
    p = new_prim()
    p.type, p.name, p.material, p.model, p.tex_frame = abi.PRIM_RECTANGLE, name, material, model, tex_frame
    p.flags = abi.FLAG_MOTION if motion else 0
    for dst, src in ((p.A, a), (p.B, b), (p.C, c), (p.D, d)):
        _v(dst, src)
    _v(p.color, color)
    _v(p.center, (np.asarray(a, float) + np.asarray(b, float) + np.asarray(c, float) + np.asarray(d, float)) / 4)
    return p


def rectangle_light(a, b, c, d, color, prim_index):
    """rectangleLight(a,b,c,d,col), geometry.cpp:2828-2843: a Rectangle that is also a light."""
    p = rectangle(a, b, c, d, color, name=abi.NAME_RECTANGLELIGHT)
    p.flags |= abi.FLAG_LIGHT
    l = abi.Light()
    l.type, l.prim_index = abi.LIGHT_RECT, prim_index
    _v(l.color, color); _v(l.center, p.center)
    for dst, src in ((l.A, a), (l.B, b), (l.C, c), (l.D, d)):
        _v(dst, src)
    return p, l


def point_light(center, color):
    l = abi.Light()
    l.type, l.prim_index = abi.LIGHT_POINT, -1
    _v(l.center, center); _v(l.color, color)
    return l


def sphere_light(center, radius, color, prim_index):
    """sphereLight(c, r, col), geometry.cpp:2756-2768."""
    p = sphere(center, radius, color)
    p.name = abi.NAME_SPHERELIGHT
    p.flags |= abi.FLAG_LIGHT
    l = abi.Light()
    l.type, l.prim_index = abi.LIGHT_SPHERE, prim_index
    _v(l.center, center); _v(l.color, color)
    l.radius = float(np.float32(radius))
    return p, l


def triangle(a, b, c, color, material=abi.MAT_NONE, motion=False, model=abi.MODEL_LAMBERT, mesh_normal=None):
    """Triangle(a,b,c,col,material,in_motion,shader), geometry.cpp:433-445; `mesh_normal` sets
    GeoPrimitive::mesh / mesh_normal the way the scene builders do for closed meshes (scene.h:3994-3996)."""
    p = new_prim()
    p.type, p.material, p.model = abi.PRIM_TRIANGLE, material, model
    p.flags = abi.FLAG_MOTION if motion else 0
    for dst, src in ((p.A, a), (p.B, b), (p.C, c)):
        _v(dst, src)
    _v(p.color, color)
    _v(p.center, (np.asarray(a, float) + np.asarray(b, float) + np.asarray(c, float)) / 3)
    if mesh_normal is not None:
        p.flags |= abi.FLAG_MESH
        _v(p.mesh_normal, mesh_normal)
    return p


def glass_block(lo, hi, color=(1.0, 1.0, 1.0)):
    """Closed axis-aligned block of 12 glass triangles with outward mesh normals."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    c = [np.array([x, y, z]) for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])]
    quads = [((0, 1, 3, 2), (-1, 0, 0)), ((4, 6, 7, 5), (1, 0, 0)), ((0, 4, 5, 1), (0, -1, 0)),
             ((2, 3, 7, 6), (0, 1, 0)), ((0, 2, 6, 4), (0, 0, -1)), ((1, 5, 7, 3), (0, 0, 1))]
    tris = []
    for (i0, i1, i2, i3), n in quads:
        tris.append(triangle(c[i0], c[i1], c[i2], color, material=abi.MAT_GLASS, mesh_normal=n))
        tris.append(triangle(c[i0], c[i2], c[i3], color, material=abi.MAT_GLASS, mesh_normal=n))
    return tris


def checkerboard(a, b, c, d, col1, col2, S, material=abi.MAT_NONE, model=abi.MODEL_LAMBERT):
    """Checkerboard(a,b,c,d,col1,col2,S), geometry.cpp:2248-2267."""
    p = rectangle(a, b, c, d, col1, material=material, model=model, name=abi.NAME_OTHER)
    p.type = abi.PRIM_CHECKERBOARD
    _v(p.color1, col1); _v(p.color2, col2)
    p.S = float(np.float32(S))
    return p


# ---------------------------------------------------------------------------
def config1():
    """BASELINE config 1: `./render test checkertexture` (render_final_project.cpp:1840-1853,
    scene.h:3052-3163) at 640x480, 1 spp.  The scene is the reference builder's own output
    (data/checkertexture_scene.npz, exported by oracle/ref_driver.cpp)."""
    scene, settings, _ = load_fixture(data_path("checkertexture_scene.npz"))
    settings.xRes, settings.yRes, settings.antialias_samples, settings.aperture = 640, 480, 1, 0.0
    return scene, settings


def config2(xres=1920, yres=1080, spp=64):
    """BASELINE config 2: the same scene at 1080p, 64 spp, with every distributed effect on:
    depth of field (aperture 0.2, focal length 10), glossy reflection on the floor as well as
    the doors, soft shadows from a rectangle light, Fresnel refraction through a glass sphere
    (SURVEY.md 8d C2)."""
    scene, settings = config1()
    prims = [abi.copy_struct(p) for p in scene.prims]
    lights = [abi.copy_struct(l) for l in scene.lights]
    for p in prims:
        if p.type == abi.PRIM_CHECKERBOARD_HOLE:
            p.flags |= abi.FLAG_GLOSSY                           # floor->reflect_params.glossy = true
    prims.extend(glass_block((2.2, 0.32, 0.7), (3.2, 1.3, 1.7)))
    lp, ll = rectangle_light((-1.5, 9.0, -1.0), (3.5, 9.0, -1.0), (3.5, 9.0, 3.0), (-1.5, 9.0, 3.0), (1.0, 1.0, 1.0),
                             prim_index=len(prims))
    prims.append(lp)
    lights.append(ll)
    settings.xRes, settings.yRes, settings.antialias_samples = xres, yres, spp
    settings.aperture, settings.focal_length = 0.2, 10.0
    return Scene(prims, lights, scene.textures), settings


def config3(xres=1920, yres=1080, spp=256):
    """BASELINE config 3: Oren-Nayar spheres (buildSceneReflectance, scene.h:3668-3694) in front of
    the value-noise cloud background (perlin_cloud), 1080p, 256 spp."""
    scene, settings, _ = load_fixture(data_path("reflectance_scene.npz"))
    prims = [abi.copy_struct(p) for p in scene.prims]
    for p in prims:
        if p.model != abi.MODEL_OREN_NAYAR and p.material == abi.MAT_NONE:
            p.model = abi.MODEL_OREN_NAYAR
            p.roughness = float(np.float32(math.sqrt(0.2)))
    settings.xRes, settings.yRes, settings.antialias_samples, settings.perlin_cloud = xres, yres, spp, 1
    return Scene(prims, scene.lights, scene.textures), settings


def config4_frame(frame, xres=1920, yres=1080, spp=16, bones=None, two_pose=False):
    """BASELINE config 4: mocap skeleton (29-30 bone cylinders, scene.h:637-659) over a
    checkerboard floor with two sphere lights (buildSceneChkpt2 layout, scene.h:3557-3666),
    bones flagged `motion` and given the velocity to their pose one frame later: one translation per bone, or
    (two_pose) each end point to its own next position (DRT_FLAG_VERTEX_MOTION)."""
    scene, settings, _ = load_fixture(data_path("chkpt2_mocap_scene.npz"))
    if bones is None:
        bones = np.load(data_path("mocap_bones_880_999.npy"))
    f0 = int(frame) % bones.shape[0]
    f1 = min(f0 + 1, bones.shape[0] - 1)
    prims, k = [], 0
    for p in scene.prims:
        q = abi.copy_struct(p)
        if q.type == abi.PRIM_CYLINDER:
            c1, c2 = bones[f0, k, 0], bones[f0, k, 1]
            n1, n2 = bones[f1, k, 0], bones[f1, k, 1]
            _v(q.c1, c1); _v(q.c2, c2); _v(q.center, (c1 + c2) / 2)
            if two_pose:
                _v(q.velocity, n1 - c1); _v(q.velocity2, n2 - c2)
                q.flags |= abi.FLAG_VERTEX_MOTION
            else:
                _v(q.velocity, ((n1 + n2) - (c1 + c2)) / 2)
            q.flags |= abi.FLAG_MOTION
            k += 1
        prims.append(q)
    settings.xRes, settings.yRes, settings.antialias_samples = xres, yres, spp
    settings.frame, settings.blur_mode, settings.blur_samples, settings.frame_range = int(frame), abi.BLUR_VELOCITY, 2, 1
    return Scene(prims, scene.lights, scene.textures), settings


# ---------------------------------------------------------------------------

def many_shapes(nx=24, nz=14, xres=160, yres=90, spp=4):
    """More analytic shapes than the shared-memory slab table holds (DRT_SMEM_GEOMS = 256): a field of nx * nz small
    spheres and short cylinders (every seventh sphere a steel mirror) over a steel floor, a rectangle light and a point
    light.  Exercises the geom-tree path of closestHit / anyHit (FT_BIG) against the oracle's reference-order walk."""
    base, settings, _ = load_fixture(data_path("checkertexture_scene.npz"))
    prims = [rectangle((-8, 0, 4), (8, 0, 4), (8, 0, -12), (-8, 0, -12), (0.7, 0.7, 0.75), material=abi.MAT_STEEL, name=abi.NAME_OTHER)]
    k = 0
    for i in range(nx):
        for j in range(nz):
            x, z = -6.9 + 0.6 * i, 2.0 - 0.9 * j
            y = 0.25 + 0.12 * ((i * 7 + j * 3) % 5)
            col = (0.3 + 0.1 * (i % 7), 0.9 - 0.1 * (j % 6), 0.4 + 0.05 * ((i + j) % 9))
            if k % 3 == 2:
                prims.append(cylinder((x - 0.15, y, z), (x + 0.15, y + 0.2, z - 0.1), 0.08, col))
            else:
                prims.append(sphere((x, y, z), 0.2, col, material=abi.MAT_STEEL if k % 7 == 0 else abi.MAT_NONE,
                                    model=abi.MODEL_OREN_NAYAR if k % 5 == 1 else abi.MODEL_LAMBERT))
            k += 1
    lp, ll = rectangle_light((-1, 5, -3), (1, 5, -3), (1, 5, -5), (-1, 5, -5), (1.0, 1.0, 0.9), len(prims))
    prims.append(lp)
    lights = [ll, point_light((3.0, 4.0, 3.0), (0.6, 0.6, 0.7))]
    s = abi.copy_struct(settings)
    s.eye[:] = [0.0, 3.5, 6.0]; s.lookingAt[:] = [0.0, 0.3, -3.0]; s.up[:] = [0, 1, 0]
    s.xRes, s.yRes, s.antialias_samples, s.aperture, s.focal_length = xres, yres, spp, 0.1, 9.0
    s.max_depth, s.brdf_samples, s.blur_samples = 4, 2, 0
    return Scene(prims, lights, base.textures), s


def terrain_mesh(n=708, size=12.0, height=1.2, origin=(-6.0, -0.5, -6.0)):
    """Procedural grid mesh for BASELINE config 5: n x n vertices -> 2 (n-1)^2 triangles
    (n = 708 gives 999 698), heights from a fixed analytic function (no RNG), UV = grid / (n-1)
    (SURVEY.md 8d C5).  Returns dict(vertices float32 (V,3), indices int32 (T,3), texcoords float32 (V,2))."""
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    u, v = i / (n - 1), j / (n - 1)
    x = origin[0] + size * u
    z = origin[2] + size * v
    y = origin[1] + height * (0.5 * np.sin(3.1 * u * np.pi) * np.cos(2.3 * v * np.pi) + 0.25 * np.sin(9.0 * u + 4.0 * v) +
                              0.1 * np.cos(23.0 * u * v))
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    tc = np.stack([u, v], axis=-1).reshape(-1, 2).astype(np.float32)
    vid = (i * n + j)
    a, b, c, d = vid[:-1, :-1], vid[1:, :-1], vid[1:, 1:], vid[:-1, 1:]
    tris = np.concatenate([np.stack([a, b, c], axis=-1).reshape(-1, 3), np.stack([a, c, d], axis=-1).reshape(-1, 3)])
    return {"vertices": verts, "indices": tris.astype(np.int32), "texcoords": tc}


def mesh_material(tex_frame=-1, model=abi.MODEL_OREN_NAYAR, roughness=0.5, color=(0.8, 0.8, 0.8)):
    """"marble" / oren-nayar roughness 0.5 as the reference's model meshes use (scene.h:564-569)."""
    p = new_prim()
    p.type, p.model, p.material = abi.PRIM_TRIANGLE, model, abi.MAT_NONE
    p.roughness = float(np.float32(roughness))
    _v(p.color, color)
    if tex_frame >= 0:
        p.tex_frame = tex_frame
        p.flags |= abi.FLAG_TEXTURE | abi.FLAG_UV_VERTS
    return p


def mesh_to_prims(mesh):
    """The same triangles as individual Triangle primitives (what the reference's loadObj path
    produces, scene.h:322-386) -- used to feed the CPU oracle, which has no mesh type."""
    V = mesh["vertices"].astype(np.float64)
    T = mesh["indices"]
    tc = mesh.get("texcoords")
    mat = mesh["material"]
    mats, mids = mesh.get("materials"), mesh.get("material_ids")
    out = []
    for ti, t in enumerate(T):
        p = abi.copy_struct(mats[mids[ti]] if mats else mat)
        p.type = abi.PRIM_TRIANGLE
        _v(p.A, V[t[0]]); _v(p.B, V[t[1]]); _v(p.C, V[t[2]])
        _v(p.center, (V[t[0]] + V[t[1]] + V[t[2]]) / 3)
        if tc is not None:
            for dst, k in ((p.uvA, 0), (p.uvB, 1), (p.uvC, 2)):
                dst[0], dst[1] = float(tc[t[k]][0]), float(tc[t[k]][1])
        out.append(p)
    return out


def config5(n=708, xres=3840, yres=2160, spp=64):
    """BASELINE config 5: textured ~1M-triangle terrain, DOF + motion blur (a moving sphere over the static mesh,
    blur_samples 2) + point light at the eye, 4K 64 spp."""
    base, settings, _ = load_fixture(data_path("checkertexture_scene.npz"))
    mesh = terrain_mesh(n)
    mesh["material"] = mesh_material(tex_frame=2)            # textures/floor.jpeg of the fixture
    ball = sphere((0.0, 1.6, 0.0), 0.7, (1.0, 0.2, 0.2), motion=True)
    _v(ball.velocity, (0.6, 0.0, 0.25))
    prims = [ball]
    s = abi.copy_struct(settings)
    s.blur_mode, s.blur_samples, s.frame_range = abi.BLUR_VELOCITY, 2, 1
    s.eye[:] = [0.0, 7.0, 9.0]; s.lookingAt[:] = [0.0, 0.0, 0.0]; s.up[:] = [0, 1, 0]
    s.xRes, s.yRes, s.antialias_samples, s.aperture, s.focal_length = xres, yres, spp, 0.2, 10.0
    lights = [point_light(tuple(s.eye), (1.0, 1.0, 1.0))]
    return Scene(prims, lights, base.textures, mesh=mesh), s
