"""ctypes mirror of include/drt.h (the C ABI of the CUDA hot path).

Field order and types must match drt.h exactly; tests/test_abi.py checks the
struct sizes against the values the shared library reports.
"""
import ctypes as C

ABI_VERSION = 2

# drt_status
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA, ERR_SCENE = 0, -1, -2, -3, -4, -5

# drt_prim_type
(PRIM_SPHERE, PRIM_CYLINDER, PRIM_TRIANGLE, PRIM_RECTANGLE, PRIM_RECTPRISMV2,
 PRIM_CHECKERBOARD, PRIM_CHECKERBOARD_HOLE, PRIM_CHECKER_CYLINDER,
 PRIM_RECTPRISM, PRIM_RECTPRISM_CYL, PRIM_RECTPRISM_HOLES) = range(11)
PRIM_TYPE_COUNT = 11
MAX_HOLES = 4
# drt_name
NAME_OTHER, NAME_RECTANGLE, NAME_SPHERELIGHT, NAME_RECTANGLELIGHT = range(4)
# drt_material
MAT_NONE, MAT_GLASS, MAT_STEEL, MAT_ALUMINUM, MAT_WATER, MAT_LINOLEUM = range(6)
# drt_model
MODEL_LAMBERT, MODEL_OREN_NAYAR, MODEL_COOK_TORRANCE, MODEL_RAW = range(4)
# drt_prim_flags
FLAG_LIGHT, FLAG_MOTION, FLAG_TEXTURE, FLAG_GLOSSY, FLAG_MESH, FLAG_UV_VERTS, FLAG_VERTEX_MOTION = (1 << i for i in range(7))
# drt_light_type
LIGHT_POINT, LIGHT_SPHERE, LIGHT_RECT = range(3)
# drt_sample_mode / drt_blur_mode
SAMPLES_KEYED = 0
BLUR_REFERENCE, BLUR_VELOCITY = 0, 1
PRECISION_REFERENCE, PRECISION_FP32 = 0, 1

D3 = C.c_double * 3
D2 = C.c_double * 2


class Hole(C.Structure):
    _fields_ = [("type", C.c_int32), ("pad_", C.c_int32), ("c1", D3), ("c2", D3), ("radius", C.c_double), ("color", D3)]


class Prim(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("name", C.c_int32), ("material", C.c_int32), ("model", C.c_int32),
        ("flags", C.c_int32), ("tex_frame", C.c_int32),
        ("color", D3), ("bordercolor", D3), ("roughness", C.c_double), ("refr", D2),
        ("center", D3), ("radius", C.c_double),
        ("A", D3), ("B", D3), ("C", D3), ("D", D3), ("E", D3), ("F", D3), ("G", D3), ("H", D3),
        ("c1", D3), ("c2", D3), ("uvA", D2), ("uvB", D2), ("uvC", D2), ("mesh_normal", D3),
        ("S", C.c_double), ("borderwidth", C.c_double), ("color1", D3), ("color2", D3),
        ("hole", D3 * 4), ("velocity", D3),
        ("n_holes", C.c_int32), ("pad_", C.c_int32), ("holes", Hole * MAX_HOLES), ("velocity2", D3),
    ]


PRIM_BYTES_V1 = 624     # sizeof(drt_prim) of ABI version 1 (no holes): fixtures written then are padded on load


class Light(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("prim_index", C.c_int32),
        ("color", D3), ("center", D3), ("radius", C.c_double), ("baxis", D3),
        ("A", D3), ("B", D3), ("C", D3), ("D", D3),
    ]


class Texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_uint8))]


class Mesh(C.Structure):
    _fields_ = [
        ("n_vertices", C.c_int64), ("n_triangles", C.c_int64),
        ("vertices", C.POINTER(C.c_float)), ("indices", C.POINTER(C.c_int32)),
        ("texcoords", C.POINTER(C.c_float)), ("material", Prim),
        ("n_materials", C.c_int32), ("pad_", C.c_int32), ("materials", C.POINTER(Prim)), ("material_ids", C.POINTER(C.c_int32)),
    ]


class SceneDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_prims", C.c_int32), ("prims", C.POINTER(Prim)),
        ("n_lights", C.c_int32), ("lights", C.POINTER(Light)),
        ("n_textures", C.c_int32), ("textures", C.POINTER(Texture)),
        ("mesh", C.POINTER(Mesh)),
    ]


class Settings(C.Structure):
    _fields_ = [
        ("xRes", C.c_int32), ("yRes", C.c_int32),
        ("eye", D3), ("lookingAt", D3), ("up", D3),
        ("aspect", C.c_float), ("near_plane", C.c_float), ("fov", C.c_float),
        ("aperture", C.c_float), ("focal_length", C.c_float),
        ("nogloss", C.c_int32), ("refr_air", C.c_float), ("refr_glass", C.c_float),
        ("max_depth", C.c_int32), ("phong", C.c_float),
        ("antialias_samples", C.c_int32), ("brdf_samples", C.c_int32), ("blur_samples", C.c_int32),
        ("frame_range", C.c_int32), ("frame_prism", C.c_int32), ("frame_cloud", C.c_int32),
        ("frame_blur", C.c_int32), ("move_per_frame", C.c_float), ("accel_t", C.c_float),
        ("sundir", D3), ("perlin_cloud", C.c_int32), ("saturation", C.c_float),
        ("clouddist", C.c_float), ("cloudhoff", C.c_float),
        ("sun_outer", D3), ("sun_inner", D3), ("sun_core", D3), ("bluesky", D3), ("redsky", D3),
        ("reflect", C.c_int32), ("frame", C.c_int32), ("seed", C.c_uint32),
        ("sample_mode", C.c_int32), ("blur_mode", C.c_int32), ("cloud_only", C.c_int32),
        ("precision", C.c_int32),
    ]


class Tile(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("device", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [
        ("collect", C.c_int32), ("kernel_ms", C.c_float), ("kernel_launches", C.c_int32), ("kernel_variant", C.c_int32),
        ("samples", C.c_uint64), ("rays", C.c_uint64), ("shadow_rays", C.c_uint64),
        ("node_tests", C.c_uint64), ("prim_tests", C.c_uint64 * PRIM_TYPE_COUNT),
        ("shade_evals", C.c_uint64), ("noise_evals", C.c_uint64),
    ]


def copy_struct(s):
    """Deep copy of a ctypes structure (by value)."""
    out = type(s)()
    C.memmove(C.byref(out), C.byref(s), C.sizeof(s))
    return out
