"""ctypes binding of libdrt.so (include/drt.h).  Thin: every call goes straight
to the C ABI.  If the CUDA library has not been built or there is no GPU, calls
fail loudly -- there is no CPU fallback in this package.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DRT_LIB") or os.path.join(_HERE, "libdrt.so")   # DRT_LIB: A/B builds during tuning
_lib = None

EXPORTS = [
    "drt_device_count", "drt_settings_default", "drt_prim_default", "drt_scene_create", "drt_scene_update_prims",
    "drt_scene_update_lights",
    "drt_scene_destroy", "drt_render", "drt_render_float", "drt_render_device", "drt_render_multi", "drt_write_ppm", "drt_last_error",
    "drt_abi_sizes", "drt_debug_rng", "drt_debug_candidate_order", "drt_debug_shuffle_j", "drt_debug_lens_index",
    "drt_skeleton_create", "drt_skeleton_load", "drt_skeleton_info", "drt_skeleton_bones", "drt_scene_pose_skeleton",
    "drt_skeleton_destroy", "drt_debug_skeleton_parse",
]


class DrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"drt error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a).  distraytracer_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.drt_last_error.restype = C.c_char_p
        L.drt_device_count.restype = C.c_int
        L.drt_settings_default.argtypes = [C.POINTER(abi.Settings)]
        L.drt_prim_default.argtypes = [C.POINTER(abi.Prim)]
        L.drt_scene_create.argtypes = [C.POINTER(abi.SceneDesc), C.c_int, C.POINTER(C.c_void_p)]
        L.drt_scene_update_prims.argtypes = [C.c_void_p, C.POINTER(abi.Prim), C.c_int32]
        L.drt_scene_update_lights.argtypes = [C.c_void_p, C.POINTER(abi.Light), C.c_int32]
        L.drt_scene_destroy.argtypes = [C.c_void_p]
        L.drt_scene_destroy.restype = None
        L.drt_render.argtypes = [C.c_void_p, C.POINTER(abi.Settings), C.POINTER(abi.Tile), C.c_void_p, C.POINTER(abi.Counters)]
        L.drt_render_float.argtypes = [C.c_void_p, C.POINTER(abi.Settings), C.POINTER(abi.Tile), C.c_void_p, C.c_void_p,
                                       C.POINTER(abi.Counters)]
        L.drt_render_device.argtypes = [C.c_void_p, C.POINTER(abi.Settings), C.POINTER(abi.Tile), C.POINTER(abi.Counters)]
        L.drt_render_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(abi.Settings), C.POINTER(abi.Tile), C.c_void_p,
                                       C.POINTER(abi.Counters)]
        L.drt_write_ppm.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_void_p]
        L.drt_abi_sizes.argtypes = [C.POINTER(C.c_int32)]
        L.drt_debug_rng.argtypes = [C.c_uint32] * 5
        L.drt_debug_rng.restype = C.c_float
        L.drt_debug_shuffle_j.argtypes = [C.c_uint32, C.c_uint32, C.c_int32]
        L.drt_debug_lens_index.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.c_int32]
        L.drt_debug_candidate_order.argtypes = [C.POINTER(abi.Prim), C.c_int32, C.POINTER(C.c_int32), C.c_int32]
        L.drt_skeleton_create.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_double, C.c_int, C.POINTER(C.c_void_p)]
        L.drt_skeleton_load.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_int, C.POINTER(C.c_void_p)]
        L.drt_skeleton_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.drt_skeleton_bones.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.drt_scene_pose_skeleton.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_int32]
        L.drt_debug_skeleton_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_double, C.POINTER(C.c_int32),
                                               C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_int32]
        L.drt_skeleton_destroy.argtypes = [C.c_void_p]
        L.drt_skeleton_destroy.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise DrtError(rc, lib().drt_last_error().decode())


def device_count():
    return lib().drt_device_count()


def default_settings():
    s = abi.Settings()
    lib().drt_settings_default(C.byref(s))
    return s


def default_prim():
    p = abi.Prim()
    lib().drt_prim_default(C.byref(p))
    return p


def check_frame_buffer(out, height, width):
    """The C side writes height*width*3 bytes through a raw pointer: anything but a C-contiguous uint8
    (height, width, 3) array would be silent memory corruption."""
    if not isinstance(out, np.ndarray) or out.dtype != np.uint8 or out.shape != (height, width, 3) or not out.flags["C_CONTIGUOUS"]:
        raise ValueError(f"output buffer must be a C-contiguous uint8 array of shape ({height}, {width}, 3)")
    if not out.flags["WRITEABLE"]:
        raise ValueError("output buffer is read-only")


class DeviceScene:
    """drt_scene handle: the scene resident in one GPU's HBM."""

    def __init__(self, scene, device=0):
        self.scene = scene
        self.device = device
        self._desc = scene.desc()
        self.handle = C.c_void_p()
        _check(lib().drt_scene_create(C.byref(self._desc), device, C.byref(self.handle)))

    def close(self):
        if getattr(self, "handle", None):
            lib().drt_scene_destroy(self.handle)
            self.handle = None

    __del__ = close

    def update_prims(self, prims):
        arr = (abi.Prim * len(prims))(*prims)
        _check(lib().drt_scene_update_prims(self.handle, arr, len(prims)))

    def update_lights(self, lights):
        arr = (abi.Light * max(1, len(lights)))(*lights)
        _check(lib().drt_scene_update_lights(self.handle, arr, len(lights)))

    def pose_skeleton(self, skeleton, frame, first_prim, drop_y=0.0, set_velocity=True):
        """drt_scene_pose_skeleton: bone cylinders prims[first_prim...] take the pose of mocap frame `frame`.
        set_velocity: 0 / False static, 1 / True one translation per bone, 2 both end points to their own next pose."""
        _check(lib().drt_scene_pose_skeleton(self.handle, skeleton.handle, int(frame), int(first_prim), float(drop_y),
                                             int(set_velocity)))

    def _tile(self, settings, tile):
        if tile is None:
            return abi.Tile(0, 0, settings.xRes, settings.yRes, self.device)
        return tile

    def render(self, settings, tile=None, out=None, counters=None):
        """drt_render: uint8 (h,w,3) in PPM row order, written into `out` (host) if given."""
        tile = self._tile(settings, tile)
        if out is None:
            out = np.empty((tile.height, tile.width, 3), dtype=np.uint8)
        check_frame_buffer(out, tile.height, tile.width)
        _check(lib().drt_render(self.handle, C.byref(settings), C.byref(tile), out.ctypes.data,
                                C.byref(counters) if counters is not None else None))
        return out

    def render_float(self, settings, tile=None, counters=None):
        tile = self._tile(settings, tile)
        f = np.empty((tile.height, tile.width, 3), dtype=np.float32)
        u = np.empty((tile.height, tile.width, 3), dtype=np.uint8)
        _check(lib().drt_render_float(self.handle, C.byref(settings), C.byref(tile), f.ctypes.data, u.ctypes.data,
                                      C.byref(counters) if counters is not None else None))
        return f, u

    def render_device(self, settings, tile=None, counters=None):
        tile = self._tile(settings, tile)
        _check(lib().drt_render_device(self.handle, C.byref(settings), C.byref(tile),
                                       C.byref(counters) if counters is not None else None))


def render_multi(handles, settings, out=None, tile=None, counters=False):
    """drt_render_multi: ONE frame on all of `handles` (DeviceScene objects of the same scene, normally one per GPU;
    handles[0]'s GPU gathers).  Returns the uint8 (h, w, 3) frame, or (frame, [abi.Counters per handle]) with counters=True."""
    if tile is None:
        tile = abi.Tile(0, 0, settings.xRes, settings.yRes, handles[0].device)
    if out is None:
        out = np.empty((tile.height, tile.width, 3), dtype=np.uint8)
    check_frame_buffer(out, tile.height, tile.width)
    arr = (C.c_void_p * len(handles))(*[h.handle for h in handles])
    cnt = (abi.Counters * len(handles))() if counters else None
    _check(lib().drt_render_multi(arr, len(handles), C.byref(settings), C.byref(tile), out.ctypes.data, cnt))
    return (out, list(cnt)) if counters else out


MOCAP_SCALE = 0.06   # types.h:6


def parse_skeleton(asf, amc, scale=MOCAP_SCALE):
    """Host half of the mocap ingest (no GPU): (n_bones incl. root, n_frames, parents[int32], dofs[int32 bit masks])."""
    nb, nf = C.c_int32(), C.c_int32()
    parents = np.full(256, -2, dtype=np.int32); dofs = np.zeros(256, dtype=np.int32)
    _check(lib().drt_debug_skeleton_parse(asf, len(asf), amc, len(amc), scale, C.byref(nb), C.byref(nf),
                                          parents.ctypes.data, dofs.ctypes.data, 256))
    return nb.value, nf.value, parents[:nb.value].copy(), dofs[:nb.value].copy()


class DeviceSkeleton:
    """drt_skeleton handle: an ASF/AMC clip posed for all frames on one GPU (table of bone cylinders in HBM)."""

    def __init__(self, asf, amc, scale=MOCAP_SCALE, device=0):
        """`asf` / `amc`: file contents as bytes, or paths (str)."""
        self.handle = C.c_void_p()
        if isinstance(asf, str) and isinstance(amc, str):
            _check(lib().drt_skeleton_load(asf.encode(), amc.encode(), scale, device, C.byref(self.handle)))
        else:
            _check(lib().drt_skeleton_create(asf, len(asf), amc, len(amc), scale, device, C.byref(self.handle)))
        nc, nf, ms = C.c_int32(), C.c_int32(), C.c_float()
        _check(lib().drt_skeleton_info(self.handle, C.byref(nc), C.byref(nf), C.byref(ms)))
        self.n_cylinders, self.n_frames, self.fk_ms = nc.value, nf.value, ms.value

    def bones(self, frame0=0, n_frames=None):
        """(n_frames, n_cylinders, 2, 3) float64 end points, copied from the device table."""
        n = self.n_frames - frame0 if n_frames is None else n_frames
        out = np.empty((n, self.n_cylinders, 2, 3), dtype=np.float64)
        _check(lib().drt_skeleton_bones(self.handle, int(frame0), int(n), out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "handle", None):
            lib().drt_skeleton_destroy(self.handle)
            self.handle = None

    __del__ = close


def write_ppm(path, rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    _check(lib().drt_write_ppm(path.encode(), rgb.shape[1], rgb.shape[0], rgb.ctypes.data))
