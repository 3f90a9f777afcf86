"""TEST INFRASTRUCTURE ONLY -- ctypes front ends for the two CPU checkers:

  * `Ref`    : oracle/_ref/libdrt_ref.so, the UNMODIFIED reference sources compiled
               against oracle/eigen_shim (see oracle/ref_driver.cpp);
  * `Oracle` : oracle/liboracle.so, the from-scratch CPU restatement
               (oracle/drt_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.  Nothing under distraytracer_b200/ may.
"""
import ctypes as C
import os
import tempfile

import numpy as np

from distraytracer_b200 import abi
from distraytracer_b200.scene import Scene

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.environ.get("DRT_REF_SO", os.path.join(HERE, "_ref", "libdrt_ref.so"))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REFERENCE_ROOT = "/root/reference"


def ref_available():
    return os.path.exists(REF_SO)


def read_ppm(path):
    with open(path, "rb") as f:
        data = f.read()
    # "P6\n<w> <h>\n255\n" exactly as helpers.h:191 writes it
    parts = data.split(b"\n", 3)
    assert parts[0] == b"P6" and parts[2] == b"255"
    w, h = (int(v) for v in parts[1].split())
    return np.frombuffer(parts[3], dtype=np.uint8, count=w * h * 3).reshape(h, w, 3).copy()


class Ref:
    """The reference renderer itself.  NOT thread-safe and not re-entrant (global
    state, like the reference); use one process per instance."""

    def __init__(self, asset_root=None, mocap=False):
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.drtref_last_error.restype = C.c_char_p
        L.drtref_rng_draws.restype = C.c_ulonglong
        L.drtref_build_scene.argtypes = [C.c_char_p, C.c_float]
        L.drtref_render_unmodified.argtypes = [C.c_char_p, C.c_int]
        L.drtref_render_cloud.argtypes = [C.c_char_p, C.c_float]
        L.drtref_render_loop.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                         C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_double)]
        L.drtref_render_loop_x.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                           C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_double)]
        L.drtref_rng_mode.argtypes = [C.c_int, C.c_uint32, C.c_uint32]
        L.drtref_mocap_bones.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int]
        if asset_root is None and os.path.isdir(REFERENCE_ROOT):
            asset_root = REFERENCE_ROOT
        self._check(L.drtref_init((asset_root or "").encode(), int(mocap)))
        L.drtref_reset_globals()

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError(f"reference harness error {rc}: {self.lib.drtref_last_error().decode()}")
        return rc

    def reset(self):
        self.lib.drtref_reset_globals()

    def build(self, name, frame=0.0):
        self._check(self.lib.drtref_build_scene(name.encode(), float(frame)))

    def settings(self):
        s = abi.Settings()
        self.lib.drtref_get_settings(C.byref(s))
        return s

    def set_settings(self, s):
        self.lib.drtref_set_settings(C.byref(s))

    def export(self):
        n, nl, nt = C.c_int(), C.c_int(), C.c_int()
        self.lib.drtref_scene_counts(C.byref(n), C.byref(nl), C.byref(nt))
        prims = (abi.Prim * max(n.value, 1))()
        lights = (abi.Light * max(nl.value, 1))()
        texs = (abi.Texture * max(nt.value, 1))()
        self._check(self.lib.drtref_export_scene(prims, lights, texs))
        textures = []
        for i in range(nt.value):
            h, w = texs[i].height, texs[i].width
            textures.append(np.ctypeslib.as_array(texs[i].rgb, shape=(h, w, 3)).copy())
        return Scene([abi.copy_struct(prims[i]) for i in range(n.value)],
                     [abi.copy_struct(lights[i]) for i in range(nl.value)], textures)

    def load(self, scene):
        d = scene.desc()
        self._check(self.lib.drtref_load_scene(C.byref(d)))

    def rng(self, mode, seed=0, stream_id=0):
        """mode 0: the reference's own mt19937(random_device); 1: deterministic stream."""
        self.lib.drtref_rng_mode(int(mode), seed, stream_id)

    def render_unmodified(self, frame=0):
        """renderImage exactly as shipped -> uint8 (yRes,xRes,3), PPM row order."""
        with tempfile.TemporaryDirectory() as td:
            p = os.path.join(td, "o.ppm")
            self._check(self.lib.drtref_render_unmodified(p.encode(), int(frame)))
            return read_ppm(p)

    def render_cloud(self, frame=0.0):
        with tempfile.TemporaryDirectory() as td:
            p = os.path.join(td, "o.ppm")
            self._check(self.lib.drtref_render_cloud(p.encode(), float(frame)))
            return read_ppm(p)

    def render_loop(self, frame=0, y0=0, y1=None, reset_policy=1, seed=0, x0=0, x1=None):
        """Pixel-loop restatement over the reference's rayColor.  Returns
        (float32 (rows,xRes,3) in PPM row order (top row first),
         bool (rows,xRes) mask of pixels where the reference itself aborts, seconds)."""
        s = self.settings()
        if y1 is None:
            y1 = s.yRes
        out = np.zeros(((y1 - y0), s.xRes, 3), dtype=np.float32)
        ab = np.zeros(((y1 - y0), s.xRes), dtype=np.uint8)
        sec = C.c_double()
        if x1 is None:
            x1 = s.xRes
        self._check(self.lib.drtref_render_loop_x(int(frame), x0, x1, y0, y1, reset_policy, seed,
                                                  out.ctypes.data_as(C.POINTER(C.c_float)),
                                                  ab.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(sec)))
        return out[::-1].copy(), ab[::-1].astype(bool), sec.value

    def mocap_bones(self, frame, max_bones=64):
        buf = (C.c_double * (6 * max_bones))()
        n = self._check(self.lib.drtref_mocap_bones(int(frame), buf, max_bones))
        return np.array(buf[:6 * n], dtype=np.float64).reshape(n, 2, 3)


def quantize(img_f32):
    """writePPM's float -> unsigned char truncation (helpers.h:178-179)."""
    a = np.asarray(img_f32, dtype=np.float32)
    # (unsigned char)NaN on x86-64 is the low byte of cvttss2si's 0x80000000 = 0
    return np.where(np.isnan(a), np.float32(0), a).astype(np.uint8)


ORACLE_KEYED, ORACLE_STREAM = 0, 1


class Oracle:
    """The CPU restatement (oracle/drt_oracle.cpp) over a POD scene."""

    def __init__(self, scene, builder=0):
        """builder 0: the reference's SAH tree (candidate order included); 1: the median-split stand-in for scenes the
        reference's O(n^2) build cannot handle (conservative gather, one primitive per leaf).  A scene's mesh is
        expanded into individual Triangle primitives on the C++ side."""
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.drt_oracle_last_error.restype = C.c_char_p
        L.drt_oracle_scene_create_ex.argtypes = [C.POINTER(abi.SceneDesc), C.c_int, C.POINTER(C.c_void_p)]
        L.drt_oracle_scene_destroy.argtypes = [C.c_void_p]
        L.drt_oracle_render.argtypes = [C.c_void_p, C.POINTER(abi.Settings), C.POINTER(abi.Tile), C.c_int,
                                        C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(abi.Counters),
                                        C.POINTER(C.c_double)]
        L.drt_oracle_value_noise.restype = C.c_double
        L.drt_oracle_value_noise.argtypes = [C.c_double] * 3
        self.scene = scene
        self._desc = scene.desc()
        self.handle = C.c_void_p()
        rc = L.drt_oracle_scene_create_ex(C.byref(self._desc), int(builder), C.byref(self.handle))
        if rc != 0:
            raise RuntimeError(f"oracle scene_create failed ({rc}): {L.drt_oracle_last_error().decode()}")

    def __del__(self):
        try:
            if self.handle:
                self.lib.drt_oracle_scene_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def render(self, settings, tile=None, mode=ORACLE_KEYED):
        """Returns (float32 (h,w,3) PPM row order in [0,255], bool (h,w) aborted mask,
        abi.Counters, seconds)."""
        if tile is None:
            tile = abi.Tile(0, 0, settings.xRes, settings.yRes, 0)
        out = np.zeros((tile.height, tile.width, 3), dtype=np.float32)
        ab = np.zeros((tile.height, tile.width), dtype=np.uint8)
        cnt = abi.Counters()
        sec = C.c_double()
        rc = self.lib.drt_oracle_render(self.handle, C.byref(settings), C.byref(tile), int(mode),
                                        out.ctypes.data_as(C.POINTER(C.c_float)),
                                        ab.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(cnt), C.byref(sec))
        if rc != 0:
            raise RuntimeError(f"oracle render failed ({rc}): {self.lib.drt_oracle_last_error().decode()}")
        return out, ab.astype(bool), cnt, sec.value

    def candidate_order(self):
        n = len(self.scene.prims)
        buf = (C.c_int * n)()
        self.lib.drt_oracle_candidate_order.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        got = self.lib.drt_oracle_candidate_order(self.handle, buf, n)
        assert got == n
        return list(buf)

    def value_noise(self, x, y, z):
        return self.lib.drt_oracle_value_noise(x, y, z)


def compare(a_f32, b_f32):
    """Parity statistics between two float images in [0,255] after writePPM quantisation."""
    qa, qb = quantize(a_f32).astype(np.int32), quantize(b_f32).astype(np.int32)
    d = np.abs(qa - qb).max(axis=-1)
    return {"frac_within_1": float((d <= 1).mean()), "max": int(d.max()), "mae": float(np.abs(qa - qb).mean()),
            "n_bad": int((d > 1).sum())}


class SkeletonOracle:
    """oracle/drt_skeleton_oracle.cpp: CPU restatement of the reference's ASF/AMC parse + forward kinematics."""

    def __init__(self, asf_bytes, amc_bytes, scale=0.06):
        L = self.lib = C.CDLL(ORACLE_SO)
        L.drt_oracle_skeleton_error.restype = C.c_char_p
        L.drt_oracle_skeleton_create.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_double, C.POINTER(C.c_void_p)]
        L.drt_oracle_skeleton_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.drt_oracle_skeleton_bones.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.drt_oracle_skeleton_structure.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.drt_oracle_skeleton_destroy.argtypes = [C.c_void_p]
        L.drt_oracle_skeleton_destroy.restype = None
        self.handle = C.c_void_p()
        if L.drt_oracle_skeleton_create(asf_bytes, len(asf_bytes), amc_bytes, len(amc_bytes), scale, C.byref(self.handle)) != 0:
            raise RuntimeError("skeleton oracle: " + L.drt_oracle_skeleton_error().decode())
        nc, nf = C.c_int(), C.c_int()
        L.drt_oracle_skeleton_info(self.handle, C.byref(nc), C.byref(nf))
        self.n_cylinders, self.n_frames = nc.value, nf.value

    def structure(self):
        parents = np.full(256, -2, dtype=np.int32); dofs = np.zeros(256, dtype=np.int32)
        n = self.lib.drt_oracle_skeleton_structure(self.handle, parents.ctypes.data_as(C.c_void_p), dofs.ctypes.data_as(C.c_void_p), 256)
        return parents[:n].copy(), dofs[:n].copy()

    def bones(self, frame):
        out = np.empty((self.n_cylinders, 2, 3), dtype=np.float64)
        if self.lib.drt_oracle_skeleton_bones(self.handle, int(frame), out.ctypes.data) < 0:
            raise RuntimeError("skeleton oracle: " + self.lib.drt_oracle_skeleton_error().decode())
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.drt_oracle_skeleton_destroy(self.handle)
            self.handle = None

    __del__ = close
