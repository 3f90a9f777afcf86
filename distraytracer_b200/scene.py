"""Host-side scene container: the POD form of the reference's `shapes`, `lights`,
`texture_frames`/`texture_dims` globals (render_final_project.cpp:70-71, 96-97)
as they cross the C ABI (include/drt.h).  Pure data; no rendering here.
"""
import ctypes as C
import io
import numpy as np

from . import abi


class Scene:
    """prims: list[abi.Prim]; lights: list[abi.Light]; textures: list[np.uint8 (H,W,3)];
    mesh: optional dict(vertices float32 (V,3), indices int32 (T,3), texcoords float32 (V,2)|None,
    material abi.Prim [, materials list[abi.Prim], material_ids int32 (T,)])."""

    def __init__(self, prims=None, lights=None, textures=None, mesh=None):
        self.prims = list(prims or [])
        self.lights = list(lights or [])
        self.textures = [np.ascontiguousarray(t, dtype=np.uint8) for t in (textures or [])]
        self.mesh = mesh
        self._keep = None

    # ---- C view --------------------------------------------------------------
    def desc(self):
        """Build a drt_scene_desc whose pointers stay valid while `self` lives."""
        n, nl, nt = len(self.prims), len(self.lights), len(self.textures)
        prims = (abi.Prim * max(n, 1))(*self.prims)
        lights = (abi.Light * max(nl, 1))(*self.lights)
        texs = (abi.Texture * max(nt, 1))()
        for i, t in enumerate(self.textures):
            assert t.ndim == 3 and t.shape[2] == 3
            texs[i].height, texs[i].width = t.shape[0], t.shape[1]
            texs[i].rgb = t.ctypes.data_as(C.POINTER(C.c_uint8))
        d = abi.SceneDesc()
        d.abi_version = abi.ABI_VERSION
        d.n_prims, d.prims = n, prims
        d.n_lights, d.lights = nl, lights
        d.n_textures, d.textures = nt, texs
        keep = [prims, lights, texs]
        if self.mesh is not None:
            m = abi.Mesh()
            v = np.ascontiguousarray(self.mesh["vertices"], dtype=np.float32)
            idx = np.ascontiguousarray(self.mesh["indices"], dtype=np.int32)
            m.n_vertices, m.n_triangles = v.shape[0], idx.shape[0]
            m.vertices = v.ctypes.data_as(C.POINTER(C.c_float))
            m.indices = idx.ctypes.data_as(C.POINTER(C.c_int32))
            tc = self.mesh.get("texcoords")
            if tc is not None:
                tc = np.ascontiguousarray(tc, dtype=np.float32)
                m.texcoords = tc.ctypes.data_as(C.POINTER(C.c_float))
            m.material = self.mesh["material"]
            mats = self.mesh.get("materials")
            mids = None
            if mats:
                # per-triangle materials: triangle t uses materials[material_ids[t]]
                marr = (abi.Prim * len(mats))(*mats)
                mids = np.ascontiguousarray(self.mesh["material_ids"], dtype=np.int32)
                assert mids.shape == (idx.shape[0],)
                m.n_materials, m.materials = len(mats), marr
                m.material_ids = mids.ctypes.data_as(C.POINTER(C.c_int32))
                keep.append(marr)
            d.mesh = C.pointer(m)
            keep += [m, v, idx, tc, mids]
        self._keep = keep
        return d

    # ---- fixtures ------------------------------------------------------------
    def to_npz_dict(self):
        out = {
            "abi_version": np.int32(abi.ABI_VERSION),
            "prims": _structs_to_bytes(self.prims, abi.Prim),
            "lights": _structs_to_bytes(self.lights, abi.Light),
            "n_textures": np.int32(len(self.textures)),
        }
        for i, t in enumerate(self.textures):
            out[f"tex{i}"] = t
        return out

    @staticmethod
    def from_npz_dict(z):
        version = int(z["abi_version"])
        assert version in (1, abi.ABI_VERSION), "fixture written for another ABI version"
        # version 1 records are a prefix of version 2's (the holes of the slab-box prisms were appended)
        prims = _bytes_to_structs(z["prims"], abi.Prim, abi.PRIM_BYTES_V1 if version == 1 else None)
        lights = _bytes_to_structs(z["lights"], abi.Light)
        texs = [z[f"tex{i}"] for i in range(int(z["n_textures"]))]
        return Scene(prims, lights, texs)


def _structs_to_bytes(items, typ):
    buf = io.BytesIO()
    for it in items:
        buf.write(bytes(it))
    return np.frombuffer(buf.getvalue(), dtype=np.uint8).copy()


def _bytes_to_structs(arr, typ, record_bytes=None):
    raw = np.asarray(arr, dtype=np.uint8).tobytes()
    sz = C.sizeof(typ)
    rec = record_bytes or sz
    assert len(raw) % rec == 0 and rec <= sz
    pad = bytes(sz - rec)
    return [typ.from_buffer_copy(raw[i * rec:(i + 1) * rec] + pad) for i in range(len(raw) // rec)]


def settings_to_bytes(s):
    return np.frombuffer(bytes(s), dtype=np.uint8).copy()


def settings_from_bytes(arr):
    return abi.Settings.from_buffer_copy(np.asarray(arr, dtype=np.uint8).tobytes())


def save_fixture(path, scene, settings, **arrays):
    d = scene.to_npz_dict()
    d["settings"] = settings_to_bytes(settings)
    d.update(arrays)
    np.savez_compressed(path, **d)


def load_fixture(path):
    z = np.load(path)
    scene = Scene.from_npz_dict(z)
    settings = settings_from_bytes(z["settings"])
    extra = {k: z[k] for k in z.files
             if k not in ("abi_version", "prims", "lights", "n_textures", "settings") and not k.startswith("tex")}
    return scene, settings, extra
