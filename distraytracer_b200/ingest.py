"""Mesh ingest: Wavefront OBJ -> the `drt_mesh` arrays.

Mirrors what the reference does on the way from an .obj file to `Triangle` primitives:

* `parse_obj`  = loadObj (objHelper.h:6-85): positions, texcoords and the per-face position /
  texcoord index triples.  The reference reads the file through tiny_obj_loader with its default
  `triangulate = true`, so polygons arrive as triangle fans; indices are 0-based, a missing
  texcoord index is -1.
* `mesh_from_obj` = the per-triangle work of the scene builders (scene.h:296-386): a 4x4 object
  transform on the positions, texture coordinates above 1 wrapped by dropping the integer part,
  the bounds check that makes the reference `throw`, and the V flip (`uv[1] = 1 - uv[1]`).
  `drt_mesh` carries one texcoord per VERTEX, so (position, texcoord) index pairs are unified
  into vertices here -- the triangles, their corner positions and corner UVs are unchanged.

Pure numpy / host side: nothing here touches the device.
"""
import numpy as np


class ObjError(ValueError):
    pass


def parse_obj(text):
    """Returns dict(vertices float32 (nv,3), texcoords float32 (nt,2), v_indices int32 (T,3),
    t_indices int32 (T,3)).  `text` is the file content (str) or a path-like to read."""
    if not isinstance(text, str) or ("\n" not in text and text.lower().endswith(".obj")):
        with open(text, "r") as f:
            text = f.read()
    verts, tcs, vi, ti = [], [], [], []
    for ln, raw in enumerate(text.splitlines(), 1):
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        tok = line.split()
        key = tok[0]
        if key == "v":
            if len(tok) < 4:
                raise ObjError(f"line {ln}: vertex needs 3 coordinates")
            verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
        elif key == "vt":
            if len(tok) < 2:
                raise ObjError(f"line {ln}: texcoord needs at least 1 coordinate")
            tcs.append([float(tok[1]), float(tok[2]) if len(tok) > 2 else 0.0])
        elif key == "f":
            corners = []
            for c in tok[1:]:
                parts = c.split("/")
                v = int(parts[0])
                t = int(parts[1]) if len(parts) > 1 and parts[1] != "" else None
                # OBJ indices are 1-based; negative ones count back from the current end
                v = v - 1 if v > 0 else len(verts) + v
                if t is not None:
                    t = t - 1 if t > 0 else len(tcs) + t
                if not (0 <= v < len(verts)) or (t is not None and not (0 <= t < len(tcs))):
                    raise ObjError(f"line {ln}: index out of range in face corner '{c}'")
                corners.append((v, -1 if t is None else t))
            if len(corners) < 3:
                raise ObjError(f"line {ln}: face with fewer than 3 corners")
            for k in range(1, len(corners) - 1):        # triangle fan, as tiny_obj_loader emits for convex polygons
                tri = (corners[0], corners[k], corners[k + 1])
                vi.append([c[0] for c in tri])
                ti.append([c[1] for c in tri])
        # vn, g, o, s, usemtl, mtllib: not used by the path (flat normals, one material per mesh)
    return {
        "vertices": np.asarray(verts, dtype=np.float32).reshape(-1, 3),
        "texcoords": np.asarray(tcs, dtype=np.float32).reshape(-1, 2),
        "v_indices": np.asarray(vi, dtype=np.int32).reshape(-1, 3),
        "t_indices": np.asarray(ti, dtype=np.int32).reshape(-1, 3),
    }


def face_roughness_from_map(obj, roughness_map):
    """Reflectance::roughness per face as the reference's model builders assign it (scene.h:372-378): the roughness map
    (uint8 array as stbi_load returns it, (H, W) or (H, W, C)) is indexed with the ORIGINAL texcoords of the three corners as
    `map[int(u * int(W-1) + v * int(H-1) * (W-1))]` -- a byte offset into the decoded buffer, channel count ignored, exactly
    as written there -- and the face gets float32 (r1 + r2 + r3) / (3 * 255): at most 766 distinct values."""
    img = np.ascontiguousarray(roughness_map, dtype=np.uint8)
    h, w = img.shape[0], img.shape[1]
    flat = img.reshape(-1)
    uv = obj["texcoords"].astype(np.float64)[obj["t_indices"]]                  # (T, 3, 2)
    at = (uv[..., 0] * int(w - 1) + uv[..., 1] * int(h - 1) * (w - 1)).astype(np.int64)   # C++ int(): truncation
    if (obj["t_indices"] < 0).any() or (at < 0).any() or (at >= flat.size).any():
        raise ObjError("roughness map lookup outside the image")               # the reference reads out of bounds there
    r = flat[at].astype(np.float32)
    return ((r[:, 0] + r[:, 1]) + r[:, 2]) / np.float32(3 * 255)


def mesh_from_obj(obj, material, transform=None, wrap_uv=True, flip_v=True, face_roughness=None):
    """OBJ arrays -> dict(vertices, indices, texcoords, material) for `Scene(mesh=...)`.

    face_roughness : optional float32 (T,), one Reflectance::roughness per face (face_roughness_from_map); the mesh then
                carries a material table with one entry per distinct value (`materials`, `material_ids`).

    transform : optional 4x4 applied to the positions as `(M * (v,1)).head<3>()` (scene.h:301-307).
    wrap_uv   : `if (uv > 1) uv -= int(uv)` per component (scene.h:335-340).
    flip_v    : `uv[1] = 1 - uv[1]` (scene.h:357-359).
    Raises ObjError where the reference prints "Texcoords out of bounds" and throws (scene.h:343-354)."""
    V = obj["vertices"].astype(np.float64)
    if transform is not None:
        M = np.asarray(transform, dtype=np.float64).reshape(4, 4)
        V = (np.concatenate([V, np.ones((len(V), 1))], axis=1) @ M.T)[:, :3]
    vi, ti = obj["v_indices"], obj["t_indices"]
    has_uv = len(obj["texcoords"]) > 0 and (ti >= 0).all()
    extra = {}
    if face_roughness is not None:
        from . import abi
        fr = np.asarray(face_roughness, dtype=np.float32)
        if fr.shape != (len(vi),):
            raise ObjError("face_roughness needs one value per face")
        values, ids = np.unique(fr, return_inverse=True)
        mats = []
        for v in values:
            m = abi.copy_struct(material)
            m.roughness = float(v)
            mats.append(m)
        extra = {"materials": mats, "material_ids": ids.astype(np.int32)}
    if not has_uv:
        return {"vertices": V.astype(np.float32), "indices": vi.astype(np.int32), "texcoords": None, "material": material, **extra}
    UV = obj["texcoords"].astype(np.float64).copy()
    if wrap_uv:
        over = UV > 1
        UV[over] = UV[over] - np.trunc(UV[over])
    used = np.unique(ti)
    # the reference's test: u >= 0 and v <= 1 for the three corners (scene.h:343-345)
    if not ((UV[used, 0] >= 0).all() and (UV[used, 1] <= 1).all()):
        raise ObjError("Texcoords out of bounds")
    if flip_v:
        UV[:, 1] = 1.0 - UV[:, 1]
    # unify (position index, texcoord index) pairs into vertices
    pairs = np.stack([vi.reshape(-1), ti.reshape(-1)], axis=1)
    uniq, inverse = np.unique(pairs, axis=0, return_inverse=True)
    return {
        "vertices": V[uniq[:, 0]].astype(np.float32),
        "indices": inverse.reshape(-1, 3).astype(np.int32),
        "texcoords": UV[uniq[:, 1]].astype(np.float32),
        "material": material, **extra,
    }


def mesh_to_obj(mesh):
    """The inverse, for fixtures and round-trip tests: one `v` / `vt` per vertex, `f a/a b/b c/c`."""
    out = []
    for v in mesh["vertices"]:
        out.append("v %.9g %.9g %.9g" % (float(v[0]), float(v[1]), float(v[2])))
    tc = mesh.get("texcoords")
    if tc is not None:
        for t in tc:
            out.append("vt %.9g %.9g" % (float(t[0]), float(t[1])))
    for a, b, c in mesh["indices"]:
        if tc is not None:
            out.append("f %d/%d %d/%d %d/%d" % (a + 1, a + 1, b + 1, b + 1, c + 1, c + 1))
        else:
            out.append("f %d %d %d" % (a + 1, b + 1, c + 1))
    return "\n".join(out) + "\n"
