"""OBJ ingest: loadObj-equivalent parsing + the scene builders' UV handling."""
import numpy as np
import pytest

from distraytracer_b200 import ingest, scenes

CUBE_FACE = """
# one quad with texcoords above 1 and one triangle addressed with negative indices
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vt 2.25 0
vt 2.25 0.5
vt 0 0.5
vn 0 0 1
f 1/1/1 2/2/1 3/3/1 4/4/1
v 0 0 1
f -1/1 -4/2 -3/3
"""


def test_parse_obj_matches_loadobj_conventions():
    o = ingest.parse_obj(CUBE_FACE)
    assert o["vertices"].shape == (5, 3) and o["texcoords"].shape == (4, 2)
    # quad -> fan of two triangles, 0-based indices (objHelper.h:63-72)
    assert o["v_indices"].tolist() == [[0, 1, 2], [0, 2, 3], [4, 1, 2]]
    assert o["t_indices"].tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]


def test_mesh_from_obj_wraps_flips_and_unifies():
    o = ingest.parse_obj(CUBE_FACE)
    m = ingest.mesh_from_obj(o, material=None, transform=[[3, 0, 0, 3], [0, 3, 0, -1], [0, 0, 3, 5], [0, 0, 0, 1]])
    V, T, UV = m["vertices"], m["indices"], m["texcoords"]
    assert V.shape[0] == UV.shape[0] == 5 and T.shape == (3, 3)
    # corner positions / UVs of every triangle survive the unification
    want_pos = (np.concatenate([o["vertices"], np.ones((5, 1), np.float32)], axis=1) @ np.array(
        [[3, 0, 0, 3], [0, 3, 0, -1], [0, 0, 3, 5], [0, 0, 0, 1]], dtype=np.float64).T)[:, :3]
    for t in range(3):
        for k in range(3):
            assert np.allclose(V[T[t, k]], want_pos[o["v_indices"][t, k]])
    # 2.25 -> 0.25 (scene.h:335-340), v -> 1 - v (scene.h:357-359)
    uv_of = {tuple(np.round(V[i], 5)): tuple(np.round(UV[i], 5)) for i in range(5) if i in T[0] or i in T[1]}
    assert uv_of[(6.0, -1.0, 5.0)] == (0.25, 1.0)
    assert uv_of[(6.0, 2.0, 5.0)] == (0.25, 0.5)


def test_mesh_from_obj_rejects_what_the_reference_throws_on():
    o = ingest.parse_obj("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt -0.5 0\nvt 1 0\nvt 0 1\nf 1/1 2/2 3/3\n")
    with pytest.raises(ingest.ObjError):
        ingest.mesh_from_obj(o, material=None)


def test_obj_round_trip_of_the_terrain_mesh():
    mesh = scenes.terrain_mesh(n=9)
    back = ingest.mesh_from_obj(ingest.parse_obj(ingest.mesh_to_obj(mesh)), material=None, wrap_uv=False, flip_v=False)
    # same triangles, corner by corner (vertex order may differ after unification)
    for key in ("vertices", "texcoords"):
        assert np.array_equal(mesh[key][mesh["indices"]], back[key][back["indices"]])


def test_parse_obj_errors():
    with pytest.raises(ingest.ObjError):
        ingest.parse_obj("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(ingest.ObjError):
        ingest.parse_obj("v 0 0\n")


def test_roughness_map_gives_a_material_table_like_the_reference_builders():
    """scene.h:372-378: each Triangle gets roughness (r1 + r2 + r3) / (3 * 255), r_k the roughness-map byte at
    int(u * int(W-1) + v * int(H-1) * (W-1)) for the ORIGINAL texcoords of its corners.  The mesh carries one material per
    distinct value (at most 766) and a per-triangle index."""
    from distraytracer_b200 import abi
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(13, 17), dtype=np.uint8)              # a grey JPEG decodes to one channel
    obj = ingest.parse_obj(CUBE_FACE.replace("vt 2.25 0\n", "vt 0.75 0\n").replace("vt 2.25 0.5", "vt 0.75 0.5"))
    fr = ingest.face_roughness_from_map(obj, img)
    flat, w, h = img.reshape(-1), 17, 13
    for t, tri in enumerate(obj["t_indices"]):
        r = [np.float32(flat[int(float(obj["texcoords"][k][0]) * int(w - 1) + float(obj["texcoords"][k][1]) * int(h - 1) * (w - 1))]) for k in tri]
        assert fr[t] == np.float32((r[0] + r[1] + r[2]) / np.float32(3 * 255))
    mat = scenes.mesh_material(model=abi.MODEL_OREN_NAYAR)
    mesh = ingest.mesh_from_obj(obj, mat, face_roughness=fr)
    assert len(mesh["materials"]) == len(np.unique(fr)) <= 766
    got = np.array([mesh["materials"][i].roughness for i in mesh["material_ids"]], dtype=np.float32)
    assert np.array_equal(got, fr)
    assert all(m.model == abi.MODEL_OREN_NAYAR and m.type == mat.type for m in mesh["materials"])
    with pytest.raises(ingest.ObjError):
        ingest.face_roughness_from_map(ingest.parse_obj(CUBE_FACE), img[:1, :2])   # texcoord 2.25 indexes past a 1x2 map (the reference reads out of bounds)
    # the flat list of Triangle primitives the oracle consumes carries the same per-face roughness
    prims = scenes.mesh_to_prims(mesh)
    assert np.array_equal(np.array([p.roughness for p in prims], dtype=np.float32), fr)
