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
