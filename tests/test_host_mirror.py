"""The C++ host-side mirror of the reference's scene surface (distraytracer_b200/host/drt_host.h):
scenes built through the mirrored constructors flatten to exactly the PODs the reference's own
builders produce (tests/golden fixtures exported from the compiled reference), and
renderImage() drives the GPU path end to end."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_case


@pytest.fixture(scope="module")
def host_bin(tmp_path_factory):
    from distraytracer_b200 import runtime
    if not os.path.exists(runtime.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    out = str(tmp_path_factory.mktemp("host") / "host_scene")
    libdir = os.path.dirname(runtime.LIB_PATH)
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "host", "host_scene.cpp"), "-o", out,
                           "-L" + libdir, "-ldrt", "-Wl,-rpath," + libdir, "-lpthread"])
    return out


@pytest.mark.parametrize("scene,case", [("hw4", "hw4"), ("reflectance", "reflectance")])
def test_flattened_scene_equals_reference_export(host_bin, tmp_path, scene, case):
    from distraytracer_b200 import abi
    out = str(tmp_path / "dump.bin")
    subprocess.check_call([host_bin, "dump", scene, out])
    raw = open(out, "rb").read()
    n_prims, n_lights = np.frombuffer(raw[:8], dtype=np.int32)
    off = 8
    prims = [abi.Prim.from_buffer_copy(raw[off + i * C.sizeof(abi.Prim): off + (i + 1) * C.sizeof(abi.Prim)]) for i in range(n_prims)]
    off += n_prims * C.sizeof(abi.Prim)
    lights = [abi.Light.from_buffer_copy(raw[off + i * C.sizeof(abi.Light): off + (i + 1) * C.sizeof(abi.Light)]) for i in range(n_lights)]
    off += n_lights * C.sizeof(abi.Light)
    st = abi.Settings.from_buffer_copy(raw[off: off + C.sizeof(abi.Settings)])
    ref_scene, ref_settings, _ = load_case(case)
    assert n_prims == len(ref_scene.prims) and n_lights == len(ref_scene.lights)
    for a, b in zip(prims, ref_scene.prims):
        assert bytes(a) == bytes(b)
    for a, b in zip(lights, ref_scene.lights):
        assert bytes(a) == bytes(b)
    assert bytes(st) == bytes(ref_settings)


@pytest.mark.gpu
def test_renderImage_through_host_mirror(host_bin, tmp_path, oracle_lib):
    from oracle.harness import Oracle, ORACLE_KEYED, quantize, read_ppm
    out = str(tmp_path / "frame.ppm")
    subprocess.check_call([host_bin, "render", "reflectance", out])
    got = read_ppm(out)
    scene, settings, _ = load_case("reflectance")
    want, _, _, _ = Oracle(scene).render(settings, mode=ORACLE_KEYED)
    d = np.abs(got.astype(int) - quantize(want).astype(int)).max(axis=-1)
    assert (d <= 1).mean() >= 0.999


def test_host_obj_ingest_matches_python_ingest(host_bin, tmp_path):
    """loadObj + setMesh of the C++ mirror and distraytracer_b200.ingest agree triangle by triangle."""
    from distraytracer_b200 import ingest
    from test_ingest import CUBE_FACE
    obj_path = str(tmp_path / "m.obj")
    open(obj_path, "w").write(CUBE_FACE)
    out = str(tmp_path / "mesh.bin")
    subprocess.check_call([host_bin, "mesh", obj_path, out])
    raw = open(out, "rb").read()
    nv, nt, has_uv = np.frombuffer(raw[:24], dtype=np.int64)
    off = 24
    V = np.frombuffer(raw[off: off + 12 * nv], dtype=np.float32).reshape(-1, 3); off += 12 * nv
    T = np.frombuffer(raw[off: off + 12 * nt], dtype=np.int32).reshape(-1, 3); off += 12 * nt
    UV = np.frombuffer(raw[off: off + 8 * nv], dtype=np.float32).reshape(-1, 2)
    want = ingest.mesh_from_obj(ingest.parse_obj(CUBE_FACE), None, transform=[[3, 0, 0, 3], [0, 3, 0, -1], [0, 0, 3, 5], [0, 0, 0, 1]])
    assert has_uv == 1 and nt == len(want["indices"]) and nv == len(want["vertices"])
    assert np.array_equal(V[T], want["vertices"][want["indices"]])
    assert np.array_equal(UV[T], want["texcoords"][want["indices"]])


@pytest.mark.gpu
def test_renderVideo_writes_the_frames_renderImage_would(host_bin, tmp_path):
    """renderVideo (frames round-robin over GPUs, one resident scene per GPU updated in place, PPMs written by a
    background thread) produces, frame for frame, the file renderImage writes for that frame."""
    from oracle.harness import read_ppm
    prefix = str(tmp_path / "vid")
    subprocess.check_call([host_bin, "video", "reflectance", prefix])
    single = str(tmp_path / "single.ppm")
    subprocess.check_call([host_bin, "render", "reflectance", single])          # renders frame 40
    assert np.array_equal(read_ppm(prefix + ".0040.ppm"), read_ppm(single))
    frames = [read_ppm(prefix + ".%04d.ppm" % f) for f in range(40, 44)]
    assert all(f.shape == (120, 160, 3) for f in frames)
    assert any(not np.array_equal(frames[0], f) for f in frames[1:])             # the light moves between frames


@pytest.mark.gpu
def test_renderFrame_in_row_blocks_is_the_same_picture(host_bin, tmp_path):
    """The multi-GPU cut of a single frame (row blocks claimed dynamically) does not change the image."""
    from oracle.harness import read_ppm
    a, b = str(tmp_path / "whole.ppm"), str(tmp_path / "blocks.ppm")
    subprocess.check_call([host_bin, "render", "reflectance", a])
    subprocess.check_call([host_bin, "render", "reflectance", b], env=dict(os.environ, DRT_HOST_BLOCKS="7"))
    assert np.array_equal(read_ppm(a), read_ppm(b))


@pytest.mark.gpu
def test_mocap_through_host_mirror(host_bin, tmp_path):
    """ASF/AMC -> device forward kinematics -> Cylinder constructors of the C++ mirror: the flattened bone cylinders and
    floor are byte for byte what the reference's buildSceneChkpt2 exported (fixture chkpt2_mocap, frame 30), and the
    mocap video loop writes the frames renderImage would."""
    from distraytracer_b200 import abi
    from oracle.harness import read_ppm
    from conftest import GOLDEN
    clip = [os.path.join(GOLDEN, "mocap_90.asf"), os.path.join(GOLDEN, "mocap_90_16_frames880_1000.amc")]
    out = str(tmp_path / "dump.bin")
    subprocess.check_call([host_bin, "dump", "mocap", out] + clip)
    raw = open(out, "rb").read()
    n_prims, n_lights = np.frombuffer(raw[:8], dtype=np.int32)
    sz = C.sizeof(abi.Prim)
    prims = [raw[8 + i * sz: 8 + (i + 1) * sz] for i in range(n_prims)]
    ref_scene, _, _ = load_case("chkpt2_mocap")
    assert n_prims == 31 and len(ref_scene.prims) == 33        # the fixture also holds the two sphere lights
    for a, b in zip(prims, ref_scene.prims[:31]):
        assert a == bytes(b)
    prefix = str(tmp_path / "vid")
    subprocess.check_call([host_bin, "video", "mocap", prefix] + clip)
    single = str(tmp_path / "single.ppm")
    subprocess.check_call([host_bin, "render", "mocap", single] + clip)
    frames = [read_ppm(prefix + ".%04d.ppm" % f) for f in range(30, 34)]
    assert np.array_equal(frames[0], read_ppm(single))
    assert any(not np.array_equal(frames[0], f) for f in frames[1:])             # the figure moves


def test_host_roughness_map_and_loadTexture_match_python_ingest(host_bin, tmp_path):
    """The per-face roughness of the reference's model builders (scene.h:372-378) through the C++ mirror -- loadTexture of a
    binary PPM, faceRoughnessFromMap, setMesh building the material table -- equals distraytracer_b200.ingest's."""
    from distraytracer_b200 import ingest, runtime
    rng = np.random.default_rng(5)
    n = 9
    lines = ["v %g %g %g" % (i, j, 0.1 * i * j) for i in range(n) for j in range(n)]
    lines += ["vt %.6f %.6f" % (i / (n - 1), j / (n - 1)) for i in range(n) for j in range(n)]
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j + 1, (i + 1) * n + j + 1, (i + 1) * n + j + 2, i * n + j + 2
            lines += ["f %d/%d %d/%d %d/%d" % (a, a, b, b, c, c), "f %d/%d %d/%d %d/%d" % (a, a, c, c, d, d)]
    text = "\n".join(lines) + "\n"
    obj_path, ppm, out = str(tmp_path / "g.obj"), str(tmp_path / "rough.ppm"), str(tmp_path / "mesh.bin")
    open(obj_path, "w").write(text)
    img = rng.integers(0, 256, size=(7, 11, 3), dtype=np.uint8)
    runtime.write_ppm(ppm, img)
    subprocess.check_call([host_bin, "mesh", obj_path, out, ppm])
    raw = open(out, "rb").read()
    nv, nt, has_uv = np.frombuffer(raw[:24], dtype=np.int64)
    off = 24 + 12 * nv + 12 * nt + 8 * nv
    nm = int(np.frombuffer(raw[off: off + 8], dtype=np.int64)[0]); off += 8
    ids = np.frombuffer(raw[off: off + 4 * nt], dtype=np.int32); off += 4 * nt
    rough = np.frombuffer(raw[off: off + 8 * nm], dtype=np.float64)
    obj = ingest.parse_obj(text)
    want = ingest.face_roughness_from_map(obj, img)
    assert nt == len(want) and 1 < nm <= 766
    assert np.array_equal(rough[ids].astype(np.float32), want)
