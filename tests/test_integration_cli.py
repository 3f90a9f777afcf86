"""integration/render_drt: the reference's own `main` and scene builders (compiled unmodified, where they lie) with every
frame rendered by libdrt.so.  The binary is built in the build container (it needs the reference tree) and
travels to the GPU box as a built artefact, like the libraries."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_case

BIN = os.path.join(ROOT, "integration", "_build", "render_drt")


def _binary():
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "integration")])
    if not os.path.exists(BIN):
        pytest.skip("integration/_build/render_drt was not built (no reference tree here)")
    return BIN


def _run_dir(tmp_path):
    """What the reference's main() opens relative to its working directory: the mocap clip (render_final_project.cpp:1388-1399)
    and ./textures (scene.h:885-905); output directories are the caller's business, as with the reference."""
    os.makedirs(tmp_path / "textures")
    for f in os.listdir(os.path.join(GOLDEN, "textures")):
        shutil.copyfile(os.path.join(GOLDEN, "textures", f), tmp_path / "textures" / f)
    shutil.copyfile(os.path.join(GOLDEN, "mocap_90.asf"), tmp_path / "90.asf")
    shutil.copyfile(os.path.join(GOLDEN, "mocap_90_16_frames880_1000.amc"), tmp_path / "90_16_v3.amc")
    for d in ("checkertexture", "prismcyl", "reflectance"):
        os.makedirs(tmp_path / "test_frames" / d)
    return str(tmp_path)


def test_cli_is_the_reference_main_routed_to_the_cuda_back_end(tmp_path):
    """No GPU needed: the binary carries the reference's mode table, its calls of renderImage land in the replacement,
    and without a device it fails loudly instead of falling back to the reference's CPU loop it also contains."""
    exe = _binary()
    syms = subprocess.run(["nm", "-C", exe], capture_output=True, text=True, check=True).stdout
    assert "renderImageCloud(char const*, int)" in syms and "drt_reference_main" in syms
    from distraytracer_b200 import runtime
    if runtime.device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu test")
    cwd = _run_dir(tmp_path)
    r = subprocess.run([exe, "test", "checkertexture"], cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr
    assert not os.listdir(os.path.join(cwd, "test_frames", "checkertexture"))


@pytest.mark.gpu
def test_cli_frame_equals_drt_render_of_the_same_scene(oracle_lib, tmp_path):
    """`./render_drt test checkertexture` (render_final_project.cpp:1840-1853: 640x480, 1 spp, aperture 0) writes the PPM
    drt_render produces for the scene the reference's builder exported (tests/golden/checkertexture.npz), byte for byte;
    `./render_drt prismcyl 7` (:1711-1723) likewise."""
    from distraytracer_b200 import abi, runtime
    from oracle.harness import read_ppm
    exe = _binary()
    cwd = _run_dir(tmp_path)
    for argv, out, case, frame in ((["test", "checkertexture"], "test_frames/checkertexture/frame.0000.ppm", "checkertexture", 0),
                                   (["prismcyl", "7"], "test_frames/prismcyl/frame.0007.ppm", "prismcyl", 7)):
        r = subprocess.run([exe] + argv, cwd=cwd, capture_output=True, text=True, env={**os.environ, "DRT_DEVICES": "1"})
        assert r.returncode == 0, r.stderr[-2000:]
        got = read_ppm(os.path.join(cwd, out))
        scene, settings, _ = load_case(case)
        s = abi.copy_struct(settings)
        s.xRes, s.yRes, s.frame, s.seed = 640, 480, frame, 0
        if case == "checkertexture":
            s.antialias_samples, s.aperture = 1, 0.0
        else:
            d = runtime.default_settings()                   # the prismcyl mode keeps the globals' defaults
            s.antialias_samples, s.aperture = d.antialias_samples, d.aperture
        want = runtime.DeviceScene(scene, 0).render(s)
        assert got.shape == want.shape == (480, 640, 3)
        assert np.array_equal(got, want), (case, int((got != want).any(-1).sum()))
        assert case == "prismcyl" or want.std() > 5
