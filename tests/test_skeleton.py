"""ASF/AMC mocap ingest + forward kinematics (BASELINE config 4).

CPU: the oracle's restatement (oracle/drt_skeleton_oracle.cpp) is pinned bit-for-bit on the bone end points
the compiled reference produced (tests/golden/mocap_bones_0_119.npy, written by make_golden.py through
drtref_mocap_bones) and, where the reference tree is present, on frames over the whole clip.
GPU: the device table of drt_skeleton_create (kernel skeleton_fk) is bit-identical to the oracle, and a scene
re-posed from it renders the same image as one given the reference's bones."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

ASF = os.path.join(GOLDEN, "mocap_90.asf")
AMC = os.path.join(GOLDEN, "mocap_90_16_first121.amc")


def _clip():
    with open(ASF, "rb") as f:
        asf = f.read()
    with open(AMC, "rb") as f:
        amc = f.read()
    return asf, amc


def test_oracle_fk_is_pinned_on_the_reference_bones(oracle_lib):
    from oracle.harness import SkeletonOracle
    sk = SkeletonOracle(*_clip())
    assert (sk.n_cylinders, sk.n_frames) == (30, 121)
    gold = np.load(os.path.join(GOLDEN, "mocap_bones_0_119.npy"))
    got = np.stack([sk.bones(f) for f in range(120)])
    assert np.array_equal(got, gold)                      # bit-exact, all 120 x 30 x 6 doubles
    # past the clip the reference clamps to the last frame (scene.h:117-121); negative frames are fatal (:111-115)
    assert np.array_equal(sk.bones(5000), sk.bones(120))
    with pytest.raises(RuntimeError):
        sk.bones(-1)


def test_oracle_fk_matches_compiled_reference_over_the_whole_clip(oracle_lib):
    from oracle.harness import SkeletonOracle, Ref, ref_available, REFERENCE_ROOT
    if not (ref_available() and os.path.exists(os.path.join(REFERENCE_ROOT, "90_16_v3.amc"))):
        pytest.skip("reference tree not present (GPU box): pinned by the golden bones instead")
    with open(os.path.join(REFERENCE_ROOT, "90.asf"), "rb") as f:
        asf = f.read()
    with open(os.path.join(REFERENCE_ROOT, "90_16_v3.amc"), "rb") as f:
        amc = f.read()
    sk = SkeletonOracle(asf, amc)
    assert sk.n_frames == 2964
    ref = Ref(mocap=True)
    for frame in list(range(0, 2964, 97)) + [2963, 4000]:
        assert np.array_equal(sk.bones(frame), ref.mocap_bones(frame)), frame


def test_skeleton_needs_a_device():
    """No CPU fallback: without a GPU the ingest fails with DRT_ERR_NO_DEVICE."""
    from distraytracer_b200 import runtime, abi
    if runtime.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(runtime.DrtError) as e:
        runtime.DeviceSkeleton(*_clip())
    assert e.value.code == abi.ERR_NO_DEVICE


@pytest.mark.gpu
def test_cuda_fk_table_is_bit_identical_to_the_oracle(oracle_lib):
    from distraytracer_b200 import runtime
    from oracle.harness import SkeletonOracle
    asf, amc = _clip()
    dev = runtime.DeviceSkeleton(asf, amc)
    orc = SkeletonOracle(asf, amc)
    assert (dev.n_cylinders, dev.n_frames) == (orc.n_cylinders, orc.n_frames) == (30, 121)
    got = dev.bones()
    want = np.stack([orc.bones(f) for f in range(orc.n_frames)])
    assert np.array_equal(got, want)
    assert np.array_equal(got[:120], np.load(os.path.join(GOLDEN, "mocap_bones_0_119.npy")))
    # file-path entry point, and a sub-range read
    dev2 = runtime.DeviceSkeleton(ASF, AMC)
    assert np.array_equal(dev2.bones(7, 3), want[7:10])
    assert dev.fk_ms > 0
    # a long clip (the 121 frames repeated to ~3000, renumbered): every frame block equals the first
    lines = amc.split(b"\n")
    head, body = lines[:3], lines[3:3 + 121 * 30]
    long_amc = b"\n".join(head + body * 25) + b"\n"
    big = runtime.DeviceSkeleton(asf, long_amc)
    assert big.n_frames == 121 * 25
    tab = big.bones()
    assert np.array_equal(tab.reshape(25, 121, 30, 2, 3), np.broadcast_to(want, (25,) + want.shape))


@pytest.mark.gpu
def test_cuda_posed_scene_renders_like_the_reference_bones(oracle_lib):
    from distraytracer_b200 import runtime, scenes, abi
    asf, amc = _clip()
    skel = runtime.DeviceSkeleton(asf, amc)
    for frame in (0, 57, 118):      # the golden table ends at 119, whose velocity needs frame 120
        scene, st = scenes.config4_frame(frame, 160, 90, 4)       # bones from the reference's golden table
        first = next(i for i, p in enumerate(scene.prims) if p.type == abi.PRIM_CYLINDER)
        want, _ = runtime.DeviceScene(scene, 0).render_float(st)
        base, _ = scenes.config4_frame(0, 160, 90, 4)
        dev = runtime.DeviceScene(base, 0)
        dev.pose_skeleton(skel, frame, first, 0.0, True)
        got, _ = dev.render_float(st)
        assert np.array_equal(np.nan_to_num(got), np.nan_to_num(want)), frame
    # clamped past the clip, rejected below it and when the target primitives are not cylinders
    dev.pose_skeleton(skel, 10_000, first)
    a, _ = dev.render_float(st)
    dev.pose_skeleton(skel, 120, first)
    b, _ = dev.render_float(st)
    assert np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
    with pytest.raises(runtime.DrtError):
        dev.pose_skeleton(skel, -1, first)
    with pytest.raises(runtime.DrtError):
        dev.pose_skeleton(skel, 0, 0 if first != 0 else len(scene.prims) - 1)
    # drop_y lowers the figure (scene.h:646-650): same as posing, then translating the bones
    dev.pose_skeleton(skel, 3, first, 0.5, False)
    lowered, _ = dev.render_float(st)
    ref_scene, _ = scenes.config4_frame(3, 160, 90, 4)
    for p in ref_scene.prims:
        if p.type == abi.PRIM_CYLINDER:
            p.c1[1] -= 0.5; p.c2[1] -= 0.5; p.center[1] = (p.c1[1] + p.c2[1]) / 2
            p.velocity[:] = [0.0, 0.0, 0.0]
    want, _ = runtime.DeviceScene(ref_scene, 0).render_float(st)
    assert np.array_equal(np.nan_to_num(lowered), np.nan_to_num(want))


@pytest.mark.gpu
def test_cuda_skeleton_rejects_malformed_clips():
    from distraytracer_b200 import runtime, abi
    asf, amc = _clip()
    with pytest.raises(runtime.DrtError) as e:
        runtime.DeviceSkeleton(b"no bone data here\n", amc)
    assert e.value.code == abi.ERR_INVALID
    with pytest.raises(runtime.DrtError):
        runtime.DeviceSkeleton(asf.replace(b"lfemur ltibia", b"lfemur nosuchbone"), amc)
    with pytest.raises(runtime.DrtError):
        runtime.DeviceSkeleton(asf, amc.replace(b"lowerback", b"lowerbach", 1))
    with pytest.raises(runtime.DrtError):
        runtime.DeviceSkeleton(asf, b":FULLY-SPECIFIED\n:DEGREES\n")
    with pytest.raises(runtime.DrtError):
        runtime.DeviceSkeleton("/nonexistent.asf", "/nonexistent.amc")
