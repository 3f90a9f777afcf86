"""The C-ABI library loads and exports every symbol include/drt.h declares; struct
layouts of the ctypes mirror match.  No compute calls (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from distraytracer_b200 import runtime
    if not os.path.exists(runtime.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return runtime.lib()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "drt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(drt_[a-z_0-9]+)\s*\(", hdr))
    assert {"drt_scene_create", "drt_render", "drt_render_float", "drt_render_device", "drt_device_count"} <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"libdrt.so does not export {name}"
    from distraytracer_b200 import runtime
    assert declared == set(runtime.EXPORTS)


def test_struct_layouts_match(lib):
    from distraytracer_b200 import abi
    sizes = (C.c_int32 * 6)()
    lib.drt_abi_sizes(sizes)
    want = [C.sizeof(t) for t in (abi.Prim, abi.Light, abi.SceneDesc, abi.Settings, abi.Tile, abi.Counters)]
    assert list(sizes) == want


def test_defaults_mirror_reference_globals(lib):
    from distraytracer_b200 import runtime
    s = runtime.default_settings()
    assert (s.xRes, s.yRes, s.antialias_samples, s.brdf_samples, s.blur_samples, s.max_depth) == (1920, 1080, 10, 2, 2, 10)
    assert abs(s.aspect - 1920 / 1080) < 1e-6 and s.aperture == pytest.approx(0.2) and s.focal_length == 10
    assert s.reflect == 1 and s.nogloss == 0 and s.perlin_cloud == 0


def test_no_cpu_fallback_without_device(lib):
    """Without a GPU scene creation must fail loudly (DRT_ERR_NO_DEVICE), never fall back."""
    from distraytracer_b200 import runtime, abi
    from conftest import load_case
    if runtime.device_count() > 0:
        pytest.skip("a GPU is present")
    scene, _, _ = load_case("hw4")
    with pytest.raises(runtime.DrtError) as e:
        runtime.DeviceScene(scene, 0)
    assert e.value.code == abi.ERR_NO_DEVICE


def test_rng_matches_oracle_copy(lib, oracle_lib):
    """The kernels' sample stream (csrc/drt_rng.cuh) and the oracle's independent copy
    (oracle/drt_rng.h) are the same integer function."""
    import numpy as np
    src = r'''
    #include "drt_rng.h"
    double probe(unsigned seed, unsigned pixel, unsigned sample, unsigned child, unsigned dim) {
      unsigned pk = drt_key_pixel(seed, pixel); unsigned sk = drt_key_sample(pk, sample);
      return drt_keyed_u01(drt_key_child(sk, child), dim); }
    '''
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as td:
        open(td + "/p.c", "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "oracle"), td + "/p.c", "-o", td + "/p.so"])
        P = C.CDLL(td + "/p.so")
        P.probe.restype = C.c_double
        P.probe.argtypes = [C.c_uint] * 5
        rng = np.random.default_rng(0)
        for _ in range(2000):
            a = [int(v) for v in rng.integers(0, 2**32, size=5, dtype=np.uint64)]
            assert float(lib.drt_debug_rng(*a)) == P.probe(*a)


def test_keyed_lens_shuffle_matches_oracle_copy_and_the_reference_algorithm(lib, oracle_lib):
    """The reference shuffles a pixel's lens samples with `for i = n-1..1: j = round(u*i); swap(v[i], v[j])`
    (helpers.h:270-279).  The kernels draw u keyed by (pixel, step), evaluate j in integers and ask which lens point ends up
    at position s (lensIndexScan / lensPermsFill); the oracle performs the swaps with `round(double)` on its own copy of
    the stream."""
    import numpy as np
    src = r'''
    #include <math.h>
    #include "drt_rng.h"
    int probe_j(unsigned seed, unsigned pixel, int i) {
      return (int)round(drt_keyed_u01(drt_key_pixel(seed, pixel), DRT_DIM_SHUFFLE(i)) * i); }
    '''
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as td:
        open(td + "/p.c", "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "oracle"), td + "/p.c", "-o", td + "/p.so", "-lm"])
        P = C.CDLL(td + "/p.so")
        P.probe_j.argtypes = [C.c_uint, C.c_uint, C.c_int]
        rng = np.random.default_rng(1)
        for n_lens in (1, 2, 3, 10, 64, 100, 256):
            seed, pixel = (int(v) for v in rng.integers(0, 2**32, size=2, dtype=np.uint64))
            js = [0] + [P.probe_j(seed, pixel, i) for i in range(1, n_lens)]
            assert js == [0] + [lib.drt_debug_shuffle_j(seed, pixel, i) for i in range(1, n_lens)]
            assert all(0 <= j <= i for i, j in enumerate(js))
            v = list(range(n_lens))
            for i in range(n_lens - 1, 0, -1):
                v[i], v[js[i]] = v[js[i]], v[i]
            assert v == [lib.drt_debug_lens_index(seed, pixel, s, n_lens) for s in range(n_lens)]
            assert sorted(v) == list(range(n_lens))


def test_host_bvh_replay_gives_the_reference_candidate_order(lib, oracle_lib):
    """drt_bvh_order.h (product, host side) replays the reference's SAH build; its leaf visiting
    order must equal the oracle's (which is pinned bit-exactly against the compiled reference)."""
    from conftest import GOLDEN_CASES, load_case
    from distraytracer_b200 import abi
    from oracle.harness import Oracle
    for case in GOLDEN_CASES:
        scene, _, _ = load_case(case)
        n = len(scene.prims)
        arr = (abi.Prim * n)(*scene.prims)
        out = (C.c_int32 * n)()
        assert lib.drt_debug_candidate_order(arr, n, out, n) == n
        assert list(out) == Oracle(scene).candidate_order(), case
