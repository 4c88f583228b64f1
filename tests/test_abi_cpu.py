"""CPU-side checks of the product boundary (no compute without a GPU):
  * libvwb200.so loads and exports every symbol include/vwb200.h declares
  * argument validation mirrors the reference's asserts
  * with no CUDA device every compute call fails loudly (no CPU fallback exists)
  * the product package never imports the oracle
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vwb():
    sys.path.insert(0, ROOT)
    from visionworkbench_b200 import build
    build.build()
    import visionworkbench_b200 as v
    return v


def test_every_declared_symbol_is_exported(vwb):
    hdr = open(os.path.join(ROOT, "include", "vwb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(vwb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 19
    so = os.path.join(ROOT, "visionworkbench_b200", "libvwb200.so")
    out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = set(re.findall(r" T (vwb200_[a-z0-9_]+)", out))
    missing = names - exported
    assert not missing, f"declared but not exported: {sorted(missing)}"
    L = vwb.lib()
    for n in names:
        assert getattr(L, n) is not None


def test_sass_is_sm100a(vwb):
    so = os.path.join(ROOT, "visionworkbench_b200", "libvwb200.so")
    out = subprocess.run(["cuobjdump", "--list-elf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_validation(vwb):
    l = np.zeros((11, 11), np.float32)
    r = np.zeros((14, 14), np.float32)
    with pytest.raises(vwb.ArgumentErr):       # even kernel (Correlation.cc:340-341)
        vwb.calc_disparity(0, l, r, (4, 4), (4, 5))
    with pytest.raises(vwb.ArgumentErr):       # zero search volume (:345-346)
        vwb.calc_disparity(0, l, r, (0, 4), (5, 5))
    with pytest.raises(vwb.ArgumentErr):       # right raster too small
        vwb.calc_disparity(0, l, r, (9, 9), (5, 5))
    with pytest.raises(vwb.ArgumentErr):       # kernel larger than the region (:342-344)
        vwb.calc_disparity(0, l, r, (2, 2), (13, 13))
    m = np.full((11, 11), 255, np.uint8)
    with pytest.raises(vwb.ArgumentErr):       # even kernel in the view
        vwb.PyramidCorrelationView(l, l, m, m, 0, 0, (0, 0, 4, 4), (4, 4), 0, 0, 0.0, -1, 0, 0, 0)
    with pytest.raises(vwb.ArgumentErr):       # unknown algorithm / cost type
        vwb.PyramidCorrelationView(l, l, m, m, 0, 0, (0, 0, 4, 4), (5, 5), 0, 0, 0.0, -1, 0, 0, 0, algorithm=7)
    with pytest.raises(vwb.ArgumentErr):
        vwb.PyramidCorrelationView(l, l, m, m, 0, 0, (0, 0, 4, 4), (5, 5), 9, 0, 0.0, -1, 0, 0, 0)


def test_no_device_means_loud_failure(vwb):
    if vwb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    l = np.zeros((11, 11), np.float32)
    r = np.zeros((14, 14), np.float32)
    with pytest.raises(vwb.NoDeviceErr):
        vwb.calc_disparity(0, l, r, (4, 4), (5, 5))
    with pytest.raises(vwb.NoDeviceErr):
        vwb.pyramid_down(l)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "visionworkbench_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "vw_oracle" not in txt and "vwo_" not in txt, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "vwo_" not in open(p).read()


def test_view_is_lazy_and_operator_call_throws(vwb):
    if vwb.device_count() > 0:
        pytest.skip("needs the no-device behaviour")
    l = np.zeros((32, 32), np.float32)
    m = np.full((32, 32), 255, np.uint8)
    # construction places inputs in HBM -> fails loudly without a device
    with pytest.raises(vwb.NoDeviceErr):
        vwb.PyramidCorrelationView(l, l, m, m, 0, 0, (0, 0, 4, 4), (5, 5), 0, 0, 0.0, -1, 0, 0, 0)


def test_sgm_boundary_without_a_device(vwb):
    """vwb200_sgm_calc_disparity: the output-size query (SGM.cc:2397-2420) and the reference's argument asserts
    (:184-188) are host logic; the compute call fails loudly without a device."""
    import ctypes as C
    L = vwb.lib()
    left = np.zeros((50, 60), np.float32)
    right = np.zeros((58, 68), np.float32)
    ow, oh = C.c_int(-1), C.c_int(-1)
    args = (left.ctypes.data, 60, 50, 60, right.ctypes.data, 68, 58, 68)
    assert L.vwb200_sgm_calc_disparity(*args, 8, 8, 5, 0, 0, None, 0, C.byref(ow), C.byref(oh), 0, None) == 0
    assert (ow.value, oh.value) == (60 - 4, 50 - 4)                  # cropped by the 5x5 census kernel only: the right raster covers the search
    small = np.zeros((50, 60), np.float32)                              # right raster without room for the search: the output shrinks
    assert L.vwb200_sgm_calc_disparity(left.ctypes.data, 60, 50, 60, small.ctypes.data, 60, 50, 60, 8, 8, 5, 0, 0, None, 0,
                                       C.byref(ow), C.byref(oh), 0, None) == 0
    assert (ow.value, oh.value) == (60 - 4 - 8, 50 - 4 - 8)
    with pytest.raises(vwb.ArgumentErr):       # even kernel (SGM.cc:184-185)
        vwb.calc_disparity_sgm(left, right, (8, 8), 4)
    with pytest.raises(vwb.ArgumentErr):       # kernel larger than the region (:187-188)
        vwb.calc_disparity_sgm(left[:3], right, (8, 8), 5)
    if vwb.device_count() == 0:
        with pytest.raises(vwb.VwError) as e:
            vwb.calc_disparity_sgm(left, right, (8, 8), 5)
        assert "CUDA" in str(e.value) or "device" in str(e.value)
