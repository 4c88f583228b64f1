"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py): inputs and expected outputs for every built
row of SURVEY section 8.  The CPU tests freeze the oracle; the GPU tests compare the CUDA engine with the same files without
running the oracle.  Integers exactly, floats within 1e-5."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _t(a):
    return tuple(int(x) for x in np.atleast_1d(a))


def _disp_equal(got, ref, what):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, what
    gv, rv = got[..., 2] != 0, ref[..., 2] != 0
    both = gv & rv
    nbad = int((gv != rv).sum()) + int(((got[..., 0] != ref[..., 0]) | (got[..., 1] != ref[..., 1]))[both].sum())
    assert nbad == 0, f"{what}: {nbad} of {gv.size} pixels differ"


def _parabola_inputs():
    v = _load("pyramid_view")
    g = _load("parabola_subpixel")
    disp = np.stack([g["disparity_x"], g["disparity_y"], g["disparity_valid"]], axis=-1).astype(np.float32)
    return g, disp, v["left"], v["right"]


# ---------------------------------------------------------------- oracle (CPU) ----------------------------------------------------------------
def test_oracle_reproduces_golden(oracle):
    g = _load("calc_disparity_12bit")
    for c in (0, 1, 2):
        _disp_equal(oracle.calc_disparity(c, g["left"], g["right"], _t(g["search"]), _t(g["kernel"])), g[f"cost{c}"], f"12bit cost {c}")
    g = _load("calc_disparity_float")
    for c in (0, 1, 2):
        _disp_equal(oracle.calc_disparity(c, g["left"], g["right"], _t(g["search"]), _t(g["kernel"])), g[f"cost{c}"], f"float cost {c}")
    g = _load("pyramid_down")
    assert np.array_equal(oracle.pyramid_down(g["img"]), g["down"]) and np.array_equal(oracle.subsample_mask_by_two(g["mask"]), g["mask_down"])
    g = _load("prefilter")
    assert np.array_equal(oracle.prefilter(g["img"], 1, 1.4), g["log"]) and np.array_equal(oracle.prefilter(g["img"], 2, 1.4), g["meansub"])
    g = _load("pyramid_view")
    p = oracle.make_params(_t(g["search"]), _t(g["kernel"]), cost=1, consistency_threshold=2.0, filter_half_kernel=3, max_pyramid_levels=3)
    for i, b in enumerate(g["bboxes"]):
        assert np.array_equal(oracle.pyramid_correlate(p, g["left"], g["right"], g["lmask"], g["rmask"], bbox=_t(b)), g[f"tile{i}"])
    g, disp, L, R = _parabola_inputs()
    sb = _t(g["bbox"])
    assert np.array_equal(oracle.parabola_subpixel(disp, L, R, _t(g["kernel"]), 0, 0.0, bbox=sb), g["refined"])
    assert np.array_equal(oracle.parabola_subpixel(disp, L, R, _t(g["kernel"]), 1, 1.4, bbox=sb), g["refined_log"])
    g = _load("sgm_core")
    si, sf = oracle.sgm_calc_disparity_subpixel(g["left"], g["right"], _t(g["search"]), int(g["kernel"]), 5)
    assert np.array_equal(si, g["disparity"]) and np.array_equal(sf, g["subpixel_lc_blend"])


# ---------------------------------------------------------------- CUDA engine (GPU) -------------------------------------------------------------
@pytest.fixture(scope="module")
def vwb():
    from visionworkbench_b200 import build
    build.build()
    import visionworkbench_b200 as v
    return v


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["calc_disparity_12bit", "calc_disparity_float"])
def test_gpu_calc_disparity_golden(vwb, name):
    g = _load(name)
    for c in (0, 1, 2):
        _disp_equal(vwb.calc_disparity(c, g["left"], g["right"], _t(g["search"]), _t(g["kernel"])),
                    g[f"cost{c}"], f"{name} cost {c}")


@pytest.mark.gpu
def test_gpu_pyramid_and_prefilter_golden(vwb):
    g = _load("pyramid_down")
    assert np.array_equal(vwb.pyramid_down(g["img"]), g["down"])
    assert np.array_equal(vwb.subsample_mask_by_two(g["mask"]), g["mask_down"])
    g = _load("prefilter")
    assert np.array_equal(vwb.prefilter_image(g["img"], 1, 1.4), g["log"])
    assert np.array_equal(vwb.prefilter_image(g["img"], 2, 1.4), g["meansub"])


@pytest.mark.gpu
def test_gpu_view_golden(vwb):
    g = _load("pyramid_view")
    search, kernel = _t(g["search"]), _t(g["kernel"])
    view = vwb.pyramid_correlate(g["left"], g["right"], g["lmask"], g["rmask"], vwb.PREFILTER_NONE, 0.0, search, kernel, 1, 0, 0.0, 2.0, 0, 3, 3)
    for i, b in enumerate(g["bboxes"]):
        _disp_equal(view.rasterize(None, _t(b)), g[f"tile{i}"], f"view tile {i}")


@pytest.mark.gpu
def test_gpu_parabola_subpixel_golden(vwb):
    g, disp, L, R = _parabola_inputs()
    sb = _t(g["bbox"])
    kernel = _t(g["kernel"])
    for mode, width, key in [(0, 0.0, "refined"), (1, 1.4, "refined_log")]:
        got = vwb.parabola_subpixel(disp, L, R, mode, width, kernel).rasterize(None, sb)
        ref = g[key]
        assert np.array_equal(got[..., 2], ref[..., 2])
        assert np.abs(got[..., :2] - ref[..., :2])[ref[..., 2] != 0].max() <= 1e-5


@pytest.mark.gpu
def test_gpu_sgm_golden(vwb):
    g = _load("sgm_core")
    gi, gf = vwb.calc_disparity_sgm_subpixel(g["left"], g["right"], _t(g["search"]), int(g["kernel"]), 5)
    assert np.array_equal(gi, g["disparity"])
    assert np.abs(gf - g["subpixel_lc_blend"]).max() <= 1e-5
