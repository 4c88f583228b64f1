"""SemiGlobalMatcher core of the oracle (oracle/vw_sgm_oracle.c): pinned by the reference's own
known-answer test (Stereo/tests/TestSGM.cxx:27-75: constant offset (2,1), search [-4,4]^2, census 3x3, > 99 % correct)
-- on the reference's fixture images when they are present, and on a synthetic equivalent that is always run."""
import os

import numpy as np
import pytest

REF_TESTS = "/root/reference/src/vw/Stereo/tests"


def _constant_offset_pair(seed, W=180, H=150, off=(2, 1), smin=(-4, -4), ssize=(9, 9)):
    """left ROI and the right ROI calc_disparity_sgm is given in TestSGM.cxx:47-52 (right = left ROI + search min,
    grown by the search size), for a right image that is the left one shifted by `off`."""
    rng = np.random.default_rng(seed)
    base = np.floor(rng.random((H + 60, W + 60)) * 256)
    # a little spatial correlation, like a real image
    base = np.floor((base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) / 4).astype(np.float32)
    x0 = y0 = 30
    left = base[y0:y0 + H, x0:x0 + W]
    ox, oy = x0 + smin[0] - off[0], y0 + smin[1] - off[1]
    right = base[oy:oy + H + ssize[1], ox:ox + W + ssize[0]]
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


@pytest.mark.parametrize("kernel", [3, 5, 7, 9])
def test_sgm_constant_offset_synthetic(oracle, kernel):
    left, right = _constant_offset_pair(7 + kernel)
    d = oracle.sgm_calc_disparity(left, right, (8, 8), kernel)      # search volume (9, 9) -> inclusive maxima (8, 8)
    hk = (kernel - 1) // 2
    assert d.shape == (left.shape[0] - 2 * hk, left.shape[1] - 2 * hk, 3)   # SGM.cc:2404-2420
    dd = d[..., :2] + np.array([-4, -4])
    correct = ((dd[..., 0] == 2) & (dd[..., 1] == 1)).mean()
    assert correct > 0.99, correct
    assert (d[..., 2] == 1).all()


def test_sgm_textureless_image_is_deterministic(oracle):
    """All costs tie: select_best_disparity's smoothing iterations (SGM.cc:1196-1288) must terminate and be repeatable."""
    left = np.full((40, 50), 100.0, np.float32)
    right = np.full((48, 58), 100.0, np.float32)
    a = oracle.sgm_calc_disparity(left, right, (8, 8), 3)
    b = oracle.sgm_calc_disparity(left, right, (8, 8), 3)
    assert np.array_equal(a, b) and a.shape[2] == 3


def test_sgm_rejects_unsupported_kernel(oracle):
    left, right = _constant_offset_pair(1, 60, 50)
    with pytest.raises(ValueError):
        oracle.sgm_calc_disparity(left, right, (8, 8), 11)      # NoImplErr in the reference (SGM.cc:1885-1888)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_TESTS, "left.tif")), reason="reference fixtures not present")
def test_sgm_reference_fixture_kat(oracle):
    cv2 = pytest.importorskip("cv2")
    L = cv2.imread(os.path.join(REF_TESTS, "left.tif"), cv2.IMREAD_UNCHANGED).astype(np.float32)
    R = cv2.imread(os.path.join(REF_TESTS, "left_const_offset.tif"), cv2.IMREAD_UNCHANGED).astype(np.float32)
    # TestSGM.cxx uses leftRoi (0,0,400,400), whose right ROI starts at (-4,-4); an interior ROI avoids reading outside the file
    x0, y0, w, h = 8, 8, 380, 380
    left = L[y0:y0 + h, x0:x0 + w]
    right = R[y0 - 4:y0 - 4 + h + 9, x0 - 4:x0 - 4 + w + 9]
    d = oracle.sgm_calc_disparity(left, right, (8, 8), 3)
    dd = d[..., :2] + np.array([-4, -4])
    assert ((dd[..., 0] == 2) & (dd[..., 1] == 1)).mean() > 0.99


def test_sgm_per_pixel_boxes_reduce_to_the_constant_case(oracle):
    """The per-pixel-box entry (what round 2's device kernels consume) with the full box everywhere must equal the
    constant-box pipeline bit for bit, integer and sub-pixel."""
    left, right = _constant_offset_pair(21, 90, 70)
    oh, ow = oracle.sgm_output_shape(left, right, (8, 8), 5)
    full = np.tile(np.array([0, 0, 8, 8], np.int32), (oh, ow, 1))
    bi, bf = oracle.sgm_calc_disparity_bounds(left, right, (8, 8), 5, full, subpixel_mode=5)
    ci, cf = oracle.sgm_calc_disparity_subpixel(left, right, (8, 8), 5, 5)
    assert np.array_equal(bi, ci) and np.array_equal(bf, cf)


def test_sgm_per_pixel_boxes_ragged(oracle):
    """Random sub-boxes around the truth, empty boxes (-> invalid pixels) and single-disparity boxes."""
    rng = np.random.default_rng(5)
    left, right = _constant_offset_pair(22, 100, 80)
    oh, ow = oracle.sgm_output_shape(left, right, (8, 8), 3)
    b = np.empty((oh, ow, 4), np.int32)
    b[..., 0] = rng.integers(0, 7, (oh, ow)); b[..., 1] = rng.integers(0, 6, (oh, ow))          # truth is (6, 5) in box coordinates
    b[..., 2] = np.minimum(8, np.maximum(b[..., 0], 6) + rng.integers(0, 3, (oh, ow)))
    b[..., 3] = np.minimum(8, np.maximum(b[..., 1], 5) + rng.integers(0, 4, (oh, ow)))
    empty = rng.random((oh, ow)) < 0.05
    b[empty] = (0, 0, -1, -1)
    single = (rng.random((oh, ow)) < 0.05) & ~empty
    b[single] = (6, 5, 6, 5)
    di, df = oracle.sgm_calc_disparity_bounds(left, right, (8, 8), 3, b, subpixel_mode=5)
    assert (di[empty] == 0).all() and (df[empty] == 0).all()
    ok = ~empty
    assert (di[ok][:, 2] == 1).all()
    inside = (di[..., 0] >= b[..., 0]) & (di[..., 0] <= b[..., 2]) & (di[..., 1] >= b[..., 1]) & (di[..., 1] <= b[..., 3])
    assert inside[ok].all()
    assert ((di[..., 0] == 6) & (di[..., 1] == 5))[ok].mean() > 0.97
    assert np.abs(df[ok][:, :2] - di[ok][:, :2]).max() <= 1.0      # sub-pixel offsets stay inside one pixel


def test_sgm_disp_bounds_from_previous_level(oracle):
    """populate_disp_bound_image / constrain_disp_bound_image (SGM.cc:241-668): trusted prior -> 2 * d +- buffer clipped to
    the search box; priors on the edge of a >= 10 wide search are not trusted; untrusted pixels take the hull of the
    trusted boxes within 10 pixels, grown by 2 (or the full box when there is none); masks zero the box."""
    oh, ow, search, buf = 40, 50, (19, 11), (2, 3)
    ok, b = oracle.sgm_disp_bounds((oh, ow), search, buf)                       # no prior: the constant box (:231-239)
    assert ok and (b == np.array([0, 0, 19, 11])).all()
    prev = np.zeros((20, 25, 3), np.int32)
    prev[..., 0] = 4; prev[..., 1] = 3; prev[..., 2] = 1                         # -> (8, 6) at this level
    prev[:, :5, 2] = 0                                                           # invalid prior on the left
    prev[10:, 20:, 0] = 0                                                        # dx = 0 is on the edge of a 20-wide search: not trusted
    ok, b = oracle.sgm_disp_bounds((oh, ow), search, buf, prev=prev)
    assert ok
    assert (b[5, 20] == np.array([6, 3, 10, 9])).all()
    # untrusted pixel within 10 px of trusted ones: hull (6,3)-(10,9) expanded by 2, cropped to the search box
    assert (b[5, 5] == np.array([4, 1, 12, 11])).all()
    # untrusted pixel farther than 10 px from any trusted one keeps the full box at conservation level 0 ...
    prev2 = prev.copy(); prev2[:, :, 2] = 0; prev2[0, 24, 2] = 1
    ok, b2 = oracle.sgm_disp_bounds((oh, ow), search, buf, prev=prev2)
    assert (b2[30, 5] == np.array([0, 0, 19, 11])).all()
    # ... and loses its search area at level 1 (:633-640)
    ok, b3 = oracle.sgm_disp_bounds((oh, ow), search, buf, prev=prev2, conserve_level=1)
    assert (b3[39, 0] == np.array([0, 0, -1, -1])).all()
    # left mask
    lm = np.full((oh, ow), 255, np.uint8); lm[:4] = 0
    ok, b4 = oracle.sgm_disp_bounds((oh, ow), search, buf, prev=prev, lmask=lm)
    assert (b4[:4] == np.array([0, 0, -1, -1])).all() and (b4[5, 20] == b[5, 20]).all()
    # the boxes feed the ragged core
    left, right = _constant_offset_pair(23, 60, 50, off=(2, 1), smin=(0, 0), ssize=(20, 12))
    shape = oracle.sgm_output_shape(left, right, search, 3)
    prev = np.zeros(((shape[0] + 1) // 2, (shape[1] + 1) // 2, 3), np.int32); prev[..., 0] = 1; prev[..., 1] = 1; prev[..., 2] = 1
    ok, bb = oracle.sgm_disp_bounds(shape, search, buf, prev=prev)
    di, _ = oracle.sgm_calc_disparity_bounds(left, right, search, 3, bb)
    assert ((di[..., 0] == 2) & (di[..., 1] == 1)).mean() > 0.99


def test_mgm_accumulation(oracle):
    """MGM (accum_mgm_multithread, SGM.cc:2619-2700): same KAT shape as TestSGM.cxx with use_mgm = true -- the constant
    offset is recovered; the accumulated costs differ from plain SGM (two predecessors averaged), so the sub-pixel part does too."""
    left, right = _constant_offset_pair(31, 120, 90)
    oh, ow = oracle.sgm_output_shape(left, right, (8, 8), 3)
    full = np.tile(np.array([0, 0, 8, 8], np.int32), (oh, ow, 1))
    mi, mf = oracle.sgm_calc_disparity_bounds(left, right, (8, 8), 3, full, subpixel_mode=5, use_mgm=True)
    si, sf = oracle.sgm_calc_disparity_bounds(left, right, (8, 8), 3, full, subpixel_mode=5)
    dd = mi[..., :2] + np.array([-4, -4])
    assert ((dd[..., 0] == 2) & (dd[..., 1] == 1)).mean() > 0.99
    assert not np.array_equal(mf, sf)
    mi2, mf2 = oracle.sgm_calc_disparity_bounds(left, right, (8, 8), 3, full, subpixel_mode=5, use_mgm=True)
    assert np.array_equal(mi, mi2) and np.array_equal(mf, mf2)
    # ragged boxes and empty pixels go through the MGM sweeps as well
    b = full.copy(); b[10:20, 10:30] = (0, 0, -1, -1); b[40:50, :, 0] = 4; b[40:50, :, 1] = 3
    ri, _ = oracle.sgm_calc_disparity_bounds(left, right, (8, 8), 3, b, use_mgm=True)
    assert (ri[10:20, 10:30] == 0).all() and (ri[45, :, 0] >= 4).all()
