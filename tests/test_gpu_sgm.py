"""GPU parity of SemiGlobalMatcher with a search box per pixel (row a10): boxes, ragged SGM, MGM.  All through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vwb():
    import visionworkbench_b200 as v
    assert v.device_count() > 0, "GPU tests need a CUDA device: the engine has no CPU path"
    return v


def _pair(seed, W, H, search, off, bits=8):
    rng = np.random.default_rng(seed)
    sx, sy = search
    base = np.floor(rng.random((H + sy + 40, W + sx + 40)) * (1 << bits))
    base = np.floor((base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) / 4).astype(np.float32)
    left = base[20:20 + H, 20:20 + W]
    right = base[20 - off[1]:20 - off[1] + H + sy, 20 - off[0]:20 - off[0] + W + sx]
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


def _prev(rng, oh, ow, search, invalid=0.15, holes=True):
    """a half-resolution prior: smooth field + noise, some invalid pixels and some invalid blocks"""
    ph, pw = (oh + 1) // 2, (ow + 1) // 2
    yy, xx = np.mgrid[0:ph, 0:pw]
    dx = (search[0] / 4 + search[0] / 8 * np.sin(xx / 9.0) + rng.integers(-1, 2, (ph, pw))).astype(np.int32)
    dy = (search[1] / 4 + search[1] / 8 * np.cos(yy / 7.0) + rng.integers(-1, 2, (ph, pw))).astype(np.int32)
    p = np.stack([np.clip(dx, 0, search[0] // 2), np.clip(dy, 0, search[1] // 2), (rng.random((ph, pw)) > invalid).astype(np.int32)], -1)
    if holes:
        p[ph // 3:ph // 3 + 14, pw // 4:pw // 4 + 17, 2] = 0          # larger than the 10-pixel look-around
        p[:3, :, 2] = 0
    return np.ascontiguousarray(p.astype(np.int32))


@pytest.mark.parametrize("level", [0, 1, 2, 3])
@pytest.mark.parametrize("masks", [False, True])
def test_sgm_disp_bounds_matches_oracle(vwb, oracle, level, masks):
    """populate_disp_bound_image + constrain_disp_bound_image (SGM.cc:241-668): integer boxes, bit-identical."""
    rng = np.random.default_rng(7 + level)
    oh, ow, search, buf = 97, 131, (40, 22), (2, 3)
    prev = _prev(rng, oh, ow, search)
    lm = rm = None
    if masks:
        lm = np.full((oh, ow), 255, np.uint8); lm[:, :6] = 0; lm[50:60, 70:90] = 0
        rm = np.full((oh + search[1] + 3, ow + search[0] + 2), 255, np.uint8)
        rm[:, -25:] = 0; rm[:5] = 0; rm[30, :] = 0; rm[-12:, :] = 0
    ok, ref = oracle.sgm_disp_bounds((oh, ow), search, buf, prev=prev, lmask=lm, rmask=rm, conserve_level=level)
    got = vwb.sgm_disp_bounds((oh, ow), search, buf, prev_disparity=prev, left_mask=lm, right_mask=rm, conserve_level=level)
    assert np.array_equal(got, ref), f"{int((got != ref).any(-1).sum())} boxes differ"
    # no prior: constant box (cropped by the right mask)
    ok, ref = oracle.sgm_disp_bounds((oh, ow), search, buf, lmask=lm, rmask=rm, conserve_level=level)
    got = vwb.sgm_disp_bounds((oh, ow), search, buf, left_mask=lm, right_mask=rm, conserve_level=level)
    assert np.array_equal(got, ref)


def test_sgm_disp_bounds_degenerate_hulls(vwb, oracle):
    """BBox2i quirks: a hull grown from identical corners counts as empty (search buffer 0), priors on the search edge."""
    rng = np.random.default_rng(3)
    oh, ow, search = 60, 70, (12, 9)
    prev = _prev(rng, oh, ow, search, invalid=0.5, holes=False)
    for buf in [(0, 0), (0, 2), (5, 5)]:
        for level in (0, 1):
            ok, ref = oracle.sgm_disp_bounds((oh, ow), search, buf, prev=prev, conserve_level=level)
            got = vwb.sgm_disp_bounds((oh, ow), search, buf, prev_disparity=prev, conserve_level=level)
            assert np.array_equal(got, ref), (buf, level)


@pytest.mark.parametrize("kernel,search,shape,buf", [(5, (24, 16), (150, 110), (2, 2)), (3, (9, 9), (90, 120), (1, 1)),
                                                     (7, (40, 6), (130, 80), (3, 1)), (9, (16, 30), (100, 90), (2, 2))])
@pytest.mark.parametrize("mgm", [False, True])
def test_sgm_ragged_matches_oracle(vwb, oracle, kernel, search, shape, buf, mgm):
    """calc_disparity_sgm with a prior (boxes of <= 32 disparities around 2 * prior, full-search pixels where the prior is
    invalid, empty pixels under the mask): bit-identical integers, floats within 1e-5."""
    W, H = shape
    rng = np.random.default_rng(kernel * 100 + W)
    left, right = _pair(11 + kernel, W, H, search, (min(5, search[0]), min(3, search[1])))
    oh, ow = oracle.sgm_output_shape(left, right, search, kernel)
    prev = _prev(rng, oh, ow, search)
    lm = np.full((oh, ow), 255, np.uint8); lm[20:30, 40:60] = 0
    ok, b = oracle.sgm_disp_bounds((oh, ow), search, buf, prev=prev, lmask=lm)
    ri, rf = oracle.sgm_calc_disparity_bounds(left, right, search, kernel, b, subpixel_mode=5, use_mgm=mgm)
    gi, gf, gb = vwb.calc_disparity_sgm_ex(vwb.CENSUS_TRANSFORM, left, right, search, kernel, use_mgm=mgm, subpixel_mode=5, search_buffer=buf,
                                           left_mask=lm, prev_disparity=prev, conserve_level=0, return_bounds=True)
    assert np.array_equal(gb, b), "boxes differ"
    assert np.array_equal(gi, ri), f"{int((gi != ri).any(-1).sum())} of {oh * ow} pixels differ"
    assert np.abs(gf - rf).max() <= 1e-5
    # explicit boxes give the same result
    gi2, gf2 = vwb.calc_disparity_sgm_ex(vwb.CENSUS_TRANSFORM, left, right, search, kernel, use_mgm=mgm, subpixel_mode=5, bounds=b)
    assert np.array_equal(gi2, ri) and np.abs(gf2 - rf).max() <= 1e-5


def test_sgm_ragged_random_boxes(vwb, oracle):
    """arbitrary boxes: every pixel its own size and position (1..7 wide), empty pixels, single-disparity pixels"""
    rng = np.random.default_rng(42)
    search, kernel = (20, 14), 5
    left, right = _pair(5, 120, 100, search, (4, 2))
    oh, ow = oracle.sgm_output_shape(left, right, search, kernel)
    b = np.zeros((oh, ow, 4), np.int32)
    b[..., 0] = rng.integers(0, search[0] + 1, (oh, ow)); b[..., 1] = rng.integers(0, search[1] + 1, (oh, ow))
    b[..., 2] = np.minimum(b[..., 0] + rng.integers(0, 7, (oh, ow)), search[0])
    b[..., 3] = np.minimum(b[..., 1] + rng.integers(0, 7, (oh, ow)), search[1])
    b[rng.random((oh, ow)) < 0.05] = (0, 0, -1, -1)
    b[30:40, 30:50] = (0, 0, search[0], search[1])        # > 32 disparities: the general step
    for mgm in (False, True):
        ri, rf = oracle.sgm_calc_disparity_bounds(left, right, search, kernel, b, subpixel_mode=2, use_mgm=mgm)
        gi, gf = vwb.calc_disparity_sgm_ex(vwb.CENSUS_TRANSFORM, left, right, search, kernel, use_mgm=mgm, subpixel_mode=2, bounds=b)
        assert np.array_equal(gi, ri), f"mgm={mgm}: {int((gi != ri).any(-1).sum())} pixels differ"
        assert np.abs(gf - rf).max() <= 1e-5


def test_sgm_wide_image_long_lines(vwb, oracle):
    """lines longer than the image is high (wrapped diagonals restart several times) and higher than wide"""
    for W, H in [(300, 40), (36, 260)]:
        search, kernel = (6, 4), 3
        left, right = _pair(W, W, H, search, (2, 1))
        oh, ow = oracle.sgm_output_shape(left, right, search, kernel)
        rng = np.random.default_rng(W)
        prev = _prev(rng, oh, ow, search, holes=False)
        ok, b = oracle.sgm_disp_bounds((oh, ow), search, (1, 1), prev=prev)
        ri, _ = oracle.sgm_calc_disparity_bounds(left, right, search, kernel, b)
        gi, _ = vwb.calc_disparity_sgm_ex(vwb.CENSUS_TRANSFORM, left, right, search, kernel, search_buffer=(1, 1), prev_disparity=prev, conserve_level=0)
        assert np.array_equal(gi, ri), (W, H, int((gi != ri).any(-1).sum()))


def test_sgm_memory_limit_retry(vwb, oracle):
    """the retry loop over the conservation levels (SGM.cc:476-497): a tight memory limit selects a higher level"""
    rng = np.random.default_rng(9)
    oh, ow, search, buf = 80, 90, (30, 20), (2, 2)
    prev = _prev(rng, oh, ow, search, invalid=0.3)
    totals = []
    for level in range(4):
        ok, b = oracle.sgm_disp_bounds((oh, ow), search, buf, prev=prev, conserve_level=level)
        n = np.where((b[..., 2] < b[..., 0]) | (b[..., 3] < b[..., 1]), 0, (b[..., 2] - b[..., 0] + 1) * (b[..., 3] - b[..., 1] + 1)).sum()
        totals.append((int(n), b))
    # the loop takes the first level whose footprint fits (main buffers: 3 bytes per entry + one thread's line buffer)
    mb = 1024.0 * 1024.0
    line = int(np.sqrt(ow * ow + oh * oh) + 1) * (search[0] + 1) * (search[1] + 1)
    need = [t[0] * 3 / mb + min(line, t[0]) * 2 / mb for t in totals]
    target = next(L for L in range(1, 4) if need[L] < min(need[:L]))
    limit = (need[target] + min(need[:target])) / 2
    got = vwb.sgm_disp_bounds((oh, ow), search, buf, prev_disparity=prev, conserve_level=-1, memory_limit_mb=limit, assumed_threads=1)
    assert np.array_equal(got, totals[target][1]), target
    got = vwb.sgm_disp_bounds((oh, ow), search, buf, prev_disparity=prev, conserve_level=-1, memory_limit_mb=1e9, assumed_threads=1)
    assert np.array_equal(got, totals[0][1])


@pytest.mark.parametrize("kernel", [3, 5, 7, 9])
def test_sgm_ternary_census_and_parabola(vwb, oracle, kernel):
    """TERNARY_CENSUS_TRANSFORM costs (incl. the 32-bit truncation of the 5x5 signature) and the 2-D parabola sub-pixel mode"""
    search = (10, 7)
    left, right = _pair(60 + kernel, 110, 90, search, (3, 2))
    ri, rf, rb = oracle.calc_disparity_sgm(left, right, search, kernel, cost_type=4, subpixel_mode=1)
    gi, gf = vwb.calc_disparity_sgm_ex(vwb.TERNARY_CENSUS_TRANSFORM, left, right, search, kernel, subpixel_mode=vwb.SUBPIXEL_PARABOLA)
    assert np.array_equal(gi, ri), int((gi != ri).any(-1).sum())
    assert np.abs(gf - rf).max() <= 1e-5
    with pytest.raises(vwb.NoImplErr):
        vwb.calc_disparity_sgm_ex(vwb.ABSOLUTE_DIFFERENCE, left, right, search, kernel)


def test_blob_filter_matches_oracle(vwb, oracle):
    """disparity_blob_filter through the view: BM, one level, blob_filter_area > 0"""
    from visionworkbench_b200.synth import make_pair
    search, kernel = (-6, -4, 7, 5), (7, 7)
    left, right, lm, rm, _ = make_pair(200, 160, search, seed=12, dropout=0.1)
    rng = np.random.default_rng(1)
    lm[rng.random(lm.shape) < 0.35] = 0                      # many small islands of valid pixels
    for area in (3, 25):
        view = vwb.pyramid_correlate(left, right, lm, rm, vwb.PREFILTER_NONE, 0.0, search, kernel, vwb.ABSOLUTE_DIFFERENCE, 0, 0.0, 2.0, 0, 2, 2,
                                     blob_filter_area=area)
        got = view.rasterize(None, (10, 8, 190, 150))
        p = oracle.make_params(search, kernel, cost=0, consistency_threshold=2.0, filter_half_kernel=2, max_pyramid_levels=2, blob_filter_area=area)
        ref = oracle.pyramid_correlate(p, left, right, lm, rm, bbox=(10, 8, 190, 150))
        assert np.array_equal(got, ref), (area, int((got != ref).any(-1).sum()))
    p0 = oracle.make_params(search, kernel, cost=0, consistency_threshold=2.0, filter_half_kernel=2, max_pyramid_levels=2)
    assert not np.array_equal(ref, oracle.pyramid_correlate(p0, left, right, lm, rm, bbox=(10, 8, 190, 150)))   # the filter did something


def test_lr_disp_diff_output(vwb, oracle):
    """lr_disp_diff (CorrelationView.cc:276-283, 669-676, 848-857): two tiles write disjoint windows of one image"""
    from visionworkbench_b200.synth import make_pair
    search, kernel = (-5, -3, 6, 4), (7, 7)
    left, right, lm, rm, _ = make_pair(220, 150, search, seed=3)
    ul = (16, 8)
    gd = np.zeros((120, 180, 2), np.float32); rd = np.zeros((120, 180, 2), np.float32)
    gd[..., 0] = -7.0; rd[..., 0] = -7.0                      # untouched pixels keep their content
    args = (vwb.PREFILTER_NONE, 0.0, search, kernel, vwb.ABSOLUTE_DIFFERENCE, 0, 0.0, 1.0, 0, 2, 1)
    view = vwb.pyramid_correlate(left, right, lm, rm, *args, lr_disp_diff=gd, region_ul=ul)
    p = oracle.make_params(search, kernel, cost=0, consistency_threshold=1.0, filter_half_kernel=2, max_pyramid_levels=1)
    for bbox in [(20, 10, 100, 100), (100, 10, 190, 120)]:
        got = view.rasterize(None, bbox)
        ref = oracle.pyramid_correlate(p, left, right, lm, rm, bbox=bbox, lr_disp_diff=rd, region_ul=ul)
        assert np.array_equal(got, ref)
    assert np.array_equal(gd, rd)
    assert (gd[..., 1] == 1).sum() > 1000 and (gd[..., 0] == -7.0).sum() > 1000
    with pytest.raises(vwb.ArgumentErr):
        view.rasterize(None, (0, 0, 64, 64))                   # not inside the diff image


@pytest.mark.parametrize("algorithm,cost,levels,thr,minlev,fhk,mode", [
    (1, 3, 2, 2.0, 0, 3, 5),        # SGM, census, R->L check at every level, filters, lc_blend
    (1, 3, 3, 1.0, 2, 0, 2),        # check only at the coarse levels, no filters, linear sub-pixel
    (2, 3, 1, -1.0, 0, 2, 1),       # MGM, no check, parabola
    (3, 4, 2, 2.0, 0, 2, 4),        # FINAL_MGM, ternary census, cosine
    (1, 3, 0, 2.0, 0, 3, 0),        # single level, no sub-pixel
])
def test_view_sgm_branch(vwb, oracle, algorithm, cost, levels, thr, minlev, fhk, mode):
    """PyramidCorrelationView with algorithm != BM (CorrelationView.cc:392-595): per-level calc_disparity_sgm seeded by the
    previous level, R->L + consistency check, filters, sub-pixel view: integer part bit-identical, floats within 1e-5."""
    from visionworkbench_b200.synth import make_pair
    search, kernel = (-9, -6, 10, 7), (5, 5)
    left, right, lm, rm, _ = make_pair(260, 200, search, seed=40 + algorithm, bits=8, dropout=0.04)
    view = vwb.pyramid_correlate(left, right, lm, rm, vwb.PREFILTER_LOG, 1.4, search, kernel, cost, 0, 0.0, thr, minlev, fhk, levels, algorithm, 0,
                                 mode, (2, 2), 6000, 4 if algorithm == 3 else 0)
    p = oracle.make_params(search, kernel, cost=cost, prefilter_mode=1, prefilter_width=1.4, consistency_threshold=thr, min_consistency_level=minlev,
                           filter_half_kernel=fhk, max_pyramid_levels=levels, algorithm=algorithm, sgm_subpixel_mode=mode,
                           blob_filter_area=4 if algorithm == 3 else 0)
    for bbox in [(0, 0, 260, 200), (33, 21, 200, 150)]:
        got = view.rasterize(None, bbox)
        ref = oracle.pyramid_correlate(p, left, right, lm, rm, bbox=bbox)
        assert np.array_equal(got[..., 2], ref[..., 2]), f"validity differs at {int((got[..., 2] != ref[..., 2]).sum())} pixels"
        assert np.abs(got - ref).max() <= 1e-5, float(np.abs(got - ref).max())
        assert np.array_equal(np.floor(got[..., :2] + 0.5), np.floor(ref[..., :2] + 0.5))
        assert (ref[..., 2] == 1).mean() > 0.5
    with pytest.raises(vwb.ArgumentErr):                       # SGM with a block-matching cost type (SGM.cc:1888-1892 via :221-226)
        vwb.pyramid_correlate(left, right, lm, rm, 0, 0.0, search, kernel, 0, 0, 0.0, thr, minlev, fhk, levels, algorithm).rasterize(None, (0, 0, 64, 64))


def test_view_prerasterize_ignores_collar(vwb, oracle):
    """prerasterize(bbox) processes exactly bbox; rasterize() adds the collar (CorrelationView.h:123-133)"""
    from visionworkbench_b200.synth import make_pair
    search, kernel = (-6, -4, 7, 5), (7, 7)
    left, right, lm, rm, _ = make_pair(200, 160, search, seed=2)
    view = vwb.pyramid_correlate(left, right, lm, rm, 0, 0.0, search, kernel, 0, 0, 0.0, 2.0, 0, 2, 2, 0, 16)
    p0 = oracle.make_params(search, kernel, cost=0, consistency_threshold=2.0, filter_half_kernel=2, max_pyramid_levels=2, collar_size=0)
    p16 = oracle.make_params(search, kernel, cost=0, consistency_threshold=2.0, filter_half_kernel=2, max_pyramid_levels=2, collar_size=16)
    bbox = (40, 30, 140, 120)
    assert np.array_equal(view.prerasterize(bbox), oracle.pyramid_correlate(p0, left, right, lm, rm, bbox=bbox))
    assert np.array_equal(view.rasterize(None, bbox), oracle.pyramid_correlate(p16, left, right, lm, rm, bbox=bbox))


@pytest.mark.parametrize("maxw,maxh", [(6, 8), (8, 8), (5, 5), (1, 1), (8, 3)])
def test_sgm_small_boxes_rows_in_lanes(vwb, oracle, maxw, maxh):
    """the four-lines-per-warp kernel (no box wider or higher than 8): every pixel its own box size and position, empty
    pixels, single-disparity pixels, boxes jumping by many disparities between neighbours -- and the same boxes through the
    lane-per-disparity kernel (VWB200_SGM_LANES_PER_ENTRY) give the oracle's result too"""
    import os
    rng = np.random.default_rng(maxw * 10 + maxh)
    search, kernel = (23, 17), 5
    for W, H in [(150, 97), (61, 140)]:
        left, right = _pair(W + maxw, W, H, search, (4, 2))
        oh, ow = oracle.sgm_output_shape(left, right, search, kernel)
        b = np.zeros((oh, ow, 4), np.int32)
        b[..., 0] = rng.integers(0, search[0] + 1, (oh, ow)); b[..., 1] = rng.integers(0, search[1] + 1, (oh, ow))
        # smooth-ish regions (boxes shared by 2 x 2 blocks, like a half-resolution prior) mixed with noisy ones
        b[: oh // 2, :, 0] = np.repeat(np.repeat(b[: oh // 2: 2, ::2, 0], 2, 0), 2, 1)[: oh // 2, :ow]
        b[: oh // 2, :, 1] = np.repeat(np.repeat(b[: oh // 2: 2, ::2, 1], 2, 0), 2, 1)[: oh // 2, :ow]
        b[..., 2] = np.minimum(b[..., 0] + rng.integers(0, maxw, (oh, ow)), search[0])
        b[..., 3] = np.minimum(b[..., 1] + rng.integers(0, maxh, (oh, ow)), search[1])
        b[rng.random((oh, ow)) < 0.04] = (0, 0, -1, -1)
        b[10:14, 20:40] = (0, 0, -1, -1)
        ri, rf = oracle.sgm_calc_disparity_bounds(left, right, search, kernel, b, subpixel_mode=5)
        for env in (None, "1"):
            if env:
                os.environ["VWB200_SGM_LANES_PER_ENTRY"] = env
            try:
                gi, gf = vwb.calc_disparity_sgm_ex(vwb.CENSUS_TRANSFORM, left, right, search, kernel, subpixel_mode=5, bounds=b)
            finally:
                os.environ.pop("VWB200_SGM_LANES_PER_ENTRY", None)
            assert np.array_equal(gi, ri), f"{W}x{H} lanes-per-entry={env}: {int((gi != ri).any(-1).sum())} pixels differ"
            assert np.abs(gf - rf).max() <= 1e-5
