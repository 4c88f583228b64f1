"""Pin the CPU oracle with the reference's own known-answer tests.

Each test names the reference test (file:line under src/vw/) whose vector it reproduces.
"""
import numpy as np
import pytest


def _shifted_pair(rng, dtype_max):
    """Stereo/tests/TestCorrelation.cxx:45-54: 25x25 noise; right = constant-edge-extended crop at
    (-3,-8) of size (25+7-1) x (35+12-1)."""
    left = np.floor(rng.random((25, 25)) * dtype_max).astype(np.float32)
    sol = (3, 8)
    w, h = 25 + 7 - 1, 35 + 12 - 1
    ys = np.clip(np.arange(h) - sol[1], 0, 24)
    xs = np.clip(np.arange(w) - sol[0], 0, 24)
    right = left[np.ix_(ys, xs)]
    return left, right, sol


@pytest.mark.parametrize("cost", [0, 1, 2])
@pytest.mark.parametrize("scale", [255.0, 32767.0, 1.0])
def test_calc_disparity_known_shift(oracle, cost, scale):
    """Stereo/tests/TestCorrelation.cxx:56-65,72-214: every pixel valid and == (3,8), 19x21 output."""
    rng = np.random.default_rng(10)
    left, right, sol = _shifted_pair(rng, scale)
    if scale == 1.0:  # the GRAYF32 case: uniform noise in [0,1)
        left = rng.random((25, 25)).astype(np.float32)
        ys = np.clip(np.arange(46) - 8, 0, 24)
        xs = np.clip(np.arange(31) - 3, 0, 24)
        right = left[np.ix_(ys, xs)]
    d = oracle.calc_disparity(cost, left, right, (7, 12), (7, 5))
    assert d.shape == (21, 19, 3)
    assert (d[..., 2] == 1).all()
    assert (d[..., 0] == sol[0]).all() and (d[..., 1] == sol[1]).all()


def test_fast_box_sum_ramp(oracle):
    """Stereo/tests/TestAlgorithms.cxx:45-71: 7x5 ramp 1..35, kernel 5x3."""
    img = np.arange(1, 36, dtype=np.float64).reshape(5, 7)
    out = oracle.fast_box_sum(img, 5, 3)
    assert out.shape == (3, 3)
    assert out[0, 0] == 150 and out[0, 1] == 165 and out[0, 2] == 180
    assert out[2, 0] == 360 and out[2, 1] == 375 and out[2, 2] == 390
    assert oracle.fast_box_sum(img, 3, 3).shape == (3, 5)


def test_fast_box_sum_double(oracle):
    """Stereo/tests/TestAlgorithms.cxx:101-123: ones, first row twos, (6,0)=3, kernel 3x3."""
    img = np.ones((5, 7))
    img[0, :] = 2
    img[0, 6] = 3
    out = oracle.fast_box_sum(img, 3, 3)
    assert out[0, 0] == 12 and out[1, 0] == 9 and out[2, 0] == 9
    assert out[0, 1] == 12 and out[0, 2] == 12 and out[0, 3] == 12 and out[0, 4] == 13


def test_fast_box_sum_char_no_overflow(oracle):
    """Stereo/tests/TestAlgorithms.cxx:125-150."""
    img = np.full((5, 7), 30.0)
    img[0, :] = 40
    img[0, 6] = 50
    out = oracle.fast_box_sum(img, 5, 3)
    assert out[2, 0] == 450 and out[2, 2] == 450 and out[0, 0] == 500 and out[0, 2] == 510 and out[1, 1] == 450
    assert oracle.fast_box_sum(np.full((5, 7), 255.0), 5, 5)[0, 0] == 6375


def test_fast_box_sum_even_kernel_rejected(oracle):
    with pytest.raises(ValueError):
        oracle.fast_box_sum(np.ones((5, 7)), 4, 3)


def test_cost_functors(oracle):
    """Stereo/tests/TestCostFunctions.cxx:36-79: {128,10,100,0} vs {128,100,10,1.0}."""
    a = [128.0, 10.0, 100.0, 0.0]
    b = [128.0, 100.0, 10.0, 1.0]
    assert [oracle.cost_pixel(0, x, y) for x, y in zip(a, b)] == [0, 90, 90, 1]
    assert [oracle.cost_pixel(1, x, y) for x, y in zip(a, b)] == [0, 8100, 8100, 1]
    assert [oracle.cost_pixel(2, x, y) for x, y in zip(a, b)] == [16384, 1000, 1000, 0]


def test_cross_corr_consistency(oracle):
    """Stereo/tests/TestCorrelate.cxx:29-55."""
    r2l = np.zeros((3, 3, 3), np.int32)
    l2r = np.zeros((3, 3, 3), np.int32)
    r2l[..., 2] = 1
    l2r[..., 2] = 1
    l2r[:, 2, 0:2] = 2          # crop(l2r,2,0,1,3) = (2,2)
    l2r[0, 0, 0:2] = 1          # l2r(0,0) = (1,1)
    r2l[1, 1, 0:2] = -1         # r2l(1,1) = (-1,-1)
    l2r[0, 1, 0:2] = 1          # l2r(1,0) = (1,1)
    out = oracle.cross_corr_consistency_check(l2r, r2l, 0)
    assert not out[0, 2, 2] and not out[1, 2, 2] and not out[2, 2, 2]
    assert out[0, 0, 2] and not out[0, 1, 2]
    out = oracle.cross_corr_consistency_check(l2r, r2l, 2)
    assert not out[0, 2, 2] and not out[1, 2, 2] and not out[2, 2, 2]
    assert out[0, 0, 2] and out[0, 1, 2]


def test_separable_convolution_semantics(oracle):
    """Image/tests/TestConvolution.cxx:131-197 and TestFilter.cxx:204-222,258-282 (zero edge)."""
    # SeparableView_0x2: y-only {1,-1} on [[1,2],[4,6]]  -> (0,0)=1 (0,1)=3 (1,0)=2 (1,1)=4  [(col,row)]
    src = np.array([[1, 2], [4, 6]], np.float32)
    d = oracle.separable_convolve(src, [], [1, -1], edge_zero=True)
    assert d[0, 0] == 1 and d[1, 0] == 3 and d[0, 1] == 2 and d[1, 1] == 4
    # SeparableView_2x0: x-only on [[1,2],[3,4]] -> (0,0)=1 (1,0)=1 (0,1)=3 (1,1)=1
    src = np.array([[1, 2], [3, 4]], np.float32)
    d = oracle.separable_convolve(src, [1, -1], [], edge_zero=True)
    assert d[0, 0] == 1 and d[0, 1] == 1 and d[1, 0] == 3 and d[1, 1] == 1
    # Filter.SepConvolution: both -> (0,0)=1 (0,1)=2 (1,0)=1 (1,1)=0
    d = oracle.separable_convolve(src, [1, -1], [1, -1], edge_zero=True)
    assert d[0, 0] == 1 and d[1, 0] == 2 and d[0, 1] == 1 and d[1, 1] == 0
    # Filter.SepConv22Box: {.5,.5} with origin (1,1)
    d = oracle.separable_convolve(src, [0.5, 0.5], [0.5, 0.5], edge_zero=True, cx=1, cy=1)
    assert d[0, 0] == 2.5 and d[1, 0] == 1.75 and d[0, 1] == 1.5 and d[1, 1] == 1
    d = oracle.separable_convolve(np.array([[1, 0], [0, 0]], np.float32), [0.5, 0.5], [0.5, 0.5], edge_zero=True, cx=1, cy=1)
    assert d[0, 0] == 0.25 and d[1, 0] == 0 and d[0, 1] == 0 and d[1, 1] == 0


def test_gaussian_kernel(oracle):
    """Image/tests/TestFilter.cxx:45-73 (sigma 1.5 default size 9; taps to 1e-7)."""
    k = oracle.gaussian_kernel(1.5)
    ref = [0.008488347404, 0.03807782601, 0.1111650246, 0.2113567063, 0.2618241916]
    assert len(k) == 9
    np.testing.assert_allclose(k[:5], ref, atol=1e-7)
    np.testing.assert_allclose(k[5:], ref[3::-1], atol=1e-7)
    assert len(oracle.gaussian_kernel(0)) == 0


def test_pyramid_kernel_is_binomial(oracle):
    """Image/Filter.h:89-99: constant image stays constant; impulse gives outer([1,4,6,4,1])/256 sampled at even pixels."""
    img = np.full((9, 11), 7.0, np.float32)
    assert (oracle.pyramid_down(img) == 7.0).all()
    imp = np.zeros((9, 9), np.float32)
    imp[4, 4] = 256.0
    d = oracle.pyramid_down(imp)
    k = np.array([1, 4, 6, 4, 1], np.float32)
    full = np.zeros((9, 9), np.float32)
    full[2:7, 2:7] = np.outer(k, k)
    np.testing.assert_array_equal(d, full[::2, ::2])
    assert d.shape == (5, 5)


def test_subsample_mask(oracle):
    """Stereo/CorrelationView.cc:38-63: >=2 of 4 on -> 255; odd edge sees zeros."""
    m = np.array([[255, 0, 255], [0, 0, 255], [255, 255, 255]], np.uint8)
    out = oracle.subsample_mask_by_two(m)
    assert out.shape == (2, 2)
    assert out.tolist() == [[0, 255], [255, 0]]


def test_pyramid_view_statistical(oracle):
    """Stereo/tests/TestPyramidCorrelationView.cxx:47-171 shape: noise image, translated copy;
    pyramid correlation with 5 levels recovers the translation for >= 90 % of pixels, >= 99 % valid."""
    rng = np.random.default_rng(7)
    H, W = 200, 300
    base = np.floor(rng.random((H + 40, W + 60)) * 255).astype(np.float32)
    # smooth a little so the pyramid has signal (the reference test uses bicubic resampling)
    base = (base[:-1, :-1] + base[1:, :-1] + base[:-1, 1:] + base[1:, 1:]) / 4
    left = base[10:10 + H, 20:20 + W]
    tx, ty = 15, 5
    right = base[10 - ty:10 - ty + H, 20 - tx:20 - tx + W]    # right(u,v) = left(u-tx, v-ty): disparity (tx,ty)
    for cost in (0, 1, 2):
        p = oracle.make_params((tx - 10, ty - 10, tx + 11, ty + 11), (7, 7), cost=cost,
                               consistency_threshold=-1, filter_half_kernel=5, max_pyramid_levels=5)
        d = oracle.pyramid_correlate(p, left, right)
        valid = d[..., 2] > 0
        # pixels whose match falls inside the right image
        inner = np.zeros((H, W), bool)
        inner[10:H - 10 - ty, 10:W - 10 - tx] = True
        assert valid[inner].mean() >= 0.99
        ok = (d[..., 0] == tx) & (d[..., 1] == ty) & valid
        assert ok[inner].mean() >= 0.90


def test_parabola_subpixel_reference_kat(oracle):
    """Stereo/tests/TestSubPixel.cxx:95-139 (a11): identical images with disparity (1,1) stay at (1,1) +- 0.1 ("null
    test"); on a 95 % horizontally stretched pair the refined disparity has a mean error below 0.6 px and below the
    rounding error of the integer input."""
    const = np.full((50, 50), 7.0, np.float32)
    d = np.zeros((50, 50, 3), np.float32)
    d[..., 0] = 1; d[..., 1] = 1; d[..., 2] = 1
    o = oracle.parabola_subpixel(d, const, const, (7, 7), 0, 0.0)
    assert np.abs(o[..., :2] - 1).max() <= 0.1 and (o[..., 2] == 1).all()
    rng = np.random.default_rng(52)
    W, H, stretch = 120, 90, 0.95
    base = np.floor(rng.random((H, W + 20)) * 1024).astype(np.float32)
    base = np.floor((base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, -1, 0) + np.roll(base, -1, 1)) / 5)
    left = np.ascontiguousarray(base[:, :W])
    xs = np.arange(W) * stretch
    x0 = np.floor(xs).astype(int)
    f = (xs - x0).astype(np.float32)
    right = np.floor(base[:, np.clip(x0, 0, W + 18)] * (1 - f) + base[:, np.clip(x0 + 1, 0, W + 18)] * f).astype(np.float32)
    true = np.arange(W) / stretch - np.arange(W)
    disp = np.zeros((H, W, 3), np.float32)
    disp[..., 0] = np.rint(true)[None, :]
    disp[..., 2] = 1
    o = oracle.parabola_subpixel(disp, left, right, (7, 7), 0, 0.0)
    err = np.abs(o[10:80, 10:100, 0] - true[None, 10:100]).mean()
    assert err < 0.6 and err < np.abs(np.rint(true) - true)[10:100].mean()
