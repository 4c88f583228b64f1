"""The integer zone kernel of the pyramid level loop (csrc/k1_zone_int.cu) against the oracle's calc_disparity
(Stereo/Correlation.cc:33-137): VWB200_K1_ZONE_INT routes calc_disparity on integer-valued rasters through it as ONE zone, so
every tile shape, kernel size, disparity-chunk split and the all-equal / tie rules are compared bit for bit.  The level loop
itself (thousands of zones per level) is covered by the view tests; the last test here pins it against the fp64 zone kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vwb():
    import visionworkbench_b200 as v
    assert v.device_count() > 0, "GPU tests need a CUDA device: the engine has no CPU path"
    return v


def rasters(rng, w, h, search, kernel, vmax=255, flat=False):
    sx, sy = search; kx, ky = kernel
    right = np.floor(rng.random((h + ky - 1 + sy - 1, w + kx - 1 + sx - 1)) * (vmax + 1)).astype(np.float32)
    dx, dy = int(rng.integers(0, sx)), int(rng.integers(0, sy))
    left = right[dy:dy + h + ky - 1, dx:dx + w + kx - 1].copy()
    left += np.floor(rng.random(left.shape) * 3) - 1
    left = np.clip(left, 0, vmax).astype(np.float32)
    if flat:                       # constant patches: every disparity has the same cost -> invalid pixels, ties everywhere
        left[: h // 2, : w // 2] = 7
        right[: h // 2 + sy, : w // 2 + sx] = 7
    return left, right


CASES = [  # (w, h), search, k
    ((18, 16), (5, 4), 15), ((19, 17), (6, 6), 15), ((61, 45), (11, 9), 15), ((8, 8), (10, 10), 15), ((40, 33), (1, 1), 15),
    ((50, 20), (1, 7), 5), ((33, 50), (9, 1), 3), ((70, 40), (20, 20), 7), ((45, 45), (17, 16), 9), ((30, 30), (13, 5), 11),
    ((64, 37), (8, 8), 13), ((25, 60), (12, 12), 17), ((41, 29), (7, 9), 19), ((36, 36), (10, 10), 21), ((20, 50), (6, 6), 23),
    ((17, 17), (5, 5), 25), ((100, 64), (40, 30), 15), ((48, 48), (128, 3), 15),
]


@pytest.mark.parametrize("cost", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_zone_int_single_zone(vwb, oracle, monkeypatch, cost, case):
    (w, h), search, k = case
    rng = np.random.default_rng(hash((w, h, search, k, cost)) % (1 << 31))
    left, right = rasters(rng, w, h, search, (k, k))
    ref = oracle.calc_disparity(cost, left, right, search, (k, k))
    monkeypatch.setenv("VWB200_K1_ZONE_INT", "1")
    n0 = vwb.kernel_launches()
    got = vwb.calc_disparity(cost, left, right, search, (k, k))
    assert vwb.kernel_launches() > n0
    assert got.shape == ref.shape == (h, w, 3)
    assert np.array_equal(got, ref), f"{int((got != ref).any(-1).sum())} of {w * h} pixels differ"


@pytest.mark.parametrize("case", [CASES[i] for i in (0, 1, 2, 3, 6, 7, 9, 12, 16, 17)])
def test_zone_int_wide_squared_cost_12bit(vwb, oracle, monkeypatch, case):
    """SquaredCost on 12-bit imagery (the synthetic pairs of SURVEY 8d): window sums need all 32 bits -> the WIDE variant"""
    (w, h), search, k = case
    rng = np.random.default_rng(hash((w, h, search, k)) % (1 << 31))
    left, right = rasters(rng, w, h, search, (k, k), vmax=4095 if k <= 15 else 2047)
    ref = oracle.calc_disparity(1, left, right, search, (k, k))
    monkeypatch.setenv("VWB200_K1_ZONE_INT", "1")
    got = vwb.calc_disparity(1, left, right, search, (k, k))
    assert np.array_equal(got, ref), f"{int((got != ref).any(-1).sum())} of {w * h} pixels differ"


@pytest.mark.parametrize("cost,vmax", [(0, 255), (1, 255), (1, 4095)])
def test_zone_int_flat_regions_and_ties(vwb, oracle, monkeypatch, cost, vmax):
    rng = np.random.default_rng(5)
    left, right = rasters(rng, 50, 44, (9, 8), (15, 15), vmax=vmax, flat=True)
    ref = oracle.calc_disparity(cost, left, right, (9, 8), (15, 15))
    assert (ref[..., 2] == 0).any() and (ref[..., 2] == 1).any()
    monkeypatch.setenv("VWB200_K1_ZONE_INT", "1")
    got = vwb.calc_disparity(cost, left, right, (9, 8), (15, 15))
    assert np.array_equal(got, ref)


def test_zone_int_wide_range_abs(vwb, oracle, monkeypatch):
    """Abs cost keeps 16-bit imagery on the integer kernel (per-pixel cost < 2^16, window sum < 2^24 with 15x15 only up to
    range 74565/225: here 12 bits)"""
    rng = np.random.default_rng(6)
    left, right = rasters(rng, 40, 40, (9, 9), (15, 15), vmax=4095)
    ref = oracle.calc_disparity(0, left, right, (9, 9), (15, 15))
    monkeypatch.setenv("VWB200_K1_ZONE_INT", "1")
    got = vwb.calc_disparity(0, left, right, (9, 9), (15, 15))
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("cost,bits", [(0, 8), (1, 8), (1, 12)])
def test_view_level0_int_zones_equal_fp64_zones(vwb, monkeypatch, cost, bits):
    """PyramidCorrelationView on integer imagery: level 0 through the integer zone kernel == through the fp64 zone kernel"""
    from visionworkbench_b200.synth import make_pair
    search = (-24, -20, 24, 20)
    left, right, lm, rm, _ = make_pair(700, 520, search, seed=31, bits=bits)
    args = (left, right, lm, rm, vwb.PREFILTER_NONE, 0.0, search, (15, 15), cost, 0, 0.0, 2.0, 0, 5, 4)
    a = vwb.pyramid_correlate(*args).rasterize()
    monkeypatch.setenv("VWB200_NO_ZONE_INT", "1")
    b = vwb.pyramid_correlate(*args).rasterize()
    assert np.array_equal(a, b)
    assert (a[..., 2] == 1).mean() > 0.5
