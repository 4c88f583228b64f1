"""Generate the golden fixtures of tests/golden/ (run from the repository root: `python tests/golden/make_golden.py`).

The reference cannot be compiled or imported here, so the vectors are produced by the CPU oracle -- which is
itself pinned by the reference's known-answer tests (tests/test_oracle_kat.py, tests/test_oracle_sgm.py).  Inputs AND
expected outputs are stored, so the fixtures do not depend on any random-number generator at test time, travel to the GPU
box, and freeze today's oracle behaviour: tests/test_zz_golden.py checks both the oracle (CPU) and the CUDA engine (GPU)
against them.  Integer disparities are compared exactly, floats within 1e-5.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from visionworkbench_b200.synth import make_pair, make_rasters  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print(name, {k: (v.shape, str(v.dtype)) for k, v in arrays.items()})


def main():
    # a3/a4/a5: calc_disparity, three cost functions, 12-bit imagery (fast / screened / general kernels on the GPU side)
    search, kernel = (16, 9), (7, 5)
    left, right = make_rasters(96, 70, search, kernel, seed=401)
    save("calc_disparity_12bit", left=left, right=right, search=np.array(search), kernel=np.array(kernel),
         **{f"cost{c}": oracle.calc_disparity(c, left, right, search, kernel) for c in (0, 1, 2)})
    # the same with non-integer imagery (general fp64 kernel only); quarter-integer values keep every box sum exact, so the
    # result does not depend on the summation order
    rng = np.random.default_rng(402)
    lf = (np.floor(rng.random((40 + 6, 50 + 6)) * 400) / 4).astype(np.float32)
    rf = (np.floor(rng.random((40 + 6 + 4, 50 + 6 + 7)) * 400) / 4).astype(np.float32)
    save("calc_disparity_float", left=lf, right=rf, search=np.array((8, 5)), kernel=np.array((7, 7)),
         **{f"cost{c}": oracle.calc_disparity(c, lf, rf, (8, 5), (7, 7)) for c in (0, 1, 2)})
    # a2: pyramid level (binomial blur + subsample) and mask subsampling
    img = np.floor(rng.random((61, 75)) * 4096).astype(np.float32)
    mask = (rng.random((61, 75)) > 0.2).astype(np.uint8) * 255
    save("pyramid_down", img=img, down=oracle.pyramid_down(img), mask=mask, mask_down=oracle.subsample_mask_by_two(mask))
    # a9: prefilters
    save("prefilter", img=img, log=oracle.prefilter(img, 1, 1.4), meansub=oracle.prefilter(img, 2, 1.4))
    # a1 (+ a6 a7 a8): the whole view on two tiles, 3 levels, L/R check, filters, masks
    vsearch, vkernel = (-10, -6, 14, 10), (9, 9)
    L, R, lm, rm, _ = make_pair(260, 200, vsearch, 403, dropout=0.03)
    p = oracle.make_params(vsearch, vkernel, cost=1, consistency_threshold=2.0, filter_half_kernel=3, max_pyramid_levels=3)
    tiles = [(0, 0, 128, 128), (100, 60, 260, 200)]
    save("pyramid_view", left=L, right=R, lmask=lm, rmask=rm, search=np.array(vsearch), kernel=np.array(vkernel),
         bboxes=np.array(tiles), **{f"tile{i}": oracle.pyramid_correlate(p, L, R, lm, rm, bbox=b) for i, b in enumerate(tiles)})
    # a11: parabola sub-pixel of an integer disparity of the same pair (images: pyramid_view.npz), one tile
    p0 = oracle.make_params(vsearch, vkernel, cost=0, consistency_threshold=-1.0, filter_half_kernel=0, max_pyramid_levels=2)
    disp = oracle.pyramid_correlate(p0, L, R, lm, rm, bbox=(0, 0, 260, 200))
    sb = (40, 30, 200, 170)
    save("parabola_subpixel", disparity_x=disp[..., 0].astype(np.int16), disparity_y=disp[..., 1].astype(np.int16),
         disparity_valid=disp[..., 2].astype(np.uint8), kernel=np.array(vkernel), bbox=np.array(sb),
         refined=oracle.parabola_subpixel(disp, L, R, vkernel, 0, 0.0, bbox=sb), refined_log=oracle.parabola_subpixel(disp, L, R, vkernel, 1, 1.4, bbox=sb))
    # a10: SGM core, census 5, search box [0,8]^2, with the lc_blend sub-pixel stage
    base = np.floor(rng.random((120, 140)) * 256)
    base = np.floor((base + np.roll(base, 1, 0) + np.roll(base, 1, 1)) / 3).astype(np.float32)
    sl = np.ascontiguousarray(base[10:90, 10:110]); sr = np.ascontiguousarray(base[10 - 2:98 - 2, 10 - 3:118 - 3])
    si, sf = oracle.sgm_calc_disparity_subpixel(sl, sr, (8, 8), 5, 5)
    save("sgm_core", left=sl, right=sr, search=np.array((8, 8)), kernel=np.array(5), disparity=si, subpixel_lc_blend=sf)


if __name__ == "__main__":
    main()
