"""ctypes binding of oracle/libvworacle.so (numpy in, numpy out).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

COST_ABS, COST_SQ, COST_NCC = 0, 1, 2
PREFILTER_NONE, PREFILTER_LOG, PREFILTER_MEANSUB = 0, 1, 2


def build(force=False):
    """Compile the oracle with gcc (Makefile next to this file)."""
    so = os.path.join(_HERE, "libvworacle.so")
    src = [os.path.join(_HERE, f) for f in ("vw_oracle.c", "vw_sgm_oracle.c", "vw_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


class CorrParams(C.Structure):
    _fields_ = [("search_x0", C.c_int32), ("search_y0", C.c_int32), ("search_x1", C.c_int32), ("search_y1", C.c_int32),
                ("kernel_x", C.c_int32), ("kernel_y", C.c_int32), ("cost_type", C.c_int32),
                ("prefilter_mode", C.c_int32), ("prefilter_width", C.c_float),
                ("consistency_threshold", C.c_float), ("min_consistency_level", C.c_int32),
                ("filter_half_kernel", C.c_int32), ("max_pyramid_levels", C.c_int32), ("collar_size", C.c_int32),
                ("algorithm", C.c_int32), ("sgm_subpixel_mode", C.c_int32), ("sgm_search_buffer_x", C.c_int32),
                ("sgm_search_buffer_y", C.c_int32), ("blob_filter_area", C.c_int32), ("sgm_threads", C.c_int32),
                ("memory_limit_mb", C.c_double)]


class CorrInputs(C.Structure):
    _fields_ = [("left", C.c_void_p), ("lcols", C.c_int), ("lrows", C.c_int), ("lpitch", C.c_int),
                ("right", C.c_void_p), ("rcols", C.c_int), ("rrows", C.c_int), ("rpitch", C.c_int),
                ("lmask", C.c_void_p), ("lmpitch", C.c_int),
                ("rmask", C.c_void_p), ("rmpitch", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.vwo_max_threads.restype = C.c_int
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fast_box_sum(img, kx, ky):
    a = np.ascontiguousarray(img, dtype=np.float64)
    h, w = a.shape
    out = np.empty((h - ky + 1, w - kx + 1), np.float64)
    rc = lib().vwo_fast_box_sum(_p(a), w, h, w, kx, ky, _p(out))
    if rc:
        raise ValueError(f"vwo_fast_box_sum rc={rc}")
    return out


def calc_disparity(cost, left, right, search, kernel):
    """left: (H+ky-1, W+kx-1) f32; right: at least (+sy-1, +sx-1).  Returns int32 (H, W, 3) {dx,dy,valid}."""
    l, r = _f32(left), _f32(right)
    sx, sy = search
    kx, ky = kernel
    H, W = l.shape[0] - ky + 1, l.shape[1] - kx + 1
    out = np.empty((H, W, 3), np.int32)
    rc = lib().vwo_calc_disparity(cost, _p(l), l.shape[1], l.shape[0], l.shape[1],
                                  _p(r), r.shape[1], r.shape[0], r.shape[1], sx, sy, kx, ky, _p(out))
    if rc:
        raise ValueError(f"vwo_calc_disparity rc={rc}")
    return out


def calc_disparity_tiled(cost, left, right, search, kernel, tile=128, nthreads=0):
    l, r = _f32(left), _f32(right)
    sx, sy = search
    kx, ky = kernel
    H, W = l.shape[0] - ky + 1, l.shape[1] - kx + 1
    out = np.empty((H, W, 3), np.int32)
    rc = lib().vwo_calc_disparity_tiled(cost, _p(l), l.shape[1], l.shape[0], l.shape[1],
                                        _p(r), r.shape[1], r.shape[0], r.shape[1], sx, sy, kx, ky, tile, nthreads, _p(out))
    if rc:
        raise ValueError(f"vwo_calc_disparity_tiled rc={rc}")
    return out


def cost_pixel(cost, a, b):
    f = lib().vwo_cost_pixel
    f.restype = C.c_double
    f.argtypes = [C.c_int, C.c_float, C.c_float]
    return f(cost, a, b)


def separable_convolve(img, kx, ky, edge_zero=False, cx=-1, cy=-1):
    a = _f32(img)
    h, w = a.shape
    kxa = np.ascontiguousarray(kx, np.float32)
    kya = np.ascontiguousarray(ky, np.float32)
    out = np.empty_like(a)
    rc = lib().vwo_separable_convolve_c(_p(a), w, h, w, _p(kxa), len(kxa), _p(kya), len(kya), cx, cy,
                                        int(edge_zero), _p(out))
    if rc:
        raise ValueError(rc)
    return out


def pyramid_down(img):
    a = _f32(img)
    h, w = a.shape
    out = np.empty((1 + (h - 1) // 2, 1 + (w - 1) // 2), np.float32)
    rc = lib().vwo_pyramid_down(_p(a), w, h, _p(out))
    if rc:
        raise ValueError(rc)
    return out


def subsample_mask_by_two(m):
    a = np.ascontiguousarray(m, np.uint8)
    h, w = a.shape
    out = np.empty((1 + (h - 1) // 2, 1 + (w - 1) // 2), np.uint8)
    lib().vwo_subsample_mask_by_two(_p(a), w, h, _p(out))
    return out


def prefilter(img, mode, width):
    a = _f32(img)
    h, w = a.shape
    out = np.empty_like(a)
    f = lib().vwo_prefilter
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
    rc = f(_p(a), w, h, mode, width, _p(out))
    if rc:
        raise ValueError(rc)
    return out


def gaussian_kernel(sigma):
    k = np.zeros(512, np.float32)
    lib().vwo_gaussian_kernel.argtypes = [C.c_double, C.c_void_p, C.c_int]
    n = lib().vwo_gaussian_kernel(float(sigma), _p(k), 512)
    return k[:n].copy()


def cross_corr_consistency_check(l2r, r2l, threshold):
    a = np.ascontiguousarray(l2r, np.int32).copy()
    b = np.ascontiguousarray(r2l, np.int32)
    f = lib().vwo_cross_corr_consistency_check
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float]
    f(_p(a), a.shape[1], a.shape[0], a.shape[1], _p(b), b.shape[1], b.shape[0], float(threshold))
    return a


def _filt(name, d, hx, hy, pt, rt):
    a = np.ascontiguousarray(d, np.int32)
    out = np.empty_like(a)
    f = getattr(lib(), name)
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
    rc = f(_p(a), a.shape[1], a.shape[0], hx, hy, pt, rt, _p(out))
    if rc:
        raise ValueError(rc)
    return out


def rm_outliers_using_thresh(d, hx, hy, pt=3.0, rt=0.5):
    return _filt("vwo_rm_outliers_using_thresh", d, hx, hy, pt, rt)


def disparity_cleanup_using_thresh(d, hx, hy, pt=3.0, rt=0.5):
    return _filt("vwo_disparity_cleanup_using_thresh", d, hx, hy, pt, rt)


def disparity_mask(d, lmask, rmask):
    a = np.ascontiguousarray(d, np.int32)
    lm = np.ascontiguousarray(lmask, np.uint8)
    rm = np.ascontiguousarray(rmask, np.uint8)
    out = np.empty_like(a)
    lib().vwo_disparity_mask(_p(a), a.shape[1], a.shape[0], _p(lm), _p(rm), rm.shape[1], rm.shape[0], _p(out))
    return out


def subdivide_regions(d, kernel, max_zones=65536):
    a = np.ascontiguousarray(d, np.int32)
    z = np.empty((max_zones, 8), np.int32)
    n = lib().vwo_subdivide_regions(_p(a), a.shape[1], a.shape[0], kernel[0], kernel[1], _p(z), max_zones)
    if n < 0:
        raise ValueError("too many zones")
    return z[:n].copy()


def make_params(search, kernel, cost=COST_ABS, prefilter_mode=PREFILTER_NONE, prefilter_width=0.0,
                consistency_threshold=-1.0, min_consistency_level=0, filter_half_kernel=0,
                max_pyramid_levels=0, collar_size=0, algorithm=0, sgm_subpixel_mode=5, sgm_search_buffer=(2, 2),
                blob_filter_area=0, sgm_threads=4, memory_limit_mb=6000.0):
    """search = (x0, y0, x1, y1) half-open BBox2i; kernel = (kx, ky)."""
    return CorrParams(search[0], search[1], search[2], search[3], kernel[0], kernel[1], cost,
                      prefilter_mode, prefilter_width, consistency_threshold, min_consistency_level,
                      filter_half_kernel, max_pyramid_levels, collar_size, algorithm, sgm_subpixel_mode,
                      sgm_search_buffer[0], sgm_search_buffer[1], blob_filter_area, sgm_threads, memory_limit_mb)


class _Inputs:
    def __init__(self, left, right, lmask=None, rmask=None):
        self.l, self.r = _f32(left), _f32(right)
        self.lm = np.ascontiguousarray(lmask if lmask is not None else np.full(self.l.shape, 255), np.uint8)
        self.rm = np.ascontiguousarray(rmask if rmask is not None else np.full(self.r.shape, 255), np.uint8)
        assert self.lm.shape == self.l.shape and self.rm.shape == self.r.shape
        self.c = CorrInputs(self.l.ctypes.data, self.l.shape[1], self.l.shape[0], self.l.shape[1],
                            self.r.ctypes.data, self.r.shape[1], self.r.shape[0], self.r.shape[1],
                            self.lm.ctypes.data, self.lm.shape[1], self.rm.ctypes.data, self.rm.shape[1])


def num_levels(params, bw, bh):
    return lib().vwo_num_levels(C.byref(params), bw, bh)


def pyramid_correlate(params, left, right, lmask=None, rmask=None, bbox=None, lr_disp_diff=None, region_ul=(0, 0)):
    """PyramidCorrelationView::rasterize over bbox (default: the whole left image).
    Returns float32 (h, w, 3) {dx, dy, valid}.  lr_disp_diff: float32 (rows, cols, 2) PixelMask<float> image, updated in place."""
    inp = _Inputs(left, right, lmask, rmask)
    if bbox is None:
        bbox = (0, 0, inp.l.shape[1], inp.l.shape[0])
    w, h = bbox[2] - bbox[0], bbox[3] - bbox[1]
    out = np.empty((h, w, 3), np.float32)
    f = lib().vwo_pyramid_correlate_rasterize_ex
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                  C.c_int, C.c_int]
    if lr_disp_diff is not None:
        assert lr_disp_diff.dtype == np.float32 and lr_disp_diff.flags.c_contiguous and lr_disp_diff.shape[2] == 2
    rc = f(C.byref(params), C.byref(inp.c), bbox[0], bbox[1], bbox[2], bbox[3], _p(out), w, None,
           _p(lr_disp_diff) if lr_disp_diff is not None else None, 0 if lr_disp_diff is None else lr_disp_diff.shape[1],
           0 if lr_disp_diff is None else lr_disp_diff.shape[0], region_ul[0], region_ul[1])
    if rc:
        raise ValueError(f"vwo_pyramid_correlate_rasterize rc={rc}")
    return out


def disparity_blob_filter(disp, area):
    """PyramidCorrelationView::disparity_blob_filter (CorrelationView.cc:242-271) at a given area threshold."""
    d = np.ascontiguousarray(disp, np.int32).copy()
    lib().vwo_disparity_blob_filter(_p(d), d.shape[1], d.shape[0], int(area))
    return d


def census_value(img, col, row, k, ternary=False, thr=5):
    a = np.ascontiguousarray(img, np.uint8)
    f = lib().vwo_census_value
    f.restype = C.c_uint64
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    return int(f(_p(a), a.shape[1], col, row, k, int(ternary), thr))


def calc_disparity_sgm(left, right, search, kernel_size, cost_type=3, use_mgm=False, subpixel_mode=0, search_buffer=(2, 2),
                       memory_limit_mb=6000.0, threads=4, lmask=None, rmask=None, prev=None, bounds=None, p1=0, p2=0):
    """The whole of vw::stereo::calc_disparity_sgm (SGM.cc:167-230).  Returns (int32 (h, w, 3), float32 (h, w, 3), boxes (h, w, 4))."""
    l, r = _f32(left), _f32(right)
    oh, ow = sgm_output_shape(l, r, search, kernel_size)
    out = np.zeros((oh, ow, 3), np.int32)
    sub = np.zeros((oh, ow, 3), np.float32)
    b = np.zeros((oh, ow, 4), np.int32) if bounds is None else np.ascontiguousarray(bounds, np.int32).copy()
    lm = None if lmask is None else np.ascontiguousarray(lmask, np.uint8)
    rm = None if rmask is None else np.ascontiguousarray(rmask, np.uint8)
    pv = None if prev is None else np.ascontiguousarray(prev, np.int32)
    f = lib().vwo_calc_disparity_sgm
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    cw, ch = C.c_int(0), C.c_int(0)
    rc = f(_p(l), l.shape[1], l.shape[0], l.shape[1], _p(r), r.shape[1], r.shape[0], r.shape[1], search[0], search[1], kernel_size, cost_type, 5,
           p1, p2, int(use_mgm), subpixel_mode, search_buffer[0], search_buffer[1], float(memory_limit_mb), threads,
           None if lm is None else _p(lm), None if rm is None else _p(rm), 0 if rm is None else rm.shape[1], 0 if rm is None else rm.shape[0],
           None if pv is None else _p(pv), 0 if pv is None else pv.shape[1], 0 if pv is None else pv.shape[0],
           _p(b), int(bounds is not None), _p(out), _p(sub), C.byref(cw), C.byref(ch))
    if rc:
        raise ValueError(f"vwo_calc_disparity_sgm rc={rc}")
    return out, sub, b


def pyramid_correlate_tiled(params, left, right, lmask=None, rmask=None, tile=1024, nthreads=0):
    inp = _Inputs(left, right, lmask, rmask)
    out = np.empty(inp.l.shape + (3,), np.float32)
    rc = lib().vwo_pyramid_correlate_tiled(C.byref(params), C.byref(inp.c), tile, nthreads, _p(out))
    if rc:
        raise ValueError(rc)
    return out


def build_pyramids(params, left, right, lmask=None, rmask=None, bbox=None, levels=None):
    """Per-tile pyramids exactly as build_image_pyramids makes them; list of dicts per level."""
    inp = _Inputs(left, right, lmask, rmask)
    if bbox is None:
        bbox = (0, 0, inp.l.shape[1], inp.l.shape[0])
    if levels is None:
        levels = num_levels(params, bbox[2] - bbox[0], bbox[3] - bbox[1])
    n = levels + 1
    PF = C.POINTER(C.c_float) * n
    PB = C.POINTER(C.c_uint8) * n
    lp, rp, lm, rm = PF(), PF(), PB(), PB()
    dims = np.zeros((n, 8), np.int32)
    rc = lib().vwo_build_pyramids(C.byref(params), C.byref(inp.c), bbox[0], bbox[1], bbox[2], bbox[3], levels,
                                  lp, rp, lm, rm, _p(dims))
    if rc <= 0:
        return None
    out = []
    fr = lib().vwo_free
    fr.argtypes = [C.c_void_p]
    for i in range(n):
        d = dims[i]
        lv = {
            "left": np.ctypeslib.as_array(lp[i], (d[1], d[0])).copy(),
            "right": np.ctypeslib.as_array(rp[i], (d[3], d[2])).copy(),
            "lmask": np.ctypeslib.as_array(lm[i], (d[5], d[4])).copy(),
            "rmask": np.ctypeslib.as_array(rm[i], (d[7], d[6])).copy(),
        }
        for ptr in (lp[i], rp[i], lm[i], rm[i]):
            fr(C.cast(ptr, C.c_void_p))
        out.append(lv)
    return out


def parabola_subpixel(disp, left, right, kernel, prefilter_mode=0, prefilter_width=0.0, bbox=None):
    d = np.ascontiguousarray(disp, np.float32)
    l, r = _f32(left), _f32(right)
    rows, cols = l.shape
    assert d.shape == (rows, cols, 3)
    if bbox is None:
        bbox = (0, 0, cols, rows)
    out = np.empty((bbox[3] - bbox[1], bbox[2] - bbox[0], 3), np.float32)
    f = lib().vwo_parabola_subpixel
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    rc = f(_p(d), cols, rows, _p(l), cols, _p(r), r.shape[1], r.shape[0], r.shape[1], kernel[0], kernel[1],
           prefilter_mode, prefilter_width, bbox[0], bbox[1], bbox[2], bbox[3], _p(out))
    if rc:
        raise ValueError(rc)
    return out


def sgm_output_shape(left, right, search, kernel_size):
    """(rows, cols) of the SGM output raster (SGM.cc:2397-2420)."""
    hk = (kernel_size - 1) // 2
    lh, lw = np.asarray(left).shape
    rh, rw = np.asarray(right).shape
    return (max(0, min(lh - 1 - hk, rh - 1 - (hk + search[1])) - hk + 1), max(0, min(lw - 1 - hk, rw - 1 - (hk + search[0])) - hk + 1))


def sgm_calc_disparity_bounds(left, right, search, kernel_size, bounds, subpixel_mode=0, p1=0, p2=0, use_mgm=False):
    """SGM core with a search box per pixel: bounds (oh, ow, 4) int32 {min_x, min_y, max_x, max_y} inclusive (max < min = none).
    Returns (int32 disparity, float32 sub-pixel disparity)."""
    l, r = _f32(left), _f32(right)
    b = np.ascontiguousarray(bounds, np.int32)
    oh, ow = sgm_output_shape(l, r, search, kernel_size)
    assert b.shape == (oh, ow, 4), (b.shape, (oh, ow, 4))
    out = np.empty((oh, ow, 3), np.int32)
    sub = np.empty((oh, ow, 3), np.float32)
    cw, ch = C.c_int(0), C.c_int(0)
    f = lib().vwo_mgm_calc_disparity_bounds if use_mgm else lib().vwo_sgm_calc_disparity_bounds
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = f(_p(l), l.shape[1], l.shape[0], l.shape[1], _p(r), r.shape[1], r.shape[0], r.shape[1], search[0], search[1], kernel_size, p1, p2,
           subpixel_mode, _p(b), _p(out), _p(sub), C.byref(cw), C.byref(ch))
    if rc:
        raise ValueError(f"vwo_sgm_calc_disparity_bounds rc={rc}")
    return out, sub


def sgm_disp_bounds(shape, search, search_buffer, prev=None, lmask=None, rmask=None, conserve_level=0):
    """populate_disp_bound_image + constrain_disp_bound_image (SGM.cc:241-668).  shape = (oh, ow) of the SGM output raster;
    prev: (ph, pw, 3) int {dx, dy, valid} at half resolution.  Returns (ok, bounds (oh, ow, 4) int32)."""
    oh, ow = shape
    b = np.empty((oh, ow, 4), np.int32)
    pv = None if prev is None else np.ascontiguousarray(prev, np.int32)
    lm = None if lmask is None else np.ascontiguousarray(lmask, np.uint8)
    rm = None if rmask is None else np.ascontiguousarray(rmask, np.uint8)
    if lm is not None:
        assert lm.shape == (oh, ow)
    f = lib().vwo_sgm_disp_bounds
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_void_p]
    rc = f(None if pv is None else _p(pv), 0 if pv is None else pv.shape[1], 0 if pv is None else pv.shape[0],
           None if lm is None else _p(lm), None if rm is None else _p(rm), 0 if rm is None else rm.shape[1], 0 if rm is None else rm.shape[0],
           ow, oh, search[0], search[1], search_buffer[0], search_buffer[1], conserve_level, _p(b))
    if rc < 0:
        raise ValueError(f"vwo_sgm_disp_bounds rc={rc}")
    return bool(rc), b


def sgm_calc_disparity_subpixel(left, right, search, kernel_size, subpixel_mode=5, p1=0, p2=0):
    """calc_disparity_sgm + create_disparity_view_subpixel (SGM.cc:1497-1614).  Returns (int32 disparity, float32 disparity)."""
    l, r = _f32(left), _f32(right)
    out = np.empty((l.shape[0], l.shape[1], 3), np.int32)
    sub = np.empty((l.shape[0], l.shape[1], 3), np.float32)
    ow, oh = C.c_int(0), C.c_int(0)
    f = lib().vwo_sgm_calc_disparity_subpixel
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = f(_p(l), l.shape[1], l.shape[0], l.shape[1], _p(r), r.shape[1], r.shape[0], r.shape[1], search[0], search[1], kernel_size, p1, p2,
           subpixel_mode, _p(out), _p(sub), C.byref(ow), C.byref(oh))
    if rc:
        raise ValueError(f"vwo_sgm_calc_disparity_subpixel rc={rc}")
    n = ow.value * oh.value * 3
    return (out.reshape(-1)[:n].reshape(oh.value, ow.value, 3).copy(), sub.reshape(-1)[:n].reshape(oh.value, ow.value, 3).copy())


def sgm_calc_disparity(left, right, search, kernel_size, p1=0, p2=0):
    """vw::stereo::calc_disparity_sgm (Stereo/SGM.cc:167-230), CENSUS_TRANSFORM, SGM, constant search box [0, search] (inclusive)
    for every pixel.  left / right: the cropped left_region / right_region rasters.  Returns int32 (oh, ow, 3) {dx, dy, valid}."""
    l, r = _f32(left), _f32(right)
    out = np.empty((l.shape[0], l.shape[1], 3), np.int32)
    ow, oh = C.c_int(0), C.c_int(0)
    f = lib().vwo_sgm_calc_disparity
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = f(_p(l), l.shape[1], l.shape[0], l.shape[1], _p(r), r.shape[1], r.shape[0], r.shape[1], search[0], search[1], kernel_size, p1, p2,
           _p(out), C.byref(ow), C.byref(oh))
    if rc:
        raise ValueError(f"vwo_sgm_calc_disparity rc={rc}")
    return out.reshape(-1)[: ow.value * oh.value * 3].reshape(oh.value, ow.value, 3).copy()


def max_threads():
    return lib().vwo_max_threads()
