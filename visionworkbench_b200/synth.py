"""Synthetic stereo pairs for tests and bench.py.

Integer-valued float imagery so that the reference's double-precision sliding sums are exact
(and therefore independent of summation order): 12-bit noise smoothed once with a 3x3 box and
re-quantised; right image = left gathered through a smooth integer disparity field inside the
search window plus +-2 integer noise; masks 255 except a seeded rectangular dropout.
"""
import numpy as np


def make_pair(width, height, search, seed, bits=12, dropout=0.03, noise=2, smooth=True):
    """search = (x0, y0, x1, y1) half-open BBox2i of disparities.  Returns left, right (float32),
    lmask, rmask (uint8) and the true integer disparity field (dx, dy) int32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    hi = (1 << bits)
    base = np.floor(rng.random((height, width), dtype=np.float32) * hi)
    if smooth:
        p = np.pad(base, 1, mode="edge")
        acc = np.zeros_like(base)
        for dy in range(3):
            for dx in range(3):
                acc += p[dy:dy + height, dx:dx + width]
        base = np.floor(acc / 9.0)
    left = base.astype(np.float32)
    # smooth integer disparity field inside the window: two low-frequency sinusoids
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    cx = 0.5 * (search[0] + search[2] - 1)
    cy = 0.5 * (search[1] + search[3] - 1)
    ax = 0.35 * (search[2] - search[0] - 1)
    ay = 0.35 * (search[3] - search[1] - 1)
    fx = 2 * np.pi / max(width, 64)
    fy = 2 * np.pi / max(height, 64)
    dxf = np.rint(cx + ax * np.sin(1.3 * fx * xx + 0.7 * fy * yy)).astype(np.int32)
    dyf = np.rint(cy + ay * np.cos(0.9 * fx * xx - 1.1 * fy * yy)).astype(np.int32)
    dxf = np.clip(dxf, search[0], search[2] - 1)
    dyf = np.clip(dyf, search[1], search[3] - 1)
    # right(u, v) ~ left(u - dx, v - dy): gather with clamp using the field evaluated at the target
    ui = np.clip(np.arange(width)[None, :] - dxf, 0, width - 1)
    vi = np.clip(np.arange(height)[:, None] - dyf, 0, height - 1)
    right = left[vi, ui]
    if noise:
        rng2 = np.random.Generator(np.random.PCG64(seed + 1))
        right = right + rng2.integers(-noise, noise + 1, size=right.shape).astype(np.float32)
        right = np.clip(right, 0, hi - 1)
    right = right.astype(np.float32)
    lmask = np.full((height, width), 255, np.uint8)
    rmask = np.full((height, width), 255, np.uint8)
    if dropout > 0:
        side = np.sqrt(dropout)
        rw, rh = max(1, int(width * side)), max(1, int(height * side))
        rng3 = np.random.Generator(np.random.PCG64(seed + 2))
        x0 = int(rng3.integers(0, max(1, width - rw)))
        y0 = int(rng3.integers(0, max(1, height - rh)))
        lmask[y0:y0 + rh, x0:x0 + rw] = 0
        x1 = int(rng3.integers(0, max(1, width - rw)))
        y1 = int(rng3.integers(0, max(1, height - rh)))
        rmask[y1:y1 + rh, x1:x1 + rw] = 0
    return left, right, lmask, rmask, (dxf, dyf)


def make_rasters(W, H, search_volume, kernel, seed, bits=12):
    """Inputs of calc_disparity: left (H+ky-1, W+kx-1), right (+sy-1, +sx-1), integer valued."""
    sx, sy = search_volume
    kx, ky = kernel
    lw, lh = W + kx - 1, H + ky - 1
    rw, rh = lw + sx - 1, lh + sy - 1
    left, right, _, _, _ = make_pair(rw, rh, (0, 0, sx, sy), seed, bits=bits, dropout=0)
    return np.ascontiguousarray(left[:lh, :lw]), right


def make_sgm_case(size, search, kernel, seed=104):
    """BASELINE config 4: the cropped left_region / right_region rasters of calc_disparity_sgm for a
    size x size output-ish pair, 8-bit-like imagery, a smooth true disparity inside [0, search]^2, and the synthetic
    half-resolution prior (true / 2, all valid) from which search_buffer (2, 2) gives 5 x 5 boxes.
    Returns left (size, size), right (size + search, size + search), prev ((oh+1)//2, (ow+1)//2, 3) int32,
    true (size, size, 2) int32 (the disparity in the left raster's coordinates)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    H = W = size
    base = np.floor(rng.random((H + search + 2, W + search + 2), dtype=np.float32) * 256)
    p = np.pad(base, 1, mode="edge")
    acc = np.zeros_like(base)
    for dy in range(3):
        for dx in range(3):
            acc += p[dy:dy + base.shape[0], dx:dx + base.shape[1]]
    right = np.floor(acc / 9.0).astype(np.float32)[:H + search, :W + search]
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    # even disparities between 8 and search - 8 so that the half-resolution prior doubles back exactly
    lo, hi = 8, max(search - 8, 10)
    dxf = 2 * np.rint((0.5 * (lo + hi) + 0.4 * (hi - lo) * np.sin(2 * np.pi * (1.3 * xx + 0.7 * yy) / max(W, 64))) / 2).astype(np.int32)
    dyf = 2 * np.rint((0.5 * (lo + hi) + 0.4 * (hi - lo) * np.cos(2 * np.pi * (0.9 * xx - 1.1 * yy) / max(H, 64))) / 2).astype(np.int32)
    dxf = np.clip(dxf, 2, search - 2); dyf = np.clip(dyf, 2, search - 2)
    left = right[np.arange(H)[:, None] + dyf, np.arange(W)[None, :] + dxf]          # left(c, r) = right(c + dx, r + dy)
    noise = np.random.Generator(np.random.PCG64(seed + 1)).integers(-2, 3, size=left.shape).astype(np.float32)
    left = np.clip(left + noise, 0, 255).astype(np.float32)
    hk = (kernel - 1) // 2
    oh, ow = H - 2 * hk, W - 2 * hk
    prev = np.zeros(((oh + 1) // 2, (ow + 1) // 2, 3), np.int32)
    prev[..., 0] = dxf[hk:hk + oh:2, hk:hk + ow:2] // 2
    prev[..., 1] = dyf[hk:hk + oh:2, hk:hk + ow:2] // 2
    prev[..., 2] = 1
    return np.ascontiguousarray(left), np.ascontiguousarray(right), prev, np.stack([dxf, dyf], -1)
