// k34_filters.cu -- the O(pixels) stages of the level loop that run between K1 launches:
//   K3  cross_corr_consistency_check        (Stereo/Correlate.cc:1441-1502)
//   K4  rm_outliers_using_thresh / disparity_cleanup_using_thresh / disparity_mask
//       (Stereo/DisparityMap.h:318-442, :97-253)
//   finalize: + search_region.min(), int -> float pixel cast (Stereo/CorrelationView.cc:880-884)
// All are integer work on {dx,dy,valid} triples; HBM-bound.
#include "common.cuh"

namespace vwb200 {

__device__ __forceinline__ int clampi3(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- K3 -------------------------------------------------------------------------------------------
__device__ __forceinline__ void consistency_pixel(vwb200_dispi* p, int c, int r, const vwb200_dispi* r2l, int rw, int rh,
                                                  ptrdiff_t rpitch, float thr) {
  const vwb200_dispi v = *p;
  const int x = c + v.dx, y = r + v.dy;
  if (x < 0 || x >= rw || y < 0 || y >= rh) { p->valid = 0; return; }
  const vwb200_dispi q = r2l[(ptrdiff_t)y * rpitch + x];
  if (!v.valid || !q.valid) { p->valid = 0; return; }
  const double a = fabs((double)(v.dx + q.dx)), b = fabs((double)(v.dy + q.dy));
  const float diff = (float)(a > b ? a : b);
  if (!(thr >= diff)) p->valid = 0;
}
__global__ void consistency_kernel(vwb200_dispi* l2r, int lw, int lh, ptrdiff_t lpitch, const vwb200_dispi* r2l, int rw,
                                   int rh, ptrdiff_t rpitch, float thr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= lw || r >= lh) return;
  consistency_pixel(l2r + (ptrdiff_t)r * lpitch + c, c, r, r2l, rw, rh, rpitch, thr);
}
int consistency_launch(vwb200_dispi* l2r, int lw, int lh, ptrdiff_t lpitch, const vwb200_dispi* r2l, int rw, int rh,
                       ptrdiff_t rpitch, float thr, cudaStream_t st) {
  if (lw <= 0 || lh <= 0) return VWB200_OK;
  dim3 b(32, 8), g((lw + 31) / 32, (lh + 7) / 8);
  consistency_kernel<<<g, b, 0, st>>>(l2r, lw, lh, lpitch, r2l, rw, rh, rpitch, thr);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// Zone-table form used by the level loop: per zone, optional L/R check against that zone's R->L
// image, then "+= zone.disparity_range().min()" (CorrelationView.cc:691-698).  One launch per level.
__global__ void zone_post_kernel(const Tile* __restrict__ tiles, const Zone* __restrict__ zones, const Zone* __restrict__ rlzones,
                                 const int2* __restrict__ post_add, vwb200_dispi* __restrict__ disp,
                                 const vwb200_dispi* __restrict__ rl, float thr, int tile) {
  const Tile t = tiles[blockIdx.x];
  const Zone z = zones[t.zone];
  const int tw = min(tile, z.w - t.tx), th = min(tile, z.h - t.ty);
  const int2 add = post_add[t.zone];
  for (int k = threadIdx.x; k < tw * th; k += blockDim.x) {
    const int x = t.tx + k % tw, y = t.ty + k / tw;
    vwb200_dispi* p = disp + z.obase + (ptrdiff_t)y * z.opitch + x;
    if (rl) {
      const Zone q = rlzones[t.zone];
      consistency_pixel(p, x, y, rl + q.obase, q.w, q.h, q.opitch, thr);
    }
    p->dx += add.x;
    p->dy += add.y;
  }
}
int zone_post_launch(const Tile* d_tiles, int ntiles, const Zone* d_zones, const Zone* d_rlzones, const int2* d_post_add,
                     vwb200_dispi* disp, const vwb200_dispi* rl, float thr, int tile, cudaStream_t st) {
  if (ntiles <= 0) return VWB200_OK;
  zone_post_kernel<<<ntiles, 256, 0, st>>>(d_tiles, d_zones, d_rlzones, d_post_add, disp, rl, thr, tile);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// ---- K4: RmOutliersUsingThreshFunc (DisparityMap.h:359-386) -----------------------------------------
// Evaluated at (cx,cy) of the ConstantEdgeExtension of `in`; the output window may extend past the
// image (second-pass quirk of disparity_cleanup_using_thresh, Image/PerPixelAccessorViews.h:63-85).
// CTA stages its (tile + halo) neighbourhood in shared memory (dx,dy packed, validity folded in).
static constexpr int RO_TW = 32, RO_TH = 8;
__global__ void __launch_bounds__(RO_TW * RO_TH)
rm_outliers_kernel(const vwb200_dispi* __restrict__ in, int w, int h, int hx, int hy, double pt, double rt,
                   int x0, int y0, int ow, int oh, vwb200_dispi* __restrict__ out) {
  extern __shared__ int sm[];                // [sh][sw] x {dx, dy}; invalid -> dx = INT_MIN
  const int sw = RO_TW + 2 * hx, sh = RO_TH + 2 * hy;
  int* sdx = sm; int* sdy = sm + sw * sh;
  const int bx = x0 + blockIdx.x * RO_TW, by = y0 + blockIdx.y * RO_TH;
  const int tid = threadIdx.y * RO_TW + threadIdx.x;
  for (int k = tid; k < sw * sh; k += RO_TW * RO_TH) {
    const int sx = k % sw, sy = k / sw;
    const vwb200_dispi v = in[(ptrdiff_t)clampi3(by - hy + sy, 0, h - 1) * w + clampi3(bx - hx + sx, 0, w - 1)];
    sdx[k] = v.valid ? v.dx : INT_MIN;
    sdy[k] = v.dy;
  }
  __syncthreads();
  const int ox = blockIdx.x * RO_TW + threadIdx.x, oy = blockIdx.y * RO_TH + threadIdx.y;
  if (ox >= ow || oy >= oh) return;
  const int cx = threadIdx.x + hx, cy = threadIdx.y + hy;
  // the centre pixel keeps its child values even when invalid ("return *acc")
  vwb200_dispi c = in[(ptrdiff_t)clampi3(y0 + oy, 0, h - 1) * w + clampi3(x0 + ox, 0, w - 1)];
  if (c.valid) {
    int matched = 0;
    const int total = (2 * hx + 1) * (2 * hy + 1);
    for (int yk = -hy; yk <= hy; ++yk) {
      const int* rdx = sdx + (cy + yk) * sw + cx;
      const int* rdy = sdy + (cy + yk) * sw + cx;
      for (int xk = -hx; xk <= hx; ++xk) {
        const int ndx = rdx[xk];
        if (ndx != INT_MIN && fabs((double)(c.dx - ndx)) <= pt && fabs((double)(c.dy - rdy[xk])) <= pt) ++matched;
      }
    }
    if (((double)matched / (double)total) < rt) { c.dx = 0; c.dy = 0; c.valid = 0; }
  }
  out[(ptrdiff_t)oy * ow + ox] = c;
}
int rm_outliers_launch(const vwb200_dispi* in, int w, int h, int hx, int hy, double pt, double rt,
                       int x0, int y0, int ow, int oh, vwb200_dispi* out, cudaStream_t st) {
  if (ow <= 0 || oh <= 0) return VWB200_OK;
  if (hx <= 0 || hy <= 0) { set_error("RmOutliersFunc: half kernel sizes must be non-zero."); return VWB200_EARG; }  // DisparityMap.h:345-346
  const size_t smem = (size_t)(RO_TW + 2 * hx) * (RO_TH + 2 * hy) * 2 * sizeof(int);
  if (smem > 200 * 1024) { set_error("outlier filter half kernel %dx%d too large", hx, hy); return VWB200_ENOIMPL; }
  VWB_CUDA(cudaFuncSetAttribute(rm_outliers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 b(RO_TW, RO_TH), g((ow + RO_TW - 1) / RO_TW, (oh + RO_TH - 1) / RO_TH);
  rm_outliers_kernel<<<g, b, smem, st>>>(in, w, h, hx, hy, pt, rt, x0, y0, ow, oh, out);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// second pass (1,1,3.0,0.20) over the padded first-pass buffer p1 of size (w+2) x (h+2)  (DisparityMap.h:440)
__global__ void cleanup_pass2_kernel(const vwb200_dispi* __restrict__ p1, int w, int h, vwb200_dispi* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int pw = w + 2;
  vwb200_dispi c = p1[(ptrdiff_t)(y + 1) * pw + (x + 1)];
  if (c.valid) {
    int matched = 0;
#pragma unroll
    for (int yk = -1; yk <= 1; ++yk)
#pragma unroll
      for (int xk = -1; xk <= 1; ++xk) {
        const vwb200_dispi n = p1[(ptrdiff_t)(y + 1 + yk) * pw + (x + 1 + xk)];
        if (n.valid && fabs((double)(c.dx - n.dx)) <= 3.0 && fabs((double)(c.dy - n.dy)) <= 3.0) ++matched;
      }
    if (((double)matched / 9.0) < 0.20) { c.dx = 0; c.dy = 0; c.valid = 0; }
  }
  out[(ptrdiff_t)y * w + x] = c;
}
int cleanup_pass2_launch(const vwb200_dispi* p1, int w, int h, vwb200_dispi* out, cudaStream_t st) {
  if (w <= 0 || h <= 0) return VWB200_OK;
  dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
  cleanup_pass2_kernel<<<g, b, 0, st>>>(p1, w, h, out);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// ---- disparity_mask (DisparityMap.h:142-162) --------------------------------------------------------------
__global__ void disparity_mask_kernel(const vwb200_dispi* __restrict__ in, int w, int h, ImgB lm, ImgB rm, vwb200_dispi* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= w || j >= h) return;
  vwb200_dispi z; z.dx = 0; z.dy = 0; z.valid = 0;
  vwb200_dispi d = in[(ptrdiff_t)j * w + i];
  bool keep = lm.p[(ptrdiff_t)j * lm.pitch + i] != 0 && d.valid;
  if (keep) {
    const int tx = i + d.dx, ty = j + d.dy;
    keep = !(tx < 0 || tx >= rm.w || ty < 0 || ty >= rm.h) && rm.p[(ptrdiff_t)ty * rm.pitch + tx] != 0;
  }
  out[(ptrdiff_t)j * w + i] = keep ? d : z;
}
int disparity_mask_launch(const vwb200_dispi* in, int w, int h, ImgB lmask, ImgB rmask, vwb200_dispi* out, cudaStream_t st) {
  if (w <= 0 || h <= 0) return VWB200_OK;
  dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
  disparity_mask_kernel<<<g, b, 0, st>>>(in, w, h, lmask, rmask, out);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// ---- finalize: disparity + search.min, pixel_cast<PixelMask<Vector2f>> (CorrelationView.cc:880-884) -------
// Writes the [ox,ox+ow) x [oy,oy+oh) window of the w x h tile result to a 12-byte-pixel destination.
__global__ void finalize_kernel(const vwb200_dispi* __restrict__ in, int w, int ax, int ay, float* __restrict__ out,
                                ptrdiff_t opitch_px, int ox, int oy, int ow, int oh) {
  // one thread per output float: fully coalesced 4-byte stores over the 12-byte pixels
  const int k = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (k >= 3 * ow || y >= oh) return;
  const int x = k / 3, c = k % 3;
  const int* src = reinterpret_cast<const int*>(in + (ptrdiff_t)(oy + y) * w + (ox + x));
  float v;
  if (c == 0) v = (float)(src[0] + ax);
  else if (c == 1) v = (float)(src[1] + ay);
  else v = src[2] ? 1.0f : 0.0f;
  out[(ptrdiff_t)y * opitch_px * 3 + k] = v;
}
int finalize_launch(const vwb200_dispi* in, int w, int h, int ax, int ay, float* out, ptrdiff_t opitch_px,
                    int ox, int oy, int ow, int oh, cudaStream_t st) {
  (void)h;
  if (ow <= 0 || oh <= 0) return VWB200_OK;
  dim3 b(256), g((3 * ow + 255) / 256, oh);
  finalize_kernel<<<g, b, 0, st>>>(in, w, ax, ay, out, opitch_px, ox, oy, ow, oh);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

}  // namespace vwb200
