// k5_sgm.cu -- SemiGlobalMatcher core on the device (first cut).
//
// vw::stereo::calc_disparity_sgm (Stereo/SGM.cc:167-230) -> SemiGlobalMatcher::semi_global_matching_func (:2387-2448) for
// CENSUS_TRANSFORM costs (kernel 3/5/7/9), plain SGM (not MGM), the same search box [0, sx] x [0, sy] for every pixel (no
// masks, no previous disparity), integer winner (create_disparity_view, :1290-1346).  Everything is integer / bit
// arithmetic except u8_convert's stretch and the tie smoothing of select_best_disparity, which use the reference's double
// operations one by one; results are bit-identical to oracle/vw_sgm_oracle.c.
//
// Data in HBM: census signatures (8 B / pixel / image), cost volume cost[pixel][d] (uint8), accumulated costs
// accum[pixel][d] (uint16), one more uint16 volume as the scratch of the tie smoothing.  d = dy * (sx + 1) + dx.
//   sgm_u8_kernel       HBM bound   4 B read + 1 B written per pixel
//   sgm_census_kernel   HBM bound   k*k L1/L2 reads, 8 B written per pixel
//   sgm_cost_kernel     HBM bound   1 B written per (pixel, d); signatures come from L2
//   sgm_path_kernel     latency bound: one CTA per scan line, threads = disparities, sequential along the line
//                       (:1014-1141 evaluate_path, SSE flavour: unsigned 16-bit min, saturating add / subtract);
//                       per (pixel, d, direction): 1 B + 2 B read, 2 B written = 40 B over the 8 directions
//   sgm_wta_kernel      HBM bound   2 B read per (pixel, d); tie smoothing (:1196-1288) only for pixels with ties
#include "common.cuh"

namespace vwb200 {

typedef uint8_t cost_t;
typedef uint16_t accum_t;

// ---- vw::u8_convert (Image/ImageThresh.h:274-286, Image/Algorithms.h:106-126) --------------------------------------
__global__ void sgm_u8_kernel(ImgF img, const float* __restrict__ stats /* min, max */, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= img.w || y >= img.h) return;
  double mn = (double)stats[0], mx = (double)stats[1];
  if (mx == mn) mx = __dadd_rn(mn, 1.0);
  const float old_min = (float)mn, old_max = (float)mx;
  const double ratio = (old_max == old_min) ? 0.0 : __ddiv_rn(255.0, (double)__fsub_rn(old_max, old_min));
  float v = img.p[(ptrdiff_t)y * img.pitch + x];
  if (v < old_min) v = old_min;
  if (v > old_max) v = old_max;
  const float n = __double2float_rn(__dadd_rn(__dmul_rn((double)__fsub_rn(v, old_min), ratio), 0.0));
  out[(size_t)y * img.w + x] = (uint8_t)n;
}

// ---- census signatures (Image/CensusTransform.h:64-160) ---------------------------------------------------------------
__constant__ int c_c9_cols[32] = {0, 4, 8, 1, 3, 5, 7, 2, 4, 6, 1, 4, 7, 0, 2, 3, 5, 6, 8, 1, 4, 7, 2, 4, 6, 1, 3, 5, 7, 0, 4, 8};
__constant__ int c_c9_rows[32] = {0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 4, 5, 5, 5, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8};

__global__ void sgm_census_kernel(const uint8_t* __restrict__ img, int w, int h, int k, unsigned long long* __restrict__ out) {
  const int hk = (k - 1) / 2, cw = w - 2 * hk, ch = h - 2 * hk;
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= cw || r >= ch) return;
  const int col = c + hk, row = r + hk;
  const int center = img[(size_t)row * w + col];
  unsigned long long sig = 0, addend = 1;
  if (k == 9) {
    for (int i = 0; i < 32; ++i) {
      if ((int)img[(size_t)(row + c_c9_rows[i] - 4) * w + (col + c_c9_cols[i] - 4)] > center) sig += addend;
      addend *= 2;
    }
  } else {
    for (int rr = row + hk; rr >= row - hk; --rr)
      for (int cc = col + hk; cc >= col - hk; --cc) {
        if (rr == row && cc == col) continue;
        if ((int)img[(size_t)rr * w + cc] > center) sig += addend;
        addend *= 2;
      }
  }
  out[(size_t)r * cw + c] = sig;
}

struct SgmGeom {
  int ndx, ndy, nd;            // disparities dx in [0, ndx), dy in [0, ndy)
  int p1, p2;
  int ow, oh, min_col, min_row;
  int lw, clw, crw, hk;        // left width, census widths, half kernel
};

// ---- Hamming costs (get_hamming_distance_costs, SGM.cc:39-73) ---------------------------------------------------------
__global__ void sgm_cost_kernel(const unsigned long long* __restrict__ lc, const unsigned long long* __restrict__ rc, SgmGeom g,
                                cost_t* __restrict__ cost) {
  const size_t total = (size_t)g.ow * g.oh * g.nd;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % g.nd);
    const size_t pix = i / g.nd;
    const int c = (int)(pix % g.ow), r = (int)(pix / g.ow);
    const int dy = d / g.ndx, dx = d - dy * g.ndx;
    const int br = r + g.min_row - g.hk, bc = c + g.min_col - g.hk;
    cost[i] = (cost_t)__popcll(lc[(size_t)br * g.clw + bc] ^ rc[(size_t)(br + dy) * g.crw + (bc + dx)]);
  }
}

// ---- one scan line per CTA (PixelPassTask, SGMAssist.h:705-815; evaluate_path SSE flavour, SGM.cc:1014-1141) -----------
__device__ __forceinline__ accum_t sat_add16(accum_t a, accum_t b) { const unsigned s = (unsigned)a + b; return (accum_t)(s > 65535u ? 65535u : s); }
__device__ __forceinline__ accum_t sat_sub16(accum_t a, accum_t b) { return (accum_t)(a > b ? a - b : 0); }

__global__ void sgm_path_kernel(const uint8_t* __restrict__ left, const cost_t* __restrict__ cost, accum_t* __restrict__ accum,
                                SgmGeom g, int sc, int sr) {
  extern __shared__ unsigned short sm[];
  accum_t* prior = sm;                  // [nd]
  accum_t* red = sm + g.nd;             // [32] per-warp minima
  // line -> first pixel: the pixels whose predecessor (c - sc, r - sr) lies outside the raster
  int c, r;
  const int line = blockIdx.x;
  if (sr == 0) { r = line; c = sc > 0 ? 0 : g.ow - 1; }
  else if (sc == 0) { c = line; r = sr > 0 ? 0 : g.oh - 1; }
  else if (line < g.ow) { c = line; r = sr > 0 ? 0 : g.oh - 1; }
  else {
    const int j = line - g.ow;          // the remaining oh - 1 rows of the side column
    c = sc > 0 ? 0 : g.ow - 1;
    r = sr > 0 ? 1 + j : j;
  }
  const int d = threadIdx.x;
  const bool act = d < g.nd;
  const int dy = act ? d / g.ndx : 0, dx = act ? d - dy * g.ndx : 0;
  // the eight adjacent disparities, clamped at the search box (populate_adjacent_disp_lookup_table, :755-800)
  const int yl = dy - 1 < 0 ? dy : dy - 1, ym = dy + 1 > g.ndy - 1 ? dy : dy + 1;
  const int xl = dx - 1 < 0 ? dx : dx - 1, xm = dx + 1 > g.ndx - 1 ? dx : dx + 1;
  const int a0 = yl * g.ndx + dx, a1 = dy * g.ndx + xl, a2 = dy * g.ndx + xm, a3 = ym * g.ndx + dx;
  const int a4 = yl * g.ndx + xl, a5 = yl * g.ndx + xm, a6 = ym * g.ndx + xl, a7 = ym * g.ndx + xm;
  const accum_t BAD = (accum_t)(255 + g.p2);                  // get_bad_accum_val (SGM.h:240)
  int last_val = -1;
  accum_t cur = 0;
  // software pipeline: the next pixel's cost / accumulated cost / grey value are requested before the current pixel is
  // evaluated (every step is a dependent chain otherwise: ncu showed 5.5 long-scoreboard stalls per issue)
  bool in = c >= 0 && c < g.ow && r >= 0 && r < g.oh;
  size_t base = in ? ((size_t)r * g.ow + c) * g.nd : 0;
  accum_t local = (in && act) ? (accum_t)cost[base + d] : (accum_t)0;
  accum_t acc_in = (in && act) ? accum[base + d] : (accum_t)0;
  int cur_val = in ? (int)left[(size_t)(r + g.min_row) * g.lw + (c + g.min_col)] : 0;
  while (in) {
    const int cn = c + sc, rn = r + sr;
    const bool in_n = cn >= 0 && cn < g.ow && rn >= 0 && rn < g.oh;
    const size_t base_n = in_n ? ((size_t)rn * g.ow + cn) * g.nd : 0;
    const accum_t local_n = (in_n && act) ? (accum_t)cost[base_n + d] : (accum_t)0;
    const accum_t acc_n = (in_n && act) ? accum[base_n + d] : (accum_t)0;
    const int val_n = in_n ? (int)left[(size_t)(rn + g.min_row) * g.lw + (cn + g.min_col)] : 0;
    if (last_val >= 0) {
      // block minimum of the previous pixel's path costs
      accum_t m = act ? prior[d] : (accum_t)65535;
      unsigned mm = m;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mm = min(mm, __shfl_xor_sync(0xffffffffu, mm, o));
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (accum_t)mm;
      __syncthreads();
      unsigned mp = BAD;
      for (int wdx = 0; wdx < (int)(blockDim.x >> 5); ++wdx) mp = min(mp, (unsigned)red[wdx]);
      const accum_t min_prior = (accum_t)mp;
      const int diff = abs(cur_val - last_val);
      accum_t p2_mod = (accum_t)g.p2;
      if (diff > 0) p2_mod = (accum_t)(p2_mod / diff);
      if (p2_mod < g.p1) p2_mod = (accum_t)g.p1;
      const accum_t dJ = (accum_t)(min_prior + p2_mod);
      if (act) {
        unsigned adj = min(min(min((unsigned)prior[a0], (unsigned)prior[a1]), min((unsigned)prior[a2], (unsigned)prior[a3])),
                           min(min((unsigned)prior[a4], (unsigned)prior[a5]), min((unsigned)prior[a6], (unsigned)prior[a7])));
        accum_t res = sat_add16((accum_t)adj, (accum_t)g.p1);
        res = (accum_t)min((unsigned)res, min((unsigned)prior[d], (unsigned)dJ));
        res = sat_add16(res, local);
        cur = sat_sub16(res, min_prior);
      }
      __syncthreads();                    // everyone has read prior[]
    } else {
      cur = local;                        // first pixel of the line (SGMAssist.h:756-759)
    }
    if (act) {
      prior[d] = cur;
      accum[base + d] = (accum_t)(acc_in + cur);                 // update_accum_buffer (SGMAssist.h:806-809), uint16 wrap
    }
    __syncthreads();
    last_val = cur_val;
    c = cn; r = rn; in = in_n; base = base_n; local = local_n; acc_in = acc_n; cur_val = val_n;
  }
}

// ---- create_disparity_view / select_best_disparity (SGM.cc:1159-1346) --------------------------------------------------
__global__ void sgm_wta_kernel(accum_t* __restrict__ accum, accum_t* __restrict__ scratch, SgmGeom g, vwb200_dispi* __restrict__ out,
                               ptrdiff_t opitch) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (size_t)g.ow * g.oh) return;
  accum_t* accum_vec = accum + pix * g.nd;
  accum_t* buffer = scratch + pix * g.nd;
  const int width = g.ndx, height = g.ndy, num = g.nd;
  int min_count = 0, min_index = 0;
  unsigned min_val = 65535;
  for (int i = 0; i < num; ++i) {
    const unsigned v = accum_vec[i];
    buffer[i] = (accum_t)v;
    if (v == min_val) ++min_count;
    if (v < min_val) { min_index = i; min_val = v; min_count = 1; }
  }
  accum_t* input_array = accum_vec;
  accum_t* output_array = buffer;
  const double third = 1.0 / 3.0;
  int iter_count = 0, index = 0;
  while (min_count > 1) {
    accum_t* sw = input_array; input_array = output_array; output_array = sw;
    index = 0; min_count = 0; min_val = 65535; min_index = 0;
    for (int row = 0; row < height; ++row)
      for (int col = 0; col < width; ++col) {
        int mn = -1, mx = 1;
        double result = 0.0, weight_total = 0.0;
        if (iter_count < 5) {
          if (mn + col < 0) mn = 0;
          if (mx + col >= width) mx = 0;
          for (int k = mn; k <= mx; ++k) { result = __dadd_rn(result, __dmul_rn((double)input_array[index + k], third)); weight_total = __dadd_rn(weight_total, third); }
        } else {
          if (mn + row < 0) mn = 0;
          if (mx + row >= height) mx = 0;
          for (int k = mn; k <= mx; ++k) { result = __dadd_rn(result, __dmul_rn((double)input_array[index + k * width], third)); weight_total = __dadd_rn(weight_total, third); }
        }
        const unsigned v = (unsigned)(accum_t)round(__ddiv_rn(result, weight_total));
        if (v == min_val) ++min_count;
        if (v < min_val) { min_index = index; min_val = v; min_count = 1; }
        output_array[index] = (accum_t)v;
        ++index;
      }
    if (++iter_count >= 6) break;
  }
  if (iter_count > 0 && iter_count % 2 == 0)
    for (int i = 0; i < index; ++i) input_array[i] = output_array[i];
  vwb200_dispi o;
  o.dy = min_index / width; o.dx = min_index - o.dy * width; o.valid = 1;        // disp_index_to_xy (:2737-2745)
  out[(ptrdiff_t)(pix / g.ow) * opitch + (pix % g.ow)] = o;
}

// ---- create_disparity_view_subpixel (SGM.cc:1402-1480, 1497-1614): the 1-D models -----------------------------------
__device__ __forceinline__ double sgm_fit(double x, int mode) {
  const double PI = 3.14159265359;
  const double lin = __ddiv_rn(x, 2.0);
  if (mode == 3) return __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(x, x), x), x), x), 4.0);      // poly4Fit
  const double cf = __dsub_rn(1.0, cos(__ddiv_rn(__dmul_rn(x, PI), 3.0)));                               // cosFit
  if (mode == 4) return cf;
  if (mode == 5) {                                                                                        // lcBlendFit
    const double factor = __dsub_rn(1.195, cos(__dmul_rn(x, PI / 2.3)));
    return __dadd_rn(__dmul_rn(cf, factor), __dmul_rn(lin, __dsub_rn(1.0, factor)));
  }
  return lin;
}
__device__ __forceinline__ double sgm_subpixel_offset(int prev, int center, int next, bool left_bound, bool right_bound, int mode) {
  const double ld = (double)(prev - center), rd = (double)(next - center);
  if (rd == 0 && ld == 0) return 0.0;
  if (left_bound) return __dmul_rn(0.5, __ddiv_rn((double)center, (double)next));                  // two_value_subpixel
  if (right_bound) return __dmul_rn(-1.0, __dmul_rn(0.5, __ddiv_rn((double)center, (double)prev)));
  double x = __ddiv_rn(rd, ld), mult = -1.0;
  if (ld < rd) { x = __ddiv_rn(ld, rd); mult = 1.0; }
  return __dmul_rn(__dsub_rn(sgm_fit(x, mode), 0.5), mult);
}
__global__ void sgm_subpixel_kernel(const accum_t* __restrict__ accum, const vwb200_dispi* __restrict__ disp, ptrdiff_t dpitch, SgmGeom g,
                                    int mode, float* __restrict__ out, ptrdiff_t opitch /* floats */) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (size_t)g.ow * g.oh) return;
  const int i = (int)(pix % g.ow), j = (int)(pix / g.ow);
  const vwb200_dispi d = disp[(ptrdiff_t)j * dpitch + i];
  float* o = out + (ptrdiff_t)j * opitch + 3 * i;
  o[2] = 1.0f;
  if (mode == 0) { o[0] = (float)d.dx; o[1] = (float)d.dy; return; }
  const int width = g.ndx, min_index = d.dy * width + d.dx;
  int x_left = -1, x_right = 1, y_up = -width, y_down = width;
  bool lb = false, rb = false, tb = false, bb = false;
  if (d.dx == 0) { x_left = 0; lb = true; }
  if (d.dx == g.ndx - 1) { x_right = 0; rb = true; }
  if (d.dy == 0) { y_up = 0; tb = true; }
  if (d.dy == g.ndy - 1) { y_down = 0; bb = true; }
  const accum_t* av = accum + pix * g.nd;
  const double ddx = sgm_subpixel_offset(av[min_index + x_left], av[min_index], av[min_index + x_right], lb, rb, mode);
  const double ddy = sgm_subpixel_offset(av[min_index + y_up], av[min_index], av[min_index + y_down], tb, bb, mode);
  o[0] = (float)__dadd_rn((double)d.dx, ddx);
  o[1] = (float)__dadd_rn((double)d.dy, ddy);
}

// ---- host side -----------------------------------------------------------------------------------------------------
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

int sgm_output_size(int lw, int lh, int rw, int rh, int sx, int sy, int k, int* ow, int* oh) {
  const int hk = (k - 1) / 2;                      // semi_global_matching_func (:2397-2420) with min_disp = 0
  const int min_row = hk, min_col = hk;
  int max_row = std::min(lh - 1 - hk, rh - 1 - (hk + sy));
  int max_col = std::min(lw - 1 - hk, rw - 1 - (hk + sx));
  max_row = std::min(max_row, lh - 1); max_col = std::min(max_col, lw - 1);
  *ow = std::max(0, max_col - min_col + 1); *oh = std::max(0, max_row - min_row + 1);
  return VWB200_OK;
}

size_t sgm_workspace_bytes(int lw, int lh, int rw, int rh, int sx, int sy, int k) {
  int ow, oh;
  sgm_output_size(lw, lh, rw, rh, sx, sy, k, &ow, &oh);
  const size_t nd = (size_t)(sx + 1) * (sy + 1), vol = (size_t)ow * oh * nd;
  return al256((size_t)lw * lh) + al256((size_t)rw * rh) + al256((size_t)lw * lh * 8) + al256((size_t)rw * rh * 8) + al256(vol) + 2 * al256(vol * 2) +
         al256(64) + 1024;
}

int sgm_launch(ImgF left, ImgF right, int sx, int sy, int k, int p1, int p2, vwb200_dispi* out, ptrdiff_t opitch, void* workspace,
               cudaStream_t st, int subpixel_mode, float* out_sub, ptrdiff_t sub_pitch) {
  if (k != 3 && k != 5 && k != 7 && k != 9) {
    set_error("Census transforms are only available in size 3, 5, 7, and 9.");       // SGM.cc:1885-1888
    return VWB200_ENOIMPL;
  }
  if (sx < 0 || sy < 0) { set_error("sgm: negative search volume"); return VWB200_EARG; }
  const long long nd = (long long)(sx + 1) * (sy + 1);
  if (nd > 1024) { set_error("sgm: %lld disparities per pixel exceed the 1024 this kernel handles", nd); return VWB200_ENOIMPL; }
  if (p1 <= 0) p1 = k == 3 ? 3 : k == 5 ? 15 : k == 7 ? 30 : 20;                       // set_parameters (:112-122)
  if (p2 <= 0) p2 = k == 3 ? 70 : k == 5 ? 750 : k == 7 ? 1500 : 1000;                 // (:141-151)
  SgmGeom g;
  g.ndx = sx + 1; g.ndy = sy + 1; g.nd = (int)nd; g.p1 = p1; g.p2 = p2;
  g.hk = (k - 1) / 2; g.min_col = g.hk; g.min_row = g.hk; g.lw = left.w;
  sgm_output_size(left.w, left.h, right.w, right.h, sx, sy, k, &g.ow, &g.oh);
  if (g.ow <= 0 || g.oh <= 0) return VWB200_OK;
  g.clw = left.w - 2 * g.hk; g.crw = right.w - 2 * g.hk;
  // carve the workspace
  unsigned char* p = static_cast<unsigned char*>(workspace);
  auto take = [&](size_t bytes) { unsigned char* q = p; p += al256(bytes); return q; };
  const size_t vol = (size_t)g.ow * g.oh * g.nd;
  uint8_t* l8 = take((size_t)left.w * left.h);
  uint8_t* r8 = take((size_t)right.w * right.h);
  unsigned long long* lc = (unsigned long long*)take((size_t)left.w * left.h * 8);
  unsigned long long* rc = (unsigned long long*)take((size_t)right.w * right.h * 8);
  cost_t* cost = take(vol);
  accum_t* accum = (accum_t*)take(vol * 2);
  accum_t* scratch = (accum_t*)take(vol * 2);
  float* stats = (float*)take(64);
  dim3 b(32, 8);
  VWB_TRY(image_stats_launch(left, stats, st));
  VWB_TRY(image_stats_launch(right, stats + 3, st));
  sgm_u8_kernel<<<dim3((left.w + 31) / 32, (left.h + 7) / 8), b, 0, st>>>(left, stats, l8);
  VWB_LAUNCH_CHECK();
  sgm_u8_kernel<<<dim3((right.w + 31) / 32, (right.h + 7) / 8), b, 0, st>>>(right, stats + 3, r8);
  VWB_LAUNCH_CHECK();
  sgm_census_kernel<<<dim3((g.clw + 31) / 32, (left.h - 2 * g.hk + 7) / 8), b, 0, st>>>(l8, left.w, left.h, k, lc);
  VWB_LAUNCH_CHECK();
  sgm_census_kernel<<<dim3((g.crw + 31) / 32, (right.h - 2 * g.hk + 7) / 8), b, 0, st>>>(r8, right.w, right.h, k, rc);
  VWB_LAUNCH_CHECK();
  sgm_cost_kernel<<<148 * 8, 256, 0, st>>>(lc, rc, g, cost);
  VWB_LAUNCH_CHECK();
  VWB_CUDA(cudaMemsetAsync(accum, 0, vol * 2, st));
  // the eight directions of accum_sgm_multithread (:2462-2611), one launch each (a pixel lies on one line per direction)
  static const int DIRS[8][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};
  const int threads = ((g.nd + 31) / 32) * 32;
  const size_t smem = (size_t)(g.nd + 32) * sizeof(accum_t);
  for (int i = 0; i < 8; ++i) {
    const int sc = DIRS[i][0], sr = DIRS[i][1];
    const int lines = sr == 0 ? g.oh : (sc == 0 ? g.ow : g.ow + g.oh - 1);
    sgm_path_kernel<<<lines, threads, smem, st>>>(l8, cost, accum, g, sc, sr);
    VWB_LAUNCH_CHECK();
  }
  const size_t npix = (size_t)g.ow * g.oh;
  sgm_wta_kernel<<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(accum, scratch, g, out, opitch);
  VWB_LAUNCH_CHECK();
  if (out_sub) {
    sgm_subpixel_kernel<<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(accum, out, opitch, g, subpixel_mode, out_sub, sub_pitch);
    VWB_LAUNCH_CHECK();
  }
  return VWB200_OK;
}

}  // namespace vwb200
