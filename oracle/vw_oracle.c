/*
 * vw_oracle.c -- CPU restatement ("oracle") of the Vision Workbench stereo-correlation
 * hot path.  TEST INFRASTRUCTURE ONLY: see vw_oracle.h.  Never linked into the product.
 *
 * Parity status: pinned by the reference's own known-answer tests
 * (tests/test_oracle_kat.py): Stereo/tests/TestCorrelation.cxx:30-66,
 * TestAlgorithms.cxx:45-150, TestCostFunctions.cxx:36-79, TestCorrelate.cxx:29-55,
 * Image/tests/TestConvolution.cxx:131-197, TestFilter.cxx:45-73.  The reference itself
 * cannot be built here (Boost/GDAL/LAPACK headers absent), so everything finer than those
 * vectors rests on this file following the cited lines loop for loop.
 *
 * Build: gcc -O3 -std=c11 -msse4.1 -ffp-contract=off -fopenmp -shared -fPIC
 * (-ffp-contract=off: the reference is an SSE4.1 build, no FMA contraction.)
 */
#include "vw_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ boxes */
/* vw::BBox2i semantics, Math/BBox.tcc:37-285 */
typedef struct { int x0, y0, x1, y1; } box_t;
#define BOX_BIG (INT_MAX - 1)
static box_t box_default(void) { box_t b = { BOX_BIG, BOX_BIG, -BOX_BIG, -BOX_BIG }; return b; }
static box_t box_xywh(int x, int y, int w, int h) { box_t b = { x, y, x + w, y + h }; return b; }
static int box_empty(box_t b) { return b.x0 >= b.x1 || b.y0 >= b.y1; }             /* :156-160 */
static int box_w(box_t b) { return box_empty(b) ? 0 : b.x1 - b.x0; }               /* :170-175 */
static int box_h(box_t b) { return box_empty(b) ? 0 : b.y1 - b.y0; }
static int box_area(box_t b) { return box_empty(b) ? 0 : (b.x1 - b.x0) * (b.y1 - b.y0); }
static int box_eq(box_t a, box_t b) { return a.x0 == b.x0 && a.y0 == b.y0 && a.x1 == b.x1 && a.y1 == b.y1; }
static box_t box_expand(box_t b, int ex, int ey) {                                  /* :228-247 */
  if (box_empty(b)) return b;
  b.x0 -= ex; b.y0 -= ey; b.x1 += ex; b.y1 += ey; return b;
}
static box_t box_shift(box_t b, int sx, int sy) {                                   /* :272-279 */
  if (box_empty(b)) return b;
  b.x0 += sx; b.x1 += sx; b.y0 += sy; b.y1 += sy; return b;
}
static box_t box_scale(box_t b, int s) {                                            /* :256-263 */
  if (box_empty(b)) return b;
  b.x0 *= s; b.y0 *= s; b.x1 *= s; b.y1 *= s; return b;
}
static box_t box_crop(box_t b, box_t c) {                                           /* :110-120 */
  if (b.x0 < c.x0) b.x0 = c.x0;
  if (b.x1 > c.x1) b.x1 = c.x1;
  if (b.y0 < c.y0) b.y0 = c.y0;
  if (b.y1 > c.y1) b.y1 = c.y1;
  return b;
}
static box_t box_grow_pt(box_t b, int x, int y) {                                   /* :83-98 */
  if (x > b.x1) b.x1 = x;
  if (x < b.x0) b.x0 = x;
  if (y > b.y1) b.y1 = y;
  if (y < b.y0) b.y0 = y;
  return b;
}
static box_t box_grow(box_t b, box_t g) {                                           /* :100-106 */
  if (box_empty(g)) return b;
  b = box_grow_pt(b, g.x0, g.y0);
  b = box_grow_pt(b, g.x1, g.y1);
  return b;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------ box sum */
/* Stereo/Algorithms.h:41-129.  Identical operation order: column sums accumulated row by
 * row; per output row a left-to-right seed, then row_sum += (front - back); afterwards
 * col += front_row; col -= back_row as two statements. */
int vwo_fast_box_sum(const double* in, int w, int h, int pitch, int kx, int ky, double* out) {
  if (kx % 2 != 1 || ky % 2 != 1) return -1;          /* :45-46 */
  if (w < kx || h < ky) return -2;
  int ow = w - kx + 1, oh = h - ky + 1;
  double* col_sum = (double*)calloc((size_t)w, sizeof(double));
  if (!col_sum) return -3;
  for (int j = 0; j < ky; ++j)                          /* :62-75 */
    for (int x = 0; x < w; ++x) col_sum[x] += in[(size_t)j * pitch + x];
  double* dst = out;
  for (int y = 0; y < oh; ++y) {
    double row_sum = 0;                                 /* :82-84 */
    for (int i = 0; i < kx; ++i) row_sum = row_sum + col_sum[i];
    const double *cback = col_sum, *cfront = col_sum + kx;
    const double* cend = col_sum + w;
    while (cfront != cend) {                            /* :88-93 */
      *dst++ = row_sum;
      row_sum += *cfront++ - *cback++;
    }
    *dst++ = row_sum;
    if (y != oh - 1) {                                  /* :97-103 */
      const double* back = in + (size_t)y * pitch;
      const double* front = in + (size_t)(y + ky) * pitch;
      for (int x = 0; x < w; ++x) {
        col_sum[x] += front[x];
        col_sum[x] -= back[x];
      }
    }
  }
  (void)ow;
  free(col_sum);
  return 0;
}

/* ------------------------------------------------------------------ calc_disparity */
/* Per-pixel functors, Stereo/CostFunctions.h:72-141: evaluated in FLOAT
 * (Core/CompoundTypes.h:131-136 dispatches on the channel type), then widened. */
static inline double cost_pixel(int cost, float a, float b) {
  switch (cost) {
    case VWO_COST_SQ:  { float d = a - b; float s = d * d; return (double)s; }       /* :90-98 */
    case VWO_COST_NCC: { float s = a * b; return (double)s; }                        /* :115-123 */
    default:           { float d = a - b; return (double)fabsf(d); }                 /* :72-88 */
  }
}
static inline int better(int cost, double c, double q) {                             /* :173-176,199-201,233-235 */
  return cost == VWO_COST_NCC ? (c > q) : (c < q);
}

/* Stereo/Correlation.cc:33-137 (best_of_search_convolution) via :330-375 (calc_disparity). */
int vwo_calc_disparity(int cost, const float* left, int lw, int lh, int lpitch,
                       const float* right, int rw, int rh, int rpitch,
                       int sx, int sy, int kx, int ky, vwo_disp_t* out) {
  if (kx % 2 != 1 || ky % 2 != 1) return -1;
  if (sx <= 0 || sy <= 0) return -2;
  if (lw < kx || lh < ky) return -3;
  if (rw < lw + sx - 1 || rh < lh + sy - 1) return -4;
  /* calc_disparity crops the right raster to region + search_volume - 1  (:355-359) */
  rw = lw + sx - 1; rh = lh + sy - 1;
  const int W = lw - kx + 1, H = lh - ky + 1;          /* :50 */
  const size_t N = (size_t)W * H;
  double* best  = (double*)malloc(N * sizeof(double));
  double* worst = (double*)malloc(N * sizeof(double));
  double* metric = (double*)malloc(N * sizeof(double));
  double* applied = (double*)malloc((size_t)lw * lh * sizeof(double));
  double *lprec = NULL, *rprec = NULL;
  int rpw = 0;
  if (!best || !worst || !metric || !applied) return -5;
  for (size_t i = 0; i < N; ++i) { out[i].dx = 0; out[i].dy = 0; out[i].valid = 1; }  /* :51-53 */

  if (cost == VWO_COST_NCC) {                          /* CostFunctions.h:214-219 */
    /* left_precision = 1.0 / box(square(left)); square() is float v*v (Math/Functors.h:316-321) */
    lprec = (double*)malloc(N * sizeof(double));
    double* sq = (double*)malloc((size_t)lw * lh * sizeof(double));
    for (int y = 0; y < lh; ++y)
      for (int x = 0; x < lw; ++x) { float v = left[(size_t)y * lpitch + x]; float s = v * v; sq[(size_t)y * lw + x] = (double)s; }
    vwo_fast_box_sum(sq, lw, lh, lw, kx, ky, lprec);
    for (size_t i = 0; i < N; ++i) lprec[i] = 1.0 / lprec[i];
    free(sq);
    /* right_precision over the WHOLE right raster */
    rpw = rw - kx + 1; int rph = rh - ky + 1;
    rprec = (double*)malloc((size_t)rpw * rph * sizeof(double));
    sq = (double*)malloc((size_t)rw * rh * sizeof(double));
    for (int y = 0; y < rh; ++y)
      for (int x = 0; x < rw; ++x) { float v = right[(size_t)y * rpitch + x]; float s = v * v; sq[(size_t)y * rw + x] = (double)s; }
    vwo_fast_box_sum(sq, rw, rh, rw, kx, ky, rprec);
    for (size_t i = 0; i < (size_t)rpw * rph; ++i) rprec[i] = 1.0 / rprec[i];
    free(sq);
  }

  for (int dy = 0; dy != sy; ++dy) {                   /* :64 */
    for (int dx = 0; dx != sx; ++dx) {                 /* :65 */
      /* right_raster_crop = crop(right, bbox(left)+d); cost_applied = cost(left,crop)  :79-80 */
      for (int y = 0; y < lh; ++y) {
        const float* lrow = left + (size_t)y * lpitch;
        const float* rrow = right + (size_t)(y + dy) * rpitch + dx;
        double* arow = applied + (size_t)y * lw;
        for (int x = 0; x < lw; ++x) arow[x] = cost_pixel(cost, lrow[x], rrow[x]);
      }
      vwo_fast_box_sum(applied, lw, lh, lw, kx, ky, metric);            /* :81 */
      if (cost == VWO_COST_NCC) {                                         /* :82, CostFunctions.h:227-231 */
        for (int y = 0; y < H; ++y)
          for (int x = 0; x < W; ++x) {
            double t = lprec[(size_t)y * W + x] * rprec[(size_t)(y + dy) * rpw + (x + dx)];
            metric[(size_t)y * W + x] *= sqrt(t);
          }
      }
      if (dx != 0 || dy != 0) {                                           /* :97-109 */
        for (size_t i = 0; i < N; ++i) {
          double c = metric[i];
          if (better(cost, c, best[i])) { best[i] = c; out[i].dx = dx; out[i].dy = dy; }
          else if (!better(cost, c, worst[i])) { worst[i] = c; }
        }
      } else {                                                            /* :110-117 */
        for (size_t i = 0; i < N; ++i) best[i] = worst[i] = metric[i];
      }
    }
  }
  for (size_t i = 0; i < N; ++i)                                          /* :121-133 */
    if (best[i] == worst[i]) out[i].valid = 0;
  free(best); free(worst); free(metric); free(applied); free(lprec); free(rprec);
  return 0;
}

/* ------------------------------------------------------------------ convolution */
/* Image/Convolution.h:275-328 with correlate_1d_at_point :51-66.  Row pass into a float
 * work image of bbox.width x child_bbox.height, then column pass.  Kernel applied via
 * rbegin() (true convolution).  Accumulation in float, in order, starting from 0. */
int vwo_separable_convolve_c(const float* in, int w, int h, int pitch,
                             const float* kxv, int nx, const float* kyv, int ny,
                             int cx, int cy, int edge_zero, float* out) {
  if (w <= 0 || h <= 0) return -1;
  if (cx < 0) cx = nx ? (nx - 1) / 2 : 0;                           /* :223 */
  if (cy < 0) cy = ny ? (ny - 1) / 2 : 0;
  /* child bbox grows by (n-c-1) on the min side and c on the max side  :281-283 */
  int padx0 = nx ? (nx - cx - 1) : 0, pady0 = ny ? (ny - cy - 1) : 0;
  int pady1 = ny ? cy : 0;
  int ch = h + pady0 + pady1;
  float* work = (float*)malloc((size_t)w * ch * sizeof(float));
  if (!work) return -2;
  /* row pass over every child row */
  for (int yy = 0; yy < ch; ++yy) {
    int sy = yy - pady0;
    int sy_c = clampi(sy, 0, h - 1);
    int row_out = edge_zero && (sy < 0 || sy >= h);
    for (int x = 0; x < w; ++x) {
      float result;
      if (nx) {
        result = 0.0f;
        for (int i = 0; i < nx; ++i) {
          int sxp = x - padx0 + i;
          float s;
          if (edge_zero) s = (row_out || sxp < 0 || sxp >= w) ? 0.0f : in[(size_t)sy_c * pitch + sxp];
          else s = in[(size_t)sy_c * pitch + clampi(sxp, 0, w - 1)];
          float prod = kxv[nx - 1 - i] * s;
          result = result + prod;
        }
      } else {
        result = row_out ? 0.0f : in[(size_t)sy_c * pitch + x];
      }
      work[(size_t)yy * w + x] = result;
    }
  }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float result;
      if (ny) {
        result = 0.0f;
        for (int j = 0; j < ny; ++j) {
          float prod = kyv[ny - 1 - j] * work[(size_t)(y + j) * w + x];
          result = result + prod;
        }
      } else result = work[(size_t)(y + pady0) * w + x];
      out[(size_t)y * w + x] = result;
    }
  free(work);
  return 0;
}

int vwo_separable_convolve(const float* in, int w, int h, int pitch,
                           const float* kxv, int nx, const float* kyv, int ny,
                           int edge_zero, float* out) {
  return vwo_separable_convolve_c(in, w, h, pitch, kxv, nx, kyv, ny, -1, -1, edge_zero, out);
}

double vwo_cost_pixel(int cost, float a, float b) { return cost_pixel(cost, a, b); }

int vwo_subsample2_f32(const float* in, int w, int h, int pitch, float* out) {   /* Manipulation.h:238-251 */
  int ow = 1 + (w - 1) / 2, oh = 1 + (h - 1) / 2;
  for (int j = 0; j < oh; ++j)
    for (int i = 0; i < ow; ++i) out[(size_t)j * ow + i] = in[(size_t)(2 * j) * pitch + 2 * i];
  return 0;
}

/* Stereo/CorrelationView.cc:38-63 over a ZeroEdgeExtension (PerPixelAccessorViews.h:131-135) */
int vwo_subsample_mask_by_two(const uint8_t* in, int w, int h, uint8_t* out) {
  int ow = 1 + (w - 1) / 2, oh = 1 + (h - 1) / 2;
  for (int j = 0; j < oh; ++j)
    for (int i = 0; i < ow; ++i) {
      int count = 0;
      for (int b = 0; b < 2; ++b)
        for (int a = 0; a < 2; ++a) {
          int x = 2 * i + a, y = 2 * j + b;
          if (x < w && y < h && in[(size_t)y * w + x]) count++;
        }
      out[(size_t)j * ow + i] = count > 1 ? 255 : 0;
    }
  return 0;
}

int vwo_pyramid_down(const float* in, int w, int h, float* out) {                /* CorrelationView.cc:210-214 */
  /* Image/Filter.h:89-99: float taps 1/16, 4/16, 6/16 */
  float k[5];
  k[0] = k[4] = (float)(1.0 / 16.0); k[1] = k[3] = (float)(4.0 / 16.0); k[2] = (float)(6.0 / 16.0);
  float* tmp = (float*)malloc((size_t)w * h * sizeof(float));
  if (!tmp) return -1;
  int rc = vwo_separable_convolve(in, w, h, w, k, 5, k, 5, 0, tmp);
  if (rc == 0) vwo_subsample2_f32(tmp, w, h, w, out);
  free(tmp);
  return rc;
}

/* ------------------------------------------------------------------ consistency check */
/* Stereo/Correlate.cc:1441-1502 */
/* diff (optional): PixelMask<float> pairs, row pitch diff_pitch pixels; pixel (c + offx, r + offy) receives disp_diff */
static int consistency_check_diff(vwo_disp_t* l2r, int lw, int lh, int lpitch, const vwo_disp_t* r2l, int rw, int rh, float threshold,
                                  float* diff, int diff_pitch, int offx, int offy) {
  for (int r = 0; r < lh; ++r)
    for (int c = 0; c < lw; ++c) {
      vwo_disp_t* p = l2r + (size_t)r * lpitch + c;
      int x = c + p->dx, y = r + p->dy;
      if (x < 0 || x >= rw || y < 0 || y >= rh) { p->valid = 0; continue; }
      const vwo_disp_t* q = r2l + (size_t)y * rw + x;
      if (!p->valid || !q->valid) { p->valid = 0; continue; }
      /* std::max(fabs(int+int), fabs(int+int)) -> float disp_diff */
      double a = fabs((double)(p->dx + q->dx)), b = fabs((double)(p->dy + q->dy));
      float dd = (float)(a > b ? a : b);
      if (!(threshold >= dd)) p->valid = 0;
      else if (diff) { float* o = diff + ((size_t)(r + offy) * diff_pitch + (c + offx)) * 2; o[0] = dd; o[1] = 1.0f; }
    }
  return 0;
}
int vwo_cross_corr_consistency_check(vwo_disp_t* l2r, int lw, int lh, int lpitch,
                                     const vwo_disp_t* r2l, int rw, int rh, float threshold) {
  return consistency_check_diff(l2r, lw, lh, lpitch, r2l, rw, rh, threshold, NULL, 0, 0, 0);
}

/* disparity_blob_filter: labels = 8-connected components of the valid pixels (the two-pass labelling of blob::BlobIndex,
 * Image/BlobIndex.h:113-305, yields exactly the connected components); blobs of size <= area are eroded. */
static int uf_find(int* parent, int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; }
int vwo_disparity_blob_filter(vwo_disp_t* d, int w, int h, int area) {
  if (area < 1 || w <= 0 || h <= 0) return 0;                                     /* CorrelationView.cc:249-250 */
  const size_t n = (size_t)w * h;
  int* parent = (int*)malloc(n * sizeof(int));
  int* size = (int*)calloc(n, sizeof(int));
  for (size_t i = 0; i < n; ++i) parent[i] = (int)i;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      if (!d[(size_t)y * w + x].valid) continue;
      const int i = y * w + x;
      static const int NX[4] = {-1, -1, 0, 1}, NY[4] = {0, -1, -1, -1};
      for (int k = 0; k < 4; ++k) {
        const int xx = x + NX[k], yy = y + NY[k];
        if (xx < 0 || xx >= w || yy < 0) continue;
        if (!d[(size_t)yy * w + xx].valid) continue;
        const int a = uf_find(parent, i), b = uf_find(parent, yy * w + xx);
        if (a != b) parent[a > b ? a : b] = a > b ? b : a;
      }
    }
  for (size_t i = 0; i < n; ++i) if (d[i].valid) size[uf_find(parent, (int)i)]++;
  for (size_t i = 0; i < n; ++i)
    if (d[i].valid && size[uf_find(parent, (int)i)] <= area) { d[i].dx = 0; d[i].dy = 0; d[i].valid = 0; }   /* result_type() */
  free(parent); free(size);
  return 0;
}

/* ------------------------------------------------------------------ outlier filters */
/* RmOutliersUsingThreshFunc applied at (cx,cy) of the ConstantEdgeExtension of `in`
 * (Stereo/DisparityMap.h:359-386); cx,cy may lie outside the image (second-pass quirk). */
static vwo_disp_t rm_outliers_at(const vwo_disp_t* in, int w, int h, int cx, int cy,
                                 int hx, int hy, double pt, double rt) {
  const vwo_disp_t c = in[(size_t)clampi(cy, 0, h - 1) * w + clampi(cx, 0, w - 1)];
  if (c.valid) {
    int matched = 0, total = 0;
    for (int yk = -hy; yk <= hy; ++yk)
      for (int xk = -hx; xk <= hx; ++xk) {
        const vwo_disp_t n = in[(size_t)clampi(cy + yk, 0, h - 1) * w + clampi(cx + xk, 0, w - 1)];
        if (n.valid && fabs((double)(c.dx - n.dx)) <= pt && fabs((double)(c.dy - n.dy)) <= pt) matched++;
        total++;
      }
    if (((double)matched / (double)total) < rt) { vwo_disp_t z = { 0, 0, 0 }; return z; }
  }
  return c;
}

int vwo_rm_outliers_using_thresh(const vwo_disp_t* in, int w, int h, int hx, int hy,
                                 double pt, double rt, vwo_disp_t* out) {
  if (hx <= 0 || hy <= 0) return -1;                   /* DisparityMap.h:345-346 */
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) out[(size_t)y * w + x] = rm_outliers_at(in, w, h, x, y, hx, hy, pt, rt);
  return 0;
}

/* Stereo/DisparityMap.h:426-442: outer (1,1,3.0,0.20) pass over the inner view WITHOUT its own
 * edge extension: border neighbours are first-pass results evaluated at out-of-image
 * coordinates (Image/PerPixelAccessorViews.h:63-85). */
int vwo_disparity_cleanup_using_thresh(const vwo_disp_t* in, int w, int h, int hx, int hy,
                                       double pt, double rt, vwo_disp_t* out) {
  if (hx <= 0 || hy <= 0) return -1;
  int pw = w + 2, ph = h + 2;
  vwo_disp_t* p1 = (vwo_disp_t*)malloc((size_t)pw * ph * sizeof(vwo_disp_t));
  if (!p1) return -2;
  for (int y = -1; y <= h; ++y)
    for (int x = -1; x <= w; ++x)
      p1[(size_t)(y + 1) * pw + (x + 1)] = rm_outliers_at(in, w, h, x, y, hx, hy, pt, rt);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const vwo_disp_t c = p1[(size_t)(y + 1) * pw + (x + 1)];
      vwo_disp_t r = c;
      if (c.valid) {
        int matched = 0, total = 0;
        for (int yk = -1; yk <= 1; ++yk)
          for (int xk = -1; xk <= 1; ++xk) {
            const vwo_disp_t n = p1[(size_t)(y + 1 + yk) * pw + (x + 1 + xk)];
            if (n.valid && fabs((double)(c.dx - n.dx)) <= 3.0 && fabs((double)(c.dy - n.dy)) <= 3.0) matched++;
            total++;
          }
        if (((double)matched / (double)total) < 0.20) { r.dx = 0; r.dy = 0; r.valid = 0; }
      }
      out[(size_t)y * w + x] = r;
    }
  free(p1);
  return 0;
}

/* Stereo/DisparityMap.h:142-162 */
int vwo_disparity_mask(const vwo_disp_t* in, int w, int h, const uint8_t* lmask,
                       const uint8_t* rmask, int rmw, int rmh, vwo_disp_t* out) {
  const vwo_disp_t z = { 0, 0, 0 };
  for (int j = 0; j < h; ++j)
    for (int i = 0; i < w; ++i) {
      size_t k = (size_t)j * w + i;
      if (!lmask[k]) { out[k] = z; continue; }
      vwo_disp_t d = in[k];
      if (!d.valid) { out[k] = z; continue; }
      int tx = i + d.dx, ty = j + d.dy;
      if (tx < 0 || tx >= rmw || ty < 0 || ty >= rmh || rmask[(size_t)ty * rmw + tx] == 0) { out[k] = z; continue; }
      out[k] = d;
    }
  return 0;
}

/* ------------------------------------------------------------------ subdivide_regions */
typedef struct { box_t img, disp; } zone_t;
typedef struct { zone_t* z; int n, cap; } zlist_t;
static void zpush(zlist_t* l, box_t img, box_t disp) {
  if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 64; l->z = (zone_t*)realloc(l->z, (size_t)l->cap * sizeof(zone_t)); }
  l->z[l->n].img = img; l->z[l->n].disp = disp; l->n++;
}
/* min/max over valid pixels of `b` (PixelAccumulator<EWMinMaxAccumulator>, Image/Statistics.h:283-290) */
static int minmax_valid(const vwo_disp_t* d, int w, box_t b, box_t* search) {
  int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN, any = 0;
  for (int y = b.y0; y < b.y1; ++y)
    for (int x = b.x0; x < b.x1; ++x) {
      const vwo_disp_t p = d[(size_t)y * w + x];
      if (!p.valid) continue;
      any = 1;
      if (p.dx < mnx) mnx = p.dx;
      if (p.dx > mxx) mxx = p.dx;
      if (p.dy < mny) mny = p.dy;
      if (p.dy > mxy) mxy = p.dy;
    }
  if (any) { search->x0 = mnx; search->y0 = mny; search->x1 = mxx + 1; search->y1 = mxy + 1; }
  return any;
}
static inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

/* Stereo/Correlation.cc:139-328 */
static int subdivide(const vwo_disp_t* disp, int w, int h, box_t cur, zlist_t* list,
                     int kx, int ky, int fail_count) {
  const int MIN_REGION_SIZE = 16;
  int cw = cur.x1 - cur.x0, chh = cur.y1 - cur.y0;     /* size() is max-min (BBox.tcc:199-204) */
  if (cw * chh <= 200 || box_w(cur) < MIN_REGION_SIZE || box_h(cur) < MIN_REGION_SIZE) {   /* :150-163 */
    box_t e = box_expand(cur, 1, 1);
    e = box_crop(e, box_xywh(0, 0, w, h));
    box_t s;
    if (!minmax_valid(disp, w, e, &s)) return 1;
    zpush(list, cur, s);
    return 1;
  }
  int spx = cw / 2, spy = chh / 2;                      /* :166-172 */
  box_t q1 = { cur.x0, cur.y0, cur.x0 + spx, cur.y0 + spy };
  box_t q4 = { cur.x0 + spx, cur.y0 + spy, cur.x1, cur.y1 };
  box_t q2 = { cur.x0 + spx, cur.y0, cur.x1, cur.y0 + spy };
  box_t q3 = { cur.x0, cur.y0 + spy, cur.x0 + spx, cur.y1 };
  box_t qs[4] = { q1, q2, q3, q4 };
  box_t ss[4] = { box_default(), box_default(), box_default(), box_default() };
  int32_t split_search = 0;                             /* :179-221, int32 arithmetic */
  for (int i = 0; i < 4; ++i) {
    box_t s;
    if (minmax_valid(disp, w, qs[i], &s)) {
      ss[i] = s;
      int32_t prod = wrap_mul((qs[i].x1 - qs[i].x0) + kx, (qs[i].y1 - qs[i].y0) + ky);
      split_search = wrap_add(split_search, wrap_mul(box_area(s), prod));
    }
  }
  box_t def = box_default();
  box_t cs = def;                                       /* :227-241 */
  if (!box_eq(ss[0], def)) cs = ss[0];
  for (int i = 1; i < 4; ++i) {
    if (!box_eq(ss[i], def) && box_eq(cs, def)) cs = ss[i];
    else cs = box_grow(cs, ss[i]);
  }
  int32_t current_search = wrap_mul(box_area(cs), wrap_mul(cw + kx, chh + ky));   /* :243 */
  const double IMPROVEMENT_RATIO = 0.8;
  if ((double)split_search > (double)current_search * IMPROVEMENT_RATIO && fail_count == 0) {  /* :247-316 */
    zone_t failed[4]; int nf = 0;
    for (int i = 0; i < 4; ++i)
      if (!subdivide(disp, w, h, qs[i], list, kx, ky, fail_count + 1)) { failed[nf].img = qs[i]; failed[nf].disp = ss[i]; nf++; }
#define MERGEABLE(A, B) (((A).img.x0 == (B).img.x0 || (A).img.y0 == (B).img.y0) && box_eq((A).disp, (B).disp))
    if (nf == 4) {
      zpush(list, cur, cs);
      return 1;
    } else if (nf == 3) {
      /* pairs tried in the order (0,1), (1,2), (0,2) :265-298 */
      if (MERGEABLE(failed[0], failed[1])) {
        zpush(list, box_grow(failed[0].img, failed[1].img), failed[0].disp);
        zpush(list, failed[2].img, failed[2].disp);
        return 1;
      }
      if (MERGEABLE(failed[1], failed[2])) {
        zpush(list, box_grow(failed[1].img, failed[2].img), failed[1].disp);
        zpush(list, failed[0].img, failed[0].disp);
        return 1;
      }
      if (MERGEABLE(failed[0], failed[2])) {
        zpush(list, box_grow(failed[0].img, failed[2].img), failed[0].disp);
        zpush(list, failed[1].img, failed[1].disp);
        return 1;
      }
      for (int i = 0; i < 3; ++i) zpush(list, failed[i].img, failed[i].disp);
    } else if (nf == 2) {
      if (MERGEABLE(failed[0], failed[1])) {
        zpush(list, box_grow(failed[0].img, failed[1].img), failed[0].disp);
        return 1;
      }
      zpush(list, failed[0].img, failed[0].disp);
      zpush(list, failed[1].img, failed[1].disp);
    } else if (nf == 1) {
      zpush(list, failed[0].img, failed[0].disp);
    }
#undef MERGEABLE
    return 1;
  } else if ((double)split_search > (double)current_search * IMPROVEMENT_RATIO && fail_count > 0) {
    return 0;                                            /* :317-319 */
  } else {                                               /* :320-326 */
    for (int i = 0; i < 4; ++i) subdivide(disp, w, h, qs[i], list, kx, ky, 0);
  }
  return 1;
}

int vwo_subdivide_regions(const vwo_disp_t* disp, int w, int h, int kx, int ky,
                          int32_t* zones_out, int max_zones) {
  zlist_t l = { 0, 0, 0 };
  subdivide(disp, w, h, box_xywh(0, 0, w, h), &l, kx, ky, 0);
  int n = l.n;
  if (n > max_zones) { free(l.z); return -1; }
  for (int i = 0; i < n; ++i) {
    int32_t* o = zones_out + 8 * i;
    o[0] = l.z[i].img.x0; o[1] = l.z[i].img.y0; o[2] = l.z[i].img.x1; o[3] = l.z[i].img.y1;
    o[4] = l.z[i].disp.x0; o[5] = l.z[i].disp.y0; o[6] = l.z[i].disp.x1; o[7] = l.z[i].disp.y1;
  }
  free(l.z);
  return n;
}

/* ------------------------------------------------------------------ prefilter */
/* Gaussian taps: Image/Filter.tcc:36-79, size Image/Filter.cc:31-37 */
static int gaussian_kernel(double sigma, float* k, int maxn) {
  if (sigma == 0) return 0;
  int size = (int)(7 * sigma);
  if (size < 3) size = 3; else if (size % 2 == 0) size -= 1;
  if (size > maxn) return -1;
  int center = size / 2;
  double sum = 0.0, tap;
  const double z = 1 / (sqrt(2.0) * sigma);
  for (int i = 1; i <= center; ++i) {
    tap = erf((i + 0.5) * z) - erf((i - 0.5) * z);
    sum += tap;
    k[center + i] = k[center - i] = (float)tap;
  }
  sum *= 2.0;
  tap = erf(0.5 * z) - erf(-0.5 * z);
  sum += tap;
  k[center] = (float)tap;
  double norm = 1.0 / sum;
  for (int i = 0; i < size; ++i) k[i] *= norm;          /* float *= double, rounds to float */
  return size;
}
int vwo_gaussian_kernel(double sigma, float* k, int maxn) { return gaussian_kernel(sigma, k, maxn); }

/* Stereo/PreFilter.h:45-95.  LoG = laplacian_filter(gaussian_filter(img, w)) : Image/Filter.h
 * gaussian_filter -> separable_convolution_filter(ConstantEdge); laplacian_filter -> 3x3
 * {{0,1,0},{1,-4,1},{0,1,0}} ConvolutionView (ConstantEdge) over the (lazy) gaussian view. */
int vwo_prefilter(const float* in, int w, int h, int mode, float width, float* out) {
  if (mode == VWO_PREFILTER_NONE) { memcpy(out, in, (size_t)w * h * sizeof(float)); return 0; }
  float k[512];
  int n = gaussian_kernel((double)width, k, 512);
  if (n < 0) return -1;
  float* g = (float*)malloc((size_t)w * h * sizeof(float));
  if (!g) return -2;
  int rc = vwo_separable_convolve(in, w, h, w, k, n, k, n, 0, g);
  if (rc) { free(g); return rc; }
  if (mode == VWO_PREFILTER_MEANSUB) {
    for (size_t i = 0; i < (size_t)w * h; ++i) out[i] = in[i] - g[i];
    free(g);
    return 0;
  }
  /* LoG: 3x3 Laplacian {{0,1,0},{1,-4,1},{0,1,0}} of the gaussian view, ConstantEdgeExtension, accumulated
   * row-major "result += k*s" in float (Image/Convolution.h:68-91, Image/Filter.h:318-326).  The zero taps
   * contribute +-0 and are skipped; the remaining order is top, left, centre, right, bottom. */
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int xm = clampi(x - 1, 0, w - 1), xp = clampi(x + 1, 0, w - 1), ym = clampi(y - 1, 0, h - 1), yp = clampi(y + 1, 0, h - 1);
      float r = 0.0f;
      r = r + 1.0f * g[(size_t)ym * w + x];
      r = r + 1.0f * g[(size_t)y * w + xm];
      r = r + -4.0f * g[(size_t)y * w + x];
      r = r + 1.0f * g[(size_t)y * w + xp];
      r = r + 1.0f * g[(size_t)yp * w + x];
      out[(size_t)y * w + x] = r;
    }
  free(g);
  return 0;
}

/* ------------------------------------------------------------------ the view */
typedef struct { float* d; int w, h; } imgf_t;
typedef struct { uint8_t* d; int w, h; } imgb_t;

static imgf_t crop_const_f(const float* src, int sw, int sh, int sp, box_t b) {     /* EdgeExtension.tcc:47-62 */
  imgf_t o; o.w = b.x1 - b.x0; o.h = b.y1 - b.y0;
  o.d = (float*)malloc((size_t)o.w * o.h * sizeof(float));
  for (int y = 0; y < o.h; ++y) {
    int syy = clampi(b.y0 + y, 0, sh - 1);
    for (int x = 0; x < o.w; ++x) o.d[(size_t)y * o.w + x] = src[(size_t)syy * sp + clampi(b.x0 + x, 0, sw - 1)];
  }
  return o;
}
static imgb_t crop_b(const uint8_t* src, int sw, int sh, int sp, box_t b, int zero) {
  imgb_t o; o.w = b.x1 - b.x0; o.h = b.y1 - b.y0;
  o.d = (uint8_t*)malloc(((size_t)o.w * o.h) > 0 ? (size_t)o.w * o.h : 1);
  for (int y = 0; y < o.h; ++y)
    for (int x = 0; x < o.w; ++x) {
      int sx = b.x0 + x, sy = b.y0 + y;
      uint8_t v;
      if (zero) v = (sx < 0 || sx >= sw || sy < 0 || sy >= sh) ? 0 : src[(size_t)sy * sp + sx];
      else v = src[(size_t)clampi(sy, 0, sh - 1) * sp + clampi(sx, 0, sw - 1)];
      o.d[(size_t)y * o.w + x] = v;
    }
  return o;
}

/* mean_pixel_value(subsample(copy_mask(img, create_mask(mask,0)),2)) : CorrelationView.cc:133-136,
 * Image/Statistics.h:363-368, Math/Functors.h:469-487 */
static int masked_mean(imgf_t img, imgb_t m, float* mean) {
  double acc = 0, cnt = 0;
  int ow = 1 + (img.w - 1) / 2, oh = 1 + (img.h - 1) / 2;
  for (int j = 0; j < oh; ++j)
    for (int i = 0; i < ow; ++i) {
      size_t k = (size_t)(2 * j) * img.w + 2 * i;
      if (m.d[k]) { acc += (double)img.d[k]; cnt += 1.0; }
    }
  if (cnt == 0) return 0;
  *mean = (float)(acc / cnt);
  return 1;
}

int vwo_num_levels(const vwo_corr_params* p, int bw, int bh) {
  /* ctor: CorrelationView.h:96-105 (float maths) */
  int sw = p->search_x1 - p->search_x0, sh = p->search_y1 - p->search_y0;
  int largest_search = sw > sh ? sw : sh;
  int max_by_search = (int)(floorf(logf((float)largest_search) / logf(2.0f)) - 1);
  if (max_by_search > p->max_pyramid_levels) max_by_search = p->max_pyramid_levels;
  if (max_by_search < 0) max_by_search = 0;
  /* prerasterize: CorrelationView.cc:301-310: log(int)->double, log(2.0f)->float */
  int smallest = bw < bh ? bw : bh;
  int largest_kernel = p->kernel_x > p->kernel_y ? p->kernel_x : p->kernel_y;
  int levels = (int)floor(log((double)smallest) / (double)logf(2.0f) - log((double)largest_kernel) / (double)logf(2.0f));
  if (max_by_search < levels) levels = max_by_search;
  if (levels < 1) levels = 0;
  return levels;
}

typedef struct {
  imgf_t* l; imgf_t* r; imgb_t* lm; imgb_t* rm; int levels;
} pyr_t;
static void pyr_free(pyr_t* p) {
  for (int i = 0; i <= p->levels; ++i) { free(p->l[i].d); free(p->r[i].d); free(p->lm[i].d); free(p->rm[i].d); }
  free(p->l); free(p->r); free(p->lm); free(p->rm);
}

/* Stereo/CorrelationView.cc:67-239 */
static int build_pyramids(const vwo_corr_params* p, const vwo_corr_inputs* in, box_t bbox, int levels, pyr_t* py) {
  int hkx = p->kernel_x / 2, hky = p->kernel_y / 2;
  int up = 1 << levels;
  py->levels = levels;
  py->l = (imgf_t*)calloc((size_t)levels + 1, sizeof(imgf_t));
  py->r = (imgf_t*)calloc((size_t)levels + 1, sizeof(imgf_t));
  py->lm = (imgb_t*)calloc((size_t)levels + 1, sizeof(imgb_t));
  py->rm = (imgb_t*)calloc((size_t)levels + 1, sizeof(imgb_t));
  box_t search = { p->search_x0, p->search_y0, p->search_x1, p->search_y1 };
  int ssx = search.x1 - search.x0, ssy = search.y1 - search.y0;
  box_t lg = box_expand(bbox, hkx * up, hky * up);                      /* :89-93 */
  box_t rg = box_shift(lg, search.x0, search.y0); rg.x1 += ssx; rg.y1 += ssy;   /* :96-97 */
  py->l[0] = crop_const_f(in->left, in->lcols, in->lrows, in->lpitch, lg);       /* :111-114 */
  py->r[0] = crop_const_f(in->right, in->rcols, in->rrows, in->rpitch, rg);
  imgb_t lm0 = crop_b(in->lmask, in->lcols, in->lrows, in->lmpitch, lg, 0);      /* :116-119 */
  imgb_t rm0 = crop_b(in->rmask, in->rcols, in->rrows, in->rmpitch, rg, 0);
  float lmean, rmean;
  int ok = masked_mean(py->l[0], lm0, &lmean) && masked_mean(py->r[0], rm0, &rmean);   /* :130-142 */
  if (ok) {                                                                     /* :144-149 */
    for (size_t i = 0; i < (size_t)py->l[0].w * py->l[0].h; ++i) if (!lm0.d[i]) py->l[0].d[i] = lmean;
    for (size_t i = 0; i < (size_t)py->r[0].w * py->r[0].h; ++i) if (!rm0.d[i]) py->r[0].d[i] = rmean;
  }
  free(lm0.d); free(rm0.d);
  if (!ok) { pyr_free(py); return 0; }
  box_t rmb = box_shift(bbox, search.x0, search.y0); rmb.x1 += ssx; rmb.y1 += ssy;   /* :192-197 */
  py->lm[0] = crop_b(in->lmask, in->lcols, in->lrows, in->lmpitch, bbox, 1);
  py->rm[0] = crop_b(in->rmask, in->rcols, in->rrows, in->rmpitch, rmb, 1);
  for (int i = 1; i <= levels; ++i) {                                            /* :209-216 */
    imgf_t a = py->l[i - 1], b = py->r[i - 1];
    py->l[i].w = 1 + (a.w - 1) / 2; py->l[i].h = 1 + (a.h - 1) / 2;
    py->l[i].d = (float*)malloc((size_t)py->l[i].w * py->l[i].h * sizeof(float));
    vwo_pyramid_down(a.d, a.w, a.h, py->l[i].d);
    py->r[i].w = 1 + (b.w - 1) / 2; py->r[i].h = 1 + (b.h - 1) / 2;
    py->r[i].d = (float*)malloc((size_t)py->r[i].w * py->r[i].h * sizeof(float));
    vwo_pyramid_down(b.d, b.w, b.h, py->r[i].d);
    imgb_t ma = py->lm[i - 1], mb = py->rm[i - 1];
    py->lm[i].w = 1 + (ma.w - 1) / 2; py->lm[i].h = 1 + (ma.h - 1) / 2;
    py->lm[i].d = (uint8_t*)malloc((size_t)py->lm[i].w * py->lm[i].h);
    vwo_subsample_mask_by_two(ma.d, ma.w, ma.h, py->lm[i].d);
    py->rm[i].w = 1 + (mb.w - 1) / 2; py->rm[i].h = 1 + (mb.h - 1) / 2;
    py->rm[i].d = (uint8_t*)malloc((size_t)py->rm[i].w * py->rm[i].h);
    vwo_subsample_mask_by_two(mb.d, mb.w, mb.h, py->rm[i].d);
  }
  if (p->prefilter_mode != VWO_PREFILTER_NONE) {                                 /* :233-236 */
    for (int i = 0; i <= levels; ++i) {
      float* t = (float*)malloc((size_t)py->l[i].w * py->l[i].h * sizeof(float));
      if (vwo_prefilter(py->l[i].d, py->l[i].w, py->l[i].h, p->prefilter_mode, p->prefilter_width, t)) { free(t); pyr_free(py); return -1; }
      free(py->l[i].d); py->l[i].d = t;
      t = (float*)malloc((size_t)py->r[i].w * py->r[i].h * sizeof(float));
      if (vwo_prefilter(py->r[i].d, py->r[i].w, py->r[i].h, p->prefilter_mode, p->prefilter_width, t)) { free(t); pyr_free(py); return -1; }
      free(py->r[i].d); py->r[i].d = t;
    }
  }
  return 1;
}

int vwo_build_pyramids(const vwo_corr_params* p, const vwo_corr_inputs* in,
                       int bx0, int by0, int bx1, int by1, int levels,
                       float** lpyr, float** rpyr, uint8_t** lmpyr, uint8_t** rmpyr, int32_t* dims) {
  pyr_t py; box_t bbox = { bx0, by0, bx1, by1 };
  int rc = build_pyramids(p, in, bbox, levels, &py);
  if (rc <= 0) return rc;
  for (int i = 0; i <= levels; ++i) {
    lpyr[i] = py.l[i].d; rpyr[i] = py.r[i].d; lmpyr[i] = py.lm[i].d; rmpyr[i] = py.rm[i].d;
    int32_t* d = dims + 8 * i;
    d[0] = py.l[i].w; d[1] = py.l[i].h; d[2] = py.r[i].w; d[3] = py.r[i].h;
    d[4] = py.lm[i].w; d[5] = py.lm[i].h; d[6] = py.rm[i].w; d[7] = py.rm[i].h;
  }
  free(py.l); free(py.r); free(py.lm); free(py.rm);
  return 1;
}
void vwo_free(void* p) { free(p); }


/* ------------------------------------------------------------------ ParabolaSubpixelView */
/* prefilter.filter(image) evaluated over an arbitrary region of the lazily edge-extended view, as
 * crop(prefilter.filter(img), region) does in ParabolaSubpixelView::prerasterize (ParabolaSubpixelView.cc:296-327):
 *   NONE    : edge_extend(img, Constant)                                     (PreFilter.h:45-51)
 *   MEANSUB : edge_extend(img) - gaussian_filter(img): the separable convolution of the replicate-extended
 *             image, also at out-of-image coordinates                        (PreFilter.h:60-70, Convolution.h:275-297)
 *   LOG     : laplacian over edge_extend(gaussian view, Constant): the stencil reads the in-image gaussian at
 *             clamped coordinates                                            (PreFilter.h:53-58, Convolution.h:154-166) */
static imgf_t filtered_region(const float* img, int w, int h, int pitch, int mode, float width, box_t reg) {
  imgf_t o = crop_const_f(img, w, h, pitch, reg);
  if (mode == VWO_PREFILTER_NONE) return o;
  float k[512];
  int n = gaussian_kernel((double)width, k, 512);
  if (mode == VWO_PREFILTER_MEANSUB) {
    if (n <= 0) { for (size_t i = 0; i < (size_t)o.w * o.h; ++i) o.d[i] = o.d[i] - o.d[i]; return o; }
    box_t big = { reg.x0 - n, reg.y0 - n, reg.x1 + n, reg.y1 + n };
    imgf_t e = crop_const_f(img, w, h, pitch, big);
    float* g = (float*)malloc((size_t)e.w * e.h * sizeof(float));
    vwo_separable_convolve(e.d, e.w, e.h, e.w, k, n, k, n, 0, g);
    for (int y = 0; y < o.h; ++y)
      for (int x = 0; x < o.w; ++x) o.d[(size_t)y * o.w + x] = o.d[(size_t)y * o.w + x] - g[(size_t)(y + n) * e.w + (x + n)];
    free(g); free(e.d);
    return o;
  }
  /* LOG */
  box_t need = { clampi(reg.x0 - 1, 0, w - 1), clampi(reg.y0 - 1, 0, h - 1), clampi(reg.x1, 0, w - 1) + 1, clampi(reg.y1, 0, h - 1) + 1 };
  int m = n > 0 ? n : 0;
  box_t big = { need.x0 - m, need.y0 - m, need.x1 + m, need.y1 + m };
  imgf_t e = crop_const_f(img, w, h, pitch, big);
  float* g = (float*)malloc((size_t)e.w * e.h * sizeof(float));
  if (n > 0) vwo_separable_convolve(e.d, e.w, e.h, e.w, k, n, k, n, 0, g);
  else memcpy(g, e.d, (size_t)e.w * e.h * sizeof(float));
#define GAT(X, Y) g[(size_t)(clampi((Y), 0, h - 1) - big.y0) * e.w + (clampi((X), 0, w - 1) - big.x0)]
  for (int y = 0; y < o.h; ++y)
    for (int x = 0; x < o.w; ++x) {
      int gx = reg.x0 + x, gy = reg.y0 + y;
      float r = 0.0f;
      r = r + 1.0f * GAT(gx, gy - 1);
      r = r + 1.0f * GAT(gx - 1, gy);
      r = r + -4.0f * GAT(gx, gy);
      r = r + 1.0f * GAT(gx + 1, gy);
      r = r + 1.0f * GAT(gx, gy + 1);
      o.d[(size_t)y * o.w + x] = r;
    }
#undef GAT
  free(g); free(e.d);
  return o;
}

/* ParabolaSubpixelView::prerasterize + evaluate (Stereo/ParabolaSubpixelView.cc:31-330).
 * disp: cols x rows x {dx,dy,valid} floats (the integer disparity as PixelMask<Vector2f>); out: bbox-sized. */
int vwo_parabola_subpixel(const float* disp, int cols, int rows, const float* left, int lpitch,
                          const float* right, int rcols, int rrows, int rpitch,
                          int kx, int ky, int prefilter_mode, float prefilter_width,
                          int bx0, int by0, int bx1, int by1, float* out) {
  const int bw = bx1 - bx0, bh = by1 - by0;
  if (bw <= 0 || bh <= 0 || bx0 < 0 || by0 < 0 || bx1 > cols || by1 > rows) return -1;
  if (kx % 2 != 1 || ky % 2 != 1) return -2;
  /* integer_disparity = crop(m_disparity, bbox) as PixelMask<Vector2i>: float -> int truncation (:292) */
  vwo_disp_t* id = (vwo_disp_t*)malloc((size_t)bw * bh * sizeof(vwo_disp_t));
  int mnx = 0, mny = 0, mxx = 0, mxy = 0, any = 0;
  for (int y = 0; y < bh; ++y)
    for (int x = 0; x < bw; ++x) {
      const float* p = disp + ((size_t)(by0 + y) * cols + (bx0 + x)) * 3;
      vwo_disp_t v = { (int)p[0], (int)p[1], p[2] != 0.0f };
      id[(size_t)y * bw + x] = v;
      if (v.valid) {      /* get_disparity_range (DisparityMap.h:52-66) */
        if (!any) { mnx = mxx = v.dx; mny = mxy = v.dy; any = 1; }
        if (v.dx < mnx) mnx = v.dx;
        if (v.dx > mxx) mxx = v.dx;
        if (v.dy < mny) mny = v.dy;
        if (v.dy > mxy) mxy = v.dy;
      }
    }
  /* entire_search_range: max += 1; expand(1)  (:296-299) */
  box_t sr = { mnx, mny, mxx + 1, mxy + 1 };
  sr = box_expand(sr, 1, 1);
  const int hkx = kx / 2, hky = ky / 2;
  box_t lreg = { bx0 - hkx, by0 - hky, bx1 + hkx, by1 + hky };                                   /* :302-305 */
  box_t rreg = { lreg.x0 + sr.x0, lreg.y0 + sr.y0, lreg.x1 + sr.x0 + (sr.x1 - sr.x0), lreg.y1 + sr.y0 + (sr.y1 - sr.y0) };
  imgf_t L = filtered_region(left, cols, rows, lpitch, prefilter_mode, prefilter_width, lreg);
  imgf_t R = filtered_region(right, rcols, rrows, rpitch, prefilter_mode, prefilter_width, rreg);
  float* patch = (float*)calloc((size_t)bw * bh * 9, sizeof(float));
  /* zones (:74-104) */
  zlist_t big = { 0, 0, 0 }, zones = { 0, 0, 0 };
  subdivide(id, bw, bh, box_xywh(0, 0, bw, bh), &big, kx, ky, 0);
  for (int zi = 0; zi < big.n; ++zi) {
    zone_t z = big.z[zi];
    double len1 = (double)box_area(z.img), len2 = (double)box_area(z.disp);
    if (len2 / len1 < 1.0) { zpush(&zones, z.img, z.disp); continue; }
    for (int x = z.img.x0; x < z.img.x1; ++x)
      for (int y = z.img.y0; y < z.img.y1; ++y) {
        vwo_disp_t v = id[(size_t)y * bw + x];
        if (!v.valid) continue;
        box_t d = { v.dx, v.dy, v.dx + 1, v.dy + 1 };
        zpush(&zones, box_xywh(x, y, 1, 1), d);
      }
  }
  for (int zi = 0; zi < zones.n; ++zi) {                                                         /* :107-224 */
    zone_t z = zones.z[zi];
    z.disp = box_expand(z.disp, 1, 1);
    const int zw = z.img.x1 - z.img.x0, zh = z.img.y1 - z.img.y0;
    const int cw = zw + kx - 1, ch = zh + ky - 1;
    double* applied = (double*)malloc((size_t)cw * ch * sizeof(double));
    double* metric = (double*)malloc((size_t)zw * zh * sizeof(double));
    for (int ddx = 0; ddx < z.disp.x1 - z.disp.x0; ++ddx)
      for (int ddy = 0; ddy < z.disp.y1 - z.disp.y0; ++ddy) {
        const int ax = ddx + z.disp.x0, ay = ddy + z.disp.y0;
        for (int y = 0; y < ch; ++y)
          for (int x = 0; x < cw; ++x) {
            float a = L.d[(size_t)(z.img.y0 + y) * L.w + (z.img.x0 + x)];
            float b = R.d[(size_t)(z.img.y0 + y + ay - sr.y0) * R.w + (z.img.x0 + x + ax - sr.x0)];
            applied[(size_t)y * cw + x] = (double)fabsf(a - b);
          }
        vwo_fast_box_sum(applied, cw, ch, cw, kx, ky, metric);
        for (int y = 0; y < zh; ++y)
          for (int x = 0; x < zw; ++x) {
            const vwo_disp_t v = id[(size_t)(z.img.y0 + y) * bw + (z.img.x0 + x)];
            const int ex = ax - v.dx, ey = ay - v.dy;
            if (ex >= -1 && ex <= 1 && ey >= -1 && ey <= 1)
              patch[((size_t)(z.img.y0 + y) * bw + (z.img.x0 + x)) * 9 + (ey + 1) * 3 + (ex + 1)] = (float)metric[(size_t)y * zw + x];
          }
      }
    free(applied); free(metric);
  }
  /* pinvA (ParabolaSubpixelView.h:82-88), float */
  static const double pd[54] = {
     1.0/6, -1.0/3,  1.0/6,  1.0/6, -1.0/3,  1.0/6,   1.0/6, -1.0/3,  1.0/6,
     1.0/6,  1.0/6,  1.0/6, -1.0/3, -1.0/3, -1.0/3,   1.0/6,  1.0/6,  1.0/6,
     1.0/4,    0.0, -1.0/4,    0.0,    0.0,    0.0,  -1.0/4,    0.0,  1.0/4,
    -1.0/6,    0.0,  1.0/6, -1.0/6,    0.0,  1.0/6,  -1.0/6,    0.0,  1.0/6,
    -1.0/6, -1.0/6, -1.0/6,    0.0,    0.0,    0.0,   1.0/6,  1.0/6,  1.0/6,
    -1.0/9,  2.0/9, -1.0/9,  2.0/9,   5.0/9, 2.0/9,  -1.0/9,  2.0/9, -1.0/9 };
  float pinv[54];
  for (int i = 0; i < 54; ++i) pinv[i] = (float)pd[i];
  for (int y = 0; y < bh; ++y)                                                                      /* :230-275 */
    for (int x = 0; x < bw; ++x) {
      const vwo_disp_t v = id[(size_t)y * bw + x];
      float* o = out + ((size_t)y * bw + x) * 3;
      const float* c = patch + ((size_t)y * bw + x) * 9;
      if (!v.valid) { o[0] = o[1] = o[2] = 0.0f; continue; }
      o[0] = (float)v.dx; o[1] = (float)v.dy; o[2] = 1.0f;
      int alleq = 1;
      for (int k = 1; k < 9; ++k) if (c[k] != c[k - 1]) alleq = 0;     /* std::equal(begin+1,end,begin) */
      if (alleq) continue;
      float xs[6];
      for (int r = 0; r < 6; ++r) { float acc = 0.0f; for (int k = 0; k < 9; ++k) { float pr = pinv[r * 9 + k] * c[k]; acc = acc + pr; } xs[r] = acc; }
      float t1 = 4 * xs[0]; t1 = t1 * xs[1];
      float t2 = xs[2] * xs[2];
      float denom = t1 - t2;
      float n1a = xs[2] * xs[4], n1b = 2 * xs[1]; n1b = n1b * xs[3];
      float n2a = xs[2] * xs[3], n2b = 2 * xs[0]; n2b = n2b * xs[4];
      float ox = (n1a - n1b) / denom, oy = (n2a - n2b) / denom;
      double nn = 0.0; { float q = ox * ox; nn += q; q = oy * oy; nn += q; }
      nn = (double)(float)nn;
      if (sqrt(nn) < 5.0) { o[0] = (float)v.dx + ox; o[1] = (float)v.dy + oy; }
    }
  free(big.z); free(zones.z); free(patch); free(L.d); free(R.d); free(id);
  return 0;
}

static int zone_cmp(const void* a, const void* b) {                                /* Correlation.h:87-91 */
  const zone_t *A = (const zone_t*)a, *B = (const zone_t*)b;
  double va = (double)box_w(A->img) * (double)box_h(A->img) * (double)box_w(A->disp) * (double)box_h(A->disp);
  double vb = (double)box_w(B->img) * (double)box_h(B->img) * (double)box_w(B->disp) * (double)box_h(B->disp);
  return va < vb ? -1 : (va > vb ? 1 : 0);
}

/* Stereo/CorrelationView.cc:273-886, block-matching branch only. out: bw*bh {dx,dy,valid} floats. */
typedef struct { float* d; int cols, rows, ulx, uly; } diff_t;
static int prerasterize_sgm(const vwo_corr_params* p, const vwo_corr_inputs* in, box_t bbox, float* out, int* levels_out, const diff_t* df);
static int prerasterize(const vwo_corr_params* p, const vwo_corr_inputs* in, box_t bbox, float* out, int* levels_out, const diff_t* df) {
  if (df && df->d) {                                                              /* :276-283 */
    if (!(bbox.x0 >= df->ulx && bbox.y0 >= df->uly && bbox.x1 <= df->ulx + df->cols && bbox.y1 <= df->uly + df->rows)) return -1;
  }
  if (p->algorithm != 0) return prerasterize_sgm(p, in, bbox, out, levels_out, df);
  const int bw = bbox.x1 - bbox.x0, bh = bbox.y1 - bbox.y0;
  const int kx = p->kernel_x, ky = p->kernel_y, hkx = kx / 2, hky = ky / 2;
  const int ssx = p->search_x1 - p->search_x0, ssy = p->search_y1 - p->search_y0;
  int levels = vwo_num_levels(p, bw, bh);
  if (levels_out) *levels_out = levels;
  const int up = 1 << levels;
  pyr_t py;
  int rc = build_pyramids(p, in, bbox, levels, &py);                              /* :320 */
  if (rc < 0) return -20;
  if (rc == 0) {                                                                  /* :321-331: all-invalid tile */
    memset(out, 0, (size_t)bw * bh * 3 * sizeof(float));
    return 0;
  }
  zlist_t zones = { 0, 0, 0 };
  zpush(&zones, box_xywh(0, 0, py.lm[levels].w, py.lm[levels].h),                 /* :338-342 */
        box_xywh(0, 0, ssx / up + 1, ssy / up + 1));
  vwo_disp_t* disparity = NULL; int dw = 0, dh = 0;
  for (int level = levels; level >= 0; --level) {                                 /* :363 */
    int scaling = 1 << level;
    dw = py.lm[level].w; dh = py.lm[level].h;                                     /* :377 */
    free(disparity);
    disparity = (vwo_disp_t*)calloc(((size_t)dw * dh) > 0 ? (size_t)dw * dh : 1, sizeof(vwo_disp_t));
    int rox = up * hkx / scaling, roy = up * hky / scaling;                       /* :381 */
    qsort(zones.z, (size_t)zones.n, sizeof(zone_t), zone_cmp);                    /* :606 (order only matters for timeouts) */
    for (int zi = 0; zi < zones.n; ++zi) {
      zone_t z = zones.z[zi];
      box_t lr = box_expand(box_shift(z.img, rox, roy), hkx, hky);                /* :611-612 */
      box_t rr = box_shift(lr, z.disp.x0, z.disp.y0);                             /* :615-616 */
      int dsx = z.disp.x1 - z.disp.x0, dsy = z.disp.y1 - z.disp.y0;
      rr.x1 += dsx; rr.y1 += dsy;
      int zw = z.img.x1 - z.img.x0, zh = z.img.y1 - z.img.y0;
      if (zw <= 0 || zh <= 0) continue;
      /* crop(left_pyramid[level], left_region): plain crops, regions lie inside by construction;
       * clamp anyway (out-of-range reads would be UB in the reference) */
      imgf_t lc = crop_const_f(py.l[level].d, py.l[level].w, py.l[level].h, py.l[level].w, lr);
      imgf_t rcq = crop_const_f(py.r[level].d, py.r[level].w, py.r[level].h, py.r[level].w, rr);
      vwo_disp_t* zd = (vwo_disp_t*)malloc((size_t)zw * zh * sizeof(vwo_disp_t));
      /* calc_disparity(cost, crop(L), crop(R), region, zone.disp.size(), kernel) :641-648.
       * right crop is lr.size + disp.size; calc_disparity uses + size - 1 of it. */
      rc = vwo_calc_disparity(p->cost_type, lc.d, lc.w, lc.h, lc.w, rcq.d, rcq.w, rcq.h, rcq.w, dsx, dsy, kx, ky, zd);
      if (rc) { free(lc.d); free(rcq.d); free(zd); free(disparity); free(zones.z); pyr_free(&py); return -30 + rc; }
      if (p->consistency_threshold >= 0 && level == 0) {                          /* :653-694 */
        /* R->L: calc_disparity(crop(edge_extend(right), right_region),
         *                      crop(edge_extend(left), left_region - disp.size()), ...) - size */
        box_t ll = box_shift(lr, -dsx, -dsy);
        /* calc_disparity crops its 2nd argument to region + search_volume - 1 */
        box_t ll_full = ll; ll_full.x1 = ll.x0 + (rr.x1 - rr.x0) + dsx - 1; ll_full.y1 = ll.y0 + (rr.y1 - rr.y0) + dsy - 1;
        imgf_t l2 = crop_const_f(py.l[level].d, py.l[level].w, py.l[level].h, py.l[level].w, ll_full);
        int rlw = (rr.x1 - rr.x0) - kx + 1, rlh = (rr.y1 - rr.y0) - ky + 1;
        vwo_disp_t* rl = (vwo_disp_t*)malloc((size_t)rlw * rlh * sizeof(vwo_disp_t));
        rc = vwo_calc_disparity(p->cost_type, rcq.d, rcq.w, rcq.h, rcq.w, l2.d, l2.w, l2.h, l2.w, dsx, dsy, kx, ky, rl);
        if (rc) { free(l2.d); free(rl); free(lc.d); free(rcq.d); free(zd); free(disparity); free(zones.z); pyr_free(&py); return -40 + rc; }
        for (size_t i = 0; i < (size_t)rlw * rlh; ++i) { rl[i].dx -= dsx; rl[i].dy -= dsy; }
        if (df && df->d)                                                          /* :669-676: offset = zone min + bbox.min - region_ul */
          consistency_check_diff(zd, zw, zh, zw, rl, rlw, rlh, p->consistency_threshold, df->d, df->cols,
                                 z.img.x0 + bbox.x0 - df->ulx, z.img.y0 + bbox.y0 - df->uly);
        else
          vwo_cross_corr_consistency_check(zd, zw, zh, zw, rl, rlw, rlh, p->consistency_threshold);
        free(l2.d); free(rl);
      }
      for (int y = 0; y < zh; ++y)                                                /* :698 */
        for (int x = 0; x < zw; ++x) {
          vwo_disp_t v = zd[(size_t)y * zw + x];
          v.dx += z.disp.x0; v.dy += z.disp.y0;
          disparity[(size_t)(z.img.y0 + y) * dw + (z.img.x0 + x)] = v;
        }
      free(lc.d); free(rcq.d); free(zd);
    }
    if (p->filter_half_kernel > 0) {                                              /* :713-744 */
      vwo_disp_t* t = (vwo_disp_t*)malloc((size_t)dw * dh * sizeof(vwo_disp_t));
      int fh = p->filter_half_kernel;
      if (level != 0) vwo_disparity_cleanup_using_thresh(disparity, dw, dh, fh, fh, 3.0, 0.5, t);
      else            vwo_rm_outliers_using_thresh(disparity, dw, dh, fh, fh, 3.0, 0.5, t);
      vwo_disparity_mask(t, dw, dh, py.lm[level].d, py.rm[level].d, py.rm[level].w, py.rm[level].h, disparity);
      free(t);
    }
    vwo_disparity_blob_filter(disparity, dw, dh, p->blob_filter_area / scaling);    /* :746-749, :242-271 */
    if (level != 0) {                                                             /* :754-799 */
      zones.n = 0;
      subdivide(disparity, dw, dh, box_xywh(0, 0, dw, dh), &zones, kx, ky, 0);
      int nl = level - 1;
      box_t scale_search = box_xywh(0, 0, py.r[nl].w - py.l[nl].w, py.r[nl].h - py.l[nl].h);
      box_t next_zone = box_xywh(0, 0, py.lm[nl].w, py.lm[nl].h);
      box_t default_range = box_xywh(0, 0, ssx, ssy);
      for (int zi = 0; zi < zones.n; ++zi) {
        zone_t* z = &zones.z[zi];
        z->img = box_scale(z->img, 2);
        z->img = box_crop(z->img, next_zone);
        z->disp = box_scale(z->disp, 2);
        z->disp = box_expand(z->disp, 2, 2);
        z->disp = box_crop(z->disp, scale_search);
        if (box_empty(z->disp)) z->disp = default_range;
      }
    }
  }
  if (dw != bw || dh != bh) { free(disparity); free(zones.z); pyr_free(&py); return -50; }   /* :832-834 */
  if (df && df->d)                                                                /* :848-857 */
    for (int r = 0; r < bh; ++r)
      for (int c = 0; c < bw; ++c)
        if (!disparity[(size_t)r * bw + c].valid) df->d[((size_t)(r + bbox.y0 - df->uly) * df->cols + (c + bbox.x0 - df->ulx)) * 2 + 1] = 0.0f;
  for (size_t i = 0; i < (size_t)bw * bh; ++i) {                                  /* :880-884 */
    out[3 * i + 0] = (float)(disparity[i].dx + p->search_x0);
    out[3 * i + 1] = (float)(disparity[i].dy + p->search_y0);
    out[3 * i + 2] = disparity[i].valid ? 1.0f : 0.0f;
  }
  free(disparity); free(zones.z); pyr_free(&py);
  return 0;
}

/* Stereo/CorrelationView.cc:273-886, the SGM / MGM / FINAL_MGM branch (:392-595, 700-750, 859-871): one
 * calc_disparity_sgm per level over the whole (padded) tile, seeded by the previous level's filtered disparity; R->L pass +
 * consistency check at the levels >= min_consistency_level; the sub-pixel view is made at level 0 BEFORE the check and the
 * filters, whose invalidations are copied onto it at the end. */
static int prerasterize_sgm(const vwo_corr_params* p, const vwo_corr_inputs* in, box_t bbox, float* out, int* levels_out, const diff_t* df) {
  const int bw = bbox.x1 - bbox.x0, bh = bbox.y1 - bbox.y0;
  const int kx = p->kernel_x, ky = p->kernel_y, hkx = kx / 2, hky = ky / 2;
  const int ssx = p->search_x1 - p->search_x0, ssy = p->search_y1 - p->search_y0;
  vwo_corr_params q = *p;
  q.prefilter_mode = VWO_PREFILTER_NONE;                                          /* CorrelationView.h:96-97 */
  int levels = vwo_num_levels(&q, bw, bh);
  if (levels_out) *levels_out = levels;
  const int up = 1 << levels;
  pyr_t py;
  int rc = build_pyramids(&q, in, bbox, levels, &py);
  if (rc < 0) return -20;
  if (rc == 0) { memset(out, 0, (size_t)bw * bh * 3 * sizeof(float)); return 0; }
  vwo_disp_t *disparity = NULL, *disparity_rl = NULL, *prev = NULL, *prev_rl = NULL;
  int dw = 0, dh = 0, pw = 0, ph = 0, rlw = 0, rlh = 0, prlw = 0, prlh = 0;
  float* sub = NULL;
  const int threads = p->sgm_threads > 0 ? p->sgm_threads : 4;
  const double mem = p->memory_limit_mb > 0 ? p->memory_limit_mb : 6000.0;
  int err = 0;
  for (int level = levels; level >= 0 && !err; --level) {
    const int use_mgm = p->algorithm == 2 || (p->algorithm == 3 && level == 0);    /* :366-367 */
    const int scaling = 1 << level;
    free(prev); prev = disparity; pw = dw; ph = dh;                                /* :373-376 */
    free(prev_rl); prev_rl = disparity_rl; prlw = rlw; prlh = rlh;
    disparity = NULL; disparity_rl = NULL;
    dw = py.lm[level].w; dh = py.lm[level].h;
    const int rox = up * hkx / scaling, roy = up * hky / scaling;
    const int dsx = ssx / scaling, dsy = ssy / scaling;                            /* :395-396: BBox2i(0,0,w/scaling,h/scaling).size() */
    box_t zone = box_xywh(0, 0, dw, dh);
    box_t lr = box_expand(box_shift(zone, rox, roy), hkx, hky);                    /* :404-405 */
    box_t rr = lr; rr.x1 += dsx; rr.y1 += dsy;                                     /* :406-407 */
    imgf_t lc = crop_const_f(py.l[level].d, py.l[level].w, py.l[level].h, py.l[level].w, lr);
    imgf_t rcq = crop_const_f(py.r[level].d, py.r[level].w, py.r[level].h, py.r[level].w, rr);
    disparity = (vwo_disp_t*)calloc((size_t)lc.w * lc.h + 1, sizeof(vwo_disp_t));
    int ow = 0, oh = 0;
    if (level == 0) sub = (float*)calloc((size_t)lc.w * lc.h * 3 + 1, sizeof(float));
    rc = vwo_calc_disparity_sgm(lc.d, lc.w, lc.h, lc.w, rcq.d, rcq.w, rcq.h, rcq.w, dsx, dsy, kx, p->cost_type, 5, 0, 0, use_mgm,
                                p->sgm_subpixel_mode, p->sgm_search_buffer_x, p->sgm_search_buffer_y, mem, threads,
                                py.lm[level].d, py.rm[level].d, py.rm[level].w, py.rm[level].h,
                                level < levels ? (const int*)prev : NULL, pw, ph, NULL, 0, (int*)disparity, level == 0 ? sub : NULL, &ow, &oh);
    if (rc || ow != dw || oh != dh) { err = rc ? rc : -60; free(lc.d); free(rcq.d); break; }
    int check_rl = 0;
    imgb_t rrm = { NULL, 0, 0 }, lrm = { NULL, 0, 0 };
    if (p->consistency_threshold >= 0.0f && level >= p->min_consistency_level) {   /* :438-590 */
      check_rl = 1;
      box_t llr = box_shift(lr, -dsx, -dsy); llr.x1 += 2 * dsx; llr.y1 += 2 * dsy; /* :453-455 */
      imgf_t l2 = crop_const_f(py.l[level].d, py.l[level].w, py.l[level].h, py.l[level].w, llr);
      /* masks of the reversed problem (:496-508), zero edge extension */
      box_t rmb = box_xywh(0, 0, (rr.x1 - rr.x0) - 2 * hkx, (rr.y1 - rr.y0) - 2 * hky);
      box_t lmb = box_shift(box_xywh(0, 0, (llr.x1 - llr.x0) - 2 * hkx, (llr.y1 - llr.y0) - 2 * hky), -dsx, -dsy);
      rrm = crop_b(py.rm[level].d, py.rm[level].w, py.rm[level].h, py.rm[level].w, rmb, 1);
      lrm = crop_b(py.lm[level].d, py.lm[level].w, py.lm[level].h, py.lm[level].w, lmb, 1);
      disparity_rl = (vwo_disp_t*)calloc((size_t)rcq.w * rcq.h + 1, sizeof(vwo_disp_t));
      rc = vwo_calc_disparity_sgm(rcq.d, rcq.w, rcq.h, rcq.w, l2.d, l2.w, l2.h, l2.w, dsx, dsy, kx, p->cost_type, 5, 0, 0, use_mgm,
                                  p->sgm_subpixel_mode, p->sgm_search_buffer_x, p->sgm_search_buffer_y, mem, threads,
                                  rrm.d, lrm.d, lrm.w, lrm.h, level < levels ? (const int*)prev_rl : NULL, prlw, prlh, NULL, 0,
                                  (int*)disparity_rl, NULL, &rlw, &rlh);
      free(l2.d);
      if (rc) { err = rc; free(lc.d); free(rcq.d); free(rrm.d); free(lrm.d); break; }
      for (size_t i = 0; i < (size_t)rlw * rlh; ++i) { disparity_rl[i].dx -= dsx; disparity_rl[i].dy -= dsy; }      /* :548-549 */
      if (level == 0 && df && df->d)
        consistency_check_diff(disparity, dw, dh, dw, disparity_rl, rlw, rlh, p->consistency_threshold, df->d, df->cols,
                               bbox.x0 - df->ulx, bbox.y0 - df->uly);
      else
        vwo_cross_corr_consistency_check(disparity, dw, dh, dw, disparity_rl, rlw, rlh, p->consistency_threshold);
      for (size_t i = 0; i < (size_t)rlw * rlh; ++i) { disparity_rl[i].dx += dsx; disparity_rl[i].dy += dsy; }      /* :586 */
    }
    free(lc.d); free(rcq.d);
    if (p->filter_half_kernel > 0) {                                              /* :713-744 */
      const int fh = p->filter_half_kernel;
      vwo_disp_t* t = (vwo_disp_t*)malloc((size_t)dw * dh * sizeof(vwo_disp_t));
      if (level != 0) vwo_disparity_cleanup_using_thresh(disparity, dw, dh, fh, fh, 3.0, 0.5, t);
      else            vwo_rm_outliers_using_thresh(disparity, dw, dh, fh, fh, 3.0, 0.5, t);
      vwo_disparity_mask(t, dw, dh, py.lm[level].d, py.rm[level].d, py.rm[level].w, py.rm[level].h, disparity);
      free(t);
      if (level != 0 && check_rl) {
        vwo_disp_t* t2 = (vwo_disp_t*)malloc((size_t)rlw * rlh * sizeof(vwo_disp_t));
        vwo_disparity_cleanup_using_thresh(disparity_rl, rlw, rlh, fh, fh, 3.0, 0.5, t2);
        vwo_disparity_mask(t2, rlw, rlh, rrm.d, lrm.d, lrm.w, lrm.h, disparity_rl);
        free(t2);
      }
    }
    vwo_disparity_blob_filter(disparity, dw, dh, p->blob_filter_area / scaling);     /* :747-749 */
    if (check_rl && level != 0) vwo_disparity_blob_filter(disparity_rl, rlw, rlh, p->blob_filter_area / scaling);
    free(rrm.d); free(lrm.d);
  }
  if (!err && (dw != bw || dh != bh)) err = -50;
  if (!err) {
    if (df && df->d)                                                              /* :848-857 */
      for (int r = 0; r < bh; ++r)
        for (int c = 0; c < bw; ++c)
          if (!disparity[(size_t)r * bw + c].valid) df->d[((size_t)(r + bbox.y0 - df->uly) * df->cols + (c + bbox.x0 - df->ulx)) * 2 + 1] = 0.0f;
    for (size_t i = 0; i < (size_t)bw * bh; ++i) {                                /* :859-871: PixelMask sum keeps the child values */
      out[3 * i + 0] = sub[3 * i + 0] + (float)p->search_x0;
      out[3 * i + 1] = sub[3 * i + 1] + (float)p->search_y0;
      out[3 * i + 2] = (sub[3 * i + 2] != 0.0f && disparity[i].valid) ? 1.0f : 0.0f;
    }
  }
  free(disparity); free(disparity_rl); free(prev); free(prev_rl); free(sub); pyr_free(&py);
  return err;
}

/* PyramidCorrelationView::rasterize, Stereo/CorrelationView.h:123-133 */
int vwo_pyramid_correlate_rasterize(const vwo_corr_params* p, const vwo_corr_inputs* in,
                                    int bx0, int by0, int bx1, int by1,
                                    float* dest, int dest_pitch, int* levels_out) {
  return vwo_pyramid_correlate_rasterize_ex(p, in, bx0, by0, bx1, by1, dest, dest_pitch, levels_out, NULL, 0, 0, 0, 0);
}
int vwo_pyramid_correlate_rasterize_ex(const vwo_corr_params* p, const vwo_corr_inputs* in,
                                       int bx0, int by0, int bx1, int by1, float* dest, int dest_pitch, int* levels_out,
                                       float* diff, int diff_cols, int diff_rows, int region_ul_x, int region_ul_y) {
  const diff_t dfv = { diff, diff_cols, diff_rows, region_ul_x, region_ul_y };
  const diff_t* df = diff ? &dfv : NULL;
  if (bx1 <= bx0 || by1 <= by0) return -1;
  box_t bbox = { bx0, by0, bx1, by1 };
  box_t proc = bbox;
  if (p->collar_size > 0) proc = box_expand(proc, p->collar_size, p->collar_size);
  int pw = proc.x1 - proc.x0, ph = proc.y1 - proc.y0;
  float* buf = (float*)malloc((size_t)pw * ph * 3 * sizeof(float));
  if (!buf) return -2;
  int rc = prerasterize(p, in, proc, buf, levels_out, df);
  if (rc == 0) {
    int ox = bbox.x0 - proc.x0, oy = bbox.y0 - proc.y0;
    for (int y = 0; y < by1 - by0; ++y)
      memcpy(dest + (size_t)y * dest_pitch * 3, buf + ((size_t)(y + oy) * pw + ox) * 3, (size_t)(bx1 - bx0) * 3 * sizeof(float));
  }
  free(buf);
  return rc;
}

int vwo_max_threads(void);
/* Tile-parallel calc_disparity for the CPU baseline: the reference parallelises over independent
 * output tiles on a thread pool (Image/ImageIO.h:289-311); within a tile best_of_search_convolution is
 * single threaded.  Rasters are shared, each tile works on its own sub-rectangles via pitches. */
int vwo_calc_disparity_tiled(int cost, const float* left, int lw, int lh, int lpitch,
                             const float* right, int rw, int rh, int rpitch,
                             int sx, int sy, int kx, int ky, int tile, int nthreads, vwo_disp_t* out) {
  const int W = lw - kx + 1, H = lh - ky + 1;
  if (W <= 0 || H <= 0 || rw < lw + sx - 1 || rh < lh + sy - 1) return -1;
  const int tx = (W + tile - 1) / tile, ty = (H + tile - 1) / tile;
  int err = 0;
  if (nthreads <= 0) nthreads = vwo_max_threads();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
  for (int t = 0; t < tx * ty; ++t) {
    const int x0 = (t % tx) * tile, y0 = (t / tx) * tile;
    const int tw = x0 + tile < W ? tile : W - x0, th = y0 + tile < H ? tile : H - y0;
    vwo_disp_t* tmp = (vwo_disp_t*)malloc((size_t)tw * th * sizeof(vwo_disp_t));
    int rc = vwo_calc_disparity(cost, left + (size_t)y0 * lpitch + x0, tw + kx - 1, th + ky - 1, lpitch,
                                right + (size_t)y0 * rpitch + x0, tw + kx - 1 + sx - 1, th + ky - 1 + sy - 1, rpitch,
                                sx, sy, kx, ky, tmp);
    if (rc) err = rc;
    else for (int y = 0; y < th; ++y) memcpy(out + (size_t)(y0 + y) * W + x0, tmp + (size_t)y * tw, (size_t)tw * sizeof(vwo_disp_t));
    free(tmp);
  }
  return err;
}

int vwo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Tile scheduler stand-in for block_write_image (Image/ImageIO.h:289-311): independent tiles
 * on a thread pool.  Baseline timing only. */
int vwo_pyramid_correlate_tiled(const vwo_corr_params* p, const vwo_corr_inputs* in,
                                int tile, int nthreads, float* dest) {
  int cols = in->lcols, rows = in->lrows;
  int tx = (cols + tile - 1) / tile, ty = (rows + tile - 1) / tile;
  int err = 0;
  if (nthreads <= 0) nthreads = vwo_max_threads();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
  for (int t = 0; t < tx * ty; ++t) {
    int x0 = (t % tx) * tile, y0 = (t / tx) * tile;
    int x1 = x0 + tile < cols ? x0 + tile : cols, y1 = y0 + tile < rows ? y0 + tile : rows;
    int rc = vwo_pyramid_correlate_rasterize(p, in, x0, y0, x1, y1, dest + ((size_t)y0 * cols + x0) * 3, cols, NULL);
    if (rc) err = rc;
  }
  return err;
}
