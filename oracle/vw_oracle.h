/*
 * vw_oracle.h -- CPU restatement ("oracle") of Vision Workbench's dense
 * block-matching stereo correlation hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke test in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (visionworkbench_b200/) never links or calls it.
 *
 * Every function cites the reference file:line (relative to the reference's
 * src/vw/) whose loop structure and arithmetic it follows.  The reference
 * cannot be compiled in this environment (no Boost / GDAL / LAPACK headers), so
 * the oracle is pinned by the reference's own known-answer tests instead
 * (tests/test_oracle_kat.py lists each test file:line it reproduces).
 *
 * Conventions: images are row-major, single plane, pitch in ELEMENTS.
 * Boxes are half-open [x0,x1) x [y0,y1) like vw::BBox2i (Math/BBox.tcc).
 * Integer disparity pixels are {dx, dy, valid} int32 triples
 * (PixelMask<Vector2i>, Image/PixelMask.h:48-160; valid is 0 / 1 here).
 */
#ifndef VW_ORACLE_H
#define VW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VWO_COST_ABS = 0, VWO_COST_SQ = 1, VWO_COST_NCC = 2 };   /* Stereo/CostFunctions.h:143-149 */
enum { VWO_PREFILTER_NONE = 0, VWO_PREFILTER_LOG = 1, VWO_PREFILTER_MEANSUB = 2 }; /* Stereo/PrefilterEnum.h:24-28 */

typedef struct { int32_t dx, dy, valid; } vwo_disp_t;

/* Stereo/Algorithms.h:41-129 : sliding kx x ky box sum, double accumulators. */
int vwo_fast_box_sum(const double* in, int w, int h, int pitch, int kx, int ky, double* out);

/* Stereo/Correlation.cc:330-375 (+ :33-137): left is (W+kx-1)x(H+ky-1), right is
 * (W+kx-1+sx-1)x(H+ky-1+sy-1); out is WxH. */
int vwo_calc_disparity(int cost, const float* left, int lw, int lh, int lpitch,
                       const float* right, int rw, int rh, int rpitch,
                       int sx, int sy, int kx, int ky, vwo_disp_t* out);

/* Image/Convolution.h:275-328 : separable convolution of the whole image with
 * ConstantEdgeExtension (or zero: edge=1), kernel origin at (n-1)/2. */
int vwo_separable_convolve(const float* in, int w, int h, int pitch,
                           const float* kx, int nx, const float* ky, int ny,
                           int edge_zero, float* out);

/* same with an explicit kernel origin (cx,cy); <0 = default (n-1)/2 */
int vwo_separable_convolve_c(const float* in, int w, int h, int pitch,
                             const float* kx, int nx, const float* ky, int ny,
                             int cx, int cy, int edge_zero, float* out);
/* per-pixel cost functor value (Stereo/CostFunctions.h:72-141), float op widened to double */
double vwo_cost_pixel(int cost, float a, float b);
int vwo_gaussian_kernel(double sigma, float* k, int maxn);

/* Image/Manipulation.h:238-251 */
int vwo_subsample2_f32(const float* in, int w, int h, int pitch, float* out /* (1+(w-1)/2) x (1+(h-1)/2) */);
/* Stereo/CorrelationView.cc:38-63 */
int vwo_subsample_mask_by_two(const uint8_t* in, int w, int h, uint8_t* out);

/* one pyramid level: subsample(separable_convolution_filter(in,{1,4,6,4,1}/16),2)
 * Stereo/CorrelationView.cc:210-214 */
int vwo_pyramid_down(const float* in, int w, int h, float* out);

/* Stereo/Correlate.cc:1441-1502 : in place on l2r */
int vwo_cross_corr_consistency_check(vwo_disp_t* l2r, int lw, int lh, int lpitch,
                                     const vwo_disp_t* r2l, int rw, int rh, float threshold);

/* Stereo/DisparityMap.h:318-442 */
int vwo_rm_outliers_using_thresh(const vwo_disp_t* in, int w, int h, int hx, int hy,
                                 double pixel_thresh, double rej_thresh, vwo_disp_t* out);
int vwo_disparity_cleanup_using_thresh(const vwo_disp_t* in, int w, int h, int hx, int hy,
                                       double pixel_thresh, double rej_thresh, vwo_disp_t* out);
/* Stereo/DisparityMap.h:97-253 */
int vwo_disparity_mask(const vwo_disp_t* in, int w, int h, const uint8_t* lmask,
                       const uint8_t* rmask, int rmw, int rmh, vwo_disp_t* out);

/* Stereo/Correlation.cc:139-328.  zones_out: 8 ints per zone
 * {img x0,y0,x1,y1, disp x0,y0,x1,y1}.  Returns number of zones (or -1 if more than max). */
int vwo_subdivide_regions(const vwo_disp_t* disp, int w, int h, int kx, int ky,
                          int32_t* zones_out, int max_zones);

/* Stereo/PreFilter.h:45-95 */
int vwo_prefilter(const float* in, int w, int h, int mode, float width, float* out);

/* ---- the lazy view: Stereo/CorrelationView.h:35-230, CorrelationView.cc:67-239,273-886 ---- */
typedef struct {
  int32_t search_x0, search_y0, search_x1, search_y1;  /* BBox2i search_region */
  int32_t kernel_x, kernel_y;
  int32_t cost_type;
  int32_t prefilter_mode; float prefilter_width;
  float   consistency_threshold;     /* < 0: no L/R check */
  int32_t min_consistency_level;     /* unused by BM */
  int32_t filter_half_kernel;
  int32_t max_pyramid_levels;
  int32_t collar_size;
  /* the rest of the constructor (CorrelationView.h:60-69) */
  int32_t algorithm;                 /* CorrelationAlgorithm: 0 BM, 1 SGM, 2 MGM, 3 FINAL_MGM */
  int32_t sgm_subpixel_mode;         /* SgmSubpixelMode (SGM.h:93-99) */
  int32_t sgm_search_buffer_x, sgm_search_buffer_y;
  int32_t blob_filter_area;
  int32_t sgm_threads;               /* vw_settings().default_num_threads() of the SGM memory estimate */
  double  memory_limit_mb;
} vwo_corr_params;

typedef struct {
  const float* left;  int lcols, lrows, lpitch;
  const float* right; int rcols, rrows, rpitch;
  const uint8_t* lmask; int lmpitch;   /* same size as left  */
  const uint8_t* rmask; int rmpitch;   /* same size as right */
} vwo_corr_inputs;

/* PyramidCorrelationView::rasterize(dest, bbox): dest is bw x bh pixels of 3 floats
 * {dx, dy, valid(0/1)} with row pitch dest_pitch (in pixels).
 * Returns 0, or <0 on error.  *levels_out (optional) = pyramid levels used. */
int vwo_pyramid_correlate_rasterize(const vwo_corr_params* p, const vwo_corr_inputs* in,
                                    int bx0, int by0, int bx1, int by1,
                                    float* dest, int dest_pitch, int* levels_out);

/* ... with the optional lr_disp_diff output (CorrelationView.h:67-68, .cc:276-283,848-857): diff is a caller-owned
 * diff_cols x diff_rows image of PixelMask<float> = {value, valid} float pairs whose (0,0) sits at region_ul in image
 * coordinates; it receives max(|dx_lr + dx_rl|, |dy_lr + dy_rl|) at the pixels that pass the L/R check (level 0) and is
 * invalidated where the final disparity is invalid.  Returns -1 (ArgumentErr) when the processed box is not inside. */
int vwo_pyramid_correlate_rasterize_ex(const vwo_corr_params* p, const vwo_corr_inputs* in,
                                       int bx0, int by0, int bx1, int by1, float* dest, int dest_pitch, int* levels_out,
                                       float* diff, int diff_cols, int diff_rows, int region_ul_x, int region_ul_y);

/* PyramidCorrelationView::disparity_blob_filter (CorrelationView.cc:242-271; BlobIndexThreaded, Image/BlobIndex.h:385-454;
 * ErodeView, Image/ErodeView.h:196-218): 8-connected blobs of valid pixels with at most `area` pixels become {0,0,invalid}. */
int vwo_disparity_blob_filter(vwo_disp_t* d, int w, int h, int area);

/* the whole of calc_disparity_sgm (vw_sgm_oracle.c) */
int vwo_calc_disparity_sgm(const float* left_f, int lw, int lh, int lpitch, const float* right_f, int rw, int rh, int rpitch,
                           int search_x, int search_y, int kernel_size, int cost_type, int ternary_threshold, int p1, int p2, int use_mgm,
                           int subpixel_mode, int buffer_x, int buffer_y, double memory_limit_mb, int threads,
                           const uint8_t* lmask, const uint8_t* rmask, int rmw, int rmh, const int* prev, int pw, int ph,
                           int* bounds_io, int use_given_bounds, int* out, float* out_sub, int* out_w, int* out_h);
uint64_t vwo_census_value(const uint8_t* img, int w, int col, int row, int k, int ternary, int thr);

/* Debug/intermediate taps for kernel-level parity tests: build the per-tile pyramids
 * exactly as build_image_pyramids does (Stereo/CorrelationView.cc:67-239).
 * Caller passes arrays of max_levels+1 pointers which the oracle mallocs (free with vwo_free). */
int vwo_build_pyramids(const vwo_corr_params* p, const vwo_corr_inputs* in,
                       int bx0, int by0, int bx1, int by1, int levels,
                       float** lpyr, float** rpyr, uint8_t** lmpyr, uint8_t** rmpyr,
                       int32_t* dims /* 8 ints per level: lw,lh,rw,rh,lmw,lmh,rmw,rmh */);
void vwo_free(void* p);

/* vw::stereo::ParabolaSubpixelView (Stereo/ParabolaSubpixelView.h:27-117, .cc:31-330): disp is the integer
 * disparity as cols x rows PixelMask<Vector2f> triples; out is the bbox-sized refined disparity. */
int vwo_parabola_subpixel(const float* disp, int cols, int rows, const float* left, int lpitch,
                          const float* right, int rcols, int rrows, int rpitch,
                          int kx, int ky, int prefilter_mode, float prefilter_width,
                          int bx0, int by0, int bx1, int by1, float* out);

/* vw::stereo::calc_disparity_sgm (Stereo/SGM.cc:167-230) -> SemiGlobalMatcher::semi_global_matching_func (:2387-2448):
 * CENSUS_TRANSFORM (kernel 3/5/7/9), SGM accumulation along 8 directions (SSE flavour of evaluate_path), integer winner of
 * select_best_disparity; the same search box [0, search_x] x [0, search_y] (inclusive) for every pixel.  See
 * vw_sgm_oracle.c for what is and is not restated.  out: out_w x out_h {dx, dy, valid} triples (caller allocates lw*lh*3). */
int vwo_sgm_calc_disparity(const float* left, int lw, int lh, int lpitch, const float* right, int rw, int rh, int rpitch,
                           int search_x, int search_y, int kernel_size, int p1, int p2, int* out, int* out_w, int* out_h);

/* calc_disparity_sgm followed by SemiGlobalMatcher::create_disparity_view_subpixel (SGM.cc:1497-1614) on its integer result;
 * subpixel_mode = SgmSubpixelMode (SGM.h:93-99) except 1 (2-D parabola, not restated).  out_sub: {dx, dy, valid} floats. */
int vwo_sgm_calc_disparity_subpixel(const float* left, int lw, int lh, int lpitch, const float* right, int rw, int rh, int rpitch,
                                    int search_x, int search_y, int kernel_size, int p1, int p2, int subpixel_mode,
                                    int* out, float* out_sub, int* out_w, int* out_h);

/* ... with a search box per pixel (m_disp_bound_image): bounds = out_w * out_h quadruples {min_x, min_y, max_x, max_y},
 * inclusive, inside [0, search]; max < min marks a pixel without search area (comes out invalid).  out_sub may be NULL. */
int vwo_sgm_calc_disparity_bounds(const float* left, int lw, int lh, int lpitch, const float* right, int rw, int rh, int rpitch,
                                  int search_x, int search_y, int kernel_size, int p1, int p2, int subpixel_mode, const int* bounds,
                                  int* out, float* out_sub, int* out_w, int* out_h);

/* ... with MGM accumulation instead of SGM (accum_mgm_multithread, SGM.cc:2619-2700; SGMAssist.h:835-1239) */
int vwo_mgm_calc_disparity_bounds(const float* left, int lw, int lh, int lpitch, const float* right, int rw, int rh, int rpitch,
                                  int search_x, int search_y, int kernel_size, int p1, int p2, int subpixel_mode, const int* bounds,
                                  int* out, float* out_sub, int* out_w, int* out_h);

/* SemiGlobalMatcher::populate_disp_bound_image (SGM.cc:241-500) + one constrain_disp_bound_image pass (:502-668) at the
 * given conservation level: per-pixel search boxes from masks and the previous half-resolution disparity. */
int vwo_sgm_disp_bounds(const int* prev, int pw, int ph, const uint8_t* lmask, const uint8_t* rmask, int rmw, int rmh,
                        int ow, int oh, int search_x, int search_y, int buffer_x, int buffer_y, int conserve_level, int* bounds);

/* number of pyramid levels prerasterize would use for this bbox (CorrelationView.cc:301-310,
 * CorrelationView.h:99-105) */
int vwo_num_levels(const vwo_corr_params* p, int bw, int bh);

/* Tile-parallel driver used only for the CPU baseline timing (the reference's
 * block_write_image thread pool, Image/ImageIO.h:289-311): rasterizes every tile
 * of size tile x tile covering [0,cols)x[0,rows) with nthreads OpenMP threads. */
int vwo_pyramid_correlate_tiled(const vwo_corr_params* p, const vwo_corr_inputs* in,
                                int tile, int nthreads, float* dest /* cols*rows*3 */);
int vwo_max_threads(void);
/* calc_disparity over independent tile x tile output tiles on nthreads OpenMP threads (baseline timing) */
int vwo_calc_disparity_tiled(int cost, const float* left, int lw, int lh, int lpitch,
                             const float* right, int rw, int rh, int rpitch,
                             int sx, int sy, int kx, int ky, int tile, int nthreads, vwo_disp_t* out);

#ifdef __cplusplus
}
#endif
#endif
