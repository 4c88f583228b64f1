// PyramidCorrelationView.h -- the C++ shim that drops the B200 engine into src/vw/Stereo as the
// rasteriser of vw::stereo::PyramidCorrelationView.
//
// It keeps the reference's lazy-view operator API (Stereo/CorrelationView.h:35-193):
//   * derives ImageViewBase<Self>, pixel_type = PixelMask<Vector2f>
//   * cols()/rows()/planes(), operator() throwing NoImplErr
//   * prerasterize_type = CropView<ImageView<result_type>>, prerasterize(bbox) returning an owning
//     buffer of bbox size wrapped as CropView(buf, -bbox.min, cols, rows)     (CorrelationView.cc:880-884)
//   * rasterize(dest, bbox) = vw::rasterize(prerasterize(bbox (+collar)), dest, bbox)  (CorrelationView.h:123-133)
//   * pixel_accessor = ProceduralPixelAccessor<Self>, origin()              (CorrelationView.h:42,113)
//   * the constructor and the factory b200_pyramid_correlate(...) take EXACTLY the reference's parameters, in its order
//     and with its defaults (:48-69, :195-230) -- tests/test_cpp_shim.py parses the reference header and compares
// and forwards the work to the C ABI of include/vwb200.h (no C++ types cross the boundary).
//
// With the real library: compile with -DVWB200_USE_REAL_VW inside src/vw/Stereo; otherwise the
// self-contained stand-ins of vw_standin.h (same names and layouts) are used.
#pragma once
#ifdef VWB200_USE_REAL_VW
#include <vw/Image/ImageView.h>
#include <vw/Image/ImageViewRef.h>
#include <vw/Image/Manipulation.h>
#include <vw/Image/PixelMask.h>
#include <vw/Stereo/CostFunctions.h>
#include <vw/Stereo/PrefilterEnum.h>
#include <vw/Stereo/CorrelationAlgorithms.h>
#include <vw/Stereo/SGM.h>
#include <vw/Image/PixelAccessors.h>
#else
#include "vw_standin.h"
#endif
#include "../vwb200.h"

namespace vw { namespace stereo {

namespace b200_detail {
// error code -> vw exception (Core/Exception.h:225-253)
inline void check(int rc) {
  if (rc == VWB200_OK) return;
  const std::string msg = vwb200_last_error();
  switch (rc) {
    case VWB200_EARG:    vw_throw(ArgumentErr() << msg);
    case VWB200_EMATH:   vw_throw(MathErr() << msg);
    case VWB200_ENOIMPL: vw_throw(NoImplErr() << msg);
    default:             vw_throw(LogicErr() << msg);
  }
}
struct Handle {                       // shared, so that view copies (views are passed by value) share the engine
  vwb200_corr* h = nullptr;
  ~Handle() { if (h) vwb200_corr_destroy(h); }
};
}  // namespace b200_detail

class B200PyramidCorrelationView : public ImageViewBase<B200PyramidCorrelationView> {
public:
  typedef PixelMask<Vector2i> pixel_typeI;
  typedef PixelMask<Vector2f> pixel_type;
  typedef PixelMask<Vector2f> result_type;
  typedef ProceduralPixelAccessor<B200PyramidCorrelationView> pixel_accessor;       // CorrelationView.h:42

  /// The argument list of PyramidCorrelationView, in its order and with its defaults (Stereo/CorrelationView.h:48-69).
  /// The input views are rasterised once and placed in HBM; tiles are then produced lazily by rasterize()/prerasterize().
  template <class LeftT, class RightT, class LMaskT, class RMaskT>
  B200PyramidCorrelationView(ImageViewBase<LeftT> const& left, ImageViewBase<RightT> const& right,
                             ImageViewBase<LMaskT> const& left_mask, ImageViewBase<RMaskT> const& right_mask,
                             PrefilterModeType prefilter_mode, float prefilter_width,
                             BBox2i const& search_region, Vector2i const& kernel_size,
                             stereo::CostFunctionType cost_type,
                             int corr_timeout, double seconds_per_op,
                             float consistency_threshold,
                             int min_consistency_level,
                             int filter_half_kernel,
                             int32 max_pyramid_levels,
                             CorrelationAlgorithm algorithm = VW_CORRELATION_BM,
                             int collar_size = 0,
                             SemiGlobalMatcher::SgmSubpixelMode sgm_subpixel_mode = SemiGlobalMatcher::SUBPIXEL_LC_BLEND,
                             Vector2i sgm_search_buffer = Vector2i(2, 2),
                             size_t memory_limit_mb = 6000,
                             int blob_filter_area = 0,
                             ImageView<PixelMask<float>>* lr_disp_diff = NULL,
                             Vector2i const& region_ul = Vector2i(0, 0),
                             bool write_debug_images = false)
    : m_handle(new b200_detail::Handle), m_collar_size(collar_size), m_lr_disp_diff(lr_disp_diff) {
    vwb200_corr_params p;
    std::memset(&p, 0, sizeof(p));
    p.search_x0 = search_region.min()[0]; p.search_y0 = search_region.min()[1];
    p.search_x1 = search_region.max()[0]; p.search_y1 = search_region.max()[1];
    p.kernel_x = kernel_size[0]; p.kernel_y = kernel_size[1];
    p.cost_type = int(cost_type);
    p.prefilter_mode = int(prefilter_mode); p.prefilter_width = prefilter_width;
    p.consistency_threshold = consistency_threshold; p.min_consistency_level = min_consistency_level;
    p.filter_half_kernel = filter_half_kernel; p.max_pyramid_levels = max_pyramid_levels;
    p.collar_size = collar_size; p.corr_timeout = corr_timeout; p.seconds_per_op = seconds_per_op;
    p.algorithm = int(algorithm); p.blob_filter_area = blob_filter_area;
    p.sgm_subpixel_mode = int(sgm_subpixel_mode);
    p.sgm_search_buffer_x = sgm_search_buffer[0]; p.sgm_search_buffer_y = sgm_search_buffer[1];
    p.memory_limit_mb = double(memory_limit_mb);
    p.region_ul_x = region_ul[0]; p.region_ul_y = region_ul[1];
    p.write_debug_images = write_debug_images ? 1 : 0;
    p.sgm_threads = 0;                    // default (4 = VW_NUM_THREADS of the reference build); see INTEGRATION.md
    b200_detail::check(vwb200_corr_create(&p, &m_handle->h));
    // rasterise the (possibly lazy) inputs once; PixelGray<float> and float share their layout
    ImageView<PixelGray<float>> l = left.impl(), r = right.impl();
    ImageView<uint8> lm = left_mask.impl(), rm = right_mask.impl();
    static_assert(sizeof(PixelGray<float>) == sizeof(float), "PixelGray<float> must be a bare float");
    if (lm.cols() != l.cols() || lm.rows() != l.rows() || rm.cols() != r.cols() || rm.rows() != r.rows())
      vw_throw(ArgumentErr() << "B200PyramidCorrelationView: masks must have the size of their images.");
    b200_detail::check(vwb200_corr_set_inputs(m_handle->h,
        reinterpret_cast<const float*>(l.data()), l.cols(), l.rows(), l.cols(),
        reinterpret_cast<const float*>(r.data()), r.cols(), r.rows(), r.cols(),
        lm.data(), lm.cols(), rm.data(), rm.cols(), /*on_device=*/0));
    if (lr_disp_diff) {                   // caller-owned, written at disjoint tile windows (CorrelationView.cc:276-283, 848-857)
      static_assert(sizeof(PixelMask<float>) == 8, "PixelMask<float> must be {value, valid} floats");
      b200_detail::check(vwb200_corr_set_lr_disp_diff(m_handle->h, reinterpret_cast<float*>(lr_disp_diff->data()), lr_disp_diff->cols(),
                                                      lr_disp_diff->rows(), lr_disp_diff->cols(), /*on_device=*/0));
    }
  }

  // Standard required ImageView interfaces (CorrelationView.h:109-117)
  inline int32 cols() const { return vwb200_corr_cols(m_handle->h); }
  inline int32 rows() const { return vwb200_corr_rows(m_handle->h); }
  inline int32 planes() const { return 1; }
  inline pixel_accessor origin() const { return pixel_accessor(*this, 0, 0); }
  inline result_type operator()(int32 /*i*/, int32 /*j*/, int32 /*p*/ = 0) const {
    vw_throw(NoImplErr() << "NewCorrelationView::operator() is not implemented.");
    return result_type();
  }

  /// Block rasterization section that does actual work: prerasterize(bbox) processes exactly bbox and returns an owning
  /// buffer wrapped as CropView(buf, -bbox.min, cols, rows) (CorrelationView.cc:880-884)
  typedef CropView<ImageView<result_type>> prerasterize_type;
  inline prerasterize_type prerasterize(BBox2i const& bbox) const {
    static_assert(sizeof(result_type) == 12, "PixelMask<Vector2f> must be {dx, dy, valid} floats");
    ImageView<result_type> buf(bbox.width(), bbox.height());
    b200_detail::check(vwb200_corr_prerasterize(m_handle->h, bbox.min()[0], bbox.min()[1], bbox.max()[0], bbox.max()[1],
                                                reinterpret_cast<float*>(buf.data()), buf.cols(), /*dest_on_device=*/0, nullptr));
    return prerasterize_type(buf, -bbox.min()[0], -bbox.min()[1], cols(), rows());
  }

  template <class DestT>
  inline void rasterize(DestT const& dest, BBox2i const& bbox) const {          // CorrelationView.h:123-133
    BBox2i proc_bbox = bbox;
    if (m_collar_size > 0) proc_bbox.expand(m_collar_size);
    vw::rasterize(prerasterize(proc_bbox), dest, bbox);
  }

private:
  std::shared_ptr<b200_detail::Handle> m_handle;
  int m_collar_size;
  ImageView<PixelMask<float>>* m_lr_disp_diff;
};

/// vw::stereo::pyramid_correlate's twin: same parameters, order and defaults (Stereo/CorrelationView.h:195-230)
template <class LeftT, class RightT, class LMaskT, class RMaskT>
inline B200PyramidCorrelationView
b200_pyramid_correlate(ImageViewBase<LeftT> const& left, ImageViewBase<RightT> const& right,
                       ImageViewBase<LMaskT> const& left_mask, ImageViewBase<RMaskT> const& right_mask,
                       PrefilterModeType prefilter_mode, float prefilter_width,
                       BBox2i const& search_region, Vector2i const& kernel_size,
                       stereo::CostFunctionType cost_type,
                       int corr_timeout, double seconds_per_op,
                       float consistency_threshold,
                       int min_consistency_level,
                       int filter_half_kernel,
                       int32 max_pyramid_levels,
                       CorrelationAlgorithm algorithm = VW_CORRELATION_BM,
                       int collar_size = 0,
                       SemiGlobalMatcher::SgmSubpixelMode sgm_subpixel_mode = SemiGlobalMatcher::SUBPIXEL_LC_BLEND,
                       Vector2i sgm_search_buffer = Vector2i(2, 2),
                       size_t memory_limit_mb = 6000,
                       int blob_filter_area = 0,
                       ImageView<PixelMask<float>>* lr_disp_diff = NULL,
                       Vector2i const& region_ul = Vector2i(0, 0),
                       bool write_debug_images = false) {
  return B200PyramidCorrelationView(left, right, left_mask, right_mask, prefilter_mode, prefilter_width, search_region,
                                    kernel_size, cost_type, corr_timeout, seconds_per_op, consistency_threshold,
                                    min_consistency_level, filter_half_kernel, max_pyramid_levels, algorithm,
                                    collar_size, sgm_subpixel_mode, sgm_search_buffer, memory_limit_mb, blob_filter_area,
                                    lr_disp_diff, region_ul, write_debug_images);
}

/// PixelMask<Vector2i> from the integer pixel of the C ABI: vwb200_dispi.valid is 0 / 1, PixelMask<Vector2i>'s valid
/// channel is ChannelRange<int32>::max() = INT_MAX (Image/PixelTypeInfo.h:95-101)
inline PixelMask<Vector2i> to_pixel_mask(vwb200_dispi const& d) {
  PixelMask<Vector2i> p(Vector2i(d.dx, d.dy));
  if (!d.valid) p.invalidate();
  return p;
}

}}  // namespace vw::stereo
