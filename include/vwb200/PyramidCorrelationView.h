// PyramidCorrelationView.h -- the C++ shim that drops the B200 engine into src/vw/Stereo as the
// rasteriser of vw::stereo::PyramidCorrelationView.
//
// It keeps the reference's lazy-view operator API (Stereo/CorrelationView.h:35-193):
//   * derives ImageViewBase<Self>, pixel_type = PixelMask<Vector2f>
//   * cols()/rows()/planes(), operator() throwing NoImplErr
//   * prerasterize_type = CropView<ImageView<result_type>>, prerasterize(bbox) returning an owning
//     buffer of bbox size wrapped as CropView(buf, -bbox.min, cols, rows)     (CorrelationView.cc:880-884)
//   * rasterize(dest, bbox) = vw::rasterize(prerasterize(bbox (+collar)), dest, bbox)  (CorrelationView.h:123-133)
//   * the factory b200_pyramid_correlate(...) with pyramid_correlate's argument list (:195-230)
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

  /// Same argument list as PyramidCorrelationView (Stereo/CorrelationView.h:48-69).  The input views are
  /// rasterised once and placed in HBM; tiles are then produced lazily by rasterize()/prerasterize().
  template <class LeftT, class RightT, class LMaskT, class RMaskT>
  B200PyramidCorrelationView(ImageViewBase<LeftT> const& left, ImageViewBase<RightT> const& right,
                             ImageViewBase<LMaskT> const& left_mask, ImageViewBase<RMaskT> const& right_mask,
                             PrefilterModeType prefilter_mode, float prefilter_width,
                             BBox2i const& search_region, Vector2i const& kernel_size,
                             CostFunctionType cost_type, int corr_timeout, double seconds_per_op,
                             float consistency_threshold, int min_consistency_level, int filter_half_kernel,
                             int32 max_pyramid_levels, CorrelationAlgorithm algorithm = VW_CORRELATION_BM,
                             int collar_size = 0, int blob_filter_area = 0)
    : m_handle(new b200_detail::Handle) {
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
    m_collar = collar_size;
  }

  // Standard required ImageView interfaces (CorrelationView.h:109-117)
  inline int32 cols() const { return vwb200_corr_cols(m_handle->h); }
  inline int32 rows() const { return vwb200_corr_rows(m_handle->h); }
  inline int32 planes() const { return 1; }
  inline result_type operator()(int32 /*i*/, int32 /*j*/, int32 /*p*/ = 0) const {
    vw_throw(NoImplErr() << "NewCorrelationView::operator() is not implemented.");
    return result_type();
  }

  /// Block rasterization section that does actual work
  typedef CropView<ImageView<result_type>> prerasterize_type;
  inline prerasterize_type prerasterize(BBox2i const& bbox) const { return run(bbox, bbox); }

  template <class DestT>
  inline void rasterize(DestT const& dest, BBox2i const& bbox) const {
    // the collar is applied inside the engine (vwb200_corr_rasterize); the returned buffer covers bbox
    vw::rasterize(run(bbox, bbox), dest, bbox);
  }

private:
  prerasterize_type run(BBox2i const& bbox, BBox2i const&) const {
    static_assert(sizeof(result_type) == 12, "PixelMask<Vector2f> must be {dx, dy, valid} floats");
    ImageView<result_type> buf(bbox.width(), bbox.height());
    b200_detail::check(vwb200_corr_rasterize(m_handle->h, bbox.min()[0], bbox.min()[1], bbox.max()[0], bbox.max()[1],
                                             reinterpret_cast<float*>(buf.data()), buf.cols(), /*dest_on_device=*/0, nullptr));
    return prerasterize_type(buf, -bbox.min()[0], -bbox.min()[1], cols(), rows());
  }
  std::shared_ptr<b200_detail::Handle> m_handle;
  int m_collar = 0;
};

/// vw::stereo::pyramid_correlate's twin (Stereo/CorrelationView.h:195-230)
template <class LeftT, class RightT, class LMaskT, class RMaskT>
inline B200PyramidCorrelationView
b200_pyramid_correlate(ImageViewBase<LeftT> const& left, ImageViewBase<RightT> const& right,
                       ImageViewBase<LMaskT> const& left_mask, ImageViewBase<RMaskT> const& right_mask,
                       PrefilterModeType prefilter_mode, float prefilter_width,
                       BBox2i const& search_region, Vector2i const& kernel_size,
                       CostFunctionType cost_type, int corr_timeout, double seconds_per_op,
                       float consistency_threshold, int min_consistency_level, int filter_half_kernel,
                       int32 max_pyramid_levels, CorrelationAlgorithm algorithm = VW_CORRELATION_BM,
                       int collar_size = 0, int blob_filter_area = 0) {
  return B200PyramidCorrelationView(left, right, left_mask, right_mask, prefilter_mode, prefilter_width, search_region,
                                    kernel_size, cost_type, corr_timeout, seconds_per_op, consistency_threshold,
                                    min_consistency_level, filter_half_kernel, max_pyramid_levels, algorithm,
                                    collar_size, blob_filter_area);
}

}}  // namespace vw::stereo
