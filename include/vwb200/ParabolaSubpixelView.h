// ParabolaSubpixelView.h -- C++ shim: vw::stereo::ParabolaSubpixelView's lazy-view interface
// (Stereo/ParabolaSubpixelView.h:28-117) over the C ABI entry vwb200_parabola_subpixel.
//
//   * derives ImageViewBase<Self>, pixel_type = PixelMask<Vector2f>, cols/rows/planes of the disparity
//   * operator() throws NoImplErr (:93-96); the constructor asserts disparity and left image sizes match (:66-68)
//   * prerasterize(bbox) returns an owning bbox-sized buffer wrapped as CropView(buf, -bbox.min, cols, rows)
//     (ParabolaSubpixelView.cc:247-330), rasterize(dest, bbox) = vw::rasterize(prerasterize(bbox), dest, bbox) (:102-105)
//   * factory b200_parabola_subpixel(...) with parabola_subpixel's argument list (:109-115)
// The inputs are rasterised once at construction (they are lazy views in the reference) and sent with every tile call;
// the engine fits the parabola on the device.  Floats agree with the reference within 1e-5.
#pragma once
#include "PyramidCorrelationView.h"

namespace vw { namespace stereo {

class B200ParabolaSubpixelView : public ImageViewBase<B200ParabolaSubpixelView> {
public:
  typedef PixelMask<Vector2f> pixel_type;
  typedef pixel_type result_type;

  template <class DispT, class LeftT, class RightT>
  B200ParabolaSubpixelView(ImageViewBase<DispT> const& disparity, ImageViewBase<LeftT> const& left_image,
                           ImageViewBase<RightT> const& right_image, PrefilterModeType prefilter_mode, float prefilter_width,
                           Vector2i const& kernel_size)
    : m_disparity(disparity.impl()), m_left(left_image.impl()), m_right(right_image.impl()), m_kernel(kernel_size),
      m_prefilter_mode(prefilter_mode), m_prefilter_width(prefilter_width) {
    static_assert(sizeof(pixel_type) == 12, "PixelMask<Vector2f> must be {dx, dy, valid} floats");
    static_assert(sizeof(PixelGray<float>) == sizeof(float), "PixelGray<float> must be a bare float");
    if (m_disparity.cols() != m_left.cols() || m_disparity.rows() != m_left.rows())
      vw_throw(ArgumentErr() << "SubpixelView: Disparity image must match left image.");
  }

  inline int32 cols() const { return m_disparity.cols(); }
  inline int32 rows() const { return m_disparity.rows(); }
  inline int32 planes() const { return 1; }
  inline pixel_type operator()(int32 /*i*/, int32 /*j*/, int32 /*p*/ = 0) const {
    vw_throw(NoImplErr() << "SubpixelView:operator() has not been implemented.");
    return pixel_type();
  }

  typedef CropView<ImageView<pixel_type>> prerasterize_type;
  inline prerasterize_type prerasterize(BBox2i const& bbox) const {
    ImageView<pixel_type> buf(bbox.width(), bbox.height());
    b200_detail::check(vwb200_parabola_subpixel(
        reinterpret_cast<const float*>(m_disparity.data()), m_disparity.cols(), m_disparity.rows(),
        reinterpret_cast<const float*>(m_left.data()), m_left.cols(),
        reinterpret_cast<const float*>(m_right.data()), m_right.cols(), m_right.rows(), m_right.cols(),
        m_kernel[0], m_kernel[1], int(m_prefilter_mode), m_prefilter_width,
        bbox.min()[0], bbox.min()[1], bbox.max()[0], bbox.max()[1],
        reinterpret_cast<float*>(buf.data()), buf.cols(), /*on_device=*/0, nullptr));
    return prerasterize_type(buf, -bbox.min()[0], -bbox.min()[1], cols(), rows());
  }
  template <class DestT>
  inline void rasterize(DestT const& dest, BBox2i const& bbox) const { vw::rasterize(prerasterize(bbox), dest, bbox); }

private:
  ImageView<pixel_type> m_disparity;
  ImageView<PixelGray<float>> m_left, m_right;
  Vector2i m_kernel;
  PrefilterModeType m_prefilter_mode;
  float m_prefilter_width;
};

/// vw::stereo::parabola_subpixel's twin (Stereo/ParabolaSubpixelView.h:109-115)
template <class DispT, class LeftT, class RightT>
inline B200ParabolaSubpixelView
b200_parabola_subpixel(ImageViewBase<DispT> const& disparity, ImageViewBase<LeftT> const& left_image,
                       ImageViewBase<RightT> const& right_image, PrefilterModeType prefilter_mode, float prefilter_width,
                       Vector2i const& kernel_size) {
  return B200ParabolaSubpixelView(disparity, left_image, right_image, prefilter_mode, prefilter_width, kernel_size);
}

}}  // namespace vw::stereo
