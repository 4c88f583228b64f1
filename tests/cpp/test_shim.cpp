// test_shim.cpp -- exercises include/vwb200/PyramidCorrelationView.h the way Vision Workbench code
// would (crop(view, bbox) rasterised into an ImageView, from several threads) and checks the result
// against the CPU oracle.  Exit codes: 0 ok, 1 mismatch, 3 no CUDA device (expected on the CPU box).
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include "vwb200/PyramidCorrelationView.h"
#include "vwb200/ParabolaSubpixelView.h"
#include <cmath>
#include "../../oracle/vw_oracle.h"

using namespace vw;
using namespace vw::stereo;

int main() {
  const int W = 256, H = 192;
  ImageView<PixelGray<float>> left(W, H), right(W, H);
  ImageView<uint8> lmask(W, H), rmask(W, H);
  uint32_t s = 12345u;
  std::vector<float> base((W + 32) * (H + 32));
  for (auto& v : base) { s = s * 1664525u + 1013904223u; v = float((s >> 20) & 0xfff); }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      left(x, y) = base[(y + 8) * (W + 32) + x + 8];
      right(x, y) = base[(y + 8 - 3) * (W + 32) + x + 8 - 5];      // disparity (5, 3)
      lmask(x, y) = (x > 40 && x < 60 && y > 30 && y < 50) ? 0 : 255;
      rmask(x, y) = 255;
    }
  const BBox2i search(-4, -6, 16, 14);
  const Vector2i kernel(7, 7);
  try {
    // bad arguments map to the reference's exception types
    bool threw = false;
    try { b200_pyramid_correlate(left, right, lmask, rmask, PREFILTER_NONE, 0.f, search, Vector2i(6, 7), ABSOLUTE_DIFFERENCE, 0, 0.0, 2.f, 0, 3, 3); }
    catch (ArgumentErr const&) { threw = true; }
    if (!threw) { std::printf("FAIL: even kernel accepted\n"); return 1; }

    // all 24 positional arguments, in the reference's order (CorrelationView.h:195-218)
    ImageView<PixelMask<float>> lr_diff(W, H);
    B200PyramidCorrelationView view = b200_pyramid_correlate(left, right, lmask, rmask, PREFILTER_NONE, 0.f, search, kernel,
                                                             SQUARED_DIFFERENCE, 0, 0.0, 2.f, 0, 3, 3, VW_CORRELATION_BM, 0,
                                                             SemiGlobalMatcher::SUBPIXEL_LC_BLEND, Vector2i(2, 2), size_t(6000), 0,
                                                             &lr_diff, Vector2i(0, 0), false);
    { B200PyramidCorrelationView::pixel_accessor acc = view.origin(); acc.next_col(); acc.advance(1, 1); (void)acc; }
    if (view.cols() != W || view.rows() != H || view.planes() != 1) { std::printf("FAIL: dims\n"); return 1; }
    threw = false;
    try { view(0, 0); } catch (NoImplErr const&) { threw = true; }
    if (!threw) { std::printf("FAIL: operator() must throw NoImplErr\n"); return 1; }

    // tiles rasterised concurrently, like block_write_image's worker threads (Image/ImageIO.h:289-311)
    std::vector<BBox2i> boxes = {BBox2i(0, 0, 128, 96), BBox2i(128, 0, 128, 96), BBox2i(0, 96, 128, 96), BBox2i(128, 96, 128, 96)};
    std::vector<ImageView<PixelMask<Vector2f>>> tiles(boxes.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < boxes.size(); ++i)
      th.emplace_back([&, i] { tiles[i] = crop(view, boxes[i]); });        // ImageView(ViewT const&) -> rasterize(dest, bbox)
    for (auto& t : th) t.join();

    vwo_corr_params p = {search.min()[0], search.min()[1], search.max()[0], search.max()[1], 7, 7, VWO_COST_SQ, 0, 0.f, 2.f, 0, 3, 3, 0};
    vwo_corr_inputs in = {reinterpret_cast<const float*>(left.data()), W, H, W, reinterpret_cast<const float*>(right.data()), W, H, W,
                          lmask.data(), W, rmask.data(), W};
    long bad = 0, valid = 0;
    std::vector<float> ref_diff(size_t(W) * H * 2, 0.f);
    for (size_t i = 0; i < boxes.size(); ++i) {
      const BBox2i& b = boxes[i];
      std::vector<float> ref(size_t(b.width()) * b.height() * 3);
      if (vwo_pyramid_correlate_rasterize_ex(&p, &in, b.min()[0], b.min()[1], b.max()[0], b.max()[1], ref.data(), b.width(), nullptr,
                                             ref_diff.data(), W, H, 0, 0)) return 1;
      for (int y = 0; y < b.height(); ++y)
        for (int x = 0; x < b.width(); ++x) {
          const PixelMask<Vector2f>& g = tiles[i](x, y);
          const float* r = &ref[(size_t(y) * b.width() + x) * 3];
          if (g.child()[0] != r[0] || g.child()[1] != r[1] || (is_valid(g) ? 1.f : 0.f) != r[2]) ++bad;
          if (is_valid(g)) ++valid;
        }
    }
    long dvalid = 0;
    for (int y = 0; y < H; ++y)                                     // the lr_disp_diff side output (CorrelationView.cc:848-857)
      for (int x = 0; x < W; ++x) {
        const PixelMask<float>& g = lr_diff(x, y);
        const float* r = &ref_diff[(size_t(y) * W + x) * 2];
        if (g.child() != r[0] || (is_valid(g) ? 1.f : 0.f) != r[1]) ++bad;
        if (is_valid(g)) ++dvalid;
      }
    std::printf("shim: %ld mismatches, %ld valid of %d, %ld lr_disp_diff pixels\n", bad, valid, W * H, dvalid);
    if (bad || !dvalid) return 1;

    // sub-pixel refinement through the second shim (Stereo/ParabolaSubpixelView.h:28-117)
    ImageView<PixelMask<Vector2f>> disp = crop(view, BBox2i(0, 0, W, H));
    threw = false;
    try { b200_parabola_subpixel(crop(view, BBox2i(0, 0, W - 1, H)), left, right, PREFILTER_NONE, 0.f, kernel); }
    catch (ArgumentErr const&) { threw = true; }
    if (!threw) { std::printf("FAIL: size mismatch accepted\n"); return 1; }
    B200ParabolaSubpixelView sub = b200_parabola_subpixel(disp, left, right, PREFILTER_LOG, 1.4f, kernel);
    const BBox2i sb(40, 30, 150, 120);
    ImageView<PixelMask<Vector2f>> fine = crop(sub, sb);
    std::vector<float> sref(size_t(sb.width()) * sb.height() * 3);
    if (vwo_parabola_subpixel(reinterpret_cast<const float*>(disp.data()), W, H, reinterpret_cast<const float*>(left.data()), W,
                              reinterpret_cast<const float*>(right.data()), W, H, W, 7, 7, 1, 1.4f, sb.min()[0], sb.min()[1],
                              sb.max()[0], sb.max()[1], sref.data())) return 1;
    long sbad = 0, moved = 0;
    for (int y = 0; y < sb.height(); ++y)
      for (int x = 0; x < sb.width(); ++x) {
        const PixelMask<Vector2f>& g = fine(x, y);
        const float* r = &sref[(size_t(y) * sb.width() + x) * 3];
        if (std::fabs(g.child()[0] - r[0]) > 1e-5f || std::fabs(g.child()[1] - r[1]) > 1e-5f || (is_valid(g) ? 1.f : 0.f) != r[2]) ++sbad;
        if (is_valid(g) && g.child()[0] != std::floor(g.child()[0])) ++moved;
      }
    std::printf("subpixel shim: %ld mismatches, %ld refined of %d\n", sbad, moved, sb.width() * sb.height());
    return sbad ? 1 : 0;
  } catch (LogicErr const& e) {
    std::printf("NODEVICE-or-CUDA: %s\n", e.what());
    return 3;
  }
}
