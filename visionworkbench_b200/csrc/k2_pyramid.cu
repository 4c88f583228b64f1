// k2_pyramid.cu -- the Gaussian-pyramid feeder of PyramidCorrelationView::build_image_pyramids
// (Stereo/CorrelationView.cc:67-239): edge-extended ROI crops, masked-mean fill, 5-tap binomial
// smoothing fused with the 2x subsample, mask reduction.
//
// Bit-exactness: the reference evaluates subsample(separable_convolution_filter(img,k,k),2) as a
// float row pass into a float work image followed by a float column pass, each output being
//   r = 0; for t in 0..4: r += k[4-t] * src[t]            (Image/Convolution.h:56-65,318)
// on an SSE build without FMA.  The kernels below issue the same __fmul_rn/__fadd_rn sequence, so
// every pyramid pixel is bit-identical -- but only the pixels the subsample keeps are computed
// (the reference computes the full-resolution convolution and throws 3/4 away,
// Image/Manipulation.h:256-284).
#include "common.cuh"

namespace vwb200 {

__device__ __forceinline__ int clampi2(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- crop(edge_extend(img, ConstantEdgeExtension()), region)  (Image/EdgeExtension.tcc:47-62) ----------
__global__ void crop_extend_f32_kernel(ImgF src, int x0, int y0, int w, int h, float* __restrict__ dst, ptrdiff_t dpitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  dst[(ptrdiff_t)y * dpitch + x] = __ldg(src.p + (ptrdiff_t)clampi2(y0 + y, 0, src.h - 1) * src.pitch + clampi2(x0 + x, 0, src.w - 1));
}
__global__ void crop_extend_u8_kernel(ImgB src, int x0, int y0, int w, int h, int zero_outside, uint8_t* __restrict__ dst, ptrdiff_t dpitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int sx = x0 + x, sy = y0 + y;
  uint8_t v;
  if (zero_outside && (sx < 0 || sx >= src.w || sy < 0 || sy >= src.h)) v = 0;       // ZeroEdgeExtension
  else v = src.p[(ptrdiff_t)clampi2(sy, 0, src.h - 1) * src.pitch + clampi2(sx, 0, src.w - 1)];
  dst[(ptrdiff_t)y * dpitch + x] = v;
}
int crop_extend_f32_launch(ImgF src, int x0, int y0, int w, int h, float* dst, ptrdiff_t dpitch, cudaStream_t st) {
  if (w <= 0 || h <= 0) return VWB200_OK;
  dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
  crop_extend_f32_kernel<<<g, b, 0, st>>>(src, x0, y0, w, h, dst, dpitch);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}
int crop_extend_u8_launch(ImgB src, int x0, int y0, int w, int h, int zero_outside, uint8_t* dst, ptrdiff_t dpitch, cudaStream_t st) {
  if (w <= 0 || h <= 0) return VWB200_OK;
  dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
  crop_extend_u8_kernel<<<g, b, 0, st>>>(src, x0, y0, w, h, zero_outside, dst, dpitch);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// ---- masked mean over every 2nd pixel (CorrelationView.cc:133-136; Math/Functors.h:469-487) ----------
// Deterministic two-stage double reduction (sum, count).  Sums of integer-valued imagery are exact.
static constexpr int MM_BLOCKS = 256;
__global__ void masked_mean_stage1(ImgF img, ImgB mask, double* __restrict__ partial) {
  const int ow = 1 + (img.w - 1) / 2, oh = 1 + (img.h - 1) / 2;
  const long long n = (long long)ow * oh;
  double s = 0.0, c = 0.0;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(k % ow), j = (int)(k / ow);
    if (mask.p[(ptrdiff_t)(2 * j) * mask.pitch + 2 * i]) { s += (double)img.p[(ptrdiff_t)(2 * j) * img.pitch + 2 * i]; c += 1.0; }
  }
  __shared__ double ss[256], sc[256];
  ss[threadIdx.x] = s; sc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { ss[threadIdx.x] += ss[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = ss[0]; partial[2 * blockIdx.x + 1] = sc[0]; }
}
__global__ void masked_mean_stage2(const double* __restrict__ partial, int n, double* __restrict__ acc2) {
  __shared__ double ss[256], sc[256];
  ss[threadIdx.x] = threadIdx.x < n ? partial[2 * threadIdx.x] : 0.0;
  sc[threadIdx.x] = threadIdx.x < n ? partial[2 * threadIdx.x + 1] : 0.0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { ss[threadIdx.x] += ss[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { acc2[0] = ss[0]; acc2[1] = sc[0]; }
}
// d_acc2 must have room for 2 + 2*MM_BLOCKS doubles: [0..1] result, rest scratch.
int masked_mean_launch(ImgF img, ImgB mask, double* d_acc2, cudaStream_t st) {
  masked_mean_stage1<<<MM_BLOCKS, 256, 0, st>>>(img, mask, d_acc2 + 2);
  VWB_LAUNCH_CHECK();
  masked_mean_stage2<<<1, 256, 0, st>>>(d_acc2 + 2, MM_BLOCKS, d_acc2);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}
// masked pixels <- (float)(sum/count)   (CorrelationView.cc:144-149).  No-op when count == 0 (the host
// checks that case and returns the all-invalid tile, :137-142).
__global__ void mean_fill_kernel(float* __restrict__ img, int w, int h, ptrdiff_t pitch, ImgB mask, const double* __restrict__ acc2) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const double cnt = acc2[1];
  if (cnt == 0.0) return;
  if (!mask.p[(ptrdiff_t)y * mask.pitch + x]) img[(ptrdiff_t)y * pitch + x] = (float)(__ddiv_rn(acc2[0], cnt));
}
int mean_fill_launch(float* img, int w, int h, ptrdiff_t pitch, ImgB mask, const double* d_acc2, cudaStream_t st) {
  dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
  mean_fill_kernel<<<g, b, 0, st>>>(img, w, h, pitch, mask, d_acc2);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// *flag = 1 if the mask holds a zero (some pixel of the tile will be mean-filled: the imagery is then not integer-valued
// everywhere even when the rasters are).  The caller zeroes the flag.
__global__ void mask_any_zero_kernel(ImgB mask, int* __restrict__ flag) {
  const long long n = (long long)mask.w * mask.h;
  bool z = false;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    z |= mask.p[(ptrdiff_t)(k / mask.w) * mask.pitch + (k % mask.w)] == 0;
  if (__any_sync(0xffffffffu, z) && (threadIdx.x & 31) == 0) *flag = 1;
}
int mask_any_zero_launch(ImgB mask, int* d_flag, cudaStream_t st) {
  mask_any_zero_kernel<<<148, 256, 0, st>>>(mask, d_flag);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// ---- fused 5-tap separable smoothing + 2x subsample ---------------------------------------------------
// CTA computes a PD_TW x PD_TH tile of OUTPUT pixels.  Stage A: row-convolved values at the even
// columns for the (2*PD_TH+3) source rows the tile needs go to shared memory (each computed once per
// CTA, coalesced reads of the source rows).  Stage B: column pass from shared memory.
static constexpr int PD_TW = 32, PD_TH = 16;
__global__ void __launch_bounds__(PD_TW * PD_TH)
pyramid_down_kernel(ImgF in, float* __restrict__ out, ptrdiff_t opitch, int ow, int oh) {
  __shared__ float rowc[2 * PD_TH + 3][PD_TW + 1];
  const float k0 = (float)(1.0 / 16.0), k1 = (float)(4.0 / 16.0), k2 = (float)(6.0 / 16.0);   // Image/Filter.h:94-98
  const int ox0 = blockIdx.x * PD_TW, oy0 = blockIdx.y * PD_TH;
  const int tid = threadIdx.y * PD_TW + threadIdx.x;
  const int nrows = 2 * PD_TH + 3;
  for (int item = tid; item < nrows * PD_TW; item += PD_TW * PD_TH) {
    const int r = item / PD_TW, c = item % PD_TW;
    const int sy = clampi2(2 * oy0 - 2 + r, 0, in.h - 1);           // ConstantEdgeExtension on the ROI image
    const int cx = 2 * (ox0 + c);
    const float* row = in.p + (ptrdiff_t)sy * in.pitch;
    const float s0 = __ldg(row + clampi2(cx - 2, 0, in.w - 1));
    const float s1 = __ldg(row + clampi2(cx - 1, 0, in.w - 1));
    const float s2 = __ldg(row + clampi2(cx, 0, in.w - 1));
    const float s3 = __ldg(row + clampi2(cx + 1, 0, in.w - 1));
    const float s4 = __ldg(row + clampi2(cx + 2, 0, in.w - 1));
    float a = __fadd_rn(0.0f, __fmul_rn(k0, s0));                   // kernel.rbegin(): k[4], k[3], ... (symmetric)
    a = __fadd_rn(a, __fmul_rn(k1, s1));
    a = __fadd_rn(a, __fmul_rn(k2, s2));
    a = __fadd_rn(a, __fmul_rn(k1, s3));
    a = __fadd_rn(a, __fmul_rn(k0, s4));
    rowc[r][c] = a;
  }
  __syncthreads();
  const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
  if (ox < ow && oy < oh) {
    const int r = 2 * threadIdx.y;
    float a = __fadd_rn(0.0f, __fmul_rn(k0, rowc[r][threadIdx.x]));
    a = __fadd_rn(a, __fmul_rn(k1, rowc[r + 1][threadIdx.x]));
    a = __fadd_rn(a, __fmul_rn(k2, rowc[r + 2][threadIdx.x]));
    a = __fadd_rn(a, __fmul_rn(k1, rowc[r + 3][threadIdx.x]));
    a = __fadd_rn(a, __fmul_rn(k0, rowc[r + 4][threadIdx.x]));
    out[(ptrdiff_t)oy * opitch + ox] = a;
  }
}
int pyramid_down_launch(ImgF in, float* out, ptrdiff_t opitch, cudaStream_t st) {
  if (in.w <= 0 || in.h <= 0) return VWB200_OK;
  const int ow = 1 + (in.w - 1) / 2, oh = 1 + (in.h - 1) / 2;     // Image/Manipulation.h:238-243
  dim3 b(PD_TW, PD_TH), g((ow + PD_TW - 1) / PD_TW, (oh + PD_TH - 1) / PD_TH);
  pyramid_down_kernel<<<g, b, 0, st>>>(in, out, opitch, ow, oh);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// ---- SubsampleMaskByTwoFunc (CorrelationView.cc:38-63) over ZeroEdgeExtension ------------------------
__global__ void subsample_mask_kernel(ImgB in, uint8_t* __restrict__ out, ptrdiff_t opitch, int ow, int oh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= ow || j >= oh) return;
  int count = 0;
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int x = 2 * i + a, y = 2 * j + b;
      if (x < in.w && y < in.h && in.p[(ptrdiff_t)y * in.pitch + x]) ++count;
    }
  out[(ptrdiff_t)j * opitch + i] = count > 1 ? 255 : 0;
}
int subsample_mask_launch(ImgB in, uint8_t* out, ptrdiff_t opitch, cudaStream_t st) {
  if (in.w <= 0 || in.h <= 0) return VWB200_OK;
  const int ow = 1 + (in.w - 1) / 2, oh = 1 + (in.h - 1) / 2;
  dim3 b(32, 8), g((ow + 31) / 32, (oh + 7) / 8);
  subsample_mask_kernel<<<g, b, 0, st>>>(in, out, opitch, ow, oh);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}


// ---- stereo pre-filters (Stereo/PreFilter.h:45-95) ----------------------------------------------------------------
// gaussian_filter = SeparableConvolutionView with ConstantEdgeExtension: row pass into a float work image, then the
// column pass, each output "r = 0; for i: r += k[n-1-i] * src[i]" in float (Image/Convolution.h:56-65,286-289,318).
// pass 0: rows (reads in, clamped x); pass 1: columns (reads the work image, clamped y).  Taps in global memory.
__global__ void sepconv_pass_kernel(ImgF in, const float* __restrict__ taps, int n, int vertical, float* __restrict__ out, ptrdiff_t opitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= in.w || y >= in.h) return;
  const int c = (n - 1) / 2, lo = n - c - 1;         // child bbox grows by (n-c-1) before, c after  (:281-283)
  float r = 0.0f;
  if (!vertical) {
    const float* row = in.p + (ptrdiff_t)y * in.pitch;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, __fmul_rn(taps[n - 1 - i], __ldg(row + clampi2(x - lo + i, 0, in.w - 1))));
  } else {
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, __fmul_rn(taps[n - 1 - i], __ldg(in.p + (ptrdiff_t)clampi2(y - lo + i, 0, in.h - 1) * in.pitch + x)));
  }
  out[(ptrdiff_t)y * opitch + x] = r;
}
int sepconv_launch(ImgF in, const float* d_taps, int n, float* work, float* out, cudaStream_t st) {
  if (in.w <= 0 || in.h <= 0) return VWB200_OK;
  dim3 b(32, 8), g((in.w + 31) / 32, (in.h + 7) / 8);
  // the reference convolves rows over the child's FULL padded height first; clamping the row index in the
  // column pass reads the same values because the row pass of a clamped row equals the clamped row's result.
  sepconv_pass_kernel<<<g, b, 0, st>>>(in, d_taps, n, 0, work, in.w);
  VWB_LAUNCH_CHECK();
  sepconv_pass_kernel<<<g, b, 0, st>>>(ImgF{work, in.w, in.h, in.w}, d_taps, n, 1, out, in.w);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}
// mode 1: 3x3 Laplacian of g (order: top, left, centre*-4, right, bottom; Image/Filter.h:318-326, Convolution.h:68-91)
// mode 2: img - g  (SubtractedMean, PreFilter.h:60-70)
__global__ void prefilter_final_kernel(ImgF img, const float* __restrict__ g, int mode, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= img.w || y >= img.h) return;
  const int w = img.w, h = img.h;
  float r;
  if (mode == 2) {
    r = __fsub_rn(img.p[(ptrdiff_t)y * img.pitch + x], g[(ptrdiff_t)y * w + x]);
  } else {
    const int xm = clampi2(x - 1, 0, w - 1), xp = clampi2(x + 1, 0, w - 1), ym = clampi2(y - 1, 0, h - 1), yp = clampi2(y + 1, 0, h - 1);
    r = __fadd_rn(0.0f, __fmul_rn(1.0f, g[(ptrdiff_t)ym * w + x]));
    r = __fadd_rn(r, __fmul_rn(1.0f, g[(ptrdiff_t)y * w + xm]));
    r = __fadd_rn(r, __fmul_rn(-4.0f, g[(ptrdiff_t)y * w + x]));
    r = __fadd_rn(r, __fmul_rn(1.0f, g[(ptrdiff_t)y * w + xp]));
    r = __fadd_rn(r, __fmul_rn(1.0f, g[(ptrdiff_t)yp * w + x]));
  }
  out[(ptrdiff_t)y * w + x] = r;
}
int prefilter_final_launch(ImgF img, const float* g, int mode, float* out, cudaStream_t st) {
  if (img.w <= 0 || img.h <= 0) return VWB200_OK;
  dim3 b(32, 8), gr((img.w + 31) / 32, (img.h + 7) / 8);
  prefilter_final_kernel<<<gr, b, 0, st>>>(img, g, mode, out);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

}  // namespace vwb200
