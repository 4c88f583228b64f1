// k1_fast.cu -- exact-integer fast path of the fused cost-volume + arg-best kernel (placeholder
// dispatch: filled in by the optimised kernel; until then every configuration reports "unsupported"
// and the generic fp64 kernel runs).
#include "common.cuh"

namespace vwb200 {

// ---- min / max / integer-valuedness reduction -------------------------------------------------------
__global__ void image_stats_kernel(ImgF img, float* __restrict__ result) {
  float mn = INFINITY, mx = -INFINITY; int allint = 1;
  const long long n = (long long)img.w * img.h;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const float v = img.p[(ptrdiff_t)(k / img.w) * img.pitch + (k % img.w)];
    mn = fminf(mn, v); mx = fmaxf(mx, v);
    if (!(v == rintf(v))) allint = 0;          // NaN/inf -> not integer
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    allint &= __shfl_xor_sync(0xffffffffu, allint, o);
  }
  if ((threadIdx.x & 31) == 0) {
    // float atomics on bit patterns: values are compared as floats via CAS-free min/max on ordered ints
    atomicMin(reinterpret_cast<int*>(result) + 0, mn >= 0 ? __float_as_int(mn) : (int)(0x80000000u - (unsigned)__float_as_int(mn)));
    atomicMax(reinterpret_cast<int*>(result) + 1, mx >= 0 ? __float_as_int(mx) : (int)(0x80000000u - (unsigned)__float_as_int(mx)));
    if (!allint) atomicExch(reinterpret_cast<int*>(result) + 2, 0);
  }
}
__global__ void image_stats_init(float* result) {
  reinterpret_cast<int*>(result)[0] = INT_MAX; reinterpret_cast<int*>(result)[1] = INT_MIN; reinterpret_cast<int*>(result)[2] = 1;
}
__global__ void image_stats_fini(float* result) {
  int* r = reinterpret_cast<int*>(result);
  for (int i = 0; i < 2; ++i) {
    const int o = r[i];
    result[i] = o >= 0 ? __int_as_float(o) : __int_as_float((int)(0x80000000u - (unsigned)o));
  }
  result[2] = r[2] ? 1.0f : 0.0f;
}
int image_stats_launch(ImgF img, float* d_result3, cudaStream_t st) {
  image_stats_init<<<1, 1, 0, st>>>(d_result3);
  VWB_LAUNCH_CHECK();
  image_stats_kernel<<<296, 256, 0, st>>>(img, d_result3);
  VWB_LAUNCH_CHECK();
  image_stats_fini<<<1, 1, 0, st>>>(d_result3);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

int k1_fast_supported(int, int, int, int, int, float, float, bool) { return VWB200_ENOIMPL; }
size_t k1_fast_workspace_bytes(int, int, int, int, int, int) { return 16; }
int k1_fast_launch(int, ImgF, ImgF, int, int, int, int, int, int, vwb200_dispi*, ptrdiff_t, void*, size_t, cudaStream_t) {
  set_error("k1_fast: not built");
  return VWB200_ENOIMPL;
}

}  // namespace vwb200
