// k1_generic.cu -- general (any float input, fp64 accumulation, zone-table driven) fused
// cost-volume + arg-best kernel.  This is the kernel the pyramid level loop uses: thousands of
// small zones with narrow search windows per level (Stereo/CorrelationView.cc:607-648), all handled
// by ONE launch per level through a tile table.
//
// Semantics follow vw::stereo::best_of_search_convolution (Stereo/Correlation.cc:33-137):
//  * per-pixel cost in FLOAT (|a-b|, (a-b)^2, a*b; Stereo/CostFunctions.h:72-141), widened to double
//  * kx x ky window sums in double
//  * NCC: cost *= sqrt(lp * rp[d]) in double (CostFunctions.h:227-231), "better" is '>'
//  * disparities visited dy-major / dx-minor; strict comparison => first disparity wins ties
//  * pixel invalid iff every disparity produced the same cost (best == worst, :121-133)
//  * NaN costs (NCC zero-energy windows) follow the reference's order-dependent best/worst state
//    machine exactly: such pixels are flagged here and re-evaluated sequentially by k1_nan_fixup.
//
// Structure per CTA (one TW x TH tile of one zone), per disparity:
//   phase 1: thread = (padded column, row segment): vertical sliding sum of the per-pixel cost,
//            stored to shared memory V[y][x']            (O(1) per pixel, coalesced global reads)
//   phase 2: thread = (row, column chunk): horizontal sliding sum over V, compare with the running
//            best kept in shared memory                  (O(1) per pixel)
// Window sums are exact (hence order independent, hence identical to the reference's sliding sums)
// whenever the cost terms are multiples of 2^-32 below 2^21 (see DESIGN.md "exactness").
#include "common.cuh"
#include <vector>
#include <algorithm>

namespace vwb200 {

static constexpr int K1G_THREADS = 256;
static constexpr int K1G_TILE = 64;
static constexpr int IDX_MASK = 0x0fffffff;
static constexpr int FLAG_DIFF = 0x40000000;
static constexpr int FLAG_NAN = 0x20000000;

int k1_generic_tile_w(int kx) { (void)kx; return 32; }
int k1_generic_tile_h(int ky) { (void)ky; return 16; }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <int COST>
__device__ __forceinline__ double pix_cost(float a, float b) {
  if (COST == VWB200_SQUARED_DIFFERENCE) { float d = __fsub_rn(a, b); return (double)__fmul_rn(d, d); }
  if (COST == VWB200_CROSS_CORRELATION) { return (double)__fmul_rn(a, b); }
  return (double)fabsf(__fsub_rn(a, b));
}
template <int COST>
__device__ __forceinline__ bool better(double c, double q) {
  return COST == VWB200_CROSS_CORRELATION ? (c > q) : (c < q);
}

__device__ __forceinline__ float ld_clamped(const ImgF& im, int x, int y) {
  return __ldg(im.p + (ptrdiff_t)clampi(y, 0, im.h - 1) * im.pitch + clampi(x, 0, im.w - 1));
}
template <bool CLAMP>
__device__ __forceinline__ float ld_img(const ImgF& im, int x, int y) {
  if (CLAMP) return ld_clamped(im, x, y);
  return __ldg(im.p + (ptrdiff_t)y * im.pitch + x);
}

// Zone kernel.  One CTA per ZT_W x ZT_H tile of a zone; the 8 warps work on different disparities at the same
// time (d = warp, warp + 8, ...) with only warp-level synchronisation inside the disparity loop:
//   phase 1  lane = padded column: vertical sliding sum of the per-pixel cost down the tile's rows -> V[warp][y][x']
//   phase 2  lane = (row, half row): horizontal sliding sum over V, compare with the warp's private running best
// After the loop the warps' private bests are merged (cost, then raster index: first disparity wins ties).
static constexpr int ZT_W = 32, ZT_H = 16, ZWARPS = 8;

// STAGE: the tile's left patch and right search patch are copied once into shared memory (as floats, with the
// constant-edge clamping applied during the copy), so the disparity loop touches no global memory.  Used whenever
// the right patch fits (ZR_MAX floats); otherwise the loop reads global memory through L1.
static constexpr int ZR_MAX = 12288;          // 48 KB of right patch
static constexpr int ZCW = 16;                // max columns per phase-2 lane (ZT_W / 2 halves)

template <int COST, bool CLAMP, bool STAGE>
__global__ void __launch_bounds__(K1G_THREADS)
k1_generic_kernel(ImgF L, ImgF R, const Zone* __restrict__ zones, const Tile* __restrict__ tiles,
                  int kx, int ky, NccMaps ncc, vwb200_dispi* __restrict__ out, double* __restrict__ scratch_cost,
                  int* __restrict__ scratch_idx, const int* __restrict__ zone_gate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Tile t = tiles[blockIdx.x];
  if (zone_gate && !zone_gate[t.zone]) return;                  // fallback pass behind the integer zone kernel: flagged zones only
  const Zone z = zones[t.zone];
  const int tw = min(ZT_W, z.w - t.tx), th = min(ZT_H, z.h - t.ty);
  const int pw = tw + kx - 1, ph = th + ky - 1;
  const int vp = (ZT_W + kx - 1) | 1;                         // odd pitch: conflict-free when lanes walk rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* Vall = reinterpret_cast<double*>(smem_raw);          // [ZWARPS][ZT_H][vp]
  float* sL = reinterpret_cast<float*>(Vall + (size_t)ZWARPS * ZT_H * vp);          // [ph][pw]      (STAGE)
  float* sR = sL + (size_t)(ZT_H + ky - 1) * (ZT_W + kx - 1);                      // [ph+sy-1][pw+sx-1]
  // the merge buffers are only used after the disparity loop and alias everything above
  double* bestall = reinterpret_cast<double*>(smem_raw);        // [ZWARPS][ZT_H*ZT_W]
  int* bidxall = reinterpret_cast<int*>(bestall + (size_t)ZWARPS * ZT_H * ZT_W);
  double* V = Vall + (size_t)warp * ZT_H * vp;
  const int lx0 = z.lx + t.tx, ly0 = z.ly + t.ty;
  const int rx0 = z.rx + t.tx, ry0 = z.ry + t.ty;
  const int nd_all = z.sx * z.sy;
  const int d_begin = t.chunk * K1G_DCHUNK;
  const int nd = z.nchunks > 1 ? min(nd_all, d_begin + K1G_DCHUNK) : nd_all;   // this CTA covers [d_begin, nd)
  const int rpw = pw + z.sx - 1;
  const int dy_lo = d_begin / z.sx;                              // first search row this CTA touches
  if (STAGE) {
    const int dy_hi = (nd - 1) / z.sx;
    const int rph = ph + (dy_hi - dy_lo);
    for (int k = threadIdx.x; k < pw * ph; k += blockDim.x) sL[k] = ld_clamped(L, lx0 + k % pw, ly0 + k / pw);
    for (int k = threadIdx.x; k < rpw * rph; k += blockDim.x) sR[k] = ld_clamped(R, rx0 + k % rpw, ry0 + dy_lo + k / rpw);
    __syncthreads();
  }
  // phase-2 work split: lane -> (row, half); each lane keeps the running best of its <= ZCW pixels in registers
  const int halves = (2 * th <= 32) ? 2 : 1;
  const int cw = (tw + halves - 1) / halves;
  const int prow = lane / halves, phalf = lane % halves;
  const int xb = phalf * cw, xe = min(tw, xb + cw);
  const bool p2 = prow < th && xb < xe;
  double rbest[ZCW];
  int ridx[ZCW];
#pragma unroll
  for (int i = 0; i < ZCW; ++i) { rbest[i] = 0.0; ridx[i] = 0; }
  bool first = true;
  for (int d = d_begin + warp; d < nd; d += ZWARPS) {
    const int dy = d / z.sx, dx = d - dy * z.sx;
    // ---- phase 1: vertical sliding sums ----
    for (int xp = lane; xp < pw; xp += 32) {
      double v = 0.0;
      if (STAGE) {
        const float* lp = sL + xp;
        const float* rp = sR + (dy - dy_lo) * rpw + xp + dx;
        for (int j = 0; j < ky; ++j) v += pix_cost<COST>(lp[j * pw], rp[j * rpw]);
        V[xp] = v;
        for (int y = 1; y < th; ++y) {
          v += pix_cost<COST>(lp[(y + ky - 1) * pw], rp[(y + ky - 1) * rpw]);
          v -= pix_cost<COST>(lp[(y - 1) * pw], rp[(y - 1) * rpw]);
          V[y * vp + xp] = v;
        }
      } else {
        const int gl = lx0 + xp, gr = rx0 + xp + dx;
        for (int j = 0; j < ky; ++j) v += pix_cost<COST>(ld_img<CLAMP>(L, gl, ly0 + j), ld_img<CLAMP>(R, gr, ry0 + j + dy));
        V[xp] = v;
        for (int y = 1; y < th; ++y) {
          v += pix_cost<COST>(ld_img<CLAMP>(L, gl, ly0 + y + ky - 1), ld_img<CLAMP>(R, gr, ry0 + y + ky - 1 + dy));
          v -= pix_cost<COST>(ld_img<CLAMP>(L, gl, ly0 + y - 1), ld_img<CLAMP>(R, gr, ry0 + y - 1 + dy));
          V[y * vp + xp] = v;
        }
      }
    }
    __syncwarp();
    // ---- phase 2: horizontal sliding sums + running best in registers ----
    if (p2) {
      const double* vr = V + prow * vp + xb;
      double h = 0.0;
      for (int i = 0; i < kx; ++i) h += vr[i];
#pragma unroll
      for (int xi = 0; xi < ZCW; ++xi) {
        if (xi < xe - xb) {
          double cost = h;
          if (COST == VWB200_CROSS_CORRELATION) {
            const int x = xb + xi;
            const double lp = ncc.inv_l[(ptrdiff_t)(ly0 + prow - ncc.l_oy) * ncc.l_w + (lx0 + x - ncc.l_ox)];
            const double rp = ncc.inv_r[(ptrdiff_t)(ry0 + prow + dy - ncc.r_oy) * ncc.r_w + (rx0 + x + dx - ncc.r_ox)];
            cost = __dmul_rn(h, sqrt(__dmul_rn(lp, rp)));
          }
          if (first) {
            rbest[xi] = cost;
            ridx[xi] = d | ((cost != cost) ? FLAG_NAN : 0);
          } else {
            int bi = ridx[xi];
            if (cost != rbest[xi]) bi |= FLAG_DIFF;
            if (cost != cost) bi |= FLAG_NAN;
            if (better<COST>(cost, rbest[xi])) { rbest[xi] = cost; bi = (bi & ~IDX_MASK) | d; }
            ridx[xi] = bi;
          }
          if (xi + 1 < xe - xb) h += vr[xi + kx] - vr[xi];
        }
      }
    }
    first = false;
    __syncwarp();
  }
  // ---- publish the warp's private bests, merge across warps; epilogue: 12-byte pixel writes ----
  const int nw = (nd - d_begin) < ZWARPS ? (nd - d_begin) : ZWARPS;
  __syncthreads();            // every warp is done with V / the staged patches: reuse the memory for the merge
  if (p2 && warp < nw) {
#pragma unroll
    for (int xi = 0; xi < ZCW; ++xi)
      if (xi < xe - xb) {
        bestall[(size_t)warp * ZT_H * ZT_W + prow * ZT_W + xb + xi] = rbest[xi];
        bidxall[(size_t)warp * ZT_H * ZT_W + prow * ZT_W + xb + xi] = ridx[xi];
      }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < tw * th; k += blockDim.x) {
    const int y = k / tw, x = k - y * tw;
    const int kk = y * ZT_W + x;
    double b = bestall[kk];
    int bi = bidxall[kk];
    int flags = bi & (FLAG_DIFF | FLAG_NAN);
    int d = bi & IDX_MASK;
    for (int w = 1; w < nw; ++w) {
      const double c = bestall[(size_t)w * ZT_H * ZT_W + kk];
      const int ci = bidxall[(size_t)w * ZT_H * ZT_W + kk];
      flags |= ci & (FLAG_DIFF | FLAG_NAN);
      if (c != b) flags |= FLAG_DIFF;
      const int cd = ci & IDX_MASK;
      if (better<COST>(c, b) || (c == b && cd < d)) { b = c; d = cd; }
    }
    if (z.nchunks > 1) {        // partial result of this disparity chunk; k1_generic_merge_kernel finishes the pixel
      const long long s = z.sbase + ((long long)t.chunk * z.h + (t.ty + y)) * z.w + (t.tx + x);
      scratch_cost[s] = b;
      scratch_idx[s] = d | flags;
      continue;
    }
    vwb200_dispi o;
    o.dx = d % z.sx + z.addx;
    o.dy = d / z.sx + z.addy;
    o.valid = (flags & FLAG_NAN) ? 2 : ((flags & FLAG_DIFF) ? 1 : 0);
    out[z.obase + (ptrdiff_t)(t.ty + y) * z.opitch + (t.tx + x)] = o;
  }
}

// merge of disparity chunks (ascending chunk order = ascending raster order of the disparities)
template <int COST>
__global__ void k1_generic_merge_kernel(const Zone* __restrict__ zones, const int* __restrict__ split, const double* __restrict__ sc,
                                        const int* __restrict__ si, vwb200_dispi* __restrict__ out) {
  const Zone z = zones[split[blockIdx.y]];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < z.w * z.h; k += gridDim.x * blockDim.x) {
    double b = sc[z.sbase + k];
    int bi = si[z.sbase + k];
    int flags = bi & (FLAG_DIFF | FLAG_NAN), d = bi & IDX_MASK;
    for (int c = 1; c < z.nchunks; ++c) {
      const double cc = sc[z.sbase + (long long)c * z.w * z.h + k];
      const int ci = si[z.sbase + (long long)c * z.w * z.h + k];
      flags |= ci & (FLAG_DIFF | FLAG_NAN);
      if (cc != b) flags |= FLAG_DIFF;
      if (better<COST>(cc, b)) { b = cc; d = ci & IDX_MASK; }     // equal cost: the earlier chunk (smaller d) stays
    }
    vwb200_dispi o;
    o.dx = d % z.sx + z.addx;
    o.dy = d / z.sx + z.addy;
    o.valid = (flags & FLAG_NAN) ? 2 : ((flags & FLAG_DIFF) ? 1 : 0);
    out[z.obase + (ptrdiff_t)(k / z.w) * z.opitch + (k % z.w)] = o;
  }
}
int k1_generic_merge_launch(int cost, const Zone* d_zones, const int* d_split, int nsplit, const double* scratch_cost,
                            const int* scratch_idx, vwb200_dispi* out, cudaStream_t st) {
  if (nsplit <= 0) return VWB200_OK;
  dim3 grid(16, nsplit);
  if (cost == VWB200_CROSS_CORRELATION) k1_generic_merge_kernel<VWB200_CROSS_CORRELATION><<<grid, 256, 0, st>>>(d_zones, d_split, scratch_cost, scratch_idx, out);
  else k1_generic_merge_kernel<VWB200_ABSOLUTE_DIFFERENCE><<<grid, 256, 0, st>>>(d_zones, d_split, scratch_cost, scratch_idx, out);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

static size_t k1g_smem_bytes(int kx, int ky, int stage_r_floats) {
  const size_t vp = (size_t)((ZT_W + kx - 1) | 1);
  size_t b = (size_t)ZWARPS * ZT_H * vp * sizeof(double);
  if (stage_r_floats > 0) b += ((size_t)(ZT_H + ky - 1) * (ZT_W + kx - 1) + (size_t)stage_r_floats) * sizeof(float);
  const size_t merge = (size_t)ZWARPS * ZT_H * ZT_W * (sizeof(double) + sizeof(int));
  return b > merge ? b : merge;
}
// floats of right search patch a CTA of this zone stages (a CTA of a split zone only touches the search rows of its
// K1G_DCHUNK disparities); > ZR_MAX: not staged
long long k1_generic_stage_floats(int kx, int ky, int sx, int sy, int nchunks) {
  const int span = nchunks > 1 ? std::min(sy, (K1G_DCHUNK + sx - 1) / sx + 1) : sy;
  return (long long)(ZT_W + kx - 1 + sx - 1) * (ZT_H + ky - 1 + span - 1);
}
bool k1_generic_can_stage(int kx, int ky, int sx, int sy, int nchunks) { return k1_generic_stage_floats(kx, ky, sx, sy, nchunks) <= ZR_MAX; }

int k1_generic_launch(int cost, ImgF left, ImgF right, const Zone* d_zones, const Tile* d_tiles, int ntiles,
                      int kx, int ky, NccMaps ncc, vwb200_dispi* out, double* scratch_cost, int* scratch_idx, bool clamp_reads,
                      int stage_r_floats, cudaStream_t st, const KEvents* ev, const int* zone_gate) {
  if (ntiles <= 0) return VWB200_OK;
  if (kx > 129 || ky > 129) { set_error("kernel size %dx%d exceeds the supported 129", kx, ky); return VWB200_ENOIMPL; }
  const bool stage = stage_r_floats > 0;
  const size_t smem = k1g_smem_bytes(kx, ky, stage_r_floats);
  void (*kern)(ImgF, ImgF, const Zone*, const Tile*, int, int, NccMaps, vwb200_dispi*, double*, int*, const int*);
#define KSEL(C) (stage ? k1_generic_kernel<C, true, true> : (clamp_reads ? k1_generic_kernel<C, true, false> : k1_generic_kernel<C, false, false>))
  switch (cost) {
    case VWB200_SQUARED_DIFFERENCE: kern = KSEL(VWB200_SQUARED_DIFFERENCE); break;
    case VWB200_CROSS_CORRELATION:  kern = KSEL(VWB200_CROSS_CORRELATION); break;
    default:                        kern = KSEL(VWB200_ABSOLUTE_DIFFERENCE); break;
  }
#undef KSEL
  // the attribute is per-function state shared by all host threads: always the device maximum, never the per-launch size
  if (smem > 227 * 1024) { set_error("zone kernel needs %zu bytes of shared memory", smem); return VWB200_ENOIMPL; }
  VWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (ev && ev->e0) cudaEventRecord(ev->e0, st);
  kern<<<ntiles, K1G_THREADS, smem, st>>>(left, right, d_zones, d_tiles, kx, ky, ncc, out, scratch_cost, scratch_idx, zone_gate);
  VWB_LAUNCH_CHECK();
  if (ev && ev->e1) cudaEventRecord(ev->e1, st);
  return VWB200_OK;
}


// ------------------------------------------------------------------------------------------------
// 1 / boxsum(v*v): NCCCost's left_precision / right_precision (Stereo/CostFunctions.h:214-219).
// square() is the float product v*v (Math/Functors.h:316-321), summed in double.
// ------------------------------------------------------------------------------------------------
// MODE 0: out = 1 / boxsum(v*v) (double).  MODE 1: out_i = boxsum(v) as int32 (integer-valued imagery: exact).
// MODE 0: v*v (float product, as the reference) -> 1/sum;  MODE 1: v -> int sum;  MODE 2: (v - cen)^2 -> int sum
template <int MODE>
__device__ __forceinline__ double box_term(float a, float cen) {
  if (MODE == 0) return (double)__fmul_rn(a, a);
  if (MODE == 1) return (double)a;
  const float d = __fsub_rn(a, cen);
  return (double)__fmul_rn(d, d);
}
template <int MODE>
__global__ void __launch_bounds__(K1G_THREADS)
box_sq_inv_kernel(ImgF img, int kx, int ky, int ox0, int oy0, int ow, int oh, double* __restrict__ out, int* __restrict__ out_i, float cen) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* V = reinterpret_cast<double*>(smem_raw);
  const int tx0 = blockIdx.x * K1G_TILE, ty0 = blockIdx.y * K1G_TILE;
  const int tw = min(K1G_TILE, ow - tx0), th = min(K1G_TILE, oh - ty0);
  const int pw = tw + kx - 1, vp = pw | 1;
  const int tid = threadIdx.x, nth = blockDim.x;
  int S = nth / pw; if (S < 1) S = 1; { int m = th / 8; if (m < 1) m = 1; if (S > m) S = m; }
  const int rps = (th + S - 1) / S;
  for (int item = tid; item < pw * S; item += nth) {
    const int xp = item % pw, s = item / pw;
    const int yb = s * rps, ye = min(th, yb + rps);
    if (yb < ye) {
      const int gx = ox0 + tx0 + xp, gy = oy0 + ty0;
      double v = 0.0;
      for (int j = 0; j < ky; ++j) { float a = ld_clamped(img, gx, gy + yb + j); v += box_term<MODE>(a, cen); }
      V[yb * vp + xp] = v;
      for (int y = yb + 1; y < ye; ++y) {
        float a = ld_clamped(img, gx, gy + y + ky - 1); v += box_term<MODE>(a, cen);
        float b = ld_clamped(img, gx, gy + y - 1); v -= box_term<MODE>(b, cen);
        V[y * vp + xp] = v;
      }
    }
  }
  __syncthreads();
  int C = nth / th; if (C < 1) C = 1; { int m = tw / 8; if (m < 1) m = 1; if (C > m) C = m; }
  const int cw = (tw + C - 1) / C;
  for (int item = tid; item < th * C; item += nth) {
    const int y = item % th, c = item / th;
    const int xb = c * cw, xe = min(tw, xb + cw);
    if (xb < xe) {
      const double* vr = V + y * vp;
      double h = 0.0;
      for (int i = 0; i < kx; ++i) h += vr[xb + i];
      for (int x = xb; x < xe; ++x) {
        if (MODE) out_i[(ptrdiff_t)(ty0 + y) * ow + (tx0 + x)] = (int)h;
        else out[(ptrdiff_t)(ty0 + y) * ow + (tx0 + x)] = __ddiv_rn(1.0, h);
        if (x + 1 < xe) h += vr[x + kx] - vr[x];
      }
    }
  }
}

int box_sq_inv_launch(ImgF img, int kx, int ky, int ox0, int oy0, int ow, int oh, double* out, cudaStream_t st) {
  if (ow <= 0 || oh <= 0) return VWB200_OK;
  size_t smem = (size_t)K1G_TILE * ((K1G_TILE + kx - 1) | 1) * sizeof(double);
  VWB_CUDA(cudaFuncSetAttribute(box_sq_inv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per-function state shared by all host threads: device maximum
  dim3 grid((ow + K1G_TILE - 1) / K1G_TILE, (oh + K1G_TILE - 1) / K1G_TILE);
  box_sq_inv_kernel<0><<<grid, K1G_THREADS, smem, st>>>(img, kx, ky, ox0, oy0, ow, oh, out, nullptr, 0.0f);
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}
int box_sum_i32_launch(ImgF img, int kx, int ky, int ox0, int oy0, int ow, int oh, int* out, cudaStream_t st, int centred, float c) {
  if (ow <= 0 || oh <= 0) return VWB200_OK;
  size_t smem = (size_t)K1G_TILE * ((K1G_TILE + kx - 1) | 1) * sizeof(double);
  dim3 grid((ow + K1G_TILE - 1) / K1G_TILE, (oh + K1G_TILE - 1) / K1G_TILE);
  if (centred) {
    VWB_CUDA(cudaFuncSetAttribute(box_sq_inv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per-function state shared by all host threads: device maximum
    box_sq_inv_kernel<2><<<grid, K1G_THREADS, smem, st>>>(img, kx, ky, ox0, oy0, ow, oh, nullptr, out, c);
  } else {
    VWB_CUDA(cudaFuncSetAttribute(box_sq_inv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per-function state shared by all host threads: device maximum
    box_sq_inv_kernel<1><<<grid, K1G_THREADS, smem, st>>>(img, kx, ky, ox0, oy0, ow, oh, nullptr, out, 0.0f);
  }
  VWB_LAUNCH_CHECK();
  return VWB200_OK;
}

// ------------------------------------------------------------------------------------------------
// NaN fix-up: pixels the main kernel marked valid==2 saw at least one NaN cost.  Replays the
// reference's sequential best/worst state machine (Stereo/Correlation.cc:97-133) for them.
// ------------------------------------------------------------------------------------------------
template <int COST>
__global__ void k1_nan_fixup_kernel(ImgF L, ImgF R, const Zone* __restrict__ zones, int kx, int ky, NccMaps ncc,
                                    vwb200_dispi* __restrict__ out) {
  const Zone z = zones[blockIdx.y];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < z.w * z.h; k += gridDim.x * blockDim.x) {
    const int x = k % z.w, y = k / z.w;
    vwb200_dispi* o = out + z.obase + (ptrdiff_t)y * z.opitch + x;
    if (o->valid != 2) continue;
    if (COST == VWB200_CROSS_CORRELATION) {
      // zero-energy LEFT window: every cost is 0 * inf = NaN, so best stays NaN at (0,0) and best == worst is false (:110-133)
      const double lp = ncc.inv_l[(ptrdiff_t)(z.ly + y - ncc.l_oy) * ncc.l_w + (z.lx + x - ncc.l_ox)];
      if (isinf(lp)) { o->dx = z.addx; o->dy = z.addy; o->valid = 1; continue; }
    }
    double best = 0, worst = 0;
    int bd = 0;
    for (int dy = 0; dy < z.sy; ++dy)
      for (int dx = 0; dx < z.sx; ++dx) {
        double h = 0.0;
        for (int j = 0; j < ky; ++j)
          for (int i = 0; i < kx; ++i)
            h += pix_cost<COST>(ld_clamped(L, z.lx + x + i, z.ly + y + j), ld_clamped(R, z.rx + x + i + dx, z.ry + y + j + dy));
        double cost = h;
        if (COST == VWB200_CROSS_CORRELATION) {
          const double lp = ncc.inv_l[(ptrdiff_t)(z.ly + y - ncc.l_oy) * ncc.l_w + (z.lx + x - ncc.l_ox)];
          const double rp = ncc.inv_r[(ptrdiff_t)(z.ry + y + dy - ncc.r_oy) * ncc.r_w + (z.rx + x + dx - ncc.r_ox)];
          cost = __dmul_rn(h, sqrt(__dmul_rn(lp, rp)));
        }
        if (dx == 0 && dy == 0) { best = worst = cost; }
        else if (better<COST>(cost, best)) { best = cost; bd = dy * z.sx + dx; }
        else if (!better<COST>(cost, worst)) { worst = cost; }
      }
    o->dx = bd % z.sx + z.addx;
    o->dy = bd / z.sx + z.addy;
    o->valid = (best == worst) ? 0 : 1;
  }
}

int k1_nan_fixup_launch(int cost, ImgF left, ImgF right, const Zone* d_zones, int nzones, int kx, int ky,
                        NccMaps ncc, vwb200_dispi* out, cudaStream_t st, int gridx) {
  for (int z0 = 0; z0 < nzones; z0 += 32768) {
    const int n = nzones - z0 < 32768 ? nzones - z0 : 32768;
    dim3 grid(gridx, n);
    switch (cost) {
      case VWB200_SQUARED_DIFFERENCE:
        k1_nan_fixup_kernel<VWB200_SQUARED_DIFFERENCE><<<grid, 128, 0, st>>>(left, right, d_zones + z0, kx, ky, ncc, out); break;
      case VWB200_CROSS_CORRELATION:
        k1_nan_fixup_kernel<VWB200_CROSS_CORRELATION><<<grid, 128, 0, st>>>(left, right, d_zones + z0, kx, ky, ncc, out); break;
      default:
        k1_nan_fixup_kernel<VWB200_ABSOLUTE_DIFFERENCE><<<grid, 128, 0, st>>>(left, right, d_zones + z0, kx, ky, ncc, out); break;
    }
    VWB_LAUNCH_CHECK();
  }
  return VWB200_OK;
}

}  // namespace vwb200
