// k1_fast_ncc.cu -- NCC (CROSS_CORRELATION) on the exact-integer fast path.
//
// Reference (Stereo/CostFunctions.h:204-236, Stereo/Correlation.cc:79-117): per disparity
//     cost(d) = double(box(l*r)) * sqrt( lp * rp[d] ),  lp = 1/box(l^2), rp = 1/box(r^2)   (all double, IEEE ops)
// and the arg-MAX with strict '>' in dy-major / dx-minor order.  The double multiply + sqrt per (pixel, disparity)
// is what makes the general kernel slow.  Here, for non-negative integer imagery (<= 12 bit):
//   * the hot loop carries the EXACT integer numerator.  Values are centred by c so that everything fits int32:
//         S' = box((l-c)(r-c))   (vertical sliding IMAD sums in registers, shuffle-based window sums)
//         box(l*r) = S' + c*(Sl + Sr[d]) - N*c^2          (Sl, Sr = box sums of l and r, precomputed)
//   * a rigorous fp32 UPPER bound of cost(d)/sqrt(lp) is formed with 5 cheap ops,
//         su = (float(S') + A[d] + B) * Q[d],   A = float(c*Sr), B = float(c*Sl - N c^2) + E,  Q = float(sqrt(rp)) rounded up
//     (E bounds every rounding in that expression), and compared with a per-pixel threshold thr that is kept just
//     below the best exact cost seen so far.  Only candidates (su >= thr: record breakers and near ties, ~10-20 per
//     pixel out of 16384) take the rare path, which evaluates the reference's double expression exactly.
//   * exactness: a disparity whose exact cost beats (or ties) the running best can never be rejected, so the set of
//     exactly-evaluated candidates always contains the reference's arg-max; ties resolve to the first in raster order.
//   * NaN costs (zero-energy windows) and the all-equal rule are settled by the same fix-up kernels as elsewhere.
// Layout / staging / decomposition are those of k1_fast.cu (TMA-staged u16 tiles, 8 columns x 8 dx per lane,
// 4 dx subsets x 2 row halves per CTA) plus two more shared-memory rings for the Q and A rows.
#include "k1_fast_common.cuh"
#include <cmath>
#include <cstdlib>
#include <algorithm>

namespace vwb200 {

static constexpr int NQ_SLOTS = F_TH + 1;

struct NccGeom {
  FastGeom g;
  int c;                 // centring constant
  int qrows;             // rows of the packed Q / A maps per strip
  float e_abs;           // E (already folded into B); kept for reference
};

static size_t ncc_smem_bytes(const FastGeom& g) {
  return (size_t)F_TH * F_COLS * 4                                  // thr
         + (size_t)g.ltile_rows * F_COLS * 2 + (size_t)g.ring_slots * g.rw * 2     // left tile, right ring (int16)
         + 2 * (size_t)NQ_SLOTS * g.rw * 4 + 64;                                    // Q ring, A ring (float)
}

static bool ncc_params(int kx, int ky, float vmin, float vmax, int* c_out, double* maxc_out) {
  if (!(vmin >= 0.0f) || !(vmax <= 4095.0f)) return false;          // products < 2^24 stay exact in the reference's float
  const int c = (int)((vmin + vmax) * 0.5f + 0.5f);
  const double maxc = std::max((double)vmax - c, (double)c - vmin);
  if (maxc > 32767.0 || maxc * maxc * kx * ky >= 2147483647.0) return false;
  *c_out = c; *maxc_out = maxc;
  return true;
}

int k1_fast_ncc_supported(int kx, int ky, int sx, int sy, float vmin, float vmax, bool integer_valued) {
  if (!integer_valued) return VWB200_ENOIMPL;
  int c; double maxc;
  if (!ncc_params(kx, ky, vmin, vmax, &c, &maxc)) return VWB200_ENOIMPL;
  if (kx < 3 || kx > 31 || ky < 1 || ky > 41) return VWB200_ENOIMPL;
  if (sx < F_B || sx > 512 || sy < 1 || (long long)sx * sy < 64) return VWB200_ENOIMPL;
  FastGeom g = make_geom(256, 32, sx, sy, kx, ky);
  if (ncc_smem_bytes(g) > 227 * 1024) return VWB200_ENOIMPL;
  return VWB200_OK;
}

// ---- workspace carving ---------------------------------------------------------------------------------------
struct NccWs {
  int16_t *L16, *R16; float *Qp, *Ap, *BEp;
  double *lp, *rp; int *Sl, *Sr;
  double* bs; int* bi;               // per-CTA private bests [148][4][32][256]
  double* pc; int* pi;               // per-dy-chunk partials (J > 1)
  unsigned char* nanflag;            // W x H
  Zone* zone;
  size_t total;
};
static inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
static NccWs carve(const FastGeom& g, void* base) {
  NccWs w;
  unsigned char* p = static_cast<unsigned char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { unsigned char* q = p ? p + off : nullptr; off += al(bytes); return q; };
  const size_t ow = (size_t)g.W + g.sx - 1, oh = (size_t)g.H + g.sy - 1;
  const size_t qrows = (size_t)g.NB * F_TH + g.sy;
  w.L16 = (int16_t*)take((size_t)g.NS * g.lrows * F_COLS * 2);
  w.R16 = (int16_t*)take((size_t)g.NS * g.rrows * g.rw * 2);
  w.Qp = (float*)take((size_t)g.NS * qrows * g.rw * 4);
  w.Ap = (float*)take((size_t)g.NS * qrows * g.rw * 4);
  w.BEp = (float*)take((size_t)g.NS * g.NB * F_TH * F_COLS * 4);
  w.lp = (double*)take((size_t)g.W * g.H * 8);
  w.rp = (double*)take(ow * oh * 8);
  w.Sl = (int*)take((size_t)g.W * g.H * 4);
  w.Sr = (int*)take(ow * oh * 4);
  w.bs = (double*)take((size_t)148 * F_SUBSETS * F_TH * F_COLS * 8);
  w.bi = (int*)take((size_t)148 * F_SUBSETS * F_TH * F_COLS * 4);
  w.pc = (double*)take(g.J > 1 ? (size_t)g.J * g.W * g.H * 8 : 8);
  w.pi = (int*)take(g.J > 1 ? (size_t)g.J * g.W * g.H * 4 : 8);
  w.nanflag = take((size_t)g.W * g.H);
  w.zone = (Zone*)take(sizeof(Zone));
  w.total = off;
  return w;
}
size_t k1_fast_ncc_workspace_bytes(int W, int H, int sx, int sy, int kx, int ky) {
  FastGeom g = make_geom(W, H, sx, sy, kx, ky);
  return carve(g, nullptr).total + 256;
}

__global__ void ncc_set_zone_kernel(Zone* dst, Zone z) { *dst = z; }

// ---- pack kernels ---------------------------------------------------------------------------------------------
__global__ void ncc_pack_img_kernel(ImgF img, int c, FastGeom g, int right, int16_t* __restrict__ out) {
  const int row = blockIdx.x, strip = blockIdx.y;
  const int s0 = strip * g.out_cols;
  const int rowlen = right ? g.rw : F_COLS, nrows = right ? g.rrows : g.lrows;
  const int lw = g.W + g.kx - 1 + (right ? g.sx - 1 : 0), lh = g.H + g.ky - 1 + (right ? g.sy - 1 : 0);
  int16_t* o = out + ((size_t)strip * nrows + row) * rowlen;
  for (int col = threadIdx.x; col < rowlen; col += blockDim.x) {
    const int gx = s0 + col;
    int16_t v = 0;
    if (row < lh && gx < lw) v = (int16_t)((int)img.p[(ptrdiff_t)row * img.pitch + gx] - c);
    o[col] = v;
  }
}
// Q = float(sqrt(rp)) rounded up, A = float(c * Sr), per strip, rows = window-origin rows of the right raster
__global__ void ncc_pack_qa_kernel(const double* __restrict__ rp, const int* __restrict__ Sr, int c, FastGeom g, int qrows, int cvt,
                                   float* __restrict__ Qp, float* __restrict__ Ap) {
  const int row = blockIdx.x, strip = blockIdx.y;
  const int s0 = strip * g.out_cols;
  const int ow = g.W + g.sx - 1, oh = g.H + g.sy - 1;
  float* q = Qp + ((size_t)strip * qrows + row) * g.rw;
  float* a = Ap + ((size_t)strip * qrows + row) * g.rw;
  for (int col = threadIdx.x; col < g.rw; col += blockDim.x) {
    const int gx = s0 + col;
    float qv = 0.0f, av = 0.0f;
    if (row < oh && gx < ow) {
      const double r = rp[(size_t)row * ow + gx];
      qv = __double2float_ru(sqrt(r));                       // +inf when the window has no energy: always a candidate
      qv = __fmul_ru(qv, 1.0000002f);                         // sqrt() itself is rounded: stay an upper bound
      const double a = (double)c * (double)Sr[(size_t)row * ow + gx];
      av = __double2float_ru(a);
      if (cvt == 2) { qv *= 2048.0f; av = __double2float_ru(a * (1.0 / 2048.0) - 12582912.0); }
    }
    q[col] = qv; a[col] = av;
  }
}
// B = float(c*Sl - N c^2 + E) rounded up, per strip/band row
__global__ void ncc_pack_be_kernel(const int* __restrict__ Sl, int c, int N, float e_abs, FastGeom g, int cvt, float* __restrict__ BEp) {
  const int row = blockIdx.x, strip = blockIdx.y;
  const int s0 = strip * g.out_cols;
  float* b = BEp + ((size_t)strip * g.NB * F_TH + row) * F_COLS;
  for (int col = threadIdx.x; col < F_COLS; col += blockDim.x) {
    const int gx = s0 + col;
    float v = 0.0f;
    if (row < g.H && gx < g.W) {
      const double b = (double)c * (double)Sl[(size_t)row * g.W + gx] - (double)N * c * c + (double)e_abs;
      v = __double2float_ru(cvt == 2 ? b * (1.0 / 2048.0) : b);
    }
    b[col] = v;
  }
}

// ---- device helpers -----------------------------------------------------------------------------------------
// prmt with sign replication (selector nibble 8|k = msb of byte k in all 8 bits); __byte_perm() only honours 3 selector bits
__device__ __forceinline__ int prmt_s(uint32_t a, uint32_t sel) {
  int d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}
__device__ __forceinline__ void unpack8s(const uint4 v, int (&o)[8]) {      // int16 pairs -> sign-extended int32 (1 PRMT each)
  o[0] = prmt_s(v.x, 0x9910u); o[1] = prmt_s(v.x, 0xBB32u);
  o[2] = prmt_s(v.y, 0x9910u); o[3] = prmt_s(v.y, 0xBB32u);
  o[4] = prmt_s(v.z, 0x9910u); o[5] = prmt_s(v.z, 0xBB32u);
  o[6] = prmt_s(v.w, 0x9910u); o[7] = prmt_s(v.w, 0xBB32u);
}
__device__ __forceinline__ void load_row_s(const int16_t* lrow, const int16_t* rrow, int (&Lv)[8], int (&Rv)[16]) {
  unpack8s(*reinterpret_cast<const uint4*>(lrow), Lv);
  int t[8];
  unpack8s(*reinterpret_cast<const uint4*>(rrow), t);
#pragma unroll
  for (int i = 0; i < 8; ++i) Rv[i] = t[i];
  unpack8s(*reinterpret_cast<const uint4*>(rrow + 8), t);
#pragma unroll
  for (int i = 0; i < 8; ++i) Rv[8 + i] = t[i];
}

// int32 window sum -> float operand of the bound.  CVT 0: rounded-up conversion; CVT 1: round-to-nearest conversion (the
// half-ulp, <= 128, is part of E); CVT 2: (v >> 11) + 1 placed in the mantissa of 1.5 * 2^23 by an integer add -- the maps are
// pre-scaled by 2^-11 and A carries the -1.5 * 2^23.
template <int CVT>
__device__ __forceinline__ float ncc_cvt_up(int v) {
  if (CVT == 0) return __int2float_ru(v);
  if (CVT == 1) return __int2float_rn(v);
  return __int_as_float((v >> 11) + 0x4B400001);
}

struct NccCtx {           // what the rare exact path needs
  const double* lp; const double* rp; const int* Sl; const int* Sr;
  unsigned char* nanflag;
  int W, H, ow, c; long long Ncc;
};

// Exact evaluation of one candidate (reference arithmetic), update of the warp-subset's private best and of the
// shared threshold.  Lane-divergent, rare.
__device__ __noinline__ void ncc_candidate(const NccCtx& cx, int sprime, int gx, int gy, int dx, int dy, int didx,
                                           double* __restrict__ bs, int* __restrict__ bi, float* __restrict__ thr) {
  if (gx >= cx.W || gy >= cx.H) return;          // (columns past the strip's own outputs have thr = +inf and Q finite or inf: harmless)
  const size_t kl = (size_t)gy * cx.W + gx, kr = (size_t)(gy + dy) * cx.ow + (gx + dx);
  const long long slr = (long long)sprime + (long long)cx.c * ((long long)cx.Sl[kl] + (long long)cx.Sr[kr]) - cx.Ncc;
  const double lpv = cx.lp[kl];
  const double s = __dmul_rn((double)slr, sqrt(__dmul_rn(lpv, cx.rp[kr])));          // Correlation.cc:82, CostFunctions.h:227-231
  if (s != s) {                                       // zero-energy window: the pixel is replayed by the NaN fix-up, skip the rest
    cx.nanflag[kl] = 1;
    atomicMax(reinterpret_cast<int*>(thr), 0x7f800000);
    return;
  }
  if (s > *bs || (s == *bs && didx < *bi)) {          // strict '>' in raster order; within a subset passes are in raster order already
    *bs = s; *bi = didx;
    // threshold in units of cost / sqrt(lp), a little below the new best
    const double t = s / sqrt(lpv);
    float tf = __double2float_rd(t * (1.0 - 1.0e-9));
    if (!(tf > 0.0f)) tf = 0.0f;
    atomicMax(reinterpret_cast<int*>(thr), __float_as_int(tf));
  }
}

template <int KX, bool FULL, int CVT>
__device__ __forceinline__ void fast_pass_ncc(const int16_t* __restrict__ ltile, const int16_t* __restrict__ rring,
                                              const float* __restrict__ qring, const float* __restrict__ aring,
                                              float* __restrict__ thr, const float* __restrict__ be_band,
                                              double* __restrict__ bs, int* __restrict__ bi, const NccCtx& cx,
                                              int lane, int g, int ky, int ring_slots, int rw, int ring_base, int qbase,
                                              int row0, int nb, int sx, int dy_abs, int s0, int y0) {
  int V[8][F_B];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < F_B; ++b) V[a][b] = 0;
  const int16_t* lp = ltile + row0 * F_COLS + 8 * lane;
  const int16_t* rp = rring + 8 * (lane + g);
  int slot_new = ring_base;
  for (int t = 0; t < ky; ++t) {
    int Lv[8], Rv[16];
    load_row_s(lp + t * F_COLS, rp + slot_new * rw, Lv, Rv);
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < F_B; ++b) V[a][b] += Lv[a] * Rv[a + b];
    if (++slot_new == ring_slots) slot_new = 0;
  }
  int slot_old = ring_base;
  int qslot = qbase;                       // Q/A ring slot of right window-origin row (dy + row0 + y)
  for (int y = 0; y < F_RH; ++y) {
    if (y > 0) {
      {
        int Lv[8], Rv[16];
        load_row_s(lp + (y + ky - 1) * F_COLS, rp + slot_new * rw, Lv, Rv);
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < F_B; ++b) V[a][b] += Lv[a] * Rv[a + b];
      }
      {
        int Lo[8], Ro[16];
        load_row_s(lp + (y - 1) * F_COLS, rp + slot_old * rw, Lo, Ro);
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < F_B; ++b) V[a][b] -= Lo[a] * Ro[a + b];
      }
      if (++slot_new == ring_slots) slot_new = 0;
      if (++slot_old == ring_slots) slot_old = 0;
    }
    // per-row operands of the bound: Q and A at the 15 right positions, B and thr at the 8 pixels
    float Qv[16], Av[16], Bv[8], Tv[8];
    {
      const float4* qp = reinterpret_cast<const float4*>(qring + qslot * rw + 8 * (lane + g));
      const float4* ap = reinterpret_cast<const float4*>(aring + qslot * rw + 8 * (lane + g));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q = qp[i], a = ap[i];
        Qv[4 * i] = q.x; Qv[4 * i + 1] = q.y; Qv[4 * i + 2] = q.z; Qv[4 * i + 3] = q.w;
        Av[4 * i] = a.x; Av[4 * i + 1] = a.y; Av[4 * i + 2] = a.z; Av[4 * i + 3] = a.w;
      }
      const float4* bp = reinterpret_cast<const float4*>(be_band + (row0 + y) * F_COLS + 8 * lane);
      const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
      Bv[0] = b0.x; Bv[1] = b0.y; Bv[2] = b0.z; Bv[3] = b0.w; Bv[4] = b1.x; Bv[5] = b1.y; Bv[6] = b1.z; Bv[7] = b1.w;
      const float* tp = thr + (row0 + y) * F_COLS + lane;
#pragma unroll
      for (int r = 0; r < 8; ++r) Tv[r] = tp[r * 32];
    }
#pragma unroll
    for (int b = 0; b < F_B; ++b) {
      if (!FULL && b >= nb) break;
      int p[8], o[8];
      p[0] = V[0][b];
#pragma unroll
      for (int a = 1; a < 8; ++a) p[a] = p[a - 1] + V[a][b];
      window_sums<KX>(p, o);
      float su[8];
      bool any = false;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        su[r] = __fmul_ru(__fadd_ru(__fadd_ru(ncc_cvt_up<CVT>(o[r]), Av[r + b]), Bv[r]), Qv[r + b]);
        any |= !(su[r] < Tv[r]);                               // candidate (also catches NaN / inf)
      }
      if (any) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float* tp = thr + (row0 + y) * F_COLS + r * 32 + lane;
          if (!(su[r] < *reinterpret_cast<volatile float*>(tp))) {        // re-read: keeps the per-pixel tests out of the hot path
            const size_t k = (size_t)(row0 + y) * F_COLS + r * 32 + lane;
            ncc_candidate(cx, o[r], s0 + 8 * lane + r, y0 + row0 + y, F_B * g + b, dy_abs, dy_abs * sx + F_B * g + b, bs + k, bi + k, tp);
            Tv[r] = *tp;
          }
        }
      }
    }
    if (++qslot == NQ_SLOTS) qslot = 0;
  }
}

template <int KX, int CVT>
__global__ void __launch_bounds__(F_THREADS, 1)
k1_fast_ncc_kernel(const int16_t* __restrict__ L16, const int16_t* __restrict__ R16, const float* __restrict__ Qp,
                   const float* __restrict__ Ap, const float* __restrict__ BEp, FastGeom G, int qrows, NccCtx cx,
                   double* __restrict__ bs_all, int* __restrict__ bi_all, vwb200_dispi* __restrict__ out, ptrdiff_t opitch,
                   double* __restrict__ part_cost, int* __restrict__ part_idx) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* thr = reinterpret_cast<float*>(smem);                                              // [32][8][32]
  int16_t* ltile = reinterpret_cast<int16_t*>(smem + (size_t)F_TH * F_COLS * 4);
  int16_t* rring = ltile + (size_t)G.ltile_rows * F_COLS;
  float* qring = reinterpret_cast<float*>(rring + (size_t)G.ring_slots * G.rw);
  float* aring = qring + (size_t)NQ_SLOTS * G.rw;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aring + (size_t)NQ_SLOTS * G.rw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int sub = w & (F_SUBSETS - 1), half = w / F_SUBSETS, row0 = half * F_RH;
  const int ngroups = (G.sx + F_B - 1) / F_B;
  double* bs_blk = bs_all + (size_t)blockIdx.x * F_SUBSETS * F_TH * F_COLS;
  int* bi_blk = bi_all + (size_t)blockIdx.x * F_SUBSETS * F_TH * F_COLS;
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  uint32_t ph0 = 0, ph1 = 0;
  const uint32_t lbytes = (uint32_t)G.ltile_rows * F_COLS * 2, rrow_bytes = (uint32_t)G.rw * 2, qrow_bytes = (uint32_t)G.rw * 4;
  for (int item = blockIdx.x; item < G.NS * G.NB * G.J; item += gridDim.x) {
    const int chunk = item % G.J, rest = item / G.J;
    const int strip = rest % G.NS, band = rest / G.NS;
    const int y0 = band * F_TH, s0 = strip * G.out_cols;
    const int dy0 = chunk * G.dy_per, ndy = min(G.sy, dy0 + G.dy_per) - dy0;
    const int16_t* lsrc = L16 + ((size_t)strip * G.lrows + y0) * F_COLS;
    const int16_t* rsrc = R16 + ((size_t)strip * G.rrows + y0 + dy0) * G.rw;
    const float* qsrc = Qp + ((size_t)strip * qrows + y0 + dy0) * G.rw;
    const float* asrc = Ap + ((size_t)strip * qrows + y0 + dy0) * G.rw;
    const float* be_band = BEp + ((size_t)strip * G.NB * F_TH + y0) * F_COLS;
    if (tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(&bars[0], lbytes + (uint32_t)G.ltile_rows * rrow_bytes + 2u * F_TH * qrow_bytes);
      tma_load_1d(ltile, lsrc, lbytes, &bars[0]);
      tma_load_1d(rring, rsrc, (uint32_t)G.ltile_rows * rrow_bytes, &bars[0]);
      tma_load_1d(qring, qsrc, (uint32_t)F_TH * qrow_bytes, &bars[0]);       // window-origin rows y0+dy0 .. +31 -> slots 0..31
      tma_load_1d(aring, asrc, (uint32_t)F_TH * qrow_bytes, &bars[0]);
    }
    // thresholds: 0 for real pixels (every cost is >= 0: the first evaluation is always a candidate), +inf elsewhere
    for (int k = tid; k < F_TH * F_COLS; k += F_THREADS) {
      const int y = k / F_COLS, r = (k % F_COLS) / 32, l = k % 32;
      const int x = 8 * l + r;
      thr[k] = (x < G.out_cols && s0 + x < G.W && y0 + y < G.H) ? 0.0f : INFINITY;
    }
    for (int k = tid; k < F_SUBSETS * F_TH * F_COLS; k += F_THREADS) { bs_blk[k] = -1.0; bi_blk[k] = 0x7fffffff; }
    __syncthreads();
    mbar_wait(&bars[0], ph0); ph0 ^= 1;
    double* wbs = bs_blk + (size_t)sub * F_TH * F_COLS;
    int* wbi = bi_blk + (size_t)sub * F_TH * F_COLS;
    for (int dy = 0; dy < ndy; ++dy) {
      const int ring_base = (dy + row0) % G.ring_slots;
      const int qbase = (dy + row0) % NQ_SLOTS;
      if (tid == 0 && dy + 1 < ndy) {
        fence_proxy_async();
        mbar_expect_tx(&bars[1], rrow_bytes + 2u * qrow_bytes);
        tma_load_1d(rring + (size_t)((dy + G.ltile_rows) % G.ring_slots) * G.rw, rsrc + (size_t)(dy + G.ltile_rows) * G.rw, rrow_bytes, &bars[1]);
        tma_load_1d(qring + (size_t)((dy + F_TH) % NQ_SLOTS) * G.rw, qsrc + (size_t)(dy + F_TH) * G.rw, qrow_bytes, &bars[1]);
        tma_load_1d(aring + (size_t)((dy + F_TH) % NQ_SLOTS) * G.rw, asrc + (size_t)(dy + F_TH) * G.rw, qrow_bytes, &bars[1]);
      }
      for (int g = sub; g < ngroups; g += F_SUBSETS) {
        if (G.sx - F_B * g >= F_B)
          fast_pass_ncc<KX, true, CVT>(ltile, rring, qring, aring, thr, be_band, wbs, wbi, cx, lane, g, G.ky, G.ring_slots, G.rw, ring_base, qbase,
                                  row0, F_B, G.sx, dy0 + dy, s0, y0);
        else
          fast_pass_ncc<KX, false, CVT>(ltile, rring, qring, aring, thr, be_band, wbs, wbi, cx, lane, g, G.ky, G.ring_slots, G.rw, ring_base, qbase,
                                   row0, G.sx - F_B * g, G.sx, dy0 + dy, s0, y0);
      }
      if (dy + 1 < ndy) { mbar_wait(&bars[1], ph1); ph1 ^= 1; }
      __syncthreads();
    }
    // ---- merge the subsets' private bests (max cost, then first in raster order) ----
    const int nw = ngroups < F_SUBSETS ? ngroups : F_SUBSETS;
    for (int pix = tid; pix < F_TH * G.out_cols; pix += F_THREADS) {
      const int x = pix % G.out_cols, y = pix / G.out_cols;
      const int gx = s0 + x, gy = y0 + y;
      if (gx >= G.W || gy >= G.H) continue;
      const int off = y * F_COLS + (x & 7) * 32 + (x >> 3);
      double best = bs_blk[off];
      int bidx = bi_blk[off];
      for (int ww = 1; ww < nw; ++ww) {
        const double c = bs_blk[(size_t)ww * F_TH * F_COLS + off];
        const int i = bi_blk[(size_t)ww * F_TH * F_COLS + off];
        if (c > best || (c == best && i < bidx)) { best = c; bidx = i; }
      }
      if (G.J > 1) {
        const size_t pk = ((size_t)chunk * G.H + gy) * G.W + gx;
        part_cost[pk] = best; part_idx[pk] = bidx;
        continue;
      }
      if (bidx == 0x7fffffff) bidx = 0;
      vwb200_dispi o;
      o.dx = bidx % G.sx + G.addx; o.dy = bidx / G.sx + G.addy;
      o.valid = cx.nanflag[(size_t)gy * G.W + gx] ? 2 : 1;
      out[(ptrdiff_t)gy * opitch + gx] = o;
    }
    __syncthreads();
  }
}

__global__ void k1_fast_ncc_merge_kernel(FastGeom G, const double* __restrict__ pc, const int* __restrict__ pi,
                                         const unsigned char* __restrict__ nanflag, vwb200_dispi* __restrict__ out, ptrdiff_t opitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= G.W || y >= G.H) return;
  const size_t plane = (size_t)G.W * G.H, k = (size_t)y * G.W + x;
  double best = pc[k];
  int bidx = pi[k];
  for (int c = 1; c < G.J; ++c) {
    const double cc = pc[c * plane + k];
    const int ci = pi[c * plane + k];
    if (cc > best || (cc == best && ci < bidx)) { best = cc; bidx = ci; }
  }
  if (bidx == 0x7fffffff) bidx = 0;
  vwb200_dispi o;
  o.dx = bidx % G.sx + G.addx; o.dy = bidx / G.sx + G.addy;
  o.valid = nanflag[k] ? 2 : 1;
  out[(ptrdiff_t)y * opitch + x] = o;
}

// "every disparity gave the same cost" for pixels whose arg-best is (0,0): exact double costs, early exit
__global__ void k1_fast_ncc_allequal_fixup(ImgF L, ImgF R, FastGeom g, const double* __restrict__ lp, const double* __restrict__ rp,
                                           vwb200_dispi* __restrict__ out, ptrdiff_t opitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= g.W || y >= g.H) return;
  vwb200_dispi* o = out + (ptrdiff_t)y * opitch + x;
  if (o->valid != 1 || o->dx != g.addx || o->dy != g.addy) return;
  const int ow = g.W + g.sx - 1;
  const double lpv = lp[(size_t)y * g.W + x];
  auto cost = [&](int dx, int dy) {
    long long s = 0;
    for (int j = 0; j < g.ky; ++j)
      for (int i = 0; i < g.kx; ++i)
        s += (long long)(int)L.p[(ptrdiff_t)(y + j) * L.pitch + x + i] * (long long)(int)R.p[(ptrdiff_t)(y + j + dy) * R.pitch + x + i + dx];
    return __dmul_rn((double)s, sqrt(__dmul_rn(lpv, rp[(size_t)(y + dy) * ow + (x + dx)])));
  };
  const double c0 = cost(0, 0);
  for (int dy = 0; dy < g.sy; ++dy)
    for (int dx = 0; dx < g.sx; ++dx) {
      if (dx == 0 && dy == 0) continue;
      if (cost(dx, dy) != c0) return;
    }
  o->valid = 0;
}

int k1_fast_ncc_launch(ImgF left, ImgF right, int W, int H, int sx, int sy, int kx, int ky, float vmin, float vmax,
                       vwb200_dispi* out, ptrdiff_t opitch, void* workspace, cudaStream_t st, const KEvents* ev) {
  FastGeom g = make_geom(W, H, sx, sy, kx, ky);
  g.scale = 1;
  int c; double maxc;
  if (!ncc_params(kx, ky, vmin, vmax, &c, &maxc)) { set_error("k1_fast_ncc: unsupported value range"); return VWB200_ENOIMPL; }
  const NccWs ws = carve(g, workspace);
  const int N = kx * ky;
  const int ow = W + sx - 1, oh = H + sy - 1;
  const int qrows = g.NB * F_TH + sy;
  // every float operation of the bound rounds up, so no slack term is needed for rounding; E only keeps the bound strictly
  // above the exact value (a tie in the bound must still be a candidate)
  static int cvt = -1;
  if (cvt < 0) { const char* e = getenv("VWB200_NCC_CVT"); cvt = e ? atoi(e) : 1; if (cvt < 0 || cvt > 2) cvt = 1; }
  const float e_abs = cvt == 1 ? 130.0f : 1.0f;
  // exact maps (reference definitions), then the packed hot-loop operands
  VWB_TRY(box_sq_inv_launch(left, kx, ky, 0, 0, W, H, ws.lp, st));
  VWB_TRY(box_sq_inv_launch(right, kx, ky, 0, 0, ow, oh, ws.rp, st));
  VWB_TRY(box_sum_i32_launch(left, kx, ky, 0, 0, W, H, ws.Sl, st));
  VWB_TRY(box_sum_i32_launch(right, kx, ky, 0, 0, ow, oh, ws.Sr, st));
  VWB_CUDA(cudaMemsetAsync(ws.nanflag, 0, (size_t)W * H, st));
  {
    dim3 gl(g.lrows, g.NS), gr(g.rrows, g.NS), gq(qrows, g.NS), gb(g.NB * F_TH, g.NS);
    ncc_pack_img_kernel<<<gl, 256, 0, st>>>(left, c, g, 0, ws.L16);
    VWB_LAUNCH_CHECK();
    ncc_pack_img_kernel<<<gr, 256, 0, st>>>(right, c, g, 1, ws.R16);
    VWB_LAUNCH_CHECK();
    ncc_pack_qa_kernel<<<gq, 256, 0, st>>>(ws.rp, ws.Sr, c, g, qrows, cvt, ws.Qp, ws.Ap);
    VWB_LAUNCH_CHECK();
    ncc_pack_be_kernel<<<gb, 256, 0, st>>>(ws.Sl, c, N, e_abs, g, cvt, ws.BEp);
    VWB_LAUNCH_CHECK();
  }
  NccCtx cx{ws.lp, ws.rp, ws.Sl, ws.Sr, ws.nanflag, W, H, ow, c, (long long)N * c * c};
  int dev = 0, nsm = 148;
  VWB_CUDA(cudaGetDevice(&dev));
  VWB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  if (nsm > 148) nsm = 148;
  const int items = g.NS * g.NB * g.J;
  const int grid = items < nsm ? items : nsm;
  const size_t smem = ncc_smem_bytes(g);
  void (*kern)(const int16_t*, const int16_t*, const float*, const float*, const float*, FastGeom, int, NccCtx, double*, int*,
               vwb200_dispi*, ptrdiff_t, double*, int*) = nullptr;
  switch (kx) {
#define KCASE(K) case K: kern = k1_fast_ncc_kernel<K, 1>; break;
    KCASE(3) KCASE(5) KCASE(7) KCASE(9) KCASE(11) KCASE(13) KCASE(15) KCASE(17) KCASE(19) KCASE(21) KCASE(23) KCASE(25)
    KCASE(27) KCASE(29) KCASE(31)
#undef KCASE
    default: set_error("k1_fast_ncc: unsupported kernel width %d", kx); return VWB200_ENOIMPL;
  }
  if (kx == 21 && cvt == 0) kern = k1_fast_ncc_kernel<21, 0>;
  if (kx == 21 && cvt == 2) kern = k1_fast_ncc_kernel<21, 2>;
  if (kx != 21 && cvt != 1) { set_error("k1_fast_ncc: VWB200_NCC_CVT variants exist for kx = 21 only"); return VWB200_EARG; }
  VWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ev && ev->e0) cudaEventRecord(ev->e0, st);
  kern<<<grid, F_THREADS, smem, st>>>(ws.L16, ws.R16, ws.Qp, ws.Ap, ws.BEp, g, qrows, cx, ws.bs, ws.bi, out, opitch, ws.pc, ws.pi);
  VWB_LAUNCH_CHECK();
  if (ev && ev->e1) cudaEventRecord(ev->e1, st);
  dim3 b(32, 8), gg((W + 31) / 32, (H + 7) / 8);
  if (g.J > 1) {
    k1_fast_ncc_merge_kernel<<<gg, b, 0, st>>>(g, ws.pc, ws.pi, ws.nanflag, out, opitch);
    VWB_LAUNCH_CHECK();
  }
  k1_fast_ncc_allequal_fixup<<<gg, b, 0, st>>>(left, right, g, ws.lp, ws.rp, out, opitch);
  VWB_LAUNCH_CHECK();
  // NaN pixels (valid == 2): sequential replay of the reference's best/worst state machine
  {
    Zone z{};
    z.obase = 0; z.opitch = (int)opitch; z.w = W; z.h = H; z.lx = 0; z.ly = 0; z.rx = 0; z.ry = 0; z.sx = sx; z.sy = sy;
    z.addx = 0; z.addy = 0; z.nchunks = 1; z.sbase = 0;
    ncc_set_zone_kernel<<<1, 1, 0, st>>>(ws.zone, z);
    VWB_LAUNCH_CHECK();
    NccMaps maps{ws.lp, 0, 0, W, H, ws.rp, 0, 0, ow, oh};
    const long long px = (long long)W * H;
    const int gridx = (int)std::min<long long>((px + 127) / 128, 148 * 16);
    VWB_TRY(k1_nan_fixup_launch(VWB200_CROSS_CORRELATION, left, right, ws.zone, 1, kx, ky, maps, out, st, gridx));
  }
  return VWB200_OK;
}

}  // namespace vwb200
