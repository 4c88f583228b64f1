// k5_sgm_paths.cu -- path accumulation of SemiGlobalMatcher with a search box per pixel.
//
// SGM (accum_sgm_multithread, Stereo/SGM.cc:2462-2611; PixelPassTask, SGMAssist.h:691-832; evaluate_path SSE flavour,
// SGM.cc:1014-1141): along each of the 8 directions every pixel lies on exactly one line; a line is a sequential chain.
//   sgm_lines_kernel: ONE WARP PER LINE, lanes = the disparities of the pixel's search box (the common case: <= 32,
//   e.g. the 5 x 5 box around a half-resolution prior).  The previous pixel's path costs stay in registers in the layout of
//   ITS box; the reference scatters them into a full-size array filled with BAD and looks up the 8 adjacent disparities
//   (+ the same one) there -- here that is a 3 x 3 minimum over the intersection with the previous box:
//     min(min8 + P1, centre) == min(min9 + P1, centre)   (P1 > 0, saturating add is monotonic)
//   so the separable form needs 4 shuffles when the box did not move and 6 when it did (the lane then gathers, from the
//   clamped position in the old box, one of {v, row-min3, column-min3, 3x3-min} according to which axis was clamped).
//   min over the previous pixel = one REDUX.  Meta records (16 B), cost bytes and accumulated costs are prefetched
//   PF / 2*PF pixels ahead in registers.  Diagonal lines are wrapped around the image edge (line L owns column
//   (L + t) mod ow of row t and restarts at the edge), which makes all lines equally long and keeps the warps of a CTA on
//   neighbouring pixels of the same row (shared sectors in L1/L2).
//   Boxes with more than 32 disparities (full-search pixels) take a general warp-strided step through a per-warp scratch
//   line in global memory.
// Algorithmic bytes per (pixel, disparity) and direction: 1 (cost) + 2 + 2 (accum read / write); the first direction
// does not read accum.  Per pixel and direction: 16 B meta.
//
// MGM (accum_mgm_multithread, SGM.cc:2619-2700; SmoothPathAccumTask, SGMAssist.h:835-1239): eight sweeps; a pixel's
// path cost is the truncating mean of evaluate_path from TWO predecessors, so a sweep is a wavefront over the image
// (rows, columns or anti-diagonals).  mgm_sweep_kernel: persistent cooperative grid, one grid barrier per wavefront,
// a warp per pixel.
#include "k5_sgm.cuh"
#include <cooperative_groups.h>

namespace vwb200 {
namespace cg = cooperative_groups;

__constant__ unsigned short c_recip[33] = {0, 1024, 512, 342, 256, 205, 171, 147, 128, 114, 103, 94, 86, 79, 74, 69, 64, 61,
                                           57, 54, 52, 49, 47, 45, 43, 41, 40, 38, 37, 36, 35, 34, 32};   // ceil(1024 / w)

struct MetaR { int b0, b1, b2, b3; unsigned start; int val; int n; };
__device__ __forceinline__ MetaR unpack_meta(uint4 q) {
  MetaR m;
  m.b0 = (short)(q.x & 0xffff); m.b1 = (short)(q.x >> 16); m.b2 = (short)(q.y & 0xffff); m.b3 = (short)(q.y >> 16);
  m.start = q.z; m.val = (int)(q.w & 255u); m.n = (int)(q.w >> 8);
  return m;
}
__device__ __forceinline__ unsigned sat_add16(unsigned a, unsigned b) { return min(a + b, 65535u); }
__device__ __forceinline__ unsigned sat_sub16(unsigned a, unsigned b) { return a > b ? a - b : 0u; }

// evaluate_path for one pixel whose box (b, n entries) and whose predecessor's box (pb, pn entries, packed costs in
// prior[]) are arbitrary: warp-strided over the entries.  emit(e, value) receives the path cost of entry e.
template <bool CG, class Emit>
__device__ __forceinline__ void eval_path_general(const MetaR& m, int pb0, int pb1, int pb2, int pb3, int pn, const sgm_accum_t* prior,
                                                  const sgm_cost_t* __restrict__ cost, unsigned p1, unsigned p2_mod, unsigned BAD, int lane,
                                                  Emit emit) {
  unsigned mp = BAD;
  auto ld = [&](int i) -> unsigned { return CG ? (unsigned)__ldcg(prior + i) : (unsigned)prior[i]; };   // CG: written by other SMs
  for (int i = lane; i < pn; i += 32) mp = min(mp, ld(i));
  mp = __reduce_min_sync(0xffffffffu, mp);
  const unsigned dJ = (mp + p2_mod) & 0xffffu;
  const int w = m.b2 - m.b0 + 1, pw = pb2 - pb0 + 1;
  for (int e = lane; e < m.n; e += 32) {
    const int y = e / w, x = e - y * w, X = m.b0 + x, Y = m.b1 + y;
    unsigned nb = BAD, centre = BAD;
    if (pn > 0) {
      const int y0 = max(Y - 1, pb1), y1 = min(Y + 1, pb3), x0 = max(X - 1, pb0), x1 = min(X + 1, pb2);
      for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) {
          const unsigned v = ld((yy - pb1) * pw + (xx - pb0));
          nb = min(nb, v);
          if (xx == X && yy == Y) centre = v;
        }
    }
    unsigned res = sat_add16(nb, p1);
    res = min(res, min(centre, dJ));
    res = sat_add16(res, (unsigned)cost[(size_t)m.start + e]);
    emit(e, sat_sub16(res, mp));
  }
}

template <int PF>
__global__ void __launch_bounds__(256) sgm_lines_kernel(const SgmMeta* __restrict__ meta, const sgm_cost_t* __restrict__ cost,
                                                        sgm_accum_t* __restrict__ accum, SgmGeom g, int sc, int sr, int first_dir,
                                                        sgm_accum_t* __restrict__ scratch, unsigned scratch_per_warp) {
  __shared__ unsigned short s_p2mod[256];
  for (unsigned d = threadIdx.x; d < 256; d += blockDim.x) {          // p2_mod as a function of the grey-value step (:1026-1031)
    const unsigned p2u = (unsigned)g.p2 & 0xffffu;                   // (accum_t)p2
    unsigned v = d ? p2u / d : p2u;
    if (v < (unsigned)g.p1) v = (unsigned)g.p1;
    s_p2mod[d] = (unsigned short)v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const bool horiz = sr == 0;
  const int NL = horiz ? g.oh : g.ow, NS = horiz ? g.ow : g.oh;
  if (line >= NL) return;
  const unsigned BAD = (unsigned)((255 + g.p2) & 0xffff);            // get_bad_accum_val (SGM.h:240) as accum_t
  const unsigned p1 = (unsigned)g.p1;
  sgm_accum_t* bufA = scratch ? scratch + (size_t)line * scratch_per_warp : nullptr;
  sgm_accum_t* bufB = bufA ? bufA + scratch_per_warp / 2 : nullptr;

  // cursors of the three pipeline stages: meta loads run 2*PF pixels ahead, data loads PF pixels ahead
  int cM, rM;
  if (horiz) { rM = line; cM = sc > 0 ? 0 : g.ow - 1; }
  else { cM = line; rM = sr > 0 ? 0 : g.oh - 1; }
  auto adv = [&](int& c, int& r) {
    r += sr; c += sc;
    if (!horiz) { if (c >= g.ow) c = 0; else if (c < 0) c = g.ow - 1; }
  };
  auto in_img = [&](int c, int r) { return c >= 0 && c < g.ow && r >= 0 && r < g.oh; };
  uint4 mA[PF], mB[PF], mC[PF];
  unsigned dcA[PF], daA[PF], dcB[PF], daB[PF];
  const uint4 MZ = make_uint4(0xffff0000u, 0xffffu, 0u, 0u);       // box (0,0,-1,-1), n = 0
  auto load_meta = [&](uint4* dst, int t0) {
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      dst[i] = (t0 + i < NS && in_img(cM, rM)) ? __ldg(reinterpret_cast<const uint4*>(meta) + ((size_t)rM * g.ow + cM)) : MZ;
      adv(cM, rM);
    }
  };
  auto load_data = [&](const uint4* m, unsigned* dc, unsigned* da) {
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const unsigned n = m[i].w >> 8;
      const bool act = (unsigned)lane < n && n <= 32u;
      dc[i] = act ? (unsigned)cost[(size_t)m[i].z + lane] : 0u;
      da[i] = (act && !first_dir) ? (unsigned)accum[(size_t)m[i].z + lane] : 0u;
    }
  };
  load_meta(mA, 0);
  load_meta(mB, PF);
  load_data(mA, dcA, daA);

  // chain state
  int last_val = -1;
  unsigned pv = 0xffffu;                 // previous pixel's path cost of this lane's entry (its box layout); 0xffff = no entry
  int pn = 0;                            // previous pixel's number of entries
  unsigned pkx = 0xffffffffu, pky = 0xffffffffu;     // previous pixel's packed box
  int pb0 = 0, pb1 = 0, pb2 = -1, pb3 = -1, x = 0, y = 0, w = 1, h = 1;   // previous box and this lane's position in it
  bool prev_in_buf = false;              // pn > 32: the previous costs live in bufA

  // compute cursor only for the "line restarts" test of wrapped diagonals
  int cC = horiz ? (sc > 0 ? 0 : g.ow - 1) : line;

  for (int t0 = 0; t0 < NS; t0 += PF) {
    load_meta(mC, t0 + 2 * PF);
    load_data(mB, dcB, daB);
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int t = t0 + i;
      if (t < NS) {                      // warp-uniform
        const MetaR m = unpack_meta(mA[i]);
        const bool restart = (t == 0) || (!horiz && sc != 0 && cC == (sc > 0 ? 0 : g.ow - 1));
        if (restart) last_val = -1;
        const bool changed = mA[i].x != pkx || mA[i].y != pky;
        int nx = x, ny = y, nw = w, nh = h;
        if (changed && m.n <= 32) {
          nw = max(m.b2 - m.b0 + 1, 1); nh = m.b3 - m.b1 + 1;
          ny = (lane * (int)c_recip[min(nw, 32)]) >> 10; nx = lane - ny * nw;
        }
        const bool act = lane < m.n;
        unsigned cur = 0xffffu;
        if (m.n <= 32 && (last_val < 0 || pn <= 32)) {
          // ---------------- register path ----------------
          const unsigned local = dcA[i];
          if (last_val < 0) {
            cur = local;                                                // first pixel of a line (SGMAssist.h:756-759)
          } else {
            const unsigned p2_mod = s_p2mod[abs(m.val - last_val)];
            const unsigned mp = min(__reduce_min_sync(0xffffffffu, pv), BAD);
            const unsigned dJ = (mp + p2_mod) & 0xffffu;
            // separable 3 x 3 minimum in the previous box's layout (lanes >= pn hold 0xffff and are never addressed)
            const unsigned vl = __shfl_up_sync(0xffffffffu, pv, 1), vr = __shfl_down_sync(0xffffffffu, pv, 1);
            const unsigned hmin = min(min(x > 0 ? vl : BAD, pv), x < w - 1 ? vr : BAD);
            unsigned nb, centre;
            if (!changed) {
              const unsigned hu = __shfl_up_sync(0xffffffffu, hmin, w), hd = __shfl_down_sync(0xffffffffu, hmin, w);
              nb = min(min(y > 0 ? hu : BAD, hmin), y < h - 1 ? hd : BAD);
              centre = pv;
            } else {
              const unsigned BB = BAD | (BAD << 16);
              const unsigned q = (pv & 0xffffu) | (hmin << 16);
              unsigned qu = __shfl_up_sync(0xffffffffu, q, w), qd = __shfl_down_sync(0xffffffffu, q, w);
              if (y == 0) qu = BB;
              if (y >= h - 1) qd = BB;
              const unsigned wm = __vimin3_u16x2(qu, q, qd);           // lo: column min3 of v, hi: 3 x 3 min
              const int X = m.b0 + nx, Y = m.b1 + ny;
              const int sxp = min(max(X, pb0), pb2), syp = min(max(Y, pb1), pb3);
              const int ddx = X - sxp, ddy = Y - syp;
              const int src = ((syp - pb1) * w + (sxp - pb0)) & 31;
              const unsigned g1 = __shfl_sync(0xffffffffu, q, src), g2 = __shfl_sync(0xffffffffu, wm, src);
              const bool reach = pn > 0 && ddx >= -1 && ddx <= 1 && ddy >= -1 && ddy <= 1;
              centre = BAD;
              if (!reach) nb = BAD;
              else if (ddx == 0 && ddy == 0) { nb = g2 >> 16; centre = g1 & 0xffffu; }
              else if (ddy == 0) nb = g2 & 0xffffu;
              else if (ddx == 0) nb = g1 >> 16;
              else nb = g1 & 0xffffu;
            }
            unsigned res = sat_add16(nb, p1);
            res = min(res, min(centre, dJ));
            res = sat_add16(res, local);
            cur = sat_sub16(res, mp);
          }
          if (act) accum[(size_t)m.start + lane] = (sgm_accum_t)(daA[i] + cur);      // update_accum_buffer, uint16 wrap
          if (!act) cur = 0xffffu;
          prev_in_buf = false;
        } else {
          // ---------------- general path: a box with more than 32 disparities is involved ----------------
          if (!prev_in_buf && last_val >= 0) { if (lane < pn) bufA[lane] = (sgm_accum_t)pv; __syncwarp(); }
          if (last_val < 0) {
            for (int e = lane; e < m.n; e += 32) {
              const unsigned c0 = cost[(size_t)m.start + e];
              bufB[e] = (sgm_accum_t)c0;
              const unsigned a0 = first_dir ? 0u : (unsigned)accum[(size_t)m.start + e];
              accum[(size_t)m.start + e] = (sgm_accum_t)(a0 + c0);
            }
          } else {
            const unsigned p2_mod = s_p2mod[abs(m.val - last_val)];
            eval_path_general<false>(m, pb0, pb1, pb2, pb3, pn, bufA, cost, p1, p2_mod, BAD, lane, [&](int e, unsigned v) {
              bufB[e] = (sgm_accum_t)v;
              const unsigned a0 = first_dir ? 0u : (unsigned)accum[(size_t)m.start + e];
              accum[(size_t)m.start + e] = (sgm_accum_t)(a0 + v);
            });
          }
          __syncwarp();
          sgm_accum_t* sw = bufA; bufA = bufB; bufB = sw;
          prev_in_buf = m.n > 32;
          cur = (m.n <= 32 && act) ? (unsigned)bufA[lane] : 0xffffu;
        }
        // this pixel becomes the predecessor
        pv = cur; pn = m.n; last_val = m.val;
        if (changed) { pkx = mA[i].x; pky = mA[i].y; pb0 = m.b0; pb1 = m.b1; pb2 = m.b2; pb3 = m.b3; x = nx; y = ny; w = nw; h = nh; }
        if (horiz) cC += sc; else { cC += sc; if (cC >= g.ow) cC = 0; else if (cC < 0) cC = g.ow - 1; }
      }
    }
#pragma unroll
    for (int i = 0; i < PF; ++i) { mA[i] = mB[i]; mB[i] = mC[i]; dcA[i] = dcB[i]; daA[i] = daB[i]; }
  }
}

int sgm_paths_launch(const SgmMeta* meta, const sgm_cost_t* cost, sgm_accum_t* accum, const SgmGeom& g, unsigned max_n, Arena& ar,
                     cudaStream_t st) {
  // the eight directions of accum_sgm_multithread (:2462-2611), one launch each (a pixel lies on one line per direction)
  static const int DIRS[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};
  sgm_accum_t* scratch = nullptr;
  unsigned per_warp = 0;
  if (max_n > 32) {
    per_warp = 2 * ((max_n + 31u) & ~31u);
    VWB_TRY(ar.alloc(&scratch, (size_t)per_warp * std::max(g.ow, g.oh)));
  }
  for (int i = 0; i < 8; ++i) {
    const int sc = DIRS[i][0], sr = DIRS[i][1];
    const int lines = sr == 0 ? g.oh : g.ow;
    const int wpb = 4;                                                 // warps per CTA
    sgm_lines_kernel<2><<<(lines + wpb - 1) / wpb, wpb * 32, 0, st>>>(meta, cost, accum, g, sc, sr, i == 0, scratch, per_warp);
    VWB_LAUNCH_CHECK();
  }
  return VWB200_OK;
}

// ---- MGM -------------------------------------------------------------------------------------------------------------
struct MgmTask { int p1c, p1r, p2c, p2r, dirx, diry, need_r_gt0, need_r_lt, need_c_gt0, need_c_lt, wf; };
// wf: wavefront index  0: c + r   1: (ow-1-c) + (oh-1-r)   2: r   3: oh-1-r   4: r + (ow-1-c)   5: (oh-1-r) + c   6: ow-1-c   7: c
__constant__ MgmTask c_mgm_tasks[8] = {
    /* L  */ {-1, 0, 0, -1, -1, 0, 1, 0, 1, 0, 0},
    /* R  */ {1, 0, 0, 1, 1, 0, 0, 1, 0, 1, 1},
    /* TL */ {-1, -1, 1, -1, -1, -1, 1, 0, 1, 1, 2},
    /* BR */ {1, 1, -1, 1, 1, 1, 0, 1, 1, 1, 3},
    /* T  */ {0, -1, 1, 0, 0, -1, 1, 0, 0, 1, 4},
    /* B  */ {0, 1, -1, 0, 0, 1, 0, 1, 1, 0, 5},
    /* TR */ {1, -1, 1, 1, 1, -1, 1, 1, 0, 1, 6},
    /* BL */ {-1, 1, -1, -1, -1, 1, 1, 1, 1, 0, 7},
};

__global__ void __launch_bounds__(256) mgm_sweep_kernel(const SgmMeta* __restrict__ meta, const sgm_cost_t* __restrict__ cost,
                                                        sgm_accum_t* __restrict__ accum, sgm_accum_t* path, const uint8_t* __restrict__ left8,
                                                        SgmGeom g, int task) {
  cg::grid_group grid = cg::this_grid();
  const MgmTask t = c_mgm_tasks[task];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned BAD = (unsigned)((255 + g.p2) & 0xffff), p1 = (unsigned)g.p1, p2u = (unsigned)g.p2 & 0xffffu;
  const int last_c = g.ow - 1, last_r = g.oh - 1;
  const int nfronts = t.wf < 2 || t.wf == 4 || t.wf == 5 ? g.ow + g.oh - 1 : (t.wf == 2 || t.wf == 3 ? g.oh : g.ow);
  for (int f = 0; f < nfronts; ++f) {
    // pixels of this wavefront: parametrised by k
    int k0 = 0, k1 = 0;                    // k range [k0, k1)
    if (t.wf == 2 || t.wf == 3) { k1 = g.ow; }
    else if (t.wf == 6 || t.wf == 7) { k1 = g.oh; }
    else { k0 = max(0, f - (g.oh - 1)); k1 = min(g.ow - 1, f) + 1; }      // k = column-like coordinate u, v = f - u
    for (int k = k0 + warp; k < k1; k += nwarps) {
      int c, r;
      switch (t.wf) {
        case 0: c = k; r = f - k; break;
        case 1: c = last_c - k; r = last_r - (f - k); break;
        case 2: c = k; r = f; break;
        case 3: c = k; r = last_r - f; break;
        case 4: c = last_c - k; r = f - k; break;
        case 5: c = k; r = last_r - (f - k); break;
        case 6: c = last_c - f; r = k; break;
        default: c = f; r = k; break;
      }
      const size_t pix = (size_t)r * g.ow + c;
      const MetaR m = unpack_meta(__ldg(reinterpret_cast<const uint4*>(meta) + pix));
      if (m.n == 0) continue;
      sgm_accum_t* out = path + m.start;
      const bool ok = (!t.need_r_gt0 || r > 0) && (!t.need_r_lt || r < last_r) && (!t.need_c_gt0 || c > 0) && (!t.need_c_lt || c < last_c);
      if (!ok) {
        for (int e = lane; e < m.n; e += 32) {
          const unsigned v = cost[(size_t)m.start + e];
          out[e] = (sgm_accum_t)v;
          accum[(size_t)m.start + e] = (sgm_accum_t)(accum[(size_t)m.start + e] + v);
        }
        continue;
      }
      // get_path_pixel_diff (SGM.cc:2715-2721) looks at the pixel OPPOSITE to the direction it is given -- kept
      // (it may lie in the kernel padding outside the output area, hence the image and not the meta record)
      const int diff = abs(m.val - (int)left8[(size_t)(r - t.diry + g.min_row) * g.lw + (c - t.dirx + g.min_col)]);
      unsigned p2_mod = diff > 0 ? p2u / (unsigned)diff : p2u;
      if (p2_mod < p1) p2_mod = p1;
      const MetaR q1 = unpack_meta(__ldg(reinterpret_cast<const uint4*>(meta) + ((size_t)(r + t.p1r) * g.ow + (c + t.p1c))));
      const MetaR q2 = unpack_meta(__ldg(reinterpret_cast<const uint4*>(meta) + ((size_t)(r + t.p2r) * g.ow + (c + t.p2c))));
      eval_path_general<true>(m, q1.b0, q1.b1, q1.b2, q1.b3, q1.n, path + q1.start, cost, p1, p2_mod, BAD, lane,
                        [&](int e, unsigned v) { out[e] = (sgm_accum_t)v; });
      __syncwarp();
      eval_path_general<true>(m, q2.b0, q2.b1, q2.b2, q2.b3, q2.n, path + q2.start, cost, p1, p2_mod, BAD, lane, [&](int e, unsigned v) {
        const unsigned mean = ((unsigned)out[e] + v) / 2u;
        out[e] = (sgm_accum_t)mean;
        accum[(size_t)m.start + e] = (sgm_accum_t)(accum[(size_t)m.start + e] + mean);
      });
    }
    __threadfence();
    grid.sync();
  }
}

int mgm_paths_launch(const SgmMeta* meta, const sgm_cost_t* cost, sgm_accum_t* accum, const uint8_t* left8, const SgmGeom& g, size_t total,
                     Arena& ar, cudaStream_t st) {
  sgm_accum_t* path;
  VWB_TRY(ar.alloc(&path, total + 64));
  int dev = 0, sms = 0, per_sm = 0;
  VWB_CUDA(cudaGetDevice(&dev));
  VWB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  VWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mgm_sweep_kernel, 256, 0));
  if (per_sm < 1) { set_error("mgm: the sweep kernel does not fit on an SM"); return VWB200_ECUDA; }
  const int blocks = sms * std::min(per_sm, 4);
  for (int task = 0; task < 8; ++task) {
    SgmGeom gg = g;
    void* args[] = {(void*)&meta, (void*)&cost, (void*)&accum, (void*)&path, (void*)&left8, (void*)&gg, (void*)&task};
    VWB_CUDA(cudaLaunchCooperativeKernel((void*)mgm_sweep_kernel, dim3(blocks), dim3(256), args, 0, st));
    VWB_LAUNCH_CHECK();
  }
  return VWB200_OK;
}

}  // namespace vwb200
