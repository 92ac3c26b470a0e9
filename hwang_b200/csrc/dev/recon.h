// Macroblock reconstruction: dequantisation + inverse transforms (H.264 clause 8.5), intra
// prediction (8.3), inter prediction with quarter-pel luma / eighth-pel chroma interpolation
// and weighted prediction (8.4.2).  One warp reconstructs one macroblock; a warp walks one
// macroblock row left to right, rows of a picture form a wavefront (see kernels.cu).
//
// This is the GPU replacement for the work libavcodec does per macroblock inside
// avcodec_send_packet in the reference (software_video_decoder.cpp:349-402 -> h264 decoder).
#pragma once
#include "ir.h"
#include "tables_gen.h"

namespace hwb {

enum { LT_STRIDE = 48, LT_OFF = 16, CT_STRIDE = 32, CT_OFF = 16 };

// Per-warp scratch ("shared memory" on the device).
struct alignas(16) ReconScratch {
  // luma tile rows -1..15 (row r at (r+1)*LT_STRIDE), column c at LT_OFF + c, c in -1..23
  uint8_t luma[17 * LT_STRIDE];
  // chroma tiles rows -1..7, column c at CT_OFF + c
  uint8_t chroma[2][9 * CT_STRIDE];
  int16_t res[24][16];  // residual per 4x4 block: 16 luma (z order; 8x8 blocks use 4 slots as 64), 4 Cb, 4 Cr
  int32_t dc[24];       // dequantised DC per block (Intra16x16 luma DC, chroma DC)
  uint32_t has_res;     // bit b: res[b] is non-zero
  // Motion compensation (staged reference tiles, BASELINE.json north_star "shared-memory staging of reference-block
  // halos"): the warp loads the (w+5) x (h+5) luma window of a partition once, as aligned 32-bit words straight into
  // `mc_luma` (row stride MC_LS bytes = 9 words: consecutive rows fall on different banks), likewise the two chroma
  // windows, and every lane interpolates its samples from shared memory.  `mc_h` holds the unrounded horizontal
  // half-sample intermediates of the positions that need the two-dimensional filter.  `pred` receives the prediction of
  // each list when two lists or weights have to be combined.
  uint8_t mc_luma[21 * 36];
  uint8_t mc_chroma[2][9 * 12];
  int16_t mc_h[21 * 16];
  uint8_t pred[2][384];
};
enum { MC_LS = 36, MC_CS = 12 };

// ---------------------------------------------------------------------------------------------
// transforms
// ---------------------------------------------------------------------------------------------
HWB_HD void idct4_1d(int &d0, int &d1, int &d2, int &d3) {
  int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
  d0 = e0 + e3; d1 = e1 + e2; d2 = e1 - e2; d3 = e0 - e3;
}

// coef: 16 raw levels (raster) or nullptr (all AC zero).  dc_override: replaces d[0] after scaling.
HWB_FN void residual4x4(const int16_t *coef, bool use_dc, int dc, const uint8_t *scaling, int qp, int16_t *out) {
  int d[16];
  const int qm = qp % 6, qs = qp / 6;
  if (coef) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int c = coef[i];
      int ls = (int)scaling[i] * dequant4_v[qm * 16 + i];
      d[i] = qs >= 4 ? (c * ls) << (qs - 4) : (c * ls + (1 << (3 - qs))) >> (4 - qs);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = 0;
  }
  if (use_dc) d[0] = dc;
#pragma unroll
  for (int r = 0; r < 4; ++r) idct4_1d(d[4 * r], d[4 * r + 1], d[4 * r + 2], d[4 * r + 3]);
#pragma unroll
  for (int c = 0; c < 4; ++c) idct4_1d(d[c], d[4 + c], d[8 + c], d[12 + c]);
#pragma unroll
  for (int i = 0; i < 16; ++i) out[i] = (int16_t)((d[i] + 32) >> 6);
}

HWB_HD void idct8_1d(int *d, int s) {
  int a0 = d[0] + d[4 * s], a4 = d[0] - d[4 * s];
  int a2 = (d[2 * s] >> 1) - d[6 * s], a6 = d[2 * s] + (d[6 * s] >> 1);
  int a1 = -d[3 * s] + d[5 * s] - d[7 * s] - (d[7 * s] >> 1);
  int a3 = d[1 * s] + d[7 * s] - d[3 * s] - (d[3 * s] >> 1);
  int a5 = -d[1 * s] + d[7 * s] + d[5 * s] + (d[5 * s] >> 1);
  int a7 = d[3 * s] + d[5 * s] + d[1 * s] + (d[1 * s] >> 1);
  int b0 = a0 + a6, b2 = a4 + a2, b4 = a4 - a2, b6 = a0 - a6;
  int b1 = a1 + (a7 >> 2), b7 = a7 - (a1 >> 2), b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5;
  d[0] = b0 + b7; d[1 * s] = b2 + b5; d[2 * s] = b4 + b3; d[3 * s] = b6 + b1;
  d[4 * s] = b6 - b1; d[5 * s] = b4 - b3; d[6 * s] = b2 - b5; d[7 * s] = b0 - b7;
}

HWB_FN void residual8x8(const int16_t *coef, const uint8_t *scaling, int qp, int16_t *out) {
  int d[64];
  const int qm = qp % 6, qs = qp / 6;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
    int c = coef[i];
    int ls = (int)scaling[i] * dequant8_v[qm * 64 + i];
    d[i] = qs >= 6 ? (c * ls) << (qs - 6) : (c * ls + (1 << (5 - qs))) >> (6 - qs);
  }
#pragma unroll 1
  for (int r = 0; r < 8; ++r) idct8_1d(d + 8 * r, 1);
#pragma unroll 1
  for (int c = 0; c < 8; ++c) idct8_1d(d + c, 8);
#pragma unroll 1
  for (int i = 0; i < 64; ++i) out[i] = (int16_t)((d[i] + 32) >> 6);
}

// Intra16x16 luma DC: 16 levels in raster order of the 4x4 DC matrix -> dequantised DC per block
// (written to dc[z] for the block at that raster position).
HWB_FN void luma_dc_transform(const int16_t *c, int ls00, int qp, int32_t *dc_by_z) {
  int f[16];
  for (int i = 0; i < 16; ++i) f[i] = c[i];
  for (int r = 0; r < 4; ++r) {
    int a = f[4 * r], b = f[4 * r + 1], cc = f[4 * r + 2], d = f[4 * r + 3];
    f[4 * r] = a + b + cc + d; f[4 * r + 1] = a + b - cc - d; f[4 * r + 2] = a - b - cc + d; f[4 * r + 3] = a - b + cc - d;
  }
  for (int k = 0; k < 4; ++k) {
    int a = f[k], b = f[4 + k], cc = f[8 + k], d = f[12 + k];
    f[k] = a + b + cc + d; f[4 + k] = a + b - cc - d; f[8 + k] = a - b - cc + d; f[12 + k] = a - b + cc - d;
  }
  const int qs = qp / 6;
  for (int i = 0; i < 16; ++i) {
    int v = qs >= 6 ? (f[i] * ls00) << (qs - 6) : (f[i] * ls00 + (1 << (5 - qs))) >> (6 - qs);
    dc_by_z[xy2z(i & 3, i >> 2)] = v;
  }
}

HWB_FN void chroma_dc_transform(const int16_t *c, int ls00, int qpc, int32_t *dc4) {
  int f0 = c[0] + c[1] + c[2] + c[3], f1 = c[0] - c[1] + c[2] - c[3];
  int f2 = c[0] + c[1] - c[2] - c[3], f3 = c[0] - c[1] - c[2] + c[3];
  const int qs = qpc / 6;
  dc4[0] = ((f0 * ls00) << qs) >> 5; dc4[1] = ((f1 * ls00) << qs) >> 5;
  dc4[2] = ((f2 * ls00) << qs) >> 5; dc4[3] = ((f3 * ls00) << qs) >> 5;
}

HWB_HD int chroma_qp(int qp, int off) { return chroma_qp_table[clip3(0, 51, qp + off)]; }

// ---------------------------------------------------------------------------------------------
// intra prediction
// ---------------------------------------------------------------------------------------------
// Directional predictors for NxN blocks (N = 4 or 8) from a linear edge array E:
//   E[N] = corner p[-1,-1], E[N-1-y] = left p[-1,y], E[N+1+x] = top p[x,-1] (x up to 2N-1).
HWB_HD int f3(const uint8_t *E, int k) { return (E[k - 1] + 2 * E[k] + E[k + 1] + 2) >> 2; }
HWB_HD int f2(const uint8_t *E, int a, int b) { return (E[a] + E[b] + 1) >> 1; }

HWB_HD int intra_dir_pred(int mode, int N, const uint8_t *E, int x, int y) {
  const int T = N + 1;  // top(x) = E[T+x]; left(y) = E[N-1-y]
  switch (mode) {
    case 0: return E[T + x];
    case 1: return E[N - 1 - y];
    case 3:  // diagonal down-left
      if (x == N - 1 && y == N - 1) return (E[T + 2 * N - 2] + 3 * E[T + 2 * N - 1] + 2) >> 2;
      return f3(E, T + x + y + 1);
    case 4:  // diagonal down-right
      return f3(E, N + x - y);
    case 5: {  // vertical-right
      int z = 2 * x - y;
      if (z >= 0) {
        int k = T + x - (y >> 1);
        return (z & 1) ? f3(E, k - 1) : f2(E, k - 1, k);
      }
      if (z == -1) return f3(E, N);
      return f3(E, N - 1 - (y - 2 * x - 2));  // left(y-2x-2) centre
    }
    case 6: {  // horizontal-down
      int z = 2 * y - x;
      if (z >= 0) {
        int k = N - 1 - (y - (x >> 1));  // left(y-(x>>1))
        return (z & 1) ? f3(E, k + 1) : f2(E, k + 1, k);
      }
      if (z == -1) return f3(E, N);
      return f3(E, T + (x - 2 * y - 2));  // top(x-2y-2) centre
    }
    case 7: {  // vertical-left
      int k = T + x + (y >> 1);
      return (y & 1) ? f3(E, k + 1) : f2(E, k, k + 1);
    }
    default: {  // 8: horizontal-up
      int z = x + 2 * y;
      if (z > 2 * N - 3) return E[0];
      if (z == 2 * N - 3) return (E[1] + 3 * E[0] + 2) >> 2;
      int k = N - 1 - (y + (x >> 1));  // left(y + (x>>1))
      return (z & 1) ? f3(E, k - 1) : f2(E, k, k - 1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// inter prediction
// ---------------------------------------------------------------------------------------------
HWB_HD int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

#if !HWB_DEVICE_BUILD
// Host emulation only: the window of the reference pictures the current macroblock was promised to stay inside
// (what the picture kernel waited for: ChunkCtx::mv_reach / mv_reach_x).  Every reference sample fetched is checked
// against it, so a reach computed too small by the entropy stage fails a CPU test instead of racing on the GPU.
struct McWindow { int max_row = 1 << 30, max_col = 1 << 30, violations = 0; };
static thread_local McWindow g_mc_window;
static inline void mc_check(int row, int col) { if (row > g_mc_window.max_row || col > g_mc_window.max_col) g_mc_window.violations++; }
#endif

// One motion-compensated partition: w x h luma samples (16x16, 8x8 or 4x4) at (x0,y0) of the picture, motion vector
// (mvx,mvy) in quarter samples, predicted from frame `rf` into dst_y (stride ds_y) and dst_cb / dst_cr (stride ds_c).
// Warp-cooperative: stage the reference windows, then lane l computes `ppl` consecutive samples of one row.  Written for
// few instructions and a small footprint (this stage is bound by instruction fetch, not by memory): lane mappings are
// shifts and masks, horizontal filters slide a six-sample window along the row (one load per sample).
HWB_FN void mc_partition(const ChunkCtx &c, int rf, int x0, int y0, int w, int h, int mvx, int mvy, uint8_t *dst_y, int ds_y,
                         uint8_t *dst_cb, uint8_t *dst_cr, int ds_c, ReconScratch *sm) {
  const int W = c.wc, H = c.hc;
  const uint8_t *ref = frame_y(c, rf);
  const int fx = mvx & 3, fy = mvy & 3;
  const int ox = x0 + (mvx >> 2) - 2, oy = y0 + (mvy >> 2) - 2;  // window origin
  const int cols = w + 5, rows = h + 5;
  const bool inside = ox >= 0 && oy >= 0 && ox + cols <= W && oy + rows <= H;
  const int sh = inside ? (ox & 3) : 0;
  const int cw = W >> 1, ch = H >> 1;
  const int cx = (x0 >> 1) + (mvx >> 3), cy = (y0 >> 1) + (mvy >> 3);
  const int ccols = (w >> 1) + 1, crows = (h >> 1) + 1;
  const bool cinside = cx >= 0 && cy >= 0 && cx + ccols <= cw && cy + crows <= ch;
  const int csh = cinside ? (cx & 3) : 0;
  // ---- stage luma (aligned words: lane = (row & 3, word)) and chroma (lane = (plane, row & 3, word))
  HWB_LANES(l)
    if (inside) {
      const int nw = (sh + cols + 3) >> 2;  // aligned words per row (<= 7); the over-read of <= 3 bytes stays inside the slab
      const int k = l & 7;
      const uint8_t *base = ref + (int64_t)oy * W + (ox - sh) + 4 * k;
      if (k < nw) {
        // all of a lane's loads (rows l>>3, +4, ...: at most 6 of the <= 21) are issued before the first store: they go
        // to L2 (ld.global.cg), and one at a time their latency was a quarter of the macroblock's time
        uint32_t v[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { const int r = (l >> 3) + 4 * i; if (r < rows) v[i] = ld_u32_cg((const uint32_t *)(base + (int64_t)r * W)); }
#pragma unroll
        for (int i = 0; i < 6; ++i) { const int r = (l >> 3) + 4 * i; if (r < rows) *(uint32_t *)(sm->mc_luma + r * MC_LS + 4 * k) = v[i]; }
      }
    } else if (l < cols) {  // window crosses the picture edge: one (clamped) column per lane
      const uint8_t *col = ref + clip3(0, W - 1, ox + l);
#pragma unroll 1
      for (int r = 0; r < rows; ++r) sm->mc_luma[r * MC_LS + l] = ld_u8_cg(col + (int64_t)clip3(0, H - 1, oy + r) * W);
    }
    {
      const int pl = l >> 4;
      const uint8_t *P = pl ? frame_cr(c, rf) : frame_cb(c, rf);
      if (cinside) {
        const int nw = (csh + ccols + 3) >> 2, k = l & 3;  // <= 3 words
        if (k < nw) {
          uint32_t v[3];  // rows (l>>2)&3, +4, +8 of the <= 9
#pragma unroll
          for (int i = 0; i < 3; ++i) { const int r = ((l >> 2) & 3) + 4 * i; if (r < crows) v[i] = ld_u32_cg((const uint32_t *)(P + (int64_t)(cy + r) * cw + (cx - csh)) + k); }
#pragma unroll
          for (int i = 0; i < 3; ++i) { const int r = ((l >> 2) & 3) + 4 * i; if (r < crows) *(uint32_t *)(sm->mc_chroma[pl] + r * MC_CS + 4 * k) = v[i]; }
        }
      } else if ((l & 15) < ccols) {
        const uint8_t *col = P + clip3(0, cw - 1, cx + (l & 15));
#pragma unroll 1
        for (int r = 0; r < crows; ++r) sm->mc_chroma[pl][r * MC_CS + (l & 15)] = ld_u8_cg(col + (int64_t)clip3(0, ch - 1, cy + r) * cw);
      }
    }
  HWB_LANES_END
#if !HWB_DEVICE_BUILD
  mc_check(clip3(0, H - 1, oy + rows - 1), clip3(0, W - 1, ox + cols - 1));
  mc_check(2 * clip3(0, ch - 1, cy + crows - 1) + 1, 2 * clip3(0, cw - 1, cx + ccols - 1) + 1);  // in luma units
#endif
  // ---- luma interpolation (8.4.2.2.1).  T(r,k): window sample, the partition's sample (0,0) is T(2,2).
  const uint8_t *T = sm->mc_luma + sh;
#define HWB_T(r, k) ((int)T[(r) * MC_LS + (k)])
  const bool need_j = (fx == 2 && fy != 0) || (fy == 2 && fx != 0);
  if (need_j) {
    // unrounded horizontal half samples of every window row (one row per lane), the vertical filter runs over them
    HWB_LANES(l)
      if (l < rows) {
        const uint8_t *row = T + l * MC_LS;
        int a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3], a4 = row[4];
#pragma unroll 1
        for (int k = 0; k < w; ++k) {
          const int a5 = row[k + 5];
          sm->mc_h[l * 16 + k] = (int16_t)tap6(a0, a1, a2, a3, a4, a5);
          a0 = a1; a1 = a2; a2 = a3; a3 = a4; a4 = a5;
        }
      }
    HWB_LANES_END
  }
  // lane -> `ppl` consecutive samples of row y: 16x16: 8 samples, 2 lanes per row; 8x8: 2 and 4; 4x4: 1 and 4 (16 lanes)
  const int lg_ppl = w == 16 ? 3 : (w == 8 ? 1 : 0), lg_lpr = w == 16 ? 1 : 2;
  // (Tried and dropped: a copy-only path for integer-sample vectors -- no gain on the bench clip: 637 vs 633 ms per
  // 9000 pictures, profiles/r2_runs/r2ar_ab.txt.)
  {
    const bool use_b = fx != 0 && !need_j;                // rounded horizontal half sample, from window row rb
    const bool use_v = fy != 0 && !(need_j && fx == 2);   // rounded vertical half sample, from window column x + cv
    const int cv = 2 + ((fx == 3) ? 1 : 0);
    HWB_LANES(l)
      const int y = l >> lg_lpr, xs = (l & ((1 << lg_lpr) - 1)) << lg_ppl;
      if (y < h) {
        const int rb = y + 2 + ((fy == 3) ? 1 : 0);
        const uint8_t *row = T + rb * MC_LS + xs;
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
        if (use_b) { a0 = row[0]; a1 = row[1]; a2 = row[2]; a3 = row[3]; a4 = row[4]; }
#pragma unroll 1
        for (int i = 0; i < (1 << lg_ppl); ++i) {
          const int x = xs + i;
          int b = 0, hh = 0, v;
          if (use_b) {
            const int a5 = row[i + 5];
            b = clip8((tap6(a0, a1, a2, a3, a4, a5) + 16) >> 5);
            if (fy == 0 && fx != 2) b = (b + (fx == 3 ? a3 : a2) + 1) >> 1;  // quarter sample next to an integer sample of the same row
            a0 = a1; a1 = a2; a2 = a3; a3 = a4; a4 = a5;
          }
          if (use_v) {
            const uint8_t *col = T + y * MC_LS + x + (fx == 0 ? 2 : cv);
            hh = clip8((tap6(col[0], col[MC_LS], col[2 * MC_LS], col[3 * MC_LS], col[4 * MC_LS], col[5 * MC_LS]) + 16) >> 5);
          }
          if (fx == 0 && fy == 0) v = HWB_T(y + 2, x + 2);
          else if (fy == 0) v = b;
          else if (fx == 0) v = fy == 2 ? hh : (hh + HWB_T(rb, x + 2) + 1) >> 1;
          else if (need_j) {
            const int16_t *hr = sm->mc_h + y * 16 + x;
            const int j = clip8((tap6(hr[0], hr[16], hr[32], hr[48], hr[64], hr[80]) + 512) >> 10);
            if (fx == 2 && fy == 2) v = j;
            else if (fx == 2) v = (j + clip8((sm->mc_h[rb * 16 + x] + 16) >> 5) + 1) >> 1;
            else v = (j + hh + 1) >> 1;
          } else v = (b + hh + 1) >> 1;
          dst_y[y * ds_y + x] = (uint8_t)v;
        }
      }
    HWB_LANES_END
  }
#undef HWB_T
  // ---- chroma: bilinear eighth-sample interpolation (8.4.2.2.2), both planes at once
  {
    const int cfx = mvx & 7, cfy = mvy & 7;
    // samples per plane 8x8 / 4x4 / 2x2: lanes per plane 16 / 16 / 4, lanes per row 2 / 4 / 2, samples per lane 4 / 1 / 1
    const int lg_lpp = w == 4 ? 2 : 4, lg_clpr = w == 8 ? 2 : 1, lg_cppl = w == 16 ? 2 : 0;
    const int wa = (8 - cfx) * (8 - cfy), wb = cfx * (8 - cfy), wc2 = (8 - cfx) * cfy, wd = cfx * cfy;
    HWB_LANES(l)
      const int pl = l >> lg_lpp, k = l & ((1 << lg_lpp) - 1);
      if (pl < 2) {
        const int y = k >> lg_clpr, xs = (k & ((1 << lg_clpr) - 1)) << lg_cppl;
        const uint8_t *t0 = sm->mc_chroma[pl] + csh + y * MC_CS + xs, *t1 = t0 + MC_CS;
        uint8_t *d = (pl ? dst_cr : dst_cb) + y * ds_c + xs;
        int a = t0[0], cc = t1[0];
#pragma unroll 1
        for (int i = 0; i < (1 << lg_cppl); ++i) {
          const int b = t0[i + 1], dd = t1[i + 1];
          d[i] = (uint8_t)((wa * a + wb * b + wc2 * cc + wd * dd + 32) >> 6);
          a = b; cc = dd;
        }
      }
    HWB_LANES_END
  }
}

// All partitions of one list of an inter macroblock.  Motion is stored per 4x4 block: a macroblock whose 16 vectors and
// 4 references agree is one 16x16 partition (P_Skip, 16x16: the bulk of inter macroblocks), otherwise every 8x8
// quadrant is one partition if its four vectors agree and four 4x4 partitions if not.
HWB_FN void mc_list(const ChunkCtx &c, const SliceDesc &sd, int list, const int16_t *mv, const int8_t *ri, int mbx, int mby,
                    uint8_t *dst_y, int ds_y, uint8_t *dst_cb, uint8_t *dst_cr, int ds_c, ReconScratch *sm) {
  const uint32_t *mw = (const uint32_t *)mv;  // [16] raster 4x4, x | y << 16
  const int r0 = ri[0];
  bool uni = r0 >= 0 && ri[1] == r0 && ri[2] == r0 && ri[3] == r0;
  const uint32_t m0 = mw[0];
#if HWB_DEVICE_BUILD
  // one load per lane and a vote instead of fifteen dependent loads by every lane
  uni = __all_sync(0xffffffffu, uni && mw[threadIdx.x & 15] == m0);
#else
  if (uni) {
#pragma unroll 1
    for (int i = 1; i < 16; ++i) uni &= mw[i] == m0;
  }
#endif
  if (uni) {
    mc_partition(c, sd.ref_frame[list][r0], mbx * 16, mby * 16, 16, 16, (int16_t)(m0 & 0xffff), (int16_t)(m0 >> 16), dst_y, ds_y, dst_cb, dst_cr, ds_c, sm);
    return;
  }
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    const int r = ri[q];
    if (r < 0) continue;
    const int bx = (q & 1) * 2, by = (q >> 1) * 2, b0 = by * 4 + bx;
    const int rf = sd.ref_frame[list][r];
    const uint32_t a = mw[b0];
    if (mw[b0 + 1] == a && mw[b0 + 4] == a && mw[b0 + 5] == a) {
      mc_partition(c, rf, mbx * 16 + bx * 4, mby * 16 + by * 4, 8, 8, (int16_t)(a & 0xffff), (int16_t)(a >> 16),
                   dst_y + by * 4 * ds_y + bx * 4, ds_y, dst_cb + by * 2 * ds_c + bx * 2, dst_cr + by * 2 * ds_c + bx * 2, ds_c, sm);
    } else {
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const int x = bx + (k & 1), y = by + (k >> 1);
        const uint32_t m = mw[y * 4 + x];
        mc_partition(c, rf, mbx * 16 + x * 4, mby * 16 + y * 4, 4, 4, (int16_t)(m & 0xffff), (int16_t)(m >> 16),
                     dst_y + y * 4 * ds_y + x * 4, ds_y, dst_cb + y * 2 * ds_c + x * 2, dst_cr + y * 2 * ds_c + x * 2, ds_c, sm);
      }
    }
  }
}

// Implicit bi-prediction weights (8.4.2.3.1): returns w1 (w0 = 64 - w1).
HWB_HD int implicit_w1(int cur_poc, int poc0, int poc1, bool any_long) {
  int tb = clip3(-128, 127, cur_poc - poc0);
  int td = clip3(-128, 127, poc1 - poc0);
  if (td == 0 || any_long) return 32;
  int tx = (16384 + iabs(td / 2)) / td;
  int dsf = clip3(-1024, 1023, (tb * tx + 32) >> 6);
  int w1 = dsf >> 2;
  if (w1 < -64 || w1 > 128) return 32;
  return w1;
}

struct WeightSel {
  int mode;  // 0 default, 1 weighted
  int logwd, w0, w1, o0, o1;
};

HWB_HD int weight_uni(const WeightSel &ws, int p, int w, int o) {
  if (!ws.mode) return p;
  int v = ws.logwd >= 1 ? ((p * w + (1 << (ws.logwd - 1))) >> ws.logwd) + o : p * w + o;
  return clip8(v);
}
HWB_HD int weight_bi(const WeightSel &ws, int p0, int p1) {
  if (!ws.mode) return (p0 + p1 + 1) >> 1;
  return clip8(((p0 * ws.w0 + p1 * ws.w1 + (1 << ws.logwd)) >> (ws.logwd + 1)) + ((ws.o0 + ws.o1 + 1) >> 1));
}

// plane: 0 luma, 1 Cb, 2 Cr.  r0/r1: reference indices (or -1).
HWB_FN WeightSel select_weights(const PicDesc &pd, const SliceDesc &sd, int plane, int r0, int r1) {
  WeightSel ws;
  ws.mode = 0; ws.logwd = 0; ws.w0 = ws.w1 = 1; ws.o0 = ws.o1 = 0;
  if (sd.use_weights == 1) {
    ws.mode = 1;
    ws.logwd = plane ? sd.chroma_log2_denom : sd.luma_log2_denom;
    if (r0 >= 0) { ws.w0 = plane ? sd.chroma_w[0][r0][plane - 1] : sd.luma_w[0][r0]; ws.o0 = plane ? sd.chroma_o[0][r0][plane - 1] : sd.luma_o[0][r0]; }
    if (r1 >= 0) { ws.w1 = plane ? sd.chroma_w[1][r1][plane - 1] : sd.luma_w[1][r1]; ws.o1 = plane ? sd.chroma_o[1][r1][plane - 1] : sd.luma_o[1][r1]; }
  } else if (sd.use_weights == 2 && r0 >= 0 && r1 >= 0) {
    bool any_long = ((sd.ref_long[0] >> r0) & 1) || ((sd.ref_long[1] >> r1) & 1);
    int w1 = implicit_w1(pd.poc, sd.ref_poc[0][r0], sd.ref_poc[1][r1], any_long);
    ws.mode = 1; ws.logwd = 5; ws.w0 = 64 - w1; ws.w1 = w1; ws.o0 = ws.o1 = 0;
  }
  return ws;
}

// ---------------------------------------------------------------------------------------------
// macroblock reconstruction (one warp)
// ---------------------------------------------------------------------------------------------
struct MbAvail {
  bool left, top, topright, topleft;  // availability for intra prediction
};

HWB_HD bool nb_avail(const ChunkCtx &c, const PicDesc &pd, const MbInfo *mbs, const MbInfo &cur, int nx, int ny) {
  if (nx < 0 || ny < 0 || nx >= c.mb_w) return false;
  const MbInfo &n = mbs[ny * c.mb_w + nx];
  if (n.slice != cur.slice) return false;
  if (pd.constrained_intra_pred && n.mbtype == MB_INTER) return false;
  return true;
}

// Reconstruct macroblock (mbx,mby) of picture `pic` into its frame buffer.  `sm` persists along the
// row: on entry its column -1 (luma/chroma) holds the unfiltered right edge of the previous macroblock.
HWB_FN void recon_mb(const ChunkCtx &c, int pic, int mbx, int mby, ReconScratch *sm) {
  const PicDesc &pd = c.pics[pic];
  const int mbaddr = mby * c.mb_w + mbx;
  const MbInfo *mbs = pic_mbinfo(c, pd.frame);
  const MbInfo mb = mbs[mbaddr];
  const SliceDesc &sd = c.slices[pd.first_slice + mb.slice];
  const int16_t *coefs = pic_coefs(c, pd.frame) + (uint64_t)mb.coef_off * 16;
  uint8_t *Y = frame_y(c, pd.frame), *Cb = frame_cb(c, pd.frame), *Cr = frame_cr(c, pd.frame);
  const int wc = c.wc, cw = c.wc >> 1;
  const bool intra = mb.mbtype != MB_INTER;
  const uint32_t nz = mb.nzmask;

  if (mb.mbtype == MB_IPCM) {
    const uint8_t *s = (const uint8_t *)coefs;
    HWB_LANES(l)
#pragma unroll 1
      for (int i = l; i < 384; i += 32) {
        if (i < 256) sm->luma[((i >> 4) + 1) * LT_STRIDE + LT_OFF + (i & 15)] = s[i];
        else { int k = i - 256; sm->chroma[k >> 6][(((k & 63) >> 3) + 1) * CT_STRIDE + CT_OFF + (k & 7)] = s[i]; }
      }
    HWB_LANES_END
  } else {
    MbAvail av;
    av.left = av.top = av.topright = av.topleft = false;
    if (intra) {
      av.left = nb_avail(c, pd, mbs, mb, mbx - 1, mby);
      av.top = nb_avail(c, pd, mbs, mb, mbx, mby - 1);
      av.topright = nb_avail(c, pd, mbs, mb, mbx + 1, mby - 1);
      av.topleft = nb_avail(c, pd, mbs, mb, mbx - 1, mby - 1);
      // top border rows (written by the warp of the row above, possibly on another SM)
      HWB_LANES(l)
        if (mby > 0) {
          int col = l - 1;  // -1..30, need -1..23
          int gx = mbx * 16 + col;
          if (col <= 23 && gx >= 0 && gx < wc) sm->luma[LT_OFF + col] = ld_u8_cg(Y + (mby * 16 - 1) * wc + gx);
          if (l < 18) {
            int pl = l / 9, cc = (l % 9) - 1;
            int gcx = mbx * 8 + cc;
            if (gcx >= 0) sm->chroma[pl][CT_OFF + cc] = ld_u8_cg((pl ? Cr : Cb) + (mby * 8 - 1) * cw + gcx);
          }
        }
      HWB_LANES_END
    }
    // ---- block phase: residuals
    const bool t8 = (mb.flags & MBF_T8x8) != 0;
    if (nz == 0 && mb.mbtype != MB_I16x16) {
      // nothing coded (skipped macroblocks, and most others at the bit rates inter pictures run at): no transform phase
      // (2 % of the picture kernel's time on the bench clip, profiles/r2_runs/r2ar_ab.txt)
      HWB_LANES(l)
        if (l == 0) sm->has_res = 0;
      HWB_LANES_END
    } else {
    const int qpc0 = chroma_qp(mb.qp, pd.chroma_qp_offset[0]), qpc1 = chroma_qp(mb.qp, pd.chroma_qp_offset[1]);
    const int sl = intra ? 0 : 3;
    HWB_LANES(l)
      if (l == 0) {
        sm->has_res = 0;
        if (mb.mbtype == MB_I16x16) {
          if (nz & (1u << NZ_LUMA_DC)) luma_dc_transform(coefs, (int)pd.scaling4[0][0] * dequant4_v[(mb.qp % 6) * 16], mb.qp, sm->dc);
#pragma unroll 1
          else for (int i = 0; i < 16; ++i) sm->dc[i] = 0;
        }
      } else if (l == 1 || l == 2) {
        int pl = l - 1;
        int bit = pl ? NZ_CR_DC : NZ_CB_DC;
        int q = pl ? qpc1 : qpc0;
        if (nz & (1u << bit))
          chroma_dc_transform(coefs + 16 * popc32(nz & ((1u << bit) - 1)), (int)pd.scaling4[sl + 1 + pl][0] * dequant4_v[(q % 6) * 16], q, sm->dc + 16 + 4 * pl);
#pragma unroll 1
        else for (int i = 0; i < 4; ++i) sm->dc[16 + 4 * pl + i] = 0;
      }
    HWB_LANES_END
    HWB_LANES(l)
      uint32_t got = 0;
      if (l < 16) {
        if (!t8) {
          int bit = NZ_LUMA0 + l;
          const int16_t *cp = (nz & (1u << bit)) ? coefs + 16 * popc32(nz & ((1u << bit) - 1)) : nullptr;
          bool usedc = mb.mbtype == MB_I16x16;
          int dcv = usedc ? sm->dc[l] : 0;
          if (cp || dcv) { residual4x4(cp, usedc, dcv, pd.scaling4[sl], mb.qp, sm->res[l]); got = 1u << l; }
        } else if ((l & 3) == 0) {
          int bit = NZ_LUMA0 + l;
          if (nz & (1u << bit)) {
            residual8x8(coefs + 16 * popc32(nz & ((1u << bit) - 1)), pd.scaling8[intra ? 0 : 1], mb.qp, sm->res[l]);
            got = 0xFu << l;
          }
        }
      } else if (l < 24) {
        int pl = (l - 16) >> 2, k = (l - 16) & 3;
        int bit = (pl ? NZ_CR0 : NZ_CB0) + k;
        const int16_t *cp = (nz & (1u << bit)) ? coefs + 16 * popc32(nz & ((1u << bit) - 1)) : nullptr;
        int dcv = sm->dc[16 + 4 * pl + k];
        if (cp || dcv) { residual4x4(cp, true, dcv, pd.scaling4[sl + 1 + pl], pl ? qpc1 : qpc0, sm->res[l]); got = 1u << l; }
      }
#if HWB_DEVICE_BUILD
      got = __reduce_or_sync(0xffffffffu, got);
      if (l == 0) sm->has_res = got;
#else
      sm->has_res |= got;
#endif
    HWB_LANES_END
    }

    if (!intra) {
      // ---- inter prediction.  One list and no weights (almost every P macroblock): the partitions are predicted
      // straight into the tile; otherwise each list goes to its own buffer and the lists are combined with the weights
      // of the macroblock's 8x8 quadrants (8.4.2.3).
      const int16_t *mv0 = pic_mv(c, pd.frame, 0) + (uint64_t)mbaddr * 32;
      const int16_t *mv1 = pic_mv(c, pd.frame, 1) + (uint64_t)mbaddr * 32;
      const int8_t *ri0 = pic_refidx(c, pd.frame, 0) + (uint64_t)mbaddr * 4;
      const int8_t *ri1 = pic_refidx(c, pd.frame, 1) + (uint64_t)mbaddr * 4;
      const bool bslice = sd.slice_type == SLICE_B;
      uint8_t *tile_y = sm->luma + LT_STRIDE + LT_OFF, *tile_cb = sm->chroma[0] + CT_STRIDE + CT_OFF, *tile_cr = sm->chroma[1] + CT_STRIDE + CT_OFF;
      if (!bslice && sd.use_weights == 0) {
        mc_list(c, sd, 0, mv0, ri0, mbx, mby, tile_y, LT_STRIDE, tile_cb, tile_cr, CT_STRIDE, sm);
      } else {
        bool any0 = false, any1 = false;
        for (int q = 0; q < 4; ++q) { any0 |= ri0[q] >= 0; any1 |= bslice && ri1[q] >= 0; }
        if (any0) mc_list(c, sd, 0, mv0, ri0, mbx, mby, sm->pred[0], 16, sm->pred[0] + 256, sm->pred[0] + 320, 8, sm);
        if (any1) mc_list(c, sd, 1, mv1, ri1, mbx, mby, sm->pred[1], 16, sm->pred[1] + 256, sm->pred[1] + 320, 8, sm);
        HWB_LANES(l)
          {  // luma: 8 consecutive samples of row l>>1 (one 8x8 quadrant: one pair of references)
            const int y = l >> 1, xs = (l & 1) * 8, q = (y >> 3) * 2 + (xs >> 3);
            const int r0 = ri0[q], r1 = bslice ? ri1[q] : -1;
            const WeightSel ws = select_weights(pd, sd, 0, r0, r1);
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int p0 = sm->pred[0][y * 16 + xs + i], p1 = sm->pred[1][y * 16 + xs + i];
              const int v = (r0 >= 0 && r1 >= 0) ? weight_bi(ws, p0, p1) : (r0 >= 0 ? weight_uni(ws, p0, ws.w0, ws.o0) : weight_uni(ws, p1, ws.w1, ws.o1));
              tile_y[y * LT_STRIDE + xs + i] = (uint8_t)v;
            }
          }
          {  // chroma: plane l>>4, 4 consecutive samples of row (l&15)>>1
            const int pl = l >> 4, y = (l & 15) >> 1, xs = (l & 1) * 4, q = (y >> 2) * 2 + (xs >> 2);
            const int r0 = ri0[q], r1 = bslice ? ri1[q] : -1;
            const WeightSel ws = select_weights(pd, sd, 1 + pl, r0, r1);
            uint8_t *t = (pl ? tile_cr : tile_cb) + y * CT_STRIDE + xs;
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int p0 = sm->pred[0][256 + pl * 64 + y * 8 + xs + i], p1 = sm->pred[1][256 + pl * 64 + y * 8 + xs + i];
              const int v = (r0 >= 0 && r1 >= 0) ? weight_bi(ws, p0, p1) : (r0 >= 0 ? weight_uni(ws, p0, ws.w0, ws.o0) : weight_uni(ws, p1, ws.w1, ws.o1));
              t[i] = (uint8_t)v;
            }
          }
        HWB_LANES_END
      }
    } else {
      // ---- intra luma
      if (mb.mbtype == MB_I16x16) {
        HWB_LANES(l)
          const uint8_t *T = sm->luma + LT_OFF;               // top row, T[x], T[-1] corner
          const uint8_t *L = sm->luma + LT_STRIDE + LT_OFF - 1;  // left col, L[y*LT_STRIDE]
          int mode = mb.imode;
          int dcv = 128, a = 0, b = 0, cc = 0;
          if (mode == 2) {
            int s = 0;
            if (av.top) for (int i = 0; i < 16; ++i) s += T[i];
            if (av.left) for (int i = 0; i < 16; ++i) s += L[i * LT_STRIDE];
            dcv = (av.top && av.left) ? (s + 16) >> 5 : ((av.top || av.left) ? (s + 8) >> 4 : 128);
          } else if (mode == 3) {
            int H = 0, V = 0;
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              H += (i + 1) * (T[8 + i] - T[6 - i]);
              V += (i + 1) * (L[(8 + i) * LT_STRIDE] - (i == 7 ? T[-1] : L[(6 - i) * LT_STRIDE]));
            }
            a = 16 * (L[15 * LT_STRIDE] + T[15]); b = (5 * H + 32) >> 6; cc = (5 * V + 32) >> 6;
          }
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            int p = l * 8 + i, x = p & 15, y = p >> 4;
            int v = mode == 0 ? T[x] : mode == 1 ? L[y * LT_STRIDE] : mode == 2 ? dcv : clip8((a + b * (x - 7) + cc * (y - 7) + 16) >> 5);
            sm->luma[(y + 1) * LT_STRIDE + LT_OFF + x] = (uint8_t)v;
          }
        HWB_LANES_END
      } else {
        // Intra4x4 / Intra8x8: blocks are serially dependent; residual is added per block here.
        const int N = mb.mbtype == MB_I8x8 ? 8 : 4;
        const int nblk = N == 8 ? 4 : 16;
#pragma unroll 1
        for (int blk = 0; blk < nblk; ++blk) {
          const int bx = N == 8 ? (blk & 1) * 8 : z2x(blk) * 4, by = N == 8 ? (blk >> 1) * 8 : z2y(blk) * 4;
          const bool aL = bx > 0 || av.left, aT = by > 0 || av.top;
          const bool aD = (bx > 0 && by > 0) ? true : (bx > 0 ? av.top : (by > 0 ? av.left : av.topleft));
          bool aC;
          if (by == 0) aC = (bx + N < 16) ? av.top : av.topright;
          else if (bx + N >= 16) aC = false;
          else aC = N == 8 ? (blk == 2) : (xy2z((bx >> 2) + 1, (by >> 2) - 1) < blk);
          const int mode = mbs[mbaddr].i4modes[blk];  // from memory: a dynamic index into the register copy `mb` would push all of it to local memory
          HWB_LANES(l)
            // each lane builds the (tiny) edge array itself; 2N+... reads from shared
            uint8_t E[26];
            const uint8_t *org = sm->luma + (by + 1) * LT_STRIDE + LT_OFF + bx;  // pixel (0,0) of block
#pragma unroll 1
            for (int i = 0; i < N; ++i) E[N - 1 - i] = aL ? org[i * LT_STRIDE - 1] : 128;
#pragma unroll 1
            for (int i = 0; i < N; ++i) E[N + 1 + i] = aT ? org[-LT_STRIDE + i] : 128;
#pragma unroll 1
            for (int i = 0; i < N; ++i) E[2 * N + 1 + i] = aC ? org[-LT_STRIDE + N + i] : E[2 * N];
            E[N] = aD ? org[-LT_STRIDE - 1] : 128;
            E[3 * N + 1] = E[3 * N];
            if (N == 8) {  // reference sample filtering (8.3.2.2.1)
              uint8_t F[26];
              if (aT) {
                F[9] = aD ? (E[8] + 2 * E[9] + E[10] + 2) >> 2 : (3 * E[9] + E[10] + 2) >> 2;
#pragma unroll 1
                for (int i = 1; i < 15; ++i) F[9 + i] = (E[8 + i] + 2 * E[9 + i] + E[10 + i] + 2) >> 2;
                F[24] = (E[23] + 3 * E[24] + 2) >> 2;
              } else for (int i = 9; i < 25; ++i) F[i] = E[i];
              if (aD) {
                if (aT && aL) F[8] = (E[9] + 2 * E[8] + E[7] + 2) >> 2;
                else if (aT) F[8] = (3 * E[8] + E[9] + 2) >> 2;
                else if (aL) F[8] = (3 * E[8] + E[7] + 2) >> 2;
                else F[8] = E[8];
              } else F[8] = E[8];
              if (aL) {
                F[7] = aD ? (E[8] + 2 * E[7] + E[6] + 2) >> 2 : (3 * E[7] + E[6] + 2) >> 2;
#pragma unroll 1
                for (int i = 1; i < 7; ++i) F[7 - i] = (E[8 - i] + 2 * E[7 - i] + E[6 - i] + 2) >> 2;
                F[0] = (E[1] + 3 * E[0] + 2) >> 2;
              } else for (int i = 0; i < 8; ++i) F[i] = E[i];
              F[25] = F[24];
#pragma unroll 1
              for (int i = 0; i < 26; ++i) E[i] = F[i];
            }
            int dcv = 128;
            if (mode == 2) {
              int s = 0;
              if (aT) for (int i = 0; i < N; ++i) s += E[N + 1 + i];
              if (aL) for (int i = 0; i < N; ++i) s += E[i];
              int sh = N == 8 ? 3 : 2;
              dcv = (aT && aL) ? (s + N) >> (sh + 1) : ((aT || aL) ? (s + (N >> 1)) >> sh : 128);
            }
            // edge samples lie outside the block being written, so lanes may write in place
            const int zb = N == 8 ? blk * 4 : blk;
            const bool hr = (sm->has_res >> zb) & 1;
            if (N == 8) {
#pragma unroll 1
              for (int k = 0; k < 2; ++k) {
                int p = l * 2 + k, x = p & 7, y = p >> 3;
                int v = mode == 2 ? dcv : intra_dir_pred(mode, 8, E, x, y);
                if (hr) v = clip8(v + sm->res[zb][y * 8 + x]);
                sm->luma[(by + y + 1) * LT_STRIDE + LT_OFF + bx + x] = (uint8_t)v;
              }
            } else if (l < 16) {
              int x = l & 3, y = l >> 2;
              int v = mode == 2 ? dcv : intra_dir_pred(mode, 4, E, x, y);
              if (hr) v = clip8(v + sm->res[zb][y * 4 + x]);
              sm->luma[(by + y + 1) * LT_STRIDE + LT_OFF + bx + x] = (uint8_t)v;
            }
          HWB_LANES_END
        }
      }
      // ---- intra chroma
      HWB_LANES(l)
        const int pl = l >> 4;
        const uint8_t *T = sm->chroma[pl] + CT_OFF;
        const uint8_t *L = sm->chroma[pl] + CT_STRIDE + CT_OFF - 1;
        const int mode = mb.cmode;  // 0 DC, 1 horizontal, 2 vertical, 3 plane
        const int p0 = (l & 15) * 4, y = p0 >> 3, x0 = p0 & 7;
        int dcv = 128, a = 0, b = 0, cc = 0;
        if (mode == 0) {
          int xo = x0 & 4, yo = y & 4;
          int st = 0, sl2 = 0;
#pragma unroll 1
          for (int i = 0; i < 4; ++i) { st += T[xo + i]; sl2 += L[(yo + i) * CT_STRIDE]; }
          bool useT = av.top, useL = av.left;
          if (xo > 0 && yo == 0) { if (useT) useL = false; }
          else if (xo == 0 && yo > 0) { if (useL) useT = false; }
          if (useT && useL) dcv = (st + sl2 + 4) >> 3;
          else if (useT) dcv = (st + 2) >> 2;
          else if (useL) dcv = (sl2 + 2) >> 2;
        } else if (mode == 3) {
          int H = 0, V = 0;
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            H += (i + 1) * (T[4 + i] - T[2 - i]);
            V += (i + 1) * (L[(4 + i) * CT_STRIDE] - (i == 3 ? T[-1] : L[(2 - i) * CT_STRIDE]));
          }
          a = 16 * (L[7 * CT_STRIDE] + T[7]); b = (34 * H + 32) >> 6; cc = (34 * V + 32) >> 6;
        }
        int vals[4];
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          int x = x0 + i;
          vals[i] = mode == 0 ? dcv : mode == 1 ? L[y * CT_STRIDE] : mode == 2 ? T[x] : clip8((a + b * (x - 3) + cc * (y - 3) + 16) >> 5);
        }
#pragma unroll 1
        for (int i = 0; i < 4; ++i) sm->chroma[pl][(y + 1) * CT_STRIDE + CT_OFF + x0 + i] = (uint8_t)vals[i];
      HWB_LANES_END
    }
    // ---- add residual (luma of Intra4x4/8x8 was done per block above)
    const bool luma_done = mb.mbtype == MB_I4x4 || mb.mbtype == MB_I8x8;
    const uint32_t hr = sm->has_res;
    HWB_LANES(l)
      if (!luma_done && (hr & 0xFFFFu)) {
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          int p = l * 8 + i, x = p & 15, y = p >> 4;
          int zb = xy2z(x >> 2, y >> 2);
          if ((hr >> zb) & 1) {
            int r = t8 ? sm->res[zb & ~3][(y & 7) * 8 + (x & 7)] : sm->res[zb][(y & 3) * 4 + (x & 3)];
            uint8_t *t = sm->luma + (y + 1) * LT_STRIDE + LT_OFF + x;
            *t = (uint8_t)clip8(*t + r);
          }
        }
      }
      if (hr >> 16) {
        const int pl = l >> 4, p0 = (l & 15) * 4, y = p0 >> 3, x0 = p0 & 7;
        int cbk = 16 + pl * 4 + (y >> 2) * 2 + (x0 >> 2);
        if ((hr >> cbk) & 1) {
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            uint8_t *t = sm->chroma[pl] + (y + 1) * CT_STRIDE + CT_OFF + x0 + i;
            *t = (uint8_t)clip8(*t + sm->res[cbk][(y & 3) * 4 + ((x0 + i) & 3)]);
          }
        }
      }
    HWB_LANES_END
  }
  // ---- write back (one 16-byte store per luma row, one 8-byte store per chroma row), and keep the right edge as the
  //      next macroblock's left border
  HWB_LANES(l)
    if (l < 16) {
      const uint8_t *s = sm->luma + (l + 1) * LT_STRIDE + LT_OFF;
      uint8_t *d = Y + (uint64_t)(mby * 16 + l) * wc + mbx * 16;
#if HWB_DEVICE_BUILD
      *(uint4 *)d = *(const uint4 *)s;
#else
      memcpy(d, s, 16);
#endif
    } else {
      int pl = (l - 16) >> 3, r = l & 7;
      const uint8_t *s = sm->chroma[pl] + (r + 1) * CT_STRIDE + CT_OFF;
      uint8_t *d = (pl ? Cr : Cb) + (uint64_t)(mby * 8 + r) * cw + mbx * 8;
#if HWB_DEVICE_BUILD
      *(uint2 *)d = *(const uint2 *)s;
#else
      memcpy(d, s, 8);
#endif
    }
  HWB_LANES_END
  HWB_LANES(l)
    if (l < 16) sm->luma[(l + 1) * LT_STRIDE + LT_OFF - 1] = sm->luma[(l + 1) * LT_STRIDE + LT_OFF + 15];
    else { int pl = (l - 16) >> 3, r = l & 7; sm->chroma[pl][(r + 1) * CT_STRIDE + CT_OFF - 1] = sm->chroma[pl][(r + 1) * CT_STRIDE + CT_OFF + 7]; }
  HWB_LANES_END
}

}  // namespace hwb
