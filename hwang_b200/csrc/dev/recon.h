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
struct ReconScratch {
  // luma tile rows -1..15 (row r at (r+1)*LT_STRIDE), column c at LT_OFF + c, c in -1..23
  uint8_t luma[17 * LT_STRIDE];
  // chroma tiles rows -1..7, column c at CT_OFF + c
  uint8_t chroma[2][9 * CT_STRIDE];
  int16_t res[24][16];  // residual per 4x4 block: 16 luma (z order; 8x8 blocks use 4 slots as 64), 4 Cb, 4 Cr
  int32_t dc[24];       // dequantised DC per block (Intra16x16 luma DC, chroma DC)
  uint32_t has_res;     // bit b: res[b] is non-zero
  // Per-lane motion-compensation scratch: the 7x9 reference window (63 bytes), the intermediate column b1[7] and
  // the two prediction arrays.  They are indexed dynamically, so as thread-local arrays they lived in local memory
  // (measured: 164 M local loads/stores per launch, 1.5 TB/s of L2 traffic, long-scoreboard the second largest
  // stall); a stride of 41 words keeps the 32 lanes on different banks.
  uint32_t lane_scratch[32][41];
};
enum { LS_WIN = 0, LS_B1 = 16, LS_P0 = 24, LS_P1 = 32 };  // word offsets inside a lane's scratch

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

// Luma prediction of a 4 (wide) x 2 (high) region whose top-left integer position (already
// displaced by mv>>2) is (px,py); fx,fy = mv&3.  ref is a coded luma plane (w x h, pitch w).
HWB_FN void mc_luma_4x2(const uint8_t *ref, int w, int h, int px, int py, int fx, int fy, int *out, uint32_t *scratch) {
  uint8_t (*win)[9] = (uint8_t (*)[9])(scratch + LS_WIN);  // rows py-2..py+4, cols px-2..px+6
  int *b1 = (int *)(scratch + LS_B1);
  const bool inside = (px >= 2) && (px + 6 < w) && (py >= 2) && (py + 4 < h);
  if (fx == 0 && fy == 0) {
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
      int yy = clip3(0, h - 1, py + r);
      for (int c = 0; c < 4; ++c) out[r * 4 + c] = ld_u8_cg(ref + yy * w + clip3(0, w - 1, px + c));
    }
#if !HWB_DEVICE_BUILD
    mc_check(clip3(0, h - 1, py + 1), clip3(0, w - 1, px + 3));
#endif
    return;
  }
  if (inside) {
#if HWB_DEVICE_BUILD
    // three aligned 32-bit loads per window row instead of nine byte loads (the row is 9 bytes at any alignment;
    // the over-read of at most 3 bytes stays inside the frame buffer).  ld.global.cg: the reference picture was
    // written by other warps of the same launch (picture kernel), the non-coherent path is not allowed here
    const uint8_t *p0 = ref + (py - 2) * w + px - 2;
    const int sh = (int)(((uintptr_t)p0) & 3) * 8;  // rows are a multiple of 16 bytes apart: same alignment for every row
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const uint32_t *q = (const uint32_t *)((uintptr_t)(p0 + r * w) & ~(uintptr_t)3);
      const uint32_t w0 = __ldcg(q), w1 = __ldcg(q + 1), w2 = __ldcg(q + 2);
      const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh), x2 = w2 >> sh;
      win[r][0] = (uint8_t)x0; win[r][1] = (uint8_t)(x0 >> 8); win[r][2] = (uint8_t)(x0 >> 16); win[r][3] = (uint8_t)(x0 >> 24);
      win[r][4] = (uint8_t)x1; win[r][5] = (uint8_t)(x1 >> 8); win[r][6] = (uint8_t)(x1 >> 16); win[r][7] = (uint8_t)(x1 >> 24);
      win[r][8] = (uint8_t)x2;
    }
#else
#pragma unroll 1
    for (int r = 0; r < 7; ++r) {
      const uint8_t *p = ref + (py - 2 + r) * w + px - 2;
      for (int c = 0; c < 9; ++c) win[r][c] = p[c];
    }
    mc_check(py + 4, px + 6);
#endif
  } else {
#pragma unroll 1
    for (int r = 0; r < 7; ++r) {
      int yy = clip3(0, h - 1, py - 2 + r);
#pragma unroll
      for (int c = 0; c < 9; ++c) win[r][c] = ld_u8_cg(ref + yy * w + clip3(0, w - 1, px - 2 + c));
    }
#if !HWB_DEVICE_BUILD
    mc_check(clip3(0, h - 1, py + 4), clip3(0, w - 1, px + 6));
#endif
  }
#define HWB_H1(r, c) tap6(win[r][(c)], win[r][(c) + 1], win[r][(c) + 2], win[r][(c) + 3], win[r][(c) + 4], win[r][(c) + 5])
#define HWB_V1(r, c) tap6(win[(r)][c], win[(r) + 1][c], win[(r) + 2][c], win[(r) + 3][c], win[(r) + 4][c], win[(r) + 5][c])
  // pixel (r,c) of the region is win[r+2][c+2]
  if (fy == 0) {
#pragma unroll 1
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 4; ++c) {
        int b = clip8((HWB_H1(r + 2, c) + 16) >> 5);
        out[r * 4 + c] = fx == 2 ? b : (b + win[r + 2][c + 2 + (fx == 3)] + 1) >> 1;
      }
  } else if (fx == 0) {
#pragma unroll 1
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 4; ++c) {
        int hh = clip8((HWB_V1(r, c + 2) + 16) >> 5);
        out[r * 4 + c] = fy == 2 ? hh : (hh + win[r + 2 + (fy == 3)][c + 2] + 1) >> 1;
      }
  } else if (fx == 2 || fy == 2) {
    for (int c = 0; c < 4; ++c) {
#pragma unroll 1
      for (int r = 0; r < 7; ++r) b1[r] = HWB_H1(r, c);
#pragma unroll 1
      for (int r = 0; r < 2; ++r) {
        int j = clip8((tap6(b1[r], b1[r + 1], b1[r + 2], b1[r + 3], b1[r + 4], b1[r + 5]) + 512) >> 10);
        int v;
        if (fx == 2 && fy == 2) v = j;
        else if (fx == 2) v = (j + clip8((b1[r + 2 + (fy == 3)] + 16) >> 5) + 1) >> 1;
        else v = (j + clip8((HWB_V1(r, c + 2 + (fx == 3)) + 16) >> 5) + 1) >> 1;
        out[r * 4 + c] = v;
      }
    }
  } else {
#pragma unroll 1
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 4; ++c) {
        int b = clip8((HWB_H1(r + 2 + (fy == 3), c) + 16) >> 5);
        int hh = clip8((HWB_V1(r, c + 2 + (fx == 3)) + 16) >> 5);
        out[r * 4 + c] = (b + hh + 1) >> 1;
      }
  }
#undef HWB_H1
#undef HWB_V1
}

// Chroma prediction of a 2x2 region at chroma position (cx,cy) (block origin, before mv); mv in
// quarter luma samples = eighth chroma samples.  ref: chroma plane (w x h).
HWB_FN void mc_chroma_2x2(const uint8_t *ref, int w, int h, int cx, int cy, int mvx, int mvy, int *out) {
  int x0 = cx + (mvx >> 3), y0 = cy + (mvy >> 3);
  int fx = mvx & 7, fy = mvy & 7;
  int v[3][3];
  for (int r = 0; r < 3; ++r) {
    int yy = clip3(0, h - 1, y0 + r);
    for (int c = 0; c < 3; ++c) v[r][c] = ld_u8_cg(ref + yy * w + clip3(0, w - 1, x0 + c));
  }
#if !HWB_DEVICE_BUILD
  mc_check(2 * clip3(0, h - 1, y0 + 2) + 1, 2 * clip3(0, w - 1, x0 + 2) + 1);  // in luma units
#endif
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c)
      out[r * 2 + c] = ((8 - fx) * (8 - fy) * v[r][c] + fx * (8 - fy) * v[r][c + 1] + (8 - fx) * fy * v[r + 1][c] +
                        fx * fy * v[r + 1][c + 1] + 32) >> 6;
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

    if (!intra) {
      // ---- inter prediction straight into the tile
      const int16_t *mv0 = pic_mv(c, pd.frame, 0) + (uint64_t)mbaddr * 32;
      const int16_t *mv1 = pic_mv(c, pd.frame, 1) + (uint64_t)mbaddr * 32;
      const int8_t *ri0 = pic_refidx(c, pd.frame, 0) + (uint64_t)mbaddr * 4;
      const int8_t *ri1 = pic_refidx(c, pd.frame, 1) + (uint64_t)mbaddr * 4;
      const bool bslice = sd.slice_type == SLICE_B;
      HWB_LANES(l)
        {  // luma: raster 4x4 block l>>1, rows (l&1)*2..+1
          int br = l >> 1, bx = br & 3, by = br >> 2, q = (by >> 1) * 2 + (bx >> 1);
          int r0 = ri0[q], r1 = bslice ? ri1[q] : -1;
          uint32_t *ls = sm->lane_scratch[l];
          int *p0 = (int *)(ls + LS_P0), *p1 = (int *)(ls + LS_P1);
          int x = mbx * 16 + bx * 4, y = mby * 16 + by * 4 + (l & 1) * 2;
          if (r0 >= 0) {
            int mx = mv0[br * 2], my = mv0[br * 2 + 1];
            mc_luma_4x2(frame_y(c, sd.ref_frame[0][r0]), c.wc, c.hc, x + (mx >> 2), y + (my >> 2), mx & 3, my & 3, p0, ls);
          }
          if (r1 >= 0) {
            int mx = mv1[br * 2], my = mv1[br * 2 + 1];
            mc_luma_4x2(frame_y(c, sd.ref_frame[1][r1]), c.wc, c.hc, x + (mx >> 2), y + (my >> 2), mx & 3, my & 3, p1, ls);
          }
          WeightSel ws = select_weights(pd, sd, 0, r0, r1);
          uint8_t *t = sm->luma + (by * 4 + (l & 1) * 2 + 1) * LT_STRIDE + LT_OFF + bx * 4;
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            int v = (r0 >= 0 && r1 >= 0) ? weight_bi(ws, p0[i], p1[i]) : (r0 >= 0 ? weight_uni(ws, p0[i], ws.w0, ws.o0) : weight_uni(ws, p1[i], ws.w1, ws.o1));
            t[(i >> 2) * LT_STRIDE + (i & 3)] = (uint8_t)v;
          }
        }
        {  // chroma: plane l>>4, luma block l&15 -> 2x2 chroma samples
          int pl = l >> 4, br = l & 15, bx = br & 3, by = br >> 2, q = (by >> 1) * 2 + (bx >> 1);
          int r0 = ri0[q], r1 = bslice ? ri1[q] : -1;
          int *p0 = (int *)(sm->lane_scratch[l] + LS_P0), *p1 = (int *)(sm->lane_scratch[l] + LS_P1);
          int cx = mbx * 8 + bx * 2, cy = mby * 8 + by * 2;
          if (r0 >= 0) {
            const uint8_t *rp = pl ? frame_cr(c, sd.ref_frame[0][r0]) : frame_cb(c, sd.ref_frame[0][r0]);
            mc_chroma_2x2(rp, cw, c.hc >> 1, cx, cy, mv0[br * 2], mv0[br * 2 + 1], p0);
          }
          if (r1 >= 0) {
            const uint8_t *rp = pl ? frame_cr(c, sd.ref_frame[1][r1]) : frame_cb(c, sd.ref_frame[1][r1]);
            mc_chroma_2x2(rp, cw, c.hc >> 1, cx, cy, mv1[br * 2], mv1[br * 2 + 1], p1);
          }
          WeightSel ws = select_weights(pd, sd, 1 + pl, r0, r1);
          uint8_t *t = sm->chroma[pl] + (by * 2 + 1) * CT_STRIDE + CT_OFF + bx * 2;
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            int v = (r0 >= 0 && r1 >= 0) ? weight_bi(ws, p0[i], p1[i]) : (r0 >= 0 ? weight_uni(ws, p0[i], ws.w0, ws.o0) : weight_uni(ws, p1[i], ws.w1, ws.o1));
            t[(i >> 1) * CT_STRIDE + (i & 1)] = (uint8_t)v;
          }
        }
      HWB_LANES_END
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
  // ---- write back, and keep the right edge as the next macroblock's left border
  HWB_LANES(l)
    if (l < 16) {
      const uint32_t *s = (const uint32_t *)(sm->luma + (l + 1) * LT_STRIDE + LT_OFF);
      uint32_t *d = (uint32_t *)(Y + (uint64_t)(mby * 16 + l) * wc + mbx * 16);
      d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
    } else {
      int pl = (l - 16) >> 3, r = l & 7;
      const uint32_t *s = (const uint32_t *)(sm->chroma[pl] + (r + 1) * CT_STRIDE + CT_OFF);
      uint32_t *d = (uint32_t *)((pl ? Cr : Cb) + (uint64_t)(mby * 8 + r) * cw + mbx * 8);
      d[0] = s[0]; d[1] = s[1];
    }
  HWB_LANES_END
  HWB_LANES(l)
    if (l < 16) sm->luma[(l + 1) * LT_STRIDE + LT_OFF - 1] = sm->luma[(l + 1) * LT_STRIDE + LT_OFF + 15];
    else { int pl = (l - 16) >> 3, r = l & 7; sm->chroma[pl][(r + 1) * CT_STRIDE + CT_OFF - 1] = sm->chroma[pl][(r + 1) * CT_STRIDE + CT_OFF + 7]; }
  HWB_LANES_END
}

}  // namespace hwb
