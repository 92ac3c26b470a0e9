// Slice-data entropy decoding (H.264 clauses 7.3.4-7.3.5, 9.1 Exp-Golomb, 9.2 CAVLC, 9.3 CABAC)
// and motion-vector derivation (8.4.1, including P_Skip and B direct prediction).
// One warp owns one slice and walks its macroblocks in raster order (the serial syntax decoding runs identically on
// all 32 lanes, bulk data movement is spread over the lanes: see "warp-cooperative execution model" below); slices of
// every picture of the chunk are decoded concurrently (entropy decoding needs no pixels).  Output: MbInfo records,
// coefficient slots, final motion vectors / reference indices per 4x4 / 8x8 block.
// The file may be included several times with different (HWB_ENT_NS, HWB_ENT_MODE) pairs: the CUDA side compiles a
// CABAC-only and a CAVLC-only copy of the slice decoder (each a fraction of the code, which matters because this kernel
// is instruction-fetch bound) next to the generic one; host builds use the generic runtime-flag version only.
#include "bits.h"
#include "ir.h"
#include "tables_gen.h"

#ifndef HWB_ENT_NS
#define HWB_ENT_NS ent
#define HWB_ENT_MODE (-1)
#endif
// HWB_ENT_NO_B: a copy of the slice decoder for chunks without B slices (no second list, no direct prediction): the
// per-macroblock path is what the instruction caches have to hold, and B support is a fifth of it.
#undef HWB_IS_B
#ifdef HWB_ENT_NO_B
#define HWB_IS_B(slice_type) false
#else
#define HWB_IS_B(slice_type) ((slice_type) == SLICE_B)
#endif
// HWB_ENT_NO_T8: a copy for batches whose pictures all have transform_8x8_mode_flag = 0 (Baseline / Main streams): no
// 8x8 residual category, no Intra8x8, no transform_size_8x8_flag -- less code on the per-macroblock path.
#undef HWB_T8_ON
#ifdef HWB_ENT_NO_T8
#define HWB_T8_ON false
#else
#define HWB_T8_ON true
#endif
#undef HWB_IS_CABAC
#if HWB_ENT_MODE == 1
#define HWB_IS_CABAC(s) true
#elif HWB_ENT_MODE == 0
#define HWB_IS_CABAC(s) false
#else
#define HWB_IS_CABAC(s) ((s).cabac)
#endif

namespace hwb {
namespace HWB_ENT_NS {

enum { NBF_INTRA = 1, NBF_IPCM = 2, NBF_SKIP = 4, NBF_DIRECT16 = 8, NBF_T8 = 16, NBF_I16 = 32, NBF_INXN = 64 };
enum { REF_UNAVAIL = -2, REF_NONE = -1 };

// What the macroblocks below (below-left / below-right) need to know about a decoded macroblock: its bottom
// edge.  The right edge never leaves the caches: the next macroblock of the row shifts it into its left column.
struct alignas(16) NbCtx {
  uint8_t flags, cbp, cmode, dirmask;  // dirmask bit x: bottom-row block x was predicted in direct mode
  uint32_t cbf;                        // nzmask-style coded flags (I_PCM: all ones)
  uint8_t cnnz_b[2][2];
  uint8_t pad[4];
  uint8_t nnz_b[4];
  int8_t imode_b[4];  // effective Intra4x4PredMode for the neighbour derivation (2 = not I_NxN, -1 = unusable)
  int8_t ref_b[2][4];
  int16_t mv_b[2][4][2];
  uint8_t mvd_b[2][4][2];
};
static_assert(sizeof(NbCtx) == 80, "NbCtx layout");
struct LeftCtx { uint8_t flags, cbp, cmode, pad; uint32_t cbf; };

// Neighbour caches: 5 rows x 8 columns per array (row 0 = blocks above, columns 4..7 = the macroblock, column 3 =
// left neighbours, column 8 == next row's column 0 = "top-right of the last column", permanently unavailable
// except for row 0 where it holds the top-right macroblock's block).  Rows of 4 blocks are word / 16-byte aligned.
#define HWB_CI(bx, by) (((by) + 1) * 8 + (bx) + 4)
enum { HWB_CACHE_N = 48 };

struct SliceDec {
  const ChunkCtx *c;
  const PicDesc *pd;
  const SliceDesc *sd;
  int slice_num;  // inside picture
  BitReader br;
  Cabac cab;
  bool cabac;
  uint32_t stop_bitpos;
  int qp;
  int last_dqp;
  int pre_mbt;  // CABAC P slices: mb_type as decoded together with mb_skip_flag (cabac_p_header), -1 = not decoded yet
  // macroblock loop state (slice_begin / slice_step / slice_end)
  int it_slice, it_first, it_addr, it_end_mb, it_run;
  bool it_end;
  LeftCtx left;
  uint32_t top_words[4];   // words 0..3 of the top neighbour's line-buffer entry (flags/cbp/cmode/dirmask, cbf, chroma nnz)
  int8_t tl_ref[2];        // top-left macroblock's bottom-right block (saved before its line entry is overwritten)
  alignas(4) int16_t tl_mv[2][2];  // copied as 32-bit words
  NbCtx *line;  // [mb_w] top context
  uint32_t coef_next;  // next free slot in the picture arena
  int32_t row_reach;   // 1 + lowest reference macroblock row read by the inter macroblocks of the current row so far (0 = none)
  int32_t row_reach_x; // reference macroblock columns needed beyond a macroblock's own column (see ChunkCtx::mv_reach_x)
  // ---- current macroblock
  int mbx, mby, mbaddr;
  bool availA, availB, availC, availD;
  alignas(16) int16_t mv_cache[2][HWB_CACHE_N][2];
  alignas(16) uint8_t mvd_cache[2][HWB_CACHE_N][2];
  alignas(16) int8_t ref_cache[2][HWB_CACHE_N];
  alignas(16) uint8_t dir_cache[HWB_CACHE_N];
  alignas(16) uint8_t nz_cache[HWB_CACHE_N];  // luma total_coeff / coded flag; 0x80 = unavailable
  alignas(16) int8_t im_cache[HWB_CACHE_N];
  uint8_t cnz_cache[2][12];  // chroma 3x4 layouts: (by+1)*4 + bx+1
  alignas(16) int16_t coef[64];  // staging for one block
  MbInfo out;
  int error;
  // Per-macroblock scratch.  It lives here (not on the stack) because on the GPU this struct is placed in
  // shared memory: with one active lane per warp, thread-local memory (interleaved across the 32 lanes)
  // touches a different cache line per word and thrashes L1.
  int8_t dref[2][4];
  int16_t dmv[2][16][2];
  int16_t level[16];
  uint8_t index[64];
  int8_t refs[2][4];
  int8_t sub[4], shape[4], pf[4];
  uint16_t grp[4];  // inter macroblock: reference groups (see decode_mb)
  // output arrays of this slice's picture (set by init_caches: saves the address arithmetic per macroblock)
  MbInfo *o_mbinfo;
  int16_t *o_mv[2];
  int8_t *o_refidx[2];
  int16_t *o_refpic[2];
  int16_t *o_coefs;
  // CABAC contexts (one fused-table entry each, see bits.h CtxE) and a copy of the 128-entry fused table live inside
  // the slice state (shared memory on the GPU): the decoder addresses both as constant offsets from the one pointer
  // it holds anyway
  alignas(16) CtxE ctxe[464];
  alignas(16) CtxE fused[128];
  // per-lane move tables of the cache maintenance (fill_tab_a/b, top_dflt, line_tab, fill_dflt_a below), copied here once
  // per slice: read as constant offsets from the slice state instead of through 64-bit global addresses
  uint32_t t_a[32], t_b[28], t_topd[20], t_line[20];
  uint8_t t_da[24];
};

HWB_HD void sd_fail(SliceDec &s, int code) { if (!s.error) s.error = code; }
HWB_HD uint32_t top_flags(const SliceDec &s) { return s.top_words[0] & 0xff; }  // valid when availB

// Engine access for the macroblock-layer syntax.  A syntax element loads the engine into registers once (HWB_CAB_ENTER),
// decodes its bins (HWB_BIN / HWB_BYP) and stores it back (HWB_CAB_LEAVE); single flags go through cabac_bin.
#define HWB_CAB_ENTER(s) CabReg cab = cab_enter((s).cab); const CtxE *const ft = (s).fused
#define HWB_CAB_LEAVE(s) cab_leave((s).cab, cab)
#define HWB_BIN_INL(s, ctx) cabac_decision(cab, (s).cab, (s).ctxe + (ctx), ft)
#define HWB_BYP(s) cabac_bypass(cab, (s).cab)
// The decision as a call that takes and returns the engine in registers: the macroblock-layer syntax has some twenty
// decision sites; inlined (30 instructions each) they were 9 KB of a per-macroblock path that has to fit a 32 KB
// instruction cache.  Only the residual loops inline the decision.
#ifndef HWB_BIN_OOL
#define HWB_BIN_OOL 1
#endif
struct BinRet { uint32_t low, range; int32_t nb; int bin; };
HWB_FN BinRet cabac_bin_reg(uint32_t low, uint32_t range, int32_t nb, SliceDec *s, int ctx) {
  CabReg c; c.low = low; c.range = range; c.nb = nb;
  BinRet r;
  r.bin = cabac_decision(c, s->cab, s->ctxe + ctx, s->fused);
  r.low = c.low; r.range = c.range; r.nb = c.nb;
  return r;
}
HWB_HD int cabac_bin_call(CabReg &c, SliceDec &s, int ctx) {
  const BinRet r = cabac_bin_reg(c.low, c.range, c.nb, &s, ctx);
  c.low = r.low; c.range = r.range; c.nb = r.nb;
  return r.bin;
}
#if HWB_BIN_OOL
#define HWB_BIN(s, ctx) cabac_bin_call(cab, s, ctx)
#else
#define HWB_BIN(s, ctx) HWB_BIN_INL(s, ctx)
#endif
HWB_FN int cabac_bin(SliceDec &s, int ctx) { HWB_CAB_ENTER(s); const int b = HWB_BIN_INL(s, ctx); HWB_CAB_LEAVE(s); return b; }
HWB_FN int cabac_term(SliceDec &s) { CabReg cab = cab_enter(s.cab); const int b = cabac_terminate(cab, s.cab); cab_leave(s.cab, cab); return b; }
HWB_FN uint32_t s_ue(SliceDec &s) { return br_ue(s.br); }
HWB_FN int32_t s_se(SliceDec &s) { return br_se(s.br); }
HWB_FN uint32_t s_get(SliceDec &s, int n) { return br_get(s.br, n); }

// ================================================================================ neighbour caches
HWB_HD void cpy4(void *d, const void *s) { *(uint32_t *)d = *(const uint32_t *)s; }
HWB_HD void set4(void *d, uint32_t v) { *(uint32_t *)d = v; }
HWB_HD void cpy8(void *d, const void *s) { *(uint64_t *)d = *(const uint64_t *)s; }
HWB_HD void cpy16(void *d, const void *s) {
#if HWB_DEVICE_BUILD
  *(uint4 *)d = *(const uint4 *)s;
#else
  memcpy(d, s, 16);
#endif
}

// ---- warp-cooperative execution model of this file
// All 32 lanes of the warp that owns a slice execute the slice decoder with identical values ("uniform" code: every
// lane reads the same shared/global addresses and writes the same values, which costs exactly what one lane would).
// Bulk data movement (neighbour caches, line buffer, outputs) is done in HWB_LANES blocks where lane l moves item l;
// a block ends with __syncwarp().  Rules: inside a block a lane touches only its own items; uniform code may follow
// a block immediately (the barrier orders it), and a block may follow uniform code immediately (every lane has
// itself written whatever uniform code wrote).  Host builds run the lanes as a loop.

// Move tables for the per-macroblock cache maintenance: lane l moves item l.  Entries are byte offsets into SliceDec,
// destination << 16 | source.  Table-driven because the hand-written version was several KB of divergent address
// arithmetic in a kernel that is bound by instruction fetch.
#define HWB_O(member) ((uint32_t)offsetof(SliceDec, member))
#define HWB_MV(dst, src) (((dst) << 16) | (src))
#define HWB_LEFT_D(y) (((y) + 1) * 8 + 3)   /* HWB_CI(-1, y) */
#define HWB_LEFT_F(y) (((y) + 1) * 8 + 7)   /* HWB_CI(3, y)  */
#define HWB_ROWS4(base, elem) \
  HWB_MV((base) + HWB_LEFT_D(0) * (elem), (base) + HWB_LEFT_F(0) * (elem)), HWB_MV((base) + HWB_LEFT_D(1) * (elem), (base) + HWB_LEFT_F(1) * (elem)), \
  HWB_MV((base) + HWB_LEFT_D(2) * (elem), (base) + HWB_LEFT_F(2) * (elem)), HWB_MV((base) + HWB_LEFT_D(3) * (elem), (base) + HWB_LEFT_F(3) * (elem))
// Called once per slice before the first macroblock: entries that never change.
HWB_FN void init_caches(SliceDec &s) {
  {
    const ChunkCtx &c = *s.c;
    const int f = s.pd->frame;
    s.o_mbinfo = pic_mbinfo(c, f); s.o_coefs = pic_coefs(c, f);
    for (int k = 0; k < 2; ++k) { s.o_mv[k] = pic_mv(c, f, k); s.o_refidx[k] = pic_refidx(c, f, k); s.o_refpic[k] = pic_refpic(c, f, k); }
  }
  HWB_LANES(l)
  for (int i = l; i < HWB_CACHE_N; i += 32) {
    for (int k = 0; k < 2; ++k) { s.ref_cache[k][i] = REF_UNAVAIL; set4(s.mv_cache[k][i], 0); s.mvd_cache[k][i][0] = s.mvd_cache[k][i][1] = 0; }
    s.nz_cache[i] = 0x80; s.im_cache[i] = -1; s.dir_cache[i] = 0;
  }
  if (l < 24) s.cnz_cache[l / 12][l % 12] = 0x80;
  HWB_LANES_END
}

// Left column, pass A: lanes 0..19 byte items (nz, intra mode, direct flag, ref list 0 / 1, rows 0..3), lanes 20..23
// chroma nnz bytes, lanes 24..31 mvd pairs (16 bits) of list 0 / 1.
HWB_TABLE uint32_t fill_tab_a[32] = {
  HWB_ROWS4(HWB_O(nz_cache), 1), HWB_ROWS4(HWB_O(im_cache), 1), HWB_ROWS4(HWB_O(dir_cache), 1),
  HWB_ROWS4(HWB_O(ref_cache[0]), 1), HWB_ROWS4(HWB_O(ref_cache[1]), 1),
  HWB_MV(HWB_O(cnz_cache[0]) + 4, HWB_O(cnz_cache[0]) + 6), HWB_MV(HWB_O(cnz_cache[0]) + 8, HWB_O(cnz_cache[0]) + 10),
  HWB_MV(HWB_O(cnz_cache[1]) + 4, HWB_O(cnz_cache[1]) + 6), HWB_MV(HWB_O(cnz_cache[1]) + 8, HWB_O(cnz_cache[1]) + 10),
  HWB_ROWS4(HWB_O(mvd_cache[0]), 2), HWB_ROWS4(HWB_O(mvd_cache[1]), 2)};
HWB_TABLE uint8_t fill_dflt_a[24] = {0x80, 0x80, 0x80, 0x80, 0xFF, 0xFF, 0xFF, 0xFF, 0, 0, 0, 0, 0xFE, 0xFE, 0xFE, 0xFE, 0xFE, 0xFE, 0xFE, 0xFE,
                                     0x80, 0x80, 0x80, 0x80};
// Pass B: lanes 0..19 = destination of word l of the top neighbour's line-buffer entry (words 0..3 -> s.top_words),
// lanes 20..27 = mv words of the left column, list 0 / 1.
#define HWB_TOPROW 4 /* HWB_CI(0, -1) */
#define HWB_BOTROW 36 /* HWB_CI(0, 3) */
#define HWB_LINE_WORDS(row) \
  HWB_O(top_words[0]), HWB_O(top_words[1]), HWB_O(top_words[2]), HWB_O(top_words[3]), HWB_O(nz_cache) + (row), HWB_O(im_cache) + (row), \
  HWB_O(ref_cache[0]) + (row), HWB_O(ref_cache[1]) + (row), \
  HWB_O(mv_cache[0]) + 4 * (row), HWB_O(mv_cache[0]) + 4 * (row) + 4, HWB_O(mv_cache[0]) + 4 * (row) + 8, HWB_O(mv_cache[0]) + 4 * (row) + 12, \
  HWB_O(mv_cache[1]) + 4 * (row), HWB_O(mv_cache[1]) + 4 * (row) + 4, HWB_O(mv_cache[1]) + 4 * (row) + 8, HWB_O(mv_cache[1]) + 4 * (row) + 12, \
  HWB_O(mvd_cache[0]) + 2 * (row), HWB_O(mvd_cache[0]) + 2 * (row) + 4, HWB_O(mvd_cache[1]) + 2 * (row), HWB_O(mvd_cache[1]) + 2 * (row) + 4
HWB_TABLE uint32_t fill_tab_b[28] = {HWB_LINE_WORDS(HWB_TOPROW), HWB_ROWS4(HWB_O(mv_cache[0]), 4), HWB_ROWS4(HWB_O(mv_cache[1]), 4)};
// what an unavailable top neighbour looks like, word by word
HWB_TABLE uint32_t top_dflt[20] = {0, 0, 0x80808080u, 0, 0x80808080u, 0xFFFFFFFFu, 0xFEFEFEFEu, 0xFEFEFEFEu, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
// finish_mb: source of word l of the line-buffer entry (the bottom row of the caches; words 0..3 staged in s.top_words)
HWB_TABLE uint32_t line_tab[20] = {HWB_LINE_WORDS(HWB_BOTROW)};
// Once per slice, with init_caches: the tables' copies inside the slice state.
HWB_HD void init_lane_tables(SliceDec &s) {
  HWB_LANES(l)
  s.t_a[l] = fill_tab_a[l];
  if (l < 28) s.t_b[l] = fill_tab_b[l];
  if (l < 20) { s.t_topd[l] = top_dflt[l]; s.t_line[l] = line_tab[l]; }
  if (l < 24) s.t_da[l] = fill_dflt_a[l];
  HWB_LANES_END
}

// Per macroblock: the left column is the previous macroblock's right column (still in the caches), the top row
// comes from the slice's line buffer, the interior is reset.
HWB_FN void fill_caches(SliceDec &s, bool unused) {
  (void)unused;
  const int nl = HWB_IS_B(s.sd->slice_type) ? 2 : (s.sd->slice_type == SLICE_P ? 1 : 0);
  const int top = HWB_CI(0, -1);
  const bool availA = s.availA, availB = s.availB, availC = s.availC, availD = s.availD;
  const NbCtx *T = s.line + s.mbx;
  uint8_t *sb = (uint8_t *)&s;
  // ---- phase 1: left column, top row (one line-buffer word per lane), corners
  HWB_LANES(l)
  {
    const uint32_t e = s.t_a[l];
    uint8_t *d = sb + (e >> 16);
    const uint8_t *f = sb + (e & 0xffff);
    if (l < 24) *d = availA ? *f : s.t_da[l];
    else *(uint16_t *)d = availA ? *(const uint16_t *)f : (uint16_t)0;
  }
  if (l < 20) {
    *(uint32_t *)(sb + s.t_b[l]) = availB ? ((const uint32_t *)T)[l] : s.t_topd[l];
  } else if (l < 28) {
    const uint32_t e = s.t_b[l];
    *(uint32_t *)(sb + (e >> 16)) = availA ? *(const uint32_t *)(sb + (e & 0xffff)) : 0u;
  } else {
    // corners: lanes 28,29 = top-right of list 0,1; lanes 30,31 = top-left
    const int k = l & 1;
    if (k < nl) {
      if (l < 30) {
        const int tr = HWB_CI(4, -1);
        if (availC) { const NbCtx &R = T[1]; s.ref_cache[k][tr] = R.ref_b[k][0]; cpy4(s.mv_cache[k][tr], R.mv_b[k][0]); }
        else { s.ref_cache[k][tr] = REF_UNAVAIL; set4(s.mv_cache[k][tr], 0); }
      } else {
        const int tl = HWB_CI(-1, -1);
        if (availD) { s.ref_cache[k][tl] = s.tl_ref[k]; cpy4(s.mv_cache[k][tl], s.tl_mv[k]); }
        else { s.ref_cache[k][tl] = REF_UNAVAIL; set4(s.mv_cache[k][tl], 0); }
      }
    }
  }
  HWB_LANES_END
  // ---- phase 2: interior (nothing coded yet, nothing direct, references "not decoded yet") and what derives from s.top
  HWB_LANES(l)
  if (l < 4) {
    const int r = HWB_CI(0, l);
    set4(s.nz_cache + r, 0); set4(s.dir_cache + r, 0);
    if (nl > 0) set4(s.ref_cache[0] + r, 0xFEFEFEFEu);
    if (nl > 1) set4(s.ref_cache[1] + r, 0xFEFEFEFEu);
  } else if (l < 8) {
    const int pl = (l - 4) >> 1, r = (l - 4) & 1;
    s.cnz_cache[pl][5 + 4 * r] = 0; s.cnz_cache[pl][6 + 4 * r] = 0;
    // top chroma nnz of plane pl, column r: byte (2 * pl + r) of s.top word 2
    s.cnz_cache[pl][1 + r] = (uint8_t)(s.top_words[2] >> (8 * (2 * pl + r)));
  } else if (l == 8) {
    const uint32_t dm = s.top_words[0] >> 24;  // dirmask
    set4(s.dir_cache + top, (dm & 1) | ((dm & 2) << 7) | ((dm & 4) << 14) | ((dm & 8) << 21));
  }
  HWB_LANES_END
}

// ================================================================================ MV prediction
struct MvRef { int ref; int mx, my; };
HWB_HD MvRef mv_at(const SliceDec &s, int l, int bx, int by) {
  const int ci = HWB_CI(bx, by);
  const uint32_t w = *(const uint32_t *)s.mv_cache[l][ci];
  MvRef r; r.ref = s.ref_cache[l][ci]; r.mx = (int16_t)(w & 0xffff); r.my = (int16_t)(w >> 16);
  return r;
}
// Median / directional prediction for the partition whose top-left 4x4 block is (bx,by), width w
// (in 4x4 units).  shape: 0 = general (median), 1 = 16x8 upper, 2 = 16x8 lower, 3 = 8x16 left, 4 = 8x16 right.
HWB_FN void pred_mv(const SliceDec &s, int l, int bx, int by, int w, int ref, int shape, int &px, int &py) {
  const MvRef A = mv_at(s, l, bx - 1, by), Bn = mv_at(s, l, bx, by - 1);
  MvRef C = mv_at(s, l, bx + w, by - 1);
  if (C.ref == REF_UNAVAIL) C = mv_at(s, l, bx - 1, by - 1);
  const bool ea = A.ref == ref, eb = Bn.ref == ref, ec = C.ref == ref;
  // directional rules first, then "exactly one neighbour uses this reference", then the median
  int pick = -1;  // 0 A, 1 B, 2 C
  if (shape == 1 && eb) pick = 1;
  else if ((shape == 2 || shape == 3) && ea) pick = 0;
  else if (shape == 4 && ec) pick = 2;
  else {
    const int n = (int)ea + (int)eb + (int)ec;
    if (n == 1) pick = ea ? 0 : (eb ? 1 : 2);
    else if (n == 0 && Bn.ref == REF_UNAVAIL && C.ref == REF_UNAVAIL && A.ref != REF_UNAVAIL) pick = 0;
  }
  if (pick < 0) { px = median3(A.mx, Bn.mx, C.mx); py = median3(A.my, Bn.my, C.my); }
  else { px = pick == 0 ? A.mx : (pick == 1 ? Bn.mx : C.mx); py = pick == 0 ? A.my : (pick == 1 ? Bn.my : C.my); }
}
HWB_FN void set_motion(SliceDec &s, int l, int bx, int by, int w, int h, int ref, int mx, int my, int amvdx, int amvdy) {
  const uint32_t mv = (uint32_t)(uint16_t)mx | ((uint32_t)(uint16_t)my << 16);
  const uint16_t mvd = (uint16_t)((amvdx & 0xff) | ((amvdy & 0xff) << 8));
  HWB_LANES(i)
  const int x = i & 3, y = (i >> 2) & 3;
  if (i < 16 && x >= bx && x < bx + w && y >= by && y < by + h) {
    const int ci = HWB_CI(x, y);
    s.ref_cache[l][ci] = (int8_t)ref;
    set4(s.mv_cache[l][ci], mv);
    *(uint16_t *)s.mvd_cache[l][ci] = mvd;
  }
  HWB_LANES_END
}
// references only (made visible for the ref_idx contexts before the motion vectors are known)
HWB_FN void set_refs(SliceDec &s, int l, int bx, int by, int w, int h, int ref) {
  HWB_LANES(i)
  const int x = i & 3, y = (i >> 2) & 3;
  if (i < 16 && x >= bx && x < bx + w && y >= by && y < by + h) s.ref_cache[l][HWB_CI(x, y)] = (int8_t)ref;
  HWB_LANES_END
}

// ================================================================================ B direct prediction
// Block until macroblock `mbaddr` of picture `colpic` has final motion data (its slice has
// published progress beyond it).  Slices take tickets in decode order, so the producer is always
// resident or finished: no deadlock.
HWB_HD void wait_col_mb(const SliceDec &s, int colpic, int mbaddr) {
#if HWB_DEVICE_BUILD
  const PicDesc &cp = s.c->pics[colpic];
  int k = cp.first_slice;
  for (int i = 1; i < cp.num_slices; ++i) if (s.c->slices[cp.first_slice + i].first_mb <= mbaddr) k = cp.first_slice + i;
  volatile int32_t *p = s.c->entropy_prog + k;
  if ((threadIdx.x & 31) == 0) {
    while (*p <= mbaddr) { __nanosleep(200); }
    __threadfence();
  }
  __syncwarp();
#else
  (void)s; (void)colpic; (void)mbaddr;
#endif
}
HWB_HD int minpos(int a, int b) { return (a >= 0 && b >= 0) ? (a < b ? a : b) : (a > b ? a : b); }

// Fills dref[2][4] / dmv[2][16][2] (raster 4x4) for the quadrants in `qmask`.
HWB_FN void direct_predict(SliceDec &s, int qmask, int8_t dref[2][4], int16_t dmv[2][16][2]) {
  const ChunkCtx &c = *s.c;
  const SliceDesc &sd = *s.sd;
  const int colf = sd.ref_frame[1][0];
  wait_col_mb(s, colf, s.mbaddr);
  const bool col_intra = ld_u8_cg(&pic_mbinfo(c, colf)[s.mbaddr].mbtype) != MB_INTER;
  const bool col_has_l1 = c.pics[colf].has_inter == 2;
  const int8_t *cr0 = pic_refidx(c, colf, 0) + (uint64_t)s.mbaddr * 4, *cr1 = pic_refidx(c, colf, 1) + (uint64_t)s.mbaddr * 4;
  const int16_t *cp0 = pic_refpic(c, colf, 0) + (uint64_t)s.mbaddr * 4, *cp1 = pic_refpic(c, colf, 1) + (uint64_t)s.mbaddr * 4;
  const int16_t *cm0 = pic_mv(c, colf, 0) + (uint64_t)s.mbaddr * 32, *cm1 = pic_mv(c, colf, 1) + (uint64_t)s.mbaddr * 32;
  const bool inf8 = s.pd->direct_8x8_inference != 0;
  int ref[2] = {0, 0}, pmx[2] = {0, 0}, pmy[2] = {0, 0};
  if (sd.direct_spatial) {
    for (int l = 0; l < 2; ++l) {
      MvRef A = mv_at(s, l, -1, 0), Bn = mv_at(s, l, 0, -1), C = mv_at(s, l, 4, -1);
      if (C.ref == REF_UNAVAIL) C = mv_at(s, l, -1, -1);
      // MinPositive chain; unavailable (-2) and unused (-1) both count as negative
      int r = minpos(A.ref, minpos(Bn.ref, C.ref));
      ref[l] = r < 0 ? -1 : r;
    }
    if (ref[0] < 0 && ref[1] < 0) { ref[0] = ref[1] = 0; }
    else {
      for (int l = 0; l < 2; ++l)
        if (ref[l] >= 0) pred_mv(s, l, 0, 0, 4, ref[l], 0, pmx[l], pmy[l]);
    }
  }
  const bool l1_short = !((sd.ref_long[1] >> 0) & 1);
  for (int q = 0; q < 4; ++q) {
    if (!((qmask >> q) & 1)) continue;
    for (int k = 0; k < 4; ++k) {
      int bx = (q & 1) * 2 + (k & 1), by = (q >> 1) * 2 + (k >> 1);
      int cbx = bx, cby = by;
      if (inf8) { cbx = (q & 1) * 3; cby = (q >> 1) * 3; }
      int cq = (cby >> 1) * 2 + (cbx >> 1), cbr = cby * 4 + cbx;
      int refcol = -1, colpicref = -1, cmx = 0, cmy = 0;
      if (!col_intra) {
        int rc0 = ld_i8_cg(cr0 + cq);
        if (rc0 >= 0) { refcol = rc0; colpicref = ld_i16_cg(cp0 + cq); cmx = ld_i16_cg(cm0 + cbr * 2); cmy = ld_i16_cg(cm0 + cbr * 2 + 1); }
        else if (col_has_l1) { refcol = ld_i8_cg(cr1 + cq); colpicref = ld_i16_cg(cp1 + cq); cmx = ld_i16_cg(cm1 + cbr * 2); cmy = ld_i16_cg(cm1 + cbr * 2 + 1); }
      }
      int br = by * 4 + bx;
      if (sd.direct_spatial) {
        bool colzero = l1_short && refcol == 0 && cmx >= -1 && cmx <= 1 && cmy >= -1 && cmy <= 1;
        for (int l = 0; l < 2; ++l) {
          dref[l][q] = (int8_t)ref[l];
          bool z = ref[l] < 0 || (ref[l] == 0 && colzero);
          dmv[l][br][0] = (int16_t)(z ? 0 : pmx[l]); dmv[l][br][1] = (int16_t)(z ? 0 : pmy[l]);
        }
      } else {
        int r0 = 0;
        if (refcol >= 0) {
          r0 = 0;
          for (int i = 0; i < sd.num_ref[0]; ++i) if (sd.ref_frame[0][i] == colpicref) { r0 = i; break; }
        }
        dref[0][q] = (int8_t)r0; dref[1][q] = 0;
        int poc0 = sd.ref_poc[0][r0], poc1 = sd.ref_poc[1][0];
        int tb = clip3(-128, 127, s.pd->poc - poc0), td = clip3(-128, 127, poc1 - poc0);
        if (((sd.ref_long[0] >> r0) & 1) || td == 0) {
          dmv[0][br][0] = (int16_t)cmx; dmv[0][br][1] = (int16_t)cmy; dmv[1][br][0] = dmv[1][br][1] = 0;
        } else {
          int tx = (16384 + iabs(td / 2)) / td;
          int dsf = clip3(-1024, 1023, (tb * tx + 32) >> 6);
          int mx0 = (dsf * cmx + 128) >> 8, my0 = (dsf * cmy + 128) >> 8;
          dmv[0][br][0] = (int16_t)mx0; dmv[0][br][1] = (int16_t)my0;
          dmv[1][br][0] = (int16_t)(mx0 - cmx); dmv[1][br][1] = (int16_t)(my0 - cmy);
        }
      }
    }
  }
}

HWB_FN void apply_direct(SliceDec &s, int l, int q, const int8_t dref[2][4], const int16_t dmv[2][16][2]) {
  HWB_LANES(k)
  if (k < 4) {
    const int bx = (q & 1) * 2 + (k & 1), by = (q >> 1) * 2 + (k >> 1), br = by * 4 + bx, ci = HWB_CI(bx, by);
    s.ref_cache[l][ci] = dref[l][q];
    s.mv_cache[l][ci][0] = dmv[l][br][0]; s.mv_cache[l][ci][1] = dmv[l][br][1];
    *(uint16_t *)s.mvd_cache[l][ci] = 0;
    s.dir_cache[ci] = 1;
  }
  HWB_LANES_END
}

// ================================================================================ CAVLC residual
HWB_HD int cavlc_nc(int na, int nb) {
  bool a = na != 0x80, b = nb != 0x80;
  if (a && b) return (na + nb + 1) >> 1;
  return a ? na : (b ? nb : 0);
}

// Decodes one residual block.  Levels are written to s.coef (zeroed first) at dezigzagged positions
// given by `scan` (scan[start+i]); for 8x8-interleaved blocks scan8 != nullptr: position =
// scan8[4*i + sub].  Returns total_coeff.
HWB_HD int cavlc_residual_impl(SliceDec &s, BitReader &b, int nC, int max_coeff, int start, const uint8_t *scan, const uint8_t *scan8, int sub) {
  int total, t1;
  int16_t *level = s.level;
  if (nC < 0) {
    uint32_t e = cavlc_chroma_dc_token[br_peek(b, 8)];
    if (!e) { sd_fail(s, 10); return 0; }
    br_skip(b, e >> 8); total = (e & 255) >> 2; t1 = e & 3;
  } else if (nC >= 8) {
    uint32_t e = cavlc_coeff_token_flc[br_peek(b, 6)];
    if (!e) { sd_fail(s, 11); return 0; }
    br_skip(b, 6); total = (e & 255) >> 2; t1 = e & 3;
  } else {
    int tab = nC < 2 ? 0 : (nC < 4 ? 1 : 2);
    uint32_t v = br_peek(b, 16);
    int lz = clz32(v) - 16;
    if (lz >= 16) { sd_fail(s, 12); return 0; }
    uint32_t suf = (v >> (12 - lz)) & 7;  // the 3 bits after the leading 1 (16-bit window, lz<=12 ok)
    if (lz > 12) suf = (v << (lz - 12)) & 7;
    uint32_t e = cavlc_coeff_token_lz[tab * 128 + lz * 8 + suf];
    if (!e) { sd_fail(s, 13); return 0; }
    br_skip(b, e >> 8); total = (e & 255) >> 2; t1 = e & 3;
  }
  if (total == 0) return 0;
  if (total > max_coeff) { sd_fail(s, 14); return 0; }
  int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
  for (int i = 0; i < total; ++i) {
    if (i < t1) { level[i] = (int16_t)(br_get1(b) ? -1 : 1); continue; }
    uint32_t v = br_peek(b, 32);
    int prefix = clz32(v);
    if (prefix > 25) { sd_fail(s, 15); return 0; }
    br_skip(b, prefix + 1);
    int code = (prefix < 15 ? prefix : 15) << suffix_len;
    int ssz = suffix_len;
    if (prefix == 14 && suffix_len == 0) ssz = 4;
    if (prefix >= 15) ssz = prefix - 3;
    if (ssz) code += (int)br_get(b, ssz);
    if (prefix >= 15 && suffix_len == 0) code += 15;
    if (prefix >= 16) code += (1 << (prefix - 3)) - 4096;
    if (i == t1 && t1 < 3) code += 2;
    int lv = (code & 1) ? (-code - 1) >> 1 : (code + 2) >> 1;
    level[i] = (int16_t)lv;
    if (suffix_len == 0) suffix_len = 1;
    if (iabs(lv) > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
  }
  int zeros_left = 0;
  if (total < max_coeff) {
    if (nC < 0) {
      uint32_t e = cavlc_chroma_dc_total_zeros[(total - 1) * 8 + br_peek(b, 3)];
      br_skip(b, e >> 8); zeros_left = e & 255;
    } else {
      uint32_t e = cavlc_total_zeros[(total - 1) * 512 + br_peek(b, 9)];
      if (!e) { sd_fail(s, 16); return 0; }
      br_skip(b, e >> 8); zeros_left = e & 255;
    }
  }
  int idx = total + zeros_left - 1;  // scan index of the highest-frequency coefficient
  if (idx >= max_coeff) { sd_fail(s, 17); return 0; }
  for (int i = 0; i < total; ++i) {
    int pos = scan8 ? scan8[4 * idx + sub] : scan[start + idx];
    s.coef[pos] = level[i];
    if (i + 1 < total) {
      int run = 0;
      if (zeros_left > 0) {
        uint32_t e = cavlc_run_before[(zeros_left < 7 ? zeros_left - 1 : 6) * 2048 + br_peek(b, 11)];
        if (!e) { sd_fail(s, 18); return 0; }
        br_skip(b, e >> 8); run = e & 255;
        zeros_left -= run;
        if (zeros_left < 0) { sd_fail(s, 19); return 0; }
      }
      idx -= 1 + run;
    }
  }
  return total;
}
// The bit reader is copied into registers for the duration of the block (s lives in shared memory on the GPU).
HWB_FN int cavlc_residual(SliceDec &s, int nC, int max_coeff, int start, const uint8_t *scan, const uint8_t *scan8, int sub) {
  BitReader b = s.br;
  int r = cavlc_residual_impl(s, b, nC, max_coeff, start, scan, scan8, sub);
  s.br = b;
  return r;
}

// ================================================================================ CABAC residual
#define HWB_CAT(sig, last, abs) ((sig) | (last) << 10 | (abs) << 20)
HWB_CTABLE uint32_t cabac_cat_ctx[6] = {HWB_CAT(105, 166, 227), HWB_CAT(120, 181, 237), HWB_CAT(134, 195, 247),
                                        HWB_CAT(149, 210, 257), HWB_CAT(152, 213, 266), HWB_CAT(402, 417, 426)};
// 4x4 zig-zag (and the identity order of the 4 chroma DC coefficients) packed 4 bits per scan position: no table load
#define HWB_ZZ4_PACKED 0xFEB7ADC963258410ull
#define HWB_IDENT_PACKED 0xFEDCBA9876543210ull

// cat: 0 I16 DC, 1 I16 AC, 2 luma 4x4, 3 chroma DC, 4 chroma AC, 5 luma 8x8.
// Coefficient staging.  Device: lane l of the warp keeps the coefficients at raster positions l and l + 32 of the block
// being decoded in two registers (every lane runs the same decoder, so a "store" is one predicated move and the block
// leaves with one coalesced store: no shared-memory staging, no clear / emit barriers).  Host: the staging array in the
// slice state.
struct CoefRegs { int lo, hi; };
HWB_HD void coef_reset(SliceDec &s, CoefRegs &r, bool big) {
#if HWB_DEVICE_BUILD
  (void)s; (void)big; r.lo = 0; r.hi = 0;
#else
  (void)r; memset(s.coef, 0, big ? 128 : 32);
#endif
}
HWB_HD void coef_put(SliceDec &s, CoefRegs &r, int pos, int v) {
#if HWB_DEVICE_BUILD
  (void)s;
  const int lane = (int)(threadIdx.x & 31);
  if (lane == (pos & 31)) { if (pos < 32) r.lo = v; else r.hi = v; }
#else
  (void)r; s.coef[pos] = (int16_t)v;
#endif
}
// Append the staged block (nslots * 16 coefficients) to the arena and mark item bits [bit, bit+nslots).
HWB_HD void coef_flush(SliceDec &s, const CoefRegs &r, int bit, int nslots) {
  int16_t *dst = s.o_coefs + (uint64_t)s.coef_next * 16;
#if HWB_DEVICE_BUILD
  const int lane = (int)(threadIdx.x & 31);
  if (nslots == 1) { if (lane < 16) dst[lane] = (int16_t)r.lo; }
  else { dst[lane] = (int16_t)r.lo; dst[lane + 32] = (int16_t)r.hi; }
#else
  (void)r; memcpy(dst, s.coef, (size_t)nslots * 32);
#endif
  s.coef_next += nslots;
  s.out.nzmask |= ((1u << nslots) - 1u) << bit;
}

// One residual block: significance map (kept as a bit mask in registers), then the levels, highest frequency first.
// `sig` / `last` / `abs` point at the category's first context of each kind.  Returns the number of non-zero
// coefficients.  The 8x8 category maps scan positions to contexts through tables and has its own map loop; the chroma DC
// increments min(i, 2) equal i for the three positions that are coded, so every other category shares the plain loop.
// The residual loops inline the decision (measured: as calls, 3000 slices took 330 ms instead of 286 and an intra
// slice 250 ms instead of 200).
#undef HWB_RBIN
#define HWB_RBIN(ptr) cabac_decision(cab, s.cab, ptr, ft)
HWB_HD int cabac_residual_impl(SliceDec &s, CabReg &cab, CoefRegs &cr, const CtxE *ft, int cat, int max_coeff, int start) {
  const uint32_t offs = cabac_cat_ctx[cat];  // ctxIdxOffset of significant_coeff_flag | last_... << 10 | coeff_abs_level_minus1 << 20
  CtxE *sig = s.ctxe + (offs & 1023), *last = s.ctxe + ((offs >> 10) & 1023), *abs_st = s.ctxe + (offs >> 20);
  uint32_t m0 = 0, m1 = 0;  // two registers, not an array: a dynamically indexed array lives in local memory
  const int lastc = max_coeff - 1;
  if (HWB_T8_ON && cat == 5) {
    int i = 0;
#pragma unroll 1
    for (; i < 63; ++i) {
      if (HWB_RBIN(sig + cabac_sig8x8_ctx[i])) {
        if (i < 32) m0 |= 1u << i; else m1 |= 1u << (i - 32);
        if (HWB_RBIN(last + cabac_last8x8_ctx[i])) break;
      }
    }
    if (i == 63) m1 |= 1u << 31;
  } else {
    uint32_t bit = 1;
    const uint32_t endbit = 1u << lastc;
    // Software pipeline: the entries of this position's last-flag context and of the next position's significance
    // context are loaded before the decision that tells which of them is needed (every position has contexts of its
    // own, so neither can be stale), which takes the shared-memory latency off the bin-to-bin dependency chain.
    CtxE e = ctxe_load(sig);
#pragma unroll 1
    while (bit != endbit) {
      const CtxE el = ctxe_load(last);
      const CtxE en = ctxe_load(sig + 1);  // one past the category's map contexts at the final position: loaded, never used
      if (cabac_decide(cab, s.cab, e, sig, ft)) {
        m0 |= bit;
        if (cabac_decide(cab, s.cab, el, last, ft)) break;
      }
      e = en; bit <<= 1; ++sig; ++last;
    }
    m0 |= bit;  // the coefficient the map ended on (last flag set, or the final position, which is inferred)
  }
  // ---- levels
  int eq1 = 1, gt1 = 0;  // eq1: 1 + number of levels equal to 1 so far, capped at 4 (the context increment while gt1 == 0)
  const int cmax = cat == 3 ? 3 : 4;
  const uint64_t scan_packed = cat == 3 ? HWB_IDENT_PACKED : (HWB_ZZ4_PACKED >> (4 * start));
#pragma unroll 1
  for (int half = (HWB_T8_ON && cat == 5) ? 1 : 0; half >= 0; --half) {
    uint32_t mask = half ? m1 : m0;
#pragma unroll 1
    while (mask) {
      const int k = 31 - clz32(mask);
      mask ^= 1u << k;
      // raster position first: the load (8x8) / shift is off the arithmetic decoder's dependency chain
      const int pos = (HWB_T8_ON && cat == 5) ? zigzag8x8[32 * half + k] : (int)((scan_packed >> (4 * k)) & 15);
      int absv = 1;
      if (!HWB_RBIN(abs_st + (gt1 ? 0 : eq1))) {
        eq1 = eq1 < 4 ? eq1 + 1 : 4;
      } else {
        CtxE *st1 = abs_st + 5 + (gt1 < cmax ? gt1 : cmax);
        absv = 2;
#pragma unroll 1
        while (absv < 15 && HWB_RBIN(st1)) absv++;
        if (absv >= 15) {
          const int esc = cabac_escape(cab, s.cab, 0);
          if (esc < 0) { sd_fail(s, 30); return 0; }
          absv += esc;
        }
        gt1++;
      }
      const int sign = cabac_bypass(cab, s.cab);
      coef_put(s, cr, pos, sign ? -absv : absv);
    }
  }
  return popc32(m0) + popc32(m1);
}

// Residual blocks in CABAC mode with the arithmetic decoder held in registers across the whole group:
//   kind 0: one block of category `cat`; coded_block_flag context increment `arg` (< 0: no flag, the block is
//           coded), nzmask item `bit0`.  Returns its number of non-zero coefficients.
//   kind 1: the four 4x4 luma blocks of quadrant q (cat 1 = Intra16x16 AC, cat 2 = other luma)
//   kind 4: the four chroma AC blocks of plane q
//   for kinds 1 and 4 `arg` is the coded_block_flag value assumed for unavailable neighbours; the neighbour
//   caches (total_coeff per block) are read and updated here.
// CAVLC staging helpers (the CAVLC decoder scatters levels into s.coef)
HWB_HD void coef_clear(SliceDec &s, int n) {
  HWB_LANES(l)
  if (2 * l < n) ((uint32_t *)s.coef)[l] = 0;
  HWB_LANES_END
}
HWB_HD void coef_emit(SliceDec &s, int bit, int nslots) {
  uint32_t *dst = (uint32_t *)(s.o_coefs + (uint64_t)s.coef_next * 16);
  HWB_LANES(l)
  if (l < nslots * 8) dst[l] = ((const uint32_t *)s.coef)[l];
  HWB_LANES_END
  s.coef_next += nslots;
  s.out.nzmask |= ((1u << nslots) - 1u) << bit;
}
// Group tables: byte offset (from the slice state) of each block's entry in its total_coeff cache; the left neighbour
// is one byte before it, the top neighbour `up` bytes.
HWB_CTABLE uint16_t blk_nz_off[24] = {
#define HWB_NZ(bx, by) (uint16_t)(HWB_O(nz_cache) + HWB_CI(bx, by))
  HWB_NZ(0, 0), HWB_NZ(1, 0), HWB_NZ(0, 1), HWB_NZ(1, 1), HWB_NZ(2, 0), HWB_NZ(3, 0), HWB_NZ(2, 1), HWB_NZ(3, 1),
  HWB_NZ(0, 2), HWB_NZ(1, 2), HWB_NZ(0, 3), HWB_NZ(1, 3), HWB_NZ(2, 2), HWB_NZ(3, 2), HWB_NZ(2, 3), HWB_NZ(3, 3),
#undef HWB_NZ
#define HWB_CNZ(p, bx, by) (uint16_t)(HWB_O(cnz_cache[p]) + ((by) + 1) * 4 + (bx) + 1)
  HWB_CNZ(0, 0, 0), HWB_CNZ(0, 1, 0), HWB_CNZ(0, 0, 1), HWB_CNZ(0, 1, 1), HWB_CNZ(1, 0, 0), HWB_CNZ(1, 1, 0), HWB_CNZ(1, 0, 1), HWB_CNZ(1, 1, 1)
#undef HWB_CNZ
};
HWB_FN int cabac_blocks(SliceDec &s, int kind, int q, int cat, int arg, int bit0) {
  HWB_CAB_ENTER(s);
  CoefRegs cr;
  const int nblk = kind == 0 ? 1 : 4;
  const bool big = HWB_T8_ON && cat == 5;
  const int maxc = big ? 64 : (cat == 3 ? 4 : ((cat == 1 || cat == 4) ? 15 : 16));
  CtxE *cbf = s.ctxe + 85 + 4 * (big ? 0 : cat);
  const int g = kind == 1 ? q * 4 : 16 + q * 4, up = kind == 1 ? 8 : 4;
  if (kind != 0) bit0 = kind == 1 ? NZ_LUMA0 + q * 4 : (q ? NZ_CR0 : NZ_CB0);
  const int unavail_mask = 0x7f | (arg << 7);  // kinds 1 and 4 only
  int n = 0;
#pragma unroll 1
  for (int k = 0; k < nblk; ++k) {
    int inc = arg;
    uint8_t *nzp = nullptr;
    if (kind != 0) {
      nzp = (uint8_t *)&s + blk_nz_off[g + k];
      // coded flag of a neighbour: its total_coeff (0..16) != 0; 0x80 marks an unavailable one, which counts as `arg`
      inc = ((nzp[-1] & unavail_mask) != 0) + 2 * ((*(nzp - up) & unavail_mask) != 0);
    }
    n = 0;
    if (inc < 0 || HWB_RBIN(cbf + inc)) {
      coef_reset(s, cr, big);
      n = cabac_residual_impl(s, cab, cr, ft, cat, maxc, (cat == 1 || cat == 4) ? 1 : 0);
      coef_flush(s, cr, bit0 + k, big ? 4 : 1);
    }
    if (nzp) *nzp = (uint8_t)n;
  }
  HWB_CAB_LEAVE(s);
  return n;
}
// Hides a constant from the compiler's interprocedural constant propagation: with literal `kind` arguments it cloned
// cabac_blocks once per call pattern (three 6.5 KB copies of the residual decoder in a fetch-bound kernel).
HWB_HD int opaque(int v) {
#if HWB_DEVICE_BUILD
  asm volatile("" : "+r"(v));
#endif
  return v;
}
HWB_HD int cabac_block(SliceDec &s, int cat, int cbf_inc, int bit) { return cabac_blocks(s, opaque(0), 0, cat, cbf_inc, bit); }

// identity scan for blocks whose coefficients are already in raster order
HWB_TABLE uint8_t scan_ident4[4] = {0, 1, 2, 3};

// ================================================================================ residual (both modes)
// nnz[] receives total_coeff per luma block (raster) and chroma block.
HWB_FN void decode_residual(SliceDec &s, bool i16, int cbp, bool t8) {
  const bool cabac = HWB_IS_CABAC(s);
  const bool intra = s.out.mbtype != MB_INTER;
  const int cbf_unavail = intra ? 1 : 0;
  if (i16) {
    if (cabac) {
      int a = s.availA ? ((s.left.flags & NBF_IPCM) ? 1 : (s.left.cbf >> NZ_LUMA_DC) & 1) : cbf_unavail;
      int bq = s.availB ? ((top_flags(s) & NBF_IPCM) ? 1 : (s.top_words[1] >> NZ_LUMA_DC) & 1) : cbf_unavail;
      cabac_block(s, 0, a + 2 * bq, NZ_LUMA_DC);
    } else {
      coef_clear(s, 16);
      int n = cavlc_residual(s, cavlc_nc(s.nz_cache[HWB_CI(-1, 0)], s.nz_cache[HWB_CI(0, -1)]), 16, 0, zigzag4x4, nullptr, 0);
      if (n) coef_emit(s, NZ_LUMA_DC, 1);
    }
  }
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    if (!((cbp >> q) & 1)) continue;
    if (HWB_T8_ON && t8) {
      int n = 0;
      if (cabac) {
        n = cabac_block(s, 5, -1, NZ_LUMA0 + q * 4);
        const uint32_t v = (uint32_t)(n > 16 ? 16 : n) * 0x0101u;
        const int bx = (q & 1) * 2, by = (q >> 1) * 2;
        *(uint16_t *)(s.nz_cache + HWB_CI(bx, by)) = (uint16_t)v; *(uint16_t *)(s.nz_cache + HWB_CI(bx, by + 1)) = (uint16_t)v;
      } else {
        coef_clear(s, 64);
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          int bx = (q & 1) * 2 + (k & 1), by = (q >> 1) * 2 + (k >> 1);
          int nc = cavlc_nc(s.nz_cache[HWB_CI(bx - 1, by)], s.nz_cache[HWB_CI(bx, by - 1)]);
          int m = cavlc_residual(s, nc, 16, 0, nullptr, zigzag8x8, k);
          s.nz_cache[HWB_CI(bx, by)] = (uint8_t)m;
          n += m;
        }
        if (n) coef_emit(s, NZ_LUMA0 + q * 4, 4);
      }
    } else {
      if (cabac) { cabac_blocks(s, opaque(1), q, i16 ? 1 : 2, cbf_unavail, 0); continue; }
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        int z = q * 4 + k, bx = z2x(z), by = z2y(z);
        int na = s.nz_cache[HWB_CI(bx - 1, by)], nb = s.nz_cache[HWB_CI(bx, by - 1)];
        coef_clear(s, 16);
        int n = cavlc_residual(s, cavlc_nc(na, nb), i16 ? 15 : 16, i16 ? 1 : 0, zigzag4x4, nullptr, 0);
        if (n) coef_emit(s, NZ_LUMA0 + z, 1);
        s.nz_cache[HWB_CI(bx, by)] = (uint8_t)n;
      }
    }
  }
  if (cbp & 0x30) {
#pragma unroll 1
    for (int p = 0; p < 2; ++p) {
      int bit = p ? NZ_CR_DC : NZ_CB_DC;
      if (cabac) {
        int a = s.availA ? ((s.left.flags & NBF_IPCM) ? 1 : (s.left.cbf >> bit) & 1) : cbf_unavail;
        int bq = s.availB ? ((top_flags(s) & NBF_IPCM) ? 1 : (s.top_words[1] >> bit) & 1) : cbf_unavail;
        cabac_block(s, 3, a + 2 * bq, bit);
      } else {
        coef_clear(s, 16);
        int n = cavlc_residual(s, -1, 4, 0, scan_ident4, nullptr, 0);
        if (n) coef_emit(s, bit, 1);
      }
    }
  }
  if (cbp & 0x20) {
#pragma unroll 1
    for (int p = 0; p < 2; ++p) {
      if (cabac) { cabac_blocks(s, opaque(4), p, opaque(4), cbf_unavail, 0); continue; }
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        int bx = k & 1, by = k >> 1;
        int na = s.cnz_cache[p][(by + 1) * 4 + bx], nb = s.cnz_cache[p][by * 4 + bx + 1];
        coef_clear(s, 16);
        int n = cavlc_residual(s, cavlc_nc(na, nb), 15, 1, zigzag4x4, nullptr, 0);
        if (n) coef_emit(s, (p ? NZ_CR0 : NZ_CB0) + k, 1);
        s.cnz_cache[p][(by + 1) * 4 + bx + 1] = (uint8_t)n;
      }
    }
  }
}

// ================================================================================ CABAC syntax elements
HWB_FN int cabac_intra_mb_type(SliceDec &s, int ctx_base, bool islice) {
  int st = ctx_base, first = ctx_base;
  if (islice) {
    if (s.availA && !(s.left.flags & NBF_INXN)) first++;
    if (s.availB && !(top_flags(s) & NBF_INXN)) first++;
    st += 2;
  }
  HWB_CAB_ENTER(s);
  int t = 0;
  if (HWB_BIN(s, first)) {
    if (cabac_terminate(cab, s.cab)) t = 25;
    else {
      // five suffix bins: context (relative to st) and weight, 4 bits each; the third only follows a set second one
      const uint32_t ctxs = islice ? 0x54321u : 0x33221u, adds = 0x1244Cu;
      t = 1;
      int prev = 1;
#pragma unroll 1
      for (int i = 0; i < 5; ++i) {
        if (i == 2 && !prev) continue;
        prev = HWB_BIN(s, st + (int)((ctxs >> (4 * i)) & 15));
        if (prev) t += (int)((adds >> (4 * i)) & 15);
      }
    }
  }
  HWB_CAB_LEAVE(s);
  return t;
}

// B slice: mb_skip_flag and mb_type in one engine session.  Returns -1 (B_Skip), 0..22, or 23 (intra: the intra mb_type
// follows, decoded by the caller).
HWB_FN int cabac_b_header(SliceDec &s, int skip_ctx) {
  int ctx = 0;
  if (s.availA && !(s.left.flags & NBF_DIRECT16)) ctx++;
  if (s.availB && !(top_flags(s) & NBF_DIRECT16)) ctx++;
  const int st = 27;
  HWB_CAB_ENTER(s);
  int r = -1;
  if (!HWB_BIN(s, 24 + skip_ctx)) {
    r = 0;
    if (HWB_BIN(s, st + ctx)) {
      if (!HWB_BIN(s, st + 3)) r = 1 + HWB_BIN(s, st + 5);
      else {
        int bits = HWB_BIN(s, st + 4);
#pragma unroll 1
        for (int i = 0; i < 3; ++i) bits = (bits << 1) | HWB_BIN(s, st + 5);
        if (bits < 8) r = bits + 3;
        else if (bits == 13) r = 23;
        else if (bits == 14) r = 11;
        else if (bits == 15) r = 22;
        else r = ((bits << 1) | HWB_BIN(s, st + 5)) - 4;
      }
    }
  }
  HWB_CAB_LEAVE(s);
  return r;
}

HWB_FN int cabac_b_sub_type(SliceDec &s) {
  const int st = 36;
  HWB_CAB_ENTER(s);
  int t = 0;
  if (HWB_BIN(s, st)) {
    if (!HWB_BIN(s, st + 1)) t = 1 + HWB_BIN(s, st + 3);
    else {
      t = 3;
      bool two = true;
      if (HWB_BIN(s, st + 2)) {
        if (HWB_BIN(s, st + 3)) { t = 11 + HWB_BIN(s, st + 3); two = false; }
        else t += 4;
      }
      if (two) { t += 2 * HWB_BIN(s, st + 3); t += HWB_BIN(s, st + 3); }
    }
  }
  HWB_CAB_LEAVE(s);
  return t;
}

HWB_FN int cabac_ref_idx(SliceDec &s, int l, int bx, int by) {
  int ra = s.ref_cache[l][HWB_CI(bx - 1, by)], rb = s.ref_cache[l][HWB_CI(bx, by - 1)];
  int ctx = 0;
  if (ra > 0 && !s.dir_cache[HWB_CI(bx - 1, by)]) ctx++;
  if (rb > 0 && !s.dir_cache[HWB_CI(bx, by - 1)]) ctx += 2;
  int ref = 0;
  HWB_CAB_ENTER(s);
#pragma unroll 1
  while (HWB_BIN(s, 54 + ctx)) {
    ref++;
    ctx = (ctx >> 2) + 4;
    if (ref >= 32) { sd_fail(s, 40); ref = 0; break; }
  }
  HWB_CAB_LEAVE(s);
  return ref;
}

// Both components of a motion vector difference in one call.  sa / sb: sums of the neighbours' absolute differences
// (context selection); returns the differences and their absolute values capped at 70 (what the neighbours-to-come see).
struct MvdPair { int dx, dy, ax, ay; };
HWB_FN MvdPair cabac_mvd_pair(SliceDec &s, int sa, int sb) {
  HWB_CAB_ENTER(s);
  MvdPair r; r.dx = r.dy = r.ax = r.ay = 0;
#pragma unroll 1
  for (int comp = 0; comp < 2; ++comp) {
    const int amvd = comp ? sb : sa, cb = comp ? 47 : 40;
    int mvd = 0, absv = 0;
    if (HWB_BIN(s, cb + (amvd < 3 ? 0 : (amvd > 32 ? 2 : 1)))) {
      mvd = 1;
      int ctx = cb + 3;
#pragma unroll 1
      while (mvd < 9 && HWB_BIN(s, ctx)) { if (mvd < 4) ctx++; mvd++; }
      if (mvd >= 9) {
        const int esc = cabac_escape(cab, s.cab, 3);
        if (esc < 0) sd_fail(s, 41); else mvd += esc;
      }
      absv = mvd < 70 ? mvd : 70;
      if (HWB_BYP(s)) mvd = -mvd;
    }
    if (comp) { r.dy = mvd; r.ay = absv; } else { r.dx = mvd; r.ax = absv; }
  }
  HWB_CAB_LEAVE(s);
  return r;
}

HWB_FN int cabac_cbp(SliceDec &s) {
  const LeftCtx &L = s.left;
  // luma: cbp bits of neighbours; unavailable / I_PCM behave as "all coded"
  int cbpa = s.availA ? ((L.flags & NBF_IPCM) ? 0x2F : L.cbp) : 0x0F;
  int cbpb = s.availB ? ((top_flags(s) & NBF_IPCM) ? 0x2F : (int)((s.top_words[0] >> 8) & 0xff)) : 0x0F;
  int cbp = 0;
  HWB_CAB_ENTER(s);
#pragma unroll 1
  for (int b8 = 0; b8 < 4; ++b8) {
    int a = (b8 & 1) ? !((cbp >> (b8 - 1)) & 1) : !((cbpa >> (b8 + 1)) & 1);
    int bq = (b8 & 2) ? !((cbp >> (b8 - 2)) & 1) : !((cbpb >> (b8 + 2)) & 1);
    cbp |= HWB_BIN(s, 73 + a + 2 * bq) << b8;
  }
  int ca = s.availA ? (cbpa >> 4) & 3 : 0, cb = s.availB ? (cbpb >> 4) & 3 : 0;
  // chroma: up to two bins; the second one's contexts sit four further on
  int ctx = 77 + (ca > 0) + 2 * (cb > 0), val = 0;
#pragma unroll 1
  for (int k = 0; k < 2; ++k) {
    if (!HWB_BIN(s, ctx)) break;
    val++;
    ctx = 77 + 4 + (ca == 2) + 2 * (cb == 2);
  }
  HWB_CAB_LEAVE(s);
  return cbp | (val << 4);
}

HWB_FN int cabac_dqp(SliceDec &s) {
  int ctx = s.last_dqp != 0, val = 0;
  HWB_CAB_ENTER(s);
#pragma unroll 1
  while (HWB_BIN(s, 60 + ctx)) {
    ctx = 2 + (ctx >> 1);
    val++;
    if (val > 104) { sd_fail(s, 42); val = 0; break; }
  }
  HWB_CAB_LEAVE(s);
  return (val & 1) ? (val + 1) >> 1 : -((val + 1) >> 1);
}

HWB_FN int cabac_chroma_mode(SliceDec &s) {
  int ctx = 64;
  if (s.availA && s.left.cmode != 0) ctx++;
  if (s.availB && ((s.top_words[0] >> 16) & 0xff) != 0) ctx++;
  HWB_CAB_ENTER(s);
  int m = 0;
#pragma unroll 1
  for (; m < 3; ++m) {  // truncated unary, maximum 3; every bin after the first uses context 67
    if (!HWB_BIN(s, ctx)) break;
    ctx = 64 + 3;
  }
  HWB_CAB_LEAVE(s);
  return m;
}

// prev_intra4x4/8x8_pred_mode_flag + rem_intra_pred_mode of every block of an I_NxN macroblock, and the mode derivation
// (8.3.1.1) that goes with them.
HWB_FN void cabac_intra_modes(SliceDec &s, bool t8) {
  MbInfo &o = s.out;
  const int nb = t8 ? 4 : 16;
  HWB_CAB_ENTER(s);
#pragma unroll 1
  for (int k = 0; k < nb; ++k) {
    const int bx = t8 ? (k & 1) * 2 : z2x(k), by = t8 ? (k >> 1) * 2 : z2y(k);
    int8_t *im = s.im_cache + HWB_CI(bx, by);
    const int ma = im[-1], mb_ = im[-8];
    const int pred = (ma < 0 || mb_ < 0) ? 2 : (ma < mb_ ? ma : mb_);
    int mode = pred;
    if (!HWB_BIN(s, 68)) {
      int rem = 0;
#pragma unroll 1
      for (int i = 0; i < 3; ++i) rem |= HWB_BIN(s, 69) << i;
      mode = rem < pred ? rem : rem + 1;
    }
    o.i4modes[k] = (uint8_t)mode;
    im[0] = (int8_t)mode;
    if (t8) { im[1] = (int8_t)mode; im[8] = (int8_t)mode; im[9] = (int8_t)mode; }
  }
  HWB_CAB_LEAVE(s);
}

// P slice: mb_skip_flag and, for a coded macroblock, mb_type up to the intra prefix.  Returns -1 (skipped), 0..3
// (P_L0_16x16, P_L0_L0_16x8, P_L0_L0_8x16, P_8x8) or 5 (intra: the intra mb_type follows).
HWB_FN int cabac_p_header(SliceDec &s, int skip_ctx) {
  HWB_CAB_ENTER(s);
  int r = -1;
  if (!HWB_BIN(s, 11 + skip_ctx)) {
    if (HWB_BIN(s, 14)) r = 5;
    else {
      const int b1 = HWB_BIN(s, 15);
      const int b2 = HWB_BIN(s, 16 + b1);
      r = b1 ? 2 - b2 : 3 * b2;
    }
  }
  HWB_CAB_LEAVE(s);
  return r;
}
// P slice: the four sub_mb_type values of a P_8x8 macroblock, 2 bits each
HWB_FN int cabac_p_sub_types(SliceDec &s) {
  HWB_CAB_ENTER(s);
  int r = 0;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    int t = 0;
    if (!HWB_BIN(s, 21)) t = !HWB_BIN(s, 22) ? 1 : (HWB_BIN(s, 23) ? 2 : 3);
    r |= t << (2 * q);
  }
  HWB_CAB_LEAVE(s);
  return r;
}

// Publish a decoded macroblock: MbInfo, final motion data, and the neighbour context for the macroblocks to come.
// The caches must hold the macroblock's total_coeff (nz/cnz), intra modes (I_NxN only), motion and direct flags.
// Shared by the decoder and by the stream generator's entropy writer.
HWB_FN void finish_mb(SliceDec &s, bool skipped, bool direct16, bool is_pcm) {
  const ChunkCtx &c = *s.c;
  const SliceDesc &sd = *s.sd;
  const bool B = HWB_IS_B(sd.slice_type);
  const int nl = B ? 2 : (sd.slice_type == SLICE_P ? 1 : 0);
  MbInfo &o = s.out;
  const int f = s.pd->frame;
  const bool inter = o.mbtype == MB_INTER;
  const bool inxn = o.mbtype == MB_I4x4 || o.mbtype == MB_I8x8;
  const bool l1_none = inter && !B && s.pd->has_inter == 2;  // P macroblock of a picture that also has B slices
  const uint32_t imv = (inter && s.pd->constrained_intra_pred) ? 0xFFFFFFFFu : 0x02020202u;
  const uint8_t flags = (uint8_t)((inter ? 0 : NBF_INTRA) | (is_pcm ? NBF_IPCM : 0) | (skipped ? NBF_SKIP : 0) | (direct16 ? NBF_DIRECT16 : 0) |
                                  ((o.flags & MBF_T8x8) ? NBF_T8 : 0) | (o.mbtype == MB_I16x16 ? NBF_I16 : 0) | (inxn ? NBF_INXN : 0));
  const uint32_t cbf = is_pcm ? 0x7FFFFFFu : o.nzmask;
  const int bot = HWB_CI(0, 3);
  NbCtx *n = s.line + s.mbx;
  MbInfo *dst = s.o_mbinfo + s.mbaddr;
  // ---- phase 1: normalise the interior so that the right column / bottom row say what neighbours must see; save
  // what the next macroblock's top-left needs from the line entry that is about to be overwritten; MbInfo out
  HWB_LANES(l)
  if (l < 16) {
    const int ci = HWB_CI(l & 3, l >> 2);
    if (!inter) {  // both lists, whatever the slice type (an unused list's cache is never read)
      s.ref_cache[0][ci] = REF_NONE; set4(s.mv_cache[0][ci], 0); *(uint16_t *)s.mvd_cache[0][ci] = 0;
      s.ref_cache[1][ci] = REF_NONE; set4(s.mv_cache[1][ci], 0); *(uint16_t *)s.mvd_cache[1][ci] = 0;
    }
    if (!inxn) s.im_cache[ci] = (int8_t)imv;
  } else if (l < 24) {
    ((uint32_t *)dst)[l - 16] = ((const uint32_t *)&o)[l - 16];
  } else if (l < 26) {
    const int k = l - 24;
    if (k < nl) { s.tl_ref[k] = n->ref_b[k][3]; cpy4(s.tl_mv[k], n->mv_b[k][3]); }
  }
  HWB_LANES_END
  // ---- phase 2: motion out (lane = 4x4 block, lanes 0..15 list 0, 16..31 list 1), line entry (lane = word), left context
  uint32_t dm = 0;  // direct-mode flags of the bottom row: only B slices ever set any
  if (B) for (int x = 0; x < 4; ++x) if (s.dir_cache[bot + x]) dm |= 1u << x;
  // words 0..3 of the line entry are staged where fill_caches keeps the top neighbour's (no longer needed)
  s.top_words[0] = flags | ((uint32_t)o.cbp << 8) | ((uint32_t)o.cmode << 16) | (dm << 24);
  s.top_words[1] = cbf;
  s.top_words[2] = s.cnz_cache[0][9] | ((uint32_t)s.cnz_cache[0][10] << 8) | ((uint32_t)s.cnz_cache[1][9] << 16) | ((uint32_t)s.cnz_cache[1][10] << 24);
  s.top_words[3] = 0;
  int reach = 0, reach_x = 0;  // per lane on the device, accumulated over the lane loop on the host
  HWB_LANES(l)
  if (inter) {
    const int k = l >> 4, i = l & 15;
    if (k < nl) {
      const uint32_t mvw = *(const uint32_t *)s.mv_cache[k][HWB_CI(i & 3, i >> 2)];
      ((uint32_t *)(s.o_mv[k] + (uint64_t)s.mbaddr * 32))[i] = mvw;
      if (s.ref_cache[k][HWB_CI(i & 3, i >> 2)] >= 0) {
        // bottom sample row of the block, displaced, plus the 3 rows the 6-tap filter reads below it
        const int row = (s.mby * 16 + (i >> 2) * 4 + 3 + ((int)(int16_t)(mvw >> 16) >> 2) + 3) >> 4;
        reach = imax(reach, 1 + clip3(0, c.mb_h - 1, row));
        // rightmost sample column read (+3 filter taps); it is final once the macroblock holding column + 3 has been
        // deblocked (the next macroblock's left edge changes up to 3 columns): progress needed = that column + 1
        const int col = s.mbx * 16 + (i & 3) * 4 + 3 + ((int)(int16_t)(mvw & 0xffff) >> 2) + 3;
        reach_x = imax(reach_x, clip3(1, c.mb_w, ((col + 3) >> 4) + 1 - s.mbx));
      }
      if (i < 4) {
        const int r = s.ref_cache[k][HWB_CI((i & 1) * 2, (i >> 1) * 2)];
        s.o_refidx[k][(uint64_t)s.mbaddr * 4 + i] = (int8_t)r;
        s.o_refpic[k][(uint64_t)s.mbaddr * 4 + i] = r >= 0 ? sd.ref_frame[k][r] : (int16_t)-1;
      }
    } else if (l1_none && i < 4) {
      s.o_refidx[1][(uint64_t)s.mbaddr * 4 + i] = -1;
      s.o_refpic[1][(uint64_t)s.mbaddr * 4 + i] = -1;
    }
  }
  if (l < 20) ((uint32_t *)n)[l] = *(const uint32_t *)((const uint8_t *)&s + s.t_line[l]);
  HWB_LANES_END
  if (inter) {
#if HWB_DEVICE_BUILD
    reach = (int)__reduce_max_sync(0xffffffffu, (unsigned)reach);  // two independent maxima
    reach_x = (int)__reduce_max_sync(0xffffffffu, (unsigned)reach_x);
#endif
    if (reach > s.row_reach) s.row_reach = reach;
    if (reach_x > s.row_reach_x) s.row_reach_x = reach_x;
  }
  s.left.flags = flags; s.left.cbp = o.cbp; s.left.cmode = o.cmode; s.left.cbf = cbf;
}

// ================================================================================ macroblock layer
// prediction flags (1 L0, 2 L1, 3 Bi) of the two partitions of B 16x8 / 8x16 types, by (mb_type-4)>>1
HWB_TABLE uint8_t b_part_pred[18] = {1, 1, 2, 2, 1, 2, 2, 1, 1, 3, 2, 3, 3, 1, 3, 2, 3, 3};

HWB_FN int read_ref(SliceDec &s, int l, int bx, int by) {
  int nref = s.sd->num_ref[l];
  if (nref <= 1) return 0;
  if (HWB_IS_CABAC(s)) return cabac_ref_idx(s, l, bx, by);
  if (nref == 2) return s_get(s, 1) ^ 1;
  return (int)s_ue(s);
}

HWB_FN void read_mvd_and_set(SliceDec &s, int l, int bx, int by, int w, int h, int ref, int shape) {
  int px, py;
  pred_mv(s, l, bx, by, w, ref, shape, px, py);
  int dx, dy, ax = 0, ay = 0;
  if (HWB_IS_CABAC(s)) {
    int sa = s.mvd_cache[l][HWB_CI(bx - 1, by)][0] + s.mvd_cache[l][HWB_CI(bx, by - 1)][0];
    int sb = s.mvd_cache[l][HWB_CI(bx - 1, by)][1] + s.mvd_cache[l][HWB_CI(bx, by - 1)][1];
    const MvdPair m = cabac_mvd_pair(s, sa, sb);
    dx = m.dx; dy = m.dy; ax = m.ax; ay = m.ay;
  } else {
    dx = s_se(s); dy = s_se(s);
  }
  set_motion(s, l, bx, by, w, h, ref, px + dx, py + dy, ax, ay);
}

// Decode one macroblock; s.mbx/mby/mbaddr and availability set by the caller, which has also called fill_caches.  `skipped`: P_Skip/B_Skip.
HWB_FN void decode_mb(SliceDec &s, bool skipped) {
  const ChunkCtx &c = *s.c;
  const SliceDesc &sd = *s.sd;
  const int st = sd.slice_type;
  const bool B = HWB_IS_B(st);
  const int nl = B ? 2 : 1;
  MbInfo &o = s.out;
  o.mbtype = MB_INTER; o.qp = (uint8_t)s.qp; o.cbp = 0; o.flags = 0; o.imode = 0; o.cmode = 0;
  o.slice = (uint16_t)s.slice_num; o.nzmask = 0; o.coef_off = s.coef_next;
  for (int i = 0; i < 4; ++i) set4(o.i4modes + 4 * i, 0x02020202u);
  bool direct16 = false;
  uint32_t dirq = 0;  // quadrants predicted in direct mode
  int8_t (*dref)[4] = s.dref;
  int16_t (*dmv)[16][2] = s.dmv;
  bool is_pcm = false;

  if (skipped) {
    o.flags |= MBF_SKIP;
    s.last_dqp = 0;
    if (!B) {
      int mx = 0, my = 0;
      MvRef A = mv_at(s, 0, -1, 0), Bn = mv_at(s, 0, 0, -1);
      if (!(A.ref == REF_UNAVAIL || Bn.ref == REF_UNAVAIL || (A.ref == 0 && A.mx == 0 && A.my == 0) || (Bn.ref == 0 && Bn.mx == 0 && Bn.my == 0)))
        pred_mv(s, 0, 0, 0, 4, 0, 0, mx, my);
      set_motion(s, 0, 0, 0, 4, 4, 0, mx, my, 0, 0);
    } else {
      direct16 = true; dirq = 15;
      direct_predict(s, 15, dref, dmv);
#pragma unroll 1
      for (int l = 0; l < 2; ++l) for (int q = 0; q < 4; ++q) apply_direct(s, l, q, dref, dmv);
    }
  } else {
    // ---------------- mb_type
    int mbt;
    if (HWB_IS_CABAC(s)) {
      if (st == SLICE_I) mbt = cabac_intra_mb_type(s, 3, true);
      else if (st == SLICE_P) {
        mbt = s.pre_mbt;  // decoded with mb_skip_flag (decode_slice)
        if (mbt == 5) mbt += cabac_intra_mb_type(s, 17, false);
      } else {
        mbt = s.pre_mbt;
        if (mbt == 23) mbt += cabac_intra_mb_type(s, 32, false);
      }
    } else mbt = (int)s_ue(s);
    int imbt = -1;  // intra mb_type 0..25
    if (st == SLICE_I) imbt = mbt;
    else if (st == SLICE_P && mbt >= 5) imbt = mbt - 5;
    else if (B && mbt >= 23) imbt = mbt - 23;
    if (imbt > 25 || (st == SLICE_P && mbt > 30) || (B && mbt > 48)) { sd_fail(s, 50); return; }

    if (imbt == 25) {
      // ---------------- I_PCM
      is_pcm = true;
      o.mbtype = MB_IPCM; o.qp = 0; o.cbp = 0x2F;
      uint8_t *dst = (uint8_t *)(s.o_coefs + (uint64_t)s.coef_next * 16);
      if (HWB_IS_CABAC(s)) {
        // pcm_alignment_zero_bits, 384 raw bytes, then the arithmetic decoder restarts (9.3.1.2)
        const uint32_t p = (cabac_bitpos(s.cab) + 7) >> 3;
        HWB_LANES(l)
        for (int i = l; i < 384; i += 32) dst[i] = s.br.base[p + i];
        HWB_LANES_END
        cabac_start(s.cab, s.br.base, p + 384);
      } else {
        br_align(s.br);
#pragma unroll 1
        for (int i = 0; i < 384; ++i) dst[i] = (uint8_t)s_get(s, 8);
      }
      s.coef_next += 12; o.nzmask = 0xFFF;
      for (int y = 0; y < 4; ++y) set4(s.nz_cache + HWB_CI(0, y), 0x10101010u);
      for (int p = 0; p < 2; ++p) { s.cnz_cache[p][5] = s.cnz_cache[p][6] = s.cnz_cache[p][9] = s.cnz_cache[p][10] = 16; }
      s.last_dqp = 0;
    } else if (imbt >= 0) {
      // ---------------- intra
      bool t8 = false;
      int cbp;
      if (imbt == 0) {
        if (HWB_T8_ON && s.pd->transform8x8_mode) {
          if (HWB_IS_CABAC(s)) {
            int ctx = (s.availA && (s.left.flags & NBF_T8)) + (s.availB && (top_flags(s) & NBF_T8));
            t8 = cabac_bin(s, 399 + ctx) != 0;
          } else t8 = s_get(s, 1) != 0;
        }
        o.mbtype = t8 ? MB_I8x8 : MB_I4x4;
        if (t8) o.flags |= MBF_T8x8;
        if (HWB_IS_CABAC(s)) cabac_intra_modes(s, t8);
        else {
          const int nb = t8 ? 4 : 16;
#pragma unroll 1
          for (int k = 0; k < nb; ++k) {
            int bx = t8 ? (k & 1) * 2 : z2x(k), by = t8 ? (k >> 1) * 2 : z2y(k);
            int ma = s.im_cache[HWB_CI(bx - 1, by)], mb_ = s.im_cache[HWB_CI(bx, by - 1)];
            int pred = (ma < 0 || mb_ < 0) ? 2 : (ma < mb_ ? ma : mb_);
            int mode;
            if (s_get(s, 1)) mode = pred;
            else { int rem = (int)s_get(s, 3); mode = rem < pred ? rem : rem + 1; }
            o.i4modes[k] = (uint8_t)mode;
            int wd = t8 ? 2 : 1;
#pragma unroll 1
            for (int y = by; y < by + wd; ++y) for (int x = bx; x < bx + wd; ++x) s.im_cache[HWB_CI(x, y)] = (int8_t)mode;
          }
        }
      } else {
        o.mbtype = MB_I16x16;
        o.imode = (uint8_t)((imbt - 1) & 3);
      }
      o.cmode = (uint8_t)(HWB_IS_CABAC(s) ? cabac_chroma_mode(s) : (int)s_ue(s));
      if (o.cmode > 3) { sd_fail(s, 51); return; }
      if (imbt == 0) {
        if (HWB_IS_CABAC(s)) cbp = cabac_cbp(s);
        else { uint32_t k = s_ue(s); if (k > 47) { sd_fail(s, 52); return; } cbp = golomb_to_intra_cbp[k]; }
      } else {
        cbp = (((imbt - 1) / 4) % 3) << 4 | ((imbt - 1) >= 12 ? 15 : 0);
      }
      o.cbp = (uint8_t)cbp;
      if (cbp || imbt > 0) {
        int dqp = HWB_IS_CABAC(s) ? cabac_dqp(s) : s_se(s);
        s.last_dqp = dqp;
        s.qp = (s.qp + dqp + 52) % 52;
      } else s.last_dqp = 0;
      o.qp = (uint8_t)s.qp;
      decode_residual(s, imbt > 0, cbp, t8);
    } else {
      // ---------------- inter
      // One code path for every partitioning: the macroblock is split into `ng` groups that share a reference
      // (the 16x16 / 16x8 / 8x16 partitions, or the four 8x8 quadrants), each group into 1, 2 or 4 motion
      // partitions.  s.grp[g] = bx | by << 2 | (w-1) << 4 | (h-1) << 6 | sub-shape << 8 | pred_mv shape << 10.
      int cbp;
      bool t8_allowed = true;
      int8_t *sub = s.sub, *pf = s.pf;
      int ng = 0;
      if (B && mbt == 0) {
        direct16 = true; dirq = 15;
        direct_predict(s, 15, dref, dmv);
#pragma unroll 1
        for (int l = 0; l < 2; ++l) for (int q = 0; q < 4; ++q) apply_direct(s, l, q, dref, dmv);
        t8_allowed = s.pd->direct_8x8_inference != 0;
      } else if ((!B && mbt >= 3) || (B && mbt == 22)) {
        // 8x8 with sub-macroblock types.  sub shapes: 0: 8x8, 1: 8x4, 2: 4x8, 3: 4x4; pred flags: 1 L0, 2 L1, 3 Bi
        ng = 4;
        const int psub = (HWB_IS_CABAC(s) && !B) ? cabac_p_sub_types(s) : 0;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          int t;
          if (HWB_IS_CABAC(s)) {
            if (B) t = cabac_b_sub_type(s);
            else t = (psub >> (2 * q)) & 3;
          } else t = (int)s_ue(s);
          if (t > (B ? 12 : 3)) { sd_fail(s, 53); return; }
          int shp = t, f = 1;
          if (B) {
            if (t == 0) { dirq |= 1u << q; shp = 0; f = 0; if (!s.pd->direct_8x8_inference) t8_allowed = false; }
            else { shp = t <= 3 ? 0 : (t >= 10 ? 3 : ((t & 1) ? 2 : 1)); f = t <= 3 ? t : (t >= 10 ? t - 9 : ((t - 4) >> 1) + 1); }
          }
          if (shp != 0) t8_allowed = false;
          sub[q] = (int8_t)t; pf[q] = (int8_t)f;
          s.grp[q] = (uint16_t)(((q & 1) * 2) | ((q >> 1) * 2) << 2 | 1 << 4 | 1 << 6 | shp << 8);
        }
        if (dirq) direct_predict(s, (int)dirq, dref, dmv);
      } else {
        // 16x16, 16x8, 8x16
        int shape, pf0, pf1;
        if (!B) { shape = mbt; pf0 = pf1 = 1; }
        else if (mbt <= 3) { shape = 0; pf0 = pf1 = mbt; }
        else {
          shape = (mbt & 1) ? 2 : 1;
          int k = (mbt - 4) >> 1;
          pf0 = b_part_pred[k * 2]; pf1 = b_part_pred[k * 2 + 1];
        }
        ng = shape == 0 ? 1 : 2;
        pf[0] = (int8_t)pf0; pf[1] = (int8_t)pf1;
        if (shape == 0) s.grp[0] = (uint16_t)(3 << 4 | 3 << 6);
        else if (shape == 1) { s.grp[0] = (uint16_t)(3 << 4 | 1 << 6 | 1 << 10); s.grp[1] = (uint16_t)(2 << 2 | 3 << 4 | 1 << 6 | 2 << 10); }
        else { s.grp[0] = (uint16_t)(1 << 4 | 3 << 6 | 3 << 10); s.grp[1] = (uint16_t)(2 | 1 << 4 | 3 << 6 | 4 << 10); }
      }
      if (ng) {
        int8_t (*refs)[4] = s.refs;
        const bool ref0_only = !B && mbt == 4 && !HWB_IS_CABAC(s);  // P_8x8ref0 (CAVLC only)
        // reference indices, list by list; each is made visible at once for the ref_idx contexts of the next ones
#pragma unroll 1
        for (int l = 0; l < nl; ++l)
#pragma unroll 1
          for (int g = 0; g < ng; ++g) {
            const int e = s.grp[g], bx = e & 3, by = (e >> 2) & 3, w = ((e >> 4) & 3) + 1, h = ((e >> 6) & 3) + 1;
            int r = -1;
            if (ng == 4 && ((dirq >> g) & 1)) { refs[l][g] = -1; continue; }
            if (pf[g] & (1 << l)) {
              r = ref0_only ? 0 : read_ref(s, l, bx, by);
              if (r >= sd.num_ref[l]) { sd_fail(s, 54); return; }
            }
            refs[l][g] = (int8_t)r;
            if (ng > 1) set_refs(s, l, bx, by, w, h, r >= 0 ? r : REF_NONE);  // a single 16x16 partition: nothing inside the macroblock looks at it
          }
        // motion vectors; references of not-yet-decoded groups must look unavailable for C-neighbour lookups
#pragma unroll 1
        for (int l = 0; l < nl; ++l) {
          if (ng > 1) set_refs(s, l, 0, 0, 4, 4, REF_UNAVAIL);
#pragma unroll 1
          for (int g = 0; g < ng; ++g) {
            const int e = s.grp[g], bx = e & 3, by = (e >> 2) & 3, w = ((e >> 4) & 3) + 1, h = ((e >> 6) & 3) + 1;
            if (ng == 4 && ((dirq >> g) & 1)) { apply_direct(s, l, g, dref, dmv); continue; }
            const int r = refs[l][g];
            if (r < 0) { set_motion(s, l, bx, by, w, h, REF_NONE, 0, 0, 0, 0); continue; }
            const int shp = (e >> 8) & 3, pshape = (e >> 10) & 7;
            const int pw = (shp & 2) ? 1 : w, ph = (shp & 1) ? 1 : h;  // sub-shapes only occur on 8x8 groups (w == h == 2)
            const int n = (w * h) / (pw * ph);
#pragma unroll 1
            for (int k = 0; k < n; ++k) {
              const int dx = pw < w ? (k & 1) : 0, dy = ph < h ? (pw < w ? k >> 1 : k) : 0;
              read_mvd_and_set(s, l, bx + dx, by + dy, pw, ph, r, pshape);
            }
          }
        }
      }
      // ---------------- cbp, transform size, qp delta, residual
      if (HWB_IS_CABAC(s)) cbp = cabac_cbp(s);
      else { uint32_t k = s_ue(s); if (k > 47) { sd_fail(s, 56); return; } cbp = golomb_to_inter_cbp[k]; }
      o.cbp = (uint8_t)cbp;
      bool t8 = false;
      if (HWB_T8_ON && (cbp & 15) && s.pd->transform8x8_mode && t8_allowed) {
        if (HWB_IS_CABAC(s)) {
          int ctx = (s.availA && (s.left.flags & NBF_T8)) + (s.availB && (top_flags(s) & NBF_T8));
          t8 = cabac_bin(s, 399 + ctx) != 0;
        } else t8 = s_get(s, 1) != 0;
      }
      if (t8) o.flags |= MBF_T8x8;
      if (cbp) {
        int dqp = HWB_IS_CABAC(s) ? cabac_dqp(s) : s_se(s);
        s.last_dqp = dqp;
        s.qp = (s.qp + dqp + 52) % 52;
      } else s.last_dqp = 0;
      o.qp = (uint8_t)s.qp;
      decode_residual(s, false, cbp, t8);
    }
  }

  finish_mb(s, skipped, direct16, is_pcm);
}

// ================================================================================ slice
// A slice is decoded in three steps so that the CUDA kernels can interleave the macroblocks of the warps of a block
// (kernels.cu, "lockstep"): slice_begin, slice_step once per macroblock until it returns false, slice_end.  The loop's
// state lives in the slice state (s.it_*).
HWB_FN void slice_begin(const ChunkCtx &c, int slice_idx, SliceDec *state) {
  SliceDec &s = *state;
  s.c = &c; s.sd = &c.slices[slice_idx]; s.pd = &c.pics[s.sd->pic];
  s.slice_num = slice_idx - s.pd->first_slice;
  s.cabac = s.pd->cabac != 0;
  s.error = 0;
  const SliceDesc &sd = *s.sd;
  const uint8_t *data = c.bitstream + sd.data_off;
  br_init(s.br, data, sd.data_size, sd.bit_off);
  // position of the rbsp_stop_one_bit
  {
    int n = (int)sd.data_size;
    while (n > 0 && data[n - 1] == 0) --n;
    uint32_t last = n > 0 ? data[n - 1] : 0x80;
    int tz = 0;
    while (!((last >> tz) & 1)) ++tz;
    s.stop_bitpos = (uint32_t)(n > 0 ? (n - 1) * 8 + (7 - tz) : 0);
  }
  s.qp = sd.qp; s.last_dqp = 0; s.row_reach = 0; s.row_reach_x = 0;
  init_caches(s);
  init_lane_tables(s);
  s.line = (NbCtx *)(c.ectx + (uint64_t)slice_idx * c.ectx_stride);
  // the arena region of a slice starts at its first macroblock's worst-case offset
  s.coef_next = (uint32_t)sd.first_mb * SLOTS_PER_MB;
  if (HWB_IS_CABAC(s)) {
    const int table = sd.slice_type == SLICE_I ? 0 : 1 + sd.cabac_init_idc;
    const CtxE *ft = (const CtxE *)cabac_fused;
    HWB_LANES(l)
    for (int i = l; i < 128; i += 32) ctxe_store(s.fused + i, ctxe_load(ft + i));
    for (int i = l; i < HWB_CABAC_NCTX; i += 32) ctxe_store(s.ctxe + i, ctxe_load(ft + cabac_init_state(table, sd.qp, i)));
    HWB_LANES_END
    cabac_start(s.cab, data, (sd.bit_off + 7) >> 3);  // cabac_alignment_one_bits, then 9 bits of codIOffset
  }
  s.it_slice = slice_idx;
  s.it_first = sd.first_mb;
  s.it_addr = sd.first_mb;
  s.mbx = sd.first_mb % c.mb_w; s.mby = sd.first_mb / c.mb_w;
  s.it_end = false;
  // CAVLC mb_skip_run state: -1 = read a new run before the next macroblock, 0 = the next
  // macroblock is coded, >0 = macroblocks still to skip
  s.it_run = -1;
  s.it_end_mb = sd.end_mb < c.nmb ? sd.end_mb : c.nmb;  // the slice must cover [first_mb, end_mb) exactly
}

// One macroblock.  Returns false when the slice is over (end of slice, error, or its last macroblock done).
HWB_FN bool slice_step(SliceDec *state) {
  SliceDec &s = *state;
  const ChunkCtx &c = *s.c;
  const SliceDesc &sd = *s.sd;
  const int first = s.it_first, end_mb = s.it_end_mb;
  int addr = s.it_addr;
  if (s.it_end || addr >= end_mb) return false;
  s.mbaddr = addr;
  s.availA = s.mbx > 0 && addr - 1 >= first;
  s.availB = addr - c.mb_w >= first;
  s.availC = s.mbx < c.mb_w - 1 && addr - c.mb_w + 1 >= first;
  s.availD = s.mbx > 0 && addr - c.mb_w - 1 >= first;
  fill_caches(s, false);  // also brings the top neighbour's flags into the slice state (mb_skip_flag context)
  bool skipped = false;
  if (sd.slice_type != SLICE_I) {
    if (HWB_IS_CABAC(s)) {
      int ctx = (s.availA && !(s.left.flags & NBF_SKIP)) + (s.availB && !(top_flags(s) & NBF_SKIP));
      s.pre_mbt = HWB_IS_B(sd.slice_type) ? cabac_b_header(s, ctx) : cabac_p_header(s, ctx);
      skipped = s.pre_mbt < 0;
    } else {
      if (s.it_run < 0) {
        s.it_run = (int)s_ue(s);
        if (s.it_run > end_mb - addr) { sd_fail(s, 60); return false; }
      }
      if (s.it_run > 0) { skipped = true; s.it_run--; }
    }
  }
  decode_mb(s, skipped);
  if (HWB_IS_CABAC(s) && s.cab.pos > sd.data_size + 8) s.br.overrun = true;  // the engine legitimately reads a few bytes ahead
  if (s.error || s.br.overrun) return false;
  bool end = false;
  if (HWB_IS_CABAC(s)) {
    end = cabac_term(s) != 0;
  } else if (skipped) {
    if (s.it_run == 0 && !br_more_rbsp_data(s.br, s.stop_bitpos)) end = true;
  } else {
    s.it_run = -1;
    if (!br_more_rbsp_data(s.br, s.stop_bitpos)) end = true;
  }
  s.it_end = end;
  addr++;
  s.it_addr = addr;
  int mbx = s.mbx + 1;
  if (mbx == c.mb_w) { mbx = 0; s.mby++; }
  s.mbx = mbx;
  if (mbx == 0 || end || addr == end_mb) {
    // the row just completed (or the part of it this slice covers): what the picture kernel must wait for in the
    // reference pictures before it predicts this row; several slices may share a row, hence the atomic maximum
    const int slice_idx = s.it_slice;
    int32_t *rr = c.mv_reach + (size_t)sd.pic * c.mb_h + (addr - 1) / c.mb_w;
    int32_t *rx = c.mv_reach_x + (size_t)sd.pic * c.mb_h + (addr - 1) / c.mb_w;
#if HWB_DEVICE_BUILD
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {  // release store: no L1 invalidation (see publish_progress in kernels.cu)
      if (s.row_reach) { atomicMax(rr, s.row_reach); atomicMax(rx, s.row_reach_x); }
      const int32_t v = (end || addr == end_mb) ? c.nmb : addr;
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(c.entropy_prog + slice_idx), "r"(v) : "memory");
    }
    __syncwarp();
#else
    if (s.row_reach > *rr) *rr = s.row_reach;
    if (s.row_reach_x > *rx) *rx = s.row_reach_x;
    c.entropy_prog[slice_idx] = (end || addr == end_mb) ? c.nmb : addr;
#endif
    s.row_reach = 0; s.row_reach_x = 0;
  }
  return !end && addr < end_mb;
}

HWB_FN void slice_end(SliceDec *state) {
  SliceDec &s = *state;
  const ChunkCtx &c = *s.c;
  const int slice_idx = s.it_slice;
  // A slice that stops before the next slice's first macroblock (or runs out of data at it without its end flag)
  // would leave macroblock records of the picture undefined: the picture is refused (error 61), nothing of the chunk
  // is reconstructed.  In CABAC mode end_of_slice_flag must also have been seen at the last macroblock.
  if (!s.error && !s.br.overrun && (s.it_addr != s.it_end_mb || (HWB_IS_CABAC(s) && !s.it_end))) s.error = 61;
  if (s.error || s.br.overrun) {
#if HWB_DEVICE_BUILD
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      atomicExch(c.error_flag, s.error ? s.error : 99);
      __threadfence();
      *((volatile int32_t *)(c.entropy_prog + slice_idx)) = c.nmb;  // never leave consumers spinning
    }
#else
    *c.error_flag = s.error ? s.error : 99;
    c.entropy_prog[slice_idx] = c.nmb;
#endif
  }
}

HWB_HD void decode_slice(const ChunkCtx &c, int slice_idx, uint8_t *cabac_states, SliceDec *state) {
  (void)cabac_states;
  slice_begin(c, slice_idx, state);
#pragma unroll 1
  while (slice_step(state)) {}
  slice_end(state);
}

}  // namespace HWB_ENT_NS
#ifndef HWB_ENT_DEFAULT_IMPORTED
#define HWB_ENT_DEFAULT_IMPORTED
using namespace ent;
#endif
}  // namespace hwb
