// Macroblock-layer entropy *writer* (CAVLC and CABAC) for the synthetic stream generator.  It is
// the mirror image of hwb::decode_mb (hwang_b200/csrc/dev/entropy.h) and deliberately reuses the
// decoder's neighbour caches, MV prediction and context derivation so that generator and decoder
// cannot drift apart silently -- what keeps both honest is libavcodec decoding the result.
#pragma once
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include "bitwriter.h"
#include "../../hwang_b200/csrc/dev/entropy.h"

namespace gen {
using namespace hwb;

inline void CabacEnc::decision(int ctx, int bin) {
  uint32_t s = st[ctx], p = s >> 1, mps = s & 1;
  uint32_t rlps = cabac_range_lps[p * 4 + ((range >> 6) & 3)];
  range -= rlps;
  if ((uint32_t)bin != mps) {
    low += range; range = rlps;
    if (p == 0) mps ^= 1;
    st[ctx] = (uint8_t)((cabac_trans_lps[p] << 1) | mps);
  } else {
    st[ctx] = (uint8_t)(((p < 62 ? p + 1 : p) << 1) | mps);
  }
  renorm();
}

// What the generator decided for one macroblock (syntax level).
struct MbEnc {
  bool skipped = false;
  int mbt = 0;    // mb_type as coded for the slice type
  int imbt = -1;  // intra mb_type 0..25, or -1
  bool t8 = false;
  int cmode = 0;
  int i4modes[16];
  int sub[4] = {0, 0, 0, 0};
  int ref[2][4];
  int mv[2][16][2];
  int cbp = 0;
  int dqp = 0;
  uint8_t pcm[384];
};

struct SliceEnc {
  SliceDec s;  // neighbour caches / context holder, shared with the decoder
  BitWriter bw;
  CabacEnc ce;
  bool cabac = false;
  int run = 0;  // pending CAVLC mb_skip_run
};

// ------------------------------------------------------------------ CAVLC residual writer
// c[]: max_coeff coefficients in scan order.
inline int cavlc_write_block(SliceEnc &e, const int *c, int max_coeff, int nC) {
  BitWriter &bw = e.bw;
  int total = 0, last = -1;
  for (int i = 0; i < max_coeff; ++i) if (c[i]) { total++; last = i; }
  int t1 = 0;
  if (total) {
    for (int i = last; i >= 0 && t1 < 3; --i) {
      if (!c[i]) continue;
      if (c[i] == 1 || c[i] == -1) t1++; else break;
    }
  }
  if (nC < 0) bw.put(cavlc_enc_chroma_dc_token_bits[total * 4 + t1], cavlc_enc_chroma_dc_token_len[total * 4 + t1]);
  else {
    int tab = nC < 2 ? 0 : (nC < 4 ? 1 : (nC < 8 ? 2 : 3));
    int len = cavlc_enc_coeff_token_len[tab * 68 + total * 4 + t1];
    assert(len > 0);
    bw.put(cavlc_enc_coeff_token_bits[tab * 68 + total * 4 + t1], len);
  }
  if (!total) return 0;
  int lev[16], runs[16], n = 0, zeros = 0, prev = last + 1;
  for (int i = last; i >= 0; --i)
    if (c[i]) { lev[n] = c[i]; if (n) runs[n - 1] = prev - i - 1; prev = i; n++; }
  runs[n - 1] = prev;  // zeros before the lowest-frequency coefficient (not coded)
  zeros = last + 1 - total;
  int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
  for (int i = 0; i < total; ++i) {
    if (i < t1) { bw.put1(lev[i] < 0); continue; }
    int code = lev[i] > 0 ? 2 * lev[i] - 2 : -2 * lev[i] - 1;
    if (i == t1 && t1 < 3) code -= 2;
    if (suffix_len == 0) {
      if (code < 14) { bw.put(1, code + 1); }
      else if (code < 30) { bw.put(1, 15); bw.put((uint32_t)(code - 14), 4); }
      else { assert(code - 30 < 4096); bw.put(1, 16); bw.put((uint32_t)(code - 30), 12); }
    } else {
      if (code < (15 << suffix_len)) { bw.put(1, (code >> suffix_len) + 1); bw.put((uint32_t)(code & ((1 << suffix_len) - 1)), suffix_len); }
      else { assert(code - (15 << suffix_len) < 4096); bw.put(1, 16); bw.put((uint32_t)(code - (15 << suffix_len)), 12); }
    }
    if (suffix_len == 0) suffix_len = 1;
    if (abs(lev[i]) > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
  }
  if (total < max_coeff) {
    if (nC < 0) bw.put(cavlc_enc_chroma_dc_total_zeros_bits[(total - 1) * 4 + zeros], cavlc_enc_chroma_dc_total_zeros_len[(total - 1) * 4 + zeros]);
    else bw.put(cavlc_enc_total_zeros_bits[(total - 1) * 16 + zeros], cavlc_enc_total_zeros_len[(total - 1) * 16 + zeros]);
  }
  int zl = zeros;
  for (int i = 0; i < total - 1 && zl > 0; ++i) {
    int r = runs[i];
    int tab = zl < 7 ? zl - 1 : 6;
    bw.put(cavlc_enc_run_bits[tab * 16 + r], cavlc_enc_run_len[tab * 16 + r]);
    zl -= r;
  }
  return total;
}

// ------------------------------------------------------------------ CABAC residual writer
inline int cabac_write_block(SliceEnc &e, const int *c, int cat, int max_coeff) {
  const int sig_off = cat == 0 ? 105 : cat == 1 ? 120 : cat == 2 ? 134 : cat == 3 ? 149 : cat == 4 ? 152 : 402;
  const int last_off = cat == 0 ? 166 : cat == 1 ? 181 : cat == 2 ? 195 : cat == 3 ? 210 : cat == 4 ? 213 : 417;
  const int abs_off = cat == 0 ? 227 : cat == 1 ? 237 : cat == 2 ? 247 : cat == 3 ? 257 : cat == 4 ? 266 : 426;
  int last = -1, n = 0;
  for (int i = 0; i < max_coeff; ++i) if (c[i]) { last = i; n++; }
  assert(n > 0);
  for (int i = 0; i < max_coeff - 1; ++i) {
    int sctx = cat == 5 ? cabac_sig8x8_ctx[i] : (cat == 3 ? (i < 2 ? i : 2) : i);
    e.ce.decision(sig_off + sctx, c[i] != 0);
    if (c[i]) {
      int lctx = cat == 5 ? cabac_last8x8_ctx[i] : (cat == 3 ? (i < 2 ? i : 2) : i);
      e.ce.decision(last_off + lctx, i == last);
      if (i == last) break;
    }
  }
  int eq1 = 0, gt1 = 0;
  for (int i = last; i >= 0; --i) {
    if (!c[i]) continue;
    int a = abs(c[i]);
    int ctx0 = gt1 ? 0 : (eq1 < 3 ? 1 + eq1 : 4);
    if (a == 1) { e.ce.decision(abs_off + ctx0, 0); eq1++; }
    else {
      e.ce.decision(abs_off + ctx0, 1);
      int cmax = cat == 3 ? 3 : 4;
      int ctx1 = 5 + (gt1 < cmax ? gt1 : cmax);
      int v = 2;
      while (v < 15 && v < a) { e.ce.decision(abs_off + ctx1, 1); v++; }
      if (a < 15) e.ce.decision(abs_off + ctx1, 0);
      else {
        int rem = a - 15, k = 0;
        while (rem >= (1 << k)) { e.ce.bypass(1); rem -= 1 << k; k++; }
        e.ce.bypass(0);
        while (k--) e.ce.bypass((rem >> k) & 1);
      }
      gt1++;
    }
    e.ce.bypass(c[i] < 0);
  }
  return n;
}

// ------------------------------------------------------------------ residual (mirrors decode_residual)
// Coefficients are read back from the arena slots written by the generator (raw levels, raster order).
inline void write_residual(SliceEnc &e, bool i16, int cbp, bool t8) {
  SliceDec &s = e.s;
  const bool cabac = e.cabac;
  const MbInfo &o = s.out;
  const bool intra = o.mbtype != MB_INTER;
  const int cbf_unavail = intra ? 1 : 0;
  const int16_t *arena = pic_coefs(*s.c, s.pd->frame) + (uint64_t)o.coef_off * 16;
  auto slot = [&](int bit) -> const int16_t * {
    return ((o.nzmask >> bit) & 1) ? arena + 16 * popc32(o.nzmask & ((1u << bit) - 1)) : nullptr;
  };
  int c[64];
  if (i16) {
    const int16_t *sl = slot(NZ_LUMA_DC);
    for (int i = 0; i < 16; ++i) c[i] = sl ? sl[zigzag4x4[i]] : 0;
    if (cabac) {
      int a = s.availA ? ((s.left.flags & NBF_IPCM) ? 1 : (s.left.cbf >> NZ_LUMA_DC) & 1) : cbf_unavail;
      int bq = s.availB ? ((s.line[s.mbx].flags & NBF_IPCM) ? 1 : (s.line[s.mbx].cbf >> NZ_LUMA_DC) & 1) : cbf_unavail;
      e.ce.decision(85 + a + 2 * bq, sl != nullptr);
      if (sl) cabac_write_block(e, c, 0, 16);
    } else cavlc_write_block(e, c, 16, cavlc_nc(s.nz_cache[HWB_CI(-1, 0)], s.nz_cache[HWB_CI(0, -1)]));
  }
  for (int q = 0; q < 4; ++q) {
    if (!((cbp >> q) & 1)) continue;
    if (t8) {
      const int16_t *sl = slot(NZ_LUMA0 + q * 4);
      if (cabac) {
        assert(sl);
        for (int i = 0; i < 64; ++i) c[i] = sl[zigzag8x8[i]];
        int n = cabac_write_block(e, c, 5, 64);
        for (int k = 0; k < 4; ++k) {
          int bx = (q & 1) * 2 + (k & 1), by = (q >> 1) * 2 + (k >> 1);
          s.nz_cache[HWB_CI(bx, by)] = (uint8_t)(n > 16 ? 16 : n);
        }
      } else {
        for (int k = 0; k < 4; ++k) {
          int bx = (q & 1) * 2 + (k & 1), by = (q >> 1) * 2 + (k >> 1);
          for (int i = 0; i < 16; ++i) c[i] = sl ? sl[zigzag8x8[4 * i + k]] : 0;
          int m = cavlc_write_block(e, c, 16, cavlc_nc(s.nz_cache[HWB_CI(bx - 1, by)], s.nz_cache[HWB_CI(bx, by - 1)]));
          s.nz_cache[HWB_CI(bx, by)] = (uint8_t)m;
        }
      }
    } else {
      for (int k = 0; k < 4; ++k) {
        int z = q * 4 + k, bx = z2x(z), by = z2y(z);
        int na = s.nz_cache[HWB_CI(bx - 1, by)], nb = s.nz_cache[HWB_CI(bx, by - 1)];
        const int16_t *sl = slot(NZ_LUMA0 + z);
        int st0 = i16 ? 1 : 0, mc = i16 ? 15 : 16, n = 0;
        for (int i = 0; i < mc; ++i) c[i] = sl ? sl[zigzag4x4[st0 + i]] : 0;
        if (cabac) {
          int a = na == 0x80 ? cbf_unavail : (na != 0), bq = nb == 0x80 ? cbf_unavail : (nb != 0);
          e.ce.decision(85 + (i16 ? 4 : 8) + a + 2 * bq, sl != nullptr);
          if (sl) n = cabac_write_block(e, c, i16 ? 1 : 2, mc);
        } else n = cavlc_write_block(e, c, mc, cavlc_nc(na, nb));
        s.nz_cache[HWB_CI(bx, by)] = (uint8_t)n;
      }
    }
  }
  if (cbp & 0x30) {
    for (int p = 0; p < 2; ++p) {
      int bit = p ? NZ_CR_DC : NZ_CB_DC;
      const int16_t *sl = slot(bit);
      for (int i = 0; i < 4; ++i) c[i] = sl ? sl[i] : 0;
      if (cabac) {
        int a = s.availA ? ((s.left.flags & NBF_IPCM) ? 1 : (s.left.cbf >> bit) & 1) : cbf_unavail;
        int bq = s.availB ? ((s.line[s.mbx].flags & NBF_IPCM) ? 1 : (s.line[s.mbx].cbf >> bit) & 1) : cbf_unavail;
        e.ce.decision(85 + 12 + a + 2 * bq, sl != nullptr);
        if (sl) cabac_write_block(e, c, 3, 4);
      } else cavlc_write_block(e, c, 4, -1);
    }
  }
  if (cbp & 0x20) {
    for (int p = 0; p < 2; ++p)
      for (int k = 0; k < 4; ++k) {
        int bx = k & 1, by = k >> 1;
        int na = s.cnz_cache[p][(by + 1) * 4 + bx], nb = s.cnz_cache[p][by * 4 + bx + 1];
        const int16_t *sl = slot((p ? NZ_CR0 : NZ_CB0) + k);
        int n = 0;
        for (int i = 0; i < 15; ++i) c[i] = sl ? sl[zigzag4x4[1 + i]] : 0;
        if (cabac) {
          int a = na == 0x80 ? cbf_unavail : (na != 0), bq = nb == 0x80 ? cbf_unavail : (nb != 0);
          e.ce.decision(85 + 16 + a + 2 * bq, sl != nullptr);
          if (sl) n = cabac_write_block(e, c, 4, 15);
        } else n = cavlc_write_block(e, c, 15, cavlc_nc(na, nb));
        s.cnz_cache[p][(by + 1) * 4 + bx + 1] = (uint8_t)n;
      }
  }
}

// ------------------------------------------------------------------ CABAC syntax element writers
inline void w_intra_mb_type(SliceEnc &e, int base, bool islice, int t) {
  SliceDec &s = e.s;
  int st = base;
  if (islice) {
    int ctx = 0;
    if (s.availA && !(s.left.flags & NBF_INXN)) ctx++;
    if (s.availB && !(s.line[s.mbx].flags & NBF_INXN)) ctx++;
    e.ce.decision(st + ctx, t != 0);
    if (!t) return;
    st += 2;
  } else {
    e.ce.decision(st, t != 0);
    if (!t) return;
  }
  if (t == 25) { e.ce.terminate(1); return; }
  e.ce.terminate(0);
  int t1 = t - 1, luma = t1 / 12, chroma = (t1 / 4) % 3, pm = t1 % 4;
  e.ce.decision(st + 1, luma);
  e.ce.decision(st + 2, chroma != 0);
  if (chroma) e.ce.decision(st + 2 + (islice ? 1 : 0), chroma == 2);
  e.ce.decision(st + 3 + (islice ? 1 : 0), pm >> 1);
  e.ce.decision(st + 3 + (islice ? 2 : 0), pm & 1);
}

inline void w_b_mb_type(SliceEnc &e, int mbt, int imbt) {
  SliceDec &s = e.s;
  int ctx = 0;
  if (s.availA && !(s.left.flags & NBF_DIRECT16)) ctx++;
  if (s.availB && !(s.line[s.mbx].flags & NBF_DIRECT16)) ctx++;
  if (mbt == 0) { e.ce.decision(27 + ctx, 0); return; }
  e.ce.decision(27 + ctx, 1);
  if (mbt <= 2) { e.ce.decision(27 + 3, 0); e.ce.decision(27 + 5, mbt - 1); return; }
  e.ce.decision(27 + 3, 1);
  int bits4, extra = -1;
  if (mbt >= 23) bits4 = 13;
  else if (mbt == 11) bits4 = 14;
  else if (mbt == 22) bits4 = 15;
  else if (mbt <= 10) bits4 = mbt - 3;
  else { int b5 = mbt + 4; bits4 = b5 >> 1; extra = b5 & 1; }
  e.ce.decision(27 + 4, (bits4 >> 3) & 1);
  e.ce.decision(27 + 5, (bits4 >> 2) & 1);
  e.ce.decision(27 + 5, (bits4 >> 1) & 1);
  e.ce.decision(27 + 5, bits4 & 1);
  if (extra >= 0) e.ce.decision(27 + 5, extra);
  if (mbt >= 23) w_intra_mb_type(e, 32, false, imbt);
}

inline void w_b_sub_type(SliceEnc &e, int t) {
  if (t == 0) { e.ce.decision(36, 0); return; }
  e.ce.decision(36, 1);
  if (t <= 2) { e.ce.decision(37, 0); e.ce.decision(39, t - 1); return; }
  e.ce.decision(37, 1);
  if (t <= 6) { e.ce.decision(38, 0); int v = t - 3; e.ce.decision(39, v >> 1); e.ce.decision(39, v & 1); }
  else if (t <= 10) { e.ce.decision(38, 1); e.ce.decision(39, 0); int v = t - 7; e.ce.decision(39, v >> 1); e.ce.decision(39, v & 1); }
  else { e.ce.decision(38, 1); e.ce.decision(39, 1); e.ce.decision(39, t - 11); }
}

inline void w_ref(SliceEnc &e, int l, int bx, int by, int ref) {
  SliceDec &s = e.s;
  int nref = s.sd->num_ref[l];
  if (nref <= 1) { assert(ref == 0); return; }
  if (!e.cabac) {
    if (nref == 2) e.bw.put1(ref ^ 1); else e.bw.ue((uint32_t)ref);
    return;
  }
  int ra = s.ref_cache[l][HWB_CI(bx - 1, by)], rb = s.ref_cache[l][HWB_CI(bx, by - 1)];
  int ctx = 0;
  if (ra > 0 && !s.dir_cache[HWB_CI(bx - 1, by)]) ctx++;
  if (rb > 0 && !s.dir_cache[HWB_CI(bx, by - 1)]) ctx += 2;
  for (int i = 0; i < ref; ++i) { e.ce.decision(54 + ctx, 1); ctx = (ctx >> 2) + 4; }
  e.ce.decision(54 + ctx, 0);
}

inline int w_mvd_comp(SliceEnc &e, int base, int amvd, int mvd) {
  int a = abs(mvd);
  int inc = amvd < 3 ? 0 : (amvd > 32 ? 2 : 1);
  if (!a) { e.ce.decision(base + inc, 0); return 0; }
  e.ce.decision(base + inc, 1);
  int ctx = base + 3, m = 1;
  while (m < 9 && m < a) { e.ce.decision(ctx, 1); if (m < 4) ctx++; m++; }
  if (a < 9) e.ce.decision(ctx, 0);
  else {
    int v = a - 9, k = 3;
    while (v >= (1 << k)) { e.ce.bypass(1); v -= 1 << k; k++; }
    e.ce.bypass(0);
    while (k--) e.ce.bypass((v >> k) & 1);
  }
  e.ce.bypass(mvd < 0);
  return a < 70 ? a : 70;
}

// Writes the mvd of a partition whose final MV is (mx,my) and updates the caches.
inline void w_mvd_and_set(SliceEnc &e, int l, int bx, int by, int w, int h, int ref, int shape, int mx, int my) {
  SliceDec &s = e.s;
  int px, py;
  pred_mv(s, l, bx, by, w, ref, shape, px, py);
  int dx = mx - px, dy = my - py, ax = 0, ay = 0;
  if (e.cabac) {
    int sa = s.mvd_cache[l][HWB_CI(bx - 1, by)][0] + s.mvd_cache[l][HWB_CI(bx, by - 1)][0];
    int sb = s.mvd_cache[l][HWB_CI(bx - 1, by)][1] + s.mvd_cache[l][HWB_CI(bx, by - 1)][1];
    ax = w_mvd_comp(e, 40, sa, dx);
    ay = w_mvd_comp(e, 47, sb, dy);
  } else { e.bw.se(dx); e.bw.se(dy); }
  set_motion(s, l, bx, by, w, h, ref, mx, my, ax, ay);
}

inline void w_cbp(SliceEnc &e, int cbp, bool intra) {
  SliceDec &s = e.s;
  if (!e.cabac) {
    const uint8_t *tab = intra ? golomb_to_intra_cbp : golomb_to_inter_cbp;
    for (int k = 0; k < 48; ++k) if (tab[k] == cbp) { e.bw.ue((uint32_t)k); return; }
    assert(0);
  }
  const LeftCtx &L = s.left;
  const NbCtx &T = s.line[s.mbx];
  int cbpa = s.availA ? ((L.flags & NBF_IPCM) ? 0x2F : L.cbp) : 0x0F;
  int cbpb = s.availB ? ((T.flags & NBF_IPCM) ? 0x2F : T.cbp) : 0x0F;
  for (int b8 = 0; b8 < 4; ++b8) {
    int a = (b8 & 1) ? !((cbp >> (b8 - 1)) & 1) : !((cbpa >> (b8 + 1)) & 1);
    int bq = (b8 & 2) ? !((cbp >> (b8 - 2)) & 1) : !((cbpb >> (b8 + 2)) & 1);
    e.ce.decision(73 + a + 2 * bq, (cbp >> b8) & 1);
  }
  int ca = s.availA ? (cbpa >> 4) & 3 : 0, cb = s.availB ? (cbpb >> 4) & 3 : 0;
  int cc = cbp >> 4;
  e.ce.decision(77 + (ca > 0) + 2 * (cb > 0), cc > 0);
  if (cc) e.ce.decision(77 + 4 + (ca == 2) + 2 * (cb == 2), cc == 2);
}

inline void w_dqp(SliceEnc &e, int dqp) {
  SliceDec &s = e.s;
  if (!e.cabac) { e.bw.se(dqp); return; }
  int val = dqp > 0 ? 2 * dqp - 1 : -2 * dqp;
  int ctx = s.last_dqp != 0;
  for (int i = 0; i < val; ++i) { e.ce.decision(60 + ctx, 1); ctx = 2 + (ctx >> 1); }
  e.ce.decision(60 + ctx, 0);
}

inline void w_t8flag(SliceEnc &e, bool t8) {
  SliceDec &s = e.s;
  if (!e.cabac) { e.bw.put1(t8); return; }
  int ctx = (s.availA && (s.left.flags & NBF_T8)) + (s.availB && (s.line[s.mbx].flags & NBF_T8));
  e.ce.decision(399 + ctx, t8);
}

// ------------------------------------------------------------------ macroblock writer (mirrors decode_mb)
// Precondition: caller set s.mbx/mby/mbaddr/avail*, called fill_caches(s), filled s.out (the MbInfo
// that reconstruction used) and the arena slots.  Emits the syntax and publishes the neighbour context.
inline void encode_mb(SliceEnc &e, const MbEnc &m) {
  SliceDec &s = e.s;
  const SliceDesc &sd = *s.sd;
  const int st = sd.slice_type;
  const bool B = st == SLICE_B;
  const int nl = B ? 2 : 1;
  MbInfo &o = s.out;
  bool direct16 = false, is_pcm = false;
  uint32_t dirq = 0;
  int8_t dref[2][4];
  int16_t dmv[2][16][2];

  // ---- skip signalling
  if (st != SLICE_I) {
    if (e.cabac) {
      int ctx = (s.availA && !(s.left.flags & NBF_SKIP)) + (s.availB && !(s.line[s.mbx].flags & NBF_SKIP));
      e.ce.decision((B ? 24 : 11) + ctx, m.skipped);
    } else {
      if (m.skipped) e.run++;
      else { e.bw.ue((uint32_t)e.run); e.run = 0; }
    }
  }
  if (m.skipped) {
    s.last_dqp = 0;
    if (!B) {
      int mx = 0, my = 0;
      MvRef A = mv_at(s, 0, -1, 0), Bn = mv_at(s, 0, 0, -1);
      if (!(A.ref == REF_UNAVAIL || Bn.ref == REF_UNAVAIL || (A.ref == 0 && A.mx == 0 && A.my == 0) || (Bn.ref == 0 && Bn.mx == 0 && Bn.my == 0)))
        pred_mv(s, 0, 0, 0, 4, 0, 0, mx, my);
      assert(mx == m.mv[0][0][0] && my == m.mv[0][0][1]);
      set_motion(s, 0, 0, 0, 4, 4, 0, mx, my, 0, 0);
    } else {
      direct16 = true;
      direct_predict(s, 15, dref, dmv);
      for (int l = 0; l < 2; ++l) for (int q = 0; q < 4; ++q) apply_direct(s, l, q, dref, dmv);
    }
  } else {
    // ---- mb_type
    if (e.cabac) {
      if (st == SLICE_I) w_intra_mb_type(e, 3, true, m.imbt);
      else if (st == SLICE_P) {
        if (m.imbt < 0) {
          e.ce.decision(14, 0);
          if (m.mbt == 0 || m.mbt == 3) { e.ce.decision(15, 0); e.ce.decision(16, m.mbt == 3); }
          else { e.ce.decision(15, 1); e.ce.decision(17, m.mbt == 1); }
        } else { e.ce.decision(14, 1); w_intra_mb_type(e, 17, false, m.imbt); }
      } else w_b_mb_type(e, m.mbt, m.imbt);
    } else e.bw.ue((uint32_t)m.mbt);

    if (m.imbt == 25) {
      is_pcm = true;
      if (e.cabac) { /* terminate(1) already flushed the arithmetic coder */ }
      e.bw.align_zero();
      for (int i = 0; i < 384; ++i) e.bw.put(m.pcm[i], 8);
      if (e.cabac) e.ce.start(&e.bw);
      for (int y = 0; y < 4; ++y) set4(s.nz_cache + HWB_CI(0, y), 0x10101010u);
      for (int p = 0; p < 2; ++p) { s.cnz_cache[p][5] = s.cnz_cache[p][6] = s.cnz_cache[p][9] = s.cnz_cache[p][10] = 16; }
      s.last_dqp = 0;
    } else if (m.imbt >= 0) {
      if (m.imbt == 0) {
        if (s.pd->transform8x8_mode) w_t8flag(e, m.t8);
        const int nb = m.t8 ? 4 : 16;
        for (int k = 0; k < nb; ++k) {
          int bx = m.t8 ? (k & 1) * 2 : z2x(k), by = m.t8 ? (k >> 1) * 2 : z2y(k);
          int ma = s.im_cache[HWB_CI(bx - 1, by)], mb_ = s.im_cache[HWB_CI(bx, by - 1)];
          int pred = (ma < 0 || mb_ < 0) ? 2 : (ma < mb_ ? ma : mb_);
          int mode = m.i4modes[k];
          if (mode == pred) { if (e.cabac) e.ce.decision(68, 1); else e.bw.put1(1); }
          else {
            int rem = mode < pred ? mode : mode - 1;
            if (e.cabac) { e.ce.decision(68, 0); e.ce.decision(69, rem & 1); e.ce.decision(69, (rem >> 1) & 1); e.ce.decision(69, (rem >> 2) & 1); }
            else { e.bw.put1(0); e.bw.put((uint32_t)rem, 3); }
          }
          int wd = m.t8 ? 2 : 1;
          for (int y = by; y < by + wd; ++y) for (int x = bx; x < bx + wd; ++x) s.im_cache[HWB_CI(x, y)] = (int8_t)mode;
        }
      }
      if (e.cabac) {
        int ctx = 0;
        if (s.availA && s.left.cmode != 0) ctx++;
        if (s.availB && s.line[s.mbx].cmode != 0) ctx++;
        e.ce.decision(64 + ctx, m.cmode != 0);
        if (m.cmode) { e.ce.decision(67, m.cmode != 1); if (m.cmode != 1) e.ce.decision(67, m.cmode == 3); }
      } else e.bw.ue((uint32_t)m.cmode);
      if (m.imbt == 0) w_cbp(e, m.cbp, true);
      if (m.cbp || m.imbt > 0) { w_dqp(e, m.dqp); s.last_dqp = m.dqp; s.qp = (s.qp + m.dqp + 52) % 52; }
      else s.last_dqp = 0;
      assert(o.qp == s.qp);
      write_residual(e, m.imbt > 0, m.cbp, m.t8);
    } else {
      bool t8_allowed = true;
      if (B && m.mbt == 0) {
        direct16 = true; dirq = 15;
        direct_predict(s, 15, dref, dmv);
        for (int l = 0; l < 2; ++l) for (int q = 0; q < 4; ++q) apply_direct(s, l, q, dref, dmv);
        t8_allowed = s.pd->direct_8x8_inference != 0;
      } else if ((!B && m.mbt >= 3) || (B && m.mbt == 22)) {
        int shape[4], pf[4];
        for (int q = 0; q < 4; ++q) {
          int t = m.sub[q];
          if (e.cabac) {
            if (B) w_b_sub_type(e, t);
            else { if (t == 0) e.ce.decision(21, 1); else { e.ce.decision(21, 0); if (t == 1) e.ce.decision(22, 0); else { e.ce.decision(22, 1); e.ce.decision(23, t == 2); } } }
          } else e.bw.ue((uint32_t)t);
          if (B && t == 0) dirq |= 1u << q;
          if (!B) { shape[q] = t; pf[q] = 1; }
          else if (t == 0) { shape[q] = 0; pf[q] = 0; }
          else { shape[q] = t <= 3 ? 0 : (t >= 10 ? 3 : ((t & 1) ? 2 : 1)); pf[q] = t <= 3 ? t : (t >= 10 ? t - 9 : ((t - 4) >> 1) + 1); }
          if (shape[q] != 0) t8_allowed = false;
          if (B && t == 0 && !s.pd->direct_8x8_inference) t8_allowed = false;
        }
        if (dirq) direct_predict(s, (int)dirq, dref, dmv);
        const bool ref0_only = !B && m.mbt == 4;
        for (int l = 0; l < nl; ++l)
          for (int q = 0; q < 4; ++q) {
            if ((dirq >> q) & 1) continue;
            int r = (pf[q] & (1 << l)) ? m.ref[l][q] : -1;
            if (r >= 0 && !ref0_only) w_ref(e, l, (q & 1) * 2, (q >> 1) * 2, r);
            int bx = (q & 1) * 2, by = (q >> 1) * 2;
            for (int y = by; y < by + 2; ++y) for (int x = bx; x < bx + 2; ++x) s.ref_cache[l][HWB_CI(x, y)] = (int8_t)(r >= 0 ? r : REF_NONE);
          }
        for (int l = 0; l < nl; ++l) {
          for (int i = 0; i < 16; ++i) s.ref_cache[l][HWB_CI(i & 3, i >> 2)] = REF_UNAVAIL;
          for (int q = 0; q < 4; ++q) {
            int bx = (q & 1) * 2, by = (q >> 1) * 2;
            if ((dirq >> q) & 1) { apply_direct(s, l, q, dref, dmv); continue; }
            if (!(pf[q] & (1 << l))) { set_motion(s, l, bx, by, 2, 2, REF_NONE, 0, 0, 0, 0); continue; }
            int r = m.ref[l][q];
#define MV(x, y) m.mv[l][(y) * 4 + (x)][0], m.mv[l][(y) * 4 + (x)][1]
            switch (shape[q]) {
              case 0: w_mvd_and_set(e, l, bx, by, 2, 2, r, 0, MV(bx, by)); break;
              case 1: w_mvd_and_set(e, l, bx, by, 2, 1, r, 0, MV(bx, by)); w_mvd_and_set(e, l, bx, by + 1, 2, 1, r, 0, MV(bx, by + 1)); break;
              case 2: w_mvd_and_set(e, l, bx, by, 1, 2, r, 0, MV(bx, by)); w_mvd_and_set(e, l, bx + 1, by, 1, 2, r, 0, MV(bx + 1, by)); break;
              default: for (int k = 0; k < 4; ++k) w_mvd_and_set(e, l, bx + (k & 1), by + (k >> 1), 1, 1, r, 0, MV(bx + (k & 1), by + (k >> 1)));
            }
          }
        }
      } else {
        int shape, pf0, pf1;
        if (!B) { shape = m.mbt; pf0 = pf1 = 1; }
        else if (m.mbt <= 3) { shape = 0; pf0 = pf1 = m.mbt; }
        else { shape = (m.mbt & 1) ? 2 : 1; int k = (m.mbt - 4) >> 1; pf0 = b_part_pred[k * 2]; pf1 = b_part_pred[k * 2 + 1]; }
        const int np = shape == 0 ? 1 : 2;
        for (int l = 0; l < nl; ++l)
          for (int p = 0; p < np; ++p) {
            int pf = p ? pf1 : pf0;
            int bx = (shape == 2 && p) ? 2 : 0, by = (shape == 1 && p) ? 2 : 0, w = shape == 2 ? 2 : 4, h = shape == 1 ? 2 : 4;
            int r = (pf & (1 << l)) ? m.ref[l][(by >> 1) * 2 + (bx >> 1)] : -1;
            if (r >= 0) w_ref(e, l, bx, by, r);
            for (int y = by; y < by + h; ++y) for (int x = bx; x < bx + w; ++x) s.ref_cache[l][HWB_CI(x, y)] = (int8_t)(r >= 0 ? r : REF_NONE);
          }
        for (int l = 0; l < nl; ++l) {
          for (int i = 0; i < 16; ++i) s.ref_cache[l][HWB_CI(i & 3, i >> 2)] = REF_UNAVAIL;
          for (int p = 0; p < np; ++p) {
            int pf = p ? pf1 : pf0;
            int bx = (shape == 2 && p) ? 2 : 0, by = (shape == 1 && p) ? 2 : 0, w = shape == 2 ? 2 : 4, h = shape == 1 ? 2 : 4;
            if (!(pf & (1 << l))) { set_motion(s, l, bx, by, w, h, REF_NONE, 0, 0, 0, 0); continue; }
            int sh = shape == 0 ? 0 : (shape == 1 ? 1 + p : 3 + p);
            w_mvd_and_set(e, l, bx, by, w, h, m.ref[l][(by >> 1) * 2 + (bx >> 1)], sh, MV(bx, by));
          }
        }
#undef MV
      }
      w_cbp(e, m.cbp, false);
      if ((m.cbp & 15) && s.pd->transform8x8_mode && t8_allowed) w_t8flag(e, m.t8);
      else assert(!m.t8);
      if (m.cbp) { w_dqp(e, m.dqp); s.last_dqp = m.dqp; s.qp = (s.qp + m.dqp + 52) % 52; }
      else s.last_dqp = 0;
      assert(o.qp == s.qp);
      write_residual(e, false, m.cbp, m.t8);
    }
  }
  // the motion data the caches now hold must be exactly what reconstruction used
  if (o.mbtype == MB_INTER) {
    for (int l = 0; l < nl; ++l) {
      const int16_t *mvo = pic_mv(*s.c, s.pd->frame, l) + (uint64_t)s.mbaddr * 32;
      const int8_t *ro = pic_refidx(*s.c, s.pd->frame, l) + (uint64_t)s.mbaddr * 4;
      for (int i = 0; i < 16; ++i) {
        int ci = HWB_CI(i & 3, i >> 2), q = ((i >> 2) >> 1) * 2 + ((i & 3) >> 1);
        if (ro[q] != s.ref_cache[l][ci] || (ro[q] >= 0 && (mvo[2 * i] != s.mv_cache[l][ci][0] || mvo[2 * i + 1] != s.mv_cache[l][ci][1]))) {
          fprintf(stderr, "h264gen: motion mismatch mb %d list %d blk %d: ir ref %d mv %d,%d  cache ref %d mv %d,%d\n", s.mbaddr, l, i,
                  ro[q], mvo[2 * i], mvo[2 * i + 1], s.ref_cache[l][ci], s.mv_cache[l][ci][0], s.mv_cache[l][ci][1]);
          abort();
        }
      }
    }
  }
  finish_mb(s, m.skipped, direct16, is_pcm);
}

}  // namespace gen
