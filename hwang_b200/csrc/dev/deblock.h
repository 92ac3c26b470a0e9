// In-loop deblocking filter (H.264 clause 8.7), one warp per macroblock, in place on the frame
// buffer.  Macroblock (x,y) may run once (x-1,y), (x,y-1) and (x+1,y-1) are finished: rows of a
// picture form a wavefront with a two-macroblock lag (see kernels.cu).
#pragma once
#include "ir.h"
#include "tables_gen.h"
#include "recon.h"

namespace hwb {

enum { DL_STRIDE = 32, DL_OFF = 16, DC_STRIDE = 16, DC_OFF = 8 };  // column 0 of every tile row is 16- (luma) / 8-byte (chroma) aligned

// 16- and 8-byte moves: the macroblock's own 16 luma / 8 chroma columns of a row travel as one 128-bit / 64-bit access
// (ld.global.cg: the samples were written by other warps of the same launch), the 4 columns left of it as one word.
struct alignas(16) Vec16 { uint32_t w[4]; };
struct alignas(8) Vec8 { uint32_t w[2]; };
HWB_HD Vec16 ld_v16_cg(const uint8_t *p) {
  Vec16 v;
#if HWB_DEVICE_BUILD
  const uint4 t = __ldcg((const uint4 *)p);
  v.w[0] = t.x; v.w[1] = t.y; v.w[2] = t.z; v.w[3] = t.w;
#else
  memcpy(&v, p, 16);
#endif
  return v;
}
HWB_HD Vec8 ld_v8_cg(const uint8_t *p) {
  Vec8 v;
#if HWB_DEVICE_BUILD
  const uint2 t = __ldcg((const uint2 *)p);
  v.w[0] = t.x; v.w[1] = t.y;
#else
  memcpy(&v, p, 8);
#endif
  return v;
}
HWB_HD void st_v16(uint8_t *p, const Vec16 &v) {
#if HWB_DEVICE_BUILD
  *(uint4 *)p = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
#else
  memcpy(p, &v, 16);
#endif
}
HWB_HD void st_v8(uint8_t *p, const Vec8 &v) {
#if HWB_DEVICE_BUILD
  *(uint2 *)p = make_uint2(v.w[0], v.w[1]);
#else
  memcpy(p, &v, 8);
#endif
}

struct alignas(16) DeblockScratch {
  uint8_t luma[20 * DL_STRIDE];       // rows -4..15 (row r at (r+4)), col c at DL_OFF + c (c = -4..15)
  uint8_t chroma[2][12 * DC_STRIDE];  // rows -4..7, col c at DC_OFF + c (c = -4..7)
  uint8_t bs[32];                     // [dir][edge][segment]
#if !HWB_DEVICE_BUILD
  Vec16 pre_l[32][2];                 // host emulation only: per-lane registers that live across lane blocks
  Vec8 pre_c[32][2];
#endif
};

// Motion vectors travel as one word (x | y << 16), as they are stored.
HWB_HD bool mv_far(uint32_t a, uint32_t b) {
  return iabs((int)(int16_t)(a & 0xffff) - (int)(int16_t)(b & 0xffff)) >= 4 || iabs((int)(int16_t)(a >> 16) - (int)(int16_t)(b >> 16)) >= 4;
}

struct BlkMotion {  // by value: as pointers into the motion arrays the structure ended up in thread-local memory
  int r0, r1;       // referenced frame per list, -1 = unused
  uint32_t m0, m1;
};

HWB_HD BlkMotion blk_motion(const ChunkCtx &c, int frame, bool two_lists, int mbaddr, int bx, int by) {
  BlkMotion m;
  int q = (by >> 1) * 2 + (bx >> 1), br = by * 4 + bx;
  m.r0 = pic_refpic(c, frame, 0)[(uint64_t)mbaddr * 4 + q];
  m.m0 = ((const uint32_t *)(pic_mv(c, frame, 0) + (uint64_t)mbaddr * 32))[br];
  if (two_lists) {
    m.r1 = pic_refpic(c, frame, 1)[(uint64_t)mbaddr * 4 + q];
    m.m1 = ((const uint32_t *)(pic_mv(c, frame, 1) + (uint64_t)mbaddr * 32))[br];
  } else {
    m.r1 = -1; m.m1 = m.m0;
  }
  return m;
}

HWB_HD int motion_bs(const BlkMotion &p, const BlkMotion &q) {
  if (!((p.r0 == q.r0 && p.r1 == q.r1) || (p.r0 == q.r1 && p.r1 == q.r0))) return 1;
  if (p.r0 != p.r1) {
    if (p.r0 == q.r0) {
      if (p.r0 >= 0 && mv_far(p.m0, q.m0)) return 1;
      if (p.r1 >= 0 && mv_far(p.m1, q.m1)) return 1;
    } else {
      if (p.r0 >= 0 && mv_far(p.m0, q.m1)) return 1;
      if (p.r1 >= 0 && mv_far(p.m1, q.m0)) return 1;
    }
    return 0;
  }
  if (p.r0 < 0) return 0;
  bool direct = mv_far(p.m0, q.m0) || mv_far(p.m1, q.m1);
  bool cross = mv_far(p.m0, q.m1) || mv_far(p.m1, q.m0);
  return (direct && cross) ? 1 : 0;
}

// Filter one line across an edge.  pix points at q0, `st` is the step across the edge.
HWB_HD void filter_luma_line(uint8_t *pix, int st, int bS, int alpha, int beta, int tc0) {
  int p0 = pix[-st], p1 = pix[-2 * st], p2 = pix[-3 * st];
  int q0 = pix[0], q1 = pix[st], q2 = pix[2 * st];
  if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
  int ap = iabs(p2 - p0), aq = iabs(q2 - q0);
  if (bS < 4) {
    int tc = tc0 + (ap < beta) + (aq < beta);
    int delta = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
    pix[-st] = (uint8_t)clip8(p0 + delta);
    pix[0] = (uint8_t)clip8(q0 - delta);
    if (ap < beta) pix[-2 * st] = (uint8_t)(p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1));
    if (aq < beta) pix[st] = (uint8_t)(q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1));
  } else {
    bool strong = iabs(p0 - q0) < ((alpha >> 2) + 2);
    if (ap < beta && strong) {
      int p3 = pix[-4 * st];
      pix[-st] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      pix[-2 * st] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
      pix[-3 * st] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
    } else {
      pix[-st] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
    }
    if (aq < beta && strong) {
      int q3 = pix[3 * st];
      pix[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      pix[st] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
      pix[2 * st] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
    } else {
      pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
  }
}

HWB_HD void filter_chroma_line(uint8_t *pix, int st, int bS, int alpha, int beta, int tc0) {
  int p0 = pix[-st], p1 = pix[-2 * st], q0 = pix[0], q1 = pix[st];
  if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
  if (bS < 4) {
    int tc = tc0 + 1;
    int delta = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
    pix[-st] = (uint8_t)clip8(p0 + delta);
    pix[0] = (uint8_t)clip8(q0 - delta);
  } else {
    pix[-st] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
    pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
  }
}

HWB_FN void deblock_mb(const ChunkCtx &c, int pic, int mbx, int mby, DeblockScratch *sm) {
  const PicDesc &pd = c.pics[pic];
  const int mbaddr = mby * c.mb_w + mbx;
  const MbInfo *mbs = pic_mbinfo(c, pd.frame);
  const MbInfo mb = mbs[mbaddr];
  const SliceDesc &sd = c.slices[pd.first_slice + mb.slice];
  if (sd.disable_deblock == 1) return;
  const bool have_left = mbx > 0 && !(sd.disable_deblock == 2 && mbs[mbaddr - 1].slice != mb.slice);
  const bool have_top = mby > 0 && !(sd.disable_deblock == 2 && mbs[mbaddr - c.mb_w].slice != mb.slice);
  const int wc = c.wc, cw = c.wc >> 1;
  uint8_t *Y = frame_y(c, pd.frame), *Cb = frame_cb(c, pd.frame), *Cr = frame_cr(c, pd.frame);
  const bool two_lists = pd.has_inter == 2;  // picture contains B slices

  // ---- boundary strengths: one lane per (direction, edge, 4-sample segment)
  HWB_LANES(l)
    const int dir = l >> 4, e = (l >> 2) & 3, s = l & 3;
    int bS = 0;
    const bool mbedge = e == 0;
    const bool exists = mbedge ? (dir ? have_top : have_left) : !((mb.flags & MBF_T8x8) && (e & 1));
    if (exists) {
      const int qbx = dir ? s : e, qby = dir ? e : s;
      int paddr = mbaddr, pbx = dir ? s : e - 1, pby = dir ? e - 1 : s;
      if (mbedge) { paddr = dir ? mbaddr - c.mb_w : mbaddr - 1; if (dir) pby = 3; else pbx = 3; }
      const MbInfo &pm = mbs[paddr];
      if (mb.mbtype != MB_INTER || pm.mbtype != MB_INTER) bS = mbedge ? 4 : 3;
      else if (((mb.nzmask >> (NZ_LUMA0 + xy2z(qbx, qby))) & 1) || ((pm.nzmask >> (NZ_LUMA0 + xy2z(pbx, pby))) & 1)) bS = 2;
      else bS = motion_bs(blk_motion(c, pd.frame, two_lists, paddr, pbx, pby), blk_motion(c, pd.frame, two_lists, mbaddr, qbx, qby));
    }
    sm->bs[l] = (uint8_t)bS;
  HWB_LANES_END
  uint32_t any = 0;
#if HWB_DEVICE_BUILD
  any = __ballot_sync(0xffffffffu, sm->bs[threadIdx.x & 31] != 0);
#else
#pragma unroll 1
  for (int i = 0; i < 32; ++i) any |= sm->bs[i] ? (1u << i) : 0;
#endif
  if (!any) return;

  // ---- tile loads (coherent loads: neighbours were written by other warps of this launch), all issued before anything
  // waits on them so that their L2 round trips overlap each other.  They come AFTER the boundary strengths: most
  // macroblocks of an inter picture have no edge to filter and leave above without touching a sample (issuing the loads
  // first, to hide their latency behind the strength computation, cost 3 % of the picture kernel's time and bound the
  // share of SMs the deblocking role needs: 40 % -> 35 %, together 600 -> 550 ms per 9000 pictures, profiles/r2_runs/r2as_ab.txt).
  // Item i of a plane: row i >> 1, part i & 1: part 0 = the macroblock's own columns (one 16-byte / 8-byte load),
  // part 1 = the 4 columns to the left (one word).
  Vec16 tl[2];
  Vec8 tc[2];
  HWB_LANES(l)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = l + 32 * k, r = (i >> 1) - 4, left = i & 1;  // rows -4..15
      const bool ok = i < 40 && !(r < 0 && mby == 0) && !(left && mbx == 0);
      const uint8_t *p = Y + (int64_t)(mby * 16 + r) * wc + mbx * 16;
      tl[k].w[0] = tl[k].w[1] = tl[k].w[2] = tl[k].w[3] = 0;
      if (ok) { if (left) tl[k].w[0] = ld_u32_cg((const uint32_t *)(p - 4)); else tl[k] = ld_v16_cg(p); }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = l + 32 * k, pl = i >= 24, j = i - 24 * pl, r = (j >> 1) - 4, left = j & 1;  // rows -4..7 of Cb, then Cr
      const bool ok = i < 48 && !(r < 0 && mby == 0) && !(left && mbx == 0);
      const uint8_t *p = (pl ? Cr : Cb) + (int64_t)(mby * 8 + r) * cw + mbx * 8;
      tc[k].w[0] = tc[k].w[1] = 0;
      if (ok) { if (left) tc[k].w[0] = ld_u32_cg((const uint32_t *)(p - 4)); else tc[k] = ld_v8_cg(p); }
    }
#if !HWB_DEVICE_BUILD
    for (int k = 0; k < 2; ++k) { sm->pre_l[l][k] = tl[k]; sm->pre_c[l][k] = tc[k]; }
#endif
  HWB_LANES_END

  // ---- tile to shared memory
  HWB_LANES(l)
#if !HWB_DEVICE_BUILD
    for (int k = 0; k < 2; ++k) { tl[k] = sm->pre_l[l][k]; tc[k] = sm->pre_c[l][k]; }
#endif
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = l + 32 * k, r = (i >> 1) - 4, left = i & 1;
      if (i < 40 && !(r < 0 && mby == 0) && !(left && mbx == 0)) {
        uint8_t *t = sm->luma + (r + 4) * DL_STRIDE + DL_OFF;
        if (left) *(uint32_t *)(t - 4) = tl[k].w[0]; else st_v16(t, tl[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = l + 32 * k, pl = i >= 24, j = i - 24 * pl, r = (j >> 1) - 4, left = j & 1;
      if (i < 48 && !(r < 0 && mby == 0) && !(left && mbx == 0)) {
        uint8_t *t = sm->chroma[pl] + (r + 4) * DC_STRIDE + DC_OFF;
        if (left) *(uint32_t *)(t - 4) = tc[k].w[0]; else st_v8(t, tc[k]);
      }
    }
  HWB_LANES_END

  const int qp_q = mb.qp;
  const int qp_l = have_left ? mbs[mbaddr - 1].qp : 0, qp_t = have_top ? mbs[mbaddr - c.mb_w].qp : 0;
  // ---- vertical edges then horizontal edges
#pragma unroll 1
  for (int dir = 0; dir < 2; ++dir) {
    const int qp_p_edge = dir ? qp_t : qp_l;
    HWB_LANES(l)
      if (l < 16) {
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
          int bS = sm->bs[dir * 16 + e * 4 + (l >> 2)];
          if (!bS) continue;
          int qpav = e == 0 ? (qp_p_edge + qp_q + 1) >> 1 : qp_q;
          int ia = clip3(0, 51, qpav + sd.alpha_off), ib = clip3(0, 51, qpav + sd.beta_off);
          int alpha = deblock_alpha[ia], beta = deblock_beta[ib];
          if (!alpha || !beta) continue;
          int tc0 = bS < 4 ? deblock_tc0[ia * 3 + bS - 1] : 0;
          uint8_t *pix = dir ? sm->luma + (e * 4 + 4) * DL_STRIDE + DL_OFF + l : sm->luma + (l + 4) * DL_STRIDE + DL_OFF + e * 4;
          filter_luma_line(pix, dir ? DL_STRIDE : 1, bS, alpha, beta, tc0);
        }
      } else {
        const int pl = (l - 16) >> 3, k = l & 7;
        const int off = pd.chroma_qp_offset[pl];
#pragma unroll 1
        for (int e = 0; e < 4; e += 2) {
          int bS = sm->bs[dir * 16 + e * 4 + (k >> 1)];
          if (!bS) continue;
          int qc = chroma_qp(qp_q, off);
          int qpav = e == 0 ? (chroma_qp(qp_p_edge, off) + qc + 1) >> 1 : qc;
          int ia = clip3(0, 51, qpav + sd.alpha_off), ib = clip3(0, 51, qpav + sd.beta_off);
          int alpha = deblock_alpha[ia], beta = deblock_beta[ib];
          if (!alpha || !beta) continue;
          int tc0 = bS < 4 ? deblock_tc0[ia * 3 + bS - 1] : 0;
          uint8_t *pix = dir ? sm->chroma[pl] + (e * 2 + 4) * DC_STRIDE + DC_OFF + k : sm->chroma[pl] + (k + 4) * DC_STRIDE + DC_OFF + e * 2;
          filter_chroma_line(pix, dir ? DC_STRIDE : 1, bS, alpha, beta, tc0);
        }
      }
    HWB_LANES_END
  }

  // ---- store rows -3..15 x columns -4..15 (luma) / rows -2..7 x columns -4..7 (chroma).  The corner above-left is not
  //      touched by this macroblock's filters and is left alone; rows above need a macroblock above, columns to the
  //      left one to the left.  Own columns: one 16-byte / 8-byte store per row.
  HWB_LANES(l)
#pragma unroll 1
    for (int i = l; i < 38; i += 32) {
      const int r = (i >> 1) - 3, left = i & 1;
      if ((r < 0 && (mby == 0 || left)) || (left && mbx == 0)) continue;
      const uint8_t *t = sm->luma + (r + 4) * DL_STRIDE + DL_OFF;
      uint8_t *p = Y + (int64_t)(mby * 16 + r) * wc + mbx * 16;
      if (left) *(uint32_t *)(p - 4) = *(const uint32_t *)(t - 4); else st_v16(p, *(const Vec16 *)t);
    }
#pragma unroll 1
    for (int i = l; i < 40; i += 32) {
      const int pl = i >= 20, j = i - 20 * pl, r = (j >> 1) - 2, left = j & 1;
      if ((r < 0 && (mby == 0 || left)) || (left && mbx == 0)) continue;
      const uint8_t *t = sm->chroma[pl] + (r + 4) * DC_STRIDE + DC_OFF;
      uint8_t *p = (pl ? Cr : Cb) + (int64_t)(mby * 8 + r) * cw + mbx * 8;
      if (left) *(uint32_t *)(p - 4) = *(const uint32_t *)(t - 4); else st_v8(p, *(const Vec8 *)t);
    }
  HWB_LANES_END
}

}  // namespace hwb
