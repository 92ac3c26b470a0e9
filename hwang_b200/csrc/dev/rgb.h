// Planar 4:2:0 -> packed RGB24 with the exact fixed-point arithmetic of swscale's unscaled
// yuv420p->rgb24 path, i.e. what the reference's get_frame produces
// (hwang/impls/software/software_video_decoder.cpp:292-325):
//   y = ((Y*8 - 128) * 9539) >> 16,  u = (U-128)*8,  v = (V-128)*8
//   R = clip8(y + ((v*13075) >> 16)); G = clip8(y + ((u*-3209) >> 16) + ((v*-6660) >> 16)); B = clip8(y + ((u*16525) >> 16))
// chroma is nearest-neighbour (one Cb/Cr sample per 2x2 luma block).
#pragma once
#include "ir.h"

namespace hwb {

HWB_HD void yuv_to_rgb(int Y, int U, int V, uint8_t *rgb) {
  int y = ((Y * 8 - 128) * 9539) >> 16, u = (U - 128) * 8, v = (V - 128) * 8;
  rgb[0] = (uint8_t)clip8(y + ((v * 13075) >> 16));
  rgb[1] = (uint8_t)clip8(y + ((u * -3209) >> 16) + ((v * -6660) >> 16));
  rgb[2] = (uint8_t)clip8(y + ((u * 16525) >> 16));
}

// 16 horizontally adjacent pixels (n < 16 at the right edge of a picture whose width is not a multiple of 16):
// Yp / Up / Vp point at the first sample, o at the first output byte.  The vector path (one 16-byte and two 8-byte
// loads, three 16-byte streaming stores) needs n == 16 and aligned pointers.  Loads go to L2 (ld.global.cg): inside the
// picture kernel the samples were written moments ago by a warp of the same launch, possibly on another SM.
// One pixel as R | G << 8 | B << 16; the chroma terms of a pixel pair are computed once (chroma is shared by two pixels).
HWB_HD uint32_t rgb_pack(int Y, int rv, int gv, int bv) {
  const int y = ((Y * 8 - 128) * 9539) >> 16;
  return (uint32_t)clip8(y + rv) | ((uint32_t)clip8(y + gv) << 8) | ((uint32_t)clip8(y + bv) << 16);
}
HWB_HD void rgb24_segment(const uint8_t *Yp, const uint8_t *Up, const uint8_t *Vp, uint8_t *o, int n) {
  const bool vec = n == 16 && ((((uintptr_t)Yp) & 15) == 0) && ((((uintptr_t)Up) & 7) == 0) && ((((uintptr_t)Vp) & 7) == 0) && ((((uintptr_t)o) & 15) == 0);
  if (vec) {
    // registers only: 16 + 8 + 8 sample bytes in 8 words, 48 output bytes in 12 (byte arrays indexed at run time by the
    // edge path used to put both paths' samples and output into thread-local memory: 80 bytes through L1 per segment)
    uint32_t yw[4], uw[2], vw[2], ow[12];
#if HWB_DEVICE_BUILD
    { const uint4 t = __ldcg((const uint4 *)Yp); yw[0] = t.x; yw[1] = t.y; yw[2] = t.z; yw[3] = t.w; }
    { const uint2 t = __ldcg((const uint2 *)Up); uw[0] = t.x; uw[1] = t.y; }
    { const uint2 t = __ldcg((const uint2 *)Vp); vw[0] = t.x; vw[1] = t.y; }
#else
    memcpy(yw, Yp, 16); memcpy(uw, Up, 8); memcpy(vw, Vp, 8);
#endif
#pragma unroll
    for (int g = 0; g < 4; ++g) {  // four pixels = two chroma samples -> three output words
      uint32_t px[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ci = 2 * g + h;  // chroma sample index 0..7
        const int u = ((int)((uw[ci >> 2] >> (8 * (ci & 3))) & 0xff) - 128) * 8, v = ((int)((vw[ci >> 2] >> (8 * (ci & 3))) & 0xff) - 128) * 8;
        const int rv = (v * 13075) >> 16, gv = ((u * -3209) >> 16) + ((v * -6660) >> 16), bv = (u * 16525) >> 16;
        px[2 * h] = rgb_pack((int)((yw[g] >> (16 * h)) & 0xff), rv, gv, bv);
        px[2 * h + 1] = rgb_pack((int)((yw[g] >> (16 * h + 8)) & 0xff), rv, gv, bv);
      }
      ow[3 * g] = px[0] | (px[1] << 24);
      ow[3 * g + 1] = (px[1] >> 8) | (px[2] << 16);
      ow[3 * g + 2] = (px[2] >> 16) | (px[3] << 8);
    }
#if HWB_DEVICE_BUILD
    uint4 *o4 = (uint4 *)o;
    __stcs(o4, make_uint4(ow[0], ow[1], ow[2], ow[3])); __stcs(o4 + 1, make_uint4(ow[4], ow[5], ow[6], ow[7])); __stcs(o4 + 2, make_uint4(ow[8], ow[9], ow[10], ow[11]));
#else
    memcpy(o, ow, 48);
#endif
    return;
  }
  // edge path (right edge of a picture whose width is not a multiple of 16, unaligned crop): byte by byte
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    uint8_t px[3];
    yuv_to_rgb(ld_u8_cg(Yp + i), ld_u8_cg(Up + (i >> 1)), ld_u8_cg(Vp + (i >> 1)), px);
    o[3 * i] = px[0]; o[3 * i + 1] = px[1]; o[3 * i + 2] = px[2];
  }
}

// One work item converts 16 horizontally adjacent pixels of the cropped frame: x16 in [0, ceil(w/16)), y in [0, h).
HWB_HD void rgb24_item(const ChunkCtx &c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst, int x16, int y) {
  (void)h;
  const uint8_t *Yp = frame_y(c, frame) + (uint64_t)(crop_y + y) * c.wc + crop_x + x16 * 16;
  const uint8_t *Up = frame_cb(c, frame) + (uint64_t)((crop_y + y) >> 1) * (c.wc >> 1) + ((crop_x + x16 * 16) >> 1);
  const uint8_t *Vp = frame_cr(c, frame) + (uint64_t)((crop_y + y) >> 1) * (c.wc >> 1) + ((crop_x + x16 * 16) >> 1);
  rgb24_segment(Yp, Up, Vp, dst + ((uint64_t)y * w + x16 * 16) * 3, w - x16 * 16 < 16 ? w - x16 * 16 : 16);
}

// RGB24 writeback fused into the deblocking pass (BASELINE.json north_star): the warp that has just deblocked a
// macroblock converts the macroblocks that became final with it (see picture_kernel) straight from L2 into the
// chunk's RGB arena.  Converts `nmb` (1 or 2) macroblocks starting at (mbx0, mby): lane = (sample row, macroblock).
// Crop offsets must be even (they are: 4:2:0 cropping units), the part outside the cropping rectangle is skipped.
HWB_FN void rgb24_macroblocks(const ChunkCtx &c, int frame, uint8_t *dst, int mbx0, int nmb, int mby) {
  HWB_LANES(l)
  for (int it = l; it < 16 * nmb; it += 32) {
    const int row = nmb == 2 ? it >> 1 : it, seg = nmb == 2 ? (it & 1) : 0;
    const int y = mby * 16 + row - c.crop_y, x0 = (mbx0 + seg) * 16 - c.crop_x;
    if (y < 0 || y >= c.out_h) continue;
    int xa = x0 < 0 ? 0 : x0, xb = x0 + 16 > c.out_w ? c.out_w : x0 + 16;
    if (xa >= xb) continue;
    const int cy = y + c.crop_y, cx = xa + c.crop_x;
    const uint8_t *Yp = frame_y(c, frame) + (uint64_t)cy * c.wc + cx;
    const uint8_t *Up = frame_cb(c, frame) + (uint64_t)(cy >> 1) * (c.wc >> 1) + (cx >> 1);
    const uint8_t *Vp = frame_cr(c, frame) + (uint64_t)(cy >> 1) * (c.wc >> 1) + (cx >> 1);
    rgb24_segment(Yp, Up, Vp, dst + ((uint64_t)y * c.out_w + xa) * 3, xb - xa);
  }
  HWB_LANES_END
}

}  // namespace hwb
