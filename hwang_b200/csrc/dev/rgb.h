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
HWB_HD void rgb24_segment(const uint8_t *Yp, const uint8_t *Up, const uint8_t *Vp, uint8_t *o, int n) {
  alignas(16) uint8_t yy[16], uu[8], vv[8], out[48];
#if HWB_DEVICE_BUILD
  const bool vec = n == 16 && ((((uintptr_t)Yp) & 15) == 0) && ((((uintptr_t)Up) & 7) == 0) && ((((uintptr_t)Vp) & 7) == 0) && ((((uintptr_t)o) & 15) == 0);
  if (vec) {
    *(uint4 *)yy = __ldcg((const uint4 *)Yp);
    *(uint2 *)uu = __ldcg((const uint2 *)Up);
    *(uint2 *)vv = __ldcg((const uint2 *)Vp);
  } else
#endif
  {
    for (int i = 0; i < n; ++i) { yy[i] = ld_u8_cg(Yp + i); uu[i >> 1] = ld_u8_cg(Up + (i >> 1)); vv[i >> 1] = ld_u8_cg(Vp + (i >> 1)); }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) yuv_to_rgb(yy[i], uu[i >> 1], vv[i >> 1], out + 3 * i);
#if HWB_DEVICE_BUILD
  if (vec) {
    uint4 *o4 = (uint4 *)o;
    const uint4 *s4 = (const uint4 *)out;
    __stcs(o4, s4[0]); __stcs(o4 + 1, s4[1]); __stcs(o4 + 2, s4[2]);
  } else
#endif
  {
    for (int i = 0; i < n * 3; ++i) o[i] = out[i];
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
