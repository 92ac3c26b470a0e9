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

// One work item converts 16 horizontally adjacent pixels: x16 in [0, ceil(w/16)), y in [0, h).
HWB_HD void rgb24_item(const ChunkCtx &c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst, int x16, int y) {
  const uint8_t *Yp = frame_y(c, frame) + (uint64_t)(crop_y + y) * c.wc + crop_x + x16 * 16;
  const uint8_t *Up = frame_cb(c, frame) + (uint64_t)((crop_y + y) >> 1) * (c.wc >> 1) + ((crop_x + x16 * 16) >> 1);
  const uint8_t *Vp = frame_cr(c, frame) + (uint64_t)((crop_y + y) >> 1) * (c.wc >> 1) + ((crop_x + x16 * 16) >> 1);
  uint8_t *o = dst + ((uint64_t)y * w + x16 * 16) * 3;
  const int n = w - x16 * 16 < 16 ? w - x16 * 16 : 16;
  alignas(16) uint8_t yy[16], uu[8], vv[8], out[48];
#if HWB_DEVICE_BUILD
  const bool vec = n == 16 && ((((uintptr_t)Yp) & 15) == 0) && ((((uintptr_t)Up) & 7) == 0) && ((((uintptr_t)Vp) & 7) == 0) && ((((uintptr_t)o) & 15) == 0);
  if (vec) {
    *(uint4 *)yy = __ldg((const uint4 *)Yp);
    *(uint2 *)uu = __ldg((const uint2 *)Up);
    *(uint2 *)vv = __ldg((const uint2 *)Vp);
  } else
#endif
  {
    for (int i = 0; i < n; ++i) { yy[i] = Yp[i]; uu[i >> 1] = Up[i >> 1]; vv[i >> 1] = Vp[i >> 1]; }
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

}  // namespace hwb
