// Intermediate representation shared by the host scheduler, the CUDA kernels and the stream
// generator: what the host hands to the GPU per picture/slice, and what the entropy stage
// hands to reconstruction / deblocking.  Plain-old-data only.
//
// Replaces (on the GPU) the state libavcodec keeps inside avcodec_send_packet for the
// reference's SoftwareVideoDecoder::feed_packet (hwang/impls/software/software_video_decoder.cpp:349-402).
#pragma once
#include "hd.h"

namespace hwb {

enum { MB_I4x4 = 0, MB_I8x8 = 1, MB_I16x16 = 2, MB_IPCM = 3, MB_INTER = 4 };
enum { MBF_T8x8 = 1, MBF_SKIP = 2 };
enum { SLICE_P = 0, SLICE_B = 1, SLICE_I = 2 };

// Coefficient slots: every coded residual block of a macroblock occupies one 16-coefficient
// slot (32 bytes, raster 4x4 order, raw levels before dequantisation) in the picture's
// coefficient arena; an 8x8-transform block occupies 4 consecutive slots (64 coefficients,
// raster 8x8 order).  nzmask bit i set <=> slot item i is present; the item's slot index is
// coef_off + popcount(nzmask & ((1<<i)-1)).  Item order:
//   bit 0       luma DC of an Intra16x16 macroblock
//   bits 1..16  luma 4x4 blocks in luma4x4BlkIdx (z) order
//   bit 17, 18  chroma DC Cb, Cr (4 coefficients used)
//   bits 19..22 Cb AC blocks, bits 23..26 Cr AC blocks (raster 2x2 order)
// An I_PCM macroblock stores its 384 raw samples in 12 slots (nzmask = 0xFFF).
enum { NZ_LUMA_DC = 0, NZ_LUMA0 = 1, NZ_CB_DC = 17, NZ_CR_DC = 18, NZ_CB0 = 19, NZ_CR0 = 23 };
enum { SLOTS_PER_MB = 27, COEFS_PER_SLOT = 16 };

struct alignas(16) MbInfo {
  uint8_t mbtype;   // MB_*
  uint8_t qp;       // QP_Y of this macroblock
  uint8_t cbp;      // coded_block_pattern (bits 0-3 luma 8x8, bits 4-5 chroma)
  uint8_t flags;    // MBF_*
  uint8_t imode;    // Intra16x16 prediction mode
  uint8_t cmode;    // intra chroma prediction mode
  uint16_t slice;   // slice number inside the picture
  uint32_t nzmask;
  uint32_t coef_off;    // first slot, relative to the picture's arena
  uint8_t i4modes[16];  // Intra4x4 (z order, 16 used) / Intra8x8 (4 used) prediction modes
};
static_assert(sizeof(MbInfo) == 32, "MbInfo layout");

struct PicDesc {
  int32_t frame;        // frame-buffer index the picture is reconstructed into
  int32_t first_slice;  // index into the chunk's SliceDesc array
  int32_t num_slices;
  int32_t poc;
  int32_t level;        // dependency level inside the chunk (host scheduling only)
  uint8_t cabac;
  uint8_t transform8x8_mode;
  uint8_t constrained_intra_pred;
  uint8_t direct_8x8_inference;
  uint8_t weighted_pred;
  uint8_t weighted_bipred_idc;
  uint8_t is_ref;
  uint8_t has_inter;
  int8_t chroma_qp_offset[2];
  uint8_t num_dep;      // distinct reference frames of all slices of the picture (dep[]), for the picture kernel's waits
  uint8_t pad[1];
  int32_t rgb_slot;     // slot of the chunk's RGB24 arena this picture is converted into by the deblocking pass, -1 = not returned
  int16_t dep[32];      // frame indices this picture predicts from
  uint8_t scaling4[6][16];  // raster order: Y-intra, Cb-intra, Cr-intra, Y-inter, Cb-inter, Cr-inter
  uint8_t scaling8[2][64];  // raster order: intra, inter
};

struct SliceDesc {
  int32_t pic;         // picture index inside the chunk
  int32_t first_mb;
  int32_t end_mb;      // first macroblock of the next slice of the picture (number of macroblocks for the last slice): the slice must end exactly here
  uint32_t data_off;   // byte offset of the slice RBSP (NAL header stripped, emulation bytes removed)
  uint32_t data_size;  // RBSP bytes
  uint32_t bit_off;    // bit offset of slice_data() inside the RBSP
  uint8_t slice_type;  // SLICE_*
  uint8_t qp;          // SliceQP_Y
  uint8_t cabac_init_idc;
  uint8_t disable_deblock;  // disable_deblocking_filter_idc
  int8_t alpha_off;         // FilterOffsetA
  int8_t beta_off;          // FilterOffsetB
  uint8_t num_ref[2];
  uint8_t direct_spatial;
  uint8_t luma_log2_denom;
  uint8_t chroma_log2_denom;
  uint8_t use_weights;      // 0 default, 1 explicit, 2 implicit
  int16_t ref_frame[2][32];  // frame-buffer index of each reference list entry
  int32_t ref_poc[2][32];
  uint32_t ref_long[2];      // bit i: list entry i is a long-term reference
  int16_t luma_w[2][32], luma_o[2][32];
  int16_t chroma_w[2][32][2], chroma_o[2][32][2];
};

// Everything a kernel needs about the chunk of pictures being decoded.  One chunk = a run of
// whole closed GOPs of one stream (constant coded size); frame buffers are never reused inside
// a chunk, so frame index == picture index and there are no write-after-read hazards.
struct ChunkCtx {
  int32_t mb_w, mb_h, nmb;
  int32_t nmb_stride;  // macroblocks per picture in the per-picture arrays below (>= nmb; a multiple of 32 on the GPU: pictures never share a 128-byte line)
  int32_t wc, hc;  // coded luma size (multiples of 16)
  int32_t num_pics, num_slices;
  uint8_t *frames;         // [frame] planar Y (wc*hc), Cb, Cr (wc/2*hc/2)
  uint64_t frame_stride;   // bytes per frame buffer
  MbInfo *mbinfo;          // [frame][nmb]
  int16_t *mv;             // [frame][list][nmb][16 (raster 4x4)][2]
  int8_t *refidx;          // [frame][list][nmb][4]
  int16_t *refpic;         // [frame][list][nmb][4]  frame index referenced, -1 = none
  int16_t *coefs;          // [frame][nmb*SLOTS_PER_MB][16]
  uint8_t *ectx;           // [slice] entropy neighbour-context scratch (per-slice line buffers)
  uint64_t ectx_stride;
  const uint8_t *bitstream;
  const PicDesc *pics;
  const SliceDesc *slices;
  const int32_t *entropy_order;  // [num_tickets] ticket -> slice index: I slices first (longest, no dependencies), then decode order
  int32_t num_tickets;           // slices to entropy-decode (slices of skipped pictures hold no ticket)
  int32_t num_intra_tickets;     // the first num_intra_tickets of them are intra slices
  int32_t intra_sms;             // SMs [intra_sm_base, intra_sm_base + intra_sms) decode the intra slices (see kernels.cu); 0 = no reservation
  int32_t intra_sm_base;         // first reserved SM (0: every launch reserves SMs from the same end of the device, so that the intra slices of batches in flight together share SMs with each other rather than with inter slices)
  int32_t *entropy_prog;   // [slice] first macroblock address not yet entropy-decoded (B direct col dependency)
  int32_t *recon_prog;     // [pic][mb_h] macroblocks reconstructed per row
  int32_t *dbl_prog;       // [pic][mb_h] macroblocks deblocked per row
  int32_t *mv_reach;       // [pic][mb_h] 1 + lowest macroblock row of a reference picture that inter prediction of this row reads (0 = none); written by the entropy stage
  int32_t *mv_reach_x;     // [pic][mb_h] how many macroblocks of those reference rows, counted from a macroblock's own column, must be deblocked before the macroblock can be predicted (>= 1)
  int32_t *error_flag;     // set non-zero by any kernel that meets an unsupported/corrupt stream
  // ---- picture kernel (reconstruction + deblocking + RGB24 writeback of a whole chunk in one launch)
  // Work items: pic << 12 | macroblock row << 1 | kind (0 reconstruct the row, 1 deblock it and emit RGB24).
  // Each list is ordered so that every item only waits on items that come earlier (in its own or the other list).
  const uint32_t *recon_items, *deblock_items;
  int32_t num_recon_items, num_deblock_items;
  uint8_t *rgb;            // [slot] packed RGB24, out_w * out_h * 3 bytes each (tight)
  uint64_t rgb_stride;
  int32_t crop_x, crop_y, out_w, out_h;  // cropping rectangle of the output frames inside the coded picture
  int32_t deblock_band;    // rows per deblocking work item (1..DEBLOCK_BAND, picture.h)
  // Completion of single pictures, so that frames travel to the host while the picture kernel is still working on the
  // rest of the batch: rows_done[pic] counts deblocked rows (device memory); the warp that completes the last row of a
  // picture sets pic_done[pic] = 1 in page-locked HOST memory (mapped into the device address space), which the host
  // polls before it copies the frame out.
  int32_t *rows_done;
  int32_t *pic_done;
  // optional (HWB_PICTURE_PROFILE=1): warp cycles of the picture kernel by activity, see picture.h PROF_*; nullptr = off
  unsigned long long *prof;
};

HWB_HD uint8_t *frame_y(const ChunkCtx &c, int f) { return c.frames + (uint64_t)f * c.frame_stride; }
HWB_HD uint8_t *frame_cb(const ChunkCtx &c, int f) { return frame_y(c, f) + (uint64_t)c.wc * c.hc; }
HWB_HD uint8_t *frame_cr(const ChunkCtx &c, int f) { return frame_cb(c, f) + (uint64_t)(c.wc >> 1) * (c.hc >> 1); }
HWB_HD MbInfo *pic_mbinfo(const ChunkCtx &c, int f) { return c.mbinfo + (uint64_t)f * c.nmb_stride; }
HWB_HD int16_t *pic_mv(const ChunkCtx &c, int f, int list) {
  return c.mv + ((uint64_t)f * 2 + list) * c.nmb_stride * 32;
}
HWB_HD int8_t *pic_refidx(const ChunkCtx &c, int f, int list) {
  return c.refidx + ((uint64_t)f * 2 + list) * c.nmb_stride * 4;
}
HWB_HD int16_t *pic_refpic(const ChunkCtx &c, int f, int list) {
  return c.refpic + ((uint64_t)f * 2 + list) * c.nmb_stride * 4;
}
HWB_HD int16_t *pic_coefs(const ChunkCtx &c, int f) {
  return c.coefs + (uint64_t)f * c.nmb_stride * SLOTS_PER_MB * COEFS_PER_SLOT;
}

}  // namespace hwb
