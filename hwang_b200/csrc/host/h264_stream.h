// Host-side H.264 syntax above the macroblock layer: NAL unit extraction from AVCC samples,
// emulation-prevention removal, SPS / PPS / slice-header parsing (7.3.2.1, 7.3.2.2, 7.3.3),
// picture order count (8.2.1), reference list construction + modification (8.2.4), reference
// marking (8.2.5).  Produces the PicDesc / SliceDesc records the GPU kernels consume.
//
// In the reference all of this happens inside libavcodec behind
// SoftwareVideoDecoder::feed (hwang/impls/software/software_video_decoder.cpp:167-248, 349-402);
// the reference's own partial helpers (hwang/util/h264.h:83-338) are unused there.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../dev/ir.h"

namespace hwb {

struct Sps {
  bool valid = false;
  int profile_idc = 0, level_idc = 0, constraint = 0;
  int chroma_format_idc = 1, bit_depth_luma = 8, bit_depth_chroma = 8;
  bool scaling_present = false;
  bool list_present[8] = {false};
  bool list_use_default[8] = {false};
  uint8_t scaling4[6][16], scaling8[2][64];  // raster
  int log2_max_frame_num = 4, poc_type = 0, log2_max_poc_lsb = 4;
  bool delta_pic_order_always_zero = false;
  int offset_for_non_ref_pic = 0, offset_for_top_to_bottom = 0;
  std::vector<int> offset_for_ref_frame;
  int max_num_ref_frames = 1;
  bool gaps_allowed = false;
  int mb_w = 0, mb_h = 0;
  bool frame_mbs_only = true, direct_8x8_inference = true;
  int crop_l = 0, crop_r = 0, crop_t = 0, crop_b = 0;
  int max_num_reorder_frames = -1, max_dec_frame_buffering = -1;
  int width() const { return mb_w * 16 - 2 * (crop_l + crop_r); }
  int height() const { return mb_h * 16 - 2 * (crop_t + crop_b); }
};

struct Pps {
  bool valid = false;
  int sps_id = 0;
  bool cabac = false, bottom_field_pic_order = false;
  int num_slice_groups = 1;
  int num_ref_default[2] = {1, 1};
  bool weighted_pred = false;
  int weighted_bipred_idc = 0;
  int init_qp = 26, chroma_qp_offset[2] = {0, 0};
  bool deblocking_control = false, constrained_intra = false, redundant_pic_cnt = false;
  bool transform8x8 = false;
  bool scaling_present = false;
  bool list_present[8] = {false};
  bool list_use_default[8] = {false};
  uint8_t scaling4[6][16], scaling8[2][64];
};

struct Mmco { int op, a, b; };

struct SliceHeader {
  int nal_type = 0, nal_ref_idc = 0;
  int first_mb = 0, slice_type = 0 /* SLICE_* */, pps_id = 0, frame_num = 0, idr_pic_id = 0;
  int poc_lsb = 0, delta_poc_bottom = 0, delta_poc[2] = {0, 0};
  bool direct_spatial = true;
  int num_ref[2] = {0, 0};
  struct Mod { int idc, val; };
  std::vector<Mod> mods[2];
  int luma_log2_denom = 0, chroma_log2_denom = 0;
  int16_t luma_w[2][32], luma_o[2][32], chroma_w[2][32][2], chroma_o[2][32][2];
  bool no_output_of_prior = false, long_term_ref = false, adaptive_marking = false;
  std::vector<Mmco> mmco;
  int cabac_init_idc = 0, qp = 26;
  int disable_deblock = 0, alpha_off = 0, beta_off = 0;
  uint32_t data_bit_off = 0;
};

// One picture planned for the GPU.
struct PlannedPic {
  PicDesc desc;
  std::vector<SliceDesc> slices;
  int64_t out_key;  // sort key for display order: (output period << 32) + POC biased
  bool idr;
};

class H264Stream {
 public:
  // avcC -> SPS/PPS tables, NAL length size.  Returns error text or "".
  std::string configure(const uint8_t *avcc, size_t n);
  // Parse one MP4 sample (one access unit).  Appends the slice RBSPs to `bitstream` and one PlannedPic to
  // `out` (frame index = `pic_index`, the picture's index inside the current chunk).  Returns error or "".
  std::string parse_sample(const uint8_t *data, size_t n, int pic_index, std::vector<uint8_t> &bitstream, PlannedPic &out);
  // Called when a chunk boundary is crossed: references into the previous chunk are forgotten
  // (chunks are cut at IDR pictures only, so nothing can refer across).
  void reset_dpb();
  bool next_is_idr(const uint8_t *data, size_t n) const;
  int mb_w() const { return mb_w_; }
  int mb_h() const { return mb_h_; }
  int width() const { return width_; }
  int height() const { return height_; }
  int crop_left() const { return crop_x_; }
  int crop_top() const { return crop_y_; }

 private:
  struct DpbEntry {
    int frame;      // chunk-local frame index
    int frame_num, frame_num_wrap, pic_num, poc;
    bool long_term;
    int long_term_idx;
  };
  std::string parse_sps(const std::vector<uint8_t> &rbsp);
  std::string parse_pps(const std::vector<uint8_t> &rbsp);
  std::string parse_slice_header(const uint8_t *rbsp, size_t rbsp_size, int nal_type, int nal_ref_idc, SliceHeader &sh);
  int compute_poc(const Sps &sps, const SliceHeader &sh);
  void build_ref_lists(const Sps &sps, const SliceHeader &sh, int cur_poc, std::vector<DpbEntry> lists[2]);
  void mark_references(const Sps &sps, const SliceHeader &sh, int cur_frame, int cur_poc);

  Sps sps_[32];
  Pps pps_[256];
  int nal_length_size_ = 4;
  int mb_w_ = 0, mb_h_ = 0, width_ = 0, height_ = 0, crop_x_ = 0, crop_y_ = 0;
  std::vector<DpbEntry> dpb_;
  int max_long_term_idx_ = -1;  // "no long-term frame indices"
  // POC state
  int prev_poc_msb_ = 0, prev_poc_lsb_ = 0, prev_frame_num_ = 0, prev_frame_num_offset_ = 0;
  bool prev_had_mmco5_ = false;
  int64_t out_period_ = 0;
  std::vector<int> levels_;  // dependency level per chunk-local frame
  int ref_level_(int f) const { return f >= 0 && f < (int)levels_.size() ? levels_[f] : 0; }
};

}  // namespace hwb
