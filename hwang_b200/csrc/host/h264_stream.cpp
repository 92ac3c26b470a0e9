#include "h264_stream.h"

#include <string.h>
#include <algorithm>

#include "../dev/bits.h"
#include "../dev/tables_gen.h"

namespace hwb {

namespace {

// Strip emulation_prevention_three_bytes (00 00 03 -> 00 00) while appending to `out`; returns the RBSP size.
// One memchr pass over the candidates (bytes equal to 3) and block copies in between: the sample bytes of a chunk
// go through here once, straight into the buffer that is uploaded to the GPU.
size_t unescape_append(std::vector<uint8_t> &out, const uint8_t *p, size_t n) {
  const size_t base = out.size();
  out.resize(base + n);
  uint8_t *d = out.data() + base;
  size_t o = 0, i = 0, scan = 0;  // i: start of the pending copy, scan: where to look for the next 03
  while (scan < n) {
    const uint8_t *z = (const uint8_t *)memchr(p + scan, 3, n - scan);
    if (!z) break;
    const size_t k = (size_t)(z - p);
    if (k >= i + 2 && p[k - 1] == 0 && p[k - 2] == 0) {  // the two zeros must follow the previous emulation byte
      memcpy(d + o, p + i, k - i); o += k - i;
      i = k + 1;
    }
    scan = k + 1;
  }
  memcpy(d + o, p + i, n - i); o += n - i;
  out.resize(base + o);
  return o;
}
std::vector<uint8_t> unescape(const uint8_t *p, size_t n) {
  std::vector<uint8_t> out;
  unescape_append(out, p, n);
  return out;
}

struct Rd {
  BitReader b;
  uint32_t stop;  // bit position of the rbsp stop bit
  explicit Rd(const std::vector<uint8_t> &v) : Rd(v.data(), v.size()) {}
  Rd(const uint8_t *v, size_t size) {
    br_init(b, v, (uint32_t)size, 0);
    int n = (int)size;
    while (n > 0 && v[n - 1] == 0) --n;
    stop = 0;
    if (n > 0) { int tz = 0; while (!((v[n - 1] >> tz) & 1)) ++tz; stop = (uint32_t)((n - 1) * 8 + 7 - tz); }
  }
  uint32_t u(int n) { return br_get(b, n); }
  uint32_t u1() { return br_get(b, 1); }
  uint32_t ue() { return br_ue(b); }
  int32_t se() { return br_se(b); }
  bool more() { return br_bitpos(b) < stop; }
  bool bad() const { return b.overrun; }
};

// scaling_list(): returns false when useDefaultScalingMatrixFlag is set
bool read_scaling_list(Rd &r, uint8_t *raster, int n) {
  const uint8_t *zz = n == 16 ? zigzag4x4 : zigzag8x8;
  int last = 8, next = 8;
  bool use_default = false;
  for (int j = 0; j < n; ++j) {
    if (next != 0) {
      int delta = r.se();
      next = (last + delta + 256) % 256;
      if (j == 0 && next == 0) use_default = true;
    }
    int v = next == 0 ? last : next;
    raster[zz[j]] = (uint8_t)v;
    last = v;
  }
  return !use_default;
}

void set_default_list(uint8_t s4[6][16], uint8_t s8[2][64], int i) {
  if (i < 6) memcpy(s4[i], i < 3 ? default_scaling4_intra : default_scaling4_inter, 16);
  else memcpy(s8[i - 6], i == 6 ? default_scaling8_intra : default_scaling8_inter, 64);
}
void copy_list(uint8_t d4[6][16], uint8_t d8[2][64], int i, const uint8_t s4[6][16], const uint8_t s8[2][64], int j) {
  if (i < 6) memcpy(d4[i], s4[j], 16); else memcpy(d8[i - 6], s8[j - 6], 64);
}

}  // namespace

std::string H264Stream::parse_sps(const std::vector<uint8_t> &rbsp) {
  Rd r(rbsp);
  Sps s;
  s.profile_idc = (int)r.u(8); s.constraint = (int)r.u(8); s.level_idc = (int)r.u(8);
  uint32_t id = r.ue();
  if (id > 31) return "sps id out of range";
  for (int i = 0; i < 6; ++i) memset(s.scaling4[i], 16, 16);
  for (int i = 0; i < 2; ++i) memset(s.scaling8[i], 16, 64);
  const int p = s.profile_idc;
  if (p == 100 || p == 110 || p == 122 || p == 244 || p == 44 || p == 83 || p == 86 || p == 118 || p == 128 || p == 138 || p == 139 || p == 134 || p == 135) {
    s.chroma_format_idc = (int)r.ue();
    if (s.chroma_format_idc == 3) r.u1();
    s.bit_depth_luma = 8 + (int)r.ue(); s.bit_depth_chroma = 8 + (int)r.ue();
    r.u1();  // qpprime_y_zero_transform_bypass
    s.scaling_present = r.u1();
    if (s.scaling_present) {
      int nl = s.chroma_format_idc != 3 ? 8 : 12;
      for (int i = 0; i < nl; ++i) {
        bool present = r.u1();
        if (i >= 8) { if (present) { uint8_t tmp[64]; read_scaling_list(r, tmp, 64); } continue; }
        if (present) {
          bool ok = i < 6 ? read_scaling_list(r, s.scaling4[i], 16) : read_scaling_list(r, s.scaling8[i - 6], 64);
          if (!ok) set_default_list(s.scaling4, s.scaling8, i);
        } else {  // fall-back rule set A
          if (i == 0 || i == 3 || i == 6 || i == 7) set_default_list(s.scaling4, s.scaling8, i);
          else copy_list(s.scaling4, s.scaling8, i, s.scaling4, s.scaling8, i - 1);
        }
      }
    }
  }
  if (s.chroma_format_idc != 1 || s.bit_depth_luma != 8 || s.bit_depth_chroma != 8) return "unsupported: only 4:2:0 8-bit";
  s.log2_max_frame_num = 4 + (int)r.ue();
  if (s.log2_max_frame_num > 16) return "SPS: log2_max_frame_num_minus4 out of range";
  s.poc_type = (int)r.ue();
  if (s.poc_type > 2) return "SPS: pic_order_cnt_type out of range";
  if (s.poc_type == 0) { s.log2_max_poc_lsb = 4 + (int)r.ue(); if (s.log2_max_poc_lsb > 16) return "SPS: log2_max_pic_order_cnt_lsb_minus4 out of range"; }
  else if (s.poc_type == 1) {
    s.delta_pic_order_always_zero = r.u1();
    s.offset_for_non_ref_pic = r.se(); s.offset_for_top_to_bottom = r.se();
    uint32_t n = r.ue();
    if (n > 255) return "bad sps";
    for (uint32_t i = 0; i < n; ++i) s.offset_for_ref_frame.push_back(r.se());
  } else if (s.poc_type != 2) return "bad pic_order_cnt_type";
  s.max_num_ref_frames = (int)r.ue();
  s.gaps_allowed = r.u1();
  s.mb_w = (int)r.ue() + 1; s.mb_h = (int)r.ue() + 1;
  s.frame_mbs_only = r.u1();
  if (!s.frame_mbs_only) return "unsupported: interlaced / MBAFF streams (frame_mbs_only_flag = 0)";
  s.direct_8x8_inference = r.u1();
  if (r.u1()) { s.crop_l = (int)r.ue(); s.crop_r = (int)r.ue(); s.crop_t = (int)r.ue(); s.crop_b = (int)r.ue(); }
  if (r.u1()) {  // VUI
    if (r.u1()) { if (r.u(8) == 255) { r.u(16); r.u(16); } }
    if (r.u1()) r.u1();
    if (r.u1()) { r.u(3); r.u1(); if (r.u1()) { r.u(8); r.u(8); r.u(8); } }
    if (r.u1()) { r.ue(); r.ue(); }
    if (r.u1()) { r.u(32); r.u(32); r.u1(); }
    bool nal_hrd = r.u1();
    auto hrd = [&]() { uint32_t c = r.ue(); r.u(4); r.u(4); for (uint32_t i = 0; i <= c && i < 32; ++i) { r.ue(); r.ue(); r.u1(); } r.u(5); r.u(5); r.u(5); r.u(5); };
    if (nal_hrd) hrd();
    bool vcl_hrd = r.u1();
    if (vcl_hrd) hrd();
    if (nal_hrd || vcl_hrd) r.u1();
    r.u1();  // pic_struct_present
    if (r.u1()) { r.u1(); r.ue(); r.ue(); r.ue(); r.ue(); s.max_num_reorder_frames = (int)r.ue(); s.max_dec_frame_buffering = (int)r.ue(); }
  }
  if (r.bad() || s.mb_w <= 0 || s.mb_h <= 0 || s.mb_w > 1024 || s.mb_h > 1024) return "corrupt sps";
  s.valid = true;
  sps_[id] = s;
  return "";
}

std::string H264Stream::parse_pps(const std::vector<uint8_t> &rbsp) {
  Rd r(rbsp);
  Pps p;
  uint32_t id = r.ue();
  if (id > 255) return "pps id out of range";
  p.sps_id = (int)r.ue();
  if (p.sps_id > 31 || !sps_[p.sps_id].valid) return "pps refers to a missing sps";
  const Sps &sps = sps_[p.sps_id];
  p.cabac = r.u1(); p.bottom_field_pic_order = r.u1();
  p.num_slice_groups = (int)r.ue() + 1;
  if (p.num_slice_groups != 1) return "unsupported: FMO (slice groups)";
  p.num_ref_default[0] = (int)r.ue() + 1; p.num_ref_default[1] = (int)r.ue() + 1;
  p.weighted_pred = r.u1(); p.weighted_bipred_idc = (int)r.u(2);
  p.init_qp = 26 + r.se(); r.se();
  p.chroma_qp_offset[0] = p.chroma_qp_offset[1] = r.se();
  p.deblocking_control = r.u1(); p.constrained_intra = r.u1(); p.redundant_pic_cnt = r.u1();
  memcpy(p.scaling4, sps.scaling4, sizeof(p.scaling4)); memcpy(p.scaling8, sps.scaling8, sizeof(p.scaling8));
  if (r.more()) {
    p.transform8x8 = r.u1();
    p.scaling_present = r.u1();
    if (p.scaling_present) {
      int nl = 6 + (p.transform8x8 ? 2 : 0);
      for (int i = 0; i < nl; ++i) {
        bool present = r.u1();
        if (present) {
          bool ok = i < 6 ? read_scaling_list(r, p.scaling4[i], 16) : read_scaling_list(r, p.scaling8[i - 6], 64);
          if (!ok) set_default_list(p.scaling4, p.scaling8, i);
        } else if (i == 0 || i == 3 || i == 6 || i == 7) {
          // fall-back rule set B (sequence-level list), or A when the SPS carries no matrices
          if (sps.scaling_present) copy_list(p.scaling4, p.scaling8, i, sps.scaling4, sps.scaling8, i);
          else set_default_list(p.scaling4, p.scaling8, i);
        } else copy_list(p.scaling4, p.scaling8, i, p.scaling4, p.scaling8, i - 1);
      }
    }
    p.chroma_qp_offset[1] = r.se();
  }
  if (r.bad()) return "corrupt pps";
  p.valid = true;
  pps_[id] = p;
  return "";
}

std::string H264Stream::configure(const uint8_t *avcc, size_t n) {
  for (auto &s : sps_) s.valid = false;
  for (auto &p : pps_) p.valid = false;
  reset_dpb();
  if (n < 7 || avcc[0] != 1) return "extradata is not an avcC record";
  nal_length_size_ = (avcc[4] & 3) + 1;
  size_t off = 5;
  int nsps = avcc[off++] & 31;
  std::string err;
  for (int i = 0; i < nsps; ++i) {
    if (off + 2 > n) return "truncated avcC";
    size_t len = ((size_t)avcc[off] << 8) | avcc[off + 1]; off += 2;
    if (off + len > n || len < 1) return "truncated avcC";
    if ((avcc[off] & 31) == 7) { err = parse_sps(unescape(avcc + off + 1, len - 1)); if (!err.empty()) return err; }
    off += len;
  }
  if (off >= n) return "truncated avcC";
  int npps = avcc[off++];
  for (int i = 0; i < npps; ++i) {
    if (off + 2 > n) return "truncated avcC";
    size_t len = ((size_t)avcc[off] << 8) | avcc[off + 1]; off += 2;
    if (off + len > n || len < 1) return "truncated avcC";
    if ((avcc[off] & 31) == 8) { err = parse_pps(unescape(avcc + off + 1, len - 1)); if (!err.empty()) return err; }
    off += len;
  }
  for (auto &s : sps_) if (s.valid) { mb_w_ = s.mb_w; mb_h_ = s.mb_h; width_ = s.width(); height_ = s.height(); crop_x_ = 2 * s.crop_l; crop_y_ = 2 * s.crop_t; break; }
  if (!mb_w_) return "avcC carries no SPS";
  return "";
}

void H264Stream::reset_dpb() {
  dpb_.clear();
  levels_.clear();
  max_long_term_idx_ = -1;
  prev_poc_msb_ = prev_poc_lsb_ = prev_frame_num_ = prev_frame_num_offset_ = 0;
  prev_had_mmco5_ = false;
}

bool H264Stream::next_is_idr(const uint8_t *data, size_t n) const {
  size_t off = 0;
  while (off + nal_length_size_ <= n) {
    size_t len = 0;
    for (int i = 0; i < nal_length_size_; ++i) len = (len << 8) | data[off + i];
    off += nal_length_size_;
    if (len == 0) continue;  // an empty NAL unit is skipped, as parse_sample skips it
    if (off + len > n) break;
    int t = data[off] & 31;
    if (t == 5) return true;
    if (t == 1) return false;
    off += len;
  }
  return false;
}

std::string H264Stream::parse_slice_header(const uint8_t *rbsp, size_t rbsp_size, int nal_type, int nal_ref_idc, SliceHeader &sh) {
  Rd r(rbsp, rbsp_size);
  sh.nal_type = nal_type; sh.nal_ref_idc = nal_ref_idc;
  sh.first_mb = (int)r.ue();
  uint32_t st = r.ue();
  if (st > 9) return "bad slice_type";
  st %= 5;
  if (st > 2) return "unsupported: SP/SI slices";
  sh.slice_type = st == 0 ? SLICE_P : (st == 1 ? SLICE_B : SLICE_I);
  sh.pps_id = (int)r.ue();
  if (sh.pps_id > 255 || !pps_[sh.pps_id].valid) return "slice refers to a missing pps";
  const Pps &pps = pps_[sh.pps_id];
  const Sps &sps = sps_[pps.sps_id];
  sh.frame_num = (int)r.u(sps.log2_max_frame_num);
  if (nal_type == 5) sh.idr_pic_id = (int)r.ue();
  if (sps.poc_type == 0) {
    sh.poc_lsb = (int)r.u(sps.log2_max_poc_lsb);
    if (pps.bottom_field_pic_order) sh.delta_poc_bottom = r.se();
  } else if (sps.poc_type == 1 && !sps.delta_pic_order_always_zero) {
    sh.delta_poc[0] = r.se();
    if (pps.bottom_field_pic_order) sh.delta_poc[1] = r.se();
  }
  if (pps.redundant_pic_cnt) { if (r.ue() != 0) return "unsupported: redundant pictures"; }
  if (sh.slice_type == SLICE_B) sh.direct_spatial = r.u1();
  sh.num_ref[0] = sh.num_ref[1] = 0;
  if (sh.slice_type != SLICE_I) {
    sh.num_ref[0] = pps.num_ref_default[0];
    if (sh.slice_type == SLICE_B) sh.num_ref[1] = pps.num_ref_default[1];
    if (r.u1()) { sh.num_ref[0] = (int)r.ue() + 1; if (sh.slice_type == SLICE_B) sh.num_ref[1] = (int)r.ue() + 1; }
    if (sh.num_ref[0] > 32 || sh.num_ref[1] > 32) return "bad num_ref_idx_active";
    for (int l = 0; l < (sh.slice_type == SLICE_B ? 2 : 1); ++l) {
      if (r.u1()) {
        for (;;) {
          uint32_t idc = r.ue();
          if (idc == 3) break;
          if (idc > 3 || r.bad() || sh.mods[l].size() > 66) return "bad ref_pic_list_modification";
          sh.mods[l].push_back({(int)idc, (int)r.ue()});
        }
      }
    }
  }
  sh.luma_log2_denom = sh.chroma_log2_denom = 0;
  const bool explicit_w = (pps.weighted_pred && sh.slice_type == SLICE_P) || (pps.weighted_bipred_idc == 1 && sh.slice_type == SLICE_B);
  if (explicit_w) {
    sh.luma_log2_denom = (int)r.ue(); sh.chroma_log2_denom = (int)r.ue();
    if (sh.luma_log2_denom > 7 || sh.chroma_log2_denom > 7) return "bad weight denominators";
    for (int l = 0; l < (sh.slice_type == SLICE_B ? 2 : 1); ++l)
      for (int i = 0; i < sh.num_ref[l]; ++i) {
        sh.luma_w[l][i] = (int16_t)(1 << sh.luma_log2_denom); sh.luma_o[l][i] = 0;
        for (int k = 0; k < 2; ++k) { sh.chroma_w[l][i][k] = (int16_t)(1 << sh.chroma_log2_denom); sh.chroma_o[l][i][k] = 0; }
        if (r.u1()) { sh.luma_w[l][i] = (int16_t)r.se(); sh.luma_o[l][i] = (int16_t)r.se(); }
        if (r.u1()) for (int k = 0; k < 2; ++k) { sh.chroma_w[l][i][k] = (int16_t)r.se(); sh.chroma_o[l][i][k] = (int16_t)r.se(); }
      }
  }
  if (nal_ref_idc) {
    if (nal_type == 5) { sh.no_output_of_prior = r.u1(); sh.long_term_ref = r.u1(); }
    else {
      sh.adaptive_marking = r.u1();
      if (sh.adaptive_marking) {
        for (;;) {
          uint32_t op = r.ue();
          if (op == 0) break;
          if (op > 6 || r.bad() || sh.mmco.size() > 66) return "bad dec_ref_pic_marking";
          Mmco m{(int)op, 0, 0};
          if (op == 1 || op == 3) m.a = (int)r.ue();
          if (op == 2) m.a = (int)r.ue();
          if (op == 3 || op == 6) m.b = (int)r.ue();
          if (op == 4) m.a = (int)r.ue();
          sh.mmco.push_back(m);
        }
      }
    }
  }
  if (pps.cabac && sh.slice_type != SLICE_I) { sh.cabac_init_idc = (int)r.ue(); if (sh.cabac_init_idc > 2) return "bad cabac_init_idc"; }
  sh.qp = pps.init_qp + r.se();
  if (sh.qp < 0 || sh.qp > 51) return "bad slice qp";
  sh.disable_deblock = 0; sh.alpha_off = sh.beta_off = 0;
  if (pps.deblocking_control) {
    sh.disable_deblock = (int)r.ue();
    if (sh.disable_deblock > 2) return "bad disable_deblocking_filter_idc";
    if (sh.disable_deblock != 1) { sh.alpha_off = 2 * r.se(); sh.beta_off = 2 * r.se(); }
  }
  if (r.bad()) return "truncated slice header";
  sh.data_bit_off = br_bitpos(r.b);
  return "";
}

int H264Stream::compute_poc(const Sps &sps, const SliceHeader &sh) {
  const bool idr = sh.nal_type == 5;
  const int max_fn = 1 << sps.log2_max_frame_num;
  int frame_num_offset = 0;
  if (!idr) {
    int prev_off = prev_had_mmco5_ ? 0 : prev_frame_num_offset_;
    int prev_fn = prev_had_mmco5_ ? 0 : prev_frame_num_;
    frame_num_offset = prev_fn > sh.frame_num ? prev_off + max_fn : prev_off;
  }
  int poc;
  if (sps.poc_type == 0) {
    int prev_msb = prev_poc_msb_, prev_lsb = prev_poc_lsb_;
    if (idr) prev_msb = prev_lsb = 0;
    const int max_lsb = 1 << sps.log2_max_poc_lsb;
    int msb;
    if (sh.poc_lsb < prev_lsb && prev_lsb - sh.poc_lsb >= max_lsb / 2) msb = prev_msb + max_lsb;
    else if (sh.poc_lsb > prev_lsb && sh.poc_lsb - prev_lsb > max_lsb / 2) msb = prev_msb - max_lsb;
    else msb = prev_msb;
    int top = msb + sh.poc_lsb, bottom = top + sh.delta_poc_bottom;
    poc = std::min(top, bottom);
    if (sh.nal_ref_idc) { prev_poc_msb_ = msb; prev_poc_lsb_ = sh.poc_lsb; }
  } else if (sps.poc_type == 1) {
    const int n = (int)sps.offset_for_ref_frame.size();
    int abs_fn = n ? frame_num_offset + sh.frame_num : 0;
    if (!sh.nal_ref_idc && abs_fn > 0) abs_fn--;
    int expected = 0;
    if (abs_fn > 0) {
      int cycle = (abs_fn - 1) / n, in_cycle = (abs_fn - 1) % n, per = 0;
      for (int v : sps.offset_for_ref_frame) per += v;
      expected = cycle * per;
      for (int i = 0; i <= in_cycle; ++i) expected += sps.offset_for_ref_frame[i];
    }
    if (!sh.nal_ref_idc) expected += sps.offset_for_non_ref_pic;
    int top = expected + sh.delta_poc[0], bottom = top + sps.offset_for_top_to_bottom + sh.delta_poc[1];
    poc = std::min(top, bottom);
  } else {
    poc = idr ? 0 : 2 * (frame_num_offset + sh.frame_num) - (sh.nal_ref_idc ? 0 : 1);
  }
  prev_frame_num_offset_ = frame_num_offset;
  prev_frame_num_ = sh.frame_num;
  return poc;
}

void H264Stream::build_ref_lists(const Sps &sps, const SliceHeader &sh, int cur_poc, std::vector<DpbEntry> lists[2]) {
  const int max_fn = 1 << sps.log2_max_frame_num;
  std::vector<DpbEntry> st, lt;
  for (auto e : dpb_) {
    if (e.long_term) { e.pic_num = e.long_term_idx; lt.push_back(e); }
    else { e.frame_num_wrap = e.frame_num > sh.frame_num ? e.frame_num - max_fn : e.frame_num; e.pic_num = e.frame_num_wrap; st.push_back(e); }
  }
  std::sort(lt.begin(), lt.end(), [](const DpbEntry &a, const DpbEntry &b) { return a.long_term_idx < b.long_term_idx; });
  if (sh.slice_type == SLICE_P) {
    std::sort(st.begin(), st.end(), [](const DpbEntry &a, const DpbEntry &b) { return a.pic_num > b.pic_num; });
    lists[0] = st; lists[0].insert(lists[0].end(), lt.begin(), lt.end());
  } else {
    std::vector<DpbEntry> before, after;
    for (auto &e : st) (e.poc < cur_poc ? before : after).push_back(e);
    std::sort(before.begin(), before.end(), [](const DpbEntry &a, const DpbEntry &b) { return a.poc > b.poc; });
    std::sort(after.begin(), after.end(), [](const DpbEntry &a, const DpbEntry &b) { return a.poc < b.poc; });
    lists[0] = before; lists[0].insert(lists[0].end(), after.begin(), after.end()); lists[0].insert(lists[0].end(), lt.begin(), lt.end());
    lists[1] = after; lists[1].insert(lists[1].end(), before.begin(), before.end()); lists[1].insert(lists[1].end(), lt.begin(), lt.end());
    if (lists[1].size() > 1 && lists[0].size() == lists[1].size()) {
      bool same = true;
      for (size_t i = 0; i < lists[0].size(); ++i) same &= lists[0][i].frame == lists[1][i].frame;
      if (same) std::swap(lists[1][0], lists[1][1]);
    }
  }
  // A slice may name more active entries than there are reference pictures (encoders that rely on the PPS default do
  // so for the first pictures of a GOP).  The standard gives such entries "no reference picture"; libavcodec, which the
  // reference delegates to, lets them stand for the initial list's first entry (h264_refs.c: default_ref), and so do we:
  // a stream that uses them decodes to the same samples.  A list with no picture at all stays short (refused by the caller).
  for (int l = 0; l < 2; ++l) {
    const int nact = sh.num_ref[l];
    const bool have_dflt = !lists[l].empty();
    const DpbEntry dflt = have_dflt ? lists[l][0] : DpbEntry{};
    auto pad = [&]() {
      if (!have_dflt) return;
      if ((int)lists[l].size() < nact) lists[l].resize(nact, dflt);
      for (auto &e : lists[l]) if (e.frame < 0) e = dflt;
    };
    if ((int)lists[l].size() > nact) lists[l].resize(nact);
    if (sh.mods[l].empty()) { pad(); continue; }
    // 8.2.4.3: the list temporarily holds one extra entry
    std::vector<DpbEntry> &L = lists[l];
    DpbEntry none{}; none.frame = -1; none.pic_num = INT32_MIN; none.long_term = false;
    L.resize(nact + 1, none);
    int pred = sh.frame_num, idx = 0;
    for (auto &m : sh.mods[l]) {
      if (idx >= nact) break;
      const DpbEntry *pick = nullptr;
      bool want_long = m.idc == 2;
      int target;
      if (!want_long) {
        int d = m.val + 1, nowrap;
        if (m.idc == 0) { nowrap = pred - d; if (nowrap < 0) nowrap += max_fn; }
        else { nowrap = pred + d; if (nowrap >= max_fn) nowrap -= max_fn; }
        pred = nowrap;
        target = nowrap > sh.frame_num ? nowrap - max_fn : nowrap;
        for (auto &e : st) if (e.pic_num == target) pick = &e;
      } else {
        target = m.val;
        for (auto &e : lt) if (e.pic_num == target) pick = &e;
      }
      if (!pick) continue;  // reference missing: leave the list as it is (corrupt stream)
      for (int c = nact; c > idx; --c) L[c] = L[c - 1];
      L[idx++] = *pick;
      int nidx = idx;
      for (int c = idx; c <= nact; ++c)
        if (!(L[c].frame >= 0 && L[c].long_term == want_long && L[c].pic_num == target)) L[nidx++] = L[c];
    }
    L.resize(nact);
    pad();
  }
}

void H264Stream::mark_references(const Sps &sps, const SliceHeader &sh, int cur_frame, int cur_poc) {
  const int max_fn = 1 << sps.log2_max_frame_num;
  DpbEntry cur{};
  cur.frame = cur_frame; cur.frame_num = sh.frame_num; cur.poc = cur_poc; cur.long_term = false; cur.long_term_idx = -1;
  prev_had_mmco5_ = false;
  if (sh.nal_type == 5) {
    dpb_.clear();
    max_long_term_idx_ = -1;
    if (sh.long_term_ref) { cur.long_term = true; cur.long_term_idx = 0; max_long_term_idx_ = 0; }
    dpb_.push_back(cur);
    return;
  }
  auto pic_num_of = [&](const DpbEntry &e) { return e.frame_num > sh.frame_num ? e.frame_num - max_fn : e.frame_num; };
  if (!sh.adaptive_marking) {
    int limit = std::max(sps.max_num_ref_frames, 1);
    if ((int)dpb_.size() >= limit) {
      int best = -1, bestw = INT32_MAX;
      for (size_t i = 0; i < dpb_.size(); ++i)
        if (!dpb_[i].long_term) { int w = pic_num_of(dpb_[i]); if (w < bestw) { bestw = w; best = (int)i; } }
      if (best >= 0) dpb_.erase(dpb_.begin() + best);
    }
  } else {
    for (auto &m : sh.mmco) {
      switch (m.op) {
        case 1: {
          int pn = sh.frame_num - (m.a + 1);
          for (size_t i = 0; i < dpb_.size(); ++i) if (!dpb_[i].long_term && pic_num_of(dpb_[i]) == pn) { dpb_.erase(dpb_.begin() + i); break; }
          break;
        }
        case 2:
          for (size_t i = 0; i < dpb_.size(); ++i) if (dpb_[i].long_term && dpb_[i].long_term_idx == m.a) { dpb_.erase(dpb_.begin() + i); break; }
          break;
        case 3: {
          int pn = sh.frame_num - (m.a + 1);
          for (size_t i = 0; i < dpb_.size(); ++i) if (dpb_[i].long_term && dpb_[i].long_term_idx == m.b) { dpb_.erase(dpb_.begin() + i); break; }
          for (auto &e : dpb_) if (!e.long_term && pic_num_of(e) == pn) { e.long_term = true; e.long_term_idx = m.b; break; }
          break;
        }
        case 4:
          max_long_term_idx_ = m.a - 1;
          for (size_t i = 0; i < dpb_.size();) { if (dpb_[i].long_term && dpb_[i].long_term_idx > max_long_term_idx_) dpb_.erase(dpb_.begin() + i); else ++i; }
          break;
        case 5:
          dpb_.clear(); max_long_term_idx_ = -1; prev_had_mmco5_ = true;
          break;
        case 6:
          for (size_t i = 0; i < dpb_.size(); ++i) if (dpb_[i].long_term && dpb_[i].long_term_idx == m.b) { dpb_.erase(dpb_.begin() + i); break; }
          cur.long_term = true; cur.long_term_idx = m.b;
          break;
      }
    }
    if (prev_had_mmco5_) {
      cur.frame_num = 0; cur.poc = 0;
      prev_frame_num_ = 0; prev_frame_num_offset_ = 0; prev_poc_msb_ = 0; prev_poc_lsb_ = 0;
    }
    // never exceed the DPB capacity even on streams that forget to free a frame
    int limit = std::max(sps.max_num_ref_frames, 1);
    while ((int)dpb_.size() >= limit) {
      int best = -1, bestw = INT32_MAX;
      for (size_t i = 0; i < dpb_.size(); ++i) if (!dpb_[i].long_term) { int w = pic_num_of(dpb_[i]); if (w < bestw) { bestw = w; best = (int)i; } }
      if (best < 0) best = 0;
      dpb_.erase(dpb_.begin() + best);
    }
  }
  dpb_.push_back(cur);
}

std::string H264Stream::parse_sample(const uint8_t *data, size_t n, int pic_index, std::vector<uint8_t> &bitstream, PlannedPic &out) {
  size_t off = 0;
  bool have_pic = false;
  SliceHeader first;
  const Pps *pps0 = nullptr;
  const Sps *sps0 = nullptr;
  int cur_poc = 0;
  out.slices.clear();
  out.idr = false;
  while (off + nal_length_size_ <= n) {
    size_t len = 0;
    for (int i = 0; i < nal_length_size_; ++i) len = (len << 8) | data[off + i];
    off += nal_length_size_;
    if (len == 0) continue;
    if (off + len > n) return "sample truncated: NAL length exceeds the sample";
    const uint8_t *nal = data + off;
    off += len;
    if (nal[0] & 0x80) return "forbidden_zero_bit set";
    const int nal_ref_idc = (nal[0] >> 5) & 3, nal_type = nal[0] & 31;
    if (nal_type == 7) { std::string e = parse_sps(unescape(nal + 1, len - 1)); if (!e.empty()) return e; continue; }
    if (nal_type == 8) { std::string e = parse_pps(unescape(nal + 1, len - 1)); if (!e.empty()) return e; continue; }
    if (nal_type >= 2 && nal_type <= 4) return "unsupported: data partitioning";
    if (nal_type != 1 && nal_type != 5) continue;
    // the slice RBSP goes straight to the chunk bitstream, 16-byte aligned (dropped again if the header is bad)
    while (bitstream.size() & 15) bitstream.push_back(0);
    const size_t rbsp_off = bitstream.size();
    const size_t rbsp_size = unescape_append(bitstream, nal + 1, len - 1);
    SliceHeader sh;
    std::string e = parse_slice_header(bitstream.data() + rbsp_off, rbsp_size, nal_type, nal_ref_idc, sh);
    if (!e.empty()) { bitstream.resize(rbsp_off); return e; }
    const Pps &pps = pps_[sh.pps_id];
    const Sps &sps = sps_[pps.sps_id];
    if (sps.mb_w != mb_w_ || sps.mb_h != mb_h_) return "unsupported: resolution change inside a stream";
    if (!have_pic) {
      have_pic = true;
      first = sh; pps0 = &pps; sps0 = &sps;
      if (nal_type == 5) { out.idr = true; out_period_++; }
      cur_poc = compute_poc(sps, sh);
      PicDesc &d = out.desc;
      memset(&d, 0, sizeof(d));
      d.frame = pic_index; d.poc = cur_poc;
      d.cabac = pps.cabac; d.transform8x8_mode = pps.transform8x8; d.constrained_intra_pred = pps.constrained_intra;
      d.direct_8x8_inference = sps.direct_8x8_inference; d.weighted_pred = pps.weighted_pred; d.weighted_bipred_idc = (uint8_t)pps.weighted_bipred_idc;
      d.is_ref = nal_ref_idc != 0; d.has_inter = 0;
      d.chroma_qp_offset[0] = (int8_t)pps.chroma_qp_offset[0]; d.chroma_qp_offset[1] = (int8_t)pps.chroma_qp_offset[1];
      memcpy(d.scaling4, pps.scaling4, sizeof(d.scaling4)); memcpy(d.scaling8, pps.scaling8, sizeof(d.scaling8));
      d.level = 0;
    } else if (sh.frame_num != first.frame_num || sh.pps_id != first.pps_id || (nal_type == 5) != (first.nal_type == 5)) {
      return "unsupported: more than one picture in a sample";
    }
    if (sh.first_mb >= mb_w_ * mb_h_) return "first_mb_in_slice out of range";
    if (sh.slice_type == SLICE_P && out.desc.has_inter < 1) out.desc.has_inter = 1;
    if (sh.slice_type == SLICE_B) out.desc.has_inter = 2;
    SliceDesc sd;
    memset(&sd, 0, sizeof(sd));
    sd.pic = pic_index; sd.first_mb = sh.first_mb;
    sd.slice_type = (uint8_t)sh.slice_type; sd.qp = (uint8_t)sh.qp; sd.cabac_init_idc = (uint8_t)sh.cabac_init_idc;
    sd.disable_deblock = (uint8_t)sh.disable_deblock; sd.alpha_off = (int8_t)sh.alpha_off; sd.beta_off = (int8_t)sh.beta_off;
    sd.num_ref[0] = (uint8_t)sh.num_ref[0]; sd.num_ref[1] = (uint8_t)sh.num_ref[1];
    sd.direct_spatial = sh.direct_spatial;
    sd.luma_log2_denom = (uint8_t)sh.luma_log2_denom; sd.chroma_log2_denom = (uint8_t)sh.chroma_log2_denom;
    sd.use_weights = 0;
    if (sh.slice_type == SLICE_P && pps.weighted_pred) sd.use_weights = 1;
    if (sh.slice_type == SLICE_B) sd.use_weights = (uint8_t)pps.weighted_bipred_idc;
    if (sd.use_weights == 1) { memcpy(sd.luma_w, sh.luma_w, sizeof(sd.luma_w)); memcpy(sd.luma_o, sh.luma_o, sizeof(sd.luma_o)); memcpy(sd.chroma_w, sh.chroma_w, sizeof(sd.chroma_w)); memcpy(sd.chroma_o, sh.chroma_o, sizeof(sd.chroma_o)); }
    if (sh.slice_type != SLICE_I) {
      std::vector<DpbEntry> lists[2];
      build_ref_lists(sps, sh, cur_poc, lists);
      for (int l = 0; l < 2; ++l) {
        if ((int)lists[l].size() < sh.num_ref[l]) return "reference picture missing (broken link / open GOP at an interval start?)";
        for (int i = 0; i < sh.num_ref[l]; ++i) {
          if (lists[l][i].frame < 0) return "reference picture missing";
          sd.ref_frame[l][i] = (int16_t)lists[l][i].frame; sd.ref_poc[l][i] = lists[l][i].poc;
          if (lists[l][i].long_term) sd.ref_long[l] |= 1u << i;
          out.desc.level = std::max(out.desc.level, ref_level_(lists[l][i].frame) + 1);
        }
      }
    }
    sd.data_off = (uint32_t)rbsp_off; sd.data_size = (uint32_t)rbsp_size; sd.bit_off = sh.data_bit_off;
    out.slices.push_back(sd);
  }
  if (!have_pic) return "sample contains no slice";
  std::sort(out.slices.begin(), out.slices.end(), [](const SliceDesc &a, const SliceDesc &b) { return a.first_mb < b.first_mb; });
  if (out.slices[0].first_mb != 0) return "unsupported: picture does not start at macroblock 0 (ASO / lost slice)";
  out.desc.num_slices = (int)out.slices.size();
  // Every macroblock must be covered by exactly one slice: reconstruction and deblocking read the entropy stage's
  // per-macroblock records of the whole picture, so a lost or duplicated slice is refused here and a slice that ends
  // early or late is caught by the entropy stage (SliceDesc::end_mb).
  for (size_t i = 0; i < out.slices.size(); ++i) {
    const int next = i + 1 < out.slices.size() ? out.slices[i + 1].first_mb : mb_w_ * mb_h_;
    if (next <= out.slices[i].first_mb) return "unsupported: two slices start at the same macroblock (redundant slices)";
    out.slices[i].end_mb = next;
  }
  // distinct reference frames of the picture (what the picture kernel waits for)
  out.desc.num_dep = 0;
  for (auto &sd : out.slices)
    for (int l = 0; l < 2; ++l)
      for (int i = 0; i < sd.num_ref[l] && sd.slice_type != SLICE_I; ++i) {
        const int16_t f = sd.ref_frame[l][i];
        bool seen = false;
        for (int k = 0; k < out.desc.num_dep; ++k) seen |= out.desc.dep[k] == f;
        if (!seen) {
          if (out.desc.num_dep >= 32) return "unsupported: more than 32 distinct reference frames in one picture";
          out.desc.dep[out.desc.num_dep++] = f;
        }
      }
  out.desc.rgb_slot = -1;
  levels_.resize(std::max<size_t>(levels_.size(), (size_t)pic_index + 1));
  levels_[pic_index] = out.desc.level;
  // mark_references leaves prev_had_mmco5_ = "THIS picture carried MMCO 5" (what compute_poc of the next picture
  // needs); a non-reference picture has no marking at all, so the flag is cleared for it here -- otherwise every
  // non-reference picture after an MMCO 5 picture would start an output period of its own
  if (first.nal_ref_idc) mark_references(*sps0, first, pic_index, cur_poc);
  else prev_had_mmco5_ = false;
  if (prev_had_mmco5_) { out_period_++; cur_poc = 0; }
  out.out_key = (out_period_ << 32) + (int64_t)cur_poc + (1ll << 31);
  (void)pps0;
  return "";
}

}  // namespace hwb
