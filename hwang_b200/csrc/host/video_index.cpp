#include "video_index.h"

namespace hwang {

namespace {
void put_varint(std::vector<uint8_t> &o, uint64_t v) {
  while (v >= 0x80) { o.push_back((uint8_t)(v | 0x80)); v >>= 7; }
  o.push_back((uint8_t)v);
}
size_t varint_size(uint64_t v) { size_t n = 1; while (v >= 0x80) { v >>= 7; ++n; } return n; }
bool get_varint(const std::vector<uint8_t> &d, size_t &pos, uint64_t &v) {
  v = 0;
  for (int shift = 0; shift < 64 && pos < d.size(); shift += 7) {
    uint8_t b = d[pos++];
    v |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) return true;
  }
  return false;
}
void put_packed(std::vector<uint8_t> &o, int field, const std::vector<uint64_t> &v) {
  if (v.empty()) return;
  put_varint(o, (uint64_t)(field << 3 | 2));
  size_t n = 0;
  for (uint64_t x : v) n += varint_size(x);
  put_varint(o, n);
  for (uint64_t x : v) put_varint(o, x);
}
}  // namespace

// proto3 canonical encoding: fields in field-number order, zero / empty values omitted
std::vector<uint8_t> VideoIndex::serialize() const {
  std::vector<uint8_t> o;
  if (frame_width_) { put_varint(o, 1 << 3 | 0); put_varint(o, frame_width_); }
  if (frame_height_) { put_varint(o, 2 << 3 | 0); put_varint(o, frame_height_); }
  put_packed(o, 3, sample_offsets_);
  put_packed(o, 4, sample_sizes_);
  put_packed(o, 5, keyframe_indices_);
  if (!metadata_bytes_.empty()) { put_varint(o, 6 << 3 | 2); put_varint(o, metadata_bytes_.size()); o.insert(o.end(), metadata_bytes_.begin(), metadata_bytes_.end()); }
  if (timescale_) { put_varint(o, 7 << 3 | 0); put_varint(o, timescale_); }
  if (duration_) { put_varint(o, 8 << 3 | 0); put_varint(o, duration_); }
  if (!format_.empty()) { put_varint(o, 9 << 3 | 2); put_varint(o, format_.size()); o.insert(o.end(), format_.begin(), format_.end()); }
  return o;
}

VideoIndex VideoIndex::deserialize(const std::vector<uint8_t> &d) {
  uint32_t timescale = 0, w = 0, h = 0;
  uint64_t duration = 0;
  std::string format;
  std::vector<uint64_t> offs, sizes, keys;
  std::vector<uint8_t> meta;
  size_t pos = 0;
  while (pos < d.size()) {
    uint64_t tag, v;
    if (!get_varint(d, pos, tag)) break;
    int field = (int)(tag >> 3), wt = (int)(tag & 7);
    if (wt == 0) {
      if (!get_varint(d, pos, v)) break;
      if (field == 1) w = (uint32_t)v; else if (field == 2) h = (uint32_t)v; else if (field == 7) timescale = (uint32_t)v; else if (field == 8) duration = v;
      else if (field == 3) offs.push_back(v); else if (field == 4) sizes.push_back(v); else if (field == 5) keys.push_back(v);
    } else if (wt == 2) {
      if (!get_varint(d, pos, v) || pos + v > d.size()) break;
      size_t end = pos + (size_t)v;
      if (field == 3 || field == 4 || field == 5) {
        std::vector<uint64_t> &dst = field == 3 ? offs : (field == 4 ? sizes : keys);
        while (pos < end) { uint64_t x; if (!get_varint(d, pos, x)) break; dst.push_back(x); }
      } else if (field == 6) meta.assign(d.begin() + pos, d.begin() + end);
      else if (field == 9) format.assign(d.begin() + pos, d.begin() + end);
      pos = end;
    } else if (wt == 1) pos += 8;
    else if (wt == 5) pos += 4;
    else break;
  }
  return VideoIndex(timescale, duration, w, h, format, offs, sizes, keys, meta);
}

VideoIntervals slice_into_video_intervals(const VideoIndex &index, const std::vector<uint64_t> &rows) {
  VideoIntervals info;
  std::vector<uint64_t> kf = index.keyframe_indices();
  kf.push_back(index.frames());
  if (kf.size() < 2 || rows.empty()) return info;
  const auto &offs = index.sample_offsets();
  const auto &sizes = index.sample_sizes();
  size_t start = 0, end = 1;
  uint64_t next_keyframe = kf[end];
  std::vector<uint64_t> valid;
  for (uint64_t row : rows) {
    if (row >= index.frames()) return VideoIntervals();
    if (row >= next_keyframe) {
      // is the GOP that starts at next_keyframe byte-adjacent to the one before it?
      bool adjacent = offs[next_keyframe - 1] + sizes[next_keyframe - 1] == offs[next_keyframe];
      next_keyframe = kf[++end];
      if (row >= next_keyframe || !adjacent) {
        // skipped a keyframe, or not adjacent: close the interval
        if (!valid.empty()) {
          info.sample_index_intervals.push_back(std::make_tuple((size_t)kf[start], (size_t)kf[end - 1]));
          info.valid_frames.push_back(valid);
        }
        while (row >= kf[end]) end++;
        valid.clear();
        start = end - 1;
        next_keyframe = kf[end];
      }
    }
    valid.push_back(row);
  }
  info.sample_index_intervals.push_back(std::make_tuple((size_t)kf[start], (size_t)kf[end]));
  info.valid_frames.push_back(valid);
  return info;
}

}  // namespace hwang
