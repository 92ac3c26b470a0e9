// ORACLE build helper: a do-nothing stand-in for the protobuf-generated header so that the reference's
// hwang/video_index.cpp (which we only want for slice_into_video_intervals) compiles without protoc.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
namespace hwang { namespace proto {
class VideoIndex {
 public:
  bool ParseFromArray(const void *, int) { return false; }
  uint32_t timescale() const { return 0; }
  uint64_t duration() const { return 0; }
  uint32_t frame_width() const { return 0; }
  uint32_t frame_height() const { return 0; }
  const std::string &format() const { return s_; }
  const std::vector<uint64_t> &sample_offsets() const { return v_; }
  const std::vector<uint64_t> &sample_sizes() const { return v_; }
  const std::vector<uint64_t> &keyframe_indices() const { return v_; }
  const std::string &metadata_bytes() const { return s_; }
  void set_timescale(uint32_t) {} void set_duration(uint64_t) {} void set_frame_width(uint32_t) {} void set_frame_height(uint32_t) {}
  void set_format(const std::string &) {} void add_sample_offsets(uint64_t) {} void add_sample_sizes(uint64_t) {} void add_keyframe_indices(uint64_t) {}
  void set_metadata_bytes(const void *, size_t) {}
  size_t ByteSizeLong() const { return 0; }
  bool SerializeToArray(void *, int) const { return true; }
 private:
  std::string s_; std::vector<uint64_t> v_;
};
}}
