// Incremental ISO-BMFF parser producing a VideoIndex.  Same pull protocol and public surface as
// hwang/mp4_index_creator.h:23-45: feed(data, size, next_offset, next_size) consumes the bytes the
// caller read at the previously returned offset and says what to read next; is_done / is_error /
// error_message / get_video_index.  Handles ftyp, moov{trak{mdia{mdhd,hdlr,minf{stbl{stsd(avc1/avcC,
// hev1|hvc1/hvcC), stsz|stz2, stsc, stco|co64, stss}}}}, mvex{trex}} and moof{traf{tfhd,trun}} fragments
// (reference: hwang/mp4_index_creator.cpp:30-755).  Differences: byte-wise readers instead of
// bit-at-a-time ones, and malformed input sets is_error() instead of exit(-1) / assert.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "video_index.h"

namespace hwang {

class MP4IndexCreator {
 public:
  explicit MP4IndexCreator(uint64_t file_size);

  // @return false if done or there was an error
  bool feed(const uint8_t *data, size_t size, uint64_t &next_offset, uint64_t &next_size);
  VideoIndex get_video_index();
  bool is_done() { return done_ || (parsed_ftyp_ && parsed_moov_ && !fragments_present_); }
  bool is_error() { return error_; }
  const std::string &error_message() { return error_message_; }

 private:
  struct Trex { uint32_t track_id, default_size, default_flags; };
  bool fail(const std::string &msg);
  bool parse_moov(const uint8_t *p, size_t n);
  bool parse_trak(const uint8_t *p, size_t n);
  bool parse_stbl(const uint8_t *p, size_t n);
  bool parse_moof(const uint8_t *p, size_t n, uint64_t moof_file_offset);

  const uint64_t file_size_;
  bool done_ = false, error_ = false;
  std::string error_message_;
  uint64_t offset_ = 0;  // file offset of the buffer the next feed() will receive
  bool parsed_ftyp_ = false, parsed_moov_ = false, fragments_present_ = false, have_video_track_ = false;
  std::vector<Trex> trex_;
  uint32_t video_track_id_ = 0;
  uint32_t timescale_ = 0;
  uint64_t duration_ = 0;
  uint32_t width_ = 0, height_ = 0;
  std::string format_;
  std::vector<uint64_t> sample_offsets_, sample_sizes_, keyframe_indices_;
  std::vector<uint8_t> extradata_;
};

}  // namespace hwang
