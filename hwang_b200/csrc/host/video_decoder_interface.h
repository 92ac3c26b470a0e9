// The decoder plugin contract: the drop-in boundary.  Same seven virtuals, argument meaning and
// error behaviour as hwang/video_decoder_interface.h:26-55 of the reference.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "common.h"

namespace hwang {

class VideoDecoderInterface {
 public:
  virtual ~VideoDecoderInterface() {}

  struct FrameInfo {
    uint32_t width;
    uint32_t height;
    std::string format;
  };
  // extradata = avcC bytes (VideoIndex::metadata_bytes)
  virtual Result configure(const FrameInfo &metadata, const std::vector<uint8_t> &extradata) = 0;
  // one MP4 sample in decode order; (NULL, 0, false) = end of interval
  virtual Result feed(const uint8_t *encoded_buffer, size_t encoded_size, bool keyframe) = 0;
  virtual Result flush() = 0;
  virtual Result discard_frame() = 0;
  // oldest decoded frame (display order) -> RGB24, width*height*3 bytes, host memory
  virtual Result get_frame(uint8_t *decoded_buffer, size_t decoded_size) = 0;
  virtual int decoded_frames_buffered() = 0;
  virtual Result wait_until_frames_copied() = 0;
};

}  // namespace hwang
