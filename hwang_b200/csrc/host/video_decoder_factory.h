// Backend selection point.  Mirrors hwang/video_decoder_factory.h:23-38; the new enumerator B200
// selects the from-scratch CUDA backend (the reference's SOFTWARE / NVIDIA / INTEL enumerators keep
// their values so serialized configs stay meaningful, but those backends are not part of this build).
#pragma once
#include <vector>

#include "common.h"
#include "video_decoder_interface.h"

namespace hwang {

enum class VideoDecoderType {
  SOFTWARE,
  NVIDIA,
  INTEL,
  B200,
};

class VideoDecoderFactory {
 public:
  static std::vector<VideoDecoderType> get_supported_decoder_types();
  static bool has_decoder_type(VideoDecoderType type);
  static VideoDecoderInterface *make_from_config(DeviceHandle device_handle, uint32_t num_devices, VideoDecoderType type);
};

}  // namespace hwang
