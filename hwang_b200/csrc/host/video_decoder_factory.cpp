#include "video_decoder_factory.h"

#include "b200_video_decoder.h"

namespace hwang {

// reference: hwang/video_decoder_factory.cpp:31-42
std::vector<VideoDecoderType> VideoDecoderFactory::get_supported_decoder_types() {
  std::vector<VideoDecoderType> t;
  t.push_back(VideoDecoderType::B200);
  return t;
}

// reference: hwang/video_decoder_factory.cpp:44-53
bool VideoDecoderFactory::has_decoder_type(VideoDecoderType type) {
  for (auto t : get_supported_decoder_types()) if (t == type) return true;
  return false;
}

// reference: hwang/video_decoder_factory.cpp:55-98.  Python maps DeviceType.GPU to NVIDIA
// (python/hwang/decoder.py:25-28); both NVIDIA and B200 therefore select this backend.  There is no
// CPU fallback: SOFTWARE / INTEL return nullptr, and B200 returns nullptr when the device cannot be opened.
VideoDecoderInterface *VideoDecoderFactory::make_from_config(DeviceHandle device_handle, uint32_t num_devices, VideoDecoderType type) {
  switch (type) {
    case VideoDecoderType::NVIDIA:
    case VideoDecoderType::B200: {
      if (device_handle.type != DeviceType::GPU) return nullptr;
      // as the reference does for its GPU backend (video_decoder_factory.cpp:76-77), the handle's type is the output type
      B200VideoDecoder *d = new B200VideoDecoder(device_handle.id, device_handle.type, num_devices);
      if (!d->ok()) { delete d; return nullptr; }
      return d;
    }
    default:
      return nullptr;
  }
}

}  // namespace hwang
