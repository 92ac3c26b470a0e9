// Sparse-frame scheduler: feeds the samples of keyframe-delimited intervals to one decoder backend on
// a feeder thread and hands the wanted frames (display order) to the caller.  Same public API and
// observable contract as hwang/decoder_automata.h:33-70 / decoder_automata.cpp:80-252 of the reference:
//  * make_instance returns nullptr if the factory cannot build the requested backend;
//  * initialize(intervals, extradata) may be called any number of times and fully resets state;
//  * successive get_frames(buf, n) calls return the wanted frames of interval 0, then interval 1, ... in
//    ascending order, n at a time (a call may straddle an interval boundary), as n tightly packed
//    width*height*3 RGB24 frames at buffer + k*frame_size;
//  * the k-th frame popped from the decoder after an interval start is absolute frame start_keyframe + k.
// Internally it is simpler than the reference (no seeking flags, no per-packet stderr print): the feeder
// streams every interval in order and never waits for the consumer except through the backend's
// decoded_frames_buffered() back-pressure, so the GPU decodes ahead while frames are being returned.
#pragma once
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "video_decoder_factory.h"
#include "video_decoder_interface.h"

namespace hwang {

class DecoderAutomata {
  DecoderAutomata() = delete;
  DecoderAutomata(const DecoderAutomata &) = delete;
  DecoderAutomata(DeviceHandle device_handle, int32_t num_devices, VideoDecoderType decoder_type, VideoDecoderInterface *decoder);

 public:
  static DecoderAutomata *make_instance(DeviceHandle device_handle, int32_t num_devices, VideoDecoderType decoder_type);
  // test hook: drive an arbitrary backend (e.g. a frame-number-stamping fake)
  static DecoderAutomata *make_with_decoder(VideoDecoderInterface *decoder);
  ~DecoderAutomata();

  struct EncodedData {
    bool operator==(const EncodedData &o) const {
      return encoded_video == o.encoded_video && width == o.width && height == o.height && start_keyframe == o.start_keyframe &&
             end_keyframe == o.end_keyframe && format == o.format && sample_offsets == o.sample_offsets &&
             sample_sizes == o.sample_sizes && keyframes == o.keyframes && valid_frames == o.valid_frames;
    }
    std::vector<uint8_t> encoded_video;
    uint32_t width = 0;
    uint32_t height = 0;
    uint64_t start_keyframe = 0;
    uint64_t end_keyframe = 0;
    std::string format;
    std::vector<uint64_t> sample_offsets;  // relative to encoded_video
    std::vector<uint64_t> sample_sizes;
    std::vector<uint64_t> keyframes;     // absolute frame numbers, first = start_keyframe
    std::vector<uint64_t> valid_frames;  // absolute, ascending
  };

  Result initialize(const std::vector<EncodedData> &encoded_data, const std::vector<uint8_t> &extradata);
  // same, taking ownership of the interval list (no copy of encoded_video)
  Result initialize(std::vector<EncodedData> &&encoded_data, const std::vector<uint8_t> &extradata);
  // Parks the feeder and hands back the intervals of the previous initialize() so that a caller about to copy new
  // ones can reuse their buffers (a fresh 90 MB vector costs ~25 ms of page faults).  The automaton is left empty.
  std::vector<EncodedData> release_intervals();
  Result get_frames(uint8_t *buffer, int32_t num_frames);
  VideoDecoderInterface *decoder() { return decoder_.get(); }

 private:
  void feeder();
  void stop_feeder();  // abort the current feed pass and wait until the feeder is parked
  static uint64_t fed_samples(const EncodedData &d);  // how many samples of an interval are fed (see the .cpp)
  Result validate(const std::vector<EncodedData> &encoded_data) const;

  const int32_t MAX_BUFFERED_FRAMES = 8;  // reference: decoder_automata.cpp:288

  DeviceHandle device_handle_;
  int32_t num_devices_;
  VideoDecoderType decoder_type_;
  std::unique_ptr<VideoDecoderInterface> decoder_;

  std::thread feeder_thread_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool quit_ = false;        // guarded by mu_
  bool work_ = false;        // a feed pass has been requested
  bool parked_ = true;       // feeder idle
  std::atomic<bool> abort_{false};

  VideoDecoderInterface::FrameInfo info_{};
  size_t frame_size_ = 0;
  std::vector<EncodedData> encoded_data_;

  // consumer cursor
  size_t interval_ = 0;       // interval the next popped frame belongs to
  uint64_t popped_ = 0;       // frames popped from the current interval
  size_t valid_idx_ = 0;      // next wanted frame of the current interval

  std::atomic<bool> result_set_{false};
  Result feeder_result_;
};

}  // namespace hwang
