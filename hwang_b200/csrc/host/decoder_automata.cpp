#include "decoder_automata.h"

#include "b200_video_decoder.h"

#include <chrono>

namespace hwang {

DecoderAutomata *DecoderAutomata::make_instance(DeviceHandle device_handle, int32_t num_devices, VideoDecoderType decoder_type) {
  // reference: decoder_automata.cpp:44-53
  VideoDecoderInterface *decoder = VideoDecoderFactory::make_from_config(device_handle, num_devices, decoder_type);
  if (decoder == nullptr) return nullptr;
  return new DecoderAutomata(device_handle, num_devices, decoder_type, decoder);
}

DecoderAutomata *DecoderAutomata::make_with_decoder(VideoDecoderInterface *decoder) {
  if (decoder == nullptr) return nullptr;
  return new DecoderAutomata(CPU_DEVICE, 1, VideoDecoderType::B200, decoder);
}

DecoderAutomata::DecoderAutomata(DeviceHandle device_handle, int32_t num_devices, VideoDecoderType decoder_type, VideoDecoderInterface *decoder)
    : device_handle_(device_handle), num_devices_(num_devices), decoder_type_(decoder_type), decoder_(decoder) {
  // our own backend applies back-pressure in bytes of device memory inside feed(): legal here, feed() runs on the feeder thread
  if (B200VideoDecoder *b = dynamic_cast<B200VideoDecoder *>(decoder_.get())) b->set_feeder_may_block(true);
  feeder_thread_ = std::thread(&DecoderAutomata::feeder, this);
}

DecoderAutomata::~DecoderAutomata() {
  // reference: decoder_automata.cpp:55-78 (drain, flush, join)
  stop_feeder();
  {
    std::unique_lock<std::mutex> lk(mu_);
    quit_ = true;
  }
  cv_.notify_all();
  feeder_thread_.join();
  decoder_->flush();
  while (decoder_->decoded_frames_buffered() > 0) { if (!decoder_->discard_frame().ok) break; }
}

void DecoderAutomata::stop_feeder() {
  abort_ = true;
  std::unique_lock<std::mutex> lk(mu_);
  cv_.wait(lk, [&] { return parked_ && !work_; });
  abort_ = false;
}

Result DecoderAutomata::validate(const std::vector<EncodedData> &encoded_data) const {
  for (auto &d : encoded_data) {
    if (d.end_keyframe < d.start_keyframe) return Result(false, "EncodedData: end_keyframe < start_keyframe");
    const uint64_t n = d.end_keyframe - d.start_keyframe;
    if (d.sample_offsets.size() < n || d.sample_sizes.size() < n) return Result(false, "EncodedData: sample tables shorter than the interval");
    for (uint64_t i = 0; i < n; ++i)
      if (d.sample_offsets[i] + d.sample_sizes[i] > d.encoded_video.size()) return Result(false, "EncodedData: sample outside encoded_video");
    uint64_t prev = 0;
    bool first = true;
    for (uint64_t v : d.valid_frames) {
      if (v < d.start_keyframe || v >= d.end_keyframe) return Result(false, "EncodedData: valid frame outside the interval");
      if (!first && v <= prev) return Result(false, "EncodedData: valid_frames must be strictly ascending");
      prev = v; first = false;
    }
    if (d.width != encoded_data[0].width || d.height != encoded_data[0].height || d.format != encoded_data[0].format)
      return Result(false, "EncodedData: all intervals must share width/height/format");
  }
  return Result();
}

Result DecoderAutomata::initialize(const std::vector<EncodedData> &encoded_data, const std::vector<uint8_t> &extradata) {
  return initialize(std::vector<EncodedData>(encoded_data), extradata);
}

Result DecoderAutomata::initialize(std::vector<EncodedData> &&encoded_data, const std::vector<uint8_t> &extradata) {
  // reference: decoder_automata.cpp:80-118
  stop_feeder();
  decoder_->flush();
  while (decoder_->decoded_frames_buffered() > 0) HWANG_RETURN_ON_ERROR(decoder_->discard_frame());
  decoder_->wait_until_frames_copied();
  result_set_ = false;
  feeder_result_ = Result();
  encoded_data_.clear();
  interval_ = 0; popped_ = 0; valid_idx_ = 0;
  if (encoded_data.empty()) return Result();
  HWANG_RETURN_ON_ERROR(validate(encoded_data));
  encoded_data_ = std::move(encoded_data);
  info_.width = encoded_data_[0].width; info_.height = encoded_data_[0].height; info_.format = encoded_data_[0].format;
  frame_size_ = (size_t)info_.width * info_.height * 3;
  {
    Result r = decoder_->configure(info_, extradata);
    if (!r.ok) {  // nothing will be fed: a later get_frames must report this instead of waiting for frames
      encoded_data_.clear();
      feeder_result_ = r; result_set_ = true;
      return r;
    }
  }
  {
    std::unique_lock<std::mutex> lk(mu_);
    work_ = true;  // consumed by the feeder; no lost wake-up even if the thread has not reached its wait yet
  }
  cv_.notify_all();
  return Result();
}

std::vector<DecoderAutomata::EncodedData> DecoderAutomata::release_intervals() {
  stop_feeder();
  std::vector<EncodedData> old = std::move(encoded_data_);
  encoded_data_.clear();
  interval_ = 0; popped_ = 0; valid_idx_ = 0;
  return old;
}

// Samples of an interval that are actually fed.  The reference's feeder stops once the consumer has its frames
// (decoder_automata.cpp:287), so the tail of a GOP after the last wanted frame is not decoded; here the cut is a
// function of the request alone, which lets the feeder and the consumer agree on it without talking: everything up
// to the last wanted frame plus the largest possible reorder depth (16 frames: a picture's position in decode order
// exceeds its position in display order by at most the DPB size), so that every picture displayed up to the last
// wanted one has been fed, pops stay contiguous from the interval start, and the few extra ones are dropped.
uint64_t DecoderAutomata::fed_samples(const EncodedData &d) {
  const uint64_t n = d.end_keyframe - d.start_keyframe;
  if (d.valid_frames.empty()) return 0;
  const uint64_t need = d.valid_frames.back() - d.start_keyframe + 1 + 16;
  return need < n ? need : n;
}

void DecoderAutomata::feeder() {
  // reference: decoder_automata.cpp:259-404
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(mu_);
      parked_ = true;
      cv_.notify_all();
      cv_.wait(lk, [&] { return quit_ || work_; });
      if (quit_) return;
      work_ = false;
      parked_ = false;
    }
    bool failed = false;
    // Our own backend collects the pictures of ALL intervals of this request into as few GPU batches as its batch size
    // allows (one entropy launch sees the slices of many short intervals: sparse requests are latency-bound per slice),
    // instead of one small batch per interval; the last interval's flush is followed by submit_pending().
    B200VideoDecoder *batching = dynamic_cast<B200VideoDecoder *>(decoder_.get());
    if (batching) batching->set_defer_submit(encoded_data_.size() > 1);
    for (size_t di = 0; di < encoded_data_.size() && !abort_ && !failed; ++di) {
      const EncodedData &d = encoded_data_[di];
      const uint64_t n = fed_samples(d);
      size_t next_kf = 0;
      // our own backend is told which frames will be fetched, so that it can leave out unrequested non-reference
      // pictures; any other VideoDecoderInterface implementation just sees the 7 reference methods
      if (B200VideoDecoder *b = dynamic_cast<B200VideoDecoder *>(decoder_.get())) b->set_interval_hint(d.start_keyframe, d.valid_frames);
      for (uint64_t i = 0; i < n && !abort_; ++i) {
        while (!abort_ && decoder_->decoded_frames_buffered() > MAX_BUFFERED_FRAMES) std::this_thread::yield();
        if (abort_) break;
        const uint64_t frame = d.start_keyframe + i;
        bool is_keyframe = false;
        while (next_kf < d.keyframes.size() && d.keyframes[next_kf] < frame) next_kf++;
        if (next_kf < d.keyframes.size() && d.keyframes[next_kf] == frame) { is_keyframe = true; next_kf++; }
        Result r = decoder_->feed(d.encoded_video.data() + d.sample_offsets[i], (size_t)d.sample_sizes[i], is_keyframe);
        if (!r.ok) { feeder_result_ = r; result_set_ = true; failed = true; break; }
      }
      if (abort_ || failed) break;
      // end of interval: everything fed must become poppable (reference :383-397)
      Result r = decoder_->feed(nullptr, 0, false);
      if (r.ok) r = decoder_->flush();
      if (r.ok && batching && di + 1 == encoded_data_.size()) r = batching->submit_pending();
      if (!r.ok) { feeder_result_ = r; result_set_ = true; failed = true; }
    }
    if (batching) {
      if (abort_ || failed) batching->submit_pending();  // nothing may stay half-collected
      batching->set_defer_submit(false);
    }
  }
}

Result DecoderAutomata::get_frames(uint8_t *buffer, int32_t num_frames) {
  // reference: decoder_automata.cpp:120-252
  int32_t got = 0;
  while (got < num_frames) {
    if (result_set_) return feeder_result_;
    // advance past exhausted intervals (their unwanted tail frames are popped and dropped)
    if (interval_ >= encoded_data_.size()) return Result(false, "get_frames: requested more frames than the intervals contain");
    const EncodedData &d = encoded_data_[interval_];
    const uint64_t total = fed_samples(d);
    if (popped_ >= total) { interval_++; popped_ = 0; valid_idx_ = 0; continue; }
    if (valid_idx_ >= d.valid_frames.size()) {
      // nothing more wanted here: make sure later intervals still hold wanted frames before waiting on this tail
      bool more = false;
      for (size_t j = interval_ + 1; j < encoded_data_.size(); ++j) more |= !encoded_data_[j].valid_frames.empty();
      if (!more) return Result(false, "get_frames: requested more frames than the intervals contain");
    }
    // nothing to pop yet: back off instead of spinning (the reference yields in a tight loop, decoder_automata.cpp:242;
    // here the poll takes the decoder's lock, which the feeder thread needs for every sample it parses)
    if (decoder_->decoded_frames_buffered() <= 0) { std::this_thread::sleep_for(std::chrono::microseconds(50)); continue; }
    const uint64_t frame = d.start_keyframe + popped_;
    if (valid_idx_ < d.valid_frames.size() && d.valid_frames[valid_idx_] == frame) {
      HWANG_RETURN_ON_ERROR(decoder_->get_frame(buffer + (size_t)got * frame_size_, frame_size_));
      valid_idx_++; got++;
    } else {
      HWANG_RETURN_ON_ERROR(decoder_->discard_frame());
    }
    popped_++;
  }
  HWANG_RETURN_ON_ERROR(decoder_->wait_until_frames_copied());
  if (result_set_) return feeder_result_;
  return Result();
}

}  // namespace hwang
