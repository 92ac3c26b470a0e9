// The B200 backend behind hwang's VideoDecoderInterface: a from-scratch H.264 decoder whose
// entropy decoding, reconstruction, deblocking and RGB conversion run as CUDA kernels (no
// libavcodec, no NVDEC, no CPU fallback).  Plays the role SoftwareVideoDecoder / NVIDIAVideoDecoder
// play in the reference (hwang/impls/software/software_video_decoder.cpp:103-457,
// hwang/impls/nvidia/nvidia_video_decoder.cpp:412-464).
//
// Threading contract (reference: decoder_automata.cpp:259-404 feeder thread, :120-252 consumer):
// feed/flush may run on one thread while get_frame/discard_frame/wait_until_frames_copied run on
// another; decoded_frames_buffered is polled from both.
#pragma once
#include <deque>
#include <memory>
#include <mutex>
#include <vector>

#include "../dev/devapi.h"
#include "h264_stream.h"
#include "video_decoder_interface.h"

namespace hwang {

struct B200Stats {
  uint64_t pictures_decoded = 0, frames_returned = 0, chunks = 0, bitstream_bytes = 0;
  uint64_t kernel_launches = 0;
  uint64_t h2d_bytes = 0, d2h_bytes = 0;
  uint64_t algorithmic_bytes = 0;  // recon write + one reference read + RGB write
  double decode_ms = 0;            // device time of the decode stages (CUDA events), summed over chunks
  double entropy_ms = 0, recon_ms = 0, deblock_ms = 0, rgb_ms = 0;
  uint64_t entropy_launches = 0, recon_launches = 0, deblock_launches = 0, rgb_launches = 0;
};

class B200VideoDecoder : public VideoDecoderInterface {
 public:
  B200VideoDecoder(int device_id, DeviceType output_type, int num_devices);
  ~B200VideoDecoder() override;
  bool ok() const { return dev_ != nullptr; }

  Result configure(const FrameInfo &metadata, const std::vector<uint8_t> &extradata) override;
  Result feed(const uint8_t *encoded_buffer, size_t encoded_size, bool keyframe) override;
  Result flush() override;
  // Optional hint from the automaton (not part of the reference's 7-method interface): the pictures fed from now
  // until the next flush() are absolute frames `start_frame`, `start_frame`+1, ... in display order, and only the
  // frames in `wanted` (ascending) will be fetched with get_frame; the rest will be dropped with discard_frame.
  // The decoder then skips unrequested NON-REFERENCE pictures altogether (the reference decodes and
  // drops them, decoder_automata.cpp:235).  Without a hint everything is decoded.
  void set_interval_hint(uint64_t start_frame, const std::vector<uint64_t> &wanted);
  Result discard_frame() override;
  Result get_frame(uint8_t *decoded_buffer, size_t decoded_size) override;
  int decoded_frames_buffered() override;
  Result wait_until_frames_copied() override;

  // ---- extensions used by the C-ABI (parity tests, benchmark)
  // oldest frame as cropped planar I420 (width*height*3/2 bytes) instead of RGB24
  Result get_frame_yuv(uint8_t *decoded_buffer, size_t decoded_size);
  // true number of poppable frames (decoded_frames_buffered saturates, see .cpp)
  int frames_ready();
  // convert the oldest frame to RGB24 into device memory only (no D2H); for kernel-only timing
  Result get_frame_device(uint8_t **device_ptr);
  void set_chunk_pictures(int n) { chunk_target_ = n < 1 ? 1 : n; }
  B200Stats stats();
  hwb_dev *dev() { return dev_; }

 private:
  struct Slab { uint8_t *base = nullptr; size_t size = 0; };
  struct Chunk {
    std::vector<uint8_t> bitstream;  // slice RBSPs of the chunk being fed; handed back to spare_bits_ once uploaded
    std::vector<hwb::PicDesc> pics;
    std::vector<hwb::SliceDesc> slices;
    std::vector<int64_t> out_keys;
    std::vector<int> order;  // display order -> frame index
    std::vector<uint8_t> skipped;  // [frame] not decoded: unrequested non-reference picture (set_interval_hint)
    size_t next_out = 0;
    Slab slab;
    hwb::ChunkCtx ctx;
    hwb_event *ev_begin = nullptr, *ev_done = nullptr;
    std::vector<hwb_event *> stage_ev;  // boundaries: [0] before entropy, [1] after entropy, then after each recon / deblock launch
    bool submitted = false, finished = false, checked = false;
    int32_t *error_dev = nullptr;
    uint64_t alg_bytes = 0;
  };
  struct PendingCopy { uint8_t *user; uint8_t *pinned; size_t size; };

  Result submit_current();
  Result finish_chunk(Chunk &c);  // wait for completion, check the device error flag
  Result pop_common(int mode, uint8_t *buf, size_t size, uint8_t **dev_out);
  void retire_front();
  void drain_copies();
  Slab take_slab(size_t n);
  void release_all();
  void reset_keep_memory();

  hwb_dev *dev_ = nullptr;
  int device_id_;
  std::mutex mu_;
  hwb::H264Stream stream_;
  bool configured_ = false;
  uint32_t width_ = 0, height_ = 0;
  // pictures per GPU batch (cut at IDR pictures).  The entropy stage is latency-bound per slice, so batches must be
  // large; several batches are in flight at once on different streams, which overlaps host parsing, the entropy
  // stage of the next batch, reconstruction of the previous one and the copies to the host.
  int chunk_target_ = 4096;  // pictures per chunk (cut at IDR pictures); measured: splitting a 3000-picture clip only loses (every stage is bound by instruction fetch, overlapped chunks share that budget)
  std::unique_ptr<Chunk> cur_;
  std::deque<std::unique_ptr<Chunk>> queue_;    // submitted chunks, oldest first
  std::vector<std::unique_ptr<Chunk>> retired_;  // fully popped, slab reusable after the next copy-stream sync
  std::vector<Slab> free_slabs_;
  size_t live_bytes_ = 0;
  size_t last_chunk_bytes_ = 0;
  // output staging
  std::vector<uint8_t> spare_bits_;
  bool hint_valid_ = false;
  uint64_t hint_start_ = 0, interval_submitted_ = 0;
  std::vector<uint64_t> hint_wanted_;
  static const int kRing = 8;
  uint8_t *rgb_dev_[kRing] = {nullptr};
  uint8_t *rgb_pinned_[kRing] = {nullptr};
  size_t ring_bytes_ = 0;
  int ring_next_ = 0;
  std::vector<PendingCopy> pending_;
  std::vector<std::pair<hwb_event *, hwb_event *>> rgb_ev_;  // per pending RGB launch
  bool profile_ = true;
  B200Stats stats_;
  std::string sticky_error_;
};

}  // namespace hwang
