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
#include <condition_variable>
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
  // Device time (CUDA events on the launching streams).  Chunks overlap (entropy decoding of later chunks runs under the
  // picture kernels of earlier ones), so the per-stage sums may exceed wall_ms: wall_ms is the device timeline from
  // "inputs of the interval's first chunk resident in HBM" to "last picture kernel of the interval done".
  double wall_ms = 0;
  double entropy_ms = 0, picture_ms = 0, aux_ms = 0;
  uint64_t entropy_launches = 0, picture_launches = 0, aux_launches = 0;
};

class B200VideoDecoder : public VideoDecoderInterface {
 public:
  // output_type (reference: SoftwareVideoDecoder's constructor, software_video_decoder.cpp:57-59; hwang/common.h:20-50):
  // where get_frame's destination buffers live.  DeviceType::GPU = caller-owned DEVICE memory (frames never cross
  // PCIe), DeviceType::CPU = host memory (page-locked buffers are written by the copy engine directly, pageable ones
  // through an internal page-locked ring).
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
  // drops them, decoder_automata.cpp:235) and converts only the wanted ones to RGB24.  Without a hint everything is
  // decoded and converted.
  void set_interval_hint(uint64_t start_frame, const std::vector<uint64_t> &wanted);
  Result discard_frame() override;
  Result get_frame(uint8_t *decoded_buffer, size_t decoded_size) override;
  int decoded_frames_buffered() override;
  Result wait_until_frames_copied() override;

  // ---- extensions used by the C-ABI (parity tests, benchmark, batch retrieval)
  // oldest frame as cropped planar I420 (width*height*3/2 bytes) instead of RGB24
  Result get_frame_yuv(uint8_t *decoded_buffer, size_t decoded_size);
  // true number of poppable frames (decoded_frames_buffered saturates, see .cpp)
  int frames_ready();
  // Zero-copy variant of get_frame for DeviceType::GPU consumers: the oldest frame's RGB24 in the decoder's own device
  // memory.  The pointer stays valid until the next wait_until_frames_copied() / configure() / destruction.
  Result get_frame_device(uint8_t **device_ptr);
  void set_chunk_pictures(int n) { chunk_target_ = n < 1 ? 1 : n; ramp_first_ = chunk_target_; }  // explicit size: no ramp
  // The caller feeds from a thread of its own (DecoderAutomata): feed() may then wait for the consumer to free device
  // memory instead of failing when the in-flight chunks reach the memory budget.
  void set_feeder_may_block(bool v) { feeder_may_block_ = v; }
  // Continue the current chunk across flush()/configure() boundaries of equal geometry (batch retrieval: slices of
  // many intervals and clips share one entropy launch).  submit_pending() closes the chunk.
  void set_defer_submit(bool v) { defer_submit_ = v; }
  Result submit_pending();
  DeviceType output_type() const { return output_type_; }
  B200Stats stats();
  hwb_dev *dev() { return dev_; }

 private:
  struct Slab { uint8_t *base = nullptr; size_t size = 0; };
  struct Chunk {
    std::vector<uint8_t> bitstream;  // slice RBSPs of the chunk being fed; handed back to spare_bits_ once uploaded
    std::vector<hwb::PicDesc> pics;
    std::vector<hwb::SliceDesc> slices;
    std::vector<int64_t> out_keys;
    std::vector<int> order;        // display order -> frame index
    std::vector<uint8_t> skipped;  // [frame] not decoded: unrequested non-reference picture (set_interval_hint)
    std::vector<int> seg_first;    // first picture of every segment (a segment = the pictures between two flushes: display order is per segment)
    std::vector<std::pair<uint64_t, std::vector<uint64_t>>> seg_hint;  // per segment: (start frame, wanted) or (~0, {}) = no hint
    size_t next_out = 0;
    Slab slab;
    hwb::ChunkCtx ctx;
    hwb_event *ev_begin = nullptr, *ev_entropy = nullptr, *ev_picture = nullptr, *ev_done = nullptr, *ev_copied = nullptr;
    bool submitted = false, finished = false, checked = false;
    bool lent = false;  // a device pointer into this chunk's RGB arena was handed out (get_frame_device)
    int32_t *error_dev = nullptr;
    volatile int32_t *done_host = nullptr;  // [pic] page-locked, written by the picture kernel when the picture is complete
    size_t done_capacity = 0;
    uint64_t alg_bytes = 0;
    int crop_x = 0, crop_y = 0;
    bool sparse = false;  // opened by a sparse request (see feed)
  };
  struct Staged { uint8_t *user = nullptr; size_t size = 0; hwb_event *done = nullptr; bool busy = false; };

  Result submit_current();
  Result finish_chunk(Chunk &c);  // wait for completion, check the device error flag
  Result wait_picture(Chunk &c, int frame);  // wait until one picture of a submitted chunk is complete
  bool picture_done(Chunk &c, int frame);
  Result pop_common(int mode, uint8_t *buf, size_t size, uint8_t **dev_out);
  Result stage_slot(int *slot);
  void retire_front();
  void poll_retired();
  void drain_copies();
  void recycle(std::unique_ptr<Chunk> &c, bool keep_slab);
  Slab take_slab(size_t n);
  void release_all();
  void reset_keep_memory();
  size_t picture_bytes() const;
  void close_segment();

  hwb_dev *dev_ = nullptr;
  int device_id_;
  DeviceType output_type_;
  std::mutex mu_;
  std::condition_variable memory_cv_;
  hwb::H264Stream stream_;
  bool configured_ = false;
  uint32_t width_ = 0, height_ = 0;
  // Pictures per GPU batch (cut at IDR pictures, also bounded by bytes: see feed()).  The entropy stage is
  // latency-bound per slice, so a batch must hold many slices; several batches are in flight at once (entropy streams
  // rotate), which overlaps host parsing, entropy decoding of later batches, the picture kernel of earlier ones and the
  // copies to the host.
  // Batches ramp up: the first batch of a cold pipeline is small (ramp_first_ pictures) so that the first frames leave
  // early -- every batch costs at least one intra slice's entropy latency before its pictures can be reconstructed --
  // and each following batch doubles up to chunk_target_ (large batches are the efficient ones).  Measured on the
  // 3000-frame benchmark clip: first frames on the host after 460 ms with equal batches, with the copy engine the
  // bottleneck from then on.
  int chunk_target_ = 960;
  int ramp_first_ = 120, ramp_target_ = 120;
  int deblock_band_ = 0;  // 0 = by batch kind (see submit_current)
  int group_target_ = 1 << 30;  // pictures per GOP group inside a batch (work order of the picture kernel); default: one group
  bool feeder_may_block_ = false, defer_submit_ = false, no_rgb_ = false, picture_profile_ = false, concurrent_ = false, intra_reserve_ = true;
  std::unique_ptr<Chunk> cur_;
  std::deque<std::unique_ptr<Chunk>> queue_;    // submitted chunks, oldest first
  std::vector<std::unique_ptr<Chunk>> retired_;  // fully popped, slab reusable after the next copy-stream sync
  std::vector<Slab> free_slabs_;
  std::vector<std::pair<int32_t *, size_t>> free_flags_;  // page-locked completion-flag arrays of recycled chunks
  size_t live_bytes_ = 0;     // device memory held (slabs in use + cached)
  size_t memory_budget_ = 0;  // in-flight limit, a fraction of what was free at configure()
  size_t last_chunk_bytes_ = 0;
  int next_entropy_stream_ = 0;
  std::vector<uint8_t> spare_bits_;
  bool hint_valid_ = false;
  uint64_t hint_start_ = 0, hint_span_ = 0;
  std::vector<uint64_t> hint_wanted_;
  hwb_event *interval_begin_ = nullptr;  // ev_begin of the first chunk since the last wall-clock reading
  // output staging for pageable destinations and the planar test output
  static const int kRing = 8;
  uint8_t *stage_dev_[kRing] = {nullptr};
  uint8_t *stage_pinned_[kRing] = {nullptr};
  Staged staged_[kRing];
  size_t ring_bytes_ = 0;
  int ring_next_ = 0;
  bool profile_ = true;
  B200Stats stats_;
  std::string sticky_error_;
};

}  // namespace hwang
