#include "b200_video_decoder.h"

#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <thread>

#include "../dev/entropy.h"  // sizeof(NbCtx)

namespace hwang {

using hwb::ChunkCtx;
using hwb::PicDesc;
using hwb::SliceDesc;

namespace {
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
const size_t kLiveBudget = (size_t)96 << 30;  // beyond this much device memory in flight, expose real back-pressure
}  // namespace

B200VideoDecoder::B200VideoDecoder(int device_id, DeviceType, int) : device_id_(device_id) {
  if (hwb_dev_open(device_id, &dev_) != 0) dev_ = nullptr;
  if (const char *e = getenv("HWB_CHUNK_PICTURES")) { int v = atoi(e); if (v > 0) chunk_target_ = v; }
}

B200VideoDecoder::~B200VideoDecoder() {
  if (!dev_) return;
  release_all();
  hwb_dev_close(dev_);
}

void B200VideoDecoder::release_all() {
  for (int i = 0; i < HWB_NUM_STREAMS; ++i) hwb_dev_stream_sync(dev_, i);
  auto drop = [&](std::unique_ptr<Chunk> &c) {
    if (!c) return;
    for (auto e : c->stage_ev) hwb_dev_event_destroy(dev_, e);
    if (c->ev_begin) hwb_dev_event_destroy(dev_, c->ev_begin);
    if (c->ev_done) hwb_dev_event_destroy(dev_, c->ev_done);
    if (c->slab.base) hwb_dev_free(dev_, c->slab.base);
    c.reset();
  };
  for (auto &c : queue_) drop(c);
  queue_.clear();
  for (auto &c : retired_) drop(c);
  retired_.clear();
  for (auto &s : free_slabs_) hwb_dev_free(dev_, s.base);
  free_slabs_.clear();
  cur_.reset();
  for (int i = 0; i < kRing; ++i) {
    if (rgb_dev_[i]) hwb_dev_free(dev_, rgb_dev_[i]);
    if (rgb_pinned_[i]) hwb_dev_free_host(dev_, rgb_pinned_[i]);
    rgb_dev_[i] = rgb_pinned_[i] = nullptr;
  }
  pending_.clear();
  for (auto &e : rgb_ev_) { hwb_dev_event_destroy(dev_, e.first); hwb_dev_event_destroy(dev_, e.second); }
  rgb_ev_.clear();
  live_bytes_ = 0;
  ring_bytes_ = 0;
}

// Drop every queued / in-flight chunk but keep slabs and staging rings for reuse.
void B200VideoDecoder::reset_keep_memory() {
  for (int i = 0; i < HWB_NUM_STREAMS; ++i) hwb_dev_stream_sync(dev_, i);
  auto recycle = [&](std::unique_ptr<Chunk> &c) {
    if (!c) return;
    for (auto e : c->stage_ev) hwb_dev_event_destroy(dev_, e);
    if (c->ev_begin) hwb_dev_event_destroy(dev_, c->ev_begin);
    if (c->ev_done) hwb_dev_event_destroy(dev_, c->ev_done);
    if (c->slab.base) free_slabs_.push_back(c->slab);
    c.reset();
  };
  for (auto &c : queue_) recycle(c);
  queue_.clear();
  for (auto &c : retired_) recycle(c);
  retired_.clear();
  cur_.reset();
  pending_.clear();
  for (auto &e : rgb_ev_) { hwb_dev_event_destroy(dev_, e.first); hwb_dev_event_destroy(dev_, e.second); }
  rgb_ev_.clear();
  ring_next_ = 0;
  // keep the two largest spare slabs
  std::sort(free_slabs_.begin(), free_slabs_.end(), [](const Slab &a, const Slab &b) { return a.size > b.size; });
  while (free_slabs_.size() > 2) { live_bytes_ -= free_slabs_.back().size; hwb_dev_free(dev_, free_slabs_.back().base); free_slabs_.pop_back(); }
}

// reference: SoftwareVideoDecoder::configure, software_video_decoder.cpp:103-165
Result B200VideoDecoder::configure(const FrameInfo &metadata, const std::vector<uint8_t> &extradata) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!dev_) return Result(false, "B200 decoder: CUDA device unavailable");
  if (!(metadata.format == "h264" || metadata.format == "avc1"))
    return Result(false, "Unsupported video codec: " + metadata.format + " (supports h264 only)");
  // hwang re-configures per interval (python/hwang/decoder.py:65): keep the device memory when the geometry is unchanged
  const bool same_geometry = configured_ && width_ == metadata.width && height_ == metadata.height;
  if (same_geometry) reset_keep_memory(); else release_all();
  sticky_error_.clear();
  hint_valid_ = false; interval_submitted_ = 0;
  std::string err = stream_.configure(extradata.data(), extradata.size());
  if (!err.empty()) { configured_ = false; return Result(false, "B200 decoder: " + err); }
  if ((uint32_t)stream_.width() != metadata.width || (uint32_t)stream_.height() != metadata.height)
    return Result(false, "B200 decoder: container size " + std::to_string(metadata.width) + "x" + std::to_string(metadata.height) +
                             " does not match the SPS (" + std::to_string(stream_.width()) + "x" + std::to_string(stream_.height()) + ")");
  width_ = metadata.width; height_ = metadata.height;
  ring_bytes_ = (size_t)width_ * height_ * 3;
  for (int i = 0; i < kRing && !rgb_dev_[i]; ++i) {
    rgb_dev_[i] = (uint8_t *)hwb_dev_malloc(dev_, ring_bytes_);
    rgb_pinned_[i] = (uint8_t *)hwb_dev_malloc_host(dev_, ring_bytes_);
    if (!rgb_dev_[i] || !rgb_pinned_[i]) return Result(false, std::string("B200 decoder: out of memory: ") + hwb_dev_error(dev_));
  }
  configured_ = true;
  return Result();
}

void B200VideoDecoder::set_interval_hint(uint64_t start_frame, const std::vector<uint64_t> &wanted) {
  std::lock_guard<std::mutex> lk(mu_);
  hint_valid_ = true; hint_start_ = start_frame; hint_wanted_ = wanted; interval_submitted_ = 0;
}

B200VideoDecoder::Slab B200VideoDecoder::take_slab(size_t n) {
  for (size_t i = 0; i < free_slabs_.size(); ++i)
    if (free_slabs_[i].size >= n) { Slab s = free_slabs_[i]; free_slabs_.erase(free_slabs_.begin() + i); return s; }
  // nothing fits: drop the cached slabs and allocate
  for (auto &s : free_slabs_) { live_bytes_ -= s.size; hwb_dev_free(dev_, s.base); }
  free_slabs_.clear();
  Slab s;
  s.base = (uint8_t *)hwb_dev_malloc(dev_, n);
  s.size = s.base ? n : 0;
  if (s.base) live_bytes_ += n;
  return s;
}

// reference: SoftwareVideoDecoder::feed, software_video_decoder.cpp:167-248
Result B200VideoDecoder::feed(const uint8_t *encoded_buffer, size_t encoded_size, bool keyframe) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!configured_) return Result(false, "B200 decoder: feed before configure");
  if (!sticky_error_.empty()) return Result(false, sticky_error_);
  if (encoded_size == 0 || encoded_buffer == nullptr) return submit_current();  // end of interval: everything fed becomes poppable
  (void)keyframe;
  const bool idr = stream_.next_is_idr(encoded_buffer, encoded_size);
  if (cur_ && idr && (int)cur_->pics.size() >= chunk_target_) HWANG_RETURN_ON_ERROR(submit_current());
  if (!cur_) {
    if (!idr) return Result(false, "B200 decoder: interval does not start with an IDR picture");
    cur_.reset(new Chunk());
    // the previous chunk's staging buffer is recycled: fresh memory would cost one page fault per 4 KB of bitstream
    cur_->bitstream = std::move(spare_bits_);
    cur_->bitstream.clear();
    cur_->bitstream.reserve(last_chunk_bytes_ + (last_chunk_bytes_ >> 2) + (1 << 20));
    stream_.reset_dpb();
  }
  hwb::PlannedPic pp;
  std::string err = stream_.parse_sample(encoded_buffer, encoded_size, (int)cur_->pics.size(), cur_->bitstream, pp);
  if (!err.empty()) { sticky_error_ = "B200 decoder: " + err; return Result(false, sticky_error_); }
  pp.desc.first_slice = (int)cur_->slices.size();
  for (auto &s : pp.slices) cur_->slices.push_back(s);
  cur_->pics.push_back(pp.desc);
  cur_->out_keys.push_back(pp.out_key);
  stats_.bitstream_bytes += encoded_size;
  return Result();
}

Result B200VideoDecoder::submit_current() {
  if (!cur_ || cur_->pics.empty()) { cur_.reset(); return Result(); }
  std::unique_ptr<Chunk> ch = std::move(cur_);
  const int P = (int)ch->pics.size(), S = (int)ch->slices.size();
  const int mb_w = stream_.mb_w(), mb_h = stream_.mb_h(), nmb = mb_w * mb_h;
  // display order
  ch->order.resize(P);
  for (int i = 0; i < P; ++i) ch->order[i] = i;
  std::stable_sort(ch->order.begin(), ch->order.end(), [&](int a, int b) { return ch->out_keys[a] < ch->out_keys[b]; });
  // unrequested non-reference pictures are not decoded at all (their frame buffers stay undefined: nothing
  // references them and the consumer drops them)
  ch->skipped.assign(P, 0);
  int nskipped = 0;
  if (hint_valid_) {
    for (int j = 0; j < P; ++j) {
      const int pic = ch->order[j];
      const uint64_t frame = hint_start_ + interval_submitted_ + (uint64_t)j;
      if (!ch->pics[pic].is_ref && !std::binary_search(hint_wanted_.begin(), hint_wanted_.end(), frame)) { ch->skipped[pic] = 1; nskipped++; }
    }
  }
  interval_submitted_ += (uint64_t)P;
  // levels
  int nlevels = 0;
  for (auto &p : ch->pics) nlevels = std::max(nlevels, p.level + 1);
  std::vector<std::vector<int32_t>> by_level(nlevels);
  for (int i = 0; i < P; ++i) if (!ch->skipped[i]) by_level[ch->pics[i].level].push_back(i);
  std::vector<int32_t> level_list;
  for (auto &v : by_level) level_list.insert(level_list.end(), v.begin(), v.end());

  // device layout
  const size_t fs = (size_t)mb_w * 16 * mb_h * 16 * 3 / 2;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 256); return o; };
  const size_t o_frames = take(fs * P), o_mbinfo = take((size_t)P * nmb * sizeof(hwb::MbInfo)), o_mv = take((size_t)P * 2 * nmb * 64),
               o_refidx = take((size_t)P * 2 * nmb * 4), o_refpic = take((size_t)P * 2 * nmb * 8),
               o_coefs = take((size_t)P * nmb * hwb::SLOTS_PER_MB * 32), o_ectx = take((size_t)S * mb_w * sizeof(hwb::NbCtx)),
               o_bits = take(ch->bitstream.size() + 64), o_pics = take((size_t)P * sizeof(PicDesc)), o_slices = take((size_t)S * sizeof(SliceDesc)),
               o_levels = take((size_t)P * 4), o_order = take((size_t)S * 4);
  const size_t n_sync = (size_t)(2 * nlevels + 1) + S + 2 * (size_t)P * mb_h + 4;
  const size_t o_sync = take(n_sync * 4);
  ch->slab = take_slab(off);
  if (!ch->slab.base) return Result(false, std::string("B200 decoder: device allocation of ") + std::to_string(off) + " bytes failed: " + hwb_dev_error(dev_));
  uint8_t *b = ch->slab.base;
  ChunkCtx &c = ch->ctx;
  memset(&c, 0, sizeof(c));
  c.mb_w = mb_w; c.mb_h = mb_h; c.nmb = nmb; c.wc = mb_w * 16; c.hc = mb_h * 16; c.num_pics = P; c.num_slices = S;
  c.frames = b + o_frames; c.frame_stride = fs;
  c.mbinfo = (hwb::MbInfo *)(b + o_mbinfo); c.mv = (int16_t *)(b + o_mv); c.refidx = (int8_t *)(b + o_refidx); c.refpic = (int16_t *)(b + o_refpic);
  c.coefs = (int16_t *)(b + o_coefs); c.ectx = b + o_ectx; c.ectx_stride = (uint64_t)mb_w * sizeof(hwb::NbCtx);
  c.bitstream = b + o_bits; c.pics = (const PicDesc *)(b + o_pics); c.slices = (const SliceDesc *)(b + o_slices);
  int32_t *sync = (int32_t *)(b + o_sync);
  int32_t *tickets = sync;
  c.entropy_prog = sync + (2 * nlevels + 1);
  c.recon_prog = c.entropy_prog + S;
  c.dbl_prog = c.recon_prog + (size_t)P * mb_h;
  c.error_flag = c.dbl_prog + (size_t)P * mb_h;
  ch->error_dev = c.error_flag;

  // Two decode streams shared by all chunks: inputs + entropy decoding on the first, reconstruction + deblocking on
  // the second.  Entropy decoding of chunk k+1 (bound by instruction fetch and by the latency of the intra slices)
  // then overlaps the level-by-level reconstruction of chunk k (bound by the wavefront latency of each level), and
  // chunk k's frames travel to the host while chunk k+1 is reconstructed.
  const int st = HWB_STREAM_DECODE, st_recon = HWB_STREAM_DECODE + 1;
  ch->ev_begin = hwb_dev_event_create(dev_);
  ch->ev_done = hwb_dev_event_create(dev_);
  int rc = 0;
  last_chunk_bytes_ = ch->bitstream.size();
  ch->bitstream.resize(ch->bitstream.size() + 64, 0);  // read-ahead padding for the bit readers
  rc |= hwb_dev_h2d(dev_, st, b + o_bits, ch->bitstream.data(), ch->bitstream.size());
  rc |= hwb_dev_h2d(dev_, st, b + o_pics, ch->pics.data(), (size_t)P * sizeof(PicDesc));
  rc |= hwb_dev_h2d(dev_, st, b + o_slices, ch->slices.data(), (size_t)S * sizeof(SliceDesc));
  if (!level_list.empty()) rc |= hwb_dev_h2d(dev_, st, b + o_levels, level_list.data(), level_list.size() * 4);
  // Entropy tickets: intra slices carry several times the bits of inter slices and wait on nothing, so they start
  // first; everything else keeps decode order (a B slice's co-located picture then always holds an earlier ticket).
  std::vector<int32_t> order;
  order.reserve(S);
  for (int i = 0; i < S; ++i) if (ch->slices[i].slice_type == hwb::SLICE_I && !ch->skipped[ch->slices[i].pic]) order.push_back(i);
  for (int i = 0; i < S; ++i) if (ch->slices[i].slice_type != hwb::SLICE_I && !ch->skipped[ch->slices[i].pic]) order.push_back(i);
  if (!order.empty()) rc |= hwb_dev_h2d(dev_, st, b + o_order, order.data(), order.size() * 4);
  c.entropy_order = (const int32_t *)(b + o_order);
  c.num_tickets = (int32_t)order.size();
  rc |= hwb_dev_memset(dev_, st, sync, 0, n_sync * 4);
  stats_.h2d_bytes += ch->bitstream.size() + (size_t)P * sizeof(PicDesc) + (size_t)S * sizeof(SliceDesc) + (size_t)P * 4;
  // a copy from pageable memory has been staged by the time cudaMemcpyAsync returns: the buffer can be reused
  spare_bits_ = std::move(ch->bitstream);
  rc |= hwb_dev_event_record(dev_, ch->ev_begin, st);  // inputs are resident in HBM from here on
  auto mark = [&]() { if (profile_) { hwb_event *e = hwb_dev_event_create(dev_); hwb_dev_event_record(dev_, e, st); ch->stage_ev.push_back(e); } };
  mark();
  int mode = ch->pics[0].cabac ? 1 : 0;
  for (auto &p : ch->pics) if ((p.cabac ? 1 : 0) != mode) mode = -1;
  if (mode == 1) {  // CABAC throughout: the copy of the kernel without B-slice support when the chunk has none
    bool has_b = false;
    for (auto &sl : ch->slices) has_b |= sl.slice_type == hwb::SLICE_B;
    if (!has_b) mode = 3;
  }
  rc |= hwb_dev_entropy(dev_, st, &c, tickets, mode);
  mark();
  {
    hwb_event *ev_entropy = hwb_dev_event_create(dev_);
    rc |= hwb_dev_event_record(dev_, ev_entropy, st);
    rc |= hwb_dev_stream_wait(dev_, st_recon, ev_entropy);
    hwb_dev_event_destroy(dev_, ev_entropy);  // the wait already enqueued keeps its own reference
  }
  auto mark_recon = [&]() { if (profile_) { hwb_event *e = hwb_dev_event_create(dev_); hwb_dev_event_record(dev_, e, st_recon); ch->stage_ev.push_back(e); } };
  mark_recon();  // stage_ev[2]: reconstruction stream free and entropy done
  size_t lo = 0;
  for (int l = 0; l < nlevels; ++l) {
    const int n = (int)by_level[l].size();
    const int32_t *pl = (const int32_t *)(b + o_levels) + lo;
    if (n > 0) rc |= hwb_dev_recon(dev_, st_recon, &c, pl, n, tickets + 1 + 2 * l);
    mark_recon();
    if (n > 0) rc |= hwb_dev_deblock(dev_, st_recon, &c, pl, n, tickets + 2 + 2 * l);
    mark_recon();
    lo += n;
  }
  rc |= hwb_dev_event_record(dev_, ch->ev_done, st_recon);
  if (rc) { sticky_error_ = std::string("B200 decoder: CUDA launch failed: ") + hwb_dev_error(dev_); return Result(false, sticky_error_); }
  for (int i = 0; i < P; ++i) if (!ch->skipped[i]) ch->alg_bytes += fs + (ch->pics[i].has_inter ? fs : 0);
  ch->submitted = true;
  stats_.chunks++;
  stats_.pictures_decoded += P - nskipped;
  queue_.push_back(std::move(ch));
  return Result();
}

// reference: SoftwareVideoDecoder::flush, software_video_decoder.cpp:250-268 (drain + reset for the next interval)
Result B200VideoDecoder::flush() {
  std::lock_guard<std::mutex> lk(mu_);
  if (!configured_) return Result();
  Result r = submit_current();
  stream_.reset_dpb();
  hint_valid_ = false; interval_submitted_ = 0;  // the hint covers one interval
  return r;
}

Result B200VideoDecoder::finish_chunk(Chunk &c) {
  if (c.checked) return Result();
  if (hwb_dev_event_sync(dev_, c.ev_done) != 0) return Result(false, std::string("B200 decoder: CUDA error: ") + hwb_dev_error(dev_));
  int32_t flag = 0;
  hwb_dev_d2h(dev_, HWB_STREAM_COPY, &flag, c.error_dev, 4);
  hwb_dev_stream_sync(dev_, HWB_STREAM_COPY);
  float ms = 0;
  if (hwb_dev_event_elapsed(dev_, c.ev_begin, c.ev_done, &ms) == 0) stats_.decode_ms += ms;
  for (size_t i = 1; i < c.stage_ev.size(); ++i) {
    float t = 0;
    if (hwb_dev_event_elapsed(dev_, c.stage_ev[i - 1], c.stage_ev[i], &t) != 0) continue;
    if (i == 1) { stats_.entropy_ms += t; stats_.entropy_launches++; }
    else if (i == 2) continue;  // waiting for the reconstruction stream (another chunk's levels)
    else if (i % 2 == 1) { stats_.recon_ms += t; stats_.recon_launches++; }
    else { stats_.deblock_ms += t; stats_.deblock_launches++; }
  }
  for (auto e : c.stage_ev) hwb_dev_event_destroy(dev_, e);
  c.stage_ev.clear();
  stats_.algorithmic_bytes += c.alg_bytes;
  c.finished = true; c.checked = true;
  if (flag) { sticky_error_ = "B200 decoder: corrupt or unsupported bitstream (device error code " + std::to_string(flag) + ")"; return Result(false, sticky_error_); }
  return Result();
}

int B200VideoDecoder::frames_ready() {
  std::lock_guard<std::mutex> lk(mu_);
  if (!sticky_error_.empty()) return -1;
  int n = 0;
  for (auto &c : queue_) {
    if (!c->finished) {
      int st = hwb_dev_event_done(dev_, c->ev_done);
      if (st < 0) { sticky_error_ = std::string("B200 decoder: CUDA error: ") + hwb_dev_error(dev_); return -1; }
      if (st == 1) c->finished = true;
      else break;  // chunks complete in order
    }
    n += (int)(c->order.size() - c->next_out);
  }
  return n;
}

// reference: SoftwareVideoDecoder::decoded_frames_buffered, software_video_decoder.cpp:341-343.
// The automaton's feeder pauses while this exceeds 8 (decoder_automata.cpp:288-293).  A batch decoder
// wants the whole interval fed while earlier chunks are still being consumed, so the value saturates
// at 8 ("at least this many frames are ready") unless a lot of device memory is already in flight.
int B200VideoDecoder::decoded_frames_buffered() {
  int n = frames_ready();
  if (n < 0) return 1;  // error state: let the consumer pop, get_frame / discard_frame report the error
  std::lock_guard<std::mutex> lk(mu_);
  if (live_bytes_ > kLiveBudget) return n;
  return n > 8 ? 8 : n;
}

void B200VideoDecoder::retire_front() {
  retired_.push_back(std::move(queue_.front()));
  queue_.pop_front();
}

void B200VideoDecoder::drain_copies() {
  hwb_dev_stream_sync(dev_, HWB_STREAM_COPY);
  for (auto &e : rgb_ev_) {
    float t = 0;
    if (hwb_dev_event_elapsed(dev_, e.first, e.second, &t) == 0) { stats_.rgb_ms += t; stats_.rgb_launches++; }
    hwb_dev_event_destroy(dev_, e.first); hwb_dev_event_destroy(dev_, e.second);
  }
  rgb_ev_.clear();
  for (auto &p : pending_) if (p.pinned) memcpy(p.user, p.pinned, p.size);
  pending_.clear();
  ring_next_ = 0;
  // every read of retired chunks has completed: their memory can be reused
  for (auto &c : retired_) {
    hwb_dev_event_destroy(dev_, c->ev_begin); hwb_dev_event_destroy(dev_, c->ev_done);
    free_slabs_.push_back(c->slab);
  }
  retired_.clear();
  // keep at most two spare slabs
  while (free_slabs_.size() > 2) { live_bytes_ -= free_slabs_.back().size; hwb_dev_free(dev_, free_slabs_.back().base); free_slabs_.pop_back(); }
}

// mode 0: RGB24 to host, 1: planar I420 to host, 2: RGB24 left in device memory, 3: discard
Result B200VideoDecoder::pop_common(int mode, uint8_t *buf, size_t size, uint8_t **dev_out) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!sticky_error_.empty()) return Result(false, sticky_error_);
  while (!queue_.empty() && queue_.front()->next_out >= queue_.front()->order.size()) retire_front();
  if (queue_.empty()) return Result(false, "B200 decoder: no decoded frame buffered");
  Chunk &c = *queue_.front();
  HWANG_RETURN_ON_ERROR(finish_chunk(c));
  const int frame = c.order[c.next_out++];
  if (mode != 3 && c.skipped[frame]) return Result(false, "B200 decoder: this frame was declared unwanted (set_interval_hint) and has not been decoded");
  if (mode == 3) {
    if (c.next_out >= c.order.size()) retire_front();
    return Result();
  }
  const size_t need = mode == 1 ? (size_t)width_ * height_ * 3 / 2 : (size_t)width_ * height_ * 3;
  if (mode != 2 && size < need) return Result(false, "B200 decoder: output buffer too small");
  if ((int)pending_.size() >= kRing || ring_next_ >= kRing) drain_copies();
  const int slot = ring_next_++;
  int rc;
  hwb_event *r0 = nullptr, *r1 = nullptr;
  // colour conversion on its own stream, the copy stream waits for it: the conversion of frame k+1 overlaps the
  // device-to-host copy of frame k (slots of the ring are recycled only after drain_copies)
  const int st_rgb = HWB_STREAM_DECODE + 2;
  if (profile_ && mode != 1) { r0 = hwb_dev_event_create(dev_); r1 = hwb_dev_event_create(dev_); hwb_dev_event_record(dev_, r0, st_rgb); }
  if (mode == 1) rc = hwb_dev_yuv(dev_, st_rgb, &c.ctx, frame, stream_.crop_left(), stream_.crop_top(), (int)width_, (int)height_, rgb_dev_[slot]);
  else rc = hwb_dev_rgb24(dev_, st_rgb, &c.ctx, frame, stream_.crop_left(), stream_.crop_top(), (int)width_, (int)height_, rgb_dev_[slot]);
  if (r1) { hwb_dev_event_record(dev_, r1, st_rgb); rgb_ev_.push_back({r0, r1}); }
  {
    hwb_event *conv = hwb_dev_event_create(dev_);
    rc |= hwb_dev_event_record(dev_, conv, st_rgb);
    rc |= hwb_dev_stream_wait(dev_, HWB_STREAM_COPY, conv);
    hwb_dev_event_destroy(dev_, conv);
  }
  if (mode == 2) {
    *dev_out = rgb_dev_[slot];
    pending_.push_back({nullptr, nullptr, 0});
  } else {
    const bool direct = hwb_dev_is_pinned(dev_, buf) == 1;
    uint8_t *dst = direct ? buf : rgb_pinned_[slot];
    rc |= hwb_dev_d2h(dev_, HWB_STREAM_COPY, dst, rgb_dev_[slot], need);
    pending_.push_back({buf, direct ? nullptr : rgb_pinned_[slot], need});
    stats_.d2h_bytes += need;
  }
  if (rc) { sticky_error_ = std::string("B200 decoder: CUDA error: ") + hwb_dev_error(dev_); return Result(false, sticky_error_); }
  if (mode != 1) stats_.algorithmic_bytes += (uint64_t)width_ * height_ * 3;
  stats_.frames_returned++;
  if (c.next_out >= c.order.size()) retire_front();
  return Result();
}

// reference: software_video_decoder.cpp:270-279
Result B200VideoDecoder::discard_frame() { return pop_common(3, nullptr, 0, nullptr); }
// reference: software_video_decoder.cpp:281-339
Result B200VideoDecoder::get_frame(uint8_t *decoded_buffer, size_t decoded_size) { return pop_common(0, decoded_buffer, decoded_size, nullptr); }
Result B200VideoDecoder::get_frame_yuv(uint8_t *decoded_buffer, size_t decoded_size) { return pop_common(1, decoded_buffer, decoded_size, nullptr); }
Result B200VideoDecoder::get_frame_device(uint8_t **device_ptr) { return pop_common(2, nullptr, 0, device_ptr); }

// reference: software_video_decoder.cpp:345-347 (no-op there; here it completes the async D2H copies)
Result B200VideoDecoder::wait_until_frames_copied() {
  std::lock_guard<std::mutex> lk(mu_);
  if (!dev_) return Result();
  drain_copies();
  if (!sticky_error_.empty()) return Result(false, sticky_error_);
  return Result();
}

B200Stats B200VideoDecoder::stats() {
  std::lock_guard<std::mutex> lk(mu_);
  stats_.kernel_launches = dev_ ? hwb_dev_launch_count(dev_) : 0;
  return stats_;
}

}  // namespace hwang
