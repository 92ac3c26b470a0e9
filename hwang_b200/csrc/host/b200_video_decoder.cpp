#include "b200_video_decoder.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <thread>

#include "../dev/entropy.h"  // sizeof(NbCtx)
#include "../dev/picture.h"  // make_item

namespace hwang {

using hwb::ChunkCtx;
using hwb::PicDesc;
using hwb::SliceDesc;

namespace {
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
const int kMaxChunkPictures = 32767;  // frame indices travel as int16 (SliceDesc::ref_frame, refpic)
}  // namespace

B200VideoDecoder::B200VideoDecoder(int device_id, DeviceType output_type, int) : device_id_(device_id), output_type_(output_type) {
  if (hwb_dev_open(device_id, &dev_) != 0) dev_ = nullptr;
  if (const char *e = getenv("HWB_CHUNK_PICTURES")) { int v = atoi(e); if (v > 0) chunk_target_ = v; }
  if (const char *e = getenv("HWB_RAMP_FIRST")) { int v = atoi(e); if (v > 0) ramp_first_ = v; }
  ramp_target_ = ramp_first_;
  if (const char *e = getenv("HWB_DEBLOCK_BAND")) { int v = atoi(e); if (v >= 1 && v <= hwb::DEBLOCK_BAND) deblock_band_ = v; }
  if (const char *e = getenv("HWB_CONCURRENT")) concurrent_ = atoi(e) != 0;
  if (const char *e = getenv("HWB_INTRA_RESERVE")) intra_reserve_ = atoi(e) != 0;
  if (const char *e = getenv("HWB_GROUP_PICTURES")) { int v = atoi(e); if (v > 0) group_target_ = v; }
  if (const char *e = getenv("HWB_PICTURE_PROFILE")) picture_profile_ = atoi(e) != 0;
  if (const char *e = getenv("HWB_NO_RGB")) no_rgb_ = atoi(e) != 0;  // experiments: no fused RGB24 writeback (frames are converted on demand)
}

B200VideoDecoder::~B200VideoDecoder() {
  if (!dev_) return;
  release_all();
  hwb_dev_close(dev_);
}

void B200VideoDecoder::recycle(std::unique_ptr<Chunk> &c, bool keep_slab) {
  if (!c) return;
  if (c->ev_begin) hwb_dev_event_destroy(dev_, c->ev_begin);
  if (c->ev_entropy) hwb_dev_event_destroy(dev_, c->ev_entropy);
  if (c->ev_picture) hwb_dev_event_destroy(dev_, c->ev_picture);
  if (c->ev_done) hwb_dev_event_destroy(dev_, c->ev_done);
  if (c->ev_copied) hwb_dev_event_destroy(dev_, c->ev_copied);
  if (c->done_host) { free_flags_.push_back({(int32_t *)c->done_host, c->done_capacity}); c->done_host = nullptr; }
  if (c->slab.base) {
    if (keep_slab) free_slabs_.push_back(c->slab);
    else { live_bytes_ -= c->slab.size; hwb_dev_free(dev_, c->slab.base); }
  }
  c.reset();
}

void B200VideoDecoder::release_all() {
  for (int i = 0; i < HWB_NUM_STREAMS; ++i) hwb_dev_stream_sync(dev_, i);
  for (auto &c : queue_) recycle(c, false);
  queue_.clear();
  for (auto &c : retired_) recycle(c, false);
  retired_.clear();
  for (auto &s : free_slabs_) { live_bytes_ -= s.size; hwb_dev_free(dev_, s.base); }
  free_slabs_.clear();
  for (auto &f : free_flags_) hwb_dev_free_host(dev_, f.first);
  free_flags_.clear();
  cur_.reset();
  for (int i = 0; i < kRing; ++i) {
    if (stage_dev_[i]) hwb_dev_free(dev_, stage_dev_[i]);
    if (stage_pinned_[i]) hwb_dev_free_host(dev_, stage_pinned_[i]);
    if (staged_[i].done) hwb_dev_event_destroy(dev_, staged_[i].done);
    stage_dev_[i] = stage_pinned_[i] = nullptr;
    staged_[i] = Staged();
  }
  if (interval_begin_) { hwb_dev_event_destroy(dev_, interval_begin_); interval_begin_ = nullptr; }
  live_bytes_ = 0;
  ring_bytes_ = 0;
  ring_next_ = 0;
}

// Drop every queued / in-flight chunk but keep slabs and staging rings for reuse.
void B200VideoDecoder::reset_keep_memory() {
  for (int i = 0; i < HWB_NUM_STREAMS; ++i) hwb_dev_stream_sync(dev_, i);
  for (auto &c : queue_) recycle(c, true);
  queue_.clear();
  for (auto &c : retired_) recycle(c, true);
  retired_.clear();
  cur_.reset();
  for (int i = 0; i < kRing; ++i) staged_[i].busy = false;
  ring_next_ = 0;
  if (interval_begin_) { hwb_dev_event_destroy(dev_, interval_begin_); interval_begin_ = nullptr; }
}

// Device bytes one picture of the configured stream costs in a chunk (frame buffer, macroblock records, motion,
// worst-case coefficient arena, RGB24 output, progress counters).
size_t B200VideoDecoder::picture_bytes() const {
  const size_t nmb = (size_t)stream_.mb_w() * stream_.mb_h();
  return nmb * (384 + sizeof(hwb::MbInfo) + 2 * 64 + 2 * 4 + 2 * 8 + hwb::SLOTS_PER_MB * 32) + (size_t)width_ * height_ * 3 +
         (size_t)stream_.mb_h() * 12 + (size_t)stream_.mb_w() * sizeof(hwb::NbCtx) + sizeof(PicDesc) + sizeof(SliceDesc) + 4096;
}

// reference: SoftwareVideoDecoder::configure, software_video_decoder.cpp:103-165
Result B200VideoDecoder::configure(const FrameInfo &metadata, const std::vector<uint8_t> &extradata) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!dev_) return Result(false, "B200 decoder: CUDA device unavailable");
  if (!(metadata.format == "h264" || metadata.format == "avc1"))
    return Result(false, "Unsupported video codec: " + metadata.format + " (supports h264 only)");
  // hwang re-configures per interval (python/hwang/decoder.py:65): keep the device memory when the geometry is unchanged
  const bool same_geometry = configured_ && width_ == metadata.width && height_ == metadata.height;
  const int old_mbw = stream_.mb_w(), old_mbh = stream_.mb_h(), old_cx = stream_.crop_left(), old_cy = stream_.crop_top();
  // batch retrieval: the pending chunk and everything queued stay alive across a re-configure of equal geometry
  const bool keep_work = defer_submit_ && same_geometry;
  if (!keep_work) {
    if (same_geometry) reset_keep_memory(); else release_all();
    sticky_error_.clear();
  }
  configured_ = false;  // every failing path below leaves the decoder unconfigured (a later feed reports it)
  hint_valid_ = false;
  std::string err = stream_.configure(extradata.data(), extradata.size());
  if (!err.empty()) return Result(false, "B200 decoder: " + err);
  if ((uint32_t)stream_.width() != metadata.width || (uint32_t)stream_.height() != metadata.height)
    return Result(false, "B200 decoder: container size " + std::to_string(metadata.width) + "x" + std::to_string(metadata.height) +
                             " does not match the SPS (" + std::to_string(stream_.width()) + "x" + std::to_string(stream_.height()) + ")");
  if (keep_work && (stream_.mb_w() != old_mbw || stream_.mb_h() != old_mbh || stream_.crop_left() != old_cx || stream_.crop_top() != old_cy)) {
    // same display size, different coded size or cropping origin: cannot share a chunk
    Result r = submit_current();
    if (!r.ok) return r;
  }
  width_ = metadata.width; height_ = metadata.height;
  ring_bytes_ = (size_t)width_ * height_ * 3;
  size_t free_b = 0, total_b = 0;
  if (hwb_dev_mem_info(dev_, &free_b, &total_b) != 0) return Result(false, std::string("B200 decoder: ") + hwb_dev_error(dev_));
  // in-flight limit: 70% of what this process can get (its own cached slabs count as available)
  memory_budget_ = (size_t)((double)(free_b + live_bytes_) * 0.7);
  if (const char *e = getenv("HWB_MEMORY_BUDGET_MB")) { long v = atol(e); if (v > 0) memory_budget_ = (size_t)v << 20; }
  configured_ = true;
  return Result();
}

void B200VideoDecoder::set_interval_hint(uint64_t start_frame, const std::vector<uint64_t> &wanted) {
  std::lock_guard<std::mutex> lk(mu_);
  hint_valid_ = true; hint_start_ = start_frame; hint_wanted_ = wanted;  // consumed when the next segment opens (feed)
  hint_span_ = wanted.empty() ? 0 : wanted.back() - start_frame + 1;
}

B200VideoDecoder::Slab B200VideoDecoder::take_slab(size_t n) {
  // best fit among the cached slabs
  int best = -1;
  for (size_t i = 0; i < free_slabs_.size(); ++i)
    if (free_slabs_[i].size >= n && (best < 0 || free_slabs_[i].size < free_slabs_[best].size)) best = (int)i;
  if (best >= 0 && free_slabs_[best].size <= 2 * n + ((size_t)64 << 20)) { Slab s = free_slabs_[best]; free_slabs_.erase(free_slabs_.begin() + best); return s; }
  Slab s;
  s.base = (uint8_t *)hwb_dev_malloc(dev_, n);
  if (!s.base) {  // make room: drop the cached slabs and retry once
    for (auto &f : free_slabs_) { live_bytes_ -= f.size; hwb_dev_free(dev_, f.base); }
    free_slabs_.clear();
    s.base = (uint8_t *)hwb_dev_malloc(dev_, n);
  }
  s.size = s.base ? n : 0;
  if (s.base) live_bytes_ += n;
  return s;
}

// reference: SoftwareVideoDecoder::feed, software_video_decoder.cpp:167-248
Result B200VideoDecoder::feed(const uint8_t *encoded_buffer, size_t encoded_size, bool keyframe) {
  std::unique_lock<std::mutex> lk(mu_);
  if (!configured_) return Result(false, "B200 decoder: feed before configure");
  if (!sticky_error_.empty()) return Result(false, sticky_error_);
  if (encoded_size == 0 || encoded_buffer == nullptr) {  // end of interval: everything fed becomes poppable
    close_segment();
    return defer_submit_ ? Result() : submit_current();
  }
  (void)keyframe;
  const bool idr = stream_.next_is_idr(encoded_buffer, encoded_size);
  // A chunk is cut at IDR pictures only (nothing refers across), once it holds chunk_target_ pictures or its device
  // footprint reaches half of the memory budget; a single GOP larger than the budget cannot be decoded.
  // A sparse request (a hint that wants less than half of the frames: long-GOP seeks, every n-th frame) is bound by
  // the latency of its GOP chains, not by throughput: all its GOPs should be in ONE batch (their chains then run side
  // by side in one picture kernel), nothing is gained by an early small batch, and RGB24 space is needed for the wanted
  // frames only.
  const bool sparse = hint_valid_ ? hint_wanted_.size() * 2 < (size_t)std::max<uint64_t>(1, hint_span_) : (cur_ && cur_->sparse);
  const size_t pb = picture_bytes() - (sparse ? (size_t)width_ * height_ * 3 * 3 / 4 : 0);
  if (cur_ && idr) {
    const size_t n = cur_->pics.size();
    if (queue_.empty()) ramp_target_ = std::min(ramp_first_, chunk_target_);  // cold pipeline: start small again
    const int target = cur_->sparse ? 4 * chunk_target_ : std::min(ramp_target_, chunk_target_);
    if ((int)n >= target || (n + 1) * pb > memory_budget_ / 2) {
      ramp_target_ = std::min(chunk_target_, ramp_target_ * 2);
      Result r = submit_current();
      if (!r.ok) return r;
    }
  }
  if (cur_ && (int)cur_->pics.size() >= kMaxChunkPictures)
    return Result(false, "B200 decoder: more than " + std::to_string(kMaxChunkPictures) + " pictures without an IDR picture");
  if (cur_ && (cur_->pics.size() + 1) * pb > memory_budget_)
    return Result(false, "B200 decoder: a GOP of more than " + std::to_string(cur_->pics.size()) + " pictures of " + std::to_string(width_) + "x" +
                             std::to_string(height_) + " does not fit the device memory budget (" + std::to_string(memory_budget_ >> 20) + " MB)");
  if (!cur_) {
    cur_.reset(new Chunk());
    // the previous chunk's staging buffer is recycled: fresh memory would cost one page fault per 4 KB of bitstream
    cur_->bitstream = std::move(spare_bits_);
    cur_->bitstream.clear();
    cur_->bitstream.reserve(last_chunk_bytes_ + (last_chunk_bytes_ >> 2) + (1 << 20));
    cur_->crop_x = stream_.crop_left(); cur_->crop_y = stream_.crop_top();
    cur_->sparse = sparse;
    stream_.reset_dpb();
  }
  // open a segment at the first picture after a flush (or of the chunk): it must be an IDR picture
  const bool at_segment_start = cur_->seg_first.empty() || cur_->seg_first.back() == -1;
  if (at_segment_start) {
    if (!idr) return Result(false, "B200 decoder: interval does not start with an IDR picture");
    if (!cur_->seg_first.empty()) cur_->seg_first.pop_back();
    cur_->seg_first.push_back((int)cur_->pics.size());
    cur_->seg_hint.push_back(hint_valid_ ? std::make_pair(hint_start_, hint_wanted_) : std::make_pair(~(uint64_t)0, std::vector<uint64_t>()));
    hint_valid_ = false;
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

// The pictures fed since the last flush form a segment (one interval of one clip): display order and the wanted-frame
// hint are per segment.  The next picture fed opens a new segment and must be an IDR picture.
void B200VideoDecoder::close_segment() {
  if (cur_ && !cur_->seg_first.empty() && cur_->seg_first.back() != -1) cur_->seg_first.push_back(-1);
}

Result B200VideoDecoder::submit_pending() {
  std::lock_guard<std::mutex> lk(mu_);
  return submit_current();
}

Result B200VideoDecoder::submit_current() {
  if (!cur_ || cur_->pics.empty()) { cur_.reset(); return Result(); }
  std::unique_ptr<Chunk> ch = std::move(cur_);
  if (!ch->seg_first.empty() && ch->seg_first.back() == -1) ch->seg_first.pop_back();
  const int P = (int)ch->pics.size(), S = (int)ch->slices.size();
  const int mb_w = stream_.mb_w(), mb_h = stream_.mb_h(), nmb = mb_w * mb_h;
  const size_t nmb_s = align_up((size_t)nmb, 32);  // per-picture stride of the macroblock arrays: pictures never share a 128-byte line
  // ---- display order per segment; unrequested non-reference pictures are not decoded at all (their frame buffers stay
  // undefined: nothing references them and the consumer drops them); wanted pictures get an RGB24 slot
  ch->order.resize(P);
  ch->skipped.assign(P, 0);
  int nskipped = 0, nrgb = 0;
  for (size_t sg = 0; sg < ch->seg_first.size(); ++sg) {
    const int a = ch->seg_first[sg], b = sg + 1 < ch->seg_first.size() ? ch->seg_first[sg + 1] : P;
    for (int i = a; i < b; ++i) ch->order[i] = i;
    std::stable_sort(ch->order.begin() + a, ch->order.begin() + b, [&](int x, int y) { return ch->out_keys[x] < ch->out_keys[y]; });
    const auto &hint = ch->seg_hint[sg];
    const bool hinted = hint.first != ~(uint64_t)0;
    for (int j = a; j < b; ++j) {
      const int pic = ch->order[j];
      bool want = true;
      if (hinted) want = std::binary_search(hint.second.begin(), hint.second.end(), hint.first + (uint64_t)(j - a));
      if (!want && !ch->pics[pic].is_ref) { ch->skipped[pic] = 1; nskipped++; }
      ch->pics[pic].rgb_slot = (want && !no_rgb_) ? nrgb++ : -1;
    }
  }
  // ---- work lists of the picture kernel, see csrc/dev/picture.h.  The batch is cut into groups of whole GOPs (about
  // group_target_ pictures); inside a group items are ordered (level, row, picture), groups follow one another.  The
  // warps flow from one group into the next without a launch boundary, while the groups -- and with them the frames, in
  // display order -- complete one after the other, so the copies to the host start long before the kernel ends.
  // Rows per deblocking item: narrow batches (few GOP chains: long-GOP seeks) are limited by resident warps waiting at
  // the reconstruction's pace, so a warp takes a band of rows; wide batches have work for every warp anyway and run
  // (measured) a little faster with more, smaller items.
  const int band = deblock_band_ > 0 ? deblock_band_ : (ch->sparse ? hwb::DEBLOCK_BAND : 1);
  std::vector<uint32_t> recon_items, deblock_items;
  recon_items.reserve((size_t)(P - nskipped) * mb_h);
  deblock_items.reserve((size_t)(P - nskipped) * mb_h);
  {
    std::vector<std::vector<int32_t>> by_level;
    int g0 = 0;
    while (g0 < P) {
      int g1 = g0 + 1;  // pictures of level 0 start a GOP (intra pictures): extend to whole GOPs until the group is large enough
      while (g1 < P && !(ch->pics[g1].level == 0 && g1 - g0 >= group_target_)) ++g1;
      int nlevels = 0;
      for (int i = g0; i < g1; ++i) nlevels = std::max(nlevels, ch->pics[i].level + 1);
      by_level.assign(nlevels, std::vector<int32_t>());
      for (int i = g0; i < g1; ++i) if (!ch->skipped[i]) by_level[ch->pics[i].level].push_back(i);
      for (auto &v : by_level)
        for (int y = 0; y < mb_h; ++y)
          for (int32_t pic : v) {
            recon_items.push_back(hwb::make_item(pic, y, 0));
            if (y % band == 0) deblock_items.push_back(hwb::make_item(pic, y, 1));  // one item per band of rows
          }
      g0 = g1;
    }
  }

  // ---- device layout
  const size_t fs = (size_t)mb_w * 16 * mb_h * 16 * 3 / 2;
  const size_t rgb_bytes = (size_t)width_ * height_ * 3;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 256); return o; };
  const size_t o_frames = take(fs * P), o_mbinfo = take((size_t)P * nmb_s * sizeof(hwb::MbInfo)), o_mv = take((size_t)P * 2 * nmb_s * 64),
               o_refidx = take((size_t)P * 2 * nmb_s * 4), o_refpic = take((size_t)P * 2 * nmb_s * 8),
               o_coefs = take((size_t)P * nmb_s * hwb::SLOTS_PER_MB * 32), o_ectx = take((size_t)S * mb_w * sizeof(hwb::NbCtx)),
               o_bits = take(ch->bitstream.size() + 64), o_pics = take((size_t)P * sizeof(PicDesc)), o_slices = take((size_t)S * sizeof(SliceDesc)),
               o_ritems = take(recon_items.size() * 4), o_ditems = take(deblock_items.size() * 4), o_order = take((size_t)S * 4),
               o_rgb = take(rgb_bytes * (size_t)nrgb);
  // counters zeroed per chunk: tickets (entropy, recon, deblock) + per-slice entropy progress + per-row progress x2 + mv reach + error flag
  const size_t n_sync = 4 + (size_t)S + 4 * (size_t)P * mb_h + (size_t)P + 4 + 2 * hwb::PROF_COUNTERS + 2;
  const size_t o_sync = take(n_sync * 4);
  if (feeder_may_block_ && memory_budget_) {
    // Back-pressure in bytes: wait for the consumer to retire chunks instead of running the device out of memory.
    // (The reference's "more than 8 frames buffered" rule counts frames of one small decoder; here a chunk is GBs.)
    // mu_ is released while waiting; only the feeder thread gets here, so `ch` and the stream state are safe.
    size_t cached = 0;
    for (auto &s : free_slabs_) cached += s.size;
    std::unique_lock<std::mutex> relock(mu_, std::adopt_lock);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(120);
    while (live_bytes_ - cached + off > memory_budget_ && !(queue_.empty() && retired_.empty()) && std::chrono::steady_clock::now() < deadline) {
      memory_cv_.wait_for(relock, std::chrono::milliseconds(2));
      poll_retired();
      cached = 0;
      for (auto &s : free_slabs_) cached += s.size;
    }
    relock.release();
  }
  ch->slab = take_slab(off);
  if (!ch->slab.base) return Result(false, std::string("B200 decoder: device allocation of ") + std::to_string(off) + " bytes failed: " + hwb_dev_error(dev_));
  uint8_t *b = ch->slab.base;
  ChunkCtx &c = ch->ctx;
  memset(&c, 0, sizeof(c));
  c.mb_w = mb_w; c.mb_h = mb_h; c.nmb = nmb; c.nmb_stride = (int32_t)nmb_s; c.wc = mb_w * 16; c.hc = mb_h * 16; c.num_pics = P; c.num_slices = S;
  c.frames = b + o_frames; c.frame_stride = fs;
  c.mbinfo = (hwb::MbInfo *)(b + o_mbinfo); c.mv = (int16_t *)(b + o_mv); c.refidx = (int8_t *)(b + o_refidx); c.refpic = (int16_t *)(b + o_refpic);
  c.coefs = (int16_t *)(b + o_coefs); c.ectx = b + o_ectx; c.ectx_stride = (uint64_t)mb_w * sizeof(hwb::NbCtx);
  c.bitstream = b + o_bits; c.pics = (const PicDesc *)(b + o_pics); c.slices = (const SliceDesc *)(b + o_slices);
  c.recon_items = (const uint32_t *)(b + o_ritems); c.deblock_items = (const uint32_t *)(b + o_ditems);
  c.num_recon_items = (int32_t)recon_items.size(); c.num_deblock_items = (int32_t)deblock_items.size();
  c.rgb = b + o_rgb; c.rgb_stride = rgb_bytes;
  c.crop_x = ch->crop_x; c.crop_y = ch->crop_y; c.out_w = (int32_t)width_; c.out_h = (int32_t)height_;
  c.deblock_band = band;
  int32_t *sync = (int32_t *)(b + o_sync);
  int32_t *tickets = sync;
  c.entropy_prog = sync + 4;
  c.recon_prog = c.entropy_prog + S;
  c.dbl_prog = c.recon_prog + (size_t)P * mb_h;
  c.mv_reach = c.dbl_prog + (size_t)P * mb_h;
  c.mv_reach_x = c.mv_reach + (size_t)P * mb_h;
  c.rows_done = c.mv_reach_x + (size_t)P * mb_h;
  c.error_flag = c.rows_done + P;
  ch->error_dev = c.error_flag;
  // completion flags in page-locked host memory (recycled between chunks)
  for (size_t i = 0; i < free_flags_.size(); ++i)
    if (free_flags_[i].second >= (size_t)P) { ch->done_host = free_flags_[i].first; ch->done_capacity = free_flags_[i].second; free_flags_.erase(free_flags_.begin() + i); break; }
  if (!ch->done_host) {
    ch->done_capacity = std::max<size_t>((size_t)P, 1024);
    ch->done_host = (int32_t *)hwb_dev_malloc_host(dev_, ch->done_capacity * 4);
    if (!ch->done_host) return Result(false, std::string("B200 decoder: page-locked allocation failed: ") + hwb_dev_error(dev_));
  }
  memset((void *)ch->done_host, 0, (size_t)P * 4);
  c.pic_done = (int32_t *)ch->done_host;
  if (picture_profile_) c.prof = (unsigned long long *)(((uintptr_t)(c.error_flag + 2) + 7) & ~(uintptr_t)7);

  // Inputs + entropy decoding on one of the rotating entropy streams, the picture kernel on the (single) picture
  // stream: entropy decoding of later chunks (latency-bound: one warp per slice) runs under the picture kernels of
  // earlier ones, whose frames travel to the host meanwhile.
  const int st = HWB_STREAM_ENTROPY0 + (next_entropy_stream_++ % HWB_NUM_ENTROPY_STREAMS), st_pic = HWB_STREAM_PICTURE;
  ch->ev_begin = hwb_dev_event_create(dev_);
  ch->ev_entropy = hwb_dev_event_create(dev_);
  ch->ev_picture = hwb_dev_event_create(dev_);
  ch->ev_done = hwb_dev_event_create(dev_);
  int rc = 0;
  last_chunk_bytes_ = ch->bitstream.size();
  ch->bitstream.resize(ch->bitstream.size() + 64, 0);  // read-ahead padding for the bit readers
  rc |= hwb_dev_h2d(dev_, st, b + o_bits, ch->bitstream.data(), ch->bitstream.size());
  rc |= hwb_dev_h2d(dev_, st, b + o_pics, ch->pics.data(), (size_t)P * sizeof(PicDesc));
  rc |= hwb_dev_h2d(dev_, st, b + o_slices, ch->slices.data(), (size_t)S * sizeof(SliceDesc));
  if (!recon_items.empty()) {
    rc |= hwb_dev_h2d(dev_, st, b + o_ritems, recon_items.data(), recon_items.size() * 4);
    rc |= hwb_dev_h2d(dev_, st, b + o_ditems, deblock_items.data(), deblock_items.size() * 4);
  }
  // Entropy tickets: intra slices carry several times the bits of inter slices and wait on nothing, so they start
  // first; everything else keeps decode order (a B slice's co-located picture then always holds an earlier ticket).
  std::vector<int32_t> order;
  order.reserve(S);
  for (int i = 0; i < S; ++i) if (ch->slices[i].slice_type == hwb::SLICE_I && !ch->skipped[ch->slices[i].pic]) order.push_back(i);
  for (int i = 0; i < S; ++i) if (ch->slices[i].slice_type != hwb::SLICE_I && !ch->skipped[ch->slices[i].pic]) order.push_back(i);
  if (!order.empty()) rc |= hwb_dev_h2d(dev_, st, b + o_order, order.data(), order.size() * 4);
  c.entropy_order = (const int32_t *)(b + o_order);
  c.num_tickets = (int32_t)order.size();
  for (int32_t i : order) if (ch->slices[i].slice_type == hwb::SLICE_I) c.num_intra_tickets++;
  // SMs reserved for the intra slices (see kernels.cu; HWB_INTRA_RESERVE=0 turns it off): an intra slice is a warp's
  // serial work for a fifth of a second, and it runs 1.4x faster on an SM whose instruction caches hold the intra path
  // only.  Every launch reserves the SMs [0, intra slices / 12): the batches of a request are in flight together, and
  // with a common base their intra slices share SMs with each other rather than with other launches' inter slices (first
  // batch of the dense benchmark request: entropy stage done after 234 ms instead of 296 with a base that rotated).
  static const int intra_per_sm = [] { const char *e = getenv("HWB_INTRA_WARPS_PER_SM"); const int v = e ? atoi(e) : 12; return v > 0 ? v : 12; }();
  if (intra_reserve_ && c.num_intra_tickets > 0 && c.num_intra_tickets < c.num_tickets) {
    c.intra_sms = std::min(37, (c.num_intra_tickets + intra_per_sm - 1) / intra_per_sm);
    c.intra_sm_base = 0;
  }
  // (Tried and removed: while intra slices of the launch are running only one or two of the four warps of a block decode
  // inter slices, so that the intra slices run at the speed they have on an idle GPU.  They finished no earlier, and the
  // batch took 391 ms instead of 281: profiles/r2_runs/r2ah_ab.txt.)
  rc |= hwb_dev_memset(dev_, st, sync, 0, n_sync * 4);
  stats_.h2d_bytes += ch->bitstream.size() + (size_t)P * sizeof(PicDesc) + (size_t)S * sizeof(SliceDesc) + (recon_items.size() * 2 + order.size()) * 4;
  // a copy from pageable memory has been staged by the time cudaMemcpyAsync returns: the buffer can be reused
  spare_bits_ = std::move(ch->bitstream);
  rc |= hwb_dev_event_record(dev_, ch->ev_begin, st);  // inputs are resident in HBM from here on
  if (!interval_begin_) { interval_begin_ = hwb_dev_event_create(dev_); if (interval_begin_) rc |= hwb_dev_event_record(dev_, interval_begin_, st); }  // a busy period of the device starts
  int mode = ch->pics[0].cabac ? 1 : 0;
  for (auto &p : ch->pics) if ((p.cabac ? 1 : 0) != mode) mode = -1;
  if (mode == 1) {  // CABAC throughout: copies of the kernel without B-slice support / without the 8x8 transform when the chunk has none
    bool has_b = false, t8 = false;  // (the per-macroblock path has to fit the SM's instruction cache: DESIGN.md 4a)
    for (auto &sl : ch->slices) has_b |= sl.slice_type == hwb::SLICE_B;
    for (auto &p : ch->pics) t8 |= p.transform8x8_mode != 0;
    static const bool no_4 = getenv("HWB_NO_IP4") != nullptr;
    if (no_4) t8 = true;
    mode = has_b ? (t8 ? 1 : 5) : (t8 ? 3 : 4);
  }
  // Default: the picture kernel starts when the batch's entropy kernel has finished (an event orders them).
  // HWB_CONCURRENT=1 launches it right behind the entropy kernel instead; it then waits for the entropy stage picture
  // by picture (csrc/dev/picture.h) and both grids are sized to fit the SMs together (hwb_dev_set_occupancy).  Measured on
  // the 3000-frame benchmark clip this gains nothing: the first pictures wait for their intra slices either way (a
  // 1080p intra slice takes 245 ms on a warp of its own and about 350 ms next to 1700 other slices), and the picture
  // kernel loses a third of its resident warps to the co-residency rule (profiles/r2_runs/r2l_sweep.jsonl).
  hwb_dev_set_occupancy(dev_, concurrent_ ? 2 : 3, concurrent_ ? 4 : 6);
  hwb_event *ev_inputs = hwb_dev_event_create(dev_);
  rc |= hwb_dev_event_record(dev_, ev_inputs, st);  // uploads and the zeroed counters
  if (c.num_tickets > 0) rc |= hwb_dev_entropy(dev_, st, &c, tickets, mode);
  rc |= hwb_dev_event_record(dev_, ch->ev_entropy, st);
  rc |= hwb_dev_stream_wait(dev_, st_pic, concurrent_ ? ev_inputs : ch->ev_entropy);
  hwb_dev_event_destroy(dev_, ev_inputs);  // the wait already enqueued keeps its own reference
  rc |= hwb_dev_event_record(dev_, ch->ev_picture, st_pic);  // previous picture kernel done (and, without concurrency, this batch's entropy stage)
  rc |= hwb_dev_picture(dev_, st_pic, &c, tickets + 1);
  rc |= hwb_dev_event_record(dev_, ch->ev_done, st_pic);
  if (rc) { sticky_error_ = std::string("B200 decoder: CUDA launch failed: ") + hwb_dev_error(dev_); return Result(false, sticky_error_); }
  for (int i = 0; i < P; ++i)
    if (!ch->skipped[i]) ch->alg_bytes += fs + (ch->pics[i].has_inter ? fs : 0) + (ch->pics[i].rgb_slot >= 0 ? rgb_bytes : 0);
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
  close_segment();
  Result r = defer_submit_ ? Result() : submit_current();
  stream_.reset_dpb();
  hint_valid_ = false;  // the hint covers one interval
  return r;
}

Result B200VideoDecoder::finish_chunk(Chunk &c) {
  if (c.checked) return Result();
  if (hwb_dev_event_sync(dev_, c.ev_done) != 0) return Result(false, std::string("B200 decoder: CUDA error: ") + hwb_dev_error(dev_));
  int32_t flag = 0;
  hwb_dev_d2h(dev_, HWB_STREAM_AUX, &flag, c.error_dev, 4);
  hwb_dev_stream_sync(dev_, HWB_STREAM_AUX);
  float ms = 0;
  static const bool trace = getenv("HWB_TRACE_BATCHES") != nullptr;  // device timeline of every batch, relative to the busy period's start
  if (trace && interval_begin_) {
    float b = 0, e = 0, p = 0, d = 0;
    hwb_dev_event_elapsed(dev_, interval_begin_, c.ev_begin, &b); hwb_dev_event_elapsed(dev_, interval_begin_, c.ev_entropy, &e);
    hwb_dev_event_elapsed(dev_, interval_begin_, c.ev_picture, &p); hwb_dev_event_elapsed(dev_, interval_begin_, c.ev_done, &d);
    fprintf(stderr, "[batch] pictures %4d  inputs resident %7.1f  entropy done %7.1f  picture kernel %7.1f .. %7.1f ms\n", (int)c.pics.size(), b, e, p, d);
  }
  if (profile_) {
    if (hwb_dev_event_elapsed(dev_, c.ev_begin, c.ev_entropy, &ms) == 0) { stats_.entropy_ms += ms; stats_.entropy_launches++; }
    if (hwb_dev_event_elapsed(dev_, c.ev_picture, c.ev_done, &ms) == 0) { stats_.picture_ms += ms; stats_.picture_launches++; }
    // device wall clock: from the first chunk since the last reading to this one (chunks finish in order)
    if (interval_begin_ && hwb_dev_event_elapsed(dev_, interval_begin_, c.ev_done, &ms) == 0) {
      bool last_in_flight = true;
      for (auto &q : queue_) if (q.get() != &c && q->submitted && !q->checked) last_in_flight = false;
      if (last_in_flight && !cur_) {
        stats_.wall_ms += ms;
        hwb_dev_event_destroy(dev_, interval_begin_);
        interval_begin_ = nullptr;
      }
    }
  }
  if (c.ctx.prof) {  // HWB_PICTURE_PROFILE=1: where the picture kernel's warps spent their cycles
    unsigned long long v[hwb::PROF_COUNTERS] = {0};
    hwb_dev_d2h(dev_, HWB_STREAM_AUX, v, c.ctx.prof, sizeof(v));
    hwb_dev_stream_sync(dev_, HWB_STREAM_AUX);
    const double life = (double)v[hwb::PROF_LIFETIME] > 0 ? (double)v[hwb::PROF_LIFETIME] : 1.0;
    fprintf(stderr, "[hwb picture profile] pictures %d: recon_mb %.1f%% deblock_mb %.1f%% rgb %.1f%% | wait ref %.1f%% intra %.1f%% recon %.1f%% deblock-above %.1f%% | pick %.1f%% (warp-cycles %.3g)\n",
            c.ctx.num_pics, 100 * v[hwb::PROF_RECON_MB] / life, 100 * v[hwb::PROF_DEBLOCK_MB] / life, 100 * v[hwb::PROF_RGB] / life,
            100 * v[hwb::PROF_WAIT_REF] / life, 100 * v[hwb::PROF_WAIT_INTRA] / life, 100 * v[hwb::PROF_WAIT_RECON] / life,
            100 * v[hwb::PROF_WAIT_DEBLOCK_ABOVE] / life, 100 * v[hwb::PROF_PICK] / life, life);
  }
  stats_.algorithmic_bytes += c.alg_bytes;
  c.finished = true; c.checked = true;
  if (flag) { sticky_error_ = "B200 decoder: corrupt or unsupported bitstream (device error code " + std::to_string(flag) + ")"; return Result(false, sticky_error_); }
  return Result();
}

bool B200VideoDecoder::picture_done(Chunk &c, int frame) {
  return c.finished || c.skipped[frame] || c.done_host[frame] != 0;
}

// Frames that can be popped without waiting: pictures complete in display order from the front of the queue.
int B200VideoDecoder::frames_ready() {
  std::lock_guard<std::mutex> lk(mu_);
  if (!sticky_error_.empty()) return -1;
  if (!retired_.empty()) poll_retired();
  int n = 0;
  for (auto &c : queue_) {
    if (!c->finished) {
      int st = hwb_dev_event_done(dev_, c->ev_done);
      if (st < 0) { sticky_error_ = std::string("B200 decoder: CUDA error: ") + hwb_dev_error(dev_); return -1; }
      if (st == 1) c->finished = true;
    }
    size_t k = c->next_out;
    while (k < c->order.size() && picture_done(*c, c->order[k])) ++k;
    n += (int)(k - c->next_out);
    if (k < c->order.size()) break;
  }
  return n;
}

// Wait for one picture (its completion flag in host memory), not for the whole batch.  The flag never comes when the
// entropy stage found the stream corrupt (the picture kernel then does nothing): the batch's end event ends the wait.
Result B200VideoDecoder::wait_picture(Chunk &c, int frame) {
  int spins = 0;
  while (!picture_done(c, frame)) {
    const int st = hwb_dev_event_done(dev_, c.ev_done);
    if (st < 0) { sticky_error_ = std::string("B200 decoder: CUDA error: ") + hwb_dev_error(dev_); return Result(false, sticky_error_); }
    if (st == 1) { c.finished = true; break; }
    if (++spins > 64) std::this_thread::sleep_for(std::chrono::microseconds(20));
  }
  if (c.finished) return finish_chunk(c);  // whole batch done: error flag, statistics
  return Result();
}

// reference: SoftwareVideoDecoder::decoded_frames_buffered, software_video_decoder.cpp:341-343.
// The automaton's feeder pauses while this exceeds 8 (decoder_automata.cpp:288-293).  A batch decoder wants the
// interval fed while earlier chunks are still being consumed, so the value saturates at 8 ("at least this many frames
// are ready"): a reference-style caller keeps feeding, and back-pressure is applied in bytes of device memory inside
// feed() instead (set_feeder_may_block), or reported as an allocation error to single-threaded callers.
int B200VideoDecoder::decoded_frames_buffered() {
  int n = frames_ready();
  if (n < 0) return 1;  // error state: let the consumer pop, get_frame / discard_frame report the error
  return n > 8 ? 8 : n;
}

// A fully popped chunk: its memory may be reused once the copies out of it have completed, which an event on the copy
// stream (ordered after the auxiliary stream's conversions) tells without blocking anybody.
void B200VideoDecoder::retire_front() {
  std::unique_ptr<Chunk> c = std::move(queue_.front());
  queue_.pop_front();
  if (!c->checked) finish_chunk(*c);  // every frame has left: the kernel is at most a few warps from its end
  c->ev_copied = hwb_dev_event_create(dev_);
  hwb_event *aux = hwb_dev_event_create(dev_);
  if (aux) { hwb_dev_event_record(dev_, aux, HWB_STREAM_AUX); hwb_dev_stream_wait(dev_, HWB_STREAM_COPY, aux); hwb_dev_event_destroy(dev_, aux); }
  if (c->ev_copied) hwb_dev_event_record(dev_, c->ev_copied, HWB_STREAM_COPY);
  retired_.push_back(std::move(c));
  poll_retired();
}

// Recycle retired chunks whose copies are done (chunks that lent out device pointers wait for drain_copies).
void B200VideoDecoder::poll_retired() {
  bool any = false;
  for (size_t i = 0; i < retired_.size();) {
    Chunk &c = *retired_[i];
    if (!c.lent && c.ev_copied && hwb_dev_event_done(dev_, c.ev_copied) == 1) {
      recycle(retired_[i], true);
      retired_.erase(retired_.begin() + i);
      any = true;
    } else ++i;
  }
  if (any) memory_cv_.notify_all();
}

// Completes every copy in flight (page-locked destinations are written by the copy engine, pageable ones get their
// memcpy from the staging ring here) and recycles the memory of fully popped chunks.
void B200VideoDecoder::drain_copies() {
  hwb_dev_stream_sync(dev_, HWB_STREAM_COPY);
  hwb_dev_stream_sync(dev_, HWB_STREAM_AUX);
  for (int i = 0; i < kRing; ++i)
    if (staged_[i].busy) { if (staged_[i].user) memcpy(staged_[i].user, stage_pinned_[i], staged_[i].size); staged_[i].busy = false; }
  for (auto &c : retired_) recycle(c, true);
  retired_.clear();
  // cached slabs go back to the driver only when this decoder holds more than its budget (cudaFree synchronises the
  // device, and a dense pass reuses every slab of the previous one: batches ramp up through the same sizes)
  while (!free_slabs_.empty() && live_bytes_ > memory_budget_) { live_bytes_ -= free_slabs_.back().size; hwb_dev_free(dev_, free_slabs_.back().base); free_slabs_.pop_back(); }
  memory_cv_.notify_all();
}

// A free slot of the page-locked staging ring (allocated on first use); completes the slot's previous copy first.
Result B200VideoDecoder::stage_slot(int *slot) {
  const int s = ring_next_;
  ring_next_ = (ring_next_ + 1) % kRing;
  if (!stage_dev_[s]) {
    stage_dev_[s] = (uint8_t *)hwb_dev_malloc(dev_, ring_bytes_);
    stage_pinned_[s] = (uint8_t *)hwb_dev_malloc_host(dev_, ring_bytes_);
    staged_[s].done = hwb_dev_event_create(dev_);
    if (!stage_dev_[s] || !stage_pinned_[s] || !staged_[s].done) return Result(false, std::string("B200 decoder: out of memory: ") + hwb_dev_error(dev_));
  }
  if (staged_[s].busy) {
    hwb_dev_event_sync(dev_, staged_[s].done);
    if (staged_[s].user) memcpy(staged_[s].user, stage_pinned_[s], staged_[s].size);
    staged_[s].busy = false;
  }
  *slot = s;
  return Result();
}

// mode 0: RGB24 to the caller's buffer, 1: planar I420 to host, 2: RGB24 left in the decoder's device memory, 3: discard
Result B200VideoDecoder::pop_common(int mode, uint8_t *buf, size_t size, uint8_t **dev_out) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!sticky_error_.empty()) return Result(false, sticky_error_);
  while (!queue_.empty() && queue_.front()->next_out >= queue_.front()->order.size()) retire_front();
  if (queue_.empty()) return Result(false, "B200 decoder: no decoded frame buffered");
  Chunk &c = *queue_.front();
  const int frame = c.order[c.next_out];
  HWANG_RETURN_ON_ERROR(wait_picture(c, frame));
  c.next_out++;
  if (mode != 3 && c.skipped[frame]) return Result(false, "B200 decoder: this frame was declared unwanted (set_interval_hint) and has not been decoded");
  if (mode == 3) {
    if (c.next_out >= c.order.size()) retire_front();
    return Result();
  }
  const size_t rgb_bytes = (size_t)width_ * height_ * 3;
  const size_t need = mode == 1 ? rgb_bytes / 2 : rgb_bytes;
  if (mode != 2 && size < need) return Result(false, "B200 decoder: output buffer too small");
  int rc = 0;
  const int slot = c.pics[frame].rgb_slot;
  const uint8_t *src = slot >= 0 ? c.ctx.rgb + (size_t)slot * c.ctx.rgb_stride : nullptr;
  if (mode == 1 || !src) {
    // planar test output, or a decoded picture nobody announced (reference picture outside the hint): converted on
    // demand by a small kernel on the auxiliary stream into a staging slot
    int s = 0;
    HWANG_RETURN_ON_ERROR(stage_slot(&s));
    if (mode == 1) rc = hwb_dev_yuv(dev_, HWB_STREAM_AUX, &c.ctx, frame, c.crop_x, c.crop_y, (int)width_, (int)height_, stage_dev_[s]);
    else rc = hwb_dev_rgb24(dev_, HWB_STREAM_AUX, &c.ctx, frame, c.crop_x, c.crop_y, (int)width_, (int)height_, stage_dev_[s]);
    stats_.aux_launches++;
    if (mode == 2) {
      rc |= hwb_dev_event_record(dev_, staged_[s].done, HWB_STREAM_AUX);
      staged_[s].user = nullptr; staged_[s].size = 0; staged_[s].busy = true;
      *dev_out = stage_dev_[s];
    } else {
      const int kind = hwb_dev_pointer_kind(dev_, buf);
      if (kind == 2) rc |= hwb_dev_d2d(dev_, HWB_STREAM_AUX, buf, stage_dev_[s], need);
      else rc |= hwb_dev_d2h(dev_, HWB_STREAM_AUX, kind == 1 ? buf : stage_pinned_[s], stage_dev_[s], need);
      rc |= hwb_dev_event_record(dev_, staged_[s].done, HWB_STREAM_AUX);
      staged_[s].user = kind == 0 ? buf : nullptr; staged_[s].size = need; staged_[s].busy = true;
      if (kind != 2) stats_.d2h_bytes += need;
    }
  } else if (mode == 2) {
    *dev_out = const_cast<uint8_t *>(src);  // stays valid until the chunk's memory is recycled (drain_copies)
    c.lent = true;
  } else {
    // The frame is already RGB24 in the chunk's arena (written by the deblocking pass).  Device destination: one
    // device-to-device copy; page-locked host destination: the copy engine writes it directly; pageable: via the ring.
    const int kind = hwb_dev_pointer_kind(dev_, buf);
    if (kind == 2) rc = hwb_dev_d2d(dev_, HWB_STREAM_COPY, buf, src, need);
    else if (kind == 1) { rc = hwb_dev_d2h(dev_, HWB_STREAM_COPY, buf, src, need); stats_.d2h_bytes += need; }
    else {
      int s = 0;
      HWANG_RETURN_ON_ERROR(stage_slot(&s));
      rc = hwb_dev_d2h(dev_, HWB_STREAM_COPY, stage_pinned_[s], src, need);
      rc |= hwb_dev_event_record(dev_, staged_[s].done, HWB_STREAM_COPY);
      staged_[s].user = buf; staged_[s].size = need; staged_[s].busy = true;
      stats_.d2h_bytes += need;
    }
  }
  if (rc) { sticky_error_ = std::string("B200 decoder: CUDA error: ") + hwb_dev_error(dev_); return Result(false, sticky_error_); }
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

// reference: software_video_decoder.cpp:345-347 (no-op there; here it completes the async copies)
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
