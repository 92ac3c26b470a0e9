// C ABI (include/hwang_b200.h) over the C++ host side.  No exceptions cross this boundary.
#include <algorithm>
#include <atomic>
#include "../../../include/hwang_b200.h"

#include <string.h>
#include <new>
#include <string>
#include <vector>

#include "b200_video_decoder.h"
#include "decoder_automata.h"
#include "mp4_index_creator.h"
#include "video_decoder_factory.h"
#include "video_index.h"

using namespace hwang;

struct hwb_decoder {
  VideoDecoderInterface *dec = nullptr;
  B200VideoDecoder *b200 = nullptr;
  std::string err;
};
struct hwb_index_creator { MP4IndexCreator c; explicit hwb_index_creator(uint64_t n) : c(n) {} };
struct hwb_video_index { VideoIndex v; };
struct hwb_automata { DecoderAutomata *a = nullptr; std::string err; };

namespace {
int ret(hwb_decoder *d, const Result &r) { if (!r.ok) { d->err = r.message; return 1; } return 0; }
int ret(hwb_automata *a, const Result &r) { if (!r.ok) { a->err = r.message; return 1; } return 0; }
hwb_dev *g_pin_dev = nullptr;
// Device the process decodes on (the last one a decoder or an automaton was created for): the page-locked output
// buffers are allocated through its context, so that a rank of a multi-GPU job never opens a context on GPU 0.
std::atomic<int> g_last_device{0};
void fill_stats(const B200Stats &s, hwb_stats *o) {
  o->pictures_decoded = s.pictures_decoded; o->frames_returned = s.frames_returned; o->chunks = s.chunks; o->bitstream_bytes = s.bitstream_bytes;
  o->kernel_launches = s.kernel_launches; o->h2d_bytes = s.h2d_bytes; o->d2h_bytes = s.d2h_bytes; o->algorithmic_bytes = s.algorithmic_bytes;
  o->wall_ms = s.wall_ms; o->entropy_ms = s.entropy_ms; o->picture_ms = s.picture_ms;
  o->entropy_launches = s.entropy_launches; o->picture_launches = s.picture_launches; o->aux_launches = s.aux_launches;
}
}  // namespace

extern "C" {

const char *hwb_version(void) { return "hwang_b200 0.1 (sm_100a)"; }
int hwb_device_count(void) { return hwb_dev_count(); }
int hwb_has_decoder_type(int t) { return VideoDecoderFactory::has_decoder_type((VideoDecoderType)t) ? 1 : 0; }

int hwb_decoder_create(int device_type, int device_id, int num_devices, int decoder_type, hwb_decoder **out) {
  *out = nullptr;
  DeviceHandle h{(DeviceType)device_type, device_id};
  VideoDecoderInterface *dec = VideoDecoderFactory::make_from_config(h, (uint32_t)num_devices, (VideoDecoderType)decoder_type);
  if (!dec) return 1;
  hwb_decoder *d = new (std::nothrow) hwb_decoder();
  if (!d) { delete dec; return 1; }
  d->dec = dec; d->b200 = dynamic_cast<B200VideoDecoder *>(dec);
  if (device_id >= 0) g_last_device = device_id;
  *out = d;
  return 0;
}
void hwb_decoder_destroy(hwb_decoder *d) { if (d) { delete d->dec; delete d; } }
int hwb_decoder_configure(hwb_decoder *d, uint32_t w, uint32_t h, const char *format, const uint8_t *extra, size_t n) {
  VideoDecoderInterface::FrameInfo fi; fi.width = w; fi.height = h; fi.format = format ? format : "";
  return ret(d, d->dec->configure(fi, std::vector<uint8_t>(extra, extra + n)));
}
int hwb_decoder_feed(hwb_decoder *d, const uint8_t *b, size_t n, int kf) { return ret(d, d->dec->feed(b, n, kf != 0)); }
int hwb_decoder_flush(hwb_decoder *d) { return ret(d, d->dec->flush()); }
int hwb_decoder_discard_frame(hwb_decoder *d) { return ret(d, d->dec->discard_frame()); }
int hwb_decoder_get_frame(hwb_decoder *d, uint8_t *b, size_t n) { return ret(d, d->dec->get_frame(b, n)); }
int hwb_decoder_decoded_frames_buffered(hwb_decoder *d) { return d->dec->decoded_frames_buffered(); }
int hwb_decoder_wait_until_frames_copied(hwb_decoder *d) { return ret(d, d->dec->wait_until_frames_copied()); }
const char *hwb_decoder_last_error(hwb_decoder *d) { return d->err.c_str(); }
int hwb_decoder_get_frame_yuv(hwb_decoder *d, uint8_t *b, size_t n) { return d->b200 ? ret(d, d->b200->get_frame_yuv(b, n)) : 1; }
int hwb_decoder_get_frame_device(hwb_decoder *d, uint8_t **p) { return d->b200 ? ret(d, d->b200->get_frame_device(p)) : 1; }
int hwb_decoder_frames_ready(hwb_decoder *d) { return d->b200 ? d->b200->frames_ready() : d->dec->decoded_frames_buffered(); }
int hwb_decoder_set_chunk_pictures(hwb_decoder *d, int n) { if (!d->b200) return 1; d->b200->set_chunk_pictures(n); return 0; }
int hwb_decoder_set_interval_hint(hwb_decoder *d, uint64_t start_frame, const uint64_t *wanted, size_t n) {
  if (!d->b200) return 1;
  d->b200->set_interval_hint(start_frame, std::vector<uint64_t>(wanted, wanted + n));
  return 0;
}
int hwb_decoder_set_defer_submit(hwb_decoder *d, int on) { if (!d->b200) return 1; d->b200->set_defer_submit(on != 0); return 0; }
int hwb_decoder_submit_pending(hwb_decoder *d) { return d->b200 ? ret(d, d->b200->submit_pending()) : 1; }
int hwb_decoder_get_stats(hwb_decoder *d, hwb_stats *out) { if (!d->b200) return 1; fill_stats(d->b200->stats(), out); return 0; }

void *hwb_alloc_pinned(size_t n) {
  if (!g_pin_dev && hwb_dev_open(g_last_device.load(), &g_pin_dev) != 0) return nullptr;
  return hwb_dev_malloc_host(g_pin_dev, n);
}
void hwb_free_pinned(void *p) { if (g_pin_dev && p) hwb_dev_free_host(g_pin_dev, p); }
// Device-memory destinations for DeviceType::GPU output (callers without a CUDA runtime of their own: tests, ctypes)
void *hwb_alloc_device(int device_id, size_t n) {
  hwb_dev *d = nullptr;
  if (hwb_dev_open(device_id, &d) != 0) return nullptr;
  void *p = hwb_dev_malloc(d, n);
  hwb_dev_close(d);
  return p;
}
void hwb_free_device(int device_id, void *p) {
  hwb_dev *d = nullptr;
  if (!p || hwb_dev_open(device_id, &d) != 0) return;
  hwb_dev_free(d, p);
  hwb_dev_close(d);
}
int hwb_copy_device_to_host(int device_id, void *dst, const void *src, size_t n) {
  hwb_dev *d = nullptr;
  if (hwb_dev_open(device_id, &d) != 0) return 1;
  int rc = hwb_dev_d2h(d, HWB_STREAM_AUX, dst, src, n);
  rc |= hwb_dev_stream_sync(d, HWB_STREAM_AUX);
  hwb_dev_close(d);
  return rc;
}

// ---------------------------------------------------------------------------------------- index
hwb_index_creator *hwb_index_creator_create(uint64_t file_size) { return new (std::nothrow) hwb_index_creator(file_size); }
void hwb_index_creator_destroy(hwb_index_creator *c) { delete c; }
int hwb_index_creator_feed(hwb_index_creator *c, const uint8_t *data, size_t size, uint64_t *next_offset, uint64_t *next_size) {
  uint64_t no = 0, ns = 0;
  bool r = c->c.feed(data, size, no, ns);
  *next_offset = no; *next_size = ns;
  return r ? 1 : 0;
}
int hwb_index_creator_is_done(hwb_index_creator *c) { return c->c.is_done(); }
int hwb_index_creator_is_error(hwb_index_creator *c) { return c->c.is_error(); }
const char *hwb_index_creator_error_message(hwb_index_creator *c) { return c->c.error_message().c_str(); }
hwb_video_index *hwb_index_creator_get_video_index(hwb_index_creator *c) { return new (std::nothrow) hwb_video_index{c->c.get_video_index()}; }

hwb_video_index *hwb_video_index_create(uint32_t timescale, uint64_t duration, uint32_t width, uint32_t height, const char *format,
                                        const uint64_t *so, const uint64_t *ss, size_t n, const uint64_t *kf, size_t nk, const uint8_t *md, size_t nmd) {
  return new (std::nothrow) hwb_video_index{VideoIndex(timescale, duration, width, height, format ? format : "", std::vector<uint64_t>(so, so + n),
                                                       std::vector<uint64_t>(ss, ss + n), std::vector<uint64_t>(kf, kf + nk), std::vector<uint8_t>(md, md + nmd))};
}
void hwb_video_index_destroy(hwb_video_index *v) { delete v; }
hwb_video_index *hwb_video_index_deserialize(const uint8_t *data, size_t size) {
  return new (std::nothrow) hwb_video_index{VideoIndex::deserialize(std::vector<uint8_t>(data, data + size))};
}
size_t hwb_video_index_serialize(const hwb_video_index *v, uint8_t *out, size_t cap) {
  std::vector<uint8_t> s = v->v.serialize();
  if (out && cap >= s.size()) memcpy(out, s.data(), s.size());
  return s.size();
}
uint32_t hwb_video_index_timescale(const hwb_video_index *v) { return v->v.timescale(); }
uint64_t hwb_video_index_duration(const hwb_video_index *v) { return v->v.duration(); }
double hwb_video_index_fps(const hwb_video_index *v) { return v->v.fps(); }
uint32_t hwb_video_index_frame_width(const hwb_video_index *v) { return v->v.frame_width(); }
uint32_t hwb_video_index_frame_height(const hwb_video_index *v) { return v->v.frame_height(); }
const char *hwb_video_index_format(const hwb_video_index *v) { return v->v.format().c_str(); }
uint64_t hwb_video_index_frames(const hwb_video_index *v) { return v->v.frames(); }
const uint64_t *hwb_video_index_sample_offsets(const hwb_video_index *v) { return v->v.sample_offsets().data(); }
const uint64_t *hwb_video_index_sample_sizes(const hwb_video_index *v) { return v->v.sample_sizes().data(); }
const uint64_t *hwb_video_index_keyframe_indices(const hwb_video_index *v, size_t *count) { *count = v->v.keyframe_indices().size(); return v->v.keyframe_indices().data(); }
const uint8_t *hwb_video_index_metadata_bytes(const hwb_video_index *v, size_t *size) { *size = v->v.metadata_bytes().size(); return v->v.metadata_bytes().data(); }

int hwb_slice_into_video_intervals(const hwb_video_index *v, const uint64_t *rows, size_t num_rows, uint64_t *starts, uint64_t *ends,
                                   uint64_t *valid_counts, size_t max_intervals, uint64_t *valid_rows) {
  VideoIntervals vi = slice_into_video_intervals(v->v, std::vector<uint64_t>(rows, rows + num_rows));
  if (vi.sample_index_intervals.empty() && num_rows) return -1;
  if (vi.sample_index_intervals.size() > max_intervals) return -1;
  size_t k = 0;
  for (size_t i = 0; i < vi.sample_index_intervals.size(); ++i) {
    starts[i] = std::get<0>(vi.sample_index_intervals[i]); ends[i] = std::get<1>(vi.sample_index_intervals[i]);
    valid_counts[i] = vi.valid_frames[i].size();
    for (uint64_t r : vi.valid_frames[i]) valid_rows[k++] = r;
  }
  return (int)vi.sample_index_intervals.size();
}

// ---------------------------------------------------------------------------------------- automata
hwb_automata *hwb_automata_create(int device_type, int device_id, int num_devices, int decoder_type) {
  DeviceHandle h{(DeviceType)device_type, device_id};
  DecoderAutomata *a = DecoderAutomata::make_instance(h, num_devices, (VideoDecoderType)decoder_type);
  if (!a) return nullptr;
  hwb_automata *w = new (std::nothrow) hwb_automata();
  if (!w) { delete a; return nullptr; }
  w->a = a;
  if (device_id >= 0) g_last_device = device_id;
  return w;
}
void hwb_automata_destroy(hwb_automata *a) { if (a) { delete a->a; delete a; } }
int hwb_automata_initialize(hwb_automata *a, const hwb_encoded_data *iv, size_t n, const uint8_t *extra, size_t nextra) {
  // the previous call's interval buffers are recycled (largest first): assign() into existing capacity touches no new pages
  std::vector<DecoderAutomata::EncodedData> old = a->a->release_intervals();
  std::sort(old.begin(), old.end(), [](const DecoderAutomata::EncodedData &x, const DecoderAutomata::EncodedData &y) {
    return x.encoded_video.capacity() > y.encoded_video.capacity(); });
  std::vector<DecoderAutomata::EncodedData> v(n);
  for (size_t i = 0; i < n; ++i) {
    auto &d = v[i];
    if (i < old.size()) d.encoded_video = std::move(old[i].encoded_video);
    d.encoded_video.assign(iv[i].encoded_video, iv[i].encoded_video + iv[i].encoded_video_size);
    d.width = iv[i].width; d.height = iv[i].height; d.start_keyframe = iv[i].start_keyframe; d.end_keyframe = iv[i].end_keyframe;
    d.format = iv[i].format ? iv[i].format : "";
    d.sample_offsets.assign(iv[i].sample_offsets, iv[i].sample_offsets + iv[i].num_samples);
    d.sample_sizes.assign(iv[i].sample_sizes, iv[i].sample_sizes + iv[i].num_samples);
    d.keyframes.assign(iv[i].keyframes, iv[i].keyframes + iv[i].num_keyframes);
    d.valid_frames.assign(iv[i].valid_frames, iv[i].valid_frames + iv[i].num_valid_frames);
  }
  return ret(a, a->a->initialize(std::move(v), std::vector<uint8_t>(extra, extra + nextra)));
}
int hwb_automata_get_frames(hwb_automata *a, uint8_t *buffer, int32_t n) { return ret(a, a->a->get_frames(buffer, n)); }
int hwb_automata_set_chunk_pictures(hwb_automata *a, int n) {
  B200VideoDecoder *b = dynamic_cast<B200VideoDecoder *>(a->a->decoder());
  if (!b) return 1;
  b->set_chunk_pictures(n);
  return 0;
}
const char *hwb_automata_last_error(hwb_automata *a) { return a->err.c_str(); }
int hwb_automata_get_stats(hwb_automata *a, hwb_stats *out) {
  B200VideoDecoder *b = dynamic_cast<B200VideoDecoder *>(a->a->decoder());
  if (!b) return 1;
  fill_stats(b->stats(), out);
  return 0;
}
}
