/* hwang_b200 -- C ABI of the B200-native H.264 frame-decode engine behind hwang's API.
 *
 * Every entry point replaces (or exposes to an FFI) one interface of the reference
 * scanner-research/hwang; the reference location is cited per group.  Conventions: opaque handles,
 * plain pointers and sizes, int status (0 = ok, non-zero = error; the message is available from the
 * matching *_last_error()).  No C++ types, no exceptions cross this boundary.
 *
 * Threading follows the reference's plugin contract (hwang/decoder_automata.cpp:259-404):
 * hwb_decoder_feed / hwb_decoder_flush may be called on one thread while hwb_decoder_get_frame /
 * hwb_decoder_discard_frame / hwb_decoder_wait_until_frames_copied run on another;
 * hwb_decoder_decoded_frames_buffered may be polled from both.
 */
#ifndef HWANG_B200_H_
#define HWANG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HWB_DEVICE_CPU 0 /* hwang::DeviceType::CPU, hwang/common.h:20-23 */
#define HWB_DEVICE_GPU 1
#define HWB_DECODER_SOFTWARE 0 /* hwang::VideoDecoderType, hwang/video_decoder_factory.h:23-27 */
#define HWB_DECODER_NVIDIA 1
#define HWB_DECODER_INTEL 2
#define HWB_DECODER_B200 3 /* new enumerator: this backend */

const char *hwb_version(void);
/* number of CUDA devices visible (0 when there is none: the library has no CPU fallback) */
int hwb_device_count(void);
/* VideoDecoderFactory::has_decoder_type, hwang/video_decoder_factory.cpp:44-53 */
int hwb_has_decoder_type(int decoder_type);

/* ---------------------------------------------------------------------------------------------
 * Decoder plugin: hwang::VideoDecoderInterface (hwang/video_decoder_interface.h:26-49), created
 * through VideoDecoderFactory::make_from_config (hwang/video_decoder_factory.cpp:55-98).
 * ------------------------------------------------------------------------------------------- */
typedef struct hwb_decoder hwb_decoder;

/* make_from_config(DeviceHandle{device_type, device_id}, num_devices, decoder_type); *out = NULL and a
 * non-zero return when the backend cannot be built (no GPU, unsupported type). */
int hwb_decoder_create(int device_type, int device_id, int num_devices, int decoder_type, hwb_decoder **out);
void hwb_decoder_destroy(hwb_decoder *d);
/* configure(FrameInfo{width,height,format}, extradata = avcC), interface.h:35 */
int hwb_decoder_configure(hwb_decoder *d, uint32_t width, uint32_t height, const char *format, const uint8_t *extradata, size_t extradata_size);
/* feed(encoded_buffer, encoded_size, keyframe); (NULL, 0, 0) = end of interval, interface.h:38 */
int hwb_decoder_feed(hwb_decoder *d, const uint8_t *encoded_buffer, size_t encoded_size, int keyframe);
int hwb_decoder_flush(hwb_decoder *d);                                                   /* interface.h:41 */
int hwb_decoder_discard_frame(hwb_decoder *d);                                           /* interface.h:43 */
/* interface.h:45, RGB24 W*H*3.  decoded_buffer may be pageable host memory, page-locked host memory (written by the
 * copy engine directly) or DEVICE memory (hwang::DeviceType::GPU output, hwang/common.h:20-50: the frame never crosses PCIe). */
int hwb_decoder_get_frame(hwb_decoder *d, uint8_t *decoded_buffer, size_t decoded_size);
int hwb_decoder_decoded_frames_buffered(hwb_decoder *d);                                 /* interface.h:47 */
int hwb_decoder_wait_until_frames_copied(hwb_decoder *d);                                /* interface.h:49 */
const char *hwb_decoder_last_error(hwb_decoder *d);
/* extensions (parity tests / benchmark): */
int hwb_decoder_get_frame_yuv(hwb_decoder *d, uint8_t *decoded_buffer, size_t decoded_size); /* cropped planar I420, W*H*3/2 */
/* zero-copy: RGB24 in the decoder's own device memory, valid until the next wait_until_frames_copied / configure */
int hwb_decoder_get_frame_device(hwb_decoder *d, uint8_t **device_rgb);
int hwb_decoder_frames_ready(hwb_decoder *d);              /* exact count (decoded_frames_buffered saturates at 8) */
int hwb_decoder_set_chunk_pictures(hwb_decoder *d, int n); /* pictures per GPU batch (cut at IDR boundaries) */
/* Optional hint, to be called before the first feed() of an interval (what DecoderAutomata::feeder knows at
 * decoder_automata.cpp:296-318): the pictures fed until the next flush() are the absolute frames start_frame,
 * start_frame+1, ... in display order, and only the `n` ascending frame numbers in `wanted` will be fetched with
 * get_frame -- the others will be dropped with discard_frame (decoder_automata.cpp:235).  Unrequested non-reference
 * pictures are then not decoded at all.  Without the call everything is decoded, as the reference backends do. */
int hwb_decoder_set_interval_hint(hwb_decoder *d, uint64_t start_frame, const uint64_t *wanted, size_t n);

/* Batch retrieval (python/hwang/decoder.py:30-69 decodes one interval of one video at a time and re-configures in
 * between): with deferred submission the pictures of consecutive intervals -- of one clip or of several clips of equal
 * geometry, re-configured in between -- are collected into ONE GPU batch, so that a single entropy launch sees all
 * their slices.  Frames still pop interval by interval, in display order.  hwb_decoder_submit_pending closes the batch. */
int hwb_decoder_set_defer_submit(hwb_decoder *d, int on);
int hwb_decoder_submit_pending(hwb_decoder *d);

typedef struct hwb_stats {
  uint64_t pictures_decoded, frames_returned, chunks, bitstream_bytes, kernel_launches, h2d_bytes, d2h_bytes, algorithmic_bytes;
  /* Device time from CUDA events on the launching streams.  wall_ms: inputs of a batch resident in HBM -> its last picture
   * kernel done, summed over busy periods.  entropy_ms / picture_ms: per-stage sums over batches; batches overlap, so
   * these may exceed wall_ms. */
  double wall_ms, entropy_ms, picture_ms;
  uint64_t entropy_launches, picture_launches, aux_launches;
} hwb_stats;
int hwb_decoder_get_stats(hwb_decoder *d, hwb_stats *out);

/* pinned host memory for output buffers (get_frame copies straight into pinned buffers) */
void *hwb_alloc_pinned(size_t n);
void hwb_free_pinned(void *p);
/* device memory for DeviceType::GPU output buffers, for callers without a CUDA runtime of their own (ctypes, tests) */
void *hwb_alloc_device(int device_id, size_t n);
void hwb_free_device(int device_id, void *p);
int hwb_copy_device_to_host(int device_id, void *dst, const void *src, size_t n);

/* ---------------------------------------------------------------------------------------------
 * Index: hwang::MP4IndexCreator (hwang/mp4_index_creator.h:23-45) and hwang::VideoIndex
 * (hwang/video_index.h:22-77).
 * ------------------------------------------------------------------------------------------- */
typedef struct hwb_index_creator hwb_index_creator;
typedef struct hwb_video_index hwb_video_index;

hwb_index_creator *hwb_index_creator_create(uint64_t file_size);
void hwb_index_creator_destroy(hwb_index_creator *c);
/* feed(): returns 1 to continue (read *next_size bytes at *next_offset), 0 when done or on error */
int hwb_index_creator_feed(hwb_index_creator *c, const uint8_t *data, size_t size, uint64_t *next_offset, uint64_t *next_size);
int hwb_index_creator_is_done(hwb_index_creator *c);
int hwb_index_creator_is_error(hwb_index_creator *c);
const char *hwb_index_creator_error_message(hwb_index_creator *c);
hwb_video_index *hwb_index_creator_get_video_index(hwb_index_creator *c);

hwb_video_index *hwb_video_index_create(uint32_t timescale, uint64_t duration, uint32_t width, uint32_t height, const char *format,
                                        const uint64_t *sample_offsets, const uint64_t *sample_sizes, size_t num_samples,
                                        const uint64_t *keyframe_indices, size_t num_keyframes, const uint8_t *metadata, size_t metadata_size);
void hwb_video_index_destroy(hwb_video_index *v);
hwb_video_index *hwb_video_index_deserialize(const uint8_t *data, size_t size);     /* VideoIndex::deserialize */
size_t hwb_video_index_serialize(const hwb_video_index *v, uint8_t *out, size_t cap); /* returns bytes needed */
uint32_t hwb_video_index_timescale(const hwb_video_index *v);
uint64_t hwb_video_index_duration(const hwb_video_index *v);
double hwb_video_index_fps(const hwb_video_index *v);
uint32_t hwb_video_index_frame_width(const hwb_video_index *v);
uint32_t hwb_video_index_frame_height(const hwb_video_index *v);
const char *hwb_video_index_format(const hwb_video_index *v);
uint64_t hwb_video_index_frames(const hwb_video_index *v);
const uint64_t *hwb_video_index_sample_offsets(const hwb_video_index *v);
const uint64_t *hwb_video_index_sample_sizes(const hwb_video_index *v);
const uint64_t *hwb_video_index_keyframe_indices(const hwb_video_index *v, size_t *count);
const uint8_t *hwb_video_index_metadata_bytes(const hwb_video_index *v, size_t *size);

/* slice_into_video_intervals (hwang/video_index.h:84-85, video_index.cpp:62-109).
 * Returns the number of intervals; starts/ends/valid_counts (capacity max_intervals) receive
 * [start_keyframe, end_keyframe) and the number of wanted rows of each interval; valid_rows (capacity
 * num_rows) receives the rows grouped per interval.  Returns -1 on malformed input. */
int hwb_slice_into_video_intervals(const hwb_video_index *v, const uint64_t *rows, size_t num_rows, uint64_t *starts, uint64_t *ends,
                                   uint64_t *valid_counts, size_t max_intervals, uint64_t *valid_rows);

/* ---------------------------------------------------------------------------------------------
 * Scheduler: hwang::DecoderAutomata (hwang/decoder_automata.h:33-70).
 * ------------------------------------------------------------------------------------------- */
typedef struct hwb_automata hwb_automata;

typedef struct hwb_encoded_data { /* DecoderAutomata::EncodedData, decoder_automata.h:43-66 */
  const uint8_t *encoded_video;
  size_t encoded_video_size;
  uint32_t width, height;
  uint64_t start_keyframe, end_keyframe;
  const char *format;
  const uint64_t *sample_offsets; /* relative to encoded_video */
  const uint64_t *sample_sizes;
  size_t num_samples;
  const uint64_t *keyframes;
  size_t num_keyframes;
  const uint64_t *valid_frames;
  size_t num_valid_frames;
} hwb_encoded_data;

/* make_instance(DeviceHandle, num_devices, decoder_type); NULL if the backend cannot be built */
hwb_automata *hwb_automata_create(int device_type, int device_id, int num_devices, int decoder_type);
void hwb_automata_destroy(hwb_automata *a);
int hwb_automata_initialize(hwb_automata *a, const hwb_encoded_data *intervals, size_t num_intervals, const uint8_t *extradata, size_t extradata_size);
/* n tightly packed W*H*3 RGB24 frames at buffer + k*W*H*3; buffer may be host (pageable / page-locked) or device memory */
int hwb_automata_get_frames(hwb_automata *a, uint8_t *buffer, int32_t num_frames);
int hwb_automata_set_chunk_pictures(hwb_automata *a, int n); /* pictures per GPU batch of the automaton's decoder */
const char *hwb_automata_last_error(hwb_automata *a);
int hwb_automata_get_stats(hwb_automata *a, hwb_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* HWANG_B200_H_ */
