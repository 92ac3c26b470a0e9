// The thin C-ABI between the C++ host side (MP4 demux, NAL / slice-header parsing, sparse-frame
// scheduling) and the CUDA side (BASELINE.json north_star, "Host side").  Plain pointers, sizes and
// int status codes only.  Implemented by csrc/cuda/kernels.cu for the product library; the unit
// tests link a host emulation of the same entry points (tests/emu) to exercise the identical
// device code on machines without a GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "ir.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hwb_dev hwb_dev;  // one per (host decoder instance, GPU)

// Streams.  The entropy kernel of a chunk is latency-bound (one warp per slice, an intra slice of a 1080p picture
// takes a quarter of a second on its own), so the entropy kernels of many chunks must be in flight at once: chunks
// rotate over HWB_NUM_ENTROPY_STREAMS streams (inputs are uploaded on the same stream).  The picture kernels of
// all chunks share one stream (chunks complete in submission order), copies to the host have their own, and the
// auxiliary stream serves the test-only planar output and the conversion of frames nobody announced.
enum { HWB_STREAM_ENTROPY0 = 0, HWB_NUM_ENTROPY_STREAMS = 16, HWB_STREAM_PICTURE = 16, HWB_STREAM_COPY = 17, HWB_STREAM_AUX = 18, HWB_NUM_STREAMS = 19 };

int hwb_dev_count(void);
int hwb_dev_open(int device, hwb_dev **out);
void hwb_dev_close(hwb_dev *d);
const char *hwb_dev_error(hwb_dev *d);

void *hwb_dev_malloc(hwb_dev *d, size_t n);
void hwb_dev_free(hwb_dev *d, void *p);
void *hwb_dev_malloc_host(hwb_dev *d, size_t n);  // pinned
void hwb_dev_free_host(hwb_dev *d, void *p);

int hwb_dev_h2d(hwb_dev *d, int stream, void *dst, const void *src, size_t n);
int hwb_dev_d2h(hwb_dev *d, int stream, void *dst, const void *src, size_t n);
int hwb_dev_d2d(hwb_dev *d, int stream, void *dst, const void *src, size_t n);
// 1 = page-locked host memory, 2 = device memory (of any device), 0 = pageable host memory
int hwb_dev_pointer_kind(hwb_dev *d, const void *p);
// bytes of device memory currently free / total
int hwb_dev_mem_info(hwb_dev *d, size_t *free_bytes, size_t *total_bytes);
int hwb_dev_memset(hwb_dev *d, int stream, void *dst, int value, size_t n);

// Decode stages.  `c` is a host copy of the chunk context (its pointers are device pointers).
// ticket: device int32[1] (entropy) / int32[2] (picture kernel) zeroed by the caller, used for ordered work distribution.
// mode: 1 = every picture of the chunk is CABAC, 3 = CABAC and no B slice, 4 = 3 and no 8x8 transform, 5 = CABAC with B slices and no 8x8
// transform, 0 = every picture is CAVLC, -1 = mixed (generic kernel)
int hwb_dev_entropy(hwb_dev *d, int stream, const hwb::ChunkCtx *c, int32_t *ticket, int mode);
// resident blocks per SM of the entropy / picture kernels launched from now on (see kernels.cu)
void hwb_dev_set_occupancy(hwb_dev *d, int entropy_blocks_per_sm, int picture_blocks_per_sm);
// Reconstruction + deblocking + RGB24 writeback of every picture of the chunk (work lists in the context).
int hwb_dev_picture(hwb_dev *d, int stream, const hwb::ChunkCtx *c, int32_t *ticket);
// Cropped planar 4:2:0 -> packed RGB24 (reference: sws_scale in SoftwareVideoDecoder::get_frame,
// software_video_decoder.cpp:292-325; arithmetic of SURVEY.md section 8a row R).
int hwb_dev_rgb24(hwb_dev *d, int stream, const hwb::ChunkCtx *c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst_dev);
// Cropped planes -> tightly packed planar I420 (parity tests compare these with libavcodec's output).
int hwb_dev_yuv(hwb_dev *d, int stream, const hwb::ChunkCtx *c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst_dev);

typedef struct hwb_event hwb_event;
hwb_event *hwb_dev_event_create(hwb_dev *d);
void hwb_dev_event_destroy(hwb_dev *d, hwb_event *e);
int hwb_dev_event_record(hwb_dev *d, hwb_event *e, int stream);
int hwb_dev_event_done(hwb_dev *d, hwb_event *e);  // 1 done, 0 pending, <0 error
int hwb_dev_event_sync(hwb_dev *d, hwb_event *e);
int hwb_dev_stream_wait(hwb_dev *d, int stream, hwb_event *e);
int hwb_dev_stream_sync(hwb_dev *d, int stream);
// milliseconds between two recorded events (device timeline)
int hwb_dev_event_elapsed(hwb_dev *d, hwb_event *a, hwb_event *b, float *ms);
// number of kernels launched through this handle so far
uint64_t hwb_dev_launch_count(hwb_dev *d);

#ifdef __cplusplus
}
#endif
