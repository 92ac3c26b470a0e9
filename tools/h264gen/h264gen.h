// Synthetic H.264 stream generator (C API).  There is no H.264 encoder, ffmpeg/x264 binary or
// sample clip in the build image, so the benchmark/parity inputs ("synthetic libx264-like clips",
// BASELINE.json) are produced by this seeded closed-loop encoder + MP4 muxer.  Test/bench tooling.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hwgen_params {
  int32_t width, height;   // display size (width % 8 == 0, height even); coded size = ceil16
  int32_t frames;          // number of pictures
  int32_t gop;             // IDR period
  int32_t profile;         // 0 constrained baseline (CAVLC), 1 main (CABAC), 2 high (CABAC + 8x8 transform)
  int32_t cabac;           // -1: per profile; 0/1 force
  int32_t bframes;         // consecutive non-reference B pictures between anchors (0 for baseline)
  int32_t num_ref;         // max_num_ref_frames (1..4)
  int32_t qp;              // base QP
  int32_t slices;          // slices per picture (>=1)
  uint32_t seed;
  int32_t weighted;        // 0 none; 1 explicit weighted prediction in P (fade); 2 also implicit bi-pred in B; 3 explicit in P and in B
  int32_t direct_spatial;  // B direct mode: 1 spatial, 0 temporal
  int32_t deblock;         // 0 on; 1 off; 2 on except slice edges; 3 on with random per-slice offsets
  int32_t constrained_intra;
  int32_t ipcm_per_100k;   // I_PCM macroblocks per 100000
  int32_t intra_in_p_pct;  // forced intra macroblocks in P/B pictures, percent
  int32_t cabac_init_idc;  // 0..2, -1 = vary per slice
  int32_t chroma_qp_offset;   // pps chroma_qp_index_offset (second offset = this - 1 in High)
  int32_t scaling_lists;   // high: 1 = custom scaling matrices in the PPS
  int32_t poc_type;        // 0, 1 (expected-delta cycle + delta_pic_order_cnt) or 2 (2 only without B pictures)
  int32_t fragmented;      // mp4: 1 = moof/trun fragments (one per GOP)
  int32_t threads;         // encoder threads (GOP parallel); 0 = hardware concurrency
  int32_t qp_jitter;       // random mb_qp_delta magnitude (0 = constant QP)
  int32_t b_pyramid;       // 1: the middle B picture of every run of B pictures is a reference picture (one B-reference level)
  int32_t rplm_pct;        // percent of P/B pictures whose slices carry ref_pic_list_modification (reorders list 0 / list 1)
  int32_t mmco;            // 1: adaptive reference marking (MMCO 1-4, 6, long-term IDR; MMCO 5 when bframes == 0) on some reference pictures
  int32_t pad_refs;        // 1: P / B slices declare num_ref_idx_active = num_ref (2 for list 1) even while fewer reference pictures exist (the
                           //    first pictures of a GOP), as encoders that rely on the PPS default do; the missing entries stand for the
                           //    initial list's first entry (how libavcodec resolves them) and macroblocks do refer to them
  int32_t mixed_slices;    // 1 (needs slices >= 2): every third slice of a P / B picture is an I slice (intra refresh by slice, as some
                           //    hardware encoders do); such pictures write slice_type 0..2 instead of 5..7
  int32_t header_variant;  // 1: parameter-set ids other than 0 (sps 3, pps 7), pic_init_qp_minus26 = -4, num_ref_idx_default_active = 2 / 2 in
                           //    the PPS (slices override only when they differ), 6-bit frame_num and 5-bit pic_order_cnt_lsb (both wrap inside a
                           //    GOP), as encoders other than this one write their headers
  int32_t direct_4x4;      // 1: direct_8x8_inference_flag = 0 (direct prediction takes the co-located motion per 4x4 block; no 8x8 transform
                           //    in macroblocks with direct parts), as Baseline-era encoders write it
  int32_t reserved[1];
} hwgen_params;

void hwgen_default_params(hwgen_params *p);

// Encode a clip.  On success returns 0 and *out_mp4 (malloc'd, free with hwgen_free).
// If recon_yuv != NULL it receives the encoder's own reconstruction of every picture in DISPLAY order
// (planar 4:2:0, cropped to width x height; frames * width*height*3/2 bytes) -- used to pin the generator
// against libavcodec.
int hwgen_encode(const hwgen_params *p, uint8_t **out_mp4, size_t *out_size, uint8_t *recon_yuv);
void hwgen_free(void *p);
const char *hwgen_last_error(void);

#ifdef __cplusplus
}
#endif
