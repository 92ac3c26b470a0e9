// Row work items of the picture kernel: reconstruction, deblocking and the RGB24 writeback of every picture of a
// chunk run in ONE launch (csrc/cuda/kernels.cu picture_kernel; the unit tests' host emulation runs the same
// functions serially).  Work items are macroblock rows: "reconstruct row y of picture p" and "deblock row y of
// picture p (and convert what became final to RGB24)".  A warp walks its row left to right; everything an item needs
// is tracked by per-row progress counters in global memory:
//   reconstruct (p,y) at x : intra macroblocks need row y-1 reconstructed up to x+1 (unfiltered top / top-right
//                            samples); inter rows need the rows of their reference pictures that the entropy stage
//                            found them to reach (ChunkCtx::mv_reach) completely deblocked;
//   deblock (p,y) at x     : (rows are handed out in bands of DEBLOCK_BAND, see deblock_band) (x,y) reconstructed; row y+1 reconstructed up to x+1 (its intra macroblocks read the
//                            UNFILTERED samples of row y, so they must be done before the filter runs in place);
//                            row y-1 deblocked up to x+1 (H.264 filters in raster order).
// Both item lists are ordered by (dependency level of the picture, row, picture), so an item only ever waits on
// items that come before it in that global order; a warp decides which list to take from when it takes (kernels.cu
// picture_kernel, with the argument why the spin waits cannot deadlock whatever the number of resident warps).
// Pictures of consecutive levels overlap as a diagonal wavefront (a P picture follows its reference at a distance of
// a few macroblocks) instead of one launch pair per level -- a GOP of 250 pictures used to cost 500 launches with a
// latency floor each.
//
// Replaces, on the GPU, the per-macroblock loop of libavcodec's h264 decoder behind avcodec_send_packet and the
// sws_scale call of SoftwareVideoDecoder::get_frame (hwang/impls/software/software_video_decoder.cpp:292-325,349-402).
#pragma once
#include "deblock.h"
#include "recon.h"
#include "rgb.h"

namespace hwb {

// Where the warps of the picture kernel spend their cycles (lane 0's clock64 deltas, summed over all warps), recorded
// only when the host asked for it (ChunkCtx::prof): the kernel is bound by dependencies and instruction issue, not by
// memory, so this is the profile that explains it.
enum { PROF_RECON_MB = 0, PROF_DEBLOCK_MB, PROF_RGB, PROF_WAIT_REF, PROF_WAIT_INTRA, PROF_WAIT_RECON, PROF_WAIT_DEBLOCK_ABOVE, PROF_PICK, PROF_LIFETIME,
       PROF_POLLS, PROF_COUNTERS = 16 };
// Compiled in only with -DHWB_PROFILE_BUILD=1 (python -m hwang_b200.build --variant ... -DHWB_PROFILE_BUILD=1): the clock
// and the counter pointer are live across the whole row loop, and the picture kernel is short of registers (80).
#ifndef HWB_PROFILE_BUILD
#define HWB_PROFILE_BUILD 0
#endif
#if HWB_DEVICE_BUILD && HWB_PROFILE_BUILD
struct ProfClock {
  unsigned long long *out;
  long long t;
  __device__ __forceinline__ explicit ProfClock(unsigned long long *p) : out(p), t(0) { if (out) t = clock64(); }
  // charge the cycles since the last mark to `what`
  __device__ __forceinline__ void mark(int what) { if (out) t = charge(out, what, t); }
  // out of line: the row loops stay small for the instruction caches; static and by value, so that the object itself
  // never has its address taken (as a member function it kept `out` and `t` in thread-local memory)
  static __device__ __noinline__ long long charge(unsigned long long *out, int what, long long t) {
    const long long now = clock64();
    if ((threadIdx.x & 31) == 0) atomicAdd(out + what, (unsigned long long)(now - t));
    return now;
  }
};
#else
#if HWB_DEVICE_BUILD
struct ProfClock { __device__ __forceinline__ explicit ProfClock(unsigned long long *) {} __device__ __forceinline__ void mark(int) {} };
#else
struct ProfClock { explicit ProfClock(unsigned long long *) {} void mark(int) {} };
#endif
#endif

struct Progress {  // a counter watched by this warp, with the last value seen (counters only grow)
  const int32_t *p;
  int seen;
};

#if HWB_DEVICE_BUILD
// The poll takes the counter's address and returns what it saw: a Progress never has its address taken and lives in
// registers (passed by reference it sat in the caller's stack frame, one thread-local load per wait_progress check:
// 218 -> 201 ms per 3000 pictures together with the other thread-local-memory removals, profiles/r2_runs/r2aq_ab.txt).
__device__ __noinline__ int poll_progress(const int32_t *p, int need) {
  int32_t v = 0;
  if ((threadIdx.x & 31) == 0) {
    for (;;) {  // relaxed GPU-scope poll; the data it guards is read with ld.global.cg (L2) by the callers
      asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
      if (v >= need) break;
      __nanosleep(64);
    }
  }
  return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ void wait_progress(Progress &g, int need) { if (g.seen < need) g.seen = poll_progress(g.p, need); }
// Release store at GPU scope: orders the warp's earlier writes (made visible to lane 0 by __syncwarp) before the
// flag.  Unlike __threadfence() + volatile store (MEMBAR.SC + CCTL.IVALL + a system-scope store) it does not
// invalidate the SM's L1 on every macroblock, which the table and MbInfo loads of the other warps live in.
__device__ __forceinline__ void publish_progress(int32_t *p, int v) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#else
// Host emulation: items run one after the other in an order the emulation has checked to be ready, so a wait that is
// not already satisfied means the host scheduler produced a wrong order.
static thread_local int g_unsatisfied_waits = 0;
static inline void wait_progress(Progress &g, int need) {
  if (g.seen >= need) return;
  g.seen = *g.p;
  if (g.seen < need) g_unsatisfied_waits++;
}
static inline void publish_progress(int32_t *p, int v) { *p = v; }
#endif

enum { ITEM_ROW_BITS = 11, ITEM_PIC_SHIFT = 12 };
HWB_HD uint32_t make_item(int pic, int row, int kind) { return ((uint32_t)pic << ITEM_PIC_SHIFT) | ((uint32_t)row << 1) | (uint32_t)kind; }
HWB_HD int item_pic(uint32_t it) { return (int)(it >> ITEM_PIC_SHIFT); }
HWB_HD int item_row(uint32_t it) { return (int)((it >> 1) & ((1u << ITEM_ROW_BITS) - 1)); }

// Macroblock row of a reference picture that must be completely deblocked before row y of `pic` may be predicted
// (-1: the row has no inter macroblock).  Rows < reach are final once row `reach` is deblocked (its top-edge filter is
// the last thing that modifies row reach-1).
HWB_HD int reference_row_needed(const ChunkCtx &c, int pic, int y) {
  // ld.global.cg: several pictures share a cache line of this array, and the entropy stage may still be writing the others'
  const int reach = (int)ld_u32_cg((const uint32_t *)c.mv_reach + (size_t)pic * c.mb_h + y);
  if (reach <= 0) return -1;
  return reach < c.mb_h ? reach : c.mb_h - 1;
}

enum { MAX_TRACKED_DEPS = 4 };

HWB_FN void recon_row(const ChunkCtx &c, int pic, int y, ReconScratch *my) {
  const PicDesc &pd = c.pics[pic];
  const MbInfo *mbs = pic_mbinfo(c, pd.frame);
  int32_t *prog = c.recon_prog + (size_t)pic * c.mb_h;
  const bool has_inter = pd.has_inter != 0;
  // Inter prediction: the reference rows this row reaches must be deblocked as far as the macroblock at hand reaches
  // to the right (frame index == picture index inside a chunk, so a reference's counters are found by its frame
  // index).  Waiting per macroblock instead of for whole rows lets a picture follow its reference at a distance of a
  // few macroblocks: the pictures of a GOP form one long diagonal wavefront.  Pictures with more references than
  // are tracked here wait for complete rows.
  // The picture kernel may be launched together with the batch's entropy kernel: a picture is touched only once all
  // its slices are entropy-decoded (the per-picture arrays are padded so that pictures never share a cache line, which
  // makes the plain loads of macroblock records, motion and coefficients safe).  A corrupt slice raises the batch's
  // error flag before it reports completion: nothing of the batch is dereferenced from then on, the rows only
  // report progress so that nobody waits for them.
  for (int i = 0; i < pd.num_slices; ++i) {
    Progress ent = {c.entropy_prog + pd.first_slice + i, -1};
    wait_progress(ent, c.nmb);
  }
  if (ld_u32_cg((const uint32_t *)c.error_flag) != 0) { publish_progress(prog + y, c.mb_w); return; }
  Progress dep[MAX_TRACKED_DEPS];
  int ndep = 0, reach_x = 0;
  if (has_inter) {
    const int row = reference_row_needed(c, pic, y);
    if (row >= 0) {
      reach_x = (int)ld_u32_cg((const uint32_t *)c.mv_reach_x + (size_t)pic * c.mb_h + y);
      if (pd.num_dep <= MAX_TRACKED_DEPS) {
        ndep = pd.num_dep;
        for (int i = 0; i < MAX_TRACKED_DEPS; ++i)
          if (i < ndep) { dep[i].p = c.dbl_prog + (size_t)pd.dep[i] * c.mb_h + row; dep[i].seen = -1; }
      } else {
        for (int i = 0; i < pd.num_dep; ++i) {
          Progress ref = {c.dbl_prog + (size_t)pd.dep[i] * c.mb_h + row, -1};
          wait_progress(ref, c.mb_w);
        }
      }
    }
  }
  Progress above = {prog + y - 1, y > 0 ? -1 : (1 << 30)};
  ProfClock pc(c.prof);
  for (int x = 0; x < c.mb_w; ++x) {
    if (mbs[y * c.mb_w + x].mbtype != MB_INTER) {
      // intra macroblocks read the unfiltered row above up to the top-right neighbour
      wait_progress(above, x + 2 < c.mb_w ? x + 2 : c.mb_w);
      pc.mark(PROF_WAIT_INTRA);
    } else if (ndep) {
      const int need = x + reach_x < c.mb_w ? x + reach_x : c.mb_w;
#pragma unroll
      for (int i = 0; i < MAX_TRACKED_DEPS; ++i) if (i < ndep) wait_progress(dep[i], need);
      pc.mark(PROF_WAIT_REF);
    }
#if !HWB_DEVICE_BUILD
    {  // what the waits above guarantee to be final in the reference pictures (see McWindow)
      const int reach = has_inter ? c.mv_reach[(size_t)pic * c.mb_h + y] : 0;
      const int X = x + reach_x < c.mb_w ? x + reach_x : c.mb_w;
      g_mc_window.max_row = reach * 16 - 1;
      g_mc_window.max_col = X >= c.mb_w ? (1 << 30) : 16 * X - 4;
    }
#endif
    recon_mb(c, pic, x, y, my);
    // consumers: intra macroblocks of the row below and the deblocking pass; pictures with inter slices have few
    // intra macroblocks, so the fence + flag store is amortised over 4 macroblocks there
    if (!has_inter || (x & 3) == 3 || x == c.mb_w - 1) publish_progress(prog + y, x + 1);
    pc.mark(PROF_RECON_MB);
  }
}

// Deblocking work item: a BAND of DEBLOCK_BAND consecutive rows, filtered by one warp in a skewed order (column x of
// the band's first row, x-2 of the second, x-4 of the third ...: exactly the lag H.264's raster-order filter needs).
// A deblocking row on its own can only advance at the pace of the reconstruction it follows (its rows are
// reconstructed side by side, all at the same column), so a warp per row spends most of its life waiting and the
// resident warps -- not the work -- limit how many pictures of a long GOP chain are in flight; a band does several
// rows' worth of filtering per step of that pace.
enum { DEBLOCK_BAND = 4 };

HWB_FN void deblock_band(const ChunkCtx &c, int pic, int y0, DeblockScratch *my) {
  const PicDesc &pd = c.pics[pic];
  int32_t *prog = c.dbl_prog + (size_t)pic * c.mb_h;
  const int32_t *rprog = c.recon_prog + (size_t)pic * c.mb_h;
  uint8_t *rgb = pd.rgb_slot >= 0 ? c.rgb + (uint64_t)pd.rgb_slot * c.rgb_stride : nullptr;
  const int band = c.deblock_band >= 1 && c.deblock_band <= DEBLOCK_BAND ? c.deblock_band : DEBLOCK_BAND;
  const int nb = y0 + band <= c.mb_h ? band : c.mb_h - y0;
  Progress rp[DEBLOCK_BAND + 1];  // reconstruction of rows y0 .. y0+nb
#pragma unroll
  for (int k = 0; k <= DEBLOCK_BAND; ++k) { rp[k].p = rprog + y0 + k; rp[k].seen = (k <= nb && y0 + k < c.mb_h) ? -1 : (1 << 30); }
  Progress above = {prog + y0 - 1, y0 > 0 ? -1 : (1 << 30)};
  ProfClock pc(c.prof);
  wait_progress(rp[0], 1);  // the row's reconstruction has started, so the entropy stage of the picture is complete
  const bool failed = ld_u32_cg((const uint32_t *)c.error_flag) != 0;
  for (int x = 0; x < c.mb_w + 2 * (nb - 1) && !failed; ++x) {
#pragma unroll
    for (int k = 0; k < DEBLOCK_BAND; ++k) {
      const int xx = x - 2 * k, y = y0 + k;
      if (k >= nb || xx < 0 || xx >= c.mb_w) continue;
      const int lag = xx + 2 < c.mb_w ? xx + 2 : c.mb_w;
      wait_progress(rp[k], xx + 1);
      wait_progress(rp[k + 1], lag);
      pc.mark(PROF_WAIT_RECON);
      if (k == 0) { wait_progress(above, lag); pc.mark(PROF_WAIT_DEBLOCK_ABOVE); }  // rows inside the band: program order
      deblock_mb(c, pic, xx, y, my);
      publish_progress(prog + y, xx + 1);
      pc.mark(PROF_DEBLOCK_MB);
      if (rgb) {
        // RGB24 writeback fused into this pass: with (xx,y) filtered, macroblock (xx,y-1) is final (its right edge was
        // filtered by (xx+1,y-1), which the wait / the skew covers; its bottom edge just now), and so is (xx-1,y) on
        // the last row.  Two macroblocks per step: 32 lanes x 16 pixels, three 16-byte stores each.
        const bool flush = xx == c.mb_w - 1, last_row = y == c.mb_h - 1;
        if (y > 0 && ((xx & 1) || flush)) rgb24_macroblocks(c, pd.frame, rgb, xx & ~1, (xx & 1) ? 2 : 1, y - 1);
        if (last_row) {
          if (xx >= 2 && !(xx & 1)) rgb24_macroblocks(c, pd.frame, rgb, xx - 2, 2, y);
          if (flush) rgb24_macroblocks(c, pd.frame, rgb, (xx & 1) ? xx - 1 : xx, (xx & 1) ? 2 : 1, y);
        }
        pc.mark(PROF_RGB);
      }
    }
  }
  // picture complete?  Every row's samples and RGB24 were written before its count (fence), so whoever counts the last
  // rows may tell the host (system-scope fence: the copy engine reads what the host was promised).
  if (failed) {  // the host learns about the error when the batch ends
    for (int k = 0; k < nb; ++k) publish_progress(prog + y0 + k, c.mb_w);
    return;
  }
#if HWB_DEVICE_BUILD
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    __threadfence();
    if (atomicAdd(c.rows_done + pic, nb) == c.mb_h - nb) {
      __threadfence_system();
      *((volatile int32_t *)c.pic_done + pic) = 1;
    }
  }
#else
  if ((c.rows_done[pic] += nb) == c.mb_h) c.pic_done[pic] = 1;
#endif
}

}  // namespace hwb
