// CUDA side of the decoder: kernels for sm_100a and the thin C-ABI (csrc/dev/devapi.h) the C++ host
// calls.  No tensor cores (there is no dense contraction on this path): the kernels are
// integer/byte work organised around the two structural constraints of H.264 decoding:
//   * entropy decoding is serial per slice  -> one warp per slice, every slice of every picture of
//     the chunk in flight at once, handed out in decode order by an atomic ticket (which also makes
//     the B-direct co-located-picture wait deadlock-free);
//   * intra prediction and deblocking depend on the left / top / top-right macroblocks
//     -> one warp per macroblock row, rows of a picture form a wavefront synchronised through
//        per-row progress counters in global memory, many pictures (all pictures of one dependency
//        level of the chunk) per launch so the wavefronts of different pictures fill the 148 SMs.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../dev/devapi.h"
#include "../dev/picture.h"
#include "../dev/entropy.h"  // generic (runtime entropy_coding_mode): namespace hwb::ent
// Specialised copies of the slice decoder (the per-macroblock path has to fit the SM's instruction cache):
// MODE 1 / 0 = CABAC / CAVLC only, NO_B = no B-slice support, NO_T8 = no 8x8 transform.
#define HWB_ENT_MODE 0
#define HWB_ENT_NS ent_cavlc
#include "../dev/entropy.h"
#undef HWB_ENT_NS
#undef HWB_ENT_MODE
#define HWB_ENT_MODE 1
#define HWB_ENT_NS ent_cabac
#include "../dev/entropy.h"
#undef HWB_ENT_NS
#define HWB_ENT_NO_T8 1
#define HWB_ENT_NS ent_cabac4
#include "../dev/entropy.h"
#undef HWB_ENT_NS
#define HWB_ENT_NO_B 1
#define HWB_ENT_NS ent_cabac_ip4
#include "../dev/entropy.h"
#undef HWB_ENT_NS
#undef HWB_ENT_NO_T8
#define HWB_ENT_NS ent_cabac_ip
#include "../dev/entropy.h"
#undef HWB_ENT_NS
#undef HWB_ENT_NO_B
#undef HWB_ENT_MODE

using namespace hwb;

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;

__device__ __forceinline__ int warp_ticket(int32_t *ticket) {
  int t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd(ticket, 1);
  return __shfl_sync(0xffffffffu, t, 0);
}

// ------------------------------------------------------------------------------------ entropy
// One lane per warp walks a slice.  All of its state (neighbour caches, CABAC contexts, scratch) sits in shared
// memory: thread-local memory is interleaved across the 32 lanes, so a single active lane would touch one cache
// line per word and thrash L1 (measured: ~4000 cycles per CABAC bin before this change).
// Ticket hand-out.  Intra slices carry several times the bits of inter slices and every one of them is a single
// warp's serial work: the intra slices of a batch set the earliest moment its first pictures can be reconstructed.  A
// warp decodes an intra slice at full speed only while the SM's instruction caches hold the intra path; next to
// inter-slice warps (a different 30 KB of code) it runs 1.5x slower.  So the SMs [intra_sm_base, intra_sm_base + intra_sms) are reserved: their
// warps take the intra tickets (ticket[3]) and nothing else until those are gone, everybody else takes inter tickets
// (ticket[0]).  Invariant kept from the single-queue version: no inter ticket is handed out before every intra ticket
// has been taken (a B slice may wait for its co-located picture's slice, which must therefore be running or done).  If
// the reserved SMs hold no block of this launch (busy device), the others take the intra tickets after a grace period
// -- no warp ever leaves while a ticket is untaken, so nothing can be stranded.
__device__ __forceinline__ int take_ticket(int32_t *counter, int limit) {
  int t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd(counter, 1);
  t = __shfl_sync(0xffffffffu, t, 0);
  return t < limit ? t : -1;
}
__device__ __forceinline__ int peek_counter(const int32_t *counter) {
  int v = 0;
  if ((threadIdx.x & 31) == 0) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
  return __shfl_sync(0xffffffffu, v, 0);
}
__device__ int next_entropy_ticket(const ChunkCtx &c, int32_t *ticket, bool reserved, long long t_start) {
  const int n_intra = c.num_intra_tickets, n_inter = c.num_tickets - c.num_intra_tickets;
  for (;;) {
    if (n_intra > 0 && peek_counter(ticket + 3) < n_intra) {
      if (reserved || c.intra_sms == 0 || clock64() - t_start > 100000) {  // ~50 us of grace for the reserved SMs
        const int t = take_ticket(ticket + 3, n_intra);
        if (t >= 0) return t;
      } else {
        __nanosleep(500);
      }
      continue;
    }
    // (Tried and dropped: the warps of a reserved SM idle instead of going on with inter slices.  The first batch's intra
    // slices finish no earlier than with every launch reserving the same SMs, which is what made the difference: 234 ms
    // instead of 296, profiles/r2_runs/r2al_*.)
    const int t = take_ticket(ticket, n_inter);
    return t >= 0 ? t + n_intra : -1;
  }
}

// (Tried and dropped: a second copy of the decoder for inter slices with every decision out of line, intra slices
// keeping the inlined one -- two copies in one launch took 357 ms per 3000 slices instead of 286: what the SMs miss in
// their own instruction caches they fetch from a cache shared by the GPC, and twice the code thrashes that one.)
// LOCKSTEP (experiment, compiled out by default: -DHWB_LOCKSTEP=1): the four warps of a block start every macroblock
// together (one block-wide barrier per macroblock).  What bounds this stage is the GPC-level instruction cache the SMs
// miss into (90 % of its peak request rate); warps that walk the per-macroblock path at the same time
// fetch each line from there once instead of four times.  Measured on the 3000-slice batch: the SM instruction-cache
// hit rate goes from 77 % to 90 % and the "no instruction" stall from 3.4 to 0.9 cycles per issued instruction -- and
// the barrier costs 3.8, because a macroblock takes anything between 800 and 5000 instructions and every round waits
// for its slowest warp: 303 ms instead of 266 (profiles/r2_runs/r2ak_*).  A warp without a slice (no ticket left) keeps
// arriving at the barrier until every warp of its block is out of work; blocks on the SMs reserved for intra slices run
// free (an intra macroblock takes ten times an inter one's time).  Not for kernels with B slices: a warp waiting for its
// co-located picture's slice (wait_col_mb) could wait for a warp of its own block that stands at the barrier.
#ifndef HWB_LOCKSTEP
#define HWB_LOCKSTEP 0
#endif
#define HWB_ENTROPY_KERNEL(NAME, NS, LOCKSTEP)                                                        \
  __global__ void __launch_bounds__(kThreads) NAME(ChunkCtx cparam, int32_t *ticket) {                \
    __shared__ NS::SliceDec sdec[kWarpsPerBlock];                                                     \
    __shared__ ChunkCtx c;                                                                            \
    if (threadIdx.x == 0) c = cparam;                                                                 \
    __syncthreads();                                                                                  \
    const int w = threadIdx.x >> 5;                                                                   \
    unsigned smid;                                                                                    \
    asm("mov.u32 %0, %%smid;" : "=r"(smid));                                                          \
    const bool reserved = smid - (unsigned)c.intra_sm_base < (unsigned)c.intra_sms;                   \
    const long long t_start = clock64();                                                              \
    if (!(LOCKSTEP) || reserved) { /* the blocks of the SMs reserved for intra slices run free */     \
      for (;;) {                                                                                      \
        const int t = next_entropy_ticket(c, ticket, reserved, t_start);                              \
        if (t < 0) return;                                                                            \
        NS::decode_slice(c, c.entropy_order[t], nullptr, &sdec[w]);                                   \
        __syncwarp();                                                                                 \
      }                                                                                               \
    }                                                                                                 \
    bool active = false, exhausted = false;                                                           \
    for (;;) {                                                                                        \
      if (!active && !exhausted) {                                                                    \
        const int t = next_entropy_ticket(c, ticket, reserved, t_start);                              \
        if (t < 0) exhausted = true;                                                                  \
        else { NS::slice_begin(c, c.entropy_order[t], &sdec[w]); active = true; }                     \
      }                                                                                               \
      if (active) {                                                                                   \
        bool more = NS::slice_step(&sdec[w]);                                                         \
        /* an intra slice that ends up here (nobody on a reserved SM took it) is decoded in one go */ \
        if (more && sdec[w].sd->slice_type == SLICE_I) { while (NS::slice_step(&sdec[w])) {} more = false; } \
        if (!more) { NS::slice_end(&sdec[w]); active = false; }                                       \
      }                                                                                               \
      if (__syncthreads_and(exhausted && !active)) return;                                            \
    }                                                                                                 \
  }
HWB_ENTROPY_KERNEL(entropy_kernel, hwb::ent, 0)                    // pictures of both entropy modes in one chunk
HWB_ENTROPY_KERNEL(entropy_cavlc_kernel, hwb::ent_cavlc, 0)        // every picture of the chunk is CAVLC
HWB_ENTROPY_KERNEL(entropy_cabac_kernel, hwb::ent_cabac, 0)        // every picture of the chunk is CABAC
HWB_ENTROPY_KERNEL(entropy_cabac4_kernel, hwb::ent_cabac4, 0)      // ... and none uses the 8x8 transform
HWB_ENTROPY_KERNEL(entropy_cabac_ip_kernel, hwb::ent_cabac_ip, HWB_LOCKSTEP)  // CABAC, no B slice in the chunk
HWB_ENTROPY_KERNEL(entropy_cabac_ip4_kernel, hwb::ent_cabac_ip4, HWB_LOCKSTEP)  // ... and no 8x8 transform (Main profile)

// ------------------------------------------------------------------------------------ picture kernel
// Reconstruction, deblocking and the RGB24 writeback of every picture of a chunk in ONE launch: see csrc/dev/picture.h
// for the work items, their dependencies and why the hand-out order makes the spin waits deadlock-free.
union PictureScratch {
  ReconScratch recon;
  DeblockScratch deblock;
};

// 80 registers, 6 blocks (24 warps) per SM.  Without static roles every warp is a generalist and chooses the kind of its next item when it
// takes it: the head of the deblocking list if the rows it consumes are already being produced (the reconstruction of
// the row below its band has started), otherwise the head of the reconstruction list, otherwise (reconstruction exhausted)
// the deblocking head whatever its state.  Deblocking first: it completes pictures, which releases the rows of the
// next level waiting for their reference and the frames waiting to be copied out.
// Why this cannot deadlock (see also csrc/dev/picture.h): let X be the earliest unfinished item in the global
// (level, row, picture) order.  Taken items only wait on earlier items, so if X is held it completes.  If X is never
// taken, look at the last pick any warp ever makes (a pick of X itself completes and is followed by another pick).
// That pick took from the other list: either
// the reconstruction head while X heads the deblocking list -- but everything X consumes is earlier than X, hence
// finished, hence started, so X was "ready" and would have been preferred; or a ready deblocking item Z while X heads
// the reconstruction list -- Z's reconstruction rows were handed out before X (same list), so they precede X and are
// finished, and the deblocking row above Z (handed out before Z) needs only such rows and its own predecessor: the whole
// chain completes, Z finishes, and its warp picks again.  Either way that pick was not the last one.
// `split` (HWB_PICTURE_SPLIT, 0 = dynamic) keeps the static assignment for experiments: warps whose global index modulo
// 5 is below it take deblocking items only until those run out.
#ifndef HWB_PICTURE_MIN_BLOCKS
#define HWB_PICTURE_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(kThreads, HWB_PICTURE_MIN_BLOCKS) picture_kernel(const __grid_constant__ ChunkCtx c, int32_t *ticket, int split) {
  __shared__ PictureScratch sm[kWarpsPerBlock];
  PictureScratch *my = &sm[threadIdx.x >> 5];
  // A corrupt or unsupported stream leaves MbInfo / coefficient offsets of the failed slice undefined: the rows check
  // the batch's error flag once the entropy stage of their picture is complete (csrc/dev/picture.h) and touch nothing then.
  // split > 0: static roles by SM: the SMs whose id falls into the first `split` percent only take deblocking items
  // (until they run out), the others only reconstruction items.  Both stages are bound by instruction fetch; an SM that
  // stays inside one of the two code bodies keeps its instruction caches for it, and the two SMs of a TPC must agree
  // (measured on the 3000-picture batch: every warp choosing dynamically 544 ms, roles per warp 484 ms, roles per SM
  // scattered over the chip 452 ms, per TPC 292 ms, one contiguous range of SM ids 276 ms).
  // The host only passes split > 0 when the grid puts blocks on every SM (both roles are then present); small grids
  // use the dynamic choice below, which needs no such assumption.
  bool fixed_deblock = false;
  if (split > 0) {
    unsigned smid, nsmid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsmid));
    const int mode = split / 1000, pct = split % 1000;  // experiments: how the SMs of the two roles are spread over the chip
    if (mode == 1) fixed_deblock = (int)(smid * 100u / nsmid) < pct;               // one contiguous range of SM ids
    else if (mode == 2) fixed_deblock = (int)(((smid >> 1) * 37u) % 100u) < pct;   // whole TPCs (SM pairs), scattered
    else if (mode == 3) fixed_deblock = (int)(smid % 5u) * 20 < pct;               // runs of 2-3 adjacent SMs
    else fixed_deblock = (int)((smid * 37u) % 100u) < pct;                         // single SMs, scattered
  }
  bool recon_left = c.num_recon_items > 0, deblock_left = c.num_deblock_items > 0;
  ProfClock life(c.prof);
  while (recon_left || deblock_left) {
    ProfClock pick(c.prof);
    bool take_deblock;
    if (!recon_left) take_deblock = true;
    else if (!deblock_left) take_deblock = false;
    else if (split > 0) take_deblock = fixed_deblock;
    else {
      // is the deblocking head ready?  (lane 0 looks, everybody follows)
      int ready = 0;
      if ((threadIdx.x & 31) == 0) {
        int32_t t;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(t) : "l"(ticket + 1) : "memory");
        if (t < c.num_deblock_items) {
          const uint32_t it = c.deblock_items[t];
          const int pic = item_pic(it), y = item_row(it);
          // the band needs its own rows and the one below reconstructed: the last of them has started => all have been handed out
          const int32_t *p = c.recon_prog + (size_t)pic * c.mb_h + (y + c.deblock_band < c.mb_h ? y + c.deblock_band : c.mb_h - 1);
          int32_t v;
          asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
          ready = v >= 1;
        }
      }
      take_deblock = __shfl_sync(0xffffffffu, ready, 0) != 0;
    }
    if (take_deblock) {
      const int t = warp_ticket(ticket + 1);
      if (t >= c.num_deblock_items) { deblock_left = false; continue; }
      const uint32_t it = c.deblock_items[t];
      pick.mark(PROF_PICK);
      deblock_band(c, item_pic(it), item_row(it), &my->deblock);
    } else {
      const int t = warp_ticket(ticket);
      if (t >= c.num_recon_items) { recon_left = false; continue; }
      const uint32_t it = c.recon_items[t];
      pick.mark(PROF_PICK);
      recon_row(c, item_pic(it), item_row(it), &my->recon);
    }
    __syncwarp();
  }
  life.mark(PROF_LIFETIME);
}

// ------------------------------------------------------------------------------------ output
__global__ void __launch_bounds__(256) rgb24_kernel(ChunkCtx c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst) {
  const int nx = (w + 15) >> 4;
  const int total = nx * h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    rgb24_item(c, frame, crop_x, crop_y, w, h, dst, i % nx, i / nx);
}

__global__ void __launch_bounds__(256) yuv_kernel(ChunkCtx c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst) {
  const int total = w * h + 2 * (w / 2) * (h / 2);
  const uint8_t *Y = frame_y(c, frame), *U = frame_cb(c, frame), *V = frame_cr(c, frame);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < w * h) dst[i] = Y[(size_t)(crop_y + i / w) * c.wc + crop_x + i % w];
    else {
      int k = i - w * h, cw = w / 2, ch = h / 2;
      const uint8_t *P = k < cw * ch ? U : V;
      if (k >= cw * ch) k -= cw * ch;
      dst[i] = P[(size_t)(crop_y / 2 + k / cw) * (c.wc / 2) + crop_x / 2 + k % cw];
    }
  }
}

}  // namespace

// ========================================================================================= C-ABI
struct hwb_dev {
  int device = 0;
  int sms = 148;
  int smem_per_sm = 228 * 1024, smem_optin = 227 * 1024;
  int entropy_bpsm = 3, picture_bpsm = 6;  // resident blocks per SM (hwb_dev_set_occupancy)
  cudaStream_t streams[HWB_NUM_STREAMS];
  std::string err;
  std::atomic<uint64_t> launches{0};
};
struct hwb_event { cudaEvent_t ev; };

#define HWB_CUDA(d, expr)                                                                             \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess) { (d)->err = std::string(#expr) + ": " + cudaGetErrorString(e__); return 1; } \
  } while (0)

extern "C" {

int hwb_dev_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int hwb_dev_open(int device, hwb_dev **out) {
  *out = nullptr;
  int n = hwb_dev_count();
  if (device < 0 || device >= n) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  hwb_dev *d = new hwb_dev();
  d->device = device;
  cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&d->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
  cudaDeviceGetAttribute(&d->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  // Priorities, highest first: copies to the host, then the picture kernel (a chunk finishes -- and its frames start
  // travelling to the host -- while later chunks are still being entropy-decoded), then entropy decoding
  int least = 0, greatest = 0;
  cudaDeviceGetStreamPriorityRange(&least, &greatest);
  for (int i = 0; i < HWB_NUM_STREAMS; ++i) {
    int prio = (i == HWB_STREAM_COPY || i == HWB_STREAM_AUX) ? greatest : (i == HWB_STREAM_PICTURE ? greatest + 1 : greatest + 2);
    if (prio > least) prio = least;
    if (cudaStreamCreateWithPriority(&d->streams[i], cudaStreamNonBlocking, prio) != cudaSuccess) { delete d; return 1; }
  }
  *out = d;
  return 0;
}
void hwb_dev_close(hwb_dev *d) {
  if (!d) return;
  cudaSetDevice(d->device);
  for (int i = 0; i < HWB_NUM_STREAMS; ++i) cudaStreamDestroy(d->streams[i]);
  delete d;
}
const char *hwb_dev_error(hwb_dev *d) { return d->err.c_str(); }

void *hwb_dev_malloc(hwb_dev *d, size_t n) {
  cudaSetDevice(d->device);
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, n ? n : 1);
  if (e != cudaSuccess) { d->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); cudaGetLastError(); return nullptr; }
  return p;
}
void hwb_dev_free(hwb_dev *d, void *p) { cudaSetDevice(d->device); cudaFree(p); }
void *hwb_dev_malloc_host(hwb_dev *d, size_t n) {
  cudaSetDevice(d->device);
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, n ? n : 1, cudaHostAllocPortable | cudaHostAllocMapped);  // mapped: the picture kernel writes completion flags into it
  if (e != cudaSuccess) { d->err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); cudaGetLastError(); return nullptr; }
  return p;
}
void hwb_dev_free_host(hwb_dev *d, void *p) { cudaSetDevice(d->device); cudaFreeHost(p); }
int hwb_dev_pointer_kind(hwb_dev *d, const void *p) {
  cudaPointerAttributes a;
  cudaSetDevice(d->device);
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  return a.type == cudaMemoryTypeHost ? 1 : ((a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? 2 : 0);
}
int hwb_dev_mem_info(hwb_dev *d, size_t *free_bytes, size_t *total_bytes) {
  cudaSetDevice(d->device);
  HWB_CUDA(d, cudaMemGetInfo(free_bytes, total_bytes));
  return 0;
}

int hwb_dev_h2d(hwb_dev *d, int s, void *dst, const void *src, size_t n) {
  cudaSetDevice(d->device);
  HWB_CUDA(d, cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, d->streams[s]));
  return 0;
}
int hwb_dev_d2h(hwb_dev *d, int s, void *dst, const void *src, size_t n) {
  cudaSetDevice(d->device);
  HWB_CUDA(d, cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, d->streams[s]));
  return 0;
}
int hwb_dev_d2d(hwb_dev *d, int s, void *dst, const void *src, size_t n) {
  cudaSetDevice(d->device);
  HWB_CUDA(d, cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, d->streams[s]));  // the destination may live on a peer GPU
  return 0;
}
int hwb_dev_memset(hwb_dev *d, int s, void *dst, int v, size_t n) {
  cudaSetDevice(d->device);
  HWB_CUDA(d, cudaMemsetAsync(dst, v, n, d->streams[s]));
  return 0;
}

static int grid_for(hwb_dev *d, int work_warps, int blocks_per_sm) {
  int blocks = (work_warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
  int cap = d->sms * blocks_per_sm;  // a multiple of the SM count; warps loop over tickets
  return blocks < cap ? (blocks < 1 ? 1 : blocks) : cap;
}

// Blocks per SM of the two kernels.  When the picture kernel of a batch is launched together with its entropy kernel
// (it waits for the entropy stage picture by picture), both grids must fit an SM at the same time whatever order the
// hardware places their blocks in -- a picture kernel that filled the machine first would wait for entropy blocks that
// can never start.  2 entropy blocks (at most 96 registers x 128 threads) + 4 picture blocks (80 x 128) = 65536 of the 65536
// registers of an SM, 19 + 75 KB of shared memory, 24 of 64 warps.
void hwb_dev_set_occupancy(hwb_dev *d, int entropy_blocks_per_sm, int picture_blocks_per_sm) {
  d->entropy_bpsm = entropy_blocks_per_sm; d->picture_bpsm = picture_blocks_per_sm;
}

int hwb_dev_entropy(hwb_dev *d, int s, const ChunkCtx *c, int32_t *ticket, int mode) {
  cudaSetDevice(d->device);
  // Resident warps per SM are capped at 12 (3 blocks of 4 warps): the slice decoder is branchy code larger than the
  // SM's instruction cache, every extra warp wandering through a different part of it costs all of them fetch
  // misses, and those misses saturate the GPC-level instruction cache (measured at 20 warps per SM: 67% of the stall
  // cycles are "no instruction", gcc instruction requests at 67% of peak).  Sweep on the 3000-slice benchmark chunk:
  // 1 block/SM 561 ms, 2: 411, 3: 394, 4: 402, 5: 409, 8: 430.  HWB_ENTROPY_BLOCKS_PER_SM overrides.
  static int bpsm_env = [] { const char *e = getenv("HWB_ENTROPY_BLOCKS_PER_SM"); int v = e ? atoi(e) : 0; return v; }();
  const int bpsm = bpsm_env > 0 ? bpsm_env : d->entropy_bpsm;
  const int grid = grid_for(d, c->num_tickets, bpsm);
  // The cap has to hold across launches too: the entropy kernels of all batches of a request are in flight together, and
  // the block scheduler would pack seven of their blocks (28 warps) onto an SM.  Unused dynamic shared memory makes a
  // block as large as 1/bpsm of what entropy blocks may take of an SM (all but 54 KB, which stay for two blocks of the
  // picture kernel), so that at most bpsm of them are resident whichever launches they come from -- earlier batches
  // first, which is the order their frames are due in.  HWB_ENTROPY_SMEM_CAP=0 turns it off.
  static const bool cap = [] { const char *e = getenv("HWB_ENTROPY_SMEM_CAP"); return !e || atoi(e) != 0; }();
  size_t pad = 0;
  auto launch = [&](auto kernel) {
    if (cap) {
      cudaFuncAttributes fa;
      if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) {
        const size_t per_block = ((size_t)d->smem_per_sm - 54 * 1024) / (size_t)bpsm - 1024;  // 1 KB per block is the system's
        if (per_block > fa.sharedSizeBytes && per_block <= (size_t)d->smem_optin) {
          pad = per_block - fa.sharedSizeBytes;
          cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
        }
      }
    }
    kernel<<<grid, kThreads, pad, d->streams[s]>>>(*c, ticket);
  };
  if (mode == 4) launch(entropy_cabac_ip4_kernel);
  else if (mode == 5) launch(entropy_cabac4_kernel);
  else if (mode == 3) launch(entropy_cabac_ip_kernel);
  else if (mode == 1) launch(entropy_cabac_kernel);
  else if (mode == 0) launch(entropy_cavlc_kernel);
  else launch(entropy_kernel);
  HWB_CUDA(d, cudaGetLastError());
  d->launches++;
  return 0;
}
int hwb_dev_picture(hwb_dev *d, int s, const ChunkCtx *c, int32_t *ticket) {
  cudaSetDevice(d->device);
  static int bpsm_env = [] { const char *e = getenv("HWB_PICTURE_BLOCKS_PER_SM"); int v = e ? atoi(e) : 0; return v; }();
  const int bpsm = bpsm_env > 0 ? bpsm_env : d->picture_bpsm;
  // percent of the SMs that deblock (0 = every warp chooses dynamically, see picture_kernel)
  static int split_pct = [] { const char *e = getenv("HWB_PICTURE_SPLIT"); int v = e ? atoi(e) : -1; return v >= 0 && v <= 90 ? v : 35; }();
  const int items = c->num_recon_items + c->num_deblock_items;
  if (items == 0) return 0;
  const int grid = grid_for(d, items, bpsm);
  static int rolemap = [] { const char *e = getenv("HWB_PICTURE_ROLEMAP"); int v = e ? atoi(e) : 1; return v >= 0 && v <= 3 ? v : 1; }();
  const int split = (grid >= 2 * d->sms && split_pct > 0) ? split_pct + 1000 * rolemap : 0;  // static roles need blocks on every SM
  picture_kernel<<<grid, kThreads, 0, d->streams[s]>>>(*c, ticket, split);
  HWB_CUDA(d, cudaGetLastError());
  d->launches++;
  return 0;
}
int hwb_dev_rgb24(hwb_dev *d, int s, const ChunkCtx *c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst) {
  cudaSetDevice(d->device);
  int items = ((w + 15) / 16) * h;
  int blocks = (items + 255) / 256;
  if (blocks > d->sms * 8) blocks = d->sms * 8;
  rgb24_kernel<<<blocks, 256, 0, d->streams[s]>>>(*c, frame, crop_x, crop_y, w, h, dst);
  HWB_CUDA(d, cudaGetLastError());
  d->launches++;
  return 0;
}
int hwb_dev_yuv(hwb_dev *d, int s, const ChunkCtx *c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst) {
  cudaSetDevice(d->device);
  yuv_kernel<<<d->sms * 4, 256, 0, d->streams[s]>>>(*c, frame, crop_x, crop_y, w, h, dst);
  HWB_CUDA(d, cudaGetLastError());
  d->launches++;
  return 0;
}

hwb_event *hwb_dev_event_create(hwb_dev *d) {
  cudaSetDevice(d->device);
  hwb_event *e = new hwb_event();
  if (cudaEventCreate(&e->ev) != cudaSuccess) { delete e; return nullptr; }
  return e;
}
void hwb_dev_event_destroy(hwb_dev *d, hwb_event *e) { if (!e) return; cudaSetDevice(d->device); cudaEventDestroy(e->ev); delete e; }
int hwb_dev_event_record(hwb_dev *d, hwb_event *e, int s) { cudaSetDevice(d->device); HWB_CUDA(d, cudaEventRecord(e->ev, d->streams[s])); return 0; }
int hwb_dev_event_done(hwb_dev *d, hwb_event *e) {
  cudaSetDevice(d->device);
  cudaError_t r = cudaEventQuery(e->ev);
  if (r == cudaSuccess) return 1;
  if (r == cudaErrorNotReady) { cudaGetLastError(); return 0; }
  d->err = std::string("cudaEventQuery: ") + cudaGetErrorString(r);
  return -1;
}
int hwb_dev_event_sync(hwb_dev *d, hwb_event *e) { cudaSetDevice(d->device); HWB_CUDA(d, cudaEventSynchronize(e->ev)); return 0; }
int hwb_dev_stream_wait(hwb_dev *d, int s, hwb_event *e) { cudaSetDevice(d->device); HWB_CUDA(d, cudaStreamWaitEvent(d->streams[s], e->ev, 0)); return 0; }
int hwb_dev_stream_sync(hwb_dev *d, int s) { cudaSetDevice(d->device); HWB_CUDA(d, cudaStreamSynchronize(d->streams[s])); return 0; }
int hwb_dev_event_elapsed(hwb_dev *d, hwb_event *a, hwb_event *b, float *ms) { cudaSetDevice(d->device); HWB_CUDA(d, cudaEventElapsedTime(ms, a->ev, b->ev)); return 0; }
uint64_t hwb_dev_launch_count(hwb_dev *d) { return d->launches.load(); }
}
