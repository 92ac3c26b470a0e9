// TEST INFRASTRUCTURE ONLY.  Host emulation of the CUDA-side C-ABI (hwang_b200/csrc/dev/devapi.h):
// the very same device functions (entropy.h / recon.h / deblock.h / rgb.h, compiled with a plain C++
// compiler, lanes run as loops) executed serially in ticket order.  It exists so that the CPU-only
// test tier can exercise the decode core and the host scheduler bit-exactly against the oracle on
// machines without a GPU.  It is built into tests/emu/libhwb_emu.so, never into the product library
// (hwang_b200/libhwang_b200.so links csrc/cuda/kernels.cu instead and fails loudly without a GPU).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

#include "../../hwang_b200/csrc/dev/devapi.h"
#include "../../hwang_b200/csrc/dev/entropy.h"
#include "../../hwang_b200/csrc/dev/picture.h"

using namespace hwb;

struct hwb_dev { std::string err; uint64_t launches = 0; };
struct hwb_event { int dummy; };
struct PictureScratchEmu { ReconScratch recon; DeblockScratch deblock; };

extern "C" {

int hwb_dev_count(void) { return 1; }
int hwb_dev_open(int, hwb_dev **out) { *out = new hwb_dev(); return 0; }
void hwb_dev_close(hwb_dev *d) { delete d; }
const char *hwb_dev_error(hwb_dev *d) { return d->err.c_str(); }
void *hwb_dev_malloc(hwb_dev *, size_t n) { return malloc(n ? n : 1); }
void hwb_dev_free(hwb_dev *, void *p) { free(p); }
void *hwb_dev_malloc_host(hwb_dev *, size_t n) { return malloc(n ? n : 1); }
void hwb_dev_free_host(hwb_dev *, void *p) { free(p); }
int hwb_dev_pointer_kind(hwb_dev *, const void *) { return 0; }
int hwb_dev_mem_info(hwb_dev *, size_t *free_bytes, size_t *total_bytes) { *free_bytes = (size_t)8 << 30; *total_bytes = (size_t)8 << 30; return 0; }
int hwb_dev_h2d(hwb_dev *, int, void *dst, const void *src, size_t n) { memcpy(dst, src, n); return 0; }
int hwb_dev_d2h(hwb_dev *, int, void *dst, const void *src, size_t n) { memcpy(dst, src, n); return 0; }
int hwb_dev_d2d(hwb_dev *, int, void *dst, const void *src, size_t n) { memcpy(dst, src, n); return 0; }
int hwb_dev_memset(hwb_dev *, int, void *dst, int v, size_t n) { memset(dst, v, n); return 0; }

void hwb_dev_set_occupancy(hwb_dev *, int, int) {}
int hwb_dev_entropy(hwb_dev *d, int, const ChunkCtx *c, int32_t *, int) {
  uint8_t states[1024];
  for (int t = 0; t < c->num_tickets; ++t) { SliceDec sd; decode_slice(*c, c->entropy_order[t], states, &sd); }
  d->launches++;
  if (getenv("HWB_EMU_MBSTATS")) {  // macroblock mix of the batch per picture kind (development aid)
    long hist[2][8] = {}, nzb[2] = {}, n[2] = {};
    for (int p = 0; p < c->num_pics; ++p) {
      const int intra = !c->pics[p].has_inter;
      const MbInfo *mi = pic_mbinfo(*c, c->pics[p].frame);
      for (int m = 0; m < c->nmb; ++m) {
        int k = mi[m].mbtype;
        if (k == MB_INTER) k = (mi[m].flags & MBF_SKIP) ? 5 : (mi[m].cbp ? 6 : 4);
        hist[intra][k]++; nzb[intra] += __builtin_popcount(mi[m].nzmask); n[intra]++;
      }
    }
    for (int i = 0; i < 2; ++i)
      fprintf(stderr, "[mbstats] %s pictures: mbs %ld  I4x4 %ld I8x8 %ld I16 %ld PCM %ld inter-nocbp %ld skip %ld inter-cbp %ld  coded blocks/mb %.2f\n",
              i ? "intra" : "inter", n[i], hist[i][0], hist[i][1], hist[i][2], hist[i][3], hist[i][4], hist[i][5], hist[i][6], n[i] ? (double)nzb[i] / n[i] : 0.0);
  }
  return 0;
}
// The picture kernel's two work lists, executed serially: an item runs once everything it waits for is complete
// (whole rows here), taking from either list; if neither head is ready the host scheduler has produced an order that
// could deadlock on the GPU, which is reported as a device error.
int hwb_dev_picture(hwb_dev *d, int, const ChunkCtx *c, int32_t *) {
  d->launches++;
  if (*c->error_flag) return 0;  // as the CUDA kernel's rows: nothing is dereferenced after an entropy error
  PictureScratchEmu *smp = new PictureScratchEmu(); PictureScratchEmu &sm = *smp;
  auto rows_done = [&](const int32_t *prog, int pic, int y) { return y < 0 || y >= c->mb_h || prog[(size_t)pic * c->mb_h + y] >= c->mb_w; };
  auto recon_ready = [&](uint32_t it) {
    const int pic = item_pic(it), y = item_row(it);
    if (!rows_done(c->recon_prog, pic, y - 1)) return false;
    const int row = c->pics[pic].has_inter ? reference_row_needed(*c, pic, y) : -1;
    if (row >= 0)
      for (int i = 0; i < c->pics[pic].num_dep; ++i) if (!rows_done(c->dbl_prog, c->pics[pic].dep[i], row)) return false;
    return true;
  };
  auto deblock_ready = [&](uint32_t it) {
    const int pic = item_pic(it), y = item_row(it);
    for (int k = 0; k <= c->deblock_band; ++k) if (!rows_done(c->recon_prog, pic, y + k)) return false;  // the band's rows and the one below
    return rows_done(c->dbl_prog, pic, y - 1);
  };
  if (getenv("HWB_EMU_REACH_STATS")) {  // how far inter prediction reaches (what the picture kernel's waits are made of)
    long hx[8] = {0}, hy[8] = {0}, rows = 0;
    for (int p = 0; p < c->num_pics; ++p)
      for (int y = 0; y < c->mb_h; ++y) {
        const int r = c->mv_reach[(size_t)p * c->mb_h + y], rx = c->mv_reach_x[(size_t)p * c->mb_h + y];
        if (r <= 0) continue;
        rows++;
        int dy = r - 1 - y; dy = dy < 0 ? 0 : (dy > 7 ? 7 : dy);
        hy[dy]++; hx[rx > 7 ? 7 : rx]++;
      }
    fprintf(stderr, "[emu reach] inter rows %ld | rows below own (0..7+):", rows);
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %ld", hy[i]);
    fprintf(stderr, " | macroblocks to the right incl. own (0..7+):");
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %ld", hx[i]);
    fprintf(stderr, "\n");
  }
  int ir = 0, id = 0;
  g_unsatisfied_waits = 0;
  g_mc_window.violations = 0;
  while (ir < c->num_recon_items || id < c->num_deblock_items) {
    bool progressed = false;
    while (ir < c->num_recon_items && recon_ready(c->recon_items[ir])) { recon_row(*c, item_pic(c->recon_items[ir]), item_row(c->recon_items[ir]), &sm.recon); ir++; progressed = true; }
    while (id < c->num_deblock_items && deblock_ready(c->deblock_items[id])) { deblock_band(*c, item_pic(c->deblock_items[id]), item_row(c->deblock_items[id]), &sm.deblock); id++; progressed = true; }
    if (!progressed) { *c->error_flag = 901; d->err = "emulation: work lists are not in dependency order"; break; }
  }
  if (g_unsatisfied_waits) { *c->error_flag = 902; d->err = "emulation: a wait was not satisfied"; }
  if (g_mc_window.violations) { *c->error_flag = 903; d->err = "emulation: inter prediction read reference samples beyond the reach the entropy stage announced"; }
  delete smp;
  return 0;
}
int hwb_dev_rgb24(hwb_dev *d, int, const ChunkCtx *c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst) {
  for (int y = 0; y < h; ++y)
    for (int x16 = 0; x16 < (w + 15) / 16; ++x16) rgb24_item(*c, frame, crop_x, crop_y, w, h, dst, x16, y);
  d->launches++;
  return 0;
}
int hwb_dev_yuv(hwb_dev *d, int, const ChunkCtx *c, int frame, int crop_x, int crop_y, int w, int h, uint8_t *dst) {
  const uint8_t *Y = frame_y(*c, frame), *U = frame_cb(*c, frame), *V = frame_cr(*c, frame);
  for (int y = 0; y < h; ++y) memcpy(dst + (size_t)y * w, Y + (size_t)(crop_y + y) * c->wc + crop_x, w);
  uint8_t *du = dst + (size_t)w * h, *dv = du + (size_t)(w / 2) * (h / 2);
  for (int y = 0; y < h / 2; ++y) {
    memcpy(du + (size_t)y * (w / 2), U + (size_t)(crop_y / 2 + y) * (c->wc / 2) + crop_x / 2, w / 2);
    memcpy(dv + (size_t)y * (w / 2), V + (size_t)(crop_y / 2 + y) * (c->wc / 2) + crop_x / 2, w / 2);
  }
  d->launches++;
  return 0;
}

hwb_event *hwb_dev_event_create(hwb_dev *) { return new hwb_event(); }
void hwb_dev_event_destroy(hwb_dev *, hwb_event *e) { delete e; }
int hwb_dev_event_record(hwb_dev *, hwb_event *, int) { return 0; }
int hwb_dev_event_done(hwb_dev *, hwb_event *) { return 1; }
int hwb_dev_event_sync(hwb_dev *, hwb_event *) { return 0; }
int hwb_dev_stream_wait(hwb_dev *, int, hwb_event *) { return 0; }
int hwb_dev_stream_sync(hwb_dev *, int) { return 0; }
int hwb_dev_event_elapsed(hwb_dev *, hwb_event *, hwb_event *, float *ms) { *ms = 0; return 0; }
uint64_t hwb_dev_launch_count(hwb_dev *d) { return d->launches; }

// ---- self-test of the arithmetic decoding engine (csrc/dev/bits.h) against a literal restatement of
// H.264 9.3.3.2 (9-bit codIRange / codIOffset, one bit read per renormalisation step, rangeTabLPS / transIdxLPS /
// transIdxMPS tables).  `ops[i]`: 0..127 = decision with context ops[i] % nctx, 128 = bypass, 129 = terminate.
// Returns the index of the first mismatch, or -1.
struct SpecCabac {
  const uint8_t *d; size_t n, bitpos; uint32_t range, offset;
  uint32_t bit() { uint32_t v = bitpos / 8 < n ? (d[bitpos / 8] >> (7 - bitpos % 8)) & 1 : 0; bitpos++; return v; }
  void start(size_t byte) { bitpos = byte * 8; range = 510; offset = 0; for (int i = 0; i < 9; ++i) offset = (offset << 1) | bit(); }
  int decision(uint8_t &st) {
    const uint32_t p = st >> 1, mps = st & 1, rlps = cabac_range_lps[p * 4 + ((range >> 6) & 3)];
    int bin;
    range -= rlps;
    if (offset >= range) { bin = (int)(mps ^ 1); offset -= range; range = rlps; st = (uint8_t)((cabac_trans_lps[p] << 1) | (p == 0 ? mps ^ 1 : mps)); }
    else { bin = (int)mps; st = (uint8_t)(((p < 62 ? p + 1 : p) << 1) | mps); }
    while (range < 256) { range <<= 1; offset = (offset << 1) | bit(); }
    return bin;
  }
  int bypass() { offset = (offset << 1) | bit(); if (offset >= range) { offset -= range; return 1; } return 0; }
  int terminate() {
    range -= 2;
    if (offset >= range) return 1;
    while (range < 256) { range <<= 1; offset = (offset << 1) | bit(); }
    return 0;
  }
};
int hwb_emu_cabac_selftest(const uint8_t *data, size_t n, size_t start_byte, const uint8_t *ops, size_t nops, int nctx) {
  const CtxE *ft = (const CtxE *)cabac_fused;
  CtxE st_a[128];
  uint8_t st_b[128];
  for (int i = 0; i < 128; ++i) { st_b[i] = (uint8_t)((i * 37 + 11) & 127); st_a[i] = ft[st_b[i]]; }
  Cabac home; cabac_start(home, data, (uint32_t)start_byte);
  CabReg c = cab_enter(home);
  SpecCabac r; r.d = data; r.n = n; r.start(start_byte);
  for (size_t i = 0; i < nops; ++i) {
    int a, b;
    if (ops[i] < 128) { const int k = ops[i] % nctx; a = cabac_decision(c, home, st_a + k, ft); b = r.decision(st_b[k]); if (ctxe_state(st_a[k]) != st_b[k]) return (int)i; }
    else if (ops[i] == 128) { a = cabac_bypass(c, home); b = r.bypass(); }
    else { a = cabac_terminate(c, home); b = r.terminate(); if (a != b) return (int)i; if (a) return -1; }
    if (a != b) return (int)i;
    cab_leave(home, c);
    if (cabac_bitpos(home) != r.bitpos) return (int)i;  // bits consumed: 9 at start + one per renormalisation shift
    if ((size_t)home.pos + 4 > n) return -1;            // ran out of test data
  }
  return -1;
}
}
