// Bit reader over an RBSP (emulation-prevention bytes already removed by the host) and the
// CABAC arithmetic decoding engine (H.264 clause 9.3.3.2).  Single-lane code: one lane of a
// warp owns a slice.
#pragma once
#include "hd.h"
#include "tables_gen.h"

namespace hwb {

struct BitReader {
  const uint8_t *base;
  uint32_t size;     // bytes
  uint32_t pos;      // next byte to load into the cache
  uint64_t cache;    // left-aligned: next bit is bit 63
  int32_t avail;     // valid bits in cache
  bool overrun;
};

HWB_HD void br_refill(BitReader &b) {
#if HWB_DEVICE_BUILD
  // Device: the byte position is always word aligned (see br_init) and this is only called with avail <= 32, so a
  // refill is one 32-bit load through the read-only path.  Slice data starts on a 16-byte boundary and the chunk
  // bitstream is padded, so reading a few bytes past the slice is safe.
  uint32_t w = __ldg((const uint32_t *)(b.base + b.pos));
  w = __byte_perm(w, 0, 0x0123);
  if (b.pos >= b.size + 8) b.overrun = true;
  b.cache |= (uint64_t)w << (32 - b.avail);
  b.avail += 32;
  b.pos += 4;
#else
  while (b.avail <= 56) {
    uint64_t v = b.pos < b.size ? b.base[b.pos] : 0;
    if (b.pos >= b.size + 8) b.overrun = true;
    b.pos++;
    b.cache |= v << (56 - b.avail);
    b.avail += 8;
  }
#endif
}
HWB_HD void br_init(BitReader &b, const uint8_t *p, uint32_t size, uint32_t bit_off) {
  b.base = p; b.size = size; b.cache = 0; b.avail = 0; b.overrun = false;
#if HWB_DEVICE_BUILD
  b.pos = (bit_off >> 3) & ~3u;
  br_refill(b);
  const int skip = (int)(bit_off - b.pos * 8 + 32);  // pos already advanced by 4
  b.cache <<= skip; b.avail -= skip;
#else
  b.pos = bit_off >> 3;
  br_refill(b);
  b.cache <<= (bit_off & 7); b.avail -= (bit_off & 7);
#endif
}
HWB_HD uint32_t br_bitpos(const BitReader &b) { return b.pos * 8 - b.avail; }
HWB_HD uint32_t br_peek(BitReader &b, int n) {  // n in 1..32
  if (b.avail < n) br_refill(b);
  return (uint32_t)(b.cache >> (64 - n));
}
HWB_HD void br_skip(BitReader &b, int n) { b.cache <<= n; b.avail -= n; }
HWB_HD uint32_t br_get(BitReader &b, int n) {
  if (n == 0) return 0;
  uint32_t v = br_peek(b, n);
  br_skip(b, n);
  return v;
}
HWB_HD uint32_t br_get1(BitReader &b) { return br_get(b, 1); }
HWB_HD uint32_t br_ue(BitReader &b) {
  uint32_t v = br_peek(b, 32);
  int lz = clz32(v);
  if (lz >= 32) { br_skip(b, 32); b.overrun = true; return 0; }
  br_skip(b, lz);
  uint32_t r = br_get(b, lz + 1);
  return r - 1;
}
HWB_HD int32_t br_se(BitReader &b) {
  uint32_t k = br_ue(b);
  return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
}
HWB_HD void br_align(BitReader &b) { br_skip(b, b.avail & 7); }
// more_rbsp_data(): true unless only the rbsp trailing bits (1 followed by zeros) remain.
HWB_HD bool br_more_rbsp_data(BitReader &b, uint32_t last_one_bitpos) { return br_bitpos(b) < last_one_bitpos; }

// ------------------------------------------------------------------------------------ CABAC
struct Cabac {
  uint32_t range;   // codIRange (9 bits)
  uint32_t offset;  // codIOffset (9 bits)
};

HWB_HD void cabac_start(Cabac &c, BitReader &b) {
  c.range = 510;
  c.offset = br_get(b, 9);
}

HWB_HD void cabac_init_states(uint8_t *st, int table, int slice_qp) {
  const int8_t *mn = cabac_init_mn + table * HWB_CABAC_NCTX * 2;
  int qp = clip3(0, 51, slice_qp);
  for (int i = 0; i < HWB_CABAC_NCTX; ++i) {
    int pre = clip3(1, 126, ((mn[2 * i] * qp) >> 4) + mn[2 * i + 1]);
    st[i] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);
  }
}

// Branch-free binary decision (9.3.3.2.1): one fused table entry gives rangeLPS and both successor states.
HWB_HD int cabac_decision(Cabac &c, BitReader &b, uint8_t *state) {
  if (b.avail < 16) br_refill(b);
  const uint32_t s = *state;
  const uint32_t e = cabac_fused[s * 4 + ((c.range >> 6) & 3)];
  const uint32_t rlps = e & 0xff;
  const uint32_t rmps = c.range - rlps;
  const bool lps = c.offset >= rmps;
  c.offset = lps ? c.offset - rmps : c.offset;
  c.range = lps ? rlps : rmps;
  *state = (uint8_t)(lps ? (e >> 8) : (e >> 16));
  const int sh = clz32(c.range) - 23;  // renormalisation shift, 0 when range >= 256
  c.range <<= sh;
  c.offset = (c.offset << sh) | (uint32_t)((b.cache >> 1) >> (63 - sh));
  b.cache <<= sh; b.avail -= sh;
  return (int)((s & 1) ^ (lps ? 1u : 0u));
}
HWB_HD int cabac_bypass(Cabac &c, BitReader &b) {
  if (b.avail < 16) br_refill(b);
  c.offset = (c.offset << 1) | (uint32_t)(b.cache >> 63);
  b.cache <<= 1; b.avail -= 1;
  const bool one = c.offset >= c.range;
  c.offset -= one ? c.range : 0u;
  return one ? 1 : 0;
}
HWB_HD int cabac_terminate(Cabac &c, BitReader &b) {
  c.range -= 2;
  if (c.offset >= c.range) return 1;
  if (c.range < 256) {
    c.range <<= 1;
    c.offset = (c.offset << 1) | br_get(b, 1);
  }
  return 0;
}

}  // namespace hwb
