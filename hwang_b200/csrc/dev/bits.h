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
// Arithmetic decoding engine (9.3.3.2) on 32-bit registers only, everything scaled by 2^23: `range` holds
// codIRange << 23 (normalised <=> bit 31 set, so the renormalisation shift is a plain count-leading-zeros), `low`
// holds codIOffset << 23 with the next `nb` (<= 16) bits of the stream left-aligned below bit 23, and the fused
// table is indexed by the context state alone (see CtxPre): a decision is  load state, load entry, pick the rangeLPS
// byte, subtract, compare, select, clz, shift.  Every ~16 consumed bits two more bytes are patched in.  (64-bit shifts cost 2-3
// instructions each on the GPU: the first engine, built on the generic 64-bit bit reader, spent 46 instructions per
// decision.)  The engine reads the slice RBSP directly: `base` is the slice's first byte (16-byte aligned on the
// device), `pos` the byte offset of the next refill (always even).
#ifndef HWB_CABAC_FUSED
#define HWB_CABAC_FUSED cabac_fused
#endif
enum { CABAC_SCALE = 23 };
struct alignas(16) Cabac {
  uint32_t low;
  uint32_t range;
  int32_t nb;      // valid stream bits below bit 23 of `low`
  uint32_t pos;
};

HWB_HD uint32_t cabac_load16(const uint8_t *base, uint32_t pos) {
#if HWB_DEVICE_BUILD
  return __byte_perm((uint32_t)__ldg((const unsigned short *)(base + pos)), 0, 0x4401);
#else
  return ((uint32_t)base[pos] << 8) | base[pos + 1];
#endif
}
// count of leading zeros of a non-zero word in one instruction (bfind.shiftamt)
HWB_HD int cabac_norm_shift(uint32_t r) {
#if HWB_DEVICE_BUILD
  int sh;
  asm("bfind.shiftamt.u32 %0, %1;" : "=r"(sh) : "r"(r));
  return sh;
#else
  return __builtin_clz(r);
#endif
}
// Start (or restart, after I_PCM samples) the engine at byte `p` of the RBSP: codIOffset = the next 9 bits.
HWB_HD void cabac_start(Cabac &c, const uint8_t *base, uint32_t p) {
  c.range = 510u << CABAC_SCALE;
  if (p & 1) {  // 3 bytes: 9 offset bits + 15 pending
    const uint32_t b0 = base[p];
    c.low = ((b0 << 16) | cabac_load16(base, p + 1)) << 8;
    c.nb = 15; c.pos = p + 3;
  } else {      // 2 bytes: 9 offset bits + 7 pending
    c.low = cabac_load16(base, p) << 16;
    c.nb = 7; c.pos = p + 2;
  }
}
// Bits of the RBSP consumed by the arithmetic decoder so far (9 at start + one per renormalisation shift).
HWB_HD uint32_t cabac_bitpos(const Cabac &c) { return c.pos * 8 - (uint32_t)c.nb; }
// Invariant between operations: nb >= 1 (a bypass decision compares before it shifts, so the next stream bit must
// already be in `low`).  nb <= 0: -nb zero bits were shifted into the offset; the top bits of the next 16 belong
// there.  The refill itself is one shared out-of-line routine (it runs once per ~20 bins; inlined at every decision
// site it was a quarter of the residual decoder's code).
struct CabacFill { uint32_t low; int32_t nb; uint32_t pos; };
HWB_FN CabacFill cabac_refill_ool(uint32_t low, int32_t nb, uint32_t pos, const uint8_t *base) {
  CabacFill f;
  f.low = low | (cabac_load16(base, pos) << (7 - nb));
  f.nb = nb + 16; f.pos = pos + 2;
  return f;
}
HWB_HD void cabac_refill(Cabac &c, const uint8_t *base) {
  if (c.nb <= 0) {
    const CabacFill f = cabac_refill_ool(c.low, c.nb, c.pos, base);
    c.low = f.low; c.nb = f.nb; c.pos = f.pos;
  }
}

HWB_HD void cabac_init_states(uint8_t *st, int table, int slice_qp) {
  const int8_t *mn = cabac_init_mn + table * HWB_CABAC_NCTX * 2;
  int qp = clip3(0, 51, slice_qp);
  for (int i = 0; i < HWB_CABAC_NCTX; ++i) {
    int pre = clip3(1, 126, ((mn[2 * i] * qp) >> 4) + mn[2 * i + 1]);
    st[i] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);
  }
}

// A context as the decision uses it: the state byte and the two table words that depend on it alone (rangeLPS for
// the four codIRange quantiser values, and both successor states), so both loads issue before codIRange is needed.
// (Fetching the *next* contexts ahead of time was tried and lost: on this latency-bound single-warp code the extra
// instructions cost more than the shared-memory round trips they hid.)
struct CtxPre { uint32_t s, rl4, ns; };
HWB_HD CtxPre cabac_prefetch(const uint8_t *state) {
  CtxPre p;
  p.s = *state;
#if HWB_DEVICE_BUILD
  const uint2 e = *(const uint2 *)(HWB_CABAC_FUSED + 2 * p.s);
  p.rl4 = e.x; p.ns = e.y;
#else
  p.rl4 = HWB_CABAC_FUSED[2 * p.s]; p.ns = HWB_CABAC_FUSED[2 * p.s + 1];
#endif
  return p;
}
// Binary decision (9.3.3.2.1) with a prefetched context.  On the chain from codIRange to the decision: pick the
// rangeLPS byte (quantiser bits 30:29 of the scaled range -> byte selector), subtract, compare.
HWB_HD int cabac_decide(Cabac &c, const uint8_t *base, const CtxPre p, uint8_t *state) {
#if HWB_DEVICE_BUILD
  const uint32_t rlps = __byte_perm(p.rl4, 0, ((c.range >> 17) & 0x3000u) | 0x0444u) >> 1;  // byte q -> bits 30:23
#else
  const uint32_t rlps = ((p.rl4 >> (8 * ((c.range >> 29) & 3))) & 0xffu) << CABAC_SCALE;
#endif
  const uint32_t rmps = c.range - rlps;
  const bool lps = c.low >= rmps;
  c.low = lps ? c.low - rmps : c.low;
  const uint32_t r = lps ? rlps : rmps;
  *state = (uint8_t)(lps ? (p.ns >> 8) : p.ns);
  const int sh = cabac_norm_shift(r);
  c.range = r << sh;
  c.low <<= sh;
  c.nb -= sh;
  cabac_refill(c, base);
  return (int)((p.s & 1) ^ (lps ? 1u : 0u));
}
HWB_HD int cabac_decision(Cabac &c, const uint8_t *base, uint8_t *state) {
  return cabac_decide(c, base, cabac_prefetch(state), state);
}
HWB_HD int cabac_bypass(Cabac &c, const uint8_t *base) {
  // compare before shifting: 2 * offset + next bit >= range  <=>  low >= range / 2 (exact: range << 22)
  const uint32_t half = c.range >> 1;
  const bool one = c.low >= half;
  c.low = (c.low - (one ? half : 0u)) << 1;
  --c.nb;
  cabac_refill(c, base);
  return one ? 1 : 0;
}
HWB_HD int cabac_terminate(Cabac &c, const uint8_t *base) {
  c.range -= 2u << CABAC_SCALE;
  if (c.low >= c.range) return 1;
  if (!(c.range >> 31)) {
    c.range <<= 1; c.low <<= 1;
    --c.nb;
    cabac_refill(c, base);
  }
  return 0;
}

}  // namespace hwb
