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
// holds codIOffset << 23 with the next `nb` (<= 16) bits of the stream left-aligned below bit 23, and a context is
// kept as its fused-table entry (see CtxE): a decision is  load entry, pick the rangeLPS byte, subtract, compare,
// select, clz, shift.  Every ~16 consumed bits two more bytes are patched in.  (64-bit shifts cost 2-3
// instructions each on the GPU: the first engine, built on the generic 64-bit bit reader, spent 46 instructions per
// decision.)  The engine reads the slice RBSP directly: `base` is the slice's first byte (16-byte aligned on the
// device), `pos` the byte offset of the next refill (always even).
enum { CABAC_SCALE = 23 };
// The engine as it is stored (slice state, shared memory on the device) ...
struct alignas(16) Cabac {
  uint32_t low;
  uint32_t range;
  int32_t nb;      // valid stream bits below bit 23 of `low`
  uint32_t pos;    // byte offset of the next refill (always even)
  const uint8_t *base;  // the slice RBSP
  uint64_t pad;
};
// ... and as the decoding loops hold it in registers.  The byte position stays in memory: only the refill (once per
// ~16 consumed bits, out of line) touches it.
struct CabReg { uint32_t low, range; int32_t nb; };
HWB_HD CabReg cab_enter(const Cabac &h) {
  CabReg c;
#if HWB_DEVICE_BUILD
  const uint4 v = *(const uint4 *)&h;
  c.low = v.x; c.range = v.y; c.nb = (int32_t)v.z;
#else
  c.low = h.low; c.range = h.range; c.nb = h.nb;
#endif
  return c;
}
HWB_HD void cab_leave(Cabac &h, const CabReg &c) {
#if HWB_DEVICE_BUILD
  *(uint2 *)&h = make_uint2(c.low, c.range);
#else
  h.low = c.low; h.range = c.range;
#endif
  h.nb = c.nb;
}

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
  c.base = base;
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
struct CabacFill { uint32_t low; int32_t nb; };
HWB_FN CabacFill cabac_refill_ool(uint32_t low, int32_t nb, Cabac *home) {
  CabacFill f;
  const uint32_t pos = home->pos;
  f.low = low | (cabac_load16(home->base, pos) << (7 - nb));
  f.nb = nb + 16;
  home->pos = pos + 2;
  return f;
}
HWB_HD void cabac_refill(CabReg &c, Cabac &home) {
  if (c.nb <= 0) {
    const CabacFill f = cabac_refill_ool(c.low, c.nb, &home);
    c.low = f.low; c.nb = f.nb;
  }
}

// Initial state (pStateIdx << 1 | valMPS) of context i (9.3.1.1).
HWB_HD int cabac_init_state(int table, int slice_qp, int i) {
  const int8_t *mn = cabac_init_mn + (table * HWB_CABAC_NCTX + i) * 2;
  const int pre = clip3(1, 126, ((mn[0] * clip3(0, 51, slice_qp)) >> 4) + mn[1]);
  return pre <= 63 ? (63 - pre) << 1 : ((pre - 64) << 1) | 1;
}
HWB_HD void cabac_init_states(uint8_t *st, int table, int slice_qp) {  // byte states: the stream generator's encoder
  for (int i = 0; i < HWB_CABAC_NCTX; ++i) st[i] = (uint8_t)cabac_init_state(table, slice_qp, i);
}

// A context as the decoder keeps it: not the state byte but the 8-byte entry of the fused table that belongs to the
// state (rangeLPS for the four codIRange quantiser values; both successor states; the state itself, bit 16 = valMPS).
// A decision loads the entry (the only load on its dependency chain), and afterwards replaces it by the successor
// state's entry -- a table lookup and a store that nothing waits for unless the next decision uses the same context.
// (Round 1 kept a state byte per context and looked the entry up inside the decision: two dependent shared-memory
// loads per bin.)
struct alignas(8) CtxE { uint32_t rl4, ns; };
HWB_HD CtxE ctxe_load(const CtxE *p) {
#if HWB_DEVICE_BUILD
  const uint2 v = *(const uint2 *)p;
  CtxE e; e.rl4 = v.x; e.ns = v.y;
  return e;
#else
  return *p;
#endif
}
HWB_HD void ctxe_store(CtxE *p, const CtxE e) {
#if HWB_DEVICE_BUILD
  *(uint2 *)p = make_uint2(e.rl4, e.ns);
#else
  *p = e;
#endif
}
HWB_HD int ctxe_state(const CtxE &e) { return (int)((e.ns >> 16) & 127); }
// Binary decision (9.3.3.2.1).  `ft` = the fused table as 128 entries (a copy near the contexts on the device).  On the
// chain from codIRange to the next codIRange: pick the rangeLPS byte (quantiser bits 30:29 of the scaled range -> byte
// selector), subtract, compare, select, count leading zeros, shift.
// `e` = the context's entry, loaded by the caller (possibly ahead of time, see the significance map loop).
HWB_HD int cabac_decide(CabReg &c, Cabac &home, const CtxE e, CtxE *ctx, const CtxE *ft) {
#if HWB_DEVICE_BUILD
  const uint32_t rlps = __byte_perm(e.rl4, 0, ((c.range >> 17) & 0x3000u) | 0x0444u) >> 1;  // byte q -> bits 30:23
#else
  const uint32_t rlps = ((e.rl4 >> (8 * ((c.range >> 29) & 3))) & 0xffu) << CABAC_SCALE;
#endif
  const uint32_t rmps = c.range - rlps;
  const bool lps = c.low >= rmps;
  c.low = lps ? c.low - rmps : c.low;
  const uint32_t r = lps ? rlps : rmps;
#if HWB_DEVICE_BUILD
  const uint32_t nsi = __byte_perm(e.ns, 0, lps ? 0x4441u : 0x4440u);
#else
  const uint32_t nsi = (lps ? e.ns >> 8 : e.ns) & 0xffu;
#endif
  ctxe_store(ctx, ctxe_load(ft + nsi));
  const int sh = cabac_norm_shift(r);
  c.range = r << sh;
  c.low <<= sh;
  c.nb -= sh;
  cabac_refill(c, home);
  return (int)(((e.ns >> 16) & 1) ^ (lps ? 1u : 0u));
}
HWB_HD int cabac_decision(CabReg &c, Cabac &home, CtxE *ctx, const CtxE *ft) {
  return cabac_decide(c, home, ctxe_load(ctx), ctx, ft);
}
HWB_HD int cabac_bypass(CabReg &c, Cabac &home) {
  // compare before shifting: 2 * offset + next bit >= range  <=>  low >= range / 2 (exact: range << 22)
  const uint32_t half = c.range >> 1;
  const bool one = c.low >= half;
  c.low = (c.low - (one ? half : 0u)) << 1;
  --c.nb;
  cabac_refill(c, home);
  return one ? 1 : 0;
}
HWB_HD int cabac_terminate(CabReg &c, Cabac &home) {
  c.range -= 2u << CABAC_SCALE;
  if (c.low >= c.range) return 1;
  if (!(c.range >> 31)) {
    c.range <<= 1; c.low <<= 1;
    --c.nb;
    cabac_refill(c, home);
  }
  return 0;
}
// Suffix of a UEGk binarisation (9.3.2.3): unary-coded exponent, then that many bits, all bypass.  Out of line and
// with the engine passed in registers: escapes are rare, and inlined they were 1.4 KB of a fetch-bound decoder.
// Returns the value to add; ok = false if the exponent runs away (corrupt stream).
struct CabEsc { uint32_t low; int32_t nb; int value; };
HWB_FN CabEsc cabac_escape_ool(uint32_t low, uint32_t range, int32_t nb, Cabac *home, int k) {
  CabReg c; c.low = low; c.range = range; c.nb = nb;
  int v = 0;
  const int kmax = k + 21;
#pragma unroll 1
  while (cabac_bypass(c, *home)) { v += 1 << k; k++; if (k > kmax) { v = -1; k = 0; break; } }
#pragma unroll 1
  while (k--) v += cabac_bypass(c, *home) << k;
  CabEsc r; r.low = c.low; r.nb = c.nb; r.value = v;
  return r;
}
HWB_HD int cabac_escape(CabReg &c, Cabac &home, int k) {
  const CabEsc r = cabac_escape_ool(c.low, c.range, c.nb, &home, k);
  c.low = r.low; c.nb = r.nb;
  return r.value;
}

}  // namespace hwb
