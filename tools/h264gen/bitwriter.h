// Bit writer, NAL packaging and the CABAC arithmetic *encoder* (H.264 clause 9.3.4.2) used by
// the synthetic stream generator.  Test/bench tooling only -- never part of the decode path.
#pragma once
#include <stdint.h>
#include <vector>

namespace gen {

struct BitWriter {
  std::vector<uint8_t> buf;
  uint32_t acc = 0;
  int nbits = 0;
  void put(uint32_t v, int n) {
    for (int i = n - 1; i >= 0; --i) {
      acc = (acc << 1) | ((v >> i) & 1);
      if (++nbits == 8) { buf.push_back((uint8_t)acc); acc = 0; nbits = 0; }
    }
  }
  void put1(int b) { put((uint32_t)b, 1); }
  void ue(uint32_t v) {
    uint32_t x = v + 1;
    int len = 32 - __builtin_clz(x);
    put(0, len - 1);
    put(x, len);
  }
  void se(int v) { ue(v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
  bool aligned() const { return nbits == 0; }
  void align_zero() { while (nbits) put1(0); }
  void trailing() { put1(1); align_zero(); }
  size_t bitpos() const { return buf.size() * 8 + nbits; }
};

// Append an RBSP as a NAL unit with emulation prevention; AVCC framing (4-byte length).
inline void append_nal_avcc(std::vector<uint8_t> &out, int nal_ref_idc, int nal_type, const std::vector<uint8_t> &rbsp) {
  std::vector<uint8_t> nal;
  nal.push_back((uint8_t)((nal_ref_idc << 5) | nal_type));
  int zeros = 0;
  for (uint8_t b : rbsp) {
    if (zeros >= 2 && b <= 3) { nal.push_back(3); zeros = 0; }
    nal.push_back(b);
    zeros = b == 0 ? zeros + 1 : 0;
  }
  uint32_t n = (uint32_t)nal.size();
  out.push_back((uint8_t)(n >> 24)); out.push_back((uint8_t)(n >> 16)); out.push_back((uint8_t)(n >> 8)); out.push_back((uint8_t)n);
  out.insert(out.end(), nal.begin(), nal.end());
}
inline std::vector<uint8_t> make_nal_raw(int nal_ref_idc, int nal_type, const std::vector<uint8_t> &rbsp) {
  std::vector<uint8_t> tmp;
  append_nal_avcc(tmp, nal_ref_idc, nal_type, rbsp);
  return std::vector<uint8_t>(tmp.begin() + 4, tmp.end());
}

struct CabacEnc {
  BitWriter *bw = nullptr;
  uint32_t low = 0, range = 510;
  int outstanding = 0;
  bool first = true;
  uint8_t st[1024];
  void start(BitWriter *w) { bw = w; low = 0; range = 510; outstanding = 0; first = true; }
  void put_bit(int b) {
    if (first) first = false; else bw->put1(b);
    while (outstanding > 0) { bw->put1(!b); outstanding--; }
  }
  void renorm() {
    while (range < 256) {
      if (low < 256) put_bit(0);
      else if (low >= 512) { low -= 512; put_bit(1); }
      else { low -= 256; outstanding++; }
      range <<= 1; low <<= 1;
    }
  }
  void decision(int ctx, int bin);
  void bypass(int bin) {
    low <<= 1;
    if (bin) low += range;
    if (low >= 1024) { put_bit(1); low -= 1024; }
    else if (low < 512) put_bit(0);
    else { low -= 512; outstanding++; }
  }
  void terminate(int bin) {
    range -= 2;
    if (bin) {
      low += range;
      range = 2;
      renorm();
      put_bit((low >> 9) & 1);
      bw->put(((low >> 7) & 3) | 1, 2);
    } else renorm();
  }
};

}  // namespace gen
