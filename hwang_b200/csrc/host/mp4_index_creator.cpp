#include "mp4_index_creator.h"

#include <string.h>
#include <functional>

namespace hwang {

namespace {

inline uint32_t rd32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint64_t rd64(const uint8_t *p) { return ((uint64_t)rd32(p) << 32) | rd32(p + 4); }
inline uint32_t fourcc(const char *s) { return rd32((const uint8_t *)s); }

struct Box {
  uint32_t type = 0;
  uint64_t size = 0;    // whole box
  uint32_t header = 0;  // header bytes
  bool ok = false;
};

// Reads a box header from [p, p+n).  size==0 ("to end of file") is resolved by the caller.
Box read_box(const uint8_t *p, size_t n) {
  Box b;
  if (n < 8) return b;
  b.size = rd32(p); b.type = rd32(p + 4); b.header = 8;
  if (b.size == 1) {
    if (n < 16) return b;
    b.size = rd64(p + 8); b.header = 16;
  }
  b.ok = true;
  return b;
}

// Iterate the child boxes inside [p, p+n); fn returns false to stop.
bool for_each_box(const uint8_t *p, size_t n, const std::function<bool(const Box &, const uint8_t *, size_t)> &fn) {
  size_t pos = 0;
  while (pos + 8 <= n) {
    Box b = read_box(p + pos, n - pos);
    if (!b.ok) return false;
    uint64_t sz = b.size == 0 ? n - pos : b.size;
    if (sz < b.header || pos + sz > n) return false;
    if (!fn(b, p + pos + b.header, (size_t)(sz - b.header))) return true;
    pos += (size_t)sz;
  }
  return true;
}

}  // namespace

MP4IndexCreator::MP4IndexCreator(uint64_t file_size) : file_size_(file_size) {}

bool MP4IndexCreator::fail(const std::string &msg) {
  if (!error_) error_message_ = msg;
  error_ = true;
  done_ = true;
  return false;
}

bool MP4IndexCreator::feed(const uint8_t *data, size_t size, uint64_t &next_offset, uint64_t &next_size) {
  if (is_done()) return false;
  size_t pos = 0;
  auto more_limit = [&](uint64_t off, uint64_t want) -> bool {
    // reference MORE_DATA_LIMIT (mp4_index_creator.cpp:88-110)
    if (off + want > file_size_) {
      want = off < file_size_ ? file_size_ - off : 0;
      if (want == 0) {
        if (parsed_ftyp_ && parsed_moov_ && fragments_present_) { done_ = true; return false; }
        return fail("Reached EOF without being done");
      }
    }
    offset_ = off; next_offset = off; next_size = want;
    return true;
  };
  while (!is_done()) {
    if (file_size_ - (offset_ + pos) < 8) {
      // trailing bytes that cannot hold a box header: same as reaching EOF
      if (parsed_ftyp_ && parsed_moov_ && fragments_present_) { done_ = true; return false; }
      return fail("Reached EOF without being done");
    }
    if (size - pos < 8) return more_limit(offset_ + pos, 1024);
    Box b = read_box(data + pos, size - pos);
    if (!b.ok) {
      if (file_size_ - (offset_ + pos) < 16) return fail("Truncated box header");
      return more_limit(offset_ + pos, 1024);
    }
    const uint64_t box_off = offset_ + pos;
    uint64_t bsize = b.size == 0 ? file_size_ - box_off : b.size;
    if (bsize < b.header) return fail("Invalid box size");
    const bool want = (!parsed_ftyp_ && b.type == fourcc("ftyp")) || (!parsed_moov_ && b.type == fourcc("moov")) || b.type == fourcc("moof");
    if (want) {
      if (size - pos < bsize) {
        // reference MORE_DATA (mp4_index_creator.cpp:78-87): need the whole box
        if (box_off + bsize > file_size_) return fail("EOF in middle of box");
        offset_ = box_off; next_offset = box_off; next_size = bsize;
        return true;
      }
      const uint8_t *payload = data + pos + b.header;
      const size_t plen = (size_t)(bsize - b.header);
      if (b.type == fourcc("ftyp")) {
        // major brand, minor version, compatible brands (reference :153-171)
        bool ok = false;
        std::string brands;
        for (size_t i = 8; i + 4 <= plen; i += 4) {
          uint32_t c = rd32(payload + i);
          if (c == fourcc("isom") || c == fourcc("iso2") || c == fourcc("avc1")) ok = true;
          brands += std::string((const char *)payload + i, 4) + ", ";
        }
        if (!ok) return fail("No supported mp4 brands: " + brands);
        parsed_ftyp_ = true;
      } else if (b.type == fourcc("moov")) {
        if (!parse_moov(payload, plen)) return false;
        parsed_moov_ = true;
      } else {
        if (!parse_moof(payload, plen, box_off)) return false;
      }
    }
    if (size - pos <= bsize) {
      // the next box starts beyond this buffer
      if (is_done()) return false;
      return more_limit(box_off + bsize, 1024);
    }
    pos += (size_t)bsize;
  }
  return false;
}

bool MP4IndexCreator::parse_moov(const uint8_t *p, size_t n) {
  bool ok = true;
  bool structure = for_each_box(p, n, [&](const Box &b, const uint8_t *q, size_t m) {
    if (b.type == fourcc("trak") && !have_video_track_) { if (!parse_trak(q, m)) { ok = false; return false; } }
    else if (b.type == fourcc("mvex")) {
      fragments_present_ = true;
      for_each_box(q, m, [&](const Box &c, const uint8_t *r, size_t k) {
        if (c.type == fourcc("trex") && k >= 24) trex_.push_back({rd32(r + 4), rd32(r + 16), rd32(r + 20)});
        return true;
      });
    }
    return true;
  });
  if (!ok) return false;
  if (!structure) return fail("Malformed 'moov' box");
  if (!have_video_track_) return fail("Could not find a video track");
  return true;
}

bool MP4IndexCreator::parse_trak(const uint8_t *p, size_t n) {
  uint32_t track_id = 0;
  bool is_video = false, ok = true;
  uint32_t timescale = 0;
  uint64_t duration = 0;
  const uint8_t *stbl = nullptr;
  size_t stbl_n = 0;
  for_each_box(p, n, [&](const Box &b, const uint8_t *q, size_t m) {
    if (b.type == fourcc("tkhd") && m >= 24) { int v = q[0]; track_id = v == 1 ? rd32(q + 20) : rd32(q + 12); }
    else if (b.type == fourcc("mdia")) {
      for_each_box(q, m, [&](const Box &c, const uint8_t *r, size_t k) {
        if (c.type == fourcc("mdhd") && k >= 24) {
          if (r[0] == 1 && k >= 36) { timescale = rd32(r + 20); duration = rd64(r + 24); }
          else { timescale = rd32(r + 12); duration = rd32(r + 16); }
        } else if (c.type == fourcc("hdlr") && k >= 12) is_video = rd32(r + 8) == fourcc("vide");
        else if (c.type == fourcc("minf")) {
          for_each_box(r, k, [&](const Box &d, const uint8_t *s, size_t l) {
            if (d.type == fourcc("stbl")) { stbl = s; stbl_n = l; }
            return true;
          });
        }
        return true;
      });
    }
    return true;
  });
  if (!is_video) return true;  // not the video track: keep looking
  if (!stbl) return fail("Could not find 'stbl'");
  timescale_ = timescale; duration_ = duration; video_track_id_ = track_id;
  ok = parse_stbl(stbl, stbl_n);
  if (ok) have_video_track_ = true;
  return ok;
}

bool MP4IndexCreator::parse_stbl(const uint8_t *p, size_t n) {
  std::vector<uint64_t> sizes, chunk_offsets;
  struct Stsc { uint32_t first, per, desc; };
  std::vector<Stsc> stsc;
  std::vector<uint64_t> sync;
  bool have_stss = false, have_stsd = false, bad = false;
  for_each_box(p, n, [&](const Box &b, const uint8_t *q, size_t m) {
    if (b.type == fourcc("stsd") && m >= 8) {
      // first sample entry (reference :420-470)
      if (rd32(q + 4) < 1 || m < 16 + 78) { bad = true; return false; }
      const uint8_t *e = q + 8;
      uint64_t esz = rd32(e);
      if (esz < 86 || 8 + esz > m) { bad = true; return false; }
      format_ = std::string((const char *)e + 4, 4);
      width_ = ((uint32_t)e[32] << 8) | e[33]; height_ = ((uint32_t)e[34] << 8) | e[35];
      for_each_box(e + 86, (size_t)esz - 86, [&](const Box &c, const uint8_t *r, size_t k) {
        if (c.type == fourcc("avcC") || c.type == fourcc("hvcC")) extradata_.assign(r, r + k);
        return true;
      });
      have_stsd = true;
    } else if (b.type == fourcc("stsz") && m >= 12) {
      uint32_t fixed = rd32(q + 4), cnt = rd32(q + 8);
      if (!fixed && (uint64_t)cnt * 4 + 12 > m) { bad = true; return false; }
      sizes.resize(cnt);
      for (uint32_t i = 0; i < cnt; ++i) sizes[i] = fixed ? fixed : rd32(q + 12 + 4 * (size_t)i);
    } else if (b.type == fourcc("stz2") && m >= 12) {
      uint32_t field = q[7], cnt = rd32(q + 8);
      if (!(field == 4 || field == 8 || field == 16) || ((uint64_t)cnt * field + 7) / 8 + 12 > m) { bad = true; return false; }
      sizes.resize(cnt);
      for (uint32_t i = 0; i < cnt; ++i) {
        if (field == 4) sizes[i] = (q[12 + i / 2] >> ((i & 1) ? 0 : 4)) & 15;
        else if (field == 8) sizes[i] = q[12 + i];
        else sizes[i] = ((uint32_t)q[12 + 2 * (size_t)i] << 8) | q[13 + 2 * (size_t)i];
      }
    } else if (b.type == fourcc("stsc") && m >= 8) {
      uint32_t cnt = rd32(q + 4);
      if ((uint64_t)cnt * 12 + 8 > m) { bad = true; return false; }
      for (uint32_t i = 0; i < cnt; ++i) stsc.push_back({rd32(q + 8 + 12 * (size_t)i), rd32(q + 12 + 12 * (size_t)i), rd32(q + 16 + 12 * (size_t)i)});
    } else if (b.type == fourcc("stco") && m >= 8) {
      uint32_t cnt = rd32(q + 4);
      if ((uint64_t)cnt * 4 + 8 > m) { bad = true; return false; }
      for (uint32_t i = 0; i < cnt; ++i) chunk_offsets.push_back(rd32(q + 8 + 4 * (size_t)i));
    } else if (b.type == fourcc("co64") && m >= 8) {
      uint32_t cnt = rd32(q + 4);
      if ((uint64_t)cnt * 8 + 8 > m) { bad = true; return false; }
      for (uint32_t i = 0; i < cnt; ++i) chunk_offsets.push_back(rd64(q + 8 + 8 * (size_t)i));
    } else if (b.type == fourcc("stss") && m >= 8) {
      uint32_t cnt = rd32(q + 4);
      if ((uint64_t)cnt * 4 + 8 > m) { bad = true; return false; }
      have_stss = true;
      for (uint32_t i = 0; i < cnt; ++i) sync.push_back(rd32(q + 8 + 4 * (size_t)i) - 1);
    }
    return true;
  });
  if (bad) return fail("Malformed sample table");
  if (!have_stsd) return fail("Could not find 'stsd'");
  // chunk -> sample offsets (reference :341-394)
  size_t si = 0;
  for (size_t ci = 0; ci < chunk_offsets.size() && si < sizes.size(); ++ci) {
    uint32_t per = 0;
    for (auto &e : stsc) if (ci + 1 >= e.first) per = e.per;
    uint64_t off = chunk_offsets[ci];
    for (uint32_t k = 0; k < per && si < sizes.size(); ++k) {
      sample_offsets_.push_back(off); sample_sizes_.push_back(sizes[si]);
      off += sizes[si++];
    }
  }
  if (si != sizes.size() && !sizes.empty()) return fail("Sample table does not cover every sample");
  if (have_stss) { for (uint64_t k : sync) if (k < sample_sizes_.size()) keyframe_indices_.push_back(k); }
  else for (size_t i = 0; i < sample_sizes_.size(); ++i) keyframe_indices_.push_back(i);  // no stss: every sample is a sync sample
  return true;
}

bool MP4IndexCreator::parse_moof(const uint8_t *p, size_t n, uint64_t moof_off) {
  bool first_traf = true, ok = true;
  uint64_t prev_traf_end = 0;
  for_each_box(p, n, [&](const Box &b, const uint8_t *q, size_t m) {
    if (b.type != fourcc("traf")) return true;
    // tfhd
    bool have_tfhd = false;
    uint32_t flags = 0, track_id = 0, def_size = 0, def_flags = 0;
    uint64_t base = 0;
    bool base_present = false, size_present = false, flags_present = false;
    for_each_box(q, m, [&](const Box &c, const uint8_t *r, size_t k) {
      if (c.type != fourcc("tfhd") || k < 8) return true;
      have_tfhd = true;
      flags = rd32(r) & 0xFFFFFF; track_id = rd32(r + 4);
      size_t o = 8;
      if (flags & 0x1) { if (o + 8 <= k) { base = rd64(r + o); base_present = true; } o += 8; }
      if (flags & 0x2) o += 4;
      if (flags & 0x8) o += 4;
      if (flags & 0x10) { if (o + 4 <= k) { def_size = rd32(r + o); size_present = true; } o += 4; }
      if (flags & 0x20) { if (o + 4 <= k) { def_flags = rd32(r + o); flags_present = true; } o += 4; }
      return false;
    });
    if (!have_tfhd) { ok = fail("Could not find 'tfhd'"); return false; }
    if (have_video_track_ && video_track_id_ && track_id != video_track_id_) { first_traf = false; return true; }
    uint64_t base_off = base_present ? base : ((flags & 0x020000) || first_traf ? moof_off : prev_traf_end);
    const Trex *trex = nullptr;
    for (auto &t : trex_) if (t.track_id == track_id) trex = &t;
    if (!trex) { ok = fail("Could not find 'trex' for track id in 'tfhd'"); return false; }
    uint64_t prev_run_end = base_off;
    for_each_box(q, m, [&](const Box &c, const uint8_t *r, size_t k) {
      if (c.type != fourcc("trun") || k < 8) return true;
      uint32_t tf = rd32(r) & 0xFFFFFF, cnt = rd32(r + 4);
      size_t o = 8;
      uint64_t data_off = prev_run_end;
      if (tf & 0x1) { if (o + 4 > k) return false; data_off = base_off + (int64_t)(int32_t)rd32(r + o); o += 4; }
      uint32_t first_flags = 0;
      bool has_first = false;
      if (tf & 0x4) { if (o + 4 > k) return false; first_flags = rd32(r + o); has_first = true; o += 4; }
      const size_t per = ((tf & 0x100) ? 4 : 0) + ((tf & 0x200) ? 4 : 0) + ((tf & 0x400) ? 4 : 0) + ((tf & 0x800) ? 4 : 0);
      if (o + (uint64_t)cnt * per > k) { ok = fail("Malformed 'trun' box"); return false; }
      uint64_t cur = data_off;
      for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t ssize = size_present ? def_size : trex->default_size;
        uint32_t sflags = flags_present ? def_flags : trex->default_flags;
        if (tf & 0x100) o += 4;
        if (tf & 0x200) { ssize = rd32(r + o); o += 4; }
        if (tf & 0x400) { sflags = rd32(r + o); o += 4; }
        else if (i == 0 && has_first) sflags = first_flags;
        if (tf & 0x800) o += 4;
        if ((sflags & 0x00010000) == 0) keyframe_indices_.push_back(sample_sizes_.size());
        sample_offsets_.push_back(cur); sample_sizes_.push_back(ssize);
        cur += ssize;
      }
      prev_run_end = cur;
      return true;
    });
    prev_traf_end = prev_run_end;
    first_traf = false;
    return ok;
  });
  return ok;
}

VideoIndex MP4IndexCreator::get_video_index() {
  return VideoIndex(timescale_, duration_, width_, height_, format_, sample_offsets_, sample_sizes_, keyframe_indices_, extradata_);
}

}  // namespace hwang
