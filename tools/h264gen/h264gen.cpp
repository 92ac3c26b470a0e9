// Synthetic H.264 stream generator: seeded content model -> closed-loop encoder (all prediction,
// transform and in-loop filter arithmetic is the decode core's own host build, so the encoder's
// reconstruction is what a conforming decoder must output) -> MP4.  See h264gen.h.
#include "h264gen.h"

#include <math.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../hwang_b200/csrc/dev/deblock.h"
#include "../../hwang_b200/csrc/dev/recon.h"
#include "entropy_enc.h"

namespace gen {

static thread_local std::string g_err;

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) { next(); next(); }
  uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); }
  int below(int n) { return n <= 1 ? 0 : (int)(next() % (uint32_t)n); }
  bool pct(int p) { return below(100) < p; }
  int range(int lo, int hi) { return lo + below(hi - lo + 1); }
};

// ------------------------------------------------------------------------------------------ content
struct Content {
  int W, H, TW, TH;
  std::vector<uint8_t> tex[3];  // background textures (luma full, chroma half)
  int pan_vx, pan_vy;           // quarter luma samples per frame
  struct Obj { int w, h, x0, y0, vx, vy, tx, ty, add; };
  std::vector<Obj> objs;
  int fade;  // 1: global brightness ramp (for weighted prediction)

  void init(int w, int h, uint32_t seed, int fade_) {
    W = w; H = h; fade = fade_;
    TW = 1; while (TW < w + 128) TW <<= 1;
    TH = 1; while (TH < h + 128) TH <<= 1;
    Rng r(seed * 77 + 5);
    for (int p = 0; p < 3; ++p) {
      int tw = p ? TW / 2 : TW, th = p ? TH / 2 : TH;
      int gs = p ? 16 : 24;  // coarse grid step
      int gw = tw / gs + 2, gh = th / gs + 2;
      std::vector<int> g(gw * gh);
      for (auto &v : g) v = p ? r.range(96, 160) : r.range(40, 215);
      // make the coarse grid wrap
      for (int y = 0; y < gh; ++y) { g[y * gw + gw - 2] = g[y * gw]; g[y * gw + gw - 1] = g[y * gw + 1]; }
      for (int x = 0; x < gw; ++x) { g[(gh - 2) * gw + x] = g[x]; g[(gh - 1) * gw + x] = g[gw + x]; }
      tex[p].resize((size_t)tw * th);
      for (int y = 0; y < th; ++y)
        for (int x = 0; x < tw; ++x) {
          int gx = x / gs, gy = y / gs, fx = x % gs, fy = y % gs;
          int a = g[gy * gw + gx], b = g[gy * gw + gx + 1], c = g[(gy + 1) * gw + gx], d = g[(gy + 1) * gw + gx + 1];
          int v = ((a * (gs - fx) + b * fx) * (gs - fy) + (c * (gs - fx) + d * fx) * fy) / (gs * gs);
          tex[p][(size_t)y * tw + x] = (uint8_t)v;
        }
      if (!p) {
        // fine grain + a few hard-edged patches so that deblocking / directional intra modes get exercised
        for (auto &v : tex[p]) v = (uint8_t)clip8(v + r.range(-3, 3));
        for (int k = 0; k < 40; ++k) {
          int pw = r.range(8, 80), ph = r.range(8, 80), px = r.below(tw - pw), py = r.below(th - ph), dv = r.range(-50, 50);
          int slope = r.range(-2, 2);
          for (int y = 0; y < ph; ++y) for (int x = 0; x < pw; ++x) {
            uint8_t &t = tex[p][(size_t)(py + y) * tw + px + x];
            t = (uint8_t)clip8(t + dv + ((slope * (x + y)) >> 2));
          }
        }
      }
    }
    pan_vx = r.range(-9, 9); pan_vy = r.range(-5, 5);
    int nobj = 3 + r.below(5);
    for (int k = 0; k < nobj; ++k) {
      Obj o;
      o.w = r.range(w / 12 + 8, w / 4 + 8); o.h = r.range(h / 12 + 8, h / 4 + 8);
      o.x0 = r.below(std::max(1, w - o.w)) * 4; o.y0 = r.below(std::max(1, h - o.h)) * 4;
      o.vx = r.range(-22, 22); o.vy = r.range(-14, 14);
      o.tx = r.below(TW) * 4; o.ty = r.below(TH) * 4; o.add = r.range(-40, 40);
      objs.push_back(o);
    }
  }
  static int bil(const uint8_t *t, int tw, int th, int qx, int qy) {  // quarter-sample coords, wrap
    int x = qx >> 2, y = qy >> 2, fx = qx & 3, fy = qy & 3;
    int x0 = x & (tw - 1), x1 = (x + 1) & (tw - 1), y0 = y & (th - 1), y1 = (y + 1) & (th - 1);
    int a = t[(size_t)y0 * tw + x0], b = t[(size_t)y0 * tw + x1], c = t[(size_t)y1 * tw + x0], d = t[(size_t)y1 * tw + x1];
    return ((a * (4 - fx) + b * fx) * (4 - fy) + (c * (4 - fx) + d * fx) * fy + 8) >> 4;
  }
  int obj_at(int t, int x, int y) const {
    for (int k = (int)objs.size() - 1; k >= 0; --k) {
      const Obj &o = objs[k];
      int ox = (o.x0 + t * o.vx) >> 2, oy = (o.y0 + t * o.vy) >> 2;
      ox = ((ox % (W + o.w)) + (W + o.w)) % (W + o.w) - o.w;  // wrap around the picture
      oy = ((oy % (H + o.h)) + (H + o.h)) % (H + o.h) - o.h;
      if (x >= ox && x < ox + o.w && y >= oy && y < oy + o.h) return k;
    }
    return -1;
  }
  // motion (quarter samples) from a picture at time t to the picture at time t-d
  void motion(int t, int x, int y, int d, int &mx, int &my) const {
    int k = obj_at(t, x, y);
    if (k < 0) { mx = d * pan_vx; my = d * pan_vy; }
    else { mx = -d * objs[k].vx; my = -d * objs[k].vy; }
  }
  int fade_add(int t) const { return fade ? (int)lround(40.0 * sin(t * 0.11)) : 0; }
  // render into coded-size planes (edges replicated)
  void render(int t, int wc, int hc, uint8_t *Y, uint8_t *U, uint8_t *V) const {
    int fa = fade_add(t);
    for (int y = 0; y < hc; ++y) {
      int sy = std::min(y, H - 1);
      for (int x = 0; x < wc; ++x) {
        int sx = std::min(x, W - 1);
        int k = obj_at(t, sx, sy), v;
        if (k < 0) v = bil(tex[0].data(), TW, TH, sx * 4 + t * pan_vx, sy * 4 + t * pan_vy);
        else { const Obj &o = objs[k]; v = bil(tex[0].data(), TW, TH, sx * 4 - t * o.vx + o.tx, sy * 4 - t * o.vy + o.ty) + o.add; }
        Y[(size_t)y * wc + x] = (uint8_t)clip8(v + fa);
      }
    }
    int cw = wc / 2, ch = hc / 2;
    for (int y = 0; y < ch; ++y) {
      int sy = std::min(y, H / 2 - 1);
      for (int x = 0; x < cw; ++x) {
        int sx = std::min(x, W / 2 - 1);
        int k = obj_at(t, sx * 2, sy * 2);
        int qx, qy;
        if (k < 0) { qx = sx * 4 + ((t * pan_vx) >> 1); qy = sy * 4 + ((t * pan_vy) >> 1); }
        else { const Obj &o = objs[k]; qx = sx * 4 - ((t * o.vx) >> 1) + (o.tx >> 1); qy = sy * 4 - ((t * o.vy) >> 1) + (o.ty >> 1); }
        U[(size_t)y * cw + x] = (uint8_t)bil(tex[1].data(), TW / 2, TH / 2, qx, qy);
        V[(size_t)y * cw + x] = (uint8_t)bil(tex[2].data(), TW / 2, TH / 2, qx, qy);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------ quantiser
// Orthogonal-projection quantiser: the pixel-domain response of each coefficient position is taken
// from a float copy of the inverse transform, so it works unchanged for 4x4 / 8x8 and any scaling list.
struct Basis {
  float g4[16][16], n4[16];
  float g8[64][64], n8[64];
  Basis() {
    for (int k = 0; k < 16; ++k) {
      float d[16] = {0}; d[k] = 1.f;
      for (int r = 0; r < 4; ++r) i4(d + 4 * r, 1);
      for (int c = 0; c < 4; ++c) i4(d + c, 4);
      n4[k] = 0;
      for (int i = 0; i < 16; ++i) { g4[k][i] = d[i] / 64.f; n4[k] += g4[k][i] * g4[k][i]; }
    }
    for (int k = 0; k < 64; ++k) {
      float d[64] = {0}; d[k] = 1.f;
      for (int r = 0; r < 8; ++r) i8(d + 8 * r, 1);
      for (int c = 0; c < 8; ++c) i8(d + c, 8);
      n8[k] = 0;
      for (int i = 0; i < 64; ++i) { g8[k][i] = d[i] / 64.f; n8[k] += g8[k][i] * g8[k][i]; }
    }
  }
  static void i4(float *d, int s) {
    float e0 = d[0] + d[2 * s], e1 = d[0] - d[2 * s], e2 = d[s] * 0.5f - d[3 * s], e3 = d[s] + d[3 * s] * 0.5f;
    d[0] = e0 + e3; d[s] = e1 + e2; d[2 * s] = e1 - e2; d[3 * s] = e0 - e3;
  }
  static void i8(float *d, int s) {
    float a0 = d[0] + d[4 * s], a4 = d[0] - d[4 * s], a2 = d[2 * s] * .5f - d[6 * s], a6 = d[2 * s] + d[6 * s] * .5f;
    float a1 = -d[3 * s] + d[5 * s] - d[7 * s] - d[7 * s] * .5f, a3 = d[s] + d[7 * s] - d[3 * s] - d[3 * s] * .5f;
    float a5 = -d[s] + d[7 * s] + d[5 * s] + d[5 * s] * .5f, a7 = d[3 * s] + d[5 * s] + d[s] + d[s] * .5f;
    float b0 = a0 + a6, b2 = a4 + a2, b4 = a4 - a2, b6 = a0 - a6;
    float b1 = a1 + a7 * .25f, b7 = a7 - a1 * .25f, b3 = a3 + a5 * .25f, b5 = a3 * .25f - a5;
    d[0] = b0 + b7; d[s] = b2 + b5; d[2 * s] = b4 + b3; d[3 * s] = b6 + b1;
    d[4 * s] = b6 - b1; d[5 * s] = b4 - b3; d[6 * s] = b2 - b5; d[7 * s] = b0 - b7;
  }
};
static const Basis g_basis;

static inline int qround(float x, float rnd) {
  int v = (int)(fabsf(x) + rnd);
  if (v > 1200) v = 1200;
  return x < 0 ? -v : v;
}

// ------------------------------------------------------------------------------------------ encoder
enum { NSLOT = 12, MAXSL = 8 };
enum { kPoc1NonRefOffset = -1 };

struct RefEntry { int slot, frame_num, poc, t; bool long_term = false; int lt_idx = 0; };
struct MmcoOp { int op, a; };

struct Encoder {
  hwgen_params P;
  // bits of frame_num and of pic_order_cnt_lsb (header_variant: 6 and 5, so that both wrap inside a GOP)
  int fn_bits() const { return P.header_variant ? 6 : 4; }
  int poc_bits() const { return P.header_variant ? 5 : 8; }
  int W, H, wc, hc, mb_w, mb_h, nmb;
  bool cabac, high;
  const Content *content;
  ChunkCtx c;
  std::vector<uint8_t> frames, ectx;
  std::vector<MbInfo> mbinfo;
  std::vector<int16_t> mv, refpic, coefs;
  std::vector<int8_t> refidx;
  std::vector<PicDesc> pics;
  std::vector<SliceDesc> slices;
  std::vector<int32_t> prog;
  int32_t errflag = 0;
  std::vector<uint8_t> srcY, srcU, srcV;
  std::vector<RefEntry> dpb;
  std::vector<MmcoOp> rplm_ops[2];  // ref_pic_list_modification of the picture being encoded (op = modification_of_pic_nums_idc)
  Rng rng;
  ReconScratch rs;
  DeblockScratch ds;
  uint8_t scaling4[6][16], scaling8[2][64];
  bool custom_scaling = false;

  explicit Encoder(const hwgen_params &p, const Content *ct) : P(p), content(ct), rng(p.seed) {
    W = p.width; H = p.height; wc = (W + 15) & ~15; hc = (H + 15) & ~15; mb_w = wc / 16; mb_h = hc / 16; nmb = mb_w * mb_h;
    high = p.profile == 2;
    cabac = p.cabac < 0 ? p.profile >= 1 : p.cabac != 0;
    size_t fs = (size_t)wc * hc * 3 / 2;
    frames.resize(fs * NSLOT); mbinfo.resize((size_t)nmb * NSLOT); mv.resize((size_t)NSLOT * 2 * nmb * 32);
    refidx.resize((size_t)NSLOT * 2 * nmb * 4); refpic.resize((size_t)NSLOT * 2 * nmb * 4);
    coefs.resize((size_t)NSLOT * nmb * SLOTS_PER_MB * 16);
    pics.resize(NSLOT); slices.resize(NSLOT * MAXSL); prog.resize(NSLOT * (MAXSL + 2 * mb_h));
    ectx.resize((size_t)NSLOT * MAXSL * mb_w * sizeof(NbCtx));
    memset(&c, 0, sizeof(c));
    c.mb_w = mb_w; c.mb_h = mb_h; c.nmb = nmb; c.nmb_stride = nmb; c.wc = wc; c.hc = hc; c.num_pics = NSLOT; c.num_slices = NSLOT * MAXSL;
    c.frames = frames.data(); c.frame_stride = fs; c.mbinfo = mbinfo.data(); c.mv = mv.data(); c.refidx = refidx.data();
    c.refpic = refpic.data(); c.coefs = coefs.data(); c.ectx = ectx.data(); c.ectx_stride = (uint64_t)mb_w * sizeof(NbCtx);
    c.bitstream = nullptr; c.pics = pics.data(); c.slices = slices.data(); c.entropy_prog = prog.data();
    c.recon_prog = prog.data() + NSLOT * MAXSL; c.dbl_prog = c.recon_prog + NSLOT * mb_h; c.error_flag = &errflag;
    srcY.resize((size_t)wc * hc); srcU.resize((size_t)wc * hc / 4); srcV.resize((size_t)wc * hc / 4);
    for (int i = 0; i < 6; ++i) for (int k = 0; k < 16; ++k) scaling4[i][k] = 16;
    for (int i = 0; i < 2; ++i) for (int k = 0; k < 64; ++k) scaling8[i][k] = 16;
    if (high && p.scaling_lists) {
      custom_scaling = true;
      Rng r(p.seed + 991);
      for (int i = 0; i < 6; ++i) for (int k = 0; k < 16; ++k) scaling4[i][k] = (uint8_t)(12 + (k & 3) * 2 + (k >> 2) * 2 + r.below(5) + (i >= 3 ? 2 : 0));
      for (int i = 0; i < 2; ++i) for (int k = 0; k < 64; ++k) scaling8[i][k] = (uint8_t)(10 + (k & 7) * 2 + (k >> 3) * 2 + r.below(4) + i * 3);
    }
  }

  // -------------------------------------------------------------------------------- parameter sets
  std::vector<uint8_t> sps_rbsp() const {
    BitWriter b;
    int prof = P.profile == 0 ? 66 : (P.profile == 1 ? 77 : 100);
    b.put((uint32_t)prof, 8);
    b.put(P.profile == 0 ? 0xC0 : (P.profile == 1 ? 0x40 : 0x00), 8);
    int level = (wc * hc <= 720 * 576) ? 30 : (wc * hc <= 1280 * 720 ? 31 : (wc * hc <= 1920 * 1088 ? 40 : 51));
    b.put((uint32_t)level, 8);
    b.ue(P.header_variant ? 3 : 0);  // seq_parameter_set_id
    if (prof == 100) { b.ue(1); b.ue(0); b.ue(0); b.put1(0); b.put1(0); }
    b.ue((uint32_t)(fn_bits() - 4));  // log2_max_frame_num_minus4
    b.ue((uint32_t)P.poc_type);
    if (P.poc_type == 0) b.ue((uint32_t)(poc_bits() - 4));  // log2_max_pic_order_cnt_lsb_minus4
    if (P.poc_type == 1) {
      // delta_pic_order_always_zero_flag (only legal here when POC follows decode order), offset_for_non_ref_pic,
      // offset_for_top_to_bottom_field, one-entry cycle: a reference frame advances the expected POC by 2
      b.put1(P.bframes == 0); b.se(kPoc1NonRefOffset); b.se(0); b.ue(1); b.se(2);
    }
    b.ue((uint32_t)P.num_ref);
    b.put1(0);
    b.ue((uint32_t)(mb_w - 1)); b.ue((uint32_t)(mb_h - 1));
    b.put1(1);  // frame_mbs_only
    b.put1(P.direct_4x4 ? 0 : 1);  // direct_8x8_inference
    bool crop = wc != W || hc != H;
    b.put1(crop);
    if (crop) { b.ue(0); b.ue((uint32_t)((wc - W) / 2)); b.ue(0); b.ue((uint32_t)((hc - H) / 2)); }
    b.put1(1);  // vui
    b.put1(0); b.put1(0); b.put1(0); b.put1(0);  // aspect, overscan, video_signal, chroma_loc
    b.put1(0);  // timing
    b.put1(0); b.put1(0);  // hrd
    b.put1(0);  // pic_struct
    b.put1(1);  // bitstream_restriction
    b.put1(1); b.ue(0); b.ue(0); b.ue(16); b.ue(16);
    const int reorder = P.bframes > 0 ? (P.b_pyramid && P.bframes >= 2 ? 2 : 1) : 0;
    b.ue((uint32_t)reorder);
    b.ue((uint32_t)std::max(P.num_ref, reorder + 1));
    b.trailing();
    return b.buf;
  }
  static void write_scaling_list(BitWriter &b, const uint8_t *raster, int n) {
    const uint8_t *zz = n == 16 ? zigzag4x4 : zigzag8x8;
    int last = 8;
    for (int i = 0; i < n; ++i) {
      int v = raster[zz[i]];
      int delta = v - last;
      if (delta > 127) delta -= 256;
      if (delta < -128) delta += 256;
      b.se(delta);
      last = v;
    }
  }
  std::vector<uint8_t> pps_rbsp() const {
    BitWriter b;
    b.ue(P.header_variant ? 7 : 0); b.ue(P.header_variant ? 3 : 0);  // pic_parameter_set_id, seq_parameter_set_id
    b.put1(cabac);
    b.put1(0);
    b.ue(0);
    b.ue(P.header_variant ? 1 : 0); b.ue(P.header_variant ? 1 : 0);  // num_ref_idx_default_active_minus1
    b.put1(P.weighted >= 1);
    b.put(P.weighted == 3 ? 1 : (P.weighted == 2 ? 2 : 0), 2);
    b.se(P.header_variant ? -4 : 0); b.se(0);  // pic_init_qp_minus26, pic_init_qs_minus26
    b.se(P.chroma_qp_offset);
    b.put1(1);  // deblocking_filter_control_present
    b.put1(P.constrained_intra != 0);
    b.put1(0);
    if (high) {
      b.put1(1);  // transform_8x8_mode
      b.put1(custom_scaling);
      if (custom_scaling) {
        for (int i = 0; i < 6; ++i) { b.put1(1); write_scaling_list(b, scaling4[i], 16); }
        for (int i = 0; i < 2; ++i) { b.put1(1); write_scaling_list(b, scaling8[i], 64); }
      }
      b.se(P.chroma_qp_offset - 1);
    }
    b.trailing();
    return b.buf;
  }

  // -------------------------------------------------------------------------------- helpers
  // Least recently used free slot: motion data of pictures still in the DPB names its references by slot (refpic), so a
  // slot must not be handed out again while such a picture can still be the co-located picture of a B slice (a
  // reference dropped by MMCO 1 and its slot reused at once made temporal direct prediction find the wrong picture).
  int slot_stamp[NSLOT] = {0};
  int stamp = 0;
  int free_slot() {
    int best = -1;
    for (int s = 0; s < NSLOT; ++s) {
      bool used = false;
      for (auto &r : dpb) if (r.slot == s) used = true;
      if (!used && (best < 0 || slot_stamp[s] < slot_stamp[best])) best = s;
    }
    if (best >= 0) {
      slot_stamp[best] = ++stamp;
      // ... and least-recently-used is not enough with a B pyramid and four references (found by tools/param_sweep.py): a
      // picture still in the DPB may name a reference that left long ago.  Whatever still names the slot that is handed
      // out now is made unmatchable, so that temporal direct prediction finds "no such picture in list 0" (index 0) exactly
      // as a decoder does, which knows pictures by identity and not by buffer.
      for (auto &r : dpb)
        for (int l = 0; l < 2; ++l) {
          int16_t *po = pic_refpic(c, r.slot, l);
          for (int i = 0; i < nmb * 4; ++i) if (po[i] == best) po[i] = -2;
        }
    }
    return best;
  }
  const uint8_t *src_plane(int p) const { return p == 0 ? srcY.data() : (p == 1 ? srcU.data() : srcV.data()); }

  struct PicState {
    int slot, t, type, frame_num, poc, gop_t0;
    bool is_ref, idr;
    int delta_poc = 0;            // poc_type 1: delta_pic_order_cnt[0]
    bool long_term_idr = false;   // IDR with long_term_reference_flag
    std::vector<MmcoOp> mmco;     // adaptive_ref_pic_marking_mode_flag = 1 when not empty
    int mmco3_lt_idx = 0;         // long_term_frame_idx of MMCO 3 / 6
    bool rplm = false;            // the slices reorder their reference lists
  };

  // quantise the residual of the current macroblock (source minus what recon_mb predicted into the
  // frame buffer) into arena slots.  Returns cbp.  `pred_*` are the prediction samples.
  struct Levels {
    int16_t dc[16]; bool has_dc;
    int16_t luma[16][16]; bool nz[16];       // z order, raster coefficients (4x4 transform)
    int16_t luma8[4][64]; bool nz8[4];
    int16_t cdc[2][4]; bool has_cdc[2];
    int16_t cac[2][4][16]; bool cnz[2][4];
  };

  void quant_luma(const int *res /*256 raster*/, bool i16, bool t8, bool intra, int qp, Levels &L) const {
    const float rnd = intra ? 0.36f : 0.22f;
    const uint8_t *sc4 = scaling4[intra ? 0 : 3];
    const int qm = qp % 6, qs = qp / 6;
    L.has_dc = false;
    for (int i = 0; i < 16; ++i) { L.nz[i] = false; L.dc[i] = 0; }
    for (int i = 0; i < 4; ++i) L.nz8[i] = false;
    if (t8) {
      const uint8_t *sc8 = scaling8[intra ? 0 : 1];
      for (int q = 0; q < 4; ++q) {
        float x[64];
        for (int i = 0; i < 64; ++i) x[i] = (float)res[((q >> 1) * 8 + (i >> 3)) * 16 + (q & 1) * 8 + (i & 7)];
        bool any = false;
        for (int k = 0; k < 64; ++k) {
          float d = 0;
          for (int i = 0; i < 64; ++i) d += x[i] * g_basis.g8[k][i];
          d /= g_basis.n8[k];
          float scale = (float)sc8[k] * dequant8_v[qm * 64 + k] * ldexpf(1.f, qs - 6);
          int lv = qround(d / scale, rnd);
          L.luma8[q][k] = (int16_t)lv; any |= lv != 0;
        }
        L.nz8[q] = any;
      }
      return;
    }
    float dcm[16];
    for (int z = 0; z < 16; ++z) {
      int bx = z2x(z) * 4, by = z2y(z) * 4;
      float x[16];
      for (int i = 0; i < 16; ++i) x[i] = (float)res[(by + (i >> 2)) * 16 + bx + (i & 3)];
      bool any = false;
      for (int k = 0; k < 16; ++k) {
        float d = 0;
        for (int i = 0; i < 16; ++i) d += x[i] * g_basis.g4[k][i];
        d /= g_basis.n4[k];
        if (i16 && k == 0) { dcm[z2y(z) * 4 + z2x(z)] = d; L.luma[z][0] = 0; continue; }
        float scale = (float)sc4[k] * dequant4_v[qm * 16 + k] * ldexpf(1.f, qs - 4);
        int lv = qround(d / scale, rnd);
        L.luma[z][k] = (int16_t)lv; any |= lv != 0;
      }
      L.nz[z] = any;
    }
    if (i16) {
      float ls = (float)sc4[0] * dequant4_v[qm * 16] * ldexpf(1.f, qs - 6);
      float f[16], t[16];
      for (int i = 0; i < 16; ++i) f[i] = dcm[i] / ls;
      static const int Hm[4][4] = {{1, 1, 1, 1}, {1, 1, -1, -1}, {1, -1, -1, 1}, {1, -1, 1, -1}};
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float a = 0; for (int k = 0; k < 4; ++k) a += Hm[i][k] * f[k * 4 + j]; t[i * 4 + j] = a; }
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
        float a = 0; for (int k = 0; k < 4; ++k) a += t[i * 4 + k] * Hm[k][j];
        int lv = qround(a / 16.f, rnd);
        L.dc[i * 4 + j] = (int16_t)lv; L.has_dc |= lv != 0;
      }
    }
  }
  void quant_chroma(const int *res /*64 raster*/, int pl, bool intra, int qpc, Levels &L) const {
    const float rnd = intra ? 0.36f : 0.22f;
    const uint8_t *sc4 = scaling4[(intra ? 0 : 3) + 1 + pl];
    const int qm = qpc % 6, qs = qpc / 6;
    float dcm[4];
    for (int b = 0; b < 4; ++b) {
      int bx = (b & 1) * 4, by = (b >> 1) * 4;
      float x[16];
      for (int i = 0; i < 16; ++i) x[i] = (float)res[(by + (i >> 2)) * 8 + bx + (i & 3)];
      bool any = false;
      for (int k = 0; k < 16; ++k) {
        float d = 0;
        for (int i = 0; i < 16; ++i) d += x[i] * g_basis.g4[k][i];
        d /= g_basis.n4[k];
        if (k == 0) { dcm[b] = d; L.cac[pl][b][0] = 0; continue; }
        float scale = (float)sc4[k] * dequant4_v[qm * 16 + k] * ldexpf(1.f, qs - 4);
        int lv = qround(d / scale, rnd);
        L.cac[pl][b][k] = (int16_t)lv; any |= lv != 0;
      }
      L.cnz[pl][b] = any;
    }
    float ls = (float)sc4[0] * dequant4_v[qm * 16] * ldexpf(1.f, qs) / 32.f;
    float f[4] = {dcm[0] / ls, dcm[1] / ls, dcm[2] / ls, dcm[3] / ls};
    float cc[4] = {f[0] + f[1] + f[2] + f[3], f[0] - f[1] + f[2] - f[3], f[0] + f[1] - f[2] - f[3], f[0] - f[1] - f[2] + f[3]};
    L.has_cdc[pl] = false;
    for (int i = 0; i < 4; ++i) { int lv = qround(cc[i] / 4.f, rnd); L.cdc[pl][i] = (int16_t)lv; L.has_cdc[pl] |= lv != 0; }
  }

  // Store levels into the arena in IR slot order; fills o.nzmask / o.cbp, advances coef_next.
  void emit_levels(SliceDec &s, MbInfo &o, bool i16, bool t8, Levels &L) {
    int16_t *dst = pic_coefs(c, s.pd->frame) + (uint64_t)s.coef_next * 16;
    int n = 0;
    uint32_t nz = 0;
    int cbp = 0;
    auto push = [&](const int16_t *src, int cnt, int bit, int nb) { memcpy(dst + n * 16, src, cnt * sizeof(int16_t)); if (cnt < nb * 16) memset(dst + n * 16 + cnt, 0, (nb * 16 - cnt) * 2); n += nb; nz |= ((1u << nb) - 1) << bit; };
    if (i16) {
      bool anyac = false;
      for (int z = 0; z < 16; ++z) anyac |= L.nz[z];
      if (L.has_dc) push(L.dc, 16, NZ_LUMA_DC, 1);
      if (anyac) { cbp |= 15; for (int z = 0; z < 16; ++z) if (L.nz[z]) push(L.luma[z], 16, NZ_LUMA0 + z, 1); }
    } else if (t8) {
      for (int q = 0; q < 4; ++q) if (L.nz8[q]) { cbp |= 1 << q; push(L.luma8[q], 64, NZ_LUMA0 + q * 4, 4); }
    } else {
      for (int q = 0; q < 4; ++q) {
        bool any = false;
        for (int k = 0; k < 4; ++k) any |= L.nz[q * 4 + k];
        if (!any) continue;
        cbp |= 1 << q;
        for (int k = 0; k < 4; ++k) if (L.nz[q * 4 + k]) push(L.luma[q * 4 + k], 16, NZ_LUMA0 + q * 4 + k, 1);
      }
    }
    bool cac = false, cdc = L.has_cdc[0] || L.has_cdc[1];
    for (int p = 0; p < 2; ++p) for (int b = 0; b < 4; ++b) cac |= L.cnz[p][b];
    if (cac) cbp |= 0x20; else if (cdc) cbp |= 0x10;
    if (cac || cdc) {
      for (int p = 0; p < 2; ++p) if (L.has_cdc[p]) push(L.cdc[p], 4, p ? NZ_CR_DC : NZ_CB_DC, 1);
      if (cac) for (int p = 0; p < 2; ++p) for (int b = 0; b < 4; ++b) if (L.cnz[p][b]) push(L.cac[p][b], 16, (p ? NZ_CR0 : NZ_CB0) + b, 1);
    }
    o.nzmask = nz; o.cbp = (uint8_t)cbp;
    s.coef_next += n;
  }

  // residual between the source and the current frame-buffer content of the macroblock
  void mb_residual(int slot, int mbx, int mby, int *ry, int *ru, int *rv) const {
    const uint8_t *Y = frame_y(c, slot), *U = frame_cb(c, slot), *V = frame_cr(c, slot);
    for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x) { size_t o = (size_t)(mby * 16 + y) * wc + mbx * 16 + x; ry[y * 16 + x] = srcY[o] - Y[o]; }
    for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) { size_t o = (size_t)(mby * 8 + y) * (wc / 2) + mbx * 8 + x; ru[y * 8 + x] = srcU[o] - U[o]; rv[y * 8 + x] = srcV[o] - V[o]; }
  }

  // ------------------------------------------------------------------------------ intra decisions
  struct IntraAvail { bool l, t, tr, tl; };
  IntraAvail intra_avail(const SliceDec &s, int slot) const {
    const MbInfo *mbs = pic_mbinfo(c, slot);
    auto ok = [&](bool av, int addr) { return av && !(P.constrained_intra && mbs[addr].mbtype == MB_INTER); };
    IntraAvail a;
    a.l = ok(s.availA, s.mbaddr - 1); a.t = ok(s.availB, s.mbaddr - mb_w);
    a.tr = ok(s.availC, s.mbaddr - mb_w + 1); a.tl = ok(s.availD, s.mbaddr - mb_w - 1);
    return a;
  }
  int pick_i16_mode(int slot, int mbx, int mby, IntraAvail a) {
    const uint8_t *Y = frame_y(c, slot);
    int best = 2, bests = 1 << 30;
    for (int mode = 0; mode < 4; ++mode) {
      if ((mode == 0 && !a.t) || (mode == 1 && !a.l) || (mode == 3 && !(a.t && a.l && a.tl))) continue;
      int sad = 0;
      int dcv = 128, pa = 0, pb = 0, pc = 0;
      const uint8_t *T = Y + (size_t)(mby * 16 - 1) * wc + mbx * 16;
      const uint8_t *Lc = Y + (size_t)(mby * 16) * wc + mbx * 16 - 1;
      if (mode == 2) {
        int sm = 0;
        if (a.t) for (int i = 0; i < 16; ++i) sm += T[i];
        if (a.l) for (int i = 0; i < 16; ++i) sm += Lc[(size_t)i * wc];
        dcv = (a.t && a.l) ? (sm + 16) >> 5 : ((a.t || a.l) ? (sm + 8) >> 4 : 128);
      } else if (mode == 3) {
        int Hh = 0, Vv = 0;
        for (int i = 0; i < 8; ++i) { Hh += (i + 1) * (T[8 + i] - T[6 - i]); Vv += (i + 1) * (Lc[(size_t)(8 + i) * wc] - (i == 7 ? T[-1] : Lc[(size_t)(6 - i) * wc])); }
        pa = 16 * (Lc[(size_t)15 * wc] + T[15]); pb = (5 * Hh + 32) >> 6; pc = (5 * Vv + 32) >> 6;
      }
      for (int y = 0; y < 16; y += 2) for (int x = 0; x < 16; x += 2) {
        int v = mode == 0 ? T[x] : mode == 1 ? Lc[(size_t)y * wc] : mode == 2 ? dcv : clip8((pa + pb * (x - 7) + pc * (y - 7) + 16) >> 5);
        sad += abs(v - srcY[(size_t)(mby * 16 + y) * wc + mbx * 16 + x]);
      }
      if (sad < bests) { bests = sad; best = mode; }
    }
    return best;
  }
  int pick_chroma_mode(IntraAvail a) {
    int cand[4], n = 0;
    cand[n++] = 0;
    if (a.l) cand[n++] = 1;
    if (a.t) cand[n++] = 2;
    if (a.l && a.t && a.tl) cand[n++] = 3;
    return cand[rng.below(n)];
  }

  // Intra NxN: sequential mode decision + quantisation on a private copy of the tile; the official
  // reconstruction is then redone by recon_mb from the IR (and must agree).
  void decide_inxn(SliceDec &s, int slot, int mbx, int mby, bool t8, int qp, IntraAvail av, MbEnc &m, Levels &L) {
    const int N = t8 ? 8 : 4, nblk = t8 ? 4 : 16;
    uint8_t tile[17][32];  // rows -1..15, cols -1..24 at [r+1][c+1]
    const uint8_t *Y = frame_y(c, slot);
    memset(tile, 128, sizeof(tile));
    for (int cc = -1; cc < 24; ++cc) { int gx = mbx * 16 + cc; if (mby > 0 && gx >= 0 && gx < wc) tile[0][cc + 1] = Y[(size_t)(mby * 16 - 1) * wc + gx]; }
    for (int r = 0; r < 16; ++r) if (mbx > 0) tile[r + 1][0] = Y[(size_t)(mby * 16 + r) * wc + mbx * 16 - 1];
    const uint8_t *sc4 = scaling4[0];
    const uint8_t *sc8 = scaling8[0];
    for (int i = 0; i < 16; ++i) L.nz[i] = false;
    for (int i = 0; i < 4; ++i) L.nz8[i] = false;
    L.has_dc = false;
    for (int blk = 0; blk < nblk; ++blk) {
      const int bx = N == 8 ? (blk & 1) * 8 : z2x(blk) * 4, by = N == 8 ? (blk >> 1) * 8 : z2y(blk) * 4;
      const bool aL = bx > 0 || av.l, aT = by > 0 || av.t;
      const bool aD = (bx > 0 && by > 0) ? true : (bx > 0 ? av.t : (by > 0 ? av.l : av.tl));
      bool aC;
      if (by == 0) aC = (bx + N < 16) ? av.t : av.tr;
      else if (bx + N >= 16) aC = false;
      else aC = N == 8 ? (blk == 2) : (xy2z((bx >> 2) + 1, (by >> 2) - 1) < blk);
      uint8_t E[26], F[26];
      for (int i = 0; i < N; ++i) E[N - 1 - i] = aL ? tile[by + i + 1][bx] : 128;
      for (int i = 0; i < N; ++i) E[N + 1 + i] = aT ? tile[by][bx + 1 + i] : 128;
      for (int i = 0; i < N; ++i) E[2 * N + 1 + i] = aC ? tile[by][bx + 1 + N + i] : E[2 * N];
      E[N] = aD ? tile[by][bx] : 128;
      E[3 * N + 1] = E[3 * N];
      if (N == 8) {
        if (aT) { F[9] = aD ? (E[8] + 2 * E[9] + E[10] + 2) >> 2 : (3 * E[9] + E[10] + 2) >> 2; for (int i = 1; i < 15; ++i) F[9 + i] = (E[8 + i] + 2 * E[9 + i] + E[10 + i] + 2) >> 2; F[24] = (E[23] + 3 * E[24] + 2) >> 2; }
        else for (int i = 9; i < 25; ++i) F[i] = E[i];
        if (aD) { if (aT && aL) F[8] = (E[9] + 2 * E[8] + E[7] + 2) >> 2; else if (aT) F[8] = (3 * E[8] + E[9] + 2) >> 2; else if (aL) F[8] = (3 * E[8] + E[7] + 2) >> 2; else F[8] = E[8]; }
        else F[8] = E[8];
        if (aL) { F[7] = aD ? (E[8] + 2 * E[7] + E[6] + 2) >> 2 : (3 * E[7] + E[6] + 2) >> 2; for (int i = 1; i < 7; ++i) F[7 - i] = (E[8 - i] + 2 * E[7 - i] + E[6 - i] + 2) >> 2; F[0] = (E[1] + 3 * E[0] + 2) >> 2; }
        else for (int i = 0; i < 8; ++i) F[i] = E[i];
        F[25] = F[24];
        memcpy(E, F, 26);
      }
      // candidate modes
      int best = 2, bests = 1 << 30;
      uint8_t bestp[64];
      for (int mode = 0; mode < 9; ++mode) {
        bool okm;
        switch (mode) {
          case 0: okm = aT; break; case 1: okm = aL; break; case 2: okm = true; break;
          case 3: case 7: okm = aT; break;
          case 8: okm = aL; break;
          default: okm = aT && aL && aD;
        }
        if (!okm) continue;
        int dcv = 128;
        if (mode == 2) {
          int sm = 0;
          if (aT) for (int i = 0; i < N; ++i) sm += E[N + 1 + i];
          if (aL) for (int i = 0; i < N; ++i) sm += E[i];
          int sh = N == 8 ? 3 : 2;
          dcv = (aT && aL) ? (sm + N) >> (sh + 1) : ((aT || aL) ? (sm + (N >> 1)) >> sh : 128);
        }
        uint8_t pr[64];
        int sad = rng.below(N * N / 2);  // noise so that near-ties exercise all modes
        for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) {
          int v = mode == 2 ? dcv : intra_dir_pred(mode, N, E, x, y);
          pr[y * N + x] = (uint8_t)v;
          sad += abs(v - srcY[(size_t)(mby * 16 + by + y) * wc + mbx * 16 + bx + x]);
        }
        if (sad < bests) { bests = sad; best = mode; memcpy(bestp, pr, N * N); }
      }
      m.i4modes[blk] = best;
      // residual -> levels -> local reconstruction
      float x[64];
      for (int y = 0; y < N; ++y) for (int xx = 0; xx < N; ++xx) x[y * N + xx] = (float)(srcY[(size_t)(mby * 16 + by + y) * wc + mbx * 16 + bx + xx] - bestp[y * N + xx]);
      const int qm = qp % 6, qs = qp / 6;
      int16_t res[64];
      bool any = false;
      if (N == 8) {
        for (int k = 0; k < 64; ++k) {
          float d = 0; for (int i = 0; i < 64; ++i) d += x[i] * g_basis.g8[k][i];
          d /= g_basis.n8[k];
          float scale = (float)sc8[k] * dequant8_v[qm * 64 + k] * ldexpf(1.f, qs - 6);
          int lv = qround(d / scale, 0.36f); L.luma8[blk][k] = (int16_t)lv; any |= lv != 0;
        }
        L.nz8[blk] = any;
        if (any) residual8x8(L.luma8[blk], sc8, qp, res);
      } else {
        for (int k = 0; k < 16; ++k) {
          float d = 0; for (int i = 0; i < 16; ++i) d += x[i] * g_basis.g4[k][i];
          d /= g_basis.n4[k];
          float scale = (float)sc4[k] * dequant4_v[qm * 16 + k] * ldexpf(1.f, qs - 4);
          int lv = qround(d / scale, 0.36f); L.luma[blk][k] = (int16_t)lv; any |= lv != 0;
        }
        L.nz[blk] = any;
        if (any) residual4x4(L.luma[blk], false, 0, sc4, qp, res);
      }
      for (int y = 0; y < N; ++y) for (int xx = 0; xx < N; ++xx)
        tile[by + y + 1][bx + xx + 1] = (uint8_t)(any ? clip8(bestp[y * N + xx] + res[y * N + xx]) : bestp[y * N + xx]);
    }
    (void)s;
  }

  // ------------------------------------------------------------------------------ picture
  void write_slice_header(BitWriter &b, const PicState &ps, const SliceDesc &sd, int first_mb, int nal_type, int idr_id,
                          int nrefs_total[2]) {
    (void)nrefs_total;
    b.ue((uint32_t)first_mb);
    // 5..7 promise that every slice of the picture has this type; a picture that mixes I slices in says 0..2
    const bool mixed = P.mixed_slices && P.slices >= 2 && ps.type != SLICE_I;
    b.ue((uint32_t)((sd.slice_type == SLICE_P ? 0 : sd.slice_type == SLICE_B ? 1 : 2) + (mixed ? 0 : 5)));
    b.ue(P.header_variant ? 7 : 0);  // pic_parameter_set_id
    b.put((uint32_t)(ps.frame_num & ((1 << fn_bits()) - 1)), fn_bits());
    if (nal_type == 5) b.ue((uint32_t)idr_id);
    if (P.poc_type == 0) b.put((uint32_t)(ps.poc & ((1 << poc_bits()) - 1)), poc_bits());
    if (P.poc_type == 1 && P.bframes != 0) b.se(ps.delta_poc);
    if (sd.slice_type == SLICE_B) b.put1(sd.direct_spatial);
    if (sd.slice_type != SLICE_I) {
      const int dflt = P.header_variant ? 2 : 1;  // the PPS's num_ref_idx_default_active
      bool ovr = sd.num_ref[0] != dflt || (sd.slice_type == SLICE_B && sd.num_ref[1] != dflt);
      b.put1(ovr);
      if (ovr) { b.ue((uint32_t)(sd.num_ref[0] - 1)); if (sd.slice_type == SLICE_B) b.ue((uint32_t)(sd.num_ref[1] - 1)); }
      for (int l = 0; l < (sd.slice_type == SLICE_B ? 2 : 1); ++l) {
        b.put1(!rplm_ops[l].empty());  // ref_pic_list_modification_flag_lX
        if (!rplm_ops[l].empty()) {
          for (auto &m : rplm_ops[l]) { b.ue((uint32_t)m.op); b.ue((uint32_t)m.a); }
          b.ue(3);
        }
      }
    }
    if (sd.use_weights == 1) {
      b.ue(sd.luma_log2_denom); b.ue(sd.chroma_log2_denom);
      for (int l = 0; l < (sd.slice_type == SLICE_B ? 2 : 1); ++l)
        for (int i = 0; i < sd.num_ref[l]; ++i) {
          bool lf = sd.luma_w[l][i] != (1 << sd.luma_log2_denom) || sd.luma_o[l][i] != 0;
          b.put1(lf);
          if (lf) { b.se(sd.luma_w[l][i]); b.se(sd.luma_o[l][i]); }
          bool cf = false;
          for (int k = 0; k < 2; ++k) cf |= sd.chroma_w[l][i][k] != (1 << sd.chroma_log2_denom) || sd.chroma_o[l][i][k] != 0;
          b.put1(cf);
          if (cf) for (int k = 0; k < 2; ++k) { b.se(sd.chroma_w[l][i][k]); b.se(sd.chroma_o[l][i][k]); }
        }
    }
    if (ps.is_ref) {
      if (nal_type == 5) { b.put1(0); b.put1(ps.long_term_idr); }
      else {
        b.put1(!ps.mmco.empty());  // adaptive_ref_pic_marking_mode_flag (0: sliding window)
        if (!ps.mmco.empty()) {
          for (auto &m : ps.mmco) { b.ue((uint32_t)m.op); if (m.op != 5) b.ue((uint32_t)m.a); if (m.op == 3) b.ue((uint32_t)ps.mmco3_lt_idx); }
          b.ue(0);
        }
      }
    }
    if (cabac && sd.slice_type != SLICE_I) b.ue(sd.cabac_init_idc);
    b.se((int)sd.qp - 26 - (P.header_variant ? -4 : 0));  // slice_qp_delta against pic_init_qp
    b.ue(sd.disable_deblock);
    if (sd.disable_deblock != 1) { b.se(sd.alpha_off / 2); b.se(sd.beta_off / 2); }
  }

  // Encode one picture; returns the AVCC sample.
  std::vector<uint8_t> encode_picture(const PicState &ps, int idr_id) {
    const int slot = ps.slot;
    content->render(ps.t, wc, hc, srcY.data(), srcU.data(), srcV.data());
    PicDesc &pd = pics[slot];
    memset(&pd, 0, sizeof(pd));
    pd.frame = slot; pd.first_slice = slot * MAXSL; pd.num_slices = P.slices; pd.poc = ps.poc;
    pd.cabac = cabac; pd.transform8x8_mode = high; pd.constrained_intra_pred = P.constrained_intra != 0;
    pd.direct_8x8_inference = P.direct_4x4 ? 0 : 1; pd.weighted_pred = P.weighted >= 1; pd.weighted_bipred_idc = P.weighted >= 2 ? 2 : 0;
    pd.is_ref = ps.is_ref; pd.has_inter = ps.type == SLICE_I ? 0 : (ps.type == SLICE_P ? 1 : 2);
    pd.chroma_qp_offset[0] = (int8_t)P.chroma_qp_offset; pd.chroma_qp_offset[1] = (int8_t)(high ? P.chroma_qp_offset - 1 : P.chroma_qp_offset);
    memcpy(pd.scaling4, scaling4, sizeof(scaling4)); memcpy(pd.scaling8, scaling8, sizeof(scaling8));

    // reference lists: initialisation (8.2.4.2: short-term by PicNum / POC, then long-term by LongTermPicNum), then the
    // optional modification (8.2.4.3) whose syntax write_slice_header emits from rplm_ops
    std::vector<RefEntry> l0, l1, lt;
    for (auto &r : dpb) if (r.long_term) lt.push_back(r);
    std::sort(lt.begin(), lt.end(), [](const RefEntry &a, const RefEntry &b) { return a.lt_idx < b.lt_idx; });
    if (ps.type == SLICE_P) {
      for (auto it = dpb.rbegin(); it != dpb.rend(); ++it) if (!it->long_term) l0.push_back(*it);  // decode order reversed = PicNum descending
      l0.insert(l0.end(), lt.begin(), lt.end());
    } else if (ps.type == SLICE_B) {
      std::vector<RefEntry> before, after;
      for (auto &r : dpb) if (!r.long_term) (r.poc < ps.poc ? before : after).push_back(r);
      std::sort(before.begin(), before.end(), [](const RefEntry &a, const RefEntry &b) { return a.poc > b.poc; });
      std::sort(after.begin(), after.end(), [](const RefEntry &a, const RefEntry &b) { return a.poc < b.poc; });
      l0 = before; l0.insert(l0.end(), after.begin(), after.end()); l0.insert(l0.end(), lt.begin(), lt.end());
      l1 = after; l1.insert(l1.end(), before.begin(), before.end()); l1.insert(l1.end(), lt.begin(), lt.end());
      if (l1.size() > 1 && l0.size() == l1.size()) {
        bool same = true;
        for (size_t i = 0; i < l0.size(); ++i) same &= l0[i].slot == l1[i].slot;
        if (same) std::swap(l1[0], l1[1]);
      }
    }
    // pad_refs: what stands in for a list entry without a reference picture (libavcodec: the initial list's first entry)
    const bool pad0 = P.pad_refs && !l0.empty(), pad1 = P.pad_refs && !l1.empty();
    const RefEntry dflt0 = pad0 ? l0[0] : RefEntry(), dflt1 = pad1 ? l1[0] : RefEntry();
    rplm_ops[0].clear(); rplm_ops[1].clear();
    if (ps.rplm) {
      const int active[2] = {std::min<int>((int)l0.size(), P.num_ref), std::min<int>((int)l1.size(), 2)};
      for (int l = 0; l < (ps.type == SLICE_B ? 2 : 1); ++l) {
        std::vector<RefEntry> &lst = l ? l1 : l0;
        if (active[l] < 2) continue;
        lst.resize(active[l]);  // entries beyond num_ref_idx_active are dropped before the modification
        // move one or two entries to the front.  picNumPred starts at CurrPicNum; a short-term target is named by the
        // difference to the prediction (idc 0 subtract / 1 add), a long-term one by its LongTermPicNum (idc 2)
        auto pic_num = [&](const RefEntry &r) { return r.frame_num > ps.frame_num ? r.frame_num - (1 << fn_bits()) : r.frame_num; };
        int pred = ps.frame_num;
        const int nops = 1 + (active[l] > 2 && rng.pct(50));
        for (int k = 0; k < nops; ++k) {
          const int src = k + 1 + rng.below(active[l] - k - 1);  // an entry behind position k
          RefEntry tgt = lst[src];
          if (tgt.long_term) rplm_ops[l].push_back({2, tgt.lt_idx});
          else {
            const int pn = pic_num(tgt);
            if (pn < pred) rplm_ops[l].push_back({0, pred - pn - 1});
            else if (pn > pred) rplm_ops[l].push_back({1, pn - pred - 1});
            else { rplm_ops[l].push_back({0, (1 << fn_bits()) - 1}); }  // same picture again: a full wrap (abs_diff_pic_num = MaxPicNum)
            pred = pn;
          }
          lst.erase(lst.begin() + src);
          lst.insert(lst.begin() + k, tgt);
        }
      }
    }
    if (pad0) while ((int)l0.size() < P.num_ref) l0.push_back(dflt0);
    if (pad1) while ((int)l1.size() < 2) l1.push_back(dflt1);
    int nrefs_total[2] = {(int)l0.size(), (int)l1.size()};

    std::vector<uint8_t> sample;
    const int nal_type = ps.idr ? 5 : 1;
    const int nal_ref_idc = ps.is_ref ? (ps.type == SLICE_I ? 3 : 2) : 0;
    for (int sl = 0; sl < P.slices; ++sl) {
      const int first_mb = (int)((int64_t)sl * nmb / P.slices), end_mb = (int)((int64_t)(sl + 1) * nmb / P.slices);
      SliceDesc &sd = slices[slot * MAXSL + sl];
      memset(&sd, 0, sizeof(sd));
      // mixed_slices: every third slice of an inter picture is an I slice
      const int stype = (P.mixed_slices && P.slices >= 2 && ps.type != SLICE_I && (sl + ps.t) % 3 == 1) ? (int)SLICE_I : (int)ps.type;
      sd.pic = slot; sd.first_mb = first_mb; sd.slice_type = (uint8_t)stype;
      sd.qp = (uint8_t)clip3(P.qp < 10 ? 0 : 10, P.qp > 44 ? 51 : 44, P.qp + (stype == SLICE_B ? 2 : (stype == SLICE_I ? -2 : 0)) + (P.qp_jitter ? rng.range(-1, 1) : 0));
      sd.cabac_init_idc = (uint8_t)(P.cabac_init_idc >= 0 ? P.cabac_init_idc : rng.below(3));
      sd.disable_deblock = (uint8_t)(P.deblock == 1 ? 1 : (P.deblock == 2 ? 2 : 0));
      if (P.deblock == 3) { sd.alpha_off = (int8_t)(rng.range(-3, 3) * 2); sd.beta_off = (int8_t)(rng.range(-3, 3) * 2); if (rng.pct(15)) sd.disable_deblock = (uint8_t)rng.range(1, 2); }
      sd.direct_spatial = (uint8_t)P.direct_spatial;
      sd.num_ref[0] = stype == SLICE_I ? 0 : (uint8_t)std::min<int>((int)l0.size(), P.num_ref);
      sd.num_ref[1] = stype != SLICE_B ? 0 : (uint8_t)std::min<int>((int)l1.size(), 2);
      for (int l = 0; l < 2; ++l) {
        const auto &lst = l ? l1 : l0;
        for (int i = 0; i < sd.num_ref[l]; ++i) {
          sd.ref_frame[l][i] = (int16_t)lst[i].slot; sd.ref_poc[l][i] = lst[i].poc;
          if (lst[i].long_term) sd.ref_long[l] |= 1u << i;
        }
      }
      sd.luma_log2_denom = 5; sd.chroma_log2_denom = 5;
      for (int l = 0; l < 2; ++l) for (int i = 0; i < 32; ++i) { sd.luma_w[l][i] = 32; sd.chroma_w[l][i][0] = sd.chroma_w[l][i][1] = 32; }
      sd.use_weights = 0;
      if (stype == SLICE_P && P.weighted >= 1) {
        sd.use_weights = 1;
        for (int i = 0; i < sd.num_ref[0]; ++i) {
          int diff = content->fade_add(ps.t) - content->fade_add(l0[i].t);
          sd.luma_o[0][i] = (int16_t)clip3(-128, 127, diff);
          if (i == 1) { sd.luma_w[0][i] = 31; sd.chroma_w[0][i][0] = 33; sd.chroma_o[0][i][1] = 1; }
        }
      } else if (stype == SLICE_B && P.weighted == 3) {
        // explicit weights in a B slice (weighted_bipred_idc = 1): per list and reference, luma and chroma, with offsets
        sd.use_weights = 1;
        for (int l = 0; l < 2; ++l)
          for (int i = 0; i < sd.num_ref[l]; ++i) {
            const auto &lst = l ? l1 : l0;
            sd.luma_o[l][i] = (int16_t)clip3(-128, 127, content->fade_add(ps.t) - content->fade_add(lst[i].t));
            sd.luma_w[l][i] = (int16_t)(32 + (l ? -3 : 2) + i); sd.chroma_w[l][i][0] = (int16_t)(31 + i); sd.chroma_w[l][i][1] = (int16_t)(33 - l); sd.chroma_o[l][i][l] = (int16_t)(l ? -2 : 1);
          }
      } else if (stype == SLICE_B && P.weighted >= 2) sd.use_weights = 2;

      SliceEnc e;
      e.cabac = cabac;
      write_slice_header(e.bw, ps, sd, first_mb, nal_type, idr_id, nrefs_total);
      SliceDec &s = e.s;
      s.c = &c; s.pd = &pd; s.sd = &sd; s.slice_num = sl; s.cabac = cabac; s.error = 0;
      s.qp = sd.qp; s.last_dqp = 0;
      s.line = (NbCtx *)(c.ectx + (uint64_t)(slot * MAXSL + sl) * c.ectx_stride);
      s.coef_next = (uint32_t)first_mb * SLOTS_PER_MB;
      init_caches(s); init_lane_tables(s);
      if (cabac) {
        while (!e.bw.aligned()) e.bw.put1(1);
        cabac_init_states(e.ce.st, stype == SLICE_I ? 0 : 1 + sd.cabac_init_idc, sd.qp);
        e.ce.start(&e.bw);
      }
      for (int addr = first_mb; addr < end_mb; ++addr) {
        s.mbaddr = addr; s.mbx = addr % mb_w; s.mby = addr / mb_w;
        s.availA = s.mbx > 0 && addr - 1 >= first_mb;
        s.availB = addr - mb_w >= first_mb;
        s.availC = s.mbx < mb_w - 1 && addr - mb_w + 1 >= first_mb;
        s.availD = s.mbx > 0 && addr - mb_w - 1 >= first_mb;
        fill_caches(s, false);
        MbEnc m;
        encode_one_mb(e, ps, sd, m);
        encode_mb(e, m);
        if (cabac) e.ce.terminate(addr == end_mb - 1);
      }
      if (cabac) e.bw.align_zero();
      else { if (e.run > 0) e.bw.ue((uint32_t)e.run); e.bw.trailing(); }
      append_nal_avcc(sample, nal_ref_idc, nal_type, e.bw.buf);
    }
    // in-loop filter over the finished picture
    for (int mby = 0; mby < mb_h; ++mby) for (int mbx = 0; mbx < mb_w; ++mbx) deblock_mb(c, slot, mbx, mby, &ds);
    return sample;
  }

  // Decide, build the IR and reconstruct one macroblock (caches already filled).
  void encode_one_mb(SliceEnc &e, const PicState &ps, const SliceDesc &sd, MbEnc &m) {
    SliceDec &s = e.s;
    const int slot = ps.slot, mbx = s.mbx, mby = s.mby, addr = s.mbaddr;
    MbInfo &o = s.out;
    memset(&o, 0, sizeof(o));
    o.mbtype = MB_INTER; o.qp = (uint8_t)s.qp; o.slice = (uint16_t)s.slice_num; o.coef_off = s.coef_next;
    for (int i = 0; i < 16; ++i) { o.i4modes[i] = 2; m.i4modes[i] = 2; }
    for (int l = 0; l < 2; ++l) { for (int q = 0; q < 4; ++q) m.ref[l][q] = -1; for (int i = 0; i < 16; ++i) m.mv[l][i][0] = m.mv[l][i][1] = 0; }
    MbInfo *mbs = pic_mbinfo(c, slot);
    const int stype = sd.slice_type;  // the slice's type (a picture may mix I slices in)
    const bool B = stype == SLICE_B;
    int want_qp = s.qp;
    // (the full QP range only for extreme base QPs: older clips keep their streams.  There the walk also crosses the 0 / 51 boundary:
    // QP_Y = (QP_Y,pred + mb_qp_delta + 52) % 52, with the short way round as the coded difference)
    const bool full_range = P.qp < 10 || P.qp > 44;
    if (P.qp_jitter && rng.pct(20)) {
      const int walk = s.qp + rng.range(-P.qp_jitter, P.qp_jitter);
      want_qp = full_range ? (walk + 52) % 52 : clip3(8, 46, walk);
    }
    auto qp_delta = [&](int want, int pred) { int d = want - pred; if (d > 25) d -= 52; if (d < -26) d += 52; return d; };
    bool intra = stype == SLICE_I || rng.pct(P.intra_in_p_pct);
    Levels L;
    memset(&L, 0, sizeof(L));
    int ry[256], ru[64], rv[64];
    ReconScratch saved = rs;

    if (intra && P.ipcm_per_100k > 0 && rng.below(100000) < P.ipcm_per_100k) {
      // ---------------- I_PCM
      m.imbt = 25; m.mbt = stype == SLICE_I ? 25 : (stype == SLICE_P ? 30 : 48);
      o.mbtype = MB_IPCM; o.qp = 0; o.cbp = 0x2F; o.nzmask = 0xFFF;
      uint8_t *dst = (uint8_t *)(pic_coefs(c, slot) + (uint64_t)s.coef_next * 16);
      for (int i = 0; i < 256; ++i) m.pcm[i] = srcY[(size_t)(mby * 16 + (i >> 4)) * wc + mbx * 16 + (i & 15)];
      for (int i = 0; i < 64; ++i) { m.pcm[256 + i] = srcU[(size_t)(mby * 8 + (i >> 3)) * (wc / 2) + mbx * 8 + (i & 7)]; m.pcm[320 + i] = srcV[(size_t)(mby * 8 + (i >> 3)) * (wc / 2) + mbx * 8 + (i & 7)]; }
      memcpy(dst, m.pcm, 384);
      s.coef_next += 12;
      mbs[addr] = o;
      recon_mb(c, slot, mbx, mby, &rs);
      return;
    }
    if (intra) {
      // ---------------- intra
      IntraAvail av = intra_avail(s, slot);
      // neighbours' MbInfo.slice must be valid for recon's availability test: they are (same picture)
      static const int nxn_pct = getenv("HWGEN_NXN_PCT") ? atoi(getenv("HWGEN_NXN_PCT")) : 45;
      bool nxn = rng.pct(nxn_pct);
      bool t8 = nxn && high && rng.pct(45);
      m.cmode = pick_chroma_mode(av);
      o.cmode = (uint8_t)m.cmode;
      const int qpc0 = chroma_qp(want_qp, pd_of(slot).chroma_qp_offset[0]), qpc1 = chroma_qp(want_qp, pd_of(slot).chroma_qp_offset[1]);
      if (nxn) {
        m.imbt = 0; m.t8 = t8;
        o.mbtype = t8 ? MB_I8x8 : MB_I4x4;
        if (t8) o.flags |= MBF_T8x8;
        decide_inxn(s, slot, mbx, mby, t8, want_qp, av, m, L);
        for (int i = 0; i < 16; ++i) o.i4modes[i] = (uint8_t)m.i4modes[i];
      } else {
        o.mbtype = MB_I16x16;
        o.imode = (uint8_t)pick_i16_mode(slot, mbx, mby, av);
      }
      // pass 1: prediction only (chroma, and luma for I16x16)
      o.qp = (uint8_t)want_qp;
      o.nzmask = 0; o.cbp = 0;
      MbInfo pass1 = o;
      if (nxn) {
        // keep NxN luma levels in pass 1 so that the luma is final; chroma residual is added in pass 2
        Levels Lc = L; Lc.has_cdc[0] = Lc.has_cdc[1] = false; memset(Lc.cnz, 0, sizeof(Lc.cnz));
        uint32_t keep = s.coef_next;
        emit_levels(s, pass1, false, t8, Lc);
        s.coef_next = keep;
      }
      mbs[addr] = pass1;
      recon_mb(c, slot, mbx, mby, &rs);
      mb_residual(slot, mbx, mby, ry, ru, rv);
      if (!nxn) quant_luma(ry, true, false, true, want_qp, L);
      quant_chroma(ru, 0, true, qpc0, L); quant_chroma(rv, 1, true, qpc1, L);
      emit_levels(s, o, !nxn, t8, L);
      if (!nxn) {
        int cbp = o.cbp;
        m.imbt = 1 + o.imode + 4 * (cbp >> 4) + ((cbp & 15) ? 12 : 0);
      }
      m.mbt = m.imbt + (stype == SLICE_I ? 0 : (stype == SLICE_P ? 5 : 23));
      m.cbp = o.cbp;
      if (o.cbp || !nxn) { m.dqp = qp_delta(want_qp, s.qp); o.qp = (uint8_t)want_qp; }
      else { m.dqp = 0; o.qp = (uint8_t)s.qp; }
      mbs[addr] = o;
      rs = saved;
      recon_mb(c, slot, mbx, mby, &rs);
      return;
    }

    // ---------------- inter
    const int nl = B ? 2 : 1;
    int8_t dref[2][4];
    int16_t dmv[2][16][2];
    int pf_q[4] = {1, 1, 1, 1};  // prediction flags per quadrant
    bool direct_all = false;
    auto true_mv = [&](int l, int ri, int px, int py, int &mx, int &my) {
      int tref = l ? -1 : -1; (void)tref;
      int d = ps.t - ref_t(sd, l, ri);
      content->motion(ps.t, std::min(px, W - 1), std::min(py, H - 1), d, mx, my);
      if (rng.pct(25)) { mx += rng.range(-2, 2); my += rng.range(-2, 2); }
      // keep the reference block within ~24 samples of the picture
      int x0 = mbx * 16 * 4, y0 = mby * 16 * 4;
      mx = clip3(-(x0 + 24 * 4), (wc - mbx * 16 + 8) * 4, mx);
      my = clip3(-(y0 + 24 * 4), (hc - mby * 16 + 8) * 4, my);
    };
    auto pick_ref = [&](int l) { int n = sd.num_ref[l]; return (n > 1 && rng.pct(22)) ? rng.below(n) : 0; };
    auto fill_part = [&](int l, int bx, int by, int w, int h, int ri) {
      int mx, my;
      true_mv(l, ri, mbx * 16 + bx * 4 + w * 2, mby * 16 + by * 4 + h * 2, mx, my);
      for (int y = by; y < by + h; ++y) for (int x = bx; x < bx + w; ++x) { m.mv[l][y * 4 + x][0] = mx; m.mv[l][y * 4 + x][1] = my; }
      for (int y = by; y < by + h; ++y) for (int x = bx; x < bx + w; ++x) m.ref[l][(y >> 1) * 2 + (x >> 1)] = ri;
    };
    int roll = rng.below(100);
    bool try_skip = false;
    if (!B) {
      if (roll < 60) {
        m.mbt = 0;
        int ri = pick_ref(0);
        fill_part(0, 0, 0, 4, 4, ri);
        // snap to the predictor / skip vector when close: mvd = 0 and P_Skip become common, as in real streams
        int px, py;
        pred_mv(s, 0, 0, 0, 4, ri, 0, px, py);
        int sx = 0, sy = 0;
        MvRef A = mv_at(s, 0, -1, 0), Bn = mv_at(s, 0, 0, -1);
        if (!(A.ref == REF_UNAVAIL || Bn.ref == REF_UNAVAIL || (A.ref == 0 && A.mx == 0 && A.my == 0) || (Bn.ref == 0 && Bn.mx == 0 && Bn.my == 0)))
          pred_mv(s, 0, 0, 0, 4, 0, 0, sx, sy);
        int mx = m.mv[0][0][0], my = m.mv[0][0][1];
        if (ri == 0 && abs(mx - sx) <= 1 && abs(my - sy) <= 1) { mx = sx; my = sy; try_skip = true; }
        else if (abs(mx - px) <= 1 && abs(my - py) <= 1) { mx = px; my = py; }
        for (int i = 0; i < 16; ++i) { m.mv[0][i][0] = mx; m.mv[0][i][1] = my; }
      } else if (roll < 70) { m.mbt = 1; fill_part(0, 0, 0, 4, 2, pick_ref(0)); fill_part(0, 0, 2, 4, 2, pick_ref(0)); }
      else if (roll < 80) { m.mbt = 2; fill_part(0, 0, 0, 2, 4, pick_ref(0)); fill_part(0, 2, 0, 2, 4, pick_ref(0)); }
      else {
        m.mbt = (!cabac && sd.num_ref[0] > 1 && rng.pct(20)) ? 4 : 3;
        for (int q = 0; q < 4; ++q) {
          int t = rng.pct(50) ? 0 : rng.range(1, 3);
          m.sub[q] = t;
          int bx = (q & 1) * 2, by = (q >> 1) * 2, ri = m.mbt == 4 ? 0 : pick_ref(0);
          if (t == 0) fill_part(0, bx, by, 2, 2, ri);
          else if (t == 1) { fill_part(0, bx, by, 2, 1, ri); fill_part(0, bx, by + 1, 2, 1, ri); }
          else if (t == 2) { fill_part(0, bx, by, 1, 2, ri); fill_part(0, bx + 1, by, 1, 2, ri); }
          else for (int k = 0; k < 4; ++k) fill_part(0, bx + (k & 1), by + (k >> 1), 1, 1, ri);
        }
      }
    } else {
      if (roll < 35) {
        m.mbt = 0; direct_all = true; try_skip = true;
        direct_predict(s, 15, dref, dmv);
        for (int l = 0; l < 2; ++l) { for (int q = 0; q < 4; ++q) m.ref[l][q] = dref[l][q]; for (int i = 0; i < 16; ++i) { m.mv[l][i][0] = dmv[l][i][0]; m.mv[l][i][1] = dmv[l][i][1]; } }
      } else if (roll < 75) {
        int pf = roll < 50 ? 1 : (roll < 60 ? 2 : 3);
        m.mbt = pf;
        for (int l = 0; l < 2; ++l) if (pf & (1 << l)) fill_part(l, 0, 0, 4, 4, pick_ref(l));
      } else if (roll < 88) {
        m.mbt = rng.range(4, 21);
        int shape = (m.mbt & 1) ? 2 : 1, k = (m.mbt - 4) >> 1;
        for (int p = 0; p < 2; ++p) {
          int pf = b_part_pred[k * 2 + p];
          int bx = (shape == 2 && p) ? 2 : 0, by = (shape == 1 && p) ? 2 : 0, w = shape == 2 ? 2 : 4, h = shape == 1 ? 2 : 4;
          for (int l = 0; l < 2; ++l) if (pf & (1 << l)) fill_part(l, bx, by, w, h, pick_ref(l));
        }
      } else {
        m.mbt = 22;
        int dq = 0;
        for (int q = 0; q < 4; ++q) { m.sub[q] = rng.pct(35) ? 0 : rng.range(1, 12); if (m.sub[q] == 0) dq |= 1 << q; }
        if (dq) direct_predict(s, dq, dref, dmv);
        for (int q = 0; q < 4; ++q) {
          int t = m.sub[q], bx = (q & 1) * 2, by = (q >> 1) * 2;
          if (t == 0) {
            for (int l = 0; l < 2; ++l) { m.ref[l][q] = dref[l][q]; for (int k = 0; k < 4; ++k) { int i = (by + (k >> 1)) * 4 + bx + (k & 1); m.mv[l][i][0] = dmv[l][i][0]; m.mv[l][i][1] = dmv[l][i][1]; } }
            continue;
          }
          int shape = t <= 3 ? 0 : (t >= 10 ? 3 : ((t & 1) ? 2 : 1)), pf = t <= 3 ? t : (t >= 10 ? t - 9 : ((t - 4) >> 1) + 1);
          for (int l = 0; l < 2; ++l) {
            if (!(pf & (1 << l))) continue;
            int ri = pick_ref(l);
            if (shape == 0) fill_part(l, bx, by, 2, 2, ri);
            else if (shape == 1) { fill_part(l, bx, by, 2, 1, ri); fill_part(l, bx, by + 1, 2, 1, ri); }
            else if (shape == 2) { fill_part(l, bx, by, 1, 2, ri); fill_part(l, bx + 1, by, 1, 2, ri); }
            else for (int k = 0; k < 4; ++k) fill_part(l, bx + (k & 1), by + (k >> 1), 1, 1, ri);
          }
        }
      }
    }
    (void)pf_q; (void)direct_all;
    // IR motion
    for (int l = 0; l < nl; ++l) {
      int16_t *mvo = pic_mv(c, slot, l) + (uint64_t)addr * 32;
      int8_t *ro = pic_refidx(c, slot, l) + (uint64_t)addr * 4;
      int16_t *po = pic_refpic(c, slot, l) + (uint64_t)addr * 4;
      for (int i = 0; i < 16; ++i) { mvo[2 * i] = (int16_t)m.mv[l][i][0]; mvo[2 * i + 1] = (int16_t)m.mv[l][i][1]; }
      for (int q = 0; q < 4; ++q) { ro[q] = (int8_t)m.ref[l][q]; po[q] = m.ref[l][q] >= 0 ? sd.ref_frame[l][m.ref[l][q]] : (int16_t)-1; }
    }
    mbs[addr] = o;
    recon_mb(c, slot, mbx, mby, &rs);
    mb_residual(slot, mbx, mby, ry, ru, rv);
    bool t8 = false;
    if (high) {
      bool allowed = true;
      if ((!B && m.mbt >= 3) || (B && m.mbt == 22)) for (int q = 0; q < 4; ++q) { int t = m.sub[q]; if (B ? (t >= 4) : (t != 0)) allowed = false; }
      // without direct_8x8_inference a macroblock with direct parts (whose motion may change every 4x4 block) has no 8x8 transform
      if (B && !pd_of(slot).direct_8x8_inference) { if (m.mbt == 0) allowed = false; if (m.mbt == 22) for (int q = 0; q < 4; ++q) if (m.sub[q] == 0) allowed = false; }
      t8 = allowed && rng.pct(50);
    }
    const int qpc0 = chroma_qp(want_qp, pd_of(slot).chroma_qp_offset[0]), qpc1 = chroma_qp(want_qp, pd_of(slot).chroma_qp_offset[1]);
    quant_luma(ry, false, t8, false, want_qp, L);
    quant_chroma(ru, 0, false, qpc0, L); quant_chroma(rv, 1, false, qpc1, L);
    emit_levels(s, o, false, t8, L);
    if (t8 && (o.cbp & 15)) { o.flags |= MBF_T8x8; m.t8 = true; }
    m.cbp = o.cbp;
    if (o.cbp) { m.dqp = qp_delta(want_qp, s.qp); o.qp = (uint8_t)want_qp; }
    else { m.dqp = 0; o.qp = (uint8_t)s.qp; }
    if (!o.cbp && try_skip) { m.skipped = true; o.flags |= MBF_SKIP; }
    mbs[addr] = o;
    if (o.cbp) { rs = saved; recon_mb(c, slot, mbx, mby, &rs); }
  }
  const PicDesc &pd_of(int slot) const { return pics[slot]; }
  int ref_t(const SliceDesc &sd, int l, int ri) const {
    int slot = sd.ref_frame[l][ri];
    for (auto &r : dpb) if (r.slot == slot) return r.t;
    return 0;
  }
};

// ------------------------------------------------------------------------------------------ MP4
static void be32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
static void be16(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 8); v.push_back(x); }
static std::vector<uint8_t> box(const char *type, const std::vector<uint8_t> &payload) {
  std::vector<uint8_t> b;
  be32(b, (uint32_t)payload.size() + 8);
  b.insert(b.end(), type, type + 4);
  b.insert(b.end(), payload.begin(), payload.end());
  return b;
}
static std::vector<uint8_t> fullbox(const char *type, uint32_t vf, const std::vector<uint8_t> &payload) {
  std::vector<uint8_t> p; be32(p, vf); p.insert(p.end(), payload.begin(), payload.end());
  return box(type, p);
}
static void cat(std::vector<uint8_t> &a, const std::vector<uint8_t> &b) { a.insert(a.end(), b.begin(), b.end()); }

static std::vector<uint8_t> mux_mp4(int W, int H, const std::vector<uint8_t> &avcc, const std::vector<std::vector<uint8_t>> &samples,
                                    const std::vector<int> &keyframes, int fragmented, int gop) {
  const uint32_t N = (uint32_t)samples.size(), timescale = 30;
  std::vector<uint8_t> ftyp;
  { std::vector<uint8_t> p; p.insert(p.end(), {'i', 's', 'o', 'm'}); be32(p, 0x200); for (const char *b : {"isom", "iso2", "avc1", "mp41"}) p.insert(p.end(), b, b + 4); ftyp = box("ftyp", p); }
  auto stsd = [&]() {
    std::vector<uint8_t> e;
    for (int i = 0; i < 6; ++i) e.push_back(0);
    be16(e, 1);
    for (int i = 0; i < 16; ++i) e.push_back(0);
    be16(e, W); be16(e, H); be32(e, 0x00480000); be32(e, 0x00480000); be32(e, 0); be16(e, 1);
    for (int i = 0; i < 32; ++i) e.push_back(0);
    be16(e, 0x18); be16(e, 0xFFFF);
    cat(e, box("avcC", avcc));
    std::vector<uint8_t> p; be32(p, 1); cat(p, box("avc1", e));
    return fullbox("stsd", 0, p);
  };
  auto mvhd = [&](uint32_t dur) { std::vector<uint8_t> p; be32(p, 0); be32(p, 0); be32(p, timescale); be32(p, dur); be32(p, 0x00010000); be16(p, 0x0100); be16(p, 0); be32(p, 0); be32(p, 0);
    uint32_t mat[9] = {0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000}; for (auto m : mat) be32(p, m); for (int i = 0; i < 6; ++i) be32(p, 0); be32(p, 2); return fullbox("mvhd", 0, p); };
  auto tkhd = [&](uint32_t dur) { std::vector<uint8_t> p; be32(p, 0); be32(p, 0); be32(p, 1); be32(p, 0); be32(p, dur); be32(p, 0); be32(p, 0); be16(p, 0); be16(p, 0); be16(p, 0); be16(p, 0);
    uint32_t mat[9] = {0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000}; for (auto m : mat) be32(p, m); be32(p, (uint32_t)W << 16); be32(p, (uint32_t)H << 16); return fullbox("tkhd", 3, p); };
  auto mdhd = [&](uint32_t dur) { std::vector<uint8_t> p; be32(p, 0); be32(p, 0); be32(p, timescale); be32(p, dur); be16(p, 0x55C4); be16(p, 0); return fullbox("mdhd", 0, p); };
  auto hdlr = [&]() { std::vector<uint8_t> p; be32(p, 0); p.insert(p.end(), {'v', 'i', 'd', 'e'}); be32(p, 0); be32(p, 0); be32(p, 0); const char *n = "VideoHandler"; p.insert(p.end(), n, n + 13); return fullbox("hdlr", 0, p); };
  auto vmhd = [&]() { std::vector<uint8_t> p; be16(p, 0); be16(p, 0); be16(p, 0); be16(p, 0); return fullbox("vmhd", 1, p); };
  auto dinf = [&]() { std::vector<uint8_t> p; be32(p, 1); cat(p, fullbox("url ", 1, {})); return box("dinf", fullbox("dref", 0, p)); };

  std::vector<uint8_t> out = ftyp;
  if (!fragmented) {
    std::vector<uint8_t> stts; { std::vector<uint8_t> p; be32(p, 1); be32(p, N); be32(p, 1); stts = fullbox("stts", 0, p); }
    std::vector<uint8_t> stsc; { std::vector<uint8_t> p; be32(p, 1); be32(p, 1); be32(p, N); be32(p, 1); stsc = fullbox("stsc", 0, p); }
    std::vector<uint8_t> stsz; { std::vector<uint8_t> p; be32(p, 0); be32(p, N); for (auto &s : samples) be32(p, (uint32_t)s.size()); stsz = fullbox("stsz", 0, p); }
    std::vector<uint8_t> stss; { std::vector<uint8_t> p; be32(p, (uint32_t)keyframes.size()); for (int k : keyframes) be32(p, (uint32_t)k + 1); stss = fullbox("stss", 0, p); }
    // two-pass: the chunk offset depends on the moov size
    uint64_t total = 0; for (auto &s : samples) total += s.size();
    bool co64 = total + (1u << 20) > 0xFFFFFFFFull;
    auto build_moov = [&](uint64_t off) {
      std::vector<uint8_t> stco;
      if (co64) { std::vector<uint8_t> p; be32(p, 1); be32(p, (uint32_t)(off >> 32)); be32(p, (uint32_t)off); stco = fullbox("co64", 0, p); }
      else { std::vector<uint8_t> p; be32(p, 1); be32(p, (uint32_t)off); stco = fullbox("stco", 0, p); }
      std::vector<uint8_t> stbl; cat(stbl, stsd()); cat(stbl, stts); cat(stbl, stsc); cat(stbl, stsz); cat(stbl, stco); cat(stbl, stss);
      std::vector<uint8_t> minf; cat(minf, vmhd()); cat(minf, dinf()); cat(minf, box("stbl", stbl));
      std::vector<uint8_t> mdia; cat(mdia, mdhd(N)); cat(mdia, hdlr()); cat(mdia, box("minf", minf));
      std::vector<uint8_t> trak; cat(trak, tkhd(N)); cat(trak, box("mdia", mdia));
      std::vector<uint8_t> moov; cat(moov, mvhd(N)); cat(moov, box("trak", trak));
      return box("moov", moov);
    };
    std::vector<uint8_t> moov = build_moov(0);
    bool big = total + 8 > 0xFFFFFFFFull;
    uint64_t off = ftyp.size() + moov.size() + (big ? 16 : 8);
    moov = build_moov(off);
    cat(out, moov);
    if (big) { be32(out, 1); out.insert(out.end(), {'m', 'd', 'a', 't'}); be32(out, (uint32_t)((total + 16) >> 32)); be32(out, (uint32_t)(total + 16)); }
    else { be32(out, (uint32_t)(total + 8)); out.insert(out.end(), {'m', 'd', 'a', 't'}); }
    for (auto &s : samples) cat(out, s);
    return out;
  }
  // fragmented: moov{mvhd, trak(empty tables), mvex{trex}} + (moof{mfhd, traf{tfhd, trun}} + mdat) per GOP
  {
    auto empty = [&](const char *t) { std::vector<uint8_t> p; be32(p, 0); return fullbox(t, 0, p); };
    std::vector<uint8_t> stsz; { std::vector<uint8_t> p; be32(p, 0); be32(p, 0); stsz = fullbox("stsz", 0, p); }
    std::vector<uint8_t> stbl; cat(stbl, stsd()); cat(stbl, empty("stts")); cat(stbl, empty("stsc")); cat(stbl, stsz); cat(stbl, empty("stco"));
    std::vector<uint8_t> minf; cat(minf, vmhd()); cat(minf, dinf()); cat(minf, box("stbl", stbl));
    std::vector<uint8_t> mdia; cat(mdia, mdhd(0)); cat(mdia, hdlr()); cat(mdia, box("minf", minf));
    std::vector<uint8_t> trak; cat(trak, tkhd(0)); cat(trak, box("mdia", mdia));
    std::vector<uint8_t> trex; { std::vector<uint8_t> p; be32(p, 1); be32(p, 1); be32(p, 1); be32(p, 0); be32(p, 0x01010000); trex = fullbox("trex", 0, p); }
    std::vector<uint8_t> moov; cat(moov, mvhd(0)); cat(moov, box("trak", trak)); cat(moov, box("mvex", trex));
    cat(out, box("moov", moov));
    uint32_t seq = 1;
    for (uint32_t g0 = 0; g0 < N; g0 += (uint32_t)gop) {
      uint32_t g1 = std::min<uint32_t>(N, g0 + (uint32_t)gop), n = g1 - g0;
      // tfhd: default-base-is-moof (0x020000); trun: data-offset (0x1) + first-sample-flags (0x4) + sample-size (0x200)
      std::vector<uint8_t> tfhd; { std::vector<uint8_t> p; be32(p, 1); tfhd = fullbox("tfhd", 0x020000, p); }
      auto build_moof = [&](uint32_t data_off) {
        std::vector<uint8_t> p; be32(p, n); be32(p, data_off); be32(p, 0x02000000);
        for (uint32_t i = g0; i < g1; ++i) be32(p, (uint32_t)samples[i].size());
        std::vector<uint8_t> trun = fullbox("trun", 0x000205, p);
        std::vector<uint8_t> traf; cat(traf, tfhd); cat(traf, trun);
        std::vector<uint8_t> mfhd; { std::vector<uint8_t> q; be32(q, seq); mfhd = fullbox("mfhd", 0, q); }
        std::vector<uint8_t> moof; cat(moof, mfhd); cat(moof, box("traf", traf));
        return box("moof", moof);
      };
      std::vector<uint8_t> moof = build_moof(0);
      moof = build_moof((uint32_t)moof.size() + 8);
      cat(out, moof);
      uint64_t tot = 8; for (uint32_t i = g0; i < g1; ++i) tot += samples[i].size();
      be32(out, (uint32_t)tot); out.insert(out.end(), {'m', 'd', 'a', 't'});
      for (uint32_t i = g0; i < g1; ++i) cat(out, samples[i]);
      seq++;
    }
    return out;
  }
}

// ------------------------------------------------------------------------------------------ driver
static int encode_clip(const hwgen_params &P, std::vector<uint8_t> &mp4, uint8_t *recon_yuv) {
  if (P.width % 8 || P.height % 2 || P.width < 16 || P.height < 16) { g_err = "width must be a multiple of 8 and height even"; return -1; }
  if (P.slices < 1 || P.slices > MAXSL) { g_err = "slices out of range"; return -1; }
  if (P.profile == 0 && (P.bframes || (P.cabac == 1))) { g_err = "baseline: no B pictures / CABAC"; return -1; }
  if (P.poc_type == 2 && P.bframes) { g_err = "poc_type 2 requires bframes == 0"; return -1; }
  if (P.poc_type < 0 || P.poc_type > 2) { g_err = "poc_type must be 0, 1 or 2"; return -1; }
  if (P.b_pyramid && P.num_ref < 3) { g_err = "b_pyramid needs num_ref >= 3 (two anchors and the reference B picture)"; return -1; }
  Content content;
  content.init(P.width, P.height, P.seed, P.weighted >= 1);
  const int ngop = (P.frames + P.gop - 1) / P.gop;
  std::vector<std::vector<uint8_t>> samples(P.frames);
  std::vector<uint8_t> avcc;
  {
    Encoder e0(P, &content);
    std::vector<uint8_t> sps = make_nal_raw(3, 7, e0.sps_rbsp()), pps = make_nal_raw(3, 8, e0.pps_rbsp());
    avcc = {1, sps[1], sps[2], sps[3], 0xFF, 0xE1};
    be16(avcc, (uint32_t)sps.size()); cat(avcc, sps);
    avcc.push_back(1); be16(avcc, (uint32_t)pps.size()); cat(avcc, pps);
  }
  int nthreads = P.threads > 0 ? P.threads : (int)std::thread::hardware_concurrency();
  nthreads = std::max(1, std::min(nthreads, ngop));
  std::vector<std::string> errs(nthreads);
  auto worker = [&](int tid) {
    Encoder enc(P, &content);
    const size_t ysz = (size_t)P.width * P.height, fsz = ysz * 3 / 2;
    for (int g = tid; g < ngop; g += nthreads) {
      const int t0 = g * P.gop, n = std::min(P.gop, P.frames - t0);
      enc.rng = Rng(P.seed * 1000003ull + g);
      enc.dpb.clear();
      // decode order within the GOP: anchors (I, then P every bframes+1 pictures), each followed by the B pictures
      // it closes; with b_pyramid the middle B picture of a run is coded first and kept as a reference
      struct Plan { int t, type; bool is_ref; };
      std::vector<Plan> plan;
      plan.push_back({0, SLICE_I, true});
      int step = P.bframes + 1, i = 0;
      while (i + step < n) {
        plan.push_back({i + step, SLICE_P, true});
        const int mid = (P.b_pyramid && P.bframes >= 2) ? (P.bframes + 1) / 2 : 0;
        if (mid) plan.push_back({i + mid, SLICE_B, true});
        for (int b = 1; b <= P.bframes; ++b) if (b != mid) plan.push_back({i + b, SLICE_B, false});
        i += step;
      }
      for (int k = i + 1; k < n; ++k) plan.push_back({k, SLICE_P, true});
      int frame_num = 0;
      // POC bookkeeping of the generator itself: `poc_base` moves when MMCO 5 restarts the numbering
      int poc_base = 0, frame_num_offset = 0, prev_frame_num = 0;
      bool prev_mmco5 = false;
      int max_lt_idx = -1;  // MaxLongTermFrameIdx ("no long-term frame indices")
      for (size_t k = 0; k < plan.size(); ++k) {
        Encoder::PicState ps;
        ps.t = t0 + plan[k].t; ps.type = plan[k].type; ps.idr = k == 0; ps.is_ref = plan[k].is_ref;
        ps.poc = 2 * plan[k].t - poc_base; ps.frame_num = frame_num; ps.gop_t0 = t0;
        ps.slot = enc.free_slot();
        if (ps.slot < 0) { errs[tid] = "generator: out of frame slots"; return; }
        ps.rplm = ps.type != SLICE_I && P.rplm_pct > 0 && enc.rng.pct(P.rplm_pct);
        if (P.poc_type == 1) {
          // 8.2.1.2 with a one-entry cycle (offset_for_ref_frame[0] = 2): expected POC from frame_num alone
          if (ps.idr) frame_num_offset = 0;
          else if (prev_mmco5) frame_num_offset = 0;
          else if (prev_frame_num > frame_num) frame_num_offset += 1 << enc.fn_bits();
          int abs_fn = frame_num_offset + frame_num;
          if (!ps.is_ref && abs_fn > 0) abs_fn--;
          int expected = abs_fn > 0 ? 2 * abs_fn : 0;
          if (!ps.is_ref) expected += kPoc1NonRefOffset;
          ps.delta_poc = ps.poc - expected;
          if (P.bframes == 0 && ps.delta_poc != 0) { errs[tid] = "generator: poc_type 1 without B pictures must follow the expected POC"; return; }
        }
        // ---- adaptive reference marking: decided before the picture is coded (it is slice-header syntax), applied after
        std::vector<RefEntry> &dpb = enc.dpb;
        auto pic_num = [&](const RefEntry &r) { return r.frame_num > frame_num ? r.frame_num - (1 << enc.fn_bits()) : r.frame_num; };
        int n_short = 0, n_long = 0;
        for (auto &r : dpb) (r.long_term ? n_long : n_short)++;
        bool cur_long = false; int cur_lt_idx = 0; bool mmco5 = false;
        if (P.mmco && ps.idr && enc.rng.pct(50) && P.num_ref >= 2) { ps.long_term_idr = true; cur_long = true; cur_lt_idx = 0; }
        if (P.mmco && ps.is_ref && !ps.idr && P.num_ref >= 2 && enc.rng.pct(40)) {
          const int choice = enc.rng.below(P.bframes == 0 ? 5 : 4);
          if (choice == 4 && ps.type == SLICE_P && n_short + n_long > 0) {
            ps.mmco.push_back({5, 0}); mmco5 = true;            // everything before this picture is forgotten
          } else if (choice == 0 && n_short >= 2) {
            const RefEntry *victim = nullptr;                     // MMCO 1: drop the OLDEST short-term reference by name
            for (auto &r : dpb) if (!r.long_term) { victim = &r; break; }
            ps.mmco.push_back({1, frame_num - pic_num(*victim) - 1});
          } else if (choice == 1 && n_long == 0 && n_short >= 2) {
            if (max_lt_idx < 0) ps.mmco.push_back({4, 1});      // MMCO 4: one long-term index; MMCO 3: oldest short-term -> long-term 0
            const RefEntry *victim = nullptr;
            for (auto &r : dpb) if (!r.long_term) { victim = &r; break; }
            ps.mmco.push_back({3, frame_num - pic_num(*victim) - 1});
            ps.mmco.push_back({-3, 0});                          // (argument 2 of MMCO 3: long_term_frame_idx, written below)
          } else if (choice == 2 && n_long > 0) {
            for (auto &r : dpb) if (r.long_term) { ps.mmco.push_back({2, r.lt_idx}); break; }  // MMCO 2: release the long-term picture
          } else if (choice == 3 && n_long == 0 && n_short >= 1) {
            if (max_lt_idx < 0) ps.mmco.push_back({4, 1});
            ps.mmco.push_back({6, 0});                            // MMCO 6: the current picture becomes long-term 0
            cur_long = true; cur_lt_idx = 0;
          }
          // adaptive marking does no sliding window: when the DPB is full, name one more short-term picture to drop
          if (!ps.mmco.empty() && !mmco5) {
            int after_short = n_short, after_long = n_long;
            for (auto &m : ps.mmco) { if (m.op == 1) after_short--; if (m.op == 3) { after_short--; after_long++; } if (m.op == 2) after_long--; }
            if (after_short + after_long + 1 > P.num_ref) {
              // the oldest short-term picture not already named
              int skip = 0;
              for (auto &m : ps.mmco) if (m.op == 1 || m.op == 3) skip++;
              const RefEntry *victim = nullptr;
              for (auto &r : dpb) if (!r.long_term) { if (skip-- == 0) { victim = &r; break; } }
              if (victim) ps.mmco.insert(ps.mmco.begin(), {1, frame_num - pic_num(*victim) - 1});
              else ps.mmco.clear();
            }
          }
        }
        // flatten MMCO 3's second argument into the op stream the slice header writer emits (op, a) pairs
        {
          std::vector<MmcoOp> flat;
          for (size_t q = 0; q < ps.mmco.size(); ++q) {
            if (ps.mmco[q].op == -3) continue;
            flat.push_back(ps.mmco[q]);
          }
          ps.mmco3_lt_idx = 0;
          ps.mmco = flat;
        }
        samples[t0 + (int)k] = enc.encode_picture(ps, g & 0xFFFF);
        if (recon_yuv) {
          uint8_t *dst = recon_yuv + (size_t)ps.t * fsz;
          const uint8_t *Y = frame_y(enc.c, ps.slot), *U = frame_cb(enc.c, ps.slot), *V = frame_cr(enc.c, ps.slot);
          for (int y = 0; y < P.height; ++y) memcpy(dst + (size_t)y * P.width, Y + (size_t)y * enc.wc, P.width);
          for (int y = 0; y < P.height / 2; ++y) { memcpy(dst + ysz + (size_t)y * P.width / 2, U + (size_t)y * enc.wc / 2, P.width / 2); memcpy(dst + ysz + ysz / 4 + (size_t)y * P.width / 2, V + (size_t)y * enc.wc / 2, P.width / 2); }
        }
        prev_frame_num = frame_num;
        prev_mmco5 = false;
        if (ps.is_ref) {
          // ---- decoded reference picture marking (8.2.5) mirrored on the generator's DPB
          if (ps.idr) {
            dpb.clear();
            max_lt_idx = ps.long_term_idr ? 0 : -1;
          } else if (ps.mmco.empty()) {
            if (n_short + n_long >= P.num_ref)  // sliding window: the oldest short-term picture goes
              for (size_t q = 0; q < dpb.size(); ++q) if (!dpb[q].long_term) { dpb.erase(dpb.begin() + q); break; }
          } else {
            for (auto &m : ps.mmco) {
              if (m.op == 1 || m.op == 3) {
                const int target = frame_num - (m.a + 1);
                for (size_t q = 0; q < dpb.size(); ++q)
                  if (!dpb[q].long_term && pic_num(dpb[q]) == target) {
                    if (m.op == 1) dpb.erase(dpb.begin() + q);
                    else { dpb[q].long_term = true; dpb[q].lt_idx = 0; }
                    break;
                  }
              } else if (m.op == 2) {
                for (size_t q = 0; q < dpb.size(); ++q) if (dpb[q].long_term && dpb[q].lt_idx == m.a) { dpb.erase(dpb.begin() + q); break; }
              } else if (m.op == 4) {
                max_lt_idx = m.a - 1;
              } else if (m.op == 5) {
                dpb.clear(); max_lt_idx = -1;
              }
            }
          }
          RefEntry cur{ps.slot, ps.frame_num, ps.poc, ps.t};
          cur.long_term = cur_long; cur.lt_idx = cur_lt_idx;
          if (mmco5) {
            // 8.2.1: after MMCO 5 the picture counts as frame_num 0 / POC 0; later pictures continue from there
            cur.frame_num = 0; cur.poc = 0;
            poc_base += ps.poc; frame_num = 0; prev_frame_num = 0; prev_mmco5 = true; frame_num_offset = 0;
          }
          dpb.push_back(cur);
          frame_num = (frame_num + 1) & ((1 << enc.fn_bits()) - 1);
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
  for (auto &t : th) t.join();
  for (auto &e : errs) if (!e.empty()) { g_err = e; return -1; }
  std::vector<int> keyframes;
  for (int g = 0; g < ngop; ++g) keyframes.push_back(g * P.gop);
  mp4 = mux_mp4(P.width, P.height, avcc, samples, keyframes, P.fragmented, P.gop);
  return 0;
}

}  // namespace gen

extern "C" {
void hwgen_default_params(hwgen_params *p) {
  memset(p, 0, sizeof(*p));
  p->width = 640; p->height = 480; p->frames = 30; p->gop = 30; p->profile = 0; p->cabac = -1; p->bframes = 0; p->num_ref = 1;
  p->qp = 26; p->slices = 1; p->seed = 1; p->direct_spatial = 1; p->cabac_init_idc = 0; p->intra_in_p_pct = 2;
}
int hwgen_encode(const hwgen_params *p, uint8_t **out_mp4, size_t *out_size, uint8_t *recon_yuv) {
  std::vector<uint8_t> mp4;
  int rc = gen::encode_clip(*p, mp4, recon_yuv);
  if (rc) return rc;
  *out_mp4 = (uint8_t *)malloc(mp4.size());
  memcpy(*out_mp4, mp4.data(), mp4.size());
  *out_size = mp4.size();
  return 0;
}
void hwgen_free(void *p) { free(p); }
const char *hwgen_last_error(void) { return gen::g_err.c_str(); }
}
