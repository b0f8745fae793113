// JPEG in: what `image::open(..)` hands the importers for .jpg textures (glTF images, Mitsuba bitmaps;
// src/pathtracer/importer/gltf.rs:39-96, src/pathtracer/importer/mitsuba.rs:104-117 through image 0.23.14 /
// jpeg-decoder, neither in the checkout).  Restated from ITU-T T.81: baseline, extended-sequential and
// progressive DCT, 8-bit samples, Huffman coding, restart intervals, 1 (grey) or 3 (YCbCr / Adobe RGB)
// components with sampling factors 1 or 2.  Arithmetic after entropy decoding follows the IJG conventions
// (13-bit fixed-point "islow" inverse DCT, triangle-filter chroma upsampling, 16-bit fixed-point YCbCr -> RGB),
// so pixels equal libjpeg's; the reference's own decoder uses a different fixed-point IDCT and may differ by a
// few levels out of 255 (parity unpinned).  Arithmetic coding, 12-bit samples, lossless and
// hierarchical modes, and 4-component (CMYK) files throw.
#include <cstring>
#include <stdexcept>

#include "image_io.hpp"

namespace ptrs_host {
namespace {

[[noreturn]] void bad(const std::string& m) { throw std::runtime_error("jpeg: " + m); }

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huffman {
  bool present = false;
  uint8_t vals[256];
  int mincode[17], maxcode[18], valptr[17];
  uint16_t look[512];  // 9-bit prefix -> (length << 8) | symbol, 0 = longer than 9 bits
  void build(const uint8_t counts[16], const uint8_t* symbols, int n_symbols) {
    present = true;
    std::memcpy(vals, symbols, (size_t)n_symbols);
    std::memset(look, 0, sizeof look);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
      valptr[len] = k;
      mincode[len] = code;
      if (code + counts[len - 1] > (1 << len)) bad("over-subscribed Huffman table");
      for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
        if (len <= 9) {
          const int first = code << (9 - len);
          for (int f = 0; f < (1 << (9 - len)); ++f) look[first + f] = (uint16_t)((len << 8) | symbols[k]);
        }
      }
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0;
  int td = 0, ta = 0;        // tables of the current scan
  int blocks_w = 0, blocks_h = 0;    // padded to whole MCUs (coefficient storage)
  int real_bw = 0, real_bh = 0;      // blocks that cover the component's own samples (non-interleaved scans)
  int samp_w = 0, samp_h = 0;        // downsampled_width / height
  std::vector<int16_t> coef;         // blocks_w * blocks_h * 64, natural order
  std::vector<uint8_t> plane;        // blocks_w * 8 wide after the inverse DCT
  int pred = 0;
};

struct Decoder {
  const uint8_t* p;
  const uint8_t* end;
  // entropy-coded segment reader
  uint32_t bitbuf = 0;
  int bitcnt = 0;
  bool hit_marker = false;

  int width = 0, height = 0, hmax = 1, vmax = 1;
  bool progressive = false, have_frame = false;
  int adobe_transform = -1;
  int restart_interval = 0;
  uint16_t quant[4][64];
  bool quant_present[4] = {false, false, false, false};
  Huffman dc[4], ac[4];
  std::vector<Component> comps;
  int eobrun = 0;

  Decoder(const uint8_t* b, size_t n) : p(b), end(b + n) {}

  int u8() {
    if (p >= end) bad("truncated file");
    return *p++;
  }
  int u16() {
    const int a = u8();
    return (a << 8) | u8();
  }

  // ---- bits ------------------------------------------------------------------------------------------
  void fill() {
    while (bitcnt <= 24) {
      int b = 0;
      if (!hit_marker && p < end) {
        b = *p;
        if (b == 0xFF) {
          const int n = p + 1 < end ? p[1] : 0xD9;
          if (n == 0x00) {
            p += 2;
          } else {
            hit_marker = true;  // leave the marker in place; feed zeros
            b = 0;
          }
        } else {
          ++p;
        }
      }
      bitbuf |= (uint32_t)b << (24 - bitcnt);
      bitcnt += 8;
    }
  }
  int get_bits(int n) {
    if (n == 0) return 0;
    if (n < 0 || n > 16) bad("bad bit count");  // never shift by 32 or more, whatever the tables say
    if (bitcnt < n) fill();
    const int v = (int)(bitbuf >> (32 - n));
    bitbuf <<= n;
    bitcnt -= n;
    return v;
  }
  int get_bit() { return get_bits(1); }
  int decode(const Huffman& h) {
    if (bitcnt < 16) fill();
    const uint16_t l = h.look[bitbuf >> 23];
    if (l) {
      const int len = l >> 8;
      bitbuf <<= len;
      bitcnt -= len;
      return l & 0xff;
    }
    int code = (int)(bitbuf >> 22), len = 10;  // first 10 bits
    for (; len <= 16; ++len) {
      if (code <= h.maxcode[len]) break;
      code = (int)(bitbuf >> (32 - len - 1));
    }
    if (len > 16) bad("bad Huffman code");
    bitbuf <<= len;
    bitcnt -= len;
    return h.vals[h.valptr[len] + code - h.mincode[len]];
  }
  static int extend(int v, int s) { return v < (1 << (s - 1)) ? v + (int)((~0u) << s) + 1 : v; }
  int receive_extend(int s) { return s ? extend(get_bits(s), s) : 0; }
  void reset_bits() {
    bitbuf = 0;
    bitcnt = 0;
    hit_marker = false;
  }

  // ---- segments --------------------------------------------------------------------------------------
  void read_dqt(int len) {
    while (len > 0) {
      const int pq_tq = u8();
      const int pq = pq_tq >> 4, tq = pq_tq & 15;
      if (tq > 3 || pq > 1) bad("bad quantisation table");
      for (int i = 0; i < 64; ++i) quant[tq][kZigzag[i]] = (uint16_t)(pq ? u16() : u8());
      quant_present[tq] = true;
      len -= 1 + 64 * (pq + 1);
    }
  }
  void read_dht(int len) {
    while (len > 0) {
      const int tc_th = u8();
      const int tc = tc_th >> 4, th = tc_th & 15;
      if (tc > 1 || th > 3) bad("bad Huffman table id");
      uint8_t counts[16], symbols[256];
      int total = 0;
      for (int i = 0; i < 16; ++i) total += counts[i] = (uint8_t)u8();
      if (total > 256) bad("bad Huffman table");
      for (int i = 0; i < total; ++i) symbols[i] = (uint8_t)u8();
      (tc ? ac[th] : dc[th]).build(counts, symbols, total);
      len -= 17 + total;
    }
  }
  void read_sof(int marker) {
    if (have_frame) bad("more than one frame");
    progressive = marker == 0xC2;
    if (u8() != 8) bad("only 8-bit samples are supported");
    height = u16();
    width = u16();
    const int nc = u8();
    if (width <= 0 || height <= 0) bad("empty image");
    if ((uint64_t)width * (uint64_t)height > (1ull << 28)) bad("image larger than 2^28 pixels");  // a corrupt header must not become a 100 GB allocation
    if (nc != 1 && nc != 3) bad("only 1- and 3-component files are supported");
    comps.resize((size_t)nc);
    for (auto& c : comps) {
      c.id = u8();
      const int hv = u8();
      c.h = hv >> 4;
      c.v = hv & 15;
      c.tq = u8();
      if (c.h < 1 || c.h > 2 || c.v < 1 || c.v > 2 || c.tq > 3) bad("unsupported sampling factors");
      hmax = std::max(hmax, c.h);
      vmax = std::max(vmax, c.v);
    }
    if (nc == 1) comps[0].h = comps[0].v = hmax = vmax = 1;  // a single component is never interleaved
    const int mcus_x = (width + 8 * hmax - 1) / (8 * hmax), mcus_y = (height + 8 * vmax - 1) / (8 * vmax);
    for (auto& c : comps) {
      c.blocks_w = mcus_x * c.h;
      c.blocks_h = mcus_y * c.v;
      c.samp_w = (width * c.h + hmax - 1) / hmax;
      c.samp_h = (height * c.v + vmax - 1) / vmax;
      c.real_bw = (c.samp_w + 7) / 8;
      c.real_bh = (c.samp_h + 7) / 8;
      c.coef.assign((size_t)c.blocks_w * c.blocks_h * 64, 0);
    }
    have_frame = true;
  }

  // ---- one block of one scan -------------------------------------------------------------------------
  void block_sequential(Component& c, int16_t* b) {
    const int t = decode(dc[c.td]);
    if (t > 11) bad("bad DC category");  // T.81 F.1.2.1: SSSS <= 11 for 8-bit samples; a corrupt DHT can hold any byte
    c.pred += receive_extend(t);
    b[0] = (int16_t)c.pred;
    for (int k = 1; k < 64;) {
      const int rs = decode(ac[c.ta]);
      const int r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (r != 15) break;
        k += 16;
      } else {
        k += r;
        if (k > 63) bad("coefficient index out of range");
        b[kZigzag[k]] = (int16_t)receive_extend(s);
        ++k;
      }
    }
  }
  void block_dc_first(Component& c, int16_t* b, int al) {
    const int t = decode(dc[c.td]);
    if (t > 11) bad("bad DC category");
    c.pred += receive_extend(t);
    b[0] = (int16_t)(c.pred * (1 << al));
  }
  void block_dc_refine(int16_t* b, int al) {
    if (get_bit()) b[0] |= (int16_t)(1 << al);
  }
  void block_ac_first(Component& c, int16_t* b, int ss, int se, int al) {
    if (eobrun > 0) {
      --eobrun;
      return;
    }
    for (int k = ss; k <= se;) {
      const int rs = decode(ac[c.ta]);
      const int r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (r < 15) {
          eobrun = (1 << r) - 1;
          if (r) eobrun += get_bits(r);
          break;
        }
        k += 16;
      } else {
        k += r;
        if (k > 63) bad("coefficient index out of range");
        b[kZigzag[k]] = (int16_t)(receive_extend(s) * (1 << al));
        ++k;
      }
    }
  }
  void block_ac_refine(Component& c, int16_t* b, int ss, int se, int al) {
    const int p1 = 1 << al, m1 = -(1 << al);
    int k = ss;
    auto correct = [&](int16_t& coef) {
      if (get_bit() && (coef & p1) == 0) coef = (int16_t)(coef + (coef >= 0 ? p1 : m1));
    };
    if (eobrun == 0) {
      for (; k <= se; ++k) {
        const int rs = decode(ac[c.ta]);
        int r = rs >> 4, s = rs & 15;
        if (s) {
          s = get_bit() ? p1 : m1;
        } else if (r != 15) {
          eobrun = 1 << r;
          if (r) eobrun += get_bits(r);
          break;
        }
        while (k <= se) {
          int16_t& coef = b[kZigzag[k]];
          if (coef != 0) {
            correct(coef);
          } else if (--r < 0) {
            break;
          }
          ++k;
        }
        if (s) {
          if (k > 63) bad("coefficient index out of range");
          b[kZigzag[k]] = (int16_t)s;
        }
      }
    }
    if (eobrun > 0) {
      for (; k <= se; ++k) {
        int16_t& coef = b[kZigzag[k]];
        if (coef != 0) correct(coef);
      }
      --eobrun;
    }
  }

  // ---- scan ------------------------------------------------------------------------------------------
  void read_sos() {
    if (!have_frame) bad("scan before frame header");
    const int ns = u8();
    if (ns < 1 || ns > (int)comps.size()) bad("bad component count in scan");
    std::vector<Component*> sc;
    for (int i = 0; i < ns; ++i) {
      const int id = u8(), tt = u8();
      Component* c = nullptr;
      for (auto& k : comps)
        if (k.id == id) c = &k;
      if (!c) bad("scan names an unknown component");
      c->td = tt >> 4;
      c->ta = tt & 15;
      if (c->td > 3 || c->ta > 3) bad("bad table selector");
      sc.push_back(c);
    }
    const int ss = u8(), se = u8(), ahal = u8();
    const int ah = ahal >> 4, al = ahal & 15;
    if (progressive) {
      if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1)) bad("bad progressive scan parameters");
    } else if (ss != 0 || se != 63 || ah != 0 || al != 0) {
      bad("bad sequential scan parameters");
    }
    for (Component* c : sc) {
      const bool need_dc = !progressive || (ss == 0 && ah == 0), need_ac = !progressive || ss > 0;
      if (need_dc && !dc[c->td].present) bad("missing DC Huffman table");
      if (need_ac && !ac[c->ta].present) bad("missing AC Huffman table");
    }
    auto do_block = [&](Component& c, int bx, int by) {
      int16_t* b = &c.coef[((size_t)by * c.blocks_w + bx) * 64];
      if (!progressive) block_sequential(c, b);
      else if (ss == 0) (ah == 0 ? block_dc_first(c, b, al) : block_dc_refine(b, al));
      else (ah == 0 ? block_ac_first(c, b, ss, se, al) : block_ac_refine(c, b, ss, se, al));
    };
    reset_bits();
    eobrun = 0;
    for (auto& c : comps) c.pred = 0;
    int until_restart = restart_interval, next_rst = 0;
    auto restart_if_due = [&](bool more) {
      if (!restart_interval || --until_restart > 0 || !more) return;
      // byte-align, expect RSTn
      reset_bits();
      while (p < end && *p != 0xFF) ++p;  // tolerate stray bytes before the marker
      while (p + 1 < end && p[0] == 0xFF && p[1] == 0xFF) ++p;
      if (p + 1 < end && p[0] == 0xFF && p[1] == 0xD0 + next_rst) p += 2;
      else if (!(p + 1 < end && p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) bad("missing restart marker");
      else p += 2;
      next_rst = (next_rst + 1) & 7;
      until_restart = restart_interval;
      eobrun = 0;
      for (auto& c : comps) c.pred = 0;
    };
    if (ns == 1) {
      Component& c = *sc[0];
      const int total = c.real_bw * c.real_bh;
      for (int i = 0; i < total; ++i) {
        do_block(c, i % c.real_bw, i / c.real_bw);
        restart_if_due(i + 1 < total);
      }
    } else {
      const int mcus_x = comps[0].blocks_w / comps[0].h, mcus_y = comps[0].blocks_h / comps[0].v;
      for (int my = 0; my < mcus_y; ++my)
        for (int mx = 0; mx < mcus_x; ++mx) {
          for (Component* c : sc)
            for (int v = 0; v < c->v; ++v)
              for (int h = 0; h < c->h; ++h) do_block(*c, mx * c->h + h, my * c->v + v);
          restart_if_due(!(my == mcus_y - 1 && mx == mcus_x - 1));
        }
    }
    // position p at the next marker (the reader stops in front of it)
    while (p < end && !(p[0] == 0xFF && p + 1 < end && p[1] != 0x00 && !(p[1] >= 0xD0 && p[1] <= 0xD7) && p[1] != 0xFF)) ++p;
  }

  // ---- reconstruction --------------------------------------------------------------------------------
  static uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
  // IJG jidctint ("islow"): 13-bit constants, 2 extra bits kept after the column pass
  static void idct(const int16_t* in, const uint16_t* q, uint8_t* out, int stride) {
    constexpr int CB = 13, P1 = 2;
    constexpr int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137,
                  F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
    int ws[64];
    auto descale = [](long x, int n) { return (int)((x + (1L << (n - 1))) >> n); };
    for (int c = 0; c < 8; ++c) {
      const int16_t* i = in + c;
      const uint16_t* qq = q + c;
      auto d = [&](int r) { return (long)i[8 * r] * qq[8 * r]; };
      int* w = ws + c;
      if (!i[8] && !i[16] && !i[24] && !i[32] && !i[40] && !i[48] && !i[56]) {
        const int dcv = (int)(d(0) * (1 << P1));
        for (int r = 0; r < 8; ++r) w[8 * r] = dcv;
        continue;
      }
      long z2 = d(2), z3 = d(6);
      long z1 = (z2 + z3) * F0541;
      long tmp2 = z1 + z3 * (-F1847), tmp3 = z1 + z2 * F0765;
      z2 = d(0);
      z3 = d(4);
      long tmp0 = (z2 + z3) * (1L << CB), tmp1 = (z2 - z3) * (1L << CB);
      const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
      tmp0 = d(7);
      tmp1 = d(5);
      tmp2 = d(3);
      tmp3 = d(1);
      z1 = tmp0 + tmp3;
      z2 = tmp1 + tmp2;
      z3 = tmp0 + tmp2;
      long z4 = tmp1 + tmp3;
      const long z5 = (z3 + z4) * F1175;
      tmp0 *= F0298;
      tmp1 *= F2053;
      tmp2 *= F3072;
      tmp3 *= F1501;
      z1 *= -F0899;
      z2 *= -F2562;
      z3 *= -F1961;
      z4 *= -F0390;
      z3 += z5;
      z4 += z5;
      tmp0 += z1 + z3;
      tmp1 += z2 + z4;
      tmp2 += z2 + z3;
      tmp3 += z1 + z4;
      w[0] = descale(tmp10 + tmp3, CB - P1);
      w[56] = descale(tmp10 - tmp3, CB - P1);
      w[8] = descale(tmp11 + tmp2, CB - P1);
      w[48] = descale(tmp11 - tmp2, CB - P1);
      w[16] = descale(tmp12 + tmp1, CB - P1);
      w[40] = descale(tmp12 - tmp1, CB - P1);
      w[24] = descale(tmp13 + tmp0, CB - P1);
      w[32] = descale(tmp13 - tmp0, CB - P1);
    }
    for (int r = 0; r < 8; ++r) {
      const int* w = ws + 8 * r;
      uint8_t* o = out + (size_t)r * stride;
      long z2 = w[2], z3 = w[6];
      long z1 = (z2 + z3) * F0541;
      long tmp2 = z1 + z3 * (-F1847), tmp3 = z1 + z2 * F0765;
      long tmp0 = ((long)w[0] + w[4]) * (1L << CB), tmp1 = ((long)w[0] - w[4]) * (1L << CB);
      const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
      tmp0 = w[7];
      tmp1 = w[5];
      tmp2 = w[3];
      tmp3 = w[1];
      z1 = tmp0 + tmp3;
      z2 = tmp1 + tmp2;
      z3 = tmp0 + tmp2;
      long z4 = tmp1 + tmp3;
      const long z5 = (z3 + z4) * F1175;
      tmp0 *= F0298;
      tmp1 *= F2053;
      tmp2 *= F3072;
      tmp3 *= F1501;
      z1 *= -F0899;
      z2 *= -F2562;
      z3 *= -F1961;
      z4 *= -F0390;
      z3 += z5;
      z4 += z5;
      tmp0 += z1 + z3;
      tmp1 += z2 + z4;
      tmp2 += z2 + z3;
      tmp3 += z1 + z4;
      constexpr int S = CB + P1 + 3;
      o[0] = clamp8(descale(tmp10 + tmp3, S) + 128);
      o[7] = clamp8(descale(tmp10 - tmp3, S) + 128);
      o[1] = clamp8(descale(tmp11 + tmp2, S) + 128);
      o[6] = clamp8(descale(tmp11 - tmp2, S) + 128);
      o[2] = clamp8(descale(tmp12 + tmp1, S) + 128);
      o[5] = clamp8(descale(tmp12 - tmp1, S) + 128);
      o[3] = clamp8(descale(tmp13 + tmp0, S) + 128);
      o[4] = clamp8(descale(tmp13 - tmp0, S) + 128);
    }
  }

  // component plane -> full resolution (width x height), IJG "fancy" (triangle filter) upsampling for factor 2
  std::vector<uint8_t> upsample(const Component& c) const {
    const int stride = c.blocks_w * 8;
    const int fx = hmax / c.h, fy = vmax / c.v;
    std::vector<uint8_t> out((size_t)width * height);
    if (fx == 1 && fy == 1) {
      for (int y = 0; y < height; ++y) std::memcpy(&out[(size_t)y * width], &c.plane[(size_t)y * stride], (size_t)width);
      return out;
    }
    const int sw = c.samp_w, sh = c.samp_h;
    if (fx == 2 && sw <= 2) {  // the IJG decoder only filters rows of more than two samples; narrower ones are replicated
      for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) out[(size_t)y * width + x] = c.plane[(size_t)(y / fy) * stride + x / 2];
      return out;
    }
    std::vector<int> colsum((size_t)sw);
    std::vector<uint8_t> row((size_t)sw * 2 + 2);
    for (int y = 0; y < height; ++y) {
      const uint8_t* near_row;
      if (fy == 2) {
        // output rows 2i, 2i+1 come from input row i with the row above / below as the far row (3:1)
        const int i = y >> 1;
        const int far = (y & 1) ? std::min(i + 1, sh - 1) : std::max(i - 1, 0);
        const uint8_t* a = &c.plane[(size_t)i * stride];
        const uint8_t* b = &c.plane[(size_t)far * stride];
        for (int x = 0; x < sw; ++x) colsum[(size_t)x] = 3 * a[x] + b[x];
        if (fx == 2) {
          // h2v2: (3 * this + neighbour + 8 | 7) >> 4 on the column sums
          {
            row[0] = (uint8_t)((colsum[0] * 4 + 8) >> 4);
            row[1] = (uint8_t)((colsum[0] * 3 + colsum[1] + 7) >> 4);
            for (int x = 1; x < sw - 1; ++x) {
              row[(size_t)2 * x] = (uint8_t)((colsum[(size_t)x] * 3 + colsum[(size_t)x - 1] + 8) >> 4);
              row[(size_t)2 * x + 1] = (uint8_t)((colsum[(size_t)x] * 3 + colsum[(size_t)x + 1] + 7) >> 4);
            }
            row[(size_t)2 * (sw - 1)] = (uint8_t)((colsum[(size_t)sw - 1] * 3 + colsum[(size_t)sw - 2] + 8) >> 4);
            row[(size_t)2 * (sw - 1) + 1] = (uint8_t)((colsum[(size_t)sw - 1] * 4 + 7) >> 4);
          }
          std::memcpy(&out[(size_t)y * width], row.data(), (size_t)width);
        } else {
          // h1v2: (3 * near + far + 1 | 2) >> 2, bias alternating by output row
          const int bias = (y & 1) ? 2 : 1;
          for (int x = 0; x < width; ++x) out[(size_t)y * width + x] = (uint8_t)((colsum[(size_t)x] + bias) >> 2);
        }
        continue;
      }
      near_row = &c.plane[(size_t)y * stride];
      // h2v1
      {
        row[0] = near_row[0];
        row[1] = (uint8_t)((near_row[0] * 3 + near_row[1] + 2) >> 2);
        for (int x = 1; x < sw - 1; ++x) {
          row[(size_t)2 * x] = (uint8_t)((near_row[x] * 3 + near_row[x - 1] + 1) >> 2);
          row[(size_t)2 * x + 1] = (uint8_t)((near_row[x] * 3 + near_row[x + 1] + 2) >> 2);
        }
        row[(size_t)2 * (sw - 1)] = (uint8_t)((near_row[sw - 1] * 3 + near_row[sw - 2] + 1) >> 2);
        row[(size_t)2 * (sw - 1) + 1] = near_row[sw - 1];
      }
      std::memcpy(&out[(size_t)y * width], row.data(), (size_t)width);
    }
    return out;
  }

  ImageU8 finish() {
    if (!have_frame) bad("no frame header");
    for (auto& c : comps) {
      if (!quant_present[c.tq]) bad("missing quantisation table");
      c.plane.assign((size_t)c.blocks_w * 8 * c.blocks_h * 8, 0);
      for (int by = 0; by < c.blocks_h; ++by)
        for (int bx = 0; bx < c.blocks_w; ++bx)
          idct(&c.coef[((size_t)by * c.blocks_w + bx) * 64], quant[c.tq], &c.plane[((size_t)by * 8 * c.blocks_w + bx) * 8], c.blocks_w * 8);
      std::vector<int16_t>().swap(c.coef);
    }
    ImageU8 img;
    img.width = width;
    img.height = height;
    if (comps.size() == 1) {
      img.channels = 1;
      img.data = upsample(comps[0]);
      return img;
    }
    img.channels = 3;
    img.data.resize((size_t)width * height * 3);
    const std::vector<uint8_t> a = upsample(comps[0]), b = upsample(comps[1]), c = upsample(comps[2]);
    // Adobe transform 0 = stored as RGB; otherwise (JFIF, or Adobe transform 1) YCbCr.  Without either marker,
    // component ids 'R','G','B' mean RGB (the IJG rule).
    bool rgb = adobe_transform == 0;
    if (adobe_transform < 0 && comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B') rgb = true;
    const size_t n = (size_t)width * height;
    if (rgb) {
      for (size_t i = 0; i < n; ++i) {
        img.data[3 * i] = a[i];
        img.data[3 * i + 1] = b[i];
        img.data[3 * i + 2] = c[i];
      }
      return img;
    }
    // IJG jdcolor: 16-bit fixed point, FIX(x) = (int)(x * 65536 + 0.5)
    constexpr int F1402 = 91881, F1772 = 116130, F0714 = 46802, F0344 = 22554, HALF = 32768;
    for (size_t i = 0; i < n; ++i) {
      const int y = a[i], cb = b[i] - 128, cr = c[i] - 128;
      img.data[3 * i] = clamp8(y + ((F1402 * cr + HALF) >> 16));
      img.data[3 * i + 1] = clamp8(y + ((-F0344 * cb + HALF - F0714 * cr) >> 16));
      img.data[3 * i + 2] = clamp8(y + ((F1772 * cb + HALF) >> 16));
    }
    return img;
  }

  ImageU8 run() {
    if (u8() != 0xFF || u8() != 0xD8) bad("not a JPEG file");
    for (;;) {
      int m = u8();
      if (m != 0xFF) continue;  // resynchronise
      do m = u8();
      while (m == 0xFF);
      if (m == 0x00 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
      if (m == 0xD9) break;  // EOI
      const int len = u16() - 2;
      if (len < 0 || p + len > end) bad("truncated segment");
      const uint8_t* next = p + len;
      switch (m) {
        case 0xDB: read_dqt(len); break;
        case 0xC4: read_dht(len); break;
        case 0xC0: case 0xC1: case 0xC2: read_sof(m); break;
        case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
          bad("lossless / hierarchical modes are not supported");
        case 0xC9: case 0xCA: case 0xCC: bad("arithmetic coding is not supported");
        case 0xDD: restart_interval = u16(); break;
        case 0xEE:
          if (len >= 12 && !std::memcmp(p, "Adobe", 5)) adobe_transform = p[11];
          break;
        case 0xDA:
          read_sos();
          next = p;  // the scan reader already stands in front of the next marker
          break;
        default: break;  // APPn, COM, DNL ...
      }
      p = next;
      if (p >= end) break;  // missing EOI: use what was decoded
    }
    return finish();
  }
};

}  // namespace

ImageU8 decode_jpeg(const uint8_t* bytes, size_t n) {
  Decoder d(bytes, n);
  return d.run();
}

ImageU8 decode_image(const uint8_t* bytes, size_t n) {
  if (n >= 3 && bytes[0] == 0xFF && bytes[1] == 0xD8) return decode_jpeg(bytes, n);
  return decode_png(bytes, n);
}
ImageU8 load_image(const std::string& path) {
  const std::vector<uint8_t> b = read_file(path);
  return decode_image(b.data(), b.size());
}

}  // namespace ptrs_host
