#include "image_io.hpp"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace ptrs_host {

std::vector<uint8_t> read_file(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<uint8_t> out;
  uint8_t buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) out.insert(out.end(), buf, buf + n);
  std::fclose(f);
  return out;
}
static void write_file(const std::string& path, const std::vector<uint8_t>& bytes) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot create " + path);
  const size_t n = std::fwrite(bytes.data(), 1, bytes.size(), f);
  std::fclose(f);
  if (n != bytes.size()) throw std::runtime_error("short write to " + path);
}

// ------------------------------------------------------------------------------------------------
// Radiance RGBE
// ------------------------------------------------------------------------------------------------
namespace {
struct Reader {
  const uint8_t* p;
  size_t n, i = 0;
  bool line(std::string* out) {
    if (i >= n) return false;
    out->clear();
    while (i < n && p[i] != '\n') out->push_back((char)p[i++]);
    if (i < n) ++i;
    if (!out->empty() && out->back() == '\r') out->pop_back();
    return true;
  }
  uint8_t byte() {
    if (i >= n) throw std::runtime_error("HDR: unexpected end of file");
    return p[i++];
  }
};

// one scanline of `w` RGBE quads: new-style per-component RLE, old-style repeat markers, or flat
void read_scanline(Reader& r, int w, uint8_t* out) {
  if (w >= 8 && w < 32768 && r.i + 4 <= r.n && r.p[r.i] == 2 && r.p[r.i + 1] == 2 && !(r.p[r.i + 2] & 0x80)) {
    const int len = (r.p[r.i + 2] << 8) | r.p[r.i + 3];
    if (len != w) throw std::runtime_error("HDR: scanline length mismatch");
    r.i += 4;
    for (int c = 0; c < 4; ++c) {
      int x = 0;
      while (x < w) {
        int count = r.byte();
        if (count > 128) {  // run
          count -= 128;
          const uint8_t v = r.byte();
          if (count == 0 || x + count > w) throw std::runtime_error("HDR: bad run length");
          for (int k = 0; k < count; ++k) out[4 * (x++) + c] = v;
        } else {  // literal
          if (count == 0 || x + count > w) throw std::runtime_error("HDR: bad literal length");
          for (int k = 0; k < count; ++k) out[4 * (x++) + c] = r.byte();
        }
      }
    }
    return;
  }
  int x = 0, shift = 0;
  while (x < w) {  // flat or old-style RLE ((1,1,1,n) repeats the previous pixel n << shift times)
    uint8_t q[4] = {r.byte(), r.byte(), r.byte(), r.byte()};
    if (q[0] == 1 && q[1] == 1 && q[2] == 1) {
      if (x == 0) throw std::runtime_error("HDR: repeat marker at the start of a scanline");
      if (shift > 24) throw std::runtime_error("HDR: bad old-style run");  // a fifth consecutive repeat marker would shift out of the int
      const int64_t count = (int64_t)q[3] << shift;
      if (x + count > w) throw std::runtime_error("HDR: bad old-style run");
      for (int64_t k = 0; k < count; ++k, ++x) std::memcpy(out + 4 * x, out + 4 * (x - 1), 4);
      shift += 8;
    } else {
      std::memcpy(out + 4 * x, q, 4);
      ++x;
      shift = 0;
    }
  }
}
}  // namespace

ImageF32 decode_hdr(const uint8_t* bytes, size_t n) {
  Reader r{bytes, n};
  std::string ln;
  if (!r.line(&ln) || (ln.rfind("#?RADIANCE", 0) != 0 && ln.rfind("#?RGBE", 0) != 0)) throw std::runtime_error("HDR: missing #?RADIANCE signature");
  bool format_ok = false;
  for (;;) {
    if (!r.line(&ln)) throw std::runtime_error("HDR: truncated header");
    if (ln.empty()) break;
    if (ln.rfind("FORMAT=", 0) == 0) {
      if (ln != "FORMAT=32-bit_rle_rgbe") throw std::runtime_error("HDR: unsupported " + ln);
      format_ok = true;
    }
  }
  (void)format_ok;  // HdrDecoder tolerates a missing FORMAT line in non-strict mode
  if (!r.line(&ln)) throw std::runtime_error("HDR: missing resolution line");
  int h = 0, w = 0;
  if (std::sscanf(ln.c_str(), "-Y %d +X %d", &h, &w) != 2 || h <= 0 || w <= 0) throw std::runtime_error("HDR: unsupported orientation '" + ln + "'");
  ImageF32 img;
  img.width = w;
  img.height = h;
  img.channels = 3;
  img.data.resize((size_t)w * h * 3);
  std::vector<uint8_t> row((size_t)w * 4);
  for (int y = 0; y < h; ++y) {
    read_scanline(r, w, row.data());
    float* dst = img.data.data() + (size_t)y * w * 3;
    for (int x = 0; x < w; ++x) {
      const uint8_t* q = row.data() + 4 * x;
      if (q[3] == 0) {
        dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = 0.0f;
      } else {
        const float e = std::exp2((float)q[3] - (128.0f + 8.0f));
        dst[3 * x] = e * (float)q[0];
        dst[3 * x + 1] = e * (float)q[1];
        dst[3 * x + 2] = e * (float)q[2];
      }
    }
  }
  return img;
}
ImageF32 load_hdr(const std::string& path) {
  const std::vector<uint8_t> b = read_file(path);
  return decode_hdr(b.data(), b.size());
}

std::vector<uint8_t> encode_hdr(const float* rgb, int width, int height) {
  std::string head = "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y " + std::to_string(height) + " +X " + std::to_string(width) + "\n";
  std::vector<uint8_t> out(head.begin(), head.end());
  out.reserve(out.size() + (size_t)width * height * 4);
  for (size_t i = 0; i < (size_t)width * height; ++i) {
    const float r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    const float v = std::fmax(r, std::fmax(g, b));
    uint8_t q[4] = {0, 0, 0, 0};
    if (v >= 1e-32f) {
      int e;
      const float m = std::frexp(v, &e) * 256.0f / v;
      q[0] = (uint8_t)(r * m);
      q[1] = (uint8_t)(g * m);
      q[2] = (uint8_t)(b * m);
      q[3] = (uint8_t)(e + 128);
    }
    // a scanline that starts with (2, 2, <128) would read as run-length coded: nudge such a pixel
    if (i % (size_t)width == 0 && q[0] == 2 && q[1] == 2 && !(q[2] & 0x80)) q[0] = 3;
    if (q[0] == 1 && q[1] == 1 && q[2] == 1) q[0] = 2;  // (1, 1, 1, n) is the old-style repeat marker
    out.insert(out.end(), q, q + 4);
  }
  return out;
}
void save_hdr(const std::string& path, const float* rgb, int width, int height) { write_file(path, encode_hdr(rgb, width, height)); }

// ------------------------------------------------------------------------------------------------
// PNG
// ------------------------------------------------------------------------------------------------
namespace {
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24));
  v.push_back((uint8_t)(x >> 16));
  v.push_back((uint8_t)(x >> 8));
  v.push_back((uint8_t)x);
}
void put_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n) {
  put32(out, (uint32_t)n);
  const size_t at = out.size();
  out.insert(out.end(), type, type + 4);
  if (n) out.insert(out.end(), data, data + n);
  put32(out, (uint32_t)crc32(0L, out.data() + at, (uInt)(n + 4)));
}
int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
}  // namespace

ImageU8 decode_png(const uint8_t* bytes, size_t n) {
  if (n < 8 || std::memcmp(bytes, kPngSig, 8) != 0) throw std::runtime_error("PNG: bad signature");
  size_t i = 8;
  int w = 0, h = 0, depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  bool end = false;
  while (!end && i + 12 <= n) {
    const uint32_t len = be32(bytes + i);
    const uint8_t* type = bytes + i + 4;
    const uint8_t* data = bytes + i + 8;
    if (i + 12 + (size_t)len > n) throw std::runtime_error("PNG: truncated chunk");
    if (be32(data + len) != (uint32_t)crc32(0L, type, (uInt)(len + 4))) throw std::runtime_error("PNG: chunk CRC mismatch");
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len != 13) throw std::runtime_error("PNG: bad IHDR");
      w = (int)be32(data);
      h = (int)be32(data + 4);
      depth = data[8];
      ctype = data[9];
      interlace = data[12];
    } else if (!std::memcmp(type, "PLTE", 4)) {
      plte.assign(data, data + len);
    } else if (!std::memcmp(type, "tRNS", 4)) {
      trns.assign(data, data + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      end = true;
    }
    i += 12 + (size_t)len;
  }
  if (w <= 0 || h <= 0 || ctype < 0) throw std::runtime_error("PNG: missing IHDR");
  if (depth != 8 || interlace != 0) throw std::runtime_error("PNG: only 8-bit non-interlaced images are supported");
  int spp;  // samples per pixel in the file
  switch (ctype) {
    case 0: spp = 1; break;
    case 2: spp = 3; break;
    case 3: spp = 1; break;
    case 4: spp = 2; break;
    case 6: spp = 4; break;
    default: throw std::runtime_error("PNG: unknown colour type");
  }
  const size_t stride = (size_t)w * spp;
  std::vector<uint8_t> raw((stride + 1) * (size_t)h);
  uLongf raw_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) throw std::runtime_error("PNG: inflate failed");
  std::vector<uint8_t> pix(stride * (size_t)h);
  for (int y = 0; y < h; ++y) {
    const uint8_t ft = raw[(stride + 1) * (size_t)y];
    const uint8_t* src = raw.data() + (stride + 1) * (size_t)y + 1;
    uint8_t* cur = pix.data() + stride * (size_t)y;
    const uint8_t* up = y ? cur - stride : nullptr;
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= (size_t)spp ? cur[x - spp] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)spp) ? up[x - spp] : 0;
      int v = src[x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: throw std::runtime_error("PNG: bad filter type");
      }
      cur[x] = (uint8_t)v;
    }
  }
  ImageU8 img;
  img.width = w;
  img.height = h;
  if (ctype == 3) {
    const bool alpha = !trns.empty();
    img.channels = alpha ? 4 : 3;
    img.data.resize((size_t)w * h * img.channels);
    for (size_t k = 0; k < (size_t)w * h; ++k) {
      const size_t idx = pix[k];
      if (3 * idx + 2 >= plte.size()) throw std::runtime_error("PNG: palette index out of range");
      uint8_t* d = img.data.data() + k * img.channels;
      d[0] = plte[3 * idx];
      d[1] = plte[3 * idx + 1];
      d[2] = plte[3 * idx + 2];
      if (alpha) d[3] = idx < trns.size() ? trns[idx] : 255;
    }
  } else {
    img.channels = spp;
    img.data = std::move(pix);
  }
  return img;
}
ImageU8 load_png(const std::string& path) {
  const std::vector<uint8_t> b = read_file(path);
  return decode_png(b.data(), b.size());
}

std::vector<uint8_t> encode_png(const uint8_t* pixels, int width, int height, int channels) {
  if (width <= 0 || height <= 0 || channels < 1 || channels > 4) throw std::runtime_error("PNG: bad image shape");
  static const uint8_t ctype_of[5] = {0, 0, 4, 2, 6};
  std::vector<uint8_t> out(kPngSig, kPngSig + 8);
  uint8_t ihdr[13];
  std::vector<uint8_t> tmp;
  put32(tmp, (uint32_t)width);
  put32(tmp, (uint32_t)height);
  std::memcpy(ihdr, tmp.data(), 8);
  ihdr[8] = 8;
  ihdr[9] = ctype_of[channels];
  ihdr[10] = ihdr[11] = ihdr[12] = 0;
  put_chunk(out, "IHDR", ihdr, 13);
  const size_t stride = (size_t)width * channels;
  std::vector<uint8_t> raw((stride + 1) * (size_t)height);
  for (int y = 0; y < height; ++y) {  // filter type 1 (Sub) compresses rendered images better than None
    uint8_t* dst = raw.data() + (stride + 1) * (size_t)y;
    const uint8_t* src = pixels + stride * (size_t)y;
    dst[0] = 1;
    for (size_t x = 0; x < stride; ++x) dst[1 + x] = (uint8_t)(src[x] - (x >= (size_t)channels ? src[x - channels] : 0));
  }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw std::runtime_error("PNG: deflate failed");
  put_chunk(out, "IDAT", comp.data(), clen);
  put_chunk(out, "IEND", nullptr, 0);
  return out;
}
void save_png(const std::string& path, const uint8_t* pixels, int width, int height, int channels) {
  write_file(path, encode_png(pixels, width, height, channels));
}

}  // namespace ptrs_host
