// Image files on either side of the rendering path:
//   Radiance .hdr (RGBE) in   what InfiniteAreaLight::new reads through image::hdr::HdrDecoder
//                              (src/pathtracer/light.rs:331-346; image 0.23.14, not in the checkout)
//   8-bit PNG / JPEG in        Mitsuba <texture type="bitmap"> / glTF images (image::open(..) -> ImageRgb8)
//   8-bit PNG out              camera.film.to_rgba_image().save("render.png") (src/headless.rs:222, 231)
// zlib does the DEFLATE part; everything else (chunks, filters, CRC, RGBE run-length coding) is here.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ptrs_host {

struct ImageF32 {  // row-major, top row first, `channels` floats per pixel
  int width = 0, height = 0, channels = 3;
  std::vector<float> data;
};
struct ImageU8 {
  int width = 0, height = 0, channels = 3;
  std::vector<uint8_t> data;
};

// Rgbe8Pixel::to_hdr of image 0.23.14: e == 0 -> black, else c * exp2(e - 136) (no +0.5 on the mantissa).
ImageF32 load_hdr(const std::string& path);
ImageF32 decode_hdr(const uint8_t* bytes, size_t n);
// flat (non run-length) RGBE; the encoder picks the largest-component exponent like Radiance's float2rgbe
std::vector<uint8_t> encode_hdr(const float* rgb, int width, int height);
void save_hdr(const std::string& path, const float* rgb, int width, int height);

// 8-bit, non-interlaced PNG of colour type 0/2/3/4/6; the result keeps the file's channel count
// (palette images expand to RGB / RGBA).  16-bit, interlaced and other formats throw.
ImageU8 load_png(const std::string& path);
ImageU8 decode_png(const uint8_t* bytes, size_t n);
std::vector<uint8_t> encode_png(const uint8_t* pixels, int width, int height, int channels);
void save_png(const std::string& path, const uint8_t* pixels, int width, int height, int channels);

// JPEG (jpeg_decode.cpp): baseline / extended-sequential / progressive Huffman, 8-bit, grey or YCbCr / RGB.
ImageU8 decode_jpeg(const uint8_t* bytes, size_t n);
// what image::open does for the importers: the format is sniffed from the first bytes (PNG or JPEG)
ImageU8 decode_image(const uint8_t* bytes, size_t n);
ImageU8 load_image(const std::string& path);

std::vector<uint8_t> read_file(const std::string& path);

}  // namespace ptrs_host
