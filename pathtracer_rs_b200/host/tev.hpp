// tev display-server messages of the headless mode (src/headless.rs:14-178): little-endian, length-prefixed.
//   CreateImage  u32 total | u8 4 | u8 grab_focus | name\0 | i32 w | i32 h | i32 3 | "r\0" "g\0" "b\0"
//   UpdateImage  u32 total | u8 3 | u8 grab_focus | name\0 | channel\0 | i32 x | i32 y | i32 w | i32 h | f32[w*h]
// The film is streamed as 100 x 100 tiles per channel, x-major like the reference's cartesian product.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ptrs_host {

std::vector<uint8_t> tev_create_image(int width, int height, const std::string& name);
// channels[k] = width * height floats, row-major (Film::to_channel_updates, film.rs:253-271)
std::vector<std::vector<uint8_t>> tev_update_image(const float* const channels[3], int width, int height, const std::string& name);

// blocking TCP client; connect() returns false when no server listens (headless.rs falls back to one-shot rendering)
class TevClient {
 public:
  ~TevClient();
  bool connect(const std::string& host_port);
  bool send(const std::vector<uint8_t>& msg);
  bool connected() const { return fd_ >= 0; }

 private:
  int fd_ = -1;
};

}  // namespace ptrs_host
