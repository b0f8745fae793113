#include "tev.hpp"

#include <arpa/inet.h>
#include <netdb.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>

namespace ptrs_host {
namespace {
void put_i32(std::vector<uint8_t>& b, int32_t v) {
  uint8_t le[4] = {(uint8_t)v, (uint8_t)((uint32_t)v >> 8), (uint8_t)((uint32_t)v >> 16), (uint8_t)((uint32_t)v >> 24)};
  b.insert(b.end(), le, le + 4);
}
void put_cstr(std::vector<uint8_t>& b, const std::string& s) {
  b.insert(b.end(), s.begin(), s.end());
  b.push_back(0);
}
void seal(std::vector<uint8_t>& b) {  // Serialize::make_message: the first four bytes hold the total length
  const uint32_t n = (uint32_t)b.size();
  b[0] = (uint8_t)n;
  b[1] = (uint8_t)(n >> 8);
  b[2] = (uint8_t)(n >> 16);
  b[3] = (uint8_t)(n >> 24);
}
}  // namespace

std::vector<uint8_t> tev_create_image(int width, int height, const std::string& name) {
  std::vector<uint8_t> b(4, 0);
  b.push_back(4);  // TevControlHeader::CreateImage
  b.push_back(1);  // grab_focus
  put_cstr(b, name);
  put_i32(b, width);
  put_i32(b, height);
  put_i32(b, 3);
  put_cstr(b, "r");
  put_cstr(b, "g");
  put_cstr(b, "b");
  seal(b);
  return b;
}

std::vector<std::vector<uint8_t>> tev_update_image(const float* const channels[3], int width, int height, const std::string& name) {
  static const char* kNames[3] = {"r", "g", "b"};
  const int kChunk = 100;
  std::vector<std::vector<uint8_t>> out;
  for (int c = 0; c < 3; ++c)
    for (int x = 0; x < width; x += kChunk)
      for (int y = 0; y < height; y += kChunk) {
        const int rows = std::min(kChunk, height - y), cols = std::min(kChunk, width - x);
        std::vector<uint8_t> b(4, 0);
        b.reserve(64 + (size_t)rows * cols * 4);
        b.push_back(3);  // TevControlHeader::UpdateImage
        b.push_back(1);
        put_cstr(b, name);
        put_cstr(b, kNames[c]);
        put_i32(b, x);
        put_i32(b, y);
        put_i32(b, cols);
        put_i32(b, rows);
        for (int r = y; r < y + rows; ++r) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(channels[c] + (size_t)r * width + x);
          b.insert(b.end(), src, src + (size_t)cols * 4);  // f32::to_le_bytes on a little-endian host
        }
        seal(b);
        out.push_back(std::move(b));
      }
  return out;
}

TevClient::~TevClient() {
  if (fd_ >= 0) ::close(fd_);
}
bool TevClient::connect(const std::string& host_port) {
  const size_t colon = host_port.rfind(':');
  if (colon == std::string::npos) return false;
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_UNSPEC;
  hints.ai_socktype = SOCK_STREAM;
  if (getaddrinfo(host_port.substr(0, colon).c_str(), host_port.substr(colon + 1).c_str(), &hints, &res) != 0) return false;
  for (addrinfo* a = res; a; a = a->ai_next) {
    const int fd = ::socket(a->ai_family, a->ai_socktype, a->ai_protocol);
    if (fd < 0) continue;
    if (::connect(fd, a->ai_addr, a->ai_addrlen) == 0) {
      fd_ = fd;
      break;
    }
    ::close(fd);
  }
  freeaddrinfo(res);
  return fd_ >= 0;
}
bool TevClient::send(const std::vector<uint8_t>& msg) {
  size_t off = 0;
  while (off < msg.size()) {
    const ssize_t n = ::send(fd_, msg.data() + off, msg.size() - off, MSG_NOSIGNAL);
    if (n <= 0) return false;
    off += (size_t)n;
  }
  return true;
}

}  // namespace ptrs_host
