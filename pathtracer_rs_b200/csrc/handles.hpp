// What the translation units behind the C ABI share about its handles and its error channel.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/ptrs_b200.h"

struct PtrsFilm {
  int device = 0;
  int width = 0, height = 0;
  float4* d = nullptr;
  bool owned = false;
};

namespace ptrs {
// records the thread-local message ptrs_last_error() returns and hands the status code back
int32_t set_error(int32_t code, const std::string& msg);
}  // namespace ptrs
