// sd_internal.h -- the few host helpers the translation units of libsafedreamer.so share (defined in sd_api.cu).
#pragma once
#include <stdint.h>

// Records the message for sd_last_error_string() and returns `code`.
int sd_fail(int code, const char* fmt, ...);
// Adds to the process-wide kernel launch counter behind sd_launch_count().
void sd_count_launches(uint64_t n);

#define SD_CUDA_TRY(expr)                                                                                         \
  do {                                                                                                            \
    cudaError_t e_ = (expr);                                                                                      \
    if (e_ != cudaSuccess)                                                                                        \
      return sd_fail(SD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);    \
  } while (0)
