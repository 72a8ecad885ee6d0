// Shared host/device helpers for libmoda_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "moda_math.h"

namespace moda {

// thread-local error text returned by moda_last_error()
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> error code (0 = ok)

#define MODA_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      moda::set_error(__VA_ARGS__);  \
      return -1;                     \
    }                                \
  } while (0)

constexpr int MAX_BONES = 64;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace moda
