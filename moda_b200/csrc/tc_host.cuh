// Host-side helpers shared by the tensor-core translation units: TMA descriptor encoding, SM count.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace moda {
namespace tc {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp16 row-major (rows, cols) with row stride ld elements; box = 64 columns x box_rows rows, 128B swizzle
static inline int make_map(CUtensorMap* map, const void* ptr, long long rows, int cols, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  MODA_REQUIRE(enc != nullptr, "tc: cuTensorMapEncodeTiled is not available from the driver");
  MODA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld % 8) == 0 && cols % 64 == 0 && ld >= cols,
               "tc: operand (ptr %p, cols %d, ld %d) violates TMA alignment (16 B base, 16 B row pitch, 64-col chunks)",
               ptr, cols, ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MODA_REQUIRE(r == CUDA_SUCCESS, "tc: cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

// SM count of the CURRENT device (cached per device: a process may drive several)
static inline int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n;   // benign race: every writer stores the same value
  }
  return cache[dev];
}


}  // namespace tc
}  // namespace moda
