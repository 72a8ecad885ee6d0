// Micro-benchmark behind DESIGN.md section 7 ("skeleton cost" of a chain step): how long after its issue does a
// tcgen05.mma burst signal completion through tcgen05.commit -> mbarrier, as seen by (a) the issuing thread and
// (b) another warp that then reads the accumulator with tcgen05.ld?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I moda_b200/csrc tools/microbench/mma_latency.cu -o mma_latency
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace moda::tc;

__global__ void __launch_bounds__(128) k(int n_mma, int N, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ long long t_issue;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x >= 32 && threadIdx.x < 64) tmem_alloc(&slot, 512);   // a warp that has not diverged (.sync.aligned)
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t a = smem_u32(base), b = smem_u32(base + 16384);
  const uint32_t idesc = make_idesc(128, N, 0, 0);
  for (int rep = 0; rep < 3; ++rep) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint64_t ad = make_desc(a + (i & 3) * 32, 16, 1024), bd = make_desc(b + (i & 3) * 32, 16, 1024);
        umma_f16(tm, ad, bd, idesc, i ? 1u : 0u);
      }
      umma_commit(&bar);
      const long long t1 = clock64();
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      t_issue = t0;
      if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (threadIdx.x >= 64 && threadIdx.x < 96) {   // a different (whole) warp: wake-up + fence + first accumulator load
      mbar_wait(&bar, rep & 1);
      tc_fence_after();
      float v[16];
      tmem_ld16_issue(tm + ((uint32_t)64 << 16), v);   // warp 2 reads TMEM lanes 64-95
      tmem_ld_wait();
      const long long t3 = clock64();
      if (v[0] == 123.456f) out[3] = 1;
      if (rep == 2 && threadIdx.x == 64) out[2] = t3;
    }
    __syncthreads();
    if (threadIdx.x == 64 && rep == 2) out[2] -= t_issue;
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x >= 32 && threadIdx.x < 64) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  printf("# tcgen05.mma M=128 K=16 kind::f16, cta_group::1, operands in shared memory (cycles, one CTA alone on the GPU)\n");
  fflush(stdout);
  printf("# N  n_mma  issue_loop  commit_seen_by_issuer  other_warp_has_first_tmem_load  ideal_tensor_cycles\n");
  for (int N : {64, 128, 256})
    for (int n : {1, 2, 4, 8, 16, 32}) {
      long long h[4] = {0, 0, 0, 0};
      k<<<1, 128, 60000>>>(n, N, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      printf("%4d %5d %10lld %10lld %10lld %10d\n", N, n, h[0], h[1], h[2], n * N / 2);
      fflush(stdout);
    }
  return 0;
}
