// One pass over the (n x m) kernel matrix of the Sinkhorn feature matching (nnutils/loss_utils.py:347-386, 20 iterations
// of c = K^T a, b = p2 / (c + d), d = K b, a = p1 / (d + d) on K = exp((f.v - 1) / 0.03), n = rays, m = 8000 lattice
// points) that serves BOTH products of an iteration:
//     y[r] = sum_j K[r][j] x[j]                       (row sums:    d = K b,   or ga = K gc in the adjoint sweep)
//     z[r] = g(y[r], r)                               (elementwise: a = p1 / (d + delta),   gd = -ga a / (d + delta))
//     w[j] += sum_r K[r][j] z[r]                      (column sums: c = K^T a,   gb = K^T gd)
// A row's z is known as soon as its row sum is, so the column product can use the row while it is still on chip: K (262 MB
// at 8192 rays) is streamed ONCE per iteration instead of twice -- the two matrix-vector products of an iteration were
// each at the HBM rate already (42 us), so halving the bytes is what is left.
//
// Persistent CTAs (two per SM); per trip a CTA loads RB = 2 rows into REGISTERS (coalesced 128-bit loads, the dot products
// with x taken on the fly), reduces the row sums, forms z, and then adds z[r] K[r][:] into per-thread column accumulators
// that live in registers for the whole launch (m <= 8192: at most 8 float4 per thread at 256 threads); they are flushed
// with one atomic per column and CTA at the end.  x is staged once in shared memory.
#include "common.cuh"

namespace moda {

constexpr int SK_THREADS = 256;
constexpr int SK_RB = 2;          // rows per trip: 16 float4 per thread in registers next to the 8 column accumulators; two
                                  // CTAs per SM overlap each other's load / reduce / column phases
constexpr int SK_MAXV = 8;        // float4 per thread and row: m <= 8 * 4 * 256 = 8192

// mode 0: z = p / (y + delta)            (forward iteration; p = 1 / n)
// mode 1: z = -y * u[r] / (v[r] + delta) (adjoint sweep; u = a_{i-1}, v = d_{i-2})
__global__ void __launch_bounds__(SK_THREADS, 2)
sinkhorn_pass_kernel(const float* __restrict__ K, int n, int m, const float* __restrict__ x, float* __restrict__ y,
                     float* __restrict__ z, float* __restrict__ w, int mode, float p, float delta,
                     const float* __restrict__ u, const float* __restrict__ v) {
  extern __shared__ __align__(16) float sk_smem[];
  const int m4 = m >> 2;                         // m % 4 == 0 (checked by the host)
  float4* xs = reinterpret_cast<float4*>(sk_smem);                    // [m4]
  __shared__ float red[SK_RB][SK_THREADS / 32];
  __shared__ float zrow[SK_RB];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  for (int i = tid; i < m4; i += SK_THREADS) xs[i] = __ldg(reinterpret_cast<const float4*>(x) + i);
  float4 acc[SK_MAXV];
#pragma unroll
  for (int k = 0; k < SK_MAXV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int nblk = (n + SK_RB - 1) / SK_RB;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int r0 = blk * SK_RB;
    float dot[SK_RB];
    float4 kreg[SK_RB][SK_MAXV];   // this thread's entries of the rows stay in registers for the column phase
#pragma unroll
    for (int rr = 0; rr < SK_RB; ++rr) dot[rr] = 0.f;
    // all loads of the trip are independent: SK_RB x SK_MAXV 128-bit loads per thread in flight
#pragma unroll
    for (int rr = 0; rr < SK_RB; ++rr) {
      const int r = r0 + rr;
      const float4* src = reinterpret_cast<const float4*>(K + (size_t)(r < n ? r : n - 1) * m);
#pragma unroll
      for (int k = 0; k < SK_MAXV; ++k) {
        const int i = tid + k * SK_THREADS;
        if (i < m4) {
          const float4 kv = __ldcs(src + i);     // streamed once: do not keep it in L1 / push x out of L2
          const float4 xv = xs[i];
          kreg[rr][k] = kv;
          dot[rr] = fmaf(kv.x, xv.x, fmaf(kv.y, xv.y, fmaf(kv.z, xv.z, fmaf(kv.w, xv.w, dot[rr]))));
        } else {
          kreg[rr][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < SK_RB; ++rr) {
      float s = dot[rr];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) red[rr][wp] = s;
    }
    __syncthreads();
    if (tid < SK_RB) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < SK_THREADS / 32; ++k) s += red[tid][k];
      const int r = r0 + tid;
      float zz = 0.f;
      if (r < n) {
        zz = (mode == 0) ? p / (s + delta) : -s * u[r] / (v[r] + delta);
        if (y) y[r] = s;
        if (z) z[r] = zz;
      }
      zrow[tid] = zz;
    }
    __syncthreads();
    if (w) {
      float zr[SK_RB];
#pragma unroll
      for (int rr = 0; rr < SK_RB; ++rr) zr[rr] = zrow[rr];
#pragma unroll
      for (int k = 0; k < SK_MAXV; ++k) {
        const int i = tid + k * SK_THREADS;
        if (i < m4) {
#pragma unroll
          for (int rr = 0; rr < SK_RB; ++rr) {
            const float4 kv = kreg[rr][k];
            acc[k].x = fmaf(zr[rr], kv.x, acc[k].x); acc[k].y = fmaf(zr[rr], kv.y, acc[k].y);
            acc[k].z = fmaf(zr[rr], kv.z, acc[k].z); acc[k].w = fmaf(zr[rr], kv.w, acc[k].w);
          }
        }
      }
    }
    __syncthreads();   // zrow / red are rewritten by the next trip
  }
  if (w) {
#pragma unroll
    for (int k = 0; k < SK_MAXV; ++k) {
      const int i = tid + k * SK_THREADS;
      if (i < m4) {
        float* o = w + 4 * (size_t)i;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(acc[k].x), "f"(acc[k].y), "f"(acc[k].z),
                     "f"(acc[k].w)
                     : "memory");
      }
    }
  }
}

}  // namespace moda

// K (n, m) row-major fp32, m % 4 == 0, m <= 8192, 16-byte aligned rows; x (m); y, z (n) may be NULL; w (m) must be zeroed by
// the caller (it is accumulated with atomics), NULL = row sums only.  mode 0: z = p / (y + delta); mode 1: z = -y u / (v +
// delta) with u, v (n).  Replaces the pair torch.mv(K, x) / torch.mv(K.t(), z) of one Sinkhorn iteration.
extern "C" int moda_sinkhorn_pass(const float* K, int n, int m, const float* x, float* y, float* z, float* w, int mode,
                                  float p, float delta, const float* u, const float* v, cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(K && x && n >= 0 && m > 0 && m % 4 == 0 && m <= SK_MAXV * 4 * SK_THREADS,
               "sinkhorn_pass: bad arguments (m %% 4 == 0, m <= %d)", SK_MAXV * 4 * SK_THREADS);
  MODA_REQUIRE(((reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0,
               "sinkhorn_pass: K, x and w must be 16-byte aligned");
  MODA_REQUIRE(mode == 0 || (mode == 1 && u && v), "sinkhorn_pass: mode 1 needs u and v");
  if (n == 0) return 0;
  const size_t smem = (size_t)m * sizeof(float);   // x only: the rows of a trip live in registers
  cudaFuncSetAttribute(sinkhorn_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int nblk = (n + SK_RB - 1) / SK_RB;
  const int grid = nblk < 2 * sms ? nblk : 2 * sms;
  sinkhorn_pass_kernel<<<grid, SK_THREADS, smem, stream>>>(K, n, m, x, y, z, w, mode, p, delta, u, v);
  return check_launch("sinkhorn_pass");
}
