// Weighted sums along rays: out[r][c] = sum_s w[r][s] * v[r][s][c].
// The "feat_final = sum(weights * feat)" of inference (nnutils/rendering.py:233) for the 16-channel nerf_feat output,
// and any other per-ray expectation under the compositing weights.  One CTA of 128 threads per ray: thread s owns the
// samples s, s + 128, ..; C <= 32 channels live in registers, then a shuffle + shared-memory reduction.
#include "common.cuh"

namespace moda {

constexpr int WS_THREADS = 128;
constexpr int WS_MAXC = 32;

template <int C4>   // channels / 4 when C % 4 == 0 (128-bit loads), 0: scalar path
__global__ void __launch_bounds__(WS_THREADS) wsum_fwd_kernel(const float* __restrict__ w, const float* __restrict__ v,
                                                              float* __restrict__ out, int R, int S, int C) {
  const int r = blockIdx.x;
  float acc[WS_MAXC];
#pragma unroll
  for (int c = 0; c < WS_MAXC; ++c) acc[c] = 0.f;
  const float* wr = w + (size_t)r * S;
  const float* vr = v + (size_t)r * S * C;
  for (int s = threadIdx.x; s < S; s += WS_THREADS) {
    const float ws = wr[s];
    if (C4 > 0) {
      const float4* p = reinterpret_cast<const float4*>(vr + (size_t)s * C);
#pragma unroll
      for (int j = 0; j < C4; ++j) {
        const float4 f = __ldg(p + j);
        acc[4 * j] = fmaf(ws, f.x, acc[4 * j]); acc[4 * j + 1] = fmaf(ws, f.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(ws, f.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(ws, f.w, acc[4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < WS_MAXC; ++c)
        if (c < C) acc[c] = fmaf(ws, vr[(size_t)s * C + c], acc[c]);
    }
  }
  __shared__ float red[WS_THREADS / 32][WS_MAXC];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < WS_MAXC; ++c) {
    if (c < C) {
      const float t = warp_sum(acc[c]);
      if (lane == 0) red[wp][c] = t;
    }
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < WS_THREADS / 32; ++k) t += red[k][threadIdx.x];
    out[(size_t)r * C + threadIdx.x] = t;
  }
}

// gw[r][s] = sum_c gout[r][c] v[r][s][c]  (NULL: skipped);  gv[r][s][c] = w[r][s] gout[r][c]  (NULL: skipped)
__global__ void __launch_bounds__(WS_THREADS) wsum_bwd_kernel(const float* __restrict__ w, const float* __restrict__ v,
                                                              const float* __restrict__ gout, float* __restrict__ gw,
                                                              float* __restrict__ gv, int R, int S, int C) {
  const int r = blockIdx.x;
  __shared__ float g[WS_MAXC];
  if (threadIdx.x < C) g[threadIdx.x] = gout[(size_t)r * C + threadIdx.x];
  __syncthreads();
  const float* wr = w + (size_t)r * S;
  const float* vr = v + (size_t)r * S * C;
  for (int s = threadIdx.x; s < S; s += WS_THREADS) {
    const float ws = wr[s];
    float t = 0.f;
    for (int c = 0; c < C; ++c) {
      if (gw) t = fmaf(g[c], vr[(size_t)s * C + c], t);
      if (gv) gv[((size_t)r * S + s) * C + c] = ws * g[c];
    }
    if (gw) gw[(size_t)r * S + s] = t;
  }
}

}  // namespace moda

using namespace moda;

extern "C" int moda_wsum_fwd(const float* w, const float* v, float* out, int R, int S, int C, cudaStream_t stream) {
  if (R == 0) return 0;
  MODA_REQUIRE(w && v && out && S > 0 && C > 0 && C <= WS_MAXC, "wsum_fwd: bad arguments (C <= %d)", WS_MAXC);
  const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
  if (vec && C == 16) wsum_fwd_kernel<4><<<R, WS_THREADS, 0, stream>>>(w, v, out, R, S, C);
  else if (vec && C == 4) wsum_fwd_kernel<1><<<R, WS_THREADS, 0, stream>>>(w, v, out, R, S, C);
  else if (vec && C == 8) wsum_fwd_kernel<2><<<R, WS_THREADS, 0, stream>>>(w, v, out, R, S, C);
  else if (vec && C == 32) wsum_fwd_kernel<8><<<R, WS_THREADS, 0, stream>>>(w, v, out, R, S, C);
  else wsum_fwd_kernel<0><<<R, WS_THREADS, 0, stream>>>(w, v, out, R, S, C);
  return check_launch("wsum_fwd");
}

extern "C" int moda_wsum_bwd(const float* w, const float* v, const float* gout, float* gw, float* gv, int R, int S, int C,
                             cudaStream_t stream) {
  if (R == 0) return 0;
  MODA_REQUIRE(w && v && gout && S > 0 && C > 0 && C <= WS_MAXC, "wsum_bwd: bad arguments (C <= %d)", WS_MAXC);
  wsum_bwd_kernel<<<R, WS_THREADS, 0, stream>>>(w, v, gout, gw, gv, R, S, C);
  return check_launch("wsum_bwd");
}
