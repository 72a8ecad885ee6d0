// Flow rendering of the third warp: the samples of a ray, warped into a paired frame (root coordinates), are projected
// with that frame's camera and their expected 2-D displacement under the compositing weights is the rendered flow
//   rendering.py:434-459, 480-499 with obj_to_cam / pinhole_cam (geom_utils.py:567-581, 654-672) and vrender_flo
//   (geom_utils.py:1704-1743):
//     Xc = R x + T;   (u, v, z) = (fx Xc.x + px Xc.z, fy Xc.y + py Xc.z, Xc.z);   xy = (u, v) / (1e-6 + z)
//     invalid = z < 1e-5  or  |xy| > 2 img_size;    w' = invalid ? 0 : w;   xy' = invalid ? 0 : xy
//     flo = sum_s w'_s / (1e-9 + sum w') (xy'_s - xys) * 2 / img_size;     valid = no sample invalid
// As tensor ops this is ~25 elementwise / batched-matmul launches over (rays, samples, 3) forward and twice that backward,
// per paired frame; here one warp per ray does it in one launch each way (the ray's camera in registers, samples strided
// over the lanes, sums by shuffles).  The backward kernel recomputes the projection and returns the gradients of the
// points, the weights and the ray's camera (R, T, K: the root pose and intrinsics are optimised by the caller).
#include "common.cuh"

namespace moda {

struct FlowCam {
  float R[9], T[3], fx, fy, px, py;
};

__device__ __forceinline__ FlowCam flow_cam(const float* __restrict__ R, const float* __restrict__ T,
                                            const float* __restrict__ K, int r) {
  FlowCam c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.R[i] = __ldg(R + (size_t)r * 9 + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) c.T[i] = __ldg(T + (size_t)r * 3 + i);
  c.fx = __ldg(K + (size_t)r * 4); c.fy = __ldg(K + (size_t)r * 4 + 1);
  c.px = __ldg(K + (size_t)r * 4 + 2); c.py = __ldg(K + (size_t)r * 4 + 3);
  return c;
}

struct FlowPt {
  float X, Y, Z, u, v, den, x, y;   // camera-frame point, homogeneous pixel, 1e-6 + z, projected pixel
  bool invalid;
};

__device__ __forceinline__ FlowPt flow_project(const FlowCam& c, const float* __restrict__ p, float lim) {
  FlowPt q;
  const float a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
  q.X = fmaf(c.R[0], a, fmaf(c.R[1], b, c.R[2] * d)) + c.T[0];
  q.Y = fmaf(c.R[3], a, fmaf(c.R[4], b, c.R[5] * d)) + c.T[1];
  q.Z = fmaf(c.R[6], a, fmaf(c.R[7], b, c.R[8] * d)) + c.T[2];
  q.u = fmaf(c.fx, q.X, c.px * q.Z);
  q.v = fmaf(c.fy, q.Y, c.py * q.Z);
  q.den = 1e-6f + q.Z;
  q.x = q.u / q.den;
  q.y = q.v / q.den;
  q.invalid = (q.Z < 1e-5f) || (sqrtf(q.x * q.x + q.y * q.y) > lim);
  return q;
}

__global__ void __launch_bounds__(256) flow_render_fwd_kernel(const float* __restrict__ xyz, const float* __restrict__ R,
                                                              const float* __restrict__ T, const float* __restrict__ K,
                                                              const float* __restrict__ w, const float* __restrict__ xys,
                                                              int N, int S, float img_size, float* __restrict__ flo,
                                                              float* __restrict__ valid) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= N) return;
  const FlowCam c = flow_cam(R, T, K, r);
  const float x0 = __ldg(xys + 2 * (size_t)r), y0 = __ldg(xys + 2 * (size_t)r + 1);
  const float lim = 2.0f * img_size;
  float W = 0.f, fx = 0.f, fy = 0.f, bad = 0.f;
  for (int s = lane; s < S; s += 32) {
    const FlowPt q = flow_project(c, xyz + ((size_t)r * S + s) * 3, lim);
    const float ws = q.invalid ? 0.f : __ldg(w + (size_t)r * S + s);
    W += ws;
    fx = fmaf(ws, (q.invalid ? 0.f : q.x) - x0, fx);
    fy = fmaf(ws, (q.invalid ? 0.f : q.y) - y0, fy);
    bad += q.invalid ? 1.f : 0.f;
  }
  W = warp_sum(W); fx = warp_sum(fx); fy = warp_sum(fy); bad = warp_sum(bad);
  if (lane == 0) {
    const float k = 2.0f / (img_size * (1e-9f + W));
    flo[2 * (size_t)r] = fx * k;
    flo[2 * (size_t)r + 1] = fy * k;
    valid[r] = bad == 0.f ? 1.f : 0.f;
  }
}

__global__ void __launch_bounds__(256) flow_render_bwd_kernel(const float* __restrict__ xyz, const float* __restrict__ R,
                                                              const float* __restrict__ T, const float* __restrict__ K,
                                                              const float* __restrict__ w, const float* __restrict__ xys,
                                                              const float* __restrict__ gflo, int N, int S, float img_size,
                                                              float* __restrict__ gxyz, float* __restrict__ gR,
                                                              float* __restrict__ gT, float* __restrict__ gK,
                                                              float* __restrict__ gw) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= N) return;
  const FlowCam c = flow_cam(R, T, K, r);
  const float x0 = __ldg(xys + 2 * (size_t)r), y0 = __ldg(xys + 2 * (size_t)r + 1);
  const float lim = 2.0f * img_size;
  const float gx = __ldg(gflo + 2 * (size_t)r) * 2.0f / img_size, gy = __ldg(gflo + 2 * (size_t)r + 1) * 2.0f / img_size;
  // pass 1: W = sum w' and A = sum_t g_wn_t w'_t with g_wn_t = g . (xy'_t - xys)   (wn = w' / (1e-9 + W))
  float W = 0.f, A = 0.f;
  for (int s = lane; s < S; s += 32) {
    const FlowPt q = flow_project(c, xyz + ((size_t)r * S + s) * 3, lim);
    if (!q.invalid) {
      const float ws = __ldg(w + (size_t)r * S + s);
      W += ws;
      A = fmaf(ws, gx * (q.x - x0) + gy * (q.y - y0), A);
    }
  }
  W = warp_sum(W); A = warp_sum(A);
  const float inv = 1.0f / (1e-9f + W);
  A *= inv;     // sum_t g_wn_t wn_t
  // (an invalid sample contributes xy' = 0, w' = 0: no gradient to its point or weight)
  float aR[9], aT[3], aK[4];
#pragma unroll
  for (int i = 0; i < 9; ++i) aR[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) aT[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) aK[i] = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float* p = xyz + ((size_t)r * S + s) * 3;
    const FlowPt q = flow_project(c, p, lim);
    float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f, gws = 0.f;
    if (!q.invalid) {
      const float ws = __ldg(w + (size_t)r * S + s);
      gws = (gx * (q.x - x0) + gy * (q.y - y0) - A) * inv;
      const float wn = ws * inv;
      // xy = (u, v) / den
      const float gu = wn * gx / q.den, gv = wn * gy / q.den;
      const float gz = -(gu * q.u + gv * q.v) / q.den;
      const float gX = c.fx * gu, gY = c.fy * gv, gZ = fmaf(c.px, gu, fmaf(c.py, gv, gz));
      aK[0] = fmaf(gu, q.X, aK[0]); aK[1] = fmaf(gv, q.Y, aK[1]); aK[2] = fmaf(gu, q.Z, aK[2]); aK[3] = fmaf(gv, q.Z, aK[3]);
      const float a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
      aR[0] = fmaf(gX, a, aR[0]); aR[1] = fmaf(gX, b, aR[1]); aR[2] = fmaf(gX, d, aR[2]);
      aR[3] = fmaf(gY, a, aR[3]); aR[4] = fmaf(gY, b, aR[4]); aR[5] = fmaf(gY, d, aR[5]);
      aR[6] = fmaf(gZ, a, aR[6]); aR[7] = fmaf(gZ, b, aR[7]); aR[8] = fmaf(gZ, d, aR[8]);
      aT[0] += gX; aT[1] += gY; aT[2] += gZ;
      gp0 = fmaf(c.R[0], gX, fmaf(c.R[3], gY, c.R[6] * gZ));
      gp1 = fmaf(c.R[1], gX, fmaf(c.R[4], gY, c.R[7] * gZ));
      gp2 = fmaf(c.R[2], gX, fmaf(c.R[5], gY, c.R[8] * gZ));
    }
    float* o = gxyz + ((size_t)r * S + s) * 3;
    o[0] = gp0; o[1] = gp1; o[2] = gp2;
    gw[(size_t)r * S + s] = gws;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) aR[i] = warp_sum(aR[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i) aT[i] = warp_sum(aT[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) aK[i] = warp_sum(aK[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) gR[(size_t)r * 9 + i] = aR[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) gT[(size_t)r * 3 + i] = aT[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) gK[(size_t)r * 4 + i] = aK[i];
  }
}

}  // namespace moda

// xyz (N, S, 3) root-frame points, R (N, 9) row-major rotation, T (N, 3), K (N, 4) = (fx, fy, px, py), w (N, S) compositing
// weights, xys (N, 2) pixel of the ray; all fp32 contiguous.  Out: flo (N, 2), valid (N).
extern "C" int moda_flow_render_fwd(const float* xyz, const float* R, const float* T, const float* K, const float* w,
                                    const float* xys, int N, int S, float img_size, float* flo, float* valid,
                                    cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(xyz && R && T && K && w && xys && flo && valid && N >= 0 && S > 0 && img_size > 0.f, "flow_render_fwd: bad arguments");
  if (N == 0) return 0;
  flow_render_fwd_kernel<<<cdiv(N, 8), 256, 0, stream>>>(xyz, R, T, K, w, xys, N, S, img_size, flo, valid);
  return check_launch("flow_render_fwd");
}

// Gradients of the above for gflo (N, 2): gxyz (N, S, 3), gR (N, 9), gT (N, 3), gK (N, 4), gw (N, S), all overwritten.
extern "C" int moda_flow_render_bwd(const float* xyz, const float* R, const float* T, const float* K, const float* w,
                                    const float* xys, const float* gflo, int N, int S, float img_size, float* gxyz, float* gR,
                                    float* gT, float* gK, float* gw, cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(xyz && R && T && K && w && xys && gflo && gxyz && gR && gT && gK && gw && N >= 0 && S > 0 && img_size > 0.f,
               "flow_render_bwd: bad arguments");
  if (N == 0) return 0;
  flow_render_bwd_kernel<<<cdiv(N, 8), 256, 0, stream>>>(xyz, R, T, K, w, xys, gflo, N, S, img_size, gxyz, gR, gT, gK, gw);
  return check_launch("flow_render_bwd");
}
