// Alpha compositing of rgb / depth / silhouette along each ray, with the cycle-distance pre-term.
//   inference()               nnutils/rendering.py:183-235   (density -> alpha -> transmittance -> sums)
//   frame_cyc_dis             nnutils/rendering.py:341, 473   (sum_s |x - x_cyc| * w.detach())
// One warp per ray; the ray is walked in chunks of 32 samples (lane = sample), the exclusive prefix
// product of (1 - alpha + 1e-10) is a 5-step shuffle scan per chunk with a running carry, and the backward
// pass walks the chunks in reverse with a suffix-sum carry.
#include "common.cuh"

namespace moda {

struct CompArgs {
  const float* rgb; int ld_rgb;      // (P,3) with row stride
  const float* sigma; int ld_sigma;  // (P) with stride
  const float* z;                    // (R,S)
  const float* d;                    // (R,3) un-normalised ray directions
  const float* beta;                 // device scalar
  const float* noise;                // (R,S) or null, already scaled by noise_std
  const unsigned char* mask;         // (R,S) or null: 1 = force alpha to 0 (rendering.py:210-215)
  const float* xa;                   // (R,S,3) frame-space samples, or null
  const float* xb;                   // (R,S,3) cycled samples, or null
  int R, S;
  // forward outputs
  float* out_rgb;    // (R,3)
  float* out_depth;  // (R)
  float* out_sil;    // (R)
  float* out_w;      // (R,S)
  float* out_vis;    // (R,S) transmittance
  float* out_cyc;    // (R) or null
  // backward
  const float* vis;                    // saved transmittance
  const float* g_rgb; const float* g_depth; const float* g_sil; const float* g_cyc;  // (R,*)
  const float* g_w;                    // (R,S) external gradient on the weights, or null
  float* g_rgbs; int ld_grgb;          // (P,3)
  float* g_sigma; int ld_gsigma;       // (P)
  float* g_beta;                       // scalar, accumulated
  float* g_nd;                         // (R) gradient w.r.t. |d| (consumed by sample_rays_bwd)
  float* g_xa; float* g_xb;            // (R,S,3) or null
};

constexpr int COMP_WARPS = 4;

__global__ void __launch_bounds__(COMP_WARPS * 32) composite_fwd_kernel(CompArgs a) {
  const int r = blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= a.R) return;
  const int S = a.S;
  const float dx = a.d[r * 3], dy = a.d[r * 3 + 1], dz = a.d[r * 3 + 2];
  const float nd = sqrtf(dx * dx + dy * dy + dz * dz);
  const float ibeta = 1.0f / (fabsf(a.beta[0]) + 1e-9f);
  float carry = 1.0f;
  float s_r = 0, s_g = 0, s_b = 0, s_d = 0, s_s = 0, s_c = 0;
  for (int c0 = 0; c0 < S; c0 += 32) {
    const int i = c0 + lane;
    const bool live = i < S;
    const size_t t = (size_t)r * S + (live ? i : 0);
    float q = 1.0f, alpha = 0.f, zi = 0.f;
    if (live) {
      zi = a.z[t];
      const float delta = (i < S - 1) ? (a.z[t + 1] - zi) * nd : 1e10f * nd;
      float sg = a.sigma[t * a.ld_sigma];
      if (a.noise) sg += a.noise[t];
      alpha = density_alpha(sg, delta, ibeta, nullptr, nullptr, nullptr);
      if (a.mask && a.mask[t]) alpha = 0.f;
      q = 1.0f - alpha + 1e-10f;
    }
    // inclusive product scan over the chunk
    float inc = q;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc *= up;
    }
    float exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 1.0f;
    const float T = carry * exc;
    carry *= __shfl_sync(0xffffffffu, inc, 31);
    if (live) {
      const float w = alpha * T;
      a.out_w[t] = w;
      a.out_vis[t] = T;
      const float* c = a.rgb + t * a.ld_rgb;
      s_r += w * c[0]; s_g += w * c[1]; s_b += w * c[2];
      s_d += w * zi;
      if (i < S - 1) s_s += w;
      if (a.out_cyc) {
        const float e0 = a.xa[t * 3] - a.xb[t * 3], e1 = a.xa[t * 3 + 1] - a.xb[t * 3 + 1],
                    e2 = a.xa[t * 3 + 2] - a.xb[t * 3 + 2];
        s_c += w * sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
      }
    }
  }
  s_r = warp_sum(s_r); s_g = warp_sum(s_g); s_b = warp_sum(s_b);
  s_d = warp_sum(s_d); s_s = warp_sum(s_s); s_c = warp_sum(s_c);
  if (lane == 0) {
    a.out_rgb[r * 3] = s_r; a.out_rgb[r * 3 + 1] = s_g; a.out_rgb[r * 3 + 2] = s_b;
    a.out_depth[r] = s_d;
    a.out_sil[r] = s_s;
    if (a.out_cyc) a.out_cyc[r] = s_c;
  }
}

__global__ void __launch_bounds__(COMP_WARPS * 32) composite_bwd_kernel(CompArgs a) {
  const int r = blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float gib = 0.f;  // d/d ibeta, reduced at the end (warps beyond R contribute 0)
  if (r < a.R) {
    const int S = a.S;
    const float dx = a.d[r * 3], dy = a.d[r * 3 + 1], dz = a.d[r * 3 + 2];
    const float nd = sqrtf(dx * dx + dy * dy + dz * dz);
    const float ibeta = 1.0f / (fabsf(a.beta[0]) + 1e-9f);
    const float gr = a.g_rgb ? a.g_rgb[r * 3] : 0.f, gg = a.g_rgb ? a.g_rgb[r * 3 + 1] : 0.f,
                gb = a.g_rgb ? a.g_rgb[r * 3 + 2] : 0.f;
    const float gdep = a.g_depth ? a.g_depth[r] : 0.f;
    const float gsil = a.g_sil ? a.g_sil[r] : 0.f;
    const float gcyc = a.g_cyc ? a.g_cyc[r] : 0.f;
    float suffix = 0.f;  // sum_{i > k} G_i w_i over the chunks already visited
    float gnd = 0.f;
    const int nchunk = (S + 31) / 32;
    for (int ch = nchunk - 1; ch >= 0; --ch) {
      const int i = ch * 32 + lane;
      const bool live = i < S;
      const size_t t = (size_t)r * S + (live ? i : 0);
      float Gw = 0.f, T = 0.f, alpha = 0.f, q = 1.f, G = 0.f;
      float da_ds = 0.f, da_dib = 0.f, da_dd = 0.f, dz_ = 0.f;
      bool masked = false;
      if (live) {
        const float zi = a.z[t];
        dz_ = (i < S - 1) ? (a.z[t + 1] - zi) : 1e10f;
        const float delta = dz_ * nd;
        float sg = a.sigma[t * a.ld_sigma];
        if (a.noise) sg += a.noise[t];
        alpha = density_alpha(sg, delta, ibeta, &da_ds, &da_dib, &da_dd);
        masked = a.mask && a.mask[t];
        if (masked) alpha = 0.f;
        q = 1.0f - alpha + 1e-10f;
        T = a.vis[t];
        const float w = alpha * T;
        const float* c = a.rgb + t * a.ld_rgb;
        G = gr * c[0] + gg * c[1] + gb * c[2] + gdep * zi + ((i < S - 1) ? gsil : 0.f);
        if (a.g_w) G += a.g_w[t];
        Gw = G * w;
        if (a.g_rgbs) {
          float* o = a.g_rgbs + t * a.ld_grgb;
          o[0] = w * gr; o[1] = w * gg; o[2] = w * gb;
        }
        if (a.g_xa || a.g_xb) {
          // frame_cyc_dis uses w.detach(): only the distance gets a gradient
          const float e0 = a.xa[t * 3] - a.xb[t * 3], e1 = a.xa[t * 3 + 1] - a.xb[t * 3 + 1],
                      e2 = a.xa[t * 3 + 2] - a.xb[t * 3 + 2];
          const float n = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
          const float k = (n > 0.f) ? gcyc * w / n : 0.f;
          if (a.g_xa) { a.g_xa[t * 3] = k * e0; a.g_xa[t * 3 + 1] = k * e1; a.g_xa[t * 3 + 2] = k * e2; }
          if (a.g_xb) { a.g_xb[t * 3] = -k * e0; a.g_xb[t * 3 + 1] = -k * e1; a.g_xb[t * 3 + 2] = -k * e2; }
        }
      }
      // suffix sum within the chunk: sfx_k = sum_{lane' > lane} Gw
      float inc = Gw;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float dn = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc += dn;
      }
      const float after = inc - Gw + suffix;
      suffix += __shfl_sync(0xffffffffu, inc, 0);
      if (live) {
        float ga = masked ? 0.f : (G * T - after / q);
        if (a.g_sigma) a.g_sigma[t * a.ld_gsigma] = ga * da_ds;
        gib += ga * da_dib;
        gnd += ga * da_dd * dz_;
      }
    }
    gnd = warp_sum(gnd);
    if (lane == 0 && a.g_nd) a.g_nd[r] = gnd;
  }
  if (a.g_beta) {
    __shared__ float red[COMP_WARPS];
    const float t = warp_sum(gib);
    if (lane == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < COMP_WARPS; ++i) tot += red[i];
      if (tot != 0.f) {
        const float bt = a.beta[0];
        const float ib = 1.0f / (fabsf(bt) + 1e-9f);
        const float sgn = (bt > 0.f) ? 1.f : ((bt < 0.f) ? -1.f : 0.f);
        atomicAdd(a.g_beta, -sgn * tot * ib * ib);
      }
    }
  }
}

}  // namespace moda

using namespace moda;

extern "C" int moda_composite_fwd(const float* rgb, int ld_rgb, const float* sigma, int ld_sigma,
                                  const float* z, const float* d, const float* beta, const float* noise,
                                  const unsigned char* mask, const float* xa, const float* xb, float* out_rgb,
                                  float* out_depth, float* out_sil, float* out_w, float* out_vis,
                                  float* out_cyc, int R, int S, cudaStream_t stream) {
  MODA_REQUIRE(rgb && sigma && z && d && beta && out_rgb && out_depth && out_sil && out_w && out_vis,
               "composite_fwd: null pointer");
  MODA_REQUIRE(!out_cyc || (xa && xb), "composite_fwd: cycle term needs both point sets");
  if (R == 0) return 0;
  CompArgs a = {};
  a.rgb = rgb; a.ld_rgb = ld_rgb; a.sigma = sigma; a.ld_sigma = ld_sigma; a.z = z; a.d = d; a.beta = beta;
  a.noise = noise; a.mask = mask; a.xa = xa; a.xb = xb; a.R = R; a.S = S;
  a.out_rgb = out_rgb; a.out_depth = out_depth; a.out_sil = out_sil; a.out_w = out_w; a.out_vis = out_vis;
  a.out_cyc = out_cyc;
  composite_fwd_kernel<<<cdiv(R, COMP_WARPS), COMP_WARPS * 32, 0, stream>>>(a);
  return check_launch("composite_fwd");
}

extern "C" int moda_composite_bwd(const float* rgb, int ld_rgb, const float* sigma, int ld_sigma,
                                  const float* z, const float* d, const float* beta, const float* noise,
                                  const unsigned char* mask, const float* xa, const float* xb,
                                  const float* vis, const float* g_rgb, const float* g_depth,
                                  const float* g_sil, const float* g_cyc, const float* g_w, float* g_rgbs,
                                  int ld_grgb, float* g_sigma, int ld_gsigma, float* g_beta, float* g_nd,
                                  float* g_xa, float* g_xb, int R, int S, cudaStream_t stream) {
  MODA_REQUIRE(rgb && sigma && z && d && beta && vis, "composite_bwd: null pointer");
  MODA_REQUIRE(!(g_xa || g_xb) || (xa && xb), "composite_bwd: cycle term needs both point sets");
  if (R == 0) return 0;
  CompArgs a = {};
  a.rgb = rgb; a.ld_rgb = ld_rgb; a.sigma = sigma; a.ld_sigma = ld_sigma; a.z = z; a.d = d; a.beta = beta;
  a.noise = noise; a.mask = mask; a.xa = xa; a.xb = xb; a.R = R; a.S = S; a.vis = vis;
  a.g_rgb = g_rgb; a.g_depth = g_depth; a.g_sil = g_sil; a.g_cyc = g_cyc; a.g_w = g_w;
  a.g_rgbs = g_rgbs; a.ld_grgb = ld_grgb; a.g_sigma = g_sigma; a.ld_gsigma = ld_gsigma;
  a.g_beta = g_beta; a.g_nd = g_nd; a.g_xa = g_xa; a.g_xb = g_xb;
  composite_bwd_kernel<<<cdiv(R, COMP_WARPS), COMP_WARPS * 32, 0, stream>>>(a);
  return check_launch("composite_bwd");
}
