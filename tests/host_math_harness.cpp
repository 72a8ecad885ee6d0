// TEST INFRASTRUCTURE ONLY.  Serial CPU driver around moda_b200/csrc/moda_math.h so that the hand-derived
// adjoints used by the CUDA kernels can be checked against the oracle's autograd without a GPU.
// Built by tests/test_host_math.py with g++; never loaded by the product.
#include <math.h>
#include <string.h>

#include <vector>

#include "moda_math.h"

using namespace moda;

struct HostEmit {
  float* acc;
  void operator()(int b, const float* v, bool any) {
    if (!any) return;
    for (int i = 0; i < ACC_STRIDE; ++i) acc[b * ACC_STRIDE + i] += v[i];
  }
};

extern "C" {

// mirrors skin_warp_fwd_kernel
void h_skin_warp_fwd(const float* pts, const float* bones, const float* rts, const float* skin_aux,
                     const float* dskin, const float* skin_in, float* y, float* skin_out, int R, int S, int B,
                     int bones_per_ray, int deform, int invert) {
  std::vector<float> ctx(B * CTX_STRIDE), used(B * 10);
  const float kappa = 1000.0f * expf(skin_aux[0]);
  for (int r = 0; r < R; ++r) {
    for (int b = 0; b < B; ++b)
      ray_bone_setup(bones + ((size_t)(bones_per_ray ? r : 0) * B + b) * 10,
                     rts ? rts + ((size_t)r * B + b) * 8 : nullptr, deform, invert, kappa, &used[b * 10],
                     &ctx[b * CTX_STRIDE]);
    for (int s = 0; s < S; ++s) {
      const size_t pi = (size_t)r * S + s;
      const float px = pts[pi * 3], py = pts[pi * 3 + 1], pz = pts[pi * 3 + 2];
      const float* dl = dskin ? dskin + pi * B : nullptr;
      const float* win = skin_in ? skin_in + pi * B : nullptr;
      float bl[8], mx, sum;
      skin_point_blend(ctx.data(), B, px, py, pz, dl, win, bl, &mx, &sum);
      if (skin_out && !win)
        for (int b = 0; b < B; ++b) {
          float l = bone_logit(&ctx[b * CTX_STRIDE], px, py, pz);
          if (dl) l += dl[b];
          skin_out[pi * B + b] = expf(l - mx) / sum;
        }
      if (y) {
        const float inv_n = 1.0f / sqrtf(bl[0] * bl[0] + bl[1] * bl[1] + bl[2] * bl[2] + bl[3] * bl[3]);
        float c[8];
        for (int i = 0; i < 8; ++i) c[i] = bl[i] * inv_n;
        dq_apply(c, px, py, pz, y + pi * 3);
      }
    }
  }
}

// mirrors skin_warp_bwd_kernel (gbones, grts, gaux accumulate into zero-initialised buffers)
void h_skin_warp_bwd(const float* pts, const float* bones, const float* rts, const float* skin_aux,
                     const float* dskin, const float* skin_in, const float* gy, const float* gskin, float* gpts,
                     float* gdskin, float* gskin_in, float* grts, float* gbones, float* gaux, int R, int S,
                     int B, int bones_per_ray, int deform, int invert) {
  std::vector<float> ctx(B * CTX_STRIDE), used(B * 10), acc(B * ACC_STRIDE);
  const float kappa = 1000.0f * expf(skin_aux[0]);
  for (int r = 0; r < R; ++r) {
    for (int b = 0; b < B; ++b)
      ray_bone_setup(bones + ((size_t)(bones_per_ray ? r : 0) * B + b) * 10,
                     rts ? rts + ((size_t)r * B + b) * 8 : nullptr, deform, invert, kappa, &used[b * 10],
                     &ctx[b * CTX_STRIDE]);
    std::fill(acc.begin(), acc.end(), 0.f);
    HostEmit emit{acc.data()};
    for (int s = 0; s < S; ++s) {
      const size_t pi = (size_t)r * S + s;
      float gp[3], gaux_pt = 0.f;
      skin_point_bwd(ctx.data(), B, pts[pi * 3], pts[pi * 3 + 1], pts[pi * 3 + 2],
                     dskin ? dskin + pi * B : nullptr, skin_in ? skin_in + pi * B : nullptr,
                     (gy && rts) ? gy + pi * 3 : nullptr, gskin ? gskin + pi * B : nullptr, true, gp,
                     gdskin ? gdskin + pi * B : nullptr, gskin_in ? gskin_in + pi * B : nullptr, &gaux_pt, emit);
      if (gaux) gaux[0] += gaux_pt;
      if (gpts) { gpts[pi * 3] = gp[0]; gpts[pi * 3 + 1] = gp[1]; gpts[pi * 3 + 2] = gp[2]; }
    }
    for (int b = 0; b < B; ++b) {
      float gbone[10], grt[8], ga = 0.f;
      ray_bone_setup_bwd(bones + ((size_t)(bones_per_ray ? r : 0) * B + b) * 10,
                         rts ? rts + ((size_t)r * B + b) * 8 : nullptr, deform, invert, kappa, &used[b * 10],
                         &acc[b * ACC_STRIDE], gbone, grt, &ga);
      if (grts && rts) for (int i = 0; i < 8; ++i) grts[((size_t)r * B + b) * 8 + i] += grt[i];
      if (gbones) for (int i = 0; i < 10; ++i) gbones[((size_t)(bones_per_ray ? r : 0) * B + b) * 10 + i] += gbone[i];
      if (gaux) gaux[0] += ga;
    }
  }
}

void h_bone_transform_fwd(const float* bones, const float* rts, float* out, int R, int B) {
  for (int i = 0; i < R * B; ++i) bone_transform_fwd(bones + (i % B) * 10, rts + (size_t)i * 8, out + (size_t)i * 10);
}

void h_bone_transform_bwd(const float* bones, const float* rts, const float* gout, float* gbones, float* grts,
                          int R, int B) {
  for (int i = 0; i < R * B; ++i)
    bone_transform_bwd(bones + (i % B) * 10, rts + (size_t)i * 8, gout + (size_t)i * 10, gbones + (i % B) * 10,
                       grts + (size_t)i * 8);
}

// serial restatement of the compositor adjoint with the same per-sample function the kernel uses
void h_density_alpha(const float* sigma, const float* delta, float ibeta, float* alpha, float* da_ds,
                     float* da_dib, float* da_dd, int n) {
  for (int i = 0; i < n; ++i) alpha[i] = density_alpha(sigma[i], delta[i], ibeta, da_ds + i, da_dib + i, da_dd + i);
}

}  // extern "C"
