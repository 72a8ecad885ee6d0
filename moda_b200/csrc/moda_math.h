// Per-element math of the articulated warp and the compositor, shared by the CUDA kernels and by the
// CPU adjoint-check harness in tests/ (compiled with g++; test infrastructure only).
//
// Everything here is a restatement of reference formulas (paths relative to /root/reference):
//   quaternion -> matrix   third_party/pytorch3d/pytorch3d/transforms/rotation_conversions.py:41-69
//   Hamilton product       rotation_conversions.py:374-392, nnutils/dual_quat.py:14-31
//   bone_transform         nnutils/geom_utils.py:73-86 (neudbs branch)
//   vec_to_sim3            nnutils/geom_utils.py:187-199
//   Gaussian skin logits   nnutils/geom_utils.py:251-276
//   DQ blend + transform   nnutils/geom_utils.py:470-491, nnutils/dual_quat.py:51-62, 87-93
// and of their hand-derived adjoints.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MODA_HD __host__ __device__ __forceinline__
#else
#define MODA_HD inline
#endif

namespace moda {

// floats per bone in the per-ray context: 3 rows of [A_k0 A_k1 A_k2 c_k], then the 8 dual-quaternion
// components used for blending.  A = diag(sqrt(kappa * exp(log_scale))) * R(q_hat)^T.
constexpr int CTX_STRIDE = 20;
// per-bone per-ray gradient accumulators: gA(9) gc(3) gdq(8)
constexpr int ACC_STRIDE = 20;

MODA_HD void quat_mul(const float* a, const float* b, float* o) {
  const float aw = a[0], ax = a[1], ay = a[2], az = a[3];
  const float bw = b[0], bx = b[1], by = b[2], bz = b[3];
  o[0] = aw * bw - ax * bx - ay * by - az * bz;
  o[1] = aw * bx + ax * bw + ay * bz - az * by;
  o[2] = aw * by - ax * bz + ay * bw + az * bx;
  o[3] = aw * bz + ax * by - ay * bx + az * bw;
}

MODA_HD void quat_conj(const float* a, float* o) {
  o[0] = a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = -a[3];
}

// o = a (x) b  =>  ga += go (x) conj(b),  gb += conj(a) (x) go
MODA_HD void quat_mul_bwd(const float* a, const float* b, const float* go, float* ga, float* gb) {
  float t[4], c[4];
  if (ga) { quat_conj(b, c); quat_mul(go, c, t); for (int i = 0; i < 4; ++i) ga[i] += t[i]; }
  if (gb) { quat_conj(a, c); quat_mul(c, go, t); for (int i = 0; i < 4; ++i) gb[i] += t[i]; }
}

// R row-major 3x3.  two_s = 2/|q|^2 so q need not be unit.
MODA_HD void quat_to_mat(const float* q, float* R) {
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  const float ts = 2.0f / (r * r + i * i + j * j + k * k);
  R[0] = 1 - ts * (j * j + k * k); R[1] = ts * (i * j - k * r);     R[2] = ts * (i * k + j * r);
  R[3] = ts * (i * j + k * r);     R[4] = 1 - ts * (i * i + k * k); R[5] = ts * (j * k - i * r);
  R[6] = ts * (i * k - j * r);     R[7] = ts * (j * k + i * r);     R[8] = 1 - ts * (i * i + j * j);
}

// gq += d(R)/d(q)^T gR, including the dependence of two_s on |q|^2
MODA_HD void quat_to_mat_bwd(const float* q, const float* gR, float* gq) {
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  const float n2 = r * r + i * i + j * j + k * k;
  const float ts = 2.0f / n2;
  const float P0 = -(j * j + k * k), P1 = i * j - k * r, P2 = i * k + j * r;
  const float P3 = i * j + k * r, P4 = -(i * i + k * k), P5 = j * k - i * r;
  const float P6 = i * k - j * r, P7 = j * k + i * r, P8 = -(i * i + j * j);
  const float gts = gR[0] * P0 + gR[1] * P1 + gR[2] * P2 + gR[3] * P3 + gR[4] * P4 + gR[5] * P5 +
                    gR[6] * P6 + gR[7] * P7 + gR[8] * P8;
  const float g0 = ts * gR[0], g1 = ts * gR[1], g2 = ts * gR[2], g3 = ts * gR[3], g4 = ts * gR[4],
              g5 = ts * gR[5], g6 = ts * gR[6], g7 = ts * gR[7], g8 = ts * gR[8];
  const float gn2 = -gts * ts / n2;
  gq[0] += -g1 * k + g2 * j + g3 * k - g5 * i - g6 * j + g7 * i + 2 * r * gn2;
  gq[1] += g1 * j + g2 * k + g3 * j - 2 * g4 * i - g5 * r + g6 * k + g7 * r - 2 * g8 * i + 2 * i * gn2;
  gq[2] += -2 * g0 * j + g1 * i + g2 * r + g3 * i + g5 * k - g6 * r + g7 * k - 2 * g8 * j + 2 * j * gn2;
  gq[3] += -2 * g0 * k - g1 * r + g2 * i + g3 * r - 2 * g4 * k + g5 * j + g6 * i + g7 * j + 2 * k * gn2;
}

// ---- bone_transform (geom_utils.py:73-86): bone (10) moved by the dual quaternion rts (8) ----------
MODA_HD void bone_transform_fwd(const float* bone, const float* rts, float* out) {
  float Rm[9], cj[4], tq[4], qo[4];
  quat_to_mat(rts, Rm);
  quat_conj(rts, cj);
  quat_mul(rts + 4, cj, tq);
  for (int a = 0; a < 3; ++a)
    out[a] = Rm[3 * a] * bone[0] + Rm[3 * a + 1] * bone[1] + Rm[3 * a + 2] * bone[2] + 2 * tq[a + 1];
  quat_mul(rts, bone + 3, qo);
  const float sg = qo[0] < 0 ? -1.0f : 1.0f;  // standardize_quaternion
  for (int a = 0; a < 4; ++a) out[3 + a] = sg * qo[a];
  for (int a = 0; a < 3; ++a) out[7 + a] = bone[7 + a];
}

// gbone (10) += ..., grts (8) += ...
MODA_HD void bone_transform_bwd(const float* bone, const float* rts, const float* gout, float* gbone,
                                float* grts) {
  float Rm[9], cj[4], qo[4];
  quat_to_mat(rts, Rm);
  quat_conj(rts, cj);
  float gRm[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) gRm[3 * a + b] = gout[a] * bone[b];
  for (int b = 0; b < 3; ++b) gbone[b] += Rm[b] * gout[0] + Rm[3 + b] * gout[1] + Rm[6 + b] * gout[2];
  quat_to_mat_bwd(rts, gRm, grts);
  // t = 2 * (d (x) conj(r))[1:]
  float gtq[4] = {0.f, 2 * gout[0], 2 * gout[1], 2 * gout[2]};
  float gcj[4] = {0.f, 0.f, 0.f, 0.f};
  quat_mul_bwd(rts + 4, cj, gtq, grts + 4, gcj);
  grts[0] += gcj[0]; grts[1] -= gcj[1]; grts[2] -= gcj[2]; grts[3] -= gcj[3];
  // orient = sign * (r (x) q)
  quat_mul(rts, bone + 3, qo);
  const float sg = qo[0] < 0 ? -1.0f : 1.0f;
  float gqo[4] = {sg * gout[3], sg * gout[4], sg * gout[5], sg * gout[6]};
  quat_mul_bwd(rts, bone + 3, gqo, grts, gbone + 3);
  for (int a = 0; a < 3; ++a) gbone[7 + a] += gout[7 + a];
}

// ---- dq_inverse (dual_quat.py:87-93) ---------------------------------------------------------------
MODA_HD void dq_inverse_fwd(const float* dq, float* out) {
  const float inv = 1.0f / (dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3]);
  out[0] = dq[0] * inv; out[1] = -dq[1] * inv; out[2] = -dq[2] * inv; out[3] = -dq[3] * inv;
  out[4] = dq[4] * inv; out[5] = -dq[5] * inv; out[6] = -dq[6] * inv; out[7] = -dq[7] * inv;
}

MODA_HD void dq_inverse_bwd(const float* dq, const float* gout, float* gdq) {
  const float n2 = dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3];
  const float inv = 1.0f / n2;
  const float sg[8] = {1, -1, -1, -1, 1, -1, -1, -1};
  float dot = 0.f;
  for (int a = 0; a < 8; ++a) {
    gdq[a] += sg[a] * gout[a] * inv;
    dot += gout[a] * sg[a] * dq[a];
  }
  const float gn2 = -dot * inv * inv;
  for (int a = 0; a < 4; ++a) gdq[a] += 2 * dq[a] * gn2;
}

// ---- per-ray bone context ---------------------------------------------------------------------------
// bone: the (possibly already deformed) Gaussian (10); dq: the dual quaternion used for blending (8);
// kappa = 1000 * exp(skin_aux[0])  (geom_utils.py:265-266: *100*exp(log_scale), then *-10).
MODA_HD void bone_ctx_fwd(const float* bone, const float* dq, float kappa, float* ctx) {
  float qh[4], R[9];
  const float nq = sqrtf(bone[3] * bone[3] + bone[4] * bone[4] + bone[5] * bone[5] + bone[6] * bone[6]);
  const float inq = 1.0f / fmaxf(nq, 1e-12f);  // F.normalize eps
  for (int a = 0; a < 4; ++a) qh[a] = bone[3 + a] * inq;
  quat_to_mat(qh, R);
  for (int k = 0; k < 3; ++k) {
    const float w = sqrtf(kappa * expf(bone[7 + k]));
    ctx[4 * k + 0] = w * R[0 + k];  // A[k][j] = w_k * R[j][k]
    ctx[4 * k + 1] = w * R[3 + k];
    ctx[4 * k + 2] = w * R[6 + k];
    ctx[4 * k + 3] = bone[k];
  }
  for (int a = 0; a < 8; ++a) ctx[12 + a] = dq[a];
}

// acc = [gA(9, row-major k,j) | gc(3) | gdq(8)] summed over the ray's points.
// gbone (10) +=, gdq (8) +=, *gaux0 += d/d(skin_aux[0]).
MODA_HD void bone_ctx_bwd(const float* bone, float kappa, const float* acc, float* gbone, float* gdq,
                          float* gaux0) {
  float qh[4], R[9], gR[9];
  const float nq = sqrtf(bone[3] * bone[3] + bone[4] * bone[4] + bone[5] * bone[5] + bone[6] * bone[6]);
  const float inq = 1.0f / fmaxf(nq, 1e-12f);
  for (int a = 0; a < 4; ++a) qh[a] = bone[3 + a] * inq;
  quat_to_mat(qh, R);
  float ga = 0.f;
  for (int k = 0; k < 3; ++k) {
    const float w = sqrtf(kappa * expf(bone[7 + k]));
    float gw = 0.f;
    for (int j = 0; j < 3; ++j) {
      gR[3 * j + k] = acc[3 * k + j] * w;
      gw += acc[3 * k + j] * R[3 * j + k];
    }
    const float gl = 0.5f * gw * w;  // d w / d log_scale = w/2 ; same for d/d aux0
    gbone[7 + k] += gl;
    ga += gl;
  }
  (void)ga;
  (void)gaux0;  // d/d skin_aux[0] is accumulated per sample in skin_point_bwd (better conditioned)
  float gqh[4] = {0.f, 0.f, 0.f, 0.f};
  quat_to_mat_bwd(qh, gR, gqh);
  const float dot = gqh[0] * qh[0] + gqh[1] * qh[1] + gqh[2] * qh[2] + gqh[3] * qh[3];
  for (int a = 0; a < 4; ++a) gbone[3 + a] += (gqh[a] - qh[a] * dot) * inq;
  for (int a = 0; a < 3; ++a) gbone[a] += acc[9 + a];
  for (int a = 0; a < 8; ++a) gdq[a] += acc[12 + a];
}

// Gaussian logit of one bone at point p (without the MLP delta)
MODA_HD float bone_logit(const float* c, float px, float py, float pz) {
  const float dx = c[3] - px, dy = c[7] - py, dz = c[11] - pz;
  const float u0 = c[0] * dx + c[1] * dy + c[2] * dz;
  const float u1 = c[4] * dx + c[5] * dy + c[6] * dz;
  const float u2 = c[8] * dx + c[9] * dy + c[10] * dz;
  return -(u0 * u0 + u1 * u1 + u2 * u2);
}

struct BlendState {
  float c[8];   // normalised blended dual quaternion
  float inv_n;  // 1/|b_real|
};

// y = DQ transform of p by the normalised blend (geom_utils.py:481-491)
MODA_HD void dq_apply(const float* c, float px, float py, float pz, float* y) {
  const float a0 = c[0], dx = c[1], dy = c[2], dz = c[3], ae = c[4], ex = c[5], ey = c[6], ez = c[7];
  // u = d0 x p + a0 p
  const float ux = dy * pz - dz * py + a0 * px;
  const float uy = dz * px - dx * pz + a0 * py;
  const float uz = dx * py - dy * px + a0 * pz;
  // d0 x u, d0 x de
  const float vx = dy * uz - dz * uy, vy = dz * ux - dx * uz, vz = dx * uy - dy * ux;
  const float wx = dy * ez - dz * ey, wy = dz * ex - dx * ez, wz = dx * ey - dy * ex;
  y[0] = px + 2 * vx + 2 * (a0 * ex - ae * dx + wx);
  y[1] = py + 2 * vy + 2 * (a0 * ey - ae * dy + wy);
  y[2] = pz + 2 * vz + 2 * (a0 * ez - ae * dz + wz);
}

// adjoint of dq_apply followed by the normalisation c = b / |b[:4]|:
// outputs gp (3, overwritten) and gb (8, overwritten) = grad w.r.t. the *unnormalised* blend b.
MODA_HD void dq_apply_bwd(const float* c, float inv_n, float px, float py, float pz, const float* gy,
                          float* gp, float* gb) {
  const float a0 = c[0], dx = c[1], dy = c[2], dz = c[3], ae = c[4], ex = c[5], ey = c[6], ez = c[7];
  const float gx = gy[0], gyy = gy[1], gz = gy[2];
  // gu = 2 (gy x d0)
  const float gux = 2 * (gyy * dz - gz * dy), guy = 2 * (gz * dx - gx * dz), guz = 2 * (gx * dy - gyy * dx);
  // gp = gy + gu x d0 + a0 gu
  gp[0] = gx + (guy * dz - guz * dy) + a0 * gux;
  gp[1] = gyy + (guz * dx - gux * dz) + a0 * guy;
  gp[2] = gz + (gux * dy - guy * dx) + a0 * guz;
  const float ux = dy * pz - dz * py + a0 * px;
  const float uy = dz * px - dx * pz + a0 * py;
  const float uz = dx * py - dy * px + a0 * pz;
  float gc[8];
  gc[0] = gux * px + guy * py + guz * pz + 2 * (gx * ex + gyy * ey + gz * ez);
  // gd0 = 2 (u x gy) + p x gu - 2 ae gy + 2 (de x gy)
  gc[1] = 2 * (uy * gz - uz * gyy) + (py * guz - pz * guy) - 2 * ae * gx + 2 * (ey * gz - ez * gyy);
  gc[2] = 2 * (uz * gx - ux * gz) + (pz * gux - px * guz) - 2 * ae * gyy + 2 * (ez * gx - ex * gz);
  gc[3] = 2 * (ux * gyy - uy * gx) + (px * guy - py * gux) - 2 * ae * gz + 2 * (ex * gyy - ey * gx);
  gc[4] = -2 * (gx * dx + gyy * dy + gz * dz);
  // gde = 2 a0 gy + 2 (gy x d0)
  gc[5] = 2 * a0 * gx + gux;
  gc[6] = 2 * a0 * gyy + guy;
  gc[7] = 2 * a0 * gz + guz;
  float dot = 0.f;
  for (int a = 0; a < 8; ++a) dot += gc[a] * c[a];
  for (int a = 0; a < 8; ++a) gb[a] = gc[a] * inv_n;
  for (int a = 0; a < 4; ++a) gb[a] -= c[a] * dot * inv_n;
}


// ---- one sample through skinning weights + blend (shared by the kernels and the CPU adjoint check) ------
// ctx: B contexts of CTX_STRIDE floats.  dl: delta logits (B) or null.  win: given weights (B) or null.
// Returns the blended (weight-normalised, not yet unit) dual quaternion in bl, softmax max / sum for reuse.
MODA_HD void skin_point_blend(const float* ctx, int B, float px, float py, float pz, const float* dl,
                              const float* win, float* bl, float* mx_out, float* sum_out, float* gref_out = nullptr) {
  for (int i = 0; i < 8; ++i) bl[i] = 0.f;
  if (win) {
    for (int b = 0; b < B; ++b) {
      const float wb = win[b];
      const float* c = ctx + b * CTX_STRIDE + 12;
      for (int i = 0; i < 8; ++i) bl[i] += wb * c[i];
    }
    *mx_out = 0.f; *sum_out = 1.f;
    if (gref_out) *gref_out = 0.f;
    return;
  }
  float mx = -INFINITY, gref = 0.f;
  for (int b = 0; b < B; ++b) {
    const float lg = bone_logit(ctx + b * CTX_STRIDE, px, py, pz);
    const float l = dl ? lg + dl[b] : lg;
    if (l > mx) { mx = l; gref = lg; }
  }
  if (gref_out) *gref_out = gref;
  float sum = 0.f;
  for (int b = 0; b < B; ++b) {
    float l = bone_logit(ctx + b * CTX_STRIDE, px, py, pz);
    if (dl) l += dl[b];
    const float e = expf(l - mx);
    sum += e;
    const float* c = ctx + b * CTX_STRIDE + 12;
    for (int i = 0; i < 8; ++i) bl[i] += e * c[i];
  }
  const float inv = 1.0f / sum;
  for (int i = 0; i < 8; ++i) bl[i] *= inv;
  *mx_out = mx; *sum_out = sum;
}

// Adjoint of one sample.  `live` = false makes every contribution an exact zero (tail lanes of a warp
// still have to take part in the reductions done by `emit`).  emit(b, v, any) receives, per bone, the 20
// per-ray accumulands [gA(9) | gc(3) | gdq(8)]; `any` says whether this sample contributes at all.
// gy: gradient on the warped point (or null); gsk: gradient on the weights (or null).
// Outputs: gp[3]; gdl[b] (if dl-mode and gdl != null); gwin[b] (if win-mode and gwin != null).
template <class Emit>
MODA_HD void skin_point_bwd(const float* ctx, int B, float px, float py, float pz, const float* dl,
                            const float* win, const float* gy, const float* gsk, bool live, float* gp,
                            float* gdl, float* gwin, float* gaux_pt, Emit& emit) {
  float bl[8], mx, sum, gref;
  skin_point_blend(ctx, B, px, py, pz, dl, win, bl, &mx, &sum, &gref);
  // d/d skin_aux[0] = sum_b gl_b * (Gaussian logit_b).  sum_b gl_b = 0, so the logits are shifted by the
  // arg-max bone's (gref) before the product: same value, without the cancellation of O(100) offsets.
  float gaux = 0.f;
  const float inv_sum = 1.0f / sum;
  float gb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  gp[0] = gp[1] = gp[2] = 0.f;
  if (gy) {
    const float n = sqrtf(bl[0] * bl[0] + bl[1] * bl[1] + bl[2] * bl[2] + bl[3] * bl[3]);
    const float inv_n = live ? 1.0f / n : 0.f;
    float c[8];
    for (int i = 0; i < 8; ++i) c[i] = bl[i] * inv_n;
    float g3[3] = {live ? gy[0] : 0.f, live ? gy[1] : 0.f, live ? gy[2] : 0.f};
    dq_apply_bwd(c, inv_n, px, py, pz, g3, gp, gb);
  }
  float gs = 0.f;  // sum_b W_b gW_b
  if (!win) {
    for (int b = 0; b < B; ++b) {
      float l = bone_logit(ctx + b * CTX_STRIDE, px, py, pz);
      if (dl) l += dl[b];
      const float w = expf(l - mx) * inv_sum;
      const float* c = ctx + b * CTX_STRIDE + 12;
      float gw = (gsk && live) ? gsk[b] : 0.f;
      for (int i = 0; i < 8; ++i) gw += gb[i] * c[i];
      gs += w * gw;
    }
  }
  for (int b = 0; b < B; ++b) {
    const float* cb = ctx + b * CTX_STRIDE;
    const float* c = cb + 12;
    float gw = (gsk && live) ? gsk[b] : 0.f;
    for (int i = 0; i < 8; ++i) gw += gb[i] * c[i];
    float w, gl = 0.f;
    float v[ACC_STRIDE];
    for (int i = 0; i < 12; ++i) v[i] = 0.f;
    if (win) {
      w = live ? win[b] : 0.f;
      if (gwin && live) gwin[b] = gw;
    } else {
      const float dx = cb[3] - px, dy = cb[7] - py, dz = cb[11] - pz;
      const float u0 = cb[0] * dx + cb[1] * dy + cb[2] * dz;
      const float u1 = cb[4] * dx + cb[5] * dy + cb[6] * dz;
      const float u2 = cb[8] * dx + cb[9] * dy + cb[10] * dz;
      const float lgauss = -(u0 * u0 + u1 * u1 + u2 * u2);
      const float l = dl ? lgauss + dl[b] : lgauss;
      w = live ? expf(l - mx) * inv_sum : 0.f;
      gl = w * (gw - gs);
      gaux = fmaf(gl, lgauss - gref, gaux);
      if (gdl && live) gdl[b] = gl;
      const float g0 = -2.f * gl * u0, g1 = -2.f * gl * u1, g2 = -2.f * gl * u2;  // gu
      v[0] = g0 * dx; v[1] = g0 * dy; v[2] = g0 * dz;
      v[3] = g1 * dx; v[4] = g1 * dy; v[5] = g1 * dz;
      v[6] = g2 * dx; v[7] = g2 * dy; v[8] = g2 * dz;
      const float gdx = cb[0] * g0 + cb[4] * g1 + cb[8] * g2;  // A^T gu
      const float gdy = cb[1] * g0 + cb[5] * g1 + cb[9] * g2;
      const float gdz = cb[2] * g0 + cb[6] * g1 + cb[10] * g2;
      v[9] = gdx; v[10] = gdy; v[11] = gdz;
      gp[0] -= gdx; gp[1] -= gdy; gp[2] -= gdz;
    }
    for (int i = 0; i < 8; ++i) v[12 + i] = w * gb[i];
    emit(b, v, (gl != 0.f) || (w != 0.f));
  }
  *gaux_pt = gaux;
}

// Everything the first B threads of a CTA do per bone, forward: rest bone + ray DQ -> bone used + context.
MODA_HD void ray_bone_setup(const float* bone, const float* rts /*or null*/, int deform, int invert,
                            float kappa, float* bone_used, float* ctx) {
  float dq[8];
  for (int i = 0; i < 8; ++i) dq[i] = 0.f;
  if (rts) {
    if (deform) bone_transform_fwd(bone, rts, bone_used);
    else for (int i = 0; i < 10; ++i) bone_used[i] = bone[i];
    if (invert) dq_inverse_fwd(rts, dq);
    else for (int i = 0; i < 8; ++i) dq[i] = rts[i];
  } else {
    for (int i = 0; i < 10; ++i) bone_used[i] = bone[i];
  }
  bone_ctx_fwd(bone_used, dq, kappa, ctx);
}

// ... and backward: accumulators -> gbone (10, overwritten), grts (8, overwritten), *gaux0 +=.
MODA_HD void ray_bone_setup_bwd(const float* bone, const float* rts, int deform, int invert, float kappa,
                                const float* bone_used, const float* acc, float* gbone, float* grts,
                                float* gaux0) {
  float gbl[10], gdq[8];
  for (int i = 0; i < 10; ++i) { gbl[i] = 0.f; gbone[i] = 0.f; }
  for (int i = 0; i < 8; ++i) { gdq[i] = 0.f; grts[i] = 0.f; }
  bone_ctx_bwd(bone_used, kappa, acc, gbl, gdq, gaux0);
  if (rts) {
    if (invert) dq_inverse_bwd(rts, gdq, grts);
    else for (int i = 0; i < 8; ++i) grts[i] = gdq[i];
    if (deform) bone_transform_bwd(bone, rts, gbl, gbone, grts);
    else for (int i = 0; i < 10; ++i) gbone[i] = gbl[i];
  } else {
    for (int i = 0; i < 10; ++i) gbone[i] = gbl[i];
  }
}

// ---- VolSDF-style density (rendering.py:199-207) ------------------------------------------------------
// returns alpha; optionally d alpha / d sigma_raw, d alpha / d ibeta, d alpha / d delta
MODA_HD float density_alpha(float sigma_raw, float delta, float ibeta, float* da_dsigma, float* da_dib,
                            float* da_ddelta) {
  const float sdf = -sigma_raw;
  const float sg = (sdf > 0.f) ? 1.f : ((sdf < 0.f) ? -1.f : 0.f);
  const float em1 = expm1f(-fabsf(sdf) * ibeta);
  const float psi = 0.5f + 0.5f * sg * em1;
  const float dens = psi * ibeta;
  const float ex = expf(-delta * dens);
  if (da_dsigma) {
    const float e = em1 + 1.0f;
    const float ddens_dsigma = (sigma_raw != 0.f) ? 0.5f * ibeta * ibeta * e : 0.f;
    const float ddens_dib = psi - 0.5f * ibeta * sdf * e * ((sdf != 0.f) ? 1.f : 0.f);
    *da_dsigma = delta * ex * ddens_dsigma;
    *da_dib = delta * ex * ddens_dib;
    *da_ddelta = dens * ex;
  }
  return 1.0f - ex;
}

}  // namespace moda
