// Small per-element kernels: positional encoding, ray sampling, dual-quaternion algebra.
//   Embedding.forward          nnutils/nerf.py:35-75
//   sample generation          nnutils/rendering.py:64-89
//   dq_* / q_*                 nnutils/dual_quat.py:4-93
#include "common.cuh"

namespace moda {

constexpr int MAX_FREQS_E = 16;
struct Window { float w[MAX_FREQS_E]; };

// out (M, C*(1+2F)) = [x | w0 sin x | w0 cos x | w1 sin 2x | ...]
__global__ void embed_fwd_kernel(const float* x, int ldx, float* out, int ldo, long long M, int C, int F,
                                 Window win) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * C) return;
  const long long m = t / C;
  const int c = (int)(t % C);
  const float v = x[m * ldx + c];
  float* o = out + m * ldo;
  o[c] = v;
  for (int k = 0; k < F; ++k) {
    float sn, cs;
    sincosf(v * (float)(1 << k), &sn, &cs);
    o[C + (2 * k) * C + c] = win.w[k] * sn;
    o[C + (2 * k + 1) * C + c] = win.w[k] * cs;
  }
}

// gx (M,C) (=|+=) d out / d x ^T gout
__global__ void embed_bwd_kernel(const float* x, int ldx, const float* gout, int ldo, float* gx, int ldg,
                                 long long M, int C, int F, Window win, int accumulate) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * C) return;
  const long long m = t / C;
  const int c = (int)(t % C);
  const float v = x[m * ldx + c];
  const float* g = gout + m * ldo;
  float acc = g[c];
  for (int k = 0; k < F; ++k) {
    const float f = (float)(1 << k);
    float sn, cs;
    sincosf(v * f, &sn, &cs);
    acc += win.w[k] * f * (cs * g[C + (2 * k) * C + c] - sn * g[C + (2 * k + 1) * C + c]);
  }
  float* o = gx + m * ldg + c;
  *o = accumulate ? (*o + acc) : acc;
}

// torch.linspace(0,1,S)[i] in fp32 (symmetric evaluation, as ATen's linspace kernel does)
__device__ __forceinline__ float linspace01(int i, int S) {
  if (S == 1) return 0.f;
  const float step = 1.0f / (float)(S - 1);
  return (i < S / 2) ? step * (float)i : 1.0f - step * (float)(S - 1 - i);
}

__device__ __forceinline__ float depth_at(float near, float far, int i, int S, int use_disp) {
  const float s = linspace01(i, S);
  if (!use_disp) return near * (1 - s) + far * s;
  return 1.0f / (1.0f / near * (1 - s) + 1.0f / far * s);
}

// z (R,S), xyz (R,S,3), dn (R,3) = d/|d|      rendering.py:64-89
__global__ void sample_rays_kernel(const float* o, const float* d, const float* near, const float* far,
                                   const float* jitter, float perturb, int use_disp, float* z, float* xyz,
                                   float* dn, int R, int S) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * S) return;
  const int r = (int)(t / S), i = (int)(t % S);
  const float nr = near[r], fr = far[r];
  float zi = depth_at(nr, fr, i, S, use_disp);
  if (perturb > 0.f) {
    const float zl = depth_at(nr, fr, max(i - 1, 0), S, use_disp);
    const float zu = depth_at(nr, fr, min(i + 1, S - 1), S, use_disp);
    const float lower = (i == 0) ? zi : 0.5f * (zl + zi);
    const float upper = (i == S - 1) ? zi : 0.5f * (zi + zu);
    zi = lower + (upper - lower) * (perturb * jitter[t]);
  }
  z[t] = zi;
  const float dx = d[r * 3], dy = d[r * 3 + 1], dz = d[r * 3 + 2];
  xyz[t * 3] = o[r * 3] + dx * zi;
  xyz[t * 3 + 1] = o[r * 3 + 1] + dy * zi;
  xyz[t * 3 + 2] = o[r * 3 + 2] + dz * zi;
  if (i == 0 && dn) {
    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    dn[r * 3] = dx * inv; dn[r * 3 + 1] = dy * inv; dn[r * 3 + 2] = dz * inv;
  }
}

// xyz (R,S,3) = o + d * z for given depths (importance-sampled second pass, rendering.py:112-113)
__global__ void points_from_depths_kernel(const float* o, const float* d, const float* z, float* xyz, int R,
                                          int S) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * S) return;
  const int r = (int)(t / S);
  const float zi = z[t];
  xyz[t * 3] = o[r * 3] + d[r * 3] * zi;
  xyz[t * 3 + 1] = o[r * 3 + 1] + d[r * 3 + 1] * zi;
  xyz[t * 3 + 2] = o[r * 3 + 2] + d[r * 3 + 2] * zi;
}

// one warp per ray: go = sum_s gxyz, gd = sum_s z gxyz + (gnd * d/|d|) + normalise-adjoint of gdn
__global__ void sample_rays_bwd_kernel(const float* d, const float* z, const float* gxyz, const float* gdn,
                                       const float* gnd, float* go, float* gd, int R, int S) {
  const int r = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
  if (gxyz) {
    for (int i = lane; i < S; i += 32) {
      const size_t t = (size_t)r * S + i;
      const float zi = z[t];
      const float g0 = gxyz[t * 3], g1 = gxyz[t * 3 + 1], g2 = gxyz[t * 3 + 2];
      a0 += g0; a1 += g1; a2 += g2;
      b0 += zi * g0; b1 += zi * g1; b2 += zi * g2;
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  b0 = warp_sum(b0); b1 = warp_sum(b1); b2 = warp_sum(b2);
  if (lane == 0) {
    const float dx = d[r * 3], dy = d[r * 3 + 1], dz = d[r * 3 + 2];
    const float n = sqrtf(dx * dx + dy * dy + dz * dz), inv = 1.0f / n;
    const float nx = dx * inv, ny = dy * inv, nz = dz * inv;
    if (gnd) { const float g = gnd[r]; b0 += g * nx; b1 += g * ny; b2 += g * nz; }
    if (gdn) {
      const float g0 = gdn[r * 3], g1 = gdn[r * 3 + 1], g2 = gdn[r * 3 + 2];
      const float dot = g0 * nx + g1 * ny + g2 * nz;
      b0 += (g0 - nx * dot) * inv; b1 += (g1 - ny * dot) * inv; b2 += (g2 - nz * dot) * inv;
    }
    if (go) { go[r * 3] = a0; go[r * 3 + 1] = a1; go[r * 3 + 2] = a2; }
    if (gd) { gd[r * 3] = b0; gd[r * 3 + 1] = b1; gd[r * 3 + 2] = b2; }
  }
}

// ---- dual quaternion algebra ------------------------------------------------------------------------
enum { DQ_QCONJ = 0, DQ_CCONJ = 1, DQ_NORMALIZE = 2, DQ_INVERSE = 3, Q_NORMALIZE = 4 };

__global__ void dq_unary_fwd_kernel(int op, const float* in, float* out, long long n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (op == Q_NORMALIZE) {
    const float* q = in + t * 4;
    const float inv = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) out[t * 4 + i] = q[i] * inv;
    return;
  }
  float a[8], o[8];
  for (int i = 0; i < 8; ++i) a[i] = in[t * 8 + i];
  if (op == DQ_QCONJ) {
    const float s[8] = {1, -1, -1, -1, 1, -1, -1, -1};
    for (int i = 0; i < 8; ++i) o[i] = a[i] * s[i];
  } else if (op == DQ_CCONJ) {
    const float s[8] = {1, -1, -1, -1, -1, 1, 1, 1};
    for (int i = 0; i < 8; ++i) o[i] = a[i] * s[i];
  } else if (op == DQ_NORMALIZE) {
    const float inv = 1.0f / sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
    for (int i = 0; i < 8; ++i) o[i] = a[i] * inv;
  } else {
    dq_inverse_fwd(a, o);
  }
  for (int i = 0; i < 8; ++i) out[t * 8 + i] = o[i];
}

__global__ void dq_unary_bwd_kernel(int op, const float* in, const float* gout, float* gin, long long n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (op == Q_NORMALIZE) {
    const float* q = in + t * 4;
    const float* g = gout + t * 4;
    const float inv = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float dot = 0.f;
    for (int i = 0; i < 4; ++i) dot += g[i] * q[i] * inv;
    for (int i = 0; i < 4; ++i) gin[t * 4 + i] = (g[i] - q[i] * inv * dot) * inv;
    return;
  }
  float a[8], g[8], o[8];
  for (int i = 0; i < 8; ++i) { a[i] = in[t * 8 + i]; g[i] = gout[t * 8 + i]; o[i] = 0.f; }
  if (op == DQ_QCONJ) {
    const float s[8] = {1, -1, -1, -1, 1, -1, -1, -1};
    for (int i = 0; i < 8; ++i) o[i] = g[i] * s[i];
  } else if (op == DQ_CCONJ) {
    const float s[8] = {1, -1, -1, -1, -1, 1, 1, 1};
    for (int i = 0; i < 8; ++i) o[i] = g[i] * s[i];
  } else if (op == DQ_NORMALIZE) {
    const float n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    const float inv = rsqrtf(n2);
    float dot = 0.f;  // sum_i g_i a_i (all 8) -> d/d|r|
    for (int i = 0; i < 8; ++i) dot += g[i] * a[i];
    for (int i = 0; i < 8; ++i) o[i] = g[i] * inv;
    for (int i = 0; i < 4; ++i) o[i] -= a[i] * dot * inv / n2;
  } else {
    dq_inverse_bwd(a, g, o);
  }
  for (int i = 0; i < 8; ++i) gin[t * 8 + i] = o[i];
}

// width 4: q_mul ; width 8: dq_mul  (dual_quat.py:14-49)
__global__ void dq_mul_fwd_kernel(const float* a, const float* b, float* out, long long n, int width) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float x[8], y[8], o[8];
  for (int i = 0; i < width; ++i) { x[i] = a[t * width + i]; y[i] = b[t * width + i]; }
  quat_mul(x, y, o);
  if (width == 8) {
    float t1[4], t2[4];
    quat_mul(x, y + 4, t1);
    quat_mul(x + 4, y, t2);
    for (int i = 0; i < 4; ++i) o[4 + i] = t1[i] + t2[i];
  }
  for (int i = 0; i < width; ++i) out[t * width + i] = o[i];
}

__global__ void dq_mul_bwd_kernel(const float* a, const float* b, const float* gout, float* ga, float* gb,
                                  long long n, int width) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float x[8], y[8], g[8], gx[8], gy[8];
  for (int i = 0; i < width; ++i) {
    x[i] = a[t * width + i]; y[i] = b[t * width + i]; g[i] = gout[t * width + i];
    gx[i] = 0.f; gy[i] = 0.f;
  }
  quat_mul_bwd(x, y, g, gx, gy);
  if (width == 8) {
    quat_mul_bwd(x, y + 4, g + 4, gx, gy + 4);
    quat_mul_bwd(x + 4, y, g + 4, gx + 4, gy);
  }
  for (int i = 0; i < width; ++i) { ga[t * width + i] = gx[i]; gb[t * width + i] = gy[i]; }
}

}  // namespace moda

using namespace moda;

static int fill_window(Window& w, const float* win, int F) {
  MODA_REQUIRE(F >= 0 && F <= MAX_FREQS_E, "embed: n_freqs=%d outside [0,%d]", F, MAX_FREQS_E);
  for (int i = 0; i < MAX_FREQS_E; ++i) w.w[i] = (win && i < F) ? win[i] : 1.0f;
  return 0;
}

extern "C" int moda_embed_fwd(const float* x, int ldx, float* out, int ldo, long long M, int C, int F,
                              const float* win, cudaStream_t stream) {
  Window w;
  if (int e = fill_window(w, win, F)) return e;
  if (M * C == 0) return 0;
  MODA_REQUIRE(x && out && ldx >= C && ldo >= C * (1 + 2 * F), "embed_fwd: bad arguments");
  embed_fwd_kernel<<<cdiv(M * C, 256), 256, 0, stream>>>(x, ldx, out, ldo, M, C, F, w);
  return check_launch("embed_fwd");
}

extern "C" int moda_embed_bwd(const float* x, int ldx, const float* gout, int ldo, float* gx, int ldg,
                              long long M, int C, int F, const float* win, int accumulate,
                              cudaStream_t stream) {
  Window w;
  if (int e = fill_window(w, win, F)) return e;
  if (M * C == 0) return 0;
  MODA_REQUIRE(x && gout && gx && ldx >= C && ldg >= C && ldo >= C * (1 + 2 * F), "embed_bwd: bad arguments");
  embed_bwd_kernel<<<cdiv(M * C, 256), 256, 0, stream>>>(x, ldx, gout, ldo, gx, ldg, M, C, F, w, accumulate);
  return check_launch("embed_bwd");
}

extern "C" int moda_sample_rays_fwd(const float* o, const float* d, const float* near, const float* far,
                                    const float* jitter, float perturb, int use_disp, float* z, float* xyz,
                                    float* dn, int R, int S, cudaStream_t stream) {
  MODA_REQUIRE(o && d && near && far && z && xyz, "sample_rays_fwd: null pointer");
  MODA_REQUIRE(perturb <= 0.f || jitter, "sample_rays_fwd: perturb>0 needs the jitter tensor");
  if ((long long)R * S == 0) return 0;
  sample_rays_kernel<<<cdiv((long long)R * S, 256), 256, 0, stream>>>(o, d, near, far, jitter, perturb, use_disp, z,
                                                                    xyz, dn, R, S);
  return check_launch("sample_rays_fwd");
}

extern "C" int moda_points_from_depths(const float* o, const float* d, const float* z, float* xyz, int R, int S,
                                       cudaStream_t stream) {
  MODA_REQUIRE(o && d && z && xyz, "points_from_depths: null pointer");
  if ((long long)R * S == 0) return 0;
  points_from_depths_kernel<<<cdiv((long long)R * S, 256), 256, 0, stream>>>(o, d, z, xyz, R, S);
  return check_launch("points_from_depths");
}

extern "C" int moda_sample_rays_bwd(const float* d, const float* z, const float* gxyz, const float* gdn,
                                    const float* gnd, float* go, float* gd, int R, int S, cudaStream_t stream) {
  MODA_REQUIRE(d && z, "sample_rays_bwd: null pointer");
  if (R == 0) return 0;
  sample_rays_bwd_kernel<<<cdiv(R, 4), 128, 0, stream>>>(d, z, gxyz, gdn, gnd, go, gd, R, S);
  return check_launch("sample_rays_bwd");
}

extern "C" int moda_dq_unary_fwd(int op, const float* in, float* out, long long n, cudaStream_t stream) {
  MODA_REQUIRE(op >= 0 && op <= 4 && in && out, "dq_unary_fwd: bad arguments");
  if (n == 0) return 0;
  dq_unary_fwd_kernel<<<cdiv(n, 128), 128, 0, stream>>>(op, in, out, n);
  return check_launch("dq_unary_fwd");
}

extern "C" int moda_dq_unary_bwd(int op, const float* in, const float* gout, float* gin, long long n,
                                 cudaStream_t stream) {
  MODA_REQUIRE(op >= 0 && op <= 4 && in && gout && gin, "dq_unary_bwd: bad arguments");
  if (n == 0) return 0;
  dq_unary_bwd_kernel<<<cdiv(n, 128), 128, 0, stream>>>(op, in, gout, gin, n);
  return check_launch("dq_unary_bwd");
}

extern "C" int moda_dq_mul_fwd(const float* a, const float* b, float* out, long long n, int width,
                               cudaStream_t stream) {
  MODA_REQUIRE((width == 4 || width == 8) && a && b && out, "dq_mul_fwd: bad arguments");
  if (n == 0) return 0;
  dq_mul_fwd_kernel<<<cdiv(n, 128), 128, 0, stream>>>(a, b, out, n, width);
  return check_launch("dq_mul_fwd");
}

extern "C" int moda_dq_mul_bwd(const float* a, const float* b, const float* gout, float* ga, float* gb,
                               long long n, int width, cudaStream_t stream) {
  MODA_REQUIRE((width == 4 || width == 8) && a && b && gout && ga && gb, "dq_mul_bwd: bad arguments");
  if (n == 0) return 0;
  dq_mul_bwd_kernel<<<cdiv(n, 128), 128, 0, stream>>>(a, b, gout, ga, gb, n, width);
  return check_launch("dq_mul_bwd");
}


// ------------------------------------------------------------------------------------------------ AdamW on the flat buffer
// The optimiser step of the reference's training loop (torch.optim.AdamW, nnutils/train_utils.py:177-222) on the ONE flat
// parameter / gradient buffer of parallel.FlatParams: decoupled weight decay, bias-corrected moments, same arithmetic
// order as torch's fused implementation.  torch's multi-tensor kernel walks a single 650k-element tensor in 64k-element
// chunks on ~10 CTAs (78 us); a plain grid over the buffer takes ~5 us, which is 4 % of a 1024-ray step.
// state (device, 3 floats): [0] step count, [1] lr / (1 - beta1^t), [2] sqrt(1 - beta2^t); updated by the prepare kernel so
// that the step can live inside a CUDA graph.
namespace moda {
__global__ void adamw_prepare_kernel(float* state, float lr, float b1, float b2) {
  const float t = state[0] + 1.0f;
  state[0] = t;
  state[1] = (float)((double)lr / (1.0 - pow((double)b1, (double)t)));
  state[2] = (float)sqrt(1.0 - pow((double)b2, (double)t));
}
__global__ void adamw_flat_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                  float4* __restrict__ v, long long n4, const float* __restrict__ state, float lr, float b1,
                                  float b2, float eps, float wd) {
  const float step_size = state[1], bc2_sqrt = state[2];
  const float decay = 1.0f - lr * wd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P_ = &pp.x; const float* G_ = &gg.x; float* M_ = &mm.x; float* V_ = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float x = P_[k] * decay;
      M_[k] = M_[k] + (1.0f - b1) * (G_[k] - M_[k]);          // lerp, as torch
      V_[k] = b2 * V_[k] + (1.0f - b2) * G_[k] * G_[k];
      const float denom = sqrtf(V_[k]) / bc2_sqrt + eps;
      P_[k] = x - step_size * (M_[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}
}  // namespace moda

// p, g, m, v: n floats, n % 4 == 0, 16-byte aligned (FlatParams pads every tensor to 64 floats); state: 3 floats
extern "C" int moda_adamw_flat(float* p, const float* g, float* m, float* v, long long n, float* state, float lr, float b1,
                               float b2, float eps, float wd, cudaStream_t stream) {
  MODA_REQUIRE(p && g && m && v && state && n % 4 == 0, "adamw_flat: bad arguments (n must be a multiple of 4)");
  MODA_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adamw_flat: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  using namespace moda;
  adamw_prepare_kernel<<<1, 1, 0, stream>>>(state, lr, b1, b2);
  const long long n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
  adamw_flat_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
                                               reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4, state, lr, b1,
                                               b2, eps, wd);
  return check_launch("adamw_flat");
}
