// Gaussian-bone skinning weights + dual-quaternion blend skinning (forward / backward warps).
//
// Replaces, fused and without materialising any (rays, samples, bones, ...) intermediate:
//   bone_transform           nnutils/geom_utils.py:59-111
//   skinning / skinning_chunk nnutils/geom_utils.py:237-302   (+ vec_to_sim3 :187-199, axis_rotate :231-235)
//   dqs_blend_skinning(_chunk) nnutils/geom_utils.py:457-517
//   neu_dbs                  nnutils/geom_utils.py:372-456    (dq_inverse, dual_quat.py:87-93)
//
// Layout: one CTA handles up to SKIN_THREADS samples of ONE ray; the ray's B bones are turned into a
// 20-float context each (moda_math.h: A = diag(sqrt(kappa*s)) R^T, centre, blend DQ) in shared memory by
// the first B threads, after which every thread owns one sample and reads the contexts as broadcast
// LDS.128.  Softmax runs in two passes that recompute the 15-flop logit instead of spilling B logits.
#include "common.cuh"

namespace moda {

constexpr int SKIN_THREADS = 128;

struct SkinArgs {
  const float* pts;       // (R,S,3)
  const float* bones;     // (B,10) shared, or (R,B,10) if bones_per_ray
  const float* rts;       // (R,B,8) or null (skin-only)
  const float* skin_aux;  // device scalar pair; [0] = log scale
  const float* dskin;     // (R,S,B) or null : MLP delta logits
  const float* skin_in;   // (R,S,B) or null : externally supplied weights (dqs_blend_skinning API)
  float* y;               // (R,S,3) or null
  float* skin_out;        // (R,S,B) or null
  int R, S, B;
  int ldd;            // row pitch of dskin / gdskin (>= B)
  int bones_per_ray;  // bones indexed by ray
  int deform;         // apply bone_transform(bones, rts) first (backward warp)
  int invert;         // blend with dq_inverse(rts) (backward warp)
  // backward only
  const float* gy;     // (R,S,3) or null
  const float* gskin;  // (R,S,B) or null : gradient arriving on skin_out
  float* gpts;         // (R,S,3) or null (overwritten)
  float* gdskin;       // (R,S,B) or null (overwritten)
  float* gskin_in;     // (R,S,B) or null (overwritten)
  float* grts;         // (R,B,8) or null (accumulated: zero-initialised by the caller)
  float* gbones;       // (B,10) or (R,B,10) (accumulated)
  float* gaux;         // (2,) accumulated ([1] never touched: unused in the reference, geom_utils.py:245)
};

__device__ __forceinline__ void build_ctx(const SkinArgs& a, int ray, float* ctx /*smem B*20*/,
                                          float* bone_s /*smem B*10: bones actually used*/) {
  const int b = threadIdx.x;
  if (b < a.B) {
    const float kappa = 1000.0f * expf(a.skin_aux[0]);
    const float* bone = a.bones + ((size_t)(a.bones_per_ray ? ray : 0) * a.B + b) * 10;
    float bn[10], rr[8], used[10], c[CTX_STRIDE];
    for (int i = 0; i < 10; ++i) bn[i] = bone[i];
    if (a.rts) {
      const float* r = a.rts + ((size_t)ray * a.B + b) * 8;
      for (int i = 0; i < 8; ++i) rr[i] = r[i];
    }
    ray_bone_setup(bn, a.rts ? rr : nullptr, a.deform, a.invert, kappa, used, c);
    for (int i = 0; i < 10; ++i) bone_s[b * 10 + i] = used[i];
    for (int i = 0; i < CTX_STRIDE; ++i) ctx[b * CTX_STRIDE + i] = c[i];
  }
}

__global__ void __launch_bounds__(SKIN_THREADS) skin_warp_fwd_kernel(SkinArgs a) {
  __shared__ __align__(16) float ctx[MAX_BONES * CTX_STRIDE];
  __shared__ float bone_s[MAX_BONES * 10];
  const int ray = blockIdx.y;
  build_ctx(a, ray, ctx, bone_s);
  __syncthreads();
  const int s = blockIdx.x * SKIN_THREADS + threadIdx.x;
  if (s >= a.S) return;
  const size_t pi = (size_t)ray * a.S + s;
  const float px = a.pts[pi * 3], py = a.pts[pi * 3 + 1], pz = a.pts[pi * 3 + 2];
  const int B = a.B;
  const float* dl = a.dskin ? a.dskin + pi * a.ldd : nullptr;
  const float* win = a.skin_in ? a.skin_in + pi * B : nullptr;
  float bl[8], mx, sum;
  skin_point_blend(ctx, B, px, py, pz, dl, win, bl, &mx, &sum);
  if (a.skin_out && !win) {
    const float inv = 1.0f / sum;
    for (int b = 0; b < B; ++b) {
      float l = bone_logit(ctx + b * CTX_STRIDE, px, py, pz);
      if (dl) l += dl[b];
      a.skin_out[pi * B + b] = expf(l - mx) * inv;
    }
  }
  if (a.y) {
    const float inv_n = 1.0f / sqrtf(bl[0] * bl[0] + bl[1] * bl[1] + bl[2] * bl[2] + bl[3] * bl[3]);
    float c[8], y[3];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = bl[i] * inv_n;
    dq_apply(c, px, py, pz, y);
    a.y[pi * 3] = y[0]; a.y[pi * 3 + 1] = y[1]; a.y[pi * 3 + 2] = y[2];
  }
}

// warp-reduces the 20 per-bone accumulands of one bone and adds them to the CTA's shared accumulators;
// a bone that no sample of the warp touches contributes exact zeros and is skipped.
struct WarpEmit {
  float* acc;
  int lane;
  __device__ __forceinline__ void operator()(int b, const float* v, bool any) {
    if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
    for (int i = 0; i < ACC_STRIDE; ++i) {
      const float t = warp_sum(v[i]);
      if (lane == 0 && t != 0.f) atomicAdd(&acc[b * ACC_STRIDE + i], t);
    }
  }
};

// Backward.  Each thread recomputes its sample's forward (skin_point_bwd), the per-ray/per-bone sums
// (gA, gc, gdq) are reduced with warp shuffles + shared-memory atomics, and the first B threads push them
// through the context / bone_transform / dq_inverse adjoints (ray_bone_setup_bwd).
__global__ void __launch_bounds__(SKIN_THREADS) skin_warp_bwd_kernel(SkinArgs a) {
  __shared__ __align__(16) float ctx[MAX_BONES * CTX_STRIDE];
  __shared__ float bone_s[MAX_BONES * 10];
  __shared__ float acc[MAX_BONES * ACC_STRIDE];
  const int ray = blockIdx.y;
  const int B = a.B;
  build_ctx(a, ray, ctx, bone_s);
  for (int i = threadIdx.x; i < B * ACC_STRIDE; i += SKIN_THREADS) acc[i] = 0.f;
  __syncthreads();
  const int s = blockIdx.x * SKIN_THREADS + threadIdx.x;
  const bool live = s < a.S;
  const size_t pi = (size_t)ray * a.S + (live ? s : 0);
  const int lane = threadIdx.x & 31;
  const float px = a.pts[pi * 3], py = a.pts[pi * 3 + 1], pz = a.pts[pi * 3 + 2];
  const float* dl = a.dskin ? a.dskin + pi * a.ldd : nullptr;
  const float* win = a.skin_in ? a.skin_in + pi * B : nullptr;
  const float* gsk = a.gskin ? a.gskin + pi * B : nullptr;
  const float* gy = (a.gy && a.rts) ? a.gy + pi * 3 : nullptr;
  float gp[3], gaux_pt;
  WarpEmit emit{acc, lane};
  skin_point_bwd(ctx, B, px, py, pz, dl, win, gy, gsk, live, gp, a.gdskin ? a.gdskin + pi * a.ldd : nullptr,
                 a.gskin_in ? a.gskin_in + pi * B : nullptr, &gaux_pt, emit);
  if (a.gpts && live) { a.gpts[pi * 3] = gp[0]; a.gpts[pi * 3 + 1] = gp[1]; a.gpts[pi * 3 + 2] = gp[2]; }
  __syncthreads();

  // ---- per-bone epilogue ----
  const int b = threadIdx.x;
  float gaux0 = gaux_pt;
  if (b < B) {
    const float kappa = 1000.0f * expf(a.skin_aux[0]);
    const float* bone = a.bones + ((size_t)(a.bones_per_ray ? ray : 0) * B + b) * 10;
    float bn[10], rr[8], gbone[10], grt[8];
    for (int i = 0; i < 10; ++i) bn[i] = bone[i];
    if (a.rts) {
      const float* r = a.rts + ((size_t)ray * B + b) * 8;
      for (int i = 0; i < 8; ++i) rr[i] = r[i];
    }
    float unused = 0.f;
    ray_bone_setup_bwd(bn, a.rts ? rr : nullptr, a.deform, a.invert, kappa, bone_s + b * 10,
                       acc + b * ACC_STRIDE, gbone, grt, &unused);
    if (a.grts && a.rts) {
      float* o = a.grts + ((size_t)ray * B + b) * 8;
      for (int i = 0; i < 8; ++i) atomicAdd(o + i, grt[i]);
    }
    if (a.gbones) {
      float* o = a.gbones + ((size_t)(a.bones_per_ray ? ray : 0) * B + b) * 10;
      for (int i = 0; i < 10; ++i)
        if (gbone[i] != 0.f) atomicAdd(o + i, gbone[i]);
    }
  }
  if (a.gaux) {
    __shared__ float red[SKIN_THREADS / 32];
    const float t = warp_sum(gaux0);
    if (lane == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < SKIN_THREADS / 32; ++i) tot += red[i];
      if (tot != 0.f) atomicAdd(a.gaux, tot);
    }
  }
}

// ---- stand-alone bone_transform (geom_utils.py:59-111) ------------------------------------------------
__global__ void bone_transform_fwd_kernel(const float* bones, const float* rts, float* out, int R, int B,
                                          int bones_per_ray) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * B) return;
  const int ray = i / B, b = i % B;
  bone_transform_fwd(bones + ((size_t)(bones_per_ray ? ray : 0) * B + b) * 10, rts + (size_t)i * 8,
                     out + (size_t)i * 10);
}

__global__ void bone_transform_bwd_kernel(const float* bones, const float* rts, const float* gout,
                                          float* gbones, float* grts, int R, int B, int bones_per_ray) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * B) return;
  const int ray = i / B, b = i % B;
  float gb[10], gr[8];
  for (int k = 0; k < 10; ++k) gb[k] = 0.f;
  for (int k = 0; k < 8; ++k) gr[k] = 0.f;
  bone_transform_bwd(bones + ((size_t)(bones_per_ray ? ray : 0) * B + b) * 10, rts + (size_t)i * 8,
                     gout + (size_t)i * 10, gb, gr);
  for (int k = 0; k < 8; ++k) grts[(size_t)i * 8 + k] = gr[k];
  float* o = gbones + ((size_t)(bones_per_ray ? ray : 0) * B + b) * 10;
  if (bones_per_ray) for (int k = 0; k < 10; ++k) o[k] = gb[k];
  else for (int k = 0; k < 10; ++k) atomicAdd(o + k, gb[k]);
}

}  // namespace moda

using namespace moda;

static int skin_check(const SkinArgs& a) {
  MODA_REQUIRE(a.B >= 1 && a.B <= MAX_BONES, "skin_warp: B=%d outside [1,%d]", a.B, MAX_BONES);
  MODA_REQUIRE(a.R >= 0 && a.S >= 1 && a.ldd >= a.B, "skin_warp: bad sizes R=%d S=%d ld_dskin=%d", a.R, a.S, a.ldd);
  MODA_REQUIRE(a.pts && a.bones && a.skin_aux, "skin_warp: null pts/bones/skin_aux");
  MODA_REQUIRE(!(a.deform || a.invert) || a.rts, "skin_warp: deform/invert need rts");
  MODA_REQUIRE(a.R <= 65535 * 1024, "skin_warp: too many rays");
  return 0;
}

extern "C" int moda_skin_warp_fwd(const float* pts, const float* bones, const float* rts,
                                  const float* skin_aux, const float* dskin, const float* skin_in, float* y,
                                  float* skin_out, int R, int S, int B, int ld_dskin, int bones_per_ray, int deform,
                                  int invert, cudaStream_t stream) {
  SkinArgs a = {};
  a.ldd = ld_dskin > 0 ? ld_dskin : B;
  a.pts = pts; a.bones = bones; a.rts = rts; a.skin_aux = skin_aux; a.dskin = dskin; a.skin_in = skin_in;
  a.y = y; a.skin_out = skin_out; a.R = R; a.S = S; a.B = B; a.bones_per_ray = bones_per_ray;
  a.deform = deform; a.invert = invert;
  if (int e = skin_check(a)) return e;
  MODA_REQUIRE(!y || rts, "skin_warp_fwd: y requested without rts");
  if (R == 0) return 0;
  // grid.y is limited to 65535: fold rays beyond that into several launches
  for (int r0 = 0; r0 < R; r0 += 65535) {
    SkinArgs c = a;
    const int rc = (R - r0 < 65535) ? R - r0 : 65535;
    const size_t po = (size_t)r0 * S;
    c.pts += po * 3; if (c.rts) c.rts += (size_t)r0 * B * 8;
    if (c.bones_per_ray) c.bones += (size_t)r0 * B * 10;
    if (c.dskin) c.dskin += po * c.ldd; if (c.skin_in) c.skin_in += po * B;
    if (c.y) c.y += po * 3; if (c.skin_out) c.skin_out += po * B;
    c.R = rc;
    dim3 grid(cdiv(S, SKIN_THREADS), rc);
    skin_warp_fwd_kernel<<<grid, SKIN_THREADS, 0, stream>>>(c);
  }
  return check_launch("skin_warp_fwd");
}

extern "C" int moda_skin_warp_bwd(const float* pts, const float* bones, const float* rts,
                                  const float* skin_aux, const float* dskin, const float* skin_in,
                                  const float* gy, const float* gskin, float* gpts, float* gdskin,
                                  float* gskin_in, float* grts, float* gbones, float* gaux, int R, int S,
                                  int B, int ld_dskin, int bones_per_ray, int deform, int invert,
                                  cudaStream_t stream) {
  SkinArgs a = {};
  a.ldd = ld_dskin > 0 ? ld_dskin : B;
  a.pts = pts; a.bones = bones; a.rts = rts; a.skin_aux = skin_aux; a.dskin = dskin; a.skin_in = skin_in;
  a.R = R; a.S = S; a.B = B; a.bones_per_ray = bones_per_ray; a.deform = deform; a.invert = invert;
  a.gy = gy; a.gskin = gskin; a.gpts = gpts; a.gdskin = gdskin; a.gskin_in = gskin_in; a.grts = grts;
  a.gbones = gbones; a.gaux = gaux;
  if (int e = skin_check(a)) return e;
  if (R == 0) return 0;
  for (int r0 = 0; r0 < R; r0 += 65535) {
    SkinArgs c = a;
    const int rc = (R - r0 < 65535) ? R - r0 : 65535;
    const size_t po = (size_t)r0 * S;
    c.pts += po * 3; if (c.rts) c.rts += (size_t)r0 * B * 8;
    if (c.bones_per_ray) { c.bones += (size_t)r0 * B * 10; if (c.gbones) c.gbones += (size_t)r0 * B * 10; }
    if (c.dskin) c.dskin += po * c.ldd; if (c.skin_in) c.skin_in += po * B;
    if (c.gy) c.gy += po * 3; if (c.gskin) c.gskin += po * B;
    if (c.gpts) c.gpts += po * 3; if (c.gdskin) c.gdskin += po * c.ldd; if (c.gskin_in) c.gskin_in += po * B;
    if (c.grts) c.grts += (size_t)r0 * B * 8;
    c.R = rc;
    dim3 grid(cdiv(S, SKIN_THREADS), rc);
    skin_warp_bwd_kernel<<<grid, SKIN_THREADS, 0, stream>>>(c);
  }
  return check_launch("skin_warp_bwd");
}

extern "C" int moda_bone_transform_fwd(const float* bones, const float* rts, float* out, int R, int B,
                                       int bones_per_ray, cudaStream_t stream) {
  MODA_REQUIRE(bones && rts && out, "bone_transform_fwd: null pointer");
  if (R * B == 0) return 0;
  bone_transform_fwd_kernel<<<cdiv((long long)R * B, 128), 128, 0, stream>>>(bones, rts, out, R, B, bones_per_ray);
  return check_launch("bone_transform_fwd");
}

extern "C" int moda_bone_transform_bwd(const float* bones, const float* rts, const float* gout,
                                       float* gbones, float* grts, int R, int B, int bones_per_ray,
                                       cudaStream_t stream) {
  MODA_REQUIRE(bones && rts && gout && gbones && grts, "bone_transform_bwd: null pointer");
  if (R * B == 0) return 0;
  bone_transform_bwd_kernel<<<cdiv((long long)R * B, 128), 128, 0, stream>>>(bones, rts, gout, gbones, grts, R, B,
                                                                  bones_per_ray);
  return check_launch("bone_transform_bwd");
}
