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
  int gcopies;        // backward: replicated (B,10) / (2) accumulators behind gbones / gaux (ray r adds into copy r % gcopies)
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
  if (a.gdskin && live)   // pad columns of a pitched row: always written, so callers need not pre-zero the buffer
    for (int b2 = B; b2 < a.ldd; ++b2) a.gdskin[pi * a.ldd + b2] = 0.f;
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
      float* o = a.gbones + ((size_t)(a.bones_per_ray ? ray : ray % a.gcopies) * B + b) * 10;
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
      if (tot != 0.f) atomicAdd(a.gaux + 2 * (ray % a.gcopies), tot);
    }
  }
}

// ---- fast path --------------------------------------------------------------------------------------------
// The configuration of the training step and of the skinning microbenchmark: weights from the Gaussian logits
// (+ MLP delta logits), B <= 32, warped point requested, no externally supplied weights, no weight output.
// What differs from the general kernels above:
//  * the delta-logit rows of a warp (32 rows x ld floats, contiguous in HBM) are staged through a per-warp
//    shared-memory tile with fully coalesced 128-bit loads; a thread then walks its own row with conflict-free
//    LDS (the general kernel's per-thread row walk costs 32 L1 sectors per instruction and is L1-bound);
//  * the tile is rewritten in place: logits -> exp(l - max) -> d loss / d logit, so no pass recomputes the
//    Gaussian logit, and the logit gradients leave through the same coalesced path (pad columns zeroed);
//  * backward: the 20 per-bone accumulands of a warp's 32 samples are reduced by transposing them through
//    shared memory (20 STS + 8 LDS.128 + 32 FADD per bone and lane) instead of 100 shuffles.
constexpr int FAST_BONES = 32;
constexpr int LT = 33;   // row pitch of the per-warp logit tile: column walks of 32 rows hit 32 banks
constexpr int RT = 36;   // row pitch of the per-warp reduction tile: 16-byte aligned rows, conflict-free LDS.128

__device__ __forceinline__ void stage_rows(const float* __restrict__ g, int nrows, int ld, int B, float* L, int lane) {
  const int n = nrows * ld;
  if ((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const int l4 = ld >> 2, n4 = n >> 2;
    for (int i = lane; i < n4; i += 32) {
      const float4 v = __ldg(g4 + i);
      const int r = i / l4, c = (i - r * l4) * 4;
      float* d = L + r * LT + c;
      if (c < B) d[0] = v.x;
      if (c + 1 < B) d[1] = v.y;
      if (c + 2 < B) d[2] = v.z;
      if (c + 3 < B) d[3] = v.w;
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      const int r = i / ld, c = i - r * ld;
      if (c < B) L[r * LT + c] = __ldg(g + i);
    }
  }
}

__device__ __forceinline__ void unstage_rows(float* __restrict__ g, int nrows, int ld, int B, const float* L, int lane) {
  const int n = nrows * ld;
  if ((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    float4* g4 = reinterpret_cast<float4*>(g);
    const int l4 = ld >> 2, n4 = n >> 2;
    for (int i = lane; i < n4; i += 32) {
      const int r = i / l4, c = (i - r * l4) * 4;
      const float* d = L + r * LT + c;
      float4 v;
      v.x = (c < B) ? d[0] : 0.f;
      v.y = (c + 1 < B) ? d[1] : 0.f;
      v.z = (c + 2 < B) ? d[2] : 0.f;
      v.w = (c + 3 < B) ? d[3] : 0.f;
      g4[i] = v;
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      const int r = i / ld, c = i - r * ld;
      g[i] = (c < B) ? L[r * LT + c] : 0.f;
    }
  }
}

__global__ void __launch_bounds__(SKIN_THREADS) skin_warp_fwd_fast_kernel(SkinArgs a) {
  __shared__ __align__(16) float ctx[FAST_BONES * CTX_STRIDE];
  __shared__ float bone_s[FAST_BONES * 10];
  __shared__ float Lt[SKIN_THREADS / 32][32 * LT];
  const int ray = blockIdx.x, B = a.B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = blockIdx.y * SKIN_THREADS + warp * 32;
  const int nrows = min(32, a.S - s0);
  float* L = Lt[warp];
  if (a.dskin && nrows > 0) stage_rows(a.dskin + ((size_t)ray * a.S + s0) * a.ldd, nrows, a.ldd, B, L, lane);
  build_ctx(a, ray, ctx, bone_s);
  __syncthreads();
  if (lane >= nrows) return;
  const size_t pi = (size_t)ray * a.S + s0 + lane;
  const float px = __ldg(a.pts + pi * 3), py = __ldg(a.pts + pi * 3 + 1), pz = __ldg(a.pts + pi * 3 + 2);
  float* Lr = L + lane * LT;
  const bool has_dl = a.dskin != nullptr;
  float mx = -INFINITY;
  for (int b = 0; b < B; ++b) {
    float l = bone_logit(ctx + b * CTX_STRIDE, px, py, pz);
    if (has_dl) l += Lr[b];
    Lr[b] = l;
    mx = fmaxf(mx, l);
  }
  float sum = 0.f, bl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int b = 0; b < B; ++b) {
    const float e = expf(Lr[b] - mx);
    sum += e;
    const float4 c0 = *reinterpret_cast<const float4*>(ctx + b * CTX_STRIDE + 12);
    const float4 c1 = *reinterpret_cast<const float4*>(ctx + b * CTX_STRIDE + 16);
    bl[0] = fmaf(e, c0.x, bl[0]); bl[1] = fmaf(e, c0.y, bl[1]); bl[2] = fmaf(e, c0.z, bl[2]); bl[3] = fmaf(e, c0.w, bl[3]);
    bl[4] = fmaf(e, c1.x, bl[4]); bl[5] = fmaf(e, c1.y, bl[5]); bl[6] = fmaf(e, c1.z, bl[6]); bl[7] = fmaf(e, c1.w, bl[7]);
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) bl[i] *= inv;
  const float inv_n = 1.0f / sqrtf(bl[0] * bl[0] + bl[1] * bl[1] + bl[2] * bl[2] + bl[3] * bl[3]);
  float c[8], y[3];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = bl[i] * inv_n;
  dq_apply(c, px, py, pz, y);
  a.y[pi * 3] = y[0]; a.y[pi * 3 + 1] = y[1]; a.y[pi * 3 + 2] = y[2];
}

__global__ void __launch_bounds__(SKIN_THREADS) skin_warp_bwd_fast_kernel(SkinArgs a) {
  __shared__ __align__(16) float ctx[FAST_BONES * CTX_STRIDE];
  __shared__ float bone_s[FAST_BONES * 10];
  // per-warp accumulators: a warp visits every bone once, so its 20 sums for bone b are a plain store (fp32 atomics
  // on shared memory are a compare-and-swap spin loop; with dense skinning weights they were the kernel's bottleneck:
  // the cycle warp's adjoint ran 2x slower than the backward warp's on the same instruction count)
  __shared__ float accw[SKIN_THREADS / 32][FAST_BONES * ACC_STRIDE];
  __shared__ float Lt[SKIN_THREADS / 32][32 * LT];
  __shared__ __align__(16) float Rt[SKIN_THREADS / 32][ACC_STRIDE * RT];
  __shared__ float red[SKIN_THREADS / 32];
  const int ray = blockIdx.x, B = a.B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = blockIdx.y * SKIN_THREADS + warp * 32;
  const int nrows = max(0, min(32, a.S - s0));
  float* L = Lt[warp];
  float* R = Rt[warp];
  if (a.dskin && nrows > 0) stage_rows(a.dskin + ((size_t)ray * a.S + s0) * a.ldd, nrows, a.ldd, B, L, lane);
  build_ctx(a, ray, ctx, bone_s);
  for (int i = threadIdx.x; i < (SKIN_THREADS / 32) * FAST_BONES * ACC_STRIDE; i += SKIN_THREADS) (&accw[0][0])[i] = 0.f;
  __syncthreads();
  const bool live = lane < nrows;
  const size_t pi = (size_t)ray * a.S + (live ? s0 + lane : 0);
  const float px = __ldg(a.pts + pi * 3), py = __ldg(a.pts + pi * 3 + 1), pz = __ldg(a.pts + pi * 3 + 2);
  float g3[3] = {0.f, 0.f, 0.f};
  if (live) { g3[0] = __ldg(a.gy + pi * 3); g3[1] = __ldg(a.gy + pi * 3 + 1); g3[2] = __ldg(a.gy + pi * 3 + 2); }
  float* Lr = L + lane * LT;
  const bool has_dl = a.dskin != nullptr;
  // pass 1: logits, their maximum, and the arg-max bone's Gaussian logit (reference point of d/d skin_aux[0])
  float mx = -INFINITY, gref = 0.f;
  for (int b = 0; b < B; ++b) {
    const float lg = bone_logit(ctx + b * CTX_STRIDE, px, py, pz);
    const float l = (has_dl && live) ? lg + Lr[b] : lg;
    Lr[b] = l;
    if (l > mx) { mx = l; gref = lg; }
  }
  // pass 2: exp, sum, blend
  float sum = 0.f, bl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int b = 0; b < B; ++b) {
    const float e = expf(Lr[b] - mx);
    Lr[b] = e;
    sum += e;
    const float4 c0 = *reinterpret_cast<const float4*>(ctx + b * CTX_STRIDE + 12);
    const float4 c1 = *reinterpret_cast<const float4*>(ctx + b * CTX_STRIDE + 16);
    bl[0] = fmaf(e, c0.x, bl[0]); bl[1] = fmaf(e, c0.y, bl[1]); bl[2] = fmaf(e, c0.z, bl[2]); bl[3] = fmaf(e, c0.w, bl[3]);
    bl[4] = fmaf(e, c1.x, bl[4]); bl[5] = fmaf(e, c1.y, bl[5]); bl[6] = fmaf(e, c1.z, bl[6]); bl[7] = fmaf(e, c1.w, bl[7]);
  }
  const float inv_sum = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) bl[i] *= inv_sum;
  float gp[3], gb[8];
  {
    const float n = sqrtf(bl[0] * bl[0] + bl[1] * bl[1] + bl[2] * bl[2] + bl[3] * bl[3]);
    const float inv_n = live ? 1.0f / n : 0.f;
    float c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = bl[i] * inv_n;
    dq_apply_bwd(c, inv_n, px, py, pz, g3, gp, gb);
  }
  // pass 3: gs = sum_b W_b gW_b with gW_b = gb . dq_b
  float gs = 0.f;
  for (int b = 0; b < B; ++b) {
    const float4 c0 = *reinterpret_cast<const float4*>(ctx + b * CTX_STRIDE + 12);
    const float4 c1 = *reinterpret_cast<const float4*>(ctx + b * CTX_STRIDE + 16);
    const float gw = gb[0] * c0.x + gb[1] * c0.y + gb[2] * c0.z + gb[3] * c0.w + gb[4] * c1.x + gb[5] * c1.y +
                     gb[6] * c1.z + gb[7] * c1.w;
    gs = fmaf(Lr[b] * inv_sum, gw, gs);
  }
  // pass 4: logit gradients and the per-ray / per-bone accumulands [gA(9) | gc(3) | gdq(8)]
  float gaux = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* cb = ctx + b * CTX_STRIDE;
    const float4 r0 = *reinterpret_cast<const float4*>(cb);
    const float4 r1 = *reinterpret_cast<const float4*>(cb + 4);
    const float4 r2 = *reinterpret_cast<const float4*>(cb + 8);
    const float4 c0 = *reinterpret_cast<const float4*>(cb + 12);
    const float4 c1 = *reinterpret_cast<const float4*>(cb + 16);
    const float gw = gb[0] * c0.x + gb[1] * c0.y + gb[2] * c0.z + gb[3] * c0.w + gb[4] * c1.x + gb[5] * c1.y +
                     gb[6] * c1.z + gb[7] * c1.w;
    const float w = live ? Lr[b] * inv_sum : 0.f;
    const float gl = w * (gw - gs);
    Lr[b] = gl;
    const bool any = (gl != 0.f) || (w != 0.f);
    if (!__any_sync(0xffffffffu, any)) continue;   // exact zeros from every sample of the warp: nothing to add
    const float dx = r0.w - px, dy = r1.w - py, dz = r2.w - pz;
    const float u0 = r0.x * dx + r0.y * dy + r0.z * dz;
    const float u1 = r1.x * dx + r1.y * dy + r1.z * dz;
    const float u2 = r2.x * dx + r2.y * dy + r2.z * dz;
    const float lgauss = -(u0 * u0 + u1 * u1 + u2 * u2);
    gaux = fmaf(gl, lgauss - gref, gaux);
    const float g0 = -2.f * gl * u0, g1 = -2.f * gl * u1, g2 = -2.f * gl * u2;
    const float gdx = r0.x * g0 + r1.x * g1 + r2.x * g2;   // A^T gu
    const float gdy = r0.y * g0 + r1.y * g1 + r2.y * g2;
    const float gdz = r0.z * g0 + r1.z * g1 + r2.z * g2;
    gp[0] -= gdx; gp[1] -= gdy; gp[2] -= gdz;
    float* Rl = R + lane;
    Rl[0 * RT] = g0 * dx; Rl[1 * RT] = g0 * dy; Rl[2 * RT] = g0 * dz;
    Rl[3 * RT] = g1 * dx; Rl[4 * RT] = g1 * dy; Rl[5 * RT] = g1 * dz;
    Rl[6 * RT] = g2 * dx; Rl[7 * RT] = g2 * dy; Rl[8 * RT] = g2 * dz;
    Rl[9 * RT] = gdx; Rl[10 * RT] = gdy; Rl[11 * RT] = gdz;
#pragma unroll
    for (int i = 0; i < 8; ++i) Rl[(12 + i) * RT] = w * gb[i];
    __syncwarp();
    if (lane < ACC_STRIDE) {
      const float4* row = reinterpret_cast<const float4*>(R + lane * RT);
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 x = row[k];
        t0 += x.x; t1 += x.y; t2 += x.z; t3 += x.w;
      }
      accw[warp][b * ACC_STRIDE + lane] = (t0 + t1) + (t2 + t3);
    }
    __syncwarp();
  }
  if (live) { a.gpts[pi * 3] = gp[0]; a.gpts[pi * 3 + 1] = gp[1]; a.gpts[pi * 3 + 2] = gp[2]; }
  if (a.gdskin && nrows > 0) {
    __syncwarp();
    unstage_rows(a.gdskin + ((size_t)ray * a.S + s0) * a.ldd, nrows, a.ldd, B, L, lane);
  }
  __syncthreads();

  // ---- per-bone epilogue (as in the general kernel) ----
  const int b = threadIdx.x;
  if (b < B) {
    const float kappa = 1000.0f * expf(a.skin_aux[0]);
    const float* bone = a.bones + ((size_t)(a.bones_per_ray ? ray : 0) * B + b) * 10;
    float bn[10], rr[8], gbone[10], grt[8];
    for (int i = 0; i < 10; ++i) bn[i] = bone[i];
    const float* r = a.rts + ((size_t)ray * B + b) * 8;
    for (int i = 0; i < 8; ++i) rr[i] = r[i];
    float unused = 0.f;
    float accb[ACC_STRIDE];
#pragma unroll
    for (int i = 0; i < ACC_STRIDE; ++i) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < SKIN_THREADS / 32; ++w) t += accw[w][b * ACC_STRIDE + i];
      accb[i] = t;
    }
    ray_bone_setup_bwd(bn, rr, a.deform, a.invert, kappa, bone_s + b * 10, accb, gbone, grt, &unused);
    if (a.grts) {
      float* o = a.grts + ((size_t)ray * B + b) * 8;
      for (int i = 0; i < 8; ++i) atomicAdd(o + i, grt[i]);
    }
    if (a.gbones) {
      float* o = a.gbones + ((size_t)(a.bones_per_ray ? ray : ray % a.gcopies) * B + b) * 10;
      for (int i = 0; i < 10; ++i)
        if (gbone[i] != 0.f) atomicAdd(o + i, gbone[i]);
    }
  }
  if (a.gaux) {
    const float t = warp_sum(gaux);
    if (lane == 0) red[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < SKIN_THREADS / 32; ++i) tot += red[i];
      if (tot != 0.f) atomicAdd(a.gaux + 2 * (ray % a.gcopies), tot);
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
  if (B <= FAST_BONES && y && rts && !skin_in && !skin_out) {
    skin_warp_fwd_fast_kernel<<<dim3(R, cdiv(S, SKIN_THREADS)), SKIN_THREADS, 0, stream>>>(a);
    return check_launch("skin_warp_fwd");
  }
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
                                  int gcopies, cudaStream_t stream) {
  SkinArgs a = {};
  a.gcopies = gcopies > 0 ? gcopies : 1;
  a.ldd = ld_dskin > 0 ? ld_dskin : B;
  a.pts = pts; a.bones = bones; a.rts = rts; a.skin_aux = skin_aux; a.dskin = dskin; a.skin_in = skin_in;
  a.R = R; a.S = S; a.B = B; a.bones_per_ray = bones_per_ray; a.deform = deform; a.invert = invert;
  a.gy = gy; a.gskin = gskin; a.gpts = gpts; a.gdskin = gdskin; a.gskin_in = gskin_in; a.grts = grts;
  a.gbones = gbones; a.gaux = gaux;
  if (int e = skin_check(a)) return e;
  if (R == 0) return 0;
  if (B <= FAST_BONES && gy && rts && gpts && !skin_in && !gskin && !gskin_in) {
    skin_warp_bwd_fast_kernel<<<dim3(R, cdiv(S, SKIN_THREADS)), SKIN_THREADS, 0, stream>>>(a);
    return check_launch("skin_warp_bwd");
  }
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
