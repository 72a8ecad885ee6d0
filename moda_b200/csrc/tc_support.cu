// SIMT helpers around the tensor-core trunk path: fp16 operand staging (positional encoding, weight packing),
// the 1- and 3-wide output heads, bias / per-ray reductions of fp16 gradients, and the dynamic loss scale that
// keeps the fp16 gradient chain inside the normal range.
//   Embedding.forward          nnutils/nerf.py:35-75      (pe16_*)
//   sigma / rgb heads          nnutils/nerf.py:178, 188-195 (head_*)
#include <cuda_fp16.h>

#include "common.cuh"

namespace moda {

struct Win16 { float w[16]; };

// out (P, 64) fp16 = [x(3) | w_k sin(2^k x) | w_k cos(2^k x)]_{k<10} | 0     (63 channels + one zero pad)
__global__ void pe16_fwd_kernel(const float* __restrict__ xyz, __half* __restrict__ out, __half* __restrict__ out_lo,
                                long long P, int F, int ldo, Win16 win) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float x[3] = {xyz[p * 3], xyz[p * 3 + 1], xyz[p * 3 + 2]};
  __half* o = out + p * ldo;
  __align__(16) __half buf[64];
  __align__(16) __half lob[64];
  auto put = [&](int i, float v) {
    const __half h = __float2half_rn(v);
    buf[i] = h;
    lob[i] = __float2half_rn(v - __half2float(h));
  };
  put(0, x[0]); put(1, x[1]); put(2, x[2]);
  for (int k = 0; k < F; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sn, cs;
      sincosf(x[c] * f, &sn, &cs);
      put(3 + 6 * k + c, win.w[k] * sn);
      put(3 + 6 * k + 3 + c, win.w[k] * cs);
    }
  }
  for (int i = 3 + 6 * F; i < 64; ++i) put(i, 0.f);
  const uint4* b4 = reinterpret_cast<const uint4*>(buf);
  uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
  for (int i = 0; i < 8; ++i) o4[i] = b4[i];
  if (out_lo) {
    const uint4* l4 = reinterpret_cast<const uint4*>(lob);
    uint4* ol4 = reinterpret_cast<uint4*>(out_lo + p * ldo);
#pragma unroll
    for (int i = 0; i < 8; ++i) ol4[i] = l4[i];
  }
}

// gxyz (P,3) (=|+=) inv_scale * dPE/dx^T g16
__global__ void pe16_bwd_kernel(const float* __restrict__ xyz, const __half* __restrict__ g16,
                                const __half* __restrict__ g16lo, int ldg, float* __restrict__ gxyz, long long P, int F,
                                Win16 win, const float* inv_scale, int accumulate) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float is = inv_scale ? *inv_scale : 1.0f;
  __align__(16) __half buf[64];
  const uint4* g4 = reinterpret_cast<const uint4*>(g16 + p * ldg);
  uint4* b4 = reinterpret_cast<uint4*>(buf);
#pragma unroll
  for (int i = 0; i < 8; ++i) b4[i] = g4[i];
  __align__(16) __half lob[64];
  if (g16lo) {
    const uint4* l4 = reinterpret_cast<const uint4*>(g16lo + p * ldg);
    uint4* lb4 = reinterpret_cast<uint4*>(lob);
#pragma unroll
    for (int i = 0; i < 8; ++i) lb4[i] = l4[i];
  }
  auto get = [&](int i) { return g16lo ? __half2float(buf[i]) + __half2float(lob[i]) : __half2float(buf[i]); };
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = xyz[p * 3 + c];
    float acc = get(c);
    // unrolled over the maximum frequency count with a guard: the gradient row then stays in registers (a run-time
    // trip count turned `buf` into a local-memory array and the kernel into a stack-traffic benchmark)
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      if (k >= F) break;
      const float f = (float)(1 << k);
      // same evaluation as the chain kernel's PE producers (chain.cu): exact two-term Cody-Waite reduction to
      // [-pi, pi], then the SFU (absolute error < 1e-6 for |2^k x| <= 160 rad); the libm sincosf it replaces made
      // this kernel instruction-bound at ~40 instructions per call
      const float r = x * f;
      const float kk = rintf(r * 0.15915494309189535f);
      float r2 = fmaf(kk, -6.2831854820251465f, r);
      r2 = fmaf(kk, 1.7484555e-7f, r2);
      const float sn = __sinf(r2), cs = __cosf(r2);
      acc += win.w[k] * f * (cs * get(3 + 6 * k + c) - sn * get(3 + 6 * k + 3 + c));
    }
    acc *= is;
    float* o = gxyz + p * 3 + c;
    *o = accumulate ? (*o + acc) : acc;
  }
}

// fp32 (rows, ld_in) columns [col0, col0+cols) -> fp16 block of out_rows x width (row pitch ld_out), zero padded.
// transpose=0: out[r][c] = in[r][col0+c]; transpose=1: out[c][r] = in[r][col0+c].  Optionally a second copy of
// the block (out2) and the low half of the split-precision pair (out_lo = fp16(v - fp16(v))).
__global__ void pack16_kernel(const float* __restrict__ in, int ld_in, int rows, int cols, int col0,
                              __half* __restrict__ out, __half* __restrict__ out2, __half* __restrict__ out_lo,
                              int ld_out, int out_rows, int width, int transpose) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= out_rows * width) return;
  const int orow = t / width, ocol = t % width;
  const int r = transpose ? ocol : orow, c = transpose ? orow : ocol;
  const float v = (r < rows && c < cols) ? in[(size_t)r * ld_in + col0 + c] : 0.f;
  const __half h = __float2half_rn(v);
  const size_t o = (size_t)orow * ld_out + ocol;
  out[o] = h;
  if (out2) out2[o] = h;
  if (out_lo) out_lo[o] = __float2half_rn(v - __half2float(h));
}

// several pack16 blocks of one packed-weight matrix in ONE launch (blockIdx.y = job): the per-step repacking of
// a network's weights is a few dozen tiny blocks, which as separate launches cost more than the copies themselves
constexpr int MAX_PACK_JOBS = 16;
struct PackJob {
  const float* in;
  __half* out;
  __half* out_lo;
  int ld_in, rows, cols, col0, out_rows, width, transpose;
};
struct PackJobs {
  PackJob j[MAX_PACK_JOBS];
  int ld_out;
};
__global__ void pack16_multi_kernel(const __grid_constant__ PackJobs js) {
  const PackJob& j = js.j[blockIdx.y];
  const int total = j.out_rows * j.width;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int orow = t / j.width, ocol = t % j.width;
    const int r = j.transpose ? ocol : orow, c = j.transpose ? orow : ocol;
    const float v = (r < j.rows && c < j.cols) ? j.in[(size_t)r * j.ld_in + j.col0 + c] : 0.f;
    const __half h = __float2half_rn(v);
    const size_t o = (size_t)orow * js.ld_out + ocol;
    j.out[o] = h;
    if (j.out_lo) j.out_lo[o] = __float2half_rn(v - __half2float(h));
  }
}

// fp32 (M, cols; row pitch ld_in) * (*scale) -> fp16 (hi, lo) pair of width `width` (zero padded), pitch ld_out
__global__ void split16_kernel(const float* __restrict__ in, int ld_in, int cols, const float* scale_p,
                               __half* __restrict__ hi, __half* __restrict__ lo, int ld_out, int width, long long M) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * width) return;
  const long long m = t / width;
  const int c = (int)(t % width);
  const float sc = scale_p ? *scale_p : 1.0f;
  const float v = (c < cols) ? in[m * ld_in + c] * sc : 0.f;
  const __half h = __float2half_rn(v);
  hi[m * ld_out + c] = h;
  if (lo) lo[m * ld_out + c] = __float2half_rn(v - __half2float(h));
}

// ---- output heads, one warp per sample ----------------------------------------------------------------
// raw (P,4) = [sigmoid(Dfe Wr^T + br) | H8 ws + bs]; H8 (P,256) fp16, Dfe (P,128) fp16
__global__ void __launch_bounds__(256) head_fwd_kernel(const __half* __restrict__ H8, const __half* __restrict__ Dfe,
                                                       const float* __restrict__ ws, const float* __restrict__ bs,
                                                       const float* __restrict__ Wr, const float* __restrict__ br,
                                                       float* __restrict__ raw, long long P) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float wsl[8], wrl[3][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) wsl[j] = ws[lane * 8 + j];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) wrl[c][j] = Wr[c * 128 + lane * 4 + j];
  const float bsv = bs[0], b0 = br[0], b1 = br[1], b2 = br[2];
  for (long long p = warp0; p < P; p += nwarps) {
    const uint4 hv = *reinterpret_cast<const uint4*>(H8 + p * 256 + lane * 8);
    const uint2 dv = *reinterpret_cast<const uint2*>(Dfe + p * 128 + lane * 4);
    const __half2* hh = reinterpret_cast<const __half2*>(&hv);
    const __half2* dh = reinterpret_cast<const __half2*>(&dv);
    float s = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(hh[j]);
      s = fmaf(f.x, wsl[2 * j], s);
      s = fmaf(f.y, wsl[2 * j + 1], s);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float2 f = __half22float2(dh[j]);
      r0 = fmaf(f.x, wrl[0][2 * j], r0); r0 = fmaf(f.y, wrl[0][2 * j + 1], r0);
      r1 = fmaf(f.x, wrl[1][2 * j], r1); r1 = fmaf(f.y, wrl[1][2 * j + 1], r1);
      r2 = fmaf(f.x, wrl[2][2 * j], r2); r2 = fmaf(f.y, wrl[2][2 * j + 1], r2);
    }
    s = warp_sum(s); r0 = warp_sum(r0); r1 = warp_sum(r1); r2 = warp_sum(r2);
    if (lane == 0) {
      float4 o;
      o.x = 1.0f / (1.0f + expf(-(r0 + b0)));
      o.y = 1.0f / (1.0f + expf(-(r1 + b1)));
      o.z = 1.0f / (1.0f + expf(-(r2 + b2)));
      o.w = s + bsv;
      *reinterpret_cast<float4*>(raw + p * 4) = o;
    }
  }
}

// Backward of both heads.  graw (P,4) is the gradient on [rgb | sigma] (unscaled fp32).
//   dDfe16 (P,128) = scale * (g_pre Wr) * (Dfe > 0)   with g_pre = g_rgb * rgb (1 - rgb)
//   gsig (P)       = graw[:,3]                         (rank-1 term of the next dgrad)
//   gWr (3,128), gbr (3), gws (256), gbs (1) accumulated with atomics
__global__ void __launch_bounds__(256) head_bwd_kernel(const __half* __restrict__ H8, const __half* __restrict__ Dfe,
                                                       const float* __restrict__ raw, const float* __restrict__ graw,
                                                       const float* __restrict__ Wr, const float* scale_p,
                                                       __half* __restrict__ dDfe, float* __restrict__ gsig,
                                                       float* gWr, float* gbr, float* gws, float* gbs, long long P) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float scale = scale_p ? *scale_p : 1.0f;
  float wrl[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) wrl[c][j] = Wr[c * 128 + lane * 4 + j];
  float a_ws[8], a_wr[3][4], a_br[3] = {0, 0, 0}, a_bs = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) a_ws[j] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) a_wr[c][j] = 0.f;
  // four consecutive rows per warp and trip, every load of the trip issued before the first use: one row per trip
  // left ~800 bytes per warp in flight and the kernel at a quarter of the HBM rate
  constexpr int U = 4;
  for (long long p0 = warp0 * U; p0 < P; p0 += nwarps * U) {
    float4 rwv[U], gv[U];
    uint4 hvv[U];
    uint2 dvv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = (p0 + u < P) ? p0 + u : P - 1;
      rwv[u] = __ldg(reinterpret_cast<const float4*>(raw + p * 4));
      gv[u] = __ldg(reinterpret_cast<const float4*>(graw + p * 4));
      hvv[u] = __ldg(reinterpret_cast<const uint4*>(H8 + p * 256 + lane * 8));
      dvv[u] = __ldg(reinterpret_cast<const uint2*>(Dfe + p * 128 + lane * 4));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u;
      if (p >= P) break;
      const float4 rw = rwv[u], g = gv[u];
      const float g0 = g.x * rw.x * (1.f - rw.x), g1 = g.y * rw.y * (1.f - rw.y), g2 = g.z * rw.z * (1.f - rw.z);
      const __half2* hh = reinterpret_cast<const __half2*>(&hvv[u]);
      const __half2* dh = reinterpret_cast<const __half2*>(&dvv[u]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hh[j]);
        a_ws[2 * j] = fmaf(g.w, f.x, a_ws[2 * j]);
        a_ws[2 * j + 1] = fmaf(g.w, f.y, a_ws[2 * j + 1]);
      }
      float d[4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float2 f = __half22float2(dh[j]);
        d[2 * j] = f.x; d[2 * j + 1] = f.y;
      }
      uint2 ov;
      __half2* oh = reinterpret_cast<__half2*>(&ov);
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a_wr[0][j] = fmaf(g0, d[j], a_wr[0][j]);
        a_wr[1][j] = fmaf(g1, d[j], a_wr[1][j]);
        a_wr[2][j] = fmaf(g2, d[j], a_wr[2][j]);
        const float v = g0 * wrl[0][j] + g1 * wrl[1][j] + g2 * wrl[2][j];
        o[j] = (d[j] > 0.f) ? v * scale : 0.f;
      }
      oh[0] = __floats2half2_rn(o[0], o[1]);
      oh[1] = __floats2half2_rn(o[2], o[3]);
      *reinterpret_cast<uint2*>(dDfe + p * 128 + lane * 4) = ov;
      if (lane == 0) {
        gsig[p] = g.w;
        a_br[0] += g0; a_br[1] += g1; a_br[2] += g2; a_bs += g.w;
      }
    }
  }
  // block reduction through shared memory, then one atomic per value per block
  __shared__ float red[8][32 * 24];
  const int w = threadIdx.x >> 5;
  float* mine = red[w] + lane * 24;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine[j] = a_ws[j];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) mine[8 + c * 4 + j] = a_wr[c][j];
  mine[20] = a_br[0]; mine[21] = a_br[1]; mine[22] = a_br[2]; mine[23] = a_bs;
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 24; i += blockDim.x) {
    float t = 0.f;
    for (int ww = 0; ww < 8; ++ww) t += red[ww][i];
    const int ln = i / 24, j = i % 24;
    if (j < 8) atomicAdd(gws + ln * 8 + j, t);
    else if (j < 20) atomicAdd(gWr + ((j - 8) / 4) * 128 + ln * 4 + (j - 8) % 4, t);
    else if (ln == 0) { if (j < 23) atomicAdd(gbr + (j - 20), t); else atomicAdd(gbs, t); }
  }
}

// out[n] += oscale * sum_m in16[m*ld + n]   (bias gradients from the fp16 gradient chain)
__global__ void colsum16_kernel(const __half* __restrict__ in, int ld, float* out, long long M, int N,
                                int rows_per_block, const float* oscale) {
  const int n2 = blockIdx.x * blockDim.x + threadIdx.x;  // pair of columns
  if (n2 * 2 >= N) return;
  const long long m0 = (long long)blockIdx.y * rows_per_block;
  const long long m1 = min(M, m0 + rows_per_block);
  float a = 0.f, b = 0.f;
  for (long long m = m0; m < m1; ++m) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(in + m * ld + n2 * 2));
    a += f.x; b += f.y;
  }
  const float s = oscale ? *oscale : 1.0f;
  atomicAdd(out + n2 * 2, a * s);
  atomicAdd(out + n2 * 2 + 1, b * s);
}

// out (R,N) fp32 = oscale * per-ray sums of S consecutive fp16 rows
__global__ void segsum16_kernel(const __half* __restrict__ in, int ld, float* out, int R, int S, int N,
                                const float* oscale, int accumulate) {
  const int n2 = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (n2 * 2 >= N) return;
  const __half* p = in + (size_t)r * S * ld + n2 * 2;
  float a = 0.f, b = 0.f;
  for (int s = 0; s < S; ++s, p += ld) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(p));
    a += f.x; b += f.y;
  }
  const float sc = oscale ? *oscale : 1.0f;
  float* o = out + (size_t)r * N + n2 * 2;
  o[0] = accumulate ? o[0] + a * sc : a * sc;
  o[1] = accumulate ? o[1] + b * sc : b * sc;
}

// amax over |g| -> power-of-two loss scale {S, 1/S} with S * amax ~ target
__global__ void amax_kernel(const float* __restrict__ g, long long n, unsigned int* amax_bits) {
  float m = 0.f;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    // 128-bit loads, four per trip in flight (at two the 134 MB delta-logit gradient streamed at 2.9 TB/s)
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const long long n4 = n >> 2;
    long long i = tid;
    for (; i + 3 * nth < n4; i += 4 * nth) {
      const float4 a = __ldcs(g4 + i), b = __ldcs(g4 + i + nth), c = __ldcs(g4 + i + 2 * nth), d = __ldcs(g4 + i + 3 * nth);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
    }
    for (; i < n4; i += nth) {
      const float4 a = __ldg(g4 + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    }
    for (long long j = (n4 << 2) + tid; j < n; j += nth) m = fmaxf(m, fabsf(g[j]));
  } else {
    for (long long i = tid; i < n; i += nth) m = fmaxf(m, fabsf(g[i]));
  }
  // one atomic per CTA (all of them hit the same word: same-address L2 atomics serialise at ~50 ns each, which made
  // the 4736 per-warp atomics of the first version most of this kernel's 35 us)
  m = warp_max(m);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
  }
}
__global__ void scale_from_amax_kernel(const unsigned int* amax_bits, float target, float* out2) {
  const float a = __uint_as_float(*amax_bits);
  float s = 1.0f;
  if (a > 0.f && isfinite(a)) {
    int e = (int)floorf(log2f(target / a));
    e = max(-40, min(40, e));
    s = exp2f((float)e);
  }
  out2[0] = s;
  out2[1] = 1.0f / s;
}

}  // namespace moda

using namespace moda;

static void fill_win16(Win16& w, const float* win, int F) {
  for (int i = 0; i < 16; ++i) w.w[i] = (win && i < F) ? win[i] : 1.0f;
}

extern "C" int moda_pe16_fwd(const float* xyz, void* out16, void* out16lo, int ldo, long long P, int F,
                             const float* win, cudaStream_t stream) {
  if (P == 0) return 0;
  MODA_REQUIRE(xyz && out16 && F >= 0 && 3 + 6 * F <= 64 && ldo >= 64 && ldo % 8 == 0, "pe16_fwd: bad arguments");
  Win16 w; fill_win16(w, win, F);
  pe16_fwd_kernel<<<cdiv(P, 128), 128, 0, stream>>>(xyz, reinterpret_cast<__half*>(out16),
                                                   reinterpret_cast<__half*>(out16lo), P, F, ldo, w);
  return check_launch("pe16_fwd");
}

extern "C" int moda_pe16_bwd(const float* xyz, const void* g16, const void* g16lo, int ldg, float* gxyz, long long P,
                             int F, const float* win, const float* inv_scale, int accumulate, cudaStream_t stream) {
  if (P == 0) return 0;
  MODA_REQUIRE(xyz && g16 && gxyz && 3 + 6 * F <= 64 && ldg >= 64 && ldg % 8 == 0, "pe16_bwd: bad arguments");
  Win16 w; fill_win16(w, win, F);
  pe16_bwd_kernel<<<cdiv(P, 128), 128, 0, stream>>>(xyz, reinterpret_cast<const __half*>(g16),
                                                   reinterpret_cast<const __half*>(g16lo), ldg, gxyz, P, F, w,
                                                   inv_scale, accumulate);
  return check_launch("pe16_bwd");
}

extern "C" int moda_pack16(const float* in, int ld_in, int rows, int cols, int col0, void* out16, void* out16_dup,
                           void* out16_lo, int ld_out, int out_rows, int width, int transpose, cudaStream_t stream) {
  MODA_REQUIRE(in && out16 && rows > 0 && cols > 0 && ld_out >= width && width > 0 && out_rows > 0,
               "pack16: bad arguments");
  MODA_REQUIRE(transpose ? (width >= rows && out_rows >= cols) : (width >= cols && out_rows >= rows),
               "pack16: output block too small");
  pack16_kernel<<<cdiv((long long)out_rows * width, 256), 256, 0, stream>>>(
      in, ld_in, rows, cols, col0, reinterpret_cast<__half*>(out16), reinterpret_cast<__half*>(out16_dup),
      reinterpret_cast<__half*>(out16_lo), ld_out, out_rows, width, transpose);
  return check_launch("pack16");
}

extern "C" int moda_pack16_multi(int n, const float* const* in, const int* ld_in, const int* rows, const int* cols,
                                 const int* col0, void* const* out16, void* const* out16_lo, int ld_out,
                                 const int* out_rows, const int* width, const int* transpose, cudaStream_t stream) {
  MODA_REQUIRE(n >= 0 && in && ld_in && rows && cols && col0 && out16 && out16_lo && out_rows && width && transpose,
               "pack16_multi: bad arguments");
  for (int i0 = 0; i0 < n; i0 += MAX_PACK_JOBS) {
    PackJobs js;
    memset(&js, 0, sizeof(js));
    js.ld_out = ld_out;
    const int m = (n - i0 < MAX_PACK_JOBS) ? n - i0 : MAX_PACK_JOBS;
    int biggest = 0;
    for (int k = 0; k < m; ++k) {
      const int i = i0 + k;
      MODA_REQUIRE(in[i] && out16[i] && rows[i] > 0 && cols[i] > 0 && ld_out >= width[i] && width[i] > 0 && out_rows[i] > 0,
                   "pack16_multi: bad job %d", i);
      MODA_REQUIRE(transpose[i] ? (width[i] >= rows[i] && out_rows[i] >= cols[i]) : (width[i] >= cols[i] && out_rows[i] >= rows[i]),
                   "pack16_multi: output block of job %d too small", i);
      PackJob& j = js.j[k];
      j.in = in[i]; j.out = reinterpret_cast<__half*>(out16[i]); j.out_lo = reinterpret_cast<__half*>(out16_lo[i]);
      j.ld_in = ld_in[i]; j.rows = rows[i]; j.cols = cols[i]; j.col0 = col0[i]; j.out_rows = out_rows[i];
      j.width = width[i]; j.transpose = transpose[i];
      if (out_rows[i] * width[i] > biggest) biggest = out_rows[i] * width[i];
    }
    pack16_multi_kernel<<<dim3(cdiv(biggest, 256), m), 256, 0, stream>>>(js);
  }
  return check_launch("pack16_multi");
}

extern "C" int moda_split16(const float* in, int ld_in, int cols, const float* scale, void* hi, void* lo, int ld_out,
                            int width, long long M, cudaStream_t stream) {
  if (M == 0) return 0;
  MODA_REQUIRE(in && hi && width >= cols && ld_out >= width, "split16: bad arguments");
  split16_kernel<<<cdiv(M * width, 256), 256, 0, stream>>>(in, ld_in, cols, scale, reinterpret_cast<__half*>(hi),
                                                         reinterpret_cast<__half*>(lo), ld_out, width, M);
  return check_launch("split16");
}

extern "C" int moda_head_fwd(const void* H8, const void* Dfe, const float* ws, const float* bs, const float* Wr,
                             const float* br, float* raw, long long P, cudaStream_t stream) {
  if (P == 0) return 0;
  MODA_REQUIRE(H8 && Dfe && ws && bs && Wr && br && raw, "head_fwd: null pointer");
  const int blocks = (int)min((long long)148 * 8, (P + 7) / 8);
  head_fwd_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const __half*>(H8), reinterpret_cast<const __half*>(Dfe),
                                             ws, bs, Wr, br, raw, P);
  return check_launch("head_fwd");
}

extern "C" int moda_head_bwd(const void* H8, const void* Dfe, const float* raw, const float* graw, const float* Wr,
                             const float* scale, void* dDfe, float* gsig, float* gWr, float* gbr, float* gws,
                             float* gbs, long long P, cudaStream_t stream) {
  if (P == 0) return 0;
  MODA_REQUIRE(H8 && Dfe && raw && graw && Wr && dDfe && gsig && gWr && gbr && gws && gbs, "head_bwd: null pointer");
  const int blocks = (int)min((long long)148 * 4, (P + 31) / 32);
  head_bwd_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const __half*>(H8), reinterpret_cast<const __half*>(Dfe),
                                             raw, graw, Wr, scale, reinterpret_cast<__half*>(dDfe), gsig, gWr, gbr,
                                             gws, gbs, P);
  return check_launch("head_bwd");
}

extern "C" int moda_colsum16(const void* in16, int ld, float* out, long long M, int N, const float* oscale,
                             cudaStream_t stream) {
  if (M == 0 || N == 0) return 0;
  MODA_REQUIRE(in16 && out && N % 2 == 0 && ld % 2 == 0, "colsum16: bad arguments");
  const int rpb = 1024;
  dim3 grid(cdiv(N / 2, 64), cdiv(M, rpb));
  colsum16_kernel<<<grid, 64, 0, stream>>>(reinterpret_cast<const __half*>(in16), ld, out, M, N, rpb, oscale);
  return check_launch("colsum16");
}

extern "C" int moda_segsum16(const void* in16, int ld, float* out, int R, int S, int N, const float* oscale,
                             int accumulate, cudaStream_t stream) {
  if (R == 0 || N == 0) return 0;
  MODA_REQUIRE(in16 && out && N % 2 == 0 && ld % 2 == 0, "segsum16: bad arguments");
  for (int r0 = 0; r0 < R; r0 += 65535) {
    const int rc = (R - r0 < 65535) ? R - r0 : 65535;
    dim3 grid(cdiv(N / 2, 64), rc);
    segsum16_kernel<<<grid, 64, 0, stream>>>(reinterpret_cast<const __half*>(in16) + (size_t)r0 * S * ld, ld,
                                            out + (size_t)r0 * N, rc, S, N, oscale, accumulate);
  }
  return check_launch("segsum16");
}

// scale2 (device, 2 floats) = {S, 1/S}; work (device, 1 uint) is scratch
extern "C" int moda_loss_scale(const float* g, long long n, float target, unsigned int* work, float* scale2,
                               cudaStream_t stream) {
  MODA_REQUIRE(g && work && scale2 && target > 0.f, "loss_scale: bad arguments");
  cudaMemsetAsync(work, 0, sizeof(unsigned int), stream);
  if (n > 0) amax_kernel<<<148 * 4, 256, 0, stream>>>(g, n, work);
  scale_from_amax_kernel<<<1, 1, 0, stream>>>(work, target, scale2);
  return check_launch("loss_scale");
}
