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
// The input vector can be formed while it is staged (xmode; the first CTA also writes it to xout for the history):
//   xmode 0: x given       1: x = xp / (xc + delta)   (b_i from the column sums c_i)
//   xmode 2: x = -xg xb / (xc + delta)                 (gc_i from gb_i, b_i, c_i in the adjoint sweep)
__global__ void __launch_bounds__(SK_THREADS, 2)
sinkhorn_pass_kernel(const float* __restrict__ K, int n, int m, const float* __restrict__ x, float* __restrict__ y,
                     float* __restrict__ z, float* __restrict__ w, int mode, float p, float delta,
                     const float* __restrict__ u, const float* __restrict__ v, int xmode, const float* __restrict__ xc,
                     const float* __restrict__ xg, const float* __restrict__ xb, float xp, float* __restrict__ xout) {
  extern __shared__ __align__(16) float sk_smem[];
  const int m4 = m >> 2;                         // m % 4 == 0 (checked by the host)
  float4* xs = reinterpret_cast<float4*>(sk_smem);                    // [m4]
  __shared__ float red[SK_RB][SK_THREADS / 32];
  __shared__ float zrow[SK_RB];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  for (int i = tid; i < m4; i += SK_THREADS) {
    float4 xv;
    if (xmode == 0) {
      xv = __ldg(reinterpret_cast<const float4*>(x) + i);
    } else {
      const float4 c = __ldg(reinterpret_cast<const float4*>(xc) + i);
      if (xmode == 1) {
        xv = make_float4(xp / (c.x + delta), xp / (c.y + delta), xp / (c.z + delta), xp / (c.w + delta));
      } else {
        const float4 g = __ldg(reinterpret_cast<const float4*>(xg) + i), b = __ldg(reinterpret_cast<const float4*>(xb) + i);
        xv = make_float4(-g.x * b.x / (c.x + delta), -g.y * b.y / (c.y + delta), -g.z * b.z / (c.z + delta),
                         -g.w * b.w / (c.w + delta));
      }
      if (blockIdx.x == 0 && xout) reinterpret_cast<float4*>(xout)[i] = xv;
    }
    xs[i] = xv;
  }
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

// ------------------------------------------------------------------------------------------------------------------
// The matrix itself: K[r][j] = exp((F_r . V_j - 1) / eps) from the D = 16 channel unit features of the rays (F, n x 16)
// and of the lattice (V, m x 16), written once, with the first column sums of the iteration (c_1 = K^T a_0, a_0 = 1/n)
// taken on the way -- instead of a GEMM, three elementwise passes over 262 MB and a matrix-vector product.
// Thread = two columns (their V rows in registers), CTA = 512 columns x a group of rows whose features are broadcast
// from shared memory; bound by the 262 MB write.
constexpr int SKM_D = 16, SKM_RB = 64;
__global__ void __launch_bounds__(256) sinkhorn_matrix_kernel(const float* __restrict__ F, const float* __restrict__ V,
                                                              int n, int m, float inv_eps, float cs_scale,
                                                              float* __restrict__ K, float* __restrict__ c_out) {
  __shared__ float4 Fs[SKM_RB][SKM_D / 4];
  // two columns per thread (j and j + 256): the broadcast reads of a feature row from shared memory, which bound the
  // one-column version (4 x 128-bit loads per 16 FMAs: 101 us), serve both
  const int j = blockIdx.x * 512 + threadIdx.x;
  const int r0 = blockIdx.y * SKM_RB;
  const int nr = min(SKM_RB, n - r0);
  for (int i = threadIdx.x; i < nr * (SKM_D / 4); i += 256)
    Fs[i / (SKM_D / 4)][i % (SKM_D / 4)] = __ldg(reinterpret_cast<const float4*>(F + (size_t)r0 * SKM_D) + i);
  __syncthreads();
  if (j >= m) return;
  const bool two = j + 256 < m;
  float4 v[SKM_D / 4], u[SKM_D / 4];
#pragma unroll
  for (int q = 0; q < SKM_D / 4; ++q) {
    v[q] = __ldg(reinterpret_cast<const float4*>(V + (size_t)j * SKM_D) + q);
    u[q] = __ldg(reinterpret_cast<const float4*>(V + (size_t)(two ? j + 256 : j) * SKM_D) + q);
  }
  float cs = 0.f, ct = 0.f;
  float* out = K + (size_t)r0 * m + j;
#pragma unroll 4
  for (int r = 0; r < nr; ++r) {
    float d = 0.f, e = 0.f;
#pragma unroll
    for (int q = 0; q < SKM_D / 4; ++q) {
      const float4 f = Fs[r][q];
      d = fmaf(f.x, v[q].x, fmaf(f.y, v[q].y, fmaf(f.z, v[q].z, fmaf(f.w, v[q].w, d))));
      e = fmaf(f.x, u[q].x, fmaf(f.y, u[q].y, fmaf(f.z, u[q].z, fmaf(f.w, u[q].w, e))));
    }
    const float k = expf((d - 1.0f) * inv_eps), l = expf((e - 1.0f) * inv_eps);
    __stcs(out + (size_t)r * m, k);
    if (two) __stcs(out + (size_t)r * m + 256, l);
    cs += k;
    ct += l;
  }
  if (c_out) {
    atomicAdd(c_out + j, cs * cs_scale);
    if (two) atomicAdd(c_out + j + 256, ct * cs_scale);
  }
}

// Row products with FOUR vectors at once: out[r][c] = sum_j K[r][j] X[j][c] (X (m, 4) as float4 in shared memory), one
// warp per row.  Serves the soft-argmax (X = b (x) [Qx Qy Qz 1]: matched point numerator and the row sum).
__global__ void __launch_bounds__(512) sinkhorn_rows4_kernel(const float* __restrict__ K, int n, int m,
                                                             const float4* __restrict__ X, float4* __restrict__ out) {
  extern __shared__ __align__(16) float sk_smem[];
  float* xs = sk_smem;   // four planes [c][m]: a lane reads 16 consecutive bytes of each, conflict-free
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const float4 x = __ldg(X + i);
    xs[i] = x.x; xs[m + i] = x.y; xs[2 * m + i] = x.z; xs[3 * m + i] = x.w;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int m4 = m >> 2;
  const float4* x0 = reinterpret_cast<const float4*>(xs);
  const float4* x1 = x0 + m4;
  const float4* x2 = x1 + m4;
  const float4* x3 = x2 + m4;
  // two rows per warp and trip: the four vector loads from shared memory (4 x the bytes of the matrix row they multiply)
  // serve both
  for (int r = 2 * (blockIdx.x * wpb + (threadIdx.x >> 5)); r < n; r += 2 * gridDim.x * wpb) {
    const float4* row0 = reinterpret_cast<const float4*>(K + (size_t)r * m);
    const float4* row1 = reinterpret_cast<const float4*>(K + (size_t)(r + 1 < n ? r + 1 : r) * m);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll 4
    for (int i = lane; i < m4; i += 32) {
      const float4 k = __ldcs(row0 + i), l = __ldcs(row1 + i);
      const float4 p = x0[i], q = x1[i], u = x2[i], v = x3[i];
      a.x = fmaf(k.x, p.x, fmaf(k.y, p.y, fmaf(k.z, p.z, fmaf(k.w, p.w, a.x))));
      a.y = fmaf(k.x, q.x, fmaf(k.y, q.y, fmaf(k.z, q.z, fmaf(k.w, q.w, a.y))));
      a.z = fmaf(k.x, u.x, fmaf(k.y, u.y, fmaf(k.z, u.z, fmaf(k.w, u.w, a.z))));
      a.w = fmaf(k.x, v.x, fmaf(k.y, v.y, fmaf(k.z, v.z, fmaf(k.w, v.w, a.w))));
      b.x = fmaf(l.x, p.x, fmaf(l.y, p.y, fmaf(l.z, p.z, fmaf(l.w, p.w, b.x))));
      b.y = fmaf(l.x, q.x, fmaf(l.y, q.y, fmaf(l.z, q.z, fmaf(l.w, q.w, b.y))));
      b.z = fmaf(l.x, u.x, fmaf(l.y, u.y, fmaf(l.z, u.z, fmaf(l.w, u.w, b.z))));
      b.w = fmaf(l.x, v.x, fmaf(l.y, v.y, fmaf(l.z, v.z, fmaf(l.w, v.w, b.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
      b.x += __shfl_xor_sync(0xffffffffu, b.x, o); b.y += __shfl_xor_sync(0xffffffffu, b.y, o);
      b.z += __shfl_xor_sync(0xffffffffu, b.z, o); b.w += __shfl_xor_sync(0xffffffffu, b.w, o);
    }
    if (lane == 0) {
      out[r] = a;
      if (r + 1 < n) out[r + 1] = b;
    }
  }
}

// Column products with four row-weight vectors at once: out[j][c] += sum_r K[r][j] Wt[r][c] (Wt (n, 4)); thread = four
// columns (one 128-bit load per row), CTA = 1024 columns x a group of rows.  Serves the direct term of the adjoint:
// gb_j = sum_r K[r][j] (alpha_r . Q_j - beta_r).
__global__ void __launch_bounds__(256) sinkhorn_cols4_kernel(const float* __restrict__ K, int n, int m, int rows_per_cta,
                                                             const float4* __restrict__ Wt, float4* __restrict__ out) {
  const int j4 = blockIdx.x * 256 + threadIdx.x;      // float4 column index
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  if (j4 >= (m >> 2)) return;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;   // accumulators of columns 4 j4 .. 4 j4 + 3
  const float4* col = reinterpret_cast<const float4*>(K) + j4;
  const size_t pitch = (size_t)(m >> 2);
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const float4 k = __ldcs(col + (size_t)r * pitch);
    const float4 w = __ldg(Wt + r);
    a0.x = fmaf(k.x, w.x, a0.x); a0.y = fmaf(k.x, w.y, a0.y); a0.z = fmaf(k.x, w.z, a0.z); a0.w = fmaf(k.x, w.w, a0.w);
    a1.x = fmaf(k.y, w.x, a1.x); a1.y = fmaf(k.y, w.y, a1.y); a1.z = fmaf(k.y, w.z, a1.z); a1.w = fmaf(k.y, w.w, a1.w);
    a2.x = fmaf(k.z, w.x, a2.x); a2.y = fmaf(k.z, w.y, a2.y); a2.z = fmaf(k.z, w.z, a2.z); a2.w = fmaf(k.z, w.w, a2.w);
    a3.x = fmaf(k.w, w.x, a3.x); a3.y = fmaf(k.w, w.y, a3.y); a3.z = fmaf(k.w, w.z, a3.z); a3.w = fmaf(k.w, w.w, a3.w);
  }
  float* o = reinterpret_cast<float*>(out + 4 * (size_t)j4);
  const float4 acc[4] = {a0, a1, a2, a3};
#pragma unroll
  for (int c = 0; c < 4; ++c)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * c), "f"(acc[c].x), "f"(acc[c].y), "f"(acc[c].z),
                 "f"(acc[c].w)
                 : "memory");
}

// The adjoint's last step without ever forming the (n x m) cost gradient: with the low-rank factors L (n, R), Rm (R, m)
// of dL/dK (direct term + the 39 outer products of the reverse sweep),
//   gcost[r][j] = K[r][j] / eps * sum_k L[r][k] Rm[k][j],     gF (n, 16) += gcost V,     gV (m, 16) += gcost^T F.
// Thread = column (its Rm column, V row and gV accumulators in registers); a CTA walks a group of rows for 256 columns,
// SKG_RB rows at a time: L and F rows broadcast from shared memory, the gcost tile parked in shared memory, then the
// threads regroup (row, channel) for the gF sums over the tile's columns.
constexpr int SKG_RB = 16, SKG_MAXR = 48;
template <int R>
__global__ void __launch_bounds__(256, 2) sinkhorn_gcost_kernel(const float* __restrict__ K, int n, int m, int rows_per_cta,
                                                                const float* __restrict__ L, const float* __restrict__ Rm,
                                                                const float* __restrict__ F, const float* __restrict__ V,
                                                                float inv_eps, float* __restrict__ gF, float* __restrict__ gV) {
  static_assert(SKM_D == 16 && R % 4 == 0 && R <= SKG_MAXR, "rank: a multiple of 4");
  __shared__ float4 Ls[SKG_RB][R / 4];
  __shared__ float4 Fs[SKG_RB][SKM_D / 4];
  // the matrix tile of the NEXT row block is fetched with cp.async while this one is multiplied (every thread copies and
  // later reads its own column only: no barrier needed, and no registers held across the ~1 us of the loads; with plain
  // loads two rows per thread were in flight and the kernel ran at the load latency, 1.4 TB/s)
  extern __shared__ __align__(16) float kbuf[];       // [2][SKG_RB][256]
  __shared__ __align__(16) float gcs[SKG_RB][260];    // row pitch 260: 128-bit reads, the two rows of a warp on different banks
  __shared__ __align__(16) float VsT[SKM_D][260];     // lattice features of the tile, channel-major
  const int tid = threadIdx.x;
  const int j = blockIdx.x * 256 + tid;
  const bool live = j < m;
  float rm[R];
#pragma unroll
  for (int k = 0; k < R; ++k) rm[k] = live ? __ldg(Rm + (size_t)k * m + j) : 0.f;
  {
    float4 v[SKM_D / 4];
#pragma unroll
    for (int q = 0; q < SKM_D / 4; ++q)
      v[q] = live ? __ldg(reinterpret_cast<const float4*>(V + (size_t)j * SKM_D) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < SKM_D / 4; ++q) {
      VsT[4 * q][tid] = v[q].x; VsT[4 * q + 1][tid] = v[q].y; VsT[4 * q + 2][tid] = v[q].z; VsT[4 * q + 3][tid] = v[q].w;
    }
  }
  float gv[SKM_D];
#pragma unroll
  for (int c = 0; c < SKM_D; ++c) gv[c] = 0.f;
  const int g_row = tid >> 4, g_c = tid & 15;         // second phase: this thread's row of the block and channel
  const int r_begin = blockIdx.y * rows_per_cta, r_end = min(n, r_begin + rows_per_cta);
  auto prefetch = [&](int buf, int r0) {
    if (live && r0 < r_end) {
      const int nr = min(SKG_RB, r_end - r0);
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(kbuf + (size_t)buf * SKG_RB * 256 + tid);
      const float* src = K + (size_t)r0 * m + j;
      for (int r = 0; r < nr; ++r)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)(r * 256 * 4)), "l"(src + (size_t)r * m)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0, r_begin);
  int buf = 0;
  for (int r0 = r_begin; r0 < r_end; r0 += SKG_RB, buf ^= 1) {
    const int nr = min(SKG_RB, r_end - r0);
    prefetch(buf ^ 1, r0 + SKG_RB);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    const float* kcur = kbuf + (size_t)buf * SKG_RB * 256 + tid;
    __syncthreads();   // the previous block's tile and rows have been consumed
    for (int i = tid; i < nr * (R / 4); i += 256)
      Ls[i / (R / 4)][i % (R / 4)] = __ldg(reinterpret_cast<const float4*>(L + (size_t)r0 * R) + i);
    for (int i = tid; i < nr * (SKM_D / 4); i += 256)
      Fs[i / (SKM_D / 4)][i % (SKM_D / 4)] = __ldg(reinterpret_cast<const float4*>(F + (size_t)r0 * SKM_D) + i);
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < nr; ++r) {
      const float k = live ? kcur[r * 256] : 0.f;
      float w0 = 0.f, w1 = 0.f;
#pragma unroll
      for (int q = 0; q < R / 4; ++q) {
        const float4 l = Ls[r][q];
        w0 = fmaf(l.x, rm[4 * q], fmaf(l.y, rm[4 * q + 1], w0));
        w1 = fmaf(l.z, rm[4 * q + 2], fmaf(l.w, rm[4 * q + 3], w1));
      }
      const float gc = k * (w0 + w1) * inv_eps;
      gcs[r][tid] = gc;
#pragma unroll
      for (int q = 0; q < SKM_D / 4; ++q) {
        const float4 f = Fs[r][q];
        gv[4 * q] = fmaf(gc, f.x, gv[4 * q]); gv[4 * q + 1] = fmaf(gc, f.y, gv[4 * q + 1]);
        gv[4 * q + 2] = fmaf(gc, f.z, gv[4 * q + 2]); gv[4 * q + 3] = fmaf(gc, f.w, gv[4 * q + 3]);
      }
    }
    __syncthreads();
    if (g_row < nr) {
      // four columns per pair of 128-bit loads (one-float reads kept the load pipe busier than the first phase's FMAs)
      const float4* gp = reinterpret_cast<const float4*>(&gcs[g_row][0]);
      const float4* vp = reinterpret_cast<const float4*>(&VsT[g_c][0]);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int c = 0; c < 64; ++c) {
        const float4 gq = gp[c], vq = vp[c];
        s0 = fmaf(gq.x, vq.x, fmaf(gq.y, vq.y, s0));
        s1 = fmaf(gq.z, vq.z, fmaf(gq.w, vq.w, s1));
      }
      atomicAdd(gF + (size_t)(r0 + g_row) * SKM_D + g_c, s0 + s1);
    }
  }
  if (live) {
#pragma unroll
    for (int q = 0; q < SKM_D / 4; ++q)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gV + (size_t)j * SKM_D + 4 * q), "f"(gv[4 * q]),
                   "f"(gv[4 * q + 1]), "f"(gv[4 * q + 2]), "f"(gv[4 * q + 3])
                   : "memory");
  }
}

}  // namespace moda

// K (n, m) row-major fp32, m % 4 == 0, m <= 8192, 16-byte aligned rows; x (m); y, z (n) may be NULL; w (m) must be zeroed by
// the caller (it is accumulated with atomics), NULL = row sums only.  mode 0: z = p / (y + delta); mode 1: z = -y u / (v +
// delta) with u, v (n).  Replaces the pair torch.mv(K, x) / torch.mv(K.t(), z) of one Sinkhorn iteration.
// xmode 0: x (m) is the input vector.  xmode 1: x = xp / (xc + delta) is formed from the column sums xc (m); xmode 2: x =
// -xg xb / (xc + delta) (the adjoint's gc from gb, b, c); in both the vector is also written to xout (m) when non-null.
extern "C" int moda_sinkhorn_pass(const float* K, int n, int m, const float* x, float* y, float* z, float* w, int mode,
                                  float p, float delta, const float* u, const float* v, int xmode, const float* xc,
                                  const float* xg, const float* xb, float xp, float* xout, cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(K && n >= 0 && m > 0 && m % 4 == 0 && m <= SK_MAXV * 4 * SK_THREADS,
               "sinkhorn_pass: bad arguments (m %% 4 == 0, m <= %d)", SK_MAXV * 4 * SK_THREADS);
  MODA_REQUIRE((xmode == 0 && x) || (xmode == 1 && xc) || (xmode == 2 && xc && xg && xb), "sinkhorn_pass: input vector missing");
  MODA_REQUIRE(((reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) |
                 reinterpret_cast<uintptr_t>(xc) | reinterpret_cast<uintptr_t>(xg) | reinterpret_cast<uintptr_t>(xb) |
                 reinterpret_cast<uintptr_t>(xout)) & 15) == 0,
               "sinkhorn_pass: K and the column vectors must be 16-byte aligned");
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
  sinkhorn_pass_kernel<<<grid, SK_THREADS, smem, stream>>>(K, n, m, x, y, z, w, mode, p, delta, u, v, xmode, xc, xg, xb, xp,
                                                           xout);
  return check_launch("sinkhorn_pass");
}

// K (n, m) = exp((F V^T - 1) / eps) from F (n, 16), V (m, 16); c_out (m, zeroed by the caller, may be NULL) += cs_scale x the
// column sums of K.
extern "C" int moda_sinkhorn_matrix(const float* F, const float* V, int n, int m, int d, float eps, float cs_scale, float* K,
                                    float* c_out, cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(F && V && K && n >= 0 && m > 0 && d == SKM_D && eps > 0.f, "sinkhorn_matrix: bad arguments (d must be %d)", SKM_D);
  MODA_REQUIRE(((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(V)) & 15) == 0,
               "sinkhorn_matrix: F and V must be 16-byte aligned");
  if (n == 0) return 0;
  dim3 grid(cdiv(m, 512), cdiv(n, SKM_RB));
  sinkhorn_matrix_kernel<<<grid, 256, 0, stream>>>(F, V, n, m, 1.0f / eps, cs_scale, K, c_out);
  return check_launch("sinkhorn_matrix");
}

// out (n, 4) = K X with X (m, 4); m % 4 == 0.
extern "C" int moda_sinkhorn_rows4(const float* K, int n, int m, const float* X, float* out, cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(K && X && out && n >= 0 && m > 0 && m % 4 == 0 && (size_t)m * 16 <= 200 * 1024, "sinkhorn_rows4: bad arguments");
  MODA_REQUIRE(((reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "sinkhorn_rows4: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  const size_t smem = (size_t)m * 16;
  cudaFuncSetAttribute(sinkhorn_rows4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = cdiv(n, 32) < sms ? cdiv(n, 32) : sms;
  sinkhorn_rows4_kernel<<<grid, 512, smem, stream>>>(K, n, m, reinterpret_cast<const float4*>(X), reinterpret_cast<float4*>(out));
  return check_launch("sinkhorn_rows4");
}

// out (m, 4, zeroed by the caller) += K^T Wt with Wt (n, 4); m % 4 == 0.
extern "C" int moda_sinkhorn_cols4(const float* K, int n, int m, const float* Wt, float* out, cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(K && Wt && out && n >= 0 && m > 0 && m % 4 == 0, "sinkhorn_cols4: bad arguments");
  MODA_REQUIRE(((reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(Wt) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "sinkhorn_cols4: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  const int gx = cdiv(m >> 2, 256);
  int groups = cdiv(148 * 4, gx);
  if (groups > n) groups = n;
  const int rows = cdiv(n, groups);
  dim3 grid(gx, cdiv(n, rows));
  sinkhorn_cols4_kernel<<<grid, 256, 0, stream>>>(K, n, m, rows, reinterpret_cast<const float4*>(Wt), reinterpret_cast<float4*>(out));
  return check_launch("sinkhorn_cols4");
}

// gF (n, 16) += gcost V and gV (m, 16) += gcost^T F with gcost = K / eps * (L Rm), L (n, R), Rm (R, m), R = 44 (zero-pad).
extern "C" int moda_sinkhorn_gcost(const float* K, int n, int m, const float* L, const float* Rm, int R, const float* F,
                                   const float* V, int d, float eps, float* gF, float* gV, cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(K && L && Rm && F && V && gF && gV && n >= 0 && m > 0 && d == SKM_D && R == 44 && eps > 0.f,
               "sinkhorn_gcost: bad arguments (d = %d, R = 44)", SKM_D);
  MODA_REQUIRE(((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(V) |
                 reinterpret_cast<uintptr_t>(gV)) & 15) == 0, "sinkhorn_gcost: L, F, V, gV must be 16-byte aligned");
  if (n == 0) return 0;
  const int gx = cdiv(m, 256);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int groups = (2 * sms) / gx;          // one wave of two CTAs per SM
  if (groups < 1) groups = 1;
  if (groups > cdiv(n, SKG_RB)) groups = cdiv(n, SKG_RB);
  int rows = cdiv(cdiv(n, groups), SKG_RB) * SKG_RB;
  dim3 grid(gx, cdiv(n, rows));
  const size_t smem = 2 * SKG_RB * 256 * sizeof(float);
  cudaFuncSetAttribute(sinkhorn_gcost_kernel<44>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sinkhorn_gcost_kernel<44><<<grid, 256, smem, stream>>>(K, n, m, rows, L, Rm, F, V, 1.0f / eps, gF, gV);
  return check_launch("sinkhorn_gcost");
}
