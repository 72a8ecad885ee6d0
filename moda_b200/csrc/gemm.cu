// fp32 SIMT linear layers for the NeRF MLPs (nnutils/nerf.py:147-198) with the operand assembly of
// evaluate_mlp (nnutils/geom_utils.py:19-57) folded into the tile loaders: the "A" operand is a virtual
// concatenation of up to three column segments, each of which is either a dense matrix, a per-ray vector
// broadcast over the ray's samples (code.repeat(1,nbins,1), geom_utils.py:43) or the positional encoding
// of a point computed on the fly (Embedding.forward, nerf.py:56-73) -- nothing is materialised.
//
// One generic kernel computes C[i,j] = sum_r P(i,r) * Q(j,r) for three roles:
//   forward  Y[m,n]  = act(sum_k A(m,k) W[n,k] + b[n])
//   dgrad    dA[m,k] = sum_n dY[m,n] W[n,k0+k]        (optionally masked by relu, or accumulated)
//   wgrad    dW[n,k0+k] += sum_m dY[m,n] A(m,k)       (split over m, fp32 atomics)
// This is the full-precision path (used for nerf_skin, whose gradients are ill-conditioned through the
// peaked skinning softmax, SURVEY.md section 7 "hard parts") and the exact mode of the 8x256 trunk; the
// tensor-core path for the trunk lives in tc_gemm.cu.
#include "common.cuh"

namespace moda {

constexpr int SEG_DENSE = 0, SEG_BCAST = 1, SEG_PE = 2;
constexpr int MAX_SEG = 3;
constexpr int MAX_FREQS = 16;

struct Seg {
  const float* p;
  int ld;     // row stride of p (for PE: stride of the point array, >= C)
  int k0;     // first column of this segment in the virtual matrix
  int k1;     // one past the last column
  int type;   // SEG_*
  int rep;    // BCAST: rows per source row (samples per ray)
  int C;      // PE: input channels
};

struct ASrc {
  Seg s[MAX_SEG];
  int nseg;
  float win[MAX_FREQS];  // PE window weights (nerf.py:63-66)
  int M;                 // rows
  int K;                 // total columns
  __device__ __forceinline__ float at(int m, int k) const {
    if (m >= M || k >= K) return 0.f;
#pragma unroll
    for (int i = 0; i < MAX_SEG; ++i) {
      if (i < nseg && k < s[i].k1) {
        const Seg& g = s[i];
        const int j = k - g.k0;
        if (g.type == SEG_DENSE) return g.p[(size_t)m * g.ld + j];
        if (g.type == SEG_BCAST) return g.p[(size_t)(m / g.rep) * g.ld + j];
        if (j < g.C) return g.p[(size_t)m * g.ld + j];
        const int jj = j - g.C;
        const int band = jj / (2 * g.C);
        const int rem = jj - band * 2 * g.C;
        const int c = rem % g.C;
        const float v = g.p[(size_t)m * g.ld + c] * (float)(1 << band);
        return win[band] * ((rem >= g.C) ? cosf(v) : sinf(v));
      }
    }
    return 0.f;
  }
};

struct Dense {  // element (i, r) = p[i*ld + r]
  const float* p;
  int ld, I, Rr;
  __device__ __forceinline__ float at(int i, int r) const {
    return (i < I && r < Rr) ? p[(size_t)i * ld + r] : 0.f;
  }
};

struct DenseT {  // element (i, r) = p[r*ld + i]
  const float* p;
  int ld, I, Rr;
  __device__ __forceinline__ float at(int i, int r) const {
    return (i < I && r < Rr) ? p[(size_t)r * ld + i] : 0.f;
  }
};

struct ASrcT {  // element (i=k, r=m) of the virtual A
  ASrc a;
  __device__ __forceinline__ float at(int i, int r) const { return a.at(r, i); }
};

constexpr int ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2;

struct EpiFwd {
  float* y;
  int ldy;
  const float* bias;  // (N) or null
  int act;
  __device__ __forceinline__ void operator()(int i, int j, float v) const {
    if (bias) v += bias[j];
    if (act == ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == ACT_SIGMOID) v = 1.0f / (1.0f + expf(-v));
    y[(size_t)i * ldy + j] = v;
  }
};

struct EpiDgrad {
  float* c;
  int ldc;
  const float* mask;  // post-relu activation of the producing layer, or null
  int ldm;
  int accumulate;
  __device__ __forceinline__ void operator()(int i, int j, float v) const {
    float* o = c + (size_t)i * ldc + j;
    if (accumulate) v += *o;
    if (mask && !(mask[(size_t)i * ldm + j] > 0.f)) v = 0.f;  // relu' applied to the accumulated total
    *o = v;
  }
};

struct EpiWgrad {
  float* w;
  int ldw;
  __device__ __forceinline__ void operator()(int i, int j, float v) const {
    if (v != 0.f) atomicAdd(w + (size_t)i * ldw + j, v);
  }
};

constexpr int BR = 16;
constexpr int GEMM_THREADS = 256;

// P_ROW: P's memory is contiguous along r (each thread loads consecutive r of one i); otherwise along i.
template <class PL, class QL, class EP, int BI, int BJ, bool P_ROW, bool Q_ROW>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_kernel(PL P, QL Q, EP ep, int I, int J, int Rdim,
                                                            int r_chunk) {
  constexpr int TI = BI / 16, TJ = BJ / 16;
  constexpr int PE_ = BI * BR / GEMM_THREADS, QE_ = BJ * BR / GEMM_THREADS;
  __shared__ __align__(16) float Ps[BR][BI];
  __shared__ __align__(16) float Qs[BR][BJ];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * BI, j0 = blockIdx.y * BJ;
  const int r_begin = blockIdx.z * r_chunk;
  const int r_end = min(Rdim, r_begin + r_chunk);
  const int ty = tid / 16, tx = tid % 16;
  float acc[TI][TJ];
#pragma unroll
  for (int a = 0; a < TI; ++a)
#pragma unroll
    for (int b = 0; b < TJ; ++b) acc[a][b] = 0.f;
  float pv[PE_], qv[QE_];

  auto fetch = [&](int r0) {
    if (P_ROW) {
      const int i = tid % BI, rr = (tid / BI) * PE_;
#pragma unroll
      for (int e = 0; e < PE_; ++e) {
        const int r = r0 + rr + e;
        pv[e] = (r < r_end) ? P.at(i0 + i, r) : 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < PE_; ++e) {
        const int lin = tid + e * GEMM_THREADS;
        const int i = lin % BI, r = r0 + lin / BI;
        pv[e] = (r < r_end) ? P.at(i0 + i, r) : 0.f;
      }
    }
    if (Q_ROW) {
      const int j = tid % BJ, rr = (tid / BJ) * QE_;
#pragma unroll
      for (int e = 0; e < QE_; ++e) {
        const int r = r0 + rr + e;
        qv[e] = (r < r_end) ? Q.at(j0 + j, r) : 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < QE_; ++e) {
        const int lin = tid + e * GEMM_THREADS;
        const int j = lin % BJ, r = r0 + lin / BJ;
        qv[e] = (r < r_end) ? Q.at(j0 + j, r) : 0.f;
      }
    }
  };
  auto stash = [&]() {
    if (P_ROW) {
      const int i = tid % BI, rr = (tid / BI) * PE_;
#pragma unroll
      for (int e = 0; e < PE_; ++e) Ps[rr + e][i] = pv[e];
    } else {
#pragma unroll
      for (int e = 0; e < PE_; ++e) {
        const int lin = tid + e * GEMM_THREADS;
        Ps[lin / BI][lin % BI] = pv[e];
      }
    }
    if (Q_ROW) {
      const int j = tid % BJ, rr = (tid / BJ) * QE_;
#pragma unroll
      for (int e = 0; e < QE_; ++e) Qs[rr + e][j] = qv[e];
    } else {
#pragma unroll
      for (int e = 0; e < QE_; ++e) {
        const int lin = tid + e * GEMM_THREADS;
        Qs[lin / BJ][lin % BJ] = qv[e];
      }
    }
  };

  if (r_begin < r_end) fetch(r_begin);
  for (int r0 = r_begin; r0 < r_end; r0 += BR) {
    __syncthreads();
    stash();
    __syncthreads();
    if (r0 + BR < r_end) fetch(r0 + BR);
#pragma unroll
    for (int kk = 0; kk < BR; ++kk) {
      float a[TI], b[TJ];
#pragma unroll
      for (int x = 0; x < TI; x += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Ps[kk][ty * TI + x]);
        a[x] = t.x; a[x + 1] = t.y; a[x + 2] = t.z; a[x + 3] = t.w;
      }
      if constexpr (TJ % 4 == 0) {
#pragma unroll
        for (int x = 0; x < TJ; x += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&Qs[kk][tx * TJ + x]);
          b[x] = t.x; b[x + 1] = t.y; b[x + 2] = t.z; b[x + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int x = 0; x < TJ; ++x) b[x] = Qs[kk][tx * TJ + x];
      }
#pragma unroll
      for (int x = 0; x < TI; ++x)
#pragma unroll
        for (int y = 0; y < TJ; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
  }
#pragma unroll
  for (int x = 0; x < TI; ++x) {
    const int i = i0 + ty * TI + x;
    if (i >= I) continue;
#pragma unroll
    for (int y = 0; y < TJ; ++y) {
      const int j = j0 + tx * TJ + y;
      if (j < J) ep(i, j, acc[x][y]);
    }
  }
}

// out[r, n] = sum_{s<S} in[(r*S+s)*ld + n]   (per-ray sums of a per-sample gradient).  When S is long
// (a code shared by every sample, e.g. rest_pose_code) the rows are split over blockIdx.z and combined with
// atomics into a zero-initialised output.
__global__ void segsum_kernel(const float* in, int ld, float* out, int R, int S, int N, int s_chunk) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (n >= N) return;
  const int s0 = blockIdx.z * s_chunk;
  const int s1 = min(S, s0 + s_chunk);
  const float* p = in + ((size_t)r * S + s0) * ld + n;
  float acc = 0.f;
  for (int s = s0; s < s1; ++s, p += ld) acc += *p;
  if (gridDim.z == 1) out[(size_t)r * N + n] = acc;
  else atomicAdd(out + (size_t)r * N + n, acc);
}

// out = g * act'(y) for the output activations (sigmoid: y(1-y); relu: y>0); row-strided (M,N) views
__global__ void act_bwd_kernel(int kind, const float* y, int ldy, const float* g, int ldg, float* out, int ldo,
                               long long M, int N) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * N) return;
  const long long m = t / N;
  const int n = (int)(t % N);
  const float yy = y[m * ldy + n], gg = g[m * ldg + n];
  out[m * ldo + n] = (kind == ACT_SIGMOID) ? gg * yy * (1.0f - yy) : ((yy > 0.f) ? gg : 0.f);
}

// out[n] += sum_m in[m*ld + n]
__global__ void colsum_kernel(const float* in, int ld, float* out, int M, int N, int rows_per_block) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float acc = 0.f;
  for (int m = m0; m < m1; ++m) acc += in[(size_t)m * ld + n];
  if (acc != 0.f) atomicAdd(out + n, acc);
}

}  // namespace moda

using namespace moda;

// Flat description of the virtual A operand as it crosses the C ABI:
//   seg_ptr[i], seg_ld[i], seg_width[i], seg_type[i], seg_aux[i] (BCAST: rows per source row; PE: channels)
static int build_asrc(ASrc& a, int M, int nseg, const float* const* seg_ptr, const int* seg_ld,
                      const int* seg_width, const int* seg_type, const int* seg_aux, const float* win,
                      int n_win) {
  MODA_REQUIRE(nseg >= 1 && nseg <= MAX_SEG, "linear: nseg=%d outside [1,%d]", nseg, MAX_SEG);
  MODA_REQUIRE(n_win <= MAX_FREQS, "linear: more than %d PE bands", MAX_FREQS);
  a.nseg = nseg;
  a.M = M;
  int k = 0;
  for (int i = 0; i < nseg; ++i) {
    MODA_REQUIRE(seg_ptr[i] != nullptr && seg_width[i] > 0, "linear: segment %d is empty", i);
    a.s[i].p = seg_ptr[i];
    a.s[i].ld = seg_ld[i];
    a.s[i].k0 = k;
    a.s[i].k1 = k + seg_width[i];
    a.s[i].type = seg_type[i];
    a.s[i].rep = (seg_type[i] == SEG_BCAST) ? seg_aux[i] : 1;
    a.s[i].C = (seg_type[i] == SEG_PE) ? seg_aux[i] : 1;
    if (seg_type[i] == SEG_BCAST) MODA_REQUIRE(seg_aux[i] >= 1, "linear: BCAST segment needs rep >= 1");
    if (seg_type[i] == SEG_PE) {
      MODA_REQUIRE(seg_aux[i] >= 1 && (seg_width[i] - seg_aux[i]) % (2 * seg_aux[i]) == 0,
                   "linear: PE segment width %d is not C*(1+2F)", seg_width[i]);
      MODA_REQUIRE((seg_width[i] / seg_aux[i] - 1) / 2 <= n_win, "linear: PE window has too few bands");
    }
    k = a.s[i].k1;
  }
  a.K = k;
  for (int i = 0; i < MAX_FREQS; ++i) a.win[i] = (i < n_win && win) ? win[i] : 1.0f;
  return 0;
}

// Y[M,N] = act(A(M,K) W[N,K]^T + bias)
extern "C" int moda_linear_fwd(int M, int N, int nseg, const float* const* seg_ptr, const int* seg_ld,
                               const int* seg_width, const int* seg_type, const int* seg_aux,
                               const float* win, int n_win, const float* W, int ldw, const float* bias,
                               int act, float* Y, int ldy, cudaStream_t stream) {
  ASrc a;
  if (int e = build_asrc(a, M, nseg, seg_ptr, seg_ld, seg_width, seg_type, seg_aux, win, n_win)) return e;
  MODA_REQUIRE(W && Y && ldw >= a.K && ldy >= N, "linear_fwd: bad W/Y");
  if (M == 0 || N == 0) return 0;
  Dense q{W, ldw, N, a.K};
  EpiFwd ep{Y, ldy, bias, act};
  if (N > 32 && cdiv(M, 128) * cdiv(N, 64) < 2 * 148) {
    // per-ray-sized problems (the hoisted per-ray bias layers, M = rays): 128-row tiles leave most SMs idle
    dim3 grid(cdiv(M, 32), cdiv(N, 64), 1);
    gemm_kernel<ASrc, Dense, EpiFwd, 32, 64, true, true><<<grid, GEMM_THREADS, 0, stream>>>(a, q, ep, M, N, a.K, a.K);
  } else if (N > 32) {
    dim3 grid(cdiv(M, 128), cdiv(N, 64), 1);
    gemm_kernel<ASrc, Dense, EpiFwd, 128, 64, true, true><<<grid, GEMM_THREADS, 0, stream>>>(a, q, ep, M, N, a.K, a.K);
  } else {
    dim3 grid(cdiv(M, 128), cdiv(N, 16), 1);
    gemm_kernel<ASrc, Dense, EpiFwd, 128, 16, true, true><<<grid, GEMM_THREADS, 0, stream>>>(a, q, ep, M, N, a.K, a.K);
  }
  return check_launch("linear_fwd");
}

// dA[M,Kseg] (=|+=) dY[M,N] W[N, k0:k0+Kseg], optionally masked by (mask > 0)
extern "C" int moda_linear_dgrad(int M, int N, int Kseg, const float* dY, int ldy, const float* W, int ldw,
                                 int k0, const float* mask, int ldm, int accumulate, float* dA, int lda,
                                 cudaStream_t stream) {
  MODA_REQUIRE(dY && W && dA && ldy >= N && lda >= Kseg, "linear_dgrad: bad arguments");
  if (M == 0 || Kseg == 0) return 0;
  Dense p{dY, ldy, M, N};
  DenseT q{W + k0, ldw, Kseg, N};
  EpiDgrad ep{dA, lda, mask, ldm, accumulate};
  if (Kseg > 32 && cdiv(M, 128) * cdiv(Kseg, 64) < 2 * 148) {
    dim3 grid(cdiv(M, 32), cdiv(Kseg, 64), 1);
    gemm_kernel<Dense, DenseT, EpiDgrad, 32, 64, true, false><<<grid, GEMM_THREADS, 0, stream>>>(p, q, ep, M, Kseg, N, N);
  } else if (Kseg > 32) {
    dim3 grid(cdiv(M, 128), cdiv(Kseg, 64), 1);
    gemm_kernel<Dense, DenseT, EpiDgrad, 128, 64, true, false><<<grid, GEMM_THREADS, 0, stream>>>(p, q, ep, M, Kseg, N, N);
  } else {
    dim3 grid(cdiv(M, 128), cdiv(Kseg, 16), 1);
    gemm_kernel<Dense, DenseT, EpiDgrad, 128, 16, true, false><<<grid, GEMM_THREADS, 0, stream>>>(p, q, ep, M, Kseg, N, N);
  }
  return check_launch("linear_dgrad");
}

// dW[N, k0:k0+K] += dY[M,N]^T A(M,K) ; dbias[N] += colsum(dY) when dbias != null
extern "C" int moda_linear_wgrad(int M, int N, int nseg, const float* const* seg_ptr, const int* seg_ld,
                                 const int* seg_width, const int* seg_type, const int* seg_aux,
                                 const float* win, int n_win, const float* dY, int ldy, float* dW, int ldw,
                                 int k0, float* dbias, cudaStream_t stream) {
  ASrcT a;
  if (int e = build_asrc(a.a, M, nseg, seg_ptr, seg_ld, seg_width, seg_type, seg_aux, win, n_win)) return e;
  MODA_REQUIRE(dY && dW && ldy >= N, "linear_wgrad: bad arguments");
  if (M == 0 || N == 0) return 0;
  const int K = a.a.K;
  DenseT p{dY, ldy, N, M};
  EpiWgrad ep{dW + k0, ldw};
  // split the reduction over rows so that the grid covers the machine a few times over
  const int tiles = cdiv(N, 64) * cdiv(K, 64);
  int splits = cdiv(148 * 4, tiles);
  int r_chunk = cdiv(M, splits);
  r_chunk = ((r_chunk + BR - 1) / BR) * BR;
  if (r_chunk < 64) r_chunk = 64;   // (256 left a 1024-ray problem on 16 CTAs: 17 us for 8 MFLOP)
  splits = cdiv(M, r_chunk);
  dim3 grid(cdiv(N, 64), cdiv(K, 64), splits);
  gemm_kernel<DenseT, ASrcT, EpiWgrad, 64, 64, false, false><<<grid, GEMM_THREADS, 0, stream>>>(p, a, ep, N, K, M, r_chunk);
  if (dbias) {
    // enough row blocks to cover the machine (the old fixed 2048 rows per block left a handful of CTAs walking
    // thousands of rows each)
    int rpb = cdiv(M, cdiv(148 * 4, cdiv(N, 64)));
    if (rpb < 16) rpb = 16;
    dim3 g2(cdiv(N, 64), cdiv(M, rpb));
    colsum_kernel<<<g2, 64, 0, stream>>>(dY, ldy, dbias, M, N, rpb);
  }
  return check_launch("linear_wgrad");
}

extern "C" int moda_segsum(const float* in, int ld, float* out, int R, int S, int N, cudaStream_t stream) {
  if (R == 0 || N == 0) return 0;
  MODA_REQUIRE(in && out && ld >= N, "segsum: bad arguments");
  MODA_REQUIRE(R <= 65535 * 1024, "segsum: too many rows");
  int splits = 1, s_chunk = S;
  if (S > 512) {
    s_chunk = 256;
    splits = cdiv(S, s_chunk);
    if (splits > 65535) { splits = 65535; s_chunk = cdiv(S, splits); splits = cdiv(S, s_chunk); }
    cudaMemsetAsync(out, 0, (size_t)R * N * sizeof(float), stream);
  }
  for (int r0 = 0; r0 < R; r0 += 65535) {
    const int rc = (R - r0 < 65535) ? R - r0 : 65535;
    dim3 grid(cdiv(N, 64), rc, splits);
    segsum_kernel<<<grid, 64, 0, stream>>>(in + (size_t)r0 * S * ld, ld, out + (size_t)r0 * N, rc, S, N, s_chunk);
  }
  return check_launch("segsum");
}

extern "C" int moda_act_bwd(int kind, const float* y, int ldy, const float* g, int ldg, float* out, int ldo,
                            long long M, int N, cudaStream_t stream) {
  MODA_REQUIRE((kind == ACT_RELU || kind == ACT_SIGMOID) && y && g && out, "act_bwd: bad arguments");
  if (M * N == 0) return 0;
  act_bwd_kernel<<<cdiv(M * N, 256), 256, 0, stream>>>(kind, y, ldy, g, ldg, out, ldo, M, N);
  return check_launch("act_bwd");
}

// ------------------------------------------------------------------------------ final layer folded into the direction layer
// The chain programs run xyz_encoding_final (no activation) and dir_encoding as one layer (see MODA_CHAIN_FOLD_FINAL in
// the header): W' (n, W) = Wd[:, :W] Wf, b' (n) = bd + Wd[:, :W] bf, and on the way back dWf (W, W) += Wd1^T dW', dbf (W) +=
// Wd1^T db', dWd[:, :W] += dW' Wf^T + db' bf^T.  O(weights) work (n <= 128, W <= 256), one launch each way instead of
// two / five launches of the general kernels above (which take ~10 us apiece on a 128 x 256 x 256 problem: 16 CTAs).
namespace moda {
constexpr int FOLD_THREADS = 256;
constexpr int FT = 32;   // tile edge: a block makes 32 x 32 outputs, K walked in chunks of 32 through shared memory

struct FoldArgs {
  const float *Wd, *Wf, *bf, *bd, *gWp, *dbp;
  float *Wp, *bp, *gWf, *gbf, *gWd;
  int ldwd, ldwf, ldgf, ldgd, n, W;
};

// The three small products as one tiled routine C(m, c) (+)= sum_k A(m, k) B(k, c); the operand views (transposes, the
// bias column / row riding along as one extra column or K index) are the only thing that differs:
//   P = 0 fold      m = i < n,  c = j <= W, k < W      A = Wd[i][k]   B = Wf[k][j] | bf[k]        C = Wp[i][j] | bp[i] (+ bd)
//   P = 1 unfold Wf m = k' < W, c = t <= W, k = i < n  A = Wd[i][k']  B = gWp[i][t] | dbp[i]      C = gWf[k'][t] | gbf[k']
//   P = 2 unfold Wd m = i < n,  c = t < W,  k = j <= W A = gWp[i][j] | dbp[i]   B = Wf[t][j] | bf[t]   C = gWd[i][t]
template <int P>
struct FoldView {
  static __device__ __forceinline__ int M(const FoldArgs& a) { return P == 1 ? a.W : a.n; }
  static __device__ __forceinline__ int N(const FoldArgs& a) { return P == 2 ? a.W : a.W + 1; }
  static __device__ __forceinline__ int K(const FoldArgs& a) { return P == 0 ? a.W : (P == 1 ? a.n : a.W + 1); }
  static constexpr bool A_K_CONTIG = (P != 1);   // which index of the operand is contiguous in memory (picks the loader's
  static constexpr bool B_K_CONTIG = (P == 2);   // thread mapping so that a warp reads whole cache lines)
  static __device__ __forceinline__ float A(const FoldArgs& a, int m, int k) {
    if (m >= M(a) || k >= K(a)) return 0.f;
    if (P == 0) return __ldg(a.Wd + (size_t)m * a.ldwd + k);
    if (P == 1) return __ldg(a.Wd + (size_t)k * a.ldwd + m);
    return k < a.W ? __ldg(a.gWp + (size_t)m * a.W + k) : __ldg(a.dbp + m);
  }
  static __device__ __forceinline__ float B(const FoldArgs& a, int k, int c) {
    if (c >= N(a) || k >= K(a)) return 0.f;
    if (P == 0) return c < a.W ? __ldg(a.Wf + (size_t)k * a.ldwf + c) : __ldg(a.bf + k);
    if (P == 1) return c < a.W ? __ldg(a.gWp + (size_t)k * a.W + c) : __ldg(a.dbp + k);
    return k < a.W ? __ldg(a.Wf + (size_t)c * a.ldwf + k) : __ldg(a.bf + c);
  }
  static __device__ __forceinline__ void C(const FoldArgs& a, int m, int c, float v) {
    if (m >= M(a) || c >= N(a)) return;
    if (P == 0) {
      if (c < a.W) a.Wp[(size_t)m * a.W + c] = v; else a.bp[m] = v + __ldg(a.bd + m);
    } else if (P == 1) {
      // atomic adds: the two evaluations of nerf_skin unfold into the same gradient buffers from different streams
      atomicAdd(c < a.W ? a.gWf + (size_t)m * a.ldgf + c : a.gbf + m, v);
    } else {
      atomicAdd(a.gWd + (size_t)m * a.ldgd + c, v);
    }
  }
};

template <int P>
__device__ __forceinline__ void fold_tile(const FoldArgs& a, int tm, int tn, float (*As)[FT + 1], float (*Bs)[FT + 1]) {
  using V = FoldView<P>;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;   // 8 warps: warp wp owns rows 4 wp .. 4 wp + 3, lane = column
  const int m0 = tm * FT, c0 = tn * FT, K = V::K(a);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float ra[4], rb[4];
  // element u of this thread in a 32 x 32 operand tile: (x = wp + 8 u, y = lane), y along the contiguous index
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = wp + 8 * u;
      ra[u] = V::A_K_CONTIG ? V::A(a, m0 + x, k0 + lane) : V::A(a, m0 + lane, k0 + x);
      rb[u] = V::B_K_CONTIG ? V::B(a, k0 + lane, c0 + x) : V::B(a, k0 + x, c0 + lane);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += FT) {
    __syncthreads();   // the previous chunk has been consumed
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = wp + 8 * u;
      if (V::A_K_CONTIG) As[lane][x] = ra[u]; else As[x][lane] = ra[u];      // As[k][m]
      if (V::B_K_CONTIG) Bs[lane][x] = rb[u]; else Bs[x][lane] = rb[u];      // Bs[k][c]
    }
    __syncthreads();
    if (k0 + FT < K) fetch(k0 + FT);   // the next chunk is in flight while this one is multiplied
#pragma unroll
    for (int k = 0; k < FT; ++k) {
      const float b = Bs[k][lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(As[k][4 * wp + u], b, acc[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) V::C(a, m0 + 4 * wp + u, c0 + lane, acc[u]);
}

// blockIdx.x enumerates the tiles of product P0 first, then (unfold) those of P1
template <int P0, int P1>
__global__ void __launch_bounds__(FOLD_THREADS) fold_final_kernel(FoldArgs a, int tiles_n0, int ntiles0, int tiles_n1) {
  __shared__ float As[FT][FT + 1];
  __shared__ float Bs[FT][FT + 1];
  int t = blockIdx.x;
  if (t < ntiles0) {
    fold_tile<P0>(a, t / tiles_n0, t % tiles_n0, As, Bs);
  } else {
    t -= ntiles0;
    fold_tile<P1>(a, t / tiles_n1, t % tiles_n1, As, Bs);
  }
}
}  // namespace moda

// Wp (n, W) = Wd[:, :W] Wf,  bp (n) = bd + Wd[:, :W] bf      (Wd (n, >= W) with row pitch ldwd, Wf (W, W) pitch ldwf)
extern "C" int moda_fold_final(const float* Wd, int ldwd, const float* Wf, int ldwf, const float* bf, const float* bd, int n,
                               int W, float* Wp, float* bp, cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(Wd && Wf && bf && bd && Wp && bp && n > 0 && W > 0 && ldwd >= W && ldwf >= W, "fold_final: bad arguments");
  FoldArgs a{};
  a.Wd = Wd; a.Wf = Wf; a.bf = bf; a.bd = bd; a.Wp = Wp; a.bp = bp; a.ldwd = ldwd; a.ldwf = ldwf; a.n = n; a.W = W;
  const int tm = cdiv(n, FT), tn = cdiv(W + 1, FT);
  fold_final_kernel<0, 0><<<tm * tn, FOLD_THREADS, 0, stream>>>(a, tn, tm * tn, 1);
  return check_launch("fold_final");
}

// gWf (W, W) += Wd1^T gWp,  gbf (W) += Wd1^T dbp,  gWd[:, :W] (n, W) += gWp Wf^T + dbp bf^T   (atomic adds)
extern "C" int moda_unfold_final(const float* gWp, const float* dbp, const float* Wd, int ldwd, const float* Wf, int ldwf,
                                 const float* bf, int n, int W, float* gWf, int ldgf, float* gbf, float* gWd, int ldgd,
                                 cudaStream_t stream) {
  using namespace moda;
  MODA_REQUIRE(gWp && dbp && Wd && Wf && bf && gWf && gbf && gWd && n > 0 && W > 0 && ldwd >= W && ldwf >= W && ldgf >= W &&
                   ldgd >= W, "unfold_final: bad arguments");
  FoldArgs a{};
  a.gWp = gWp; a.dbp = dbp; a.Wd = Wd; a.Wf = Wf; a.bf = bf; a.gWf = gWf; a.gbf = gbf; a.gWd = gWd;
  a.ldwd = ldwd; a.ldwf = ldwf; a.ldgf = ldgf; a.ldgd = ldgd; a.n = n; a.W = W;
  const int tn1 = cdiv(W + 1, FT), nt1 = cdiv(W, FT) * tn1;      // gWf | gbf
  const int tn2 = cdiv(W, FT), nt2 = cdiv(n, FT) * tn2;          // gWd[:, :W]
  fold_final_kernel<1, 2><<<nt1 + nt2, FOLD_THREADS, 0, stream>>>(a, tn1, nt1, tn2);
  return check_launch("unfold_final");
}
