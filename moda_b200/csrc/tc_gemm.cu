// Tensor-core (tcgen05 / TMEM / TMA) linear layers for the 8x256 trunk MLP (nnutils/nerf.py:147-198).
//
//   tc_linear : Y[M,N] = epi( [A1 | A2][M,K] * B[N,K]^T )   fp16 operands, fp32 accumulation in TMEM
//               used for the forward layers (B = W) and the data-gradient chain (B = W^T), with bias /
//               per-ray bias / ReLU / ReLU-mask / rank-1 term fused into the epilogue.
//   tc_wgrad  : dW[N,K] += dY[M,N]^T * X[M,K]                 both operands MN-major straight from the same
//               row-major activation tiles (no transposes in memory).
//
// Structure of tc_linear (one CTA per SM, persistent over 128-row tiles):
//   warp 0   : TMA producer.  Loads the whole weight matrix once (it stays resident in shared memory for the
//              life of the CTA: <= 160 KB), then streams 128x64 fp16 activation chunks through a 4-deep ring.
//   warp 1   : allocates TMEM (2 accumulators x N columns) and issues tcgen05.mma (one elected lane).
//   warps 2-5: epilogue.  tcgen05.ld the finished accumulator (32 columns at a time), apply the fused
//              epilogue and write fp16 / fp32 rows, while warp 1 already fills the other accumulator.
// All shared-memory operand tiles use the 128-byte swizzle that TMA writes and UMMA descriptors read.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace moda {
namespace tc {

constexpr int TILE_M = 128;
constexpr int CHUNK_K = 64;                        // fp16 elements per 128-byte swizzle row
constexpr int A_STAGE_BYTES = TILE_M * CHUNK_K * 2;  // 16 KB
constexpr int LIN_MAX_STAGES = 4;
constexpr int LIN_THREADS = 320;  // TMA warp + MMA warp + 8 epilogue warps

// Where the K chunks of the A operand come from: chunk kc is columns [64*col[kc], 64*col[kc]+64) of source
// src[kc].  This is how concatenated inputs ([PE | h]) and split-precision operands ([hi | lo | hi]) are fed.
constexpr int MAX_A_SRC = 4;
constexpr int MAX_CHUNKS = 8;
struct AMaps { CUtensorMap m[MAX_A_SRC]; };
struct ChunkTable {
  int n;
  unsigned char src[MAX_CHUNKS];
  unsigned char col[MAX_CHUNKS];
};

struct LinEpi {
  const float* bias;      // (N) or null
  const float* rowbias;   // (M/rep, N) or null: per-ray bias (hoisted per-ray-constant inputs)
  int rep;
  int relu;
  const __half* mask;     // (M, ldm) or null: output is zeroed where mask <= 0 (ReLU of the producing layer)
  int ldm;
  const float* rv;        // (M) or null  } rank-1 term rv[m] * cv[n] * (*rscale) added before masking
  const float* cv;        // (N)          } (the sigma head's data gradient)
  const float* rscale;    // device scalar or null (=1)
  __half* y16;            // (M, ldy16) or null
  __half* y16lo;          // (M, ldy16) or null: fp16(value - fp16(value)), the low half of a split-precision pair
  int ldy16;
  int acc16;              // y16 += result instead of =
  float* y32;             // (M, ldy32) or null, multiplied by *oscale when given
  int ldy32;
  const float* oscale;
};

constexpr int EPI_WARPS = 8;                    // two warps per TMEM lane quarter, each owning half the columns
constexpr int STG_PITCH = 80;                   // bytes per staged row: 32 fp16 (64 B) + 16 B pad (bank spread)
constexpr int STG_BYTES = 32 * STG_PITCH;       // per-warp transposition buffer

// Each epilogue thread owns one accumulator row (that is how tcgen05.ld hands out TMEM lanes), but global
// memory wants a warp to touch whole rows.  So every 32x32 chunk goes through a per-warp shared-memory
// transposition: thread-per-row on the register side, 8 rows x 64 B per instruction on the memory side.
template <int N_TILE, bool ROWBIAS, bool RANK1, bool MASK>
__global__ void __launch_bounds__(LIN_THREADS, 1)
tc_linear_kernel(const __grid_constant__ AMaps mapsA, const __grid_constant__ CUtensorMap mapB, int M,
                 ChunkTable ct, int LIN_STAGES, LinEpi ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KC = ct.n;
  constexpr int B_CHUNK_BYTES = N_TILE * CHUNK_K * 2;
  constexpr int N_EPI = (N_TILE >= 64) ? EPI_WARPS : 4;   // active epilogue warps
  constexpr int COLS_PER_WARP = N_TILE / (N_EPI / 4);
  uint8_t* sB = smem;
  uint8_t* sA = smem + (size_t)KC * B_CHUNK_BYTES;
  uint8_t* sStage = sA + LIN_STAGES * A_STAGE_BYTES;        // EPI_WARPS x STG_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + EPI_WARPS * STG_BYTES);
  uint64_t* full = bars;                        // [LIN_STAGES]
  uint64_t* empty = bars + LIN_MAX_STAGES;      // [LIN_STAGES]
  uint64_t* b_full = bars + 2 * LIN_MAX_STAGES; // [1]
  uint64_t* t_full = b_full + 1;                // [2]
  uint64_t* t_empty = t_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);  // [N_TILE]
  float* s_cv = s_bias + N_TILE;                             // [N_TILE]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (M + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    for (int i = 0; i < LIN_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], N_EPI); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < N_TILE; i += LIN_THREADS) {
    s_bias[i] = ep.bias ? ep.bias[i] : 0.f;
    s_cv[i] = ep.cv ? ep.cv[i] : 0.f;
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * N_TILE);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      prefetch_tmap(&mapsA.m[0]);
      prefetch_tmap(&mapB);
      mbar_expect_tx(b_full, (uint32_t)(KC * B_CHUNK_BYTES));
      for (int kc = 0; kc < KC; ++kc) tma_load_2d(sB + (size_t)kc * B_CHUNK_BYTES, &mapB, b_full, kc * CHUNK_K, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], A_STAGE_BYTES);
          tma_load_2d(sA + stage * A_STAGE_BYTES, &mapsA.m[ct.src[kc]], &full[stage], (int)ct.col[kc] * CHUNK_K,
                      t * TILE_M);
          if (++stage == LIN_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE_M, N_TILE, 0, 0);
      mbar_wait(b_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        mbar_wait(&t_empty[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N_TILE);
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)kc * B_CHUNK_BYTES);
#pragma unroll
          for (int k = 0; k < CHUNK_K / 16; ++k) {
            const uint64_t ad = make_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = make_desc(b_addr + k * 32, 16, 1024);
            umma_f16(d_tmem, ad, bd, idesc, (kc | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the A stage when these MMAs retire
          if (++stage == LIN_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&t_full[buf]);     // accumulator complete
      }
    }
  } else if (warp - 2 < N_EPI) {
    // ------------------------------------------------------------------ epilogue (warps 2..2+N_EPI)
    const int ew = warp - 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int col_base = (ew >> 2) * COLS_PER_WARP;
    uint8_t* stg = sStage + ew * STG_BYTES;
    const float rscale = ep.rscale ? *ep.rscale : 1.0f;
    const float oscale = ep.oscale ? *ep.oscale : 1.0f;
    const float relu_floor = ep.relu ? 0.f : -INFINITY;
    // memory-side mapping of a 32x32 fp16 chunk: lane -> (row = lane/4 + 8*i, 16-byte piece = lane%4)
    const int mrow = lane >> 2, mpc = lane & 3;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      mbar_wait(&t_full[buf], bphase);
      tc_fence_after();
      const int row0 = t * TILE_M + q * 32;       // first row of this warp's 32-row slab
      const int row = row0 + lane;
      const bool live = row < M;
      const float rvv = (RANK1 && live) ? ep.rv[row] * rscale : 0.f;
      const float* rb = (ROWBIAS && live) ? ep.rowbias + (size_t)(row / ep.rep) * N_TILE : nullptr;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * N_TILE + col_base);
      float va[32], vb[32];
      uint4 mnext[4];
      if (MASK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = mrow + 8 * i;
          mnext[i] = (row0 + r < M)
                         ? *reinterpret_cast<const uint4*>(ep.mask + (size_t)(row0 + r) * ep.ldm + col_base + mpc * 8)
                         : make_uint4(0, 0, 0, 0);
        }
      }
      tmem_ld32_issue(t_addr, va);
#pragma unroll 1
      for (int cc = 0; cc < COLS_PER_WARP; cc += 64) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (cc + half * 32 >= COLS_PER_WARP) break;
          float* v = half ? vb : va;
          float* vn = half ? va : vb;
          const int c0 = col_base + cc + half * 32;
          tmem_ld_wait();
          if (cc + half * 32 + 32 < COLS_PER_WARP) tmem_ld32_issue(t_addr + (uint32_t)(cc + half * 32 + 32), vn);
          // ---- mask chunk: coalesced global read (prefetched one chunk ahead) -> shared -> this thread's row
          if (MASK) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(stg + (mrow + 8 * i) * STG_PITCH + mpc * 16) = mnext[i];
            __syncwarp();
            if (cc + half * 32 + 32 < COLS_PER_WARP) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = mrow + 8 * i;
                mnext[i] = (row0 + r < M)
                               ? *reinterpret_cast<const uint4*>(ep.mask + (size_t)(row0 + r) * ep.ldm + c0 + 32 + mpc * 8)
                               : make_uint4(0, 0, 0, 0);
              }
            }
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + j4 * 4);
            v[j4 * 4] += b4.x; v[j4 * 4 + 1] += b4.y; v[j4 * 4 + 2] += b4.z; v[j4 * 4 + 3] += b4.w;
          }
          if (ROWBIAS) {
            if (rb) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b4 = *reinterpret_cast<const float4*>(rb + c0 + j4 * 4);
                v[j4 * 4] += b4.x; v[j4 * 4 + 1] += b4.y; v[j4 * 4 + 2] += b4.z; v[j4 * 4 + 3] += b4.w;
              }
            }
          }
          if (RANK1) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 c4 = *reinterpret_cast<const float4*>(s_cv + c0 + j4 * 4);
              v[j4 * 4] = fmaf(rvv, c4.x, v[j4 * 4]); v[j4 * 4 + 1] = fmaf(rvv, c4.y, v[j4 * 4 + 1]);
              v[j4 * 4 + 2] = fmaf(rvv, c4.z, v[j4 * 4 + 2]); v[j4 * 4 + 3] = fmaf(rvv, c4.w, v[j4 * 4 + 3]);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], relu_floor);
          if (MASK) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 mv = *reinterpret_cast<const uint4*>(stg + lane * STG_PITCH + j4 * 16);
              const __half2* mh = reinterpret_cast<const __half2*>(&mv);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 mf = __half22float2(mh[e]);
                if (!(mf.x > 0.f)) v[j4 * 8 + 2 * e] = 0.f;
                if (!(mf.y > 0.f)) v[j4 * 8 + 2 * e + 1] = 0.f;
              }
            }
            __syncwarp();
          }
          if (ep.y32 && live) {
            float4* yp = reinterpret_cast<float4*>(ep.y32 + (size_t)row * ep.ldy32 + c0);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              yp[j4] = make_float4(v[j4 * 4] * oscale, v[j4 * 4 + 1] * oscale, v[j4 * 4 + 2] * oscale,
                                   v[j4 * 4 + 3] * oscale);
          }
          if (ep.y16) {
            // hi (and optionally lo) halves: registers -> shared (thread = row) -> global (8 rows x 64 B / instr)
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
              if (pass == 1 && !ep.y16lo) break;
              __half* dst = pass ? ep.y16lo : ep.y16;
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                uint4 o;
                __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float x0 = v[j4 * 8 + 2 * e], x1 = v[j4 * 8 + 2 * e + 1];
                  const __half2 h = __floats2half2_rn(x0, x1);
                  if (pass == 0) oh[e] = h;
                  else {
                    const float2 hf = __half22float2(h);
                    oh[e] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                  }
                }
                *reinterpret_cast<uint4*>(stg + lane * STG_PITCH + j4 * 16) = o;
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = mrow + 8 * i;
                if (row0 + r < M) {
                  uint4 o = *reinterpret_cast<const uint4*>(stg + r * STG_PITCH + mpc * 16);
                  uint4* gp = reinterpret_cast<uint4*>(dst + (size_t)(row0 + r) * ep.ldy16 + c0 + mpc * 8);
                  if (ep.acc16 && pass == 0) {
                    const uint4 old = *gp;
                    const __half2* ph = reinterpret_cast<const __half2*>(&old);
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) oh[e] = __hadd2(oh[e], ph[e]);
                  }
                  *gp = o;
                }
              }
              __syncwarp();
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * N_TILE);
  }
}

// -------------------------------------------------------------------------------------------- wgrad
// dW[n, k] += oscale * sum over pairs p of sum_m dY_p[m, n] X_p[m, k].  Per pipeline item: 64 rows (samples)
// of one pair's dY (NOUT columns) and X (KIN columns), each as 64-column TMA boxes of 64 rows x 128 B (8 KB).
// As UMMA operands both are MN-major: the 128-byte rows run along M (resp. N) and the row index is K.
// Several pairs implement split precision: (dYhi, Xhi), (dYlo, Xhi), (dYhi, Xlo).
// NOUT = 64 is padded to the UMMA M of 128 by pointing the second 64-channel atom at a zeroed block.
constexpr int WG_ROWS = 64;
constexpr int WG_BOX_BYTES = WG_ROWS * 128;  // 8 KB
constexpr int WG_STAGES = 3;
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_PAIRS = 3;
struct WgMaps { CUtensorMap y[WG_MAX_PAIRS]; CUtensorMap x[WG_MAX_PAIRS]; };

template <int NOUT, int KIN>
__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ WgMaps maps, int npair, int M, float* dW, int ldw, int n_valid, int k_valid,
                const float* oscale_p, float* dbias) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int YB = NOUT / 64, XB = KIN / 64;       // boxes per operand per stage
  constexpr int STAGE_BYTES = (YB + XB) * WG_BOX_BYTES;
  constexpr int MB = (NOUT + 127) / 128;             // 128-row blocks of the accumulator
  constexpr bool PAD_M = (NOUT == 64);
  constexpr int TCOLS = MB * KIN <= 32 ? 32 : (MB * KIN <= 64 ? 64 : (MB * KIN <= 128 ? 128 : (MB * KIN <= 256 ? 256 : 512)));
  static_assert(MB * KIN <= 512, "accumulator does not fit TMEM");
  static_assert(NOUT == 64 || NOUT % 128 == 0, "output channels: 64 or blocks of 128");
  uint8_t* zero_blk = smem + WG_STAGES * STAGE_BYTES;  // 8 KB of zeros (only used when PAD_M)
  uint64_t* bars = reinterpret_cast<uint64_t*>(zero_blk + WG_BOX_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + WG_STAGES;
  uint64_t* done = bars + 2 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_chunks = (M + WG_ROWS - 1) / WG_ROWS;

  if (threadIdx.x == 0) {
    // a stage is released by the MMA commit and, when bias gradients are wanted, by the 4 column-sum warps
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], dbias ? 5 : 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (PAD_M) {
    for (int i = threadIdx.x; i < WG_BOX_BYTES / 16; i += WG_THREADS)
      reinterpret_cast<uint4*>(zero_blk)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();  // make the generic-proxy zeros visible to the tensor core's async-proxy reads
  }
  if (warp == 1) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_work = (int)blockIdx.x < num_chunks;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&maps.y[0]);
      prefetch_tmap(&maps.x[0]);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = blockIdx.x; c < num_chunks; c += gridDim.x) {
        for (int p = 0; p < npair; ++p) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* s = smem + stage * STAGE_BYTES;
          for (int b = 0; b < YB; ++b) tma_load_2d(s + b * WG_BOX_BYTES, &maps.y[p], &full[stage], b * 64, c * WG_ROWS);
          for (int b = 0; b < XB; ++b)
            tma_load_2d(s + (YB + b) * WG_BOX_BYTES, &maps.x[p], &full[stage], b * 64, c * WG_ROWS);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && has_work) {
      constexpr uint32_t idesc = make_idesc(128, KIN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int c = blockIdx.x; c < num_chunks; c += gridDim.x) {
        for (int p = 0; p < npair; ++p) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t y_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t x_addr = y_addr + YB * WG_BOX_BYTES;
          // LBO = distance to the next 64 MN elements: the next box, or the zero block when padding 64 -> 128
          const uint32_t y_lbo = PAD_M ? (smem_u32(zero_blk) - y_addr) : (uint32_t)WG_BOX_BYTES;
#pragma unroll
          for (int k = 0; k < WG_ROWS / 16; ++k) {
            // K advances by 16 rows = 2048 B; SBO = next 8 K rows
            const uint64_t bd = make_desc(x_addr + k * 2048, WG_BOX_BYTES, 1024);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              const uint64_t ad = make_desc(y_addr + mb * 2 * WG_BOX_BYTES + k * 2048, y_lbo, 1024);
              umma_f16(tmem_base + (uint32_t)(mb * KIN), ad, bd, idesc, (first && k == 0) ? 0u : 1u);
            }
          }
          first = false;
          umma_commit(&empty[stage]);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit(done);
    }
  } else if (has_work) {
    const int q = warp & 3;
    const float oscale = oscale_p ? *oscale_p : 1.0f;
    if (dbias) {
      // bias gradient = column sums of dY, taken from the operand boxes while the tensor core consumes them
      // (pairs 0 and 1 carry dY's hi and lo halves; pair 2 repeats hi).  Box = 64 rows x 128 B, 128-byte swizzle:
      // element (r, c) lives at r*128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2.
      constexpr int PAIRS = NOUT / 2;            // column pairs
      constexpr int GROUPS = 128 / PAIRS;        // row groups sharing the 128 threads
      constexpr int RPG = WG_ROWS / GROUPS;      // rows per group
      const int t = threadIdx.x - 64;
      const int cp = t % PAIRS, grp = t / PAIRS;
      const int box = cp >> 5, j = cp & 31;      // 32 column pairs per 64-column box
      float s0 = 0.f, s1 = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int c = blockIdx.x; c < num_chunks; c += gridDim.x) {
        for (int p = 0; p < npair; ++p) {
          mbar_wait(&full[stage], phase);
          if (p < 2) {
            const uint8_t* yb = smem + stage * STAGE_BYTES + box * WG_BOX_BYTES;
#pragma unroll 8
            for (int rr = 0; rr < RPG; ++rr) {
              const int r = grp * RPG + rr;
              const __half2 h = *reinterpret_cast<const __half2*>(yb + r * 128 + ((((j >> 2) ^ (r & 7))) << 4) + (j & 3) * 4);
              const float2 f = __half22float2(h);
              s0 += f.x; s1 += f.y;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      const int col = cp * 2;
      if (col < n_valid) atomicAdd(dbias + col, s0 * oscale);
      if (col + 1 < n_valid) atomicAdd(dbias + col + 1, s1 * oscale);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    for (int mb = 0; mb < MB; ++mb) {
      const int n = mb * 128 + q * 32 + lane;  // output channel (row of dW)
#pragma unroll 1
      for (int c0 = 0; c0 < KIN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * KIN + c0), v);
        if (n < n_valid) {
          float* o = dW + (size_t)n * ldw + c0;
          // 148 CTAs add their partial sums into the same (N, K) block: 128-bit reductions quarter the number of
          // L2 atomic operations of the tail (rows whose start is not 16-byte aligned fall back to scalar adds)
          if (c0 + 32 <= k_valid && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(v[j] * oscale),
                           "f"(v[j + 1] * oscale), "f"(v[j + 2] * oscale), "f"(v[j + 3] * oscale)
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < k_valid) atomicAdd(o + j, v[j] * oscale);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

// Several weight gradients of ONE shape in one launch (a trunk pass has eight 256 x 256 ones, a nerf_skin pass nine
// 64 x 64 ones): the grid is partitioned among the jobs, CTAs [cta0, cta0 + nctas) stream the rows of job j.  Against one
// launch per layer this removes the ramp-up / tail of every launch and -- because a job's partial sums now come from
// ~148 / njobs CTAs instead of 148 -- most of the red.global traffic of the final flush (256 KB per CTA for a 256 x 256
// block), which is what kept the weight gradients at ~33 us per launch however small the batch (profiles/r02: 1.22 ms
// of a 2.1 ms step at 1024 rays per GPU).
constexpr int WG_MAX_JOBS = 9;
// ring depth of the multi-job kernel by shape: as many stages as ~192 KB hold, at most 8.  Three 16 KB stages of a 64 x 64
// job keep 48 KB per SM in flight -- 7 MB over the machine, about what the HBM latency x bandwidth product needs and no
// more: the 64 x 64 launch ran at 5.0 TB/s against 7.1 TB/s of the 256 x 256 one (three 64 KB stages).
constexpr int wg_multi_stages(int nout, int kin) {
  const int stage = (nout + kin) / 64 * WG_BOX_BYTES;
  const int n = (192 * 1024) / stage;
  return n < 3 ? 3 : (n > 8 ? 8 : n);
}
struct WgJob { CUtensorMap y, x; float* dW; float* dbias; int ldw, n_valid, k_valid, cta0, nctas, pad_; };
struct WgJobs { int njobs; int pad_[15]; WgJob job[WG_MAX_JOBS]; };

template <int NOUT, int KIN>
__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_multi_kernel(const __grid_constant__ WgJobs jobs, int M, const float* oscale_p) {
  // which job this CTA works on: CTAs [cta0, cta0 + nctas) of the grid belong to one job and split its row chunks
  int jid = 0;
  for (int j = 1; j < jobs.njobs; ++j)
    if ((int)blockIdx.x >= jobs.job[j].cta0) jid = j;
  const WgJob& J = jobs.job[jid];
  const int cta = (int)blockIdx.x - J.cta0, nctas = J.nctas;
  float* const dW = J.dW;
  float* const dbias = J.dbias;
  const int ldw = J.ldw, n_valid = J.n_valid, k_valid = J.k_valid;
  constexpr int npair = 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int YB = NOUT / 64, XB = KIN / 64;       // boxes per operand per stage
  constexpr int STAGE_BYTES = (YB + XB) * WG_BOX_BYTES;
  constexpr int NST = wg_multi_stages(NOUT, KIN);   // ring depth
  constexpr int MB = (NOUT + 127) / 128;             // 128-row blocks of the accumulator
  constexpr bool PAD_M = (NOUT == 64);
  constexpr int TCOLS = MB * KIN <= 32 ? 32 : (MB * KIN <= 64 ? 64 : (MB * KIN <= 128 ? 128 : (MB * KIN <= 256 ? 256 : 512)));
  static_assert(MB * KIN <= 512, "accumulator does not fit TMEM");
  static_assert(NOUT == 64 || NOUT % 128 == 0, "output channels: 64 or blocks of 128");
  uint8_t* zero_blk = smem + NST * STAGE_BYTES;  // 8 KB of zeros (only used when PAD_M)
  uint64_t* bars = reinterpret_cast<uint64_t*>(zero_blk + WG_BOX_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + NST;
  uint64_t* done = bars + 2 * NST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_chunks = (M + WG_ROWS - 1) / WG_ROWS;

  if (threadIdx.x == 0) {
    // a stage is released by the MMA commit and, when bias gradients are wanted, by the 4 column-sum warps
    for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], dbias ? 5 : 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (PAD_M) {
    for (int i = threadIdx.x; i < WG_BOX_BYTES / 16; i += WG_THREADS)
      reinterpret_cast<uint4*>(zero_blk)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();  // make the generic-proxy zeros visible to the tensor core's async-proxy reads
  }
  if (warp == 1) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_work = cta < num_chunks;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&J.y);
      prefetch_tmap(&J.x);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = cta; c < num_chunks; c += nctas) {
        for (int p = 0; p < npair; ++p) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* s = smem + stage * STAGE_BYTES;
          for (int b = 0; b < YB; ++b) tma_load_2d(s + b * WG_BOX_BYTES, &J.y, &full[stage], b * 64, c * WG_ROWS);
          for (int b = 0; b < XB; ++b)
            tma_load_2d(s + (YB + b) * WG_BOX_BYTES, &J.x, &full[stage], b * 64, c * WG_ROWS);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && has_work) {
      constexpr uint32_t idesc = make_idesc(128, KIN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int c = cta; c < num_chunks; c += nctas) {
        for (int p = 0; p < npair; ++p) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t y_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t x_addr = y_addr + YB * WG_BOX_BYTES;
          // LBO = distance to the next 64 MN elements: the next box, or the zero block when padding 64 -> 128
          const uint32_t y_lbo = PAD_M ? (smem_u32(zero_blk) - y_addr) : (uint32_t)WG_BOX_BYTES;
#pragma unroll
          for (int k = 0; k < WG_ROWS / 16; ++k) {
            // K advances by 16 rows = 2048 B; SBO = next 8 K rows
            const uint64_t bd = make_desc(x_addr + k * 2048, WG_BOX_BYTES, 1024);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              const uint64_t ad = make_desc(y_addr + mb * 2 * WG_BOX_BYTES + k * 2048, y_lbo, 1024);
              umma_f16(tmem_base + (uint32_t)(mb * KIN), ad, bd, idesc, (first && k == 0) ? 0u : 1u);
            }
          }
          first = false;
          umma_commit(&empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit(done);
    }
  } else if (has_work) {
    const int q = warp & 3;
    const float oscale = oscale_p ? *oscale_p : 1.0f;
    if (dbias) {
      // bias gradient = column sums of dY, taken from the operand boxes while the tensor core consumes them
      // (pairs 0 and 1 carry dY's hi and lo halves; pair 2 repeats hi).  Box = 64 rows x 128 B, 128-byte swizzle:
      // element (r, c) lives at r*128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2.
      constexpr int PAIRS = NOUT / 2;            // column pairs
      constexpr int GROUPS = 128 / PAIRS;        // row groups sharing the 128 threads
      constexpr int RPG = WG_ROWS / GROUPS;      // rows per group
      const int t = threadIdx.x - 64;
      const int cp = t % PAIRS, grp = t / PAIRS;
      const int box = cp >> 5, j = cp & 31;      // 32 column pairs per 64-column box
      float s0 = 0.f, s1 = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int c = cta; c < num_chunks; c += nctas) {
        for (int p = 0; p < npair; ++p) {
          mbar_wait(&full[stage], phase);
          if (p < 2) {
            const uint8_t* yb = smem + stage * STAGE_BYTES + box * WG_BOX_BYTES;
#pragma unroll 8
            for (int rr = 0; rr < RPG; ++rr) {
              const int r = grp * RPG + rr;
              const __half2 h = *reinterpret_cast<const __half2*>(yb + r * 128 + ((((j >> 2) ^ (r & 7))) << 4) + (j & 3) * 4);
              const float2 f = __half22float2(h);
              s0 += f.x; s1 += f.y;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
      const int col = cp * 2;
      if (col < n_valid) atomicAdd(dbias + col, s0 * oscale);
      if (col + 1 < n_valid) atomicAdd(dbias + col + 1, s1 * oscale);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    for (int mb = 0; mb < MB; ++mb) {
      const int n = mb * 128 + q * 32 + lane;  // output channel (row of dW)
#pragma unroll 1
      for (int c0 = 0; c0 < KIN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * KIN + c0), v);
        if (n < n_valid) {
          float* o = dW + (size_t)n * ldw + c0;
          // 148 CTAs add their partial sums into the same (N, K) block: 128-bit reductions quarter the number of
          // L2 atomic operations of the tail (rows whose start is not 16-byte aligned fall back to scalar adds)
          if (c0 + 32 <= k_valid && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(v[j] * oscale),
                           "f"(v[j + 1] * oscale), "f"(v[j + 2] * oscale), "f"(v[j + 3] * oscale)
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < k_valid) atomicAdd(o + j, v[j] * oscale);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

}  // namespace tc
}  // namespace moda

using namespace moda;
using namespace moda::tc;

// ------------------------------------------------------------------------------------------ host side
#include "tc_host.cuh"

template <int N_TILE, bool ROWBIAS, bool RANK1, bool MASK>
static int launch_linear(const AMaps& a, const CUtensorMap& b, int M, const ChunkTable& ct, const LinEpi& ep,
                         cudaStream_t stream) {
  const int KC = ct.n;
  int stages = LIN_MAX_STAGES;
  size_t smem = 0;
  for (; stages >= 2; --stages) {
    smem = 1024 + (size_t)KC * N_TILE * CHUNK_K * 2 + (size_t)stages * A_STAGE_BYTES + EPI_WARPS * STG_BYTES + 256 +
           2 * N_TILE * 4;
    if (smem <= 232448) break;
  }
  MODA_REQUIRE(smem <= 232448, "tc_linear: K=%d N=%d needs %zu B of shared memory", KC * 64, N_TILE, smem);
  // a per-DEVICE function attribute: set on every launch (cheap) so that a second device of the process gets it too
  cudaFuncSetAttribute(tc_linear_kernel<N_TILE, ROWBIAS, RANK1, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  const int tiles = (M + TILE_M - 1) / TILE_M;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  tc_linear_kernel<N_TILE, ROWBIAS, RANK1, MASK><<<grid, LIN_THREADS, smem, stream>>>(a, b, M, ct, stages, ep);
  return check_launch("tc_linear");
}

template <int N_TILE>
static int dispatch_epi(const AMaps& a, const CUtensorMap& b, int M, const ChunkTable& ct, const LinEpi& ep,
                        cudaStream_t stream) {
  const bool rb = ep.rowbias != nullptr, r1 = ep.rv != nullptr, mk = ep.mask != nullptr;
  if (!rb && !r1 && !mk) return launch_linear<N_TILE, false, false, false>(a, b, M, ct, ep, stream);
  if (rb && !r1 && !mk) return launch_linear<N_TILE, true, false, false>(a, b, M, ct, ep, stream);
  if (!rb && !r1 && mk) return launch_linear<N_TILE, false, false, true>(a, b, M, ct, ep, stream);
  if (!rb && r1) return launch_linear<N_TILE, false, true, true>(a, b, M, ct, ep, stream);
  MODA_REQUIRE(false, "tc_linear: epilogue combination (rowbias=%d rank1=%d mask=%d) not instantiated", rb, r1, mk);
  return -1;
}

static int dispatch_linear(const AMaps& a, const CUtensorMap& b, int M, int N, const ChunkTable& ct, const LinEpi& ep,
                           cudaStream_t stream) {
  if (N == 256) return dispatch_epi<256>(a, b, M, ct, ep, stream);
  if (N == 128) return dispatch_epi<128>(a, b, M, ct, ep, stream);
  if (N == 64) return dispatch_epi<64>(a, b, M, ct, ep, stream);
  return dispatch_epi<32>(a, b, M, ct, ep, stream);
}

// Y = epi([A1 | A2] B^T).  A1 (M,K1) ld lda1, A2 (M,K2) ld lda2 (K2 may be 0), B (N, K1+K2) ld ldb; all fp16.
extern "C" int moda_tc_linear(const void* A1, int lda1, int K1, const void* A2, int lda2, int K2, const void* B,
                              int ldb, int M, int N, const float* bias, const float* rowbias, int rep, int relu,
                              const void* mask, int ldm, const float* rv, const float* cv, const float* rscale,
                              void* y16, int ldy16, int acc16, float* y32, int ldy32, const float* oscale,
                              cudaStream_t stream) {
  if (M == 0) return 0;
  MODA_REQUIRE(A1 && B && K1 > 0 && K1 % 64 == 0 && K2 % 64 == 0 && (K2 == 0 || A2), "tc_linear: bad operands");
  MODA_REQUIRE(N == 32 || N == 64 || N == 128 || N == 256, "tc_linear: N=%d (supported: 32, 64, 128, 256)", N);
  MODA_REQUIRE(y16 || y32, "tc_linear: no output");
  MODA_REQUIRE((!y16 || (ldy16 % 8 == 0)) && (!y32 || (ldy32 % 4 == 0)) && (!mask || (ldm % 8 == 0)),
               "tc_linear: output / mask row pitch must keep 16-byte alignment");
  AMaps a;
  CUtensorMap b;
  if (int e = make_map(&a.m[0], A1, M, K1, lda1, TILE_M)) return e;
  if (K2 > 0) { if (int e = make_map(&a.m[1], A2, M, K2, lda2, TILE_M)) return e; } else a.m[1] = a.m[0];
  a.m[2] = a.m[0]; a.m[3] = a.m[0];
  if (int e = make_map(&b, B, N, K1 + K2, ldb, N)) return e;
  ChunkTable ct;
  ct.n = (K1 + K2) / 64;
  MODA_REQUIRE(ct.n <= MAX_CHUNKS, "tc_linear: K=%d exceeds %d chunks", K1 + K2, MAX_CHUNKS);
  for (int i = 0; i < ct.n; ++i) {
    ct.src[i] = (unsigned char)(i < K1 / 64 ? 0 : 1);
    ct.col[i] = (unsigned char)(i < K1 / 64 ? i : i - K1 / 64);
  }
  LinEpi ep;
  ep.bias = bias; ep.rowbias = rowbias; ep.rep = rep > 0 ? rep : 1; ep.relu = relu;
  ep.mask = reinterpret_cast<const __half*>(mask); ep.ldm = ldm; ep.rv = rv; ep.cv = cv; ep.rscale = rscale;
  ep.y16 = reinterpret_cast<__half*>(y16); ep.y16lo = nullptr; ep.ldy16 = ldy16; ep.acc16 = acc16; ep.y32 = y32;
  ep.ldy32 = ldy32; ep.oscale = oscale;
  MODA_REQUIRE(!rv || cv, "tc_linear: rank-1 term needs both vectors");
  return dispatch_linear(a, b, M, N, ct, ep, stream);
}

// Split-precision variant for nerf_skin (needs ~fp32 accuracy): every operand is an fp16
// (hi, lo) pair with value = hi + lo, and  x W^T ~= hi Whi^T + lo Whi^T + hi Wlo^T  is evaluated as ONE GEMM
// whose K dimension is the concatenation [hi | lo | hi] against B = [Whi | Whi | Wlo] (B is laid out that way
// by the caller: (N, 3*K)).  Up to two concatenated inputs (x = [x1 | x2]).  The result is written as an fp16
// (hi, lo) pair and/or fp32.
extern "C" int moda_tc_linear_split(const void* A1hi, const void* A1lo, int lda1, int K1, const void* A2hi,
                                    const void* A2lo, int lda2, int K2, const void* B3, int ldb, int M, int N,
                                    const float* bias, const float* rowbias, int rep, int relu, const void* mask,
                                    int ldm, void* yhi, void* ylo, int ldy16, int acc16, float* y32, int ldy32,
                                    const float* oscale, cudaStream_t stream) {
  if (M == 0) return 0;
  MODA_REQUIRE(A1hi && A1lo && B3 && K1 > 0 && K1 % 64 == 0 && K2 % 64 == 0 && (K2 == 0 || (A2hi && A2lo)),
               "tc_linear_split: bad operands");
  MODA_REQUIRE(N == 32 || N == 64 || N == 128 || N == 256, "tc_linear_split: N=%d unsupported", N);
  MODA_REQUIRE(yhi || y32, "tc_linear_split: no output");
  MODA_REQUIRE(!acc16 || !ylo, "tc_linear_split: accumulation is only defined for a plain fp16 output");
  const int c1 = K1 / 64, c2 = K2 / 64, K = K1 + K2;
  ChunkTable ct;
  ct.n = 3 * (c1 + c2);
  MODA_REQUIRE(ct.n <= MAX_CHUNKS, "tc_linear_split: K=%d exceeds %d chunks", 3 * K, MAX_CHUNKS);
  AMaps a;
  CUtensorMap b;
  if (int e = make_map(&a.m[0], A1hi, M, K1, lda1, TILE_M)) return e;
  if (int e = make_map(&a.m[1], A1lo, M, K1, lda1, TILE_M)) return e;
  if (K2 > 0) {
    if (int e = make_map(&a.m[2], A2hi, M, K2, lda2, TILE_M)) return e;
    if (int e = make_map(&a.m[3], A2lo, M, K2, lda2, TILE_M)) return e;
  } else { a.m[2] = a.m[0]; a.m[3] = a.m[1]; }
  if (int e = make_map(&b, B3, N, 3 * K, ldb, N)) return e;
  int i = 0;
  for (int part = 0; part < 3; ++part) {   // [hi | lo | hi]
    const int lo = (part == 1);
    for (int c = 0; c < c1; ++c, ++i) { ct.src[i] = (unsigned char)(lo ? 1 : 0); ct.col[i] = (unsigned char)c; }
    for (int c = 0; c < c2; ++c, ++i) { ct.src[i] = (unsigned char)(lo ? 3 : 2); ct.col[i] = (unsigned char)c; }
  }
  LinEpi ep;
  ep.bias = bias; ep.rowbias = rowbias; ep.rep = rep > 0 ? rep : 1; ep.relu = relu;
  ep.mask = reinterpret_cast<const __half*>(mask); ep.ldm = ldm; ep.rv = nullptr; ep.cv = nullptr; ep.rscale = nullptr;
  ep.y16 = reinterpret_cast<__half*>(yhi); ep.y16lo = reinterpret_cast<__half*>(ylo); ep.ldy16 = ldy16;
  ep.acc16 = acc16; ep.y32 = y32; ep.ldy32 = ldy32; ep.oscale = oscale;
  return dispatch_linear(a, b, M, N, ct, ep, stream);
}

template <int NOUT, int KIN>
static int launch_wgrad(const WgMaps& maps, int npair, int M, float* dW, int ldw, int n_valid, int k_valid,
                        const float* oscale, float* dbias, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)WG_STAGES * ((NOUT + KIN) / 64) * WG_BOX_BYTES + WG_BOX_BYTES + 256;
  cudaFuncSetAttribute(tc_wgrad_kernel<NOUT, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);   // per device: every launch
  const int chunks = (M + WG_ROWS - 1) / WG_ROWS;
  const int grid = chunks < sm_count() ? chunks : sm_count();
  tc_wgrad_kernel<NOUT, KIN><<<grid, WG_THREADS, smem, stream>>>(maps, npair, M, dW, ldw, n_valid, k_valid, oscale,
                                                                    dbias);
  return check_launch("tc_wgrad");
}

static int dispatch_wgrad(const WgMaps& maps, int npair, int N, int K, int M, float* dW, int ldw, int n_valid,
                          int k_valid, const float* oscale, float* dbias, cudaStream_t stream) {
#define MODA_WG(NN, KK) if (N == NN && K == KK) return launch_wgrad<NN, KK>(maps, npair, M, dW, ldw, n_valid, k_valid, oscale, dbias, stream)
  MODA_WG(256, 256); MODA_WG(256, 64); MODA_WG(256, 128); MODA_WG(128, 256); MODA_WG(128, 128); MODA_WG(128, 64);
  MODA_WG(64, 64); MODA_WG(64, 128);
#undef MODA_WG
  MODA_REQUIRE(false, "tc_wgrad: shape N=%d K=%d not instantiated", N, K);
  return -1;
}

// dW (N, ldw) fp32 += oscale * dY (M,N)^T X (M,K); dY, X fp16 row-major.  Accumulates (zero dW first).
// Only the first n_valid rows / k_valid columns are written (operands may carry zero padding).  dbias (N) or NULL:
// += oscale * column sums of dY (the bias gradient), taken from the operand tiles already in shared memory.
extern "C" int moda_tc_wgrad(const void* dY, int ldy, int N, const void* X, int ldx, int K, int M, float* dW, int ldw,
                             int n_valid, int k_valid, const float* oscale, float* dbias, cudaStream_t stream) {
  if (M == 0) return 0;
  MODA_REQUIRE(dY && X && dW, "tc_wgrad: null pointer");
  WgMaps maps;
  if (int e = make_map(&maps.y[0], dY, M, N, ldy, WG_ROWS)) return e;
  if (int e = make_map(&maps.x[0], X, M, K, ldx, WG_ROWS)) return e;
  maps.y[1] = maps.y[2] = maps.y[0];
  maps.x[1] = maps.x[2] = maps.x[0];
  return dispatch_wgrad(maps, 1, N, K, M, dW, ldw, n_valid, k_valid, oscale, dbias, stream);
}

// Split-precision weight gradient: dW += oscale * (dYhi + dYlo)^T (Xhi + Xlo) without the lo*lo term.
// Rows >= n_valid and columns >= k_valid of the padded 64-wide operands are not written.
extern "C" int moda_tc_wgrad_split(const void* dYhi, const void* dYlo, int ldy, int N, const void* Xhi,
                                   const void* Xlo, int ldx, int K, int M, float* dW, int ldw, int n_valid,
                                   int k_valid, const float* oscale, float* dbias, cudaStream_t stream) {
  if (M == 0) return 0;
  MODA_REQUIRE(dYhi && dYlo && Xhi && Xlo && dW, "tc_wgrad_split: null pointer");
  WgMaps maps;
  if (int e = make_map(&maps.y[0], dYhi, M, N, ldy, WG_ROWS)) return e;
  if (int e = make_map(&maps.y[1], dYlo, M, N, ldy, WG_ROWS)) return e;
  maps.y[2] = maps.y[0];
  if (int e = make_map(&maps.x[0], Xhi, M, K, ldx, WG_ROWS)) return e;
  maps.x[1] = maps.x[0];
  if (int e = make_map(&maps.x[2], Xlo, M, K, ldx, WG_ROWS)) return e;
  return dispatch_wgrad(maps, 3, N, K, M, dW, ldw, n_valid, k_valid, oscale, dbias, stream);
}

template <int NOUT, int KIN>
static int launch_wgrad_multi(const tc::WgJobs& jobs, int grid, int M, const float* oscale, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)tc::wg_multi_stages(NOUT, KIN) * ((NOUT + KIN) / 64) * WG_BOX_BYTES + WG_BOX_BYTES + 256;
  cudaFuncSetAttribute(tc_wgrad_multi_kernel<NOUT, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  tc_wgrad_multi_kernel<NOUT, KIN><<<grid, WG_THREADS, smem, stream>>>(jobs, M, oscale);
  return check_launch("tc_wgrad_multi");
}

// njobs (<= 9) weight gradients of the same shape (N, K) over the same M rows in one launch:
//   dW[j] (N, ldw[j]) += (*oscale) dY[j]^T X[j]  (first n_valid[j] rows / k_valid[j] columns),  dbias[j] += column sums.
// Host arrays of length njobs; dbias[j] may be NULL.  Same operand rules as moda_tc_wgrad.
extern "C" int moda_tc_wgrad_multi(int njobs, const void* const* dY, const int* ldy, const void* const* X, const int* ldx,
                                   float* const* dW, const int* ldw, const int* n_valid, const int* k_valid,
                                   float* const* dbias, int N, int K, int M, const float* oscale, cudaStream_t stream) {
  if (M == 0 || njobs == 0) return 0;
  MODA_REQUIRE(njobs > 0 && njobs <= WG_MAX_JOBS && dY && X && dW && ldy && ldx && ldw && n_valid && k_valid,
               "tc_wgrad_multi: bad arguments (1..%d jobs)", WG_MAX_JOBS);
  alignas(64) WgJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  jobs.njobs = njobs;
  const int chunks = (M + WG_ROWS - 1) / WG_ROWS;
  int total = sm_count();
  if (total > chunks * njobs) total = chunks * njobs;
  if (total < njobs) total = njobs;
  int c0 = 0;
  for (int j = 0; j < njobs; ++j) {
    MODA_REQUIRE(dY[j] && X[j] && dW[j], "tc_wgrad_multi: null pointer in job %d", j);
    WgJob& J = jobs.job[j];
    if (int e = make_map(&J.y, dY[j], M, N, ldy[j], WG_ROWS)) return e;
    if (int e = make_map(&J.x, X[j], M, K, ldx[j], WG_ROWS)) return e;
    J.dW = dW[j]; J.dbias = dbias ? dbias[j] : nullptr; J.ldw = ldw[j]; J.n_valid = n_valid[j]; J.k_valid = k_valid[j];
    J.cta0 = c0;
    J.nctas = total / njobs + (j < total % njobs ? 1 : 0);
    c0 += J.nctas;
  }
#define MODA_WGM(NN, KK) if (N == NN && K == KK) return launch_wgrad_multi<NN, KK>(jobs, c0, M, oscale, stream)
  MODA_WGM(256, 256); MODA_WGM(256, 64); MODA_WGM(64, 64); MODA_WGM(128, 128); MODA_WGM(128, 64); MODA_WGM(64, 128);
#undef MODA_WGM
  MODA_REQUIRE(false, "tc_wgrad_multi: shape N=%d K=%d not instantiated", N, K);
  return -1;
}
