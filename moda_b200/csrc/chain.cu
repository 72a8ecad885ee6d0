// Fused layer chains on tcgen05: a whole MLP pass over a 128-sample tile runs inside ONE persistent CTA, with the
// activations living in shared memory (as the next layer's A operand) and TMEM (accumulators), never in HBM
// except for what the other pass needs (fp16 copies for the weight gradients, ReLU sign bits for the adjoint).
//
//   nerf_coarse forward  : PE(xyz) -> 8 x (256, ReLU, skip at layer 5) -> [final ->] dir layer (+ per-ray bias) -> heads
//                          (nnutils/nerf.py:147-198 on the input evaluate_mlp assembles, geom_utils.py:19-57; the final
//                          layer, a linear map without activation, is folded into the dir layer: MODE_FOLD_FINAL)
//   nerf_coarse adjoint  : d_dfe -> [d_fin ->] dY[7] ... dY[0] and dPE, ReLU masks from the saved sign bits
//   nerf_skin forward    : the same engine at width 64 with split-precision (hi, lo) fp16 operands (also nerf_vis)
//   nerf_skin adjoint    : plain fp16 chain
//   nerf_feat (5 x 128)  : forward and adjoint programs on the 256-wide engine (N = 128 / 64 steps)
//
// One engine, table driven ("Program": a list of Steps).  Warp roles of a CTA, one tile in flight (SLOTS = 1):
//   warp 0       weight producer: streams the packed fp16 weight chunks (K = 64 columns each) of every step through
//                a shared-memory ring with TMA; it depends on nothing but ring slots, so it runs far ahead.
//   warp 1       MMA issuer: for every step, waits for the A chunks it needs (written by the previous step's
//                epilogue or by the PE producers), issues tcgen05.mma into one of two TMEM accumulators.
//   warps 2..    epilogue (4 or 8 warps): tcgen05.ld -> bias / per-ray bias / rank-1 term / ReLU / sign-bit mask ->
//                fp16 -> (a) the 128-byte-swizzled shared-memory chunk that is the next step's A operand, written IN
//                PLACE (the step that read it has completed), (b) a TMA store of that same chunk to HBM for the other
//                pass, (c) heads.  Chunk by chunk: the next step's MMAs start on chunk 0 while chunks 1-3 of the
//                current step are still being drained.
//   then         positional-encoding producer warp (forward programs): sincos in fp32, fp16 (hi[, lo]) rows written
//                straight into the swizzled A chunk of the NEXT tile while the current tile is in flight; the last
//                warp saves chunks with TMA stores.
// Two tiles in flight (SLOTS = 2, the default of the 256-wide programs, launched as CTA pairs): warps 0 and 1 are the
// MMA issuers of slot 0 and slot 1, warps 2-17 two groups of eight epilogue warps (which also write the PE chunk of
// their slot's next tile), warp 18 the store warp, warp 19 the weight producer -- see the comment at chain_kernel.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace moda {
namespace chain {

using namespace moda::tc;

constexpr int TILE_M = 128;
constexpr int CHUNK_BYTES = TILE_M * 64 * 2;  // one A chunk: 128 rows x 64 fp16 (128-byte rows, SWIZZLE_128B)
constexpr int MAX_STEPS = 14;
constexpr int MAX_KC = 6;
constexpr int MAX_CHUNKS = 8;
constexpr int MAX_SAVE_MAPS = 14;
constexpr int MAX_STAGES = 8;
constexpr int MAX_DUTY = 48;
constexpr int head_smem(int nh) { return (nh > 1 ? nh - 1 : 1) * 128 * 4 * 4; }   // [warp-of-quarter - 1][row][4] floats (per tile slot)
// the adjoint programs have no heads: their 2 x 2 KB go to the weight ring (a fourth 16 KB stage in two-slot mode)
constexpr int head_bytes(int prog, int slots, int nh) { return prog == 0 ? slots * head_smem(nh) : 0; }

// Epilogue flavour bits (tested at run time: every branch is uniform across the CTA).
enum : int {
  E_RELU = 1,          // max(v, 0)
  E_RANK1 = 2,         // v += gsig[row] * rscale * cvec[col]
  E_MASK_IN = 4,       // v = bit ? v : 0 with the saved sign bits of slot mask_slot
  E_MASK_OUT = 8,      // save sign bits (v > 0) into slot mask_slot
  E_ADD_SX = 16,       // v += fp16 partial parked in the out chunk by an earlier step (same thread, same columns)
  E_HEAD_SIGMA = 32,   // sigma = v . ws + bs (kept until the rgb step)
  E_HEAD_RGB = 64,     // raw[row] = (sigmoid(v Wr^T + br), sigma)
  E_OUT_F32 = 128,     // write columns < ld_y32 of the result to y32 (fp32, global)
  E_LOAD16 = 256,      // no MMA: the tile's rows of an fp16 (M, n) global matrix become the out chunks
  E_LOAD32 = 512,      // no MMA: fp32 (M, load_cols) * (*load_scale) -> fp16, zero padded to n
  E_BIAS = 1024,       // + bias[col] (staged in shared memory)
  E_ROWBIAS = 2048,    // + rowbias[row / rep][col]
  E_SMEM = 4096,       // fp16 result -> shared-memory chunk(s) out_chunk.. (next step's A operand / TMA store source)
  E_LO = 8192,         // also fp16(v - fp16(v)) -> out_lo_chunk.. (split precision)
};

struct Step {
  int n;               // result width (accumulator columns): 64, 128 or 256
  int kc;              // K chunks (0: load step)
  int flags;
  int out_chunk;       // first shared-memory chunk receiving the fp16 result, -1: none
  int out_lo_chunk;    // chunk receiving fp16(v - fp16(v)) (split precision), -1: none
  int save_map;        // TMA-store descriptor for the fp16 result, -1: not saved
  int mask_slot;
  int release_pe;      // last step of the tile that reads the PE chunks
  int bias_off;        // float offset of `bias` (or of the staged per-tile rowbias, one copy per tile slot) in the bias table
  int sd_wait;         // stores_done phases (steps with store-warp work) of this tile that precede this step
  int tile_bias;       // rowbias is constant over a tile (rep % 128 == 0): staged at bias_off (+ slot * n) at every tile start
  const float* bias;     // (n) or null
  const float* rowbias;  // (M / rep, n) or null
  // per K chunk, packed (one constant-bank read on the MMA thread): bits 0-7 shared-memory chunk feeding it, bits 8-15
  // which write of that chunk within the tile it must see, bit 16: that write was made by the epilogue of an earlier
  // MMA step of the tile (with a single-buffered accumulator, i.e. two tile slots, the acc_free wait already covers it)
  unsigned int a_info[MAX_KC];
  unsigned short b_col[MAX_KC];    // chunk index (64-column block) in the packed weights
};

struct Program {
  int nsteps, num_tiles, rep, nchunks, stages;
  int bias_floats;       // size of the shared-memory bias / head-vector table
  int mask_tiles;        // tile stride of the sign-bit buffer: the tile count rounded up to even (same in both launch modes)
  int has_tile_bias;     // some step has tile_bias set
  // head vectors staged behind the biases in shared memory: vec0 (ws, or cvec of the adjoint) at float offset vec_off0,
  // vec1 (Wr, (3, n)) at vec_off1; -1: none
  int vec_off0, vec_off1, vec_len0, vec_len1;
  const float* vec0; const float* vec1;
  long long M;
  int wpt[MAX_CHUNKS];   // writes per tile of each chunk (ready-barrier phases per tile)
  int from_pe[MAX_CHUNKS];  // chunk is written by the PE producer warps (else by the epilogue warps)
  // work list of the store warp, in the order the chunks get written: wait for write `gen` of `chunk`, then (map >= 0)
  // TMA-store it to columns [64 col, 64 col + 64) of save[map].  Every write of every hi chunk is listed (the
  // parity waits need to observe each phase), grouped by step.
  int nduty, sd_per_tile;
  struct Duty { unsigned char chunk, gen, step, col; signed char map; } duty[MAX_DUTY];
  // positional-encoding producers
  int pe_chunk;          // chunk of the hi half (-1: program has no PE producers), lo half in pe_chunk + 1
  int pe_lo, pe_save_map, F;
  float win[10];
  const float* xyz;
  // heads (nerf_coarse forward)
  const float* ws; const float* bs; const float* Wr; const float* br; float* raw;
  int sigma_only;        // density-grid mode: the program ends at the sigma head and raw is (M) = sigma
  // rank-1 term (nerf_coarse adjoint): gsig (M), cvec (n), rscale (device scalar)
  const float* gsig; const float* cvec; const float* rscale;
  // load steps
  const void* load_src; int load_ld; int load_cols; const float* load_scale;
  // fp32 output
  float* y32; int ld_y32;
  unsigned int* maskbits;
  long long* trace;      // optional event trace of block 0's third tile (tools/chain_trace.py), null: off
  Step st[MAX_STEPS];
};

struct Maps {
  CUtensorMap w;                     // packed weights, box = 64 columns x BOX_ROWS rows
  CUtensorMap save[MAX_SAVE_MAPS];   // fp16 outputs, box = 64 columns x 128 rows
};

// progress words of the roles (for the timeout report): [0] weight producer, [1] MMA, [2] epilogue warp 0,
// [3] PE warp 0, [4] store warp duty, [5] store warp state
// (they live in the last 32 bytes of the CTA's dynamic shared memory: see launch())
__device__ __forceinline__ int* dbg_words() {
  extern __shared__ __align__(1024) uint8_t smem_raw_dbg[];
  uint32_t total;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(total));
  return reinterpret_cast<int*>(smem_raw_dbg + total - 32);
}
#define s_dbg (dbg_words())

// spin on an mbarrier phase; a wait that lasts seconds means a protocol bug: trap instead of hanging the GPU
// CLUSTER: the arrivals come from the peer CTA of a pair too (acquire at cluster scope)
template <int SLEEP_NS = 0, int CLUSTER = 0>
__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  if (CLUSTER)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  if (ok) return;
  const long long t0 = clock64();
  do {
    if (CLUSTER)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    else
      asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (SLEEP_NS > 0 && !ok) __nanosleep(SLEEP_NS);   // service warps: do not compete with the epilogue for issue slots
    if (!ok && clock64() - t0 > 4000000000LL) {
      if ((threadIdx.x & 31) == 0)
        printf("moda chain: mbarrier wait timed out (block %d thread %d bar %u parity %u) progress: prod %d mma %d epi %d "
               "pe %d store %d/%d\n", blockIdx.x, threadIdx.x, addr, parity, s_dbg[0], s_dbg[1], s_dbg[2], s_dbg[3],
               s_dbg[4], s_dbg[5]);
      __trap();
    }
  } while (!ok);
}

// debug timeline: every traced thread appends records of 4 x int64 (event, a, b, SM clock) to its own region
// (region = event / 10 < 32, 1000 records each; tr[region] counts them).  Plain stores: the trace costs a few cycles.
// The record count of a region lives in a register of the (single) thread that writes it (`tr_n` in the caller's
// scope): nothing is read back from memory, so an event costs a handful of store instructions.
__device__ __forceinline__ void trace_rec(long long* tr, int ev, int a, int b, int& n) {
  const int region = ev / 10;
  if (n < 1000) {
    long long* r = tr + 32 + 4 * (region * 1000 + n);
    r[0] = ev; r[1] = a; r[2] = b; r[3] = clock64();
    tr[region] = ++n;
  }
}
#ifdef MODA_CHAIN_TRACE
#define MODA_TR(on, ev, a, b) do { if (on) trace_rec(pg.trace, (ev), (a), (b), tr_n); } while (0)
#else
#define MODA_TR(on, ev, a, b) do { } while (0)   // product build: no trace code in the hot loops (tools/build_variant.sh trace -DMODA_CHAIN_TRACE)
#endif

// fp32 pair -> packed fp16 pair, round to nearest, SATURATING to +-65504: an overflowing gradient in the adjoint chains
// (fp16 under a single power-of-two loss scale) is clamped instead of turning into inf and, one layer later, NaN in
// the flat gradient buffer that is all-reduced to every rank
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b));   // d = {hi: first source, lo: second}
  return r;
}
// the same with the ReLU folded into the conversion (max(x, 0) then round): one F2FP instead of two FMNMX + one F2FP
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint32_t pack2_lo(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  return *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void tma_store_2d_u32(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// timing experiment MODA_EXP_ALIAS_PE (results are WRONG): the 256-wide programs keep their fifth (PE) chunk on top of
// chunk 0, which frees 2 x 16 KB of shared memory for two more weight stages -- does the depth of the ring matter?
#ifdef MODA_EXP_ALIAS_PE
#define MODA_CHUNK_OFF(c, wide) ((uint32_t)((wide) ? ((c) & 3) : (c)) * CHUNK_BYTES)
#define MODA_NCHUNKS(n, wide) ((wide) ? (n) - 1 : (n))
#else
#define MODA_CHUNK_OFF(c, wide) ((uint32_t)(c) * CHUNK_BYTES)
#define MODA_NCHUNKS(n, wide) (n)
#endif

// per-thread state of an epilogue warp that does not change within a tile
struct EpiCtx {
  uint32_t sA;          // shared-window address of chunk 0
  uint32_t s_bias;      // shared-window address of the bias table
  uint32_t s_vec0, s_vec1;  // shared-window addresses of the staged head vectors
  uint32_t tmem_row;    // TMEM address of this warp's lane quarter, column 0
  int trow, sw, h;      // row within the tile, its swizzle phase, first sub-block of this warp within a chunk
  int tile, T;
  long long row;
  bool live, lane0;
  float rscale, lscale;
  int tr;               // trace event base (0: off)
  int slot;             // tile slot of this epilogue group (two tiles in flight per CTA: 0 / 1)
  mutable int trn;      // trace records written by this thread
  uint32_t ready2_remote;   // CTA pairs: cluster address of the leader's ready2[slot][0] (0: single-CTA kernel)
  bool step_fence;          // two tile slots: one proxy fence + chunk arrivals per STEP instead of per 64-column chunk
};

// One step's epilogue for this warp (flavour F = st.flags).  Sub-blocks of 16 columns are processed
// chunk-major with the TMEM load of the next sub-block in flight; after each completed 64-column chunk:
// proxy fence, then the warp counts itself in on the chunk's mbarrier (no CTA-wide barrier: warps drift freely).  NH warps share a TMEM lane quarter: warp h owns the
// sub-blocks {h, h + NH, ..} < 4 of every chunk.
template <int CF, int CN, int NH, int ACC_STRIDE, int EPI_THREADS>
__device__ __forceinline__ void run_step_body(const int Frt, const Program& pg, const Maps& maps, const Step& st, const EpiCtx& cx,
                                         uint32_t acc_col, uint64_t* ready, float& hs0, float& hs1, float& hs2,
                                         unsigned long long pm0, unsigned long long pm1) {
  constexpr int SPC = 4 / NH;                  // sub-blocks per chunk for this warp
  const int F = (CF >= 0) ? CF : Frt;   // CF >= 0: flavour known at compile time (straight-line hot paths)
  int& tr_n = cx.trn;
  (void)tr_n;
  const bool MMA = !(F & (E_LOAD16 | E_LOAD32));
  const int n = (CN > 0) ? CN : st.n;   // CN > 0: width known at compile time -> the sub-block loop unrolls completely
  const int nch = n >> 6;
  const int nsub = nch * SPC;
  const uint32_t t_addr = cx.tmem_row + acc_col;
  auto col_of = [&](int i) { return (i / SPC) * 64 + (cx.h + NH * (i % SPC)) * 16; };
  float va[16], vb[16];
#ifdef MODA_EXP_NO_TMEMLD   // timing experiment: the epilogue does not read the accumulator
#define MODA_TMEM_LD_ISSUE(addr, v) do { for (int j_ = 0; j_ < 16; ++j_) (v)[j_] = __uint_as_float((addr) + j_); } while (0)
#define MODA_TMEM_LD_WAIT() do { } while (0)
#else
#define MODA_TMEM_LD_ISSUE(addr, v) tmem_ld16_issue((addr), (v))
#define MODA_TMEM_LD_WAIT() tmem_ld_wait()
#endif
#ifdef MODA_EXP_NO_STS      // timing experiment: the epilogue does not write the A chunks
#define MODA_STS128(...) do { } while (0)
#else
#define MODA_STS128(...) sts128(__VA_ARGS__)
#endif
  if (MMA) MODA_TMEM_LD_ISSUE(t_addr + (uint32_t)col_of(0), va);
  // ReLU sign bits: 16 per sub-block (column j <-> bit 15 - j, set = pre-activation >= 0); this thread's sub-blocks
  // of the step are packed four to a 64-bit word (at most two words) -> at most two global accesses per thread
  // and step instead of one per sub-block
  constexpr int MW = (ACC_STRIDE / 16 / NH + 3) / 4;
  static_assert(MW <= 2, "mask words per thread");
  unsigned long long m0 = 0, m1 = 0;
  unsigned long long* mp = nullptr;
  if (F & (E_MASK_IN | E_MASK_OUT)) {
    mp = reinterpret_cast<unsigned long long*>(pg.maskbits) +
         ((((size_t)st.mask_slot * cx.T + cx.tile) * NH + cx.h) * MW) * TILE_M + cx.trow;
    if (F & E_MASK_IN) { m0 = pm0; m1 = pm1; }   // fetched by the caller before it waited for the accumulator
  }
  float rvv = 0.f;
  if (F & E_RANK1) rvv = cx.live ? pg.gsig[cx.row] * cx.rscale : 0.f;
  const float* rb = nullptr;
  if (F & E_ROWBIAS) rb = st.rowbias + (size_t)((cx.live ? cx.row : 0) / pg.rep) * n;
  const uint32_t sb = cx.s_bias + (uint32_t)((st.bias_off + (st.tile_bias ? cx.slot * st.n : 0)) * 4);
  auto sub = [&](const int i, float* __restrict__ v, float* __restrict__ vn) {
    const int cc = col_of(i);                // first of this thread's 16 columns
    const int c64 = cc >> 6;                 // chunk within the result
    const int p0 = (cc & 63) >> 3;           // first 16-byte piece within the chunk row (2 pieces per sub-block)
    const int chunk = st.out_chunk + c64;
    const uint32_t crow = cx.sA + MODA_CHUNK_OFF(chunk, ACC_STRIDE == 256) + (uint32_t)(cx.trow * 128);
    const bool last_of_chunk = (i % SPC) == SPC - 1;
    if (F & E_LOAD16) {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(pg.load_src) +
                                                        (size_t)cx.row * pg.load_ld + cc);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint4 o = cx.live ? __ldg(src + j) : make_uint4(0, 0, 0, 0);
        sts128(crow + (((p0 + j) ^ cx.sw) << 4), o.x, o.y, o.z, o.w);
      }
    } else {
      if (F & E_LOAD32) {
        const bool in = cx.live && cc < pg.load_cols;
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pg.load_src) +
                                                            (size_t)cx.row * pg.load_ld + cc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = in ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * j] = f.x * cx.lscale; v[4 * j + 1] = f.y * cx.lscale;
          v[4 * j + 2] = f.z * cx.lscale; v[4 * j + 3] = f.w * cx.lscale;
        }
      } else {
        MODA_TMEM_LD_WAIT();
        if (i + 1 < nsub) MODA_TMEM_LD_ISSUE(t_addr + (uint32_t)col_of(i + 1), vn);
      }
      if (F & E_BIAS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = lds128f(sb + (uint32_t)((cc + 4 * j) * 4));
          v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
        }
      }
      if (F & E_ROWBIAS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(rb + cc) + j);
          v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
        }
      }
      if (F & E_RANK1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = lds128f(cx.s_vec0 + (uint32_t)((cc + 4 * j) * 4));
          v[4 * j] = fmaf(rvv, f.x, v[4 * j]); v[4 * j + 1] = fmaf(rvv, f.y, v[4 * j + 1]);
          v[4 * j + 2] = fmaf(rvv, f.z, v[4 * j + 2]); v[4 * j + 3] = fmaf(rvv, f.w, v[4 * j + 3]);
        }
      }
      if (F & E_MASK_IN) {
        const unsigned int mbits = (unsigned int)(((MW == 2 && i >= 4) ? m1 : m0) >> (16 * (i & 3)));
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (mbits & (0x8000u >> j)) ? v[j] : 0.f;
      }
      if (F & E_MASK_OUT) {
        // sign bits by funnel shift, two independent 8-element chains (one instruction per element)
        unsigned int n0 = 0, n1 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          n0 = __funnelshift_l(__float_as_uint(v[j]), n0, 1);
          n1 = __funnelshift_l(__float_as_uint(v[8 + j]), n1, 1);
        }
        const unsigned long long w16 = (unsigned long long)((~((n0 << 8) | n1)) & 0xFFFFu) << (16 * (i & 3));
        if (MW == 2 && i >= 4) m1 |= w16; else m0 |= w16;
      }
      // ReLU: when the fp16 chunk is the only consumer of the activated values, the conversion clamps (pack2_relu below)
      const bool relu_in_cvt = (F & E_RELU) && (F & E_SMEM) &&
                               !(F & (E_ADD_SX | E_HEAD_SIGMA | E_HEAD_RGB | E_OUT_F32 | E_LO));
      if ((F & E_RELU) && !relu_in_cvt) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (F & E_ADD_SX) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint4 o = lds128(crow + (((p0 + j) ^ cx.sw) << 4));
          const __half2* oh = reinterpret_cast<const __half2*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(oh[e]);
            v[8 * j + 2 * e] += f.x; v[8 * j + 2 * e + 1] += f.y;
          }
        }
      }
      if (F & E_HEAD_SIGMA) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = lds128f(cx.s_vec0 + (uint32_t)((cc + 4 * j) * 4));
          hs0 = fmaf(v[4 * j], f.x, hs0); hs0 = fmaf(v[4 * j + 1], f.y, hs0);
          hs0 = fmaf(v[4 * j + 2], f.z, hs0); hs0 = fmaf(v[4 * j + 3], f.w, hs0);
        }
      }
      if (F & E_HEAD_RGB) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f0 = lds128f(cx.s_vec1 + (uint32_t)((cc + 4 * j) * 4));
          const float4 f1 = lds128f(cx.s_vec1 + (uint32_t)((n + cc + 4 * j) * 4));
          const float4 f2 = lds128f(cx.s_vec1 + (uint32_t)((2 * n + cc + 4 * j) * 4));
          hs0 = fmaf(v[4 * j], f0.x, hs0); hs0 = fmaf(v[4 * j + 1], f0.y, hs0);
          hs0 = fmaf(v[4 * j + 2], f0.z, hs0); hs0 = fmaf(v[4 * j + 3], f0.w, hs0);
          hs1 = fmaf(v[4 * j], f1.x, hs1); hs1 = fmaf(v[4 * j + 1], f1.y, hs1);
          hs1 = fmaf(v[4 * j + 2], f1.z, hs1); hs1 = fmaf(v[4 * j + 3], f1.w, hs1);
          hs2 = fmaf(v[4 * j], f2.x, hs2); hs2 = fmaf(v[4 * j + 1], f2.y, hs2);
          hs2 = fmaf(v[4 * j + 2], f2.z, hs2); hs2 = fmaf(v[4 * j + 3], f2.w, hs2);
        }
      }
      if (F & E_OUT_F32) {
        if (cx.live && cc < pg.ld_y32) {
          float4* yp = reinterpret_cast<float4*>(pg.y32 + (size_t)cx.row * pg.ld_y32 + cc);
#pragma unroll
          for (int j = 0; j < 4; ++j) yp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      if (F & E_SMEM) {
        if (relu_in_cvt) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            MODA_STS128(crow + (((p0 + j) ^ cx.sw) << 4), pack2_relu(v[8 * j], v[8 * j + 1]),
                        pack2_relu(v[8 * j + 2], v[8 * j + 3]), pack2_relu(v[8 * j + 4], v[8 * j + 5]),
                        pack2_relu(v[8 * j + 6], v[8 * j + 7]));
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            MODA_STS128(crow + (((p0 + j) ^ cx.sw) << 4), pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
                        pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
        }
        if (F & E_LO) {
          const uint32_t crow_lo = cx.sA + (uint32_t)((st.out_lo_chunk + c64) * CHUNK_BYTES + cx.trow * 128);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            sts128(crow_lo + (((p0 + j) ^ cx.sw) << 4), pack2_lo(v[8 * j], v[8 * j + 1]),
                   pack2_lo(v[8 * j + 2], v[8 * j + 3]), pack2_lo(v[8 * j + 4], v[8 * j + 5]),
                   pack2_lo(v[8 * j + 6], v[8 * j + 7]));
        }
      }
    }
    if ((F & E_SMEM) && last_of_chunk && !cx.step_fence) {
      // this warp's part of the chunk is complete: make the generic-proxy writes visible to the async proxy
      // (tensor core, TMA) and count the warp in; the MMA thread and the store warp wait on the chunk barrier
#ifndef MODA_EXP_NO_FENCE
      fence_async_smem();
#endif
      __syncwarp();
      if (cx.lane0) {
        mbar_arrive(&ready[chunk]);
        if (F & E_LO) mbar_arrive(&ready[st.out_lo_chunk + c64]);
        if (cx.ready2_remote) {   // CTA pair: the leader's MMA thread counts both CTAs' writers
          mbar_arrive_remote(cx.ready2_remote + 8u * (uint32_t)chunk);
          if (F & E_LO) mbar_arrive_remote(cx.ready2_remote + 8u * (uint32_t)(st.out_lo_chunk + c64));
        }
      }
    }
    if ((F & E_SMEM) && cx.step_fence && i == nsub - 1) {
      // two tile slots: the next MMA step of this tile starts on acc_free (after the whole step), so nothing gains from
      // chunk-granular hand-over: ONE proxy fence per step, then the arrivals for all of its chunks (the store warp and,
      // for chunks a load step wrote, the MMA warp wait on them)
#ifndef MODA_EXP_NO_FENCE
      fence_async_smem();
#endif
      __syncwarp();
      if (cx.lane0) {
        for (int c = 0; c < nch; ++c) {
          mbar_arrive(&ready[st.out_chunk + c]);
          if (cx.ready2_remote) mbar_arrive_remote(cx.ready2_remote + 8u * (uint32_t)(st.out_chunk + c));
        }
      }
    }
  };
  if constexpr (CN > 0) {
    // sub-block index, columns, chunk addresses and mask-word positions all fold to constants
    constexpr int NSUB = (CN >> 6) * SPC;
#pragma unroll
    for (int i2 = 0; i2 < NSUB; i2 += 2) {
      sub(i2, va, vb);
      if (i2 + 1 < NSUB) sub(i2 + 1, vb, va);
    }
  } else {
#pragma unroll 1
    for (int i2 = 0; i2 < nsub; i2 += 2) {
      sub(i2, va, vb);
      if (i2 + 1 < nsub) sub(i2 + 1, vb, va);
    }
  }
#ifndef MODA_EXP_NO_MASKW
  if (F & E_MASK_OUT) {
#else
  if (false) {
#endif
    *mp = m0;
    if (MW == 2 && nsub > 4) mp[TILE_M] = m1;
  }
}

// hot flavours inline (straight-line code inside the kernel's register allocation) ...
template <int CF, int CN, int NH, int ACC_STRIDE, int EPI_THREADS>
__device__ __forceinline__ void run_step(const int Frt, const Program& pg, const Maps& maps, const Step& st,
                                         const EpiCtx& cx, uint32_t acc_col, uint64_t* ready, float& hs0, float& hs1,
                                         float& hs2, unsigned long long pm0, unsigned long long pm1) {
  run_step_body<CF, CN, NH, ACC_STRIDE, EPI_THREADS>(Frt, pg, maps, st, cx, acc_col, ready, hs0, hs1, hs2, pm0, pm1);
}
// ... the run-time-flag version (a few steps per tile) out of line, so that it does not weigh on them
template <int NH, int ACC_STRIDE, int EPI_THREADS>
__device__ __noinline__ void run_step_generic(const int Frt, const Program& pg, const Maps& maps, const Step& st,
                                              const EpiCtx& cx, uint32_t acc_col, uint64_t* ready, float& hs0,
                                              float& hs1, float& hs2, unsigned long long pm0, unsigned long long pm1) {
  run_step_body<-1, 0, NH, ACC_STRIDE, EPI_THREADS>(Frt, pg, maps, st, cx, acc_col, ready, hs0, hs1, hs2, pm0, pm1);
}
// the flavours that run once or twice per tile: straight-line too, but out of line, each with its own register
// allocation (inlined next to the hot flavour they push it into spilling)
template <int CF, int NH, int ACC_STRIDE, int EPI_THREADS>
__device__ __noinline__ void run_step_cold(const Program& pg, const Maps& maps, const Step& st, const EpiCtx& cx,
                                           uint32_t acc_col, uint64_t* ready, float* hs, unsigned long long pm0,
                                           unsigned long long pm1) {
  float hs0 = 0.f, hs1 = 0.f, hs2 = 0.f;
  run_step_body<CF, 0, NH, ACC_STRIDE, EPI_THREADS>(CF, pg, maps, st, cx, acc_col, ready, hs0, hs1, hs2, pm0, pm1);
  if (CF & (E_HEAD_SIGMA | E_HEAD_RGB)) { hs[0] = hs0; hs[1] = hs1; hs[2] = hs2; }
}

// One row of the positional encoding, fp16 (hi [, lo for the split-precision 64-wide programs]) straight into the swizzled
// A chunk: hrow = shared-window address of the row, sw = its swizzle phase (row & 7), x = the point.
template <int BOX_ROWS>
__device__ __forceinline__ void pe_row(const Program& pg, const uint32_t hrow, const int sw, const float (&x)[3]) {
    // channels come out in index order; eight neighbours (one 16-byte piece of the swizzled row) share one
  // 128-bit store (idx is a compile-time constant after unrolling, so the tests and the address arithmetic
  // fold away)
  float pend = 0.f;
  uint32_t pk[4], pkl[4];
  auto put = [&](const int idx, const float v) {
    if ((idx & 1) == 0) { pend = v; return; }
    const __half2 hv = __floats2half2_rn(pend, v);
    pk[(idx & 7) >> 1] = *reinterpret_cast<const uint32_t*>(&hv);
    if (BOX_ROWS == 64) {   // split precision: low halves into the next chunk
      const float2 hf = __half22float2(hv);
      const __half2 lv = __floats2half2_rn(pend - hf.x, v - hf.y);
      pkl[(idx & 7) >> 1] = *reinterpret_cast<const uint32_t*>(&lv);
    }
    if ((idx & 7) == 7) {
      const uint32_t a = hrow + (uint32_t)(((idx >> 3) ^ sw) << 4);
      sts128(a, pk[0], pk[1], pk[2], pk[3]);
      if (BOX_ROWS == 64) sts128(a + CHUNK_BYTES, pkl[0], pkl[1], pkl[2], pkl[3]);
    }
  };
  put(0, x[0]); put(1, x[1]); put(2, x[2]);
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const float f = (float)(1 << k);
    const float w = (k < pg.F) ? pg.win[k] : 0.f;
    float sn[3], cs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // SFU sin/cos after an exact two-term Cody-Waite reduction to [-pi, pi]: absolute error below 1e-6
      // for |2^k x| <= 160 rad, far inside the fp16 rounding of the plain path and at the level of the
      // (hi, lo) pair's 2^-22 of the split-precision path
      const float r = x[c] * f;
      const float kk = rintf(r * 0.15915494309189535f);
      float r2 = fmaf(kk, -6.2831854820251465f, r);
      r2 = fmaf(kk, 1.7484555e-7f, r2);
      sn[c] = w * __sinf(r2);
      cs[c] = w * __cosf(r2);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) put(3 + 6 * k + c, sn[c]);
#pragma unroll
    for (int c = 0; c < 3; ++c) put(3 + 6 * k + 3 + c, cs[c]);
  }
  put(63, 0.f);
}

// BOX_ROWS: rows of one weight TMA box = widest accumulator half.  128: nerf_coarse (N <= 256, ring stage 32 KB),
// 64: nerf_skin (N = 64, ring stage 8 KB).  EPI_WARPS (16 or 8) epilogue warps and PE_WARPS (4 or 2) producer warps;
// the 64-wide configuration is sized so that TWO CTAs fit one SM (independent tiles hide each other's latencies).
#ifndef MODA_TRUNK_EPI
#define MODA_TRUNK_EPI 8   // epilogue warps of the 256-wide chains (8 or 16)
#endif
#ifndef MODA_SKIN_BWD_CTAS
#define MODA_SKIN_BWD_CTAS 3   // resident CTAs per SM of the 64-wide adjoint chain (2 chunks: room for 3; 0.363 -> 0.347 ms)
#endif
#ifndef MODA_SKIN_BWD_STAGES
#define MODA_SKIN_BWD_STAGES 8
#endif
#ifndef MODA_SKIN_EPI
#define MODA_SKIN_EPI 4    // epilogue warps of the 64-wide chains (4 or 8)
#endif
constexpr int P_FWD = 0, P_BWD = 1;   // which set of straight-line epilogue flavours a kernel instance carries

// PAIR = 1 (256-wide chains only): launched as clusters of two CTAs.  Each CTA runs the whole protocol on its own tile
// (own A chunks, epilogue, PE producers, store warp), but every layer step is ONE tcgen05.mma.cta_group::2 stream issued
// by the leader CTA (cluster rank 0): M = 256 = both tiles, and each CTA stages and feeds only HALF of every weight
// chunk (N/2 rows of B).  That halves the weight traffic through each SM's shared memory (TMA writes + UMMA reads:
// 256 KB -> 128 KB of the 448 KB a layer step moves), which is what bounds these kernels.
// Cross-CTA signalling: the peer's TMA loads count their bytes on the leader's w_full; writers of both CTAs arrive on
// the leader's ready2 / acc_free2 / stores_done2 (remote mbarrier arrives); tcgen05.commit multicasts acc_full,
// w_empty and pe_free to both CTAs.
//
// SLOTS = 2 (CTA pairs only): TWO tiles in flight per CTA.  Each tile slot has its own A chunks, its own TMEM
// accumulator (256 columns), its own group of EPI_WARPS epilogue warps and its own barriers; the MMA thread alternates
// between the slots step by step, so the tensor core works on one tile's layer while the other tile's epilogue drains
// its accumulator and rewrites its A chunks: the accumulator-full -> epilogue -> chunk-ready -> MMA round trip of a
// tile (which left the tensor pipe idle for ~2/3 of every step with one tile per CTA) is hidden
// behind the other tile's MMAs.  A slot's accumulator is single-buffered: the next step's MMAs start when the
// epilogue has drained it completely.  Per-tile arithmetic is unchanged, so results are bit-identical.
// Tiles of CTA b: b, b + G, b + 2G, ... (G = grid); the j-th of them runs in slot j % SLOTS as that slot's tile j / SLOTS.
//
// Two tile slots, ONE MMA-issuing warp PER SLOT (MODA_DUAL_ISSUE, default): a slot-step costs its issuing warp ~4.2k cycles
// (per K chunk a weight-stage wait, descriptors, four tcgen05.mma and a commit: ~690 cycles against 512 tensor cycles;
// then the stores-done wait, the accumulator commit, the next step's table entry and the acc_free wait: ~1.1k cycles
// during which the tensor pipe, whose queue holds ~4 MMAs, runs dry).  With a single warp alternating between the slots
// those costs serialise: 4.2k cycles per tile-step although the epilogue warps idle 60 % of the time waiting for an
// accumulator (profiles/r02_s20_trace_*).  With one issuer per slot the fixed costs of one slot's step overlap the other
// slot's MMAs; the weight ring is shared, each issuer tracking the ring position of its own chunks in the global order
// (tile iteration, step, slot, K chunk) the producer fills it in.  The producer moves to the last warp.
#ifndef MODA_DUAL_ISSUE
#define MODA_DUAL_ISSUE 1
#endif
// (with two issuers the positional encoding is produced by the epilogue warps, which wait for an accumulator 60 % of the
// time, instead of a dedicated warp: warps are allocated in groups of four, so a 21st warp would cost the registers of
// 24 -- 80 per thread instead of 96)
constexpr int chain_threads(int slots, int epi_warps, int pe_warps) {
  return ((slots == 2 && MODA_DUAL_ISSUE) ? (4 + slots * epi_warps) : (3 + slots * epi_warps + pe_warps)) * 32;
}
template <int BOX_ROWS, int EPI_WARPS, int PE_WARPS, int MIN_CTAS, int PROG, int PAIR, int SLOTS>
__global__ void __launch_bounds__(chain_threads(SLOTS, EPI_WARPS, PE_WARPS), MIN_CTAS)
chain_kernel(const __grid_constant__ Program pg, const __grid_constant__ Maps maps) {
  static_assert(!PAIR || BOX_ROWS == 128, "CTA pairs: 256-wide chains only");
  static_assert(SLOTS == 1 || (SLOTS == 2 && PAIR == 1 && PE_WARPS == 1), "two tile slots: CTA-pair kernels only");
  constexpr int ACC_STRIDE = (BOX_ROWS == 128) ? 256 : 64;       // TMEM columns per accumulator
  constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  constexpr int BOX_BYTES = BOX_ROWS * 128;
  constexpr int STAGE_BYTES = PAIR ? BOX_BYTES : ((BOX_ROWS == 128) ? 2 * BOX_BYTES : BOX_BYTES);
  constexpr int PB_ROWS = 32;                                    // PAIR: rows of one weight TMA box (4 KB)
  constexpr int NH = EPI_WARPS / 4;                              // warps sharing one TMEM lane quarter
  constexpr int EPI_THREADS = EPI_WARPS * 32;
  constexpr int PE_THREADS = PE_WARPS * 32;
  (void)PE_THREADS;
  constexpr int PE_ROWS = TILE_M / PE_THREADS;                   // rows per producer thread
  constexpr bool DUAL = (SLOTS == 2) && (MODA_DUAL_ISSUE != 0);  // one MMA-issuing warp per tile slot (warps 0 and 1)
  constexpr int W_PROD = DUAL ? (3 + SLOTS * EPI_WARPS) : 0;     // warp that streams the weights
  // EPI_PE: the positional encoding of a slot's next tile is written by its epilogue warps (one row per thread) instead
  // of the PE warp.  Always with two issuers (no warp left for it); for the 64-wide forward program because its single
  // PE warp needs ~19k cycles per tile (split hi/lo rows) and the PE chunk is single-buffered, so that time adds to the
  // 20k cycles of the five steps that read the chunk: 39k cycles per tile and CTA (profiles/r02_s12_trace_skin_fwd.txt).
  // Four warps write the same rows in a quarter of the time, right after the last step that reads the old chunk.
#ifndef MODA_SKIN_EPI_PE
#define MODA_SKIN_EPI_PE 1
#endif
  constexpr bool EPI_PE = DUAL || (BOX_ROWS == 64 && EPI_WARPS == 4 && MODA_SKIN_EPI_PE != 0);
  constexpr int PE_ARRIVALS = EPI_PE ? 4 : PE_WARPS;             // warps that write (and arrive on) a PE chunk

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                            // SLOTS x nchunks x 16 KB
  uint8_t* sB = sA + (size_t)SLOTS * MODA_NCHUNKS(pg.nchunks, BOX_ROWS == 128) * CHUNK_BYTES;   // stages x STAGE_BYTES
  float* s_head = reinterpret_cast<float*>(sB + (size_t)pg.stages * STAGE_BYTES);   // SLOTS x head_smem
  float* s_bias = s_head + head_bytes(PROG, SLOTS, NH) / 4;      // bias_floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + (size_t)pg.bias_floats);
  uint64_t* w_full = bars;                       // [MAX_STAGES]
  uint64_t* w_empty = w_full + MAX_STAGES;       // [MAX_STAGES]
  uint64_t* ready = w_empty + MAX_STAGES;        // [2][MAX_CHUNKS] chunk (re)written and visible to the tensor core (per slot)
  uint64_t* acc_full = ready + 2 * MAX_CHUNKS;   // [2] per accumulator
  uint64_t* acc_free = acc_full + 2;             // [2]
  uint64_t* pe_free = acc_free + 2;              // [2] per slot
  uint64_t* stores_done = pe_free + 2;           // [2] per slot, one phase per step: its TMA stores have read their chunks
  // CTA pairs, used in the leader CTA only: the same three conditions counted over both CTAs
  uint64_t* ready2 = stores_done + 2;            // [2][MAX_CHUNKS]
  uint64_t* acc_free2 = ready2 + 2 * MAX_CHUNKS; // [2]
  uint64_t* stores_done2 = acc_free2 + 2;        // [2]
  // two issuers: turn[x] hands the weight ring to issuer x (parity waits on a ring barrier are only sound for a consumer
  // that is less than one ring revolution ahead of the fills, which holds when chunks are claimed in their global order)
  uint64_t* turn = stores_done2 + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + 2);
  constexpr int SLOT_CHUNK_BYTES_MAX = MAX_CHUNKS * CHUNK_BYTES;
  (void)SLOT_CHUNK_BYTES_MAX;
  const uint32_t slot_bytes = (uint32_t)MODA_NCHUNKS(pg.nchunks, BOX_ROWS == 128) * CHUNK_BYTES;   // A chunks of one tile slot
  // accumulator of (slot, n-th MMA step of that slot) and the phase its barriers are in: one tile per CTA ping-pongs
  // over both accumulators step by step, two tile slots own one accumulator each
  auto acc_of = [](int slot, uint32_t ctr) { return SLOTS == 2 ? slot : (int)(ctr & 1); };
  auto acc_phase = [](uint32_t ctr) { return SLOTS == 2 ? (ctr & 1) : ((ctr >> 1) & 1); };
  // Scope of the MMA thread's waits on barriers that the PEER CTA's warps arrive on.  What those arrivals order is the
  // peer's shared memory, already visible to the peer's async proxy (fence.proxy.async before the arrive) and read only
  // by the peer's tensor core / TMA once this thread issues the MMA; this thread reads none of it.  A cluster-scope
  // acquire makes ptxas add an L1 invalidation (CCTL.IVALL) after every wait, on the kernel's serial thread.
#ifdef MODA_MMA_WAIT_CLUSTER
  constexpr int MMA_WAIT_CLUSTER = PAIR;
#else
  constexpr int MMA_WAIT_CLUSTER = 0;
#endif
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  (void)leader;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = pg.num_tiles;
  const int ntile = (T - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA (grid <= T)

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int k = 0; k < 2; ++k) {
      for (int i = 0; i < MAX_CHUNKS; ++i) mbar_init(&ready[k * MAX_CHUNKS + i], pg.from_pe[i] ? PE_ARRIVALS : EPI_WARPS);
      mbar_init(&stores_done[k], 1);
      mbar_init(&acc_full[k], 1); mbar_init(&acc_free[k], EPI_WARPS);
      mbar_init(&pe_free[k], 1);
      mbar_init(&turn[k], 1);
      if (PAIR) {
        for (int i = 0; i < MAX_CHUNKS; ++i) mbar_init(&ready2[k * MAX_CHUNKS + i], 2 * (pg.from_pe[i] ? PE_ARRIVALS : EPI_WARPS));
        mbar_init(&acc_free2[k], 2 * EPI_WARPS);
        mbar_init(&stores_done2[k], 2);
      }
    }
    fence_barrier_init();
  }
  // biases of every step, staged once (step s reads the bias_off[s] region of the table)
  for (int s = 0; s < pg.nsteps; ++s) {
    const Step& st = pg.st[s];
    if (st.bias)
      for (int i = threadIdx.x; i < st.n; i += blockDim.x) s_bias[st.bias_off + i] = st.bias[i];
  }
  if (pg.vec_off0 >= 0)
    for (int i = threadIdx.x; i < pg.vec_len0; i += blockDim.x) s_bias[pg.vec_off0 + i] = pg.vec0[i];
  if (pg.vec_off1 >= 0)
    for (int i = threadIdx.x; i < pg.vec_len1; i += blockDim.x) s_bias[pg.vec_off1 + i] = pg.vec1[i];
  if (warp == 1) { if (PAIR) tmem_alloc_pair(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_PROD) {
    // ================================================================== weight producer
    if (lane == 0) {
      prefetch_tmap(&maps.w);
      int stage = 0;
      uint32_t phase = 0;
      // same order as the MMA thread consumes: tile iteration, step, slot (active slots of the iteration), K chunk
      for (int it = 0; it * SLOTS < ntile; ++it) {
        const int nact = (ntile - it * SLOTS < SLOTS) ? ntile - it * SLOTS : SLOTS;
        for (int s = 0; s < pg.nsteps; ++s) {
          const Step& st = pg.st[s];
          const int boxes = st.n > BOX_ROWS ? st.n / BOX_ROWS : 1;
          s_dbg[0] = s;
          for (int slot = 0; slot < nact; ++slot)
          for (int kc = 0; kc < st.kc; ++kc) {
            wait_or_trap<0>(&w_empty[stage], phase ^ 1);
            if (PAIR) {
              // this CTA's half of the chunk: rows [crank n/2, (crank + 1) n/2) of the packed weights, in 32-row
              // boxes; the bytes of both halves are counted on the leader's barrier
              const int half = st.n >> 1;
              if (leader) mbar_expect_tx(&w_full[stage], (uint32_t)(st.n * 128));
              const uint32_t bar = mapa_u32(smem_u32(&w_full[stage]), 0);
              const uint32_t dst = smem_u32(sB + (size_t)stage * STAGE_BYTES);
              for (int bx = 0; bx * PB_ROWS < half; ++bx)
                tma_load_2d_pair(dst + (uint32_t)(bx * PB_ROWS * 128), &maps.w, bar, (int)st.b_col[kc] * 64,
                                 (int)crank * half + bx * PB_ROWS);
            } else {
              mbar_expect_tx(&w_full[stage], (uint32_t)(boxes * BOX_BYTES));
              for (int bx = 0; bx < boxes; ++bx)
                tma_load_2d(sB + (size_t)stage * STAGE_BYTES + bx * BOX_BYTES, &maps.w, &w_full[stage],
                            (int)st.b_col[kc] * 64, bx * BOX_ROWS);
            }
            if (++stage == pg.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || (DUAL && warp == 0)) {
    // ================================================================== MMA issuer (CTA pairs: the leader's only)
    // The WHOLE warp runs the loop converged (every lane polls the barriers) and one elected lane issues the tensor
    // core instructions: the loop state is then warp-uniform and ptxas keeps descriptors, barrier addresses and counters
    // in uniform registers.  Run by a single diverged lane, every tcgen05.mma cost an ELECT plus seven R2UR moves and
    // the ~100 dependent instructions per K chunk, arbitrated against four busy epilogue warps on the same scheduler,
    // took ~900 cycles against the 512 tensor cycles of the chunk (profiles/r02_*: tensor pipe 37 % active whatever
    // the epilogue did): this thread, not the tensor core or the epilogue, bounded the kernel.
    if (leader) {
      const int my_slot = warp;   // DUAL: this warp issues for one tile slot only
      (void)my_slot;
      uint32_t turn_ctr = 0;      // DUAL: hand-overs of the weight ring received so far
      bool turn_owed = false;     // DUAL, slot 0: the other issuer will hand the ring back before my next chunks
      (void)turn_ctr; (void)turn_owed;
      int stage = 0;
      uint32_t phase = 0, gstep = 0;
      uint32_t mma_ctr0 = 0, mma_ctr1 = 0;               // MMA steps issued so far, per slot
      int tr_n = 0;
      (void)tr_n;
      const uint32_t dhi = (uint32_t)(make_desc(0, 16, 1024) >> 32);   // everything but the start address
      const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
      const uint32_t w_full_u = smem_u32(w_full), w_empty_u = smem_u32(w_empty), acc_full_u = smem_u32(acc_full),
                     pe_free_u = smem_u32(pe_free);
      for (int it = 0; it * SLOTS < ntile; ++it) {
        const int nact = (ntile - it * SLOTS < SLOTS) ? ntile - it * SLOTS : SLOTS;
        for (int s = 0; s < pg.nsteps; ++s) {
          const Step& st = pg.st[s];
          const int nkc = st.kc, n = st.n, rel_pe = st.release_pe;
          // stores_done completes one phase per step that has store-warp work; `sd_need` of them precede this
          // step.  It grows by at most one per step and is waited for at every step, so no phase is skipped
          // (parity waits must observe each one), and the store warp cannot run ahead: it only arrives after it
          // has seen chunks written by an epilogue that was itself released by this thread.
          const uint32_t sd_need = (uint32_t)it * (uint32_t)pg.sd_per_tile + (uint32_t)st.sd_wait;
          const uint32_t idesc = make_idesc(PAIR ? 2 * TILE_M : TILE_M, n, 0, 0);
#pragma unroll 1
          for (int slot = 0; slot < nact; ++slot, ++gstep) {
            if (DUAL && slot != my_slot) {
              // the other issuer's chunks: only the ring position moves on
              stage += nkc;
              while (stage >= pg.stages) { stage -= pg.stages; phase ^= 1; }
              continue;
            }
            if (lane == 0) s_dbg[1] = (int)gstep;
            uint64_t* const sd_bar = PAIR ? &stores_done2[slot] : &stores_done[slot];
            if (nkc == 0) {
              if (sd_need > 0) wait_or_trap<0, MMA_WAIT_CLUSTER>(sd_bar, (sd_need - 1) & 1);
              continue;
            }
            const uint32_t ctr = slot ? mma_ctr1 : mma_ctr0;
            const int b = acc_of(slot, ctr);
            // trace regions: 0 for the single issuer / slot 0's, 20 for slot 1's
            const int tr = (pg.trace && blockIdx.x == 0 && it == 2 && lane == 0) ? ((DUAL && slot) ? 201 : 1) : 0;
            (void)tr;
            MODA_TR(tr, tr - 1 + 0, s, slot);
            // the epilogue that last read this accumulator has drained it
            wait_or_trap<0, MMA_WAIT_CLUSTER>(PAIR ? &acc_free2[b] : &acc_free[b], acc_phase(ctr) ^ 1);
            tc_fence_after();
            MODA_TR(tr, tr - 1 + 1, s, slot);
            const uint32_t d_tmem = tmem_base + (uint32_t)(b * ACC_STRIDE);
            const uint32_t sA_slot = sA_u + (uint32_t)slot * slot_bytes;
            if (DUAL) {
              // claim the ring: the chunks before mine in the global order have been seen by their issuer
              if (my_slot == 1 || turn_owed) {
                wait_or_trap(&turn[my_slot], turn_ctr & 1);
                ++turn_ctr;
                turn_owed = false;
              }
            }
            uint64_t* const rdy = (PAIR ? ready2 : ready) + slot * MAX_CHUNKS;
#pragma unroll 1
            for (int kc = 0; kc < nkc; ++kc) {
              const uint32_t info = st.a_info[kc];
              const int c = (int)(info & 0xFF);
              // weights first: the producer runs far ahead, so this check completes while the epilogue is still writing
              // the A chunk, and nothing but the descriptors stands between the chunk's arrival and the MMA issue
              wait_or_trap(&w_full[stage], phase);
              MODA_TR(tr, tr - 1 + 3, s, kc);
              // two tile slots: chunks written by this tile's earlier MMA-step epilogues are complete once the
              // accumulator has been handed back (their warps arrived on the chunk barriers before acc_free)
              if (!(SLOTS == 2 && (info & 0x10000))) {
                const uint32_t gen = (uint32_t)it * (uint32_t)pg.wpt[c] + ((info >> 8) & 0xFF);
                wait_or_trap<0, MMA_WAIT_CLUSTER>(&rdy[c], gen & 1);
              }
              tc_fence_after();   // (measured: dropping this per-chunk fence changes nothing, 1.32 / 1.29 ms either way)
              MODA_TR(tr, tr - 1 + 5, s, kc);
              // the descriptors of a chunk are built once; the three K = 16 advances are plain additions (32 bytes = 2
              // units of the 16-byte address field, which cannot carry out of its 14 bits within a 16 KB chunk)
              const uint32_t alo = ((sA_slot + MODA_CHUNK_OFF(c, BOX_ROWS == 128)) & 0x3FFFF) >> 4;
              const uint32_t blo = ((sB_u + (uint32_t)stage * STAGE_BYTES) & 0x3FFFF) >> 4;
              if (elect_one()) {
                if (PAIR) {
                  umma_f16_pair_lohi(d_tmem, alo, blo, dhi, idesc, kc ? 1u : 0u);
                  umma_f16_pair_lohi(d_tmem, alo + 2, blo + 2, dhi, idesc, 1u);
                  umma_f16_pair_lohi(d_tmem, alo + 4, blo + 4, dhi, idesc, 1u);
                  umma_f16_pair_lohi(d_tmem, alo + 6, blo + 6, dhi, idesc, 1u);
                  umma_commit_pair_u32(w_empty_u + 8u * (uint32_t)stage, 3);
                } else {
                  umma_f16_lohi(d_tmem, alo, blo, dhi, idesc, kc ? 1u : 0u);
                  umma_f16_lohi(d_tmem, alo + 2, blo + 2, dhi, idesc, 1u);
                  umma_f16_lohi(d_tmem, alo + 4, blo + 4, dhi, idesc, 1u);
                  umma_f16_lohi(d_tmem, alo + 6, blo + 6, dhi, idesc, 1u);
                  umma_commit_u32(w_empty_u + 8u * (uint32_t)stage);
                }
              }
              __syncwarp();
              MODA_TR(tr, tr - 1 + 6, s, kc);
              if (++stage == pg.stages) { stage = 0; phase ^= 1; }
            }
            if (DUAL && nact == 2) {
              // every weight stage of this step has been observed full: the ring goes to the other issuer
              if (lane == 0) mbar_arrive(&turn[my_slot ^ 1]);
              if (my_slot == 0) turn_owed = true;
            }
            MODA_TR(tr, tr - 1 + 2, s, slot);
            // In-place rewrite guarantee: the epilogue of THIS step overwrites chunks that TMA stores of earlier steps
            // read; it starts on acc_full, so that is only signalled once those stores have read their source.
            if (sd_need > 0) wait_or_trap<0, MMA_WAIT_CLUSTER>(sd_bar, (sd_need - 1) & 1);
            MODA_TR(tr, tr - 1 + 7, s, slot);
            if (elect_one()) {
              if (PAIR) {
                umma_commit_pair_u32(acc_full_u + 8u * (uint32_t)b, 3);
                if (rel_pe) umma_commit_pair_u32(pe_free_u + 8u * (uint32_t)slot, 3);
              } else {
                umma_commit_u32(acc_full_u + 8u * (uint32_t)b);
                if (rel_pe) umma_commit_u32(pe_free_u + 8u * (uint32_t)slot);
              }
            }
            __syncwarp();
            MODA_TR(tr, tr - 1 + 4, s, slot);
            if (slot) mma_ctr1 = ctr + 1; else mma_ctr0 = ctr + 1;
          }
        }
      }
      (void)w_full_u;
    }
  } else if (warp < 2 + SLOTS * EPI_WARPS) {
    // ================================================================== epilogue (one group of EPI_WARPS per tile slot)
    // Column assignment is chunk-major: for every 64-column chunk of the result, warp (q, h) owns the 32-column
    // sub-blocks {h, h + NH, ..} < 2 of lane quarter q.  All warps therefore finish chunk 0 first, and the next
    // step's MMAs start on it while chunks 1.. are still being drained.
    const int slot = (warp - 2) / EPI_WARPS;
    const int ew = (warp - 2) - slot * EPI_WARPS;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    float* const s_head_slot = s_head + slot * (head_smem(NH) / 4);
    uint64_t* const ready_slot = ready + slot * MAX_CHUNKS;
    const int bar_id = 4 + 2 * slot;        // named barriers of this group: bar_id (heads), bar_id + 1 (tile biases)
    EpiCtx cx;
    cx.slot = slot;
    cx.trn = 0;
    int& tr_n = cx.trn;
    (void)tr_n;
    cx.sA = smem_u32(sA) + (uint32_t)slot * slot_bytes;
    cx.s_bias = smem_u32(s_bias);
    cx.s_vec0 = cx.s_bias + (uint32_t)((pg.vec_off0 >= 0 ? pg.vec_off0 : 0) * 4);
    cx.s_vec1 = cx.s_bias + (uint32_t)((pg.vec_off1 >= 0 ? pg.vec_off1 : 0) * 4);
    cx.tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
    cx.h = ew >> 2;
    cx.trow = q * 32 + lane;
    cx.sw = cx.trow & 7;
    cx.lane0 = lane == 0;
    cx.rscale = pg.rscale ? *pg.rscale : 1.0f;
    cx.lscale = pg.load_scale ? *pg.load_scale : 1.0f;
    cx.T = pg.mask_tiles;
    cx.ready2_remote = PAIR ? mapa_u32(smem_u32(&ready2[slot * MAX_CHUNKS]), 0) : 0u;
#ifdef MODA_EXP_CHUNK_FENCE
    cx.step_fence = false;
#else
    // forward programs only: measured 1.45 -> 1.41 ms for the forward and 1.51 -> 1.58 ms for the adjoint, whose
    // stores-done wait on the MMA warp is exposed when the TMA stores of a step start late
    cx.step_fence = SLOTS == 2 && PROG == P_FWD;
#endif
    const uint32_t acc_free2_remote = PAIR ? mapa_u32(smem_u32(&acc_free2[0]), 0) : 0u;
    (void)acc_free2_remote;
    const int h = cx.h, trow = cx.trow;
    uint32_t mma_ctr = 0;
    float sig_keep = 0.f;
    // DUAL: the four h == 0 warps of the slot write the PE chunk of the slot's NEXT tile (thread = row, as in the
    // epilogue) right after the epilogue of the last step that reads the current one -- its acc_full says those MMAs
    // have completed, and the chunk's TMA save was waited for by the issuer steps earlier
    const bool pe_duty = EPI_PE && pg.pe_chunk >= 0 && h == 0;
    auto pe_fetch = [&](int tile_n, float (&x)[3]) {
      const long long row = (long long)tile_n * TILE_M + trow;
#pragma unroll
      for (int c = 0; c < 3; ++c) x[c] = (row < pg.M) ? __ldg(pg.xyz + row * 3 + c) : 0.f;
    };
    auto pe_write = [&](const float (&x)[3]) {
      pe_row<BOX_ROWS>(pg, cx.sA + MODA_CHUNK_OFF(pg.pe_chunk, BOX_ROWS == 128) + (uint32_t)(trow * 128), cx.sw, x);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ready_slot[pg.pe_chunk]);
        if (pg.pe_lo) mbar_arrive(&ready_slot[pg.pe_chunk + 1]);
        if (PAIR) mbar_arrive_remote(cx.ready2_remote + 8u * (uint32_t)pg.pe_chunk);
      }
    };
    if (pe_duty && slot < ntile) {
      float x0[3];
      pe_fetch((int)blockIdx.x + slot * (int)gridDim.x, x0);
      pe_write(x0);
    }
    for (int it = 0; it * SLOTS + slot < ntile; ++it) {
      const int tile = (int)blockIdx.x + (it * SLOTS + slot) * (int)gridDim.x;
      cx.tile = tile;
      cx.tr = (pg.trace && blockIdx.x == 0 && it == 2 && lane == 0) ? 40 + 10 * ew + 80 * slot : 0;   // regions 4..: one per warp
      cx.row = (long long)tile * TILE_M + trow;
      cx.live = cx.row < pg.M;
      if (pg.has_tile_bias) {
        // per-ray bias rows (the hoisted per-ray-constant inputs) when a tile never straddles two rays: staged once
        // per tile and then read like any bias.  First barrier: every warp is done with the previous tile's rows.
        named_bar(bar_id + 1, EPI_THREADS);
        const int et = threadIdx.x - 64 - slot * EPI_THREADS;
        long long row0 = (long long)tile * TILE_M;
        if (row0 >= pg.M) row0 = pg.M - 1;   // dead tile of an odd tile count (CTA pairs): any valid ray
        const size_t ray = (size_t)(row0 / pg.rep);
        for (int s = 0; s < pg.nsteps; ++s) {
          const Step& st = pg.st[s];
          if (st.tile_bias)
            for (int i = et; i < st.n; i += EPI_THREADS) s_bias[st.bias_off + slot * st.n + i] = st.rowbias[ray * st.n + i];
        }
        named_bar(bar_id + 1, EPI_THREADS);
      }
      for (int s = 0; s < pg.nsteps; ++s) {
        const Step& st = pg.st[s];
        const int flags = st.flags;
        if (ew == 0 && lane == 0) s_dbg[2] = it * 100 + s;
        int b = 0;
        const bool pe_now = pe_duty && st.release_pe && (it + 1) * SLOTS + slot < ntile;
        float xn[3] = {0.f, 0.f, 0.f};
        if (pe_now) pe_fetch(tile + SLOTS * (int)gridDim.x, xn);   // in flight while this step's accumulator is awaited
        // the saved ReLU sign words of this step (adjoint programs) come from HBM: ask for them before waiting for
        // the accumulator, so that their latency is off the step's critical path
        unsigned long long pm0 = 0, pm1 = 0;
        if (flags & E_MASK_IN) {
          constexpr int MWc = (ACC_STRIDE / 16 / NH + 3) / 4;
          const unsigned long long* mp = reinterpret_cast<const unsigned long long*>(pg.maskbits) +
              ((((size_t)st.mask_slot * cx.T + cx.tile) * NH + cx.h) * MWc) * TILE_M + cx.trow;
          asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(pm0) : "l"(mp));
          if (MWc == 2 && (st.n >> 6) * (4 / NH) > 4) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(pm1) : "l"(mp + TILE_M));
        }
        if (st.kc > 0) {
          b = acc_of(slot, mma_ctr);
          wait_or_trap(&acc_full[b], acc_phase(mma_ctr));
          tc_fence_after();
        }
        MODA_TR(cx.tr, cx.tr, s, 0);
        float hs0 = 0.f, hs1 = 0.f, hs2 = 0.f;
        const uint32_t acc_col = (uint32_t)(b * ACC_STRIDE);
        // every flavour a program uses runs as a straight-line specialisation (a taken branch costs an
        // instruction-fetch bubble: ~30 of them per sub-block in the run-time-flag version, which stays as the
        // out-of-line fallback for the rare combinations, e.g. a forward without saved sign bits)
        constexpr int LO = (BOX_ROWS == 128) ? 0 : E_LO;
        constexpr int HOT_FWD = E_BIAS | E_RELU | E_MASK_OUT | E_SMEM | LO;
        constexpr int HOT_BWD = E_MASK_IN | E_SMEM;
#define MODA_FLAVOUR(FL) \
  if (flags == (FL) && st.n == ACC_STRIDE) run_step<(FL), ACC_STRIDE, NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1); else
#define MODA_FLAVOUR_N(FL, NN) \
  if (flags == (FL) && st.n == (NN)) run_step<(FL), (NN), NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1); else
#define MODA_COLD(FL) \
  if (flags == (FL)) { float hs[3]; run_step_cold<(FL), NH, ACC_STRIDE, EPI_THREADS>(pg, maps, st, cx, acc_col, ready_slot, hs, pm0, pm1); \
                       if ((FL) & (E_HEAD_SIGMA | E_HEAD_RGB)) { hs0 = hs[0]; hs1 = hs[1]; hs2 = hs[2]; } } else
        // Dispatch modes (per direction): 2 = every flavour a program uses inline, unrolled over its compile-time
        // width (default; measured on B200 against mode 0: trunk fwd 2.03 -> 1.65 ms, trunk bwd 2.08 -> 1.72 ms,
        // skin fwd 0.80 -> 0.71 ms, skin bwd 0.62 -> 0.39 ms per pass); 1 = hot flavour inline, the others as
        // out-of-line straight-line functions; 0 = hot flavour inline, the others through the run-time-flag version.
#ifndef MODA_CHAIN_DISPATCH_FWD
#define MODA_CHAIN_DISPATCH_FWD 2
#endif
#ifndef MODA_CHAIN_DISPATCH_BWD
#define MODA_CHAIN_DISPATCH_BWD 2
#endif
        if constexpr (PROG == P_FWD && MODA_CHAIN_DISPATCH_FWD == 2) {
          // every flavour of the forward programs inline and unrolled over its compile-time width
          MODA_FLAVOUR(HOT_FWD)
          MODA_FLAVOUR(HOT_FWD & ~E_MASK_OUT)
          if constexpr (BOX_ROWS == 128) {
            MODA_FLAVOUR(HOT_FWD | E_HEAD_SIGMA)
            MODA_FLAVOUR(E_BIAS | E_RELU | E_HEAD_SIGMA)                 // sigma-only program: last layer feeds the head only
            MODA_FLAVOUR(E_BIAS | E_SMEM)
            MODA_FLAVOUR_N(E_BIAS | E_RELU | E_HEAD_RGB | E_SMEM, 128)
            MODA_FLAVOUR_N(HOT_FWD, 128)                                 // the 5 x 128 program (nerf_feat)
            MODA_FLAVOUR_N(HOT_FWD, 64)
            MODA_FLAVOUR_N(E_BIAS | E_OUT_F32, 64)
            run_step<-1, 0, NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
          } else {
            MODA_FLAVOUR(E_BIAS | E_SMEM | E_LO)
            MODA_FLAVOUR(E_BIAS | E_OUT_F32)
            run_step<-1, 0, NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
          }
        } else if constexpr (PROG == P_BWD && MODA_CHAIN_DISPATCH_BWD == 2) {
          MODA_FLAVOUR(HOT_BWD)
          MODA_FLAVOUR(E_SMEM)
          if constexpr (BOX_ROWS == 128) {
            MODA_FLAVOUR(E_RANK1 | E_MASK_IN | E_SMEM)
            MODA_FLAVOUR_N(E_LOAD16 | E_SMEM, 128)
            MODA_FLAVOUR_N(E_SMEM, 64)
            MODA_FLAVOUR_N(E_ADD_SX | E_SMEM, 64)
            MODA_FLAVOUR_N(HOT_BWD, 128)                                 // the 5 x 128 program (nerf_feat)
            MODA_FLAVOUR_N(HOT_BWD, 64)
            MODA_FLAVOUR_N(E_LOAD32 | E_SMEM, 64)
            run_step<-1, 0, NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
          } else {
            MODA_FLAVOUR(E_LOAD32 | E_SMEM)
            MODA_FLAVOUR(E_ADD_SX | E_SMEM)
            run_step<-1, 0, NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
          }
        } else if constexpr (PROG == P_FWD && !MODA_CHAIN_DISPATCH_FWD) {
          MODA_FLAVOUR(HOT_FWD)
          MODA_FLAVOUR(HOT_FWD & ~E_MASK_OUT)                            // no sign bits wanted: inference, grid queries
          run_step<-1, 0, NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
        } else if constexpr (PROG == P_BWD && !MODA_CHAIN_DISPATCH_BWD) {
          MODA_FLAVOUR(HOT_BWD)
          run_step<-1, 0, NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
        } else if constexpr (PROG == P_FWD && BOX_ROWS == 128) {
          MODA_FLAVOUR(HOT_FWD)
          MODA_COLD(HOT_FWD | E_HEAD_SIGMA)
          MODA_COLD(E_BIAS | E_SMEM)                                   // xyz_encoding_final
          MODA_COLD(E_BIAS | E_RELU | E_HEAD_RGB | E_SMEM)             // dir layer + rgb head
          run_step_generic<NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
        } else if constexpr (PROG == P_FWD) {
          MODA_FLAVOUR(HOT_FWD)
          MODA_COLD(E_BIAS | E_SMEM | E_LO)
          MODA_COLD(E_BIAS | E_OUT_F32)
          run_step_generic<NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
        } else if constexpr (BOX_ROWS == 128) {
          MODA_FLAVOUR(HOT_BWD)
          MODA_COLD(E_SMEM)
          MODA_COLD(E_LOAD16 | E_SMEM)
          MODA_COLD(E_ADD_SX | E_SMEM)
          MODA_COLD(E_RANK1 | E_MASK_IN | E_SMEM)
          run_step_generic<NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
        } else {
          MODA_FLAVOUR(HOT_BWD)
          MODA_COLD(E_SMEM)
          MODA_COLD(E_LOAD32 | E_SMEM)
          MODA_COLD(E_ADD_SX | E_SMEM)
          run_step_generic<NH, ACC_STRIDE, EPI_THREADS>(flags, pg, maps, st, cx, acc_col, ready_slot, hs0, hs1, hs2, pm0, pm1);
        }
#undef MODA_COLD
#undef MODA_FLAVOUR_N
#undef MODA_FLAVOUR
        if (st.kc > 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_remote(acc_free2_remote + 8u * (uint32_t)b); else mbar_arrive(&acc_free[b]); }
          MODA_TR(cx.tr, cx.tr + 2, s, 0);
          ++mma_ctr;
        }
        if (pe_now) pe_write(xn);
        // head partial sums: the warps h >= 1 of a lane quarter hand theirs to warp h = 0 through shared memory
        // (summation order h = 0, 1, ..: as before)
        if (flags & E_HEAD_SIGMA) {
          if (h > 0) s_head_slot[((h - 1) * 128 + trow) * 4 + 3] = hs0;
          if (NH > 1) named_bar(bar_id, EPI_THREADS);
          if (h == 0) {
            sig_keep = pg.bs[0] + hs0;
#pragma unroll
            for (int k = 1; k < NH; ++k) sig_keep += s_head_slot[((k - 1) * 128 + trow) * 4 + 3];
            if (pg.sigma_only && cx.live) pg.raw[cx.row] = sig_keep;
          }
          if (pg.sigma_only && NH > 1) named_bar(bar_id, EPI_THREADS);   // s_head is reused by the next tile's partials
        }
        if (flags & E_HEAD_RGB) {
          if (h > 0) {
            float* mine = s_head_slot + ((h - 1) * 128 + trow) * 4;
            mine[0] = hs0; mine[1] = hs1; mine[2] = hs2;
          }
          if (NH > 1) named_bar(bar_id, EPI_THREADS);
          if (h == 0 && cx.live) {
            float o0 = pg.br[0] + hs0, o1 = pg.br[1] + hs1, o2 = pg.br[2] + hs2;
#pragma unroll
            for (int k = 1; k < NH; ++k) {
              const float* part = s_head_slot + ((k - 1) * 128 + trow) * 4;
              o0 += part[0]; o1 += part[1]; o2 += part[2];
            }
            float4 o;
            o.x = 1.0f / (1.0f + expf(-o0));
            o.y = 1.0f / (1.0f + expf(-o1));
            o.z = 1.0f / (1.0f + expf(-o2));
            o.w = sig_keep;
            *reinterpret_cast<float4*>(pg.raw + cx.row * 4) = o;
          }
          if (NH > 1) named_bar(bar_id, EPI_THREADS);   // s_head is reused by the next tile's sigma partials
        }
      }
    }
  } else if (!DUAL && warp < 2 + SLOTS * EPI_WARPS + PE_WARPS) {
    // ================================================================== positional-encoding producers
    // (two tile slots: the one producer warp serves them alternately, in the order the tiles start)
    // EPI_PE: the epilogue warps write the PE chunks; this warp only keeps the thread layout
    if (pg.pe_chunk >= 0 && !EPI_PE) {
      const int pw = warp - 2 - SLOTS * EPI_WARPS;
      const bool lead = (pw == 0) && lane == 0;
      (void)lead;
      int tr_n = 0;
      (void)tr_n;
      for (int j = 0; j < ntile; ++j) {
        const int slot = j % SLOTS, it = j / SLOTS;
        const int tile = (int)blockIdx.x + j * (int)gridDim.x;
        const uint32_t pe_base = smem_u32(sA) + (uint32_t)slot * slot_bytes + (uint32_t)(pg.pe_chunk * CHUNK_BYTES);
        // channels: [x(3) | w_k sin(2^k x)(3) | w_k cos(2^k x)(3)]_k, zero padded to 64 (nnutils/nerf.py:35-75);
        // values go straight to the swizzled chunk rows as they are produced (PE_ROWS rows per thread)
        const bool tr = pg.trace && blockIdx.x == 0 && it == 2 && slot == 0 && lead;
        (void)tr;
        MODA_TR(tr, 30, 0, 0);
        if (lead) s_dbg[3] = j;
        // this thread's PE_ROWS consecutive rows of xyz: fetched before the wait, so the global-load latency hides
        // behind it
        float xs[PE_ROWS][3];
#pragma unroll
        for (int rr = 0; rr < PE_ROWS; ++rr) {
          const long long row = (long long)tile * TILE_M + pw * 32 * PE_ROWS + rr * 32 + lane;
          const bool in = row < pg.M;
#pragma unroll
          for (int c = 0; c < 3; ++c) xs[rr][c] = in ? __ldg(pg.xyz + row * 3 + c) : 0.f;
        }
        wait_or_trap<0>(&pe_free[slot], (it & 1) ^ 1);   // the slot's previous tile's last reader of the PE chunk(s) has completed
        MODA_TR(tr, 31, 0, 0);
#pragma unroll 1
        for (int rr = 0; rr < PE_ROWS; ++rr) {
          // lanes own consecutive rows: their swizzle phases differ, so the 128-bit stores below are conflict-free
          // (the earlier lane -> 4 consecutive rows layout with 32-bit stores was a 16-way bank conflict, ~5k shared-
          // memory wavefronts per tile taken from the tensor core's operand fetches)
          const int trow = pw * 32 * PE_ROWS + rr * 32 + lane;
          const int sw = trow & 7;
          float x[3] = {xs[0][0], xs[0][1], xs[0][2]};
#pragma unroll
          for (int j = 1; j < PE_ROWS; ++j)   // register select (the row loop stays rolled: code size)
            if (rr == j) { x[0] = xs[j][0]; x[1] = xs[j][1]; x[2] = xs[j][2]; }
          pe_row<BOX_ROWS>(pg, pe_base + (uint32_t)(trow * 128), sw, x);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ready[slot * MAX_CHUNKS + pg.pe_chunk]);
          if (pg.pe_lo) mbar_arrive(&ready[slot * MAX_CHUNKS + pg.pe_chunk + 1]);
          if (PAIR) mbar_arrive_remote(mapa_u32(smem_u32(&ready2[slot * MAX_CHUNKS + pg.pe_chunk]), 0));
        }
      }
    }
  } else {
    // ================================================================== store warp
    // Saves the chunks the other pass needs with TMA, in the order they get written.  The chunk barriers are
    // shared with the MMA thread (multiple waiters); the PE chunk's save is safe against the next tile's rewrite
    // because pe_free is committed steps later, after the MMA thread has seen this step's stores_done.
    if (lane == 0) {
      const uint32_t sA_u32 = smem_u32(sA);
      int tr_n = 0;
      (void)tr_n;
      for (int it = 0; it * SLOTS < ntile; ++it) {
        const int nact = (ntile - it * SLOTS < SLOTS) ? ntile - it * SLOTS : SLOTS;
        int d0 = 0;
        for (int s = 0; s < pg.nsteps; ++s) {
          int d_end = d0;
          // two tile slots: the slots' epilogues of a step run one after the other (the MMA thread alternates), so
          // their chunks are served in that order
          for (int slot = 0; slot < nact; ++slot) {
            const int tile = (int)blockIdx.x + (it * SLOTS + slot) * (int)gridDim.x;
            bool any = false, work = false;
            int d = d0;
            while (d < pg.nduty && pg.duty[d].step == s) {
              work = true;
              const int c = pg.duty[d].chunk;
              const uint32_t gen = (uint32_t)it * (uint32_t)pg.wpt[c] + pg.duty[d].gen;
              s_dbg[4] = it * 1000 + s * 10 + (d % 10); s_dbg[5] = 0;
              wait_or_trap<0>(&ready[slot * MAX_CHUNKS + c], gen & 1);
              s_dbg[5] = 1;
              MODA_TR(pg.trace && blockIdx.x == 0 && it == 2 && slot == 0, 20, s, c);
#ifdef MODA_EXP_NO_STORE
              if (false) {
#else
              if (pg.duty[d].map >= 0) {
#endif
#ifdef MODA_EXP_STORE_L2
                const int trow0 = (tile & 63) * TILE_M;   // experiment: every store lands in the same L2-resident rows
#else
                const int trow0 = tile * TILE_M;
#endif
                tma_store_2d_u32(&maps.save[pg.duty[d].map], sA_u32 + (uint32_t)slot * slot_bytes + MODA_CHUNK_OFF(c, BOX_ROWS == 128),
                                 (int)pg.duty[d].col * 64, trow0);
                bulk_commit();
                any = true;
              }
              ++d;
            }
            d_end = d;
            s_dbg[5] = 2;
            if (any) bulk_wait_read0();
            MODA_TR(pg.trace && blockIdx.x == 0 && it == 2 && slot == 0 && any, 21, s, 0);
            if (work) { if (PAIR) mbar_arrive_remote(mapa_u32(smem_u32(&stores_done2[slot]), 0)); else mbar_arrive(&stores_done[slot]); }
            s_dbg[5] = 3;
          }
          d0 = d_end;
        }
      }
      bulk_wait_all0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // no CTA of the pair exits while the other may still signal it or use its shared memory
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (PAIR) cluster_sync_all();
}

}  // namespace chain
}  // namespace moda

// ====================================================================================================== host side
#include "tc_host.cuh"

using namespace moda;
using namespace moda::chain;

#include <atomic>
// Launch mode of the 256-wide chains is a PER-CALL argument (`mode`): bit 0 = CTA pairs (tcgen05 cta_group::2), bit 1 =
// two tiles in flight per CTA (needs bit 0).  The only process-wide state is (a) a sticky flag that a cluster launch
// failed once on this process's device/partition (then every later call uses the single-CTA kernels) and (b) the debug
// trace hook below, both atomics.
// bit 2 = xyz_encoding_final folded into dir_encoding: the reference applies final (no activation) and then the direction
// layer to [final | dir | env] (nerf.py:182-190), i.e. two linear maps in a row; the caller hands in the product weights
// W' = Wdir[:, :256] Wfinal (and the bias Wdir[:, :256] bfinal inside the per-ray bias), and the programs run one
// 256-wide step less per pass (and save / re-read one (P,256) activation and one gradient less).
constexpr int MODE_PAIR = 1, MODE_TWO_SLOTS = 2, MODE_FOLD_FINAL = 4;
static std::atomic<int> g_pair_unavailable{0};
static std::atomic<long long*> g_trace{nullptr};
constexpr int PAIR_UNAVAILABLE = -77;   // launch<..., PAIR = 1> could not place the cluster: the caller relaunches single-CTA
extern "C" int moda_chain_pair_available(void) { return g_pair_unavailable.load() ? 0 : 1; }
// debug: device buffer of >= 4 + 4 * 4000 int64 (zeroed by the caller) that the next chain launches fill with a timeline
extern "C" int moda_chain_set_trace(long long* buf) { g_trace.store(buf); return 0; }

namespace {

struct Builder {
  Program pg;
  Maps maps;
  int nsave = 0;
  int gen[MAX_CHUNKS];   // writes issued so far to each chunk within the tile
  int by_mma_step[MAX_CHUNKS];   // the latest write of the chunk comes from the epilogue of an MMA step (not PE / load)
  int err = 0;

  Builder() {
    memset(&pg, 0, sizeof(pg));
    memset(&maps, 0, sizeof(maps));
    memset(gen, 0, sizeof(gen));
    memset(by_mma_step, 0, sizeof(by_mma_step));
    pg.pe_chunk = -1;
    pg.pe_save_map = -1;
    pg.vec_off0 = pg.vec_off1 = -1;
  }
  // registers an fp16 (rows, cols) output for TMA stores; returns its descriptor index, -1 when ptr is null
  int save(const void* ptr, long long rows, int cols) {
    if (!ptr) return -1;
    if (nsave >= MAX_SAVE_MAPS) { set_error("chain: too many saved outputs"); err = -1; return -1; }
    if (int e = make_map(&maps.save[nsave], ptr, rows, cols, cols, TILE_M)) { err = e; return -1; }
    return nsave++;
  }
  Step& add(int n, int flags) {
    Step& st = pg.st[pg.nsteps++];
    st.n = n; st.flags = flags; st.kc = 0;
    st.out_chunk = st.out_lo_chunk = st.save_map = -1;
    st.mask_slot = 0; st.release_pe = 0; st.bias_off = 0; st.bias = nullptr; st.rowbias = nullptr; st.tile_bias = 0;
    return st;
  }
  // K chunk: A from shared-memory chunk `a` (its latest write), B from packed-weight chunk `bcol`
  void k(Step& st, int a, int bcol) {
    st.a_info[st.kc] = (unsigned int)a | ((unsigned int)(gen[a] - 1) << 8) | ((unsigned int)(by_mma_step[a] ? 1 : 0) << 16);
    st.b_col[st.kc] = (unsigned short)bcol;
    ++st.kc;
  }
  // the step's epilogue writes `count` chunks starting at `first` (and optionally their lo halves)
  void add_duty(int chunk, int g, int step, int col, int map) {
    if (pg.nduty >= MAX_DUTY) { set_error("chain: store work list overflow"); err = -1; return; }
    Program::Duty& d = pg.duty[pg.nduty++];
    d.chunk = (unsigned char)chunk; d.gen = (unsigned char)g; d.step = (unsigned char)step;
    d.col = (unsigned char)col; d.map = (signed char)map;
  }
  // call after st.save_map is set: the step's epilogue writes `count` chunks starting at `first` (and optionally
  // their lo halves)
  void out(Step& st, int first, int count, int lo_first = -1) {
    st.out_chunk = first;
    st.out_lo_chunk = lo_first;
    st.flags |= E_SMEM | (lo_first >= 0 ? E_LO : 0);
    const int step = (int)(&st - pg.st);
    const int mma_step = (st.flags & (E_LOAD16 | E_LOAD32)) ? 0 : 1;   // k() calls precede out() for MMA steps
    for (int i = 0; i < count; ++i) {
      add_duty(first + i, gen[first + i], step, i, st.save_map);
      ++gen[first + i];
      by_mma_step[first + i] = mma_step;
      if (lo_first >= 0) { ++gen[lo_first + i]; by_mma_step[lo_first + i] = mma_step; }
    }
  }
  void pe(int chunk, bool lo, int save_map) {
    pg.pe_chunk = chunk; pg.pe_lo = lo ? 1 : 0; pg.pe_save_map = save_map;
    pg.from_pe[chunk] = 1;
    add_duty(chunk, gen[chunk], 0, 0, save_map);
    ++gen[chunk];
    by_mma_step[chunk] = 0;
    if (lo) { pg.from_pe[chunk + 1] = 1; ++gen[chunk + 1]; by_mma_step[chunk + 1] = 0; }
  }
  // slots: tile slots of the kernel that will run the program (per-tile bias rows get one copy per slot).  Idempotent.
  void finish(int slots) {
    for (int i = 0; i < MAX_CHUNKS; ++i) pg.wpt[i] = gen[i];
    int phases = 0, d = 0;
    for (int s = 0; s < pg.nsteps; ++s) {
      pg.st[s].sd_wait = phases;
      bool work = false;
      while (d < pg.nduty && pg.duty[d].step == s) { work = true; ++d; }
      if (work) ++phases;
    }
    pg.sd_per_tile = phases;
    // bias / head-vector table: regions packed back to back (every width is a multiple of 64 floats, so the 128-bit
    // reads of the epilogue stay aligned)
    int off = 0;
    for (int s = 0; s < pg.nsteps; ++s) {
      Step& st = pg.st[s];
      if (st.bias) { st.bias_off = off; off += st.n; st.flags |= E_BIAS; }
      if (st.rowbias) {
        if (pg.rep % TILE_M == 0) { st.bias_off = off; off += slots * st.n; st.flags |= E_BIAS; st.tile_bias = 1; pg.has_tile_bias = 1; }
        else st.flags |= E_ROWBIAS;
      }
    }
    if (pg.vec0) { pg.vec_off0 = off; off += (pg.vec_len0 + 63) & ~63; }
    if (pg.vec1) { pg.vec_off1 = off; off += (pg.vec_len1 + 63) & ~63; }
    pg.bias_floats = off;
  }
};

template <int BOX_ROWS, int EPI_WARPS, int PE_WARPS, int MIN_CTAS, int PROG, int PAIR = 0, int SLOTS = 1>
int launch(Builder& bld, const void* wpack, int wrows, int wcols, cudaStream_t stream) {
  if (bld.err) return bld.err;
  Builder b = bld;   // the caller's program stays untouched: a fallback relaunches it in another mode
  b.finish(SLOTS);
  b.pg.trace = g_trace.load();
  b.pg.mask_tiles = (b.pg.num_tiles + 1) & ~1;
  if (PAIR) b.pg.num_tiles = (b.pg.num_tiles + 1) & ~1;   // both CTAs of a pair run the same number of tiles
  constexpr int STAGE_BYTES = PAIR ? BOX_ROWS * 128 : ((BOX_ROWS == 128) ? 2 * BOX_ROWS * 128 : BOX_ROWS * 128);
  constexpr int THREADS = chain_threads(SLOTS, EPI_WARPS, PE_WARPS);
  const size_t cap = 232448 / MIN_CTAS - (MIN_CTAS > 1 ? 1024 : 0);   // 1 KB per CTA is reserved by the system
  // weight ring: as many stages as fit (16 KB stages for CTA pairs, 32 KB otherwise), at most the program's request
  const size_t fixed = 1024 + (size_t)SLOTS * MODA_NCHUNKS(b.pg.nchunks, BOX_ROWS == 128) * CHUNK_BYTES + (size_t)head_bytes(PROG, SLOTS, EPI_WARPS / 4) +
                       (size_t)b.pg.bias_floats * 4 + 512 + 32;
  MODA_REQUIRE(fixed + 2 * STAGE_BYTES <= cap, "chain: needs %zu B of shared memory (limit %zu)", fixed + 2 * STAGE_BYTES, cap);
  int stages = (int)((cap - fixed) / STAGE_BYTES);
  if (stages > b.pg.stages) stages = b.pg.stages;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  b.pg.stages = stages;
  const size_t smem = fixed + (size_t)stages * STAGE_BYTES;
  if (int e = make_map(&b.maps.w, wpack, wrows, wcols, wcols, PAIR ? 32 : BOX_ROWS)) return e;
  auto kern = chain_kernel<BOX_ROWS, EPI_WARPS, PE_WARPS, MIN_CTAS, PROG, PAIR, SLOTS>;
  // per-device function attribute: set on every launch (cheap), so a second device in the same process gets it too
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
  const int slots = PAIR ? ((sm_count() * MIN_CTAS) & ~1) : sm_count() * MIN_CTAS;
  const int grid = b.pg.num_tiles < slots ? b.pg.num_tiles : slots;
  if (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, b.pg, b.maps);
    if (e != cudaSuccess) {
      // a device / partition that cannot co-schedule the CTA pair: fall back to the single-CTA kernels for good
      (void)cudaGetLastError();
      g_pair_unavailable.store(1);
      return PAIR_UNAVAILABLE;
    }
  } else {
    kern<<<grid, THREADS, smem, stream>>>(b.pg, b.maps);
  }
  return check_launch("chain");
}

// launches a 256-wide program in the requested mode (see MODE_*), falling back to the single-CTA kernel when a cluster
// cannot be placed.  Two tile slots only pay when every CTA gets at least two tiles.
template <int PROG>
int launch_trunk(Builder& b, int mode, const void* wpack, int wcols, cudaStream_t stream, int wrows = 256) {
  const bool pair = (mode & MODE_PAIR) && !g_pair_unavailable.load();
  if (pair) {
    const int ctas = sm_count() & ~1;
    const int tiles = (b.pg.num_tiles + 1) & ~1;
    int e;
    if ((mode & MODE_TWO_SLOTS) && tiles >= 2 * ctas) {
      b.pg.stages = 5;   // as many 16 KB stages as fit: 3 for the forward programs, 4 for the adjoint
      e = launch<128, MODA_TRUNK_EPI, 1, 1, PROG, 1, 2>(b, wpack, wrows, wcols, stream);
    } else {
      b.pg.stages = 8;
      e = launch<128, MODA_TRUNK_EPI, 1, 1, PROG, 1, 1>(b, wpack, wrows, wcols, stream);
    }
    if (e != PAIR_UNAVAILABLE) return e;
  }
  b.pg.stages = 4;   // single-CTA weight ring: 4 stages of 32 KB
  return launch<128, MODA_TRUNK_EPI, 1, 1, PROG, 0, 1>(b, wpack, wrows, wcols, stream);
}

// the epilogue reads these operands with 128-bit loads
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

void fill_win(Program& pg, int F, const float* win) {
  pg.F = F;
  for (int i = 0; i < 10; ++i) pg.win[i] = (win && i < F) ? win[i] : 1.0f;
}

}  // namespace

// ---------------------------------------------------------------------------------------------- nerf_coarse
// Packed weights wpack: fp16 (256, 38*64), 64-column chunks in this order (rows = output channel, zero padded):
//   0: W1[:, :63]   1-4: W2   5-8: W3   9-12: W4   13: W5[:, :63]   14-17: W5[:, 63:]   18-21: W6   22-25: W7
//   26-29: W8   30-33: Wfinal   34-37: Wdir[:, :256] (128 rows)
// biases: b1..b8, bfinal.  rowbias (P / rep, 128) = Wdir[:, 256:] [dir | env] + bdir per ray.
// Saved for the adjoint / weight gradients when non-null: A0 (P,64) = fp16 PE, H (8,P,256), fin (P,256),
// dfe (P,128), maskbits (8, tiles, 8, 128) uint32 sign bits of H.  raw (P,4) = [sigmoid(rgb) | sigma].
extern "C" int moda_chain_trunk_fwd(const float* xyz, long long P, int rep, int F, const float* win, const void* wpack,
                                    const float* const* biases, const float* rowbias, const float* ws, const float* bs,
                                    const float* Wr, const float* br, void* A0, void* H, void* fin, void* dfe,
                                    unsigned int* maskbits, float* raw, int mode, cudaStream_t stream) {
  if (P == 0) return 0;
  MODA_REQUIRE(xyz && wpack && biases && rowbias && ws && bs && Wr && br && raw && rep > 0 && F >= 0 && F <= 10,
               "chain_trunk_fwd: bad arguments");
  MODA_REQUIRE(al16(rowbias) && al16(raw) && al16(wpack), "chain_trunk_fwd: rowbias, raw and wpack must be 16-byte aligned");
  Builder b;
  Program& pg = b.pg;
  pg.M = P; pg.num_tiles = (int)((P + TILE_M - 1) / TILE_M); pg.rep = rep; pg.nchunks = 5;
  pg.xyz = xyz; fill_win(pg, F, win);
  pg.ws = ws; pg.bs = bs; pg.Wr = Wr; pg.br = br; pg.raw = raw; pg.maskbits = maskbits;
  pg.vec0 = ws; pg.vec_len0 = 256; pg.vec1 = Wr; pg.vec_len1 = 3 * 128;
  const int PE = 4;
  const int order[4] = {0, 1, 2, 3};   // the epilogue finishes the chunks of a result in this order
  b.pe(PE, false, b.save(A0, P, 64));
  int col = 0;
  for (int l = 0; l < 8; ++l) {
    Step& st = b.add(256, E_RELU | (maskbits ? E_MASK_OUT : 0));
    st.bias = biases[l]; st.mask_slot = l;
    if (l == 0) { b.k(st, PE, col); col += 1; }
    else {
      if (l == 4) { b.k(st, PE, col); col += 1; st.release_pe = 1; }
      for (int i = 0; i < 4; ++i) b.k(st, order[i], col + order[i]);
      col += 4;
    }
    if (l == 7) st.flags |= E_HEAD_SIGMA;
    st.save_map = b.save(H ? (const char*)H + (size_t)l * P * 256 * 2 : nullptr, P, 256);
    b.out(st, 0, 4);
  }
  if (!(mode & MODE_FOLD_FINAL)) {
    Step& st = b.add(256, 0);           // xyz_encoding_final (no activation)
    st.bias = biases[8];
    for (int i = 0; i < 4; ++i) b.k(st, order[i], col + order[i]);
    col += 4;
    st.save_map = b.save(fin, P, 256);
    b.out(st, 0, 4);
  }
  {
    // dir_encoding on [fin | per-ray constant part as a bias]; folded: on the layer-8 activations with W' (chunks 30-33)
    Step& st = b.add(128, E_RELU | E_HEAD_RGB);
    st.rowbias = rowbias;
    for (int i = 0; i < 4; ++i) b.k(st, order[i], col + order[i]);
    col += 4;
    st.save_map = b.save(dfe, P, 128);
    b.out(st, 0, 2);
  }
  return launch_trunk<P_FWD>(b, mode, wpack, col * 64, stream);
}

// Density only (the grid query of mesh extraction, nnutils/train_utils.py:1377-1404 -> nerf.py:176-180 with
// sigma_only=True): the first 30 weight chunks of the forward packing, layers 1-8 and the sigma head; nothing is
// saved.  sigma (P) fp32.
extern "C" int moda_chain_trunk_sigma(const float* xyz, long long P, int F, const float* win, const void* wpack,
                                      const float* const* biases, const float* ws, const float* bs, float* sigma,
                                      int mode, cudaStream_t stream) {
  if (P == 0) return 0;
  MODA_REQUIRE(xyz && wpack && biases && ws && bs && sigma && F >= 0 && F <= 10, "chain_trunk_sigma: bad arguments");
  MODA_REQUIRE(al16(wpack) && al16(ws), "chain_trunk_sigma: wpack and ws must be 16-byte aligned");
  Builder b;
  Program& pg = b.pg;
  pg.M = P; pg.num_tiles = (int)((P + TILE_M - 1) / TILE_M); pg.rep = 1; pg.nchunks = 5;
  pg.xyz = xyz; fill_win(pg, F, win);
  pg.ws = ws; pg.bs = bs; pg.raw = sigma; pg.sigma_only = 1;
  pg.vec0 = ws; pg.vec_len0 = 256;
  const int PE = 4;
  b.pe(PE, false, -1);
  int col = 0;
  for (int l = 0; l < 8; ++l) {
    Step& st = b.add(256, E_RELU);
    st.bias = biases[l];
    if (l == 0) { b.k(st, PE, col); col += 1; }
    else {
      if (l == 4) { b.k(st, PE, col); col += 1; st.release_pe = 1; }
      for (int i = 0; i < 4; ++i) b.k(st, i, col + i);
      col += 4;
    }
    if (l == 7) st.flags |= E_HEAD_SIGMA;   // the last layer's activations only feed the head: not written back
    else b.out(st, 0, 4);
  }
  return launch_trunk<P_FWD>(b, mode, wpack, 38 * 64, stream);
}

// Adjoint chain of nerf_coarse.  Packed transposed weights wpackT: fp16 (256, 42*64); rows = input channel of the
// layer (the N of the data-gradient GEMM), columns = its output channel (K):
//   0-1: Wdir[:, :256]^T (K = 128)   2-5: Wfinal^T   6-9: W8^T   10-13: W7^T   14-17: W6^T
//   18-21: W5[:, :63]^T (64 rows)   22-25: W5[:, 63:]^T   26-29: W4^T   30-33: W3^T   34-37: W2^T
//   38-41: W1[:, :63]^T (64 rows)
// In: d_dfe (P,128) fp16 = scaled gradient at the dir layer's pre-activation (already ReLU-masked), gsig (P),
// ws (256), rscale (device scalar): dY[7] gets + gsig * ws * rscale before its mask.
// Out (fp16): d_fin (P,256), dY (8,P,256) with dY[i] = gradient at layer i's pre-activation, d_pe (P,64).
extern "C" int moda_chain_trunk_bwd(const void* d_dfe, const float* gsig, const float* ws, const float* rscale,
                                    const void* wpackT, const unsigned int* maskbits, long long P, void* d_fin,
                                    void* dY, void* d_pe, int mode, cudaStream_t stream) {
  if (P == 0) return 0;
  const bool fold = (mode & MODE_FOLD_FINAL) != 0;   // wpackT: chunks 0-1 = W'^T, then W8^T ...: no d_fin step
  MODA_REQUIRE(d_dfe && gsig && ws && wpackT && maskbits && (d_fin || fold) && dY && d_pe, "chain_trunk_bwd: null pointer");
  MODA_REQUIRE(al16(d_dfe) && al16(wpackT), "chain_trunk_bwd: d_dfe and wpackT must be 16-byte aligned");
  Builder b;
  Program& pg = b.pg;
  pg.M = P; pg.num_tiles = (int)((P + TILE_M - 1) / TILE_M); pg.rep = 1; pg.nchunks = 5;
  pg.gsig = gsig; pg.cvec = ws; pg.rscale = rscale; pg.maskbits = const_cast<unsigned int*>(maskbits);
  pg.vec0 = ws; pg.vec_len0 = 256;
  pg.load_src = d_dfe; pg.load_ld = 128; pg.load_cols = 128;
  const int SX = 4;
  const int order[4] = {0, 1, 2, 3};
  auto dy = [&](int i) { return (const char*)dY + (size_t)i * P * 256 * 2; };
  { Step& st = b.add(128, E_LOAD16); b.out(st, 0, 2); }
  if (!fold) {
    Step& st = b.add(256, 0);                       // d_fin = d_dfe Wdir[:, :256]
    b.k(st, 0, 0); b.k(st, 1, 1);
    st.save_map = b.save(d_fin, P, 256);
    b.out(st, 0, 4);
  }
  {
    Step& st = b.add(256, E_RANK1 | E_MASK_IN);     // dY[7] = (d_fin Wfinal + gsig ws) . [H8 > 0]; folded: d_dfe W'
    if (fold) { b.k(st, 0, 0); b.k(st, 1, 1); }
    else for (int i = 0; i < 4; ++i) b.k(st, order[i], 2 + order[i]);
    st.mask_slot = 7; st.save_map = b.save(dy(7), P, 256);
    b.out(st, 0, 4);
  }
  int col = fold ? 2 : 6;
  for (int l = 7; l >= 1; --l) {
    if (l == 4) {
      Step& sx = b.add(64, 0);                      // dPE partial = dY[4] W5[:, :63], parked in the SX chunk
      for (int i = 0; i < 4; ++i) b.k(sx, order[i], col + order[i]);
      col += 4;
      b.out(sx, SX, 1);
    }
    Step& st = b.add(256, E_MASK_IN);               // dY[l-1] = (dY[l] W_{l+1}[:, hidden]) . [H_l > 0]
    for (int i = 0; i < 4; ++i) b.k(st, order[i], col + order[i]);
    col += 4;
    st.mask_slot = l - 1; st.save_map = b.save(dy(l - 1), P, 256);
    b.out(st, 0, 4);
  }
  {
    Step& st = b.add(64, E_ADD_SX);                 // dPE = dY[0] W1[:, :63] + partial
    for (int i = 0; i < 4; ++i) b.k(st, order[i], col + order[i]);
    col += 4;
    st.save_map = b.save(d_pe, P, 64);
    b.out(st, SX, 1);
  }
  return launch_trunk<P_BWD>(b, mode, wpackT, col * 64, stream);
}

// ------------------------------------------------------------------------------------------- 5 x 128 raw-feature MLP
// nerf_feat (nnutils/moda.py:447-449: NeRF(D=5, W=128, in 63, raw_feat, out 16)) on a plain PE(xyz) input, on the
// 256-wide engine (128-row weight boxes, CTA pairs / two tile slots as requested by `mode`), plain fp16 operands, the
// final layer always folded into the direction layer (see MODE_FOLD_FINAL).
// wpack fp16 (128, 13*64), 64-column chunks (rows = output channel, zero padded):
//   0: W1[:, :63]   1-2: W2   3-4: W3   5-6: W4   7: W5[:, :63]   8-9: W5[:, 63:]   10-11: W' = Wdir Wfinal (64 rows)
//   12: Wrgb (out_channels rows, K = 64)
// biases: b1..b5 (128), b' = bdir + Wdir bfinal (64), brgb zero padded to 64.  y32 (P,32) fp32, columns >= out_channels 0.
// Saved when non-null: A0 (P,64), H (5,P,128), dfe (P,64), maskbits (6, tiles2, 4, 128) u64.
extern "C" int moda_chain_feat_fwd(const float* xyz, long long P, int F, const float* win, const void* wpack,
                                   const float* const* biases, void* A0, void* H, void* dfe, unsigned int* maskbits,
                                   float* y32, int mode, cudaStream_t stream) {
  if (P == 0) return 0;
  MODA_REQUIRE(xyz && wpack && biases && y32 && F >= 0 && F <= 10, "chain_feat_fwd: bad arguments");
  MODA_REQUIRE(al16(y32) && al16(wpack), "chain_feat_fwd: y32 and wpack must be 16-byte aligned");
  Builder b;
  Program& pg = b.pg;
  pg.M = P; pg.num_tiles = (int)((P + TILE_M - 1) / TILE_M); pg.rep = 1; pg.nchunks = 3;
  pg.xyz = xyz; fill_win(pg, F, win);
  pg.maskbits = maskbits; pg.y32 = y32; pg.ld_y32 = 32;
  const int PE = 2;
  const int mo = maskbits ? E_MASK_OUT : 0;
  b.pe(PE, false, b.save(A0, P, 64));
  int col = 0;
  for (int l = 0; l < 5; ++l) {
    Step& st = b.add(128, E_RELU | mo);
    st.bias = biases[l]; st.mask_slot = l;
    if (l == 0) { b.k(st, PE, col); col += 1; }
    else {
      if (l == 4) { b.k(st, PE, col); col += 1; st.release_pe = 1; }
      b.k(st, 0, col); b.k(st, 1, col + 1); col += 2;
    }
    st.save_map = b.save(H ? (const char*)H + (size_t)l * P * 128 * 2 : nullptr, P, 128);
    b.out(st, 0, 2);
  }
  { Step& st = b.add(64, E_RELU | mo); st.bias = biases[5]; st.mask_slot = 5; b.k(st, 0, col); b.k(st, 1, col + 1); col += 2;
    st.save_map = b.save(dfe, P, 64); b.out(st, 0, 1); }
  { Step& st = b.add(64, E_OUT_F32); st.bias = biases[6]; b.k(st, 0, col); col += 1; }
  return launch_trunk<P_FWD>(b, mode, wpack, col * 64, stream, 128);
}

// Adjoint of the above.  wpackT fp16 (128, 14*64): 0: Wrgb^T (64 dfe rows, K = out_channels)   1: W'^T (128 rows, K = 64)
//   2-3: W5[:, :63]^T (64 rows, K = 128)   4-5: W5[:, 63:]^T   6-7: W4^T   8-9: W3^T   10-11: W2^T   12-13: W1[:, :63]^T
// In: gout (P,32) fp32, scale (device scalar), maskbits.  Out fp16: G (P,64) = scale gout, d_dfe (P,64), dY (5,P,128),
// d_pe (P,64).
extern "C" int moda_chain_feat_bwd(const float* gout, const float* scale, const void* wpackT, const unsigned int* maskbits,
                                   long long P, void* G, void* d_dfe, void* dY, void* d_pe, int mode, cudaStream_t stream) {
  if (P == 0) return 0;
  MODA_REQUIRE(gout && wpackT && maskbits && G && d_dfe && dY && d_pe, "chain_feat_bwd: null pointer");
  MODA_REQUIRE(al16(gout) && al16(wpackT), "chain_feat_bwd: gout and wpackT must be 16-byte aligned");
  Builder b;
  Program& pg = b.pg;
  pg.M = P; pg.num_tiles = (int)((P + TILE_M - 1) / TILE_M); pg.rep = 1; pg.nchunks = 3;
  pg.maskbits = const_cast<unsigned int*>(maskbits);
  pg.load_src = gout; pg.load_ld = 32; pg.load_cols = 32; pg.load_scale = scale;
  const int SX = 2;
  auto dy = [&](int i) { return (const char*)dY + (size_t)i * P * 128 * 2; };
  { Step& st = b.add(64, E_LOAD32); st.save_map = b.save(G, P, 64); b.out(st, 0, 1); }
  { Step& st = b.add(64, E_MASK_IN); b.k(st, 0, 0); st.mask_slot = 5; st.save_map = b.save(d_dfe, P, 64); b.out(st, 0, 1); }
  { Step& st = b.add(128, E_MASK_IN); b.k(st, 0, 1); st.mask_slot = 4; st.save_map = b.save(dy(4), P, 128); b.out(st, 0, 2); }
  { Step& st = b.add(64, 0); b.k(st, 0, 2); b.k(st, 1, 3); b.out(st, SX, 1); }
  int col = 4;
  for (int l = 4; l >= 1; --l) {
    Step& st = b.add(128, E_MASK_IN);
    b.k(st, 0, col); b.k(st, 1, col + 1); col += 2;
    st.mask_slot = l - 1; st.save_map = b.save(dy(l - 1), P, 128);
    b.out(st, 0, 2);
  }
  { Step& st = b.add(64, E_ADD_SX); b.k(st, 0, col); b.k(st, 1, col + 1); col += 2; st.save_map = b.save(d_pe, P, 64); b.out(st, SX, 1); }
  return launch_trunk<P_BWD>(b, mode, wpackT, col * 64, stream, 128);
}

// ------------------------------------------------------------------------------------------------ nerf_skin
// Split precision: every operand is an fp16 (hi, lo) pair and  x W^T ~ hi Whi^T + lo Whi^T + hi Wlo^T.
// Packed weights wpack: fp16 (64, 18*64); per layer the chunks [Whi | Wlo] (rows = output channel, zero padded to 64):
//   0-1: W1[:, :63]   2-3: W2   4-5: W3   6-7: W4   8-9: W5[:, :63]   10-11: W5[:, 63+nc:]   12-13: Wfinal
//   14-15: Wdir (32 rows)   16-17: Wrgb (K = 32, rows = out_channels)
// biases[8]: rb1 (rowbias (P/rep,64): pose-code part of layer 1 + b1), b2, b3, b4, rb5 (rowbias), bfinal,
// bdir (padded to 64), brgb (padded to 64).  Out: y32 (P,32) fp32 delta logits; saved (hi halves, fp16) when
// non-null: A0 (P,64), H (5,P,64), fin (P,64), dfe (P,64), maskbits (6, tiles, 2, 128): H1..H5, dfe.
extern "C" int moda_chain_skin_fwd(const float* xyz, long long P, int rep, int F, const float* win, const void* wpack,
                                   const float* const* biases, void* A0, void* H, void* fin, void* dfe,
                                   unsigned int* maskbits, float* y32, int fold, cudaStream_t stream) {
  if (P == 0) return 0;
  MODA_REQUIRE(xyz && wpack && biases && y32 && rep > 0 && F >= 0 && F <= 10, "chain_skin_fwd: bad arguments");
  MODA_REQUIRE(al16(biases[0]) && al16(biases[4]) && al16(y32) && al16(wpack),
               "chain_skin_fwd: row biases, y32 and wpack must be 16-byte aligned");
  Builder b;
  Program& pg = b.pg;
  pg.M = P; pg.num_tiles = (int)((P + TILE_M - 1) / TILE_M); pg.rep = rep; pg.nchunks = 4; pg.stages = 4;
  pg.xyz = xyz; fill_win(pg, F, win);
  pg.maskbits = maskbits; pg.y32 = y32; pg.ld_y32 = 32;
  const int PEH = 0, PEL = 1, AH = 2, AL = 3;
  b.pe(PEH, true, b.save(A0, P, 64));
  const int mo = maskbits ? E_MASK_OUT : 0;
  auto split = [&](Step& st, int hi, int lo, int wc) { b.k(st, hi, wc); b.k(st, lo, wc); b.k(st, hi, wc + 1); };
  for (int l = 0; l < 5; ++l) {
    Step& st = b.add(64, E_RELU | mo);
    st.mask_slot = l;
    if (l == 0) { st.rowbias = biases[0]; split(st, PEH, PEL, 0); }
    else if (l == 4) { st.rowbias = biases[4]; split(st, PEH, PEL, 8); split(st, AH, AL, 10); st.release_pe = 1; }
    else { st.bias = biases[l]; split(st, AH, AL, 2 * l); }
    st.save_map = b.save(H ? (const char*)H + (size_t)l * P * 64 * 2 : nullptr, P, 64);
    b.out(st, AH, 1, AL);
  }
  // fold: xyz_encoding_final folded into dir_encoding (see MODE_FOLD_FINAL): wpack holds W' in place of Wfinal and
  // Wdir (16 chunks), biases[5] is ignored and biases[6] = bdir + Wdir bfinal
  int wc = 12;
  if (!fold) {
    Step& st = b.add(64, 0); st.bias = biases[5]; split(st, AH, AL, wc); wc += 2;
    st.save_map = b.save(fin, P, 64); b.out(st, AH, 1, AL);
  }
  { Step& st = b.add(64, E_RELU | mo); st.mask_slot = 5; st.bias = biases[6]; split(st, AH, AL, wc); wc += 2;
    st.save_map = b.save(dfe, P, 64); b.out(st, AH, 1, AL); }
  { Step& st = b.add(64, E_OUT_F32); st.bias = biases[7]; split(st, AH, AL, wc); wc += 2; }
  return launch<64, MODA_SKIN_EPI, 1, 2, P_FWD>(b, wpack, 64, wc * 64, stream);
}

// Adjoint chain of nerf_skin on plain fp16 operands.  wpackT: fp16 (64, 9*64), rows = input channel, cols = output:
//   0: Wrgb^T (rows: 32 dfe channels; K: out_channels)   1: Wdir^T (K = 32)   2: Wfinal^T   3: W5[:, :63]^T
//   4: W5[:, 63+nc:]^T   5: W4^T   6: W3^T   7: W2^T   8: W1[:, :63]^T
// In: gout (P,32) fp32 (columns >= out_channels zero), scale (device scalar), maskbits from the forward.
// Out (fp16, 64 columns): G = scale * gout, d_dfe, d_fin, dY (5,P,64), d_pe.
extern "C" int moda_chain_skin_bwd(const float* gout, const float* scale, const void* wpackT,
                                   const unsigned int* maskbits, long long P, void* G, void* d_dfe, void* d_fin,
                                   void* dY, void* d_pe, int fold, cudaStream_t stream) {
  if (P == 0) return 0;
  // fold: wpackT = [Wrgb^T, W'^T, W5[:, :63]^T, ...] (8 chunks), no d_fin step (see MODE_FOLD_FINAL)
  MODA_REQUIRE(gout && wpackT && maskbits && G && d_dfe && (d_fin || fold) && dY && d_pe, "chain_skin_bwd: null pointer");
  MODA_REQUIRE(al16(gout) && al16(wpackT), "chain_skin_bwd: gout and wpackT must be 16-byte aligned");
  Builder b;
  Program& pg = b.pg;
  pg.M = P; pg.num_tiles = (int)((P + TILE_M - 1) / TILE_M); pg.rep = 1; pg.nchunks = 2; pg.stages = MODA_SKIN_BWD_STAGES;
  pg.maskbits = const_cast<unsigned int*>(maskbits);
  pg.load_src = gout; pg.load_ld = 32; pg.load_cols = 32; pg.load_scale = scale;
  const int A = 0, SX = 1;
  auto dy = [&](int i) { return (const char*)dY + (size_t)i * P * 64 * 2; };
  { Step& st = b.add(64, E_LOAD32); st.save_map = b.save(G, P, 64); b.out(st, A, 1); }
  { Step& st = b.add(64, E_MASK_IN); b.k(st, A, 0); st.mask_slot = 5; st.save_map = b.save(d_dfe, P, 64); b.out(st, A, 1); }
  int wc = 1;
  if (!fold) { Step& st = b.add(64, 0); b.k(st, A, wc++); st.save_map = b.save(d_fin, P, 64); b.out(st, A, 1); }
  { Step& st = b.add(64, E_MASK_IN); b.k(st, A, wc++); st.mask_slot = 4; st.save_map = b.save(dy(4), P, 64); b.out(st, A, 1); }
  { Step& st = b.add(64, 0); b.k(st, A, wc++); b.out(st, SX, 1); }
  for (int l = 4; l >= 1; --l) {
    Step& st = b.add(64, E_MASK_IN);
    b.k(st, A, wc++);
    st.mask_slot = l - 1; st.save_map = b.save(dy(l - 1), P, 64);
    b.out(st, A, 1);
  }
  { Step& st = b.add(64, E_ADD_SX); b.k(st, A, wc++); st.save_map = b.save(d_pe, P, 64); b.out(st, SX, 1); }
  return launch<64, MODA_SKIN_EPI, 1, MODA_SKIN_BWD_CTAS, P_BWD>(b, wpackT, 64, wc * 64, stream);
}
