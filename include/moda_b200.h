/* libmoda_b200.so -- C ABI of the B200-native articulated volume renderer.
 *
 * The reference (ChaoyueSong/MoDA) has no FFI: its boundary for this path is a Python function-level API
 * (nnutils/rendering.py, nnutils/nerf.py, nnutils/geom_utils.py, nnutils/dual_quat.py).  Each entry point
 * below names the reference function(s) it replaces; the Python side that keeps the reference's names and
 * signatures on top of this ABI is moda_b200/{rendering,nerf,geom_utils,dual_quat}.py, and the binding a
 * MoDA maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless stated otherwise ("host"); pointers are
 *     borrowed for the duration of the stream-ordered call; the library allocates nothing;
 *   - "ld*" arguments are row strides in elements, so column slices of wider matrices can be passed;
 *   - return value 0 = success; >0 = cudaError_t of the launch; <0 = argument check failed;
 *     moda_last_error() returns the thread-local message; nothing synchronises the device;
 *   - re-entrant: no mutable state except two atomics (a sticky "cluster launch unavailable" flag and the debug trace
 *     hook moda_chain_set_trace); per-device attributes are set on every launch; work is issued only on the stream argument;
 *   - optional pointers may be NULL where stated.
 */
#ifndef MODA_B200_H
#define MODA_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* moda_version(void);
const char* moda_last_error(void);
/* 0 when the current device is sm_100 (the only target this library is built for) */
int moda_device_check(void);

/* ---- positional encoding: Embedding.forward, nnutils/nerf.py:35-75 -----------------------------------
 * out (M, C*(1+2F)) = [x | w_k sin(2^k x) | w_k cos(2^k x)]_k ; win: HOST array of F window weights
 * (nerf.py:63-66) or NULL for all-ones. */
int moda_embed_fwd(const float* x, int ldx, float* out, int ldo, long long M, int C, int F, const float* win,
                   cudaStream_t stream);
/* gx (M,C) (= or += when accumulate) d out/d x ^T gout */
int moda_embed_bwd(const float* x, int ldx, const float* gout, int ldo, float* gx, int ldg, long long M, int C,
                   int F, const float* win, int accumulate, cudaStream_t stream);

/* ---- ray sampling: render_rays, nnutils/rendering.py:64-89 -------------------------------------------
 * z (R,S) = near(1-s)+far s [or disparity], stratified jitter with mid-point bins when perturb>0
 * (jitter = the U[0,1) tensor of rendering.py:82); xyz (R,S,3) = o + d z; dn (R,3) = d/|d| (may be NULL). */
int moda_sample_rays_fwd(const float* o, const float* d, const float* near_, const float* far_,
                         const float* jitter, float perturb, int use_disp, float* z, float* xyz, float* dn, int R,
                         int S, cudaStream_t stream);
/* xyz = o + d z for given depths (second pass of the coarse->fine scheme, rendering.py:112-113) */
int moda_points_from_depths(const float* o, const float* d, const float* z, float* xyz, int R, int S,
                            cudaStream_t stream);
/* go (R,3) = sum_s gxyz ; gd (R,3) = sum_s z gxyz + gnd d/|d| + normalise-adjoint(gdn); gxyz/gdn/gnd/go may be NULL */
int moda_sample_rays_bwd(const float* d, const float* z, const float* gxyz, const float* gdn, const float* gnd,
                         float* go, float* gd, int R, int S, cudaStream_t stream);
/* sample_pdf, rendering.py:582-623.  merge=0: bins (R,n+1), weights (R,n) -> out (R,NI).
 * merge=1: weights = compositing weights (R,S); bins are the mid-points of z (R,S), the pdf uses
 * weights[:,1:-1] (n = S-2) and out (R,S+NI) is the sorted union with z (rendering.py:103-110).
 * det: u = linspace(0,1,NI); else u (R,NI) is the uniform draw. */
int moda_sample_pdf(const float* z, const float* bins, const float* weights, const float* u, float* out, int R,
                    int S, int n, int NI, int det, float eps, int merge, cudaStream_t stream);

/* ---- dual quaternions: nnutils/dual_quat.py ------------------------------------------------------------
 * op: 0 dq_quaternion_conjugate (:65-74), 1 dq_combined_conjugate (:76-85), 2 dq_normalize (:51-62),
 *     3 dq_inverse (:87-93), 4 q_normalize (:4-12, 4-wide).  n = number of (dual) quaternions. */
int moda_dq_unary_fwd(int op, const float* in, float* out, long long n, cudaStream_t stream);
int moda_dq_unary_bwd(int op, const float* in, const float* gout, float* gin, long long n, cudaStream_t stream);
/* width 4: q_mul (:14-31); width 8: dq_mul (:33-49) */
int moda_dq_mul_fwd(const float* a, const float* b, float* out, long long n, int width, cudaStream_t stream);
int moda_dq_mul_bwd(const float* a, const float* b, const float* gout, float* ga, float* gb, long long n,
                    int width, cudaStream_t stream);

/* ---- Gaussian bones: bone_transform, nnutils/geom_utils.py:59-111 (dual-quaternion branch) -------------
 * bones (B,10) or (R,B,10) when bones_per_ray; rts (R,B,8); out (R,B,10). */
int moda_bone_transform_fwd(const float* bones, const float* rts, float* out, int R, int B, int bones_per_ray,
                            cudaStream_t stream);
/* gbones accumulated (zero it first) when shared, overwritten when per-ray; grts overwritten */
int moda_bone_transform_bwd(const float* bones, const float* rts, const float* gout, float* gbones, float* grts,
                            int R, int B, int bones_per_ray, cudaStream_t stream);

/* ---- skinning weights + dual-quaternion blend skinning ------------------------------------------------
 * One kernel covers skinning / skinning_chunk (geom_utils.py:237-302), dqs_blend_skinning(_chunk)
 * (:457-517) and the warps of neu_dbs (:372-456) fused with gauss_mlp_skinning's weight computation:
 *   weights  W = softmax_b(-kappa sum_k s_k (R_b^T (c_b - p))_k^2 + dskin_b),  kappa = 1000 exp(skin_aux[0])
 *            (or W = skin_in when given);   blend b = sum_b W_b dq_b;   y = DQ-transform(p; b/|b_real|).
 *   deform: bones are first moved by bone_transform(bones, rts) (backward warp);
 *   invert: dq = dq_inverse(rts) (backward warp) instead of rts (forward warp).
 * pts (R,S,3); rts (R,B,8) or NULL (weights only); dskin / skin_in / y / skin_out (R,S,*) optional. */
int moda_skin_warp_fwd(const float* pts, const float* bones, const float* rts, const float* skin_aux,
                       const float* dskin, const float* skin_in, float* y, float* skin_out, int R, int S, int B,
                       int ld_dskin /* row pitch of dskin, 0 = B */, int bones_per_ray, int deform, int invert,
                       cudaStream_t stream);
/* gy (R,S,3) / gskin (R,S,B): incoming gradients (either may be NULL).  gpts, gdskin, gskin_in overwritten;
 * grts (R,B,8), gbones, gaux accumulated (zero them first).  gcopies >= 1: with bones shared by every ray, gbones is
 * (gcopies,B,10) and gaux (gcopies,2): ray r accumulates into copy r % gcopies and the caller sums the copies.  All rays
 * adding into ONE (B,10) block (gcopies = 1) serialises in L2: same-address atomics retire at ~20 M/s per address on
 * B200, which tripled this kernel's time at 8192 rays. */
int moda_skin_warp_bwd(const float* pts, const float* bones, const float* rts, const float* skin_aux,
                       const float* dskin, const float* skin_in, const float* gy, const float* gskin, float* gpts,
                       float* gdskin, float* gskin_in, float* grts, float* gbones, float* gaux, int R, int S,
                       int B, int ld_dskin, int bones_per_ray, int deform, int invert, int gcopies,
                       cudaStream_t stream);

/* ---- compositing: inference, nnutils/rendering.py:183-235 (+ frame_cyc_dis :341,:473) ------------------
 * rgb (P,3; row stride ld_rgb), sigma (P; stride ld_sigma), z (R,S), d (R,3), beta (1); noise (R,S) and
 * mask (R,S uint8, 1 = alpha forced to 0, :210-215) optional; xa/xb (R,S,3) optional pair for the cycle
 * term out_cyc = sum_s |xa-xb| w.  Outputs: out_rgb (R,3), out_depth (R), out_sil (R), out_w (R,S),
 * out_vis (R,S) = transmittance. */
int moda_composite_fwd(const float* rgb, int ld_rgb, const float* sigma, int ld_sigma, const float* z,
                       const float* d, const float* beta, const float* noise, const unsigned char* mask,
                       const float* xa, const float* xb, float* out_rgb, float* out_depth, float* out_sil,
                       float* out_w, float* out_vis, float* out_cyc, int R, int S, cudaStream_t stream);
/* vis = saved transmittance; g_* incoming (NULL = zero); g_rgbs/g_sigma/g_nd/g_xa/g_xb overwritten,
 * g_beta accumulated.  g_nd (R) is the gradient w.r.t. |d| (feed it to moda_sample_rays_bwd). */
int moda_composite_bwd(const float* rgb, int ld_rgb, const float* sigma, int ld_sigma, const float* z,
                       const float* d, const float* beta, const float* noise, const unsigned char* mask,
                       const float* xa, const float* xb, const float* vis, const float* g_rgb,
                       const float* g_depth, const float* g_sil, const float* g_cyc, const float* g_w,
                       float* g_rgbs, int ld_grgb, float* g_sigma, int ld_gsigma, float* g_beta, float* g_nd,
                       float* g_xa, float* g_xb, int R, int S, cudaStream_t stream);

/* ---- per-ray expectations under the compositing weights: feat_final = sum_s w f (nnutils/rendering.py:233) ----
 * w (R,S), v (R,S,C), C <= 32 -> out (R,C).  Adjoint: gw (R,S) = sum_c gout v, gv (R,S,C) = w gout (either may be NULL). */
int moda_wsum_fwd(const float* w, const float* v, float* out, int R, int S, int C, cudaStream_t stream);
int moda_wsum_bwd(const float* w, const float* v, const float* gout, float* gw, float* gv, int R, int S, int C,
                  cudaStream_t stream);

/* ---- linear layers of NeRF.forward (nnutils/nerf.py:147-198) with evaluate_mlp's input assembly
 * (nnutils/geom_utils.py:19-57) folded in.  The A operand is a virtual row-wise concatenation of nseg
 * (<=3) column segments described by HOST arrays: seg_ptr[i] device pointer, seg_ld[i] row stride,
 * seg_width[i] columns, seg_type[i] (0 dense matrix; 1 per-ray rows broadcast over seg_aux[i] samples;
 * 2 positional encoding of points with seg_aux[i] channels, computed on the fly), win = HOST PE window.
 * W is nn.Linear's (N, ldw) row-major weight, read in place. */
int moda_linear_fwd(int M, int N, int nseg, const float* const* seg_ptr, const int* seg_ld, const int* seg_width,
                    const int* seg_type, const int* seg_aux, const float* win, int n_win, const float* W, int ldw,
                    const float* bias, int act /* 0 none, 1 relu, 2 sigmoid */, float* Y, int ldy,
                    cudaStream_t stream);
/* dA (M,Kseg) (= or +=) dY (M,N) W[:, k0:k0+Kseg], then zeroed where mask<=0 (relu of the producer) */
int moda_linear_dgrad(int M, int N, int Kseg, const float* dY, int ldy, const float* W, int ldw, int k0,
                      const float* mask, int ldm, int accumulate, float* dA, int lda, cudaStream_t stream);
/* dW[:, k0:k0+K] += dY^T A ; dbias += colsum(dY) (dbias may be NULL).  Accumulates: zero dW first. */
int moda_linear_wgrad(int M, int N, int nseg, const float* const* seg_ptr, const int* seg_ld,
                      const int* seg_width, const int* seg_type, const int* seg_aux, const float* win, int n_win,
                      const float* dY, int ldy, float* dW, int ldw, int k0, float* dbias, cudaStream_t stream);
/* out (R,N) = per-ray sums over S consecutive rows of in (R*S, N) */
int moda_segsum(const float* in, int ld, float* out, int R, int S, int N, cudaStream_t stream);
/* out = g * act'(y): kind 1 relu, 2 sigmoid; row-strided (M,N) views */
int moda_act_bwd(int kind, const float* y, int ldy, const float* g, int ldg, float* out, int ldo, long long M,
                 int N, cudaStream_t stream);

/* ---- tensor-core (tcgen05 / TMEM / TMA) linear layers for the 8x256 trunk: NeRF.forward, nerf.py:147-198 --
 * fp16 operands (row-major, 16-byte aligned, row pitch multiple of 8 elements, K multiple of 64), fp32
 * accumulation.  Y[M,N] = epi([A1 | A2][M,K1+K2] B[N,K1+K2]^T), N in {64,128,256}.  Epilogue, in order:
 * + bias[n] + rowbias[m/rep][n] + rv[m] cv[n] (*rscale);  ReLU;  zero where mask[m][n] <= 0;  then
 * y16 (=|+= when acc16) and/or y32 = value * (*oscale).  bias/rowbias/rv/cv/rscale/oscale are fp32 device
 * pointers (scalars for rscale/oscale), any of them may be NULL. */
int moda_tc_linear(const void* A1, int lda1, int K1, const void* A2, int lda2, int K2, const void* B, int ldb, int M,
                   int N, const float* bias, const float* rowbias, int rep, int relu, const void* mask, int ldm,
                   const float* rv, const float* cv, const float* rscale, void* y16, int ldy16, int acc16, float* y32,
                   int ldy32, const float* oscale, cudaStream_t stream);
/* dW (N, ldw) fp32 += (*oscale) dY[M,N]^T X[M,K], first k_valid columns only; (N,K) in {256,128} x {64,128,256}.
 * Both operands are consumed MN-major straight from their row-major fp16 storage. */
int moda_tc_wgrad(const void* dY, int ldy, int N, const void* X, int ldx, int K, int M, float* dW, int ldw,
                  int n_valid, int k_valid, const float* oscale, float* dbias /* NULL, or (N) += colsum(dY) */,
                  cudaStream_t stream);
/* njobs (<= 9) weight gradients of ONE shape (N, K) over the same M rows in one launch (HOST arrays of length njobs;
 * dbias[j] may be NULL): the grid is partitioned among the jobs, which removes the ramp-up / tail of a launch per layer
 * and most of the red.global traffic of the final flush.  (N, K) in {256,128,64} x {256,128,64} as instantiated. */
int moda_tc_wgrad_multi(int njobs, const void* const* dY, const int* ldy, const void* const* X, const int* ldx,
                        float* const* dW, const int* ldw, const int* n_valid, const int* k_valid, float* const* dbias,
                        int N, int K, int M, const float* oscale, cudaStream_t stream);
/* fp16 operand staging for the trunk: positional encoding of (P,3) points into (P,64) [63 channels + zero pad]
 * (Embedding.forward, nerf.py:35-75) and its adjoint (gxyz (=|+=) (*inv_scale) J^T g16) */
int moda_pe16_fwd(const float* xyz, void* out16, void* out16lo /* NULL, or low half of the split pair */, int ldo,
                  long long P, int F, const float* win, cudaStream_t stream);
int moda_pe16_bwd(const float* xyz, const void* g16, const void* g16lo, int ldg, float* gxyz, long long P, int F,
                  const float* win, const float* inv_scale, int accumulate, cudaStream_t stream);
/* fp32 weight block in[:, col0:col0+cols] (rows x cols) -> fp16 block of out_rows x width (row pitch ld_out),
 * zero padded, optionally transposed */
int moda_pack16(const float* in, int ld_in, int rows, int cols, int col0, void* out16, void* out16_dup,
                void* out16_lo, int ld_out, int out_rows, int width, int transpose, cudaStream_t stream);
/* n pack16 blocks of one packed matrix (shared row pitch ld_out) in one launch; out16_lo[i] may be NULL */
int moda_pack16_multi(int n, const float* const* in, const int* ld_in, const int* rows, const int* cols,
                      const int* col0, void* const* out16, void* const* out16_lo, int ld_out, const int* out_rows,
                      const int* width, const int* transpose, cudaStream_t stream);
/* fp32 (M, cols) * (*scale) -> fp16 (hi, lo) pair, zero padded to `width` columns (lo may be NULL) */
int moda_split16(const float* in, int ld_in, int cols, const float* scale, void* hi, void* lo, int ld_out, int width,
                 long long M, cudaStream_t stream);
/* Split-precision (fp32-class) variants used for nerf_skin: operands are fp16 (hi, lo) pairs, value = hi + lo;
 * x W^T ~= hi Whi^T + lo Whi^T + hi Wlo^T runs as one GEMM over K = [hi | lo | hi] against B3 = [Whi | Whi | Wlo]
 * (N, 3K).  Outputs: fp16 (hi, lo) pair and/or fp32. */
int moda_tc_linear_split(const void* A1hi, const void* A1lo, int lda1, int K1, const void* A2hi, const void* A2lo,
                         int lda2, int K2, const void* B3, int ldb, int M, int N, const float* bias,
                         const float* rowbias, int rep, int relu, const void* mask, int ldm, void* yhi, void* ylo,
                         int ldy16, int acc16, float* y32, int ldy32, const float* oscale, cudaStream_t stream);
/* dW += (*oscale) (dYhi + dYlo)^T (Xhi + Xlo) (lo*lo dropped); rows >= n_valid / cols >= k_valid not written */
int moda_tc_wgrad_split(const void* dYhi, const void* dYlo, int ldy, int N, const void* Xhi, const void* Xlo, int ldx,
                        int K, int M, float* dW, int ldw, int n_valid, int k_valid, const float* oscale,
                        float* dbias, cudaStream_t stream);
/* sigma (256->1) and rgb (128->3, sigmoid) heads (nerf.py:178, 188-195) on fp16 activations; raw (P,4) fp32 */
int moda_head_fwd(const void* H8, const void* Dfe, const float* ws, const float* bs, const float* Wr,
                  const float* br, float* raw, long long P, cudaStream_t stream);
/* their adjoint: dDfe16 = (*scale) (g_pre Wr) relu'(Dfe), gsig = graw[:,3]; gWr/gbr/gws/gbs accumulated */
int moda_head_bwd(const void* H8, const void* Dfe, const float* raw, const float* graw, const float* Wr,
                  const float* scale, void* dDfe, float* gsig, float* gWr, float* gbr, float* gws, float* gbs,
                  long long P, cudaStream_t stream);
/* out[n] += (*oscale) sum_m in16[m][n]   /   out (R,N) = (*oscale) per-ray sums of S consecutive rows */
int moda_colsum16(const void* in16, int ld, float* out, long long M, int N, const float* oscale, cudaStream_t stream);
int moda_segsum16(const void* in16, int ld, float* out, int R, int S, int N, const float* oscale, int accumulate,
                  cudaStream_t stream);
/* scale2 = {S, 1/S}, S = 2^floor(log2(target / max|g|)): loss scale of the fp16 gradient chain; work: 1 uint */
int moda_loss_scale(const float* g, long long n, float target, unsigned int* work, float* scale2,
                    cudaStream_t stream);

/* ---- fused layer chains (csrc/chain.cu): one persistent tcgen05 kernel per MLP pass; activations stay in shared
 * memory / TMEM across all layers.  Replaces, per 128-sample tile, the whole of NeRF.forward (nnutils/nerf.py:147-198)
 * on the input evaluate_mlp assembles (nnutils/geom_utils.py:19-57), resp. its adjoint.
 * Packed-weight chunk orders are documented at the definitions.  Saved outputs may be NULL (inference). */
int moda_chain_trunk_fwd(const float* xyz, long long P, int rep /* samples per ray */, int F, const float* win,
                         const void* wpack /* fp16 (256, 38*64) */, const float* const* biases /* b1..b8, bfinal */,
                         const float* rowbias /* (P/rep,128) per-ray part of the dir layer */, const float* ws,
                         const float* bs, const float* Wr, const float* br, void* A0 /* (P,64) fp16 PE */,
                         void* H /* (8,P,256) */, void* fin /* (P,256) */, void* dfe /* (P,128) */,
                         unsigned int* maskbits /* (8,tiles2,8,128) ReLU sign bits, tiles2 = ceil(P/128) rounded up to even */,
                         float* raw /* (P,4) */, int mode /* MODA_CHAIN_* bits */, cudaStream_t stream);
/* Density-only pass of nerf_coarse for grid queries (extract_mesh, nnutils/train_utils.py:1377-1404 with
 * nerf.py:176-180 sigma_only=True): layers 1-8 + sigma head on tensor cores, nothing saved.  wpack as for
 * moda_chain_trunk_fwd; sigma (P) fp32. */
int moda_chain_trunk_sigma(const float* xyz, long long P, int F, const float* win, const void* wpack,
                           const float* const* biases, const float* ws, const float* bs, float* sigma, int mode,
                           cudaStream_t stream);
int moda_chain_trunk_bwd(const void* d_dfe /* (P,128) fp16 */, const float* gsig /* (P) */, const float* ws,
                         const float* rscale, const void* wpackT /* fp16 (256, 42*64) */,
                         const unsigned int* maskbits, long long P, void* d_fin /* (P,256) */,
                         void* dY /* (8,P,256) */, void* d_pe /* (P,64) */, int mode, cudaStream_t stream);
int moda_chain_skin_fwd(const float* xyz, long long P, int rep, int F, const float* win,
                        const void* wpack /* fp16 (64, 18*64): [Whi | Wlo] per layer */,
                        const float* const* biases /* rb1, b2, b3, b4, rb5, bfinal, bdir64, brgb64 */, void* A0,
                        void* H /* (5,P,64) */, void* fin, void* dfe, unsigned int* maskbits /* (6,tiles2,4,128), tiles2 as above */,
                        float* y32 /* (P,32) delta skinning logits */,
                        int fold /* 1: xyz_encoding_final folded into dir_encoding, see MODA_CHAIN_FOLD_FINAL */,
                        cudaStream_t stream);
/* nerf_feat (nnutils/moda.py:447-449: NeRF(D=5, W=128, in 63, raw_feat=True, out 16), evaluated by rendering.py:174-178
 * and loss_utils.py:318-320) as one chain kernel per pass on the 256-wide engine; the final layer is always folded into
 * the direction layer (MODA_CHAIN_FOLD_FINAL).  Packed-weight chunk orders at the definitions (csrc/chain.cu). */
int moda_chain_feat_fwd(const float* xyz, long long P, int F, const float* win, const void* wpack /* fp16 (128, 13*64) */,
                        const float* const* biases /* b1..b5 (128), b' (64), brgb (64, zero padded) */, void* A0 /* (P,64) */,
                        void* H /* (5,P,128) */, void* dfe /* (P,64) */, unsigned int* maskbits /* (6,tiles2,4,128) u64 */,
                        float* y32 /* (P,32) */, int mode /* MODA_CHAIN_PAIR | MODA_CHAIN_TWO_SLOTS */, cudaStream_t stream);
int moda_chain_feat_bwd(const float* gout /* (P,32) */, const float* scale, const void* wpackT /* fp16 (128, 14*64) */,
                        const unsigned int* maskbits, long long P, void* G /* (P,64) */, void* d_dfe /* (P,64) */,
                        void* dY /* (5,P,128) */, void* d_pe /* (P,64) */, int mode, cudaStream_t stream);
/* debug: device buffer (>= 16004 int64, zeroed) that subsequent chain launches fill with an event timeline of
 * block 0's third tile (tools/chain_trace.py); NULL switches tracing off */
int moda_chain_set_trace(long long* buf);
/* `mode` of the 256-wide chains (moda_chain_trunk_fwd / _sigma / _bwd), a per-call argument:
 *   MODA_CHAIN_PAIR       launched as clusters of two CTAs sharing one tcgen05.mma.cta_group::2 stream (M = 256 = two
 *                         tiles; each CTA stages half of every weight chunk);
 *   MODA_CHAIN_TWO_SLOTS  (with MODA_CHAIN_PAIR) two tiles in flight per CTA: the tensor core works on one tile's layer
 *                         while the other tile's epilogue drains its accumulator (used when every CTA gets >= 2 tiles).
 * Results are bit-identical in every mode.  If a cluster cannot be launched the library falls back to the single-CTA
 * kernels for the rest of the process (moda_chain_pair_available() then returns 0).  The sign-bit buffers must be
 * sized for an EVEN tile count in every mode. */
#define MODA_CHAIN_PAIR 1
#define MODA_CHAIN_TWO_SLOTS 2
/*   MODA_CHAIN_FOLD_FINAL xyz_encoding_final (no activation, nerf.py:182) and dir_encoding (nerf.py:186-190) are two
 *                         linear maps in a row: the caller passes their product W' = Wdir[:, :W] Wfinal in the packed
 *                         weights (forward: in place of [Wfinal, Wdir]; adjoint: W'^T in place of [Wdir^T, Wfinal^T]) and
 *                         bdir + Wdir[:, :W] bfinal in the bias, and every pass runs one step less; fin / d_fin are not
 *                         produced (may be NULL).  The caller maps dW' back: dWdir = dW' Wfinal^T, dWfinal = Wdir^T dW'.
 *                         The 64-wide programs take the same switch as their `fold` argument. */
#define MODA_CHAIN_FOLD_FINAL 4
int moda_chain_pair_available(void);
int moda_chain_skin_bwd(const float* gout /* (P,32) */, const float* scale, const void* wpackT /* fp16 (64, 9*64) */,
                        const unsigned int* maskbits, long long P, void* G, void* d_dfe, void* d_fin,
                        void* dY /* (5,P,64) */, void* d_pe, int fold, cudaStream_t stream);

/* xyz_encoding_final folded into dir_encoding (MODA_CHAIN_FOLD_FINAL): the product weights Wp (n, W) = Wd[:, :W] Wf and
 * bias bp = bd + Wd[:, :W] bf, and the way back gWf += Wd1^T gWp, gbf += Wd1^T dbp, gWd[:, :W] += gWp Wf^T + dbp bf^T
 * (nerf.py:182-190: two linear maps in a row).  O(weights) work, fp32. */
int moda_fold_final(const float* Wd, int ldwd, const float* Wf, int ldwf, const float* bf, const float* bd, int n, int W,
                    float* Wp, float* bp, cudaStream_t stream);
int moda_unfold_final(const float* gWp, const float* dbp, const float* Wd, int ldwd, const float* Wf, int ldwf,
                      const float* bf, int n, int W, float* gWf, int ldgf, float* gbf, float* gWd, int ldgd,
                      cudaStream_t stream);
/* One pass over the (n, m) Sinkhorn kernel matrix of feat_match (nnutils/loss_utils.py:347-386) serving both products of an
 * iteration: y = K x (row sums), z = g(y) elementwise (mode 0: p / (y + delta); mode 1: -y u / (v + delta)), w += K^T z
 * (column sums, w zeroed by the caller; NULL: row sums only).  K row-major fp32, m % 4 == 0, m <= 8192.
 * The input vector is either given (xmode 0: x) or formed while it is staged and written to xout (when non-null):
 * xmode 1: x = xp / (xc + delta) (b_i of loss_utils.py:372 from the column sums), xmode 2: x = -xg xb / (xc + delta) (the
 * adjoint's gc_i). */
int moda_sinkhorn_pass(const float* K, int n, int m, const float* x, float* y, float* z, float* w, int mode, float p,
                       float delta, const float* u, const float* v, int xmode, const float* xc, const float* xg,
                       const float* xb, float xp, float* xout, cudaStream_t stream);
/* The kernel matrix itself, K (n, m) = exp((F V^T - 1) / eps) from unit features F (n, d), V (m, d), d = 16
 * (loss_utils.py:325-332, 359); c_out (m, zeroed by the caller, may be NULL) += cs_scale x column sums of K. */
int moda_sinkhorn_matrix(const float* F, const float* V, int n, int m, int d, float eps, float cs_scale, float* K,
                         float* c_out, cudaStream_t stream);
/* out (n, 4) = K X, X (m, 4): the soft-argmax numerator and row sum of loss_utils.py:383-386 in one pass over K. */
int moda_sinkhorn_rows4(const float* K, int n, int m, const float* X, float* out, cudaStream_t stream);
/* out (m, 4, zeroed by the caller) += K^T Wt, Wt (n, 4): the direct term of the adjoint in one pass over K. */
int moda_sinkhorn_cols4(const float* K, int n, int m, const float* Wt, float* out, cudaStream_t stream);
/* gF (n, d) += gcost V, gV (m, d) += gcost^T F with gcost = K / eps * (L Rm) never materialised: L (n, R), Rm (R, m) are the
 * low-rank factors of dLoss/dK (R = 44: 4 direct + 39 sweep terms + 1 zero), d = 16. */
int moda_sinkhorn_gcost(const float* K, int n, int m, const float* L, const float* Rm, int R, const float* F, const float* V,
                        int d, float eps, float* gF, float* gV, cudaStream_t stream);
/* Flow rendering of the third warp (rendering.py:434-459, 480-499: obj_to_cam, pinhole_cam, vrender_flo of
 * geom_utils.py:567-581, 654-672, 1704-1743) for one paired frame: the warped samples xyz (N, S, 3) are projected with the
 * ray's camera R (N, 9 row-major), T (N, 3), K (N, 4) = (fx, fy, px, py) and flo (N, 2) = sum_s w'_s / (1e-9 + sum w') (xy'_s -
 * xys) * 2 / img_size over the valid samples (z >= 1e-5 and |xy| <= 2 img_size), valid (N) = 1 when every sample is.  The
 * backward entry recomputes the projection and overwrites gxyz (N, S, 3), gR (N, 9), gT (N, 3), gK (N, 4), gw (N, S). */
int moda_flow_render_fwd(const float* xyz, const float* R, const float* T, const float* K, const float* w, const float* xys,
                         int N, int S, float img_size, float* flo, float* valid, cudaStream_t stream);
int moda_flow_render_bwd(const float* xyz, const float* R, const float* T, const float* K, const float* w, const float* xys,
                         const float* gflo, int N, int S, float img_size, float* gxyz, float* gR, float* gT, float* gK,
                         float* gw, cudaStream_t stream);
/* AdamW step (torch.optim.AdamW as the reference's training loop uses it, nnutils/train_utils.py:177-222) on the flat
 * parameter / gradient buffers of the data-parallel path; m, v: moment buffers, state: 3 device floats {step count,
 * lr / (1 - beta1^t), sqrt(1 - beta2^t)}, advanced on the device so that the call can be captured in a CUDA graph.
 * n % 4 == 0, 16-byte aligned buffers. */
int moda_adamw_flat(float* p, const float* g, float* m, float* v, long long n, float* state, float lr, float b1, float b2,
                    float eps, float wd, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MODA_B200_H */
