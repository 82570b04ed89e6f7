/* mcd_b200.h -- C ABI of the B200 (sm_100a) neuron->concept scoring kernels.
 *
 * This is the drop-in boundary for the scoring path of Mammo-CLIP Dissect.  The
 * reference implements the path as PyTorch calls inside
 *   concept_vit/similarity.py            (soft_wpmi :49-73, wpmi :75-97, cos* :7-47)
 *   concept_vit/utils.py:27-52           (get_activation pooling hook)
 *   concept_vit/utils.py:570-594         (row-normalise + I @ T.T)
 * and selects it with eval("similarity.<name>") (describe_clip_neurons.py:41).  The host
 * mirror of that Python surface lives in mammo_clip_dissect_b200/similarity.py and binds
 * these entry points with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless named host_*;
 *   - matrices are row-major fp32 with an explicit leading dimension in ELEMENTS;
 *   - `stream` is a cudaStream_t (torch.cuda.current_stream().cuda_stream); calls only
 *     enqueue work: no allocation, no synchronisation, re-entrant per stream;
 *   - return value 0 = success, negative = error (mcd_strerror), never throws;
 *   - N = probe images, K = neurons (columns of the activation matrix), C = concepts,
 *     k = top_k, D = embedding width.
 */
#ifndef MCD_B200_H
#define MCD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCD_ABI_VERSION 1

#define MCD_OK 0
#define MCD_ERR_INVALID_ARGUMENT (-1)
#define MCD_ERR_UNSUPPORTED (-2)
#define MCD_ERR_WORKSPACE (-3)
#define MCD_ERR_CUDA (-4)
#define MCD_ERR_NO_DEVICE (-5)

#define MCD_LSE_BLOCK 256 /* neurons per log-sum-exp partial; fixed so results do not depend on sharding */

typedef void *mcd_stream_t; /* cudaStream_t */

typedef enum { MCD_F32 = 0, MCD_F16 = 1, MCD_BF16 = 2 } mcd_dtype_t;
typedef enum { MCD_POOL_MEAN = 0, MCD_POOL_MAX = 1 } mcd_pool_t;

/* ---- library ------------------------------------------------------------------------- */
int mcd_abi_version(void);
const char *mcd_strerror(int code);
const char *mcd_build_info(void);      /* "sm_100a nvcc <ver> ..." */
uint64_t mcd_launch_count(void);       /* kernels launched by this library so far (monotonic) */
int mcd_device_check(void);            /* 0 iff the current device is compute capability 10.x */
int mcd_set_tunable(const char *name, int64_t value); /* bench/test knobs: "topk_splits", "accum_tile", ... */

/* ---- K1b: S = softmax(a * P, dim=1)          replaces similarity.py:54 / :80 ---------- */
/* P [n_rows, n_cols] (ldp), S [n_rows, lds]; columns n_cols..lds-1 of S are written as 0. */
int mcd_softmax_rows_f32(const float *P, int64_t ldp, float *S, int64_t lds,
                         int64_t n_rows, int64_t n_cols, float a, mcd_stream_t stream);

/* ---- K1: row-normalise I and T, P = I T^T, S = softmax(a P)    replaces utils.py:577-594
 *      (+ similarity.py:54).  I [N,D] (ldi), T [C,D] (ldt).  P_out and/or S_out may be NULL.
 *      Tensor-core path: tcgen05 kind::tf32 with a 3-term hi/lo split (fp32-grade result). */
size_t mcd_gemm_nt_softmax_workspace_bytes(int64_t N, int64_t C, int64_t D);
int mcd_last_gemm_path(void);          /* what the last K1 call ran: 1 tcgen05 3xTF32, 2 tcgen05 with fused softmax, 3 fp32 FFMA */
int mcd_gemm_nt_softmax_f32(const float *I, int64_t ldi, const float *T, int64_t ldt,
                            int64_t N, int64_t C, int64_t D, int normalize_rows, float a,
                            float *P_out, int64_t ldp, float *S_out, int64_t lds,
                            void *workspace, size_t workspace_bytes, mcd_stream_t stream);

/* ---- K2: per-column top-k over the probe-image axis   replaces torch.topk(A, dim=0, k)
 *      (similarity.py:55, :82, :107).  Total order: value desc, image index asc, NaN largest,
 *      -0.0 == +0.0.  A [N,K] (lda).  Any of idx64_out / idx32_out / vals_out [k,K] may be NULL. */
size_t mcd_topk_cols_workspace_bytes(int64_t N, int64_t K, int64_t k);
int mcd_topk_cols_f32(const float *A, int64_t lda, int64_t N, int64_t K, int64_t k,
                      int64_t *idx64_out, int32_t *idx32_out, float *vals_out,
                      void *workspace, size_t workspace_bytes, mcd_stream_t stream);

/* ---- K3: L[j,c] = sum_r log(1 + p[r] (S[idx[r,j],c] - 1) + min_prob)   (soft-WPMI body,
 *      similarity.py:59-65); p == NULL gives sum_r log(S[idx[r,j],c] + min_prob) (WPMI body,
 *      similarity.py:85-89).  S [N,C] (lds), idx [k,K] int32 (ld = K), p [k], L [K,C] (ldl).
 *      Limits: k <= 512, N * lds < 2^30.
 *      mcd_wpmi_accum_f32      any S: the reference's operation order per term (sub, mul, add, add, log -- a term <= 0
 *                              gives NaN / -inf exactly as in the reference), one MUFU lg2 per term;
 *      mcd_wpmi_accum_prob_f32 the caller vouches that S is a probability matrix (finite entries in [0, 1], what
 *                              mcd_softmax_rows_f32 / mcd_gemm_nt_softmax_f32 write; NaN propagates): a term is one FMA,
 *                              the terms of 4 consecutive ranks are multiplied before one lg2 (4x fewer MUFU ops).
 *                              Differs from the reference order by ~5e-8 relative on L (stated tolerance 1e-5); with
 *                              entries outside [0, 1] the result is undefined (two negative terms multiply to a
 *                              positive product).  This is what the fused calls below and the Python surface use. */
int mcd_wpmi_accum_f32(const float *S, int64_t lds, int64_t N, int64_t C,
                       const int32_t *idx, int64_t K, int64_t k, const float *p, float min_prob,
                       float *L, int64_t ldl, mcd_stream_t stream);
int mcd_wpmi_accum_prob_f32(const float *S, int64_t lds, int64_t N, int64_t C,
                            const int32_t *idx, int64_t K, int64_t k, const float *p, float min_prob,
                            float *L, int64_t ldl, mcd_stream_t stream);

/* ---- K3b: log p(d) and the final subtraction          replaces similarity.py:67-72 / :91-96
 *      partials [ceil(K/256), 2, C]: per 256-neuron block (max_c, sum_j exp(L[j,c]-max_c)).
 *      finalize combines `n_blocks_total` partials IN ORDER (fp64), so a neuron-sharded run
 *      that concatenates every shard's partials in global block order is bit-identical to
 *      the unsharded run:  out = L - lam * (lse - log(K_total)).  out may alias L. */
int mcd_col_lse_partials_f32(const float *L, int64_t ldl, int64_t K, int64_t C,
                             float *partials, mcd_stream_t stream);
int mcd_pmi_finalize_f32(const float *L, int64_t ldl, int64_t K, int64_t C,
                         const float *partials_all, int64_t n_blocks_total, int64_t K_total,
                         float lam, float *prob_d_out /* [C] scratch+result */,
                         float *out, int64_t ldo, mcd_stream_t stream);

/* ---- the whole soft_wpmi / wpmi call (similarity.py:49-73 / :75-97) behind one entry point:
 *      softmax(a P) -> column top-k of A -> gather / log-sum -> block LSE -> out = L - lam log p(d),
 *      the same kernels as the functions above with the intermediates in `workspace` (256-byte
 *      aligned).  p: device ramp [k] for soft_wpmi, NULL for wpmi.  out [K, C] (ldo). */
size_t mcd_pmi_scores_workspace_bytes(int64_t N, int64_t K, int64_t C, int64_t k);
int mcd_pmi_scores_f32(const float *P, int64_t ldp, const float *A, int64_t lda, int64_t N, int64_t K,
                       int64_t C, int64_t k, float a, float lam, const float *p, float min_prob,
                       float *out, int64_t ldo, void *workspace, size_t workspace_bytes,
                       mcd_stream_t stream);

/* ---- the same call up to the log-sums (what one rank of the neuron-sharded multi-GPU call runs before the partials
 *      are exchanged): L [K, C] (ldl) and the 256-neuron block partials [ceil(K/256), 2, C].
 *      Workspace as for mcd_pmi_scores_f32.  For long columns (N >= 8192) and K >= 8192 the neurons are cut into a few
 *      column chunks and chunk q's select + K3 + partials run on a library-owned side stream under K2's scan of chunk
 *      q + 1 (the side stream and its events are created once per device on first use; the call still only enqueues
 *      work and everything is ordered behind `stream` on return).  mcd_pmi_scores_f32 = this + mcd_pmi_finalize_f32. */
int mcd_pmi_logsums_f32(const float *P, int64_t ldp, const float *A, int64_t lda, int64_t N, int64_t K,
                        int64_t C, int64_t k, float a, const float *p, float min_prob,
                        float *L, int64_t ldl, float *partials, void *workspace, size_t workspace_bytes,
                        mcd_stream_t stream);

/* ---- K3b for several layers stacked along the neuron axis of one [sum K_l, C] matrix (SURVEY.md 8 f1:
 *      the 12-39 layers of a real job in one pass of every kernel).  Same arithmetic, block by
 *      block, as the single-layer functions, so each layer's rows get exactly the bits a separate
 *      call would produce.  Device tables: block_tab [n_blocks][3] = (first row, rows <= 256,
 *      segment) with every layer's blocks starting at its first row; seg_tab [n_seg][2] = (first
 *      block, blocks); seg_log_count [n_seg] = log(K_l) in fp64.  prob_d_out [n_seg][C]. */
int mcd_col_lse_partials_seg_f32(const float *L, int64_t ldl, int64_t C, const int32_t *block_tab,
                                 int64_t n_blocks, float *partials, mcd_stream_t stream);
int mcd_pmi_finalize_seg_f32(const float *L, int64_t ldl, int64_t C, const float *partials,
                             const int32_t *block_tab, int64_t n_blocks, const int32_t *seg_tab,
                             const double *seg_log_count, int64_t n_seg, float lam,
                             float *prob_d_out, float *out, int64_t ldo, mcd_stream_t stream);

/* ---- K3b fused with the all-gather of the scores (neuron-sharded multi-GPU call):
 *      same arithmetic as mcd_pmi_finalize_f32 on this rank's contiguous [K, C] log-sums, but the
 *      finalized slice is stored into rows [row_offset, row_offset + K) of EVERY destination
 *      matrix dest_bases[0..n_dest) (contiguous [K_total, C] fp32, 16-byte aligned slices; peer
 *      GPUs' memory mapped into this process plus the rank's own copy).  dest_bases is a HOST
 *      array of device pointers.  The caller orders the exchange (a barrier across the ranks
 *      before and after, on the same stream). */
#define MCD_MAX_PEERS 16
int mcd_pmi_finalize_bcast_f32(const float *L, int64_t K, int64_t C, const float *partials_all,
                               int64_t n_blocks_total, int64_t K_total, float lam,
                               float *prob_d_out, float *const *dest_bases, int n_dest,
                               int64_t row_offset, mcd_stream_t stream);
/* n floats of src to dest_bases[p] + dest_offset for every p (device pointers: peer-mapped buffers and / or local ones).
 * The multi-GPU exchange of the LSE partials: every rank stores its [blocks_g, 2, C] slice into all ranks' tables, then
 * the ranks meet at a symmetric-memory barrier.  No reference counterpart (the reference is single-GPU). */
int mcd_bcast_f32(const float *src, int64_t n, float *const *dest_bases, int n_dest, int64_t dest_offset,
                  mcd_stream_t stream);

/* ---- K3b's finalize fused with the per-neuron top concepts: out = L - lam log p(d) as above, and in the same pass
 *      (a warp per neuron row, the finalized values still in registers) the t <= 64 best (value, concept) pairs of
 *      every row, sorted descending (value desc, concept index asc, NaN largest) -- what the callers compute next with
 *      torch.topk(sim, 10, dim=1) / torch.max(sim, 1) (describe_broad_neurons.py:101, describe_clip_neurons.py:64).
 *      C <= 1024.  The _seg_ form serves several layers in one matrix: row_seg [K] names the layer of every row. */
int mcd_pmi_finalize_topk_f32(const float *L, int64_t ldl, int64_t K, int64_t C, const float *partials_all,
                              int64_t n_blocks_total, int64_t K_total, float lam, float *prob_d_out,
                              float *out, int64_t ldo, int64_t t, float *top_vals_out, int64_t *top_idx_out,
                              mcd_stream_t stream);
int mcd_pmi_finalize_seg_topk_f32(const float *L, int64_t ldl, int64_t K, int64_t C, const float *partials,
                                  const int32_t *row_seg, int64_t n_blocks, const int32_t *seg_tab,
                                  const double *seg_log_count, int64_t n_seg, float lam, float *prob_d_out,
                                  float *out, int64_t ldo, int64_t t, float *top_vals_out, int64_t *top_idx_out,
                                  mcd_stream_t stream);

/* ---- per-neuron top concepts: the t largest entries of every row of the score matrix, sorted
 *      descending (value desc, concept index asc, NaN largest)   replaces torch.topk(sim, 10, dim=1)
 *      / torch.max(sim, 1) in the callers (describe_broad_neurons.py:101, describe_clip_neurons.py:64).
 *      X [n_rows, n_cols] (ldx), n_cols <= 1024, t <= 64; vals_out / idx_out [n_rows, t]. */
int mcd_row_topk_f32(const float *X, int64_t ldx, int64_t n_rows, int64_t n_cols, int64_t t,
                     float *vals_out, int64_t *idx_out, mcd_stream_t stream);

/* ---- K4: spatial pooling of a hooked NCHW activation   replaces utils.py:38 / :47 -------
 *      x [B,C,H,W] contiguous, dtype f32/f16/bf16; out [B,C] same dtype (fp32 accumulation).
 *      workspace: mcd_pool_nchw_workspace_bytes (partials for planes split across CTAs). */
size_t mcd_pool_nchw_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W);
int mcd_pool_nchw(const void *x, mcd_dtype_t dtype, int64_t B, int64_t C, int64_t H, int64_t W,
                  mcd_pool_t mode, void *out, void *workspace, size_t workspace_bytes,
                  mcd_stream_t stream);

/* the same with the result row of image b at out[b * out_ld + c] (out_ld >= C elements), as `dtype` or as fp32
 * (out_dtype) -- the hook writes straight into the [n_images, sum K_l] activation matrix -- and, with channels_last != 0,
 * for activations stored in channels-last order ([B, H, W, C] in memory) without repacking them.  Same workspace. */
int mcd_pool_nchw_to(const void *x, mcd_dtype_t dtype, int64_t B, int64_t C, int64_t H, int64_t W,
                     int channels_last, mcd_pool_t mode, void *out, mcd_dtype_t out_dtype, int64_t out_ld,
                     void *workspace, size_t workspace_bytes, mcd_stream_t stream);

/* ---- cosine similarities (similarity.py:7-47), next rows of the scope table -------------
 *      column statistics of X [N,M]: cubed: mean[m] and norm[m] = max(||(x-mean)^3||_2, min_norm);
 *      plain: norm[m] = ||x||_2 (mean_out may be NULL). */
int mcd_col_stats_f32(const float *X, int64_t ldx, int64_t N, int64_t M, int cubed, float min_norm,
                      float *mean_out, float *norm_out, mcd_stream_t stream);
/* out[j,c] = sum_i f(A[i,j]) * f(P[i,c]),  f(x) = (x-mean)^3 / norm (cubed) or x / norm */
int mcd_cos_matmul_f32(const float *A, int64_t lda, const float *meanA, const float *normA,
                       const float *P, int64_t ldp, const float *meanP, const float *normP,
                       int64_t N, int64_t K, int64_t C, int cubed,
                       float *out, int64_t ldo, mcd_stream_t stream);

/* the whole cos_similarity / cos_similarity_cubed call (similarity.py:7-47) behind one entry point, on the tensor cores:
 * row-parallel column statistics of P [N,C] and A [N,K]; f(A), f(P) written transposed as hi + lo fp32 pairs into
 * `workspace` (neurons in slabs of 8192); out [K,C] = f(A)^T f(P) by a tcgen05 kind::tf32 3-term-split GEMM whose TMEM
 * accumulators are flushed into fp32 registers every 8 k-blocks (fp32-grade result for any N), split-K when the output
 * tiles alone would leave SMs idle.  min_norm is used by the cubed form only.  mcd_last_cos_path(): 1 tensor cores,
 * 3 the CUDA-core kernels above (tensor-map encoder unavailable, or tunable gemm_variant = 1). */
size_t mcd_cos_similarity_workspace_bytes(int64_t N, int64_t K, int64_t C);
int mcd_cos_similarity_f32(const float *P, int64_t ldp, const float *A, int64_t lda, int64_t N, int64_t K, int64_t C,
                           int cubed, float min_norm, float *out, int64_t ldo,
                           void *workspace, size_t workspace_bytes, mcd_stream_t stream);
int mcd_last_cos_path(void);

/* ---- rank_reorder (similarity.py:99-132).  idx / vals [top_n, K] from mcd_topk_cols_f32 (top_n = int(0.05 N) <= 8192).
 *      out[j,c] = -(mean_r |t_r - asc[rank_rc]|^p / baseline_j) / (mean_r P[idx_r, c])^scale_p, ranks by sorting
 *      (registers for top_n <= 512, shared memory beyond), ties among the gathered cosines by row position.
 *      Three steps sharing one workspace (mcd_rank_reorder_workspace_bytes), so that the baseline -- which needs the
 *      reference's random permutations -- can be produced on another stream while the ranks are being computed:
 *        mcd_rank_baseline_draws_f32 / _perms_f32   baseline_j = mean over the reference's 5 x torch.randperm(top_n) per
 *              neuron (similarity.py:119) of |asc_r - asc_perm(r)|^p, from draws [K, 5, top_n - 1] uint32 = the CPU
 *              generator's raw 32-bit outputs (torch.randperm = Fisher-Yates with z = draw % (n - i); shuffles on the
 *              device), or from perms [K, 5, top_n] int32 drawn on the host;
 *        mcd_rank_errors_f32    the gather + rank pass: leaves mean_r |.|^p in out and (mean cosine)^scale_p in the workspace;
 *        mcd_rank_finish_f32    out = -((e / baseline) / den), the reference's operation order (similarity.py:128-129).
 *      mcd_rank_reorder_f32 / mcd_rank_reorder_draws_f32 run the three on one stream.
 *      mcd_mt19937_draws produces the raw draws themselves: it continues MT19937 (the engine of torch's CPU generator) on
 *      the device.  state_io: 624 state words + 1 word "index of the next output" (624 = twist before the next draw); on
 *      return it holds the state after `count` outputs, for the host to write back into the generator. */
size_t mcd_rank_reorder_workspace_bytes(int64_t K, int64_t top_n, int64_t C);
int mcd_mt19937_draws(uint32_t *state_io, int64_t count, uint32_t *draws, mcd_stream_t stream);
int mcd_rank_baseline_draws_f32(const float *vals, int64_t K, int64_t top_n, int64_t C, const uint32_t *draws, float p,
                                void *workspace, size_t workspace_bytes, mcd_stream_t stream);
int mcd_rank_baseline_perms_f32(const float *vals, int64_t K, int64_t top_n, int64_t C, const int32_t *perms, float p,
                                void *workspace, size_t workspace_bytes, mcd_stream_t stream);
int mcd_rank_errors_f32(const float *P, int64_t ldp, int64_t N, int64_t C, const int32_t *idx, const float *vals,
                        int64_t K, int64_t top_n, float p, float scale_p, void *workspace, size_t workspace_bytes,
                        float *out, int64_t ldo, mcd_stream_t stream);
int mcd_rank_finish_f32(int64_t K, int64_t C, int64_t top_n, void *workspace, size_t workspace_bytes, float *out,
                        int64_t ldo, mcd_stream_t stream);
int mcd_rank_reorder_f32(const float *P, int64_t ldp, int64_t N, int64_t C, const int32_t *idx, const float *vals,
                         int64_t K, int64_t top_n, const int32_t *perms, float p, float scale_p,
                         void *workspace, size_t workspace_bytes, float *out, int64_t ldo, mcd_stream_t stream);
int mcd_rank_reorder_draws_f32(const float *P, int64_t ldp, int64_t N, int64_t C, const int32_t *idx, const float *vals,
                               int64_t K, int64_t top_n, const uint32_t *draws, float p, float scale_p,
                               void *workspace, size_t workspace_bytes, float *out, int64_t ldo, mcd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MCD_B200_H */
