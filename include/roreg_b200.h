/* roreg_b200.h - C ABI of the B200-native RoReg per-pair registration hot path.
 *
 * Drop-in boundary: the reference has no FFI of its own - its hot path is
 * stock PyTorch / NumPy called from test/{matcher,estimator}.py - so each entry point below names the
 * reference function (file:line, relative to the reference root) whose arithmetic it replaces.  The
 * Python mirror of the reference's plugin classes (the modules in roreg_b200/test) binds these with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 = roreg_status; nothing throws across the boundary;
 *   - all data pointers are DEVICE pointers unless the name ends in _host; the caller owns every
 *     buffer; the library owns only the context (group tables in constant/global memory + workspace);
 *   - calls enqueue on the given cudaStream_t (passed as void*) and do not synchronise;
 *   - one context per (device, host thread); calls on one context are not re-entrant;
 *   - descriptors are float32 [n,32,60] (channel-major, group element fastest) exactly as the
 *     reference's .npy cache stores them (test/extractor.py:60); keypoints are float64 [n,3].
 */
#ifndef ROREG_B200_H
#define ROREG_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ROREG_F 32          /* descriptor channels            */
#define ROREG_G 60          /* icosahedral group order        */
#define ROREG_NEI 13        /* group-conv neighbourhood size  */

typedef enum {
  ROREG_OK = 0,
  ROREG_ERR_ARG = -1,        /* bad argument                                     */
  ROREG_ERR_CUDA = -2,       /* CUDA runtime error (see roreg_last_error)        */
  ROREG_ERR_NOMEM = -3,      /* workspace allocation failed                      */
  ROREG_ERR_UNSUPPORTED = -4,/* size outside the compiled limits                 */
  ROREG_ERR_IO = -5          /* a result file could not be written               */
} roreg_status;

typedef struct roreg_ctx roreg_ctx;

int roreg_version(void);
/* perm60x60: utils/group_related/60_60.npy as int32 (P[a][b] = idx(R_b R_a));  nei60x13:
 * Nei_Index_in_SO3_ordered_13.npy as int32;  rot60x3x3: Rotation.npy float64.  HOST pointers. */
int roreg_ctx_create(int device, const int32_t* perm60x60_host, const int32_t* nei60x13_host,
                     const double* rot60x3x3_host, roreg_ctx** out);
int roreg_ctx_destroy(roreg_ctx* ctx);
const char* roreg_last_error(roreg_ctx* ctx);
/* number of kernel launches issued through this context since creation (bench.py gpu_launches) */
int64_t roreg_launch_count(roreg_ctx* ctx);

/* ---- a13  test/matcher.py:69-72 (normalise=1) / network/rot_coh_match.py:346-347 (normalise=0)
 * out[i,:] = mean_g eqv[sample[i],:,g]  (then x/(||x||+1e-5)).  sample may be NULL (identity).      */
int roreg_inv_pool(roreg_ctx* ctx, const float* eqv, const int32_t* sample, int n_out, int normalise,
                   float* out, void* stream);

/* ---- a15  utils/knn_search.py:138-162  modified_knn_matcher.__call__(target, source):
 * for every SOURCE row the k nearest TARGET rows under d = sqrt(sum (a-b)^2 + 1e-7) (difference form,
 * float32), ascending, ties -> lower index.  target [n,f], source [m,f] row-major, f <= 32, k <= 16.  */
int roreg_knn(roreg_ctx* ctx, const float* target, int n, const float* source, int m, int f, int k,
              float* dist, int32_t* idx, void* stream);

/* ---- a13  test/matcher.py:94-106: 1-NN both ways on [n0,32] / [n1,32] invariant features + the mutual
 * check; matches come out in increasing row of f0 as (row in f0, row in f1).  n_matches: device int32[1].
 * nn01 [n0] / nn10 [n1] optional outputs (may be NULL).  mode 0 = float32 difference form (reference
 * arithmetic), mode 1 / 2 / 3 = tensor-core Gram form (tcgen05, 3xTF32; 2 = 8 epilogue warps, 3 = one Gram serves
 * both directions), mode 4 = one tcgen05 Gram per pair in the fp16 two-accumulator split (x = hi + 2^-11 lo',
 * float32-class products), norms folded into the contraction, column direction reduced in registers, row direction
 * finished by an exact float32 re-evaluation of the 8 best candidates in the reference's arithmetic - the fast one.
 * Modes >= 1 need n0 == n1 here (the batched engine has no such restriction on its arena) and |x| <= ~1e4 (fp16 range;
 * the matcher's inputs are L2-normalised).                                                                          */
int roreg_mutual_match(roreg_ctx* ctx, const float* f0, int n0, const float* f1, int n1, int mode,
                       int32_t* matches, int32_t* n_matches, int32_t* nn01, int32_t* nn10, void* stream);

/* ---- a4 / a5  equivariant correlation on K (X row, Y row) pairs.
 * variant 1: test/estimator.py:85-89 Batch_Des2R_torch  cor[a] = sum_{f,g} X[f,P[a,g]] Y[f,g]
 * variant 2: network/rot_coh_match.py:158-163 R-indicator  cor[h] = sum_{f,g} X[f,P[g,h]] Y[f,g]
 * X,Y: descriptor arrays [*,32,60]; idxX/idxY: int32 row indices [K] (NULL = 0..K-1).
 * cor_out [K,60] float32 and argmax_out [K] int32 (first maximal index) are each optional.           */
int roreg_group_corr(roreg_ctx* ctx, const float* X, const int32_t* idxX, const float* Y,
                     const int32_t* idxY, int K, int variant, float* cor_out, int32_t* argmax_out,
                     void* stream);

/* Arithmetic of the 60x60 Gram inside roreg_group_corr / roreg_register_batch: 0 = float32 FMA on CUDA cores
 * (default), 1 = tcgen05 tensor cores with the 3xTF32 split (float32-class products, different summation order),
 * two matches per pipeline item and one CTA per SM, 2 = the same arithmetic, one match per item and two co-resident
 * CTAs per SM (results identical to mode 1), 3 = float16 two-accumulator split (x = hi + 2^-11 lo', same accuracy class),
 * operands converted straight from registers into MN-major tiles - the shared-memory-lean, fastest one.            */
int roreg_set_corr_mode(roreg_ctx* ctx, int mode);

/* ---- a17  test/estimator.py:349-366 + utils/r_eval.py:90-106:
 * R = quat2mat(q)(float32 products) @ float32(Rgroup[pre_idx]),  t = key0 - key1 @ R^T  -> [K,3,4] f64 */
int roreg_hypotheses_from_quat(roreg_ctx* ctx, const float* quat, const int32_t* pre_idx,
                               const double* keys0_m, const double* keys1_m, int K, double* trans,
                               void* stream);

/* ---- a18  test/estimator.py:377-382,426-436  one-shot RANSAC: overlap[h] = sum_i s_i [|k0_i - T_h k1_i|^2
 * < ird^2] / K for h = 0..H-1 with T_h = trans[order[h]] (order NULL = identity); best = first maximum
 * under strict '>' starting from 0 (best_id = -1 if every overlap is 0).  scores: float32 (scores_f64=0)
 * or float64 (=1).  overlaps [H] optional.  best_id int32[1], best_overlap float64[1] on the device.   */
int roreg_ransac_oneshot(roreg_ctx* ctx, const double* k0, const double* k1, const void* scores,
                         int scores_f64, int K, const double* trans, const int32_t* order, int H,
                         double ird, double* overlaps, int32_t* best_id, double* best_overlap,
                         void* stream);

/* ---- a19  test/estimator.py:28-72 refiner.Refine_trans applied at radius ird*2 then ird
 * (:438-439): weighted Kabsch on the inliers, R = U V^T without reflection fix.  T_in: [3,4] device
 * pointer, or, if T_index != NULL, trans[order[*T_index]] as in roreg_ransac_oneshot.  T_out [4,4];
 * inlier_mask [K] uint8 optional = inliers of T_out's predecessor at radius ird (the set the last
 * Kabsch used).                                                                                      */
int roreg_refine(roreg_ctx* ctx, const double* k0, const double* k1, const void* scores, int scores_f64,
                 int K, const double* T_in, const int32_t* order, const int32_t* T_index, double ird,
                 double* T_out, uint8_t* inlier_mask, void* stream);

/* ---- a19  one refiner.Refine_trans call (test/estimator.py:60-72) at the given radius; T_in [3,4].
 * inlier_mask [K] optional = the inliers of T_in at `radius`.                                           */
int roreg_refine_once(roreg_ctx* ctx, const double* k0, const double* k1, const void* scores, int scores_f64,
                      int K, const double* T_in, double radius, double* T_out, uint8_t* inlier_mask,
                      void* stream);

/* ---- a20  test/estimator.py:139-147 Threepps2Tran on device, proper-rotation branch (see DESIGN.md
 * "rank-2 Kabsch"): triplets [H,3] int32 index the K_sel selected matches.  -> trans [H,3,4]          */
int roreg_kabsch3(roreg_ctx* ctx, const double* k0_sel, const double* k1_sel, const int32_t* triplets,
                  int H, double* trans, void* stream);

/* ---- a1-a3 / a22  group-convolution networks (GF: network/group_feat.py:26-45, ET: network/eqv_trans.py:119-138,
 * RD: network/rot_detect.py:43-55).  Activations are channel-last rows [(item*60+g)][C] split into tf32 hi/lo
 * parts; a group convolution gathers the 13 group neighbours (data_process, group_feat.py:20-24) inside the operand
 * load of the tcgen05 GEMM, which has a fused bias / residual / eval-BN / ReLU epilogue.                           */
/* descriptors [*,32,60] -> rows [(item*60+g)][n_src*32]; src_host / rows_host / permute_host are HOST arrays of
 * n_src device pointers / flags; flagged sources are read through P[pre_idx[item]] (eqv_trans.py:126-128).      */
int roreg_pack_descriptors(roreg_ctx* ctx, int n_src, const float* const* src_host, const int32_t* const* rows_host,
                           const int32_t* permute_host, const int32_t* pre_idx, int n_items, const float* bn_scale,
                           const float* bn_shift, int relu, float* out_hi, float* out_lo, void* stream);
/* Group convolution as an IMPLICIT GEMM (data_process + Conv2d(C, O, (1,13)): network/group_feat.py:20-33, ops.py:45-51):
 * out[(item*n_gout+j)][o] = sum_{k,c} act[(item*60 + N[gset[j]][k])][c] W[o][k*C+c] (+ epilogue as roreg_gemm); the
 * 13-neighbour gather happens in the GEMM's operand load (cp.async row copies into the swizzled tile), nothing is materialised.  C % 32 == 0;
 * gset NULL = all 60 group elements (n_gout = 60); act_lo / W_lo / out_lo may be NULL when npass == 1; act_* and W_* bases
 * 16-byte aligned (ROREG_ERR_ARG otherwise).                                                                       */
int roreg_gconv_gemm(roreg_ctx* ctx, const float* act_hi, const float* act_lo, int n_items, int C, const int32_t* gset,
                     int n_gout, const float* W_hi, const float* W_lo, int w_rows, int O, int NT, int npass,
                     const float* bias, const float* residual, int res_ld, float* raw_out, int raw_ld, float* out_hi,
                     float* out_lo, int out_ld, const float* bn_scale, const float* bn_shift, int relu, void* stream);
/* out[r][o] = sum_c A[r][c] W[o][c]; v = out + bias (+ residual[r*res_ld+o]); raw_out = v; act = relu?(v*bn_scale
 * + bn_shift) split hi/lo.  Kdim % 32 == 0; W has w_rows >= ceil(O/NT)*NT rows; npass 1 (TF32) or 3 (3xTF32).    */
int roreg_gemm(roreg_ctx* ctx, const float* A_hi, const float* A_lo, int R, int Kdim, const float* W_hi,
               const float* W_lo, int w_rows, int O, int NT, int npass, const float* bias, const float* residual,
               int res_ld, float* raw_out, int raw_ld, float* act_hi, float* act_lo, int act_ld,
               const float* bn_scale, const float* bn_shift, int relu, void* stream);
int roreg_gf_finalize(roreg_ctx* ctx, const float* conv_out, const float* x, int n, float* eqv_out, void* stream);
int roreg_rd_finalize(roreg_ctx* ctx, const float* raw, int n, float* feat_out, void* stream);
int roreg_row_std60(roreg_ctx* ctx, const float* cor, int n, float* out, void* stream);
int roreg_quat_normalize(roreg_ctx* ctx, const float* q_in, int ld, int K, float* q_out, void* stream);

/* ---- all-pairs 60-rotation correlation (north_star kernel 1; a strict superset of what the reference evaluates,
 * SURVEY.md section 8(0)):  best[n][m] = max_a cor_a(n,m), best_a = argmax_a (first maximum) with
 * cor_a(n,m) = sum_{f,g} X[n,f,P[a,g]] Y[m,f,g]  (test/estimator.py:85-89 on every pair), and per row n the
 * rotation-invariant nearest neighbour nn[n] = argmin_m |X_n|^2 + |Y_m|^2 - 2 best[n][m] with its rotation nn_a and
 * distance nn_dist (nn / nn_a / nn_dist may be NULL).  X_*, Y_* are the channel-last tf32 hi/lo rows [(n*60+g)][32]
 * that roreg_pack_descriptors writes (viewed as [n][1920]); 60 tcgen05 GEMMs whose A-operand K-chunks are permuted
 * through the TMA coordinate, running (max, argmax) in the epilogue.  best [N][M] float32, best_a [N][M] uint8.       */
int roreg_group_corr_allpairs(roreg_ctx* ctx, const float* X_hi, const float* X_lo, int N, const float* Y_hi,
                              const float* Y_lo, int M, int npass, float* best, uint8_t* best_a, int32_t* nn,
                              int32_t* nn_a, float* nn_dist, void* stream);

/* ---- a6-a11  glue of the rotation-coherence matcher Match_ot (network/rot_coh_match.py); the dense layers and the
 * [m,n] score matrices (score_mat :8-12) use roreg_gemm, the R-indicator (:154-163) roreg_group_corr variant 2.   */
/* Knn_index_extract :34-45: first k columns of the descending argsort of every row (ties -> lower column).        */
int roreg_topk_rows(roreg_ctx* ctx, const float* S, int m, int n, int ld, int k, int32_t* idx, void* stream);
/* Knn_feat_extract :48-60: out[r][:] = src[idx[r]][:]                                                              */
int roreg_gather_rows(roreg_ctx* ctx, const float* src, const int32_t* idx, long long n_out, int C, float* out,
                      void* stream);
/* :194-195 knn_coor - coor on coordinates divided by `step` (coor_norm_step :328,:342-343); rows [(i*k+j)][32],
 * 3 used, zero padded                                                                                              */
int roreg_rel_coor(roreg_ctx* ctx, const float* coor, const int32_t* idx, int m, int k, float step, float* out,
                   void* stream);
/* nn.InstanceNorm2d(affine=False) statistics over all P positions of x [P][C] (:18,:68)                           */
int roreg_chan_stats(roreg_ctx* ctx, const float* x, long long P, int C, float* mean, float* rstd, void* stream);
/* GEMM operand preparation: channel concat of <= 3 sources (row broadcast by row_div, optional L2 normalisation
 * over the source's channels :204-206), optional instance norm + ReLU, zero pad to Kout, tf32 hi/lo split.        */
int roreg_prep_rows(roreg_ctx* ctx, int n_src, const float* const* src_host, const int32_t* C_host,
                    const int32_t* rowdiv_host, const int32_t* l2_host, const float* mean, const float* rstd, int relu,
                    long long P, int Kout, float* hi, float* lo, float* plain, void* stream);
/* attention :84-92 + head layout of :105-119 on projected Q [m][32], K / V [m*k][32]                              */
int roreg_mha(roreg_ctx* ctx, const float* Q, const float* Kp, const float* Vp, int m, int k, float* out, void* stream);
/* :203  [R_ind ; max over the points] -> rows [m][128] (120 used)                                                  */
int roreg_rind_rows(roreg_ctx* ctx, const float* rind, int m, float* out128, void* stream);
/* sinkhorn_ot :277-319 (dustbin alpha, `iters` log-domain iterations; the (m+1)x(n+1) coupling matrix is never
 * materialised) + the mutual assignment of :369-378.  u [m+1], v [n+1] are the final potentials.                  */
int roreg_sinkhorn_match(roreg_ctx* ctx, const float* S, int m, int n, int ld, float alpha, int iters, float* u,
                         float* v, int32_t* matches0, float* mscores0, void* stream);

/* ---- batched engine: B independent pairs per call (the throughput path bench.py times) ------------
 * Clouds live in one arena: desc [n_clouds][n][32][60] float32, keys [n_clouds][n][3] float64.
 * pair_cloud [B][2] int32 = (cloud id0, cloud id1).  sample [B][2][keynum] int32 or NULL (identity,
 * keynum = n).  Pipeline = mutual matcher -> Des2R -> (yohoc | yohoo-with-given-hypotheses) -> refine. */
typedef struct {
  int32_t n_clouds, n, keynum, B;
  const float* desc;
  const double* keys;
  const int32_t* pair_cloud;
  const int32_t* sample;
  int32_t nn_mode;            /* roreg_mutual_match mode                                            */
  int32_t estimator;          /* 0 = yohoc (coarse-rotation-guided), 1 = hypotheses given (hyp_host_svd),
                                 2 = stop after Des2R (then roreg_estimate_batch with ET hypotheses)  */
  int32_t max_iter;           /* RANSAC iterations (hypotheses scored)                              */
  double ird;                 /* inlier radius, cfg.ransac_ird                                      */
  uint64_t seed;              /* device RNG seed (estimator 0 with triplets == NULL)                */
  const int32_t* triplets;    /* [B][max_iter][3] host-drawn triplets (parity mode) or NULL         */
  const double* hyp_host_svd; /* [B][max_iter][3][4] hypotheses computed by the caller or NULL      */
  /* outputs */
  int32_t* matches;           /* [B][keynum][2] original keypoint indices (id0, id1)                */
  int32_t* n_matches;         /* [B]                                                                */
  int32_t* dr_index;          /* [B][keynum]                                                        */
  double* poses;              /* [B][4][4]                                                          */
  int32_t* recall;            /* [B] 0-based index of the winning hypothesis in the scored order, -1 = none */
  double* best_overlap;       /* [B]                                                                */
} roreg_batch;

int roreg_register_batch(roreg_ctx* ctx, const roreg_batch* batch, void* stream);
/* Pipelined form of roreg_register_batch for back-to-back batches (same per-pair loops, test/matcher.py:64-109 +
 * test/estimator.py:102-111,163-242): a call enqueues the pooling, NN and Des2R of `batch` and, beside the pooling, the RANSAC
 * tail (hypotheses, scoring, refinement) of the batch given to the PREVIOUS call.  The outputs poses / recall / best_overlap
 * of a batch are therefore complete only after the next roreg_register_batch_pipelined or roreg_register_batch_flush on the
 * same stream; its keys, pair_cloud, matches, n_matches and dr_index arrays must stay untouched until then (use two sets of
 * buffers).  estimator 0 or 1; per-stage timing is not recorded.  Results equal roreg_register_batch's.               */
int roreg_register_batch_pipelined(roreg_ctx* ctx, const roreg_batch* batch, void* stream);
int roreg_register_batch_flush(roreg_ctx* ctx, void* stream);
/* one-shot RANSAC + refinement (test/estimator.py:426-439) of every pair of the batch on caller-provided
 * hypotheses [B][max_iter][3][4] float64 (n_hyp [B] valid per pair, NULL = all), after estimator = 2.               */
int roreg_estimate_batch(roreg_ctx* ctx, const roreg_batch* batch, const double* hyps, const int32_t* n_hyp,
                         void* stream);

/* Per-stage device timing of the last roreg_register_batch call (measurement hook for bench.py; events are
 * recorded on the call's stream).  Stages: 0 inv_pool, 1 nn (both directions), 2 mutual compaction,
 * 3 group correlation (Des2R), 4 hypothesis generation, 5 scoring + selection, 6 refine.
 * roreg_get_stage_ms synchronises on the last event and writes ROREG_N_STAGES floats (milliseconds). */
#define ROREG_N_STAGES 7
int roreg_set_timing(roreg_ctx* ctx, int enable);

/* Schedule of roreg_register_batch.  0 (default): every stage on the caller's stream.  1: a batch of >= 8 pairs is split in two halves whose stages overlap on two
 * internal streams (pooling of the second half beside the NN of the first, RANSAC tail of the first beside the NN of the
 * second); the call still only returns work ordered after / before the caller's stream.  Per-stage timing
 * (roreg_set_timing) and estimator == 2 always use the serial schedule.  Results are identical; on a B200 the serial
 * schedule measured faster, hence the default.                                                          */
int roreg_set_overlap(roreg_ctx* ctx, int enable);
/* Arithmetic of the one-shot scoring inside roreg_ransac_oneshot / roreg_register_batch (test/estimator.py:377-382).
 * 0 (default): every point test in float64, as the reference.  1: float32 pre-filter - a test is decided in float32 when its
 * squared distance is outside [ird^2 - m, ird^2 + m] (m = a rigorous bound of the float32 evaluation error, kernels_ransac.cuh)
 * and repeated in float64 otherwise; decisions and sums are identical to mode 0 by construction.                      */
int roreg_set_score_mode(roreg_ctx* ctx, int mode);
int roreg_get_stage_ms(roreg_ctx* ctx, float* ms_host);

/* ---- stage I/O: the four per-pair files the reference's plugins leave for its evaluator, written from C
 * so that writer threads run without the Python GIL (no GPU involved; any path may be NULL = skip that file):
 *   match_path     int64 [K,2]  .npy   (test/matcher.py:108, np.save of match_pps)
 *   scores_path    float64 [K]  .npy   = ones (test/matcher.py:109)
 *   dr_index_path  int64 [K]    .npy   (test/estimator.py:111)
 *   npz_path       np.savez(trans = pose [4,4] float64, recalltime = int64 scalar)   (test/estimator.py:242)
 * NPY format 1.0 / ZIP of stored members: exactly what np.load reads back.                                              */
int roreg_write_pair_files(const char* match_path, const char* scores_path, const char* dr_index_path, const char* npz_path,
                           const int64_t* matches, const int64_t* dr_index, int K, const double* pose4x4,
                           long long recalltime);

#ifdef __cplusplus
}
#endif
#endif
