// roreg_capi.cu - extern "C" entry points of libroreg_b200.so (see include/roreg_b200.h).
#include "common.cuh"
#include "kernels_match.cuh"
#include "kernels_corr.cuh"
#include "kernels_ransac.cuh"
#include "kernels_nn_tc.cuh"
#include "kernels_corr_tc.cuh"
#include "kernels_corr_tc2.cuh"
#include "kernels_nn_tc4.cuh"
#include "kernels_corr_tc3.cuh"
#include "kernels_gemm_tc.cuh"
#include "kernels_allpairs_tc.cuh"
#include "kernels_gconv.cuh"
#include "kernels_matchot.cuh"
#include "host_io.inl"

using namespace roreg;

static inline unsigned rr_blocks(roreg_ctx* c, long long work, int per_block) {
  long long b = (work + per_block - 1) / per_block;
  const long long cap = (long long)c->sm_count * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" {

int roreg_version(void) { return 100; }

int roreg_ctx_create(int device, const int32_t* perm, const int32_t* nei, const double* rot, roreg_ctx** out) {
  if (!perm || !nei || !rot || !out) return ROREG_ERR_ARG;
  roreg_ctx* c = new roreg_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->overlap = 0;
  c->score_mode = 0;
  c->pipe = nullptr;
  *out = c;
  RR_CUDA(c, cudaSetDevice(device));
  cudaDeviceProp prop;
  RR_CUDA(c, cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  uint8_t p8[3600], pT8[3600];
  float r32[540];
  for (int a = 0; a < 60; ++a)
    for (int g = 0; g < 60; ++g) {
      const int v = perm[a * 60 + g];
      if (v < 0 || v >= 60) { snprintf(c->err, sizeof(c->err), "perm table entry out of range"); return ROREG_ERR_ARG; }
      p8[a * 60 + g] = (uint8_t)v;
      pT8[g * 60 + a] = (uint8_t)v;       // pT8[h][g] = P[g][h]
    }
  for (int i = 0; i < 540; ++i) r32[i] = (float)rot[i];
  memcpy(c->h_perm8, p8, 3600);
  RR_CUDA(c, cudaMalloc(&c->d_perm8, 3600));
  RR_CUDA(c, cudaMalloc(&c->d_permT8, 3600));
  RR_CUDA(c, cudaMalloc(&c->d_nei, 60 * 13 * sizeof(int32_t)));
  RR_CUDA(c, cudaMalloc(&c->d_rot32, 540 * sizeof(float)));
  RR_CUDA(c, cudaMalloc(&c->d_rot64, 540 * sizeof(double)));
  RR_CUDA(c, cudaMemcpy(c->d_perm8, p8, 3600, cudaMemcpyHostToDevice));
  RR_CUDA(c, cudaMemcpy(c->d_permT8, pT8, 3600, cudaMemcpyHostToDevice));
  RR_CUDA(c, cudaMemcpy(c->d_nei, nei, 60 * 13 * sizeof(int32_t), cudaMemcpyHostToDevice));
  RR_CUDA(c, cudaMemcpy(c->d_rot32, r32, sizeof(r32), cudaMemcpyHostToDevice));
  RR_CUDA(c, cudaMemcpy(c->d_rot64, rot, 540 * sizeof(double), cudaMemcpyHostToDevice));
  return ROREG_OK;
}

static void pipe_destroy(roreg_ctx* c);
int roreg_ctx_destroy(roreg_ctx* c) {
  if (!c) return ROREG_ERR_ARG;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->ev[0]) for (int i = 0; i <= ROREG_N_STAGES; ++i) cudaEventDestroy(c->ev[i]);
  if (c->s_tc) {
    cudaStreamDestroy(c->s_tc); cudaStreamDestroy(c->s_light); cudaEventDestroy(c->ev_fork);
    for (int h = 0; h < 2; ++h) { cudaEventDestroy(c->ev_pool[h]); cudaEventDestroy(c->ev_corr[h]); cudaEventDestroy(c->ev_join[h]); }
  }
  pipe_destroy(c);
  cudaFree(c->d_perm8); cudaFree(c->d_permT8); cudaFree(c->d_nei); cudaFree(c->d_rot32); cudaFree(c->d_rot64);
  if (c->ws) cudaFree(c->ws);
  delete c;
  return ROREG_OK;
}

int roreg_set_corr_mode(roreg_ctx* c, int mode) {
  RR_ARG(c, mode >= 0 && mode <= 3);
  c->corr_mode = mode;
  return ROREG_OK;
}

int roreg_set_score_mode(roreg_ctx* c, int mode) {
  if (!c) return ROREG_ERR_ARG;
  RR_ARG(c, mode == 0 || mode == 1);
  c->score_mode = mode;
  return ROREG_OK;
}

int roreg_set_timing(roreg_ctx* c, int enable) {
  if (!c) return ROREG_ERR_ARG;
  if (enable && !c->ev[0])
    for (int i = 0; i <= ROREG_N_STAGES; ++i) RR_CUDA(c, cudaEventCreate(&c->ev[i]));
  c->timing = enable; c->ev_valid = 0;
  return ROREG_OK;
}

int roreg_get_stage_ms(roreg_ctx* c, float* ms) {
  RR_ARG(c, ms != nullptr);
  if (!c->timing || !c->ev_valid) { snprintf(c->err, sizeof(c->err), "no timed batch call recorded"); return ROREG_ERR_ARG; }
  RR_CUDA(c, cudaEventSynchronize(c->ev[ROREG_N_STAGES]));
  for (int i = 0; i < ROREG_N_STAGES; ++i) RR_CUDA(c, cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
  return ROREG_OK;
}

const char* roreg_last_error(roreg_ctx* c) { return c ? c->err : "null context"; }
int64_t roreg_launch_count(roreg_ctx* c) { return c ? c->launches : -1; }

// ------------------------------------------------------------------------------------------------
int roreg_inv_pool(roreg_ctx* c, const float* eqv, const int32_t* sample, int n_out, int normalise,
                   float* out, void* stream) {
  RR_ARG(c, n_out >= 0);
  if (n_out == 0) return ROREG_OK;                 // empty input: nothing to do (pointers may be NULL)
  RR_ARG(c, eqv && out);
  PoolArgs a{eqv, nullptr, sample, 0, n_out, n_out, normalise, out};
  inv_pool_kernel<<<(n_out + 7) / 8, 256, 0, (cudaStream_t)stream>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

static int launch_nn(roreg_ctx* c, int mode, const float* src, const float* tgt, long long src_ps, long long tgt_ps,
                     int n_src, int n_tgt, int32_t* out_idx, float* out_dist, long long out_ps, int B, cudaStream_t st) {
  if (n_src == 0) return ROREG_OK;
  (void)mode;
  NNArgs a{src, tgt, src_ps, tgt_ps, n_src, n_tgt, out_idx, out_dist, out_ps};
  nn_diff_kernel<<<dim3((n_src + 63) / 64, B), 256, 0, st>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_knn(roreg_ctx* c, const float* target, int n, const float* source, int m, int f, int k,
              float* dist, int32_t* idx, void* stream) {
  RR_ARG(c, target && source && idx && dist);
  RR_ARG(c, n >= 1 && m >= 0 && f >= 1 && f <= 32 && k >= 1 && k <= 16 && k <= n);
  if (m == 0) return ROREG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (f == 32 && k == 1) return launch_nn(c, 0, source, target, 0, 0, m, n, idx, dist, 0, 1, st);
  knn_small_kernel<16><<<(m + 127) / 128, 128, 128 * f * sizeof(float), st>>>(target, n, source, m, f, k, dist, idx);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_mutual_match(roreg_ctx* c, const float* f0, int n0, const float* f1, int n1, int mode,
                       int32_t* matches, int32_t* n_matches, int32_t* nn01, int32_t* nn10, void* stream) {
  RR_ARG(c, f0 && f1 && matches && n_matches && n0 >= 1 && n1 >= 1 && (mode >= 0 && mode <= 4));
  cudaStream_t st = (cudaStream_t)stream;
  if (mode >= 1 && n0 != n1) {
    snprintf(c->err, sizeof(c->err), "nn mode 1 (tcgen05 Gram) needs n0 == n1 in the single-pair entry");
    return ROREG_ERR_UNSUPPORTED;
  }
  size_t need = rr_align(sizeof(int32_t) * n0) + rr_align(sizeof(int32_t) * n1) + 4096;
  if (mode == 4) need += rr_align(sizeof(float) * 2 * (size_t)n0 * RR_F) + nn_tc4_workspace_bytes(1, n0) + 2048;
  else if (mode >= 1) need += rr_align(sizeof(float) * 2 * (size_t)n0 * RR_F) + nn_tc_workspace_bytes(2LL * n0);
  int rc = rr_ws_reserve(c, need);
  if (rc) return rc;
  rr_arena ar{(char*)c->ws, 0};
  int32_t* w01 = nn01 ? nn01 : ar.take<int32_t>(n0);
  int32_t* w10 = nn10 ? nn10 : ar.take<int32_t>(n1);
  if (mode == 4) {
    float* inv2 = ar.take<float>(2 * (size_t)n0 * RR_F);
    const size_t NT = (size_t)(n0 + TC_BM - 1) / TC_BM;
    ar.off = (ar.off + 1023) & ~size_t(1023);
    uint8_t* img = ar.take<uint8_t>(2 * NT * T4_TILE_BYTES);
    float* rowval = ar.take<float>(2 * NT * NT * TC_BM);
    uint8_t* rowgid = ar.take<uint8_t>(2 * NT * NT * TC_BM);
    int32_t* bchunk = ar.take<int32_t>((size_t)n0);
    RR_CUDA(c, cudaMemcpyAsync(inv2, f0, sizeof(float) * (size_t)n0 * RR_F, cudaMemcpyDeviceToDevice, st));
    RR_CUDA(c, cudaMemcpyAsync(inv2 + (size_t)n0 * RR_F, f1, sizeof(float) * (size_t)n0 * RR_F, cudaMemcpyDeviceToDevice, st));
    if ((rc = nn_tc4_launch_both(c, inv2, n0, 1, img, rowval, rowgid, bchunk, w01, w10, st))) return rc;
  } else if (mode >= 1) {
    float* inv2 = ar.take<float>(2 * (size_t)n0 * RR_F);
    float* Ahat = ar.take<float>(2 * (size_t)n0 * TC_KEXT);
    float* Bhat = ar.take<float>(2 * (size_t)n0 * TC_KEXT);
    float* nh = ar.take<float>(2 * (size_t)n0);
    unsigned long long* cb = ar.take<unsigned long long>((size_t)n0);
    RR_CUDA(c, cudaMemcpyAsync(inv2, f0, sizeof(float) * (size_t)n0 * RR_F, cudaMemcpyDeviceToDevice, st));
    RR_CUDA(c, cudaMemcpyAsync(inv2 + (size_t)n0 * RR_F, f1, sizeof(float) * (size_t)n0 * RR_F, cudaMemcpyDeviceToDevice, st));
    if (mode == 3) { if ((rc = nn_tc3_launch_both(c, inv2, n0, 1, Ahat, nh, cb, w01, w10, st))) return rc; }
    else if (mode == 2) { if ((rc = nn_tc2_launch_both(c, inv2, n0, 1, Ahat, nh, w01, w10, st))) return rc; }
    else if ((rc = nn_tc_launch_both(c, inv2, n0, 1, Ahat, Bhat, nh, w01, w10, st))) return rc;
  } else {
    if ((rc = launch_nn(c, mode, f0, f1, 0, 0, n0, n1, w01, nullptr, 0, 1, st))) return rc;   // KNN(feats1, feats0): rows of cloud0 search cloud1
    if ((rc = launch_nn(c, mode, f1, f0, 0, 0, n1, n0, w10, nullptr, 0, 1, st))) return rc;
  }
  CompactArgs ca{w01, w10, n0, n1, 0, nullptr, 0, matches, n0 < n1 ? n0 : n1, n_matches};
  mutual_compact_kernel<<<1, 1024, 0, st>>>(ca);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_group_corr(roreg_ctx* c, const float* X, const int32_t* idxX, const float* Y, const int32_t* idxY,
                     int K, int variant, float* cor_out, int32_t* argmax_out, void* stream) {
  RR_ARG(c, X && Y && K >= 0 && (variant == 1 || variant == 2) && (cor_out || argmax_out));
  if (K == 0) return ROREG_OK;
  if (c->corr_mode >= 1) {
    CorrTcArgs t{nullptr, nullptr, idxX, idxY, 1, nullptr, 0, nullptr, K, 1, (variant == 1) ? c->d_perm8 : c->d_permT8, cor_out, argmax_out, 3, 0, nullptr};
    return c->corr_mode == 3 ? group_corr_tc3_launch(c, X, Y, t, (cudaStream_t)stream)
         : c->corr_mode == 2 ? group_corr_tc2_launch(c, X, Y, t, (cudaStream_t)stream)
                             : group_corr_tc_launch(c, X, Y, t, (cudaStream_t)stream);   // indices are trusted
  }
  CorrArgs a{};
  a.X = X; a.Y = Y; a.idxX = idxX; a.idxY = idxY; a.idx_stride = 1; a.pair_cloud = nullptr; a.n = 0;
  a.n_matches = nullptr; a.K = K; a.B = 1; a.tab = (variant == 1) ? c->d_perm8 : c->d_permT8;
  a.cor_out = cor_out; a.argmax_out = argmax_out;
  const int grid = K < c->sm_count * 8 ? K : c->sm_count * 8;
  group_corr_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_hypotheses_from_quat(roreg_ctx* c, const float* quat, const int32_t* pre_idx, const double* k0,
                               const double* k1, int K, double* trans, void* stream) {
  RR_ARG(c, quat && pre_idx && k0 && k1 && trans && K >= 0);
  if (K == 0) return ROREG_OK;
  hyp_from_quat_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(quat, pre_idx, k0, k1, K, c->d_rot32, trans);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

static MatchView single_view(const double* k0, const double* k1, const void* scores, int scores_f64, int K) {
  MatchView v{};
  v.keys0 = k0; v.keys1 = k1; v.pair_cloud = nullptr; v.n = 0; v.matches = nullptr; v.cap = K;
  v.n_matches = nullptr; v.K = K; v.scores = scores; v.scores_f64 = scores_f64; v.scores_pair_stride = 0;
  return v;
}

static int score_and_select(roreg_ctx* c, const MatchView& mv, int cap, const double* hyps, long long hyp_ps,
                            const int32_t* order, const int32_t* n_hyp, int H, double ird, double* partial,
                            double* overlaps, int32_t* best_id, double* best_overlap, int B, cudaStream_t st) {
  const int tiles = (cap + RR_SCORE_TILE - 1) / RR_SCORE_TILE;
  ScoreArgs sa{mv, hyps, hyp_ps, order, n_hyp, H, ird * ird, partial, tiles, ird};
  if (c->score_mode == 1) {
    const int hx = (H + 255) / 256;
    const long long items = (long long)hx * tiles * B;
    long long grid = items;
    // resident CTAs of the persistent scoring kernel: 2 per SM by default (run c1/c2: 24.9 k pairs/s against 24.1 k uncapped
    // and 24.2 / 24.4 k at 1 / 3 - it leaves room for the pooling kernel that co-runs in the pipelined schedule); 0 = uncapped
    long long per_sm = 2;
    if (const char* e = getenv("ROREG_SCORE_CTAS_PER_SM")) per_sm = atoi(e);
    if (per_sm >= 1 && per_sm * c->sm_count < grid) grid = per_sm * c->sm_count;
    if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;
    ransac_score_pre_kernel<<<(unsigned)grid, 256, 0, st>>>(sa, hx, items);
  } else ransac_score_kernel<<<dim3((H + 255) / 256, tiles, B), 256, 0, st>>>(sa);
  RR_LAUNCH_CHECK(c);
  SelectArgs se{partial, tiles, H, mv.n_matches, mv.K, overlaps, best_id, best_overlap};
  ransac_select_kernel<<<B, 256, 0, st>>>(se);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_ransac_oneshot(roreg_ctx* c, const double* k0, const double* k1, const void* scores, int scores_f64,
                         int K, const double* trans, const int32_t* order, int H, double ird, double* overlaps,
                         int32_t* best_id, double* best_overlap, void* stream) {
  RR_ARG(c, k0 && k1 && trans && best_id && best_overlap && K >= 1 && H >= 1 && ird > 0);
  const int tiles = (K + RR_SCORE_TILE - 1) / RR_SCORE_TILE;
  int rc = rr_ws_reserve(c, rr_align(sizeof(double) * (size_t)tiles * H) + 4096);
  if (rc) return rc;
  rr_arena ar{(char*)c->ws, 0};
  double* partial = ar.take<double>((size_t)tiles * H);
  MatchView mv = single_view(k0, k1, scores, scores_f64, K);
  return score_and_select(c, mv, K, trans, 0, order, nullptr, H, ird, partial, overlaps, best_id, best_overlap, 1,
                          (cudaStream_t)stream);
}

int roreg_refine(roreg_ctx* c, const double* k0, const double* k1, const void* scores, int scores_f64, int K,
                 const double* T_in, const int32_t* order, const int32_t* T_index, double ird, double* T_out,
                 uint8_t* inlier_mask, void* stream) {
  RR_ARG(c, k0 && k1 && T_in && T_out && K >= 1 && ird > 0);
  RefineArgs a{};
  a.mv = single_view(k0, k1, scores, scores_f64, K);
  if (T_index) { a.T_in = nullptr; a.hyps = T_in; a.hyp_pair_stride = 0; a.order = order; a.best_id = T_index; }
  else { a.T_in = T_in; a.T_pair_stride = 0; }
  a.rad0 = ird * 2.0; a.rad1 = ird; a.rounds = 2; a.T_out = T_out; a.inlier_mask = inlier_mask; a.mask_stride = K;
  refine_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_refine_once(roreg_ctx* c, const double* k0, const double* k1, const void* scores, int scores_f64, int K,
                      const double* T_in, double radius, double* T_out, uint8_t* inlier_mask, void* stream) {
  RR_ARG(c, k0 && k1 && T_in && T_out && K >= 1 && radius > 0);
  RefineArgs a{};
  a.mv = single_view(k0, k1, scores, scores_f64, K);
  a.T_in = T_in; a.T_pair_stride = 0;
  a.rad0 = radius; a.rad1 = radius; a.rounds = 1; a.T_out = T_out; a.inlier_mask = inlier_mask; a.mask_stride = K;
  refine_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_kabsch3(roreg_ctx* c, const double* k0s, const double* k1s, const int32_t* triplets, int H,
                  double* trans, void* stream) {
  RR_ARG(c, k0s && k1s && triplets && trans && H >= 0);
  if (H == 0) return ROREG_OK;
  kabsch3_kernel<<<(H + 127) / 128, 128, 0, (cudaStream_t)stream>>>(k0s, k1s, triplets, H, trans);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

// ------------------------------------------------------------------------------------------------
// group-convolution networks (GF / ET / RD): pack, implicit group-convolution GEMM, dense GEMM, tails
// ------------------------------------------------------------------------------------------------
int roreg_pack_descriptors(roreg_ctx* c, int n_src, const float* const* src_host, const int32_t* const* rows_host,
                           const int32_t* permute_host, const int32_t* pre_idx, int n_items, const float* bn_scale,
                           const float* bn_shift, int relu, float* out_hi, float* out_lo, void* stream) {
  RR_ARG(c, n_src >= 1 && n_src <= 4 && src_host && out_hi && n_items >= 0);
  if (n_items == 0) return ROREG_OK;
  PackArgs a{};
  for (int s = 0; s < n_src; ++s) {
    RR_ARG(c, src_host[s] != nullptr);
    a.src[s] = src_host[s]; a.rows[s] = rows_host ? rows_host[s] : nullptr; a.permute[s] = permute_host ? permute_host[s] : 0;
  }
  a.n_src = n_src; a.pre_idx = pre_idx; a.perm = c->d_perm8; a.n_items = n_items;
  a.bn_scale = bn_scale; a.bn_shift = bn_shift; a.relu = relu; a.out_hi = out_hi; a.out_lo = out_lo;
  pack_desc_kernel<<<dim3(n_items, n_src), 256, 0, (cudaStream_t)stream>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_gemm(roreg_ctx* c, const float* A_hi, const float* A_lo, int R, int Kdim, const float* W_hi, const float* W_lo,
               int w_rows, int O, int NT, int npass, const float* bias, const float* residual, int res_ld, float* raw_out,
               int raw_ld, float* act_hi, float* act_lo, int act_ld, const float* bn_scale, const float* bn_shift, int relu,
               void* stream) {
  RR_ARG(c, A_hi && W_hi && R >= 0 && O >= 1 && NT >= 16 && w_rows >= NT && (raw_out || act_hi));
  GemmArgs a{};
  a.R = R; a.Kdim = Kdim; a.O = O; a.NT = NT; a.n_ntiles = (O + NT - 1) / NT; a.npass = npass;
  RR_ARG(c, (long long)a.n_ntiles * NT <= w_rows);
  a.bias = bias; a.residual = residual; a.res_ld = res_ld; a.raw_out = raw_out; a.raw_ld = raw_ld;
  a.act_hi = act_hi; a.act_lo = act_lo; a.act_ld = act_ld; a.bn_scale = bn_scale; a.bn_shift = bn_shift; a.relu = relu;
  RR_ARG(c, (bn_scale == nullptr) == (bn_shift == nullptr));
  return gemm_tc_launch(c, A_hi, A_lo, W_hi, W_lo, w_rows, a, (cudaStream_t)stream);
}

int roreg_gconv_gemm(roreg_ctx* c, const float* act_hi, const float* act_lo, int n_items, int C, const int32_t* gset, int n_gout,
                     const float* W_hi, const float* W_lo, int w_rows, int O, int NT, int npass, const float* bias,
                     const float* residual, int res_ld, float* raw_out, int raw_ld, float* out_hi, float* out_lo, int out_ld,
                     const float* bn_scale, const float* bn_shift, int relu, void* stream) {
  RR_ARG(c, act_hi && W_hi && n_items >= 0 && C >= 32 && (C % 32) == 0 && n_gout >= 1 && n_gout <= 60 && O >= 1 && NT >= 16 && w_rows >= NT);
  RR_ARG(c, (raw_out || out_hi) && (long long)n_items * 60 < (1LL << 31));
  // 16-byte row copies (cp.async) and TMA both need 16-byte aligned activation / weight bases
  RR_ARG(c, (reinterpret_cast<uintptr_t>(act_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(act_lo) & 15) == 0 && (reinterpret_cast<uintptr_t>(W_hi) & 15) == 0);
  if (n_items == 0) return ROREG_OK;
  GemmArgs a{};
  a.R = n_items * n_gout; a.Kdim = 13 * C; a.O = O; a.NT = NT; a.n_ntiles = (O + NT - 1) / NT; a.npass = npass;
  RR_ARG(c, (long long)a.n_ntiles * NT <= w_rows);
  a.bias = bias; a.residual = residual; a.res_ld = res_ld; a.raw_out = raw_out; a.raw_ld = raw_ld;
  a.act_hi = out_hi; a.act_lo = out_lo; a.act_ld = out_ld; a.bn_scale = bn_scale; a.bn_shift = bn_shift; a.relu = relu;
  RR_ARG(c, (bn_scale == nullptr) == (bn_shift == nullptr));
  a.g_C = C; a.g_ng = n_gout; a.g_nei = c->d_nei; a.g_set = gset;
  return gemm_tc_launch(c, act_hi, act_lo, W_hi, W_lo, w_rows, a, (cudaStream_t)stream, (long long)n_items * 60);
}

int roreg_gf_finalize(roreg_ctx* c, const float* conv_out, const float* x, int n, float* eqv_out, void* stream) {
  RR_ARG(c, conv_out && x && eqv_out && n >= 0);
  if (n == 0) return ROREG_OK;
  gf_finalize_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(conv_out, x, n, eqv_out);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_rd_finalize(roreg_ctx* c, const float* raw, int n, float* feat_out, void* stream) {
  RR_ARG(c, raw && feat_out && n >= 0);
  if (n == 0) return ROREG_OK;
  rd_finalize_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(raw, n, feat_out);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_row_std60(roreg_ctx* c, const float* cor, int n, float* out, void* stream) {
  RR_ARG(c, cor && out && n >= 0);
  if (n == 0) return ROREG_OK;
  row_std60_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cor, n, out);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_quat_normalize(roreg_ctx* c, const float* q_in, int ld, int K, float* q_out, void* stream) {
  RR_ARG(c, q_in && q_out && ld >= 4 && K >= 0);
  if (K == 0) return ROREG_OK;
  quat_normalize_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(q_in, ld, K, q_out);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

// all-pairs 60-rotation correlation (north_star kernel 1): ONE persistent tcgen05 kernel with the rotation
// loop inside and the running (max, argmax) in registers (kernels_allpairs_tc.cuh); X / Y are the channel-last tf32-split
// descriptors from roreg_pack_descriptors.  ROREG_ALLPAIRS_LAUNCHES=1 selects round 1's 60 K-permuted GEMM launches (A/B).
int roreg_group_corr_allpairs(roreg_ctx* c, const float* X_hi, const float* X_lo, int N, const float* Y_hi, const float* Y_lo, int M,
                              int npass, float* best, uint8_t* best_a, int32_t* nn, int32_t* nn_a, float* nn_dist, void* stream) {
  RR_ARG(c, X_hi && Y_hi && best && best_a && N >= 1 && M >= 1 && (npass == 1 || (npass == 3 && X_lo && Y_lo)));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = rr_ws_reserve(c, rr_align(sizeof(int32_t) * 60 * 60) + rr_align(sizeof(float) * N) + rr_align(sizeof(float) * M) + 4096);
  if (rc) return rc;
  rr_arena ar{(char*)c->ws, 0};
  int32_t* cols = ar.take<int32_t>(3600); float* nx = ar.take<float>(N); float* ny = ar.take<float>(M);
  if (!getenv("ROREG_ALLPAIRS_LAUNCHES")) {
    if ((rc = allpairs_tc_launch(c, X_hi, X_lo, N, Y_hi, Y_lo, M, npass, best, best_a, st))) return rc;
  } else {
    // column coordinate of k-chunk g for rotation a: P[a][g]*32   (cor_a = sum_g <X[:,P[a,g]], Y[:,g]>, test/estimator.py:85-89)
    {
      static int32_t h_cols[3600];
      for (int i = 0; i < 3600; ++i) h_cols[i] = (int32_t)c->h_perm8[i] * 32;
      RR_CUDA(c, cudaMemcpyAsync(cols, h_cols, sizeof(h_cols), cudaMemcpyHostToDevice, st));
    }
    fill_f32_kernel<<<rr_blocks(c, (long long)N * M, 256), 256, 0, st>>>(best, (long long)N * M, -INFINITY);
    RR_LAUNCH_CHECK(c);
    for (int a_id = 0; a_id < 60; ++a_id) {
      GemmArgs g{};
      g.R = N; g.Kdim = 1920; g.O = M; g.NT = M >= 256 ? 256 : ((M + 15) / 16) * 16; g.n_ntiles = (M + g.NT - 1) / g.NT; g.npass = npass;
      g.raw_out = best; g.raw_ld = M; g.a_cols = cols + a_id * 60; g.amax_arg = best_a; g.amax_id = a_id;
      if ((rc = gemm_tc_launch(c, X_hi, X_lo, Y_hi, Y_lo, M, g, st))) return rc;
    }
  }
  if (nn) {
    RR_ARG(c, nn_a && nn_dist);
    row_sqnorm_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(X_hi, X_lo, N, 1920, nx);
    RR_LAUNCH_CHECK(c);
    row_sqnorm_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(Y_hi, Y_lo, M, 1920, ny);
    RR_LAUNCH_CHECK(c);
    allpairs_rowmin_kernel<<<N, 256, 0, st>>>(best, best_a, N, M, nx, ny, nn, nn_a, nn_dist);
    RR_LAUNCH_CHECK(c);
  }
  return ROREG_OK;
}

// ------------------------------------------------------------------------------------------------
// Match_ot glue (network/rot_coh_match.py)
// ------------------------------------------------------------------------------------------------

int roreg_topk_rows(roreg_ctx* c, const float* S, int m, int n, int ld, int k, int32_t* idx, void* stream) {
  RR_ARG(c, S && idx && m >= 0 && n >= 1 && ld >= n && k >= 1 && k <= 16 && k <= n && n <= 11000);
  if (m == 0) return ROREG_OK;
  topk_rows_kernel<<<m, 256, n * sizeof(float), (cudaStream_t)stream>>>(S, n, ld, k, idx);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_gather_rows(roreg_ctx* c, const float* src, const int32_t* idx, long long n_out, int C, float* out, void* stream) {
  RR_ARG(c, src && idx && out && n_out >= 0 && C >= 4 && (C % 4) == 0);
  if (n_out == 0) return ROREG_OK;
  gather_rows_kernel<<<rr_blocks(c, n_out * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n_out, C, out);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_rel_coor(roreg_ctx* c, const float* coor, const int32_t* idx, int m, int k, float step, float* out, void* stream) {
  RR_ARG(c, coor && idx && out && m >= 0 && k >= 1 && step > 0);
  if (m == 0) return ROREG_OK;
  rel_coor_kernel<<<rr_blocks(c, (long long)m * k * 32, 256), 256, 0, (cudaStream_t)stream>>>(coor, idx, m, k, step, out);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_chan_stats(roreg_ctx* c, const float* x, long long P, int C, float* mean, float* rstd, void* stream) {
  RR_ARG(c, x && mean && rstd && P >= 1 && C >= 1 && C <= 256 && (256 % C) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = rr_ws_reserve(c, 4096 + sizeof(double) * 2 * 256);
  if (rc) return rc;
  double* acc = reinterpret_cast<double*>(c->ws);
  RR_CUDA(c, cudaMemsetAsync(acc, 0, sizeof(double) * 2 * C, st));
  chan_stats_kernel<<<rr_blocks(c, P, 256 / C * 8), 256, 0, st>>>(x, P, C, acc);
  RR_LAUNCH_CHECK(c);
  chan_stats_finish_kernel<<<(C + 63) / 64, 64, 0, st>>>(acc, P, C, mean, rstd);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_prep_rows(roreg_ctx* c, int n_src, const float* const* src_host, const int32_t* C_host, const int32_t* rowdiv_host,
                    const int32_t* l2_host, const float* mean, const float* rstd, int relu, long long P, int Kout, float* hi,
                    float* lo, float* plain, void* stream) {
  RR_ARG(c, n_src >= 1 && n_src <= 3 && src_host && C_host && hi && lo && P >= 0 && Kout >= 1);
  RR_ARG(c, (mean == nullptr) == (rstd == nullptr));
  if (P == 0) return ROREG_OK;
  PrepArgs a{};
  int tot = 0;
  for (int s = 0; s < n_src; ++s) {
    RR_ARG(c, src_host[s] != nullptr && C_host[s] >= 1);
    a.src[s] = src_host[s]; a.C[s] = C_host[s]; a.row_div[s] = rowdiv_host ? rowdiv_host[s] : 1; a.l2norm[s] = l2_host ? l2_host[s] : 0;
    RR_ARG(c, a.row_div[s] >= 1);
    tot += C_host[s];
  }
  RR_ARG(c, tot <= Kout && (mean == nullptr || n_src == 1));
  a.n_src = n_src; a.mean = mean; a.rstd = rstd; a.relu = relu; a.P = P; a.Kout = Kout; a.hi = hi; a.lo = lo; a.plain = plain;
  prep_rows_kernel<<<(unsigned)((P + 7) / 8), 256, 0, (cudaStream_t)stream>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_mha(roreg_ctx* c, const float* Q, const float* Kp, const float* Vp, int m, int k, float* out, void* stream) {
  RR_ARG(c, Q && Kp && Vp && out && m >= 0 && k >= 1 && k <= 16);
  if (m == 0) return ROREG_OK;
  mha_kernel<<<(m + 7) / 8, 256, 0, (cudaStream_t)stream>>>(Q, Kp, Vp, m, k, out);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_rind_rows(roreg_ctx* c, const float* rind, int m, float* out128, void* stream) {
  RR_ARG(c, rind && out128 && m >= 1);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = rr_ws_reserve(c, 4096);
  if (rc) return rc;
  float* cmax = reinterpret_cast<float*>(c->ws);
  colmax60_kernel<<<60, 256, 0, st>>>(rind, m, cmax);
  RR_LAUNCH_CHECK(c);
  rind_rows_kernel<<<(unsigned)(((long long)m * 128 + 255) / 256), 256, 0, st>>>(rind, cmax, m, out128);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

int roreg_sinkhorn_match(roreg_ctx* c, const float* S, int m, int n, int ld, float alpha, int iters, float* u, float* v,
                         int32_t* matches0, float* mscores0, void* stream) {
  RR_ARG(c, S && u && v && matches0 && mscores0 && m >= 1 && n >= 1 && ld >= n && iters >= 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int G = c->sm_count;
  // the persistent kernel reads the interior block with 16 / 8-byte loads: n, ld multiples of 4 and a 16-byte aligned S
  const bool fused = ((m + G) / G <= 64) && ((size_t)(n + 4) * sizeof(float) <= 200 * 1024) && (n % 4 == 0) && (ld % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(S) & 15) == 0) && !getenv("ROREG_SINKHORN_LAUNCHES");
  int rc = rr_ws_reserve(c, rr_align(sizeof(int32_t) * m) + rr_align(sizeof(float) * m) + rr_align(sizeof(int32_t) * n) +
                            (fused ? rr_align(sizeof(float2) * (size_t)G * (n + 1)) + rr_align(256) : 0) + 4096);
  if (rc) return rc;
  rr_arena ar{(char*)c->ws, 0};
  int32_t* idx0 = ar.take<int32_t>(m); float* max0 = ar.take<float>(m); int32_t* idx1 = ar.take<int32_t>(n);
  if (fused) {
    // one persistent cooperative kernel: 100 iterations + the assignment (kernels_matchot.cuh, sinkhorn_fused_kernel)
    float2* part = ar.take<float2>((size_t)G * (n + 1));
    unsigned int* bar = ar.take<unsigned int>(64);
    RR_CUDA(c, cudaMemsetAsync(bar, 0, 256, st));
    SinkFusedArgs fa{S, m, n, ld, alpha, -logf((float)(m + n)), iters, u, v, part, idx0, max0, idx1, matches0, mscores0, 0};
    if (const char* e = getenv("ROREG_DEBUG_SINK")) fa.dbg = atoi(e);
    const size_t smem = (size_t)(n + 4) * sizeof(float);
    static unsigned long long attr_mask = 0;
    if (rr_first_use_on_device(&attr_mask, c->device))
      RR_CUDA(c, cudaFuncSetAttribute(sinkhorn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    void* args[] = {&fa, &bar};
    RR_CUDA(c, cudaLaunchCooperativeKernel((const void*)sinkhorn_fused_kernel, dim3(G), dim3(1024), args, smem, st));
    c->launches += 1;
    return ROREG_OK;
  }
  RR_CUDA(c, cudaMemsetAsync(u, 0, sizeof(float) * (m + 1), st));
  RR_CUDA(c, cudaMemsetAsync(v, 0, sizeof(float) * (n + 1), st));
  SinkArgs a{S, m, n, ld, alpha, -logf((float)(m + n)), u, v};
  for (int it = 0; it < iters; ++it) {
    sinkhorn_rows_kernel<<<m + 1, 256, 0, st>>>(a);
    RR_LAUNCH_CHECK(c);
    sinkhorn_cols_kernel<<<(n + 1 + 31) / 32, 1024, 0, st>>>(a);
    RR_LAUNCH_CHECK(c);
  }
  ot_row_argmax_kernel<<<m, 256, 0, st>>>(a, idx0, max0);
  RR_LAUNCH_CHECK(c);
  ot_col_argmax_kernel<<<(n + 31) / 32, 256, 0, st>>>(a, idx1);
  RR_LAUNCH_CHECK(c);
  ot_mutual_kernel<<<(m + 127) / 128, 128, 0, st>>>(idx0, max0, idx1, m, matches0, mscores0);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

// ------------------------------------------------------------------------------------------------
// batched engine
// ------------------------------------------------------------------------------------------------
// ---- the batched engine in three phases per (sub-)batch -----------------------------------------------------------
// P  pooling (HBM-bound, SMs mostly idle)            TC  NN + mutual check + Des2R (tensor-core kernels, one CTA per SM)
// T  hypotheses + scoring + refinement (FP64 / latency-bound small kernels)
// Serial schedule (default): P, TC, T on the caller's stream.  Overlapped schedule (roreg_set_overlap(1)): the batch is split in
// two halves; the TC chains of both halves run back to back on a high-priority stream while P of the second half runs beside
// the NN of the first and T of the first beside the NN of the second on a low-priority stream - the kernels' bottlenecks
// are complementary and one pooling / scoring CTA fits on an SM next to the persistent NN CTA.  Measured (run 47, 64 pairs): 19.9 k
// pairs/s overlapped against 21.5 k serial - the co-resident kernels slow each other more than the overlap hides - so it is off by default.
struct BatchPlan {            // one (sub-)batch: views into the caller's arrays + its slice of the workspace
  roreg_batch b;              // pointers already offset to the first pair of the slice, b.B = pairs in the slice
  int pair_base;
  float* inv; int32_t *nn01, *nn10; double *hyps, *partial; int32_t *scratch, *n_hyp, *bucket; double* cdf;
  uint8_t* img; float* rowval; uint8_t* rowgid; int32_t* bgroup;          // nn mode 4
  float *Ahat, *Bhat, *nh; unsigned long long* cb;                        // nn modes 1-3
};

static size_t batch_ws_bytes(const roreg_batch* b, int B) {
  const int S = b->keynum, H = b->max_iter;
  const int tiles = (S + RR_SCORE_TILE - 1) / RR_SCORE_TILE;
  size_t need = rr_align(sizeof(float) * (size_t)B * 2 * S * RR_F) + 2 * rr_align(sizeof(int32_t) * (size_t)B * S) +
                rr_align(sizeof(double) * (size_t)B * H * 12) + rr_align(sizeof(double) * (size_t)B * tiles * H) +
                rr_align(sizeof(int32_t) * (size_t)B * S) + 4 * rr_align(sizeof(int32_t) * (size_t)B) +
                rr_align(sizeof(int32_t) * (size_t)B * 128) + rr_align(sizeof(double) * (size_t)B * 60) + 8192;
  if (b->nn_mode == 4) need += nn_tc4_workspace_bytes(B, S) + 2048;
  else if (b->nn_mode >= 1) need += nn_tc_workspace_bytes((long long)B * 2 * S);
  return rr_align(need);
}

static void batch_plan(const roreg_batch* b, int p0, int B, char* ws, BatchPlan* pl) {
  const int S = b->keynum, H = b->max_iter;
  const int tiles = (S + RR_SCORE_TILE - 1) / RR_SCORE_TILE;
  pl->b = *b; pl->b.B = B; pl->pair_base = p0;
  pl->b.pair_cloud = b->pair_cloud + 2 * (size_t)p0;
  if (b->sample) pl->b.sample = b->sample + (size_t)p0 * 2 * S;
  if (b->triplets) pl->b.triplets = b->triplets + (size_t)p0 * H * 3;
  if (b->hyp_host_svd) pl->b.hyp_host_svd = b->hyp_host_svd + (size_t)p0 * H * 12;
  pl->b.matches = b->matches + (size_t)p0 * S * 2; pl->b.n_matches = b->n_matches + p0; pl->b.dr_index = b->dr_index + (size_t)p0 * S;
  pl->b.poses = b->poses + (size_t)p0 * 16; pl->b.recall = b->recall + p0; pl->b.best_overlap = b->best_overlap + p0;
  rr_arena ar{ws, 0};
  pl->inv = ar.take<float>((size_t)B * 2 * S * RR_F);
  pl->nn01 = ar.take<int32_t>((size_t)B * S);
  pl->nn10 = ar.take<int32_t>((size_t)B * S);
  pl->hyps = ar.take<double>((size_t)B * H * 12);
  pl->partial = ar.take<double>((size_t)B * tiles * H);
  pl->scratch = ar.take<int32_t>((size_t)B * S);
  pl->n_hyp = ar.take<int32_t>(B);
  pl->bucket = ar.take<int32_t>((size_t)B * 128);
  pl->cdf = ar.take<double>((size_t)B * 60);
  pl->img = nullptr; pl->rowval = nullptr; pl->rowgid = nullptr; pl->bgroup = nullptr;
  pl->Ahat = pl->Bhat = pl->nh = nullptr; pl->cb = nullptr;
  if (b->nn_mode == 4) {
    const size_t NT = (size_t)(S + TC_BM - 1) / TC_BM;
    ar.off = (ar.off + 1023) & ~size_t(1023);
    pl->img = ar.take<uint8_t>((size_t)B * 2 * NT * T4_TILE_BYTES);
    pl->rowval = ar.take<float>((size_t)B * 2 * NT * NT * TC_BM);
    pl->rowgid = ar.take<uint8_t>((size_t)B * 2 * NT * NT * TC_BM);
    pl->bgroup = ar.take<int32_t>((size_t)B * S);
  } else if (b->nn_mode >= 1) {
    pl->Ahat = ar.take<float>((size_t)B * 2 * S * TC_KEXT);
    pl->Bhat = ar.take<float>((size_t)B * 2 * S * TC_KEXT);
    pl->nh = ar.take<float>((size_t)B * 2 * S);
    pl->cb = ar.take<unsigned long long>((size_t)B * S);
  }
}

#define RR_MARK(i) do { if (mark) RR_CUDA(c, cudaEventRecord(c->ev[i], st)); } while (0)

// P: invariant pooling + normalisation of both sides of every pair  (test/matcher.py:69-72)
static int batch_phase_pool(roreg_ctx* c, const BatchPlan& pl, cudaStream_t st, bool mark) {
  const roreg_batch* b = &pl.b;
  const int B = b->B, S = b->keynum;
  RR_MARK(0);
  PoolArgs pa{b->desc, b->pair_cloud, b->sample, b->n, S, B * 2 * S, 1, pl.inv};
  if (b->nn_mode == 4) {
    const int NT = (S + TC_BM - 1) / TC_BM;
    inv_pool_t4_kernel<<<(pa.rows + 7) / 8, 256, 0, st>>>(pa, NT, pl.img);      // pooling fused with the fp16 operand image
  } else {
    inv_pool_kernel<<<(pa.rows + 7) / 8, 256, 0, st>>>(pa);
  }
  RR_LAUNCH_CHECK(c);
  RR_MARK(1);
  return ROREG_OK;
}

// TC: 1-NN both ways (test/matcher.py:94-97), mutual check (:98-107), coarse rotation of every match (test/estimator.py:105-111)
static int batch_phase_tc(roreg_ctx* c, const BatchPlan& pl, cudaStream_t st, bool mark) {
  const roreg_batch* b = &pl.b;
  const int B = b->B, S = b->keynum;
  int rc;
  const long long ps = 2LL * S * RR_F;
  if (b->nn_mode == 4) {
    if ((rc = nn_tc4_launch_both(c, pl.inv, S, B, pl.img, pl.rowval, pl.rowgid, pl.bgroup, pl.nn01, pl.nn10, st, true))) return rc;
  } else if (b->nn_mode == 3) {
    if ((rc = nn_tc3_launch_both(c, pl.inv, S, B, pl.Ahat, pl.nh, pl.cb, pl.nn01, pl.nn10, st))) return rc;
  } else if (b->nn_mode == 2) {
    if ((rc = nn_tc2_launch_both(c, pl.inv, S, B, pl.Ahat, pl.nh, pl.nn01, pl.nn10, st))) return rc;
  } else if (b->nn_mode == 1) {
    if ((rc = nn_tc_launch_both(c, pl.inv, S, B, pl.Ahat, pl.Bhat, pl.nh, pl.nn01, pl.nn10, st))) return rc;
  } else {
    if ((rc = launch_nn(c, 0, pl.inv, pl.inv + (size_t)S * RR_F, ps, ps, S, S, pl.nn01, nullptr, S, B, st))) return rc;
    if ((rc = launch_nn(c, 0, pl.inv + (size_t)S * RR_F, pl.inv, ps, ps, S, S, pl.nn10, nullptr, S, B, st))) return rc;
  }
  RR_MARK(2);
  CompactArgs ca{pl.nn01, pl.nn10, S, S, S, b->sample, S, b->matches, S, b->n_matches};
  mutual_compact_kernel<<<B, 1024, 0, st>>>(ca);
  RR_LAUNCH_CHECK(c);
  RR_MARK(3);
  // X = cloud id1, Y = cloud id0
  if (c->corr_mode >= 1) {
    CorrTcArgs t{nullptr, nullptr, b->matches + 1, b->matches, 2, b->pair_cloud, b->n, b->n_matches, S, B, c->d_perm8, nullptr, b->dr_index, 3, 0, nullptr};
    if (const char* e = getenv("ROREG_DEBUG_CORR_PASSES")) { const int v = atoi(e); if (v >= 1 && v <= 3) t.dbg_passes = v; }
    if (const char* e = getenv("ROREG_DEBUG_CORR_SKIP")) t.dbg_skip = atoi(e);
    if ((rc = (c->corr_mode == 3 ? group_corr_tc3_launch(c, b->desc, b->desc, t, st) : c->corr_mode == 2 ? group_corr_tc2_launch(c, b->desc, b->desc, t, st) : group_corr_tc_launch(c, b->desc, b->desc, t, st)))) return rc;
  } else {
    CorrArgs co{};
    co.X = b->desc; co.Y = b->desc; co.idxX = b->matches + 1; co.idxY = b->matches; co.idx_stride = 2;
    co.pair_cloud = b->pair_cloud; co.n = b->n; co.n_matches = b->n_matches; co.K = S; co.B = B;
    co.tab = c->d_perm8; co.cor_out = nullptr; co.argmax_out = b->dr_index;
    const long long total = (long long)B * S;
    const int grid = (int)(total < (long long)c->sm_count * 8 ? total : (long long)c->sm_count * 8);
    group_corr_kernel<<<grid, 128, 0, st>>>(co);
    RR_LAUNCH_CHECK(c);
  }
  RR_MARK(4);
  return ROREG_OK;
}

// T: hypotheses, one-shot scoring (test/estimator.py:149-154,232-238 / :426-436), two refinements (:240-241 / :438-439)
static int batch_phase_tail(roreg_ctx* c, const BatchPlan& pl, cudaStream_t st, bool mark) {
  const roreg_batch* b = &pl.b;
  const int B = b->B, S = b->keynum, H = b->max_iter;
  int rc;
  MatchView mv{};
  mv.keys0 = b->keys; mv.keys1 = b->keys; mv.pair_cloud = b->pair_cloud; mv.n = b->n; mv.matches = b->matches;
  mv.cap = S; mv.n_matches = b->n_matches; mv.K = S; mv.scores = nullptr; mv.scores_f64 = 0; mv.scores_pair_stride = 0;
  const double* hyp_src = pl.hyps;
  const int32_t* n_hyp_src = pl.n_hyp;
  if (b->hyp_host_svd) { hyp_src = b->hyp_host_svd; n_hyp_src = nullptr; }
  else {
    CoarseArgs cg{mv, b->dr_index, S, b->triplets, H, b->seed, pl.pair_base, pl.hyps, pl.n_hyp, pl.scratch, pl.bucket, pl.cdf};
    if (!b->triplets) {
      coarse_bucket_kernel<<<B, 256, 0, st>>>(cg);
      RR_LAUNCH_CHECK(c);
    } else {
      RR_CUDA(c, cudaMemsetAsync(pl.n_hyp, 0x7f, sizeof(int32_t) * B, st));     // "all H hypotheses valid"
    }
    coarse_hyp_kernel<<<dim3((H + 63) / 64, B), 64, 0, st>>>(cg);
    RR_LAUNCH_CHECK(c);
  }
  RR_MARK(5);
  if ((rc = score_and_select(c, mv, S, hyp_src, (long long)H * 12, nullptr, n_hyp_src, H, b->ird, pl.partial, nullptr,
                             b->recall, b->best_overlap, B, st))) return rc;
  RR_MARK(6);
  RefineArgs ra{};
  ra.mv = mv; ra.T_in = nullptr; ra.hyps = hyp_src; ra.hyp_pair_stride = (long long)H * 12; ra.order = nullptr;
  ra.best_id = b->recall; ra.rad0 = b->ird * 2.0; ra.rad1 = b->ird; ra.rounds = 2; ra.T_out = b->poses; ra.inlier_mask = nullptr; ra.mask_stride = S;
  refine_kernel<<<B, 512, 0, st>>>(ra);
  RR_LAUNCH_CHECK(c);
  RR_MARK(7);
  return ROREG_OK;
}
#undef RR_MARK

int roreg_set_overlap(roreg_ctx* c, int enable) {
  if (!c) return ROREG_ERR_ARG;
  c->overlap = enable ? 1 : 0;
  return ROREG_OK;
}

static int overlap_streams(roreg_ctx* c) {
  if (c->s_tc) return ROREG_OK;
  int lo = 0, hi = 0;
  RR_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));          // lo = numerically largest = lowest priority
  RR_CUDA(c, cudaStreamCreateWithPriority(&c->s_tc, cudaStreamNonBlocking, hi));
  RR_CUDA(c, cudaStreamCreateWithPriority(&c->s_light, cudaStreamNonBlocking, lo));
  RR_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  for (int h = 0; h < 2; ++h) {
    RR_CUDA(c, cudaEventCreateWithFlags(&c->ev_pool[h], cudaEventDisableTiming));
    RR_CUDA(c, cudaEventCreateWithFlags(&c->ev_corr[h], cudaEventDisableTiming));
    RR_CUDA(c, cudaEventCreateWithFlags(&c->ev_join[h], cudaEventDisableTiming));
  }
  return ROREG_OK;
}

int roreg_register_batch(roreg_ctx* c, const roreg_batch* b, void* stream) {
  RR_ARG(c, b && b->desc && b->keys && b->pair_cloud && b->matches && b->n_matches && b->dr_index && b->poses &&
                b->recall && b->best_overlap);
  RR_ARG(c, b->B >= 1 && b->n >= 1 && b->keynum >= 1 && b->keynum <= b->n && b->max_iter >= 1 && b->ird > 0);
  RR_ARG(c, b->sample || b->keynum == b->n);
  RR_ARG(c, b->estimator == 0 || (b->estimator == 1 && b->hyp_host_svd) || b->estimator == 2);
  RR_ARG(c, b->nn_mode >= 0 && b->nn_mode <= 4);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // the two-stream schedule needs two sub-batches worth the launch overhead, no per-stage timing and the full pipeline
  const bool overlapped = c->overlap && !c->timing && b->estimator != 2 && b->B >= 8 && !getenv("ROREG_NO_OVERLAP");
  if (!overlapped) {
    if ((rc = rr_ws_reserve(c, batch_ws_bytes(b, b->B)))) return rc;
    BatchPlan pl;
    batch_plan(b, 0, b->B, (char*)c->ws, &pl);
    const bool mark = c->timing != 0;
    if ((rc = batch_phase_pool(c, pl, st, mark))) return rc;
    if ((rc = batch_phase_tc(c, pl, st, mark))) return rc;
    if (b->estimator == 2) {                 // matcher + Des2R only: the caller generates hypotheses (ET network) and calls roreg_estimate_batch
      if (mark) { for (int i = 5; i <= ROREG_N_STAGES; ++i) RR_CUDA(c, cudaEventRecord(c->ev[i], st)); c->ev_valid = 1; }
      return ROREG_OK;
    }
    if ((rc = batch_phase_tail(c, pl, st, mark))) return rc;
    if (mark) c->ev_valid = 1;
    return ROREG_OK;
  }
  const int B0 = (b->B + 1) / 2, B1 = b->B - B0;
  const size_t w0 = batch_ws_bytes(b, B0), w1 = batch_ws_bytes(b, B1);
  if ((rc = rr_ws_reserve(c, w0 + w1 + 4096))) return rc;
  if ((rc = overlap_streams(c))) return rc;
  BatchPlan pl[2];
  batch_plan(b, 0, B0, (char*)c->ws, &pl[0]);
  batch_plan(b, B0, B1, (char*)c->ws + w0, &pl[1]);
  // fork: both internal streams start after everything already enqueued on the caller's stream
  RR_CUDA(c, cudaEventRecord(c->ev_fork, st));
  RR_CUDA(c, cudaStreamWaitEvent(c->s_light, c->ev_fork, 0));
  RR_CUDA(c, cudaStreamWaitEvent(c->s_tc, c->ev_fork, 0));
  // s_light: P0 P1 ... ; s_tc: (P0) TC0 (P1) TC1 ; s_light: ... (TC0) T0 (TC1) T1
  for (int h = 0; h < 2; ++h) {
    if ((rc = batch_phase_pool(c, pl[h], c->s_light, false))) return rc;
    RR_CUDA(c, cudaEventRecord(c->ev_pool[h], c->s_light));
  }
  for (int h = 0; h < 2; ++h) {
    RR_CUDA(c, cudaStreamWaitEvent(c->s_tc, c->ev_pool[h], 0));
    if ((rc = batch_phase_tc(c, pl[h], c->s_tc, false))) return rc;
    RR_CUDA(c, cudaEventRecord(c->ev_corr[h], c->s_tc));
  }
  for (int h = 0; h < 2; ++h) {
    RR_CUDA(c, cudaStreamWaitEvent(c->s_light, c->ev_corr[h], 0));
    if ((rc = batch_phase_tail(c, pl[h], c->s_light, false))) return rc;
  }
  // join: later work on the caller's stream sees every result
  RR_CUDA(c, cudaEventRecord(c->ev_join[0], c->s_tc));
  RR_CUDA(c, cudaEventRecord(c->ev_join[1], c->s_light));
  RR_CUDA(c, cudaStreamWaitEvent(st, c->ev_join[0], 0));
  RR_CUDA(c, cudaStreamWaitEvent(st, c->ev_join[1], 0));
  return ROREG_OK;
}

// ------------------------------------------------------------------------------------------------
// Pipelined form for back-to-back batches (throughput runs): the RANSAC tail T of batch i-1 (FP64 / latency-bound kernels with
// small footprints) runs on an internal high-priority stream BESIDE the pooling P of batch i (HBM-bound, SMs idle) - the only two
// phases whose CTAs fit on one SM together; the tensor-core phase TC owns the SMs and runs alone.  Per call, on the caller's stream:
//     [ T(i-1) on the internal stream  ||  P(i) ]  ->  TC(i)  ->  wait for T(i-1)
// Every phase processes a WHOLE batch (no half-batches: run 47 showed their fixed costs eat the overlap).  Two workspace slots
// alternate, so batch i+1 reuses the slot of batch i-1 only after its tail was joined.
struct PipeState {
  BatchPlan pl;                 // the batch whose P and TC are enqueued and whose T is still owed
  bool pending;
  int slot;                     // workspace slot of the NEXT batch
  void* ws[2]; size_t ws_bytes[2];
  cudaStream_t s_tail; cudaEvent_t ev_tc, ev_tail;
};

static void pipe_destroy(roreg_ctx* c) {
  PipeState* ps = reinterpret_cast<PipeState*>(c->pipe);
  if (!ps) return;
  cudaDeviceSynchronize();
  for (int i = 0; i < 2; ++i) if (ps->ws[i]) cudaFree(ps->ws[i]);
  if (ps->s_tail) cudaStreamDestroy(ps->s_tail);
  if (ps->ev_tc) cudaEventDestroy(ps->ev_tc);
  if (ps->ev_tail) cudaEventDestroy(ps->ev_tail);
  delete ps;
  c->pipe = nullptr;
}

static int pipe_get(roreg_ctx* c, PipeState** out) {
  if (!c->pipe) {
    PipeState* ps = new PipeState();
    ps->pending = false; ps->slot = 0; ps->ws[0] = ps->ws[1] = nullptr; ps->ws_bytes[0] = ps->ws_bytes[1] = 0;
    ps->s_tail = nullptr; ps->ev_tc = nullptr; ps->ev_tail = nullptr;
    c->pipe = ps;
    int lo = 0, hi = 0;
    RR_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // which of the two co-running phases gets freed CTA slots first is a tuning knob (ROREG_PIPE_TAIL_PRIO = hi | lo | same):
    // hi (default) lets the small tail kernels in as pooling CTAs retire; lo runs them in whatever the pooling grid leaves
    int prio = hi;
    if (const char* e = getenv("ROREG_PIPE_TAIL_PRIO")) prio = !strcmp(e, "lo") ? lo : !strcmp(e, "same") ? 0 : hi;
    RR_CUDA(c, cudaStreamCreateWithPriority(&ps->s_tail, cudaStreamNonBlocking, prio));
    RR_CUDA(c, cudaEventCreateWithFlags(&ps->ev_tc, cudaEventDisableTiming));
    RR_CUDA(c, cudaEventCreateWithFlags(&ps->ev_tail, cudaEventDisableTiming));
  }
  *out = reinterpret_cast<PipeState*>(c->pipe);
  return ROREG_OK;
}

// enqueue the owed tail on the internal stream, ordered after everything on `st` so far; the caller joins with ev_tail
static int pipe_tail(roreg_ctx* c, PipeState* ps, cudaStream_t st) {
  int rc;
  RR_CUDA(c, cudaEventRecord(ps->ev_tc, st));
  RR_CUDA(c, cudaStreamWaitEvent(ps->s_tail, ps->ev_tc, 0));
  if ((rc = batch_phase_tail(c, ps->pl, ps->s_tail, false))) return rc;
  RR_CUDA(c, cudaEventRecord(ps->ev_tail, ps->s_tail));
  return ROREG_OK;
}

int roreg_register_batch_flush(roreg_ctx* c, void* stream) {
  if (!c) return ROREG_ERR_ARG;
  PipeState* ps = reinterpret_cast<PipeState*>(c->pipe);
  if (!ps || !ps->pending) return ROREG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = pipe_tail(c, ps, st))) return rc;
  RR_CUDA(c, cudaStreamWaitEvent(st, ps->ev_tail, 0));
  ps->pending = false;
  return ROREG_OK;
}

int roreg_register_batch_pipelined(roreg_ctx* c, const roreg_batch* b, void* stream) {
  RR_ARG(c, b && b->desc && b->keys && b->pair_cloud && b->matches && b->n_matches && b->dr_index && b->poses &&
                b->recall && b->best_overlap);
  RR_ARG(c, b->B >= 1 && b->n >= 1 && b->keynum >= 1 && b->keynum <= b->n && b->max_iter >= 1 && b->ird > 0);
  RR_ARG(c, b->sample || b->keynum == b->n);
  RR_ARG(c, b->estimator == 0 || (b->estimator == 1 && b->hyp_host_svd));
  RR_ARG(c, b->nn_mode >= 0 && b->nn_mode <= 4);
  cudaStream_t st = (cudaStream_t)stream;
  PipeState* ps;
  int rc;
  if ((rc = pipe_get(c, &ps))) return rc;
  const int slot = ps->slot;
  const size_t need = batch_ws_bytes(b, b->B);
  if (ps->ws_bytes[slot] < need) {            // growth only: the slot's last user was joined into `st` one call ago, drain before freeing
    RR_CUDA(c, cudaDeviceSynchronize());
    if (ps->ws[slot]) { cudaFree(ps->ws[slot]); ps->ws[slot] = nullptr; ps->ws_bytes[slot] = 0; }
    const size_t want = need + need / 8 + (1 << 20);
    if (cudaMalloc(&ps->ws[slot], want) != cudaSuccess) {
      cudaGetLastError();
      snprintf(c->err, sizeof(c->err), "pipelined workspace cudaMalloc(%zu) failed", want);
      return ROREG_ERR_NOMEM;
    }
    ps->ws_bytes[slot] = want;
  }
  const bool owed = ps->pending;
  if (owed && (rc = pipe_tail(c, ps, st))) return rc;            // T(i-1) on the internal stream ...
  BatchPlan pl;
  batch_plan(b, 0, b->B, (char*)ps->ws[slot], &pl);
  if ((rc = batch_phase_pool(c, pl, st, false))) return rc;      // ... beside P(i)
  if ((rc = batch_phase_tc(c, pl, st, false))) return rc;
  if (owed) RR_CUDA(c, cudaStreamWaitEvent(st, ps->ev_tail, 0)); // later work on `st` (and the slot's next user) sees T(i-1) done
  ps->pl = pl; ps->pending = true; ps->slot = slot ^ 1;
  return ROREG_OK;
}

// steps 6-7 of the batched engine on caller-provided hypotheses (yohoo: ET network + test/estimator.py:349-366):
// hyps [B][H][3][4] float64, n_hyp [B] valid hypotheses per pair (NULL = H); uses batch->matches / n_matches from a
// preceding roreg_register_batch(estimator = 2) and writes poses / recall / best_overlap.
int roreg_estimate_batch(roreg_ctx* c, const roreg_batch* b, const double* hyps, const int32_t* n_hyp, void* stream) {
  RR_ARG(c, b && hyps && b->keys && b->pair_cloud && b->matches && b->n_matches && b->poses && b->recall && b->best_overlap);
  RR_ARG(c, b->B >= 1 && b->keynum >= 1 && b->max_iter >= 1 && b->ird > 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = b->B, S = b->keynum, H = b->max_iter;
  const int tiles = (S + RR_SCORE_TILE - 1) / RR_SCORE_TILE;
  int rc = rr_ws_reserve(c, rr_align(sizeof(double) * (size_t)B * tiles * H) + 4096);
  if (rc) return rc;
  rr_arena ar{(char*)c->ws, 0};
  double* partial = ar.take<double>((size_t)B * tiles * H);
  MatchView mv{};
  mv.keys0 = b->keys; mv.keys1 = b->keys; mv.pair_cloud = b->pair_cloud; mv.n = b->n; mv.matches = b->matches;
  mv.cap = S; mv.n_matches = b->n_matches; mv.K = S; mv.scores = nullptr; mv.scores_f64 = 0; mv.scores_pair_stride = 0;
  if ((rc = score_and_select(c, mv, S, hyps, (long long)H * 12, nullptr, n_hyp, H, b->ird, partial, nullptr, b->recall,
                             b->best_overlap, B, st))) return rc;
  RefineArgs ra{};
  ra.mv = mv; ra.T_in = nullptr; ra.hyps = hyps; ra.hyp_pair_stride = (long long)H * 12; ra.order = nullptr;
  ra.best_id = b->recall; ra.rad0 = b->ird * 2.0; ra.rad1 = b->ird; ra.rounds = 2; ra.T_out = b->poses; ra.inlier_mask = nullptr; ra.mask_stride = S;
  refine_kernel<<<B, 512, 0, st>>>(ra);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

}  // extern "C"
