// kernels_ransac.cuh - hypothesis generation, one-shot RANSAC scoring, first-best selection and the
// weighted-Kabsch refiner.  Reference arithmetic: test/estimator.py:28-72 (refiner), :139-154
// (Threepps2Tran, overlap_cal), :349-366 (hypotheses from quaternions), :377-382,:426-439 (one-shot
// RANSAC + refine).  All geometry is float64, as in the reference (keypoints and Trans are float64).
#pragma once
#include "common.cuh"
#include "math3.cuh"

namespace roreg {

// Matched keypoints of pair p.  Batched: keys arena + match rows [B][cap][2] + device counts.
// Single call: already-gathered k0/k1 [K][3] (matches == NULL), K from the host.
struct MatchView {
  const double* keys0; const double* keys1;    // single call: [K][3] each.  batched: keys arena [n_clouds][n][3] (both equal)
  const int32_t* pair_cloud; int n;            // batched only
  const int32_t* matches; int cap;             // batched only
  const int32_t* n_matches; int K;             // device counts (batched) or host K
  const void* scores; int scores_f64;          // NULL -> 1.0 (mutual matcher, test/matcher.py:109)
  long long scores_pair_stride;
};

__device__ __forceinline__ int mv_count(const MatchView& v, int p) { return v.n_matches ? v.n_matches[p] : v.K; }
__device__ __forceinline__ void mv_load(const MatchView& v, int p, int k, double a[3], double b[3], double& s) {
  const double* p0; const double* p1;
  if (v.matches) {
    const int32_t* m = v.matches + ((long long)p * v.cap + k) * 2;
    p0 = v.keys0 + ((long long)v.pair_cloud[2 * p] * v.n + m[0]) * 3;
    p1 = v.keys1 + ((long long)v.pair_cloud[2 * p + 1] * v.n + m[1]) * 3;
  } else {
    p0 = v.keys0 + (long long)k * 3; p1 = v.keys1 + (long long)k * 3;
  }
  a[0] = p0[0]; a[1] = p0[1]; a[2] = p0[2];
  b[0] = p1[0]; b[1] = p1[1]; b[2] = p1[2];
  if (!v.scores) s = 1.0;
  else if (v.scores_f64) s = reinterpret_cast<const double*>(v.scores)[p * v.scores_pair_stride + k];
  else s = (double)reinterpret_cast<const float*>(v.scores)[p * v.scores_pair_stride + k];
}

__device__ __forceinline__ bool is_inlier(const double T[12], const double a[3], const double b[3], double r2) {
  // transform_points (utils/utils.py:42): pts @ R^T + t ; diff = k0 - k1' ; sum of squares < r^2
  const double x = T[0] * b[0] + T[1] * b[1] + T[2] * b[2] + T[3];
  const double y = T[4] * b[0] + T[5] * b[1] + T[6] * b[2] + T[7];
  const double z = T[8] * b[0] + T[9] * b[1] + T[10] * b[2] + T[11];
  const double dx = a[0] - x, dy = a[1] - y, dz = a[2] - z;
  return (dx * dx + dy * dy + dz * dz) < r2;
}

// ---------------------------------------------------------------------------------------------
// hypotheses from (quaternion, coarse rotation id): one thread per match.
// ---------------------------------------------------------------------------------------------
__global__ void hyp_from_quat_kernel(const float* __restrict__ quat, const int32_t* __restrict__ pre_idx,
                                     const double* __restrict__ k0, const double* __restrict__ k1, int K,
                                     const float* __restrict__ rot32, double* __restrict__ trans) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  float q[4] = {quat[4 * i], quat[4 * i + 1], quat[4 * i + 2], quat[4 * i + 3]};
  float Rg[9];
  const int a = pre_idx[i];
#pragma unroll
  for (int j = 0; j < 9; ++j) Rg[j] = rot32[a * 9 + j];
  double R[9];
  quat_times_anchor(q, Rg, R);
  double* T = trans + (long long)i * 12;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    // t = key0 - key1 @ R.T   (test/estimator.py:364)
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(k1[3 * i], R[3 * r]), __dmul_rn(k1[3 * i + 1], R[3 * r + 1])),
                                 __dmul_rn(k1[3 * i + 2], R[3 * r + 2]));
    T[4 * r] = R[3 * r]; T[4 * r + 1] = R[3 * r + 1]; T[4 * r + 2] = R[3 * r + 2];
    T[4 * r + 3] = k0[3 * i + r] - dot;
  }
}

// ---------------------------------------------------------------------------------------------
// yohoc hypothesis generation (test/estimator.py:119-147,221-232).  One CTA per pair.
//  * histogram of DR_index over the pair's matches, p_r ~ n(n-0.01)(n-0.02), n = count/100 (0 if count<2)
//  * per iteration: rotation id r ~ p, then 3 matches drawn with replacement from bucket r
//    (device counter-based RNG, or host-drawn triplets when `triplets` != NULL - parity mode)
//  * 3-point Kabsch (proper-rotation branch, math3.cuh) -> hyps[p][it][3][4]
// If every bucket has < 2 members (reference: degenerate, :215-218) n_hyp[p] = 0.
// ---------------------------------------------------------------------------------------------
struct CoarseArgs {
  MatchView mv;
  const int32_t* dr_index; long long dr_pair_stride;
  const int32_t* triplets;       // [B][H][3] or NULL
  int H; uint64_t seed;
  int pair_base;                 // index of this launch's first pair in the caller's batch (keeps the draws independent of how a batch is split)
  double* hyps;                  // [B][H][12]
  int32_t* n_hyp;                // [B]
  int32_t* scratch;              // [B][cap] bucket-sorted match ids
  int32_t* bucket;               // [B][128]: cnt[60] | start[61]  (written by coarse_bucket_kernel)
  double* cdf;                   // [B][60]
};

// stage 1: per-pair histogram / cdf / bucket sort (one CTA per pair)
__global__ void __launch_bounds__(256) coarse_bucket_kernel(CoarseArgs a) {
  __shared__ int cnt[60], start[61], fill[60];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int K = mv_count(a.mv, p);
  const int32_t* dr = a.dr_index + p * a.dr_pair_stride;
  int32_t* sorted = a.scratch + (long long)p * a.mv.cap;
  if (tid < 60) { cnt[tid] = 0; fill[tid] = 0; }
  __syncthreads();
  for (int k = tid; k < K; k += 256) atomicAdd(&cnt[dr[k]], 1);
  __syncthreads();
  if (tid == 0) {
    double tot = 0; int s = 0;
    double c[60];
    for (int r = 0; r < 60; ++r) {
      start[r] = s; s += cnt[r];
      double pr = 0;
      if (cnt[r] >= 2) { const double num = (double)cnt[r] / 100.0; pr = num * (num - 0.01) * (num - 0.02); }
      tot += pr; c[r] = tot;
    }
    start[60] = s;
    for (int r = 0; r < 60; ++r) a.cdf[p * 60 + r] = (tot > 0) ? c[r] / tot : 0.0;
    a.n_hyp[p] = (tot > 0) ? a.H : 0;
  }
  __syncthreads();
  if (tid < 60) a.bucket[p * 128 + tid] = cnt[tid];
  if (tid < 61) a.bucket[p * 128 + 60 + tid] = start[tid];
  // stable counting sort by one warp, 32 matches per step: a match's slot = start[bucket] + matches of the same bucket seen before
  // (fill) + its rank among the same-bucket lanes of this step (match_any).  Deterministic - ascending k inside a bucket - so the
  // device-side draws pick the same triplets on every run (an atomicAdd scatter made them depend on the arrival order).
  if (tid < 32) {
    for (int base = 0; base < K; base += 32) {
      const int k = base + tid;
      const bool valid = k < K;
      const int r = valid ? dr[k] : 64 + tid;                          // invalid lanes get unique keys
      const unsigned m = __match_any_sync(0xffffffffu, r);
      const int rank = __popc(m & ((1u << tid) - 1u));
      const int before = valid ? fill[r] : 0;
      __syncwarp();
      if (valid) {
        sorted[start[r] + before + rank] = k;
        if ((m >> tid) == 1u) fill[r] = before + __popc(m);           // the highest lane of the group publishes the new count
      }
      __syncwarp();
    }
  }
}

// stage 2: one thread per hypothesis, grid = (ceil(H/64), B)
__global__ void __launch_bounds__(64) coarse_hyp_kernel(CoarseArgs a) {
  __shared__ int cnt[60], start[61];
  __shared__ double cdf[60];
  const int p = blockIdx.y, tid = threadIdx.x;
  const int it = blockIdx.x * 64 + tid;
  const int32_t* sorted = a.scratch + (long long)p * a.mv.cap;
  if (!a.triplets) {
    if (a.n_hyp[p] == 0) return;                       // degenerate pair: no bucket with >= 2 members
    if (tid < 60) { cnt[tid] = a.bucket[p * 128 + tid]; cdf[tid] = a.cdf[p * 60 + tid]; }
    if (tid < 61) start[tid] = a.bucket[p * 128 + 60 + tid];
    __syncthreads();
  }
  if (it >= a.H) return;
  int idx[3];
  if (a.triplets) {
    const int32_t* t = a.triplets + ((long long)p * a.H + it) * 3;
    idx[0] = t[0]; idx[1] = t[1]; idx[2] = t[2];
  } else {
    const double u = rr_u01(a.seed, p + a.pair_base, it, 0);
    int r = 0;
    while (r < 59 && !(u < cdf[r])) ++r;
    while (r > 0 && cnt[r] < 2) --r;                // guard against cdf round-off landing on an empty bucket
    if (cnt[r] < 2) { for (r = 0; r < 59 && cnt[r] < 2; ++r) {} }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      int o = (int)(rr_u01(a.seed, p + a.pair_base, it, 1 + j) * cnt[r]);
      o = min(o, cnt[r] - 1);
      idx[j] = sorted[start[r] + o];
    }
  }
  double k0[9], k1[9], s;
#pragma unroll
  for (int j = 0; j < 3; ++j) mv_load(a.mv, p, idx[j], &k0[3 * j], &k1[3 * j], s);
  three_point_transform(k0, k1, a.hyps + ((long long)p * a.H + it) * 12);
}

// kabsch3: the 3-point solver alone on gathered (selected) matches - single-call C-ABI entry.
__global__ void kabsch3_kernel(const double* __restrict__ k0s, const double* __restrict__ k1s,
                               const int32_t* __restrict__ trip, int H, double* __restrict__ trans) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  double a[9], b[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int m = trip[3 * h + j];
#pragma unroll
    for (int c = 0; c < 3; ++c) { a[3 * j + c] = k0s[3 * m + c]; b[3 * j + c] = k1s[3 * m + c]; }
  }
  three_point_transform(a, b, trans + (long long)h * 12);
}

// ---------------------------------------------------------------------------------------------
// one-shot scoring: thread = hypothesis, CTA tile = 256 matches staged in shared memory (broadcast
// reads), grid = (hyp chunks, match tiles, pairs).  partial[p][tile][h] = sum of inlier scores.
// ---------------------------------------------------------------------------------------------
#define RR_SCORE_TILE 256
struct ScoreArgs {
  MatchView mv;
  const double* hyps; long long hyp_pair_stride;   // [B][*][12]
  const int32_t* order;                            // [H] indices into hyps or NULL (single call)
  const int32_t* n_hyp;                            // [B] device counts or NULL -> H
  int H; double r2;
  double* partial;                                 // [B][tiles][H]
  int tiles;
  double r;                                        // sqrt(r2), score mode 1 only
};

__global__ void __launch_bounds__(256) ransac_score_kernel(ScoreArgs a) {
  __shared__ double sk[RR_SCORE_TILE][7];
  const int p = blockIdx.z, tile = blockIdx.y, tid = threadIdx.x;
  const int K = mv_count(a.mv, p);
  const int k_begin = tile * RR_SCORE_TILE;
  const int h = blockIdx.x * 256 + tid;
  const int H = a.n_hyp ? min(a.H, a.n_hyp[p]) : a.H;
  if (k_begin >= K) {                       // empty tile: still publish zeros so that the reduction can read it
    if (h < a.H) a.partial[((long long)p * a.tiles + tile) * a.H + h] = 0.0;
    return;
  }
  const int cnt = min(RR_SCORE_TILE, K - k_begin);
  if (tid < cnt) {
    double x[3], y[3], s;
    mv_load(a.mv, p, k_begin + tid, x, y, s);
    sk[tid][0] = x[0]; sk[tid][1] = x[1]; sk[tid][2] = x[2];
    sk[tid][3] = y[0]; sk[tid][4] = y[1]; sk[tid][5] = y[2]; sk[tid][6] = s;
  }
  __syncthreads();
  if (h >= a.H) return;
  double acc = 0.0;
  if (h < H) {
    const long long src = a.order ? a.order[h] : h;
    const double* Tp = a.hyps + p * a.hyp_pair_stride + src * 12;
    double T[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) T[j] = Tp[j];
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const bool in = is_inlier(T, &sk[k][0], &sk[k][3], a.r2);
      acc += in ? sk[k][6] : 0.0;
    }
  }
  a.partial[((long long)p * a.tiles + tile) * a.H + h] = acc;
}

// ---------------------------------------------------------------------------------------------
// score mode 1 (roreg_set_score_mode, opt-in): float32 pre-filter with an exact float64 re-check.  Same grid, same partial sums,
// same decisions as ransac_score_kernel: every point test is first evaluated in float32 (15 FP32-pipe operations instead of 20
// FP64-pipe ones) and compared against [r2 - m, r2 + m]; only a test that lands inside the band is repeated in float64 with
// is_inlier().  m bounds |s32 - d2| for every d2 on the wrong side of r2 (u = 2^-24, T = [R|t], a = k0_i, b = k1_i):
//   x32_c = fma(R_c0,b_0, fma(R_c1,b_1, fma(R_c2,b_2,t_c)))  =>  |x32_c - x_c| <= 6u (rowsum|R| * max|b| + max|t|)   (3 roundings + 3 input
//   d32_c = a_c - x32_c                                       =>  |d32_c - d_c| <= E + u|d_c|,  E = u (max|a| + 6 (...))    conversions)
//   s32   = fma(d_z,d_z, fma(d_y,d_y, d_x*d_x))               =>  |s32 - d2|    <= 2 sqrt(3) E d + 4 E^2 + 6 u d2
// The right-hand side grows more slowly than d2 itself, so evaluated at d = r it covers both directions (a true inlier cannot
// read above r2 + m, a true outlier cannot read below r2 - m); the kernel uses 2 m.  NaN / inf hypotheses fail both float
// comparisons and take the float64 path.  max|a|, max|b| are per tile (block reduction while staging), the T terms per thread.
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ bool is_inlier_recheck(const double* __restrict__ Tp, const double* pt, double r2) {
  double T[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) T[j] = Tp[j];
  return is_inlier(T, pt, pt + 3, r2);
}

// Work items = (hypothesis block, match tile, pair), hypothesis block fastest; the grid may be smaller than the item count (a cap on
// resident CTAs leaves room for a co-running kernel, roreg_register_batch_pipelined): every CTA strides over the items.
__global__ void __launch_bounds__(256) ransac_score_pre_kernel(ScoreArgs a, int hx_blocks, long long items) {
  __shared__ double sk[RR_SCORE_TILE][7];
  __shared__ float4 sa[RR_SCORE_TILE], sb[RR_SCORE_TILE];
  __shared__ uint32_t smax[2][8];
  const int tid = threadIdx.x;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
  const int hx = (int)(item % hx_blocks), tile = (int)((item / hx_blocks) % a.tiles), p = (int)(item / ((long long)hx_blocks * a.tiles));
  const int K = mv_count(a.mv, p);
  const int k_begin = tile * RR_SCORE_TILE;
  const int h = hx * 256 + tid;
  const int H = a.n_hyp ? min(a.H, a.n_hyp[p]) : a.H;
  if (k_begin >= K) {                       // CTA-uniform
    if (h < a.H) a.partial[((long long)p * a.tiles + tile) * a.H + h] = 0.0;
    continue;
  }
  const int cnt = min(RR_SCORE_TILE, K - k_begin);
  float am = 0.f, bm = 0.f;
  if (tid < cnt) {
    double x[3], y[3], s;
    mv_load(a.mv, p, k_begin + tid, x, y, s);
    sk[tid][0] = x[0]; sk[tid][1] = x[1]; sk[tid][2] = x[2];
    sk[tid][3] = y[0]; sk[tid][4] = y[1]; sk[tid][5] = y[2]; sk[tid][6] = s;
    sa[tid] = make_float4((float)x[0], (float)x[1], (float)x[2], 0.f);
    sb[tid] = make_float4((float)y[0], (float)y[1], (float)y[2], 0.f);
    am = f32_at_or_above(fmax(fabs(x[0]), fmax(fabs(x[1]), fabs(x[2]))));
    bm = f32_at_or_above(fmax(fabs(y[0]), fmax(fabs(y[1]), fabs(y[2]))));
    if (!(am == am)) am = __int_as_float(0x7f800000);          // NaN coordinates: widen the band to everything
    if (!(bm == bm)) bm = __int_as_float(0x7f800000);
  }
  // non-negative floats order like their bit patterns
  const uint32_t wa = __reduce_max_sync(0xffffffffu, __float_as_uint(am)), wb = __reduce_max_sync(0xffffffffu, __float_as_uint(bm));
  if ((tid & 31) == 0) { smax[0][tid >> 5] = wa; smax[1][tid >> 5] = wb; }
  __syncthreads();
  double acc = 0.0;
  if (h < H) {
    uint32_t ua = 0, ub = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { ua = max(ua, smax[0][w]); ub = max(ub, smax[1][w]); }
    const double Amax = (double)__uint_as_float(ua), Bmax = (double)__uint_as_float(ub);
    const long long src = a.order ? a.order[h] : h;
    const double* Tp = a.hyps + p * a.hyp_pair_stride + src * 12;
    float T[12];
    double Td[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) { Td[j] = Tp[j]; T[j] = (float)Td[j]; }
    const double m = prefilter_band(Td, Amax, Bmax, a.r, a.r2);                 // math3.cuh (shared with the host test)
    // NaN in T (fmax drops NaN operands) must reach the float64 path: the float products below are NaN then, both comparisons false
    const float lo = f32_at_or_below(a.r2 - m), hi = f32_at_or_above(a.r2 + m);
    const bool weighted = a.mv.scores != nullptr;
    int n_in = 0;
    auto dist2 = [&](int k) -> float {
      const float4 pa = sa[k], pb = sb[k];
      return prefilter_dist2(T, pa.x, pa.y, pa.z, pb.x, pb.y, pb.z);
    };
    auto count = [&](int k, bool in) {
      if (weighted) { if (in) acc += sk[k][6]; }      // same order of float64 additions as ransac_score_kernel
      else n_in += in ? 1 : 0;
    };
    int k = 0;
    for (; k + 4 <= cnt; k += 4) {                    // four tests per branch: the band is hit by ~1e-4 of the tests
      float s2[4]; bool in[4]; bool amb = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) { s2[j] = dist2(k + j); in[j] = s2[j] < lo; amb |= !in[j] && !(s2[j] > hi); }
      if (amb) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (!in[j] && !(s2[j] > hi)) in[j] = is_inlier_recheck(Tp, &sk[k + j][0], a.r2);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) count(k + j, in[j]);
    }
    for (; k < cnt; ++k) {
      const float s2 = dist2(k);
      bool in = s2 < lo;
      if (!in && !(s2 > hi)) in = is_inlier_recheck(Tp, &sk[k][0], a.r2);
      count(k, in);
    }
    if (!weighted) acc = (double)n_in;            // scores == NULL: every weight is 1.0, the float64 running sum is this integer
  }
  if (h < a.H) a.partial[((long long)p * a.tiles + tile) * a.H + h] = acc;
  __syncthreads();                          // the staged tile is overwritten by the next item
  }
}

// first-best selection (strict '>' from 0, test/estimator.py:427-436).  One CTA per pair.
struct SelectArgs {
  const double* partial; int tiles; int H;
  const int32_t* n_matches; int K;
  double* overlaps;            // [B][H] optional
  int32_t* best_id;            // [B]
  double* best_overlap;        // [B]
};

__global__ void __launch_bounds__(256) ransac_select_kernel(SelectArgs a) {
  __shared__ double sv[8]; __shared__ int si[8];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int K = a.n_matches ? a.n_matches[p] : a.K;
  double bv = 0.0; int bi = 0x7fffffff;
  for (int h = tid; h < a.H; h += 256) {
    double s = 0.0;
    for (int t = 0; t < a.tiles; ++t) s += a.partial[((long long)p * a.tiles + t) * a.H + h];
    const double ov = s / (double)K;                 // np.sum(scores[overlap]) / scores.shape[0]
    if (a.overlaps) a.overlaps[(long long)p * a.H + h] = ov;
    if (ov > bv) { bv = ov; bi = h; }                // increasing h per thread: keeps the first maximum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double vo = __shfl_xor_sync(0xffffffffu, bv, o);
    const int io = __shfl_xor_sync(0xffffffffu, bi, o);
    if (vo > bv || (vo == bv && io < bi)) { bv = vo; bi = io; }
  }
  if ((tid & 31) == 0) { sv[tid >> 5] = bv; si[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
    const bool none = !(bv > 0.0);
    a.best_id[p] = none ? -1 : bi;
    a.best_overlap[p] = none ? 0.0 : bv;
  }
}

// ---------------------------------------------------------------------------------------------
// refiner: two rounds (radius 2*ird, then ird) of weighted Kabsch on the inliers.  One CTA per pair.
//   w_i = s_i / sum(s) ; c0 = sum w k0 ; c1 = sum w k1 ; H = (k0-c0)^T diag(w) (k1-c1)
//   R = U V^T (no reflection fix) ; t = c0 - R c1
// Block reductions run in a fixed order (warp shuffle tree, then warps 0..15 sequentially): the
// result is deterministic run to run.
// ---------------------------------------------------------------------------------------------
struct RefineArgs {
  MatchView mv;
  const double* T_in; long long T_pair_stride;     // explicit [3][4] per pair, or
  const double* hyps; long long hyp_pair_stride; const int32_t* order; const int32_t* best_id;  // hyps[order[best]]
  double rad0, rad1; int rounds;   // radius of round 0 / round 1; rounds = 1 or 2
  double* T_out;               // [B][16]
  uint8_t* inlier_mask;        // [B][cap] optional
  int mask_stride;
};

template <int NV>
__device__ __forceinline__ void block_reduce(double v[NV], double* sm /*[16][NV]*/, double out[NV]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const double r = warp_sum_d(v[j]);
    if (lane == 0) sm[warp * NV + j] = r;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double s = 0.0;
    for (int w = 0; w < 16; ++w) s += sm[w * NV + j];
    out[j] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) refine_kernel(RefineArgs a) {
  __shared__ double red[16 * 9];
  __shared__ double Tcur[12];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int K = mv_count(a.mv, p);
  if (tid < 12) {
    double v;
    if (a.T_in) v = a.T_in[p * a.T_pair_stride + tid];
    else {
      const int b = a.best_id[p];
      if (b < 0) v = (tid % 5 == 0) ? 1.0 : 0.0;      // no hypothesis scored: identity in, flagged by best_id = -1
      else { const long long src = a.order ? a.order[b] : b; v = a.hyps[p * a.hyp_pair_stride + src * 12 + tid]; }
    }
    Tcur[tid] = v;
  }
  __syncthreads();
  for (int round = 0; round < a.rounds; ++round) {
    const double rad = (round == 0) ? a.rad0 : a.rad1;
    const double r2 = rad * rad;
    double T[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) T[j] = Tcur[j];
    double acc7[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int k = tid; k < K; k += 512) {
      double x[3], y[3], s;
      mv_load(a.mv, p, k, x, y, s);
      const bool in = is_inlier(T, x, y, r2);
      if (round == a.rounds - 1 && a.inlier_mask) a.inlier_mask[(long long)p * a.mask_stride + k] = in ? 1 : 0;
      if (in) {
        acc7[0] += s;
        acc7[1] += s * x[0]; acc7[2] += s * x[1]; acc7[3] += s * x[2];
        acc7[4] += s * y[0]; acc7[5] += s * y[1]; acc7[6] += s * y[2];
      }
    }
    double tot[7];
    block_reduce<7>(acc7, red, tot);
    const double inv = 1.0 / tot[0];
    const double c0[3] = {tot[1] * inv, tot[2] * inv, tot[3] * inv};
    const double c1[3] = {tot[4] * inv, tot[5] * inv, tot[6] * inv};
    double h9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = tid; k < K; k += 512) {
      double x[3], y[3], s;
      mv_load(a.mv, p, k, x, y, s);
      if (is_inlier(T, x, y, r2)) {
        const double w = s * inv;
        const double ax = x[0] - c0[0], ay = x[1] - c0[1], az = x[2] - c0[2];
        const double bx = (y[0] - c1[0]) * w, by = (y[1] - c1[1]) * w, bz = (y[2] - c1[2]) * w;
        h9[0] += ax * bx; h9[1] += ax * by; h9[2] += ax * bz;
        h9[3] += ay * bx; h9[4] += ay * by; h9[5] += ay * bz;
        h9[6] += az * bx; h9[7] += az * by; h9[8] += az * bz;
      }
    }
    double Hm[9];
    block_reduce<9>(h9, red, Hm);
    if (tid == 0) {
      double R[9];
      polar_uvt(Hm, R);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        Tcur[4 * r] = R[3 * r]; Tcur[4 * r + 1] = R[3 * r + 1]; Tcur[4 * r + 2] = R[3 * r + 2];
        Tcur[4 * r + 3] = c0[r] - (c1[0] * R[3 * r] + c1[1] * R[3 * r + 1] + c1[2] * R[3 * r + 2]);
      }
    }
    __syncthreads();
  }
  if (tid < 16) {
    double v = (tid < 12) ? Tcur[tid] : ((tid == 15) ? 1.0 : 0.0);
    a.T_out[(long long)p * 16 + tid] = v;
  }
}

}  // namespace roreg
