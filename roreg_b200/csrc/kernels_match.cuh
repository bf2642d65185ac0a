// kernels_match.cuh - invariant pooling, brute-force NN / kNN, mutual-check compaction.
// Reference arithmetic: test/matcher.py:69-72,94-106 and utils/knn_search.py:17-66.
#pragma once
#include "common.cuh"

namespace roreg {

// ---------------------------------------------------------------------------------------------
// inv_pool: one warp per keypoint.  The 7680-byte descriptor row is read as 480 float4, 15 per lane,
// fully coalesced (each warp-wide request covers 512 contiguous bytes); a float4 never straddles a
// channel because 60 % 4 == 0.  HBM-bound: 7680 B in, 128 B out per keypoint.
//   batched addressing: row r -> (pair, side, i); cloud = pair_cloud[2*pair+side];
//   src keypoint = sample ? sample[r] : i.   pair_cloud == NULL -> single cloud at `desc`.
// ---------------------------------------------------------------------------------------------
struct PoolArgs {
  const float* desc; const int32_t* pair_cloud; const int32_t* sample;
  int n;            // keypoints per cloud in the arena
  int S;            // rows per (pair, side)
  int rows;         // total rows
  int normalise;
  float* out;       // [rows][32]
};

__global__ void __launch_bounds__(256) inv_pool_kernel(PoolArgs a) {
  __shared__ float part[8][480];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= a.rows) return;
  int i = r % a.S;
  long long cloud = 0;
  if (a.pair_cloud) cloud = a.pair_cloud[r / a.S];
  const int srcrow = a.sample ? a.sample[r] : i;
  const float4* src = reinterpret_cast<const float4*>(a.desc + (cloud * a.n + srcrow) * (long long)RR_ROW);
  float4 v[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) v[k] = ldg_stream4(src + k * 32 + lane);
#pragma unroll
  for (int k = 0; k < 15; ++k) part[warp][k * 32 + lane] = (v[k].x + v[k].y) + (v[k].z + v[k].w);
  __syncwarp();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 15; ++j) s += part[warp][lane * 15 + j];
  float m = s / 60.0f;
  if (a.normalise) {
    const float ss = warp_sum(m * m);
    m = m / (sqrtf(ss) + 1e-5f);
  }
  a.out[(long long)r * RR_F + lane] = m;
}

// ---------------------------------------------------------------------------------------------
// nn_diff_kernel: 1-NN of every source row among the target rows, float32 difference form
// d = sqrt(sum_f (a-b)^2 + 1e-7), result = lexicographic min of (d, index) which is exactly
// torch's `dist.min(dim=1)` (first minimal index) on the reference's distance values.
// CTA = 64 source rows x all targets (64-column tiles), 256 threads, 4x4 register micro-tile.
// blockIdx.y = pair; src/tgt/out advance by the given per-pair strides.
// ---------------------------------------------------------------------------------------------
struct NNArgs {
  const float* src; const float* tgt; long long src_pair_stride, tgt_pair_stride;
  int n_src, n_tgt;
  int32_t* out_idx; float* out_dist; long long out_pair_stride;
};

__global__ void __launch_bounds__(256) nn_diff_kernel(NNArgs a) {
  __shared__ __align__(16) float As[RR_F][64];
  __shared__ __align__(16) float Bs[RR_F][64];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const float* src = a.src + blockIdx.y * a.src_pair_stride;
  const float* tgt = a.tgt + blockIdx.y * a.tgt_pair_stride;
  const int row0 = blockIdx.x * 64;
  {  // A tile: thread -> (row = tid%64, 8 channels starting at (tid/64)*8)
    const int r = tid & 63, c0 = (tid >> 6) * 8;
    float4 u = make_float4(0, 0, 0, 0), w = u;
    if (row0 + r < a.n_src) {
      const float4* p = reinterpret_cast<const float4*>(src + (long long)(row0 + r) * RR_F + c0);
      u = __ldg(p); w = __ldg(p + 1);
    }
    As[c0 + 0][r] = u.x; As[c0 + 1][r] = u.y; As[c0 + 2][r] = u.z; As[c0 + 3][r] = u.w;
    As[c0 + 4][r] = w.x; As[c0 + 5][r] = w.y; As[c0 + 6][r] = w.z; As[c0 + 7][r] = w.w;
  }
  float best_s[4], best_d2[4]; int best_i[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best_s[i] = INFINITY; best_d2[i] = INFINITY; best_i[i] = 0x7fffffff; }

  for (int col0 = 0; col0 < a.n_tgt; col0 += 64) {
    __syncthreads();     // previous tile fully consumed (and As visible on the first pass)
    {
      const int r = tid & 63, c0 = (tid >> 6) * 8;
      float4 u = make_float4(0, 0, 0, 0), w = u;
      if (col0 + r < a.n_tgt) {
        const float4* p = reinterpret_cast<const float4*>(tgt + (long long)(col0 + r) * RR_F + c0);
        u = __ldg(p); w = __ldg(p + 1);
      }
      Bs[c0 + 0][r] = u.x; Bs[c0 + 1][r] = u.y; Bs[c0 + 2][r] = u.z; Bs[c0 + 3][r] = u.w;
      Bs[c0 + 4][r] = w.x; Bs[c0 + 5][r] = w.y; Bs[c0 + 6][r] = w.z; Bs[c0 + 7][r] = w.w;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int f = 0; f < RR_F; ++f) {
      const float4 av = *reinterpret_cast<const float4*>(&As[f][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[f][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float d = aa[i] - bb[j]; acc[i][j] = fmaf(d, d, acc[i][j]); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = col0 + tx * 4 + j;
      if (gc < a.n_tgt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float d2 = acc[i][j];
          if (d2 < best_d2[i]) {            // necessary for sqrt(d2+1e-7) < best_s (sqrt is monotone)
            const float s = sqrtf(d2 + 1e-7f);
            if (s < best_s[i]) { best_s[i] = s; best_i[i] = gc; best_d2[i] = d2; }
          }
        }
      }
    }
  }
  // lexicographic (s, idx) min across the 16 tx lanes that share these 4 rows
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = best_s[i]; int ix = best_i[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float so = __shfl_xor_sync(0xffffffffu, s, o);
      const int io = __shfl_xor_sync(0xffffffffu, ix, o);
      if (so < s || (so == s && io < ix)) { s = so; ix = io; }
    }
    const int row = row0 + ty * 4 + i;
    if (tx == 0 && row < a.n_src) {
      a.out_idx[blockIdx.y * a.out_pair_stride + row] = ix;
      if (a.out_dist) a.out_dist[blockIdx.y * a.out_pair_stride + row] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// knn_small_kernel: generic k <= 16, f <= 32 brute force (NMS_sample's 5-NN on xyz, test/matcher.py:22).
// One thread per source row, targets streamed through shared memory, sorted insertion under the
// lexicographic (d, index) order.
// ---------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(128) knn_small_kernel(const float* __restrict__ tgt, int n,
                                                        const float* __restrict__ src, int m, int f, int k,
                                                        float* __restrict__ dist, int32_t* __restrict__ idx) {
  extern __shared__ float tile[];            // [128][f]
  const int row = blockIdx.x * 128 + threadIdx.x;
  float q[RR_F];
#pragma unroll
  for (int c = 0; c < RR_F; ++c) q[c] = (row < m && c < f) ? src[(long long)row * f + c] : 0.f;
  float bd[KMAX]; int bi[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { bd[j] = INFINITY; bi[j] = 0x7fffffff; }
  for (int t0 = 0; t0 < n; t0 += 128) {
    __syncthreads();
    const int cnt = min(128, n - t0);
    for (int e = threadIdx.x; e < cnt * f; e += 128) tile[e] = tgt[(long long)t0 * f + e];
    __syncthreads();
    if (row < m) {
      for (int t = 0; t < cnt; ++t) {
        float d2 = 0.f;
#pragma unroll
        for (int c = 0; c < RR_F; ++c)
          if (c < f) { const float d = q[c] - tile[t * f + c]; d2 = fmaf(d, d, d2); }
        const float s = sqrtf(d2 + 1e-7f);
        float kth = INFINITY;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) if (j == k - 1) kth = bd[j];
        if (s < kth) {                       // strictly better than the current k-th (ties keep the lower index)
          float cs = s; int ci = t0 + t;
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
            if (j < k && (cs < bd[j])) { const float ts = bd[j]; const int ti = bi[j]; bd[j] = cs; bi[j] = ci; cs = ts; ci = ti; }
          }
        }
      }
    }
  }
  if (row < m)
    for (int j = 0; j < k; ++j) { dist[(long long)row * k + j] = bd[j]; idx[(long long)row * k + j] = bi[j]; }
}

// ---------------------------------------------------------------------------------------------
// mutual_compact_kernel: the Python loop of test/matcher.py:98-105 as an ordered stream compaction.
// One CTA per pair; matches are emitted in increasing row of cloud 0 as ORIGINAL keypoint indices
// (sample0[i], sample1[nn01[i]]).
// ---------------------------------------------------------------------------------------------
struct CompactArgs {
  const int32_t* nn01; const int32_t* nn10; int n0, n1; long long nn_pair_stride;
  const int32_t* sample;       // [B][2][S] or NULL
  int S;
  int32_t* matches;            // [B][cap][2]
  int cap;
  int32_t* n_matches;          // [B]
};

__global__ void __launch_bounds__(1024) mutual_compact_kernel(CompactArgs a) {
  __shared__ int warp_tot[32];
  __shared__ int running;
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t* nn01 = a.nn01 + p * a.nn_pair_stride;
  const int32_t* nn10 = a.nn10 + p * a.nn_pair_stride;
  if (tid == 0) running = 0;
  __syncthreads();
  for (int base = 0; base < a.n0; base += 1024) {
    const int i = base + tid;
    int j = -1; bool keep = false;
    if (i < a.n0) { j = nn01[i]; keep = (j >= 0 && j < a.n1 && nn10[j] == i); }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int wpre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int t = warp_tot[w]; if (w < warp) wbase += t; total += t; }
    const int start = running;
    if (keep) {
      const int pos = start + wbase + wpre;
      int o0 = i, o1 = j;
      if (a.sample) { o0 = a.sample[(long long)(2 * p) * a.S + i]; o1 = a.sample[(long long)(2 * p + 1) * a.S + j]; }
      a.matches[((long long)p * a.cap + pos) * 2 + 0] = o0;
      a.matches[((long long)p * a.cap + pos) * 2 + 1] = o1;
    }
    __syncthreads();
    if (tid == 0) running = start + total;
    __syncthreads();
  }
  if (tid == 0) a.n_matches[p] = running;
}

}  // namespace roreg
