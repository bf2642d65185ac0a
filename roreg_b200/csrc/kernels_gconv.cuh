// kernels_gconv.cuh - data movement around the tcgen05 GEMM for the icosahedral group convolutions
// (network/group_feat.py:20-24 data_process + network/ops.py:11-63 Comb_Conv / Residual_Comb_Conv,
//  network/eqv_trans.py:103-138, network/rot_detect.py:43-55).
//
// Activations live channel-LAST: act[(item*60 + g)][c], split into tf32 hi / lo parts by the producing
// epilogue.  A group convolution  out[o,g] = b[o] + sum_{c,k} W[o,c,k] act[c, N[g,k]]  is the GEMM of
//   A[(item, g)][(k, c)] = act[(item*60 + N[g,k])][c]        (the 13 group neighbours of g)
// with W_flat[o][(k,c)].  A is IMPLICIT: the GEMM's loader warps gather the activation rows with
// cp.async (kernels_gemm_tc.cuh, GemmArgs.g_*) - round 1 materialised it with an im2col
// pass (13x the activation bytes written and read back).  Only the output group elements a later stage needs are
// computed (`gset`), which is how ET's "only g = 0 of the head is used" (network/eqv_trans.py:136) prunes the two last layers.
#pragma once
#include "common.cuh"

namespace roreg {

// ---- descriptors [*,32,60] (channel-first) -> channel-last activation rows, BN + ReLU + tf32 split --------
struct PackArgs {
  const float* src[4]; const int32_t* rows[4]; int permute[4]; int n_src;
  const int32_t* pre_idx;        // [n_items] group element whose P row permutes the flagged sources (eqv_trans.py:126-128)
  const uint8_t* perm;           // [60][60]
  int n_items;
  const float* bn_scale; const float* bn_shift; int relu;   // per output channel [n_src*32] or NULL
  float* out_hi; float* out_lo;  // [n_items*60][n_src*32]
};

__global__ void __launch_bounds__(256) pack_desc_kernel(PackArgs a) {
  __shared__ float tile[32][61];
  const int item = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const long long row = a.rows[s] ? a.rows[s][item] : item;
  const float* src = a.src[s] + row * (long long)RR_ROW;
  for (int e = tid; e < RR_ROW; e += 256) tile[e / 60][e % 60] = src[e];
  __syncthreads();
  const uint8_t* pr = (a.permute[s] && a.pre_idx) ? a.perm + a.pre_idx[item] * 60 : nullptr;
  const int C = a.n_src * 32;
  for (int e = tid; e < RR_ROW; e += 256) {
    const int g = e >> 5, c = e & 31;                       // consecutive threads -> consecutive channels
    float y = tile[c][pr ? pr[g] : g];
    const int oc = s * 32 + c;
    if (a.bn_scale) y = fmaf(y, a.bn_scale[oc], a.bn_shift[oc]);
    if (a.relu) y = fmaxf(y, 0.f);
    uint32_t tb; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tb) : "f"(y));
    const float hi = __uint_as_float(tb);
    const long long o = ((long long)item * 60 + g) * C + oc;
    a.out_hi[o] = hi;
    if (a.out_lo) a.out_lo[o] = y - hi;
  }
}

// ---- GF tail: eqv = (conv + x) / max(||.||_c, 1e-4), back to channel-first  (network/group_feat.py:38-42) ----
__global__ void __launch_bounds__(256) gf_finalize_kernel(const float* __restrict__ conv, const float* __restrict__ x, int n,
                                                          float* __restrict__ eqv) {
  __shared__ float tile[60][33];
  const int item = blockIdx.x, tid = threadIdx.x;
  for (int e = tid; e < RR_ROW; e += 256) { const int g = e >> 5, c = e & 31; tile[g][c] = conv[((long long)item * 60 + g) * 32 + c]; }
  __syncthreads();
  for (int e = tid; e < RR_ROW; e += 256) { const int c = e / 60, g = e % 60; tile[g][c] += x[(long long)item * RR_ROW + e]; }
  __syncthreads();
  __shared__ float inv_norm[60];
  if (tid < 60) {
    float ss = 0.f;
    for (int c = 0; c < 32; ++c) ss += tile[tid][c] * tile[tid][c];
    inv_norm[tid] = 1.0f / fmaxf(sqrtf(ss), 1e-4f);
  }
  __syncthreads();
  for (int e = tid; e < RR_ROW; e += 256) { const int c = e / 60, g = e % 60; eqv[(long long)item * RR_ROW + e] = tile[g][c] * inv_norm[g]; }
}

// ---- RD tail: feat = raw / ||raw||_c (16 channels), channel-first, zero-padded to 32 channels so that the
// autocorrelation reuses the group-correlation kernels  (network/rot_detect.py:46-51)
__global__ void __launch_bounds__(256) rd_finalize_kernel(const float* __restrict__ raw, int n, float* __restrict__ feat) {
  __shared__ float tile[60][17];
  __shared__ float inv_norm[60];
  const int item = blockIdx.x, tid = threadIdx.x;
  for (int e = tid; e < 960; e += 256) { const int g = e >> 4, c = e & 15; tile[g][c] = raw[((long long)item * 60 + g) * 16 + c]; }
  __syncthreads();
  if (tid < 60) {
    float ss = 0.f;
    for (int c = 0; c < 16; ++c) ss += tile[tid][c] * tile[tid][c];
    inv_norm[tid] = 1.0f / sqrtf(ss);
  }
  __syncthreads();
  for (int e = tid; e < RR_ROW; e += 256) {
    const int c = e / 60, g = e % 60;
    feat[(long long)item * RR_ROW + e] = (c < 16) ? tile[g][c] * inv_norm[g] : 0.f;
  }
}

// unbiased std over the 60 correlation values of each row (torch.std default, rot_detect.py:52)
__global__ void row_std60_kernel(const float* __restrict__ cor, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = cor + (long long)i * 60;
  float mean = 0.f;
  for (int j = 0; j < 60; ++j) mean += p[j];
  mean /= 60.f;
  float ss = 0.f;
  for (int j = 0; j < 60; ++j) { const float d = p[j] - mean; ss += d * d; }
  out[i] = sqrtf(ss / 59.f);
}

// quaternion_pre / ||quaternion_pre||  (network/eqv_trans.py:137)
__global__ void quat_normalize_kernel(const float* __restrict__ q_in, int ld, int K, float* __restrict__ q_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const float w = q_in[(long long)i * ld], x = q_in[(long long)i * ld + 1], y = q_in[(long long)i * ld + 2], z = q_in[(long long)i * ld + 3];
  const float nrm = sqrtf(w * w + x * x + y * y + z * z);
  q_out[4 * i] = w / nrm; q_out[4 * i + 1] = x / nrm; q_out[4 * i + 2] = y / nrm; q_out[4 * i + 3] = z / nrm;
}

// ---- all-pairs tail: squared norms of the channel-last rows and the per-row minimum of
// dist(n,m) = |X_n|^2 + |Y_m|^2 - 2 max_a cor_a(n,m)    ---------------------------------------
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ hi, const float* __restrict__ lo, long long rows, int K,
                                                         float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long r = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < K; c += 32) { const float v = hi[r * K + c] + (lo ? lo[r * K + c] : 0.f); s += v * v; }
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}

__global__ void __launch_bounds__(256) allpairs_rowmin_kernel(const float* __restrict__ best, const uint8_t* __restrict__ besta, int N, int M,
                                                              const float* __restrict__ nx, const float* __restrict__ ny,
                                                              int32_t* __restrict__ nn, int32_t* __restrict__ nn_a, float* __restrict__ nn_dist) {
  __shared__ float sv[8]; __shared__ int si[8];
  const int n = blockIdx.x, tid = threadIdx.x;
  float bv = INFINITY; int bi = 0x7fffffff;
  for (int m = tid; m < M; m += 256) {
    const float d = nx[n] + ny[m] - 2.f * best[(long long)n * M + m];
    if (d < bv) { bv = d; bi = m; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float vo = __shfl_xor_sync(0xffffffffu, bv, o); const int io = __shfl_xor_sync(0xffffffffu, bi, o);
    if (vo < bv || (vo == bv && io < bi)) { bv = vo; bi = io; }
  }
  if ((tid & 31) == 0) { sv[tid >> 5] = bv; si[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) if (sv[w] < bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
    nn[n] = bi; nn_dist[n] = bv; nn_a[n] = besta[(long long)n * M + bi];
  }
}

__global__ void fill_f32_kernel(float* __restrict__ p, long long n, float v) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) p[i] = v;
}

}  // namespace roreg
