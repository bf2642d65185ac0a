// kernels_corr.cuh - equivariant (icosahedral) correlation on (X row, Y row) pairs.
//   variant 1  test/estimator.py:85-89      cor[a] = sum_{f,g} X[f,P[a,g]] * Y[f,g]
//   variant 2  network/rot_coh_match.py:158-163  cor[h] = sum_{f,g} X[f,P[g,h]] * Y[f,g]
// Both are generalised diagonal sums of the 60x60 Gram matrix G[h,g] = sum_f X[f,h] Y[f,g]:
//   cor[a] = sum_g G[tab[a][g], g]   with tab = P (variant 1) or P^T (variant 2).
// The reference materialises X[:, :, P.flat] = [K,32,3600] (2.3 GB at K = 5000); here each match
// costs one 7680-byte read per side and the Gram never leaves the SM.
#pragma once
#include "common.cuh"

namespace roreg {

struct CorrArgs {
  const float* X; const float* Y;
  const int32_t* idxX; const int32_t* idxY; int idx_stride;   // element stride of the index arrays (1, or 2 for [K][2] match rows)
  const int32_t* pair_cloud;       // batched: [B][2] (cloud id0, cloud id1); X rows come from id1, Y rows from id0
  int n;                            // keypoints per cloud (batched addressing)
  const int32_t* n_matches;         // batched: [B] device counts; NULL -> K
  int K;                            // rows per pair (capacity when batched)
  int B;
  const uint8_t* tab;               // [60][60] gather table
  float* cor_out; int32_t* argmax_out;
};

// v1: FP32 CUDA-core Gram, exact float32 products, sequential-f FMA accumulation.
// 128 threads: thread t owns the 4x8 block G[h0..h0+3][g0..g0+7], h0 = (t/8)*4, g0 = (t%8)*8.
__global__ void __launch_bounds__(128) group_corr_kernel(CorrArgs a) {
  __shared__ __align__(16) float Xs[RR_F][64];
  __shared__ __align__(16) float Ys[RR_F][64];
  __shared__ float Gs[64][65];               // transposed: Gs[g][h]
  __shared__ uint8_t tabs[3600];
  __shared__ float red_v[2]; __shared__ int red_i[2];
  const int tid = threadIdx.x;
  for (int e = tid; e < 3600; e += 128) tabs[e] = a.tab[e];
  for (int e = tid; e < RR_F * 4; e += 128) { Xs[e >> 2][60 + (e & 3)] = 0.f; Ys[e >> 2][60 + (e & 3)] = 0.f; }
  const long long total = (long long)a.B * a.K;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const int p = (int)(w / a.K), k = (int)(w % a.K);
    if (a.n_matches && k >= a.n_matches[p]) continue;      // uniform across the CTA
    long long rx = a.idxX ? a.idxX[w * a.idx_stride] : k;
    long long ry = a.idxY ? a.idxY[w * a.idx_stride] : k;
    if (a.pair_cloud) { rx += (long long)a.pair_cloud[2 * p + 1] * a.n; ry += (long long)a.pair_cloud[2 * p] * a.n; }
    const float4* xs = reinterpret_cast<const float4*>(a.X + rx * RR_ROW);
    const float4* ys = reinterpret_cast<const float4*>(a.Y + ry * RR_ROW);
    __syncthreads();                         // previous match fully consumed
#pragma unroll
    for (int q = tid; q < 480; q += 128) {
      const float4 u = ldg_stream4(xs + q), v = ldg_stream4(ys + q);
      const int e = q * 4, f = e / 60, g = e % 60;
      *reinterpret_cast<float4*>(&Xs[f][g]) = u;
      *reinterpret_cast<float4*>(&Ys[f][g]) = v;
    }
    __syncthreads();
    const int h0 = (tid >> 3) * 4, g0 = (tid & 7) * 8;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int f = 0; f < RR_F; ++f) {
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[f][h0]);
      const float4 y0 = *reinterpret_cast<const float4*>(&Ys[f][g0]);
      const float4 y1 = *reinterpret_cast<const float4*>(&Ys[f][g0 + 4]);
      const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
      const float yy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xx[i], yy[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) Gs[g0 + j][h0 + i] = acc[i][j];
    __syncthreads();
    float c = -INFINITY;
    if (tid < RR_G) {
      c = 0.f;
      const uint8_t* t = tabs + tid * 60;
#pragma unroll 10
      for (int g = 0; g < RR_G; ++g) c += Gs[g][t[g]];
      if (a.cor_out) a.cor_out[w * RR_G + tid] = c;
    }
    if (a.argmax_out) {
      // first maximal index (torch.argmax): lexicographic (max value, min index)
      if (tid < 64) {
        float v = c; int ix = (tid < RR_G) ? tid : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float vo = __shfl_xor_sync(0xffffffffu, v, o);
          const int io = __shfl_xor_sync(0xffffffffu, ix, o);
          if (vo > v || (vo == v && io < ix)) { v = vo; ix = io; }
        }
        if ((tid & 31) == 0) { red_v[tid >> 5] = v; red_i[tid >> 5] = ix; }
      }
      __syncthreads();
      if (tid == 0) {
        int ix = red_i[0];
        if (red_v[1] > red_v[0]) ix = red_i[1];      // warp 1 holds the larger indices
        a.argmax_out[w] = ix;
      }
    }
  }
}

}  // namespace roreg
