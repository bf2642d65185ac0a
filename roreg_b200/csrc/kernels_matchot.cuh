// kernels_matchot.cuh - small kernels of the rotation-coherence matcher Match_ot
// (network/rot_coh_match.py:8-390).  The dense 1x1 layers and the [m,n] score matrices run on the tcgen05
// GEMM (kernels_gemm_tc.cuh) and the R-indicator on the group-correlation kernels (variant 2); what is
// here is the glue the reference expresses with argsort / advanced indexing / softmax / InstanceNorm /
// logsumexp:  top-k of a score row, row gathers, the 4-head neighbourhood attention, per-channel instance
// statistics, operand preparation (concat + normalise + tf32 split) and the log-domain Sinkhorn passes.
// Activations are channel-last rows [positions][C] float32.
#pragma once
#include "common.cuh"

namespace roreg {

// ---- top-k of each row of S [m][n] (descending value, ties -> lower column), k <= 16 ---------------------------
// Knn_index_extract (rot_coh_match.py:34-45) sorts the whole row; only the first k columns are ever used.
__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ S, int n, int ld, int k, int32_t* __restrict__ idx) {
  extern __shared__ float row[];                 // [n]
  __shared__ float rv[8]; __shared__ int ri[8];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* src = S + (long long)r * ld;
  for (int j = tid; j < n; j += 256) row[j] = src[j];
  __syncthreads();
  for (int t = 0; t < k; ++t) {
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int j = tid; j < n; j += 256) { const float v = row[j]; if (v > bv) { bv = v; bi = j; } }   // increasing j: first max per thread
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float vo = __shfl_xor_sync(0xffffffffu, bv, o); const int io = __shfl_xor_sync(0xffffffffu, bi, o);
      if (vo > bv || (vo == bv && io < bi)) { bv = vo; bi = io; }
    }
    if ((tid & 31) == 0) { rv[tid >> 5] = bv; ri[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w) if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
      idx[(long long)r * k + t] = bi;
      if (bi < n) row[bi] = -INFINITY;
    }
    __syncthreads();
  }
}

// ---- out[(i*k + j)][c] = src[idx[i*k + j]][c]   (C % 4 == 0) -----------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                          long long n_out, int C, float* __restrict__ out) {
  const int c4 = C >> 2;
  const long long total = n_out * c4;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const long long r = e / c4; const int q = (int)(e % c4);
    reinterpret_cast<float4*>(out)[e] = reinterpret_cast<const float4*>(src)[(long long)idx[r] * c4 + q];
  }
}

// ---- relative neighbour coordinates: out[(i*k+j)][0..2] = (coor[idx[i*k+j]] - coor[i]) in units of `step`, zero-padded to 32
__global__ void __launch_bounds__(256) rel_coor_kernel(const float* __restrict__ coor /*[m][3]*/, const int32_t* __restrict__ idx,
                                                       int m, int k, float step, float* __restrict__ out /*[m*k][32]*/) {
  const long long total = (long long)m * k * 32;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const long long r = e >> 5; const int c = (int)(e & 31);
    float v = 0.f;
    if (c < 3) v = __fdiv_rn(coor[(long long)idx[r] * 3 + c], step) - __fdiv_rn(coor[(r / k) * 3 + c], step);   // keys / coor_norm_step (:342-343)
    out[e] = v;
  }
}

// ---- per-channel instance statistics over P rows of x [P][C] (InstanceNorm2d, affine=False, biased variance) ------
__global__ void __launch_bounds__(256) chan_stats_kernel(const float* __restrict__ x, long long P, int C, double* __restrict__ acc /*[C][2]*/) {
  // thread -> channel c = tid % C, row lane = tid / C ; requires C <= 256 and 256 % C == 0
  const int c = threadIdx.x % C, rl = threadIdx.x / C, rpb = 256 / C;
  double s = 0, ss = 0;
  for (long long r = (long long)blockIdx.x * rpb + rl; r < P; r += (long long)gridDim.x * rpb) {
    const double v = x[r * C + c]; s += v; ss += v * v;
  }
  // one atomic pair per channel and CTA (run c6: 256 contended double atomics per CTA made this 85 us per call)
  __shared__ double sh[256][2];
  sh[threadIdx.x][0] = s; sh[threadIdx.x][1] = ss;
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < rpb; ++k) { s += sh[k * C + c][0]; ss += sh[k * C + c][1]; }
    atomicAdd(&acc[2 * c], s); atomicAdd(&acc[2 * c + 1], ss);
  }
}
__global__ void chan_stats_finish_kernel(const double* __restrict__ acc, long long P, int C, float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc[2 * c] / (double)P;
  double v = acc[2 * c + 1] / (double)P - m * m;
  if (v < 0) v = 0;
  mean[c] = (float)m; rstd[c] = (float)(1.0 / sqrt(v + 1e-5));
}

// ---- GEMM operand preparation: concat up to 3 sources along channels, optional per-source row broadcast and
// L2 normalisation over its channels, optional instance-norm + ReLU, zero pad to Kout, tf32 hi/lo split ------------
struct PrepArgs {
  const float* src[3]; int C[3]; int row_div[3]; int l2norm[3]; int n_src;
  const float* mean; const float* rstd; int relu;      // instance norm (applied to the single source) or NULL
  long long P; int Kout;
  float* hi; float* lo; float* plain;                  // plain (optional): the float32 value before the split
};
__global__ void __launch_bounds__(256) prep_rows_kernel(PrepArgs a) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= a.P) return;
  int off = 0;
  for (int s = 0; s < a.n_src; ++s) {
    const float* p = a.src[s] + (r / a.row_div[s]) * a.C[s];
    float scale = 1.f;
    if (a.l2norm[s]) {
      float ss = 0.f;
      for (int c = lane; c < a.C[s]; c += 32) ss += p[c] * p[c];
      scale = 1.0f / sqrtf(warp_sum(ss));
    }
    for (int c = lane; c < a.C[s]; c += 32) {
      float y = p[c] * scale;
      if (a.mean) { y = (y - a.mean[c]) * a.rstd[c]; }
      if (a.relu) y = fmaxf(y, 0.f);
      uint32_t tb; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tb) : "f"(y));
      const float h = __uint_as_float(tb);
      a.hi[r * a.Kout + off + c] = h; a.lo[r * a.Kout + off + c] = y - h;
      if (a.plain) a.plain[r * a.Kout + off + c] = y;
    }
    off += a.C[s];
  }
  for (int c = off + lane; c < a.Kout; c += 32) { a.hi[r * a.Kout + c] = 0.f; a.lo[r * a.Kout + c] = 0.f; if (a.plain) a.plain[r * a.Kout + c] = 0.f; }
}

// ---- 4-head attention of each point over its k neighbours (rot_coh_match.py:84-119), projections already applied ---
// Q [m][32], Kp / Vp [m*k][32]; channel c = d*4 + h.  One warp per point, lane = channel.
__global__ void __launch_bounds__(256) mha_kernel(const float* __restrict__ Q, const float* __restrict__ Kp, const float* __restrict__ Vp,
                                                  int m, int k, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= m) return;
  const float q = Q[(long long)i * 32 + lane];
  float sc[16];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float v = -INFINITY;
    if (j < k) {
      v = q * Kp[((long long)i * k + j) * 32 + lane];
      v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);  // sum over d (lanes with equal h)
      v = v / 2.8284271247461903f;            // / dim**.5, dim = 8
    }
    sc[j] = v; mx = fmaxf(mx, v);
  }
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) { sc[j] = (j < k) ? expf(sc[j] - mx) : 0.f; den += sc[j]; }
  float o = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) if (j < k) o += (sc[j] / den) * Vp[((long long)i * k + j) * 32 + lane];
  out[(long long)i * 32 + lane] = o;
}

// ---- [rind | column max over the points] -> [m][128] rows (120 used), rot_coh_match.py:203 ---------------------------
__global__ void colmax60_kernel(const float* __restrict__ rind /*[m][60]*/, int m, float* __restrict__ cmax /*[60]*/) {
  const int h = blockIdx.x; float v = -INFINITY;
  for (int i = threadIdx.x; i < m; i += blockDim.x) v = fmaxf(v, rind[(long long)i * 60 + h]);
  __shared__ float sm[256];
  sm[threadIdx.x] = v; __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + s]); __syncthreads(); }
  if (threadIdx.x == 0) cmax[h] = sm[0];
}
__global__ void rind_rows_kernel(const float* __restrict__ rind, const float* __restrict__ cmax, int m, float* __restrict__ out /*[m][128]*/) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= (long long)m * 128) return;
  const long long i = e >> 7; const int c = (int)(e & 127);
  out[e] = (c < 60) ? rind[i * 60 + c] : (c < 120 ? cmax[c - 60] : 0.f);
}

// ---- Sinkhorn (rot_coh_match.py:277-319).  Z = [[S, alpha],[alpha, alpha]] is never materialised: S [m][n] + the bin.
// rows:  u[i] = log_mu[i] - logsumexp_j(Z[i][j] + v[j])     (one CTA per row, coalesced)
// cols:  v[j] = log_nu[j] - logsumexp_i(Z[i][j] + u[i])     (CTA = 32 columns x 8 row-lanes, online max/sum)
struct SinkArgs { const float* S; int m, n, ld; float alpha; float norm; float* u; float* v; };

__device__ __forceinline__ void lse_merge(float& mx, float& sm, float omx, float osm) {
  const float nm = fmaxf(mx, omx);
  if (nm == -INFINITY) { mx = nm; sm = 0.f; return; }
  sm = sm * expf(mx - nm) + osm * expf(omx - nm); mx = nm;
}

__global__ void __launch_bounds__(256) sinkhorn_rows_kernel(SinkArgs a) {
  __shared__ float smx[8], ssm[8];
  const int i = blockIdx.x, tid = threadIdx.x;          // i in [0, m]  (row m = the dustbin row)
  float mx = -INFINITY, sm = 0.f;
  for (int j = tid; j <= a.n; j += 256) {
    const float z = (i < a.m && j < a.n) ? a.S[(long long)i * a.ld + j] : a.alpha;
    const float t = z + a.v[j];
    if (t > mx) { sm = sm * expf(mx - t) + 1.f; mx = t; } else sm += expf(t - mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, sm, o); lse_merge(mx, sm, om, os); }
  if ((tid & 31) == 0) { smx[tid >> 5] = mx; ssm[tid >> 5] = sm; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) lse_merge(mx, sm, smx[w], ssm[w]);
    const float log_mu = (i < a.m) ? a.norm : (logf((float)a.n) + a.norm);
    a.u[i] = log_mu - (mx + logf(sm));
  }
}

// 32 columns x 32 row-lanes per CTA; every thread keeps 4 independent online (max, sum) pairs so that 4 loads are
// in flight per thread (the first version - 8 row-lanes, one dependent chain - was L2-latency bound: 200 us per pass).
__global__ void __launch_bounds__(1024) sinkhorn_cols_kernel(SinkArgs a) {
  __shared__ float smx[32][33], ssm[32][33];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;                 // j in [0, n]
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, sm[4] = {0.f, 0.f, 0.f, 0.f};
  if (j <= a.n) {
    for (int i0 = rl; i0 <= a.m; i0 += 128) {
      float t[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + 32 * q;
        t[q] = -INFINITY;
        if (i <= a.m) t[q] = ((i < a.m && j < a.n) ? a.S[(long long)i * a.ld + j] : a.alpha) + a.u[i];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (t[q] > mx[q]) { sm[q] = sm[q] * expf(mx[q] - t[q]) + 1.f; mx[q] = t[q]; }
        else if (t[q] > -INFINITY) sm[q] += expf(t[q] - mx[q]);
      }
    }
  }
  lse_merge(mx[0], sm[0], mx[1], sm[1]); lse_merge(mx[2], sm[2], mx[3], sm[3]); lse_merge(mx[0], sm[0], mx[2], sm[2]);
  smx[rl][lane] = mx[0]; ssm[rl][lane] = sm[0];
  __syncthreads();
  if (rl == 0 && j <= a.n) {
    float m0 = mx[0], s0 = sm[0];
    for (int w = 1; w < 32; ++w) lse_merge(m0, s0, smx[w][lane], ssm[w][lane]);
    const float log_nu = (j < a.n) ? a.norm : (logf((float)a.m) + a.norm);
    a.v[j] = log_nu - (m0 + logf(s0));
  }
}

// ---- Sinkhorn + final assignment as ONE persistent kernel ------------------------------------------------
// Round 1 ran the 100 iterations as 200 launches, each streaming S from HBM.  Here the grid is one CTA per SM, launched
// cooperatively; CTA c OWNS a block of consecutive rows of Z for the whole run:
//   row pass     u[i] = log_mu - LSE_j(Z[i][j] + v[j]) for its rows (v staged in shared memory);
//   column pass  for every column j the partial LSE over ITS rows of Z[i][j] + u[i] (its own, just computed u: no global read)
//                -> part[c][j] = (max, sum); grid barrier; CTA c merges the partials of ITS columns in CTA order -> v[j]; barrier.
// S (100 MB at 5000 x 5000) is read twice per iteration from L2 - it fits the 126 MB L2 - in the same float32 online
// max / sum arithmetic as before; per iteration two grid barriers instead of two launches.  The assignment
// (rot_coh_match.py:369-379) reuses the ownership: row argmax locally, column argmax through the same partial / merge step.
// N more terms of an online logsumexp, branch-free: one rescale of the running sum per group, exponentials as ex2.approx of
// an FFMA (runs c6-c8: with expf and a data-dependent branch per element the passes were arithmetic-bound; with four loads in
// flight per thread they were L2-LATENCY-bound - lts throughput 7 % - so the passes below keep 8 / 16 loads in flight).
template <int N>
__device__ __forceinline__ void lse_acc(float& mx, float& sm, const float (&t)[N]) {
  constexpr float L2E = 1.4426950408889634f;
  float mN = t[0];
#pragma unroll
  for (int q = 1; q < N; ++q) mN = fmaxf(mN, t[q]);
  if (mN == -INFINITY) return;
  const float nm = fmaxf(mx, mN), nml = nm * L2E;
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < N; ++q) acc += exp2f(fmaf(t[q], L2E, -nml));     // -inf terms give 0
  sm = sm * exp2f(fmaf(mx, L2E, -nml)) + acc;
  mx = nm;
}

struct SinkFusedArgs {
  const float* S; int m, n, ld; float alpha, norm; int iters;
  float* u; float* v;             // [m+1], [n+1]
  float2* part;                   // [gridDim.x][n+1] column partials (max, sum) / (best value, row index bits)
  int32_t* idx0; float* max0; int32_t* idx1; int32_t* matches0; float* mscores0;
  int dbg;                        // ROREG_DEBUG_SINK (bottleneck experiments only, results then WRONG): 1 = no grid barriers, 2 = barriers without fences
};

__device__ __forceinline__ void sink_grid_sync(unsigned int* bar, unsigned int& gen, int dbg = 0) {
  // sense-free generation barrier on a global counter: every CTA is resident (cooperative launch, one CTA per SM)
  __syncthreads();
  if (dbg == 1) return;
  if (threadIdx.x == 0) {
    ++gen;
    if (dbg != 2) __threadfence();
    atomicAdd(bar, 1u);
    const unsigned int target = gen * gridDim.x;
    unsigned int spins = 0;
    while (*reinterpret_cast<volatile unsigned int*>(bar) < target)
      if (++spins > (1u << 24)) { printf("roreg: sinkhorn grid barrier timed out (block %d, generation %u)\n", blockIdx.x, gen); __trap(); }   // never hang the GPU
    if (dbg != 2) __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024, 1) sinkhorn_fused_kernel(SinkFusedArgs a, unsigned int* bar) {
  extern __shared__ __align__(16) float vs[];           // [n+1] current v (row pass) / v for the assignment
  __shared__ float us[64];                              // u of the rows this CTA owns
  __shared__ float rmx[32], rsm[32]; __shared__ int rix[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, c = blockIdx.x;
  const int rpc = (a.m + 1 + G - 1) / G;                // rows per CTA (<= 64, checked by the launcher)
  const int r0 = c * rpc, r1 = min(r0 + rpc, a.m + 1);
  const int cpc = (a.n + 1 + G - 1) / G;                // columns per CTA in the merge step
  const int c0 = c * cpc, c1 = min(c0 + cpc, a.n + 1);
  unsigned int gen = 0;
  const float log_mu_bin = logf((float)a.n) + a.norm, log_nu_bin = logf((float)a.m) + a.norm;
  for (int j = tid; j <= a.n; j += 1024) vs[j] = 0.f;  // v = 0 (rot_coh_match.py:302)
  __syncthreads();
  for (int it = 0; it < a.iters; ++it) {
    // ---------------- row pass: 4 rows at a time, 256 threads per row, 16-byte loads ----------------
    // (run c9: 36 instructions per element - 64-bit index arithmetic, bounds predicates and the dustbin select on every load -
    // made the passes issue-bound at ~50 us per iteration; the interior [m][n] block is now read as float4 / float2 without
    // predicates and the dustbin row / column are added as separate terms.  Requires n % 4 == 0, ld % 4 == 0: the launcher checks.)
    const int nq = a.n >> 2;                             // float4 per interior row
    for (int rb = r0; rb < r1; rb += 4) {
      const int i = rb + (tid >> 8), p = tid & 255;
      float mx = -INFINITY, sm = 0.f;
      if (i < r1) {
        const float4* vq = reinterpret_cast<const float4*>(vs);
        if (i < a.m) {
          const float4* row = reinterpret_cast<const float4*>(a.S + (long long)i * a.ld);
          // up to five float4 per thread in flight: at n = 5000 the whole row is ONE wait per group of four rows
          for (int j0 = p; j0 < nq; j0 += 5 * 256) {
            float4 sv[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) { const int j4 = j0 + 256 * q; sv[q] = (j4 < nq) ? __ldg(row + j4) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); }
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              const int j4 = j0 + 256 * q;
              if (j4 < nq) {
                const float4 v0 = vq[j4];
                const float t[4] = {sv[q].x + v0.x, sv[q].y + v0.y, sv[q].z + v0.z, sv[q].w + v0.w};
                lse_acc<4>(mx, sm, t);
              }
            }
          }
        } else {                                          // the dustbin row: Z = alpha everywhere
          for (int j4 = p; j4 < nq; j4 += 256) {
            const float4 v0 = vq[j4];
            const float t[4] = {a.alpha + v0.x, a.alpha + v0.y, a.alpha + v0.z, a.alpha + v0.w};
            lse_acc<4>(mx, sm, t);
          }
        }
        if (p == 0) { const float t[1] = {a.alpha + vs[a.n]}; lse_acc<1>(mx, sm, t); }     // the dustbin column
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, sm, o); lse_merge(mx, sm, om, os); }
      if (lane == 0) { rmx[warp] = mx; rsm[warp] = sm; }
      __syncthreads();
      if (p == 0 && i < r1) {
        float m0 = rmx[warp], s0 = rsm[warp];
        for (int w = 1; w < 8; ++w) lse_merge(m0, s0, rmx[warp + w], rsm[warp + w]);
        const float uu = ((i < a.m) ? a.norm : log_mu_bin) - (m0 + logf(s0));
        us[i - r0] = uu; a.u[i] = uu;
      }
      __syncthreads();
    }
    // ---------------- column pass over the rows this CTA owns: a thread takes column PAIRS, 16 rows in flight ----------------
    {
      const int ri1 = min(r1, a.m);                       // interior rows of this CTA
      const bool owns_bin = (r1 == a.m + 1) && (r0 <= a.m);
      for (int jp = tid; jp < (a.n >> 1); jp += 1024) {
        float mx0 = -INFINITY, sm0 = 0.f, mx1 = -INFINITY, sm1 = 0.f;
        const float2* col = reinterpret_cast<const float2*>(a.S + (long long)r0 * a.ld) + jp;
        const int ldp = a.ld >> 1;
        int i = r0;
        for (; i + 8 <= ri1; i += 8) {
          float t0[8], t1[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) { const float2 z = __ldg(col + (long long)(i - r0 + q) * ldp); const float uu = us[i - r0 + q]; t0[q] = z.x + uu; t1[q] = z.y + uu; }
          lse_acc<8>(mx0, sm0, t0); lse_acc<8>(mx1, sm1, t1);
        }
        for (; i < ri1; ++i) {
          const float2 z = __ldg(col + (long long)(i - r0) * ldp); const float uu = us[i - r0];
          const float t0[1] = {z.x + uu}, t1[1] = {z.y + uu};
          lse_acc<1>(mx0, sm0, t0); lse_acc<1>(mx1, sm1, t1);
        }
        if (owns_bin) { const float t[1] = {a.alpha + us[a.m - r0]}; lse_acc<1>(mx0, sm0, t); lse_acc<1>(mx1, sm1, t); }
        float2* pp = a.part + (long long)c * (a.n + 1) + 2 * jp;
        pp[0] = make_float2(mx0, sm0); pp[1] = make_float2(mx1, sm1);
      }
      if (tid == 0) {                                     // the dustbin column: alpha + u over all own rows
        float mx = -INFINITY, sm = 0.f;
        for (int i = r0; i < r1; ++i) { const float t[1] = {a.alpha + us[i - r0]}; lse_acc<1>(mx, sm, t); }
        a.part[(long long)c * (a.n + 1) + a.n] = make_float2(mx, sm);
      }
    }
    sink_grid_sync(bar, gen, a.dbg);
    // ---------------- merge the partials of this CTA's columns (fixed order: lane l takes CTAs l, l+32, ...) ----------------
    for (int j = c0 + warp; j < c1; j += 32) {
      float mx = -INFINITY, sm = 0.f;
      for (int k = lane; k < G; k += 32) { const float2 pr = __ldcg(a.part + (long long)k * (a.n + 1) + j); lse_merge(mx, sm, pr.x, pr.y); }   // written by other SMs: L2 loads
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, sm, o); lse_merge(mx, sm, om, os); }
      if (lane == 0) a.v[j] = ((j < a.n) ? a.norm : log_nu_bin) - (mx + logf(sm));
    }
    sink_grid_sync(bar, gen, a.dbg);
    for (int j = tid; j <= a.n; j += 1024) vs[j] = __ldcg(a.v + j);
    __syncthreads();
  }
  if (a.iters == 0) {                                   // u = v = 0
    for (int i = r0 + tid; i < r1; i += 1024) { us[i - r0] = 0.f; a.u[i] = 0.f; }
    for (int j = c0 + tid; j < c1; j += 1024) a.v[j] = 0.f;
    __syncthreads();
  }
  // ---------------- assignment: row argmax (own rows), column argmax partials, merge, mutual check ----------------
  for (int rb = r0; rb < min(r1, a.m); rb += 4) {
    const int i = rb + (tid >> 8), p = tid & 255;
    float bv = -INFINITY; int bi = 0x7fffffff;
    if (i < min(r1, a.m)) {
      const float* row = a.S + (long long)i * a.ld; const float ui = us[i - r0];
      for (int j = p; j < a.n; j += 256) { const float t = __ldg(row + j) + ui + vs[j] - a.norm; if (t > bv) { bv = t; bi = j; } }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float vo = __shfl_xor_sync(0xffffffffu, bv, o); const int io = __shfl_xor_sync(0xffffffffu, bi, o);
      if (vo > bv || (vo == bv && io < bi)) { bv = vo; bi = io; }
    }
    if (lane == 0) { rmx[warp] = bv; rix[warp] = bi; }
    __syncthreads();
    if (p == 0 && i < min(r1, a.m)) {
      float v0 = rmx[warp]; int i0 = rix[warp];
      for (int w = 1; w < 8; ++w) if (rmx[warp + w] > v0 || (rmx[warp + w] == v0 && rix[warp + w] < i0)) { v0 = rmx[warp + w]; i0 = rix[warp + w]; }
      a.idx0[i] = i0; a.max0[i] = v0;
    }
    __syncthreads();
  }
  for (int j = tid; j < a.n; j += 1024) {               // first maximal row of this CTA's block (rows ascend, strict >)
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = r0; i < min(r1, a.m); ++i) { const float t = __ldg(a.S + (long long)i * a.ld + j) + us[i - r0] + vs[j] - a.norm; if (t > bv) { bv = t; bi = i; } }
    a.part[(long long)c * (a.n + 1) + j] = make_float2(bv, __int_as_float(bi));
  }
  sink_grid_sync(bar, gen);
  for (int j = c0 + warp; j < min(c1, a.n); j += 32) {
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int k = lane; k < G; k += 32) {
      const float2 pr = __ldcg(a.part + (long long)k * (a.n + 1) + j); const int ri = __float_as_int(pr.y);
      if (pr.x > bv || (pr.x == bv && ri < bi)) { bv = pr.x; bi = ri; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float vo = __shfl_xor_sync(0xffffffffu, bv, o); const int io = __shfl_xor_sync(0xffffffffu, bi, o);
      if (vo > bv || (vo == bv && io < bi)) { bv = vo; bi = io; }
    }
    if (lane == 0) a.idx1[j] = bi;
  }
  sink_grid_sync(bar, gen);
  for (int i = r0 + tid; i < min(r1, a.m); i += 1024) {  // matches0 / matching_scores0 (rot_coh_match.py:371-378)
    const int j = a.idx0[i];
    const bool mutual = (__ldcg(a.idx1 + j) == i);
    a.matches0[i] = mutual ? j : -1;
    a.mscores0[i] = mutual ? expf(a.max0[i]) : 0.f;
  }
}

// final assignment (rot_coh_match.py:369-379): row / column argmax of Z + u + v - norm on the inner [m][n] block
__global__ void __launch_bounds__(256) ot_row_argmax_kernel(SinkArgs a, int32_t* __restrict__ idx0, float* __restrict__ max0) {
  __shared__ float sv[8]; __shared__ int si[8];
  const int i = blockIdx.x, tid = threadIdx.x;
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int j = tid; j < a.n; j += 256) {
    const float t = a.S[(long long)i * a.ld + j] + a.u[i] + a.v[j] - a.norm;
    if (t > bv) { bv = t; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float vo = __shfl_xor_sync(0xffffffffu, bv, o); const int io = __shfl_xor_sync(0xffffffffu, bi, o);
    if (vo > bv || (vo == bv && io < bi)) { bv = vo; bi = io; }
  }
  if ((tid & 31) == 0) { sv[tid >> 5] = bv; si[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
    idx0[i] = bi; max0[i] = bv;
  }
}
__global__ void __launch_bounds__(256) ot_col_argmax_kernel(SinkArgs a, int32_t* __restrict__ idx1) {
  __shared__ float sv[8][33]; __shared__ int si[8][33];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float bv = -INFINITY; int bi = 0x7fffffff;
  if (j < a.n)
    for (int i = rl; i < a.m; i += 8) {
      const float t = a.S[(long long)i * a.ld + j] + a.u[i] + a.v[j] - a.norm;
      if (t > bv) { bv = t; bi = i; }
    }
  sv[rl][lane] = bv; si[rl][lane] = bi;
  __syncthreads();
  if (rl == 0 && j < a.n) {
    for (int w = 1; w < 8; ++w) if (sv[w][lane] > bv || (sv[w][lane] == bv && si[w][lane] < bi)) { bv = sv[w][lane]; bi = si[w][lane]; }
    idx1[j] = bi;
  }
}
// matches0 / matching_scores0 (rot_coh_match.py:371-378): mutual0 = (i == indices1[indices0[i]])
__global__ void ot_mutual_kernel(const int32_t* __restrict__ idx0, const float* __restrict__ max0, const int32_t* __restrict__ idx1,
                                 int m, int32_t* __restrict__ matches0, float* __restrict__ mscores0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int j = idx0[i];
  const bool mutual = (idx1[j] == i);
  matches0[i] = mutual ? j : -1;
  mscores0[i] = mutual ? expf(max0[i]) : 0.f;
}

}  // namespace roreg
