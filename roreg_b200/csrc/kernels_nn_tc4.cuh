// kernels_nn_tc4.cuh - mutual-NN search, nn mode 4: ONE Gram per pair, both directions from the same accumulators,
// norms folded into the contraction, epilogue free of shared-memory traffic.
//
// What runs 12/13/23 established about modes 1-3:
//   * reading accumulators out of TMEM is not the limit (818 B/clk/SM with 8 warps, scripts/tmem_bench.cu);
//   * a K = 8 tf32 MMA with both operands in shared memory reads 8 KB per 64 clk = the SM's whole shared-memory
//     bandwidth, so every LDS the epilogue issues (column norms) competes with the tensor pipe (mode-2 inner loop:
//     430 -> 814 clk per tile once MMAs run beside it);
//   * modes 1/2 compute the Gram twice (once per direction): 2 x 1600 tiles of 832 tensor-pipe clocks per pair and
//     3.3 GB of operand tiles streamed L2 -> SM per 32 pairs.
// Mode 4 therefore
//   * folds -|a|^2/2 - |b|^2/2 into the MMA as a 13th K-step (extended operand columns [-n_hi, -n_lo, 1, 1] x
//     [1, 1, -n_hi, -n_lo]), so an accumulator element IS -d^2/2 and the epilogue is compares on registers only;
//   * makes a CTA own a 128-column block of cloud 1 (stationary B tile) and sweep all row tiles of cloud 0 through
//     it: the column direction (nn10) is a per-thread running maximum over the sweep, kept in registers
//     (cmax[64] + packed tile indices) and reduced across lanes once per item; the row direction (nn01) is the
//     per-tile maximum of the 64 registers a thread holds, published as a packed (value, column) key per
//     (row, column block) and reduced by a small finishing kernel;
//   * reads operands from a tile-major, pre-swizzled image written by the prep kernel (exactly the bytes of the
//     SWIZZLE_128B K-major shared-memory layout), so a tile is ONE contiguous cp.async.bulk - no tensor maps.
//
//   warp 0      producer   48 KB bulk copies: stationary tile once per item, streaming tiles through 3 stages
//   warp 1      MMA        13 x tcgen05.mma kind::tf32 (M = N = 128, K = 8) per tile into 4 TMEM accumulators
//   warps 2-9   epilogue   tcgen05.ld 32x32b.x32 x2 -> column running max / row max + first index
#pragma once
#include "kernels_nn_tc.cuh"

namespace roreg {

constexpr int T4_TILE_BYTES = 3 * TC_BOX_BYTES;                  // [hi | lo | ext] = 48 KB per 128-row tile
constexpr int T4_STAGES = 3;
constexpr int T4_ACC = 4;                                        // TMEM accumulators (4 x 128 columns = all 512)
constexpr int T4_THREADS = 64 + 256;
constexpr int T4_SMEM_BYTES = T4_TILE_BYTES * (1 + T4_STAGES) + 128 * 8 /*row merge*/ + 2 * 4 * 64 * 8 /*column merge*/ + 1024 + 256;
constexpr float T4_PAD_NORM = 1e30f;                             // half-norm of a padding row: its distances lose every comparison

// element (r, k) of a [128 x 32 f32] K-major SWIZZLE_128B box: 16-byte chunk k/4 of row r sits at chunk (k/4) ^ (r % 8)
__device__ __forceinline__ int t4_sw128(int r, int k) { return r * 128 + ((((k >> 2) ^ (r & 7))) << 4) + ((k & 3) << 2); }

// ---- prep: pooled features [B*2][S][32] -> tile images [B*2][NT][48 KB] ---------------------------------------
__global__ void __launch_bounds__(256) nn_tc4_prep_kernel(const float* __restrict__ inv, int S, int NT, long long total_rows,
                                                          uint8_t* __restrict__ img) {
  const long long gr = blockIdx.x * 8LL + (threadIdx.x >> 5);    // padded row number over all (pair, side)
  const int lane = threadIdx.x & 31;
  if (gr >= total_rows) return;
  const int Sp = NT * TC_BM;
  const long long ps = gr / Sp; const int rr = (int)(gr % Sp);   // (pair, side), row within the padded cloud
  const int t = rr >> 7, r = rr & 127;
  uint8_t* tile = img + (ps * NT + t) * (long long)T4_TILE_BYTES;
  float x = 0.f;
  if (rr < S) x = inv[(ps * S + rr) * 32 + lane];
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
  const float hi = __uint_as_float(hb);
  const float lo = x - hi;
  const float nh = (rr < S) ? 0.5f * warp_sum(x * x) : T4_PAD_NORM;
  uint32_t nb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(nb) : "f"(nh));
  const float n_hi = __uint_as_float(nb);
  uint32_t lb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(nh - n_hi));
  const float n_lo = __uint_as_float(lb);
  // extended columns: 0..7 = row-role vector (A operand), 8..15 = column-role vector (B operand), 16..31 unused
  float e = 0.f;
  if (lane == 0 || lane == 10) e = -n_hi;
  else if (lane == 1 || lane == 11) e = -n_lo;
  else if (lane == 2 || lane == 3 || lane == 8 || lane == 9) e = 1.f;
  const int off = t4_sw128(r, lane);
  *reinterpret_cast<float*>(tile + off) = hi;
  *reinterpret_cast<float*>(tile + TC_BOX_BYTES + off) = lo;
  *reinterpret_cast<float*>(tile + 2 * TC_BOX_BYTES + off) = e;
}

struct NNTc4Args {
  const uint8_t* img;                  // [B*2][NT] tile images
  int S, NT, B;
  unsigned long long* rowpart;         // [B][NT column blocks][NT*128 rows]  packed (ordered value << 32 | ~column)
  int32_t* nn10;                       // [B][S]
};

__device__ __forceinline__ uint32_t t4_ord(float v) {            // monotone float -> uint
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void t4_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int PASSES>
__global__ void __launch_bounds__(T4_THREADS, 1) nn_tc4_kernel(NNTc4Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                   // stationary tile (cloud 1 block), 48 KB
  uint8_t* sA = smem + T4_TILE_BYTES;                   // T4_STAGES x 48 KB streaming tiles (cloud 0)
  unsigned long long* mrg = reinterpret_cast<unsigned long long*>(smem + T4_TILE_BYTES * (1 + T4_STAGES));   // [128] row keys of the upper column half
  unsigned long long* cmg = mrg + 128;                  // [2 halves][4 quadrants][64 columns] column candidates
  uint64_t* bars = reinterpret_cast<uint64_t*>(cmg + 2 * 4 * 64);
  // barriers: 0 b_full, 1 b_empty, 2..4 a_full, 5..7 a_empty, 8..11 tmem_full, 12..15 tmem_empty
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1); mbar_init(BAR(1), 1);
    for (int s = 0; s < T4_STAGES; ++s) { mbar_init(BAR(2 + s), 1); mbar_init(BAR(5 + s), 1); }
    for (int s = 0; s < T4_ACC; ++s) { mbar_init(BAR(8 + s), 1); mbar_init(BAR(12 + s), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const int NT = a.NT;
  const int n_items = a.B * NT;                         // item = (pair, 128-column block of cloud 1)

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it_a = 0, b_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int j = item % NT, p = item / NT;
        const uint8_t* imgB = a.img + ((long long)(p * 2 + 1) * NT + j) * T4_TILE_BYTES;
        const uint8_t* imgA = a.img + (long long)(p * 2) * NT * T4_TILE_BYTES;
        mbar_wait(BAR(1), b_phase ^ 1);                 // the previous item's MMAs no longer read the stationary tile
        mbar_expect_tx(BAR(0), T4_TILE_BYTES);
        t4_bulk(smem_u32(sB), imgB, T4_TILE_BYTES, BAR(0));
        b_phase ^= 1;
        for (int t = 0; t < NT; ++t, ++it_a) {
          const int st = it_a % T4_STAGES; const uint32_t ph = (it_a / T4_STAGES) & 1;
          mbar_wait(BAR(5 + st), ph ^ 1);
          mbar_expect_tx(BAR(2 + st), T4_TILE_BYTES);
          t4_bulk(smem_u32(sA + st * T4_TILE_BYTES), imgA + (long long)t * T4_TILE_BYTES, T4_TILE_BYTES, BAR(2 + st));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it_a = 0, it_t = 0, b_phase = 0;
      const uint32_t bhi = smem_u32(sB), blo = bhi + TC_BOX_BYTES, bex = bhi + 2 * TC_BOX_BYTES + 32;   // column-role ext = logical columns 8..15
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(BAR(0), b_phase); b_phase ^= 1;
        for (int t = 0; t < NT; ++t, ++it_a, ++it_t) {
          const int st = it_a % T4_STAGES; const uint32_t ph = (it_a / T4_STAGES) & 1;
          const int acc = it_t % T4_ACC; const uint32_t tph = (it_t / T4_ACC) & 1;
          mbar_wait(BAR(2 + st), ph);
          mbar_wait(BAR(12 + acc), tph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ahi = smem_u32(sA + st * T4_TILE_BYTES), alo = ahi + TC_BOX_BYTES, aex = ahi + 2 * TC_BOX_BYTES;
          const uint32_t d_tmem = tmem_base + acc * TC_BN;
          // the norm step first: it is the one that must never be dropped (passes < 3 is a debug knob)
          umma_tf32(d_tmem, umma_desc_sw128(aex), umma_desc_sw128(bex), TC_IDESC, 0u);
          const uint32_t aop[3] = {ahi, alo, ahi}, bop[3] = {bhi, bhi, blo};
#pragma unroll
          for (int c = 0; c < PASSES; ++c)                // compile-time count: no run-time predicate near the descriptor moves
#pragma unroll
            for (int kk = 0; kk < TC_KC / 8; ++kk)
              umma_tf32(d_tmem, umma_desc_sw128(aop[c] + kk * 32), umma_desc_sw128(bop[c] + kk * 32), TC_IDESC, 1u);
          umma_commit(BAR(5 + st));
          umma_commit(BAR(8 + acc));
        }
        umma_commit(BAR(1));
      }
    }
  } else {
    // ===================== epilogue: 8 warps, warp -> (lane quadrant q, column half hf) =====================
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;                     // warps 2..5 -> columns 0..63, warps 6..9 -> 64..127
    const int row_in_tile = q * 32 + lane;
    const int Sp = NT * TC_BM;
    uint32_t it_t = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int j = item % NT, p = item / NT;
      float cmax[64]; uint32_t ctile[16];
#pragma unroll
      for (int e = 0; e < 64; ++e) cmax[e] = -INFINITY;
#pragma unroll
      for (int e = 0; e < 16; ++e) ctile[e] = 0;
      unsigned long long* rp = a.rowpart + ((long long)p * NT + j) * Sp;
      for (int t = 0; t < NT; ++t, ++it_t) {
        const int acc = it_t % T4_ACC; const uint32_t tph = (it_t / T4_ACC) & 1;
        mbar_wait(BAR(8 + acc), tph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TC_BN + hf * 64;
        uint32_t r[64];
        RR_TMEM_LD32(r, taddr);
        { uint32_t* r2 = r + 32; RR_TMEM_LD32(r2, taddr + 32); }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(BAR(12 + acc));                     // accumulator free: the MMA of tile t+4 may start
        // column direction: running maximum of -d^2/2 over the rows this lane sees; strict '>' + increasing t = first row wins ties
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 64; ++e) {
          const float v = __uint_as_float(r[e]);
          if (v > cmax[e]) ctile[e >> 2] = __byte_perm(ctile[e >> 2], (uint32_t)t, (0x3210 & ~(0xF << (4 * (e & 3)))) | (4 << (4 * (e & 3))));
          cmax[e] = fmaxf(cmax[e], v);
          if (e & 1) m1 = fmaxf(m1, v); else m0 = fmaxf(m0, v);
        }
        const float m = fmaxf(m0, m1);
        // row direction: first column attaining the tile-row maximum
        int j4[4] = {64, 64, 64, 64};                   // four independent select chains (e mod 4), merged by min
#pragma unroll
        for (int e = 63; e >= 0; --e) if (__uint_as_float(r[e]) == m) j4[e & 3] = e;
        const int jj = min(min(j4[0], j4[1]), min(j4[2], j4[3]));
        const unsigned long long key = ((unsigned long long)t4_ord(m) << 32) | (0xffffffffu - (uint32_t)(j * TC_BN + hf * 64 + jj));
        if (hf == 1) mrg[row_in_tile] = key;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        if (hf == 0) {
          const unsigned long long o = mrg[row_in_tile];
          rp[t * TC_BM + row_in_tile] = o > key ? o : key;       // equal values: the larger key is the smaller column
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
      }
      // ---- end of the sweep: reduce the column candidates over the 128 lanes -------------------------------
      // lexicographic (largest value, smallest row); row = tile * 128 + row_in_tile
#pragma unroll
      for (int e = 0; e < 64; ++e) {
        float v = cmax[e];
        uint32_t row = ((ctile[e >> 2] >> (8 * (e & 3))) & 0xff) * TC_BM + row_in_tile;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float vo = __shfl_xor_sync(0xffffffffu, v, o);
          const uint32_t ro = __shfl_xor_sync(0xffffffffu, row, o);
          if (vo > v || (vo == v && ro < row)) { v = vo; row = ro; }
        }
        if (lane == (e & 31)) cmg[(hf * 4 + q) * 64 + e] = ((unsigned long long)t4_ord(v) << 32) | (0xffffffffu - row);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      {
        const int et = threadIdx.x - 64;                // 0..255
        if (et < 128) {
          const int h2 = et >> 6, e = et & 63;
          unsigned long long best = 0;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) { const unsigned long long c = cmg[(h2 * 4 + qq) * 64 + e]; best = c > best ? c : best; }
          const int col = j * TC_BN + et;
          if (col < a.S) a.nn10[(long long)p * a.S + col] = (int32_t)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// ---- finish: nn01[p][row] = column of the best key over the NT column blocks ---------------------------------
__global__ void __launch_bounds__(256) nn_tc4_finish_kernel(const unsigned long long* __restrict__ rowpart, int S, int NT, int B,
                                                            int32_t* __restrict__ nn01) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= (long long)B * S) return;
  const int p = (int)(i / S), row = (int)(i % S);
  const long long Sp = (long long)NT * TC_BM;
  const unsigned long long* rp = rowpart + (long long)p * NT * Sp + row;
  unsigned long long best = 0;
  for (int j = 0; j < NT; ++j) { const unsigned long long c = rp[j * Sp]; best = c > best ? c : best; }
  nn01[i] = (int32_t)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
}

static inline size_t nn_tc4_workspace_bytes(int B, int S) {
  const size_t NT = (size_t)(S + TC_BM - 1) / TC_BM;
  return rr_align((size_t)B * 2 * NT * T4_TILE_BYTES) + rr_align(sizeof(unsigned long long) * (size_t)B * NT * NT * TC_BM) + 2048;
}

// inv: [B][2][S][32] pooled features; img / rowpart: workspace of nn_tc4_workspace_bytes(B, S)
static inline int nn_tc4_launch_both(roreg_ctx* c, const float* inv, int S, int B, uint8_t* img, unsigned long long* rowpart,
                                     int32_t* nn01, int32_t* nn10, cudaStream_t st) {
  const int NT = (S + TC_BM - 1) / TC_BM;
  RR_ARG(c, NT <= 256);                                  // tile indices of the column direction are kept as bytes
  RR_ARG(c, (reinterpret_cast<uintptr_t>(img) & 1023) == 0);
  const long long total_rows = (long long)B * 2 * NT * TC_BM;
  nn_tc4_prep_kernel<<<(unsigned)((total_rows + 7) / 8), 256, 0, st>>>(inv, S, NT, total_rows, img);
  RR_LAUNCH_CHECK(c);
  static int passes = 0;
  if (!passes) {
    RR_CUDA(c, cudaFuncSetAttribute(nn_tc4_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, T4_SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(nn_tc4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, T4_SMEM_BYTES));
    passes = 3;
    if (const char* e = getenv("ROREG_DEBUG_NN_PASSES")) if (atoi(e) == 1) passes = 1;      // bottleneck experiments only
  }
  const int items = B * NT;
  const int grid = items < c->sm_count ? items : c->sm_count;
  NNTc4Args a{img, S, NT, B, rowpart, nn10};
  if (passes == 1) nn_tc4_kernel<1><<<grid, T4_THREADS, T4_SMEM_BYTES, st>>>(a);
  else nn_tc4_kernel<3><<<grid, T4_THREADS, T4_SMEM_BYTES, st>>>(a);
  RR_LAUNCH_CHECK(c);
  const long long n = (long long)B * S;
  nn_tc4_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rowpart, S, NT, B, nn01);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

}  // namespace roreg
