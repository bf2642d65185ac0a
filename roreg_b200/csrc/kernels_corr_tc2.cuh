// kernels_corr_tc2.cuh - equivariant correlation (Des2R / R-indicator), corr mode 2: ONE match per pipeline
// item and TWO (or more) co-resident CTAs per SM.
//
// Why (runs 23/24): the bare 7680-byte row gather reaches 7.1 TB/s with three landing stages per SM
// (scripts/gather_bench.cu), yet the mode-1 kernel - one CTA per SM, two matches per item, a single set of
// operand tiles - ran at 2.1 TB/s: its per-item chain  land -> convert -> 12 MMAs -> commit -> next convert
// is serial (the operand tiles are single-buffered; double-buffering them does not fit next to four 30 KB
// landing stages).  Mode 2 halves the item (one match: 15 KB landed, 32 KB of operand tiles, a 64-column
// accumulator) so that a CTA needs 95 KB of shared memory, and lets the hardware interleave the serial chains
// of two CTAs on every SM: while one converts, the other's MMAs and loads are in flight.
//
// Per match  G[h][g] = sum_f X[f][h] * Y[f][g]  (60 x 60, K = 32) as tcgen05.mma kind::tf32 M = 128, N = 64:
// A = the 64-row X tile (rows 60..63 zero; rows 64..127 of the instruction read whatever follows the tile -
// their accumulator lanes are never looked at), B = the 64-row Y tile; 3xTF32 = (hi,hi) + (lo,hi) + (hi,lo).
//
//   warp 0      producer   2 x cp.async.bulk of 7680 B (X row, Y row) per stage, 3 stages
//   warp 1      MMA        3 x 4 tcgen05.mma (K = 8), commit -> tiles_free, mma_done[acc]
//   warps 2-3   convert    hi/lo split + transpose into K-major SWIZZLE_128B operand tiles (thread = h)
//   warps 4-5   epilogue   tcgen05.ld (lanes 0..63) -> smem transpose -> 60 generalised-diagonal sums -> argmax
#pragma once
#include "kernels_corr_tc.cuh"

namespace roreg {

constexpr int C2_STAGES = 3;
constexpr int C2_RAW_BYTES = 2 * CT_RAW_BOX;              // X | Y = 15,360 B per stage
constexpr int C2_OPER_BYTES = 64 * 32 * 4;                // one K-major operand: 64 rows x 32 f = 8 KB
constexpr int C2_TILES_BYTES = 4 * C2_OPER_BYTES;         // Xhi | Xlo | Yhi | Ylo = 32 KB
constexpr int C2_GS_BYTES = 60 * 64 * 4;                  // transposed Gram [60 g][64 h]
constexpr int C2_SMEM_BYTES = C2_STAGES * C2_RAW_BYTES + C2_TILES_BYTES + C2_GS_BYTES + 3600 + 16 + 256 + 1024;   // 98,400 B -> two CTAs per SM
constexpr int C2_THREADS = 192;
// kind::tf32, A/B K-major, D = f32, M = 128, N = 64
constexpr uint32_t C2_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

template <bool TRACE, int PASSES>
__global__ void __launch_bounds__(C2_THREADS, 2) group_corr_tc2_kernel(CorrTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* tiles = smem + C2_STAGES * C2_RAW_BYTES;                                        // 45 KB: 1024-aligned
  float* Gs = reinterpret_cast<float*>(tiles + C2_TILES_BYTES);                            // [60 g][64 h]
  uint8_t* tabs = reinterpret_cast<uint8_t*>(Gs) + C2_GS_BYTES;                            // 3600 B
  float* red_v = reinterpret_cast<float*>(tabs + 3600); int* red_i = reinterpret_cast<int*>(red_v + 1);
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(tabs + 3600 + 16) + 7) & ~uintptr_t(7));
  // barriers: 0..2 raw_full[s], 3..5 raw_free[s], 6 conv_done, 7 tiles_free, 8..9 mma_done[acc], 10..11 acc_free[acc]
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  for (int e = threadIdx.x; e < 3600; e += C2_THREADS) tabs[e] = a.tab[e];
  if (threadIdx.x == 0) {
    for (int s = 0; s < C2_STAGES; ++s) { mbar_init(BAR(0 + s), 1); mbar_init(BAR(3 + s), 64); }
    mbar_init(BAR(6), 64); mbar_init(BAR(7), 1);
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(8 + s), 1); mbar_init(BAR(10 + s), 64); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const long long n_items = (long long)a.B * a.K;
  // timeline instrumentation (TRACE instantiation only): event e of this CTA's it-th live item.  Unpredicated on purpose:
  // a store guarded by (a.trace && blockIdx.x == 0 && it < 256) made ptxas 12.9 guard the R2UR moves of the UMMA
  // descriptors with the same predicate (run 26/27: "out-of-range shared address" in every CTA but the traced one).
#define C2_TRACE(e) do { if (TRACE) a.trace[((size_t)blockIdx.x * 256 + (it < 255u ? it : 255u)) * 12 + (e)] = clock64(); } while (0)
  // every role walks the same item sequence and skips the same items (device-side match counts)
  auto live = [&](long long item, int& p, int& k) -> bool {
    p = (int)(item / a.K); k = (int)(item % a.K);
    return k < (a.n_matches ? a.n_matches[p] : a.K);
  };

  if (warp == 0) {
    // ===================== producer: all 32 lanes resolve the rows of 32 upcoming items, lane 0 issues =====================
    uint32_t it = 0;
    for (long long base = blockIdx.x; base < n_items; base += 32LL * gridDim.x) {
      const long long item = base + (long long)lane * gridDim.x;
      bool ok = false; long long rx = 0, ry = 0;
      if (item < n_items) {
        int p, k; ok = live(item, p, k);
        if (ok) {
          const long long w = (long long)p * a.K + k;
          rx = a.idxX ? a.idxX[w * a.idx_stride] : k;
          ry = a.idxY ? a.idxY[w * a.idx_stride] : k;
          if (a.pair_cloud) { rx += (long long)a.pair_cloud[2 * p + 1] * a.n; ry += (long long)a.pair_cloud[2 * p] * a.n; }
        }
      }
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
      for (int l = 0; l < 32; ++l) {
        if (!((okmask >> l) & 1)) continue;              // uniform across the warp
        const long long x = __shfl_sync(0xffffffffu, rx, l), y = __shfl_sync(0xffffffffu, ry, l);
        if (lane == 0) {
          const int st = it % C2_STAGES; const uint32_t ph = (it / C2_STAGES) & 1;
          mbar_wait(BAR(3 + st), ph ^ 1);                // convert warps have consumed this landing buffer
          C2_TRACE(0);
          uint8_t* sb = smem + st * C2_RAW_BYTES;
          mbar_expect_tx(BAR(0 + st), C2_RAW_BYTES);
          bulk_load_1d(smem_u32(sb), a.X + x * RR_ROW, CT_RAW_BOX, BAR(0 + st));
          bulk_load_1d(smem_u32(sb + CT_RAW_BOX), a.Y + y * RR_ROW, CT_RAW_BOX, BAR(0 + st));
        }
        ++it;
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t tb = smem_u32(tiles);
      const uint32_t xhi = tb, xlo = tb + C2_OPER_BYTES, yhi = tb + 2 * C2_OPER_BYTES, ylo = tb + 3 * C2_OPER_BYTES;
      const uint32_t aop[3] = {xhi, xlo, xhi}, bop[3] = {yhi, yhi, ylo};
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        int p, k; if (!live(item, p, k)) continue;
        const int acc = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(BAR(6), it & 1);                       // operand tiles written and visible to the async proxy
        C2_TRACE(4);
        mbar_wait(BAR(10 + acc), aph ^ 1);               // accumulator drained
        C2_TRACE(5);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll
        for (int c = 0; c < PASSES; ++c)                 // compile-time count: a run-time guard here predicates the descriptor moves (see C2_TRACE)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_tf32(d_tmem, umma_desc_sw128(aop[c] + kk * 32), umma_desc_sw128(bop[c] + kk * 32), C2_IDESC, (c | kk) ? 1u : 0u);
        umma_commit(BAR(7));                             // operand tiles reusable by the convert warps
        umma_commit(BAR(8 + acc));                       // accumulator ready for the epilogue
        C2_TRACE(6);
        ++it;
      }
    }
  } else if (warp < 4) {
    // ===================== convert: split hi/lo and transpose into K-major SW128 operand tiles ============
    // raw[f][h] (h contiguous, pitch 60) -> tile row h, 128 B of f per row, 16-B chunk c = f/4 stored at chunk
    // position c ^ (h % 8)  (the 128-byte swizzle TMA / UMMA use).  Rows h = 60..63 are zeros.
    const int h = threadIdx.x - 64;                      // 0..63
    const bool real = h < RR_G;
    uint32_t it = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      int p, k; if (!live(item, p, k)) continue;
      const int st = it % C2_STAGES; const uint32_t ph = (it / C2_STAGES) & 1;
      mbar_wait(BAR(0 + st), ph);                        // rows landed
      if (h == 0) C2_TRACE(1);
      mbar_wait(BAR(7), (it & 1) ^ 1);                   // the MMAs of the previous item no longer read the operand tiles
      if (h == 0) C2_TRACE(2);
      const float* raw = reinterpret_cast<const float*>(smem + st * C2_RAW_BYTES);
      if (!(a.dbg_skip & 2))
#pragma unroll
      for (int op = 0; op < 2; ++op) {                   // 0: X, 1: Y
        const float* src = raw + op * (CT_RAW_BOX / 4) + (real ? h : 0);
        uint8_t* thi = tiles + (op * 2) * C2_OPER_BYTES + h * 128;
        uint8_t* tlo = thi + C2_OPER_BYTES;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 hv, lv; float x; uint32_t t;
          x = real ? src[(4 * c + 0) * RR_G] : 0.f; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x)); hv.x = __uint_as_float(t); lv.x = x - hv.x;
          x = real ? src[(4 * c + 1) * RR_G] : 0.f; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x)); hv.y = __uint_as_float(t); lv.y = x - hv.y;
          x = real ? src[(4 * c + 2) * RR_G] : 0.f; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x)); hv.z = __uint_as_float(t); lv.z = x - hv.z;
          x = real ? src[(4 * c + 3) * RR_G] : 0.f; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x)); hv.w = __uint_as_float(t); lv.w = x - hv.w;
          const int pos = (c ^ (h & 7)) * 16;
          *reinterpret_cast<float4*>(thi + pos) = hv;
          *reinterpret_cast<float4*>(tlo + pos) = lv;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05
      mbar_arrive(BAR(6));                               // conv_done
      mbar_arrive(BAR(3 + st));                          // landing buffer free for the next copy
      if (h == 0) C2_TRACE(3);
      ++it;
    }
  } else {
    // ===================== epilogue: warps 4,5 own TMEM lanes 0..31 / 32..63 =====================
    const int h = (warp - 4) * 32 + lane;                // Gram row (h) == the 'a' this thread later sums
    // this thread always sums the generalised diagonal a = h: keep its 60 table bytes in registers
    uint32_t trow[15];
#pragma unroll
    for (int w4 = 0; w4 < 15; ++w4) {
      const uint8_t* t = tabs + (h < RR_G ? h : 0) * 60 + 4 * w4;
      trow[w4] = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    }
    uint32_t it = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      int p, k; if (!live(item, p, k)) continue;
      const int acc = it & 1; const uint32_t ph = (it >> 1) & 1;
      mbar_wait(BAR(8 + acc), ph);
      if (h == 0) C2_TRACE(7);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + acc * 64;
      uint32_t r[64];
      RR_TMEM_LD32(r, taddr);
      { uint32_t* r2 = r + 32; RR_TMEM_LD32(r2, taddr + 32); }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(BAR(10 + acc));                        // accumulator free as soon as it sits in registers
      if (h == 0) C2_TRACE(8);
      // transposed store: Gs[g][h]; a warp writes 32 consecutive h -> conflict-free
#pragma unroll
      for (int g = 0; g < 60; ++g) Gs[g * 64 + h] = __uint_as_float(r[g]);
      asm volatile("bar.sync 2, 64;" ::: "memory");
      float c = -INFINITY;
      if (a.dbg_skip & 1) c = Gs[h];
      else if (h < RR_G) {
        float c4[4] = {0.f, 0.f, 0.f, 0.f};              // four independent chains (the sum order differs from g = 0..59 only in rounding)
#pragma unroll
        for (int w4 = 0; w4 < 15; ++w4) {
          const uint32_t tw = trow[w4];
          c4[0] += Gs[(4 * w4 + 0) * 64 + (tw & 0xff)];
          c4[1] += Gs[(4 * w4 + 1) * 64 + ((tw >> 8) & 0xff)];
          c4[2] += Gs[(4 * w4 + 2) * 64 + ((tw >> 16) & 0xff)];
          c4[3] += Gs[(4 * w4 + 3) * 64 + (tw >> 24)];
        }
        c = (c4[0] + c4[1]) + (c4[2] + c4[3]);
      }
      const long long w = (long long)p * a.K + k;
      if (h < RR_G && a.cor_out) a.cor_out[w * RR_G + h] = c;
      float v = c; int ix = (h < RR_G) ? h : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float vo = __shfl_xor_sync(0xffffffffu, v, o);
        const int io = __shfl_xor_sync(0xffffffffu, ix, o);
        if (vo > v || (vo == v && io < ix)) { v = vo; ix = io; }
      }
      if (warp == 5 && lane == 0) { red_v[0] = v; red_i[0] = ix; }          // upper half-row warp (h 32..63) publishes
      asm volatile("bar.sync 2, 64;" ::: "memory");
      if (warp == 4 && lane == 0 && a.argmax_out) {
        int best = ix;                                                      // the lower warp holds the smaller indices
        if (red_v[0] > v) best = red_i[0];
        a.argmax_out[w] = best;
      }
      asm volatile("bar.sync 2, 64;" ::: "memory");                         // Gs / red reusable
      if (h == 0) C2_TRACE(9);
      ++it;
    }
  }
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
}

static inline int group_corr_tc2_launch(roreg_ctx* c, const float* X, const float* Y, CorrTcArgs a, cudaStream_t st) {
  RR_ARG(c, (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  a.X = X; a.Y = Y;
  static unsigned long long attr_mask = 0;
  static int ctas_per_sm = 2;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(group_corr_tc2_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES));
    if (const char* e = getenv("ROREG_DEBUG_CORR_CTAS")) { const int v = atoi(e); if (v >= 1 && v <= 2) ctas_per_sm = v; }
  }
  const long long items = (long long)a.B * a.K;
  const long long cap = (long long)c->sm_count * ctas_per_sm;
  const int grid = (int)(items < cap ? items : cap);
  const char* trace_fn = getenv("ROREG_DEBUG_CORR_TRACE");
  static bool traced = false;
  if (trace_fn && !traced && items >= 100000) {          // one-off timeline dump of CTA 0 (debug only; synchronises)
    traced = true;
    const size_t nb = (size_t)grid * 256 * 12 * sizeof(long long);
    RR_CUDA(c, cudaFuncSetAttribute(group_corr_tc2_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES));
    RR_CUDA(c, cudaMalloc(&a.trace, nb));
    RR_CUDA(c, cudaMemsetAsync(a.trace, 0, nb, st));
    group_corr_tc2_kernel<true, 3><<<grid, C2_THREADS, C2_SMEM_BYTES, st>>>(a);
    RR_LAUNCH_CHECK(c);
    RR_CUDA(c, cudaStreamSynchronize(st));
    long long* h = (long long*)malloc(256 * 12 * sizeof(long long));
    RR_CUDA(c, cudaMemcpy(h, a.trace, 256 * 12 * sizeof(long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_fn, "w")) {
      fprintf(f, "# it P_issue C_land C_tilesfree C_done M_convdone M_accfree M_committed E_mmadone E_loaded E_end (clock64 - first)\n");
      long long t0 = h[0];
      for (int i = 0; i < 255; ++i) {
        fprintf(f, "%d", i);
        for (int e = 0; e < 10; ++e) fprintf(f, " %lld", h[i * 12 + e] ? h[i * 12 + e] - t0 : -1);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    free(h); cudaFree(a.trace);
    return ROREG_OK;
  }
  a.trace = nullptr;
  if (a.dbg_passes == 1) {                               // bottleneck experiments only (ROREG_DEBUG_CORR_PASSES=1)
    RR_CUDA(c, cudaFuncSetAttribute(group_corr_tc2_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES));
    group_corr_tc2_kernel<false, 1><<<grid, C2_THREADS, C2_SMEM_BYTES, st>>>(a);
  } else
    group_corr_tc2_kernel<false, 3><<<grid, C2_THREADS, C2_SMEM_BYTES, st>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

}  // namespace roreg
