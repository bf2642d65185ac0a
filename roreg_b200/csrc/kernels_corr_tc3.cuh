// kernels_corr_tc3.cuh - equivariant correlation (Des2R / R-indicator), corr mode 3: the shared-memory-lean pipeline.
//
// What runs 23-28 established about modes 1/2 (profiles/r01_run26_corr2_timeline.txt): the bare 7680-byte row gather
// reaches 7.1 TB/s, but the kernels ran at 2.1 TB/s because they are SHARED-MEMORY-bandwidth bound: per match the
// landed rows are written (15 KB) and re-read (15 KB) by the convert stage, the 3xTF32 operand tiles are written
// (32 KB) and read by 12 tf32 MMAs (48 KB), and the 60x60 Gram goes through shared memory once more (30 KB) for the
// generalised-diagonal sums: ~140-165 KB per match against 128 B/clk.  Mode 3 removes what can be removed:
//   * no landing buffers: loader warps LDG a descriptor row straight into registers (15 x 16 B per lane, 12 warps =
//     90 KB in flight per SM) and write the operands from there;
//   * float16 two-accumulator arithmetic (see kernels_nn_tc4.cuh): x = hi + 2^-11 lo', D1 = Xhi.Yhi,
//     D2 = Xlo'.Yhi + Xhi.Ylo', G = D1 + 2^-11 D2 - float32-class products, but the operand tiles are half the bytes
//     (16 KB per match) and the contraction is 6 kind::f16 MMAs per TWO matches (24 KB of operand reads per match);
//   * MN-major operands: a descriptor row [32 f][60 h] is already "h contiguous", so a lane converts its float4
//     (4 consecutive h of one f) and stores 8 bytes - no transpose, conflict-free.  Canonical layout (CUTLASS
//     cute/atom/mma_traits_sm100.hpp, Major-MN / SWIZZLE_128B, in 16-byte units ((8,n),(8,k)):((1,LBO),(8,SBO))):
//     8 f-rows x 128 B (64 h) atoms, chunk16 ^= f % 8, SBO = 1024 B between 8-f groups, LBO = 4096 B between the two
//     matches stacked along M (resp. N).
// Shared-memory traffic per match: 16 KB (operand writes) + 24 KB (MMA reads) + 30 KB (Gram transpose) = 70 KB.
//
//   warps 0-11   loaders    3 groups (one tile set each) x {X m0, X m1, Y m0, Y m1}: LDG row -> regs -> fp16 hi / lo' -> STS.64
//   warps 12-15  epilogue   tcgen05.ld D1, D2 -> FFMA combine -> smem transpose -> 60 generalised-diagonal sums -> argmax
//   warp 16      MMA        6 x tcgen05.mma kind::f16 (M = N = 128, K = 16, A/B MN-major), 2 x (D1 | D2) in TMEM
#pragma once
#include <cuda_fp16.h>
#include "kernels_corr_tc.cuh"
#include "kernels_nn_tc4.cuh"

namespace roreg {

constexpr int C3_GROUPS = 3;
constexpr int C3_THREADS = (4 * C3_GROUPS + 4 + 1) * 32;  // 544
constexpr int C3_OP_BYTES = 32 * 128;                     // one match's operand tile: 32 f-rows x 64 h fp16 = 4 KB
constexpr int C3_BUF_BYTES = 8 * C3_OP_BYTES;             // Xhi m0|m1, Xlo m0|m1, Yhi m0|m1, Ylo m0|m1 = 32 KB
constexpr int C3_GS_BYTES = 2 * 60 * 64 * 4;              // transposed Gram of both matches [2][60 g][64 h]
constexpr int C3_SMEM_BYTES = C3_GROUPS * C3_BUF_BYTES + C3_GS_BYTES + 3600 + 16 + 256 + 1024;   // 134 KB
// kind::f16, A and B MN-major, D = f32, M = N = 128
constexpr uint32_t C3_IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
// MN-major SWIZZLE_128B descriptor: LBO = 4096 B, SBO = 1024 B, version 1, layout type 2
__device__ __forceinline__ uint64_t c3_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// spin without reading the clock on the success path; a lost arrive still ends in a trap, never a hung GPU
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 26)) {
      printf("roreg: mbarrier wait timed out: block %d warp %d lane %d barrier@%u parity %u\n", blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31, bar, parity);
      __trap();
    }
  }
}
// tcgen05.mma issued by one elected lane of a converged warp (no divergent region around the instruction)
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p, e;\nelect.sync _|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(bar) : "memory");
}

template <bool TRACE>
__global__ void __launch_bounds__(C3_THREADS, 1) group_corr_tc3_kernel(CorrTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bufs = smem;                                                                    // one 32 KB operand-tile set per loader group
  float* Gs = reinterpret_cast<float*>(smem + C3_GROUPS * C3_BUF_BYTES);                           // [2][60 g][64 h]
  uint8_t* tabs = reinterpret_cast<uint8_t*>(Gs) + C3_GS_BYTES;                            // 3600 B
  float* red_v = reinterpret_cast<float*>(tabs + 3600); int* red_i = reinterpret_cast<int*>(red_v + 2);   // [2] each
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(tabs + 3600 + 16) + 7) & ~uintptr_t(7));
  // barriers: 0..2 conv_done[grp], 3..5 tiles_free[grp], 6..7 mma_done[acc], 8..9 acc_free[acc].
  // A loader group owns one tile set: every waiter then only ever distinguishes ADJACENT phases of a barrier (run 34: with
  // three groups sharing two tile sets a group could run two phases ahead of tiles_free and pass the parity test early).
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  for (int e = threadIdx.x; e < 3600; e += C3_THREADS) tabs[e] = a.tab[e];
  // operand tiles start as zeros: the loaders never write the padding columns h = 60..63
  for (int e = threadIdx.x; e < C3_GROUPS * C3_BUF_BYTES / 16; e += C3_THREADS) reinterpret_cast<uint4*>(bufs)[e] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < C3_GROUPS; ++s) { mbar_init(BAR(0 + s), 4); mbar_init(BAR(3 + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(6 + s), 1); mbar_init(BAR(8 + s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // timeline instrumentation (TRACE instantiation only, unpredicated stores - see kernels_corr_tc2.cuh): event e of this CTA's i-th live item
#define C3_TRACE(i, e) do { if (TRACE) a.trace[((size_t)blockIdx.x * 256 + ((i) < 255u ? (i) : 255u)) * 12 + (e)] = clock64(); } while (0)
  const int items_per_pair = (a.K + 1) / 2;
  const long long n_items = (long long)a.B * items_per_pair;
  // every role walks the same item sequence and skips the same items (device-side match counts)
  auto item_count = [&](long long item, int& p, int& k0) -> int {
    p = (int)(item / items_per_pair); k0 = (int)(item % items_per_pair) * 2;
    const int cnt = a.n_matches ? a.n_matches[p] : a.K;
    return cnt - k0;                                   // <= 0: nothing, 1: one match, >= 2: two matches
  };

  if (warp < 4 * C3_GROUPS) {
    // ===================== loaders =====================
    const int grp = warp >> 2, role = warp & 3;        // role: 0 X m0, 1 X m1, 2 Y m0, 3 Y m1
    const int isY = role >> 1, m = role & 1;
    // byte offset of this lane's j-th float4 (flat element e = (32 j + lane) * 4 -> f = e / 60, h = e % 60) inside a tile
    int off[15];
#pragma unroll
    for (int j = 0; j < 15; ++j) {
      const int e = (j * 32 + lane) * 4, f = e / 60, h = e % 60;
      off[j] = (f >> 3) * 1024 + (f & 7) * 128 + ((((h * 2) >> 4) ^ (f & 7)) << 4) + ((h * 2) & 15);
    }
    const float* base = isY ? a.Y : a.X;
    const int32_t* idx = isY ? a.idxY : a.idxX;
    // row of this warp's (side, slot) for an item; the odd tail's second slot re-reads the first match
    auto row_ptr = [&](int p, int k0, int avail) -> const float4* {
      const int k = k0 + ((m < avail) ? m : 0);
      const long long w = (long long)p * a.K + k;
      long long r = idx ? idx[w * a.idx_stride] : k;
      if (a.pair_cloud) r += (long long)a.pair_cloud[2 * p + (isY ? 0 : 1)] * a.n;
      return reinterpret_cast<const float4*>(base + r * RR_ROW);
    };
    // walk to this group's next live item
    long long item = blockIdx.x; uint32_t live = 0;    // `live` = index of the next live item of this CTA
    const float4* next = nullptr; uint32_t next_it = 0;
    auto advance = [&]() {
      next = nullptr;
      for (; item < n_items; item += gridDim.x) {
        int p, k0; const int avail = item_count(item, p, k0);
        if (avail <= 0) continue;
        const uint32_t my = live++;
        if ((int)(my % C3_GROUPS) == grp) { next = row_ptr(p, k0, avail); next_it = my; item += gridDim.x; return; }
      }
    };
    advance();
    while (next) {
      const float4* src = next; const uint32_t it = next_it;
      float4 v[15];
#pragma unroll
      for (int j = 0; j < 15; ++j) v[j] = ldg_stream4(src + j * 32 + lane);
      if (role == 0 && lane == 0) C3_TRACE(it, 0);
      advance();                                       // resolve the next row while this one is in flight
      const uint32_t use = it / C3_GROUPS;             // how often this group's tile set has been filled before
      if (role == 0 && lane == 0) C3_TRACE(it, 1);
      mbar_wait_lean(BAR(3 + grp), (use & 1) ^ 1);     // the MMAs of this group's previous item no longer read the tile set
      if (role == 0 && lane == 0) C3_TRACE(it, 2);
      uint8_t* thi = bufs + grp * C3_BUF_BYTES + (isY * 4 + m) * C3_OP_BYTES;
      uint8_t* tlo = thi + 2 * C3_OP_BYTES;
#pragma unroll
      for (int j = 0; j < 15; ++j) {
        const __half2 h01 = __floats2half2_rn(v[j].x, v[j].y), h23 = __floats2half2_rn(v[j].z, v[j].w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((v[j].x - f01.x) * T4_LO_SCALE, (v[j].y - f01.y) * T4_LO_SCALE);
        const __half2 l23 = __floats2half2_rn((v[j].z - f23.x) * T4_LO_SCALE, (v[j].w - f23.y) * T4_LO_SCALE);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(thi + off[j]) = hv;
        *reinterpret_cast<uint2*>(tlo + off[j]) = lv;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(0 + grp));        // conv_done: 4 warps per item
      if (role == 0 && lane == 0) C3_TRACE(it, 3);
    }
  } else if (warp == 16) {
    // ===================== MMA issuer: the whole warp walks the loop, one elected lane issues =====================
    uint32_t it = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      int p, k0; if (item_count(item, p, k0) <= 0) continue;
      const int g = it % C3_GROUPS, acc = it & 1; const uint32_t gph = (it / C3_GROUPS) & 1, aph = (it >> 1) & 1;
      mbar_wait_lean(BAR(0 + g), gph);                 // operand tiles written and visible to the async proxy
      if (lane == 0) C3_TRACE(it, 4);
      mbar_wait_lean(BAR(8 + acc), aph ^ 1);           // accumulators drained
      if (lane == 0) C3_TRACE(it, 5);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tb = smem_u32(bufs + g * C3_BUF_BYTES);
      const uint32_t xhi = tb, xlo = tb + 2 * C3_OP_BYTES, yhi = tb + 4 * C3_OP_BYTES, ylo = tb + 6 * C3_OP_BYTES;
      const uint32_t d1 = tmem_base + acc * 256, d2 = d1 + 128;
      umma_f16_elect(d1, c3_desc(xhi), c3_desc(yhi), C3_IDESC, 0u);                 // hi.hi, f = 0..15
      umma_f16_elect(d1, c3_desc(xhi + 2048), c3_desc(yhi + 2048), C3_IDESC, 1u);   //        f = 16..31
      umma_f16_elect(d2, c3_desc(xlo), c3_desc(yhi), C3_IDESC, 0u);                 // lo'.hi
      umma_f16_elect(d2, c3_desc(xlo + 2048), c3_desc(yhi + 2048), C3_IDESC, 1u);
      umma_f16_elect(d2, c3_desc(xhi), c3_desc(ylo), C3_IDESC, 1u);                 // hi.lo'
      umma_f16_elect(d2, c3_desc(xhi + 2048), c3_desc(ylo + 2048), C3_IDESC, 1u);
      umma_commit_elect(BAR(3 + g));                   // operand tiles reusable by their loader group
      umma_commit_elect(BAR(6 + acc));                 // accumulators ready for the epilogue
      if (lane == 0) C3_TRACE(it, 6);
      ++it;
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                            // TMEM lane quadrant of this warp (warps 12..15 -> 0..3)
    const int m = q >> 1;                              // match slot: lanes 0..63 -> 0, 64..127 -> 1
    const int h = (q & 1) * 32 + lane;                 // Gram row (h) == the 'a' this thread later sums
    float* G = Gs + m * 60 * 64;
    // this thread always sums the generalised diagonal a = h: keep its 60 table bytes in registers
    uint32_t trow[15];
#pragma unroll
    for (int w4 = 0; w4 < 15; ++w4) {
      const uint8_t* t = tabs + (h < RR_G ? h : 0) * 60 + 4 * w4;
      trow[w4] = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    }
    uint32_t it = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      int p, k0; const int avail = item_count(item, p, k0);
      if (avail <= 0) continue;
      const int buf = it & 1; const uint32_t ph = (it >> 1) & 1;
      mbar_wait_lean(BAR(6 + buf), ph);
      if (warp == 12 && lane == 0) C3_TRACE(it, 7);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + m * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r1[32], r2[32];
        RR_TMEM_LD32(r1, taddr + half * 32);
        RR_TMEM_LD32(r2, taddr + 128 + half * 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (half == 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(8 + buf));    // accumulators free as soon as they sit in registers
          if (warp == 12 && lane == 0) C3_TRACE(it, 8);
        }
        // transposed store: Gs[m][g][h]; a warp writes 32 consecutive h -> conflict-free
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const int g = half * 32 + u;
          if (g < 60) G[g * 64 + h] = fmaf(__uint_as_float(r2[u]), T4_LO_UNSCALE, __uint_as_float(r1[u]));
        }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + m) : "memory");
      float c = -INFINITY;
      if (h < RR_G) {
        float c4[4] = {0.f, 0.f, 0.f, 0.f};            // four independent chains (the sum order differs from g = 0..59 only in rounding)
#pragma unroll
        for (int w4 = 0; w4 < 15; ++w4) {
          const uint32_t tw = trow[w4];
          c4[0] += G[(4 * w4 + 0) * 64 + (tw & 0xff)];
          c4[1] += G[(4 * w4 + 1) * 64 + ((tw >> 8) & 0xff)];
          c4[2] += G[(4 * w4 + 2) * 64 + ((tw >> 16) & 0xff)];
          c4[3] += G[(4 * w4 + 3) * 64 + (tw >> 24)];
        }
        c = (c4[0] + c4[1]) + (c4[2] + c4[3]);
      }
      const bool valid = m < avail;
      const long long w = (long long)p * a.K + k0 + m;
      if (valid && h < RR_G && a.cor_out) a.cor_out[w * RR_G + h] = c;
      float v = c; int ix = (h < RR_G) ? h : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float vo = __shfl_xor_sync(0xffffffffu, v, o);
        const int io = __shfl_xor_sync(0xffffffffu, ix, o);
        if (vo > v || (vo == v && io < ix)) { v = vo; ix = io; }
      }
      if ((q & 1) == 1 && lane == 0) { red_v[m] = v; red_i[m] = ix; }       // upper half-row warp (h 32..63) publishes
      asm volatile("bar.sync %0, 64;" ::"r"(2 + m) : "memory");
      if ((q & 1) == 0 && lane == 0 && valid && a.argmax_out) {
        int best = ix;                                                      // lower warp holds the smaller indices
        if (red_v[m] > v) best = red_i[m];
        a.argmax_out[w] = best;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + m) : "memory");             // Gs / red reusable
      if (warp == 12 && lane == 0) C3_TRACE(it, 9);
      ++it;
    }
  }
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// X, Y: descriptor arrays [rows][32][60] float32, 16-byte aligned (a row is 7680 B)
static inline int group_corr_tc3_launch(roreg_ctx* c, const float* X, const float* Y, CorrTcArgs a, cudaStream_t st) {
  RR_ARG(c, (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  a.X = X; a.Y = Y; a.trace = nullptr;
  static bool attr_set = false;
  if (!attr_set) {
    RR_CUDA(c, cudaFuncSetAttribute(group_corr_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(group_corr_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM_BYTES));
    attr_set = true;
  }
  const long long items = (long long)a.B * ((a.K + 1) / 2);
  const int grid = (int)(items < c->sm_count ? items : c->sm_count);
  const char* trace_fn = getenv("ROREG_DEBUG_CORR_TRACE");
  static bool traced = false;
  if (trace_fn && !traced && items >= 50000) {           // one-off timeline dump of CTA 0 (debug only; synchronises)
    traced = true;
    const size_t nb = (size_t)grid * 256 * 12 * sizeof(long long);
    RR_CUDA(c, cudaMalloc(&a.trace, nb));
    RR_CUDA(c, cudaMemsetAsync(a.trace, 0, nb, st));
    group_corr_tc3_kernel<true><<<grid, C3_THREADS, C3_SMEM_BYTES, st>>>(a);
    RR_LAUNCH_CHECK(c);
    RR_CUDA(c, cudaStreamSynchronize(st));
    long long* h = (long long*)malloc(256 * 12 * sizeof(long long));
    RR_CUDA(c, cudaMemcpy(h, a.trace, 256 * 12 * sizeof(long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_fn, "w")) {
      fprintf(f, "# it L_issued L_advanced L_tilesfree L_done M_convdone M_accfree M_committed E_mmadone E_loaded E_end (clock64 - first; loader stamps: the X-m0 warp of the group that owns the item)\n");
      const long long t0 = h[0];
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
  group_corr_tc3_kernel<false><<<grid, C3_THREADS, C3_SMEM_BYTES, st>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

}  // namespace roreg
