// kernels_corr_tc3.cuh - equivariant correlation (Des2R / R-indicator), corr mode 3: the shared-memory-lean pipeline.
//
// What runs 23-28 established about modes 1/2 (profiles/r01_run26_corr2_timeline.txt): the bare 7680-byte row gather
// reaches 7.1 TB/s, but the kernels ran at 2.1 TB/s because they are SHARED-MEMORY-bandwidth bound: per match the
// landed rows are written (15 KB) and re-read (15 KB) by the convert stage, the 3xTF32 operand tiles are written
// (32 KB) and read by 12 tf32 MMAs (48 KB), and the 60x60 Gram goes through shared memory once more (30 KB) for the
// generalised-diagonal sums: ~140-165 KB per match against 128 B/clk.  Mode 3 removes what can be removed:
//   * no landing buffers: loader warps LDG a descriptor row straight into registers (15 x 16 B per lane, 11 warps =
//     82 KB in flight per SM) and write the operands from there;
//   * float16 two-accumulator arithmetic (see kernels_nn_tc4.cuh): x = hi + 2^-11 lo', D1 = Xhi.Yhi,
//     D2 = Xlo'.Yhi + Xhi.Ylo', G = D1 + 2^-11 D2 - float32-class products, but the operand tiles are half the bytes
//     (16 KB per match) and the contraction is 6 kind::f16 MMAs per TWO matches (24 KB of operand reads per match);
//   * MN-major operands: a descriptor row [32 f][60 h] is already "h contiguous", so a lane converts its float4
//     (4 consecutive h of one f) and stores 8 bytes - no transpose, conflict-free.  Canonical layout (CUTLASS
//     cute/atom/mma_traits_sm100.hpp, Major-MN / SWIZZLE_128B, in 16-byte units ((8,n),(8,k)):((1,LBO),(8,SBO))):
//     8 f-rows x 128 B (64 h) atoms, chunk16 ^= f % 8, SBO = 1024 B between 8-f groups, LBO = 4096 B between the two
//     matches stacked along M (resp. N).
// Shared-memory traffic per match: 16 KB (operand writes) + 24 KB (MMA reads) + 30 KB (Gram store + diagonal reads) = 70 KB.
// Round 2 (runs c2-c4): the Gram is stored ROW-wise with its columns at coset slots so that both the stores and the table-
// addressed diagonal reads are bank-conflict-free (round 1: 2.3 wavefronts per read); the read table lives in shared memory
// (ptxas had spilled 40 pre-extracted offsets: 40 local loads per item on the epilogue's critical chain); the item walk goes
// pair by pair without per-item count loads.  0.975 -> 0.835 ms per 64 pairs; the kernel is now l1tex-pipe bound (73 %).
//
//   warps 0-10   loaders    a pool over the row tasks {X m0, X m1, Y m0, Y m1} of every item: LDG row -> regs -> fp16 hi / lo' -> STS.64
//   warp 11      MMA        6 x tcgen05.mma kind::f16 (M = N = 128, K = 16, A/B MN-major), 2 x (D1 | D2) in TMEM
//   warps 12-19  epilogue   (two sets of 4 warps alternating over the items)  tcgen05.ld D1, D2 -> FFMA combine -> row-wise 16-byte stores
//                           at the coset slots of icosa_cosets.cuh -> 60 conflict-free loads per generalised diagonal -> argmax
#pragma once
#include <cuda_fp16.h>
#include "kernels_corr_tc.cuh"
#include "kernels_nn_tc4.cuh"
#include "icosa_cosets.cuh"

namespace roreg {

constexpr int C3_GROUPS = 3;                              // operand-tile sets (items in flight between the loaders and the tensor pipe)
// Warp roles are template parameters: LOADERS loader warps (a pool that takes the (item, row) tasks round-robin), one MMA warp,
// ESETS epilogue sets of 4 warps.  Instantiated: <11, 2> = 640 threads, 96 registers per thread.  (<11, 3> = 768 threads leaves 80
// registers per thread - registers are allocated per 4 warps, 21-24 warps cost the same - and ran slower, see the launcher.)
__host__ __device__ constexpr int c3_threads(int loaders, int esets) { return (loaders + 1 + 4 * esets) * 32; }
constexpr int C3_OP_BYTES = 32 * 128;                     // one match's operand tile: 32 f-rows x 64 h fp16 = 4 KB
constexpr int C3_BUF_BYTES = 8 * C3_OP_BYTES;             // Xhi m0|m1, Xlo m0|m1, Yhi m0|m1, Ylo m0|m1 = 32 KB
constexpr int C3_GS_STRIDE = 68;                          // floats per Gram row: 64 slots + 4 -> 16-byte stores of 8 consecutive rows hit 8 distinct bank quads (17 odd)
constexpr int C3_GS_BYTES = 2 * 64 * C3_GS_STRIDE * 4;    // Gram of both matches [2][64 h][68], columns at their coset slots (icosa_cosets.cuh), one per epilogue set
__host__ __device__ constexpr int c3_smem_bytes(int esets) { return C3_GROUPS * C3_BUF_BYTES + esets * C3_GS_BYTES + 3840 + 64 + 256 + 1024; }   // 169 KB (2 sets) / 203 KB (3 sets)
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

template <bool TRACE, int VARIANT, int LOADERS, int ESETS>
__global__ void __launch_bounds__(c3_threads(LOADERS, ESETS), 1) group_corr_tc3_kernel(CorrTcArgs a) {
  constexpr int C3_LOADERS = LOADERS, C3_THREADS = c3_threads(LOADERS, ESETS);
  constexpr int NSLOT = (ESETS == 2) ? 2 : 2 * ESETS;      // mma_done barriers: one per (accumulator, epilogue set) combination an item can have
  static_assert(ESETS >= 2 && ESETS <= 3 && LOADERS >= 4, "role layout");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* bufs = smem;                                                                    // one 32 KB operand-tile set per loader group
  float* Gs = reinterpret_cast<float*>(smem + C3_GROUPS * C3_BUF_BYTES);                   // [ESETS epilogue sets][2 matches][64 h][68]
  uint32_t* tabw = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(Gs) + ESETS * C3_GS_BYTES);   // [2 warp halves][15][32 lanes] packed byte offsets, 3840 B
  uint32_t* red_k = tabw + 2 * 15 * 32; int* red_i = reinterpret_cast<int*>(red_k + 8);   // [ESETS sets][2 matches] each (room for 4 sets)
  uint64_t* bars = reinterpret_cast<uint64_t*>(red_i + 8);                                 // 8-byte aligned: every size above is a multiple of 8
  // barriers: 0..2 conv_done[set], 3..5 tiles_free[set], 6..11 mma_done[it % NSLOT], 12..13 acc_free[acc].
  // mma_done has one barrier per (accumulator, epilogue set) pairing so that its only waiters - the 4 warps of ONE set - see its
  // phases strictly in order (with 3 sets sharing 2 accumulators a per-accumulator barrier could be a phase behind its waiter).
  // Every waiter must only ever have to distinguish ADJACENT phases of a barrier (run 34: with three loader groups sharing
  // two tile sets a group ran two phases ahead of tiles_free and passed the parity test early -> deadlock).
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  // read table of the diagonal sums, by ROW of the Gram: lane l of warp half w sums a = LANE_A[w][l] and needs, in row h, the
  // column g with tab[a][g] = h, stored at slot(g).  Kept in shared memory as packed words tabw[w][h / 4][l] (byte h % 4 =
  // 4 * slot: a byte offset into the row) - consecutive lanes read consecutive words, and nothing table-sized lives in
  // registers across the item loop (run c2: ptxas had spilled 40 pre-extracted offsets, 40 LDL per item on the critical chain).
  {
    const uint8_t* slot = VARIANT == 1 ? C3_SLOT_V1 : C3_SLOT_V2;
    uint8_t* tmp = reinterpret_cast<uint8_t*>(Gs);                       // [60 a][60 h] slot bytes, aliased on the (still unused) Gram buffers
    for (int e = threadIdx.x; e < 3600; e += C3_THREADS) { const int aa = e / 60, g = e - aa * 60; tmp[aa * 60 + a.tab[e]] = slot[g]; }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * 32 * 60; e += C3_THREADS) {
      const int w = e / (32 * 60), l = (e / 60) & 31, hh = e % 60;
      const int aa = (VARIANT == 1 ? C3_LANE_A_V1 : C3_LANE_A_V2)[w][l];
      reinterpret_cast<uint8_t*>(tabw)[((w * 15 + (hh >> 2)) * 32 + l) * 4 + (hh & 3)] = aa < RR_G ? (uint8_t)(4 * tmp[aa * 60 + hh]) : (uint8_t)0;
    }
  }
  // operand tiles start as zeros: the loaders never write the padding columns h = 60..63
  for (int e = threadIdx.x; e < C3_GROUPS * C3_BUF_BYTES / 16; e += C3_THREADS) reinterpret_cast<uint4*>(bufs)[e] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < C3_GROUPS; ++s) { mbar_init(BAR(0 + s), 4); mbar_init(BAR(3 + s), 1); }
    for (int s = 0; s < NSLOT; ++s) mbar_init(BAR(6 + s), 1);
    for (int s = 0; s < 2; ++s) mbar_init(BAR(12 + s), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == C3_LOADERS) {
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
  const int items_per_pair = (a.K + 1) / 2;           // items are numbered p * items_per_pair + j (< 2^31, checked by the launcher)
  // Every role walks the same item sequence: CTA b owns the items congruent to b modulo the grid; item (p, j) covers the matches
  // 2 j, 2 j + 1 of pair p and is live while 2 j < n_matches[p].  The live items of a pair are its FIRST ones, so the walk goes
  // pair by pair: one count load and one modulo per pair, then j advances by the grid size - no per-item global load, and the
  // dead tail of each pair (a third of the item space at 3400 of 5000 matches) is never visited (run c3: the walk was 18 % of
  // the epilogue warps' time and 25 % of the loaders').
  struct Walk { int p, j, cn; };                       // cn = match count of pair p;  p >= B: end
  auto walk_pair = [&](Walk& w) {                      // settle on the first live item of pair w.p or a later pair
    for (; w.p < a.B; ++w.p) {
      w.cn = a.n_matches ? a.n_matches[w.p] : a.K;
      const int first = (int)(((long long)w.p * items_per_pair) % (int)gridDim.x);       // residue of the pair's item 0
      w.j = (int)blockIdx.x - first; if (w.j < 0) w.j += (int)gridDim.x;
      if (2 * w.j < w.cn) return;
    }
  };
  auto walk_begin = [&]() -> Walk { Walk w{0, 0, 0}; walk_pair(w); return w; };
  auto walk_live = [&](const Walk& w) -> bool { return w.p < a.B; };
  auto walk_next = [&](Walk& w) { w.j += (int)gridDim.x; if (2 * w.j >= w.cn) { ++w.p; walk_pair(w); } };
  auto walk_avail = [&](const Walk& w) -> int { return w.cn - 2 * w.j; };   // 1: one match (odd tail), >= 2: two matches

  if (warp < C3_LOADERS) {
    // ===================== loaders =====================
    // Row task T = 4 * (live item index) + role, role: 0 X m0, 1 X m1, 2 Y m0, 3 Y m1; warp w takes the tasks T = w (mod LOADERS).
    // A warp's successive items are at most 3 apart and the MMAs retire in item order, so when it waits for tiles_free of item L
    // (the MMAs of item L-3) the barrier is at most one phase behind - the parity test stays unambiguous.
    // row of a (side, slot) for an item; the odd tail's second slot re-reads the first match
    auto row_ptr = [&](int p, int k0, int avail, int role) -> const float4* {
      const int isY = role >> 1, m = role & 1;
      const int k = k0 + ((m < avail) ? m : 0);
      const long long w = (long long)p * a.K + k;
      const int32_t* idx = isY ? a.idxY : a.idxX;
      long long r = idx ? idx[w * a.idx_stride] : k;
      if (a.pair_cloud) r += (long long)a.pair_cloud[2 * p + (isY ? 0 : 1)] * a.n;
      return reinterpret_cast<const float4*>((isY ? a.Y : a.X) + r * RR_ROW);
    };
    Walk wk = walk_begin(); uint32_t live = 0;         // `live` = index of the next live item of this CTA
    int role_next = warp;                              // (warp - 4 * live) mod 11, kept incrementally
    const float4* next = nullptr; uint32_t next_it = 0; int next_role = 0;
    auto advance = [&]() {
      next = nullptr;
      for (; walk_live(wk); walk_next(wk)) {
        const int avail = walk_avail(wk);
        const uint32_t my = live++;
        const int role = role_next;
        role_next = role_next >= 4 ? role_next - 4 : role_next + C3_LOADERS - 4;
        if (role < 4) { next = row_ptr(wk.p, 2 * wk.j, avail, role); next_it = my; next_role = role; walk_next(wk); return; }
      }
    };
    advance();
    while (next) {
      const float4* src = next; const uint32_t it = next_it; const int role = next_role;
      // lanes 0..14 take the 15 float4 of descriptor row f = 2 j, lanes 16..30 those of row f = 2 j + 1 (lanes 15, 31 idle): a
      // half-warp then stores into ONE 128-byte f-row of the tile - conflict-free - where the flat mapping (32 consecutive
      // float4 per instruction) straddled two or three f-rows and doubled the store wavefronts (run 42: LSU data pipe 81 % busy)
      const int hw = lane >> 4, ql = lane & 15;
      const bool act = ql < 15;
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = act ? ldg_stream4(src + 30 * j + 15 * hw + ql) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (role == 0 && lane == 0) C3_TRACE(it, 0);
      advance();                                       // resolve the next row while this one is in flight
      const int set = it % C3_GROUPS; const uint32_t use = it / C3_GROUPS;
      if (role == 0 && lane == 0) C3_TRACE(it, 1);
      mbar_wait_lean(BAR(3 + set), (use & 1) ^ 1);     // the MMAs of the set's previous item no longer read it
      if (role == 0 && lane == 0) C3_TRACE(it, 2);
      uint8_t* thi = bufs + set * C3_BUF_BYTES + ((role >> 1) * 4 + (role & 1)) * C3_OP_BYTES;
      uint8_t* tlo = thi + 2 * C3_OP_BYTES;
      const int qh = ql >> 1, qb = (ql & 1) * 8;       // h = 4 ql: 16-byte chunk h / 8 and the 8-byte half inside it
      if (act)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __half2 h01 = __floats2half2_rn(v[j].x, v[j].y), h23 = __floats2half2_rn(v[j].z, v[j].w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((v[j].x - f01.x) * T4_LO_SCALE, (v[j].y - f01.y) * T4_LO_SCALE);
        const __half2 l23 = __floats2half2_rn((v[j].z - f23.x) * T4_LO_SCALE, (v[j].w - f23.y) * T4_LO_SCALE);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        const int fr = ((2 * j) & 7) | hw;               // f % 8 with f = 2 j + hw;  f / 8 = j / 4
        const int off = (j >> 2) * 1024 + fr * 128 + ((qh ^ fr) << 4) + qb;
        *reinterpret_cast<uint2*>(thi + off) = hv;
        *reinterpret_cast<uint2*>(tlo + off) = lv;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(0 + set));        // conv_done: 4 row tasks per item
      if (role == 0 && lane == 0) C3_TRACE(it, 3);
    }
  } else if (warp == C3_LOADERS) {
    // ===================== MMA issuer: the whole warp walks the loop, one elected lane issues =====================
    uint32_t it = 0;
    for (Walk wk = walk_begin(); walk_live(wk); walk_next(wk)) {
      const int g = it % C3_GROUPS, acc = it & 1; const uint32_t gph = (it / C3_GROUPS) & 1, aph = (it >> 1) & 1;
      mbar_wait_lean(BAR(0 + g), gph);                 // operand tiles written and visible to the async proxy
      if (lane == 0) C3_TRACE(it, 4);
      mbar_wait_lean(BAR(12 + acc), aph ^ 1);          // accumulators drained
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
      umma_commit_elect(BAR(6 + it % NSLOT));          // accumulators ready for this item's epilogue set
      if (lane == 0) C3_TRACE(it, 6);
      ++it;
    }
  } else {
    // ===================== epilogue =====================
    // ESETS sets of 4 warps: set s takes the items with it % ESETS == s (accumulator buffer it & 1) - a warp is alone on its
    // latency chain (run 39/40: 2100 clk per item), so ESETS items are drained side by side
    const int eset = (warp - (C3_LOADERS + 1)) >> 2;
    const int q = warp & 3;                            // TMEM lane quadrant of this warp
    const int m = q >> 1;                              // match slot: lanes 0..63 -> 0, 64..127 -> 1
    const int h = (q & 1) * 32 + lane;                 // Gram row (h) this thread owns in TMEM and stores
    float* G = Gs + (eset * 2 + m) * 64 * C3_GS_STRIDE;
    float* grow = G + h * C3_GS_STRIDE;
    const int nbar = 2 + eset * 2 + m;                 // named barrier of this (set, match) warp pair
    uint32_t* rk = red_k + eset * 2; int* ri = red_i + eset * 2;
    // the generalised diagonal this thread sums: the two warps of a match take the two lane lists of icosa_cosets.cuh (30 lanes
    // each), whose columns are a transversal of the cosets in every row -> each LDS below touches 30 distinct banks
    const int my_a = (VARIANT == 1 ? C3_LANE_A_V1 : C3_LANE_A_V2)[q & 1][lane];
    const bool has_a = my_a < RR_G;
    const uint32_t* tw_lane = tabw + (q & 1) * 15 * 32 + lane;
    const uint8_t* Gb = reinterpret_cast<const uint8_t*>(G);
    uint32_t it = 0;
    for (Walk wk = walk_begin(); walk_live(wk); walk_next(wk)) {
      const int avail = walk_avail(wk);
      if ((int)(it % ESETS) != eset) { ++it; continue; }
      const int p = wk.p, k0 = 2 * wk.j;
      const int buf = it & 1; const uint32_t ph = (it / NSLOT) & 1;
      mbar_wait_lean(BAR(6 + it % NSLOT), ph);
      if ((warp & 3) == 0 && lane == 0) C3_TRACE(it, 7);
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
          if (lane == 0) mbar_arrive(BAR(12 + buf));   // accumulators free as soon as they sit in registers
          if ((warp & 3) == 0 && lane == 0) C3_TRACE(it, 8);
        }
        // row store: this thread's row h, column g at slot(g), as 16-byte stores (8 consecutive rows = 8 distinct bank quads)
        float v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = fmaf(__uint_as_float(r2[u]), T4_LO_UNSCALE, __uint_as_float(r1[u]));
        if (VARIANT == 1) { if (half == 0) C3_STORE_V1_H0(grow, v); else C3_STORE_V1_H1(grow, v); }
        else              { if (half == 0) C3_STORE_V2_H0(grow, v); else C3_STORE_V2_H1(grow, v); }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(nbar) : "memory");
      if ((warp & 3) == 0 && lane == 0) C3_TRACE(it, 10);
      float c = -INFINITY;
      if (has_a) {
        // 20 loads in flight at a time (independent addresses straight from the table bytes), then a fixed-shape tree: the warp is
        // alone on its latency chain, so the shared-memory latency must be paid three times, not 60 (the sum runs over the rows
        // h = 0..59 of the Gram instead of g = 0..59: same terms, different rounding order)
        float c8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int part = 0; part < 3; ++part) {           // 20 loads in flight at a time (register budget: 96 per thread)
          float gv[20];
#pragma unroll
          for (int w = 0; w < 5; ++w) {
            const int w4 = part * 5 + w;
            const uint32_t tw = tw_lane[w4 * 32];
            gv[4 * w + 0] = *reinterpret_cast<const float*>(Gb + (4 * w4 + 0) * (C3_GS_STRIDE * 4) + __byte_perm(tw, 0, 0x4440));
            gv[4 * w + 1] = *reinterpret_cast<const float*>(Gb + (4 * w4 + 1) * (C3_GS_STRIDE * 4) + __byte_perm(tw, 0, 0x4441));
            gv[4 * w + 2] = *reinterpret_cast<const float*>(Gb + (4 * w4 + 2) * (C3_GS_STRIDE * 4) + __byte_perm(tw, 0, 0x4442));
            gv[4 * w + 3] = *reinterpret_cast<const float*>(Gb + (4 * w4 + 3) * (C3_GS_STRIDE * 4) + __byte_perm(tw, 0, 0x4443));
          }
#pragma unroll
          for (int k = 0; k < 20; ++k) c8[(part * 20 + k) & 7] += gv[k];
        }
        c = ((c8[0] + c8[1]) + (c8[2] + c8[3])) + ((c8[4] + c8[5]) + (c8[6] + c8[7]));
      }
      if ((warp & 3) == 0 && lane == 0) C3_TRACE(it, 11);
      const bool valid = m < avail;
      const long long w = (long long)p * a.K + k0 + m;
      if (valid && has_a && a.cor_out) a.cor_out[w * RR_G + my_a] = c;
      // first maximal index (torch.argmax): warp-wide integer max of the order-preserving key, then the smallest a attaining it
      const uint32_t key = has_a ? t4_ord(c + 0.f) : 0u;                   // + 0.f: -0 and +0 compare equal, as in float arithmetic
      const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
      const int ix = (int)__reduce_min_sync(0xffffffffu, (has_a && key == kmax) ? (uint32_t)my_a : 0x7fffffffu);
      if ((q & 1) == 1 && lane == 0) { rk[m] = kmax; ri[m] = ix; }          // the second warp of the match publishes
      asm volatile("bar.sync %0, 64;" ::"r"(nbar) : "memory");
      if ((q & 1) == 0 && lane == 0 && valid && a.argmax_out) {
        int best = ix;                                                      // the lane lists interleave: equal keys -> the smaller a
        if (rk[m] > kmax || (rk[m] == kmax && ri[m] < ix)) best = ri[m];
        a.argmax_out[w] = best;
      }
      // no third barrier: the partner warp finished reading Gs before the barrier above, and it reaches the next item's first
      // barrier (after which `red` is rewritten) only after this warp has read `red`
      if ((warp & 3) == 0 && lane == 0) C3_TRACE(it, 9);
      ++it;
    }
  }
  __syncthreads();
  if (warp == C3_LOADERS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// X, Y: descriptor arrays [rows][32][60] float32, 16-byte aligned (a row is 7680 B)
static inline int group_corr_tc3_launch(roreg_ctx* c, const float* X, const float* Y, CorrTcArgs a, cudaStream_t st) {
  const bool v2 = (a.tab == c->d_permT8);               // variant 2 (R-indicator convention): its own coset layout
  RR_ARG(c, (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  a.X = X; a.Y = Y; a.trace = nullptr;
  // role layout <11 loaders, 2 epilogue sets>.  A third epilogue set (<11, 3>, 768 threads) was measured in run c4: 80 registers
  // per thread force spills in the loaders and the epilogue, Des2R 1.07 ms against 0.835 ms - not instantiated.
  static unsigned long long attr_mask = 0;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(group_corr_tc3_kernel<false, 1, 11, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, c3_smem_bytes(2)));
    RR_CUDA(c, cudaFuncSetAttribute(group_corr_tc3_kernel<false, 2, 11, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, c3_smem_bytes(2)));
    RR_CUDA(c, cudaFuncSetAttribute(group_corr_tc3_kernel<true, 1, 11, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, c3_smem_bytes(2)));
  }
  const long long items = (long long)a.B * ((a.K + 1) / 2);
  RR_ARG(c, items < (1LL << 31));
  const int grid = (int)(items < c->sm_count ? items : c->sm_count);
  const char* trace_fn = getenv("ROREG_DEBUG_CORR_TRACE");
  static bool traced = false;
  if (trace_fn && !traced && !v2 && items >= 50000) {    // one-off timeline dump of CTA 0 (debug only; synchronises)
    traced = true;
    const size_t nb = (size_t)grid * 256 * 12 * sizeof(long long);
    RR_CUDA(c, cudaMalloc(&a.trace, nb));
    RR_CUDA(c, cudaMemsetAsync(a.trace, 0, nb, st));
    group_corr_tc3_kernel<true, 1, 11, 2><<<grid, c3_threads(11, 2), c3_smem_bytes(2), st>>>(a);
    RR_LAUNCH_CHECK(c);
    RR_CUDA(c, cudaStreamSynchronize(st));
    long long* h = (long long*)malloc(256 * 12 * sizeof(long long));
    RR_CUDA(c, cudaMemcpy(h, a.trace, 256 * 12 * sizeof(long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_fn, "w")) {
      fprintf(f, "# it L_issued L_advanced L_tilesfree L_done M_convdone M_accfree M_committed E_mmadone E_loaded E_end E_stored E_summed (clock64 - first; loader stamps: the X-m0 warp of the group that owns the item)\n");
      const long long t0 = h[0];
      for (int i = 0; i < 255; ++i) {
        fprintf(f, "%d", i);
        for (int e = 0; e < 12; ++e) fprintf(f, " %lld", h[i * 12 + e] ? h[i * 12 + e] - t0 : -1);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    free(h); cudaFree(a.trace);
    return ROREG_OK;
  }
  if (v2) group_corr_tc3_kernel<false, 2, 11, 2><<<grid, c3_threads(11, 2), c3_smem_bytes(2), st>>>(a);
  else group_corr_tc3_kernel<false, 1, 11, 2><<<grid, c3_threads(11, 2), c3_smem_bytes(2), st>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

}  // namespace roreg
