// kernels_gemm_tc.cuh - dense layer GEMM for the group-convolution networks (GF, ET, RD) on tcgen05.
//
//   out[r][o] = sum_c A[r][c] * W[o][c]         A: [R][Kdim] activations (rows = (item, group element)),
//                                                W: [O][Kdim] weights, both K-major float32 in HBM.
// A group convolution (network/ops.py:11-20) is this GEMM on the neighbour-gathered activation matrix, gathered by the
// producer warp itself (Kdim = 13*Cin, GemmArgs.g_*, see kernels_gconv.cuh); the 1x1 FC head of ET_test (network/eqv_trans.py:93-101) is it
// directly.  Precision: npass = 1 -> one TF32 pass (what the reference's own cuDNN path does on Ampere+,
// torch.backends.cudnn.allow_tf32 defaults to True), npass = 3 -> hi/lo split of both operands, three
// passes (hi.hi + lo.hi + hi.lo) into the same accumulator: float32-class, used for parity with the oracle.
// Fused epilogue (one thread per output row, TMEM -> registers):
//   v = acc + bias[o] (+ residual[r][o]);  raw_out[r][o] = v;  y = v*bn_scale[o] + bn_shift[o];  relu;
//   act_hi[r][o] = tf32(y), act_lo[r][o] = y - tf32(y)      (the next layer's operands)
// Pipeline: warps 0-3 TMA producers (SWIZZLE_128B boxes, or tile::gather4 row gathers for the implicit group convolution),
// warp 4 MMA issuer, warps 5-8 epilogue; every operand image of a k-chunk is loaded ONCE per stage and all passes run from
// it; two TMEM accumulators of up to 256 columns so the epilogue of tile t overlaps the MMAs of tile t+1.
#pragma once
#include "kernels_nn_tc.cuh"

namespace roreg {

constexpr int GM_BM = 128, GM_KC = 32;
constexpr int GM_A_BYTES = GM_BM * GM_KC * 4;            // 16 KB
constexpr int GM_W_BYTES = 256 * GM_KC * 4;              // up to 32 KB (NT <= 256 weight rows)
// One stage holds EVERY operand image the passes of a k-chunk need, so each image is read from L2 once per k-chunk
// (round 1 re-loaded A_hi and W_hi for the second / third pass: 144 KB per k-chunk against the ~42 B/clk/SM the L2 delivers):
//   1 pass : A | W                 48 KB x 4 stages
//   3 pass : A_hi | A_lo | W_hi | W_lo   96 KB x 2 stages (a stage is 12 MMAs = ~1500 tensor clocks: two are enough to prefetch)
template <int NPASS> struct GemmCfg {
  static constexpr int STAGES = NPASS == 3 ? 2 : 4;
  static constexpr int A_IMAGES = NPASS == 3 ? 2 : 1;
  static constexpr int STAGE_BYTES = A_IMAGES * (GM_A_BYTES + GM_W_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;      // the dynamic array is declared 1024-byte aligned: no slack needed
  static constexpr int OFF_ALO = GM_A_BYTES;                         // 3-pass only
  static constexpr int OFF_WHI = A_IMAGES * GM_A_BYTES;
  static constexpr int OFF_WLO = OFF_WHI + GM_W_BYTES;               // 3-pass only
};

struct GemmArgs {
  int R, Kdim, O, NT;          // rows, contraction length (multiple of 32), valid output channels, N tile (16..256, %16 == 0)
  int n_ntiles, npass;
  const float* bias;           // [O] or NULL
  const float* residual; int res_ld;
  float* raw_out; int raw_ld;
  float* act_hi; float* act_lo; int act_ld;
  const float* bn_scale; const float* bn_shift; int relu;
  // all-pairs group correlation (roreg_group_corr_allpairs): the K axis is 60 chunks of 32 channels, one per group
  // element; rotation `a` pairs A-chunk P[a][g] with W-chunk g, i.e. the A operand's column coordinate of k-chunk g is
  // a_cols[g] - a permutation applied by the TMA coordinate alone, no data movement.
  const int32_t* a_cols;       // [Kdim/32] A column (element) coordinate per k-chunk, or NULL = kc*32
  uint8_t* amax_arg; int amax_id;   // if set: raw_out / amax_arg hold a running (max, argmax id) instead of being overwritten
  // IMPLICIT GROUP CONVOLUTION (network/group_feat.py:20-24 data_process folded into the operand load): when g_C > 0 the A
  // operand is never materialised.  The tensor maps describe the channel-last ACTIVATION [n_items*60][g_C]; GEMM row
  // r = item * g_ng + j (output group element g = g_set ? g_set[j] : j) and k-chunk kc (tap k = kc*32 / g_C, channels
  // c0 = kc*32 % g_C) read activation row item*60 + g_nei[g][k], columns c0..c0+31 - fetched four rows per instruction with
  // cp.async.bulk.tensor ... tile::gather4 straight into the SWIZZLE_128B operand tile (Kdim = 13 * g_C).
  int g_C, g_ng; const int32_t* g_nei; const int32_t* g_set;
  // g_ldg != 0: the gathered operand is loaded by the producer warps with LDG.128 and stored into the swizzled tile with STS.128
  // (needs the raw activation pointers and NPROD == 8); 0: TMA tile::gather4 through the tensor maps.
  int g_ldg; const float* g_act_hi; const float* g_act_lo;
  int dbg;   // ROREG_DEBUG_GEMM (bottleneck experiments, results WRONG): 1 = epilogue without its global loads / stores, 2 = LDG loader without the loads
};

// four rows of a 2-D tensor (box = 32 columns x 1 row) -> 4 x 128 B at dst, swizzled on the absolute shared-memory address
// exactly like a 4-row box (scripts/gather4_test.cu, profiles/r02_gather4_test.txt)
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// Warp roles: NPROD producer warps, then the MMA warp, then four epilogue warps.  NPROD is a template parameter because the gather
// issue is serialised per lane (see the producer section): 4 / 8 / 16 producer warps issue 8 / 4 / 2 gathers each per k-chunk.
__host__ __device__ constexpr int gm_threads(int nprod) { return (nprod + 1 + 4) * 32; }

template <int NPASS, int NPROD>
__global__ void __launch_bounds__(gm_threads(NPROD), 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                                                         const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo,
                                                         GemmArgs a) {
  using Cfg = GemmCfg<NPASS>;
  constexpr int GM_PRODUCERS = NPROD, LANES = 32 / NPROD;     // gather lanes per producer warp
  // Shared-memory budget: static (~1.4 KB, rounded up to the array's 1 KB alignment) + 192 KB of stages + barriers = 194.3 KiB,
  // i.e. under the 196 KiB carve-out step (+1 KiB the system reserves per CTA) - see the note at nei_s.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) { if (threadIdx.x == 0) printf("roreg: gemm_tc_kernel: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  // barriers: 0..3 full, 4..7 empty, 8..9 tmem_full, 10..11 tmem_empty
  __shared__ uint32_t tmem_base_s;
  __shared__ uint16_t acols_s[256];
  // byte tables: together with the dynamic buffers the CTA must stay under the 196 KiB shared-memory carve-out - the next step
  // (228 KiB) halves the L1 that the epilogue's row-strided loads live in (run c7: 3.4 KB of int32 tables cost the all-pairs
  // GEMMs 45 %)
  __shared__ uint8_t nei_s[60 * 13];
  __shared__ uint8_t gset_s[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const bool gather = a.g_C > 0;
  for (int i = threadIdx.x; i < a.Kdim / GM_KC && i < 256; i += blockDim.x) acols_s[i] = (uint16_t)(a.a_cols ? a.a_cols[i] : i * GM_KC);   // < Kdim <= 8192
  if (gather) {
    for (int i = threadIdx.x; i < 60 * 13; i += blockDim.x) nei_s[i] = (uint8_t)a.g_nei[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) gset_s[i] = (uint8_t)((a.g_set && i < a.g_ng) ? a.g_set[i] : i);
  }
  if (threadIdx.x == 0) {
    // full[st]: one arrival from the TMA-issuing lane (arrive.expect_tx); with the LDG loader also one per warp of the group that
    // filled the stage's A tile (4)
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(BAR(s), (gather && a.g_ldg) ? 5 : 1); mbar_init(BAR(4 + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(8 + s), 1); mbar_init(BAR(10 + s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GM_PRODUCERS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const int n_mt = (a.R + GM_BM - 1) / GM_BM;
  const int n_tiles = n_mt * a.n_ntiles;
  const int n_kc = a.Kdim / GM_KC;
  const uint32_t w_bytes = (uint32_t)a.NT * GM_KC * 4;
  const uint32_t stage_tx = Cfg::A_IMAGES * (GM_A_BYTES + w_bytes);
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.NT >> 3) << 17) | ((uint32_t)(GM_BM >> 4) << 24);

  if (warp < GM_PRODUCERS) {
    // ===================== producers =====================
    // Plain mode: lane 0 of warp 0 issues two (four) box loads per stage, the other producer warps have nothing to do.
    // Gather mode: a tile::gather4 instruction takes its four row coordinates from uniform registers, so the compiler
    // serialises lanes that hold different rows (ELECT + 6 x R2UR + UTMALDG per lane: run c5 - one warp issuing all 32 gathers
    // of a stage took ~1300 clk per k-chunk and made GF slower than the im2col version).  The 32 gathers of a stage are
    // therefore spread over NPROD warps: warp w, lanes 0..LANES-1 fetch the row groups w * LANES + l of the tile.  Run c11 (4 warps
    // x 8 lanes): ~2000 clk per k-chunk in the issue loop (ELECT, 6 x R2UR.BROADCAST, UTMALDG, branch: ~250 clk per gather)
    // against 512 clk of MMA work - the producers, not the tensor pipe or L2, paced the gather layers (tensor pipe 33 % active).
    const int cpk = gather ? a.g_C / GM_KC : 1;          // k-chunks per tap
    if (gather && a.g_ldg) {
      // ---- LDG / STS loader (run c16: the plain-mode pipeline reaches 94 % tensor-pipe activity in the all-pairs kernel, so the
      // gather layers' 33 % was the TMA gather itself).  Two groups of four warps take alternate stages.  In a group, thread
      // (row lane rl = t / 8, chunk q = t % 8) loads the 16-byte chunk q of rows rl + 16 j (j = 0..7): a warp instruction covers
      // four full 128-byte activation rows, and stores it at the SWIZZLE_128B position the MMA descriptor expects (row r at
      // (r / 8) * 1024 + (r % 8) * 128, chunk q ^ (r % 8)): four rows = four distinct 128-byte lines per STS.128, conflict-free.
      // One pass: the loads of the group's NEXT stage are issued before the current one is stored (registers, no barrier needed).
      if (NPROD == 8) {
        const int grp = warp >> 2, tg = (warp & 3) * 32 + lane;
        const int q = tg & 7, rl = tg >> 3;
        const uint32_t sw_off = (uint32_t)((rl >> 3) * 1024 + (rl & 7) * 128 + ((q ^ (rl & 7)) << 4));     // + 2048 j
        uint32_t it0 = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, it0 += (uint32_t)n_kc) {
          const int mt = t / a.n_ntiles, nt = t % a.n_ntiles;
          int base[8], tap0[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            int r = mt * GM_BM + rl + 16 * j;
            if (r >= a.R) r = a.R - 1;                   // rows past the end: any valid row (their outputs are masked)
            const int item = r / a.g_ng, jj = r - item * a.g_ng;
            base[j] = item * 60; tap0[j] = (int)gset_s[jj] * 13;
          }
          auto issue = [&](int kc, const float* act, float4 (&buf)[8]) {
            if (a.dbg == 2) return;
            const int k_tap = kc / cpk, c0 = (kc - k_tap * cpk) * GM_KC + 4 * q;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              buf[j] = __ldg(reinterpret_cast<const float4*>(act + (long long)(base[j] + (int)nei_s[tap0[j] + k_tap]) * a.g_C + c0));
          };
          const int first = (int)((grp - (int)(it0 & 1u)) & 1);
          float4 cur[8], nxt[8], lo[8];
          if (first < n_kc) issue(first, a.g_act_hi, cur);
          for (int kc = first; kc < n_kc; kc += 2) {
            if (NPASS == 1 && kc + 2 < n_kc) issue(kc + 2, a.g_act_hi, nxt);
            if (NPASS == 3) issue(kc, a.g_act_lo, lo);
            const uint32_t it = it0 + (uint32_t)kc;
            const int st = it % Cfg::STAGES; const uint32_t ph = (it / Cfg::STAGES) & 1;
            if (lane == 0) mbar_wait(BAR(4 + st), ph ^ 1);
            __syncwarp();
            uint8_t* sbp = smem + st * Cfg::STAGE_BYTES;
            if ((warp & 3) == 0 && lane == 0) {
              const uint32_t sb = smem_u32(sbp);
              mbar_expect_tx(BAR(st), Cfg::A_IMAGES * w_bytes);
              tma_load_2d(sb + Cfg::OFF_WHI, &mapWhi, kc * GM_KC, nt * a.NT, BAR(st));
              if (NPASS == 3) tma_load_2d(sb + Cfg::OFF_WLO, &mapWlo, kc * GM_KC, nt * a.NT, BAR(st));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(sbp + sw_off + 2048 * j) = cur[j];
            if (NPASS == 3) {
#pragma unroll
              for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(sbp + Cfg::OFF_ALO + sw_off + 2048 * j) = lo[j];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to tcgen05
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(st));
            if (NPASS == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
            } else if (kc + 2 < n_kc) {
              issue(kc + 2, a.g_act_hi, cur);
            }
          }
        }
      }
    } else if (gather || warp == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int mt = t / a.n_ntiles, nt = t % a.n_ntiles;
        int base[4] = {0, 0, 0, 0}, tap0[4] = {0, 0, 0, 0};
        if (gather && lane < LANES) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int r = mt * GM_BM + 4 * (warp * LANES + lane) + i;
            if (r >= a.R) r = a.R - 1;                   // rows past the end: any valid row (their outputs are masked)
            const int item = r / a.g_ng, j = r - item * a.g_ng;
            base[i] = item * 60; tap0[i] = (int)gset_s[j] * 13;
          }
        }
        int k_tap = 0, k_sub = 0;                        // kc = k_tap * cpk + k_sub
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int st = it % Cfg::STAGES; const uint32_t ph = (it / Cfg::STAGES) & 1;
          if (lane == 0) mbar_wait(BAR(4 + st), ph ^ 1);  // one poller per warp
          __syncwarp();
          const uint32_t sb = smem_u32(smem + st * Cfg::STAGE_BYTES);
          if (warp == 0 && lane == 0) {
            mbar_expect_tx(BAR(st), stage_tx);
            tma_load_2d(sb + Cfg::OFF_WHI, &mapWhi, kc * GM_KC, nt * a.NT, BAR(st));
            if (NPASS == 3) tma_load_2d(sb + Cfg::OFF_WLO, &mapWlo, kc * GM_KC, nt * a.NT, BAR(st));
            if (!gather) {
              tma_load_2d(sb, &mapAhi, (int)acols_s[kc], mt * GM_BM, BAR(st));
              if (NPASS == 3) tma_load_2d(sb + Cfg::OFF_ALO, &mapAlo, (int)acols_s[kc], mt * GM_BM, BAR(st));
            }
          }
          if (gather) {
            if (lane < LANES) {
              const int c0 = k_sub * GM_KC;
              const int r0 = base[0] + (int)nei_s[tap0[0] + k_tap], r1 = base[1] + (int)nei_s[tap0[1] + k_tap];
              const int r2 = base[2] + (int)nei_s[tap0[2] + k_tap], r3 = base[3] + (int)nei_s[tap0[3] + k_tap];
              const uint32_t dst = sb + (warp * LANES + lane) * 512;
              tma_gather4(dst, &mapAhi, c0, r0, r1, r2, r3, BAR(st));
              if (NPASS == 3) tma_gather4(dst + Cfg::OFF_ALO, &mapAlo, c0, r0, r1, r2, r3, BAR(st));
            }
            if (++k_sub == cpk) { k_sub = 0; ++k_tap; }
          }
        }
      }
    }
  } else if (warp == GM_PRODUCERS) {
    if (lane == 0) {
      uint32_t it = 0, it_t = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it_t) {
        const int acc = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
        mbar_wait(BAR(10 + acc), tph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int st = it % Cfg::STAGES; const uint32_t ph = (it / Cfg::STAGES) & 1;
          mbar_wait(BAR(st), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + st * Cfg::STAGE_BYTES);
          const uint32_t whi = sa + Cfg::OFF_WHI;
#pragma unroll
          for (int kk = 0; kk < GM_KC / 8; ++kk)          // A_hi . W_hi
            umma_tf32(d_tmem, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(whi + kk * 32), idesc, (kc | kk) ? 1u : 0u);
          if (NPASS == 3) {
            const uint32_t alo = sa + Cfg::OFF_ALO, wlo = sa + Cfg::OFF_WLO;
#pragma unroll
            for (int kk = 0; kk < GM_KC / 8; ++kk)        // A_lo . W_hi
              umma_tf32(d_tmem, umma_desc_sw128(alo + kk * 32), umma_desc_sw128(whi + kk * 32), idesc, 1u);
#pragma unroll
            for (int kk = 0; kk < GM_KC / 8; ++kk)        // A_hi . W_lo
              umma_tf32(d_tmem, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(wlo + kk * 32), idesc, 1u);
          }
          umma_commit(BAR(4 + st));
        }
        umma_commit(BAR(8 + acc));
      }
    }
  } else {
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    uint32_t it_t = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it_t) {
      const int mt = t / a.n_ntiles, nt = t % a.n_ntiles;
      const int acc = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
      mbar_wait(BAR(8 + acc), tph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long r = (long long)mt * GM_BM + row_in_tile;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      for (int c0 = 0; c0 < a.NT; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (r < a.R && a.dbg != 1) {
          // running-maximum mode: all 16 previous maxima of this chunk are loaded BEFORE any update.  Round 1 read them one by one
          // between the stores (same array: the compiler may not hoist a load above a store), i.e. 256 serialised L2 round trips
          // per thread and tile - the all-pairs GEMMs were bound by that chain, not by the tensor pipe (run c7: 355 us per launch
          // against 94 us of MMA work).
          float old[16], res[16];
          // 16-byte loads where the row chunk allows it: one L2 request per four values, no reliance on L1 reuse (with ~200 KB
          // of shared memory per CTA only ~28 KB of L1 are left: the epilogue's per-thread sectors do not survive there)
          const bool vec_res = a.residual && ((a.res_ld & 3) == 0) && (nt * a.NT + c0 + 15 < a.O);
          const bool vec_old = a.amax_arg && ((a.raw_ld & 3) == 0) && (nt * a.NT + c0 + 15 < a.O);
          if (a.residual) {                               // the residual row chunk in one go, ahead of the stores
            if (vec_res) {
              const float4* pr = reinterpret_cast<const float4*>(a.residual + r * a.res_ld + nt * a.NT + c0);
#pragma unroll
              for (int q = 0; q < 4; ++q) { const float4 w = __ldcg(pr + q); res[4 * q] = w.x; res[4 * q + 1] = w.y; res[4 * q + 2] = w.z; res[4 * q + 3] = w.w; }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) { const int oj = nt * a.NT + c0 + j; res[j] = (oj < a.O) ? a.residual[r * a.res_ld + oj] : 0.f; }
            }
          }
          if (a.amax_arg) {
            if (vec_old) {
              const float4* po = reinterpret_cast<const float4*>(a.raw_out + r * a.raw_ld + nt * a.NT + c0);
#pragma unroll
              for (int q = 0; q < 4; ++q) { const float4 w = __ldcg(po + q); old[4 * q] = w.x; old[4 * q + 1] = w.y; old[4 * q + 2] = w.z; old[4 * q + 3] = w.w; }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) { const int oj = nt * a.NT + c0 + j; old[j] = (oj < a.O) ? a.raw_out[r * a.raw_ld + oj] : INFINITY; }
            }
          }
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int o = nt * a.NT + c0 + 4 * j4;
            if (o >= a.O) break;
            float x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int oj = o + j;
              float val = __uint_as_float(v[4 * j4 + j]);
              if (oj < a.O) {
                if (a.bias) val += __ldg(a.bias + oj);
                if (a.residual) val += res[4 * j4 + j];
              }
              x[j] = val;
            }
            const bool full = (o + 3 < a.O);
            if (a.amax_arg) {          // running (max, argmax) over successive launches; strict '>' keeps the first maximum
              const long long ix0 = r * a.raw_ld + o;
              const bool u0 = x[0] > old[4 * j4], u1 = x[1] > old[4 * j4 + 1], u2 = x[2] > old[4 * j4 + 2], u3 = x[3] > old[4 * j4 + 3];
              if (vec_old) {
                if (u0 | u1 | u2 | u3) {                  // one 16-byte + one 4-byte store per group of four instead of up to eight scalar ones
                  *reinterpret_cast<float4*>(a.raw_out + ix0) = make_float4(u0 ? x[0] : old[4 * j4], u1 ? x[1] : old[4 * j4 + 1],
                                                                           u2 ? x[2] : old[4 * j4 + 2], u3 ? x[3] : old[4 * j4 + 3]);
                  uchar4* pa = reinterpret_cast<uchar4*>(a.amax_arg + ix0);
                  uchar4 ar = (a.amax_id == 0) ? make_uchar4(0, 0, 0, 0) : *pa;      // first rotation: nothing to keep
                  const uint8_t id = (uint8_t)a.amax_id;
                  if (u0) ar.x = id; if (u1) ar.y = id; if (u2) ar.z = id; if (u3) ar.w = id;
                  *pa = ar;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (o + j < a.O && x[j] > old[4 * j4 + j]) { a.raw_out[ix0 + j] = x[j]; a.amax_arg[ix0 + j] = (uint8_t)a.amax_id; }
              }
            } else if (a.raw_out) {
              float* p = a.raw_out + r * a.raw_ld + o;
              if (full && ((a.raw_ld & 3) == 0)) *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
              else for (int j = 0; j < 4 && o + j < a.O; ++j) p[j] = x[j];
            }
            if (a.act_hi) {
              float hi[4], lo[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float y = x[j];
                if (a.bn_scale && o + j < a.O) y = fmaf(y, __ldg(a.bn_scale + o + j), __ldg(a.bn_shift + o + j));
                if (a.relu) y = fmaxf(y, 0.f);
                uint32_t tb; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tb) : "f"(y));
                hi[j] = __uint_as_float(tb); lo[j] = y - hi[j];
              }
              float* ph = a.act_hi + r * a.act_ld + o;
              if (full && ((a.act_ld & 3) == 0)) {
                *reinterpret_cast<float4*>(ph) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                if (a.act_lo) *reinterpret_cast<float4*>(a.act_lo + r * a.act_ld + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
              } else {
                for (int j = 0; j < 4 && o + j < a.O; ++j) { ph[j] = hi[j]; if (a.act_lo) a.act_lo[r * a.act_ld + o + j] = lo[j]; }
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(BAR(10 + acc));
    }
  }
  __syncthreads();
  if (warp == GM_PRODUCERS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

static inline int gemm_make_map(roreg_ctx* c, CUtensorMap* m, const float* base, long long rows, int kdim, int box_rows) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !p || qres != cudaDriverEntryPointSuccess) { snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled entry point unavailable"); return ROREG_ERR_CUDA; }
    fn = (PFN_encodeTiled)p;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kdim * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)GM_KC, (cuuint32_t)box_rows};      // box_rows = 1: tile::gather4 source (rows chosen per instruction)
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled(gemm) failed (%d) rows=%lld kdim=%d box=%d", (int)r, rows, kdim, box_rows); return ROREG_ERR_CUDA; }
  return ROREG_OK;
}

// A_hi/A_lo: [R][Kdim] - or, in gather mode (a.g_C > 0), the channel-last activation [n_act_rows][g_C];
// W_hi/W_lo: [Opad][Kdim] with Opad = n_ntiles*NT rows allocated (rows >= O may be anything finite: masked).
static inline int gemm_tc_launch(roreg_ctx* c, const float* A_hi, const float* A_lo, const float* W_hi, const float* W_lo,
                                 long long w_rows, GemmArgs a, cudaStream_t st, long long n_act_rows = 0) {
  if (a.R <= 0) return ROREG_OK;
  if (a.Kdim / GM_KC > 256 || a.Kdim % GM_KC || a.NT % 16 || a.NT < 16 || a.NT > 256 || (a.npass != 1 && a.npass != 3) || (a.npass == 3 && (!A_lo || !W_lo))) {
    snprintf(c->err, sizeof(c->err), "gemm_tc_launch: unsupported shape Kdim=%d NT=%d npass=%d", a.Kdim, a.NT, a.npass);
    return ROREG_ERR_UNSUPPORTED;
  }
  const bool gather = a.g_C > 0;
  if (gather && (a.g_C % GM_KC || a.Kdim != 13 * a.g_C || a.g_ng < 1 || a.g_ng > 60 || !a.g_nei || a.a_cols || n_act_rows < 60)) {
    snprintf(c->err, sizeof(c->err), "gemm_tc_launch: unsupported implicit group convolution C=%d Kdim=%d n_g=%d", a.g_C, a.Kdim, a.g_ng);
    return ROREG_ERR_UNSUPPORTED;
  }
  CUtensorMap mAh, mAl, mWh, mWl;
  int rc;
  if (gather) {
    if ((rc = gemm_make_map(c, &mAh, A_hi, n_act_rows, a.g_C, 1))) return rc;
    if ((rc = gemm_make_map(c, &mAl, A_lo ? A_lo : A_hi, n_act_rows, a.g_C, 1))) return rc;
  } else {
    if ((rc = gemm_make_map(c, &mAh, A_hi, a.R, a.Kdim, GM_BM))) return rc;
    if ((rc = gemm_make_map(c, &mAl, A_lo ? A_lo : A_hi, a.R, a.Kdim, GM_BM))) return rc;
  }
  if ((rc = gemm_make_map(c, &mWh, W_hi, w_rows, a.Kdim, a.NT))) return rc;
  if ((rc = gemm_make_map(c, &mWl, W_lo ? W_lo : W_hi, w_rows, a.Kdim, a.NT))) return rc;
  static unsigned long long attr_mask = 0;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1>::SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel<3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<3>::SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1>::SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel<3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<3>::SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1>::SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<3>::SMEM_BYTES));
  }
  const int n_mt = (a.R + GM_BM - 1) / GM_BM;
  const long long tiles = (long long)n_mt * a.n_ntiles;
  const int grid = (int)(tiles < c->sm_count ? tiles : c->sm_count);
  // producer warps: plain mode needs one issuing lane (4 = the smallest instantiation); gather mode: ROREG_GEMM_PRODUCERS = 4 | 8 | 16
  static int gprod = 0;
  if (!gprod) { const char* e = getenv("ROREG_GEMM_PRODUCERS"); const int v = e ? atoi(e) : 8; gprod = (v == 4 || v == 16) ? v : 8; }
  // gathered operand: ROREG_GEMM_GATHER=tma selects the TMA tile::gather4 producer, default = the LDG / STS loader (8 producer warps)
  static int gldg = -1;
  if (gldg < 0) { const char* e = getenv("ROREG_GEMM_GATHER"); gldg = (e && !strcmp(e, "tma")) ? 0 : 1; }
  a.g_ldg = gather ? gldg : 0; a.g_act_hi = A_hi; a.g_act_lo = A_lo ? A_lo : A_hi;
  if (const char* e = getenv("ROREG_DEBUG_GEMM")) a.dbg = atoi(e);
  const int nprod = gather ? (a.g_ldg ? 8 : gprod) : 4;
#define GM_LAUNCH(NP, NPR) gemm_tc_kernel<NP, NPR><<<grid, gm_threads(NPR), GemmCfg<NP>::SMEM_BYTES, st>>>(mAh, mAl, mWh, mWl, a)
  if (a.npass == 3) { if (nprod == 4) GM_LAUNCH(3, 4); else if (nprod == 8) GM_LAUNCH(3, 8); else GM_LAUNCH(3, 16); }
  else              { if (nprod == 4) GM_LAUNCH(1, 4); else if (nprod == 8) GM_LAUNCH(1, 8); else GM_LAUNCH(1, 16); }
#undef GM_LAUNCH
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

}  // namespace roreg
