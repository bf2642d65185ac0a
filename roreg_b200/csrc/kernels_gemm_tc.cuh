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
// Pipeline: one box-producer lane (SWIZZLE_128B TMA boxes), eight loader warps that gather the implicit group convolution's
// rows with cp.async, one MMA-issuing lane, eight epilogue warps (roles and register budgets at GM_LOADERS); every operand
// image of a k-chunk is loaded ONCE per stage and all passes run from it; two TMEM accumulators of up to 256 columns so the
// epilogue of tile t overlaps the MMAs of tile t+1.
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
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 8 * 2048;      // stages | barriers | 8 epilogue staging blocks (the dynamic array is declared 1024-byte aligned: no slack needed)
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
  // g_cpasync != 0: the gathered operand is copied by the producer warps with cp.async (16 bytes per thread and row, needs the raw
  // activation pointers); 0: TMA tile::gather4 through the tensor maps.
  int g_cpasync; const float* g_act_hi; const float* g_act_lo;
  long long* trace;   // ROREG_DEBUG_GEMM_TRACE=<file>: clock64 stamps [1024 k-chunk iterations][8 events] of CTA 0 (one chosen launch)
};

// four rows of a 2-D tensor (box = 32 columns x 1 row) -> 4 x 128 B at dst, swizzled on the absolute shared-memory address
// exactly like a 4-row box (scripts/gather4_test.cu, profiles/r02_gather4_test.txt)
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// Warp roles, by warpgroup so that setmaxnreg can move registers to the epilogue (run c20: the same epilogue took 59 k clk per tile
// at 128 registers per thread and 104 k clk at 96, where ptxas spills 324 bytes):
//   warps 0-7   A loaders (implicit group convolution only; idle in plain mode)            40 registers
//   warps 8-15  epilogue, two per TMEM lane quadrant (warp id % 4), alternate 16-column chunks   168 registers
//   warp 16     MMA issuer (one lane); warp 17 box producer (one lane: W, and A in plain mode); 18-19 idle   56 registers
constexpr int GM_LOADERS = 8, GM_EPI_WARPS = 8, GM_EPI0 = GM_LOADERS, GM_MMA_WARP = GM_LOADERS + GM_EPI_WARPS, GM_BOX_WARP = GM_MMA_WARP + 1;
constexpr int GM_THREADS = (GM_LOADERS + GM_EPI_WARPS + 4) * 32;      // 640
constexpr int GM_STG_BYTES = 32 * 64;                                 // per epilogue warp: 32 rows x 16 floats
static_assert(GM_EPI_WARPS * GM_STG_BYTES == 8 * 2048, "GemmCfg::SMEM_BYTES reserves 8 staging blocks of 2 KB");

template <int NPASS>
__global__ void __launch_bounds__(GM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                                                         const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo,
                                                         GemmArgs a) {
  using Cfg = GemmCfg<NPASS>;
  constexpr int LANES = 32 / GM_LOADERS;                      // gather4 lanes per loader warp
  // Shared-memory budget: static (~1.4 KB, rounded up to the array's 1 KB alignment) + 192 KB of stages + barriers + 16 KB of
  // epilogue staging = 210.3 KiB.  That is past the 196 KiB carve-out step, i.e. the L1 is ~28 KB instead of ~60 KB: acceptable
  // since the epilogue no longer leans on L1 (residual / running-maximum rows come with 16-byte ld.cg, stores are staged); while it
  // did (round-2 run c7) crossing the step cost the all-pairs launches 45 %.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) { if (threadIdx.x == 0) printf("roreg: gemm_tc_kernel: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  // barriers: 0..3 full, 4..7 empty, 8..9 tmem_full, 10..11 tmem_empty
  __shared__ uint32_t tmem_base_s;
  __shared__ uint16_t acols_s[256];
  // byte tables (the CTA's total must stay under the 227 KiB limit)
  __shared__ uint8_t nei_s[60 * 13];
  __shared__ uint8_t gset_s[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const bool gather = a.g_C > 0;
#ifdef ROREG_GEMM_TRACE       // debugging builds only (nvcc -DROREG_GEMM_TRACE): the guard perturbs the MMA warp's code
#define GM_TRACE(i, e) do { if (a.trace && blockIdx.x == 0 && (i) < 1024u) a.trace[(size_t)(i) * 8 + (e)] = clock64(); } while (0)
#else
#define GM_TRACE(i, e) do { } while (0)
#endif
  for (int i = threadIdx.x; i < a.Kdim / GM_KC && i < 256; i += blockDim.x) acols_s[i] = (uint16_t)(a.a_cols ? a.a_cols[i] : i * GM_KC);   // < Kdim <= 8192
  if (gather) {
    for (int i = threadIdx.x; i < 60 * 13; i += blockDim.x) nei_s[i] = (uint8_t)a.g_nei[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) gset_s[i] = (uint8_t)((a.g_set && i < a.g_ng) ? a.g_set[i] : i);
  }
  if (threadIdx.x == 0) {
    // full[st]: one arrival from the box producer (arrive.expect_tx); with the cp.async loader also one per loader thread
    // (cp.async.mbarrier.arrive.noinc: the arrival is counted here, not added by the instruction)
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(BAR(s), (gather && a.g_cpasync) ? GM_LOADERS * 32 + 1 : 1); mbar_init(BAR(4 + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(8 + s), 1); mbar_init(BAR(10 + s), GM_EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GM_MMA_WARP) {
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
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.NT >> 3) << 17) | ((uint32_t)(GM_BM >> 4) << 24);

  if (warp < GM_LOADERS) {
    // ===================== A loaders (implicit group convolution) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const int cpk = gather ? a.g_C / GM_KC : 1;          // k-chunks per tap
    if (gather && a.g_cpasync) {
      // ---- cp.async loader.  Timeline of run c19 (clock64 trace, profiles/r02_gemm_timeline.txt): a tile::gather4 stage
      // completes ~4500 clk after its issue and the four stages in flight turn over every ~1400-1570 clk against ~800 clk of MMA
      // work - the TMA unit spends ~35 clk per gather4 (4 rows) where a tiled box costs ~2 clk per row.  A register-staged
      // LDG.128 -> STS.128 loader was slower still (its L2 round trip sits on the loader's critical path).  cp.async needs no
      // registers and no waiting: every loader thread copies the 16-byte chunk q = t % 8 of rows t / 8 + 32 j straight to
      // the SWIZZLE_128B position the MMA descriptor expects (row r at (r / 8) * 1024 + (r % 8) * 128, chunk q ^ (r % 8): a
      // warp instruction reads four whole 128-byte activation rows and writes four distinct 128-byte lines), then posts
      // cp.async.mbarrier.arrive.noinc on the stage's full barrier, which fires when its copies have landed.  The loaders run as
      // far ahead as the empty barriers allow (4 / 2 stages).
      constexpr int RSTEP = GM_LOADERS * 4, NJ = GM_BM / RSTEP;     // rows covered per pass of the loader threads; passes per tile
      const int tg = warp * 32 + lane;
      const int q = tg & 7, rl = tg >> 3;
      const uint32_t sw_off = (uint32_t)((rl >> 3) * 1024 + (rl & 7) * 128 + ((q ^ (rl & 7)) << 4));     // + (RSTEP / 8) * 1024 * j
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int mt = t / a.n_ntiles;
        int base[NJ], tap0[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          int r = mt * GM_BM + rl + RSTEP * j;
          if (r >= a.R) r = a.R - 1;                   // rows past the end: any valid row (their outputs are masked)
          const int item = r / a.g_ng, jj = r - item * a.g_ng;
          base[j] = item * 60; tap0[j] = (int)gset_s[jj] * 13;
        }
        int k_tap = 0, k_sub = 0;                        // kc = k_tap * cpk + k_sub
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int st = it % Cfg::STAGES; const uint32_t ph = (it / Cfg::STAGES) & 1;
          if (lane == 0) mbar_wait(BAR(4 + st), ph ^ 1);  // one poller per warp
          __syncwarp();
          if (tg == 0) GM_TRACE(it, 0);
          const uint32_t sb = smem_u32(smem + st * Cfg::STAGE_BYTES);
          const int c0 = k_sub * GM_KC + 4 * q;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const long long off = (long long)(base[j] + (int)nei_s[tap0[j] + k_tap]) * a.g_C + c0;
            const uint32_t dst = sb + sw_off + (uint32_t)(RSTEP / 8 * 1024 * j);
            cp_async16(dst, a.g_act_hi + off);
            if (NPASS == 3) cp_async16(dst + Cfg::OFF_ALO, a.g_act_lo + off);
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(BAR(st)) : "memory");
          if (tg == 0) GM_TRACE(it, 1);
          if (++k_sub == cpk) { k_sub = 0; ++k_tap; }
        }
      }
    } else if (gather) {
      // ---- TMA tile::gather4 loader (ROREG_GEMM_GATHER=tma, kept for A/B runs).  The instruction takes its four row coordinates
      // from uniform registers, so the compiler serialises lanes that hold different rows (ELECT + 6 x R2UR + UTMALDG per lane):
      // the 32 gathers of a stage are spread over the 8 loader warps, 4 lanes each.
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int mt = t / a.n_ntiles;
        int base[4] = {0, 0, 0, 0}, tap0[4] = {0, 0, 0, 0};
        if (lane < LANES) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int r = mt * GM_BM + 4 * (warp * LANES + lane) + i;
            if (r >= a.R) r = a.R - 1;
            const int item = r / a.g_ng, j = r - item * a.g_ng;
            base[i] = item * 60; tap0[i] = (int)gset_s[j] * 13;
          }
        }
        int k_tap = 0, k_sub = 0;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int st = it % Cfg::STAGES; const uint32_t ph = (it / Cfg::STAGES) & 1;
          if (lane == 0) mbar_wait(BAR(4 + st), ph ^ 1);
          __syncwarp();
          if (warp == 0 && lane == 0) GM_TRACE(it, 0);
          const uint32_t sb = smem_u32(smem + st * Cfg::STAGE_BYTES);
          if (lane < LANES) {
            const int c0 = k_sub * GM_KC;
            const int r0 = base[0] + (int)nei_s[tap0[0] + k_tap], r1 = base[1] + (int)nei_s[tap0[1] + k_tap];
            const int r2 = base[2] + (int)nei_s[tap0[2] + k_tap], r3 = base[3] + (int)nei_s[tap0[3] + k_tap];
            const uint32_t dst = sb + (warp * LANES + lane) * 512;
            tma_gather4(dst, &mapAhi, c0, r0, r1, r2, r3, BAR(st));
            if (NPASS == 3) tma_gather4(dst + Cfg::OFF_ALO, &mapAlo, c0, r0, r1, r2, r3, BAR(st));
          }
          if (warp == 0 && lane == 0) GM_TRACE(it, 1);
          if (++k_sub == cpk) { k_sub = 0; ++k_tap; }
        }
      }
    }
  } else if (warp >= GM_MMA_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == GM_BOX_WARP && lane == 0) {
      // ===================== box producer: W every stage; A too in plain mode =====================
      // expected bytes of a stage: everything that completes on the barrier by TMA (W; A unless cp.async brings it)
      const uint32_t stage_tx = Cfg::A_IMAGES * (w_bytes + ((gather && a.g_cpasync) ? 0u : (uint32_t)GM_A_BYTES));
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int mt = t / a.n_ntiles, nt = t % a.n_ntiles;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int st = it % Cfg::STAGES; const uint32_t ph = (it / Cfg::STAGES) & 1;
          mbar_wait(BAR(4 + st), ph ^ 1);
          const uint32_t sb = smem_u32(smem + st * Cfg::STAGE_BYTES);
          mbar_expect_tx(BAR(st), stage_tx);
          tma_load_2d(sb + Cfg::OFF_WHI, &mapWhi, kc * GM_KC, nt * a.NT, BAR(st));
          if (NPASS == 3) tma_load_2d(sb + Cfg::OFF_WLO, &mapWlo, kc * GM_KC, nt * a.NT, BAR(st));
          if (!gather) {
            tma_load_2d(sb, &mapAhi, (int)acols_s[kc], mt * GM_BM, BAR(st));
            if (NPASS == 3) tma_load_2d(sb + Cfg::OFF_ALO, &mapAlo, (int)acols_s[kc], mt * GM_BM, BAR(st));
          }
        }
      }
    } else if (warp == GM_MMA_WARP && lane == 0) {
      uint32_t it = 0, it_t = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it_t) {
        const int acc = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
        mbar_wait(BAR(10 + acc), tph ^ 1);
        GM_TRACE(it_t, 6);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int st = it % Cfg::STAGES; const uint32_t ph = (it / Cfg::STAGES) & 1;
          mbar_wait(BAR(st), ph);
          GM_TRACE(it, 2);
          if (gather && a.g_cpasync) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) writes -> tcgen05 reads
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
          GM_TRACE(it, 3);
        }
        umma_commit(BAR(8 + acc));
      }
    }
  } else {
    // ===================== epilogue =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int q = warp & 3, half = (warp - GM_EPI0) >> 2;     // TMEM lane quadrant of this warp (warp id % 4); chunk parity
    const int row_in_tile = q * 32 + lane;
    uint8_t* stg = smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256 + (warp - GM_EPI0) * GM_STG_BYTES;     // this warp's 32 x 64-byte store staging block
    // 16-byte accesses of whole chunks are possible when every row pitch is a multiple of 4 floats and every base is 16-byte aligned
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const bool vec_ok = !a.amax_arg && ((a.NT & 15) == 0) &&
                        (!a.bias || al16(a.bias)) && (!a.residual || (al16(a.residual) && (a.res_ld & 3) == 0)) &&
                        (!a.raw_out || (al16(a.raw_out) && (a.raw_ld & 3) == 0)) &&
                        (!a.act_hi || (al16(a.act_hi) && (a.act_ld & 3) == 0 && (!a.act_lo || al16(a.act_lo)))) &&
                        (!a.bn_scale || (al16(a.bn_scale) && al16(a.bn_shift)));
    uint32_t it_t = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it_t) {
      const int mt = t / a.n_ntiles, nt = t % a.n_ntiles;
      const int acc = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
      mbar_wait(BAR(8 + acc), tph);
      if (warp == GM_EPI0 && lane == 0) GM_TRACE(it_t, 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long r = (long long)mt * GM_BM + row_in_tile;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      // one 16-column chunk of the accumulator row (v) -> bias / residual / running max / BN / ReLU / split -> global
      auto process = [&](const uint32_t (&v)[16], const int c0) {
        const int o0 = nt * a.NT + c0;
        if (vec_ok && o0 + 16 <= a.O) {
          // Fast path (whole chunk in range, 16-byte aligned rows, no running maximum; warp-uniform): ~200 instructions.  The general
          // path below tests every element and loads every bias / BN parameter on its own - 1840 instructions per chunk (ncu source
          // view, run c23), which made the epilogue (60-76 k clk per tile) the limit of every layer with fewer than ~90 k-chunks.
          // Stores go through a 32-row x 64-byte staging block per warp so that four lanes write one row's 64 contiguous bytes (two
          // full sectors): thread-per-row stores are 32 half-sector writes per instruction, and with every SM in its epilogue
          // those (~100 sector writes per clock chip-wide) were what a 24 k-clk tile epilogue waited for (run c24 / c29).
          float x[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]);
          if (a.bias) {
            const float4* pb = reinterpret_cast<const float4*>(a.bias + o0);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float4 w = __ldg(pb + k); x[4 * k] += w.x; x[4 * k + 1] += w.y; x[4 * k + 2] += w.z; x[4 * k + 3] += w.w; }
          }
          if (a.residual && r < a.R) {
            const float4* pr = reinterpret_cast<const float4*>(a.residual + r * a.res_ld + o0);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float4 w = __ldcg(pr + k); x[4 * k] += w.x; x[4 * k + 1] += w.y; x[4 * k + 2] += w.z; x[4 * k + 3] += w.w; }
          }
          // rows of the warp's block -> global, transposed through the staging block (XOR on the 16-byte piece index: conflict-free
          // for the row-wise writes and for the 4-lanes-per-row reads)
          auto stage_store = [&](float* base, int ld, const float (&y)[16]) {
            float4* srow = reinterpret_cast<float4*>(stg + lane * 64);
#pragma unroll
            for (int k = 0; k < 4; ++k) srow[k ^ ((lane >> 1) & 3)] = make_float4(y[4 * k], y[4 * k + 1], y[4 * k + 2], y[4 * k + 3]);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int rr = 8 * k + (lane >> 2), pc = lane & 3;
              const float4 w = reinterpret_cast<const float4*>(stg + rr * 64)[pc ^ ((rr >> 1) & 3)];
              const long long grow = r - lane + rr;
              if (grow < a.R) *reinterpret_cast<float4*>(base + grow * ld + o0 + 4 * pc) = w;
            }
            __syncwarp();
          };
          if (a.raw_out) stage_store(a.raw_out, a.raw_ld, x);
          if (a.act_hi) {
            if (a.bn_scale) {
              const float4* ps = reinterpret_cast<const float4*>(a.bn_scale + o0);
              const float4* pt = reinterpret_cast<const float4*>(a.bn_shift + o0);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float4 sc = __ldg(ps + k), sh = __ldg(pt + k);
                x[4 * k] = fmaf(x[4 * k], sc.x, sh.x); x[4 * k + 1] = fmaf(x[4 * k + 1], sc.y, sh.y);
                x[4 * k + 2] = fmaf(x[4 * k + 2], sc.z, sh.z); x[4 * k + 3] = fmaf(x[4 * k + 3], sc.w, sh.w);
              }
            }
            float hi[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (a.relu) x[j] = fmaxf(x[j], 0.f);
              uint32_t tb; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tb) : "f"(x[j]));
              hi[j] = __uint_as_float(tb);
            }
            stage_store(a.act_hi, a.act_ld, hi);
            if (a.act_lo) {
#pragma unroll
              for (int j = 0; j < 16; ++j) x[j] -= hi[j];
              stage_store(a.act_lo, a.act_ld, x);
            }
          }
        } else if (r < a.R) {
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
      };
      // The two warps of a lane quadrant take alternate chunks; the tcgen05.ld of a warp's next chunk is in flight while it works on
      // the current one (run c19: 84 k clk of epilogue per 128 x 256 tile with four warps and serial loads - more than the tile's
      // 78 k clk of MMA work in the 256-channel layers).
#define GM_LDTM(V, C0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                     : "=r"(V[0]), "=r"(V[1]), "=r"(V[2]), "=r"(V[3]), "=r"(V[4]), "=r"(V[5]), "=r"(V[6]), "=r"(V[7]), \
                       "=r"(V[8]), "=r"(V[9]), "=r"(V[10]), "=r"(V[11]), "=r"(V[12]), "=r"(V[13]), "=r"(V[14]), "=r"(V[15]) \
                     : "r"(taddr + (C0)) : "memory")
      uint32_t va[16], vb[16];
      int c0 = 16 * half;
      if (c0 < a.NT) GM_LDTM(va, c0);
      while (c0 < a.NT) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 < a.NT) GM_LDTM(vb, c0 + 32);
        process(va, c0);
        c0 += 32;
        if (c0 >= a.NT) break;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 < a.NT) GM_LDTM(va, c0 + 32);
        process(vb, c0);
        c0 += 32;
      }
#undef GM_LDTM
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(BAR(10 + acc));
      if (warp == GM_EPI0 && lane == 0) GM_TRACE(it_t, 5);
    }
  }
#undef GM_TRACE
  __syncthreads();
  if (warp == GM_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
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
    RR_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1>::SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<3>::SMEM_BYTES));
  }
  const int n_mt = (a.R + GM_BM - 1) / GM_BM;
  const long long tiles = (long long)n_mt * a.n_ntiles;
  const int grid = (int)(tiles < c->sm_count ? tiles : c->sm_count);
  // gathered operand: cp.async loader by default; ROREG_GEMM_GATHER=tma selects the TMA tile::gather4 producer (A/B runs)
  const char* gsel = gather ? getenv("ROREG_GEMM_GATHER") : nullptr;        // read per launch: the tests flip it
  a.g_cpasync = (gather && !(gsel && !strcmp(gsel, "tma"))) ? 1 : 0; a.g_act_hi = A_hi; a.g_act_lo = A_lo ? A_lo : A_hi;
  // timeline of one gather launch (debugging only): ROREG_DEBUG_GEMM_TRACE=<file>, ROREG_DEBUG_GEMM_TRACE_LAUNCH=<index among the gather launches>
  static int trace_seen = 0; long long* d_trace = nullptr; const char* trace_fn = getenv("ROREG_DEBUG_GEMM_TRACE");
  if (trace_fn && gather) {
    const char* e = getenv("ROREG_DEBUG_GEMM_TRACE_LAUNCH");
    if (trace_seen++ == (e ? atoi(e) : 0)) { RR_CUDA(c, cudaMalloc(&d_trace, 1024 * 8 * sizeof(long long))); RR_CUDA(c, cudaMemset(d_trace, 0, 1024 * 8 * sizeof(long long))); }
  }
  a.trace = d_trace;
  if (a.npass == 3) gemm_tc_kernel<3><<<grid, GM_THREADS, GemmCfg<3>::SMEM_BYTES, st>>>(mAh, mAl, mWh, mWl, a);
  else              gemm_tc_kernel<1><<<grid, GM_THREADS, GemmCfg<1>::SMEM_BYTES, st>>>(mAh, mAl, mWh, mWl, a);
  RR_LAUNCH_CHECK(c);
  if (d_trace) {
    RR_CUDA(c, cudaStreamSynchronize(st));
    static long long h[1024 * 8];
    RR_CUDA(c, cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost)); cudaFree(d_trace);
    if (FILE* f = fopen(trace_fn, "w")) {
      fprintf(f, "# R=%d Kdim=%d O=%d NT=%d npass=%d C=%d ng=%d cpasync=%d grid=%d\n", a.R, a.Kdim, a.O, a.NT, a.npass, a.g_C, a.g_ng, a.g_cpasync, grid);
      for (int i = 0; i < 1024; ++i) { for (int e2 = 0; e2 < 8; ++e2) fprintf(f, "%lld ", h[i * 8 + e2]); fprintf(f, "\n"); }
      fclose(f);
    }
  }
  return ROREG_OK;
}

}  // namespace roreg
