// kernels_allpairs_tc.cuh - all-pairs 60-rotation group correlation as ONE persistent tcgen05 kernel (north_star kernel 1).
//
//   best[n][m] = max_a cor_a(n,m),  best_a[n][m] = argmax_a (first maximum),
//   cor_a(n,m) = sum_{f,g} X[n,f,P[a,g]] * Y[m,f,g]                      (test/estimator.py:85-89 on EVERY pair (n,m))
//
// With channel-last descriptors [(n*60+g)][32] a keypoint is one K = 1920 row and rotation `a` is a permutation of its sixty
// 32-float chunks: the A operand of k-chunk g is X's chunk P[a][g] - applied by the TMA column coordinate alone.  Round 1
// launched 60 GEMMs whose epilogues read-modify-wrote best / best_a in HBM.  Here a CTA owns a 128 x 256 tile of (n, m) for all
// 60 rotations: the rotation loop runs INSIDE the kernel, each rotation's accumulator (TMEM, double-buffered) is compared against
// a running (max, argmax) held in REGISTERS (thread = one row x 128 columns: 128 floats + 32 packed index words), and the tile is
// written once.  Operand tiles are TMA-staged (SWIZZLE_128B boxes, mbarrier ring), the same single-load-per-k-chunk stages
// as kernels_gemm_tc.cuh (1 pass: A | W x 4 stages; 3xTF32: A_hi | A_lo | W_hi | W_lo x 2 stages).
//   warp 0      TMA producer (one lane)        warp 1   MMA issuer (one lane)
//   warps 2-5   epilogue, columns   0..127 of the tile (TMEM lane quadrant = warp % 4)
//   warps 6-9   epilogue, columns 128..255
// Bound: L2 -> SM operand traffic, as for every 128 x 256 fp32 tile: both operands are re-streamed for each rotation.
#pragma once
#include "kernels_gemm_tc.cuh"

namespace roreg {

constexpr int AP_THREADS = 10 * 32;
struct RotCols { uint8_t p[3600]; };             // P[a][g]: which 32-float chunk of X pairs with chunk g of Y under rotation a

struct AllPairsArgs {
  int N, M, npass;
  float* best; uint8_t* best_a;                  // [N][M]
};

template <int NPASS>
__global__ void __launch_bounds__(AP_THREADS, 1) allpairs_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                                                                    const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo,
                                                                    const __grid_constant__ RotCols rot, AllPairsArgs a) {
  using Cfg = GemmCfg<NPASS>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) { if (threadIdx.x == 0) printf("roreg: allpairs_tc_kernel: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  // barriers: 0..3 full, 4..7 empty, 8..9 tmem_full, 10..11 tmem_empty
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(BAR(s), 1); mbar_init(BAR(4 + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(8 + s), 1); mbar_init(BAR(10 + s), 256); }
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

  const int n_mt = (a.N + 127) / 128, n_nt = (a.M + 255) / 256;
  const int n_tiles = n_mt * n_nt;
  constexpr int NKC = 60;
  constexpr uint32_t stage_tx = Cfg::A_IMAGES * (GM_A_BYTES + GM_W_BYTES);
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int mt = t / n_nt, nt = t % n_nt;
        for (int r = 0; r < 60; ++r)
          for (int kc = 0; kc < NKC; ++kc, ++it) {
            const int st = it % Cfg::STAGES; const uint32_t ph = (it / Cfg::STAGES) & 1;
            mbar_wait(BAR(4 + st), ph ^ 1);
            mbar_expect_tx(BAR(st), stage_tx);
            const uint32_t sb = smem_u32(smem + st * Cfg::STAGE_BYTES);
            const int acol = (int)rot.p[r * 60 + kc] * GM_KC;
            tma_load_2d(sb, &mapAhi, acol, mt * 128, BAR(st));
            tma_load_2d(sb + Cfg::OFF_WHI, &mapWhi, kc * GM_KC, nt * 256, BAR(st));
            if (NPASS == 3) {
              tma_load_2d(sb + Cfg::OFF_ALO, &mapAlo, acol, mt * 128, BAR(st));
              tma_load_2d(sb + Cfg::OFF_WLO, &mapWlo, kc * GM_KC, nt * 256, BAR(st));
            }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, ia = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
        for (int r = 0; r < 60; ++r, ++ia) {
          const int acc = ia & 1; const uint32_t tph = (ia >> 1) & 1;
          mbar_wait(BAR(10 + acc), tph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_tmem = tmem_base + acc * 256;
          for (int kc = 0; kc < NKC; ++kc, ++it) {
            const int st = it % Cfg::STAGES; const uint32_t ph = (it / Cfg::STAGES) & 1;
            mbar_wait(BAR(st), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(smem + st * Cfg::STAGE_BYTES), whi = sa + Cfg::OFF_WHI;
#pragma unroll
            for (int kk = 0; kk < GM_KC / 8; ++kk)
              umma_tf32(d_tmem, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(whi + kk * 32), idesc, (kc | kk) ? 1u : 0u);
            if (NPASS == 3) {
              const uint32_t alo = sa + Cfg::OFF_ALO, wlo = sa + Cfg::OFF_WLO;
#pragma unroll
              for (int kk = 0; kk < GM_KC / 8; ++kk) umma_tf32(d_tmem, umma_desc_sw128(alo + kk * 32), umma_desc_sw128(whi + kk * 32), idesc, 1u);
#pragma unroll
              for (int kk = 0; kk < GM_KC / 8; ++kk) umma_tf32(d_tmem, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(wlo + kk * 32), idesc, 1u);
            }
            umma_commit(BAR(4 + st));
          }
          umma_commit(BAR(8 + acc));
        }
    }
  } else {
    // ===================== epilogue: running (max, argmax) over the 60 rotations in registers =====================
    const int q = warp & 3;                              // TMEM lane quadrant
    const int chalf = (warp - 2) >> 2;                   // 0: columns 0..127, 1: columns 128..255 of the tile
    const int row_in_tile = q * 32 + lane;
    uint32_t ia = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int mt = t / n_nt, nt = t % n_nt;
      float best[128]; uint32_t argw[32];
#pragma unroll
      for (int c = 0; c < 128; ++c) best[c] = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) argw[c] = 0u;
      for (int r = 0; r < 60; ++r, ++ia) {
        const int acc = ia & 1; const uint32_t tph = (ia >> 1) & 1;
        mbar_wait(BAR(8 + acc), tph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + chalf * 128;
        const uint32_t rb = (uint32_t)r * 0x01010101u;   // the rotation index in every byte
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 16) {
          uint32_t v[16];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                         "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(taddr + c0) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = c0 + j;
            const float x = __uint_as_float(v[j]);
            if (x > best[c]) {                            // strict '>': the first maximal rotation is kept (torch.argmax)
              best[c] = x;
              const uint32_t m = 0xFFu << (8 * (c & 3));
              argw[c >> 2] = (argw[c >> 2] & ~m) | (rb & m);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(BAR(10 + acc));
      }
      // the tile's outputs, once
      const long long row = (long long)mt * 128 + row_in_tile;
      const int col0 = nt * 256 + chalf * 128;
      if (row < a.N) {
        float* pb = a.best + row * a.M + col0; uint8_t* pa = a.best_a + row * a.M + col0;
        const bool vec = ((a.M & 3) == 0) && (col0 + 127 < a.M);
        if (vec) {
#pragma unroll
          for (int c = 0; c < 128; c += 4) *reinterpret_cast<float4*>(pb + c) = make_float4(best[c], best[c + 1], best[c + 2], best[c + 3]);
          if ((a.M & 15) == 0) {                          // byte rows: 16-byte stores only when the row pitch keeps them aligned
#pragma unroll
            for (int c = 0; c < 32; c += 4) *reinterpret_cast<uint4*>(pa + 4 * c) = make_uint4(argw[c], argw[c + 1], argw[c + 2], argw[c + 3]);
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) *reinterpret_cast<uint32_t*>(pa + 4 * c) = argw[c];
          }
        } else {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (col0 + c < a.M) { pb[c] = best[c]; pa[c] = (uint8_t)((argw[c >> 2] >> (8 * (c & 3))) & 0xFFu); }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// X_hi/X_lo: [N][1920], Y_hi/Y_lo: [M][1920] channel-last tf32-split descriptors (roreg_pack_descriptors); perm8 = P[a][g] (host copy)
static inline int allpairs_tc_launch(roreg_ctx* c, const float* X_hi, const float* X_lo, int N, const float* Y_hi, const float* Y_lo, int M,
                                     int npass, float* best, uint8_t* best_a, cudaStream_t st) {
  CUtensorMap mAh, mAl, mWh, mWl;
  int rc;
  if ((rc = gemm_make_map(c, &mAh, X_hi, N, 1920, 128))) return rc;
  if ((rc = gemm_make_map(c, &mAl, X_lo ? X_lo : X_hi, N, 1920, 128))) return rc;
  if ((rc = gemm_make_map(c, &mWh, Y_hi, M, 1920, 256))) return rc;
  if ((rc = gemm_make_map(c, &mWl, Y_lo ? Y_lo : Y_hi, M, 1920, 256))) return rc;
  static unsigned long long attr_mask = 0;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(allpairs_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1>::SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(allpairs_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<3>::SMEM_BYTES));
  }
  RotCols rot;
  memcpy(rot.p, c->h_perm8, 3600);
  AllPairsArgs a{N, M, npass, best, best_a};
  const long long tiles = (long long)((N + 127) / 128) * ((M + 255) / 256);
  const int grid = (int)(tiles < c->sm_count ? tiles : c->sm_count);
  if (npass == 3) allpairs_tc_kernel<3><<<grid, AP_THREADS, GemmCfg<3>::SMEM_BYTES, st>>>(mAh, mAl, mWh, mWl, rot, a);
  else allpairs_tc_kernel<1><<<grid, AP_THREADS, GemmCfg<1>::SMEM_BYTES, st>>>(mAh, mAl, mWh, mWl, rot, a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

}  // namespace roreg
