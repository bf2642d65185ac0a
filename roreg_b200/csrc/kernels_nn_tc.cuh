// kernels_nn_tc.cuh - mutual-NN search as a tensor-core Gram (nn mode 1).
//
// d2(i,j) = |a_i|^2 + |b_j|^2 - 2 a_i.b_j with the 32-d inner product on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM).  One TF32 pass would give ~1e-3 absolute error on the
// Gram - too coarse for the 1e-4 distance tolerance - so the operands are split x = hi + lo
// (hi = tf32(x), lo = x - hi, exact) and the contraction is run over an extended K = 96:
//     A' = [a_hi | a_lo | a_hi],  B' = [b_hi | b_hi | b_lo]   =>   A'.B'^T = hi.hi + lo.hi + hi.lo
// which drops only the lo.lo term (~2^-22 relative).  The [N,M] distance matrix is never materialised:
// the epilogue warps read each 128x128 accumulator tile straight from TMEM and keep a running
// (min, argmin) per row in registers.  The column direction is the same kernel with the operand roles
// swapped (blockIdx-independent work item = (pair, direction, 128-row block)).
//
// Pipeline (warp-specialised, persistent, one CTA per SM):
//   warp 0      TMA producer   cp.async.bulk.tensor.2d, SWIZZLE_128B boxes [128 rows x 32 floats]
//   warp 1      MMA issuer     12 x tcgen05.mma (M=128,N=128,K=8) per tile, tcgen05.commit -> mbarriers
//   warps 2-5   epilogue       tcgen05.ld 32x32b.x32 -> running argmin; TMEM double-buffered (2 x 128 cols)
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "kernels_match.cuh"

namespace roreg {

constexpr int TC_BM = 128, TC_BN = 128, TC_KC = 32, TC_NCHUNK = 3, TC_KEXT = TC_KC * TC_NCHUNK;   // K' = 96
constexpr int TC_STAGES = 3;
constexpr int TC_BOX_BYTES = TC_BM * TC_KC * 4;                  // 16 KB per [128 x 32 f32] box
constexpr int TC_TILE_BYTES = TC_BOX_BYTES * TC_NCHUNK;          // 48 KB per operand tile
constexpr int TC_SMEM_BYTES = TC_TILE_BYTES * (1 + TC_STAGES) + 2 * TC_BN * 4 + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr long long TC_WAIT_CYCLES = 3000000000LL;     // ~1.5-2 s of SM clocks: a lost arrive becomes a trap, never a hung GPU

// ---- prep: split the pooled features into the extended-K operands ------------------------------------
__global__ void __launch_bounds__(256) nn_tc_prep_kernel(const float* __restrict__ inv, int rows,
                                                         float* __restrict__ Ahat, float* __restrict__ Bhat,
                                                         float* __restrict__ nrm_half) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float x = inv[(long long)r * 32 + lane];
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
  const float hi = __uint_as_float(hb);
  const float lo = x - hi;
  float* a = Ahat + (long long)r * TC_KEXT; float* b = Bhat + (long long)r * TC_KEXT;
  a[lane] = hi; a[32 + lane] = lo; a[64 + lane] = hi;
  b[lane] = hi; b[32 + lane] = hi; b[64 + lane] = lo;
  const float ss = warp_sum(x * x);
  if (lane == 0) nrm_half[r] = 0.5f * ss;
}

// ---- small PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > TC_WAIT_CYCLES) __trap();     // never hang the GPU: a lost arrive becomes a launch failure
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_128B canonical layout: 8-row atoms of 1024 B (SBO = 1024), LBO unused (=1), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, A/B K-major, D = f32, M = 128, N = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

struct NNTcArgs {
  const float* nrm_half;      // [rows] 0.5*|x|^2
  int S, B;                   // rows per (pair, side); pairs
  int32_t* nn01; int32_t* nn10;  // [B][S]
  int passes;                  // 3 = hi.hi + lo.hi + hi.lo (default); 1 / 2 only for bottleneck experiments (ROREG_DEBUG_NN_PASSES)
};

__global__ void __launch_bounds__(192, 1) nn_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                       const __grid_constant__ CUtensorMap mapB, NNTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // 48 KB
  uint8_t* sB = smem + TC_TILE_BYTES;                   // TC_STAGES x 48 KB
  float* sNb = reinterpret_cast<float*>(smem + TC_TILE_BYTES * (1 + TC_STAGES));   // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sNb + 2 * TC_BN);
  // barrier map: 0 a_full, 1 a_empty, 2..4 b_full, 5..7 b_empty, 8..9 tmem_full, 10..11 tmem_empty
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1); mbar_init(BAR(1), 1);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(BAR(2 + s), 1); mbar_init(BAR(5 + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(8 + s), 1); mbar_init(BAR(10 + s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const int nrb = (a.S + TC_BM - 1) / TC_BM;            // row blocks per (pair, dir)
  const int nct = (a.S + TC_BN - 1) / TC_BN;            // column tiles
  const int n_items = a.B * 2 * nrb;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it_b = 0, a_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rb = item % nrb, d = (item / nrb) & 1, p = item / (2 * nrb);
        const int a_row0 = (p * 2 + d) * a.S + rb * TC_BM;
        const int b_row_base = (p * 2 + (1 - d)) * a.S;
        mbar_wait(BAR(1), a_phase ^ 1);                 // A tile free (previous item's MMAs retired)
        mbar_expect_tx(BAR(0), TC_TILE_BYTES);
        for (int c = 0; c < TC_NCHUNK; ++c) tma_load_2d(smem_u32(sA + c * TC_BOX_BYTES), &mapA, c * TC_KC, a_row0, BAR(0));
        a_phase ^= 1;
        for (int ct = 0; ct < nct; ++ct, ++it_b) {
          const int st = it_b % TC_STAGES; const uint32_t ph = (it_b / TC_STAGES) & 1;
          mbar_wait(BAR(5 + st), ph ^ 1);
          mbar_expect_tx(BAR(2 + st), TC_TILE_BYTES);
          for (int c = 0; c < TC_NCHUNK; ++c)
            tma_load_2d(smem_u32(sB + st * TC_TILE_BYTES + c * TC_BOX_BYTES), &mapB, c * TC_KC, b_row_base + ct * TC_BN, BAR(2 + st));
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it_b = 0, it_t = 0, a_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(BAR(0), a_phase); a_phase ^= 1;
        for (int ct = 0; ct < nct; ++ct, ++it_b, ++it_t) {
          const int st = it_b % TC_STAGES; const uint32_t ph = (it_b / TC_STAGES) & 1;
          const int acc = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
          mbar_wait(BAR(2 + st), ph);                   // B tile landed
          mbar_wait(BAR(10 + acc), tph ^ 1);            // accumulator drained by the epilogue
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_tmem = tmem_base + acc * TC_BN;
#pragma unroll
          for (int c = 0; c < TC_NCHUNK; ++c)
#pragma unroll
            for (int kk = 0; kk < TC_KC / 8; ++kk) {
              const uint64_t ad = umma_desc_sw128(smem_u32(sA + c * TC_BOX_BYTES) + kk * 32);
              const uint64_t bd = umma_desc_sw128(smem_u32(sB + st * TC_TILE_BYTES + c * TC_BOX_BYTES) + kk * 32);
              umma_tf32(d_tmem, ad, bd, TC_IDESC, (c | kk) ? 1u : 0u);
            }
          umma_commit(BAR(5 + st));                     // smem stage reusable once these MMAs retire
          umma_commit(BAR(8 + acc));                    // accumulator ready for the epilogue
        }
        umma_commit(BAR(1));                            // A tile reusable
      }
    }
  } else {
    // ===================== epilogue: running (min, argmin) per row =====================
    const int q = warp & 3;                             // TMEM lane quadrant this warp may read
    const int row_in_tile = q * 32 + lane;
    const int et = threadIdx.x - 64;                    // 0..127
    uint32_t it_t = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rb = item % nrb, d = (item / nrb) & 1, p = item / (2 * nrb);
      const int b_row_base = (p * 2 + (1 - d)) * a.S;
      float best_v = INFINITY; int best_j = 0x7fffffff;
      for (int ct = 0; ct < nct; ++ct, ++it_t) {
        const int acc = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
        // column half-norms of this tile (+inf for columns past the end of the target cloud)
        {
          const int j = ct * TC_BN + et;
          sNb[acc * TC_BN + et] = (j < a.S) ? a.nrm_half[b_row_base + j] : INFINITY;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(BAR(8 + acc), tph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TC_BN;
#pragma unroll 1
        for (int cc = 0; cc < TC_BN / 32; ++cc) {
          uint32_t r[32];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                       "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                         "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                         "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                         "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                       : "r"(taddr + cc * 32) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const float* nb = sNb + acc * TC_BN + cc * 32;
          float v[32]; float m = INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = nb[j] - __uint_as_float(r[j]); m = fminf(m, v[j]); }   // (d2 - |a|^2)/2
          if (m < best_v) {
            int jj = 31;
#pragma unroll
            for (int j = 31; j >= 0; --j) if (v[j] == m) jj = j;       // first column attaining the chunk minimum
            best_v = m; best_j = ct * TC_BN + cc * 32 + jj;
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(BAR(10 + acc));                     // this thread is done with the accumulator
      }
      const int row = rb * TC_BM + row_in_tile;
      if (row < a.S) (d == 0 ? a.nn01 : a.nn10)[(long long)p * a.S + row] = best_j;
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
  }
}

// ---- host side: tensor maps + launch -------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline int nn_tc_make_map(roreg_ctx* c, CUtensorMap* m, const float* base, long long rows) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !p || qres != cudaDriverEntryPointSuccess) {
      snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled entry point unavailable");
      return ROREG_ERR_CUDA;
    }
    fn = (PFN_encodeTiled)p;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)TC_KEXT, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)TC_KEXT * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)TC_KC, (cuuint32_t)TC_BM};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled failed (%d)", (int)r); return ROREG_ERR_CUDA; }
  return ROREG_OK;
}

// Both directions of B pairs in one launch.  inv: [B][2][S][32] pooled features; Ahat/Bhat/nrm: workspace.
static inline int nn_tc_launch_both(roreg_ctx* c, const float* inv, int S, int B, float* Ahat, float* Bhat, float* nrm_half,
                                    int32_t* nn01, int32_t* nn10, cudaStream_t st) {
  const long long rows = (long long)B * 2 * S;
  nn_tc_prep_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(inv, (int)rows, Ahat, Bhat, nrm_half);
  RR_LAUNCH_CHECK(c);
  CUtensorMap mA, mB;
  int rc;
  if ((rc = nn_tc_make_map(c, &mA, Ahat, rows))) return rc;
  if ((rc = nn_tc_make_map(c, &mB, Bhat, rows))) return rc;
  static unsigned long long attr_mask = 0;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(nn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  }
  const int nrb = (S + TC_BM - 1) / TC_BM;
  const int items = B * 2 * nrb;
  const int grid = items < c->sm_count ? items : c->sm_count;
  NNTcArgs a{nrm_half, S, B, nn01, nn10, 3};
  nn_tc_kernel<<<grid, 192, TC_SMEM_BYTES, st>>>(mA, mB, a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}


// =====================================================================================================
// v2 (nn mode 2).  ncu on v1 (profiles/r01_run3_*): tensor pipe 36 % active, L2 12 %, one epilogue warp per
// SMSP serialised on  tcgen05.ld -> wait -> 32 columns of math  => the EPILOGUE, not the MMA, paces the tile.
// v2 keeps the tile shape and changes what v1 was slow at:
//   * 8 epilogue warps (two per TMEM lane quadrant, each owning 64 of the 128 columns); both 32-column
//     tcgen05.ld of a tile are issued back to back and waited once; the accumulator is released as soon as
//     it sits in registers, before the min/argmin math;
//   * the min is taken first (FADD+FMNMX per element), the index is only searched when the chunk improves
//     the running minimum (rare after the first tiles);
//   * one operand array H[rows][64] = [hi | lo]: the three products (hi,hi),(lo,hi),(hi,lo) address chunks
//     of the same tiles, so an operand tile is 2 TMA boxes (32 KB) instead of 3, and 4 stages fit.
// =====================================================================================================
constexpr int T2_STAGES = 4;
constexpr int T2_TILE_BYTES = 2 * TC_BOX_BYTES;                                  // [hi | lo] = 32 KB
constexpr int T2_SMEM_BYTES = T2_TILE_BYTES * (1 + T2_STAGES) + 2 * TC_BN * 4 + 128 * 8 + 1024 + 256;
constexpr int T2_THREADS = 64 + 256;

__global__ void __launch_bounds__(256) nn_tc2_prep_kernel(const float* __restrict__ inv, int rows,
                                                          float* __restrict__ H, float* __restrict__ nrm_half) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float x = inv[(long long)r * 32 + lane];
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
  const float hi = __uint_as_float(hb);
  H[(long long)r * 64 + lane] = hi;
  H[(long long)r * 64 + 32 + lane] = x - hi;
  const float ss = warp_sum(x * x);
  if (lane == 0) nrm_half[r] = 0.5f * ss;
}

#define RR_TMEM_LD32(rr, addr)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
               : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),     \
                 "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15]), \
                 "=r"(rr[16]), "=r"(rr[17]), "=r"(rr[18]), "=r"(rr[19]), "=r"(rr[20]), "=r"(rr[21]), "=r"(rr[22]), "=r"(rr[23]), \
                 "=r"(rr[24]), "=r"(rr[25]), "=r"(rr[26]), "=r"(rr[27]), "=r"(rr[28]), "=r"(rr[29]), "=r"(rr[30]), "=r"(rr[31])  \
               : "r"(addr) : "memory")

__global__ void __launch_bounds__(T2_THREADS, 1) nn_tc2_kernel(const __grid_constant__ CUtensorMap mapH, NNTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // 32 KB
  uint8_t* sB = smem + T2_TILE_BYTES;                   // T2_STAGES x 32 KB
  float* sNb = reinterpret_cast<float*>(smem + T2_TILE_BYTES * (1 + T2_STAGES));   // [2][128]
  float* mrg_v = sNb + 2 * TC_BN; int* mrg_j = reinterpret_cast<int*>(mrg_v + 128);  // [128] each
  uint64_t* bars = reinterpret_cast<uint64_t*>(mrg_j + 128);
  // barriers: 0 a_full, 1 a_empty, 2..5 b_full, 6..9 b_empty, 10..11 tmem_full, 12..13 tmem_empty
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sNbW[8 * 2 * 64];      // [epilogue warp][tile parity][64 columns]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1); mbar_init(BAR(1), 1);
    for (int s = 0; s < T2_STAGES; ++s) { mbar_init(BAR(2 + s), 1); mbar_init(BAR(6 + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(10 + s), 1); mbar_init(BAR(12 + s), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const int nrb = (a.S + TC_BM - 1) / TC_BM;
  const int nct = (a.S + TC_BN - 1) / TC_BN;
  const int n_items = a.B * 2 * nrb;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it_b = 0, a_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rb = item % nrb, d = (item / nrb) & 1, p = item / (2 * nrb);
        const int a_row0 = (p * 2 + d) * a.S + rb * TC_BM;
        const int b_row_base = (p * 2 + (1 - d)) * a.S;
        mbar_wait(BAR(1), a_phase ^ 1);
        mbar_expect_tx(BAR(0), T2_TILE_BYTES);
        for (int c = 0; c < 2; ++c) tma_load_2d(smem_u32(sA + c * TC_BOX_BYTES), &mapH, c * TC_KC, a_row0, BAR(0));
        a_phase ^= 1;
        for (int ct = 0; ct < nct; ++ct, ++it_b) {
          const int st = it_b % T2_STAGES; const uint32_t ph = (it_b / T2_STAGES) & 1;
          mbar_wait(BAR(6 + st), ph ^ 1);
          mbar_expect_tx(BAR(2 + st), T2_TILE_BYTES);
          for (int c = 0; c < 2; ++c)
            tma_load_2d(smem_u32(sB + st * T2_TILE_BYTES + c * TC_BOX_BYTES), &mapH, c * TC_KC, b_row_base + ct * TC_BN, BAR(2 + st));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it_b = 0, it_t = 0, a_phase = 0;
      const uint32_t ahi = smem_u32(sA), alo = ahi + TC_BOX_BYTES;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(BAR(0), a_phase); a_phase ^= 1;
        for (int ct = 0; ct < nct; ++ct, ++it_b, ++it_t) {
          const int st = it_b % T2_STAGES; const uint32_t ph = (it_b / T2_STAGES) & 1;
          const int par = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
          mbar_wait(BAR(2 + st), ph);
          mbar_wait(BAR(12 + par), tph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t bhi = smem_u32(sB + st * T2_TILE_BYTES), blo = bhi + TC_BOX_BYTES;
          const uint32_t d_tmem = tmem_base + par * TC_BN;
          const uint32_t aop[3] = {ahi, alo, ahi}, bop[3] = {bhi, bhi, blo};
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c < a.passes) {
#pragma unroll
              for (int kk = 0; kk < TC_KC / 8; ++kk)
                umma_tf32(d_tmem, umma_desc_sw128(aop[c] + kk * 32), umma_desc_sw128(bop[c] + kk * 32), TC_IDESC, (c | kk) ? 1u : 0u);
            }
          umma_commit(BAR(6 + st));
          umma_commit(BAR(10 + par));
        }
        umma_commit(BAR(1));
      }
    }
  } else {
    // ===================== epilogue: 8 warps, warp -> (lane quadrant q, column half hf) =====================
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;                     // warps 2..5 -> columns 0..63, warps 6..9 -> 64..127
    const int row_in_tile = q * 32 + lane;
    uint32_t it_t = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rb = item % nrb, d = (item / nrb) & 1, p = item / (2 * nrb);
      const int b_row_base = (p * 2 + (1 - d)) * a.S;
      float best_v = INFINITY; int best_j = 0x7fffffff;
      for (int ct = 0; ct < nct; ++ct, ++it_t) {
        const int par = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
        // per-warp staging of the 64 column half-norms this warp needs (no CTA-wide barrier per tile)
        float* nbw = sNbW + ((warp - 2) * 2 + par) * 64;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = ct * TC_BN + hf * 64 + u * 32 + lane;
          nbw[u * 32 + lane] = (j < a.S) ? __ldg(a.nrm_half + b_row_base + j) : INFINITY;
        }
        __syncwarp();
        mbar_wait(BAR(10 + par), tph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + par * TC_BN + hf * 64;
        uint32_t r0[32], r1[32];
        RR_TMEM_LD32(r0, taddr);
        RR_TMEM_LD32(r1, taddr + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(BAR(12 + par));                     // accumulator free: the MMA of tile t+2 may start
        const float4* nb4 = reinterpret_cast<const float4*>(nbw);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t* rg = half ? r1 : r0;
          float m = INFINITY;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 nb = nb4[half * 8 + j4];
            m = fminf(m, fminf(fminf(nb.x - __uint_as_float(rg[4 * j4]), nb.y - __uint_as_float(rg[4 * j4 + 1])),
                               fminf(nb.z - __uint_as_float(rg[4 * j4 + 2]), nb.w - __uint_as_float(rg[4 * j4 + 3]))));
          }
          if (m < best_v) {                             // rare: find the first column attaining the new minimum
            const float* nbs = nbw + half * 32;
            int jj = 31;
#pragma unroll
            for (int j = 31; j >= 0; --j) if (nbs[j] - __uint_as_float(rg[j]) == m) jj = j;
            best_v = m; best_j = ct * TC_BN + hf * 64 + half * 32 + jj;
          }
        }
      }
      // merge the two column halves of each row (lexicographic (value, index); half 1 has the larger indices)
      if (hf == 1) { mrg_v[row_in_tile] = best_v; mrg_j[row_in_tile] = best_j; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (hf == 0) {
        if (mrg_v[row_in_tile] < best_v) { best_v = mrg_v[row_in_tile]; best_j = mrg_j[row_in_tile]; }
        const int row = rb * TC_BM + row_in_tile;
        if (row < a.S) (d == 0 ? a.nn01 : a.nn10)[(long long)p * a.S + row] = best_j;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

static inline int nn_tc2_launch_both(roreg_ctx* c, const float* inv, int S, int B, float* H, float* nrm_half,
                                     int32_t* nn01, int32_t* nn10, cudaStream_t st) {
  const long long rows = (long long)B * 2 * S;
  nn_tc2_prep_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(inv, (int)rows, H, nrm_half);
  RR_LAUNCH_CHECK(c);
  CUtensorMap mH;
  {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
      void* p = nullptr; cudaDriverEntryPointQueryResult qres;
      cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
      if (e != cudaSuccess || !p || qres != cudaDriverEntryPointSuccess) { snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled entry point unavailable"); return ROREG_ERR_CUDA; }
      fn = (PFN_encodeTiled)p;
    }
    const cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {64 * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_KC, (cuuint32_t)TC_BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&mH, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)H, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled(H) failed (%d)", (int)r); return ROREG_ERR_CUDA; }
  }
  static unsigned long long attr_mask = 0;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(nn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
  }
  const int nrb = (S + TC_BM - 1) / TC_BM;
  const int items = B * 2 * nrb;
  const int grid = items < c->sm_count ? items : c->sm_count;
  NNTcArgs a{nrm_half, S, B, nn01, nn10, 3};
  if (const char* e = getenv("ROREG_DEBUG_NN_PASSES")) { const int v = atoi(e); if (v >= 1 && v <= 3) a.passes = v; }
  nn_tc2_kernel<<<grid, T2_THREADS, T2_SMEM_BYTES, st>>>(mH, a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}


// =====================================================================================================
// v3 (nn mode 3): ONE Gram per pair.  Run 13 showed v2's time does not change when 8 of its 12 MMAs are
// removed: the kernel is paced by reading the accumulators out of TMEM (~55 B/clk/SM), and v1/v2 read every
// element twice because the column direction re-computes the Gram with the operand roles swapped.  v3 computes
// each 128x128 tile once and takes both minima from the same registers:
//   rows    : running (min, argmin) per thread over the column tiles, as before;
//   columns : warp-wide integer min (redux.sync) of the order-preserving bit pattern of d2/2 over the 32 rows a
//             warp holds, a second redux for the smallest row attaining it, the four lane-quadrants are merged in
//             shared memory and one 64-bit atomicMin per column publishes (key << 32 | row): the lexicographic
//             (distance, row) minimum, i.e. torch.min's first-index tie rule, independent of the arrival order.
// =====================================================================================================
struct NNTc3Args {
  const float* nrm_half; int S, B;
  int32_t* nn01;                       // [B][S]
  unsigned long long* col_best;        // [B][S] packed (key << 32 | row), pre-set to all ones
};

__global__ void nn_tc3_unpack_kernel(const unsigned long long* __restrict__ col_best, long long n, int32_t* __restrict__ nn10) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i < n) nn10[i] = (int32_t)(col_best[i] & 0xffffffffull);
}

__global__ void __launch_bounds__(T2_THREADS, 1) nn_tc3_kernel(const __grid_constant__ CUtensorMap mapH, NNTc3Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + T2_TILE_BYTES;
  float* sNb = reinterpret_cast<float*>(smem + T2_TILE_BYTES * (1 + T2_STAGES));   // [2][128]
  float* mrg_v = sNb + 2 * TC_BN; int* mrg_j = reinterpret_cast<int*>(mrg_v + 128);  // [128] each
  uint64_t* bars = reinterpret_cast<uint64_t*>(mrg_j + 128);
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t col_key[2][4][128];      // [tile parity][lane quadrant][column]
  __shared__ uint32_t col_row[2][4][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1); mbar_init(BAR(1), 1);
    for (int s = 0; s < T2_STAGES; ++s) { mbar_init(BAR(2 + s), 1); mbar_init(BAR(6 + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(10 + s), 1); mbar_init(BAR(12 + s), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  const int nrb = (a.S + TC_BM - 1) / TC_BM;
  const int nct = (a.S + TC_BN - 1) / TC_BN;
  const int n_items = a.B * nrb;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it_b = 0, a_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rb = item % nrb, p = item / nrb;
        const int a_row0 = (p * 2) * a.S + rb * TC_BM;          // rows of cloud 0 search ...
        const int b_row_base = (p * 2 + 1) * a.S;               // ... the rows of cloud 1
        mbar_wait(BAR(1), a_phase ^ 1);
        mbar_expect_tx(BAR(0), T2_TILE_BYTES);
        for (int c = 0; c < 2; ++c) tma_load_2d(smem_u32(sA + c * TC_BOX_BYTES), &mapH, c * TC_KC, a_row0, BAR(0));
        a_phase ^= 1;
        for (int ct = 0; ct < nct; ++ct, ++it_b) {
          const int st = it_b % T2_STAGES; const uint32_t ph = (it_b / T2_STAGES) & 1;
          mbar_wait(BAR(6 + st), ph ^ 1);
          mbar_expect_tx(BAR(2 + st), T2_TILE_BYTES);
          for (int c = 0; c < 2; ++c)
            tma_load_2d(smem_u32(sB + st * T2_TILE_BYTES + c * TC_BOX_BYTES), &mapH, c * TC_KC, b_row_base + ct * TC_BN, BAR(2 + st));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it_b = 0, it_t = 0, a_phase = 0;
      const uint32_t ahi = smem_u32(sA), alo = ahi + TC_BOX_BYTES;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(BAR(0), a_phase); a_phase ^= 1;
        for (int ct = 0; ct < nct; ++ct, ++it_b, ++it_t) {
          const int st = it_b % T2_STAGES; const uint32_t ph = (it_b / T2_STAGES) & 1;
          const int par = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
          mbar_wait(BAR(2 + st), ph);
          mbar_wait(BAR(12 + par), tph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t bhi = smem_u32(sB + st * T2_TILE_BYTES), blo = bhi + TC_BOX_BYTES;
          const uint32_t d_tmem = tmem_base + par * TC_BN;
          const uint32_t aop[3] = {ahi, alo, ahi}, bop[3] = {bhi, bhi, blo};
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int kk = 0; kk < TC_KC / 8; ++kk)
              umma_tf32(d_tmem, umma_desc_sw128(aop[c] + kk * 32), umma_desc_sw128(bop[c] + kk * 32), TC_IDESC, (c | kk) ? 1u : 0u);
          umma_commit(BAR(6 + st));
          umma_commit(BAR(10 + par));
        }
        umma_commit(BAR(1));
      }
    }
  } else {
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int row_in_tile = q * 32 + lane;
    const int et = threadIdx.x - 64;
    uint32_t it_t = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rb = item % nrb, p = item / nrb;
      const int b_row_base = (p * 2 + 1) * a.S;
      const int row = rb * TC_BM + row_in_tile;
      const bool row_ok = row < a.S;
      const float nah = row_ok ? a.nrm_half[(p * 2) * a.S + row] : 0.f;
      float best_v = INFINITY; int best_j = 0x7fffffff;
      for (int ct = 0; ct < nct; ++ct, ++it_t) {
        const int par = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
        if (et < TC_BN) {
          const int j = ct * TC_BN + et;
          sNb[par * TC_BN + et] = (j < a.S) ? a.nrm_half[b_row_base + j] : INFINITY;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mbar_wait(BAR(10 + par), tph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + par * TC_BN + hf * 64;
        uint32_t r0[32], r1[32];
        RR_TMEM_LD32(r0, taddr);
        RR_TMEM_LD32(r1, taddr + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(BAR(12 + par));
        const float* nbs = sNb + par * TC_BN + hf * 64;
        uint32_t my_key[2] = {0xffffffffu, 0xffffffffu}, my_row[2] = {0x7fffffffu, 0x7fffffffu};
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t* rg = half ? r1 : r0;
          float m = INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = nbs[half * 32 + j] - __uint_as_float(rg[j]);          // (d2 - |a|^2) / 2 : row-direction key
            m = fminf(m, v);
            // column direction: d2/2 = |a|^2/2 + v >= 0 -> its bit pattern orders like the value
            const uint32_t kb = row_ok ? __float_as_uint(fmaxf(nah + v, 0.f)) : 0xffffffffu;
            const uint32_t mn = __reduce_min_sync(0xffffffffu, kb);
            const uint32_t rw = __reduce_min_sync(0xffffffffu, (kb == mn) ? (uint32_t)row : 0x7fffffffu);
            if (lane == j) { my_key[half] = mn; my_row[half] = rw; }
          }
          if (m < best_v) {
            int jj = 31;
#pragma unroll
            for (int j = 31; j >= 0; --j) if (nbs[half * 32 + j] - __uint_as_float(rg[j]) == m) jj = j;
            best_v = m; best_j = ct * TC_BN + hf * 64 + half * 32 + jj;
          }
        }
        col_key[par][q][hf * 64 + lane] = my_key[0]; col_row[par][q][hf * 64 + lane] = my_row[0];
        col_key[par][q][hf * 64 + 32 + lane] = my_key[1]; col_row[par][q][hf * 64 + 32 + lane] = my_row[1];
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (et < TC_BN) {                                  // one thread per column: merge the 4 lane quadrants, publish
          const int j = ct * TC_BN + et;
          if (j < a.S) {
            unsigned long long best = 0xffffffffffffffffull;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const unsigned long long c = ((unsigned long long)col_key[par][qq][et] << 32) | col_row[par][qq][et];
              best = c < best ? c : best;
            }
            atomicMin(a.col_best + (long long)p * a.S + j, best);
          }
        }
      }
      if (hf == 1) { mrg_v[row_in_tile] = best_v; mrg_j[row_in_tile] = best_j; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (hf == 0) {
        if (mrg_v[row_in_tile] < best_v) { best_v = mrg_v[row_in_tile]; best_j = mrg_j[row_in_tile]; }
        if (row_ok) a.nn01[(long long)p * a.S + row] = best_j;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

static inline int nn_tc3_launch_both(roreg_ctx* c, const float* inv, int S, int B, float* H, float* nrm_half,
                                     unsigned long long* col_best, int32_t* nn01, int32_t* nn10, cudaStream_t st) {
  const long long rows = (long long)B * 2 * S;
  nn_tc2_prep_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(inv, (int)rows, H, nrm_half);
  RR_LAUNCH_CHECK(c);
  RR_CUDA(c, cudaMemsetAsync(col_best, 0xff, sizeof(unsigned long long) * (size_t)B * S, st));
  CUtensorMap mH;
  {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
      void* p = nullptr; cudaDriverEntryPointQueryResult qres;
      cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
      if (e != cudaSuccess || !p || qres != cudaDriverEntryPointSuccess) { snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled entry point unavailable"); return ROREG_ERR_CUDA; }
      fn = (PFN_encodeTiled)p;
    }
    const cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {64 * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_KC, (cuuint32_t)TC_BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&mH, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)H, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled(H) failed (%d)", (int)r); return ROREG_ERR_CUDA; }
  }
  static unsigned long long attr_mask = 0;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(nn_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
  }
  const int nrb = (S + TC_BM - 1) / TC_BM;
  const int items = B * nrb;
  const int grid = items < c->sm_count ? items : c->sm_count;
  NNTc3Args a{nrm_half, S, B, nn01, col_best};
  nn_tc3_kernel<<<grid, T2_THREADS, T2_SMEM_BYTES, st>>>(mH, a);
  RR_LAUNCH_CHECK(c);
  const long long n = (long long)B * S;
  nn_tc3_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(col_best, n, nn10);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

static inline size_t nn_tc_workspace_bytes(long long rows) {
  return 2 * rr_align(sizeof(float) * rows * TC_KEXT) + rr_align(sizeof(float) * rows) + rr_align(sizeof(unsigned long long) * rows) + 1024;
}

}  // namespace roreg
