// kernels_corr_tc.cuh - equivariant correlation (Des2R / R-indicator) with the 60x60 Gram on tcgen05.
//
// Per match  G[h][g] = sum_f X[f][h] * Y[f][g]  is a dense [60x32]x[32x60] contraction.  Two matches are
// stacked per MMA (M = N = 128: rows = (match, h), columns = (match, g); the two off-diagonal 64x64
// blocks are unused - the tensor pipe has the head-room, the kernel is HBM-bound on the 2 x 7680 B it
// must gather per match).  A descriptor row [32 f][60 h] lands in shared memory exactly as it lies in
// HBM: ONE 7680-byte cp.async.bulk per row (a 256-byte-aligned contiguous run = 60 full 128-byte lines).
// (Rounds 1-21 used a 2-D tensor-map box [32 f] x [64 h]; its 240-byte, sector-straddling box rows were
// issued one by one and capped the bare load skeleton at ~4 TB/s - run 16.)
// Accuracy: 3xTF32.  A "convert" warp group reads the landed tile, splits hi = tf32(x), lo = x - hi and
// writes both TRANSPOSED into the K-major SWIZZLE_128B operand layout the NN kernel already uses
// ([128 rows = (match,h)] x [32 f]), and the issuer runs (hi,hi) + (lo,hi) + (hi,lo) into one TMEM
// accumulator: float32-class products.  (A first version fed the MN-major tiles to the MMA directly; it
// ran but produced zeros on the B200 - run 4 - so the transposition moved into the convert stage, which
// touches every element anyway.)
// Epilogue: each thread owns one Gram row in TMEM, writes it transposed to shared memory and thread a
// sums the generalised diagonal  cor[a] = sum_g G[tab[a][g]][g]  (tab = P: variant 1, P^T: variant 2).
//
//   warp 0      TMA producer     4 bulk copies (2 matches x {X,Y}) of 7680 B per stage
//   warp 1      MMA issuer       3 x 4 tcgen05.mma kind::tf32 (M=N=128, K=8), commit -> mbarrier
//   next CW     convert          hi/lo split + transpose in shared memory, fence.proxy.async, arrive   (CW = 4 or 8 warps)
//   last 4      epilogue         tcgen05.ld -> smem transpose -> diagonal sums -> argmax
#pragma once
#include "kernels_nn_tc.cuh"
#include "kernels_corr.cuh"

namespace roreg {

constexpr int CT_STAGES = 4;                              // raw landing buffers: 4 x 32 KB in flight per SM hide the HBM latency (run 13: 2 stages capped the kernel at 2.1 TB/s)
constexpr int CT_RAW_BOX = 32 * 60 * 4;                   // [32 f][60 h] f32 = 7680 B: one descriptor row exactly as in HBM
constexpr int CT_RAW_BYTES = 4 * CT_RAW_BOX;              // X0 | X1 | Y0 | Y1 = 30 KB per stage
constexpr int CT_OPER_BYTES = 128 * 32 * 4;               // one K-major operand: 128 rows x 32 f = 16 KB
constexpr int CT_TILES_BYTES = 4 * CT_OPER_BYTES;         // Xhi | Xlo | Yhi | Ylo = 64 KB, single-buffered (double-buffering them bought nothing in run 13)
constexpr int CT_GS_BYTES = 2 * 60 * 64 * 4;              // transposed Gram of both matches [2][60 g][64 h]
constexpr int CT_SMEM_BYTES = CT_STAGES * CT_RAW_BYTES + CT_TILES_BYTES + CT_GS_BYTES + 3600 + 16 + 256 + 1024;   // 224,032 B of the 232,448 B limit
template <int CW> struct CtThreads { static constexpr int value = 64 + 32 * CW + 128; };   // TMA, MMA, CW convert warps, 4 epilogue warps

struct CorrTcArgs {
  const float* X; const float* Y;    // descriptor arrays [rows][32][60]; filled in by group_corr_tc_launch
  const int32_t* idxX; const int32_t* idxY; int idx_stride;
  const int32_t* pair_cloud; int n;
  const int32_t* n_matches; int K, B;
  const uint8_t* tab;
  float* cor_out; int32_t* argmax_out;
  int dbg_passes, dbg_skip;          // bottleneck experiments only (ROREG_DEBUG_CORR_PASSES / ROREG_DEBUG_CORR_SKIP): defaults 3 / 0
  long long* trace;                  // mode 2 only, ROREG_DEBUG_CORR_TRACE=<file>: clock64 stamps [256 items][12 events] of CTA 0
};

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int CW>
__global__ void __launch_bounds__(CtThreads<CW>::value, 1) group_corr_tc_kernel(CorrTcArgs a) {
  constexpr int CT_THREADS = CtThreads<CW>::value;
  constexpr int CONV_THREADS = 32 * CW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* tiles0 = smem + CT_STAGES * CT_RAW_BYTES;                                       // 2 x (Xhi | Xlo | Yhi | Ylo)
  float* Gs = reinterpret_cast<float*>(tiles0 + CT_TILES_BYTES);                           // [2][60 g][64 h]
  uint8_t* tabs = reinterpret_cast<uint8_t*>(Gs) + CT_GS_BYTES;                            // 3600 B
  float* red_v = reinterpret_cast<float*>(tabs + 3600); int* red_i = reinterpret_cast<int*>(red_v + 2);   // [2] each
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(tabs + 3600 + 16) + 7) & ~uintptr_t(7));
  // barriers: 0..3 raw_full[s], 4..7 raw_free[s], 8 conv_done, 9 tiles_free, 10..11 mma_done[acc], 12..13 acc_free[acc]
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  for (int e = threadIdx.x; e < 3600; e += CT_THREADS) tabs[e] = a.tab[e];
  if (threadIdx.x == 0) {
    for (int s = 0; s < CT_STAGES; ++s) { mbar_init(BAR(0 + s), 1); mbar_init(BAR(4 + s), CONV_THREADS); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(10 + s), 1); mbar_init(BAR(12 + s), 128); }
    mbar_init(BAR(8), CONV_THREADS); mbar_init(BAR(9), 1);
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

  const int items_per_pair = (a.K + 1) / 2;
  const long long n_items = (long long)a.B * items_per_pair;
  // every role walks the same item sequence and skips the same items (device-side match counts)
  auto item_count = [&](long long item, int& p, int& k0) -> int {
    p = (int)(item / items_per_pair); k0 = (int)(item % items_per_pair) * 2;
    const int cnt = a.n_matches ? a.n_matches[p] : a.K;
    return cnt - k0;                                   // <= 0: nothing, 1: one match, >= 2: two matches
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The row indices of a match come from global memory (match list -> keypoint rows); one thread chasing them
    // item by item costs ~4 dependent L2/HBM round trips per item (run 16: 0.41 of 0.77 ms was this skeleton).
    // All 32 lanes therefore resolve the coordinates of 32 upcoming items in parallel; lane 0 then issues the TMAs.
    {
      uint32_t it = 0;
      for (long long base = blockIdx.x; base < n_items; base += 32LL * gridDim.x) {
        const long long item = base + (long long)lane * gridDim.x;
        int avail = 0; long long cx[2] = {0, 0}, cy[2] = {0, 0};   // row numbers in X / Y
        if (item < n_items) {
          int p, k0; avail = item_count(item, p, k0);
          if (avail > 0) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              const int k = k0 + ((m < avail) ? m : 0);    // odd tail: the second slot re-reads the first match
              const long long w = (long long)p * a.K + k;
              long long rx = a.idxX ? a.idxX[w * a.idx_stride] : k;
              long long ry = a.idxY ? a.idxY[w * a.idx_stride] : k;
              if (a.pair_cloud) { rx += (long long)a.pair_cloud[2 * p + 1] * a.n; ry += (long long)a.pair_cloud[2 * p] * a.n; }
              cx[m] = rx; cy[m] = ry;
            }
          }
        }
        for (int l = 0; l < 32; ++l) {
          const int av = __shfl_sync(0xffffffffu, avail, l);
          const long long x0 = __shfl_sync(0xffffffffu, cx[0], l), x1 = __shfl_sync(0xffffffffu, cx[1], l);
          const long long y0 = __shfl_sync(0xffffffffu, cy[0], l), y1 = __shfl_sync(0xffffffffu, cy[1], l);
          if (av <= 0) continue;                         // uniform across the warp (shuffled value)
          if (lane == 0) {
            const int st = it % CT_STAGES; const uint32_t ph = (it / CT_STAGES) & 1;
            mbar_wait(BAR(4 + st), ph ^ 1);              // convert warps have consumed this raw buffer
            uint8_t* sb = smem + st * CT_RAW_BYTES;
            mbar_expect_tx(BAR(0 + st), CT_RAW_BYTES);
            bulk_load_1d(smem_u32(sb + 0 * CT_RAW_BOX), a.X + x0 * RR_ROW, CT_RAW_BOX, BAR(0 + st));
            bulk_load_1d(smem_u32(sb + 2 * CT_RAW_BOX), a.Y + y0 * RR_ROW, CT_RAW_BOX, BAR(0 + st));
            bulk_load_1d(smem_u32(sb + 1 * CT_RAW_BOX), a.X + x1 * RR_ROW, CT_RAW_BOX, BAR(0 + st));
            bulk_load_1d(smem_u32(sb + 3 * CT_RAW_BOX), a.Y + y1 * RR_ROW, CT_RAW_BOX, BAR(0 + st));
          }
          ++it;
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        int p, k0; if (item_count(item, p, k0) <= 0) continue;
        const int acc = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(BAR(8), it & 1);                     // operand tiles written and visible to the async proxy
        mbar_wait(BAR(12 + acc), aph ^ 1);             // accumulator drained
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tb = smem_u32(tiles0);
        const uint32_t xhi = tb, xlo = tb + CT_OPER_BYTES, yhi = tb + 2 * CT_OPER_BYTES, ylo = tb + 3 * CT_OPER_BYTES;
        const uint32_t d_tmem = tmem_base + acc * 128;
        const uint32_t aop[3] = {xhi, xlo, xhi}, bop[3] = {yhi, yhi, ylo};
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < a.dbg_passes) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_tf32(d_tmem, umma_desc_sw128(aop[c] + kk * 32), umma_desc_sw128(bop[c] + kk * 32), TC_IDESC, (c | kk) ? 1u : 0u);
          }
        umma_commit(BAR(9));                           // operand tiles reusable by the convert warps
        umma_commit(BAR(10 + acc));                    // accumulator ready for the epilogue
        ++it;
      }
    }
  } else if (warp < 2 + CW) {
    // ===================== convert: split hi/lo and transpose into K-major SW128 operand tiles ============
    // raw[f][h] (h contiguous, pitch 60) -> tile row r = (match, h), 128 B of f per row, 16-B chunk c = f/4
    // stored at chunk position c ^ (r % 8)  (the 128-byte swizzle TMA / UMMA use).  Rows h = 60..63 are zeros.
    // CW = 4: a thread converts its row of X and of Y; CW = 8: threads 0..127 take X, 128..255 take Y.
    const int cid = threadIdx.x - 64;
    const int ct = cid & 127;                          // operand row (match = ct/64, h = ct%64)
    uint32_t it = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      int p, k0; if (item_count(item, p, k0) <= 0) continue;
      const int st = it % CT_STAGES; const uint32_t ph = (it / CT_STAGES) & 1;
      mbar_wait(BAR(0 + st), ph);                      // raw tile landed
      mbar_wait(BAR(9), (it & 1) ^ 1);                 // the MMAs of the previous item no longer read the operand tiles
      uint8_t* tiles = tiles0;
      const float* raw = reinterpret_cast<const float*>(smem + st * CT_RAW_BYTES);
      const int m = ct >> 6, h = ct & 63;
      const bool live = h < RR_G;
      if (!(a.dbg_skip & 2))
#pragma unroll
      for (int o = 0; o < (CW == 8 ? 1 : 2); ++o) {    // 0: X, 1: Y
        const int op = (CW == 8) ? (cid >> 7) : o;
        const float* src = raw + (op * 2 + m) * (CT_RAW_BOX / 4) + (live ? h : 0);
        uint8_t* thi = tiles + (op * 2) * CT_OPER_BYTES + ct * 128;
        uint8_t* tlo = thi + CT_OPER_BYTES;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 hv, lv; float x; uint32_t t;
          x = live ? src[(4 * c + 0) * RR_G] : 0.f; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x)); hv.x = __uint_as_float(t); lv.x = x - hv.x;
          x = live ? src[(4 * c + 1) * RR_G] : 0.f; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x)); hv.y = __uint_as_float(t); lv.y = x - hv.y;
          x = live ? src[(4 * c + 2) * RR_G] : 0.f; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x)); hv.z = __uint_as_float(t); lv.z = x - hv.z;
          x = live ? src[(4 * c + 3) * RR_G] : 0.f; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x)); hv.w = __uint_as_float(t); lv.w = x - hv.w;
          const int pos = (c ^ (ct & 7)) * 16;
          *reinterpret_cast<float4*>(thi + pos) = hv;
          *reinterpret_cast<float4*>(tlo + pos) = lv;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05
      mbar_arrive(BAR(8));                             // conv_done
      mbar_arrive(BAR(4 + st));                        // raw buffer free for the next TMA
      ++it;
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                            // TMEM lane quadrant of this warp (warp id mod 4, for CW = 4 and 8 alike)
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
      const int acc = it & 1; const uint32_t ph = (it >> 1) & 1;
      mbar_wait(BAR(10 + acc), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128 + m * 64;
      uint32_t r[64];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t* rr = r + half * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
                       "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15]),
                       "=r"(rr[16]), "=r"(rr[17]), "=r"(rr[18]), "=r"(rr[19]), "=r"(rr[20]), "=r"(rr[21]), "=r"(rr[22]), "=r"(rr[23]),
                       "=r"(rr[24]), "=r"(rr[25]), "=r"(rr[26]), "=r"(rr[27]), "=r"(rr[28]), "=r"(rr[29]), "=r"(rr[30]), "=r"(rr[31])
                     : "r"(taddr + half * 32) : "memory");
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(BAR(12 + acc));                      // accumulator free as soon as it sits in registers
      // transposed store: Gs[m][g][h]; a warp writes 32 consecutive h -> conflict-free
#pragma unroll
      for (int g = 0; g < 60; ++g) G[g * 64 + h] = __uint_as_float(r[g]);
      asm volatile("bar.sync %0, 64;" ::"r"(2 + m) : "memory");
      float c = -INFINITY;
      if (a.dbg_skip & 1) c = G[h];
      else if (h < RR_G) {
        float c4[4] = {0.f, 0.f, 0.f, 0.f};                  // four independent chains (the sum order differs from g = 0..59 only in rounding)
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
      ++it;
    }
  }
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

template <int CW>
static inline int group_corr_tc_launch_cw(roreg_ctx* c, const CorrTcArgs& a, int grid, cudaStream_t st) {
  static unsigned long long attr_mask = 0;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(group_corr_tc_kernel<CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_BYTES));
  }
  group_corr_tc_kernel<CW><<<grid, CtThreads<CW>::value, CT_SMEM_BYTES, st>>>(a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

// X, Y: descriptor arrays [rows][32][60] float32 (a row = 7680 B, so any 16-byte aligned base keeps every bulk copy aligned)
static inline int group_corr_tc_launch(roreg_ctx* c, const float* X, const float* Y, CorrTcArgs a, cudaStream_t st) {
  RR_ARG(c, (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  a.X = X; a.Y = Y;
  const long long items = (long long)a.B * ((a.K + 1) / 2);
  const int grid = (int)(items < c->sm_count ? items : c->sm_count);
  static int cw = 0;
  if (!cw) { cw = 4; if (const char* e = getenv("ROREG_DEBUG_CORR_CW")) if (atoi(e) == 8) cw = 8; }
  return cw == 8 ? group_corr_tc_launch_cw<8>(c, a, grid, st) : group_corr_tc_launch_cw<4>(c, a, grid, st);
}

}  // namespace roreg
