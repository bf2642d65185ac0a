// kernels_nn_tc4.cuh - mutual-NN search, nn mode 4: ONE Gram per pair, both directions from the same accumulators,
// norms folded into the contraction, float16 two-accumulator arithmetic, epilogue free of shared-memory traffic.
//
// What runs 12/13/23/26/28 established about modes 1-3 and the first (3xTF32) cut of this kernel:
//   * reading accumulators out of TMEM is not the limit (818 B/clk/SM with 8 warps, scripts/tmem_bench.cu);
//   * a K = 8 tf32 MMA with both operands in shared memory reads 8 KB per 64 clk = the SM's whole shared-memory
//     bandwidth, so every LDS the epilogue issues (column norms) competes with the tensor pipe;
//   * modes 1/2 compute the Gram twice (once per direction) and stream 3.3 GB of operand tiles L2 -> SM per 32 pairs;
//   * FSETP / FMNMX / SEL / PRMT issue at half rate: at 5.5 such instructions per accumulator element the epilogue,
//     not the tensor pipe, paced the 3xTF32 cut (1460 of 2300 clk per tile, profiles/r01_run28_nn4_timeline.txt), and
//     its 48 KB tiles made the L2 -> SM stream (2.5 GB per 32 pairs) the next wall.
// Mode 4 therefore
//   * splits x = hi + lo with hi = fp16(x), lo' = fp16((x - hi) * 2^11) and keeps TWO accumulators per tile,
//       D1 = hi.hi,   D2 = lo'.hi + hi.lo',   value = D1 + 2^-11 * D2
//     (products of 11-bit mantissas are exact in the f32 accumulator; only lo.lo ~ 2^-22 is dropped, as with 3xTF32):
//     operand rows are 128 B instead of 256 B and a tile costs 8 kind::f16 MMAs (512 clk) instead of 13 tf32 ones;
//   * folds -|a|^2/2 - |b|^2/2 into both accumulators through one extra K-step each (the half-norm is carried as three
//     fp16 pieces n1 + 2^-11 (n2 + n3)), so  value = -d^2/2  and the epilogue never touches shared memory;
//   * makes a CTA own a 128-column block of cloud 1 (stationary B tile) and sweep all row tiles of cloud 0 through
//     it: the column direction (nn10) is a per-thread running maximum over the sweep, kept in registers
//     (cmax[64] + packed tile indices) and reduced across lanes once per item; the row direction (nn01) only
//     records the maximum VALUE of every (row, 64-column chunk) and which of its eight 8-column groups holds it
//     - 0.8 instruction per element - and two small kernels finish it: best chunk per row, then the exact index
//     among that group's 8 columns in the REFERENCE's float32 difference-form arithmetic (utils/knn_search.py:33-38);
//   * reads operands from a tile-major, pre-swizzled image written by the prep kernel (exactly the bytes of the
//     K-major shared-memory layouts: a SWIZZLE_128B box + un-swizzled extension blocks), so a tile is ONE contiguous
//     cp.async.bulk - no tensor maps; a streamed tile only moves the 20 KB its role needs, which leaves room for 9 stages
//     (run 40: with 5 stages of 32 KB the 6300-clock copy latency under load bounded the tile rate at ~1300 clk).
//
//   warp 0      producer   bulk copies: stationary tile (28 KB) once per item, streaming tiles (20 KB prefix) through 9 stages
//   warps 1,10  MMA        8 x tcgen05.mma kind::f16 (M = N = 128, K = 16) per tile, 2 x (D1 | D2) in TMEM; one issuing warp per accumulator pair
//   warps 2-9   epilogue   tcgen05.ld 32x32b.x32 -> FFMA combine -> column running max / chunk-row max
#pragma once
#include <cuda_fp16.h>
#include "kernels_nn_tc.cuh"

namespace roreg {

constexpr int T4_EXT_BYTES = 128 * 32;                           // one extension block: 128 rows x 16 fp16 (K = 16), un-swizzled K-major
constexpr int T4_TILE_BYTES = TC_BOX_BYTES + 3 * T4_EXT_BYTES;   // [hi | lo'] box (16 KB) + XA | XB1 | XB2 (4 KB each) = 28 KB per 128-row tile
constexpr int T4_STREAM_BYTES = TC_BOX_BYTES + T4_EXT_BYTES;     // a streamed (row-role) tile needs only the prefix [hi | lo'] + XA = 20 KB
constexpr int T4_STAGES = 9;
constexpr int T4_THREADS = 64 + 256 + 32;                        // producer, MMA issuer A, 8 epilogue warps, MMA issuer B
constexpr int T4_SMEM_BYTES = T4_TILE_BYTES + T4_STREAM_BYTES * T4_STAGES + 2 * 4 * 64 * 8 /*column merge*/ + 1024 + 256;   // 213 KB
constexpr float T4_PAD_NORM = 30000.f;                           // half-norm of a padding row (fp16-representable): loses every comparison
constexpr float T4_LO_SCALE = 2048.f, T4_LO_UNSCALE = 1.f / 2048.f;
// kind::f16 (A, B = F16, K-major), D = f32, M = 128, N = 128
constexpr uint32_t T4_IDESC = (1u << 4) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// K-major, no swizzle (CUTLASS cute/atom/mma_traits_sm100.hpp: Major-K INTERLEAVE = ((8,m),(T,2)):((1T,SBO),(1,LBO))): a K = 16 fp16
// slice of an extension block is two 16-byte chunks per row; 8-row core matrices of 128 B, chunk 1 at +LBO = 128 B, the next 8
// rows at +SBO = 256 B.  Descriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 | layout type 0.
__device__ __forceinline__ uint64_t t4_desc_ext(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
// byte offset of 32-bit word w (fp16 columns 2w, 2w+1; w < 8) of row r inside an extension block
__device__ __forceinline__ int t4_ext_off(int r, int w) { return (r >> 3) * 256 + (w >> 2) * 128 + (r & 7) * 16 + (w & 3) * 4; }

// byte offset of fp16 element (r, k), k < 64, in a [128 x 128 B] K-major SWIZZLE_128B box:
// 16-byte chunk (2k / 16) of row r sits at chunk position (2k / 16) ^ (r % 8)
__device__ __forceinline__ int t4_sw128_h(int r, int k) { return r * 128 + ((((k >> 3) ^ (r & 7))) << 4) + ((k & 7) << 1); }

// ---- prep: pooled features [B*2][S][32] -> tile images [B*2][NT][32 KB] ---------------------------------------
// tile = box 0 (SWIZZLE_128B, row r: hi[0..31] | lo'[0..31]) + three un-swizzled extension blocks of 16 fp16 columns:
//   XA   row-role extension     [-n1, 1, -n2, -n3, 1, 1, 0...]
//   XB1  column-role, for D1    [ 1, -n1, 0...]
//   XB2  column-role, for D2    [ 0, 0, 1, 1, -n2, -n3, 0...]
// so that  XA . XB1 = -na1 - nb1  and  XA . XB2 = -(na2 + na3) - (nb2 + nb3).  A streamed (row-role) tile is the 20 KB prefix
// box 0 + XA; only the stationary (column-role) tile needs all 28 KB.
// one warp emits one padded row rr of (pair, side) ps; lane holds x = feature[lane] (0 for a padding row rr >= S)
__device__ __forceinline__ void t4_emit_row(uint8_t* __restrict__ img, long long ps, int rr, int S, int NT, int lane, float x) {
  const int t = rr >> 7, r = rr & 127;
  uint8_t* tile = img + (ps * NT + t) * (long long)T4_TILE_BYTES;
  const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f);
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn((x - __half2float(hi)) * T4_LO_SCALE);
  const float nh = (rr < S) ? 0.5f * warp_sum(x * x) : T4_PAD_NORM;
  const __half n1 = __float2half_rn(nh);
  const float r1 = (nh - __half2float(n1)) * T4_LO_SCALE;
  const __half n2 = __float2half_rn(r1);
  const __half n3 = __float2half_rn(r1 - __half2float(n2));
  // lane L stores 32-bit word L of the 128-byte row: words 0..15 = hi pairs, 16..31 = lo' pairs (one coalesced row per warp)
  {
    const uint32_t hb = __half_as_ushort(hi), lb = __half_as_ushort(lo);
    const int s0 = (2 * lane) & 31;
    const uint32_t h0 = __shfl_sync(0xffffffffu, hb, s0), h1 = __shfl_sync(0xffffffffu, hb, s0 + 1);
    const uint32_t l0 = __shfl_sync(0xffffffffu, lb, s0), l1 = __shfl_sync(0xffffffffu, lb, s0 + 1);
    const uint32_t word = lane < 16 ? (h0 | (h1 << 16)) : (l0 | (l1 << 16));
    *reinterpret_cast<uint32_t*>(tile + t4_sw128_h(r, 2 * lane)) = word;
  }
  // extension blocks (un-swizzled): lanes 0..23 write word (lane & 7) of block lane >> 3:
  //   XA  (row role)            [-n1, 1 | -n2, -n3 | 1, 1 | 0 ...]
  //   XB1 (column role, for D1) [ 1, -n1 | 0 ...]
  //   XB2 (column role, for D2) [ 0, 0 | 1, 1 | -n2, -n3 | 0 ...]
  if (lane < 24) {
    const int blk = lane >> 3, w = lane & 7;
    __half e0 = zero, e1 = zero;
    if (blk == 0) {
      if (w == 0) { e0 = __hneg(n1); e1 = one; } else if (w == 1) { e0 = __hneg(n2); e1 = __hneg(n3); } else if (w == 2) { e0 = one; e1 = one; }
    } else if (blk == 1) {
      if (w == 0) { e0 = one; e1 = __hneg(n1); }
    } else {
      if (w == 1) { e0 = one; e1 = one; } else if (w == 2) { e0 = __hneg(n2); e1 = __hneg(n3); }
    }
    *reinterpret_cast<__half2*>(tile + TC_BOX_BYTES + blk * T4_EXT_BYTES + t4_ext_off(r, w)) = __halves2half2(e0, e1);
  }
}

__global__ void __launch_bounds__(256) nn_tc4_prep_kernel(const float* __restrict__ inv, int S, int NT, long long total_rows,
                                                          uint8_t* __restrict__ img) {
  // a warp converts 4 consecutive padded rows: the 4 loads are issued before anything is used (the kernel is latency-bound)
  const long long gr0 = (blockIdx.x * 8LL + (threadIdx.x >> 5)) * 4;    // padded row number over all (pair, side)
  const int lane = threadIdx.x & 31;
  if (gr0 >= total_rows) return;
  const int Sp = NT * TC_BM;                                             // multiple of 4: the 4 rows share (pair, side) and tile
  const long long ps = gr0 / Sp; const int rr0 = (int)(gr0 % Sp);
  float xs[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) xs[u] = (rr0 + u < S) ? inv[(ps * S + rr0 + u) * 32 + lane] : 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) t4_emit_row(img, ps, rr0 + u, S, NT, lane, xs[u]);
}

// the padding rows S .. NT*128-1 of every (pair, side) only (the fused pooling kernel writes the real rows)
__global__ void __launch_bounds__(256) nn_tc4_pad_kernel(int S, int NT, int n_ps, uint8_t* __restrict__ img) {
  const int pad = NT * TC_BM - S;
  const long long w = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (w >= (long long)n_ps * pad) return;
  t4_emit_row(img, w / pad, S + (int)(w % pad), S, NT, threadIdx.x & 31, 0.f);
}

// invariant pooling (kernels_match.cuh: inv_pool_kernel, same arithmetic) fused with the operand-image row: the pooled,
// normalised feature is written once as float32 (the resolve kernel and the reference-arithmetic paths read it) and once as
// the fp16 hi / lo' / extension row of its tile - saves the prep kernel's pass over the features.
__global__ void __launch_bounds__(256) inv_pool_t4_kernel(PoolArgs a, int NT, uint8_t* __restrict__ img) {
  __shared__ float part[8][480];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= a.rows) return;
  const int ps = r / a.S, i = r - ps * a.S;
  long long cloud = 0;
  if (a.pair_cloud) cloud = a.pair_cloud[ps];
  const int srcrow = a.sample ? a.sample[r] : i;
  const float4* src = reinterpret_cast<const float4*>(a.desc + (cloud * a.n + srcrow) * (long long)RR_ROW);
  float4 v[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) v[k] = ldg_stream4(src + k * 32 + lane);
#pragma unroll
  for (int k = 0; k < 15; ++k) part[warp][k * 32 + lane] = (v[k].x + v[k].y) + (v[k].z + v[k].w);
  __syncwarp();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 15; ++j) s += part[warp][lane * 15 + j];
  float m = s / 60.0f;
  if (a.normalise) {
    const float ss = warp_sum(m * m);
    m = m / (sqrtf(ss) + 1e-5f);
  }
  a.out[(long long)r * RR_F + lane] = m;
  t4_emit_row(img, ps, i, a.S, NT, lane, m);
}

struct NNTc4Args {
  const uint8_t* img;                  // [B*2][NT] tile images
  int S, NT, B;
  float* rowval;                       // [B][2*NT chunks][NT*128 rows]  max over the chunk's 64 columns of -d^2/2
  uint8_t* rowgid;                     // same shape: which of the chunk's eight 8-column groups holds that maximum (first one)
  int32_t* nn10;                       // [B][S]
  long long* trace;                    // TRACE instantiation only (ROREG_DEBUG_NN_TRACE=<file>): clock64 stamps [CTA][256 tiles][8 events]
};

__device__ __forceinline__ uint32_t t4_ord(float v) {            // monotone float -> uint
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void t4_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// spin without reading the clock on the success path; a lost arrive still ends in a trap, never a hung GPU
__device__ __forceinline__ void t4_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();
  }
}
// tcgen05.mma / commit issued by one elected lane of a converged warp
__device__ __forceinline__ void t4_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p, e;\nelect.sync _|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(T4_IDESC), "r"(accum) : "memory");
}
__device__ __forceinline__ void t4_commit(uint32_t bar) {
  asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(bar) : "memory");
}

// if (v > cmax) { cmax = v; byte `slot` of ctile = t; }   as FSETP + @p FADD (an FMA-pipe move: v + -0) + @p PRMT.
// (Run 43 tried the tile index as an fp16 pair updated by a predicated HFMA2 to take the PRMT off the half-rate ALU pipe:
//  the kernel got 16 % slower - fp16 FMAs issue on the "heavy" half of the FMA pipe only - so the byte-packed form stays.)
__device__ __forceinline__ void t4_col_update(float& cmax, uint32_t& ctile, float v, uint32_t t, int slot) {
#define T4_CU(SEL) asm("{\n.reg .pred p;\nsetp.gt.f32 p, %2, %0;\n@p add.f32 %0, %2, 0f80000000;\n@p prmt.b32 %1, %1, %3, " #SEL ";\n}" \
                       : "+f"(cmax), "+r"(ctile) : "f"(v), "r"(t))
  switch (slot) {
    case 0: T4_CU(0x3214); break;
    case 1: T4_CU(0x3240); break;
    case 2: T4_CU(0x3410); break;
    default: T4_CU(0x4210); break;
  }
#undef T4_CU
}

template <bool TRACE>
__global__ void __launch_bounds__(T4_THREADS, 1) nn_tc4_kernel(NNTc4Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                   // stationary tile (cloud 1 block), 28 KB
  uint8_t* sA = smem + T4_TILE_BYTES;                   // T4_STAGES x 20 KB streaming tiles (cloud 0)
  unsigned long long* cmg = reinterpret_cast<unsigned long long*>(smem + T4_TILE_BYTES + T4_STREAM_BYTES * T4_STAGES);   // [2 halves][4 quadrants][64 columns]
  uint64_t* bars = reinterpret_cast<uint64_t*>(cmg + 2 * 4 * 64);
  // barriers: 0 b_full, 1 b_empty, 2..10 a_full, 11..19 a_empty, 20..21 tmem_full, 22..23 tmem_empty
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1); mbar_init(BAR(1), 2);         // b_empty: one commit per MMA-issuing warp
    for (int s = 0; s < T4_STAGES; ++s) { mbar_init(BAR(2 + s), 1); mbar_init(BAR(2 + T4_STAGES + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(2 + 2 * T4_STAGES + s), 1); mbar_init(BAR(4 + 2 * T4_STAGES + s), 256); }
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

  // unpredicated on purpose (see kernels_corr_tc2.cuh): event e of this CTA's i-th tile
#define T4_TRACE(i, e) do { if (TRACE) a.trace[((size_t)blockIdx.x * 256 + ((i) < 255u ? (i) : 255u)) * 8 + (e)] = clock64(); } while (0)
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
          t4_wait(BAR(2 + T4_STAGES + st), ph ^ 1);
          T4_TRACE(it_a, 0);
          mbar_expect_tx(BAR(2 + st), T4_STREAM_BYTES);
          t4_bulk(smem_u32(sA + st * T4_STREAM_BYTES), imgA + (long long)t * T4_TILE_BYTES, T4_STREAM_BYTES, BAR(2 + st));
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // Two MMA-issuing warps, one per accumulator pair: warp 1 takes the even tiles of the CTA's tile sequence, warp 10 the odd
    // ones.  The serial per-tile loop of a single issuer (waits, descriptor set-up, 8 issues that each block for the 64
    // tensor clocks of the previous one, 2 commits) paced the kernel at ~1300 clk per tile (run 40 timeline) while the tensor
    // pipe needs 512; with two issuers the loops overlap.  The whole warp walks the loop (no divergent region around the
    // tcgen05 instructions), one elected lane issues.
    const uint32_t mine = (warp == 1) ? 0u : 1u;
    uint32_t it_a = 0, it_t = 0, b_phase = 0;
    // box-0 descriptors: low word (start address >> 4 | LBO field) + a constant high word (SBO = 1024 B, version 1, SWIZZLE_128B);
    // hi at +0/+32 B, lo' at +64/+96 B of a row.  Extension blocks: un-swizzled descriptors (t4_desc_ext).
    constexpr uint64_t DHI = ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t b0 = ((smem_u32(sB) >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t a00 = ((smem_u32(sA) >> 4) & 0x3FFF) | (1u << 16);
    const uint64_t bx1 = t4_desc_ext(smem_u32(sB) + TC_BOX_BYTES + T4_EXT_BYTES), bx2 = t4_desc_ext(smem_u32(sB) + TC_BOX_BYTES + 2 * T4_EXT_BYTES);
    auto D = [&](uint32_t lo) -> uint64_t { return DHI | lo; };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      t4_wait(BAR(0), b_phase); b_phase ^= 1;
      for (int t = 0; t < NT; ++t, ++it_a, ++it_t) {
        if ((it_t & 1u) != mine) continue;
        const int st = it_a % T4_STAGES; const uint32_t ph = (it_a / T4_STAGES) & 1;
        const int acc = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
        t4_wait(BAR(2 + st), ph);
        if (lane == 0) T4_TRACE(it_t, 1);
        t4_wait(BAR(4 + 2 * T4_STAGES + acc), tph ^ 1);
        if (lane == 0) T4_TRACE(it_t, 2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = a00 + st * (T4_STREAM_BYTES >> 4);          // 16-byte units: hi +0/+2, lo' +4/+6
        const uint64_t ax = t4_desc_ext(smem_u32(sA) + st * T4_STREAM_BYTES + TC_BOX_BYTES);
        const uint32_t d1 = tmem_base + acc * 256, d2 = d1 + 128;
        t4_umma(d1, ax, bx1, 0u);                                       // -na1 - nb1
        t4_umma(d1, D(a0), D(b0), 1u);                                  // hi.hi
        t4_umma(d1, D(a0 + 2), D(b0 + 2), 1u);
        t4_umma(d2, ax, bx2, 0u);                                       // -(na2 + na3) - (nb2 + nb3)
        t4_umma(d2, D(a0 + 4), D(b0), 1u);                              // lo'.hi
        t4_umma(d2, D(a0 + 6), D(b0 + 2), 1u);
        t4_umma(d2, D(a0), D(b0 + 4), 1u);                              // hi.lo'
        t4_umma(d2, D(a0 + 2), D(b0 + 6), 1u);
        t4_commit(BAR(2 + T4_STAGES + st));
        t4_commit(BAR(2 + 2 * T4_STAGES + acc));
        if (lane == 0) T4_TRACE(it_t, 3);
      }
      t4_commit(BAR(1));
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
      float cmax[64]; uint32_t ctile[16];             // running column maxima; tile index of each as a byte, four per register
#pragma unroll
      for (int e = 0; e < 64; ++e) cmax[e] = -INFINITY;
#pragma unroll
      for (int e = 0; e < 16; ++e) ctile[e] = 0;
      float* rv = a.rowval + ((long long)p * 2 * NT + 2 * j + hf) * Sp;
      uint8_t* rg = a.rowgid + ((long long)p * 2 * NT + 2 * j + hf) * Sp;
      for (int t = 0; t < NT; ++t, ++it_t) {
        const int acc = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
        t4_wait(BAR(2 + 2 * T4_STAGES + acc), tph);
        if (threadIdx.x == 64) T4_TRACE(it_t, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + hf * 64;
        float g8[8];                                    // maxima of the eight 8-column groups of this thread's 64 columns
#pragma unroll
        for (int k = 0; k < 8; ++k) g8[k] = -INFINITY;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r1[32], r2[32];
          RR_TMEM_LD32(r1, taddr + half * 32);
          RR_TMEM_LD32(r2, taddr + 128 + half * 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (half == 1) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(BAR(4 + 2 * T4_STAGES + acc));  // accumulators free: the MMAs of tile t+2 may start
            if (threadIdx.x == 64) T4_TRACE(it_t, 5);
          }
          // column direction: running maximum of -d^2/2 over the rows this lane sees; strict '>' + increasing t = first row wins ties.
          // Per element: FFMA (combine) + FSETP + @p FADD + @p PRMT, and half an FMNMX3 for the row direction.
#pragma unroll
          for (int u = 0; u < 32; u += 2) {
            const int e = half * 32 + u;
            const float v0 = fmaf(__uint_as_float(r2[u]), T4_LO_UNSCALE, __uint_as_float(r1[u]));
            const float v1 = fmaf(__uint_as_float(r2[u + 1]), T4_LO_UNSCALE, __uint_as_float(r1[u + 1]));
            t4_col_update(cmax[e], ctile[e >> 2], v0, (uint32_t)t, e & 3);
            t4_col_update(cmax[e + 1], ctile[e >> 2], v1, (uint32_t)t, (e + 1) & 3);
            g8[e >> 3] = fmaxf(fmaxf(g8[e >> 3], v0), v1);
          }
        }
        if (threadIdx.x == 64) T4_TRACE(it_t, 6);
        // row direction: only the chunk maximum and the first 8-column group attaining it; the index is resolved by nn_tc4_resolve_kernel
        const float m = fmaxf(fmaxf(fmaxf(g8[0], g8[1]), fmaxf(g8[2], g8[3])), fmaxf(fmaxf(g8[4], g8[5]), fmaxf(g8[6], g8[7])));
        int gid = 7;
#pragma unroll
        for (int k = 6; k >= 0; --k) if (g8[k] == m) gid = k;
        rv[t * TC_BM + row_in_tile] = m;
        rg[t * TC_BM + row_in_tile] = (uint8_t)gid;
        if (threadIdx.x == 64) T4_TRACE(it_t, 7);
      }
      // ---- end of the sweep: reduce the column candidates over the 128 lanes -------------------------------
      // lexicographic (largest value, smallest row); row = tile * 128 + row_in_tile.  Two warp-wide integer reductions per column
      // (redux.sync on the order-preserving key, then on the rows attaining it): the shuffle butterfly this replaces cost
      // 18-33 k clocks per item - a third of the sweep itself (profiles/r01_run40_nn4_timeline.txt, tiles 39 -> 40).
#pragma unroll
      for (int e = 0; e < 64; ++e) {
        const uint32_t key = t4_ord(cmax[e] + 0.f);     // + 0.f: -0 and +0 compare equal, as in float arithmetic
        const uint32_t row = ((ctile[e >> 2] >> (8 * (e & 3))) & 0xff) * TC_BM + row_in_tile;
        const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
        const uint32_t rmin = __reduce_min_sync(0xffffffffu, key == kmax ? row : 0xffffffffu);
        if (lane == (e & 31)) cmg[(hf * 4 + q) * 64 + e] = ((unsigned long long)kmax << 32) | (0xffffffffu - rmin);
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

// ---- finish A: best 64-column chunk of every row (first chunk on equal values) -> its first maximal 8-column group ------
__global__ void __launch_bounds__(256) nn_tc4_best_group_kernel(const float* __restrict__ rowval, const uint8_t* __restrict__ rowgid,
                                                                int S, int NT, int B, int32_t* __restrict__ best_group) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= (long long)B * S) return;
  const int p = (int)(i / S), row = (int)(i % S);
  const long long Sp = (long long)NT * TC_BM;
  const float* rv = rowval + (long long)p * 2 * NT * Sp + row;
  float best = -INFINITY; int bc = 0;
  for (int c = 0; c < 2 * NT; ++c) { const float v = rv[c * Sp]; if (v > best) { best = v; bc = c; } }
  best_group[i] = bc * 8 + rowgid[(long long)p * 2 * NT * Sp + bc * Sp + row];
}

// ---- finish B: exact index inside the winning 8-column group, in the reference's arithmetic ------------------------
// d = sqrt(sum_f (a_f - b_f)^2 + 1e-7) accumulated f = 0..31 with FMA, lexicographic (d, index) minimum = torch's
// dist.min(dim) on utils/knn_search.py:33-38's values (the same arithmetic as nn mode 0 / nn_diff_kernel).
// One thread per (row, candidate column): 8 consecutive lanes resolve one row.
__global__ void __launch_bounds__(256) nn_tc4_resolve_kernel(const float* __restrict__ inv, const int32_t* __restrict__ best_group,
                                                             int S, int B, int32_t* __restrict__ nn01) {
  const long long gt = blockIdx.x * 256LL + threadIdx.x;
  const long long i = gt >> 3; const int cnd = (int)(gt & 7);
  const bool live = i < (long long)B * S;                 // B * S is not always a multiple of 32: keep the shuffles warp-uniform
  float s = INFINITY; int col = 0x7fffffff;
  if (live) {
    const int p = (int)(i / S), row = (int)(i % S);
    col = best_group[i] * 8 + cnd;
    if (col < S) {
      const float4* a4 = reinterpret_cast<const float4*>(inv + ((long long)(p * 2) * S + row) * 32);
      const float4* b4 = reinterpret_cast<const float4*>(inv + ((long long)(p * 2 + 1) * S + col) * 32);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 av = __ldg(a4 + k), bv = __ldg(b4 + k);
        float d;
        d = av.x - bv.x; acc = fmaf(d, d, acc);
        d = av.y - bv.y; acc = fmaf(d, d, acc);
        d = av.z - bv.z; acc = fmaf(d, d, acc);
        d = av.w - bv.w; acc = fmaf(d, d, acc);
      }
      s = sqrtf(acc + 1e-7f);
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    const float so = __shfl_xor_sync(0xffffffffu, s, o);
    const int co = __shfl_xor_sync(0xffffffffu, col, o);
    if (so < s || (so == s && co < col)) { s = so; col = co; }
  }
  if (live && cnd == 0) nn01[i] = col;
}

static inline size_t nn_tc4_workspace_bytes(int B, int S) {
  const size_t NT = (size_t)(S + TC_BM - 1) / TC_BM;
  return rr_align((size_t)B * 2 * NT * T4_TILE_BYTES) + rr_align(sizeof(float) * (size_t)B * 2 * NT * NT * TC_BM) +
         rr_align((size_t)B * 2 * NT * NT * TC_BM) + rr_align(sizeof(int32_t) * (size_t)B * S) + 4096;
}

// inv: [B][2][S][32] pooled features; img / rowval / rowgid / best_group: workspace of nn_tc4_workspace_bytes(B, S)
// img_ready: the real rows of the image were already written by inv_pool_t4_kernel (only the padding rows are missing)
static inline int nn_tc4_launch_both(roreg_ctx* c, const float* inv, int S, int B, uint8_t* img, float* rowval, uint8_t* rowgid, int32_t* best_group,
                                     int32_t* nn01, int32_t* nn10, cudaStream_t st, bool img_ready = false) {
  const int NT = (S + TC_BM - 1) / TC_BM;
  RR_ARG(c, NT <= 256);                                  // tile indices of the column direction are kept as bytes
  RR_ARG(c, (reinterpret_cast<uintptr_t>(img) & 15) == 0);    // cp.async.bulk source: 16-byte aligned (the swizzle only constrains the shared-memory side)
  const long long total_rows = (long long)B * 2 * NT * TC_BM;
  if (img_ready) {
    const long long pad_rows = (long long)B * 2 * (NT * TC_BM - S);
    if (pad_rows > 0) {
      nn_tc4_pad_kernel<<<(unsigned)((pad_rows + 7) / 8), 256, 0, st>>>(S, NT, B * 2, img);
      RR_LAUNCH_CHECK(c);
    }
  } else {
    nn_tc4_prep_kernel<<<(unsigned)((total_rows / 4 + 7) / 8), 256, 0, st>>>(inv, S, NT, total_rows, img);
    RR_LAUNCH_CHECK(c);
  }
  static unsigned long long attr_mask = 0;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(nn_tc4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T4_SMEM_BYTES));
    RR_CUDA(c, cudaFuncSetAttribute(nn_tc4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T4_SMEM_BYTES));
  }
  const int items = B * NT;
  const int grid = items < c->sm_count ? items : c->sm_count;
  NNTc4Args a{img, S, NT, B, rowval, rowgid, nn10, nullptr};
  const char* trace_fn = getenv("ROREG_DEBUG_NN_TRACE");
  static bool traced = false;
  if (trace_fn && !traced && items >= 1000) {            // one-off timeline dump of CTA 0 (debug only; synchronises)
    traced = true;
    const size_t nb = (size_t)grid * 256 * 8 * sizeof(long long);
    RR_CUDA(c, cudaMalloc(&a.trace, nb));
    RR_CUDA(c, cudaMemsetAsync(a.trace, 0, nb, st));
    nn_tc4_kernel<true><<<grid, T4_THREADS, T4_SMEM_BYTES, st>>>(a);
    RR_LAUNCH_CHECK(c);
    RR_CUDA(c, cudaStreamSynchronize(st));
    long long* h = (long long*)malloc(256 * 8 * sizeof(long long));
    RR_CUDA(c, cudaMemcpy(h, a.trace, 256 * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_fn, "w")) {
      fprintf(f, "# tile P_issue M_afull M_accfree M_committed E_full E_loaded E_math E_end (clock64 - first)\n");
      const long long t0 = h[0];
      for (int i = 0; i < 255; ++i) {
        fprintf(f, "%d", i);
        for (int e = 0; e < 8; ++e) fprintf(f, " %lld", h[i * 8 + e] ? h[i * 8 + e] - t0 : -1);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    free(h); cudaFree(a.trace);
  } else {
    nn_tc4_kernel<false><<<grid, T4_THREADS, T4_SMEM_BYTES, st>>>(a);
    RR_LAUNCH_CHECK(c);
  }
  const long long n = (long long)B * S;
  nn_tc4_best_group_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rowval, rowgid, S, NT, B, best_group);
  RR_LAUNCH_CHECK(c);
  nn_tc4_resolve_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, st>>>(inv, best_group, S, B, nn01);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

}  // namespace roreg
