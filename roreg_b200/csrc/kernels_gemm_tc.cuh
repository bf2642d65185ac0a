// kernels_gemm_tc.cuh - dense layer GEMM for the group-convolution networks (GF, ET, RD) on tcgen05.
//
//   out[r][o] = sum_c A[r][c] * W[o][c]         A: [R][Kdim] activations (rows = (item, group element)),
//                                                W: [O][Kdim] weights, both K-major float32 in HBM.
// A group convolution (network/ops.py:11-20) is this GEMM on the im2col-gathered activation matrix
// (Kdim = 13*Cin, see kernels_gconv.cuh); the 1x1 FC head of ET_test (network/eqv_trans.py:93-101) is it
// directly.  Precision: npass = 1 -> one TF32 pass (what the reference's own cuDNN path does on Ampere+,
// torch.backends.cudnn.allow_tf32 defaults to True), npass = 3 -> hi/lo split of both operands, three
// passes (hi.hi + lo.hi + hi.lo) into the same accumulator: float32-class, used for parity with the oracle.
// Fused epilogue (one thread per output row, TMEM -> registers):
//   v = acc + bias[o] (+ residual[r][o]);  raw_out[r][o] = v;  y = v*bn_scale[o] + bn_shift[o];  relu;
//   act_hi[r][o] = tf32(y), act_lo[r][o] = y - tf32(y)      (the next layer's operands)
// Pipeline: warp 0 TMA producer (SWIZZLE_128B boxes), warp 1 MMA issuer, warps 2-5 epilogue; 4 smem stages;
// two TMEM accumulators of up to 256 columns so the epilogue of tile t overlaps the MMAs of tile t+1.
#pragma once
#include "kernels_nn_tc.cuh"

namespace roreg {

constexpr int GM_BM = 128, GM_KC = 32, GM_STAGES = 4;
constexpr int GM_A_BYTES = GM_BM * GM_KC * 4;            // 16 KB
constexpr int GM_W_BYTES = 256 * GM_KC * 4;              // up to 32 KB (NT <= 256 weight rows)
constexpr int GM_STAGE_BYTES = GM_A_BYTES + GM_W_BYTES;  // 48 KB
constexpr int GM_SMEM_BYTES = GM_STAGES * GM_STAGE_BYTES + 1024 + 256;

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
};

__global__ void __launch_bounds__(192, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                                                         const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo,
                                                         GemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GM_STAGES * GM_STAGE_BYTES);
  // barriers: 0..3 full, 4..7 empty, 8..9 tmem_full, 10..11 tmem_empty
  __shared__ uint32_t tmem_base_s;
  __shared__ int32_t acols_s[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  for (int i = threadIdx.x; i < a.Kdim / GM_KC && i < 256; i += blockDim.x) acols_s[i] = a.a_cols ? a.a_cols[i] : i * GM_KC;
  if (threadIdx.x == 0) {
    for (int s = 0; s < GM_STAGES; ++s) { mbar_init(BAR(s), 1); mbar_init(BAR(4 + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(8 + s), 1); mbar_init(BAR(10 + s), 128); }
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

  const int n_mt = (a.R + GM_BM - 1) / GM_BM;
  const int n_tiles = n_mt * a.n_ntiles;
  const int n_kc = a.Kdim / GM_KC;
  const int n_k = n_kc * a.npass;
  const uint32_t stage_tx = GM_A_BYTES + (uint32_t)a.NT * GM_KC * 4;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.NT >> 3) << 17) | ((uint32_t)(GM_BM >> 4) << 24);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int mt = t / a.n_ntiles, nt = t % a.n_ntiles;
        for (int k = 0; k < n_k; ++k, ++it) {
          const int pass = k / n_kc, kc = k % n_kc;
          const int st = it % GM_STAGES; const uint32_t ph = (it / GM_STAGES) & 1;
          mbar_wait(BAR(4 + st), ph ^ 1);
          mbar_expect_tx(BAR(st), stage_tx);
          uint8_t* sb = smem + st * GM_STAGE_BYTES;
          // pass 0: A_hi.W_hi   pass 1: A_lo.W_hi   pass 2: A_hi.W_lo
          tma_load_2d(smem_u32(sb), (pass == 1) ? &mapAlo : &mapAhi, acols_s[kc], mt * GM_BM, BAR(st));
          tma_load_2d(smem_u32(sb + GM_A_BYTES), (pass == 2) ? &mapWlo : &mapWhi, kc * GM_KC, nt * a.NT, BAR(st));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, it_t = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it_t) {
        const int acc = it_t & 1; const uint32_t tph = (it_t >> 1) & 1;
        mbar_wait(BAR(10 + acc), tph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int k = 0; k < n_k; ++k, ++it) {
          const int st = it % GM_STAGES; const uint32_t ph = (it / GM_STAGES) & 1;
          mbar_wait(BAR(st), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + st * GM_STAGE_BYTES), sw = sa + GM_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < GM_KC / 8; ++kk)
            umma_tf32(d_tmem, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sw + kk * 32), idesc, (k | kk) ? 1u : 0u);
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
        if (r < a.R) {
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
                if (a.residual) val += a.residual[r * a.res_ld + oj];
              }
              x[j] = val;
            }
            const bool full = (o + 3 < a.O);
            if (a.amax_arg) {          // running (max, argmax) over successive launches; strict '>' keeps the first maximum
              for (int j = 0; j < 4 && o + j < a.O; ++j) {
                const long long ix = r * a.raw_ld + o + j;
                if (x[j] > a.raw_out[ix]) { a.raw_out[ix] = x[j]; a.amax_arg[ix] = (uint8_t)a.amax_id; }
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
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
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
  const cuuint32_t box[2] = {(cuuint32_t)GM_KC, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled(gemm) failed (%d) rows=%lld kdim=%d box=%d", (int)r, rows, kdim, box_rows); return ROREG_ERR_CUDA; }
  return ROREG_OK;
}

// A_hi/A_lo: [R][Kdim]; W_hi/W_lo: [Opad][Kdim] with Opad = n_ntiles*NT rows allocated (rows >= O may be anything finite: masked).
static inline int gemm_tc_launch(roreg_ctx* c, const float* A_hi, const float* A_lo, const float* W_hi, const float* W_lo,
                                 long long w_rows, GemmArgs a, cudaStream_t st) {
  if (a.R <= 0) return ROREG_OK;
  if (a.Kdim / GM_KC > 256 || a.Kdim % GM_KC || a.NT % 16 || a.NT < 16 || a.NT > 256 || (a.npass != 1 && a.npass != 3) || (a.npass == 3 && (!A_lo || !W_lo))) {
    snprintf(c->err, sizeof(c->err), "gemm_tc_launch: unsupported shape Kdim=%d NT=%d npass=%d", a.Kdim, a.NT, a.npass);
    return ROREG_ERR_UNSUPPORTED;
  }
  CUtensorMap mAh, mAl, mWh, mWl;
  int rc;
  if ((rc = gemm_make_map(c, &mAh, A_hi, a.R, a.Kdim, GM_BM))) return rc;
  if ((rc = gemm_make_map(c, &mAl, A_lo ? A_lo : A_hi, a.R, a.Kdim, GM_BM))) return rc;
  if ((rc = gemm_make_map(c, &mWh, W_hi, w_rows, a.Kdim, a.NT))) return rc;
  if ((rc = gemm_make_map(c, &mWl, W_lo ? W_lo : W_hi, w_rows, a.Kdim, a.NT))) return rc;
  static unsigned long long attr_mask = 0;
  if (rr_first_use_on_device(&attr_mask, c->device)) {
    RR_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES));
  }
  const int n_mt = (a.R + GM_BM - 1) / GM_BM;
  const long long tiles = (long long)n_mt * a.n_ntiles;
  const int grid = (int)(tiles < c->sm_count ? tiles : c->sm_count);
  gemm_tc_kernel<<<grid, 192, GM_SMEM_BYTES, st>>>(mAh, mAl, mWh, mWl, a);
  RR_LAUNCH_CHECK(c);
  return ROREG_OK;
}

}  // namespace roreg
