// tmem_bench.cu - microbenchmark behind the NN-epilogue design: how fast can epilogue warps read accumulators
// out of TMEM (tcgen05.ld 32x32b.x32) on a B200 SM, alone / with min-reduction math / with the tensor pipe
// issuing 128x128x8 tf32 MMAs at the same time.  Prints bytes/clk/SM for every variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/tmem_bench scripts/tmem_bench.cu
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../roreg_b200/csrc/kernels_nn_tc.cuh"
using namespace roreg;

// OPS: 0 = load only, 1 = row min (FMNMX3 tree), 2 = FADD + row min (nn mode 2's inner loop),
//      4 = v = na - g; column running min + predicated tile index; row min of v + nb  (the one-Gram sweep epilogue)
template <int NW, int OPS, bool MMA>
__global__ void __launch_bounds__(NW * 32 + 32, 1) tmem_bench_kernel(long long* clk, float* sink, unsigned* mma_tiles, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  __shared__ volatile int stop;
  __shared__ __align__(16) float nb_s[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 64) nb_s[threadIdx.x] = 0.001f * threadIdx.x;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f + (i & 15);
  if (warp == NW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const long long t0 = clock64();
  if (warp == NW) {
    if (MMA && lane == 0) {
      unsigned tiles = 0; uint32_t ph = 0;
      const uint32_t a0 = smem_u32(smem), b0 = a0 + 16384;
      while (!stop) {
        const uint32_t d = tmem_base + 256 + (tiles & 1) * 128;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_tf32(d, umma_desc_sw128(a0 + kk * 32), umma_desc_sw128(b0 + kk * 32), TC_IDESC, (c | kk) ? 1u : 0u);
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), ph); ph ^= 1;
        ++tiles;
      }
      mma_tiles[blockIdx.x] = tiles;
    }
  } else {
    const int q = warp & 3, cg = warp >> 2;
    float best = INFINITY; int bestj = 0;
    float cmin[OPS == 4 ? 64 : 1]; uint32_t ctile[OPS == 4 ? 16 : 1];
    if (OPS == 4) {
#pragma unroll
      for (int e = 0; e < 64; ++e) cmin[e] = INFINITY;
#pragma unroll
      for (int e = 0; e < 16; ++e) ctile[e] = 0;
    }
    for (int it = 0; it < iters; ++it) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (((it & 1) * 128 + cg * 64) & 255);
      uint32_t r0[32], r1[32];
      RR_TMEM_LD32(r0, taddr);
      RR_TMEM_LD32(r1, taddr + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (OPS == 1) {
        float m = INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fminf(m, fminf(__uint_as_float(r0[j]), __uint_as_float(r1[j])));
        if (m < best) { best = m; bestj = it; }
      } else if (OPS == 2) {
        const float4* nb4 = reinterpret_cast<const float4*>(nb_s);
        float m = INFINITY;
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 nb = nb4[j4];
          const uint32_t* rg = (j4 < 8) ? (r0 + 4 * j4) : (r1 + 4 * (j4 - 8));
          m = fminf(m, fminf(fminf(nb.x - __uint_as_float(rg[0]), nb.y - __uint_as_float(rg[1])),
                             fminf(nb.z - __uint_as_float(rg[2]), nb.w - __uint_as_float(rg[3]))));
        }
        if (m < best) { best = m; bestj = it; }
      } else if (OPS == 4) {
        const float na = 0.25f + 1e-3f * it;
        const float4* nb4 = reinterpret_cast<const float4*>(nb_s);
        float m = INFINITY;
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 nb = nb4[j4];
          const uint32_t* rg = (j4 < 8) ? (r0 + 4 * j4) : (r1 + 4 * (j4 - 8));
          float v[4]; const float nbv[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            v[u] = na - __uint_as_float(rg[u]);
            const int e = 4 * j4 + u;
            if (v[u] < cmin[e]) { cmin[e] = v[u]; ctile[e >> 2] = __byte_perm(ctile[e >> 2], (uint32_t)it, (0x3210 & ~(0xF << (4 * u))) | (4 << (4 * u))); }
            m = fminf(m, v[u] + nbv[u]);
          }
        }
        if (m < best) { best = m; bestj = it; }
      } else {
        best += __uint_as_float(r0[lane & 31 ? 3 : 5]) + __uint_as_float(r1[7]);
      }
    }
    float acc = best + bestj;
    if (OPS == 4) {
#pragma unroll
      for (int e = 0; e < 64; ++e) acc += cmin[e];
#pragma unroll
      for (int e = 0; e < 16; ++e) acc += (float)ctile[e];
    }
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");
    if (threadIdx.x == 0) { clk[blockIdx.x] = clock64() - t0; stop = 1; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == NW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

template <int NW, int OPS, bool MMA>
static void run(const char* name, int sms) {
  const int iters = 4000;
  long long* clk; float* sink; unsigned* tiles;
  cudaMalloc(&clk, sms * sizeof(long long)); cudaMalloc(&sink, (size_t)sms * 1024 * sizeof(float)); cudaMalloc(&tiles, sms * sizeof(unsigned));
  cudaMemset(tiles, 0, sms * sizeof(unsigned));
  auto kern = tmem_bench_kernel<NW, OPS, MMA>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024);
  for (int rep = 0; rep < 2; ++rep) kern<<<sms, NW * 32 + 32, 34 * 1024>>>(clk, sink, tiles, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-40s FAILED: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(sms); std::vector<unsigned> ht(sms);
  cudaMemcpy(h.data(), clk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemcpy(ht.data(), tiles, sms * sizeof(unsigned), cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double med = (double)h[sms / 2];
  const double bytes = (double)NW * iters * 64 * 32 * 4;           // per SM
  printf("%-44s warps %2d  clk/iter %8.1f  TMEM read %7.1f B/clk/SM", name, NW, med / iters, bytes / med);
  if (MMA) printf("  | concurrent MMA: %.1f clk per 128x128x96 tile", med / std::max(1u, ht[sms / 2]));
  printf("\n");
  cudaFree(clk); cudaFree(sink); cudaFree(tiles);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs; each reader warp: 2 x tcgen05.ld.32x32b.x32 (8 KB) per iteration\n", p.name, sms);
  run<4, 0, false>("load only", sms);
  run<8, 0, false>("load only", sms);
  run<16, 0, false>("load only", sms);
  run<4, 1, false>("row min (FMNMX3)", sms);
  run<8, 1, false>("row min (FMNMX3)", sms);
  run<8, 2, false>("FADD + row min (mode-2 inner loop)", sms);
  run<8, 4, false>("sweep epilogue (col state + row min)", sms);
  run<8, 0, true>("load only + MMA", sms);
  run<8, 1, true>("row min + MMA", sms);
  run<8, 2, true>("FADD + row min + MMA", sms);
  run<8, 4, true>("sweep epilogue + MMA", sms);
  run<0 + 4, 4, true>("sweep epilogue + MMA", sms);
  return 0;
}
