// gather_bench.cu - what bandwidth can an SM-resident TMA pipeline reach when it GATHERS 7680-byte descriptor rows
// (the Des2R access pattern: per match one row of cloud id1 in random order + one row of cloud id0 in ascending
// order)?  Bare skeleton: one producer thread issues cp.async.bulk copies into a ring of shared-memory stages, one
// consumer thread releases every stage the moment it lands.  No math, no tensor cores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/gather_bench scripts/gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>
#include "../roreg_b200/csrc/kernels_nn_tc.cuh"
using namespace roreg;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// rows[item][4] : the four descriptor rows of an item (X0, Y0, X1, Y1).  mode 0: TMA bulk copies; mode 1: LDG.128 by 4 warps
__global__ void __launch_bounds__(192, 1) gather_kernel(const float* __restrict__ desc, const int* __restrict__ rows, int n_items,
                                                        int stages, int mode, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[16];
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto FREE = [&](int s) { return bar0 + 8u * (8 + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(FULL(s), 1); mbar_init(FREE(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mode == 0) {
    if (warp == 0) {
      uint32_t it = 0;
      for (int base = blockIdx.x; base < n_items; base += 32 * gridDim.x) {
        const int item = base + lane * gridDim.x;
        int4 r = make_int4(0, 0, 0, 0);
        if (item < n_items) r = reinterpret_cast<const int4*>(rows)[item];
        for (int l = 0; l < 32; ++l) {
          const int x0 = __shfl_sync(0xffffffffu, r.x, l), y0 = __shfl_sync(0xffffffffu, r.y, l);
          const int x1 = __shfl_sync(0xffffffffu, r.z, l), y1 = __shfl_sync(0xffffffffu, r.w, l);
          if (base + l * (int)gridDim.x >= n_items) break;
          if (lane == 0) {
            const int st = it % stages; const uint32_t ph = (it / stages) & 1;
            mbar_wait(FREE(st), ph ^ 1);
            uint8_t* sb = smem + st * 30720;
            mbar_expect_tx(FULL(st), 30720);
            bulk_load_1d(smem_u32(sb), desc + (long long)x0 * 1920, 7680, FULL(st));
            bulk_load_1d(smem_u32(sb + 7680), desc + (long long)y0 * 1920, 7680, FULL(st));
            bulk_load_1d(smem_u32(sb + 15360), desc + (long long)x1 * 1920, 7680, FULL(st));
            bulk_load_1d(smem_u32(sb + 23040), desc + (long long)y1 * 1920, 7680, FULL(st));
          }
          ++it;
          __syncwarp();
        }
      }
    } else if (warp == 1 && lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int st = it % stages; const uint32_t ph = (it / stages) & 1;
        mbar_wait(FULL(st), ph);
        mbar_arrive(FREE(st));
      }
    }
  } else {
    // LDG gather: warps 0..3 each stream one of the item's four rows (15 x 16 B per lane), `stages` items unrolled in flight
    if (warp < 4) {
      float acc = 0.f;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int row = rows[item * 4 + warp];
        const float4* src = reinterpret_cast<const float4*>(desc + (long long)row * 1920);
        float4 v[15];
#pragma unroll
        for (int k = 0; k < 15; ++k) v[k] = ldg_stream4(src + k * 32 + lane);
#pragma unroll
        for (int k = 0; k < 15; ++k) acc += v[k].x + v[k].w;
      }
      sink[blockIdx.x * 192 + threadIdx.x] = acc;
    }
  }
}

// L2-resident tile stream (the nn mode-4 pattern): groups of `share` CTAs stream the same 40 x tile_bytes image again and again.
__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* __restrict__ img, int tile_bytes, int n_tiles, int reps, int share, int stages, int split) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[16];
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto FREE = [&](int s) { return bar0 + 8u * (8 + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(FULL(s), 1); mbar_init(FREE(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* base = img + (size_t)(blockIdx.x / share) * n_tiles * tile_bytes;
  const int total = n_tiles * reps;
  if (threadIdx.x == 0) {
    for (int it = 0; it < total; ++it) {
      const int st = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(FREE(st), ph ^ 1);
      mbar_expect_tx(FULL(st), tile_bytes);
      const uint8_t* src = base + (size_t)(it % n_tiles) * tile_bytes;
      const int piece = tile_bytes / split;
      for (int k = 0; k < split; ++k) bulk_load_1d(smem_u32(smem + st * tile_bytes + k * piece), src + k * piece, piece, FULL(st));
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < total; ++it) {
      const int st = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(FULL(st), ph);
      mbar_arrive(FREE(st));
    }
  }
}

static void run_stream(int sms) {
  const int n_tiles = 40, reps = 20;
  uint8_t* img; cudaMalloc(&img, (size_t)148 * n_tiles * 49152); cudaMemset(img, 1, (size_t)148 * n_tiles * 49152);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int tbs[2] = {32768, 16384};
  for (int tb : tbs)
    for (int share : {1, 8, 40, 148})
      for (int stages : {3, 5})
        for (int split : {1, 4}) {
          if (stages * tb > 196608) continue;
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          const size_t sm = (size_t)stages * tb + 1024;
          stream_kernel<<<sms, 64, sm>>>(img, tb, n_tiles, 2, share, stages, split);
          cudaEventRecord(e0);
          stream_kernel<<<sms, 64, sm>>>(img, tb, n_tiles, reps, share, stages, split);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("FAILED: %s\n", cudaGetErrorString(e)); exit(1); }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          const double bytes = (double)sms * n_tiles * reps * tb;
          printf("stream tile %5d B  share %3d CTAs/image  stages %d  copies/tile %d : %.3f ms  %.0f GB/s  %.1f B/clk/SM (at 1.965 GHz)\n", tb, share, stages,
                 split, ms, bytes / ms / 1e6, bytes / ms / 1e6 / sms / 1.965);
        }
  cudaFree(img);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 's') { cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); run_stream(p.multiProcessorCount); return 0; }
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const int B = 32, n = 5000, K = 3400;
  const long long rows_total = (long long)B * 2 * n;
  float* desc; cudaMalloc(&desc, rows_total * 7680);
  cudaMemset(desc, 0, rows_total * 7680);
  std::mt19937 rng(1);
  const int items_per_pair = K / 2, n_items = B * items_per_pair;
  auto build = [&](int order) {   // order 0: Des2R pattern (X random, Y ascending); 1: both ascending; 2: both random
    std::vector<int> rows((size_t)n_items * 4);
    for (int b = 0; b < B; ++b) {
      std::vector<int> a(n), c(n);
      std::iota(a.begin(), a.end(), 0); std::iota(c.begin(), c.end(), 0);
      std::shuffle(a.begin(), a.end(), rng); std::shuffle(c.begin(), c.end(), rng);
      a.resize(K); c.resize(K);
      if (order != 2) std::sort(a.begin(), a.end());
      if (order == 1) std::sort(c.begin(), c.end());
      for (int i = 0; i < items_per_pair; ++i) {
        int* r = &rows[((size_t)b * items_per_pair + i) * 4];
        r[0] = (2 * b + 1) * n + c[2 * i]; r[1] = (2 * b) * n + a[2 * i];
        r[2] = (2 * b + 1) * n + c[2 * i + 1]; r[3] = (2 * b) * n + a[2 * i + 1];
      }
    }
    return rows;
  };
  int* d_rows; cudaMalloc(&d_rows, (size_t)n_items * 16);
  float* sink; cudaMalloc(&sink, sms * 192 * 4 * 4);
  cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 30720 + 1024);
  printf("%s, %d SMs; %d items x 4 rows x 7680 B = %.2f GB per launch\n", p.name, sms, n_items, n_items * 30720.0 / 1e9);
  const char* on[3] = {"X random / Y ascending (Des2R)", "both ascending", "both random"};
  for (int order = 0; order < 3; ++order) {
    std::vector<int> rows = build(order);
    cudaMemcpy(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode)
      for (int stages = (mode ? 1 : 1); stages <= (mode ? 1 : 7); ++stages) {
        for (int grid_mul = 1; grid_mul <= (mode ? 4 : 1); grid_mul *= 2) {
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          const size_t sm = mode ? 1024 : (size_t)stages * 30720 + 1024;
          gather_kernel<<<sms * grid_mul, 192, sm>>>(desc, d_rows, n_items, stages, mode, sink);
          cudaEventRecord(e0);
          for (int rep = 0; rep < 5; ++rep) gather_kernel<<<sms * grid_mul, 192, sm>>>(desc, d_rows, n_items, stages, mode, sink);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("FAILED: %s\n", cudaGetErrorString(e)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
          printf("%-32s %s stages %d ctas/SM %d : %.3f ms  %.0f GB/s\n", on[order], mode ? "LDG.128 " : "TMA bulk", stages, grid_mul, ms,
                 n_items * 30720.0 / ms / 1e6);
        }
      }
  }
  return 0;
}
