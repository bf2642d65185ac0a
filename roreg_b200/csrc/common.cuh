// common.cuh - context, error handling and small device helpers for libroreg_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/roreg_b200.h"

#define RR_F 32
#define RR_G 60
#define RR_ROW (RR_F * RR_G)      // floats per keypoint descriptor (7680 B)

struct roreg_ctx {
  int device;
  int sm_count;
  uint8_t h_perm8[3600]; // host copy of P
  uint8_t* d_perm8;     // [60][60]  P[a][g]            (variant 1 table)
  uint8_t* d_permT8;    // [60][60]  P[g][h] stored [h][g]  (variant 2 table)
  int32_t* d_nei;       // [60][13]
  float* d_rot32;       // [60][9]
  double* d_rot64;      // [60][9]
  void* ws;             // workspace arena (grown on demand)
  size_t ws_bytes;
  int64_t launches;
  int corr_mode;                  // 0 = FP32 CUDA-core Gram, 1 / 2 = tcgen05 3xTF32 Gram, 2-match / 1-match items (roreg_set_corr_mode)
  int score_mode;                 // 0 = float64 one-shot scoring, 1 = float32 pre-filter + exact float64 re-check (roreg_set_score_mode)
  int timing;                     // roreg_set_timing: record an event after every stage of roreg_register_batch
  cudaEvent_t ev[ROREG_N_STAGES + 1];
  int ev_valid;
  // two-stream schedule of roreg_register_batch (roreg_set_overlap): tensor-core chain vs pooling / RANSAC tail
  int overlap;                    // 1 (default): split a batch in two halves and overlap their stages on s_tc / s_light
  cudaStream_t s_tc, s_light;     // created on first use: s_tc high priority, s_light low priority
  cudaEvent_t ev_fork, ev_pool[2], ev_corr[2], ev_join[2];
  void* pipe;                     // state of roreg_register_batch_pipelined (PipeState, roreg_capi.cu), created on first use
  char err[512];
};

#define RR_CUDA(ctx, expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s -> %s", __FILE__, __LINE__, #expr,     \
               cudaGetErrorString(_e));                                                         \
      return ROREG_ERR_CUDA;                                                                    \
    }                                                                                           \
  } while (0)

#define RR_ARG(ctx, cond)                                                                       \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d bad argument: %s", __FILE__, __LINE__, #cond); \
      return ROREG_ERR_ARG;                                                                     \
    }                                                                                           \
  } while (0)

#define RR_LAUNCH_CHECK(ctx)                                                                    \
  do {                                                                                          \
    (ctx)->launches++;                                                                          \
    RR_CUDA(ctx, cudaGetLastError());                                                           \
  } while (0)

static inline int rr_ws_reserve(roreg_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return ROREG_OK;
  if (ctx->ws) {
    // the arena may still be in use by enqueued work: drain before freeing
    cudaDeviceSynchronize();
    cudaFree(ctx->ws);
    ctx->ws = nullptr; ctx->ws_bytes = 0;
  }
  size_t want = bytes + bytes / 4 + (1 << 20);
  cudaError_t e = cudaMalloc(&ctx->ws, want);
  if (e != cudaSuccess) {
    snprintf(ctx->err, sizeof(ctx->err), "workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    return ROREG_ERR_NOMEM;
  }
  ctx->ws_bytes = want;
  return ROREG_OK;
}

// bump allocator over the workspace (256-byte aligned slices)
struct rr_arena {
  char* base; size_t off;
  template <typename T> T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
};
static inline size_t rr_align(size_t b) { return (b + 255) & ~size_t(255); }

// cudaFuncSetAttribute applies to the CURRENT device only: remember per device (bit d of a process-wide mask), not per process
static inline bool rr_first_use_on_device(unsigned long long* mask, int device) {
  const unsigned long long bit = 1ull << (device & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}


#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;   // valid in lane 0
}
// splitmix64: counter-based stream for the device-side RANSAC draws
__host__ __device__ __forceinline__ uint64_t rr_mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ double rr_u01(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  uint64_t h = rr_mix64(seed ^ rr_mix64(a * 0x100000001B3ull + rr_mix64(b * 0x9E3779B1ull + c)));
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
#endif
