// gather4_test.cu - what does cp.async.bulk.tensor.2d ... tile::gather4 write, and where?  (B200 micro-test, not part of the product)
// Source: float matrix [ROWS][64].  Tensor map: 2-D, SWIZZLE_128B, box = {32 floats, BOXR rows} with BOXR in {1, 4}.
// Kernel: lane 0 issues two gather4 copies (rows {5,99,3,42} -> smem offset 0, rows {7,1,250,8} -> smem offset 512) of column 32,
// waits on the mbarrier (bounded) and dumps 1024 B of shared memory.  The host prints which (row, 16-byte chunk) each chunk of
// the dump holds, i.e. whether rows land contiguously at 128 B pitch with the SWIZZLE_128B XOR applied on absolute address bits.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int ROWS = 256, COLS = 64;
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int* status) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(smem);
  for (int i = threadIdx.x; i < 512; i += 32) reinterpret_cast<float*>(smem)[i] = -1.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(1024) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(d), "l"(&map), "r"(32), "r"(5), "r"(99), "r"(3), "r"(42), "r"(b) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(d + 512), "l"(&map), "r"(32), "r"(7), "r"(1), "r"(250), "r"(8), "r"(b) : "memory");
    uint32_t done = 0; long long t0 = clock64();
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(0) : "memory");
      if (clock64() - t0 > 200000000LL) break;
    }
    *status = done;
  }
  __syncwarp();
  for (int i = threadIdx.x; i < 256; i += 32) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)p;
  std::vector<float> h(ROWS * COLS);
  for (int r = 0; r < ROWS; ++r) for (int c = 0; c < COLS; ++c) h[r * COLS + c] = r * 100.f + c;     // value encodes (row, col)
  float *d, *out; int* st;
  cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 1024); cudaMalloc(&st, 4);
  for (int boxr : {1, 4}) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {COLS, ROWS}; const cuuint64_t strides[1] = {COLS * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)boxr}; const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box rows %d: encode rc=%d\n", boxr, (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaMemset(st, 0xff, 4);
    k<<<1, 32, 4096>>>(m, out, st);
    cudaError_t e = cudaDeviceSynchronize();
    int hs = -7; float ho[256];
    if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, 1024, cudaMemcpyDeviceToHost);
    printf("  barrier completed: %d\n", hs);
    for (int chunk = 0; chunk < 64; ++chunk) {          // 16-byte chunks of the 1024-byte dump
      const float v = ho[chunk * 4];
      if (chunk % 8 == 0) printf("  smem row %d (offset %4d):", chunk / 8, chunk * 16);
      if (v < 0) printf(" [ -- ]"); else printf(" [r%3d c%2d]", (int)(v / 100.f), (int)v % 100);
      if (chunk % 8 == 7) printf("\n");
    }
  }
  return 0;
}
