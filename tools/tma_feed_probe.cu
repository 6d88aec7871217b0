// Probe (GPU box), NOT part of the library: how many bytes per clock can ONE SM pull from L2 into shared memory, and does
// it depend on how the copy is described?  Motivation: every tcgen05 GEMM here settles at 31-38 B/clk of
// operand feed per SM whatever the number of active SMs, i.e. ~3.5 cycles per 128-byte box row.
//
//   mode 0: 2-D tensor loads, box 64 bf16 x 128 rows, SWIZZLE_128B (the operand tiles of the GEMMs: 128 rows of 128 B)
//   mode 1: 1-D bulk copies (cp.async.bulk.shared::cluster.global), 16 KB contiguous per instruction
//   mode 2: 2-D tensor loads, box 256 B x 64 rows, no swizzle (half as many, longer rows)
//   mode 3: 1-D bulk copies of 2 KB (8 per 16 KB)
//
// One thread per CTA keeps a ring of STAGES x 48 KB loads in flight over an L2-resident buffer; nothing reads the data.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I regennet_b200/csrc -o /tmp/tma_feed_probe tools/tma_feed_probe.cu -lcuda
//   /tmp/tma_feed_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace regen;

constexpr int STAGES = 4, STAGE_BYTES = 3 * 16384, ITERS = 400;

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ptx::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm_sw, const __grid_constant__ CUtensorMap tm_wide, const uint8_t* g,
             size_t g_bytes, int mode, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(&bar[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  // the buffer is viewed as rows of 128 bytes (mode 0: [rows, 64 bf16]) or rows of 256 bytes (mode 2)
  const size_t tiles = g_bytes / 16384;
  size_t tile = ((size_t)blockIdx.x * 977) % tiles;
  const unsigned long long t0 = clock64(), n0 = ptx::globaltimer_ns();
  for (int it = 0; it < ITERS; ++it) {
    const int s = it % STAGES;
    if (it >= STAGES) ptx::mbar_wait(&bar[s], ((it / STAGES) - 1) & 1);
    ptx::mbar_expect_tx(&bar[s], STAGE_BYTES);
    for (int j = 0; j < 3; ++j) {
      uint8_t* dst = smem + s * STAGE_BYTES + j * 16384;
      if (mode == 0) {
        ptx::tma_load_2d(dst, &tm_sw, &bar[s], 0, (int)(tile * 128));
      } else if (mode == 1) {
        bulk_load_1d(dst, g + tile * 16384, 16384, &bar[s]);
      } else if (mode == 2) {
        ptx::tma_load_2d(dst, &tm_wide, &bar[s], 0, (int)(tile * 64));
      } else {
        for (int c = 0; c < 8; ++c) bulk_load_1d(dst + c * 2048, g + tile * 16384 + c * 2048, 2048, &bar[s]);
      }
      tile += 131;
      if (tile >= tiles) tile -= tiles;
    }
  }
  for (int it = ITERS - STAGES; it < ITERS; ++it) ptx::mbar_wait(&bar[it % STAGES], (it / STAGES) & 1);
  const unsigned long long t1 = clock64(), n1 = ptx::globaltimer_ns();
  out[2 * blockIdx.x] = t1 - t0;
  out[2 * blockIdx.x + 1] = n1 - n0;
}

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

int main() {
  const size_t g_bytes = 48u << 20;  // L2-resident after the first pass
  uint8_t* g;
  CK(cudaMalloc(&g, g_bytes));
  CK(cudaMemset(g, 1, g_bytes));
  unsigned long long* out;
  CK(cudaMalloc(&out, 148 * 2 * 8));
  CUtensorMap tm_sw, tm_wide;
  {
    cuuint64_t dims[2] = {64, g_bytes / 128};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm_sw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode sw failed %d\n", (int)r); return 1; }
    cuuint64_t dims2[2] = {128, g_bytes / 256};
    cuuint64_t strides2[1] = {256};
    cuuint32_t box2[2] = {128, 64};
    r = cuTensorMapEncodeTiled(&tm_wide, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims2, strides2, box2, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode wide failed %d\n", (int)r); return 1; }
  }
  const int smem_bytes = STAGES * STAGE_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const char* names[4] = {"2-D tensor, 128 rows x 128 B, SWIZZLE_128B", "1-D bulk, 16 KB per copy",
                          "2-D tensor, 64 rows x 256 B, no swizzle", "1-D bulk, 2 KB per copy"};
  for (int grid : {8, 148}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {  // rep 0 warms L2
        probe_kernel<<<grid, 128, smem_bytes>>>(tm_sw, tm_wide, g, g_bytes, mode, out);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
      }
      std::vector<unsigned long long> h(2 * grid);
      CK(cudaMemcpy(h.data(), out, 2 * grid * 8, cudaMemcpyDeviceToHost));
      double cyc = 0, ns = 0;
      for (int i = 0; i < grid; ++i) { cyc += h[2 * i]; ns += h[2 * i + 1]; }
      cyc /= grid; ns /= grid;
      const double bytes = (double)ITERS * STAGE_BYTES;
      printf("grid %3d  %-44s %6.1f B/clk/SM  %6.1f GB/s/SM  (%.2f cycles per 128 B; chip %.2f TB/s)\n", grid, names[mode],
             bytes / cyc, bytes / ns, cyc / (bytes / 128), bytes / ns * grid / 1e3);
    }
  }
  return 0;
}
