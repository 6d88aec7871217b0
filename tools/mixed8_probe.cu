// Probe (GPU box), NOT part of the library: does the "fp16 main term + two plain-e4m3 correction products" scheme of
// DESIGN.md section 3 work on the tensor core the way tools/precision_probe.py emulates it on the CPU?
//
//   D = [ (A_lo * 2^11) . (W_hi * 2^4)^T  +  A_hi . (W_lo * 2^15)^T ]      kind::f8f6f4, e4m3 x e4m3, K = 32 per MMA
//   D = A_hi . W_hi^T + D * 2^-15                                          first kind::f16 MMA, scale-input-d = 15
//   D += A_hi . W_hi^T                                                     remaining kind::f16 MMAs
//
// i.e. 2 bf16-MMA equivalents per product instead of the 3 of bf16x3 / fp16x3, at the same operand bytes (2 + 1 + 1 per
// element).  What the probe checks on the device: (1) the semantics of the scale-input-d immediate, (2) fp8 operands
// in the K-major SWIZZLE_128B layout (128 elements per 128-byte row, +32 bytes per K = 32 step), (3) that f8f6f4 and f16
// MMAs may accumulate into the same TMEM columns.  One CTA, one 128 x 128 tile, K = 512, no TMA and no pipelining
// (operand tiles are written to shared memory by the threads): a correctness probe, not a benchmark.
//
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I regennet_b200/csrc -o /tmp/mixed8_probe tools/mixed8_probe.cu
//   /tmp/mixed8_probe          # prints the max abs error vs float64 of: fp16 x1, fp16 x3, mixed8
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace regen;

constexpr int M = 128, N = 128, K = 512;

// kind::f16 with the scale-input-d immediate: D = A.B + D * 2^-15
__device__ __forceinline__ void mma_f16_ss_scale15(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 15;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
// kind::f8f6f4 (plain, not block scaled): e4m3 x e4m3 -> fp32, K = 32 per instruction
__device__ __forceinline__ void mma_f8_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// c_format F32, a_format = b_format = 0 (F16 for kind::f16, E4M3 for kind::f8f6f4), both K-major
__host__ __device__ constexpr uint32_t idesc_fmt0(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// 128 rows x 128 bytes, SWIZZLE_128B: 16-byte chunk c of row r lives at r * 128 + ((c ^ (r & 7)) << 4)
__device__ __forceinline__ void fill_tile(uint8_t* tile, const uint8_t* src, size_t row_pitch_bytes, size_t col_byte0) {
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)r * row_pitch_bytes + col_byte0 + 16 * c);
    *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
}

// mode 0: fp16 x1    mode 1: fp16 x3 (A_lo.W_hi + A_hi.W_lo + A_hi.W_hi, all fp16)    mode 2: mixed8
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __half* a16, const __half* w16, const __half* alo16, const __half* wlo16, const uint8_t* a_hi8,
             const uint8_t* a_lo8, const uint8_t* w_hi8, const uint8_t* w_lo8, float* out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* t0 = smem;             // four 16 KB operand tiles
  uint8_t* t1 = smem + 16384;
  uint8_t* t2 = smem + 2 * 16384;
  uint8_t* t3 = smem + 3 * 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * 16384);
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_base_smem, 128);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t acc = *tmem_base_smem;
  constexpr uint32_t idesc = idesc_fmt0(M, N);
  uint32_t phase = 0;
  bool first = true;   // no MMA has written the accumulator yet

  auto round_done = [&]() {   // all MMAs issued so far have completed: the operand tiles may be overwritten
    if (threadIdx.x == 0) ptx::tcgen05_commit(bar);
    ptx::mbar_wait(bar, phase);
    phase ^= 1;
    ptx::tcgen05_fence_after();
  };

  if (mode == 2) {
    // ---- phase 1: correction products in e4m3, 128 K elements (= one 128-byte row) per round
    for (int kb = 0; kb < K / 128; ++kb) {
      fill_tile(t0, a_lo8, K, (size_t)kb * 128);
      fill_tile(t1, w_hi8, K, (size_t)kb * 128);
      fill_tile(t2, a_hi8, K, (size_t)kb * 128);
      fill_tile(t3, w_lo8, K, (size_t)kb * 128);
      ptx::fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        ptx::tcgen05_fence_after();
        for (int k = 0; k < 4; ++k) {   // K = 32 e4m3 elements = 32 bytes per step
          const uint32_t adv = (uint32_t)k * 32;
          mma_f8_ss(acc, ptx::umma_desc_k_sw128(ptx::smem_u32(t0) + adv), ptx::umma_desc_k_sw128(ptx::smem_u32(t1) + adv),
                    idesc, first ? 0u : 1u);
          first = false;
          mma_f8_ss(acc, ptx::umma_desc_k_sw128(ptx::smem_u32(t2) + adv), ptx::umma_desc_k_sw128(ptx::smem_u32(t3) + adv),
                    idesc, 1u);
        }
      }
      first = false;
      round_done();
    }
  }
  // ---- main term (and, in mode 1, the fp16 correction products): 64 K elements per round
  bool scale_pending = mode == 2;   // the first main-term MMA folds the scaled corrections in: D = A.B + D * 2^-15
  for (int kb = 0; kb < K / 64; ++kb) {
    fill_tile(t0, reinterpret_cast<const uint8_t*>(a16), (size_t)K * 2, (size_t)kb * 128);
    fill_tile(t1, reinterpret_cast<const uint8_t*>(w16), (size_t)K * 2, (size_t)kb * 128);
    if (mode == 1) {
      fill_tile(t2, reinterpret_cast<const uint8_t*>(alo16), (size_t)K * 2, (size_t)kb * 128);
      fill_tile(t3, reinterpret_cast<const uint8_t*>(wlo16), (size_t)K * 2, (size_t)kb * 128);
    }
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      ptx::tcgen05_fence_after();
      for (int k = 0; k < 4; ++k) {     // K = 16 fp16 elements = 32 bytes per step
        const uint32_t adv = (uint32_t)k * 32;
        const uint64_t a_hi = ptx::umma_desc_k_sw128(ptx::smem_u32(t0) + adv);
        const uint64_t w_hi = ptx::umma_desc_k_sw128(ptx::smem_u32(t1) + adv);
        if (mode == 1) {
          const uint64_t a_lo = ptx::umma_desc_k_sw128(ptx::smem_u32(t2) + adv);
          const uint64_t w_lo = ptx::umma_desc_k_sw128(ptx::smem_u32(t3) + adv);
          ptx::mma_f16_ss(acc, a_lo, w_hi, idesc, first ? 0u : 1u);
          first = false;
          ptx::mma_f16_ss(acc, a_hi, w_lo, idesc, 1u);
        }
        if (scale_pending) {
          mma_f16_ss_scale15(acc, a_hi, w_hi, idesc);
          scale_pending = false;
        } else {
          ptx::mma_f16_ss(acc, a_hi, w_hi, idesc, first ? 0u : 1u);
        }
        first = false;
      }
    }
    first = false;
    scale_pending = false;
    round_done();
  }

  // ---- read the accumulator: warp w owns TMEM lanes 32 w .. 32 w + 31 (= rows), 128 columns
  const uint32_t lane_addr = acc + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    __syncwarp();
    ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)c0, v);
    ptx::tmem_ld_wait(v);
    for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(acc, 128);
  }
}

static uint8_t e4m3(float v) {
  if (v > 448.f) v = 448.f;      // the fp8 copies are clamped
  if (v < -448.f) v = -448.f;
  return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
}

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

int main() {
  std::vector<float> A((size_t)M * K), W((size_t)N * K);
  srand(1);
  auto gauss = []() {
    float u = (rand() + 1.0f) / (RAND_MAX + 2.0f), v = (rand() + 1.0f) / (RAND_MAX + 2.0f);
    return sqrtf(-2.f * logf(u)) * cosf(6.2831853f * v);
  };
  for (auto& x : A) x = gauss();
  for (auto& x : W) x = 0.04f * gauss();
  std::vector<__half> a16(A.size()), w16(W.size()), alo16(A.size()), wlo16(W.size());
  std::vector<uint8_t> a_hi8(A.size()), a_lo8(A.size()), w_hi8(W.size()), w_lo8(W.size());
  for (size_t i = 0; i < A.size(); ++i) {
    a16[i] = __float2half_rn(A[i]);
    const float hi = __half2float(a16[i]), lo = A[i] - hi;
    alo16[i] = __float2half_rn(lo);
    a_hi8[i] = e4m3(hi);
    a_lo8[i] = e4m3(lo * 2048.f);          // 2^11
  }
  for (size_t i = 0; i < W.size(); ++i) {
    w16[i] = __float2half_rn(W[i]);
    const float hi = __half2float(w16[i]), lo = W[i] - hi;
    wlo16[i] = __float2half_rn(lo);
    w_hi8[i] = e4m3(hi * 16.f);            // 2^4
    w_lo8[i] = e4m3(lo * 32768.f);         // 2^15
  }
  std::vector<double> ref((size_t)M * N);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(size_t)i * K + k] * (double)W[(size_t)j * K + k];
      ref[(size_t)i * N + j] = s;
    }
  __half *d_a16, *d_w16, *d_alo, *d_wlo;
  uint8_t *d_ah8, *d_al8, *d_wh8, *d_wl8;
  float* d_out;
  CK(cudaMalloc(&d_a16, A.size() * 2)); CK(cudaMalloc(&d_w16, W.size() * 2));
  CK(cudaMalloc(&d_alo, A.size() * 2)); CK(cudaMalloc(&d_wlo, W.size() * 2));
  CK(cudaMalloc(&d_ah8, A.size())); CK(cudaMalloc(&d_al8, A.size()));
  CK(cudaMalloc(&d_wh8, W.size())); CK(cudaMalloc(&d_wl8, W.size()));
  CK(cudaMalloc(&d_out, (size_t)M * N * 4));
  CK(cudaMemcpy(d_a16, a16.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w16, w16.data(), W.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_alo, alo16.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wlo, wlo16.data(), W.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ah8, a_hi8.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_al8, a_lo8.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wh8, w_hi8.data(), W.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wl8, w_lo8.data(), W.size(), cudaMemcpyHostToDevice));
  const int smem_bytes = 4 * 16384 + 1024 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const char* names[3] = {"fp16 x1 (1 MMA / product)", "fp16 x3 (3 MMAs / product)", "fp16 + e4m3 corrections (2 MMA equivalents)"};
  std::vector<float> out((size_t)M * N);
  for (int mode = 0; mode < 3; ++mode) {
    CK(cudaMemset(d_out, 0xff, (size_t)M * N * 4));
    probe_kernel<<<1, 128, smem_bytes>>>(d_a16, d_w16, d_alo, d_wlo, d_ah8, d_al8, d_wh8, d_wl8, d_out, mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), d_out, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    double err = 0, mag = 0;
    for (size_t i = 0; i < out.size(); ++i) {
      err = fmax(err, fabs((double)out[i] - ref[i]));
      mag = fmax(mag, fabs(ref[i]));
    }
    printf("%-48s max abs err vs float64 %.3e  (|result| max %.3f)\n", names[mode], err, mag);
  }
  printf("expected from the CPU emulation: x1 ~1e-3 relative, x3 ~1e-6, mixed8 within a small factor of x3\n");
  return 0;
}
