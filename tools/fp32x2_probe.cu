// Microbenchmark (GPU box): issue throughput of scalar FFMA / FADD against the packed FFMA2 / FADD2 (fma.rn.f32x2,
// add.rn.f32x2) of sm_100a, at the occupancy of the fused GEMM+LN epilogue (8 or 16 warps per SM, one CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp32x2_probe.bin tools/fp32x2_probe.cu && tools/fp32x2_probe.bin
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float add1(float a, float b) {
  float d;
  asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

constexpr int ILP = 8, ITERS = 2048;

template <int MODE>
__global__ void probe(float* out, long long* cycles, float seed) {
  float a[ILP], b = seed, c = seed * 0.5f;
  unsigned long long p[ILP], pb, pc;
  for (int i = 0; i < ILP; ++i) { a[i] = seed + i; float2 t = make_float2(seed + i, seed - i); p[i] = *reinterpret_cast<unsigned long long*>(&t); }
  { float2 t = make_float2(b, c); pb = *reinterpret_cast<unsigned long long*>(&t); pc = pb; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) a[i] = fma1(a[i], b, c);
      if (MODE == 1) a[i] = add1(a[i], b);
      if (MODE == 2) p[i] = fma2(p[i], pb, pc);
      if (MODE == 3) p[i] = add2(p[i], pb);
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
  for (int i = 0; i < ILP; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); acc += a[i] + t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const char* names[4] = {"FFMA  (scalar)", "FADD  (scalar)", "FFMA2 (f32x2)", "FADD2 (f32x2)"};
  for (int warps : {4, 8, 16, 32}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<148, warps * 32>>>(out, cyc, 1.0f);
        if (mode == 1) probe<1><<<148, warps * 32>>>(out, cyc, 1.0f);
        if (mode == 2) probe<2><<<148, warps * 32>>>(out, cyc, 1.0f);
        if (mode == 3) probe<3><<<148, warps * 32>>>(out, cyc, 1.0f);
      }
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double instr_per_smsp = (double)ITERS * ILP * (warps / 4.0);
      const double lanes = (mode >= 2 ? 2.0 : 1.0);
      printf("%2d warps/SM  %-15s %8lld cycles  %.2f cycles/instr/SMSP  %.1f fp32 lane-ops/clk/SM\n", warps, names[mode], c,
             c / instr_per_smsp, instr_per_smsp * 4 * 32 * lanes / c);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
