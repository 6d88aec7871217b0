// Shared helpers for libregen_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/regen_sm100.h"

namespace regen {

// thread-local last-error text returned by regen_last_error()
void set_error(const char* fmt, ...);

#define REGEN_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      ::regen::set_error(__VA_ARGS__);        \
      return REGEN_EINVAL;                    \
    }                                         \
  } while (0)

#define REGEN_CUDA(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::regen::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                         __LINE__);                                                           \
      return REGEN_ECUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define REGEN_LAUNCH_CHECK()                                                                  \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      ::regen::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),          \
                         __FILE__, __LINE__);                                                 \
      return REGEN_ECUDA;                                                                     \
    }                                                                                         \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// RAII: make `device` current for the scope and restore the caller's device afterwards, so a handle that lives on
// cuda:1 can be created / used / destroyed (possibly from a garbage collector) without flipping the thread's device.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) changed = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// number of kernels this library has launched in this process (regen_launch_count)
extern long long g_launches;
inline void count_launch(int n = 1) { g_launches += n; }

// Launch with programmatic stream serialization (PDL): the kernel's CTAs may be scheduled while the preceding kernel of
// the stream is still draining, so its prologue (barrier init, TMEM allocation, tensor-map prefetch) and the launch
// latency overlap that tail.  EVERY kernel launched through here executes ptx::griddep_wait() before its first global
// memory access.  REGEN_DEBUG_NO_PDL=1 turns the attribute off (A/B measurements).
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("REGEN_DEBUG_NO_PDL");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid size for a grid-stride kernel: enough blocks for the work, capped at 8 resident CTAs per SM
inline int grid_cap(int64_t blocks) {
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// fp32 -> (hi, lo) bf16 pair with hi + lo ~= v to 16 significand bits (bf16x3 operand split)
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace regen
