// Evaluation feature extractor (ST-GCN) behind the C ABI -- SURVEY.md 8f row 3.
//
// Reference: eval/a2m/recognition/models/stgcn.py:76-126 (STGCN.forward), :145-213 (st_gcn block) and
// eval/a2m/recognition/models/stgcnutils/tgcn.py:55-64 (1x1 conv to K*C_out channels + einsum with the adjacency
// partitions), inference mode (BatchNorm running statistics, dropout = identity).  The whole extractor is ~3 GFLOP per
// (sample, person) -- five orders of magnitude below one sampling loop -- so it is written as plain fp32 CUDA-core
// kernels (one thread per output element, the innermost tensor axis v across the lanes so that activation reads
// coalesce and weight reads broadcast); it is NOT on a roofline.  Samples are processed in chunks so that the
// intermediate of the graph convolution (K * C_out channels at the input length) stays bounded.
//
// STATUS: written against the pinned oracle (oracle/stgcn_ref.py) in the round whose GPU budget was already spent; its
// GPU parity tests (tests/test_gpu_stgcn.py) are marked xfail(strict=False) until they have been seen green on a B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/regen_sm100.h"
#include "common.cuh"
#include "stgcn_elems.cuh"

using namespace regen;
using namespace regen::stgcn;

namespace {

constexpr int kThreadsS = 256;
constexpr int kChunk = 64;   // (sample, person) rows per pass

#define STGCN_GRID_STRIDE(total)                                                                       \
  for (int64_t i = (int64_t)blockIdx.x * kThreadsS + threadIdx.x; i < (total); i += (int64_t)gridDim.x * kThreadsS)

__global__ void __launch_bounds__(kThreadsS) prep_kernel(const float* __restrict__ out_in, float* __restrict__ x, Bn bn,
                                                        int n0, int NM, int V, int C, int P, int T) {
  STGCN_GRID_STRIDE((int64_t)NM * C * T * V) x[i] = prep_elem(i, out_in, bn, n0, V, C, P, T);
}
__global__ void __launch_bounds__(kThreadsS) conv1x1_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ b, float* __restrict__ y, Bn bn,
                                                           int use_bn, int NM, int Cin, int Cout, int T, int Tout,
                                                           int V, int stride) {
  STGCN_GRID_STRIDE((int64_t)NM * Cout * Tout * V) y[i] = conv1x1_elem(i, x, W, b, bn, use_bn, Cin, Cout, T, Tout, V, stride);
}
__global__ void __launch_bounds__(kThreadsS) graph_bn_relu_kernel(const float* __restrict__ y, const float* __restrict__ A,
                                                                 float* __restrict__ h, Bn bn, int NM, int K, int Cout,
                                                                 int T, int V) {
  STGCN_GRID_STRIDE((int64_t)NM * Cout * T * V) h[i] = graph_elem(i, y, A, bn, K, Cout, T, V);
}
__global__ void __launch_bounds__(kThreadsS) tconv_bn_res_relu_kernel(const float* __restrict__ h, const float* __restrict__ W,
                                                                     const float* __restrict__ b, const float* res,
                                                                     float* out, Bn bn, int NM, int C, int T, int Tout,
                                                                     int V, int stride) {
  STGCN_GRID_STRIDE((int64_t)NM * C * Tout * V) out[i] = tconv_elem(i, h, W, b, res, bn, C, T, Tout, V, stride);
}
// feat rows n0 .. n0 + Nc of the chunk
__global__ void __launch_bounds__(kThreadsS) pool_kernel(const float* __restrict__ x, float* __restrict__ feat, int n0,
                                                        int Nc, int P, int C, int TV) {
  STGCN_GRID_STRIDE((int64_t)Nc * C) feat[(int64_t)n0 * C + i] = pool_elem(i, x, P, C, TV);
}
__global__ void __launch_bounds__(kThreadsS) fc_kernel(const float* __restrict__ feat, const float* __restrict__ Wf,
                                                      const float* __restrict__ bf, float* __restrict__ yhat, int N, int C,
                                                      int NC) {
  STGCN_GRID_STRIDE((int64_t)N * NC) yhat[i] = fc_elem(i, feat, Wf, bf, C, NC);
}
// Aeff[i] = A * edge_importance_i      (stgcn.py:106-107), once per load
__global__ void __launch_bounds__(kThreadsS) mul_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       float* __restrict__ o, int n) {
  STGCN_GRID_STRIDE(n) o[i] = a[i] * b[i];
}

inline int grid_of(int64_t total) { return grid_cap(ceil_div(total, kThreadsS)); }

}  // namespace

struct regen_stgcn {
  regen_stgcn_desc d;
  int device = 0;
  float* packed = nullptr;   // device copy of the packed weights
  float* aeff = nullptr;     // [10][K, V, V]
  Weights w;
  bool loaded = false;
  // workspaces for one chunk at length T_ws
  float *x0 = nullptr, *x1 = nullptr, *y = nullptr, *hbuf = nullptr, *res = nullptr;
  int T_ws = 0;
};

namespace {

void free_ws(regen_stgcn* h) {
  cudaFree(h->x0); cudaFree(h->x1); cudaFree(h->y); cudaFree(h->hbuf); cudaFree(h->res);
  h->x0 = h->x1 = h->y = h->hbuf = h->res = nullptr;
  h->T_ws = 0;
}

}  // namespace

extern "C" {

int64_t regen_stgcn_packed_size(const regen_stgcn_desc* d) {
  if (!desc_ok(d)) return -1;
  return walk(*d, nullptr, nullptr);
}

int regen_stgcn_create(regen_stgcn** out, int32_t device, const regen_stgcn_desc* d) {
  REGEN_CHECK_ARG(out, "regen_stgcn_create: null handle pointer");
  REGEN_CHECK_ARG(desc_ok(d), "regen_stgcn_create: bad descriptor");
  regen_stgcn* h = new regen_stgcn();
  h->d = *d;
  h->device = device;
  memset(&h->w, 0, sizeof(h->w));
  *out = h;
  return REGEN_OK;
}

void regen_stgcn_destroy(regen_stgcn* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  free_ws(h);
  cudaFree(h->packed);
  cudaFree(h->aeff);
  delete h;
}

int regen_stgcn_load_weights(regen_stgcn* h, const float* packed, int64_t n_floats, void* stream) {
  REGEN_CHECK_ARG(h && packed, "regen_stgcn_load_weights: null argument");
  DeviceGuard guard(h->device);
  const int64_t need = walk(h->d, nullptr, nullptr);
  REGEN_CHECK_ARG(n_floats == need, "regen_stgcn_load_weights: packed buffer has %lld floats, the descriptor needs %lld",
                  (long long)n_floats, (long long)need);
  cudaStream_t s = (cudaStream_t)stream;
  REGEN_CUDA(cudaSetDevice(h->device));
  const int kvv = h->d.num_part * h->d.num_node * h->d.num_node;
  if (!h->packed) {
    REGEN_CUDA(cudaMalloc(&h->packed, need * sizeof(float)));
    REGEN_CUDA(cudaMalloc(&h->aeff, (size_t)kBlocks * kvv * sizeof(float)));
  }
  REGEN_CUDA(cudaMemcpyAsync(h->packed, packed, need * sizeof(float), cudaMemcpyDeviceToDevice, s));
  walk(h->d, &h->w, h->packed);
  for (int i = 0; i < kBlocks; ++i) {
    mul_kernel<<<grid_of(kvv), kThreadsS, 0, s>>>(h->w.A, h->w.blk[i].imp, h->aeff + (size_t)i * kvv, kvv);
    REGEN_LAUNCH_CHECK();
    count_launch();
  }
  h->loaded = true;
  return REGEN_OK;
}

int regen_stgcn_forward(regen_stgcn* h, const float* output, int32_t N, int32_t T, float* features, float* yhat,
                        void* stream) {
  REGEN_CHECK_ARG(h && output && features && yhat, "regen_stgcn_forward: null argument");
  REGEN_CHECK_ARG(N >= 0 && T >= 1, "regen_stgcn_forward: bad sizes N=%d T=%d", N, T);
  if (!h->loaded) {
    set_error("regen_stgcn_forward: regen_stgcn_load_weights has not been called");
    return REGEN_ESTATE;
  }
  if (N == 0) return REGEN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceGuard guard(h->device);
  const int P = h->d.num_person, V = h->d.num_node, K = h->d.num_part, C = h->d.in_channels / P;
  if (h->T_ws < T) {
    REGEN_CUDA(cudaStreamSynchronize(s));
    free_ws(h);
    // channels x length of the blocks: 64 x T, 128 x ceil(T/2), 256 x ceil(T/4) <= 64 (T + 4); the graph convolution of
    // the two stride-2 blocks runs at the INPUT length with the output channels: 128 x T, 256 x ceil(T/2) <= 128 (T + 2)
    const size_t act = (size_t)kChunk * 64 * (T + 4) * V;
    const size_t act0 = act > (size_t)kChunk * C * T * V ? act : (size_t)kChunk * C * T * V;
    REGEN_CUDA(cudaMalloc(&h->x0, act0 * sizeof(float)));
    REGEN_CUDA(cudaMalloc(&h->x1, act0 * sizeof(float)));
    REGEN_CUDA(cudaMalloc(&h->hbuf, (size_t)kChunk * 128 * (T + 2) * V * sizeof(float)));
    REGEN_CUDA(cudaMalloc(&h->res, act * sizeof(float)));
    REGEN_CUDA(cudaMalloc(&h->y, (size_t)kChunk * K * 128 * (T + 2) * V * sizeof(float)));
    h->T_ws = T;
  }
  const int samples_per_chunk = kChunk / P;
  const int kvv = K * V * V;
  for (int n0 = 0; n0 < N; n0 += samples_per_chunk) {
    const int Nc = N - n0 < samples_per_chunk ? N - n0 : samples_per_chunk;
    const int NM = Nc * P;
    prep_kernel<<<grid_of((int64_t)NM * C * T * V), kThreadsS, 0, s>>>(output, h->x0, h->w.data_bn, n0, NM, V, C, P, T);
    REGEN_LAUNCH_CHECK();
    count_launch();
    float *cur = h->x0, *nxt = h->x1;
    cudaError_t err = cudaSuccess;
    const int T_last = for_each_block(h->d, T, [&](int i, int cin, int cout, int st, int Tc, int Tout) {
      const BlockW& b = h->w.blk[i];
      const float* resp = nullptr;
      if (i > 0) {
        if (b.res_conv) {
          conv1x1_kernel<<<grid_of((int64_t)NM * cout * Tout * V), kThreadsS, 0, s>>>(cur, b.res_w, b.res_b, h->res, b.bnr, 1,
                                                                                     NM, cin, cout, Tc, Tout, V, st);
          count_launch();
          resp = h->res;
        } else {
          resp = cur;
        }
      }
      conv1x1_kernel<<<grid_of((int64_t)NM * K * cout * Tc * V), kThreadsS, 0, s>>>(cur, b.gcn_w, b.gcn_b, h->y, b.bn0, 0, NM,
                                                                                   cin, K * cout, Tc, Tc, V, 1);
      graph_bn_relu_kernel<<<grid_of((int64_t)NM * cout * Tc * V), kThreadsS, 0, s>>>(h->y, h->aeff + (size_t)i * kvv, h->hbuf,
                                                                                     b.bn0, NM, K, cout, Tc, V);
      tconv_bn_res_relu_kernel<<<grid_of((int64_t)NM * cout * Tout * V), kThreadsS, 0, s>>>(h->hbuf, b.t_w, b.t_b, resp, nxt,
                                                                                           b.bn3, NM, cout, Tc, Tout, V, st);
      count_launch(3);
      if (err == cudaSuccess) err = cudaGetLastError();
      float* tmp = cur; cur = nxt; nxt = tmp;
    });
    if (err != cudaSuccess) {
      set_error("regen_stgcn_forward: kernel launch failed: %s", cudaGetErrorString(err));
      return REGEN_ECUDA;
    }
    pool_kernel<<<grid_of((int64_t)Nc * 256), kThreadsS, 0, s>>>(cur, features, n0, Nc, P, 256, T_last * V);
    REGEN_LAUNCH_CHECK();
    count_launch();
  }
  fc_kernel<<<grid_of((int64_t)N * h->d.num_class), kThreadsS, 0, s>>>(features, h->w.fc_w, h->w.fc_b, yhat, N, 256,
                                                                      h->d.num_class);
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

}  // extern "C"
