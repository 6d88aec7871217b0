// Evaluation feature extractor (ST-GCN) behind the C ABI -- SURVEY.md 8f row 3.
//
// Reference: eval/a2m/recognition/models/stgcn.py:76-126 (STGCN.forward), :145-213 (st_gcn block) and
// eval/a2m/recognition/models/stgcnutils/tgcn.py:55-64 (1x1 conv to K*C_out channels + einsum with the adjacency
// partitions), inference mode (BatchNorm running statistics, dropout = identity).  The whole extractor is ~3 GFLOP per
// (sample, person) -- five orders of magnitude below one sampling loop -- but eval_cmdm pushes 20 x 2 x 1000 samples
// through it, so the three convolution-shaped stages run as shared-memory-tiled fp32 CUDA-core GEMMs (exact fp32
// arithmetic: the extractor feeds FID / accuracy numbers that are compared across papers, so no reduced-precision
// tensor-core path): 64 x 64 output tiles, 16-deep k slabs, 4 x 4 register micro-tiles per thread:
//   * 1x1 convolutions (graph-conv projection to K * C_out channels, residual projection) and the 9-tap temporal
//     convolution are ONE kernel, conv_tiled_kernel<TAPS>: per sample Y[C_out x (T' V)] = sum_tap W_tap[C_out x C_in] .
//     X_tap[C_in x (T' V)] with the tap / stride shift folded into the column addressing (temporal weights are re-packed
//     to [tap][C_out][C_in] at load time);
//   * the adjacency einsum nkctv,kvw->nctw is graph_tiled_kernel: per sample H[(c, t) x w] = sum_k Y_k[(c, t) x v] . A_k[v x w]
//     where Y_k is a CONTIGUOUS [(c, t) x V] slab of the projection output;
// BatchNorm / bias / residual / ReLU are fused into the tile epilogues.  Graphs with more than 64 nodes fall back to the
// one-thread-per-output kernels (the per-element arithmetic in stgcn_elems.cuh, which the host check also exercises).
// Samples are processed in chunks so that the intermediate of the graph convolution (K * C_out channels at the input
// length) stays bounded.
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


// ---------------------------------------------------------------------------------------------- tiled fp32 kernels
constexpr int kTM = 64, kTN = 64, kTK = 16;   // output tile and k-slab of the tiled kernels (256 threads, 4 x 4 each)

struct ConvEpi {      // epilogue of conv_tiled_kernel: out = relu?( bn?(acc + bias) + res? )
  const float* bias;
  const float* res;   // same layout as out, or null
  Bn bn;
  int use_bn, relu;
};

// out[n, co, t', v] = epi( sum_tap sum_ci W[tap][co][ci] * x[n, ci, t' * stride + tap - PAD, v] ),  PAD = 4 for 9 taps, 0 for 1
template <int TAPS>
__global__ void __launch_bounds__(256) conv_tiled_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                         float* __restrict__ out, ConvEpi epi, int Cin, int Cout, int T,
                                                         int Tout, int V, int stride) {
  __shared__ __align__(16) float Ws[kTK][kTM + 4];
  __shared__ __align__(16) float Xs[kTK][kTN + 4];
  constexpr int PAD = TAPS == 9 ? 4 : 0;
  const int n = blockIdx.z, co0 = blockIdx.y * kTM, col0 = blockIdx.x * kTN;
  const int Nc = Tout * V;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // loader roles: X -- column lc of the tile, k rows lk, lk + 4, lk + 8, lk + 12;  W -- output channel wm, 4 consecutive ci
  const int lc = tid & 63, lk = tid >> 6;
  const int wm = tid >> 2, wk = (tid & 3) * 4;
  const int col = col0 + lc;
  const int tq = col < Nc ? col / V : 0, v = col < Nc ? col - tq * V : 0;
  const float* xn = x + (int64_t)n * Cin * T * V;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int tap = 0; tap < TAPS; ++tap) {
    const int ts = tq * stride + tap - PAD;
    const bool colok = col < Nc && ts >= 0 && ts < T;
    const float* xcol = xn + (int64_t)ts * V + v;
    const float* wt = W + (int64_t)tap * Cout * Cin;
    for (int ci0 = 0; ci0 < Cin; ci0 += kTK) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int ci = ci0 + lk + 4 * r;
        Xs[lk + 4 * r][lc] = (colok && ci < Cin) ? __ldg(xcol + (int64_t)ci * T * V) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int ci = ci0 + wk + r, co = co0 + wm;
        Ws[wk + r][wm] = (co < Cout && ci < Cin) ? __ldg(wt + (int64_t)co * Cin + ci) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kTK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&Ws[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c >= Nc) continue;
      const int64_t idx = ((int64_t)n * Cout + co) * Nc + c;
      float r = acc[i][j] + epi.bias[co];
      if (epi.use_bn) r = bn_apply(r, epi.bn, co);
      if (epi.res) r += epi.res[idx];
      out[idx] = epi.relu ? fmaxf(r, 0.f) : r;
    }
  }
}

// h[n, co, t, w] = relu(bn(sum_k sum_v y[n, k * Cout + co, t, v] * A[k, v, w])),  V <= 64.  Rows r = (co, t): for a fixed k the
// rows [r0, r0 + 64) of the projection output are one contiguous slab of 64 * V floats.
__global__ void __launch_bounds__(256) graph_tiled_kernel(const float* __restrict__ y, const float* __restrict__ A,
                                                          float* __restrict__ h, Bn bn, int K, int Cout, int T, int V) {
  __shared__ __align__(16) float Yt[kTN][kTM + 4];   // [v][row]
  __shared__ __align__(16) float As[kTN][kTN + 4];   // [v][w], zero-padded to 64 columns
  const int n = blockIdx.y, r0 = blockIdx.x * kTM;
  const int R = Cout * T;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int rows = R - r0 < kTM ? R - r0 : kTM;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float* slab = y + (((int64_t)n * K + k) * R + r0) * V;
    for (int e = tid; e < kTM * V; e += 256) {
      const int r = e / V, v = e - r * V;
      Yt[v][r] = r < rows ? __ldg(slab + e) : 0.f;
    }
    const float* ak = A + (int64_t)k * V * V;
    for (int e = tid; e < V * kTN; e += 256) {
      const int v = e >> 6, w = e & 63;
      As[v][w] = w < V ? __ldg(ak + v * V + w) : 0.f;
    }
    __syncthreads();
    for (int v = 0; v < V; ++v) {
      const float4 a = *reinterpret_cast<const float4*>(&Yt[v][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&As[v][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= R) continue;
    const int co = r / T;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int w = tx * 4 + j;
      if (w < V) h[((int64_t)n * R + r) * V + w] = fmaxf(bn_apply(acc[i][j], bn, co), 0.f);
    }
  }
}

// temporal-convolution weights [co][ci][9] -> [tap][co][ci] (once per load)
__global__ void __launch_bounds__(kThreadsS) repack_tconv_kernel(const float* __restrict__ w, float* __restrict__ o, int C) {
  STGCN_GRID_STRIDE((int64_t)9 * C * C) {
    const int ci = (int)(i % C), co = (int)((i / C) % C), tap = (int)(i / ((int64_t)C * C));
    o[i] = w[((int64_t)co * C + ci) * 9 + tap];
  }
}

// feat[n, c] = mean over persons of the mean over (t, v): one warp per (n, c)
__global__ void __launch_bounds__(kThreadsS) pool_warp_kernel(const float* __restrict__ x, float* __restrict__ feat, int n0,
                                                              int Nc, int P, int C, int TV) {
  const int warp = (blockIdx.x * kThreadsS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Nc * C) return;
  const int n = warp / C, c = warp - n * C;
  float total = 0.f;
  for (int m = 0; m < P; ++m) {
    const float* xp = x + ((int64_t)(n * P + m) * C + c) * TV;
    float s = 0.f;
    for (int j = lane; j < TV; j += 32) s += xp[j];
    total += warp_sum(s) / (float)TV;
  }
  if (lane == 0) feat[(int64_t)n0 * C + warp] = total / (float)P;
}

inline int grid_of(int64_t total) { return grid_cap(ceil_div(total, kThreadsS)); }

}  // namespace

struct regen_stgcn {
  regen_stgcn_desc d;
  int device = 0;
  float* packed = nullptr;   // device copy of the packed weights
  float* aeff = nullptr;     // [10][K, V, V]
  float* twp = nullptr;      // temporal-convolution weights of the ten blocks re-packed to [tap][C][C]
  size_t twp_off[kBlocks] = {0};
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
  cudaFree(h->twp);
  delete h;
}

int regen_stgcn_load_weights(regen_stgcn* h, const float* packed, int64_t n_floats, void* stream) {
  REGEN_CHECK_ARG(h && packed, "regen_stgcn_load_weights: null argument");
  DeviceGuard guard(h->device);
  const int64_t need = walk(h->d, nullptr, nullptr);
  REGEN_CHECK_ARG(n_floats == need, "regen_stgcn_load_weights: packed buffer has %lld floats, the descriptor needs %lld",
                  (long long)n_floats, (long long)need);
  cudaStream_t s = (cudaStream_t)stream;
  const int kvv = h->d.num_part * h->d.num_node * h->d.num_node;
  size_t twp_total = 0;
  for (int i = 0; i < kBlocks; ++i) {
    h->twp_off[i] = twp_total;
    twp_total += (size_t)9 * kCoutTab[i] * kCoutTab[i];
  }
  if (!h->packed) {
    REGEN_CUDA(cudaMalloc(&h->packed, need * sizeof(float)));
    REGEN_CUDA(cudaMalloc(&h->aeff, (size_t)kBlocks * kvv * sizeof(float)));
    REGEN_CUDA(cudaMalloc(&h->twp, twp_total * sizeof(float)));
  }
  REGEN_CUDA(cudaMemcpyAsync(h->packed, packed, need * sizeof(float), cudaMemcpyDeviceToDevice, s));
  walk(h->d, &h->w, h->packed);
  for (int i = 0; i < kBlocks; ++i) {
    mul_kernel<<<grid_of(kvv), kThreadsS, 0, s>>>(h->w.A, h->w.blk[i].imp, h->aeff + (size_t)i * kvv, kvv);
    const int c = kCoutTab[i];
    repack_tconv_kernel<<<grid_of((int64_t)9 * c * c), kThreadsS, 0, s>>>(h->w.blk[i].t_w, h->twp + h->twp_off[i], c);
    REGEN_LAUNCH_CHECK();
    count_launch(2);
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
    const bool tiled = V <= kTN;   // the graph kernel keeps a [V x 64] adjacency tile in shared memory
    const int T_last = for_each_block(h->d, T, [&](int i, int cin, int cout, int st, int Tc, int Tout) {
      const BlockW& b = h->w.blk[i];
      const float* resp = nullptr;
      if (tiled) {
        auto conv1 = [&](const float* src, const float* w, const float* bias, float* dst, const Bn* bn, int ci, int co,
                         int Tin, int To, int stride) {
          ConvEpi e;
          e.bias = bias; e.res = nullptr; e.use_bn = bn ? 1 : 0; e.relu = 0;
          if (bn) e.bn = *bn; else memset(&e.bn, 0, sizeof(e.bn));
          dim3 grid((unsigned)ceil_div((int64_t)To * V, kTN), (unsigned)ceil_div(co, kTM), (unsigned)NM);
          conv_tiled_kernel<1><<<grid, 256, 0, s>>>(src, w, dst, e, ci, co, Tin, To, V, stride);
          count_launch();
        };
        if (i > 0) {
          if (b.res_conv) {
            conv1(cur, b.res_w, b.res_b, h->res, &b.bnr, cin, cout, Tc, Tout, st);
            resp = h->res;
          } else {
            resp = cur;
          }
        }
        conv1(cur, b.gcn_w, b.gcn_b, h->y, nullptr, cin, K * cout, Tc, Tc, 1);
        graph_tiled_kernel<<<dim3((unsigned)ceil_div((int64_t)cout * Tc, kTM), (unsigned)NM), 256, 0, s>>>(
            h->y, h->aeff + (size_t)i * kvv, h->hbuf, b.bn0, K, cout, Tc, V);
        ConvEpi e;
        e.bias = b.t_b; e.res = resp; e.bn = b.bn3; e.use_bn = 1; e.relu = 1;
        dim3 grid((unsigned)ceil_div((int64_t)Tout * V, kTN), (unsigned)ceil_div(cout, kTM), (unsigned)NM);
        conv_tiled_kernel<9><<<grid, 256, 0, s>>>(h->hbuf, h->twp + h->twp_off[i], nxt, e, cout, cout, Tc, Tout, V, st);
        count_launch(2);
      } else {
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
      }
      if (err == cudaSuccess) err = cudaGetLastError();
      float* tmp = cur; cur = nxt; nxt = tmp;
    });
    if (err != cudaSuccess) {
      set_error("regen_stgcn_forward: kernel launch failed: %s", cudaGetErrorString(err));
      return REGEN_ECUDA;
    }
    pool_warp_kernel<<<(unsigned)ceil_div((int64_t)Nc * 256 * 32, kThreadsS), kThreadsS, 0, s>>>(cur, features, n0, Nc, P, 256,
                                                                                                  T_last * V);
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
