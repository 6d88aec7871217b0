// ST-GCN feature extractor: per-output-element arithmetic and the packed-weight walk, shared by the CUDA kernels
// (stgcn.cu) and the host-side check of the same code (tests/stgcn_hostcheck.cpp, compiled with g++ and run on the CPU
// against the oracle: the index arithmetic of every kernel is exercised without a GPU).  Plain C++ when not compiled by nvcc.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/regen_sm100.h"

#ifdef __CUDACC__
#define STGCN_HD __host__ __device__ __forceinline__
#else
#define STGCN_HD inline
#endif

namespace regen {
namespace stgcn {

constexpr int kBlocks = 10;
constexpr int kCinTab[kBlocks] = {0, 64, 64, 64, 64, 128, 128, 128, 256, 256};   // [0] = in_channels / num_person
constexpr int kCoutTab[kBlocks] = {64, 64, 64, 64, 128, 128, 128, 256, 256, 256};
constexpr int kStrideTab[kBlocks] = {1, 1, 1, 1, 2, 1, 1, 2, 1, 1};
constexpr float kBnEps = 1e-5f;

struct Bn {  // pointers into the packed weight buffer, [C] each
  const float *w, *b, *mean, *var;
};

struct BlockW {
  const float *gcn_w, *gcn_b, *t_w, *t_b, *res_w, *res_b, *imp;
  Bn bn0, bn3, bnr;
  bool res_conv;
};

struct Weights {
  const float* A;
  Bn data_bn;
  BlockW blk[kBlocks];
  const float *fc_w, *fc_b;
};

inline int block_cin(const regen_stgcn_desc& d, int i) { return i == 0 ? d.in_channels / d.num_person : kCinTab[i]; }

inline bool desc_ok(const regen_stgcn_desc* d) {
  return d && (d->num_person == 1 || d->num_person == 2) && d->in_channels >= d->num_person &&
         d->in_channels % d->num_person == 0 && d->num_class >= 1 && d->num_node >= 1 && d->num_node <= 1024 &&
         d->num_part >= 1 && d->num_part <= 8;
}

// Walk the packed layout documented in include/regen_sm100.h: returns its size in floats; fills `w` (pointers into
// `base`) when w != nullptr.
inline int64_t walk(const regen_stgcn_desc& d, Weights* w, const float* base) {
  int64_t off = 0;
  auto take = [&](int64_t n) {
    const float* p = base ? base + off : nullptr;
    off += n;
    return p;
  };
  auto take_bn = [&](int c) {
    Bn b;
    b.w = take(c); b.b = take(c); b.mean = take(c); b.var = take(c);
    return b;
  };
  const int V = d.num_node, K = d.num_part;
  const float* A = take((int64_t)K * V * V);
  Bn dbn = take_bn(d.in_channels * V);
  if (w) { w->A = A; w->data_bn = dbn; }
  for (int i = 0; i < kBlocks; ++i) {
    const int cin = block_cin(d, i), cout = kCoutTab[i];
    BlockW b;
    memset(&b, 0, sizeof(b));
    b.gcn_w = take((int64_t)K * cout * cin);
    b.gcn_b = take((int64_t)K * cout);
    b.bn0 = take_bn(cout);
    b.t_w = take((int64_t)cout * cout * 9);
    b.t_b = take(cout);
    b.bn3 = take_bn(cout);
    b.res_conv = i > 0 && !(cin == cout && kStrideTab[i] == 1);
    if (b.res_conv) {
      b.res_w = take((int64_t)cout * cin);
      b.res_b = take(cout);
      b.bnr = take_bn(cout);
    }
    b.imp = take((int64_t)K * V * V);
    if (w) w->blk[i] = b;
  }
  const float* fw = take((int64_t)d.num_class * 256);
  const float* fb = take(d.num_class);
  if (w) { w->fc_w = fw; w->fc_b = fb; }
  return off;
}

STGCN_HD float bn_apply(float x, const Bn& bn, int c) {
  return (x - bn.mean[c]) / sqrtf(bn.var[c] + kBnEps) * bn.w[c] + bn.b[c];
}

// output [N, V, C*P, T] -> data_bn -> x [NM, C, T, V] element i      (stgcn.py:81-103)
// data_bn channel of (m, v, c): (m*V + v)*C + c  (x.view(N, M*V*C, T) for P == 2; v*C + c for P == 1, m = 0)
STGCN_HD float prep_elem(int64_t i, const float* out_in, const Bn& bn, int n0, int V, int C, int P, int T) {
  const int v = (int)(i % V);
  const int t = (int)((i / V) % T);
  const int c = (int)((i / ((int64_t)V * T)) % C);
  const int nm = (int)(i / ((int64_t)V * T * C));
  const int n = n0 + nm / P, m = nm % P;
  const float val = out_in[(((int64_t)n * V + v) * (C * P) + m * C + c) * T + t];
  return bn_apply(val, bn, (m * V + v) * C + c);
}

// y[n, co, t', v] = b[co] + sum_ci W[co, ci] * x[n, ci, t' * stride, v]   (+ optional BatchNorm: the residual branch)
STGCN_HD float conv1x1_elem(int64_t i, const float* x, const float* W, const float* b, const Bn& bn, int use_bn, int Cin,
                            int Cout, int T, int Tout, int V, int stride) {
  const int v = (int)(i % V);
  const int t = (int)((i / V) % Tout);
  const int co = (int)((i / ((int64_t)V * Tout)) % Cout);
  const int n = (int)(i / ((int64_t)V * Tout * Cout));
  const float* xp = x + ((int64_t)n * Cin * T + (int64_t)t * stride) * V + v;
  const float* wp = W + (int64_t)co * Cin;
  float acc = 0.f;
  for (int ci = 0; ci < Cin; ++ci) acc += wp[ci] * xp[(int64_t)ci * T * V];
  acc += b[co];
  return use_bn ? bn_apply(acc, bn, co) : acc;
}

// h[n, co, t, w] = relu(bn(sum_k sum_v y[n, k*Cout + co, t, v] * A[k, v, w]))       (tgcn.py:60-62, tcn.0, tcn.1)
STGCN_HD float graph_elem(int64_t i, const float* y, const float* A, const Bn& bn, int K, int Cout, int T, int V) {
  const int w = (int)(i % V);
  const int t = (int)((i / V) % T);
  const int co = (int)((i / ((int64_t)V * T)) % Cout);
  const int n = (int)(i / ((int64_t)V * T * Cout));
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float* yp = y + (((int64_t)n * K * Cout + (int64_t)k * Cout + co) * T + t) * V;
    const float* ap = A + (int64_t)k * V * V + w;
    for (int v = 0; v < V; ++v) acc += yp[v] * ap[(int64_t)v * V];
  }
  return fmaxf(bn_apply(acc, bn, co), 0.f);
}

// out[n, co, t', v] = relu(bn(b[co] + sum_k sum_ci W[co, ci, k] * h[n, ci, t'*stride + k - 4, v]) + res[n, co, t', v])
// (tcn.2 .. tcn.4, residual, relu: stgcn.py:206-211); res == nullptr for the first block
STGCN_HD float tconv_elem(int64_t i, const float* h, const float* W, const float* b, const float* res, const Bn& bn, int C,
                          int T, int Tout, int V, int stride) {
  const int v = (int)(i % V);
  const int t = (int)((i / V) % Tout);
  const int co = (int)((i / ((int64_t)V * Tout)) % C);
  const int n = (int)(i / ((int64_t)V * Tout * C));
  const float* hp = h + (int64_t)n * C * T * V + v;
  const float* wp = W + (int64_t)co * C * 9;
  float acc = 0.f;
  for (int k = 0; k < 9; ++k) {  // tap-outer, like the oracle's explicit tap sum
    const int ts = t * stride + k - 4;
    if (ts < 0 || ts >= T) continue;
    float part = 0.f;
    for (int ci = 0; ci < C; ++ci) part += wp[ci * 9 + k] * hp[((int64_t)ci * T + ts) * V];
    acc += part;
  }
  float r = bn_apply(acc + b[co], bn, co);
  if (res) r += res[i];
  return fmaxf(r, 0.f);
}

// feat[n, c] = mean over persons of the mean over (t, v) of x[(n, m), c, t, v]     (stgcn.py:113-117); i = n * C + c
STGCN_HD float pool_elem(int64_t i, const float* x, int P, int C, int TV) {
  const int n = (int)(i / C), c = (int)(i % C);
  float total = 0.f;
  for (int m = 0; m < P; ++m) {
    const float* xp = x + ((int64_t)(n * P + m) * C + c) * TV;
    float s = 0.f;
    for (int j = 0; j < TV; ++j) s += xp[j];
    total += s / (float)TV;
  }
  return total / (float)P;
}

// yhat[n, j] = bf[j] + sum_c feat[n, c] * Wf[j, c]      (fcn, a 1x1 convolution on a 1x1 map); i = n * NC + j
STGCN_HD float fc_elem(int64_t i, const float* feat, const float* Wf, const float* bf, int C, int NC) {
  const int n = (int)(i / NC), j = (int)(i % NC);
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc += feat[(int64_t)n * C + c] * Wf[(int64_t)j * C + c];
  return acc + bf[j];
}

// Block schedule shared by the device driver and the host check: calls f(i, cin, cout, stride, T_in, T_out) per block.
template <class F>
inline int for_each_block(const regen_stgcn_desc& d, int T, F&& f) {
  int Tc = T;
  for (int i = 0; i < kBlocks; ++i) {
    const int st = kStrideTab[i];
    const int Tout = (Tc + 8 - 9) / st + 1;
    f(i, block_cin(d, i), kCoutTab[i], st, Tc, Tout);
    Tc = Tout;
  }
  return Tc;
}

}  // namespace stgcn
}  // namespace regen
