// Test infrastructure (CPU): runs the ST-GCN kernels' per-element functions (regennet_b200/csrc/stgcn_elems.cuh -- the very
// code the CUDA kernels of stgcn.cu execute per thread), the packed-weight walk and the block schedule on the host, with the
// same chunking and workspace sizing as regen_stgcn_forward, so that tests/test_stgcn_hostcheck.py can compare them with the
// oracle without a GPU.  Built by the test with g++; never part of the product library.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../regennet_b200/csrc/stgcn_elems.cuh"

using namespace regen::stgcn;

#define NEED(cond)                                                          \
  do {                                                                      \
    if (!(cond)) {                                                          \
      fprintf(stderr, "stgcn_hostcheck: %s failed (line %d)\n", #cond, __LINE__); \
      return -100;                                                          \
    }                                                                       \
  } while (0)

extern "C" long long stgcn_host_packed_size(const regen_stgcn_desc* d) { return desc_ok(d) ? walk(*d, nullptr, nullptr) : -1; }

extern "C" int stgcn_host_forward(const regen_stgcn_desc* d, const float* packed, long long n_packed, const float* output,
                                  int N, int T, float* features, float* yhat) {
  NEED(desc_ok(d));
  NEED(walk(*d, nullptr, nullptr) == n_packed);
  Weights w;
  walk(*d, &w, packed);
  const int kChunk = 64;
  const int P = d->num_person, V = d->num_node, K = d->num_part, C = d->in_channels / P;
  const size_t kvv = (size_t)K * V * V;
  std::vector<float> aeff((size_t)kBlocks * kvv);
  for (int i = 0; i < kBlocks; ++i)
    for (size_t j = 0; j < kvv; ++j) aeff[i * kvv + j] = w.A[j] * w.blk[i].imp[j];
  // workspace sizes exactly as regen_stgcn_forward allocates them
  const size_t act = (size_t)kChunk * 64 * (T + 4) * V;
  const size_t act0 = act > (size_t)kChunk * C * T * V ? act : (size_t)kChunk * C * T * V;
  const size_t n_h = (size_t)kChunk * 128 * (T + 2) * V, n_y = (size_t)kChunk * K * 128 * (T + 2) * V;
  std::vector<float> x0(act0), x1(act0), hbuf(n_h), res(act), y(n_y);
  const int samples_per_chunk = kChunk / P;
  int rc = 0;
  for (int n0 = 0; n0 < N; n0 += samples_per_chunk) {
    const int Nc = N - n0 < samples_per_chunk ? N - n0 : samples_per_chunk;
    const int NM = Nc * P;
    NEED((size_t)NM * C * T * V <= act0);
    for (int64_t i = 0; i < (int64_t)NM * C * T * V; ++i) x0[i] = prep_elem(i, output, w.data_bn, n0, V, C, P, T);
    float *cur = x0.data(), *nxt = x1.data();
    const int T_last = for_each_block(*d, T, [&](int i, int cin, int cout, int st, int Tc, int Tout) {
      const BlockW& b = w.blk[i];
      const float* resp = nullptr;
      if ((size_t)NM * cout * Tout * V > act0 || (size_t)NM * cout * Tout * V > act || (size_t)NM * cout * Tc * V > n_h ||
          (size_t)NM * K * cout * Tc * V > n_y) {
        rc = -200 - i;   // a workspace would overflow
        return;
      }
      if (i > 0) {
        if (b.res_conv) {
          for (int64_t e = 0; e < (int64_t)NM * cout * Tout * V; ++e)
            res[e] = conv1x1_elem(e, cur, b.res_w, b.res_b, b.bnr, 1, cin, cout, Tc, Tout, V, st);
          resp = res.data();
        } else {
          resp = cur;
        }
      }
      for (int64_t e = 0; e < (int64_t)NM * K * cout * Tc * V; ++e)
        y[e] = conv1x1_elem(e, cur, b.gcn_w, b.gcn_b, b.bn0, 0, cin, K * cout, Tc, Tc, V, 1);
      for (int64_t e = 0; e < (int64_t)NM * cout * Tc * V; ++e)
        hbuf[e] = graph_elem(e, y.data(), aeff.data() + (size_t)i * kvv, b.bn0, K, cout, Tc, V);
      for (int64_t e = 0; e < (int64_t)NM * cout * Tout * V; ++e)
        nxt[e] = tconv_elem(e, hbuf.data(), b.t_w, b.t_b, resp, b.bn3, cout, Tc, Tout, V, st);
      float* tmp = cur; cur = nxt; nxt = tmp;
    });
    if (rc) return rc;
    for (int64_t e = 0; e < (int64_t)Nc * 256; ++e) features[(int64_t)n0 * 256 + e] = pool_elem(e, cur, P, 256, T_last * V);
  }
  for (int64_t e = 0; e < (int64_t)N * d->num_class; ++e) yhat[e] = fc_elem(e, features, w.fc_w, w.fc_b, 256, d->num_class);
  return 0;
}
