// Handle-free elementwise operators of the sampling hot path (HBM-bound).
//
// Every kernel here moves each byte exactly once: float4-vectorised, coalesced, grid sized as a
// multiple of the SM count.  Arithmetic mirrors the reference's fp32 operation order with explicit
// round-to-nearest intrinsics (no FMA contraction), so the update is bit-comparable with the
// eager PyTorch sequence it replaces.
#include "common.cuh"
#include "ptx.cuh"

namespace regen {

long long g_launches = 0;
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

constexpr int kThreads = 256;
constexpr int kBlocksPerSM = 8;

inline int grid_for(int64_t work_items) {
  int64_t blocks = ceil_div(work_items, kThreads);
  int64_t cap = (int64_t)kNumSMs * kBlocksPerSM;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float clampf(float v, bool clip) { return clip ? fminf(fmaxf(v, -1.f), 1.f) : v; }

struct PSampleCoef {
  float c1, c2, sig;
};

// diffusion/gaussian_diffusion.py:273-276, :549-559
__device__ __forceinline__ PSampleCoef load_psample(const int64_t* t, const float* coef1, const float* coef2,
                                                    const float* logvar, int b) {
  int64_t tb = t[b];
  PSampleCoef c;
  c.c1 = __ldg(coef1 + tb);
  c.c2 = __ldg(coef2 + tb);
  float nz = tb != 0 ? 1.f : 0.f;
  c.sig = __fmul_rn(nz, expf(__fmul_rn(0.5f, __ldg(logvar + tb))));
  return c;
}

__device__ __forceinline__ float psample_one(float x, float x0, float nz, const PSampleCoef& c) {
  float mean = __fadd_rn(__fmul_rn(c.c1, x0), __fmul_rn(c.c2, x));
  return __fadd_rn(mean, __fmul_rn(c.sig, nz));
}

// Plain vector loads / stores.  Evict-first streaming hints (ld.global.cs / st.global.cs, -DREGEN_UPD_STREAM_HINTS) were
// measured and are slower on B200: 16.0 us against 15.2 us per launch at config 2, cold L2.
#ifdef REGEN_UPD_STREAM_HINTS
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
#else
__device__ __forceinline__ float4 ld_stream(const float4* p) { return *p; }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { *p = v; }
#endif

constexpr int kUpdUnroll = 4;        // float4 items per thread per block iteration (12 x 16 B loads in flight per thread)
constexpr int kUpdCoefCap = 1024;    // per-sample coefficient table in shared memory up to this batch

template <bool VEC>
__global__ void __launch_bounds__(kThreads) p_sample_update_kernel(
    const float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ noise,
    float* __restrict__ out, float* __restrict__ pred, const int64_t* __restrict__ t,
    const float* __restrict__ coef1, const float* __restrict__ coef2, const float* __restrict__ logvar,
    uint32_t n_items, uint32_t inner_items, uint32_t B, int clip) {
  // one item = one float4 (VEC) or one float; inner_items = items per (sample, outer index)
  ptx::griddep_wait();    // PDL: launched while the producer of x0 / noise drains; its results are needed from here on
  ptx::griddep_launch();
  if (VEC) {
    // Per-sample coefficients once per block (the gather chain t[b] -> table -> exp is two dependent global loads; per
    // item it sat in front of every store), then kUpdUnroll independent float4 triples per thread and iteration.
    __shared__ float s_c1[kUpdCoefCap], s_c2[kUpdCoefCap], s_sig[kUpdCoefCap];
    const bool tab = B <= (uint32_t)kUpdCoefCap;
    if (tab) {
      for (uint32_t b = threadIdx.x; b < B; b += kThreads) {
        PSampleCoef c = load_psample(t, coef1, coef2, logvar, (int)b);
        s_c1[b] = c.c1; s_c2[b] = c.c2; s_sig[b] = c.sig;
      }
      __syncthreads();
    }
    // grid-stride order: at any moment the whole grid reads ONE contiguous window of each operand (DRAM-page friendly; a
    // per-block contiguous partition -- 1184 separate streams per operand -- measured 18.0 us against 13.4 us under ncu)
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t base = blockIdx.x * kThreads + threadIdx.x; base < n_items; base += stride * kUpdUnroll) {
      float4 vx[kUpdUnroll], v0[kUpdUnroll], vn[kUpdUnroll];
#pragma unroll
      for (int u = 0; u < kUpdUnroll; ++u) {
        const uint32_t i = base + u * stride;
        if (i < n_items) {
          vx[u] = ld_stream(reinterpret_cast<const float4*>(x) + i);
          v0[u] = ld_stream(reinterpret_cast<const float4*>(x0) + i);
          vn[u] = noise ? ld_stream(reinterpret_cast<const float4*>(noise) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < kUpdUnroll; ++u) {
        const uint32_t i = base + u * stride;
        if (i >= n_items) continue;
        const int b = (int)((i / inner_items) % B);
        PSampleCoef c;
        if (tab) { c.c1 = s_c1[b]; c.c2 = s_c2[b]; c.sig = s_sig[b]; }
        else c = load_psample(t, coef1, coef2, logvar, b);
        float4 p0 = v0[u];
        p0.x = clampf(p0.x, clip); p0.y = clampf(p0.y, clip); p0.z = clampf(p0.z, clip); p0.w = clampf(p0.w, clip);
        float4 o;
        o.x = psample_one(vx[u].x, p0.x, vn[u].x, c);
        o.y = psample_one(vx[u].y, p0.y, vn[u].y, c);
        o.z = psample_one(vx[u].z, p0.z, vn[u].z, c);
        o.w = psample_one(vx[u].w, p0.w, vn[u].w, c);
        st_stream(reinterpret_cast<float4*>(out) + i, o);
        if (pred) st_stream(reinterpret_cast<float4*>(pred) + i, p0);
      }
    }
  } else {
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_items; i += gridDim.x * kThreads) {
      int b = (int)((i / inner_items) % B);
      PSampleCoef c = load_psample(t, coef1, coef2, logvar, b);
      float v0 = clampf(x0[i], clip);
      out[i] = psample_one(x[i], v0, noise ? noise[i] : 0.f, c);
      if (pred) pred[i] = v0;
    }
  }
}

struct DdimCoef {
  float sra, srm1, sqrt_abp, dir, sig;
};

// diffusion/gaussian_diffusion.py:769-793 with :418-423
__device__ __forceinline__ DdimCoef load_ddim(const int64_t* t, const float* sra, const float* srm1,
                                              const float* ac, const float* acp, float eta, int b) {
  int64_t tb = t[b];
  DdimCoef c;
  c.sra = __ldg(sra + tb);
  c.srm1 = __ldg(srm1 + tb);
  float ab = __ldg(ac + tb), abp = __ldg(acp + tb);
  float sigma = __fmul_rn(__fmul_rn(eta, __fsqrt_rn(__fdiv_rn(__fsub_rn(1.f, abp), __fsub_rn(1.f, ab)))),
                          __fsqrt_rn(__fsub_rn(1.f, __fdiv_rn(ab, abp))));
  c.sqrt_abp = __fsqrt_rn(abp);
  c.dir = __fsqrt_rn(__fsub_rn(__fsub_rn(1.f, abp), __fmul_rn(sigma, sigma)));
  float nz = tb != 0 ? 1.f : 0.f;
  c.sig = __fmul_rn(nz, sigma);
  return c;
}

__device__ __forceinline__ float ddim_one(float x, float x0, float nz, const DdimCoef& c) {
  float eps = __fdiv_rn(__fsub_rn(__fmul_rn(c.sra, x), x0), c.srm1);
  float mean = __fadd_rn(__fmul_rn(x0, c.sqrt_abp), __fmul_rn(c.dir, eps));
  return __fadd_rn(mean, __fmul_rn(c.sig, nz));
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads) ddim_update_kernel(
    const float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ noise,
    float* __restrict__ out, float* __restrict__ pred, const int64_t* __restrict__ t,
    const float* __restrict__ sra, const float* __restrict__ srm1, const float* __restrict__ ac,
    const float* __restrict__ acp, float eta, uint32_t n_items, uint32_t inner_items, uint32_t B, int clip) {
  ptx::griddep_wait();
  ptx::griddep_launch();
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_items; i += gridDim.x * kThreads) {
    int b = (int)((i / inner_items) % B);
    DdimCoef c = load_ddim(t, sra, srm1, ac, acp, eta, b);
    if (VEC) {
      float4 vx = reinterpret_cast<const float4*>(x)[i];
      float4 v0 = reinterpret_cast<const float4*>(x0)[i];
      float4 vn = noise ? reinterpret_cast<const float4*>(noise)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      v0.x = clampf(v0.x, clip); v0.y = clampf(v0.y, clip); v0.z = clampf(v0.z, clip); v0.w = clampf(v0.w, clip);
      float4 o;
      o.x = ddim_one(vx.x, v0.x, vn.x, c);
      o.y = ddim_one(vx.y, v0.y, vn.y, c);
      o.z = ddim_one(vx.z, v0.z, vn.z, c);
      o.w = ddim_one(vx.w, v0.w, vn.w, c);
      reinterpret_cast<float4*>(out)[i] = o;
      if (pred) reinterpret_cast<float4*>(pred)[i] = v0;
    } else {
      float v0 = clampf(x0[i], clip);
      out[i] = ddim_one(x[i], v0, noise ? noise[i] : 0.f, c);
      if (pred) pred[i] = v0;
    }
  }
}

// model/cfg_sampler.py:31
template <bool VEC>
__global__ void __launch_bounds__(kThreads) cfg_combine_kernel(const float* __restrict__ cond,
                                                               const float* __restrict__ uncond,
                                                               const float* __restrict__ scale,
                                                               float* __restrict__ out, uint32_t n_items,
                                                               uint32_t inner_items, uint32_t B) {
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_items; i += gridDim.x * kThreads) {
    float s = __ldg(scale + (i / inner_items) % B);
    if (VEC) {
      float4 c = reinterpret_cast<const float4*>(cond)[i];
      float4 u = reinterpret_cast<const float4*>(uncond)[i];
      float4 o;
      o.x = __fadd_rn(u.x, __fmul_rn(s, __fsub_rn(c.x, u.x)));
      o.y = __fadd_rn(u.y, __fmul_rn(s, __fsub_rn(c.y, u.y)));
      o.z = __fadd_rn(u.z, __fmul_rn(s, __fsub_rn(c.z, u.z)));
      o.w = __fadd_rn(u.w, __fmul_rn(s, __fsub_rn(c.w, u.w)));
      reinterpret_cast<float4*>(out)[i] = o;
    } else {
      out[i] = __fadd_rn(uncond[i], __fmul_rn(s, __fsub_rn(cond[i], uncond[i])));
    }
  }
}

// diffusion/gaussian_diffusion.py:319-323 (motion editing / in-betweening): the model output is overwritten where the
// mask is set,  out = x0 * ~mask + motion * mask, with torch's operation order (two products, one sum).  mask holds 0.0 / 1.0.
template <bool VEC>
__global__ void __launch_bounds__(kThreads) inpaint_blend_kernel(float* __restrict__ x0, const float* __restrict__ motion,
                                                                 const float* __restrict__ mask, uint32_t n_items) {
  ptx::griddep_wait();  // PDL: x0 is the output projection's (or the guidance combine's) result
  ptx::griddep_launch();
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_items; i += gridDim.x * kThreads) {
    if (VEC) {
      float4 v = reinterpret_cast<float4*>(x0)[i];
      const float4 m = reinterpret_cast<const float4*>(mask)[i];
      const float4 w = reinterpret_cast<const float4*>(motion)[i];
      v.x = __fadd_rn(__fmul_rn(v.x, 1.f - m.x), __fmul_rn(w.x, m.x));
      v.y = __fadd_rn(__fmul_rn(v.y, 1.f - m.y), __fmul_rn(w.y, m.y));
      v.z = __fadd_rn(__fmul_rn(v.z, 1.f - m.z), __fmul_rn(w.z, m.z));
      v.w = __fadd_rn(__fmul_rn(v.w, 1.f - m.w), __fmul_rn(w.w, m.w));
      reinterpret_cast<float4*>(x0)[i] = v;
    } else {
      x0[i] = __fadd_rn(__fmul_rn(x0[i], 1.f - mask[i]), __fmul_rn(motion[i], mask[i]));
    }
  }
}

// ------------------------------------------------------------------------------------------
// PLMS (pseudo linear multistep) sampler pieces, diffusion/gaussian_diffusion.py:1007-1098.  Off the hot path (no caller
// in the reference uses it); three small HBM-bound kernels keep its arithmetic in the library, in torch's operation order.
// Tables are indexed with t[b] + t_shift, negative indices wrapping like Python's (the corrector evaluates at t - 1).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int plms_index(const int64_t* t, int b, int shift, int n_table) {
  int64_t i = t[b] + shift;
  if (i < 0) i += n_table;
  return (int)(i < 0 ? 0 : (i >= n_table ? n_table - 1 : i));
}

// eps = (sqrt_recip_ac[t] * x - clip(x0)) / sqrt_recipm1_ac[t]      (_predict_eps_from_xstart, :418-423); pred = clip(x0)
__global__ void __launch_bounds__(kThreads) plms_eps_kernel(const float* __restrict__ x, const float* __restrict__ x0,
                                                            float* __restrict__ eps, float* __restrict__ pred,
                                                            const int64_t* __restrict__ t, const float* __restrict__ sra,
                                                            const float* __restrict__ srm1, uint32_t n, uint32_t inner,
                                                            uint32_t B, int n_table, int t_shift, int clip) {
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const int k = plms_index(t, (int)((i / inner) % B), t_shift, n_table);
    const float v0 = clampf(x0[i], clip);
    eps[i] = __fdiv_rn(__fsub_rn(__fmul_rn(__ldg(sra + k), x[i]), v0), __ldg(srm1 + k));
    if (pred) pred[i] = v0;
  }
}

// eps' from the history (e0 newest): order 1..4 = Adams-Bashforth (:1075-1086), 5 = improved Euler (e0 + e1) / 2 (:1066)
__global__ void __launch_bounds__(kThreads) plms_combine_kernel(const float* __restrict__ e0, const float* __restrict__ e1,
                                                                const float* __restrict__ e2, const float* __restrict__ e3,
                                                                float* __restrict__ out, uint32_t n, int order) {
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    float r;
    if (order == 1) r = e0[i];
    else if (order == 2) r = __fdiv_rn(__fsub_rn(__fmul_rn(3.f, e0[i]), e1[i]), 2.f);
    else if (order == 3)
      r = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.f, e0[i]), __fmul_rn(16.f, e1[i])), __fmul_rn(5.f, e2[i])), 12.f);
    else if (order == 4)
      r = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(55.f, e0[i]), __fmul_rn(59.f, e1[i])), __fmul_rn(37.f, e2[i])),
                              __fmul_rn(9.f, e3[i])), 24.f);
    else r = __fdiv_rn(__fadd_rn(e0[i], e1[i]), 2.f);
    out[i] = r;
  }
}

// mode 0: pred' = sra[t] x - srm1[t] eps';  mean = pred' sqrt(abar_prev[t]) + sqrt(1 - abar_prev[t]) eps';
//         out = mean * (t != 0) + pred * (1 - (t != 0))                                             (:1068-1095)
// mode 1: out = pred sqrt(abar_prev[t]) + sqrt(1 - abar_prev[t]) eps'     (improved-Euler predictor, :1064)
__global__ void __launch_bounds__(kThreads) plms_finish_kernel(const float* __restrict__ x, const float* __restrict__ epsp,
                                                               const float* __restrict__ pred, float* __restrict__ out,
                                                               const int64_t* __restrict__ t, const float* __restrict__ sra,
                                                               const float* __restrict__ srm1, const float* __restrict__ acp,
                                                               uint32_t n, uint32_t inner, uint32_t B, int n_table, int mode) {
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const int b = (int)((i / inner) % B);
    const int k = plms_index(t, b, 0, n_table);
    const float ap = __ldg(acp + k);
    const float s1 = __fsqrt_rn(ap), s2 = __fsqrt_rn(__fsub_rn(1.f, ap));
    const float e = epsp[i];
    if (mode == 1) {
      out[i] = __fadd_rn(__fmul_rn(pred[i], s1), __fmul_rn(s2, e));
    } else {
      const float pp = __fsub_rn(__fmul_rn(__ldg(sra + k), x[i]), __fmul_rn(__ldg(srm1 + k), e));
      const float mean = __fadd_rn(__fmul_rn(pp, s1), __fmul_rn(s2, e));
      const float nz = t[b] != 0 ? 1.f : 0.f;
      out[i] = __fadd_rn(__fmul_rn(mean, nz), __fmul_rn(pred[i], __fsub_rn(1.f, nz)));
    }
  }
}

// utils/rotation_conversions.py:529-534.  Warp-private staging: each warp streams kRotPerWarp rotations per iteration
// through its own slice of shared memory (float4-coalesced global traffic in both directions, 24 B in / 36 B out per
// rotation) and synchronises with __syncwarp only -- no block barrier sits between a block's loads and its stores, so the
// warps of an SM stay spread over the load / compute / store phases and keep bytes in flight.
constexpr int kRotPerWarp = 128, kRotWarps = 4, kRotPerBlock = kRotPerWarp * kRotWarps;
__device__ __forceinline__ void rot6d_one(const float* a, float* o) {
  // three 8-byte shared loads: lanes are 24 bytes apart, which is conflict-free for 64-bit accesses (stride 3, odd)
  const float2 p0 = reinterpret_cast<const float2*>(a)[0], p1 = reinterpret_cast<const float2*>(a)[1],
               p2 = reinterpret_cast<const float2*>(a)[2];
  float a1x = p0.x, a1y = p0.y, a1z = p1.x, a2x = p1.y, a2y = p2.x, a2z = p2.y;
  // F.normalize: v / max(||v||, 1e-12)
  float n1 = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a1x, a1x), __fmul_rn(a1y, a1y)), __fmul_rn(a1z, a1z))), 1e-12f);
  float b1x = __fdiv_rn(a1x, n1), b1y = __fdiv_rn(a1y, n1), b1z = __fdiv_rn(a1z, n1);
  float d = __fadd_rn(__fadd_rn(__fmul_rn(b1x, a2x), __fmul_rn(b1y, a2y)), __fmul_rn(b1z, a2z));
  float b2x = __fsub_rn(a2x, __fmul_rn(d, b1x)), b2y = __fsub_rn(a2y, __fmul_rn(d, b1y)), b2z = __fsub_rn(a2z, __fmul_rn(d, b1z));
  float n2 = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(b2x, b2x), __fmul_rn(b2y, b2y)), __fmul_rn(b2z, b2z))), 1e-12f);
  b2x = __fdiv_rn(b2x, n2); b2y = __fdiv_rn(b2y, n2); b2z = __fdiv_rn(b2z, n2);
  o[0] = b1x; o[1] = b1y; o[2] = b1z;
  o[3] = b2x; o[4] = b2y; o[5] = b2z;
  o[6] = __fsub_rn(__fmul_rn(b1y, b2z), __fmul_rn(b1z, b2y));
  o[7] = __fsub_rn(__fmul_rn(b1z, b2x), __fmul_rn(b1x, b2z));
  o[8] = __fsub_rn(__fmul_rn(b1x, b2y), __fmul_rn(b1y, b2x));
}

__global__ void __launch_bounds__(kRotWarps * 32) rot6d_kernel(const float* __restrict__ d6, float* __restrict__ R,
                                                                int64_t n) {
  __shared__ __align__(16) float s_in[kRotWarps][kRotPerWarp * 6];
  __shared__ __align__(16) float s_out[kRotWarps][kRotPerWarp * 9];
  ptx::griddep_wait();
  ptx::griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* in = s_in[warp];
  float* ob = s_out[warp];
  const int64_t n_chunks = (n + kRotPerWarp - 1) / kRotPerWarp;
  for (int64_t chunk = (int64_t)blockIdx.x * kRotWarps + warp; chunk < n_chunks; chunk += (int64_t)gridDim.x * kRotWarps) {
    const int64_t base = chunk * kRotPerWarp;
    const int cnt = (int)min((int64_t)kRotPerWarp, n - base);
    const float* src = d6 + base * 6;  // 24 * base bytes: 16 B aligned because kRotPerWarp * 24 % 16 == 0
    const int nin = cnt * 6;
    if (cnt == kRotPerWarp) {
      float4 v[kRotPerWarp * 6 / 128];   // all of the warp's loads are issued before the first is consumed
#pragma unroll
      for (int k = 0; k < kRotPerWarp * 6 / 128; ++k) v[k] = ld_stream(reinterpret_cast<const float4*>(src) + k * 32 + lane);
#pragma unroll
      for (int k = 0; k < kRotPerWarp * 6 / 128; ++k) reinterpret_cast<float4*>(in)[k * 32 + lane] = v[k];
    } else {
      for (int i = lane; i < nin; i += 32) in[i] = src[i];
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kRotPerWarp / 32; ++r) {
      const int j = r * 32 + lane;       // lane-contiguous rotations: stride 6 / 9 words between lanes
      if (j < cnt) rot6d_one(in + j * 6, ob + j * 9);
    }
    __syncwarp();
    float* dst = R + base * 9;  // 36 * base bytes: 16 B aligned because kRotPerWarp * 36 % 16 == 0
    const int nout = cnt * 9;
    if (cnt == kRotPerWarp) {
#pragma unroll
      for (int k = 0; k < kRotPerWarp * 9 / 128; ++k)
        st_stream(reinterpret_cast<float4*>(dst) + k * 32 + lane, reinterpret_cast<const float4*>(ob)[k * 32 + lane]);
    } else {
      for (int i = lane; i < nout; i += 32) dst[i] = ob[i];
    }
    __syncwarp();
  }
}

// [B, I, T] <-> [T, B, I] through a padded shared-memory tile (per b: an I x T matrix transpose).
constexpr int kTile = 32;
template <bool TO_TBI>
__global__ void __launch_bounds__(kTile * 8) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              int B, int I, int T) {
  __shared__ float tile[kTile][kTile + 1];
  int b = blockIdx.z;
  int i0 = blockIdx.y * kTile, t0 = blockIdx.x * kTile;
  int tx = threadIdx.x % kTile, ty = threadIdx.x / kTile;  // 32 x 8
  if (TO_TBI) {
    // read src[b, i, t] with t fastest; write dst[t, b, i] with i fastest
    for (int r = ty; r < kTile; r += 8) {
      int i = i0 + r, t = t0 + tx;
      if (i < I && t < T) tile[r][tx] = src[((int64_t)b * I + i) * T + t];
    }
    __syncthreads();
    for (int r = ty; r < kTile; r += 8) {
      int t = t0 + r, i = i0 + tx;
      if (i < I && t < T) dst[((int64_t)t * B + b) * I + i] = tile[tx][r];
    }
  } else {
    for (int r = ty; r < kTile; r += 8) {
      int t = t0 + r, i = i0 + tx;
      if (i < I && t < T) tile[r][tx] = src[((int64_t)t * B + b) * I + i];
    }
    __syncthreads();
    for (int r = ty; r < kTile; r += 8) {
      int i = i0 + r, t = t0 + tx;
      if (i < I && t < T) dst[((int64_t)b * I + i) * T + t] = tile[tx][r];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Temporal Gaussian smoothing of generated motions -- scipy.ndimage.gaussian_filter1d(x, sigma, axis=-1) as called at
// sample/cgenerate.py:142 (sigma=1) and render/crendermotion.py:79 (sigma=3): weights exp(-k^2 / 2 sigma^2) / sum over
// |k| <= radius = int(truncate*sigma + 0.5), 'reflect' boundary (d c b a | a b c d | d c b a), double accumulation.
// Element (col, t) lives at col_base(col) + t*t_stride: BJFT: col*T + t;  TBI: col + t*n_cols.
// ------------------------------------------------------------------------------------------------------------
constexpr int kMaxRadius = 64;
__device__ __forceinline__ int reflect_index(int i, int n) {
  // scipy 'reflect' (half-sample symmetric), valid for any offset
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

__device__ __forceinline__ void gaussian_weights(double* sw, int radius, double sigma) {
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = -radius; k <= radius; ++k) {
      const double w = exp(-0.5 * (double)(k * k) / (sigma * sigma));
      sw[k + radius] = w;
      tot += w;
    }
    for (int k = 0; k <= 2 * radius; ++k) sw[k] /= tot;
  }
  __syncthreads();
}

template <bool TBI>
__global__ void __launch_bounds__(256) gaussian_filter_time_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                   int64_t n_cols, int T, int radius, double sigma) {
  __shared__ double sw[2 * kMaxRadius + 1];
  gaussian_weights(sw, radius, sigma);
  const int64_t total = n_cols * T;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    // consecutive threads walk the contiguous dimension: t for BJFT, col for TBI
    const int64_t col = TBI ? e % n_cols : e / T;
    const int t = (int)(TBI ? e / n_cols : e % T);
    const int64_t base = TBI ? col : col * T;
    const int64_t ts = TBI ? n_cols : 1;
    double acc = 0.0;
    for (int k = -radius; k <= radius; ++k) acc += sw[k + radius] * (double)src[base + (int64_t)reflect_index(t + k, T) * ts];
    dst[e] = (float)acc;
  }
}

// Fused tail of sample/cgenerate.py:142-156 + model/rotation2xyz.py:253-270 on the sampler's native layout:
// x [T, B, J, 6] (TBI) -> temporal Gaussian filter -> drop the last `drop` joints (the translation row) ->
// rotation_6d_to_matrix -> R [B, T, J - drop, 3, 3].  One thread per (b, t, joint).
__global__ void __launch_bounds__(256) smooth_rot6d_kernel(const float* __restrict__ x, float* __restrict__ R, int T, int B,
                                                           int J, int drop, int radius, double sigma) {
  __shared__ double sw[2 * kMaxRadius + 1];
  gaussian_weights(sw, radius, sigma);
  const int Jr = J - drop;
  const int64_t total = (int64_t)B * T * Jr;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int j = (int)(e % Jr);
    const int t = (int)((e / Jr) % T);
    const int b = (int)(e / ((int64_t)Jr * T));
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int k = -radius; k <= radius; ++k) {
      const float* p = x + (((int64_t)reflect_index(t + k, T) * B + b) * J + j) * 6;
      const double w = sw[k + radius];
#pragma unroll
      for (int f = 0; f < 6; ++f) a[f] += w * (double)p[f];
    }
    const float a1x = (float)a[0], a1y = (float)a[1], a1z = (float)a[2], a2x = (float)a[3], a2y = (float)a[4], a2z = (float)a[5];
    float n1 = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a1x, a1x), __fmul_rn(a1y, a1y)), __fmul_rn(a1z, a1z))), 1e-12f);
    float b1x = __fdiv_rn(a1x, n1), b1y = __fdiv_rn(a1y, n1), b1z = __fdiv_rn(a1z, n1);
    float d = __fadd_rn(__fadd_rn(__fmul_rn(b1x, a2x), __fmul_rn(b1y, a2y)), __fmul_rn(b1z, a2z));
    float b2x = __fsub_rn(a2x, __fmul_rn(d, b1x)), b2y = __fsub_rn(a2y, __fmul_rn(d, b1y)), b2z = __fsub_rn(a2z, __fmul_rn(d, b1z));
    float n2 = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(b2x, b2x), __fmul_rn(b2y, b2y)), __fmul_rn(b2z, b2z))), 1e-12f);
    b2x = __fdiv_rn(b2x, n2); b2y = __fdiv_rn(b2y, n2); b2z = __fdiv_rn(b2z, n2);
    float* o = R + e * 9;  // e enumerates (b, t, j) in output order
    o[0] = b1x; o[1] = b1y; o[2] = b1z;
    o[3] = b2x; o[4] = b2y; o[5] = b2z;
    o[6] = __fsub_rn(__fmul_rn(b1y, b2z), __fmul_rn(b1z, b2y));
    o[7] = __fsub_rn(__fmul_rn(b1z, b2x), __fmul_rn(b1x, b2z));
    o[8] = __fsub_rn(__fmul_rn(b1x, b2y), __fmul_rn(b1y, b2x));
  }
}

// Device-side step bookkeeping of a CUDA-graph-captured sampling loop: step number pos[0] selects the loop index
// (table index of the update) and the model timestep (respace.py:125-126) from per-loop sequences, broadcasts both to
// the int64[B] vectors the denoiser / update kernels read, and advances pos -- one launch per step, no host values.
__global__ void step_tables_kernel(const int64_t* __restrict__ seq_idx, const int64_t* __restrict__ seq_model,
                                   int64_t* pos, int64_t* __restrict__ t_idx, int64_t* __restrict__ t_model, int B,
                                   int n_seq) {
  ptx::griddep_wait();  // PDL: the previous step's update kernel still reads t_idx
  ptx::griddep_launch();
  int64_t p = pos[0];
  if (p < 0) p = 0;
  if (p >= n_seq) p = n_seq - 1;
  const int64_t a = seq_idx[p], b = seq_model[p];
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    t_idx[i] = a;
    t_model[i] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) pos[0] = p + 1;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace regen

using namespace regen;

extern "C" {

const char* regen_version(void) { return "regen_sm100 0.1.0 sm_100a"; }
const char* regen_last_error(void) { return regen::g_err; }
int64_t regen_launch_count(void) { return (int64_t)regen::g_launches; }
void regen_launch_count_add(int64_t n) { regen::g_launches += n; }

int regen_step_tables(const int64_t* seq_idx, const int64_t* seq_model, int64_t* pos, int64_t* t_idx, int64_t* t_model,
                      int32_t B, int32_t n_seq, void* stream) {
  REGEN_CHECK_ARG(seq_idx && seq_model && pos && t_idx && t_model, "step_tables: null pointer");
  REGEN_CHECK_ARG(B >= 1 && n_seq >= 1, "step_tables: bad sizes");
  REGEN_CUDA(launch_pdl(step_tables_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, seq_idx, seq_model, pos, t_idx,
                        t_model, B, n_seq));
  count_launch();
  return REGEN_OK;
}

int regen_p_sample_update(const float* x, const float* x0, const float* noise, float* out, float* pred_xstart,
                          const int64_t* t, const float* coef1, const float* coef2, const float* logvar,
                          int64_t n_elem, int64_t inner, int32_t B, int32_t clip_denoised, void* stream) {
  REGEN_CHECK_ARG(x && x0 && out && t && coef1 && coef2 && logvar, "p_sample_update: null pointer");
  REGEN_CHECK_ARG(n_elem >= 0 && inner > 0 && B > 0, "p_sample_update: bad sizes");
  REGEN_CHECK_ARG(n_elem < (int64_t)1 << 32, "p_sample_update: n_elem too large");
  if (n_elem == 0) return REGEN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = (inner % 4 == 0) && (n_elem % 4 == 0) && aligned16(x) && aligned16(x0) && (!noise || aligned16(noise)) &&
             aligned16(out) && (!pred_xstart || aligned16(pred_xstart));
  if (vec) {
    uint32_t items = (uint32_t)(n_elem / 4);
    REGEN_CUDA(launch_pdl(p_sample_update_kernel<true>, dim3(grid_for(ceil_div(items, kUpdUnroll))), dim3(kThreads), 0, s, x,
                          x0, noise, out, pred_xstart, t, coef1, coef2, logvar, items, (uint32_t)(inner / 4), (uint32_t)B,
                          (int)clip_denoised));
  } else {
    REGEN_CUDA(launch_pdl(p_sample_update_kernel<false>, dim3(grid_for(n_elem)), dim3(kThreads), 0, s, x, x0, noise, out,
                          pred_xstart, t, coef1, coef2, logvar, (uint32_t)n_elem, (uint32_t)inner, (uint32_t)B,
                          (int)clip_denoised));
  }
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

int regen_ddim_update(const float* x, const float* x0, const float* noise, float* out, float* pred_xstart,
                      const int64_t* t, const float* sqrt_recip_ac, const float* sqrt_recipm1_ac, const float* ac,
                      const float* ac_prev, float eta, int64_t n_elem, int64_t inner, int32_t B, int32_t clip_denoised,
                      void* stream) {
  REGEN_CHECK_ARG(x && x0 && out && t && sqrt_recip_ac && sqrt_recipm1_ac && ac && ac_prev,
                  "ddim_update: null pointer");
  REGEN_CHECK_ARG(n_elem >= 0 && inner > 0 && B > 0, "ddim_update: bad sizes");
  REGEN_CHECK_ARG(n_elem < (int64_t)1 << 32, "ddim_update: n_elem too large");
  if (n_elem == 0) return REGEN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = (inner % 4 == 0) && (n_elem % 4 == 0) && aligned16(x) && aligned16(x0) && (!noise || aligned16(noise)) &&
             aligned16(out) && (!pred_xstart || aligned16(pred_xstart));
  if (vec) {
    uint32_t items = (uint32_t)(n_elem / 4);
    REGEN_CUDA(launch_pdl(ddim_update_kernel<true>, dim3(grid_for(items)), dim3(kThreads), 0, s, x, x0, noise, out,
                          pred_xstart, t, sqrt_recip_ac, sqrt_recipm1_ac, ac, ac_prev, eta, items, (uint32_t)(inner / 4),
                          (uint32_t)B, (int)clip_denoised));
  } else {
    REGEN_CUDA(launch_pdl(ddim_update_kernel<false>, dim3(grid_for(n_elem)), dim3(kThreads), 0, s, x, x0, noise, out,
                          pred_xstart, t, sqrt_recip_ac, sqrt_recipm1_ac, ac, ac_prev, eta, (uint32_t)n_elem, (uint32_t)inner,
                          (uint32_t)B, (int)clip_denoised));
  }
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

int regen_cfg_combine(const float* cond, const float* uncond, const float* scale, float* out, int64_t n_elem,
                      int64_t inner, int32_t B, void* stream) {
  REGEN_CHECK_ARG(cond && uncond && scale && out, "cfg_combine: null pointer");
  REGEN_CHECK_ARG(n_elem >= 0 && inner > 0 && B > 0, "cfg_combine: bad sizes");
  REGEN_CHECK_ARG(n_elem < (int64_t)1 << 32, "cfg_combine: n_elem too large");
  if (n_elem == 0) return REGEN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = (inner % 4 == 0) && (n_elem % 4 == 0) && aligned16(cond) && aligned16(uncond) && aligned16(out);
  if (vec) {
    uint32_t items = (uint32_t)(n_elem / 4);
    cfg_combine_kernel<true><<<grid_for(items), kThreads, 0, s>>>(cond, uncond, scale, out, items,
                                                                  (uint32_t)(inner / 4), (uint32_t)B);
  } else {
    cfg_combine_kernel<false><<<grid_for(n_elem), kThreads, 0, s>>>(cond, uncond, scale, out, (uint32_t)n_elem,
                                                                    (uint32_t)inner, (uint32_t)B);
  }
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

int regen_inpaint_blend(float* x0, const float* motion, const float* mask01, int64_t n_elem, void* stream) {
  REGEN_CHECK_ARG(n_elem >= 0 && n_elem < (int64_t)1 << 32, "inpaint_blend: bad n_elem");
  if (n_elem == 0) return REGEN_OK;
  REGEN_CHECK_ARG(x0 && motion && mask01, "inpaint_blend: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if ((n_elem % 4 == 0) && aligned16(x0) && aligned16(motion) && aligned16(mask01)) {
    const uint32_t items = (uint32_t)(n_elem / 4);
    REGEN_CUDA(launch_pdl(inpaint_blend_kernel<true>, dim3(grid_for(items)), dim3(kThreads), 0, s, x0, motion, mask01, items));
  } else {
    REGEN_CUDA(launch_pdl(inpaint_blend_kernel<false>, dim3(grid_for(n_elem)), dim3(kThreads), 0, s, x0, motion, mask01,
                          (uint32_t)n_elem));
  }
  count_launch();
  return REGEN_OK;
}

int regen_plms_eps(const float* x, const float* x0, float* eps, float* pred, const int64_t* t, const float* sqrt_recip_ac,
                   const float* sqrt_recipm1_ac, int64_t n_elem, int64_t inner, int32_t B, int32_t n_table, int32_t t_shift,
                   int32_t clip_denoised, void* stream) {
  REGEN_CHECK_ARG(x && x0 && eps && t && sqrt_recip_ac && sqrt_recipm1_ac, "plms_eps: null pointer");
  REGEN_CHECK_ARG(n_elem >= 0 && n_elem < (int64_t)1 << 32 && inner > 0 && B > 0 && n_table > 0, "plms_eps: bad sizes");
  if (n_elem == 0) return REGEN_OK;
  plms_eps_kernel<<<grid_for(n_elem), kThreads, 0, (cudaStream_t)stream>>>(x, x0, eps, pred, t, sqrt_recip_ac,
                                                                          sqrt_recipm1_ac, (uint32_t)n_elem,
                                                                          (uint32_t)inner, (uint32_t)B, n_table, t_shift,
                                                                          clip_denoised);
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

int regen_plms_combine(const float* e0, const float* e1, const float* e2, const float* e3, float* out, int64_t n_elem,
                       int32_t order, void* stream) {
  REGEN_CHECK_ARG(e0 && out && order >= 1 && order <= 5, "plms_combine: bad argument");
  REGEN_CHECK_ARG((order == 1) || e1, "plms_combine: history too short for the order");
  REGEN_CHECK_ARG((order != 3 && order != 4) || e2, "plms_combine: history too short for the order");
  REGEN_CHECK_ARG(order != 4 || e3, "plms_combine: history too short for the order");
  REGEN_CHECK_ARG(n_elem >= 0 && n_elem < (int64_t)1 << 32, "plms_combine: bad n_elem");
  if (n_elem == 0) return REGEN_OK;
  plms_combine_kernel<<<grid_for(n_elem), kThreads, 0, (cudaStream_t)stream>>>(e0, e1, e2, e3, out, (uint32_t)n_elem, order);
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

int regen_plms_finish(const float* x, const float* eps_prime, const float* pred, float* out, const int64_t* t,
                      const float* sqrt_recip_ac, const float* sqrt_recipm1_ac, const float* ac_prev, int64_t n_elem,
                      int64_t inner, int32_t B, int32_t n_table, int32_t mode, void* stream) {
  REGEN_CHECK_ARG(eps_prime && pred && out && t && ac_prev && (mode == 0 || mode == 1), "plms_finish: bad argument");
  REGEN_CHECK_ARG(mode == 1 || (x && sqrt_recip_ac && sqrt_recipm1_ac), "plms_finish: null pointer");
  REGEN_CHECK_ARG(n_elem >= 0 && n_elem < (int64_t)1 << 32 && inner > 0 && B > 0 && n_table > 0, "plms_finish: bad sizes");
  if (n_elem == 0) return REGEN_OK;
  plms_finish_kernel<<<grid_for(n_elem), kThreads, 0, (cudaStream_t)stream>>>(x, eps_prime, pred, out, t, sqrt_recip_ac,
                                                                             sqrt_recipm1_ac, ac_prev, (uint32_t)n_elem,
                                                                             (uint32_t)inner, (uint32_t)B, n_table, mode);
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

int regen_rot6d_to_matrix(const float* d6, float* R, int64_t n, void* stream) {
  REGEN_CHECK_ARG(n >= 0, "rot6d_to_matrix: negative n");
  if (n == 0) return REGEN_OK;
  REGEN_CHECK_ARG(d6 && R, "rot6d_to_matrix: null pointer");
  REGEN_CHECK_ARG(aligned16(d6) && aligned16(R), "rot6d_to_matrix: pointers must be 16-byte aligned");
  int64_t blocks = ceil_div(n, kRotPerBlock);
  int64_t cap = (int64_t)kNumSMs * 12;   // 7.5 KB of staging per warp: 12 blocks of 4 warps per SM
  if (blocks > cap) blocks = cap;
  REGEN_CUDA(launch_pdl(rot6d_kernel, dim3((unsigned)blocks), dim3(kRotWarps * 32), 0, (cudaStream_t)stream, d6, R, n));
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

static int launch_transpose(bool to_tbi, const float* src, float* dst, int32_t B, int32_t I, int32_t T, void* stream) {
  REGEN_CHECK_ARG(B >= 0 && I >= 0 && T >= 0, "layout: negative size");
  if (B == 0 || I == 0 || T == 0) return REGEN_OK;
  REGEN_CHECK_ARG(src && dst, "layout: null pointer");
  REGEN_CHECK_ARG(B <= 65535 && ceil_div(I, kTile) <= 65535, "layout: B or I too large for one launch");
  dim3 grid((unsigned)ceil_div(T, kTile), (unsigned)ceil_div(I, kTile), (unsigned)B);
  if (to_tbi)
    transpose_kernel<true><<<grid, kTile * 8, 0, (cudaStream_t)stream>>>(src, dst, B, I, T);
  else
    transpose_kernel<false><<<grid, kTile * 8, 0, (cudaStream_t)stream>>>(src, dst, B, I, T);
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

int regen_bjft_to_tbi(const float* src, float* dst, int32_t B, int32_t I, int32_t T, void* stream) {
  return launch_transpose(true, src, dst, B, I, T, stream);
}
int regen_tbi_to_bjft(const float* src, float* dst, int32_t B, int32_t I, int32_t T, void* stream) {
  return launch_transpose(false, src, dst, B, I, T, stream);
}

static int gaussian_radius(double sigma, double truncate) { return (int)(truncate * sigma + 0.5); }

int regen_gaussian_filter1d_time(const float* src, float* dst, int64_t n_cols, int32_t T, int32_t layout, double sigma,
                                 double truncate, void* stream) {
  REGEN_CHECK_ARG(n_cols >= 0 && T >= 0, "gaussian_filter1d_time: negative size");
  if (n_cols == 0 || T == 0) return REGEN_OK;
  REGEN_CHECK_ARG(src && dst && src != dst, "gaussian_filter1d_time: null or aliased pointers");
  REGEN_CHECK_ARG(sigma > 0.0 && truncate > 0.0, "gaussian_filter1d_time: sigma and truncate must be positive");
  const int radius = gaussian_radius(sigma, truncate);
  REGEN_CHECK_ARG(radius <= kMaxRadius, "gaussian_filter1d_time: radius %d exceeds %d", radius, kMaxRadius);
  REGEN_CHECK_ARG(layout == 0 || layout == 1, "gaussian_filter1d_time: layout must be 0 (BJFT) or 1 (TBI)");
  const int blocks = grid_for(n_cols * T);
  if (layout == 1)
    gaussian_filter_time_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n_cols, T, radius, sigma);
  else
    gaussian_filter_time_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n_cols, T, radius, sigma);
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

int regen_smooth_rot6d_to_matrix(const float* x_tbi, float* R, int32_t T, int32_t B, int32_t J, int32_t drop_joints,
                                 double sigma, double truncate, void* stream) {
  REGEN_CHECK_ARG(T >= 0 && B >= 0 && J >= 0 && drop_joints >= 0 && drop_joints <= J, "smooth_rot6d: bad sizes");
  if (T == 0 || B == 0 || J == drop_joints) return REGEN_OK;
  REGEN_CHECK_ARG(x_tbi && R, "smooth_rot6d: null pointer");
  REGEN_CHECK_ARG(sigma > 0.0 && truncate > 0.0, "smooth_rot6d: sigma and truncate must be positive");
  const int radius = gaussian_radius(sigma, truncate);
  REGEN_CHECK_ARG(radius <= kMaxRadius, "smooth_rot6d: radius %d exceeds %d", radius, kMaxRadius);
  const int blocks = grid_for((int64_t)B * T * (J - drop_joints));
  smooth_rot6d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x_tbi, R, T, B, J, drop_joints, radius, sigma);
  REGEN_LAUNCH_CHECK();
  count_launch();
  return REGEN_OK;
}

}  // extern "C"
