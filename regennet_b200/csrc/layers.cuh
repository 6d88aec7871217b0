// Non-GEMM kernels of the CMDM denoiser: operand packing, LayerNorm (single and chained
// LN1 -> +cross-attention constant -> LN2), causal self-attention, exact-fp32 setup GEMM, CFG combine.
// Token rows are seq-first: row = t * B + b (the reference's [T, B, D] layout, model/cmdm.py:312-313).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace regen {
namespace layers {

constexpr int D = 512;       // latent_dim
constexpr int H = 4;         // heads
constexpr int HD = 128;      // head dim
constexpr int FF = 1024;     // ff_size
constexpr float LN_EPS = 1e-5f;

// ------------------------------------------------------------------------------------------
// fp32 [R, C] -> bf16 (hi, lo) [R_out, Cpad], zero padded columns.  dup > 1 replicates the batch:
// out row (t, b') with b' in [0, B*dup) reads src row (t, b' % B)   (classifier-free guidance).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ src, int ld_src,
                                                         __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, int Cpad, int C, int R_out,
                                                         int B, int dup) {
  ptx::griddep_wait();  // PDL (launch_pdl below): src may be the previous kernel's output
  ptx::griddep_launch();
  const int c4n = Cpad / 4;
  const int64_t total = (int64_t)R_out * c4n;
  const int Beff = B * dup;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int r = (int)(i / c4n), c = (int)(i % c4n) * 4;
    int sr = dup == 1 ? r : (r / Beff) * B + (r % Beff) % B;
    float v[4];
    const float* s = src + (size_t)sr * ld_src + c;
    if (c + 3 < C && (ld_src & 3) == 0) {
      float4 f = *reinterpret_cast<const float4*>(s);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = (c + e < C) ? s[e] : 0.f;
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_bf16(v[e], h[e], l[e]);
    *reinterpret_cast<uint2*>(hi + (size_t)r * Cpad + c) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(lo + (size_t)r * Cpad + c) = *reinterpret_cast<uint2*>(l);
  }
}

inline void launch_split_rows(const float* src, int ld_src, __nv_bfloat16* hi, __nv_bfloat16* lo, int Cpad, int C,
                              int R_out, int B, int dup, cudaStream_t s) {
  int64_t total = (int64_t)R_out * (Cpad / 4);
  int blocks = grid_cap(ceil_div(total, 256));
  launch_pdl(split_rows_kernel, dim3(blocks), dim3(256), 0, s, src, ld_src, hi, lo, Cpad, C, R_out, B, dup);
  count_launch();
}

// ------------------------------------------------------------------------------------------
// fp32 [R, C] -> the mixed8 operand pack of gemm_ln_sm100.cuh / gemm_sm100.cuh: x16 fp16 [R, C] and bytes [R, 2 C], per group
// of 64 columns.  Weights (act = 0, setup only): 64 bytes e4m3(fp16(w) * 2^6) | 64 bytes e4m3((w - fp16(w)) * 2^17).
// Activations (act = 1; test hook and the split of the sampler state x -- everywhere else the producing epilogues write
// the pack themselves): 64 bytes e4m3((a - fp16(a)) * 2^9) | 64 bytes e4m3(fp16(a) * 2^-2).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_m8_kernel(const float* __restrict__ src, uint16_t* __restrict__ x16,
                                                      uint8_t* __restrict__ x8, int R, int C, int act) {
  const int64_t total = (int64_t)R * (C / 4);
  const float s_hi = act ? 0.25f : 64.f, s_lo = act ? 512.f : 131072.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int r = (int)(i / (C / 4)), c = (int)(i % (C / 4)) * 4;
    const float4 f = *reinterpret_cast<const float4*>(src + (size_t)r * C + c);
    const uint32_t h0 = ptx::pack_f16x2_sat(f.x, f.y), h1 = ptx::pack_f16x2_sat(f.z, f.w);
    float a0, a1, a2, a3;
    ptx::upk2(ptx::f16x2_to_f32x2(h0), a0, a1);
    ptx::upk2(ptx::f16x2_to_f32x2(h1), a2, a3);
    *reinterpret_cast<uint2*>(x16 + (size_t)r * C + c) = make_uint2(h0, h1);
    uint8_t* grp = x8 + (size_t)r * 2 * C + (c >> 6) * 128 + (c & 63);
    const uint32_t hi8 = ptx::pack_e4m3x4(a0 * s_hi, a1 * s_hi, a2 * s_hi, a3 * s_hi);
    const uint32_t lo8 = ptx::pack_e4m3x4((f.x - a0) * s_lo, (f.y - a1) * s_lo, (f.z - a2) * s_lo, (f.w - a3) * s_lo);
    *reinterpret_cast<uint32_t*>(grp) = act ? lo8 : hi8;
    *reinterpret_cast<uint32_t*>(grp + 64) = act ? hi8 : lo8;
  }
}

// ------------------------------------------------------------------------------------------
// Exact fp32 GEMM on CUDA cores for one-off setup work (weight folding, timestep table, the
// loop-invariant cmotion embedding):
//   C[m,n] = act( sum_k A[m*sam + k*sak] * Bm[k*sbk + n*sbn] + bias_n[n] + bias_m[m] )
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                    const float* __restrict__ Bm, int64_t sbk, int64_t sbn,
                                                    const float* __restrict__ bias_n,
                                                    const float* __restrict__ bias_m, float* __restrict__ Cout,
                                                    int64_t ldc, int M, int N, int K, int act) {
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int mm = m0 + ty + r * 8, kk = k0 + tx;
      As[ty + r * 8][tx] = (mm < M && kk < K) ? A[mm * sam + kk * sak] : 0.f;
      int kb = k0 + ty + r * 8, nn = n0 + tx;
      Bs[ty + r * 8][tx] = (kb < K && nn < N) ? Bm[kb * sbk + nn * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      float b = Bs[kk][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[ty + r * 8][kk], b, acc[r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int mm = m0 + ty + r * 8, nn = n0 + tx;
    if (mm < M && nn < N) {
      float v = acc[r];
      if (bias_n) v += bias_n[nn];
      if (bias_m) v += bias_m[mm];
      if (act == 1) v = v / (1.f + expf(-v));  // SiLU
      Cout[mm * ldc + nn] = v;
    }
  }
}

inline void launch_sgemm(const float* A, int64_t sam, int64_t sak, const float* Bm, int64_t sbk, int64_t sbn,
                         const float* bias_n, const float* bias_m, float* C, int64_t ldc, int M, int N, int K, int act,
                         cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, 32));
  sgemm_kernel<<<grid, 256, 0, s>>>(A, sam, sak, Bm, sbk, sbn, bias_n, bias_m, C, ldc, M, N, K, act);
  count_launch();
}

// out[t, b', :] = in[t, b' % B, :] + pe[t, :]    (b' in [0, B*dup)) -- finishes the loop-invariant
// conditioning bias: fuse(cmotion embedding) + biases + positional encoding (model/cmdm.py:207-218)
__global__ void __launch_bounds__(256) finalize_condbias_kernel(const float* __restrict__ in,
                                                                const float* __restrict__ pe,
                                                                float* __restrict__ out, int T, int B, int dup) {
  const int Beff = B * dup;
  const int64_t total = (int64_t)T * Beff * (D / 4);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int c4 = (int)(i % (D / 4));
    int64_t r = i / (D / 4);
    int t = (int)(r / Beff), b = (int)(r % Beff) % B;
    float4 v = reinterpret_cast<const float4*>(in)[((int64_t)t * B + b) * (D / 4) + c4];
    float4 p = reinterpret_cast<const float4*>(pe)[(int64_t)t * (D / 4) + c4];
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// out = float(hi) + float(lo)   (test hooks: recombine a bf16 pair)
__global__ void __launch_bounds__(256) merge_split_kernel(const __nv_bfloat16* __restrict__ hi,
                                                          const __nv_bfloat16* __restrict__ lo,
                                                          float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    out[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}

// out[b, :] = vec[:]   (D = 512; one block of 128 threads per row)
__global__ void __launch_bounds__(128) broadcast_rows_kernel(const float* __restrict__ vec, float* __restrict__ out) {
  reinterpret_cast<float4*>(out + (size_t)blockIdx.x * D)[threadIdx.x] =
      __ldg(reinterpret_cast<const float4*>(vec) + threadIdx.x);
}

// out[b, :] (+)= table[clamp(idx[b]), :]   -- EmbedAction row gather (model/cmdm.py:363-366), D = 512
__global__ void __launch_bounds__(128) gather_rows_kernel(const float* __restrict__ table,
                                                          const int64_t* __restrict__ idx, float* __restrict__ out,
                                                          int num_rows, int accumulate) {
  const int b = blockIdx.x;
  int64_t r = idx[b];
  r = r < 0 ? 0 : (r >= num_rows ? num_rows - 1 : r);
  float4 v = reinterpret_cast<const float4*>(table + r * D)[threadIdx.x];
  float4* o = reinterpret_cast<float4*>(out + (size_t)b * D) + threadIdx.x;
  if (accumulate) {
    float4 a = *o;
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  *o = v;
}

// arch 'offline' (model/cmdm.py:234-235): token 0 of sample b' is the condition embedding,
//   h[b', :] = (e2tab[t[b' % B]] + cond_emb[b']) + pe[0]     -> fp32 row + bf16 (hi, lo) split, rows [0, Beff)
__global__ void __launch_bounds__(128) build_emb_rows_kernel(const float* __restrict__ e2tab,
                                                             const float* __restrict__ cond_emb,
                                                             const float* __restrict__ pe0, const int64_t* __restrict__ t,
                                                             float* __restrict__ h, __nv_bfloat16* __restrict__ hi,
                                                             __nv_bfloat16* __restrict__ lo, int B, int n_table) {
  ptx::griddep_wait();  // PDL: t comes from the step-bookkeeping kernel; h rows are read by the previous step's GEMMs
  ptx::griddep_launch();
  const int be = blockIdx.x;
  int64_t tb = t[be % B];
  tb = tb < 0 ? 0 : (tb >= n_table ? n_table - 1 : tb);
  float4 v = __ldg(reinterpret_cast<const float4*>(e2tab + (size_t)tb * D) + threadIdx.x);
  if (cond_emb) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(cond_emb + (size_t)be * D) + threadIdx.x);
    v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
  }
  const float4 p = __ldg(reinterpret_cast<const float4*>(pe0) + threadIdx.x);
  v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
  reinterpret_cast<float4*>(h + (size_t)be * D)[threadIdx.x] = v;
  __nv_bfloat16 hh[4], ll[4];
  split_bf16(v.x, hh[0], ll[0]); split_bf16(v.y, hh[1], ll[1]); split_bf16(v.z, hh[2], ll[2]); split_bf16(v.w, hh[3], ll[3]);
  *reinterpret_cast<uint2*>(hi + (size_t)be * D + 4 * threadIdx.x) = *reinterpret_cast<uint2*>(hh);
  *reinterpret_cast<uint2*>(lo + (size_t)be * D + 4 * threadIdx.x) = *reinterpret_cast<uint2*>(ll);
}

// cyc[l][r][:] = ctab[t[(r % Beff) % B]][l][:] (+ ccond[r % Beff][l][:])  for r in [0, Beff + 32):
// the folded cross-attention constant of every layer as a row-cyclic table, so that the 32 consecutive token rows
// (t, b0 .. b0+31 wrapping into t+1) of an epilogue slice are ONE 2-D TMA box starting at row (row0 % Beff).
__global__ void __launch_bounds__(128) build_cyc_kernel(const float* __restrict__ ctab, const float* __restrict__ ccond,
                                                        const int64_t* __restrict__ t, float* __restrict__ cyc, int L,
                                                        int B, int Beff, int n_table) {
  ptx::griddep_wait();  // PDL: t is written by the step-bookkeeping kernel, cyc is read by the previous step's GEMMs
  ptx::griddep_launch();
  const int r = blockIdx.x, l = blockIdx.y;
  const int be = r % Beff;
  int64_t tb = t[be % B];
  tb = tb < 0 ? 0 : (tb >= n_table ? n_table - 1 : tb);
  float4 v = __ldg(reinterpret_cast<const float4*>(ctab + ((size_t)tb * L + l) * D) + threadIdx.x);
  if (ccond) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(ccond + ((size_t)be * L + l) * D) + threadIdx.x);
    v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
  }
  reinterpret_cast<float4*>(cyc + ((size_t)l * (Beff + 32) + r) * D)[threadIdx.x] = v;
}

// ------------------------------------------------------------------------------------------
// LayerNorm over D=512, one warp per token row.
//   CHAIN = false:  y = LN(in; g1, b1)                                      (norm3)
//   CHAIN = true :  y = LN( LN(in; g1, b1) + c[b]; g2, b2 )                 (norm1 -> cross-attn -> norm2)
// where c[b] = ctab[t[b % B]] (+ ccond[b]) is the folded 1-token cross-attention output
// out_proj(v_proj(emb_b)) (model/cmdm.py:224-227 with a length-1 memory: softmax over one key == 1).
// Writes the fp32 residual stream and its bf16 (hi, lo) split (A operand of the next GEMM).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ln_row(float (&v)[16], const float* __restrict__ g, const float* __restrict__ b,
                                       int lane) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) s += v[e];
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    v[e] -= mean;
    q += v[e] * v[e];
  }
  const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / D) + LN_EPS);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + c * 32 + lane);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(b) + c * 32 + lane);
    v[c * 4 + 0] = v[c * 4 + 0] * rstd * g4.x + b4.x;
    v[c * 4 + 1] = v[c * 4 + 1] * rstd * g4.y + b4.y;
    v[c * 4 + 2] = v[c * 4 + 2] * rstd * g4.z + b4.z;
    v[c * 4 + 3] = v[c * 4 + 3] * rstd * g4.w + b4.w;
  }
}

struct LnParams {
  const float* in;      // [M, D]
  const float *g1, *b1, *g2, *b2;
  const float* ctab;    // [n_table, ld_c] + layer offset applied by the caller
  const float* ccond;   // [Beff, ld_c] + layer offset, or null
  const int64_t* t;     // [B] original timesteps
  int ld_c;
  int B, Beff, n_table;
  float* out_f32;       // [M, D]
  __nv_bfloat16 *out_hi, *out_lo;
  int M;
};

template <bool CHAIN>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
  ptx::griddep_wait();  // PDL: p.in is the preceding GEMM's output
  ptx::griddep_launch();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= p.M) return;
  float v[16];
  const float4* src = reinterpret_cast<const float4*>(p.in + (size_t)row * D);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float4 f = src[c * 32 + lane];
    v[c * 4 + 0] = f.x; v[c * 4 + 1] = f.y; v[c * 4 + 2] = f.z; v[c * 4 + 3] = f.w;
  }
  ln_row(v, p.g1, p.b1, lane);
  if (CHAIN) {
    const int be = row % p.Beff;
    int64_t tb = p.t[be % p.B];
    tb = tb < 0 ? 0 : (tb >= p.n_table ? p.n_table - 1 : tb);
    const float4* ct = reinterpret_cast<const float4*>(p.ctab + (size_t)tb * p.ld_c);
    const float4* cc = p.ccond ? reinterpret_cast<const float4*>(p.ccond + (size_t)be * p.ld_c) : nullptr;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 a = __ldg(ct + c * 32 + lane);
      if (cc) {
        float4 b = __ldg(cc + c * 32 + lane);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      v[c * 4 + 0] += a.x; v[c * 4 + 1] += a.y; v[c * 4 + 2] += a.z; v[c * 4 + 3] += a.w;
    }
    ln_row(v, p.g2, p.b2, lane);
  }
  float4* of = reinterpret_cast<float4*>(p.out_f32 + (size_t)row * D);
  uint2* oh = reinterpret_cast<uint2*>(p.out_hi + (size_t)row * D);
  uint2* ol = reinterpret_cast<uint2*>(p.out_lo + (size_t)row * D);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    of[c * 32 + lane] = make_float4(v[c * 4 + 0], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_bf16(v[c * 4 + e], h[e], l[e]);
    oh[c * 32 + lane] = *reinterpret_cast<uint2*>(h);
    ol[c * 32 + lane] = *reinterpret_cast<uint2*>(l);
  }
}

// ------------------------------------------------------------------------------------------
// Causal self-attention, fp32 on CUDA cores (bring-up version; one CTA per (sample, head)).
// qkv: [T*Beff, 3D] fp32 (q | k | v), rows seq-first.  P = softmax(q k^T / sqrt(HD) + causal mask)
// (model/cmdm.py:168-171, 220-227), out = P v written as a bf16 (hi, lo) pair [T*Beff, D].
// ------------------------------------------------------------------------------------------
constexpr int ATT_WARPS = 8;
constexpr int KSTR = HD + 4;  // padded K row stride: conflict-free float4 reads across keys

inline size_t attention_smem_bytes(int T) {
  int Tp = (T + 31) / 32 * 32;
  return ((size_t)T * KSTR + (size_t)T * HD + ATT_WARPS * (HD + Tp)) * sizeof(float);
}

__global__ void __launch_bounds__(ATT_WARPS * 32) attention_simt_kernel(const float* __restrict__ qkv,
                                                                        __nv_bfloat16* __restrict__ out_hi,
                                                                        __nv_bfloat16* __restrict__ out_lo, int T,
                                                                        int Beff) {
  extern __shared__ __align__(16) float att_smem[];
  const int Tp = (T + 31) / 32 * 32;
  float* sK = att_smem;
  float* sV = sK + (size_t)T * KSTR;
  float* sQ = sV + (size_t)T * HD;
  float* sP = sQ + ATT_WARPS * HD;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < T * (HD / 4); i += ATT_WARPS * 32) {
    int t = i / (HD / 4), c = (i % (HD / 4)) * 4;
    const float* base = qkv + ((size_t)t * Beff + b) * (3 * D) + h * HD + c;
    *reinterpret_cast<float4*>(sK + (size_t)t * KSTR + c) = *reinterpret_cast<const float4*>(base + D);
    *reinterpret_cast<float4*>(sV + (size_t)t * HD + c) = *reinterpret_cast<const float4*>(base + 2 * D);
  }
  __syncthreads();

  const float scale = 0.08838834764831845f;  // 1/sqrt(128)
  float* q = sQ + warp * HD;
  float* pr = sP + warp * Tp;
  for (int i = warp; i < T; i += ATT_WARPS) {
    const size_t row = (size_t)i * Beff + b;
    *reinterpret_cast<float4*>(q + lane * 4) = *reinterpret_cast<const float4*>(qkv + row * (3 * D) + h * HD + lane * 4);
    __syncwarp();
    float mx = -INFINITY;
    for (int j0 = 0; j0 <= i; j0 += 32) {
      const int j = j0 + lane;
      float s = -INFINITY;
      if (j <= i) {
        const float4* kr = reinterpret_cast<const float4*>(sK + (size_t)j * KSTR);
        const float4* qr = reinterpret_cast<const float4*>(q);
        float a = 0.f;
#pragma unroll 8
        for (int d = 0; d < HD / 4; ++d) {
          const float4 kk = kr[d], qq = qr[d];
          a = fmaf(qq.x, kk.x, a); a = fmaf(qq.y, kk.y, a); a = fmaf(qq.z, kk.z, a); a = fmaf(qq.w, kk.w, a);
        }
        s = a * scale;
      }
      pr[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j <= i; j += 32) {
      const float e = expf(pr[j] - mx);
      pr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    __syncwarp();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j <= i; ++j) {
      const float pj = pr[j];
      const float4 vv = *reinterpret_cast<const float4*>(sV + (size_t)j * HD + lane * 4);
      acc.x = fmaf(pj, vv.x, acc.x); acc.y = fmaf(pj, vv.y, acc.y);
      acc.z = fmaf(pj, vv.z, acc.z); acc.w = fmaf(pj, vv.w, acc.w);
    }
    float o[4] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};
    __nv_bfloat16 hh[4], ll[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_bf16(o[e], hh[e], ll[e]);
    *reinterpret_cast<uint2*>(out_hi + row * D + h * HD + lane * 4) = *reinterpret_cast<uint2*>(hh);
    *reinterpret_cast<uint2*>(out_lo + row * D + h * HD + lane * 4) = *reinterpret_cast<uint2*>(ll);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Classifier-free guidance on the doubled batch: rows (t, b) conditional, (t, B + b) unconditional.
//   out[t, b, :] = u + scale[b] * (c - u)                                   (model/cfg_sampler.py:31)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cfg_rows_kernel(const float* __restrict__ x0e, const float* __restrict__ scale,
                                                       float* __restrict__ out, int T, int B, int I) {
  ptx::griddep_wait();  // PDL: x0e comes from the output-projection GEMM
  ptx::griddep_launch();
  const int64_t total = (int64_t)T * B * I;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int c = (int)(i % I);
    int64_t r = i / I;
    int t = (int)(r / B), b = (int)(r % B);
    const float cv = x0e[((int64_t)t * 2 * B + b) * I + c];
    const float uv = x0e[((int64_t)t * 2 * B + B + b) * I + c];
    out[i] = __fadd_rn(uv, __fmul_rn(__ldg(scale + b), __fsub_rn(cv, uv)));
  }
}

}  // namespace layers
}  // namespace regen
