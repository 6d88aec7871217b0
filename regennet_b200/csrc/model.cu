// CMDM denoiser (arch='online') on sm_100a: handle, weight packing, loop-invariant conditioning and
// the per-step forward.  See include/regen_sm100.h for the ABI and DESIGN.md for the data layout.
#include <vector>

#include "common.cuh"
#include "attention_sm100.cuh"
#include "gemm_sm100.cuh"
#include "gemm_ln_sm100.cuh"
#include "layers.cuh"
#include "tmap.cuh"

using namespace regen;
using layers::D;
using layers::FF;
typedef __nv_bfloat16 bf16;

namespace {

struct SplitBuf {  // bf16 (hi, lo) operand pair + its TMA maps
  bf16 *hi = nullptr, *lo = nullptr;
  CUtensorMap tm_hi, tm_lo;      // box rows: 128 for activations (A operand), 256 for weights (single-CTA kernel)
  CUtensorMap tm_hi2, tm_lo2;    // weights only: box rows 128 = one CTA's half of the W tile in the CTA-pair kernel
  CUtensorMap tm_hi3, tm_lo3;    // weights only: box rows 64 = the W tile of the small-batch single-CTA kernel (BN = 64)
                                 //               and of a half-width tail slice of the CTA-pair kernel
  CUtensorMap tm_hi4, tm_lo4;    // weights only: box rows 32 = one CTA's half of a quarter-width tail slice (pair kernel)
  CUtensorMap st_hi, st_lo;      // activations only: store-side maps (box 32 x 16), rebuilt per prepare_cond with rows = M
  CUtensorMap st32_hi, st32_lo;  // same with box 32 x 32 (16-warp GEMM epilogue)
  CUtensorMap st64_hi, st64_lo;  // same with box 32 x 64, 128-byte rows (fused GEMM+LayerNorm epilogue)
  size_t cols = 0;
  bool has64 = false;            // tm_hi3 / tm_lo3 (64-row boxes) are valid
};

// mixed8 pack of one weight matrix [rows, K] for the pair GEMM's mixed8 main loop: fp16 halves + byte rows [rows, 2K], with
// the 128- (tile), 64- and 32-row (tail slice) boxes of both
struct M8W {
  uint16_t* w16 = nullptr;
  uint8_t* w8 = nullptr;
  CUtensorMap tm16, tm8, tm16_64, tm8_64, tm16_32, tm8_32;
};

struct LayerDev {
  SplitBuf wqkv, wo, w1, w2;
  M8W m_qkv, m_w1;   // precision 'mixed8h' only
  // precision 'mixed8' only: linear2 weights as fp16 [512, 1024] + e4m3 bytes [512, 2048] (per 64 columns: hi * 2^6 | lo * 2^17) for the
  // fused linear2 + LayerNorm kernel (the bf16 pair above still serves the small-batch route)
  uint16_t *w2_16 = nullptr, *wo_16 = nullptr;
  uint8_t *w2_8 = nullptr, *wo_8 = nullptr;
  CUtensorMap tm_w2_16, tm_w2_8, tm_wo_16, tm_wo_8;   // wo_*: the same pack of the attention output projection [512, 512]
  float *bqkv, *bo, *b1, *b2, *n1w, *n1b, *n2w, *n2b, *n3w, *n3b;
};

}  // namespace

struct regen_handle {
  regen_model_desc desc;
  int device = 0;
  int L = 0, I = 0, Kin = 0, Mmax = 0;
  bool loaded = false, cond_ready = false;
  bool offline = false;  // arch 1: encoder over S = T + 1 tokens (condition token first)
  // current problem (set by prepare_cond): S tokens per sample (T, or T + 1 offline), M = S * Beff token rows
  int B = 0, Beff = 0, T = 0, S = 0, M = 0;
  bool guidance = false, has_cond = false;

  std::vector<void*> allocs;
  // optional per-kernel-class CUDA-event profiling (regen_profile_begin / regen_profile_end)
  bool prof_on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[4];
  // weights
  SplitBuf w_in, w_out;
  float *b_out = nullptr, *w_c = nullptr, *b_c = nullptr, *P = nullptr, *qvec = nullptr, *ctab = nullptr, *pe = nullptr;
  float *action_emb = nullptr, *text_w = nullptr, *text_b = nullptr, *cond_emb = nullptr;
  int pe_len = 0, num_actions = 0, clip_dim = 0;
  LayerDev layer[REGEN_MAX_LAYERS];
  // activations
  SplitBuf a_in, h_s, att, ffn, qkv_s;
  SplitBuf h_fr;                     // offline: view of h_s without the condition-token rows (A operand of the output
                                     // projection, bf16 outputs of the input projection); maps rebuilt per prepare_cond
  CUtensorMap st_h_fr, st32_h_fr;    // offline: fp32 h without the condition-token rows (stores of the input projection)
  CUtensorMap ld32_condbias;         // conditioning bias [T*Beff, 512] as 32 x 32 boxes (residual load of the input projection)
  float* e2tab = nullptr;            // offline: timestep-embedding table [num_table_steps, 512] (model/cmdm.py:291-298)
  CUtensorMap tm_qkv_hi, tm_qkv_lo;  // 3-D [T, Beff, 1536] views of qkv_s for the attention kernel (per prepare_cond)
  CUtensorMap tm_att_hi, tm_att_lo;  // 3-D [T, Beff, 512] store views of the attention output (box 32 frames x 64 d)
  // precision 'mixed8', fused route: the FFN activations leave the FFN1 epilogue as fp16 in ffn.hi's memory and as e4m3
  // bytes [M, 2048] (per 64 columns: (v - fp16(v)) * 2^9 | fp16(v) / 4) in ffn.lo's memory
  CUtensorMap tm_ffn8, st_ffn8;      // load map (box 128 rows x 128 B) / store map (box 32 rows x 128 B, rows = M)
  // the attention output the same way: fp16 in att.hi's memory, bytes [M, 1024] in att.lo's memory
  CUtensorMap tm_att8, st_att8;      // 2-D load map (box 128 rows x 128 B) / 3-D store map [S, Beff, 1024] (box 32 frames x 128 B)
  // precision 'mixed8h', fused route, arch 'online': the residual stream h the same way (fp16 in h_s.hi's memory, bytes
  // [M, 1024] in h_s.lo's memory), read by the mixed8 main loop of the QKV / FFN1 / output GEMMs (A operand) and as the residual
  CUtensorMap tm_h8, st_h8;          // load map (box 128 rows x 128 B) / store + residual map (box 32 rows x 128 B, rows = M)
  M8W m_out;                         // output projection weights in the mixed8 pack
  CUtensorMap st_h, st_tmp, st_x0e;  // store-side maps of the fp32 activation buffers (rows = M)
  CUtensorMap st32_h;                // h with box 32 x 32: residual load + store of the fused GEMM+LayerNorm kernel
  bool tma_store = true;             // REGEN_DEBUG_NO_TMA_STORE=1: st.global epilogue (A/B measurements)
  bool fused_ln = true;              // REGEN_DEBUG_NO_FUSED_LN=1: GEMM -> tmp -> LayerNorm kernels
  unsigned long long* steplog = nullptr;  // regen_test_step_log: whole-step timeline buffer (2 words per launch slot)
  int steplog_slot = 0, steplog_cap = 0;
  int attn_mc_mode = 2;              // multi-chunk compact attention kernel (2 CTAs / SM) for 64 < S <= 256: 2 = auto (when the
                                     // last 128-query block would be at most half full), REGEN_ATTN_MC=0 never, =1 always
  bool attn_mc = false;              // decision for the current S (set by regen_prepare_cond)
  bool narrow_slices = true;         // REGEN_DEBUG_WIDE_SLICES=1: tail slices of the pair GEMM load the full-height W box (A/B)
  bool exit_wait_full = true;        // REGEN_DEBUG_EXIT_WAIT_READ=1: kernels only wait until their bulk stores have READ the staging
                                     // tiles before exit (measured: no difference, so the conservative full wait stays the default)
  int nob16 = 2;                     // REGEN_DEBUG_LN_NOB=3: three output staging tile pairs in the R16 final pass (A/B)
  bool store64 = true;               // REGEN_DEBUG_NO_STORE64=1: 32-column store boxes in the 16-warp GEMM epilogue (A/B)
  bool res16 = true;                 // REGEN_DEBUG_F32_RESIDUAL=1: fused GEMM+LN kernels keep the fp32 copy of h (A/B)
  bool prefetch_res = true;          // REGEN_DEBUG_NO_RES_PREFETCH=1: no L2 prefetch of the residual tile (A/B)
  float* cyc = nullptr;              // [L][max_batch + 32][512] row-cyclic cross-attention constants (per denoise)
  CUtensorMap tm_cyc[REGEN_MAX_LAYERS];
  bool simt_attention = false;       // REGEN_DEBUG_SIMT_ATTENTION=1: fp32 CUDA-core attention for A/B debugging
  // L2 eviction-priority hints on the TMA traffic of the layer pipeline (REGEN_L2_HINTS bitmask, A/B switch):
  //   1 attention reads q | k | v evict_first   2 bf16 (hi, lo) outputs of every kernel evict_last
  //   4 weight tiles evict_last                  8 fused GEMM+LN A tiles (attention output / FFN activations) evict_first
  int l2_hints = 0;
  unsigned long long pol(int bit, unsigned long long policy) const { return (l2_hints & bit) ? policy : 0ull; }
  float *h = nullptr, *qkv = nullptr, *tmp = nullptr, *x0e = nullptr, *condbias = nullptr, *cmo_tbi = nullptr,
        *ccond = nullptr, *scratch = nullptr;

  template <typename T>
  int alloc(T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
      return REGEN_ECUDA;
    }
    allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return REGEN_OK;
  }
};

static unsigned long long* g_test_timeline = nullptr;  // device buffer [128], set by regen_test_gemm_timeline

namespace {

enum { CLS_GEMM = 0, CLS_ATTN = 1, CLS_LN = 2, CLS_OTHER = 3 };

// RAII scope: records a CUDA event pair around the launches of one kernel class when profiling is on
struct ProfScope {
  regen_handle* h;
  cudaStream_t s;
  int cls;
  cudaEvent_t e1 = nullptr;
  ProfScope(regen_handle* h_, int cls_, cudaStream_t s_) : h(h_), s(s_), cls(cls_) {
    if (h->prof_on) {
      cudaEvent_t e0;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, s);
      h->prof_ev[cls].push_back({e0, e1});
    }
  }
  ~ProfScope() {
    if (e1) cudaEventRecord(e1, s);
  }
};

#define TRY(expr)                \
  do {                           \
    int _rc = (expr);            \
    if (_rc != REGEN_OK) return _rc; \
  } while (0)

int alloc_split(regen_handle* h, SplitBuf* s, size_t rows, size_t cols, uint32_t box_rows) {
  TRY(h->alloc(&s->hi, rows * cols));
  TRY(h->alloc(&s->lo, rows * cols));
  REGEN_CUDA(cudaMemset(s->hi, 0, rows * cols * sizeof(bf16)));
  REGEN_CUDA(cudaMemset(s->lo, 0, rows * cols * sizeof(bf16)));
  TRY(make_tmap_bf16_2d(&s->tm_hi, s->hi, rows, cols, cols, box_rows));
  TRY(make_tmap_bf16_2d(&s->tm_lo, s->lo, rows, cols, cols, box_rows));
  TRY(make_tmap_bf16_2d(&s->tm_hi2, s->hi, rows, cols, cols, 128));
  TRY(make_tmap_bf16_2d(&s->tm_lo2, s->lo, rows, cols, cols, 128));
  TRY(make_tmap_bf16_2d(&s->tm_hi3, s->hi, rows, cols, cols, 64));
  TRY(make_tmap_bf16_2d(&s->tm_lo3, s->lo, rows, cols, cols, 64));
  TRY(make_tmap_bf16_2d(&s->tm_hi4, s->hi, rows, cols, cols, 32));
  TRY(make_tmap_bf16_2d(&s->tm_lo4, s->lo, rows, cols, cols, 32));
  s->cols = cols;
  s->has64 = true;
  return REGEN_OK;
}

int alloc_m8w(regen_handle* h, M8W* w, size_t rows, size_t K) {
  TRY(h->alloc(&w->w16, rows * K));
  TRY(h->alloc(&w->w8, rows * 2 * K));
  REGEN_CUDA(cudaMemset(w->w16, 0, rows * K * 2));
  REGEN_CUDA(cudaMemset(w->w8, 0, rows * 2 * K));
  TRY(make_tmap_bf16_2d(&w->tm16, w->w16, rows, K, K, 128));
  TRY(make_tmap_bf16_2d(&w->tm16_64, w->w16, rows, K, K, 64));
  TRY(make_tmap_bf16_2d(&w->tm16_32, w->w16, rows, K, K, 32));
  TRY(make_tmap_u8_2d(&w->tm8, w->w8, rows, 2 * K, 2 * K, 128, 128));
  TRY(make_tmap_u8_2d(&w->tm8_64, w->w8, rows, 2 * K, 2 * K, 64, 128));
  TRY(make_tmap_u8_2d(&w->tm8_32, w->w8, rows, 2 * K, 2 * K, 32, 128));
  return REGEN_OK;
}

void pack_m8w(const float* src, M8W* w, int rows, int K, cudaStream_t s) {
  layers::pack_m8_kernel<<<grid_cap(ceil_div((int64_t)rows * K / 4, 256)), 256, 0, s>>>(src, w->w16, w->w8, rows, K, 0);
  count_launch();
}

int copy_vec(regen_handle* h, float** dst, const float* src, size_t n, cudaStream_t s) {
  TRY(h->alloc(dst, n));
  REGEN_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return REGEN_OK;
}

// CTA-pair (cta_group::2, 256-row tiles) kernel for anything larger than one 128-row tile; the single-CTA kernel
// serves tiny batches.  REGEN_DEBUG_GEMM_1CTA=1 forces the single-CTA kernel (A/B measurements).
bool use_pair_kernel(int M) {
  static int force1 = -1;
  if (force1 < 0) {
    const char* e = getenv("REGEN_DEBUG_GEMM_1CTA");
    force1 = (e && e[0] == '1') ? 1 : 0;
  }
  return !force1 && M > 128;
}

// REGEN_DEBUG_QKV_TIMELINE=1: the regen_test_gemm_timeline buffer receives the stamps of the (last) QKV GEMM of a forward
// instead of those of the fused GEMM+LN kernels (tools/qkv_timeline.py)
bool qkv_timeline() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("REGEN_DEBUG_QKV_TIMELINE");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

// o32: store-side maps {box 16, box 32} of the fp32 output named in p; osplit: the bf16-pair output buffer
// (null -> st.global epilogue)
// m8w != null: A (a.tm_hi = fp16 halves, *a8 = byte rows) and W (*m8w) are the mixed8 operand pack: pair kernel, mixed8 main loop
int run_gemm(regen_handle* h, const SplitBuf& a, const SplitBuf& w, gemm::Params p, const CUtensorMap* o32,
             const SplitBuf* osplit, cudaStream_t s, const CUtensorMap* a8 = nullptr, const M8W* m8w = nullptr) {
  ProfScope prof(h, CLS_GEMM, s);
  gemm::OutMaps om;
  p.tma_store = 0;
  p.exit_wait_full = h->exit_wait_full ? 1 : 0;
  p.pol_w = h->pol(4, ptx::kL2EvictLast);
  p.pol_store = h->pol(2, ptx::kL2EvictLast);
  if (h->tma_store && (p.N & 3) == 0 && (!p.out_f32 || o32) && (!p.out_hi || osplit)) {
    p.tma_store = 1;
    // the 16-epilogue-warp pair kernel (no residual, one kind of output) stores 32-column boxes, everything else 16
    const bool wide = use_pair_kernel(p.M) && !p.residual && !(p.out_f32 && p.out_hi);
    if (o32) om.f32 = o32[0];  // fp32 outputs always use 16-column boxes
    if (osplit) {
      om.hi = wide ? osplit->st32_hi : osplit->st_hi;
      om.lo = wide ? osplit->st32_lo : osplit->st_lo;
      // bf16-pair outputs only (QKV, FFN1): 64-column boxes (128-byte rows) through warp pairs, half the store row requests
      if (wide && h->store64 && !p.out_f32 && (p.N & 63) == 0) {
        p.pair64 = 1;
        om.hi = osplit->st64_hi;
        om.lo = osplit->st64_lo;
        if (p.m8) om.lo = h->st_ffn8;  // mixed8 operand bytes (the fp16 halves use the 16-bit 32 x 64 box map)
      }
    }
  }
  if (p.m8 && !p.pair64) {
    set_error("run_gemm: the mixed8 output format needs the 64-column store path");
    return REGEN_EINVAL;
  }
  if (h->steplog && h->steplog_slot < h->steplog_cap) {
    p.steplog = h->steplog;
    p.steplog_slot = h->steplog_slot++;
    p.steplog_cta = 2 * h->steplog_cap;
  }
  if (qkv_timeline() && p.N == 3 * D) p.timeline = g_test_timeline;  // bring-up: pipeline stamps of the QKV GEMM inside a real forward
  cudaError_t e;
  if (m8w) {
    if (!use_pair_kernel(p.M) || !p.tma_store) {
      set_error("run_gemm: the mixed8 main loop needs the pair kernel with TMA stores");
      return REGEN_EINVAL;
    }
    gemm::SliceMaps sm;
    if (h->narrow_slices) { sm.hi32 = &m8w->tm16_32; sm.lo32 = &m8w->tm8_32; sm.hi64 = &m8w->tm16_64; sm.lo64 = &m8w->tm8_64; }
    e = gemm::launch2_m8<256>(a.tm_hi, *a8, m8w->tm16, m8w->tm8, om, p, s, sm);
  } else if (use_pair_kernel(p.M)) {
    gemm::SliceMaps sm;
    if (h->narrow_slices) { sm.hi32 = &w.tm_hi4; sm.lo32 = &w.tm_lo4; sm.hi64 = &w.tm_hi3; sm.lo64 = &w.tm_lo3; }
    e = h->desc.precision != 1 ? gemm::launch2<256, true>(a.tm_hi, a.tm_lo, w.tm_hi2, w.tm_lo2, om, p, s, sm)
                               : gemm::launch2<256, false>(a.tm_hi, a.tm_lo, w.tm_hi2, w.tm_lo2, om, p, s, sm);
  }
  else {
    // at most 64 token rows (a single sample of T <= 64 frames): the GEMM is bound by the bytes one SM can pull per
    // k-block (measured ~50 B/clk), two thirds of which were the 128-row A box -- load the 64-row box instead
    const bool a64 = p.M <= 64 && a.has64;
    p.a_rows = a64 ? 64 : 0;
    const CUtensorMap& ah = a64 ? a.tm_hi3 : a.tm_hi;
    const CUtensorMap& al = a64 ? a.tm_lo3 : a.tm_lo;
    e = h->desc.precision != 1 ? gemm::launch<64, true>(ah, al, w.tm_hi3, w.tm_lo3, om, p, s)
                               : gemm::launch<64, false>(ah, al, w.tm_hi3, w.tm_lo3, om, p, s);
  }
  if (e != cudaSuccess) {
    set_error("gemm launch (M=%d N=%d K=%d) failed: %s", p.M, p.N, p.K, cudaGetErrorString(e));
    return REGEN_ECUDA;
  }
  count_launch();
  return REGEN_OK;
}

gemm::Params gp(int M, int N, int K) {
  gemm::Params p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  return p;
}

}  // namespace

extern "C" {

int regen_create(regen_handle** out, int32_t device, const regen_model_desc* d) {
  REGEN_CHECK_ARG(out && d, "regen_create: null argument");
  REGEN_CHECK_ARG(d->latent_dim == 512 && d->num_heads == 4 && d->ff_size == 1024,
                  "regen_create: kernels are specialised for latent_dim=512, num_heads=4, ff_size=1024 "
                  "(utils/model_util.py:69-70); got %d/%d/%d", d->latent_dim, d->num_heads, d->ff_size);
  REGEN_CHECK_ARG(d->num_layers >= 1 && d->num_layers <= REGEN_MAX_LAYERS, "regen_create: bad num_layers %d",
                  d->num_layers);
  REGEN_CHECK_ARG(d->input_feats >= 1 && d->input_feats <= 4096, "regen_create: bad input_feats %d", d->input_feats);
  REGEN_CHECK_ARG(d->cm_mode == 0 || d->cm_mode == 1, "regen_create: cm_mode must be 0 (add) or 1 (concat)");
  REGEN_CHECK_ARG(d->precision >= 0 && d->precision <= 3,
                  "regen_create: precision must be 0 (bf16x3), 1 (bf16), 2 (mixed8) or 3 (mixed8h)");
  REGEN_CHECK_ARG(d->arch == 0 || d->arch == 1, "regen_create: arch must be 0 ('online') or 1 ('offline')");
  REGEN_CHECK_ARG(d->max_batch >= 1 && d->max_frames >= 1 && d->num_table_steps >= 1, "regen_create: bad sizes");
  // more than 256 tokens per sample run the streaming CUDA-core attention (attn::attention_long_kernel); the positional
  // table bounds the length (checked against pe_len in regen_load_weights)
  REGEN_CHECK_ARG((int64_t)d->max_batch * d->max_frames < (1 << 24), "regen_create: max_batch*max_frames too large");
  {
    int ndev = 0;
    REGEN_CUDA(cudaGetDeviceCount(&ndev));
    REGEN_CHECK_ARG(device >= 0 && device < ndev, "regen_create: device %d out of range (%d devices)", device, ndev);
  }
  DeviceGuard guard(device);
  regen_handle* h = new regen_handle();
  h->desc = *d;
  h->device = device;
  h->L = d->num_layers;
  h->I = d->input_feats;
  h->Kin = (int)ceil_div(h->I, 64) * 64;
  h->offline = d->arch == 1;
  h->Mmax = d->max_batch * (d->max_frames + d->arch);
  {
    const char* e = getenv("REGEN_DEBUG_SIMT_ATTENTION");
    h->simt_attention = e && e[0] == '1' && layers::attention_smem_bytes(d->max_frames) <= 227 * 1024;
    const char* e2 = getenv("REGEN_DEBUG_NO_TMA_STORE");
    h->tma_store = !(e2 && e2[0] == '1');
    const char* e3 = getenv("REGEN_DEBUG_NO_FUSED_LN");
    h->fused_ln = h->tma_store && !(e3 && e3[0] == '1');
    const char* e4 = getenv("REGEN_DEBUG_NO_RES_PREFETCH");
    h->prefetch_res = !(e4 && e4[0] == '1');
    const char* e4b = getenv("REGEN_DEBUG_F32_RESIDUAL");
    h->res16 = !(e4b && e4b[0] == '1');
    const char* e4g = getenv("REGEN_ATTN_MC");
    if (e4g && (e4g[0] == '0' || e4g[0] == '1')) h->attn_mc_mode = e4g[0] - '0';
    const char* e4f = getenv("REGEN_DEBUG_WIDE_SLICES");
    h->narrow_slices = !(e4f && e4f[0] == '1');
    const char* e4e = getenv("REGEN_DEBUG_EXIT_WAIT_READ");
    h->exit_wait_full = !(e4e && e4e[0] == '1');
    const char* e4d = getenv("REGEN_DEBUG_LN_NOB");
    h->nob16 = (e4d && e4d[0] == '3') ? 3 : 2;
    const char* e4c = getenv("REGEN_DEBUG_NO_STORE64");
    h->store64 = !(e4c && e4c[0] == '1');
    const char* e5 = getenv("REGEN_L2_HINTS");
    if (e5) h->l2_hints = atoi(e5);
  }
  const size_t Mx = (size_t)h->Mmax;
  int rc = REGEN_OK;
  do {
    // packed weights
    if ((rc = alloc_split(h, &h->w_in, D, h->Kin, 256))) break;
    if ((rc = alloc_split(h, &h->w_out, h->I, D, 256))) break;
    for (int l = 0; l < h->L && rc == REGEN_OK; ++l) {
      LayerDev& ld = h->layer[l];
      if ((rc = alloc_split(h, &ld.wqkv, 3 * D, D, 256))) break;
      if ((rc = alloc_split(h, &ld.wo, D, D, 256))) break;
      if ((rc = alloc_split(h, &ld.w1, FF, D, 256))) break;
      if ((rc = alloc_split(h, &ld.w2, D, FF, 256))) break;
      if (d->precision >= 2) {
        if ((rc = h->alloc(&ld.w2_16, (size_t)D * FF))) break;
        if ((rc = h->alloc(&ld.w2_8, (size_t)D * 2 * FF))) break;
        if ((rc = make_tmap_bf16_2d(&ld.tm_w2_16, ld.w2_16, D, FF, FF, 128))) break;
        if ((rc = make_tmap_u8_2d(&ld.tm_w2_8, ld.w2_8, D, 2 * FF, 2 * FF, 128, 128))) break;
        if ((rc = h->alloc(&ld.wo_16, (size_t)D * D))) break;
        if ((rc = h->alloc(&ld.wo_8, (size_t)D * 2 * D))) break;
        if ((rc = make_tmap_bf16_2d(&ld.tm_wo_16, ld.wo_16, D, D, D, 128))) break;
        if ((rc = make_tmap_u8_2d(&ld.tm_wo_8, ld.wo_8, D, 2 * D, 2 * D, 128, 128))) break;
      }
      if (d->precision == 3) {
        if ((rc = alloc_m8w(h, &ld.m_qkv, 3 * D, D))) break;
        if ((rc = alloc_m8w(h, &ld.m_w1, FF, D))) break;
      }
    }
    if (rc) break;
    if ((rc = h->alloc(&h->w_c, (size_t)D * h->I))) break;
    if ((rc = h->alloc(&h->b_c, D))) break;
    if ((rc = h->alloc(&h->P, (size_t)h->L * D * D))) break;
    if ((rc = h->alloc(&h->qvec, (size_t)h->L * D))) break;
    if (!h->offline && (rc = h->alloc(&h->ctab, (size_t)d->num_table_steps * h->L * D))) break;
    if (h->offline && (rc = h->alloc(&h->e2tab, (size_t)d->num_table_steps * D))) break;
    // activations
    if ((rc = alloc_split(h, &h->a_in, Mx, h->Kin, 128))) break;
    if ((rc = alloc_split(h, &h->h_s, Mx, D, 128))) break;
    if ((rc = alloc_split(h, &h->att, Mx, D, 128))) break;
    if ((rc = alloc_split(h, &h->ffn, Mx, FF, 128))) break;
    if (d->precision >= 2 && (rc = make_tmap_u8_2d(&h->tm_ffn8, h->ffn.lo, Mx, 2 * FF, 2 * FF, 128, 128))) break;
    if (d->precision >= 2 && (rc = make_tmap_u8_2d(&h->tm_att8, h->att.lo, Mx, 2 * D, 2 * D, 128, 128))) break;
    if (d->precision == 3 && (rc = make_tmap_u8_2d(&h->tm_h8, h->h_s.lo, Mx, 2 * D, 2 * D, 128, 128))) break;
    if (d->precision == 3 && (rc = alloc_m8w(h, &h->m_out, h->I, D))) break;
    if ((rc = alloc_split(h, &h->qkv_s, Mx, 3 * D, 128))) break;
    if ((rc = h->alloc(&h->h, Mx * D))) break;
    if ((rc = h->alloc(&h->qkv, Mx * 3 * D))) break;
    if ((rc = h->alloc(&h->tmp, Mx * D))) break;
    if ((rc = h->alloc(&h->x0e, Mx * h->I))) break;
    if ((rc = h->alloc(&h->condbias, Mx * D))) break;
    if ((rc = h->alloc(&h->cmo_tbi, Mx * h->I))) break;
    if ((rc = h->alloc(&h->scratch, Mx * D))) break;
    if ((rc = h->alloc(&h->ccond, (size_t)d->max_batch * h->L * D))) break;
    if ((rc = h->alloc(&h->cond_emb, (size_t)d->max_batch * D))) break;
    if ((rc = h->alloc(&h->cyc, (size_t)h->L * (d->max_batch + 32) * D))) break;
    // the attribute is per function, not per handle: always allow the full 227 KB so that handles with
    // different max_frames can coexist in one process
    cudaError_t e = cudaFuncSetAttribute(layers::attention_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attention) failed: %s", cudaGetErrorString(e));
      rc = REGEN_ECUDA;
    }
  } while (0);
  if (rc != REGEN_OK) {
    regen_destroy(h);
    return rc;
  }
  *out = h;
  return REGEN_OK;
}

void regen_destroy(regen_handle* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

int regen_load_weights(regen_handle* h, const regen_weight_ptrs* w, void* stream) {
  REGEN_CHECK_ARG(h && w, "regen_load_weights: null argument");
  REGEN_CHECK_ARG(!h->loaded, "regen_load_weights: weights already loaded (create a new handle to reload)");
  DeviceGuard guard(h->device);
  REGEN_CHECK_ARG(w->in_w && w->in_b && w->cmo_w && w->cmo_b && w->t0_w && w->t0_b && w->t2_w && w->t2_b && w->pe &&
                      w->out_w && w->out_b, "regen_load_weights: null weight pointer");
  REGEN_CHECK_ARG(h->desc.cm_mode == 0 || (w->fuse_w && w->fuse_b), "regen_load_weights: cm_mode=concat needs fuse_process");
  REGEN_CHECK_ARG(w->pe_len >= h->desc.num_table_steps && w->pe_len >= h->desc.max_frames,
                  "regen_load_weights: positional table (%d rows) shorter than num_table_steps / max_frames", w->pe_len);
  cudaStream_t s = (cudaStream_t)stream;
  const int I = h->I, L = h->L;
  float* fold = nullptr;
  TRY(h->alloc(&fold, (size_t)D * (I > D ? I : D)));
  float* bvec = nullptr;
  TRY(h->alloc(&bvec, D));

  TRY(copy_vec(h, &h->pe, w->pe, (size_t)w->pe_len * D, s));
  h->pe_len = w->pe_len;
  TRY(copy_vec(h, &h->b_out, w->out_b, I, s));
  if (w->action_emb) {
    REGEN_CHECK_ARG(w->num_actions >= 1, "regen_load_weights: action_emb given but num_actions=%d", w->num_actions);
    TRY(copy_vec(h, &h->action_emb, w->action_emb, (size_t)w->num_actions * D, s));
    h->num_actions = w->num_actions;
  }
  if (w->text_w) {
    REGEN_CHECK_ARG(w->text_b && w->clip_dim >= 1, "regen_load_weights: embed_text needs bias and clip_dim");
    TRY(copy_vec(h, &h->text_w, w->text_w, (size_t)D * w->clip_dim, s));
    TRY(copy_vec(h, &h->text_b, w->text_b, D, s));
    h->clip_dim = w->clip_dim;
  }

  if (h->desc.cm_mode == 1) {
    // h = W_f [hx ; hc] + b_f with hx = W_in x + b_in, hc = W_cmo c + b_cmo  (model/cmdm.py:201-211)
    //   => h = (W_f1 W_in) x + (W_f2 W_cmo) c + (W_f1 b_in + W_f2 b_cmo + b_f)
    // W_in' = W_f1 . W_in : [D, I];  A = W_f1 (row stride 2D), B = W_in [D(k), I(n)]
    layers::launch_sgemm(w->fuse_w, 2 * D, 1, w->in_w, I, 1, nullptr, nullptr, fold, I, D, I, D, 0, s);
    layers::launch_split_rows(fold, I, h->w_in.hi, h->w_in.lo, h->Kin, I, D, 1, 1, s);
    // W_c = W_f2 . W_cmo
    layers::launch_sgemm(w->fuse_w + D, 2 * D, 1, w->cmo_w, I, 1, nullptr, nullptr, h->w_c, I, D, I, D, 0, s);
    // b_c = W_f1 b_in + (W_f2 b_cmo + b_f)
    layers::launch_sgemm(w->fuse_w + D, 2 * D, 1, w->cmo_b, 1, 1, nullptr, w->fuse_b, bvec, 1, D, 1, D, 0, s);
    layers::launch_sgemm(w->fuse_w, 2 * D, 1, w->in_b, 1, 1, nullptr, bvec, h->b_c, 1, D, 1, D, 0, s);
  } else {
    // 'add': h = W_in x + W_cmo c + (b_in + b_cmo)
    layers::launch_split_rows(w->in_w, I, h->w_in.hi, h->w_in.lo, h->Kin, I, D, 1, 1, s);
    REGEN_CUDA(cudaMemcpyAsync(h->w_c, w->cmo_w, (size_t)D * I * sizeof(float), cudaMemcpyDeviceToDevice, s));
    // b_c[m] = 1 * b_in[m] + b_cmo[m]   (K = 1 GEMM with A = b_in as a [D,1] matrix and B = [[1]])
    float one = 1.f;
    float* one_d = nullptr;
    TRY(h->alloc(&one_d, 1));
    REGEN_CUDA(cudaMemcpyAsync(one_d, &one, sizeof(float), cudaMemcpyHostToDevice, s));
    layers::launch_sgemm(w->in_b, 1, 1, one_d, 1, 1, nullptr, w->cmo_b, h->b_c, 1, D, 1, 1, 0, s);
  }
  layers::launch_split_rows(w->out_w, D, h->w_out.hi, h->w_out.lo, D, D, I, 1, 1, s);
  if (h->desc.precision == 3) pack_m8w(w->out_w, &h->m_out, I, D, s);

  for (int l = 0; l < L; ++l) {
    const regen_layer_weights& lw = w->layers[l];
    REGEN_CHECK_ARG(lw.qkv_w && lw.qkv_b && lw.o_w && lw.o_b && lw.l1_w && lw.l1_b && lw.l2_w && lw.l2_b && lw.n1_w &&
                        lw.n1_b && lw.n2_w && lw.n2_b,
                    "regen_load_weights: null pointer in layer %d", l);
    REGEN_CHECK_ARG(h->offline || (lw.xv_w && lw.xv_b && lw.xo_w && lw.xo_b && lw.n3_w && lw.n3_b),
                    "regen_load_weights: null cross-attention / norm3 pointer in decoder layer %d", l);
    LayerDev& ld = h->layer[l];
    layers::launch_split_rows(lw.qkv_w, D, ld.wqkv.hi, ld.wqkv.lo, D, D, 3 * D, 1, 1, s);
    layers::launch_split_rows(lw.o_w, D, ld.wo.hi, ld.wo.lo, D, D, D, 1, 1, s);
    layers::launch_split_rows(lw.l1_w, D, ld.w1.hi, ld.w1.lo, D, D, FF, 1, 1, s);
    layers::launch_split_rows(lw.l2_w, FF, ld.w2.hi, ld.w2.lo, FF, FF, D, 1, 1, s);
    if (h->desc.precision >= 2) {
      layers::pack_m8_kernel<<<grid_cap(ceil_div((int64_t)D * FF / 4, 256)), 256, 0, s>>>(lw.l2_w, ld.w2_16, ld.w2_8, D, FF, 0);
      layers::pack_m8_kernel<<<grid_cap(ceil_div((int64_t)D * D / 4, 256)), 256, 0, s>>>(lw.o_w, ld.wo_16, ld.wo_8, D, D, 0);
      count_launch();
    }
    if (h->desc.precision == 3) {
      pack_m8w(lw.qkv_w, &ld.m_qkv, 3 * D, D, s);
      pack_m8w(lw.l1_w, &ld.m_w1, FF, D, s);
    }
    TRY(copy_vec(h, &ld.bqkv, lw.qkv_b, 3 * D, s));
    TRY(copy_vec(h, &ld.bo, lw.o_b, D, s));
    TRY(copy_vec(h, &ld.b1, lw.l1_b, FF, s));
    TRY(copy_vec(h, &ld.b2, lw.l2_b, D, s));
    TRY(copy_vec(h, &ld.n1w, lw.n1_w, D, s));
    TRY(copy_vec(h, &ld.n1b, lw.n1_b, D, s));
    TRY(copy_vec(h, &ld.n2w, lw.n2_w, D, s));
    TRY(copy_vec(h, &ld.n2b, lw.n2_b, D, s));
    if (h->offline) continue;  // encoder layer: no cross-attention, no norm3
    TRY(copy_vec(h, &ld.n3w, lw.n3_w, D, s));
    TRY(copy_vec(h, &ld.n3b, lw.n3_b, D, s));
    // 1-token cross-attention: c = W_xo (W_xv e + b_xv) + b_xo = P_l e + q_l
    //   P_l = W_xo . W_xv : A = W_xo [D, D(k)], B = W_xv [D(k), D(n)]
    layers::launch_sgemm(lw.xo_w, D, 1, lw.xv_w, D, 1, nullptr, nullptr, h->P + (size_t)l * D * D, D, D, D, D, 0, s);
    //   q_l = W_xo b_xv + b_xo
    layers::launch_sgemm(lw.xo_w, D, 1, lw.xv_b, 1, 1, nullptr, lw.xo_b, h->qvec + (size_t)l * D, 1, D, 1, D, 0, s);
  }

  // timestep table: ctab[t] = P (W_t2 silu(W_t0 pe[t] + b_t0) + b_t2) + q   for t < num_table_steps
  // (model/cmdm.py:291-298 followed by the folded cross-attention of every layer)
  {
    const int nt = h->desc.num_table_steps;
    float *e1 = nullptr, *e2 = nullptr;
    TRY(h->alloc(&e1, (size_t)nt * D));
    TRY(h->alloc(&e2, (size_t)nt * D));
    // e1 = silu(pe[0:nt] . W_t0^T + b_t0):  B[k, n] = W_t0[n, k] -> sbk = 1, sbn = D
    layers::launch_sgemm(h->pe, D, 1, w->t0_w, 1, D, w->t0_b, nullptr, e1, D, nt, D, D, 1, s);
    layers::launch_sgemm(e1, D, 1, w->t2_w, 1, D, w->t2_b, nullptr, h->offline ? h->e2tab : e2, D, nt, D, D, 0, s);
    // ctab = e2 . P^T + q with P stacked [L*D, D]   (online; the offline model uses e2 itself as token 0)
    if (!h->offline)
      layers::launch_sgemm(e2, D, 1, h->P, 1, D, h->qvec, nullptr, h->ctab, (int64_t)L * D, nt, L * D, D, 0, s);
  }
  REGEN_LAUNCH_CHECK();
  h->loaded = true;
  return REGEN_OK;
}

int regen_prepare_cond(regen_handle* h, const float* cmotion_bjft, const int64_t* action, const float* text_feat,
                       int32_t B, int32_t T, int32_t guidance, int32_t uncond, void* stream) {
  REGEN_CHECK_ARG(h && cmotion_bjft, "regen_prepare_cond: null argument");
  DeviceGuard guard(h->device);
  if (!h->loaded) {
    set_error("regen_prepare_cond: weights not loaded");
    return REGEN_ESTATE;
  }
  const int Beff = guidance ? 2 * B : B;
  const int S = T + (h->offline ? 1 : 0);  // tokens per sample
  REGEN_CHECK_ARG(B >= 1 && T >= 1 && Beff <= h->desc.max_batch && T <= h->desc.max_frames &&
                      (int64_t)Beff * S <= h->Mmax,
                  "regen_prepare_cond: B=%d (effective %d) T=%d exceed the handle's max_batch=%d / max_frames=%d", B,
                  Beff, T, h->desc.max_batch, h->desc.max_frames);
  REGEN_CHECK_ARG(!action || h->action_emb, "regen_prepare_cond: action indices given but the model has no embed_action");
  REGEN_CHECK_ARG(!text_feat || h->text_w, "regen_prepare_cond: text features given but the model has no embed_text");
  cudaStream_t s = (cudaStream_t)stream;
  const int I = h->I, L = h->L;
  h->B = B; h->Beff = Beff; h->T = T; h->S = S; h->M = S * Beff;
  const int Mf = T * Beff;  // frame rows (the input / output projections never see the condition token)
  h->guidance = guidance != 0;
  // 'text' models keep a contribution even when unconditional: mask_cond zeroes the CLIP FEATURES, so
  // embed_text still adds its bias (model/cmdm.py:182-184); 'action' embeddings are zeroed themselves (:185-187)
  const bool text_model = h->text_w != nullptr;
  REGEN_CHECK_ARG(!text_model || uncond || text_feat, "regen_prepare_cond: text-conditioned model needs text features");
  h->has_cond = text_model || (action && !uncond);
  // cmotion [B,I,T] -> [T,B,I]; hc' = cmotion . W_c^T + b_c; condbias = dup(hc') + pe[t]
  TRY(regen_bjft_to_tbi(cmotion_bjft, h->cmo_tbi, B, I, T, stream));
  layers::launch_sgemm(h->cmo_tbi, I, 1, h->w_c, 1, I, h->b_c, nullptr, h->scratch, D, T * B, D, I, 0, s);
  {
    int64_t total = (int64_t)Mf * (D / 4);
    // offline: frame f is token f + 1 and gets pe[f + 1] (model/cmdm.py:234-236)
    layers::finalize_condbias_kernel<<<grid_cap(ceil_div(total, 256)), 256, 0, s>>>(
        h->scratch, h->pe + (h->offline ? D : 0), h->condbias, T, B, guidance ? 2 : 1);
    count_launch();
  }
  // store-side maps clip at the logical extents, so they are rebuilt for the current M = T * Beff
  for (SplitBuf* sb : {&h->h_s, &h->ffn, &h->qkv_s}) {
    TRY(make_tmap_store_2d(&sb->st_hi, sb->hi, true, h->M, sb->cols, sb->cols));
    TRY(make_tmap_store_2d(&sb->st_lo, sb->lo, true, h->M, sb->cols, sb->cols));
    TRY(make_tmap_store_2d(&sb->st32_hi, sb->hi, true, h->M, sb->cols, sb->cols, 32));
    TRY(make_tmap_store_2d(&sb->st32_lo, sb->lo, true, h->M, sb->cols, sb->cols, 32));
    TRY(make_tmap_store_2d(&sb->st64_hi, sb->hi, true, h->M, sb->cols, sb->cols, 64));
    TRY(make_tmap_store_2d(&sb->st64_lo, sb->lo, true, h->M, sb->cols, sb->cols, 64));
  }
  if (h->desc.precision >= 2) TRY(make_tmap_u8_2d(&h->st_ffn8, h->ffn.lo, h->M, 2 * FF, 2 * FF, 32, 128));
  TRY(make_tmap_store_2d(&h->st32_h, h->h, false, h->M, D, D, 32));
  TRY(make_tmap_store_2d(&h->ld32_condbias, h->condbias, false, Mf, D, D, 32));
  if (h->offline) {
    // views without the Beff condition-token rows: outputs of the input projection, A operand of the output projection
    const size_t off = (size_t)Beff * D;
    SplitBuf& v = h->h_fr;
    v.hi = h->h_s.hi + off; v.lo = h->h_s.lo + off; v.cols = D;
    TRY(make_tmap_bf16_2d(&v.tm_hi, v.hi, Mf, D, D, 128));
    TRY(make_tmap_bf16_2d(&v.tm_lo, v.lo, Mf, D, D, 128));
    TRY(make_tmap_store_2d(&v.st_hi, v.hi, true, Mf, D, D));
    TRY(make_tmap_store_2d(&v.st_lo, v.lo, true, Mf, D, D));
    TRY(make_tmap_store_2d(&v.st32_hi, v.hi, true, Mf, D, D, 32));
    TRY(make_tmap_store_2d(&v.st32_lo, v.lo, true, Mf, D, D, 32));
    TRY(make_tmap_store_2d(&v.st64_hi, v.hi, true, Mf, D, D, 64));
    TRY(make_tmap_store_2d(&v.st64_lo, v.lo, true, Mf, D, D, 64));
    TRY(make_tmap_store_2d(&h->st_h_fr, h->h + off, false, Mf, D, D));
    TRY(make_tmap_store_2d(&h->st32_h_fr, h->h + off, false, Mf, D, D, 32));
  }
  for (int l = 0; l < L && !h->offline; ++l)
    TRY(make_tmap_store_2d(&h->tm_cyc[l], h->cyc + (size_t)l * (Beff + 32) * D, false, Beff + 32, D, D, 32));
  TRY(make_tmap_store_2d(&h->st_h, h->h, false, h->M, D, D));
  TRY(make_tmap_store_2d(&h->st_tmp, h->tmp, false, h->M, D, D));
  if ((I & 3) == 0) TRY(make_tmap_store_2d(&h->st_x0e, h->x0e, false, Mf, I, I));
  // 64-frame boxes for the compact (S <= 64) and the multi-chunk compact (64 < S <= 256) attention kernels
  // Measured (config 3, S = 150: attention 1.54 -> 1.11 ms per step; config 5, S = 196: 0.84 -> 0.86): the multi-chunk
  // kernel wins when the 128-query-block kernel would run a mostly empty last block, and ties otherwise.
  h->attn_mc = S > 64 && S <= 256 &&
               (h->attn_mc_mode == 1 || (h->attn_mc_mode == 2 && ceil_div(S, 64) < 2 * ceil_div(S, 128)));
  const uint32_t qkv_box = (S <= 64 || h->attn_mc) ? 64 : 128;
  TRY(make_tmap_bf16_3d(&h->tm_qkv_hi, h->qkv_s.hi, 3 * D, Beff, S, qkv_box));
  TRY(make_tmap_bf16_3d(&h->tm_qkv_lo, h->qkv_s.lo, 3 * D, Beff, S, qkv_box));
  TRY(make_tmap_bf16_3d(&h->tm_att_hi, h->att.hi, D, Beff, S, 32));
  TRY(make_tmap_bf16_3d(&h->tm_att_lo, h->att.lo, D, Beff, S, 32));
  if (h->desc.precision == 3) TRY(make_tmap_u8_2d(&h->st_h8, h->h_s.lo, h->M, 2 * D, 2 * D, 32, 128));
  if (h->desc.precision >= 2) TRY(make_tmap_u8_3d(&h->st_att8, h->att.lo, 2 * D, Beff, S, 32, 128));
  if (h->has_cond) {
    // cond_emb[b'] for b' in [0, Beff): conditional rows [0,B), unconditional rows [B,2B) under guidance
    if (text_model) {
      layers::broadcast_rows_kernel<<<Beff, 128, 0, s>>>(h->text_b, h->cond_emb);
      count_launch();
      if (!uncond)
        layers::launch_sgemm(text_feat, h->clip_dim, 1, h->text_w, 1, h->clip_dim, h->text_b, nullptr, h->cond_emb, D,
                             B, D, h->clip_dim, 0, s);
    } else {
      REGEN_CUDA(cudaMemsetAsync(h->cond_emb, 0, (size_t)Beff * D * sizeof(float), s));
    }
    if (action && !uncond) {
      layers::gather_rows_kernel<<<B, 128, 0, s>>>(h->action_emb, action, h->cond_emb, h->num_actions, 1);
      count_launch();
    }
    // ccond[b'] = P . cond_emb[b']   (folded 1-token cross-attention of every layer; offline: cond_emb joins token 0)
    if (!h->offline)
      layers::launch_sgemm(h->cond_emb, D, 1, h->P, 1, D, nullptr, nullptr, h->ccond, (int64_t)L * D, Beff, L * D, D, 0,
                           s);
  }
  REGEN_LAUNCH_CHECK();
  h->cond_ready = true;
  return REGEN_OK;
}

int regen_denoise(regen_handle* h, const float* x_tbi, const int64_t* t, const float* cfg_scale, float* x0_tbi,
                  int32_t B, int32_t T, void* stream) {
  REGEN_CHECK_ARG(h && x_tbi && t && x0_tbi, "regen_denoise: null argument");
  DeviceGuard guard(h->device);
  if (!h->cond_ready) {
    set_error("regen_denoise: regen_prepare_cond has not been called");
    return REGEN_ESTATE;
  }
  REGEN_CHECK_ARG(B == h->B && T == h->T, "regen_denoise: B=%d T=%d differ from prepare_cond's B=%d T=%d", B, T, h->B,
                  h->T);
  REGEN_CHECK_ARG(!h->guidance || cfg_scale, "regen_denoise: guidance was requested but cfg_scale is null");
  cudaStream_t s = (cudaStream_t)stream;
  const int M = h->M, I = h->I, L = h->L, Beff = h->Beff, S = h->S;
  const bool offline = h->offline;
  // The fused GEMM+LayerNorm kernel gives one 256-row tile to a CTA pair, so it uses 2 * ceil(M / 256) SMs.  Below half
  // of the machine the N-tiled GEMM + a separate warp-per-row LayerNorm kernel is faster (measured crossover at M ~ 9 500:
  // B = 32, T = 60 runs 0.705 instead of 1.022 ms per step; single samples use the 64-column single-CTA GEMM tiles).
  static int force_fused = -1;  // REGEN_DEBUG_FORCE_FUSED=1: fused route for any M > 128 (feed-bandwidth experiments)
  if (force_fused < 0) {
    const char* e = getenv("REGEN_DEBUG_FORCE_FUSED");
    force_fused = (e && e[0] == '1') ? 1 : 0;
  }
  const bool fused = h->fused_ln && (ceil_div(M, 256) >= kNumSMs / 4 || (force_fused && M > 128));
  // precision 'mixed8': on the fused route linear2 runs as one fp16 MMA + two e4m3 correction MMAs per product (2 instead
  // of 3 bf16-MMA equivalents); FFN1's epilogue writes its activations in that operand format.  Everything else, and
  // the whole small-batch route, is bf16x3.
  const bool m8 = h->desc.precision >= 2 && fused && h->tma_store && h->store64 && h->res16;
  // ... and the attention output projection the same way when a tcgen05 attention kernel (which can write that format) runs
  const bool m8_att = m8 && !h->simt_attention && S <= 256;
  const int Mf = T * Beff;                              // frame rows
  // precision 'mixed8h' (arch 'online'): the residual stream h is the mixed8 pack as well, so the input projection writes it,
  // both fused kernels load / store it, and QKV / FFN1 / the output projection run the mixed8 main loop on it
  const bool m8h = m8_att && h->desc.precision == 3 && !offline && Mf > 128 && (I & 3) == 0;
  const size_t fr_off = offline ? (size_t)Beff * D : 0;  // offline: the frames follow the Beff condition-token rows

  // A operand of the input projection: split x into bf16 (hi, lo), K padded to a multiple of 64,
  // batch duplicated under guidance
  {
    ProfScope prof(h, CLS_OTHER, s);
    layers::launch_split_rows(x_tbi, I, h->a_in.hi, h->a_in.lo, h->Kin, I, Mf, B, h->guidance ? 2 : 1, s);
    if (offline) {  // token 0 = timestep embedding + condition embedding + pe[0]   (model/cmdm.py:234-235)
      launch_pdl(layers::build_emb_rows_kernel, dim3(Beff), dim3(128), 0, s, (const float*)h->e2tab,
                 h->has_cond ? (const float*)h->cond_emb : (const float*)nullptr, (const float*)h->pe, t, h->h, h->h_s.hi,
                 h->h_s.lo, B, h->desc.num_table_steps);
      count_launch();
    }
  }

  // h = x . W_in'^T + condbias      (input_process + fuse_process + positional encoding, hoisted parts in condbias)
  {
    gemm::Params p = gp(Mf, D, h->Kin);
    p.residual = h->condbias; p.ld_res = D;
    p.out_f32 = h->h + fr_off; p.ld_out = D;
    p.out_hi = h->h_s.hi + fr_off; p.out_lo = h->h_s.lo + fr_off; p.ld_split = D;
    const CUtensorMap hmaps[2] = {offline ? h->st_h_fr : h->st_h, offline ? h->st32_h_fr : h->st32_h};
    if (fused && Mf > 128) {
      // full-row (256 x 512) tiles through the fused kernel's machinery without LayerNorm: the residual arrives by
      // TMA and the outputs leave in one sweep (34.6 -> ~24 us at config 2; the generic residual epilogue is
      // bound by LSU round trips)
      ProfScope prof(h, CLS_GEMM, s);
      gemmln::Params q{};
      memset(&q, 0, sizeof(q));
      q.M = Mf; q.K = h->Kin; q.Beff = Beff; q.ln_eps = layers::LN_EPS;
      q.prefetch_res = h->prefetch_res ? 1 : 0;
      q.pol_a = h->pol(8, ptx::kL2EvictFirst); q.pol_w = h->pol(4, ptx::kL2EvictLast); q.pol_store = h->pol(2, ptx::kL2EvictLast);
      q.store_f32 = h->res16 ? 0 : 1;  // the fused consumers rebuild the residual from (hi, lo)
      q.exit_wait_full = h->exit_wait_full ? 1 : 0;
      if (h->steplog && h->steplog_slot < h->steplog_cap) {
        q.steplog = h->steplog; q.steplog_slot = h->steplog_slot++; q.steplog_cta = 2 * h->steplog_cap;
      }
      const SplitBuf& ob = offline ? h->h_fr : h->h_s;
      cudaError_t e = m8h ? gemmln::launch_noln_h8(h->a_in.tm_hi, h->a_in.tm_lo, h->w_in.tm_hi2, h->w_in.tm_lo2,
                                                     h->ld32_condbias, hmaps[1], ob.st64_hi, h->st_h8, q, s)
          : h->desc.precision != 1
          ? gemmln::launch_noln<true>(h->a_in.tm_hi, h->a_in.tm_lo, h->w_in.tm_hi2, h->w_in.tm_lo2, h->ld32_condbias,
                                      hmaps[1], ob.st64_hi, ob.st64_lo, q, s)
          : gemmln::launch_noln<false>(h->a_in.tm_hi, h->a_in.tm_lo, h->w_in.tm_hi2, h->w_in.tm_lo2, h->ld32_condbias,
                                       hmaps[1], ob.st64_hi, ob.st64_lo, q, s);
      if (e != cudaSuccess) {
        set_error("input projection launch failed: %s", cudaGetErrorString(e));
        return REGEN_ECUDA;
      }
      count_launch();
    } else {
      TRY(run_gemm(h, h->a_in, h->w_in, p, hmaps, offline ? &h->h_fr : &h->h_s, s));
    }
  }
  if (fused && !offline) {
    ProfScope prof(h, CLS_OTHER, s);
    launch_pdl(layers::build_cyc_kernel, dim3(Beff + 32, L), dim3(128), 0, s, h->ctab,
               h->has_cond ? (const float*)h->ccond : (const float*)nullptr, t, h->cyc, L, B, Beff,
               h->desc.num_table_steps);
    count_launch();
  }
  for (int l = 0; l < L; ++l) {
    LayerDev& ld = h->layer[l];
    {  // q | k | v  -> bf16 (hi, lo) operands of the attention kernel
      gemm::Params p = gp(M, 3 * D, D);
      p.bias = ld.bqkv;
      p.out_hi = h->qkv_s.hi; p.out_lo = h->qkv_s.lo; p.ld_split = 3 * D;
      if (h->simt_attention) { p.out_f32 = h->qkv; p.ld_out = 3 * D; }
      TRY(run_gemm(h, h->h_s, ld.wqkv, p, nullptr, h->simt_attention ? nullptr : &h->qkv_s, s, m8h ? &h->tm_h8 : nullptr,
                   m8h ? &ld.m_qkv : nullptr));
    }
    {
      ProfScope prof(h, CLS_ATTN, s);
      if (h->simt_attention) {
        layers::attention_simt_kernel<<<Beff * layers::H, layers::ATT_WARPS * 32, layers::attention_smem_bytes(T), s>>>(
            h->qkv, h->att.hi, h->att.lo, T, Beff);
      } else {
        attn::Params ap;
        ap.out_hi = h->att.hi; ap.out_lo = h->att.lo; ap.T = S; ap.Beff = Beff; ap.causal = offline ? 0 : 1; ap.dbg = h->exit_wait_full ? 4 : 0;
        ap.m8 = m8_att ? 1 : 0;
        const CUtensorMap& att_lo_map = m8_att ? h->st_att8 : h->tm_att_lo;
        ap.timeline = nullptr;
        ap.steplog = nullptr; ap.steplog_slot = 0; ap.steplog_cta = 0;
        ap.pol_load = h->pol(1, ptx::kL2EvictFirst);
        ap.pol_store = h->pol(2, ptx::kL2EvictLast);
        if (h->steplog && h->steplog_slot < h->steplog_cap) { ap.steplog = h->steplog; ap.steplog_slot = h->steplog_slot++; }
        cudaError_t e = S > 256 ? attn::launch_long(h->qkv_s.hi, h->qkv_s.lo, ap, s)
                        : S <= 64 ? attn::launch<64>(h->tm_qkv_hi, h->tm_qkv_lo, h->tm_att_hi, att_lo_map, ap, s)
                        : h->attn_mc
                                ? attn::launch_mc(h->tm_qkv_hi, h->tm_qkv_lo, h->tm_att_hi, att_lo_map, ap, s)
                                : attn::launch<128>(h->tm_qkv_hi, h->tm_qkv_lo, h->tm_att_hi, att_lo_map, ap, s);
        if (e != cudaSuccess) {
          set_error("attention launch (T=%d Beff=%d) failed: %s", T, Beff, cudaGetErrorString(e));
          return REGEN_ECUDA;
        }
      }
      count_launch();
    }
    if (offline && fused) {  // encoder layer: h = LN1(h + attn . W_o^T + b_o)   (nn.TransformerEncoderLayer, post-norm)
      ProfScope prof(h, CLS_GEMM, s);
      gemmln::Params q{};
      q.M = M; q.K = D; q.Beff = Beff; q.bias = ld.bo; q.g1 = ld.n1w; q.b1 = ld.n1b; q.g2 = nullptr; q.b2 = nullptr;
      q.ln_eps = layers::LN_EPS; q.store_f32 = 0; q.nob16 = h->nob16; q.exit_wait_full = h->exit_wait_full ? 1 : 0;
      q.prefetch_res = h->prefetch_res ? 1 : 0;
      q.pol_a = h->pol(8, ptx::kL2EvictFirst); q.pol_w = h->pol(4, ptx::kL2EvictLast); q.pol_store = h->pol(2, ptx::kL2EvictLast);
      q.timeline = qkv_timeline() ? nullptr : g_test_timeline;
      q.steplog = nullptr; q.steplog_slot = 0; q.steplog_cta = 0;
      if (h->steplog && h->steplog_slot < h->steplog_cap) {
        q.steplog = h->steplog; q.steplog_slot = h->steplog_slot++; q.steplog_cta = 2 * h->steplog_cap;
      }
      cudaError_t e = m8_att ? gemmln::launch_m8<false>(h->att.tm_hi, h->tm_att8, ld.tm_wo_16, ld.tm_wo_8, h->st32_h,
                                                        h->h_s.st64_hi, h->h_s.st64_lo, q, s)
          : h->desc.precision != 1
          ? gemmln::launch<true, false>(h->att.tm_hi, h->att.tm_lo, ld.wo.tm_hi2, ld.wo.tm_lo2, h->st32_h, h->st32_h,
                                        h->h_s.st64_hi, h->h_s.st64_lo, q, s, h->res16)
          : gemmln::launch<false, false>(h->att.tm_hi, h->att.tm_lo, ld.wo.tm_hi2, ld.wo.tm_lo2, h->st32_h, h->st32_h,
                                         h->h_s.st64_hi, h->h_s.st64_lo, q, s, h->res16);
      if (e != cudaSuccess) {
        set_error("fused out_proj+LayerNorm launch failed: %s", cudaGetErrorString(e));
        return REGEN_ECUDA;
      }
      count_launch();
    } else if (fused) {  // h = LN2( LN1(h + attn . W_o^T + b_o) + c_l[b] )   -- one kernel
      ProfScope prof(h, CLS_GEMM, s);
      gemmln::Params q{};
      q.M = M; q.K = D; q.Beff = Beff; q.bias = ld.bo; q.g1 = ld.n1w; q.b1 = ld.n1b; q.g2 = ld.n2w; q.b2 = ld.n2b;
      q.ln_eps = layers::LN_EPS; q.store_f32 = 0; q.nob16 = h->nob16; q.exit_wait_full = h->exit_wait_full ? 1 : 0;
      q.prefetch_res = h->prefetch_res ? 1 : 0;
      q.pol_a = h->pol(8, ptx::kL2EvictFirst); q.pol_w = h->pol(4, ptx::kL2EvictLast); q.pol_store = h->pol(2, ptx::kL2EvictLast);
      q.timeline = qkv_timeline() ? nullptr : g_test_timeline;
      q.steplog = nullptr; q.steplog_slot = 0; q.steplog_cta = 0;
      if (h->steplog && h->steplog_slot < h->steplog_cap) {
        q.steplog = h->steplog; q.steplog_slot = h->steplog_slot++; q.steplog_cta = 2 * h->steplog_cap;
      }
      cudaError_t e = m8_att ? gemmln::launch_m8<true>(h->att.tm_hi, h->tm_att8, ld.tm_wo_16, ld.tm_wo_8, h->tm_cyc[l],
                                                       h->h_s.st64_hi, m8h ? h->st_h8 : h->h_s.st64_lo, q, s, m8h)
          : h->desc.precision != 1
          ? gemmln::launch<true, true>(h->att.tm_hi, h->att.tm_lo, ld.wo.tm_hi2, ld.wo.tm_lo2, h->st32_h, h->tm_cyc[l],
                                       h->h_s.st64_hi, h->h_s.st64_lo, q, s, h->res16)
          : gemmln::launch<false, true>(h->att.tm_hi, h->att.tm_lo, ld.wo.tm_hi2, ld.wo.tm_lo2, h->st32_h, h->tm_cyc[l],
                                        h->h_s.st64_hi, h->h_s.st64_lo, q, s, h->res16);
      if (e != cudaSuccess) {
        set_error("fused out_proj+LayerNorm launch failed: %s", cudaGetErrorString(e));
        return REGEN_ECUDA;
      }
      count_launch();
    } else {
    {  // tmp = h + attn . W_o^T + b_o
      gemm::Params p = gp(M, D, D);
      p.bias = ld.bo;
      p.residual = h->h; p.ld_res = D;
      p.out_f32 = h->tmp; p.ld_out = D;
      const CUtensorMap tmaps[2] = {h->st_tmp, h->st_tmp};
      TRY(run_gemm(h, h->att, ld.wo, p, tmaps, nullptr, s));
    }
    if (offline) {  // h = LN1(tmp)   (encoder layer)
      layers::LnParams q;
      memset(&q, 0, sizeof(q));
      q.in = h->tmp; q.g1 = ld.n1w; q.b1 = ld.n1b;
      q.out_f32 = h->h; q.out_hi = h->h_s.hi; q.out_lo = h->h_s.lo; q.M = M;
      q.B = B; q.Beff = Beff;
      ProfScope prof(h, CLS_LN, s);
      launch_pdl(layers::layernorm_kernel<false>, dim3((unsigned)ceil_div(M, 8)), dim3(256), 0, s, q);
      count_launch();
    } else {  // h = LN2( LN1(tmp) + c_l[b] )
      layers::LnParams q;
      q.in = h->tmp; q.g1 = ld.n1w; q.b1 = ld.n1b; q.g2 = ld.n2w; q.b2 = ld.n2b;
      q.ctab = h->ctab + (size_t)l * D;
      q.ccond = h->has_cond ? h->ccond + (size_t)l * D : nullptr;
      q.t = t; q.ld_c = L * D; q.B = B; q.Beff = Beff; q.n_table = h->desc.num_table_steps;
      q.out_f32 = h->h; q.out_hi = h->h_s.hi; q.out_lo = h->h_s.lo; q.M = M;
      ProfScope prof(h, CLS_LN, s);
      launch_pdl(layers::layernorm_kernel<true>, dim3((unsigned)ceil_div(M, 8)), dim3(256), 0, s, q);
      count_launch();
    }
    }
    {  // ffn = gelu(h . W_1^T + b_1)
      gemm::Params p = gp(M, FF, D);
      p.bias = ld.b1; p.gelu = 1; p.m8 = m8 ? 1 : 0;
      p.out_hi = h->ffn.hi; p.out_lo = h->ffn.lo; p.ld_split = FF;
      TRY(run_gemm(h, h->h_s, ld.w1, p, nullptr, &h->ffn, s, m8h ? &h->tm_h8 : nullptr, m8h ? &ld.m_w1 : nullptr));
    }
    if (fused) {  // h = LN3(h + ffn . W_2^T + b_2)   -- one kernel (encoder layer: norm2)
      ProfScope prof(h, CLS_GEMM, s);
      gemmln::Params q{};
      q.M = M; q.K = FF; q.Beff = Beff; q.bias = ld.b2; q.g1 = offline ? ld.n2w : ld.n3w; q.b1 = offline ? ld.n2b : ld.n3b;
      q.g2 = nullptr; q.b2 = nullptr;
      q.ln_eps = layers::LN_EPS; q.store_f32 = 0; q.nob16 = h->nob16; q.exit_wait_full = h->exit_wait_full ? 1 : 0;
      q.prefetch_res = h->prefetch_res ? 1 : 0;
      q.pol_a = h->pol(8, ptx::kL2EvictFirst); q.pol_w = h->pol(4, ptx::kL2EvictLast); q.pol_store = h->pol(2, ptx::kL2EvictLast);
      q.timeline = qkv_timeline() ? nullptr : g_test_timeline;
      q.steplog = nullptr; q.steplog_slot = 0; q.steplog_cta = 0;
      if (h->steplog && h->steplog_slot < h->steplog_cap) {
        q.steplog = h->steplog; q.steplog_slot = h->steplog_slot++; q.steplog_cta = 2 * h->steplog_cap;
      }
      cudaError_t e = m8 ? gemmln::launch_m8<false>(h->ffn.tm_hi, h->tm_ffn8, ld.tm_w2_16, ld.tm_w2_8, h->st32_h,
                                                    h->h_s.st64_hi, m8h ? h->st_h8 : h->h_s.st64_lo, q, s, m8h)
          : h->desc.precision != 1
          ? gemmln::launch<true, false>(h->ffn.tm_hi, h->ffn.tm_lo, ld.w2.tm_hi2, ld.w2.tm_lo2, h->st32_h, h->st32_h,
                                        h->h_s.st64_hi, h->h_s.st64_lo, q, s, h->res16)
          : gemmln::launch<false, false>(h->ffn.tm_hi, h->ffn.tm_lo, ld.w2.tm_hi2, ld.w2.tm_lo2, h->st32_h, h->st32_h,
                                         h->h_s.st64_hi, h->h_s.st64_lo, q, s, h->res16);
      if (e != cudaSuccess) {
        set_error("fused linear2+LayerNorm launch failed: %s", cudaGetErrorString(e));
        return REGEN_ECUDA;
      }
      count_launch();
    } else {
    {  // tmp = h + ffn . W_2^T + b_2
      gemm::Params p = gp(M, D, FF);
      p.bias = ld.b2;
      p.residual = h->h; p.ld_res = D;
      p.out_f32 = h->tmp; p.ld_out = D;
      const CUtensorMap tmaps[2] = {h->st_tmp, h->st_tmp};
      TRY(run_gemm(h, h->ffn, ld.w2, p, tmaps, nullptr, s));
    }
    {  // h = LN3(tmp)
      layers::LnParams q;
      memset(&q, 0, sizeof(q));
      q.in = h->tmp; q.g1 = offline ? ld.n2w : ld.n3w; q.b1 = offline ? ld.n2b : ld.n3b;
      q.out_f32 = h->h; q.out_hi = h->h_s.hi; q.out_lo = h->h_s.lo; q.M = M;
      q.B = B; q.Beff = Beff;
      ProfScope prof(h, CLS_LN, s);
      launch_pdl(layers::layernorm_kernel<false>, dim3((unsigned)ceil_div(M, 8)), dim3(256), 0, s, q);
      count_launch();
    }
    }
  }
  {  // x0 = h . W_out^T + b_out   (output_process; rows already in [T,B,I] order)
    gemm::Params p = gp(Mf, I, D);
    p.bias = h->b_out;
    p.out_f32 = h->guidance ? h->x0e : x0_tbi; p.ld_out = I;
    CUtensorMap st_x0[2];
    const CUtensorMap* om = nullptr;
    if ((I & 3) == 0 && h->tma_store) {
      // store maps (box 16 and box 32) of the output buffer; a caller-owned output is encoded per call (~1 us each)
      float* dst = h->guidance ? h->x0e : x0_tbi;
      TRY(make_tmap_store_2d(&st_x0[0], dst, false, Mf, I, I, 16));
      TRY(make_tmap_store_2d(&st_x0[1], dst, false, Mf, I, I, 32));
      om = st_x0;
    }
    TRY(run_gemm(h, offline ? h->h_fr : h->h_s, h->w_out, p, om, nullptr, s, m8h ? &h->tm_h8 : nullptr,
                 m8h ? &h->m_out : nullptr));
  }
  if (h->guidance) {
    int64_t total = (int64_t)T * B * I;
    int blocks = grid_cap(ceil_div(total, 256));
    ProfScope prof(h, CLS_OTHER, s);
    launch_pdl(layers::cfg_rows_kernel, dim3(blocks), dim3(256), 0, s, h->x0e, cfg_scale, x0_tbi, T, B, I);
    count_launch();
  }
  REGEN_LAUNCH_CHECK();
  return REGEN_OK;
}

int regen_profile_begin(regen_handle* h) {
  REGEN_CHECK_ARG(h, "regen_profile_begin: null handle");
  for (auto& v : h->prof_ev) {
    for (auto& p : v) {
      cudaEventDestroy(p.first);
      cudaEventDestroy(p.second);
    }
    v.clear();
  }
  h->prof_on = true;
  return REGEN_OK;
}

int regen_profile_end(regen_handle* h, float* ms, int32_t* launches) {
  REGEN_CHECK_ARG(h && ms && launches, "regen_profile_end: null argument");
  h->prof_on = false;
  REGEN_CUDA(cudaDeviceSynchronize());
  for (int c = 0; c < 4; ++c) {
    double tot = 0.0;
    for (auto& p : h->prof_ev[c]) {
      float t = 0.f;
      REGEN_CUDA(cudaEventElapsedTime(&t, p.first, p.second));
      tot += t;
      cudaEventDestroy(p.first);
      cudaEventDestroy(p.second);
    }
    ms[c] = (float)tot;
    launches[c] = (int32_t)h->prof_ev[c].size();
    h->prof_ev[c].clear();
  }
  return REGEN_OK;
}

// Kernel-level test hook: C = A . W^T (+bias)(+residual)(gelu) through the tcgen05 GEMM, fp32 in / out.

int regen_test_step_log(regen_handle* h, unsigned long long* device_buf, int32_t capacity_slots) {
  REGEN_CHECK_ARG(h && capacity_slots >= 0, "regen_test_step_log: bad argument");
  h->steplog = capacity_slots > 0 ? device_buf : nullptr;
  h->steplog_cap = capacity_slots;
  h->steplog_slot = 0;
  return REGEN_OK;
}

int regen_test_gemm_timeline(unsigned long long* device_buf128) {
  g_test_timeline = device_buf128;
  return REGEN_OK;
}

int regen_test_gemm(const float* A, const float* W, const float* bias, const float* residual, float* out, int32_t M,
                    int32_t N, int32_t K, int32_t gelu, int32_t precision, void* stream) {
  REGEN_CHECK_ARG(A && W && out, "regen_test_gemm: null argument");
  REGEN_CHECK_ARG(M >= 1 && N >= 1 && K >= 1, "regen_test_gemm: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  if (precision == 2) {
    // mixed8 main loop of the pair kernel (fp16 + e4m3 operand pack, packed here from the fp32 inputs)
    REGEN_CHECK_ARG(use_pair_kernel(M) && (K & 127) == 0 && (N & 3) == 0 && !residual,
                    "regen_test_gemm: precision 2 needs M > 128, K %% 128 == 0, N %% 4 == 0 and no residual");
    uint16_t *a16, *w16;
    uint8_t *a8, *w8;
    REGEN_CUDA(cudaMalloc(&a16, (size_t)M * K * 2));
    REGEN_CUDA(cudaMalloc(&a8, (size_t)M * K * 2));
    REGEN_CUDA(cudaMalloc(&w16, (size_t)N * K * 2));
    REGEN_CUDA(cudaMalloc(&w8, (size_t)N * K * 2));
    layers::pack_m8_kernel<<<grid_cap(ceil_div((int64_t)M * K / 4, 256)), 256, 0, s>>>(A, a16, a8, M, K, 1);
    layers::pack_m8_kernel<<<grid_cap(ceil_div((int64_t)N * K / 4, 256)), 256, 0, s>>>(W, w16, w8, N, K, 0);
    CUtensorMap ta16, ta8, tw16, tw8;
    int rc = make_tmap_bf16_2d(&ta16, a16, M, K, K, 128);
    if (!rc) rc = make_tmap_u8_2d(&ta8, a8, M, 2 * K, 2 * K, 128, 128);
    if (!rc) rc = make_tmap_bf16_2d(&tw16, w16, N, K, K, 128);
    if (!rc) rc = make_tmap_u8_2d(&tw8, w8, N, 2 * K, 2 * K, 128, 128);
    if (!rc) {
      gemm::Params p = gp(M, N, K);
      p.bias = bias; p.out_f32 = out; p.ld_out = N; p.gelu = gelu; p.tma_store = 1; p.exit_wait_full = 1;
      p.timeline = g_test_timeline;
      gemm::OutMaps om;
      rc = make_tmap_store_2d(&om.f32, out, false, M, N, N, 16);
      if (!rc) {
        cudaError_t e = gemm::launch2_m8<256>(ta16, ta8, tw16, tw8, om, p, s, gemm::SliceMaps());
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
          set_error("regen_test_gemm(mixed8): %s", cudaGetErrorString(e));
          rc = REGEN_ECUDA;
        }
      }
    }
    cudaFree(a16); cudaFree(a8); cudaFree(w16); cudaFree(w8);
    return rc;
  }
  const int Kp = (int)ceil_div(K, 64) * 64;
  bf16 *ah, *al, *wh, *wl;
  REGEN_CUDA(cudaMalloc(&ah, (size_t)M * Kp * 2));
  REGEN_CUDA(cudaMalloc(&al, (size_t)M * Kp * 2));
  REGEN_CUDA(cudaMalloc(&wh, (size_t)N * Kp * 2));
  REGEN_CUDA(cudaMalloc(&wl, (size_t)N * Kp * 2));
  layers::launch_split_rows(A, K, ah, al, Kp, K, M, 1, 1, s);
  layers::launch_split_rows(W, K, wh, wl, Kp, K, N, 1, 1, s);
  CUtensorMap ta_h, ta_l, tw_h, tw_l;
  const bool pair = use_pair_kernel(M);
  int rc = make_tmap_bf16_2d(&ta_h, ah, M, Kp, Kp, 128);
  if (!rc) rc = make_tmap_bf16_2d(&ta_l, al, M, Kp, Kp, 128);
  if (!rc) rc = make_tmap_bf16_2d(&tw_h, wh, N, Kp, Kp, pair ? 128 : 256);
  if (!rc) rc = make_tmap_bf16_2d(&tw_l, wl, N, Kp, Kp, pair ? 128 : 256);
  if (!rc) {
    gemm::Params p = gp(M, N, Kp);
    p.bias = bias; p.residual = residual; p.ld_res = N; p.out_f32 = out; p.ld_out = N; p.gelu = gelu;
    p.timeline = g_test_timeline;
    gemm::OutMaps om;  // (gp() zeroed steplog)
    const char* nt = getenv("REGEN_DEBUG_NO_TMA_STORE");
    if ((N & 3) == 0 && !(nt && nt[0] == '1') &&
        make_tmap_store_2d(&om.f32, out, false, M, N, N, 16) == REGEN_OK)
      p.tma_store = 1;
    cudaError_t e;
    if (pair)
      e = precision == 0 ? gemm::launch2<256, true>(ta_h, ta_l, tw_h, tw_l, om, p, s)
                         : gemm::launch2<256, false>(ta_h, ta_l, tw_h, tw_l, om, p, s);
    else
      e = precision == 0 ? gemm::launch<256, true>(ta_h, ta_l, tw_h, tw_l, om, p, s)
                         : gemm::launch<256, false>(ta_h, ta_l, tw_h, tw_l, om, p, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      set_error("regen_test_gemm: %s", cudaGetErrorString(e));
      rc = REGEN_ECUDA;
    }
  }
  cudaFree(ah); cudaFree(al); cudaFree(wh); cudaFree(wl);
  return rc;
}

// Kernel-level test hook: causal multi-head self-attention on a seq-first q|k|v tensor through the tcgen05
// attention kernel.  qkv fp32 [T*B, 1536] -> out fp32 [T*B, 512] (hi + lo recombined).  Synchronises.
int regen_test_attention(const float* qkv, float* out, int32_t B, int32_t T, int32_t dbg, void* stream) {
  REGEN_CHECK_ARG(qkv && out && B >= 1 && T >= 1 && T <= 4096, "regen_test_attention: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t M = (size_t)B * T;
  bf16 *qh, *ql, *oh, *ol;
  REGEN_CUDA(cudaMalloc(&qh, M * 3 * D * 2));
  REGEN_CUDA(cudaMalloc(&ql, M * 3 * D * 2));
  REGEN_CUDA(cudaMalloc(&oh, M * D * 2));
  REGEN_CUDA(cudaMalloc(&ol, M * D * 2));
  layers::launch_split_rows(qkv, 3 * D, qh, ql, 3 * D, 3 * D, (int)M, 1, 1, s);
  CUtensorMap th, tl, toh, tol;
  const bool mc = (dbg & 8) && T > 64 && T <= 256;   // bit 3: multi-chunk compact kernel
  int rc = make_tmap_bf16_3d(&th, qh, 3 * D, B, T, (T <= 64 || mc) ? 64 : 128);
  if (!rc) rc = make_tmap_bf16_3d(&tl, ql, 3 * D, B, T, (T <= 64 || mc) ? 64 : 128);
  if (!rc) rc = make_tmap_bf16_3d(&toh, oh, D, B, T, 32);
  if (!rc) rc = make_tmap_bf16_3d(&tol, ol, D, B, T, 32);
  if (!rc) {
    attn::Params ap;
    ap.out_hi = oh; ap.out_lo = ol; ap.T = T; ap.Beff = B; ap.causal = (dbg & 2) ? 0 : 1; ap.dbg = dbg & 1; ap.m8 = 0;
    ap.timeline = g_test_timeline;
    ap.steplog = nullptr; ap.steplog_slot = 0; ap.steplog_cta = 0;
    ap.pol_load = 0; ap.pol_store = 0;
    cudaError_t e = T > 256 ? attn::launch_long(qh, ql, ap, s)
                    : T <= 64 ? attn::launch<64>(th, tl, toh, tol, ap, s)
                    : mc    ? attn::launch_mc(th, tl, toh, tol, ap, s)
                            : attn::launch<128>(th, tl, toh, tol, ap, s);
    if (e == cudaSuccess) {
      layers::merge_split_kernel<<<grid_cap(ceil_div((int64_t)M * D, 256)), 256, 0, s>>>(oh, ol, out, (int64_t)M * D);
      e = cudaStreamSynchronize(s);
    }
    if (e != cudaSuccess) {
      set_error("regen_test_attention: %s", cudaGetErrorString(e));
      rc = REGEN_ECUDA;
    }
  }
  cudaFree(qh); cudaFree(ql); cudaFree(oh); cudaFree(ol);
  return rc;
}

}  // extern "C"
