/*
 * regen_sm100.h -- C ABI of libregen_sm100.so, the B200 (sm_100a) implementation of the
 * ReGenNet diffusion-sampling hot path.
 *
 * The reference (liangxuy/ReGenNet) is pure Python/PyTorch and has no FFI layer; the drop-in
 * boundary is its Python API (CMDM.forward, SpacedDiffusion.p_sample_loop / ddim_sample_loop,
 * ClassifierFreeSampleModel.forward, rotation_6d_to_matrix).  The Python classes in
 * regennet_b200/ keep those signatures and bind the entry points below with ctypes; each
 * entry point names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, a negative REGEN_E* code otherwise; no C++ exception
 *     crosses the ABI.  regen_last_error() returns a thread-local message for the last failure.
 *   - all tensor pointers are DEVICE pointers owned by the caller (contiguous, fp32 / int64),
 *     unless a parameter is documented as host memory.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises the device, everything is CUDA-graph capturable.
 *   - layouts:  BJFT = [B, J, F, T] (T innermost, the reference's user-facing layout)
 *               TBI  = [T, B, I=J*F] (token-major, seq-first; the reference's internal
 *                      layout, model/cmdm.py:312-313, and the memory order of the permuted
 *                      tensors its sampler carries from step 2 on)
 */
#ifndef REGEN_SM100_H
#define REGEN_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REGEN_OK 0
#define REGEN_EINVAL (-1)   /* bad argument */
#define REGEN_ECUDA (-2)    /* CUDA runtime / driver error */
#define REGEN_ESTATE (-3)   /* call order violated (e.g. denoise before load_weights) */
#define REGEN_EUNSUPPORTED (-4)

#define REGEN_MAX_LAYERS 16

/* library / build info: "regen_sm100 <version> sm_100a" */
const char* regen_version(void);
const char* regen_last_error(void);
/* kernels launched by this library in this process so far (evidence for bench.py's gpu_launches) */
int64_t regen_launch_count(void);
/* credit n kernel launches executed by a CUDA-graph replay (the counter above only sees the launches
 * made while the graph was captured); called by the host layer after cudaGraphLaunch */
void regen_launch_count_add(int64_t n);

/* Step bookkeeping for a CUDA-graph-captured sampling loop.  Replaces the per-step host code of
 * diffusion/gaussian_diffusion.py:726-728 (`t = th.tensor([i] * shape[0])`) and the integer remap of
 * diffusion/respace.py:125-126 (`map_tensor[ts]`), bit-exact: with p = pos[0] (clamped to
 * [0, n_seq)), t_idx[0..B) = seq_idx[p], t_model[0..B) = seq_model[p], then pos[0] = p + 1.
 * All pointers are device int64; one launch, capturable. */
int regen_step_tables(const int64_t* seq_idx, const int64_t* seq_model, int64_t* pos, int64_t* t_idx,
                      int64_t* t_model, int32_t B, int32_t n_seq, void* stream);

/* ------------------------------------------------------------------------------------------
 * Handle-free elementwise operators (HBM-bound; coalesced, vectorised)
 * ---------------------------------------------------------------------------------------- */

/* Ancestral posterior update.  Replaces diffusion/gaussian_diffusion.py:265-287
 * (q_posterior_mean_variance), :366-388 (process_xstart + mean) and :544-559 (p_sample):
 *     x0c  = clip ? clamp(x0,-1,1) : x0
 *     out  = coef1[t_b]*x0c + coef2[t_b]*x + (t_b != 0) * exp(0.5*logvar[t_b]) * noise
 * Tables are fp32 device arrays built on the host in fp64 and then cast, exactly as
 * _extract_into_tensor does (:1604-1617).  t is int64[B] on the device (indices into the
 * tables).  Element e belongs to sample b = (e / inner) % B: inner = J*F*T for BJFT,
 * inner = I for TBI.  pred_xstart (nullable) receives x0c.  noise == NULL gives the posterior
 * mean only (p_mean_variance's "mean").  out may alias x. */
int regen_p_sample_update(const float* x, const float* x0, const float* noise, float* out,
                          float* pred_xstart, const int64_t* t, const float* coef1,
                          const float* coef2, const float* logvar, int64_t n_elem, int64_t inner,
                          int32_t B, int32_t clip_denoised, void* stream);

/* DDIM update.  Replaces diffusion/gaussian_diffusion.py:744-794 (ddim_sample) with
 * :418-423 (_predict_eps_from_xstart):
 *     eps   = (sqrt_recip_ac[t]*x - x0c) / sqrt_recipm1_ac[t]
 *     sigma = eta*sqrt((1-ac_prev[t])/(1-ac[t]))*sqrt(1-ac[t]/ac_prev[t])
 *     out   = x0c*sqrt(ac_prev[t]) + sqrt(1-ac_prev[t]-sigma^2)*eps + (t!=0)*sigma*noise    */
int regen_ddim_update(const float* x, const float* x0, const float* noise, float* out,
                      float* pred_xstart, const int64_t* t, const float* sqrt_recip_ac,
                      const float* sqrt_recipm1_ac, const float* ac, const float* ac_prev,
                      float eta, int64_t n_elem, int64_t inner, int32_t B, int32_t clip_denoised,
                      void* stream);

/* Classifier-free guidance combine.  Replaces model/cfg_sampler.py:31:
 *     out = uncond + scale[b]*(cond - uncond)                                            */
int regen_cfg_combine(const float* cond, const float* uncond, const float* scale, float* out,
                      int64_t n_elem, int64_t inner, int32_t B, void* stream);

/* Inpainting blend of the model output (motion editing, sample/edit.py:75-90).  Replaces
 * diffusion/gaussian_diffusion.py:319-323:  x0 = x0 * ~mask + motion * mask, in place, torch's operation order.
 * mask01 holds 0.0f / 1.0f; all three arrays share one memory layout (the sampler uses [T,B,I]). */
int regen_inpaint_blend(float* x0, const float* motion, const float* mask01, int64_t n_elem, void* stream);

/* PLMS sampler (diffusion/gaussian_diffusion.py:1007-1098; no caller in the reference, off the hot path).  Element e
 * belongs to sample b = (e / inner) % B; tables are fp32 device arrays of n_table entries indexed by t[b] + t_shift
 * (negative indices wrap).  regen_plms_eps: eps = (sqrt_recip_ac*x - clip(x0)) / sqrt_recipm1_ac, pred = clip(x0) (nullable).
 * regen_plms_combine: eps' from the history, e0 newest: order 1..4 Adams-Bashforth (:1075-1086), 5 = (e0 + e1) / 2 (:1066).
 * regen_plms_finish: mode 0  pred' = sra*x - srm1*eps', mean = pred'*sqrt(ac_prev) + sqrt(1 - ac_prev)*eps',
 *                            out = mean*(t != 0) + pred*(1 - (t != 0));   mode 1  out = pred*sqrt(ac_prev) + sqrt(1 - ac_prev)*eps'. */
int regen_plms_eps(const float* x, const float* x0, float* eps, float* pred, const int64_t* t,
                   const float* sqrt_recip_ac, const float* sqrt_recipm1_ac, int64_t n_elem, int64_t inner,
                   int32_t B, int32_t n_table, int32_t t_shift, int32_t clip_denoised, void* stream);
int regen_plms_combine(const float* e0, const float* e1, const float* e2, const float* e3, float* out,
                       int64_t n_elem, int32_t order, void* stream);
int regen_plms_finish(const float* x, const float* eps_prime, const float* pred, float* out, const int64_t* t,
                      const float* sqrt_recip_ac, const float* sqrt_recipm1_ac, const float* ac_prev, int64_t n_elem,
                      int64_t inner, int32_t B, int32_t n_table, int32_t mode, void* stream);

/* rot6d -> rotation matrix (Gram-Schmidt).  Replaces utils/rotation_conversions.py:513-534
 * (call sites model/rotation2xyz.py:56, 202, 270):
 * d6 [n,6] contiguous -> R [n,3,3] contiguous, rows (b1,b2,b3); 60 bytes of HBM traffic per rotation. */
int regen_rot6d_to_matrix(const float* d6, float* R, int64_t n, void* stream);

/* Post-sampling tail.  Temporal Gaussian smoothing = scipy.ndimage.gaussian_filter1d(x, sigma,
 * axis=-1, mode='reflect', truncate) as called at sample/cgenerate.py:142 (sigma 1) and render/crendermotion.py:79
 * (sigma 3), double accumulation like scipy.  layout 0: BJFT ([n_cols, T], T contiguous), 1: TBI ([T, n_cols]). */
int regen_gaussian_filter1d_time(const float* src, float* dst, int64_t n_cols, int32_t T, int32_t layout,
                                 double sigma, double truncate, void* stream);
/* Fused tail on the sampler's native layout: x TBI [T,B,J,6] -> Gaussian filter along T -> drop the last
 * drop_joints joints (translation row, model/rotation2xyz.py:253-255) -> rotation_6d_to_matrix ->
 * R [B,T,J-drop_joints,3,3] (the tensor model/rotation2xyz.py:270 hands to the body model). */
int regen_smooth_rot6d_to_matrix(const float* x_tbi, float* R, int32_t T, int32_t B, int32_t J,
                                 int32_t drop_joints, double sigma, double truncate, void* stream);

/* Layout conversion between the user-facing BJFT and the internal TBI layout.  Replaces the
 * permute/reshape of model/cmdm.py:312-313 (InputProcess) and :353-354 (OutputProcess). */
int regen_bjft_to_tbi(const float* src, float* dst, int32_t B, int32_t I, int32_t T, void* stream);
int regen_tbi_to_bjft(const float* src, float* dst, int32_t B, int32_t I, int32_t T, void* stream);

/* ------------------------------------------------------------------------------------------
 * The denoiser: CMDM.forward, arch='online' / 'offline'  (model/cmdm.py:173-252)
 * ---------------------------------------------------------------------------------------- */

typedef struct regen_handle regen_handle;

typedef struct {
  int32_t latent_dim;      /* D, must be 512 */
  int32_t num_heads;       /* H, must be 4 (head_dim 128) */
  int32_t ff_size;         /* F, must be 1024 */
  int32_t num_layers;      /* <= REGEN_MAX_LAYERS */
  int32_t input_feats;     /* I = njoints*nfeats */
  int32_t cm_mode;         /* 0 = 'add', 1 = 'concat' (model/cmdm.py:207-211) */
  int32_t max_batch;       /* largest effective batch (2x the user batch under CFG) */
  int32_t max_frames;      /* largest T */
  int32_t num_table_steps; /* size of the timestep-embedding table; timesteps must be < this */
  int32_t precision;       /* 0 = bf16x3 split (parity mode), 1 = single-pass bf16 (fast), 2 = mixed8: bf16x3 except the two
                              fused GEMM+LayerNorm kernels of the large-batch route, which run one fp16 + two e4m3 correction
                              MMAs per product, 3 = mixed8h: all GEMMs of that route (arch 0 only; otherwise like 2) */
  int32_t arch;            /* 0 = 'online': causal nn.TransformerDecoder, 1-token memory (model/cmdm.py:75-81, 203-227);
                              1 = 'offline': nn.TransformerEncoder over [condition token | frames], no mask (:63-71, 228-238) */
} regen_model_desc;

typedef struct {
  const float *qkv_w, *qkv_b;   /* self_attn.in_proj_{weight[3D,D],bias[3D]} */
  const float *o_w, *o_b;       /* self_attn.out_proj */
  const float *xv_w, *xv_b;     /* multihead_attn.in_proj_weight[2D:3D], in_proj_bias[2D:3D]  (arch 0 only) */
  const float *xo_w, *xo_b;     /* multihead_attn.out_proj                                    (arch 0 only) */
  const float *l1_w, *l1_b;     /* linear1 [F,D] */
  const float *l2_w, *l2_b;     /* linear2 [D,F] */
  const float *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b; /* norm1..3 (arch 1: norm1, norm2 only) */
} regen_layer_weights;

typedef struct {
  const float *in_w, *in_b;     /* input_process.poseEmbedding [D,I] */
  const float *cmo_w, *cmo_b;   /* cmo_process.poseEmbedding [D,I] */
  const float *fuse_w, *fuse_b; /* fuse_process [D,2D]; NULL for cm_mode 'add' */
  const float *t0_w, *t0_b;     /* embed_timestep.time_embed.0 */
  const float *t2_w, *t2_b;     /* embed_timestep.time_embed.2 */
  const float *pe;              /* sequence_pos_encoder.pe [pe_len, D] */
  int32_t pe_len;
  const float *out_w, *out_b;   /* output_process.poseFinal [I,D] */
  const float *action_emb;      /* embed_action.action_embedding [num_actions, D] or NULL */
  int32_t num_actions;
  const float *text_w, *text_b; /* embed_text [D, clip_dim], [D] or NULL */
  int32_t clip_dim;
  regen_layer_weights layers[REGEN_MAX_LAYERS];
} regen_weight_ptrs;

/* Replaces CMDM.__init__ (model/cmdm.py:13-111) for the device-side state. */
int regen_create(regen_handle** out, int32_t device, const regen_model_desc* desc);
void regen_destroy(regen_handle* h);

/* Replaces load_state_dict (utils/model_util.py:5-8): packs fp32 weights into the library's own
 * bf16 hi/lo operand buffers, folds fuse_process into the input projection, folds the 1-token
 * cross-attention (value+out projection) and builds the timestep-embedding table. */
int regen_load_weights(regen_handle* h, const regen_weight_ptrs* w, void* stream);

/* Loop-invariant conditioning, once per sampling loop (or per forward in the generic route).
 * cmotion: actor motion, BJFT [B,J,F,T].  action: int64[B] class indices (EmbedAction row gather,
 * model/cmdm.py:358-366) or NULL.  text_feat: [B, clip_dim] CLIP text features (embed_text Linear,
 * model/cmdm.py:182-184) or NULL.  uncond != 0 zeroes both (mask_cond force_mask, :129-132).
 * guidance != 0 doubles the batch internally: rows [0,B) conditional, rows [B,2B) unconditional
 * -- model/cfg_sampler.py:24-31.  Replaces model/cmdm.py:181-187, :202 (cmo_process), the cmotion
 * half of :207-211 and the positional-encoding add :218. */
int regen_prepare_cond(regen_handle* h, const float* cmotion_bjft, const int64_t* action,
                       const float* text_feat, int32_t B, int32_t T, int32_t guidance,
                       int32_t uncond, void* stream);

/* CMDM.forward (model/cmdm.py:173-252) after regen_prepare_cond.  x_tbi [T,B,I]; t int64[B] device,
 * original (un-respaced) timesteps; x0_tbi [T,B,I] (with guidance: the guided combination
 * uncond + scale[b]*(cond-uncond), scale = cfg_scale[B] device). */
int regen_denoise(regen_handle* h, const float* x_tbi, const int64_t* t, const float* cfg_scale,
                  float* x0_tbi, int32_t B, int32_t T, void* stream);

/* Measurement support: between begin and end every kernel the denoiser launches is bracketed by a
 * CUDA event pair on its stream; end synchronises and returns the summed device time (ms) and the
 * launch count per kernel class: [0] tcgen05 GEMMs, [1] attention, [2] LayerNorm, [3] other. */
int regen_profile_begin(regen_handle* h);
int regen_profile_end(regen_handle* h, float* ms, int32_t* launches);

/* ------------------------------------------------------------------------------------------
 * Evaluation feature extractor: ST-GCN inference.  Replaces STGCN.forward of
 * eval/a2m/recognition/models/stgcn.py:76-126 (st_gcn blocks :145-213, graph convolution
 * eval/a2m/recognition/models/stgcnutils/tgcn.py:55-64) in model.eval() mode.
 * STATUS: built against the pinned oracle without GPU time left in its round; see regennet_b200/csrc/stgcn.cu.
 *
 * Packed weights: ONE fp32 device buffer, tensors in this order, each contiguous in its state-dict shape:
 *   A [K,V,V];  data_bn {weight, bias, running_mean, running_var} [in_channels*V];
 *   for block i = 0..9 (channels 64,64,64,64,128,128,128,256,256,256; temporal stride 2 at i = 4 and 7):
 *     gcn.conv.weight [K*Cout,Cin], gcn.conv.bias [K*Cout], tcn.0 {weight,bias,running_mean,running_var} [Cout],
 *     tcn.2.weight [Cout,Cout,9], tcn.2.bias [Cout], tcn.3 {weight,bias,running_mean,running_var} [Cout],
 *     (i = 4, 7 only) residual.0.weight [Cout,Cin], residual.0.bias [Cout], residual.1 {4 x [Cout]},
 *     edge_importance.i [K,V,V];
 *   fcn.weight [num_class,256], fcn.bias [num_class].
 * regen_stgcn_packed_size returns the number of floats (negative for a bad descriptor).
 * regen_stgcn_forward: output fp32 [N, V, in_channels, T] (the sampler's batch["output"]; in_channels =
 * C * num_person with the persons stacked along that axis) -> features fp32 [N,256], yhat fp32 [N,num_class].
 * ---------------------------------------------------------------------------------------- */
typedef struct regen_stgcn regen_stgcn;
typedef struct {
  int32_t in_channels; /* C * num_person */
  int32_t num_person;  /* 1 or 2 */
  int32_t num_class;
  int32_t num_node;    /* V */
  int32_t num_part;    /* K adjacency partitions */
} regen_stgcn_desc;
int64_t regen_stgcn_packed_size(const regen_stgcn_desc* d);
int regen_stgcn_create(regen_stgcn** h, int32_t device, const regen_stgcn_desc* d);
int regen_stgcn_load_weights(regen_stgcn* h, const float* packed, int64_t n_floats, void* stream);
int regen_stgcn_forward(regen_stgcn* h, const float* output, int32_t N, int32_t T, float* features, float* yhat,
                        void* stream);
void regen_stgcn_destroy(regen_stgcn* h);

/* ------------------------------------------------------------------------------------------
 * Kernel-level test hook (used by tests/ only): out[M,N] = act(A[M,K] . W[N,K]^T + bias + residual)
 * through the same tcgen05/TMA GEMM kernel the denoiser uses (the arithmetic of every nn.Linear on
 * the path, torch F.linear).  fp32 device buffers in and out; synchronises the stream.
 * ---------------------------------------------------------------------------------------- */
int regen_test_gemm(const float* A, const float* W, const float* bias, const float* residual,
                    float* out, int32_t M, int32_t N, int32_t K, int32_t gelu, int32_t precision,
                    void* stream);

/* Bring-up instrumentation for regen_test_gemm: when set to a device buffer of 128 uint64, CTA 0 of the next test
 * GEMMs records SM clock values at pipeline events (see gemm_sm100.cuh Params::timeline).  NULL disables. */
int regen_test_gemm_timeline(unsigned long long* device_buf128);
/* Bring-up instrumentation: from now on the tcgen05 kernels regen_denoise launches (GEMM, fused GEMM+LN, attention) log
 * into device_buf[2*slot] = earliest "inputs available" time and [2*slot+1] = latest CTA exit time (nanoseconds of the
 * GPU global timer, atomicMin / atomicMax -- initialise to ~0 / 0), slot = launch order since this call.  The slots are
 * baked into captured graphs; capacity_slots = 0 turns logging off.  The GEMM kernels also store every CTA's exit time
 * at device_buf[2*capacity_slots + slot*160 + blockIdx.x] (the buffer must hold 2*cap + 160*cap words).
 * Used by tools/step_timeline.py only. */
int regen_test_step_log(regen_handle* h, unsigned long long* device_buf, int32_t capacity_slots);

/* Kernel-level test hook (tests/ only): causal 4-head self-attention (head_dim 128) of a seq-first q|k|v
 * tensor through the tcgen05 attention kernel -- the arithmetic of nn.MultiheadAttention with the causal
 * mask of model/cmdm.py:168-171.  qkv fp32 [T*B, 1536] (row = t*B + b) -> out fp32 [T*B, 512].  dbg = 0. */
int regen_test_attention(const float* qkv, float* out, int32_t B, int32_t T, int32_t dbg, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REGEN_SM100_H */
