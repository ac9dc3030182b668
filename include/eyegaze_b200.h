/*
 * eyegaze_b200.h -- C ABI of libeyegaze_b200.so: the sm_100a kernels under the drop-in
 * nn.Module classes (DualEEGTransformer / EarlyFusionViT / LateFusionViT / FuzzyGatingFusion).
 *
 * The reference (roseDwayane/EyeGaze-Multimodal) has no native boundary: its hot path is
 * ATen ops issued from Python.  Each entry point below replaces the group of ATen calls
 * cited beside it (paths relative to the reference root); INTEGRATION.md shows the ctypes
 * binding the Python host side uses.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named h_*; sizes are element counts;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - dtype codes: EGB_F32 = 0, EGB_BF16 = 1 (HBM storage type of activations);
 *   - return value 0 = success; otherwise egb_last_error() holds a message.  There is no CPU
 *     fallback anywhere in this library.
 */
#ifndef EYEGAZE_B200_H_
#define EYEGAZE_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGB_F32 0
#define EGB_BF16 1

const char* egb_last_error(void);
int egb_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t egb_launch_count(void);
/* optional CUDA-event timing of every tensor-core GEMM launch (bench.py roofline): enable, run, read.
 * egb_prof_read: kind 0 = tcgen05 GEMM; out[4] = {launches, total ms, total FLOPs, total algorithmic bytes} */
int egb_prof_enable(int on);
int egb_prof_read(int kind, double* out, int reset);
/* per-launch records in launch order: out[i*8..] = {ms, FLOPs, bytes, tag0..tag3 (GEMM: M, N, K, variant), 0} */
int egb_prof_dump(int kind, double* out, int max_records, int* n_out);

/* Device-resident seed epoch for CUDA-graph replays.  Dropout seeds are launch arguments and therefore frozen into a
 * captured graph; after egb_seed_epoch_enable every mask-drawing kernel of this library mixes the current value of one
 * device word into its seed, and egb_seed_epoch_advance (one 1-thread launch, capturable) increments that word -- once
 * per replay, so forward and backward of a replay agree and successive replays draw fresh masks. */
int egb_seed_epoch_enable(void** device_word_out);
int egb_seed_epoch_advance(void* stream);

/* ---------------------------------------------------------------------------------------------
 * Generalised GEMM   C[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T )
 *
 * Replaces every nn.Linear / nn.Conv1d / nn.Conv2d(32->64) / patch-embed Conv2d on the path:
 *   art.py:203-205,213,272 (q/k/v/out projections, FFN), dual_eeg_transformer.py:81-86,154-171,
 *   863-868,923,1074-1079,1100-1105, and timm's qkv/proj/fc1/fc2/patch_embed (early_fusion_vit.py:86).
 *
 * Operands are strided 3-level views so convolutions run as implicit GEMMs without im2col:
 *   flat row index i  ->  group g = i / rows_per_group, r = i % rows_per_group,
 *   element (i, j)    ->  base[g * group_stride + r * row_stride + j]      (inner stride is 1)
 * `major` = 0: rows are the M (or N) index and the inner index is K (reduction)   ("K-major")
 * `major` = 1: rows are the K (reduction) index and the inner index is M (or N)   ("MN-major")
 * Overlapping rows (row_stride < inner extent) are legal: that is the conv-as-GEMM view.
 * The 3x3 Conv2d of the spectrogram CNN uses inner segmentation (one segment per kernel row) over a
 * zero-padded channels-last image, so it also runs without an im2col buffer.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;
  int32_t major;          /* 0 = K-major, 1 = MN-major */
  int32_t rows_per_group; /* <= 0: a single group holding all rows */
  int64_t row_stride;     /* elements */
  int64_t group_stride;   /* elements */
  /* Optional inner-index segmentation (2-D convolutions): inner index i is split as
   * seg = i / seg_len, i' = i % seg_len and addresses row (row + seg * seg_row_shift), inner i'.
   * seg_len = 0 disables it; seg_len must be a multiple of 64.  Single-group operands only. */
  int32_t seg_len;
  int32_t seg_row_shift;
} egb_operand;

typedef struct {
  void* ptr;
  int32_t dtype;          /* EGB_F32 / EGB_BF16 */
  int32_t rows_per_group; /* <= 0: single group */
  int64_t row_stride;
  int64_t group_stride;
} egb_matrix;

#define EGB_ACT_NONE 0
#define EGB_ACT_RELU 1
#define EGB_ACT_GELU 2
#define EGB_ACT_GELU_DGRAD 3   /* GELU; c_pre (required) receives gelu'(pre-activation) instead of the pre-activation */
/* backward-side epilogues: multiply the accumulator by act'(.) read from `aux` */
#define EGB_ACTBWD_NONE 0
#define EGB_ACTBWD_RELU_MASK 1 /* aux = forward OUTPUT y (post relu [+dropout]); x *= (y != 0) * aux_scale */
#define EGB_ACTBWD_GELU 2      /* aux = forward PRE-activation; x *= gelu'(aux) */
#define EGB_ACTBWD_MUL 3       /* aux = stored activation derivative (EGB_ACT_GELU_DGRAD); x *= aux */

typedef struct {
  int32_t M, N, K;
  int32_t in_dtype;    /* dtype of A and B: EGB_BF16 -> tcgen05 tensor-core kernel; EGB_F32 -> FP32 FMA kernel */
  egb_operand a, b;
  egb_matrix c;        /* output */
  egb_matrix c_pre;    /* optional (ptr may be NULL): pre-activation copy (bias added, before act) */
  egb_matrix residual; /* optional: added after activation/dropout */
  egb_matrix aux;      /* optional: operand of the act_bwd epilogue */
  const float* bias;   /* optional, length N (fp32) */
  float alpha;
  int32_t act;
  int32_t act_bwd;
  float aux_scale;
  float dropout_p;     /* applied after activation; 0 disables */
  uint64_t dropout_seed;
  int32_t accumulate;  /* 1: C += result (C must be fp32, split-K with red.global.add); 2: zero C first, then accumulate */
  int32_t split_k;     /* accumulating launches: number of K splits, 0 = auto.  Non-accumulating fp32 launches of small problems:
                          0 = the library may split K over a thread-block cluster (sum of the partial sums in rank order),
                          1 = one CTA walks K in order (the summation order the fp32 parity mode is pinned with) */
  float* c_colsum;     /* optional [N] fp32, ACCUMULATED (caller zeroes): column sums of the stored C -- the bias gradient
                          of the layer whose pre-activation gradient this GEMM produces (F.linear backward), taken in the
                          epilogue instead of a second pass over C */
} egb_gemm_desc;

int egb_gemm(const egb_gemm_desc* d, void* stream);
/* debug aid: [grid][8] int64 device buffer receiving per-CTA barrier-wait cycle totals of the following tensor-core
 * GEMM launches (NULL disables) */
int egb_debug_gemm_timing(long long* device_buf);

/* ---------------------------------------------------------------------------------------------
 * dtype casts / weight re-layout (per-step bf16 copies of the fp32 master parameters)
 * ------------------------------------------------------------------------------------------- */
int egb_cast_from_f32(const float* src, void* dst, int dtype, int64_t n, void* stream);
int egb_cast_to_f32(const void* src, int dtype, float* dst, int64_t n, void* stream);
/* dst[i0*d0+i1*d1+i2*d2+i3*d3] = src[i0*s0+i1*s1+i2*s2+i3*s3], any dtype pair */
int egb_copy_strided4(const void* src, int src_dtype, void* dst, int dst_dtype, const int32_t* sizes,
                      const int64_t* src_strides, const int64_t* dst_strides, void* stream);
int egb_zero(void* ptr, int64_t nbytes, void* stream); /* cudaMemsetAsync */

/* (B,C,T) fp32 x2 -> [2B, Tp, C] channels-last, `pad` zero rows in front (the Conv1d padding,
 * dual_eeg_transformer.py:154), zero rows behind up to Tp.  Operand layout of the implicit-GEMM conv. */
int egb_eeg_pack(const float* eeg1, const float* eeg2, void* out, int dtype, int B, int C, int T, int pad, int Tp,
                 void* stream);

/* Token sequence [cls | ibs tokens (shared by both players) | spectrogram tokens | temporal tokens] + learned
 * positional embedding (dual_eeg_transformer.py:1157-1179, art.py:120-126).  S = 2B stacked players. */
int egb_seq_assemble_fwd(const float* cls, const float* pos, const void* ibs, const void* spec, const void* h, void* out,
                         int dtype, int S, int B, int L, int D, int n_ibs, int n_spec, int n_h, void* stream);
/* dpos[L,D] (+=) = sum_s dx[s]; dibs[B,n_ibs,D] = dx[b,1+t] + dx[B+b,1+t] (may be NULL) */
int egb_seq_assemble_bwd(const void* dx, float* dpos, void* dibs, int dtype, int S, int B, int L, int D, int n_ibs,
                         void* stream);

/* out[b,t,:] = x[b,t,:] + e[t,:] (IBS type embedding, dual_eeg_transformer.py:909) */
int egb_add_rows_broadcast(const void* x, const float* e, void* out, int dtype, int64_t rows, int NT, int D, void* stream);

/* CLS slice, temporal mean-pool, IBS-token mean-pool, symmetric features (dual_eeg_transformer.py:1193-1225,933-938).
 * zf is the (B,3D) classifier input; columns [0,D) are filled later by the SymmetricFusion GEMM. */
int egb_tail_pool_fwd(const void* z, int dtype, float* cls1, float* cls2, float* sym, float* zf, float* ibs_pool, int B,
                      int L, int D, int n_ibs, int offset, int ibs_single, void* stream);
int egb_tail_pool_bwd(const void* z, int dtype, const float* dcls1, const float* dcls2, const float* dsym,
                      const float* dzf, const float* dibs_pool, void* dz, int B, int L, int D, int n_ibs, int offset,
                      int ibs_single, void* stream);

/* out[n] (+)= sum_m x[m,n]  -- bias gradients */
int egb_colsum(const egb_matrix* x, int M, int N, float* out, int zero_first, void* stream);
/* out = dy * act'(aux) (mode 1: aux = relu output, factor (aux!=0)*scale; mode 2: aux = pre-activation, gelu') */
int egb_act_bwd(const egb_matrix* dy, const void* aux, const egb_matrix* out, int M, int N, int mode, float scale,
                void* stream);
/* out = dy * mask(seed) / (1-p): regenerates the mask a GEMM epilogue applied (index m*N+n) */
int egb_dropout_bwd(const void* dy, void* out, int dtype, int64_t n_elems, float p, uint64_t seed, void* stream);
/* same over [M, N] rows (N % 8 == 0, N <= 1024) with colsum (optional [N] fp32, ACCUMULATED) = column sums of `out`:
 * the bias gradient of the layer whose epilogue applied the mask, in the same pass */
int egb_dropout_bwd_colsum(const void* dy, void* out, int dtype, int M, int N, float p, uint64_t seed, float* colsum,
                           void* stream);
/* mean cross entropy (F.cross_entropy) and d(loss)/d(logits) in one pass; labels are int64 */
int egb_cross_entropy(const float* logits, const int64_t* labels, float* loss, float* dlogits, int B, int C, void* stream);
int egb_scale_by_device_scalar(const float* x, const float* g, float* out, int64_t n, void* stream);

/* LayerNorm over the last dim (art.py:283-296,306; timm eps 1e-6).  bwd ACCUMULATES into dgamma/dbeta. */
int egb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int dtype,
                      int M, int D, float eps, void* stream);
int egb_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                      float* dgamma, float* dbeta, int dtype, int M, int D, void* stream);
/* same, with dx += dres: the gradient that reaches x around the normalisation (pre-norm residual block,
 * timm Block.forward `x = x + attn(norm1(x))`), so autograd's separate accumulation pass disappears */
int egb_layernorm_bwd_res(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                          void* dx, float* dgamma, float* dbeta, const void* dres, int dtype, int M, int D, void* stream);
/* same, plus dx_colsum (optional [D] fp32, ACCUMULATED): column sums of the stored dx = the bias gradient of the Linear
 * whose output was added into the stream this LayerNorm reads (art.py:293-295 / timm Block), so F.linear's bias-gradient
 * pass over dY disappears */
int egb_layernorm_bwd_ex(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                         void* dx, float* dgamma, float* dbeta, const void* dres, float* dx_colsum, int dtype, int M,
                         int D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused multi-head attention (art.py:203-213; timm Attention): softmax(QK^T*scale) [dropout] V without
 * materialising the probabilities.  Tensors are addressed base[b*bs + row*rs + head*head_dim + d].
 * kv_shift: query batch s uses keys/values of batch (s+kv_shift)%S  (CrossBrainAttention, both directions
 * in one launch).  `probs` (optional, fp32 [S,H,Lq,Lk]) exports the softmax for the analysis hooks.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void *q, *k, *v;
  void* o;
  int64_t q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;
  const void* d_o;
  int64_t do_bs, do_rs;
  void *dq, *dk, *dv;
  int64_t dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs;
  float* lse;   /* [S,H,Lq] saved by fwd */
  float* delta; /* [S,H,Lq] scratch of bwd */
  float* probs;
  int32_t dtype, S, H, Lq, Lk, head_dim, kv_shift;
  float scale, dropout_p;
  uint64_t seed;
  /* backward only, optional (all three or none): [H*head_dim] fp32 each, ACCUMULATED with the column sums of the stored
   * dq / dk / dv over all (batch, row) -- the bias gradients of the q / k / v projections (art.py:203-205, timm qkv),
   * taken while the gradient tiles leave the kernel instead of a second pass over them */
  float *dq_colsum, *dk_colsum, *dv_colsum;
} egb_attention_desc;
int egb_attention_fwd(const egb_attention_desc* d, void* stream);
/* debug aid: 8 x int64 device buffer receiving clock64() phase stamps of one CTA of the following tensor-core
 * attention launches (NULL disables) */
int egb_debug_attention_timing(long long* device_buf);
int egb_attention_bwd(const egb_attention_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Spectrogram tokens (dual_eeg_transformer.py:88-135)
 * ------------------------------------------------------------------------------------------- */
/* STFT(n_fft, hop, hann window, centre/reflect) -> |.| -> first `bins` -> log(+1e-8); out fp32 [2*n_sig, bins, 1+T/hop] */
int egb_stft_logmag(const float* eeg1, const float* eeg2, const float* window, float* out, int n_sig_per_stream, int T,
                    int n_fft, int hop, int bins, void* stream);
/* Conv2d(1->32,3x3,p1)+ReLU+MaxPool2 fused; out = zero-bordered channels-last [N, Hh/2+2, Ww/2+2, 32].
   amax (optional, [N, (Hh/2)*(Ww/2), 8] uint16): per pooled output 3 bits -- which of the 2x2 conv outputs won the pool
   (0..3, first maximum as in max_pool2d) or 4 = ReLU inactive; four channels per word.  The backward call that gets the
   record skips the recomputation of the convolution. */
int egb_spec_conv1_pool_fwd(const float* img, const float* w, const float* bias, void* out, int dtype, int N, int Hh,
                            int Ww, int64_t out_elems, uint16_t* amax, void* stream);
int egb_spec_conv1_pool_bwd(const float* img, const float* w, const float* bias, const void* dout, int dtype, float* dw,
                            float* db, int N, int Hh, int Ww, const uint16_t* amax, void* stream);
/* 3x3 convolution, 64 -> 32 channels, over zero-bordered channels-last images as an implicit GEMM that stages every input
   tile once (data gradient of spec_conv[3], dual_eeg_transformer.py:81-86):
       y[(m + out_shift) * 32 + c] = sum_{a, b < 3, o < 64} x[(m + a * Wp + b) * 64 + o] * w[c * 768 + a * 256 + b * 64 + o]
   for m in [0, M).  x: bf16 [x_rows, 64]; w: bf16 [32, 768]; y: bf16 rows of 32 channels. */
int egb_conv3x3_c64_c32(const void* x, long long x_rows, const void* w, void* y, long long M, long long out_shift, int Wp,
                        void* stream);
/* the forward convolution spec_conv[3] (32 -> 64 channels, + bias) by the same kernel:
       y[(m + out_shift) * 64 + o] = bias[o] + sum_{a, b < 3, c < 32} x[(m + a * Wp + b) * 32 + c] * w[o * 384 + a * 128 + b * 32 + c]
   x: bf16 [x_rows, 32]; w: bf16 [64, 384]; bias: fp32 [64]; y: bf16 rows of 64 channels. */
int egb_conv3x3_c32_c64(const void* x, long long x_rows, const void* w, const float* bias, void* y, long long M,
                        long long out_shift, int Wp, void* stream);
/* weight gradient of spec_conv[3], transposed and in the four-slot segment layout of the forward weight matrix:
       dwt[(a * 128 + j * 32 + c) * 64 + o] += sum_{m < M} x[(m + a * Wp + j) * 32 + c] * dy[(m + dy_shift) * 64 + o]   a, j < 3
   x: bf16 [x_rows, 32]; dy: bf16 [dy_rows, 64], zero at border positions; dwt: fp32 [384, 64], ZEROED BY THE CALLER (rows
   j = 3 stay zero).  Both operands cross L2 -> shared memory once; partial sums are added with fp32 atomics. */
int egb_conv3x3_dw_c32_c64(const void* x, long long x_rows, const void* dy, long long dy_rows, float* dwt, long long M,
                           long long dy_shift, int Wp, void* stream);
/* ReLU + AdaptiveAvgPool2d(4,4) + flatten over the padded conv-2 output [N, H1+2, W1+2, 64] */
int egb_relu_avgpool_fwd(const void* y, void* out, int dtype, int N, int H1, int W1, void* stream);
/* backward: dy (zero on the border and where ReLU is inactive); db, when not NULL, receives (+=, fp32 atomics; zeroed by
   the caller) the column sums of dy = the bias gradient of spec_conv[3], so dy is not read again for it */
int egb_relu_avgpool_bwd(const void* y, const void* dpool, void* dy, int dtype, int N, int H1, int W1, float* db, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Inter-brain-synchrony connectivity (dual_eeg_transformer.py:473-819), parameter-free, forward only.
 * out fp32 [B, n_bands, n_out, C, C].  Scratch (floats): phase, xb: B*n_bands*2*C*T each; stats:
 * B*n_bands*2*C*8; pspec: B*2*C*(max(hi)-min(lo)+1).  twiddle: T/2 complex exp(-2 pi i k/T).
 * ------------------------------------------------------------------------------------------- */
int egb_ibs_connectivity(const float* eeg1, const float* eeg2, const float* twiddle, float* phase, float* xb,
                         float* stats, float* pspec, float* out, int B, int C, int T, int n_bands, const int32_t* band_lo,
                         const int32_t* band_hi, const int32_t* slot_of, int n_out, void* stream);
/* Legacy scalar IBS features of IBSTokenGenerator (dual_eeg_transformer.py:418-470, `ibs_mode: scalar`): n_bands x 7 global
 * scalars per trial -> out fp32 [B, n_bands*7].  Scratch as above plus cspec: B*2*C*nbins complex values (2 floats each). */
int egb_ibs_scalar_features(const float* eeg1, const float* eeg2, const float* twiddle, float* phase, float* xb,
                            float* stats, float* pspec, float* cspec, float* out, int B, int C, int T, int n_bands,
                            const int32_t* band_lo, const int32_t* band_hi, void* stream);
/* InstanceNorm1d over the token axis per (trial, matrix cell) (dual_eeg_transformer.py:893-901) */
int egb_instnorm_tokens_fwd(const float* x, const float* gamma, const float* beta, void* y, int dtype, int B, int NT,
                            int P, float eps, int apply_norm, void* stream);
int egb_instnorm_tokens_bwd(const float* x, const void* dy, int dtype, float* dgamma, float* dbeta, int B, int NT, int P,
                            float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FuzzyGatingFusion (fuzzy_gating_fusion.py:297-390).  mode: 0 full, 1 no_temperature,
 * 2 no_fuzzification, 3 fixed_weights.  aux: (B+1) rows of 16 floats; row b = {H_img,H_eeg,mu[4],w[4],alpha,T_img,T_eeg},
 * row B = {sigma_rel_img, sigma_rel_eeg, sigma_unrel_img, sigma_unrel_eeg, theta[4], T_img, T_eeg}.
 * dparams (12 floats, accumulated): tau_img, tau_eeg, c_unrel_img, c_unrel_eeg, ls_rel_img, ls_rel_eeg,
 * ls_unrel_img, ls_unrel_eeg, beta[4].
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float *tau_img, *tau_eeg, *c_reliable, *c_unreliable_img, *c_unreliable_eeg;
  const float *log_sigma_reliable_img, *log_sigma_reliable_eeg, *log_sigma_unreliable_img, *log_sigma_unreliable_eeg;
  const float* beta;
  int32_t mode, B, num_classes;
  float eps_temp, eps_log, eps_div;
} egb_fuzzy_desc;
int egb_fuzzy_fwd(const egb_fuzzy_desc* d, const float* img, const float* eeg, float* fused, float* alpha, float* aux,
                  void* stream);
int egb_fuzzy_bwd(const egb_fuzzy_desc* d, const float* img, const float* eeg, const float* g_fused,
                  const float* g_alpha, float* d_img, float* d_eeg, float* dparams, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Gaze branch prologue (early_fusion_vit.py:149-196 + timm PatchEmbed): fuse the two heat-maps and emit
 * the [B*n_patches, C*ps*ps] patch matrix.  mode: 0 concat 1 add 2 subtract 3 subtract_abs 4 multiply
 * 5 single image.  Images are (B,3,H,W) fp32 with an explicit batch stride (channel-sliced views are legal).
 * egb_fill_row0: out[s,0,:] = cls + pos[0,:].
 * ------------------------------------------------------------------------------------------- */
int egb_vit_patchify(const float* img_a, const float* img_b, int64_t a_batch_stride, int64_t b_batch_stride, void* out,
                     float* stats_scratch, int dtype, int B, int H, int W, int ps, int mode, void* stream);
int egb_fill_row0(const float* cls, const float* pos, void* out, int dtype, int S, int L, int D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-step tail (the step right after the path): global-norm gradient clipping and AdamW
 * over all parameters in two launches -- torch.nn.utils.clip_grad_norm_(params, max_norm) + torch.optim.AdamW.step()
 * (train_art.py:221-229, train_multimodal_fuzzy_fusion.py:464-472).  fp32 parameters, gradients and moments.
 *   tensor_table : device array of { float* p; const float* g; float* exp_avg; float* exp_avg_sq; int64 step; }
 *                  (g NULL = skipped; step = the tensor's own 1-based update count, used for the bias corrections)
 *   chunk_table  : device array of { int32 tensor; int32 n; int64 offset; }, one CTA per chunk
 * egb_multi_tensor_sqnorm ACCUMULATES sum(g^2) into out_sqnorm (caller zeroes); egb_multi_tensor_adamw reads it on the
 * device (sqnorm NULL or max_norm <= 0: no clipping).
 * ------------------------------------------------------------------------------------------- */
/* fp32 master parameters -> bf16 compute copies, every tensor of the table in ONE launch.  cast_table: device array of
   {const float* src, bf16* dst} records; chunk_table: {int32 tensor, int32 n, int64 offset} records, one CTA each. */
int egb_multi_tensor_cast_bf16(const void* cast_table, const void* chunk_table, int n_chunks, void* stream);
int egb_multi_tensor_sqnorm(const void* tensor_table, const void* chunk_table, int n_chunks, float* out_sqnorm,
                            void* stream);
int egb_multi_tensor_adamw(const void* tensor_table, const void* chunk_table, int n_chunks, float lr, float beta1,
                           float beta2, float eps, float weight_decay, float max_norm, const float* sqnorm, void* stream);
/* Same update with DEVICE-resident step state, so that no launch argument changes from step to step (CUDA-graph
 * capturable, no host read / write per step).  Every pointer may be NULL (then the host value / default applies):
 *   lr         : this group's learning rate (egb_lr_schedule_step writes it; replaces the CosineAnnealingLR / LambdaLR
 *                host schedulers of train_art.py:401-409, train_multimodal_fuzzy_fusion.py:197-214,743-750)
 *   ctrl       : 4 floats written by egb_adamw_prepare: {skip this step, steps skipped so far, device step count, -}
 *   grad_scale : gradients are divided by it first -- torch.amp.GradScaler's scale (train_multimodal_fuzzy_fusion.py:462)
 *   use_device_step : bias corrections use ctrl[2] instead of the per-tensor step column of the table
 * egb_adamw_prepare (one thread, between the norm pass and the update): skip = (*found_inf != 0) [GradScaler's verdict,
 * the `_step_supports_amp_scaling` optimizer protocol] or (skip_nonfinite and sqrt(*sqnorm) / scale is inf / nan) [the
 * finite check taken from our own norm pass].  A skipped step changes nothing and does not count towards the bias
 * corrections, like torch's fused AdamW; advance_step != 0 also advances ctrl[2] on a step that is not skipped. */
typedef struct {
  const float* lr;
  const float* ctrl;
  const float* grad_scale;
  int32_t use_device_step;
} egb_adamw_state;
int egb_adamw_prepare(float* ctrl, const float* sqnorm, const float* grad_scale, const float* found_inf,
                      int skip_nonfinite, int advance_step, void* stream);
int egb_multi_tensor_adamw_ex(const void* tensor_table, const void* chunk_table, int n_chunks, float lr, float beta1,
                              float beta2, float eps, float weight_decay, float max_norm, const float* sqnorm,
                              const egb_adamw_state* state, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-side training-step state (rest): LR schedules, loss / metric accumulation.
 * egb_lr_schedule_step: sched[0] (the scheduler's step counter, float) += 1, then for each of the n groups
 *   lr_out[g] = base_lr[g] * factor(sched[0]);  opt_step (may be NULL) += 1 as well when advance_opt != 0.
 *   kind 0 constant; kind 1 CosineAnnealingLR(T_max = p0, eta_min = p1) in closed form (train_art.py:401-409; for
 *   eta_min != 0 lr_out = eta_min + (base - eta_min) * factor); kind 2 linear warm-up for p0 steps then cosine decay to 0
 *   at p1 total steps (train_multimodal_fuzzy_fusion.py:197-214).  `advance` = 0 recomputes lr_out without stepping.
 * egb_accum_scalars: acc[i] += *src[i] for i < n (n <= 8 device scalars, e.g. the six losses train_art.py:224-229 reads
 *   with .item() every step); acc[n] += 1 (batch count).
 * egb_argmax_count: acc[0] += #(argmax(logits[b]) == labels[b]), acc[1] += B; optional preds[b] (int64) is written
 *   (train_multimodal_fuzzy_fusion.py:507-509 copies predictions to the host every step).
 * ------------------------------------------------------------------------------------------- */
int egb_lr_schedule_step(float* sched, float* opt_step, const float* base_lr, float* lr_out, int n_groups, int kind,
                         float p0, float p1, int advance, int advance_opt, void* stream);
int egb_accum_scalars(const float* const* h_src, int n, float* acc, void* stream);
int egb_argmax_count(const float* logits, const int64_t* labels, float* acc, int64_t* preds, int B, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input side (the step before the path), on a whole device batch instead of per __getitem__:
 * egb_eeg_window_normalize: x, out [B, C, T] fp32.  mode 0 = common average reference then per-channel z-score
 *   (population std + 1e-8; 1_Data/processed/dual_eeg_dataset.py:158-166), mode 1 = whole-window z-score (:196-198).
 * egb_image_u8_normalize: uint8 [B, H, W, 3] -> fp32 [B, 3, H, W]: ToTensor (/255) + Normalize(mean, std)
 *   (multimodal_dataset.py:73-83); mean3 / std3 are HOST arrays of three floats.
 * ------------------------------------------------------------------------------------------- */
int egb_eeg_window_normalize(const float* x, float* out, int B, int C, int T, int mode, void* stream);
int egb_image_u8_normalize(const uint8_t* hwc, float* chw, int B, int H, int W, const float* mean3, const float* std3,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batch-level auxiliary losses of DualEEGTransformer (dual_eeg_transformer.py:1255-1371); the similarity matrices
 * themselves are egb_gemm calls.  Every reduction stays on the device (no host read, CUDA-graph capturable).
 * egb_l2norm_rows_*  : F.normalize(x, dim=-1) (eps 1e-12) and its backward; inv_norm[r] < 0 marks a clamped row.
 * egb_infonce_rows   : F.cross_entropy(sim[B,N], arange(B)) -> *loss; sim is OVERWRITTEN with d loss / d sim (:1301-1302).
 * egb_supcon_rows    : supervised contrastive loss over sim[B,B] (exp without max subtraction, self pairs masked,
 *                      -log(pos/(all+1e-8)+1e-8), mean over the rows that have a positive, 0 if none; :1336-1371);
 *                      sim is OVERWRITTEN with d loss / d sim; stats: 3*B floats, acc2: 2 floats of scratch.
 * egb_mse_loss       : F.mse_loss(a, b) -> *loss, da = d loss / d a (= -d loss / d b) (:1255-1260).
 * ------------------------------------------------------------------------------------------- */
int egb_l2norm_rows_fwd(const float* x, float* y, float* inv_norm, int rows, int D, float eps, void* stream);
int egb_l2norm_rows_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int rows, int D, void* stream);
int egb_infonce_rows(float* sim, float* loss, int B, int N, void* stream);
int egb_supcon_rows(float* sim, const int64_t* labels, float* stats, float* acc2, float* loss, int B, void* stream);
int egb_mse_loss(const float* a, const float* b, float* da, float* loss, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EYEGAZE_B200_H_ */
