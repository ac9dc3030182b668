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
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;
  int32_t major;          /* 0 = K-major, 1 = MN-major */
  int32_t rows_per_group; /* <= 0: a single group holding all rows */
  int64_t row_stride;     /* elements */
  int64_t group_stride;   /* elements */
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
/* backward-side epilogues: multiply the accumulator by act'(.) read from `aux` */
#define EGB_ACTBWD_NONE 0
#define EGB_ACTBWD_RELU_MASK 1 /* aux = forward OUTPUT y (post relu [+dropout]); x *= (y != 0) * aux_scale */
#define EGB_ACTBWD_GELU 2      /* aux = forward PRE-activation; x *= gelu'(aux) */

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
  int32_t accumulate;  /* 1: C += result (C must be fp32); enables split-K with red.global.add */
  int32_t split_k;     /* 0 = auto */
} egb_gemm_desc;

int egb_gemm(const egb_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EYEGAZE_B200_H_ */
