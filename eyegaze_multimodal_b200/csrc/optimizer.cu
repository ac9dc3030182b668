// Training-step tail: global-norm gradient clipping + AdamW over ALL parameters in two launches.
// Replaces  torch.nn.utils.clip_grad_norm_(params, max_norm)  followed by  torch.optim.AdamW.step()
// (train_art.py:221-229, train_multimodal_fuzzy_fusion.py:464-472): ~8 foreach passes and a host read of the norm.
// HBM-bound: 16 B read + 12 B written per parameter; the clip coefficient never leaves the device.
//
// Tensors are addressed through two small device tables built by the host (eyegaze_multimodal_b200/optim.py):
//   tensor record t : parameter, gradient (NULL = no gradient this step: skipped, like torch), exp_avg, exp_avg_sq,
//                     the tensor's own step count (bias corrections)
//   chunk record  c : tensor index, element offset, element count (<= 16384)       -- one CTA per chunk
#include "common.cuh"
#include "../../include/eyegaze_b200.h"
#include <string.h>

extern void egb_count_launch(int n);

namespace {

struct TensorRec {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long step;   // 1-based update count of THIS tensor (torch keeps it per parameter: one without gradient lags)
};
struct ChunkRec {
  int32_t tensor;
  int32_t n;
  int64_t offset;
};

__global__ void __launch_bounds__(256) multi_sqnorm_kernel(const TensorRec* __restrict__ T, const ChunkRec* __restrict__ C,
                                                           float* __restrict__ out) {
  __shared__ float red[8];
  const ChunkRec c = C[blockIdx.x];
  const float* g = T[c.tensor].g;
  float acc = 0.f;
  if (g != nullptr) {
    g += c.offset;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
      const int n4 = c.n >> 2;
      for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 v = *reinterpret_cast<const float4*>(g + 4 * i);
        acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
      }
      for (int i = 4 * n4 + threadIdx.x; i < c.n; i += blockDim.x) acc = fmaf(g[i], g[i], acc);
    } else {
      for (int i = threadIdx.x; i < c.n; i += blockDim.x) acc = fmaf(g[i], g[i], acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(out, s);
  }
}

// fp32 master parameters -> their bf16 compute copies, ALL tensors in one launch (same chunk table layout as the
// optimizer kernels).  A training step re-derives ~80 weight copies after every optimizer update; as one cast launch per
// tensor that was ~0.7 ms of launch-bound work inside the captured step for 0.09 ms of traffic.
struct CastRec {
  const float* src;
  bf16* dst;
};
__global__ void __launch_bounds__(256) multi_cast_bf16_kernel(const CastRec* __restrict__ T, const ChunkRec* __restrict__ C) {
  const ChunkRec c = C[blockIdx.x];
  const float* src = T[c.tensor].src + c.offset;
  bf16* dst = T[c.tensor].dst + c.offset;
  if (((reinterpret_cast<uintptr_t>(src) & 15u) | (reinterpret_cast<uintptr_t>(dst) & 7u)) == 0) {
    const int n4 = c.n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(dst + 4 * i) = o;
    }
    for (int i = 4 * n4 + threadIdx.x; i < c.n; i += blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
  } else {
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, max_norm;
  // optional device-resident step state (all may be NULL): with these the launch arguments never change from step to
  // step, so the optimiser tail can live inside a captured CUDA graph and needs no host read or write per step
  const float* lr_dev;       // learning rate of this group (written by egb_lr_schedule_step)
  const float* ctrl;         // egb_adamw_prepare's verdict: {skip this step, steps skipped so far, device step count}
  const float* grad_scale;   // gradients are divided by this before use (torch.amp.GradScaler's scale)
  int use_dev_step;          // bias corrections from ctrl[2] instead of the per-tensor table column
};

// One thread, launched between the norm pass and the update: decides ONCE whether this step is skipped (GradScaler's
// found_inf, or a non-finite global gradient norm) and keeps the counts that make skipped steps invisible to the bias
// corrections, exactly like torch's fused AdamW under a GradScaler (a skipped step does not advance `step`).
__global__ void adamw_prepare_kernel(float* ctrl, const float* sqnorm, const float* grad_scale, const float* found_inf,
                                     int skip_nonfinite, int advance_step) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  bool skip = found_inf != nullptr && *found_inf != 0.f;
  if (!skip && skip_nonfinite && sqnorm != nullptr) {
    const float inv = grad_scale != nullptr ? 1.f / *grad_scale : 1.f;
    skip = !isfinite(sqrtf(*sqnorm) * inv);
  }
  ctrl[0] = skip ? 1.f : 0.f;
  if (skip) ctrl[1] += 1.f;
  else if (advance_step) ctrl[2] += 1.f;
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamArgs& a, float clip, float decay,
                                      float step_size, float bc2_sqrt) {
  g *= clip;
  p *= decay;                                    // decoupled weight decay: p *= 1 - lr * wd
  m = fmaf(a.beta1, m, (1.f - a.beta1) * g);     // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.beta2, v, (1.f - a.beta2) * g * g);
  const float denom = sqrtf(v) / bc2_sqrt + a.eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) multi_adamw_kernel(const TensorRec* __restrict__ T, const ChunkRec* __restrict__ C,
                                                          AdamArgs a, const float* __restrict__ sqnorm) {
  const ChunkRec c = C[blockIdx.x];
  const TensorRec t = T[c.tensor];
  if (t.g == nullptr) return;
  float skipped = 0.f;
  if (a.ctrl != nullptr) {
    if (a.ctrl[0] != 0.f) return;                                     // this step is skipped (egb_adamw_prepare)
    skipped = a.ctrl[1];
  }
  const float inv_scale = a.grad_scale != nullptr ? 1.f / *a.grad_scale : 1.f;
  float clip = inv_scale;
  if (sqnorm != nullptr && a.max_norm > 0.f) {
    const float nrm = sqrtf(*sqnorm) * inv_scale;                     // norm of the UNSCALED gradients
    clip *= fminf(1.f, a.max_norm / (nrm + 1e-6f));                   // clip_grad_norm_
  }
  const float lr = a.lr_dev != nullptr ? *a.lr_dev : a.lr;
  const float decay = 1.f - lr * a.weight_decay;
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {   // bias corrections of this tensor's step, in double like torch's host code
    const double step = a.use_dev_step ? (double)a.ctrl[2] : (double)t.step - (double)skipped;
    s_bc[0] = (float)(1.0 - pow((double)a.beta1, step));
    s_bc[1] = (float)sqrt(1.0 - pow((double)a.beta2, step));
  }
  __syncthreads();
  const float step_size = lr / s_bc[0];
  const float bc2_sqrt = s_bc[1];
  float* p = t.p + c.offset;
  const float* g = t.g + c.offset;
  float* m = t.m + c.offset;
  float* v = t.v + c.offset;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  int done = 0;
  if (vec) {
    const int n4 = c.n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 pv = *reinterpret_cast<float4*>(p + 4 * i);
      const float4 gv = *reinterpret_cast<const float4*>(g + 4 * i);
      float4 mv = *reinterpret_cast<float4*>(m + 4 * i);
      float4 vv = *reinterpret_cast<float4*>(v + 4 * i);
      adam1(pv.x, gv.x, mv.x, vv.x, a, clip, decay, step_size, bc2_sqrt);
      adam1(pv.y, gv.y, mv.y, vv.y, a, clip, decay, step_size, bc2_sqrt);
      adam1(pv.z, gv.z, mv.z, vv.z, a, clip, decay, step_size, bc2_sqrt);
      adam1(pv.w, gv.w, mv.w, vv.w, a, clip, decay, step_size, bc2_sqrt);
      *reinterpret_cast<float4*>(p + 4 * i) = pv;
      *reinterpret_cast<float4*>(m + 4 * i) = mv;
      *reinterpret_cast<float4*>(v + 4 * i) = vv;
    }
    done = 4 * n4;
  }
  for (int i = done + threadIdx.x; i < c.n; i += blockDim.x) {
    float pv = p[i], mv = m[i], vv = v[i];
    adam1(pv, g[i], mv, vv, a, clip, decay, step_size, bc2_sqrt);
    p[i] = pv; m[i] = mv; v[i] = vv;
  }
}

}  // namespace

extern "C" {

/* cast_table: device array of {const float* src, bf16* dst}; chunk_table as for the optimizer kernels */
int egb_multi_tensor_cast_bf16(const void* cast_table, const void* chunk_table, int n_chunks, void* stream) {
  if (n_chunks <= 0) return 0;
  multi_cast_bf16_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const CastRec*)cast_table, (const ChunkRec*)chunk_table);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_multi_tensor_sqnorm(const void* tensor_table, const void* chunk_table, int n_chunks, float* out_sqnorm,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(tensor_table && chunk_table && out_sqnorm && n_chunks > 0, "multi_tensor_sqnorm: bad arguments");
  multi_sqnorm_kernel<<<n_chunks, 256, 0, st>>>((const TensorRec*)tensor_table, (const ChunkRec*)chunk_table, out_sqnorm);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_multi_tensor_adamw(const void* tensor_table, const void* chunk_table, int n_chunks, float lr, float beta1,
                           float beta2, float eps, float weight_decay, float max_norm, const float* sqnorm,
                           void* stream) {
  egb_adamw_state none;
  memset(&none, 0, sizeof(none));
  return egb_multi_tensor_adamw_ex(tensor_table, chunk_table, n_chunks, lr, beta1, beta2, eps, weight_decay, max_norm,
                                   sqnorm, &none, stream);
}

int egb_adamw_prepare(float* ctrl, const float* sqnorm, const float* grad_scale, const float* found_inf,
                      int skip_nonfinite, int advance_step, void* stream) {
  EGB_CHECK(ctrl != nullptr, "adamw_prepare: ctrl is required");
  adamw_prepare_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ctrl, sqnorm, grad_scale, found_inf, skip_nonfinite, advance_step);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_multi_tensor_adamw_ex(const void* tensor_table, const void* chunk_table, int n_chunks, float lr, float beta1,
                              float beta2, float eps, float weight_decay, float max_norm, const float* sqnorm,
                              const egb_adamw_state* state, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(tensor_table && chunk_table && state && n_chunks > 0, "multi_tensor_adamw: bad arguments");
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
  a.lr_dev = state->lr; a.ctrl = state->ctrl; a.grad_scale = state->grad_scale;
  a.use_dev_step = state->use_device_step && state->ctrl != nullptr;
  multi_adamw_kernel<<<n_chunks, 256, 0, st>>>((const TensorRec*)tensor_table, (const ChunkRec*)chunk_table, a, sqnorm);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
