// Device-side state of the training step around the hot path: learning-rate schedules and the
// loss / metric bookkeeping that the reference's loops do on the host with a .item() synchronisation per value and step
// (train_art.py:224-229: six per step; train_multimodal_fuzzy_fusion.py:507-517: predictions, labels, alphas and five
// losses per step).  Everything here is a few threads of arithmetic on device scalars: the point is that the step needs
// NO host read, so a whole epoch can be enqueued (or replayed as a CUDA graph) without draining the GPU.
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);

namespace {

struct SrcPtrs {
  const float* p[8];
};

__global__ void lr_schedule_kernel(float* sched, float* opt_step, const float* __restrict__ base_lr, float* lr_out, int n,
                                   int kind, float p0, float p1, int advance, int advance_opt) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float t = sched[0];
  if (advance) { t += 1.f; sched[0] = t; }
  if (advance_opt && opt_step != nullptr) opt_step[0] += 1.f;
  for (int g = 0; g < n; ++g) {
    float lr = base_lr[g];
    if (kind == 1) {          // CosineAnnealingLR closed form: eta_min + (base - eta_min) (1 + cos(pi t / T_max)) / 2
      const double f = 0.5 * (1.0 + cos(3.14159265358979323846 * (double)t / (double)p0));
      lr = p1 + (base_lr[g] - p1) * (float)f;
    } else if (kind == 2) {   // linear warm-up over p0 steps, cosine to zero at p1 steps (LambdaLR)
      const float ws = fmaxf(1.f, p0);
      double f;
      if (t < p0) f = (double)t / (double)ws;
      else {
        const double prog = (double)(t - p0) / (double)fmaxf(1.f, p1 - p0);
        f = fmax(0.0, 0.5 * (1.0 + cos(3.14159265358979323846 * prog)));
      }
      lr = base_lr[g] * (float)f;
    }
    lr_out[g] = lr;
  }
}

__global__ void accum_scalars_kernel(SrcPtrs s, int n, float* acc) {
  const int i = threadIdx.x;
  if (i < n) acc[i] += *s.p[i];
  if (i == n) acc[n] += 1.f;
}

__global__ void __launch_bounds__(256) argmax_count_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                           float* acc, long long* preds, int B, int C) {
  __shared__ float red[8];
  float hit = 0.f;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const float* r = logits + (long long)b * C;
    int best = 0;
    float bv = r[0];
    for (int c = 1; c < C; ++c)
      if (r[c] > bv) { bv = r[c]; best = c; }           // first maximum wins, like torch.argmax
    if (preds != nullptr) preds[b] = best;
    hit += (labels[b] == (long long)best) ? 1.f : 0.f;
  }
  hit = warp_sum(hit);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = hit;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    if (s != 0.f) atomicAdd(&acc[0], s);
    if (blockIdx.x == 0) atomicAdd(&acc[1], (float)B);
  }
}

}  // namespace

extern "C" {

int egb_lr_schedule_step(float* sched, float* opt_step, const float* base_lr, float* lr_out, int n_groups, int kind,
                         float p0, float p1, int advance, int advance_opt, void* stream) {
  EGB_CHECK(sched && base_lr && lr_out && n_groups > 0 && kind >= 0 && kind <= 2, "lr_schedule_step: bad arguments");
  lr_schedule_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sched, opt_step, base_lr, lr_out, n_groups, kind, p0, p1, advance,
                                                         advance_opt);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_accum_scalars(const float* const* h_src, int n, float* acc, void* stream) {
  EGB_CHECK(h_src && acc && n > 0 && n <= 8, "accum_scalars: 1..8 device scalars");
  SrcPtrs s;
  for (int i = 0; i < 8; ++i) s.p[i] = i < n ? h_src[i] : nullptr;
  accum_scalars_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(s, n, acc);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_argmax_count(const float* logits, const int64_t* labels, float* acc, int64_t* preds, int B, int C, void* stream) {
  EGB_CHECK(logits && labels && acc && B > 0 && C > 0, "argmax_count: bad arguments");
  const int grid = (B + 255) / 256 < 148 ? (B + 255) / 256 : 148;
  argmax_count_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, (const long long*)labels, acc, (long long*)preds, B, C);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
