// Batch-level auxiliary losses of DualEEGTransformer (dual_eeg_transformer.py:1255-1371), fused row kernels:
//   * compute_symmetry_loss            : mse(cls1, cls2)                                             (:1255-1260)
//   * compute_ibs_alignment_loss       : InfoNCE, cross_entropy(ibs_n . [cls1_n; cls2_n]^T / tau, arange(B))  (:1262-1303)
//   * compute_ibs_contrastive_loss     : supervised contrastive loss over the B x B cosine similarities     (:1305-1371)
// The similarity matrices are GEMMs (egb_gemm, fp32 FFMA kernel: the (B, d) tokens are fp32 tail outputs); what is
// fused here is everything the reference does AROUND them with ~25 ATen launches and five B x B temporaries:
// F.normalize (+ its backward), the row soft-max / masks / log-ratio reductions, and -- in the same pass -- the
// gradient of the loss with respect to the similarity matrix, written in place over it, so the backward pass is
// two GEMMs and nothing else.  All reductions stay on the device (the reference's `has_pos.sum() == 0` early return
// becomes a guarded division), so the losses are CUDA-graph capturable and need no host synchronisation.
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = -INFINITY;
  for (int i = 0; i < nw; ++i) s = fmaxf(s, red[i]);
  return s;
}

// y = x / max(||x||_2, eps) per row (F.normalize); inv[r] = 1 / max(||x||, eps), negated when the clamp was active
__global__ void __launch_bounds__(128) l2norm_rows_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              float* __restrict__ inv, int D, float eps) {
  __shared__ float red[4];
  const float* xr = x + (long long)blockIdx.x * D;
  float* yr = y + (long long)blockIdx.x * D;
  float acc = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) acc = fmaf(xr[i], xr[i], acc);
  const float nrm = sqrtf(block_sum(acc, red));
  const bool clamped = nrm < eps;
  const float s = 1.f / fmaxf(nrm, eps);
  for (int i = threadIdx.x; i < D; i += blockDim.x) yr[i] = xr[i] * s;
  if (threadIdx.x == 0) inv[blockIdx.x] = clamped ? -s : s;
}
// dx = (dy - y (y . dy)) / ||x||   (dy / eps where the clamp was active)
__global__ void __launch_bounds__(128) l2norm_rows_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                              const float* __restrict__ inv, float* __restrict__ dx,
                                                              int D) {
  __shared__ float red[4];
  const long long o = (long long)blockIdx.x * D;
  const float s = inv[blockIdx.x];
  float acc = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) acc = fmaf(dy[o + i], y[o + i], acc);
  const float dot = block_sum(acc, red);
  for (int i = threadIdx.x; i < D; i += blockDim.x)
    dx[o + i] = s < 0.f ? dy[o + i] * (-s) : (dy[o + i] - y[o + i] * dot) * s;
}

// cross_entropy(sim, arange(B)), mean over rows: loss += (lse_i - sim[i][i]) / B;  sim[i][:] <- (softmax_i - onehot_i) / B
__global__ void __launch_bounds__(256) infonce_rows_kernel(float* __restrict__ sim, float* __restrict__ loss, int B, int N) {
  __shared__ float red[8];
  float* row = sim + (long long)blockIdx.x * N;
  const int i = blockIdx.x;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < N; j += blockDim.x) mx = fmaxf(mx, row[j]);
  mx = block_max(mx, red);
  float se = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) se += __expf(row[j] - mx);
  se = block_sum(se, red);
  const float lse = mx + __logf(se);
  const float target = row[i];
  __syncthreads();
  const float invB = 1.f / (float)B, inv_se = 1.f / se;
  for (int j = threadIdx.x; j < N; j += blockDim.x)
    row[j] = (__expf(row[j] - mx) * inv_se - (j == i ? 1.f : 0.f)) * invB;
  if (threadIdx.x == 0) atomicAdd(loss, (lse - target) * invB);
}

// Supervised contrastive loss, pass 1: per row  pos_i = sum_{j != i, y_j == y_i} e^{s_ij},  all_i = sum_{j != i} e^{s_ij}
// (no max subtraction, exactly as the reference: s <= 1 / tau),  l_i = -log(pos_i / (all_i + 1e-8) + 1e-8);
// acc[0] += has_i * l_i, acc[1] += has_i;  stats[i] = {pos_i, all_i, has_i}
__global__ void __launch_bounds__(256) supcon_stats_kernel(const float* __restrict__ sim, const long long* __restrict__ labels,
                                                           float* __restrict__ stats, float* __restrict__ acc, int B) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float* row = sim + (long long)i * B;
  const long long yi = labels[i];
  float pos = 0.f, all = 0.f, npos = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    if (j == i) continue;
    const float e = __expf(row[j]);
    all += e;
    if (labels[j] == yi) { pos += e; npos += 1.f; }
  }
  pos = block_sum(pos, red);
  all = block_sum(all, red);
  npos = block_sum(npos, red);
  if (threadIdx.x == 0) {
    const float has = npos > 0.f ? 1.f : 0.f;
    stats[3 * i] = pos; stats[3 * i + 1] = all; stats[3 * i + 2] = has;
    if (has > 0.f) {
      atomicAdd(&acc[0], -__logf(pos / (all + 1e-8f) + 1e-8f));
      atomicAdd(&acc[1], 1.f);
    }
  }
}
// pass 2: loss = acc[0] / max(acc[1], 1) (0 when no row has a positive);  sim <- dLoss/dsim in place
__global__ void __launch_bounds__(256) supcon_grad_kernel(float* __restrict__ sim, const long long* __restrict__ labels,
                                                          const float* __restrict__ stats, const float* __restrict__ acc,
                                                          float* __restrict__ loss, int B) {
  const int i = blockIdx.x;
  float* row = sim + (long long)i * B;
  const float cnt = fmaxf(acc[1], 1.f);
  if (i == 0 && threadIdx.x == 0) *loss = acc[0] / cnt;
  const float pos = stats[3 * i], all = stats[3 * i + 1], has = stats[3 * i + 2];
  const float den = all + 1e-8f, r = pos / den + 1e-8f;
  const float w = has / cnt;
  const float c_all = w * pos / (r * den * den);     // d l_i / d all_i
  const float c_pos = w / (r * den);                 // -d l_i / d pos_i
  const long long yi = labels[i];
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    float g = 0.f;
    if (j != i) {
      const float e = __expf(row[j]);
      g = e * (c_all - (labels[j] == yi ? c_pos : 0.f));
    }
    row[j] = g;
  }
}

// mean((a - b)^2): loss (accumulated) and d loss / d a = 2 (a - b) / n
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                  float* __restrict__ da, float* __restrict__ loss, long long n) {
  __shared__ float red[8];
  const float inv = 1.f / (float)n;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc = fmaf(d, d, acc);
    da[i] = 2.f * d * inv;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv);
}

}  // namespace

extern "C" {

int egb_l2norm_rows_fwd(const float* x, float* y, float* inv_norm, int rows, int D, float eps, void* stream) {
  EGB_CHECK(x && y && inv_norm && rows > 0 && D > 0, "l2norm_rows_fwd: bad arguments");
  l2norm_rows_fwd_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(x, y, inv_norm, D, eps);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_l2norm_rows_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int rows, int D, void* stream) {
  EGB_CHECK(dy && y && inv_norm && dx && rows > 0 && D > 0, "l2norm_rows_bwd: bad arguments");
  l2norm_rows_bwd_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(dy, y, inv_norm, dx, D);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_infonce_rows(float* sim, float* loss, int B, int N, void* stream) {
  EGB_CHECK(sim && loss && B > 0 && N >= B, "infonce_rows: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  infonce_rows_kernel<<<B, 256, 0, st>>>(sim, loss, B, N);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_supcon_rows(float* sim, const int64_t* labels, float* stats, float* acc2, float* loss, int B, void* stream) {
  EGB_CHECK(sim && labels && stats && acc2 && loss && B > 0, "supcon_rows: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CUDA(cudaMemsetAsync(acc2, 0, 2 * sizeof(float), st));
  supcon_stats_kernel<<<B, 256, 0, st>>>(sim, (const long long*)labels, stats, acc2, B);
  supcon_grad_kernel<<<B, 256, 0, st>>>(sim, (const long long*)labels, stats, acc2, loss, B);
  egb_count_launch(2);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_mse_loss(const float* a, const float* b, float* da, float* loss, int64_t n, void* stream) {
  EGB_CHECK(a && b && da && loss && n > 0, "mse_loss: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  mse_kernel<<<grid, 256, 0, st>>>(a, b, da, loss, (long long)n);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
