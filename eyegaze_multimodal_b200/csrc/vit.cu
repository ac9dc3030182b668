// Gaze-branch prologue kernels (early_fusion_vit.py:149-196 `_fuse_inputs` + the timm patch embedding):
// the two fp32 heat-maps are fused (concat / add / subtract / |a-b| / instance-normalised product), cut into
// 16x16 patches and written once, in the GEMM operand dtype, as the [B*n_patches, C*ps*ps] matrix the
// patch-embedding GEMM consumes -- the 6-channel concatenated image is never materialised.
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);

namespace {

// stats[(b*3+c)*2 + {0,1}] = mean, 1/(unbiased std + 1e-6) of a*b over H*W   (early_fusion_vit.py:189-193)
__global__ void __launch_bounds__(256) prod_stats_kernel(const float* __restrict__ a0, const float* __restrict__ b0,
                                                         float* __restrict__ stats, int HW, long long a_bs, long long b_bs) {
  __shared__ float red[8];
  const int bi = blockIdx.x / 3, ci = blockIdx.x % 3;
  const float* a = a0 + (long long)bi * a_bs + (long long)ci * HW;
  const float* b = b0 + (long long)bi * b_bs + (long long)ci * HW;
  const long long base = 0;
  float s = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) s += a[base + i] * b[base + i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float mean = tot / (float)HW;
  __syncthreads();
  float q = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) { const float d = a[base + i] * b[base + i] - mean; q += d * d; }
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tq = 0.f;
    for (int w = 0; w < 8; ++w) tq += red[w];
    stats[blockIdx.x * 2] = mean;
    stats[blockIdx.x * 2 + 1] = 1.f / (sqrtf(tq / (float)(HW - 1)) + 1e-6f);
  }
}

// mode: 0 concat(6ch) 1 add 2 subtract 3 subtract_abs 4 multiply 5 single image (a only)
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ stats,
                                T* __restrict__ out, int B, int H, int W, int ps, int mode, long long a_bs,
                                long long b_bs) {
  const int gh = H / ps, gw = W / ps, np = gh * gw;
  const int cin = mode == 0 ? 6 : 3;
  const int K = cin * ps * ps;
  const long long total = (long long)B * np * (K / 4);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % (K / 4)) * 4;
    const long long row = idx / (K / 4);
    const int patch = (int)(row % np), bi = (int)(row / np);
    const int c = k / (ps * ps), kh = (k % (ps * ps)) / ps, kw = k % ps;
    const int y = (patch / gw) * ps + kh, x = (patch % gw) * ps + kw;
    const int ci = c % 3;
    const long long in_img = ((long long)ci * H + y) * W + x;
    const long long off_a = (long long)bi * a_bs + in_img, off_b = (long long)bi * b_bs + in_img;
    float va[4], vb[4], v[4];
    if (mode == 0) {
      if (c < 3) ld4(a + off_a, v); else ld4(b + off_b, v);
    } else if (mode == 5) {
      ld4(a + off_a, v);
    } else {
      ld4(a + off_a, va);
      ld4(b + off_b, vb);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (mode == 1) v[j] = (va[j] + vb[j]) * 0.5f;
        else if (mode == 2) v[j] = (va[j] - vb[j]) * 0.5f;
        else if (mode == 3) v[j] = fabsf(va[j] - vb[j]);
        else v[j] = (va[j] * vb[j] - stats[(bi * 3 + ci) * 2]) * stats[(bi * 3 + ci) * 2 + 1];
      }
    }
    st4(out + row * K + k, v);
  }
}

// out[s, 0, :] = cls + pos[0, :]
template <typename T>
__global__ void fill_row0_kernel(const float* __restrict__ cls, const float* __restrict__ pos, T* __restrict__ out, int S,
                                 int L, int D) {
  const long long total = (long long)S * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D), s = (int)(i / D);
    out[(long long)s * L * D + d] = from_f<T>(cls[d] + pos[d]);
  }
}

}  // namespace

extern "C" {

int egb_vit_patchify(const float* img_a, const float* img_b, int64_t a_batch_stride, int64_t b_batch_stride, void* out,
                     float* stats_scratch, int dtype, int B, int H, int W, int ps, int mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(H % ps == 0 && W % ps == 0 && ps % 4 == 0 && W % 4 == 0, "patchify: image %dx%d / patch %d unsupported", H, W, ps);
  EGB_CHECK(mode >= 0 && mode <= 5, "patchify: bad mode");
  EGB_CHECK(a_batch_stride % 4 == 0 && b_batch_stride % 4 == 0, "patchify: batch strides must be multiples of 4");
  if (mode == 4) {
    EGB_CHECK(stats_scratch != nullptr, "patchify: multiply mode needs a stats buffer of B*3*2 floats");
    prod_stats_kernel<<<B * 3, 256, 0, st>>>(img_a, img_b, stats_scratch, H * W, a_batch_stride, b_batch_stride);
    egb_count_launch(1);
    EGB_LAUNCH_CHECK();
  }
  const int cin = mode == 0 ? 6 : 3;
  const long long total = (long long)B * (H / ps) * (W / ps) * (cin * ps * ps / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (dtype == EGB_BF16)
    patchify_kernel<bf16><<<(int)blocks, 256, 0, st>>>(img_a, img_b, stats_scratch, (bf16*)out, B, H, W, ps, mode,
                                                       a_batch_stride, b_batch_stride);
  else
    patchify_kernel<float><<<(int)blocks, 256, 0, st>>>(img_a, img_b, stats_scratch, (float*)out, B, H, W, ps, mode,
                                                        a_batch_stride, b_batch_stride);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_fill_row0(const float* cls, const float* pos, void* out, int dtype, int S, int L, int D, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)S * D;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (dtype == EGB_BF16)
    fill_row0_kernel<bf16><<<blocks, 256, 0, st>>>(cls, pos, (bf16*)out, S, L, D);
  else
    fill_row0_kernel<float><<<blocks, 256, 0, st>>>(cls, pos, (float*)out, S, L, D);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
