// Input side of the path: the per-window EEG normalisation and the image ToTensor + Normalize that
// the reference's datasets run per sample on the host (numpy / PIL, one __getitem__ at a time), as two HBM-bound
// kernels over a whole device batch.  The host then ships raw windows (and uint8 images: 4x fewer H2D bytes).
//   egb_eeg_window_normalize  mode 0: common average reference + per-channel z-score
//                                     (1_Data/processed/dual_eeg_dataset.py:158-166)
//                             mode 1: whole-window z-score (:196-198)
//   egb_image_u8_normalize    uint8 HWC -> float CHW / 255, (x - mean_c) / std_c  (multimodal_dataset.py:73-83)
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);

namespace {

constexpr int IN_THREADS = 512;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5;
  __syncthreads();                       // red may still be read from a previous call
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < IN_THREADS / 32; ++i) s += red[i];
  return s;
}

// one CTA per window [C, T]; the window is read from L2 after the first pass
__global__ void __launch_bounds__(IN_THREADS) eeg_window_normalize_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                        int C, int T, int mode) {
  extern __shared__ float car[];         // [T] (mode 0)
  __shared__ float red[IN_THREADS / 32];
  const float* xw = x + (long long)blockIdx.x * C * T;
  float* ow = out + (long long)blockIdx.x * C * T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (mode == 0) {
    for (int t = threadIdx.x; t < T; t += IN_THREADS) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += xw[(long long)c * T + t];
      car[t] = s / (float)C;
    }
    __syncthreads();
    for (int c = warp; c < C; c += IN_THREADS / 32) {     // a warp per channel
      const float* xc = xw + (long long)c * T;
      float s = 0.f;
      for (int t = lane; t < T; t += 32) s += xc[t] - car[t];
      const float mean = warp_sum(s) / (float)T;
      float q = 0.f;
      for (int t = lane; t < T; t += 32) {
        const float d = xc[t] - car[t] - mean;
        q = fmaf(d, d, q);
      }
      const float inv = 1.f / (sqrtf(warp_sum(q) / (float)T) + 1e-8f);   // numpy std: population (ddof = 0)
      for (int t = lane; t < T; t += 32) ow[(long long)c * T + t] = (xc[t] - car[t] - mean) * inv;
    }
  } else {
    const long long n = (long long)C * T;
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += IN_THREADS) s += xw[i];
    const float mean = block_sum(s, red) / (float)n;
    float q = 0.f;
    for (long long i = threadIdx.x; i < n; i += IN_THREADS) {
      const float d = xw[i] - mean;
      q = fmaf(d, d, q);
    }
    const float inv = 1.f / (sqrtf(block_sum(q, red) / (float)n) + 1e-8f);
    for (long long i = threadIdx.x; i < n; i += IN_THREADS) ow[i] = (xw[i] - mean) * inv;
  }
}

struct NormArgs {
  float scale[3], shift[3];              // out = u8 * scale_c + shift_c  with scale = 1 / (255 std), shift = -mean / std
};

// thread = 4 consecutive pixels of one image row-major plane position: 12 input bytes, one float4 per output plane
__global__ void __launch_bounds__(256) image_u8_normalize_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                                                 long long pixels_per_image, long long total_quads,
                                                                 NormArgs a) {
  const long long qpi = pixels_per_image >> 2;           // quads per image (pixels_per_image % 4 == 0 is checked)
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total_quads;
       q += (long long)gridDim.x * blockDim.x) {
    const long long b = q / qpi, r = q - b * qpi;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (b * pixels_per_image + 4 * r) * 3);
    const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];        // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
    const float r0 = (float)(w0 & 255u), g0 = (float)((w0 >> 8) & 255u), b0 = (float)((w0 >> 16) & 255u);
    const float r1 = (float)(w0 >> 24), g1 = (float)(w1 & 255u), b1 = (float)((w1 >> 8) & 255u);
    const float r2 = (float)((w1 >> 16) & 255u), g2 = (float)(w1 >> 24), b2 = (float)(w2 & 255u);
    const float r3 = (float)((w2 >> 8) & 255u), g3 = (float)((w2 >> 16) & 255u), b3 = (float)(w2 >> 24);
    float* o = out + b * pixels_per_image * 3 + 4 * r;
    *reinterpret_cast<float4*>(o) = make_float4(fmaf(r0, a.scale[0], a.shift[0]), fmaf(r1, a.scale[0], a.shift[0]),
                                                fmaf(r2, a.scale[0], a.shift[0]), fmaf(r3, a.scale[0], a.shift[0]));
    *reinterpret_cast<float4*>(o + pixels_per_image) =
        make_float4(fmaf(g0, a.scale[1], a.shift[1]), fmaf(g1, a.scale[1], a.shift[1]), fmaf(g2, a.scale[1], a.shift[1]),
                    fmaf(g3, a.scale[1], a.shift[1]));
    *reinterpret_cast<float4*>(o + 2 * pixels_per_image) =
        make_float4(fmaf(b0, a.scale[2], a.shift[2]), fmaf(b1, a.scale[2], a.shift[2]), fmaf(b2, a.scale[2], a.shift[2]),
                    fmaf(b3, a.scale[2], a.shift[2]));
  }
}

}  // namespace

extern "C" {

int egb_eeg_window_normalize(const float* x, float* out, int B, int C, int T, int mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(x && out && B > 0 && C > 0 && T > 0, "eeg_window_normalize: bad arguments");
  EGB_CHECK(mode == 0 || mode == 1, "eeg_window_normalize: mode %d (0 = CAR + channel z-score, 1 = window z-score)", mode);
  const size_t smem = mode == 0 ? (size_t)T * sizeof(float) : 0;
  EGB_CHECK(smem <= 200 * 1024, "eeg_window_normalize: T=%d does not fit the per-window reference buffer", T);
  if (smem > 48 * 1024)
    EGB_CUDA(cudaFuncSetAttribute(eeg_window_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eeg_window_normalize_kernel<<<B, IN_THREADS, smem, st>>>(x, out, C, T, mode);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_image_u8_normalize(const uint8_t* hwc, float* chw, int B, int H, int W, const float* mean3, const float* std3,
                           void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(hwc && chw && B > 0 && H > 0 && W > 0 && mean3 && std3, "image_u8_normalize: bad arguments");
  const long long ppi = (long long)H * W;
  EGB_CHECK(ppi % 4 == 0, "image_u8_normalize: H*W=%lld must be a multiple of 4", ppi);
  EGB_CHECK(((uintptr_t)hwc % 4) == 0 && ((uintptr_t)chw % 16) == 0, "image_u8_normalize: unaligned buffers");
  NormArgs a;
  for (int c = 0; c < 3; ++c) {
    EGB_CHECK(std3[c] > 0.f, "image_u8_normalize: std must be positive");
    a.scale[c] = 1.f / (255.f * std3[c]);
    a.shift[c] = -mean3[c] / std3[c];
  }
  const long long quads = (long long)B * (ppi / 4);
  long long blocks = (quads + 255) / 256;
  const long long cap = (long long)egb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  image_u8_normalize_kernel<<<(int)blocks, 256, 0, st>>>(hwc, chw, ppi, quads, a);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
