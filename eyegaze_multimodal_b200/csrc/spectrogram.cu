// Spectrogram-token prologue (dual_eeg_transformer.py:88-135), memory-bound stages:
//   1. STFT (n_fft window, hop, centre/reflect padding) -> |.| -> first `bins` bins -> log(.+1e-8)
//      as a direct windowed DFT from shared memory (frames are 128 samples: a table-driven DFT beats
//      a library FFT round trip and fuses the magnitude/log).
//   2. Conv2d(1->32,3x3,pad 1) + ReLU + MaxPool2d(2) fused: the 32-channel full-resolution activation
//      (4.46 MB per trial and stream in the reference) is never written; output is channels-last with a
//      zero border, i.e. exactly the operand layout of the implicit-GEMM 3x3 convolution that follows.
//   3. ReLU + AdaptiveAvgPool2d(4,4) (+ flatten in (c, ph, pw) order) after the tensor-core conv.
// Backward kernels produce the conv-1 weight gradient (recomputing the pooled arg-max from the stored
// log-magnitude image) and the pre-activation gradient of conv 2 in the padded layout.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);

namespace {

// ------------------------------------------------------------------------------------------ 1. STFT log-magnitude
// grid = (signals), block = 256.  The signal (T samples) is staged in smem with reflect padding.
// Thread (k, g): frequency bin k = tid % bins, frame group g = tid / bins; the thread accumulates up to STFT_FPT
// frames (g, g + G, g + 2G, ...) of bin k in registers.  Within a warp all lanes share the frames (sample loads are
// smem broadcasts) and differ in k (twiddle loads spread over the banks); the window is folded into the twiddle
// once per sample and reused by all frames of the thread.
constexpr int STFT_FPT = 5;

__global__ void __launch_bounds__(256) stft_logmag_kernel(const float* __restrict__ e1, const float* __restrict__ e2,
                                                          const float* __restrict__ window, float* __restrict__ out,
                                                          int n_sig_per_stream, int T, int n_fft, int hop, int bins,
                                                          int frames) {
  extern __shared__ float sm[];
  float* xs = sm;                       // [T + n_fft]  reflect padded
  float* ct = xs + T + n_fft;           // [n_fft] cos table
  float* st = ct + n_fft;               // [n_fft] sin table
  float* ws = st + n_fft;               // [n_fft] window
  const int sig = blockIdx.x;
  const float* src = sig < n_sig_per_stream ? e1 + (long long)sig * T : e2 + (long long)(sig - n_sig_per_stream) * T;
  const int half = n_fft / 2;
  for (int i = threadIdx.x; i < T + n_fft; i += blockDim.x) {
    int t = i - half;                   // torch.stft(center=True, pad_mode='reflect')
    if (t < 0) t = -t;
    if (t >= T) t = 2 * (T - 1) - t;
    xs[i] = src[t];
  }
  for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)i / (float)n_fft, &s, &c);
    ct[i] = c;
    st[i] = s;
    ws[i] = window[i];
  }
  __syncthreads();
  float* o = out + (long long)sig * bins * frames;
  const int groups = blockDim.x / bins;          // frame groups (4 for 64 bins)
  const int k = threadIdx.x % bins, g = threadIdx.x / bins;
  if (g >= groups) return;
  for (int f0 = g; f0 < frames; f0 += groups * STFT_FPT) {
    float re[STFT_FPT], im[STFT_FPT];
    const float* xf[STFT_FPT];
#pragma unroll
    for (int j = 0; j < STFT_FPT; ++j) {
      re[j] = 0.f;
      im[j] = 0.f;
      const int f = f0 + j * groups;
      xf[j] = xs + (f < frames ? f : f0) * hop;   // out-of-range slots recompute frame f0 and are not stored
    }
    int ph = 0;                                   // (k * n) mod n_fft
#pragma unroll 4
    for (int n = 0; n < n_fft; ++n) {
      const float w = ws[n];
      const float cw = ct[ph] * w, sw = st[ph] * w;
#pragma unroll
      for (int j = 0; j < STFT_FPT; ++j) {
        const float v = xf[j][n];
        re[j] = fmaf(v, cw, re[j]);
        im[j] = fmaf(-v, sw, im[j]);
      }
      ph += k;
      if (ph >= n_fft) ph -= n_fft;
    }
#pragma unroll
    for (int j = 0; j < STFT_FPT; ++j) {
      const int f = f0 + j * groups;
      if (f < frames) o[k * frames + f] = logf(sqrtf(re[j] * re[j] + im[j] * im[j]) + 1e-8f);
    }
  }
}

// Register-tiled variant: a thread owns FOUR bins (k0 + j * bins/4) x STFT_FPT frames, so per sample index it issues
// 1 window + 8 twiddle + 5 sample loads for 40 FMAs (0.35 shared-memory loads per FMA instead of 0.8: ncu had the first
// kernel at 79 % LSU-pipe utilisation, 97 % l1tex throughput).  Every output is accumulated in the same order as above
// (n = 0 .. n_fft-1, one fmaf each), so the two kernels are bit-identical.  blockDim = (bins / 4) x frame groups.
__global__ void __launch_bounds__(256) stft_logmag4_kernel(const float* __restrict__ e1, const float* __restrict__ e2,
                                                           const float* __restrict__ window, float* __restrict__ out,
                                                           int n_sig_per_stream, int T, int n_fft, int hop, int bins,
                                                           int frames) {
  extern __shared__ float sm[];
  float* xs = sm;                       // [T + n_fft]  reflect padded
  float* ct = xs + T + n_fft;           // [n_fft] cos table
  float* st = ct + n_fft;               // [n_fft] sin table
  float* ws = st + n_fft;               // [n_fft] window
  const int sig = blockIdx.x;
  const float* src = sig < n_sig_per_stream ? e1 + (long long)sig * T : e2 + (long long)(sig - n_sig_per_stream) * T;
  const int half = n_fft / 2;
  for (int i = threadIdx.x; i < T + n_fft; i += blockDim.x) {
    int t = i - half;                   // torch.stft(center=True, pad_mode='reflect')
    if (t < 0) t = -t;
    if (t >= T) t = 2 * (T - 1) - t;
    xs[i] = src[t];
  }
  for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)i / (float)n_fft, &s, &c);
    ct[i] = c;
    st[i] = s;
    ws[i] = window[i];
  }
  __syncthreads();
  float* o = out + (long long)sig * bins * frames;
  const int bq = bins >> 2;                        // bin threads
  const int groups = blockDim.x / bq;              // frame groups
  const int k0 = threadIdx.x % bq, g = threadIdx.x / bq;
  if (g >= groups) return;
  for (int f0 = g; f0 < frames; f0 += groups * STFT_FPT) {
    float re[4][STFT_FPT], im[4][STFT_FPT];
    const float* xf[STFT_FPT];
#pragma unroll
    for (int j = 0; j < STFT_FPT; ++j) {
      const int f = f0 + j * groups;
      xf[j] = xs + (f < frames ? f : f0) * hop;   // out-of-range slots recompute frame f0 and are not stored
#pragma unroll
      for (int b = 0; b < 4; ++b) { re[b][j] = 0.f; im[b][j] = 0.f; }
    }
    int ph[4] = {0, 0, 0, 0};                     // (k * n) mod n_fft per bin
#pragma unroll 2
    for (int n = 0; n < n_fft; ++n) {
      const float w = ws[n];
      float v[STFT_FPT];
#pragma unroll
      for (int j = 0; j < STFT_FPT; ++j) v[j] = xf[j][n];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float cw = ct[ph[b]] * w, sw = st[ph[b]] * w;
#pragma unroll
        for (int j = 0; j < STFT_FPT; ++j) {
          re[b][j] = fmaf(v[j], cw, re[b][j]);
          im[b][j] = fmaf(-v[j], sw, im[b][j]);
        }
        ph[b] += k0 + b * bq;
        if (ph[b] >= n_fft) ph[b] -= n_fft;
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = k0 + b * bq;
#pragma unroll
      for (int j = 0; j < STFT_FPT; ++j) {
        const int f = f0 + j * groups;
        if (f < frames) o[k * frames + f] = logf(sqrtf(re[b][j] * re[b][j] + im[b][j] * im[b][j]) + 1e-8f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ 2. conv1 + relu + maxpool
// img: [N, Hh, Ww] fp32.  out: [N, Hh/2 + 2, Wp, 32] channels-last, Wp = Ww/2 + 2, zero border.
// One thread per (pooled position, channel quad); the 4x4 input patch feeding a 2x2 pool window is held
// in registers and reused across the 4 conv outputs.
template <typename T>
__global__ void __launch_bounds__(256) spec_conv1_pool_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                              const float* __restrict__ bias, T* __restrict__ out, int Hh,
                                                              int Ww, int H1, int W1, int Wp, long long out_img_stride,
                                                              unsigned short* __restrict__ amax) {
  extern __shared__ float sm[];
  float* im = sm;                    // [(Hh+2)][(Ww+2)] zero padded
  float* wsm = im + (Hh + 2) * (Ww + 2);  // [32][9] + [32]
  const int n = blockIdx.x;
  const int ldw = Ww + 2;
  for (int i = threadIdx.x; i < (Hh + 2) * ldw; i += blockDim.x) {
    const int y = i / ldw - 1, x = i % ldw - 1;
    im[i] = (y >= 0 && y < Hh && x >= 0 && x < Ww) ? img[(long long)n * Hh * Ww + y * Ww + x] : 0.f;
  }
  for (int i = threadIdx.x; i < 32 * 9 + 32; i += blockDim.x) wsm[i] = i < 288 ? w[i] : bias[i - 288];
  __syncthreads();
  T* o = out + (long long)n * out_img_stride;
  // blockDim is a multiple of 8, so a thread serves the same channel quad for all its positions: its 36 weights and 4
  // biases are read into registers once (with the weights re-read from shared memory for every FMA the kernel was
  // bound by the shared-memory pipe: 355 us for 290 MB)
  // The 144 FMAs per (position, channel quad) bound the kernel by instruction issue; as packed pairs (two channels per
  // fma.rn.f32x2, the same fma per lane in the same order: bit-identical) they are 72.
  const int cq = threadIdx.x & 7;
  float2 wr[2][9], br[2];                        // channel pairs (4 cq, 4 cq + 1), (4 cq + 2, 4 cq + 3)
#pragma unroll
  for (int cp = 0; cp < 2; ++cp) {
    br[cp] = make_float2(wsm[288 + cq * 4 + 2 * cp], wsm[288 + cq * 4 + 2 * cp + 1]);
#pragma unroll
    for (int i = 0; i < 9; ++i) wr[cp][i] = make_float2(wsm[(cq * 4 + 2 * cp) * 9 + i], wsm[(cq * 4 + 2 * cp + 1) * 9 + i]);
  }
  for (int idx = threadIdx.x; idx < H1 * W1 * 8; idx += blockDim.x) {
    const int pos = idx >> 3;
    const int ph = pos / W1, pw = pos % W1;
    float2 patch[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) patch[a][b] = splat2(im[(2 * ph + a) * ldw + 2 * pw + b]);
    float res[4];
    unsigned code = 0;   // per channel 3 bits: 0..3 = which of the 2 x 2 conv outputs won the pool, 4 = relu inactive
#pragma unroll
    for (int cp = 0; cp < 2; ++cp) {
      float best0 = 0.f, best1 = 0.f;  // relu(max(.)) == max(0, .)
      unsigned sel0 = 4u, sel1 = 4u;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float2 a = br[cp];
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) a = fma2(wr[cp][kh * 3 + kw], patch[dy + kh][dx + kw], a);
          if (a.x > best0) { best0 = a.x; sel0 = (unsigned)(dy * 2 + dx); }   // first maximum wins, as in max_pool2d
          if (a.y > best1) { best1 = a.y; sel1 = (unsigned)(dy * 2 + dx); }
        }
      res[2 * cp] = best0;
      res[2 * cp + 1] = best1;
      code |= (sel0 << (6 * cp)) | (sel1 << (6 * cp + 3));
    }
    st4(o + ((long long)(ph + 1) * Wp + (pw + 1)) * 32 + cq * 4, res);
    if (amax != nullptr) amax[((long long)n * H1 * W1 + pos) * 8 + cq] = (unsigned short)code;
  }
  // the zero border of this image (top and bottom rows, first and last column), so that the buffer needs no memset
  {
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
    const int nb = 2 * Wp + 2 * H1;
    for (int idx = threadIdx.x; idx < nb * 8; idx += blockDim.x) {
      const int b = idx >> 3;
      int pos;
      if (b < Wp) pos = b;
      else if (b < 2 * Wp) pos = (H1 + 1) * Wp + (b - Wp);
      else { const int r = (b - 2 * Wp) >> 1; pos = (r + 1) * Wp + (((b - 2 * Wp) & 1) ? Wp - 1 : 0); }
      st4(o + (long long)pos * 32 + (idx & 7) * 4, z);
    }
  }
}

// dW1[c,kh,kw] += sum dP1 * img(arg-max patch); db1[c] += sum dP1 (only where the pooled relu is active).
// dP1: [N, H1+2, Wp, 32] padded layout (values at (h+1, w+1)).  Per-block partial sums -> global atomics.
template <typename T>
__global__ void __launch_bounds__(256) spec_conv1_pool_bwd_kernel(const float* __restrict__ img,
                                                                  const float* __restrict__ w,
                                                                  const float* __restrict__ bias,
                                                                  const T* __restrict__ dout, float* __restrict__ dw,
                                                                  float* __restrict__ db, int N, int Hh, int Ww, int H1,
                                                                  int W1, int Wp, long long out_img_stride) {
  extern __shared__ float sm[];
  const int ldw = Ww + 2;
  float* im = sm;
  float* wsm = im + (Hh + 2) * ldw;   // 288 + 32
  float* acc = wsm + 320;             // 288 + 32 block accumulators
  // this image's dP1 tile, (H1+2) x Wp x 32 in its storage type (16-byte aligned: the floats before it are a multiple of 4)
  T* gsm = reinterpret_cast<T*>(acc + 320 + ((4 - (((Hh + 2) * ldw) & 3)) & 3));
  const int g_vec = (int)(out_img_stride * sizeof(T) / 16);
  for (int i = threadIdx.x; i < 320; i += blockDim.x) {
    wsm[i] = i < 288 ? w[i] : bias[i - 288];
    acc[i] = 0.f;
  }
  __syncthreads();
  // partial sums of this thread's channel stay in registers across all images of the CTA
  float lw[9], lb = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) lw[i] = 0.f;
  const int c = threadIdx.x & 31;
  float wr[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wr[i] = wsm[c * 9 + i];
  const float bc = wsm[288 + c];
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < (Hh + 2) * ldw; i += blockDim.x) {
      const int y = i / ldw - 1, x = i % ldw - 1;
      im[i] = (y >= 0 && y < Hh && x >= 0 && x < Ww) ? img[(long long)n * Hh * Ww + y * Ww + x] : 0.f;
    }
    // the whole gradient tile in one sweep of 16-byte loads (all in flight at once): read position by position from
    // global memory, every warp waited a full DRAM round trip per position (1.39 ms for 268 MB)
    {
      const uint4* src = reinterpret_cast<const uint4*>(dout + (long long)n * out_img_stride);
      uint4* dst = reinterpret_cast<uint4*>(gsm);
      for (int i = threadIdx.x; i < g_vec; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const T* g = gsm;
    // thread -> channel c = tid & 31 (so the 32 lanes of a warp hit 32 different accumulators).  The lane's nine
    // weights live in registers and the 4 x 4 input patch of a pooled position is read ONCE (16 broadcast loads per
    // warp; the first version issued two shared loads per FMA -- weight and sample -- and was bound by them: 991 us).
    for (int pos = threadIdx.x >> 5; pos < H1 * W1; pos += blockDim.x >> 5) {
      const int ph = pos / W1, pw = pos % W1;
      const float go = to_f(g[((long long)(ph + 1) * Wp + (pw + 1)) * 32 + c]);
      float patch[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) patch[a][b] = im[(2 * ph + a) * ldw + 2 * pw + b];
      float best = 0.f;
      int sel = -1;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float a = bc;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) a = fmaf(wr[kh * 3 + kw], patch[dy + kh][dx + kw], a);
          if (a > best) { best = a; sel = dy * 2 + dx; }   // first maximum wins, as in max_pool2d
        }
      const float gs = sel < 0 ? 0.f : go;                // relu inactive: no contribution
      lb += gs;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          // sample under tap (kh, kw) of the winning conv output: a select over the four candidates (registers only)
          const float v0 = patch[kh][kw], v1 = patch[kh][kw + 1], v2 = patch[kh + 1][kw], v3 = patch[kh + 1][kw + 1];
          const float v = sel == 0 ? v0 : sel == 1 ? v1 : sel == 2 ? v2 : v3;
          lw[kh * 3 + kw] = fmaf(gs, v, lw[kh * 3 + kw]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) atomicAdd(&acc[c * 9 + i], lw[i]);
  atomicAdd(&acc[288 + c], lb);
  __syncthreads();
  for (int i = threadIdx.x; i < 320; i += blockDim.x) {
    if (i < 288) atomicAdd(dw + i, acc[i]);
    else atomicAdd(db + (i - 288), acc[i]);
  }
}

// Same gradients from the arg-max record the forward kernel wrote (3 bits per pooled output): no conv recomputation --
// the recomputing kernel above executes 935 M warp instructions per cfg2 step (36 FMAs + 27 selects per pooled output
// to find the winner again) and was bound by instruction issue at 1.5 ms.  Lane = channel; the nine samples under the
// winning conv output are read from the shared-memory image at a lane-dependent offset (<= 4 distinct addresses per
// warp instruction).
template <typename T>
__global__ void __launch_bounds__(256) spec_conv1_pool_bwd_rec_kernel(const float* __restrict__ img, const T* __restrict__ dout,
                                                                      const unsigned short* __restrict__ amax,
                                                                      float* __restrict__ dw, float* __restrict__ db, int N,
                                                                      int Hh, int Ww, int H1, int W1, int Wp,
                                                                      long long out_img_stride) {
  extern __shared__ float sm[];
  const int ldw = Ww + 2;
  float* im = sm;
  float* acc = im + (Hh + 2) * ldw;
  T* gsm = reinterpret_cast<T*>(acc + 320 + ((4 - (((Hh + 2) * ldw) & 3)) & 3));
  unsigned short* asm_ = reinterpret_cast<unsigned short*>(gsm + out_img_stride);
  const int g_vec = (int)(out_img_stride * sizeof(T) / 16);
  const int a_vec = H1 * W1;                     // 8 x uint16 = one 16-byte vector per position
  for (int i = threadIdx.x; i < 320; i += blockDim.x) acc[i] = 0.f;
  float lw[9], lb = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) lw[i] = 0.f;
  const int c = threadIdx.x & 31;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < (Hh + 2) * ldw; i += blockDim.x) {
      const int y = i / ldw - 1, x = i % ldw - 1;
      im[i] = (y >= 0 && y < Hh && x >= 0 && x < Ww) ? img[(long long)n * Hh * Ww + y * Ww + x] : 0.f;
    }
    {
      const uint4* src = reinterpret_cast<const uint4*>(dout + (long long)n * out_img_stride);
      uint4* dst = reinterpret_cast<uint4*>(gsm);
      for (int i = threadIdx.x; i < g_vec; i += blockDim.x) dst[i] = src[i];
      const uint4* sa = reinterpret_cast<const uint4*>(amax + (long long)n * H1 * W1 * 8);
      uint4* da = reinterpret_cast<uint4*>(asm_);
      for (int i = threadIdx.x; i < a_vec; i += blockDim.x) da[i] = sa[i];
    }
    __syncthreads();
    for (int pos = threadIdx.x >> 5; pos < H1 * W1; pos += blockDim.x >> 5) {
      const int ph = pos / W1, pw = pos - ph * W1;
      const unsigned sel = ((unsigned)asm_[pos * 8 + (c >> 2)] >> (3 * (c & 3))) & 7u;
      const float go = sel < 4u ? to_f(gsm[((ph + 1) * Wp + (pw + 1)) * 32 + c]) : 0.f;
      const float* src = im + (2 * ph + (int)((sel >> 1) & 1u)) * ldw + 2 * pw + (int)(sel & 1u);
      lb += go;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) lw[kh * 3 + kw] = fmaf(go, src[kh * ldw + kw], lw[kh * 3 + kw]);
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) atomicAdd(&acc[c * 9 + i], lw[i]);
  atomicAdd(&acc[288 + c], lb);
  __syncthreads();
  for (int i = threadIdx.x; i < 320; i += blockDim.x) {
    if (i < 288) atomicAdd(dw + i, acc[i]);
    else atomicAdd(db + (i - 288), acc[i]);
  }
}

// ------------------------------------------------------------------------------------------ 3. relu + adaptive avgpool
// y: conv-2 pre-activation in the padded layout [N, H1+2, Wp, 64] (value of (h,w) at (h+1,w+1)).
// out: [N, 64*16] with index c*16 + ph*4 + pw  (== flatten of (64,4,4)).
// 8 channels per thread (128-bit accesses for bf16, 2 x 128-bit for fp32); grid-stride over (image, bin, channel group)
template <typename T>
__global__ void __launch_bounds__(256) relu_avgpool_kernel(const T* __restrict__ y, T* __restrict__ out, int N, int H1,
                                                           int W1, int Wp, long long img_stride) {
  const long long total = (long long)N * 16 * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx & 7), bin = (int)((idx >> 3) & 15);
    const long long n = idx >> 7;
    const int ph = bin >> 2, pw = bin & 3;
    const int h0 = (ph * H1) / 4, h1 = ((ph + 1) * H1 + 3) / 4;
    const int w0 = (pw * W1) / 4, w1 = ((pw + 1) * W1 + 3) / 4;
    const T* yi = y + n * img_stride + cg * 8;
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.f;
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) {
        float v[8];
        ld8(yi + ((long long)(h + 1) * Wp + (w + 1)) * 64, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += fmaxf(v[i], 0.f);
      }
    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
    T* o = out + n * 1024 + (cg * 8) * 16 + bin;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i * 16] = from_f<T>(a[i] * inv);
  }
}

// dy (padded layout, zero on the border and where relu is inactive) from dpooled [N, 1024].
// One CTA per image: its 1024 pooled gradients are staged in shared memory as [bin][channel] fp32, already divided by
// the bin's element count, so the per-position work is one 16-byte load of y, <= 4 x 2 float4 shared loads and one
// 16-byte store (the first version gathered the 2-byte pooled gradients straight from global memory with a 32-byte
// stride, eight per bin and thread: 778 us for 1.07 GB).
template <typename T> __device__ __forceinline__ float round_to(float x);
template <> __device__ __forceinline__ float round_to<float>(float x) { return x; }
template <> __device__ __forceinline__ float round_to<bf16>(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <typename T>
__global__ void __launch_bounds__(256) relu_avgpool_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dpool,
                                                               T* __restrict__ dy, int N, int H1, int W1, int Wp, int Hp,
                                                               long long img_stride, float* __restrict__ db) {
  __shared__ __align__(16) float g_s[16 * 64];   // [ph*4+pw][c], scaled by 1 / bin size
  __shared__ float cs_s[8][64];                  // per-warp channel sums of the stored gradient (bias gradient of the conv)
  __shared__ int hb[4][2], wb[4][2];
  const int n = blockIdx.x;
  if (threadIdx.x < 4) {
    hb[threadIdx.x][0] = (threadIdx.x * H1) / 4;
    hb[threadIdx.x][1] = ((threadIdx.x + 1) * H1 + 3) / 4;
    wb[threadIdx.x][0] = (threadIdx.x * W1) / 4;
    wb[threadIdx.x][1] = ((threadIdx.x + 1) * W1 + 3) / 4;
  }
  __syncthreads();
  {
    // thread t reads 4 consecutive pooled gradients (channel c = t / 4, bins 4 (t % 4) ..): coalesced
    float v[4];
    ld4(dpool + (long long)n * 1024 + threadIdx.x * 4, v);
    const int c = threadIdx.x >> 2, ph = threadIdx.x & 3;
#pragma unroll
    for (int pw = 0; pw < 4; ++pw)
      g_s[(ph * 4 + pw) * 64 + c] = v[pw] / (float)((hb[ph][1] - hb[ph][0]) * (wb[pw][1] - wb[pw][0]));
  }
  __syncthreads();
  const int per_img = Hp * Wp * 8;
  float cs[8];                                   // a thread's channel group (rem & 7) is fixed: blockDim.x % 8 == 0
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = 0.f;
  for (int rem = threadIdx.x; rem < per_img; rem += blockDim.x) {
    const int cg = rem & 7, pos = rem >> 3;
    const int h = pos / Wp - 1, w = pos % Wp - 1;
    const long long off = (long long)n * img_stride + (long long)pos * 64 + cg * 8;
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = 0.f;
    if (h >= 0 && h < H1 && w >= 0 && w < W1) {
      float v[8];
      ld8(y + off, v);
      // adaptive bins may overlap when H1 % 4 != 0: sum over every bin containing (h, w)
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) {
        if (h < hb[ph][0] || h >= hb[ph][1]) continue;
#pragma unroll
        for (int pw = 0; pw < 4; ++pw) {
          if (w < wb[pw][0] || w >= wb[pw][1]) continue;
          const float4 a = *reinterpret_cast<const float4*>(&g_s[(ph * 4 + pw) * 64 + cg * 8]);
          const float4 b = *reinterpret_cast<const float4*>(&g_s[(ph * 4 + pw) * 64 + cg * 8 + 4]);
          g[0] += a.x; g[1] += a.y; g[2] += a.z; g[3] += a.w;
          g[4] += b.x; g[5] += b.y; g[6] += b.z; g[7] += b.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        g[i] = v[i] > 0.f ? round_to<T>(g[i]) : 0.f;       // the column sums are those of the STORED values
        cs[i] += g[i];
      }
    }
    st8(dy + off, g);
  }
  if (db != nullptr) {
    // bias gradient of spec_conv[3] = column sums of dy: lanes with equal (lane & 7) hold the same channel group
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
      cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) cs_s[warp][lane * 8 + i] = cs[i];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += cs_s[w][threadIdx.x];
      atomicAdd(db + threadIdx.x, t);
    }
  }
}

}  // namespace

extern "C" {

int egb_stft_logmag(const float* eeg1, const float* eeg2, const float* window, float* out, int n_sig_per_stream, int T,
                    int n_fft, int hop, int bins, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(n_fft >= 2 && n_fft % 2 == 0 && bins <= n_fft / 2 + 1 && hop > 0 && T > n_fft / 2,
            "stft: unsupported configuration n_fft=%d hop=%d bins=%d T=%d", n_fft, hop, bins, T);
  const int frames = 1 + T / hop;
  const size_t smem = sizeof(float) * ((size_t)T + 4 * n_fft);
  EGB_CHECK(smem <= 200 * 1024, "stft: window too long for shared memory");
  EGB_CHECK(bins <= 256, "stft: at most 256 frequency bins");
  static size_t smem_set = 0;
  if (smem > smem_set) {
    EGB_CUDA(cudaFuncSetAttribute(stft_logmag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  static const int tiled = getenv("EGB_STFT_TILED") ? atoi(getenv("EGB_STFT_TILED")) : 1;
  const int groups4 = (frames + STFT_FPT - 1) / STFT_FPT;
  if (tiled && bins % 4 == 0 && (bins / 4) * groups4 <= 256 && (bins / 4) * groups4 >= 32) {
    static size_t smem_set4 = 0;
    if (smem > smem_set4) {
      EGB_CUDA(cudaFuncSetAttribute(stft_logmag4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set4 = smem;
    }
    stft_logmag4_kernel<<<2 * n_sig_per_stream, (bins / 4) * groups4, smem, st>>>(eeg1, eeg2, window, out, n_sig_per_stream, T,
                                                                                n_fft, hop, bins, frames);
  } else {
    stft_logmag_kernel<<<2 * n_sig_per_stream, 256, smem, st>>>(eeg1, eeg2, window, out, n_sig_per_stream, T, n_fft, hop,
                                                                bins, frames);
  }
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

/* out must hold N * (H1+2) * Wp * 32 elements (+ slack rows for the implicit-GEMM over-read); this call zeroes it. */
int egb_spec_conv1_pool_fwd(const float* img, const float* w, const float* bias, void* out, int dtype, int N, int Hh,
                            int Ww, int64_t out_elems, uint16_t* amax, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int H1 = Hh / 2, W1 = Ww / 2, Wp = W1 + 2;
  EGB_CHECK(H1 > 0 && W1 > 0, "spec_conv1: image too small");
  const size_t esz = dtype == EGB_BF16 ? 2 : 4;
  const size_t smem = sizeof(float) * ((size_t)(Hh + 2) * (Ww + 2) + 320);
  const long long istr = (long long)(H1 + 2) * Wp * 32;
  // the kernel writes every position of every image (borders included); only the slack rows behind the last image are cleared
  EGB_CHECK(out_elems >= (long long)N * istr, "spec_conv1: output buffer smaller than N padded images");
  if (out_elems > (long long)N * istr)
    EGB_CUDA(cudaMemsetAsync((char*)out + (size_t)N * istr * esz, 0, (size_t)(out_elems - (long long)N * istr) * esz, st));
  if (dtype == EGB_BF16)
    spec_conv1_pool_kernel<bf16><<<N, 256, smem, st>>>(img, w, bias, (bf16*)out, Hh, Ww, H1, W1, Wp, istr, amax);
  else
    spec_conv1_pool_kernel<float><<<N, 256, smem, st>>>(img, w, bias, (float*)out, Hh, Ww, H1, W1, Wp, istr, amax);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

/* dw (288) and db (32) are accumulated into. */
int egb_spec_conv1_pool_bwd(const float* img, const float* w, const float* bias, const void* dout, int dtype, float* dw,
                            float* db, int N, int Hh, int Ww, const uint16_t* amax, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int H1 = Hh / 2, W1 = Ww / 2, Wp = W1 + 2;
  const long long istr = (long long)(H1 + 2) * Wp * 32;
  if (amax != nullptr) {
    const size_t smem_r = sizeof(float) * ((size_t)(Hh + 2) * (Ww + 2) + 320 + 4) + (size_t)istr * (dtype == EGB_BF16 ? 2 : 4) +
                          (size_t)H1 * W1 * 16;
    EGB_CHECK(smem_r <= 200 * 1024, "spec_conv1_pool_bwd: image too large for shared memory");
    static bool attr_r = false;
    if (!attr_r) {
      EGB_CUDA(cudaFuncSetAttribute(spec_conv1_pool_bwd_rec_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      EGB_CUDA(cudaFuncSetAttribute(spec_conv1_pool_bwd_rec_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_r = true;
    }
    const int blocks_r = N < 6 * egb_num_sms() ? N : 6 * egb_num_sms();
    if (dtype == EGB_BF16)
      spec_conv1_pool_bwd_rec_kernel<bf16><<<blocks_r, 256, smem_r, st>>>(img, (const bf16*)dout, amax, dw, db, N, Hh, Ww, H1, W1,
                                                                       Wp, istr);
    else
      spec_conv1_pool_bwd_rec_kernel<float><<<blocks_r, 256, smem_r, st>>>(img, (const float*)dout, amax, dw, db, N, Hh, Ww, H1,
                                                                        W1, Wp, istr);
    egb_count_launch(1);
    EGB_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = sizeof(float) * ((size_t)(Hh + 2) * (Ww + 2) + 640 + 4) + (size_t)istr * (dtype == EGB_BF16 ? 2 : 4);
  EGB_CHECK(smem <= 200 * 1024, "spec_conv1_pool_bwd: image too large for shared memory");
  static bool attr = false;
  if (!attr) {
    EGB_CUDA(cudaFuncSetAttribute(spec_conv1_pool_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    EGB_CUDA(cudaFuncSetAttribute(spec_conv1_pool_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  int blocks = N < 6 * egb_num_sms() ? N : 6 * egb_num_sms();
  if (dtype == EGB_BF16)
    spec_conv1_pool_bwd_kernel<bf16><<<blocks, 256, smem, st>>>(img, w, bias, (const bf16*)dout, dw, db, N, Hh, Ww, H1,
                                                               W1, Wp, istr);
  else
    spec_conv1_pool_bwd_kernel<float><<<blocks, 256, smem, st>>>(img, w, bias, (const float*)dout, dw, db, N, Hh, Ww, H1,
                                                                W1, Wp, istr);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_relu_avgpool_fwd(const void* y, void* out, int dtype, int N, int H1, int W1, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int Wp = W1 + 2;
  const long long istr = (long long)(H1 + 2) * Wp * 64;
  long long want = ((long long)N * 128 + 255) / 256;
  const int grid_fwd = (int)(want < 1 ? 1 : (want > 16LL * egb_num_sms() ? 16LL * egb_num_sms() : want));
  if (dtype == EGB_BF16)
    relu_avgpool_kernel<bf16><<<grid_fwd, 256, 0, st>>>((const bf16*)y, (bf16*)out, N, H1, W1, Wp, istr);
  else
    relu_avgpool_kernel<float><<<grid_fwd, 256, 0, st>>>((const float*)y, (float*)out, N, H1, W1, Wp, istr);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_relu_avgpool_bwd(const void* y, const void* dpool, void* dy, int dtype, int N, int H1, int W1, float* db, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int Wp = W1 + 2, Hp = H1 + 2;
  const long long istr = (long long)Hp * Wp * 64;
  const int grid_bwd = N;                          // one CTA per image
  if (dtype == EGB_BF16)
    relu_avgpool_bwd_kernel<bf16><<<grid_bwd, 256, 0, st>>>((const bf16*)y, (const bf16*)dpool, (bf16*)dy, N, H1, W1, Wp, Hp, istr, db);
  else
    relu_avgpool_bwd_kernel<float><<<grid_bwd, 256, 0, st>>>((const float*)y, (const float*)dpool, (float*)dy, N, H1, W1, Wp, Hp,
                                                      istr, db);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
