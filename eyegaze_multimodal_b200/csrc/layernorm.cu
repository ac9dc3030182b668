// LayerNorm forward / backward over the last dimension (art.py:283-296,306,328 post-LN blocks, eps 1e-5;
// timm ViT pre-LN, eps 1e-6).  One warp per row, 128-bit loads, row kept in registers (D <= 1024) so the
// tensor is read exactly once per pass; mean / rstd are saved in fp32 for the backward.
#include <stdlib.h>
#include <stdint.h>
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);

namespace {

constexpr int LN_MAX_CHUNKS = 4;  // 4 chunks x 32 lanes x 8 elements = D up to 1024

// eight consecutive elements in their storage format (bf16: one 16-byte register quad; fp32: two)
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 q;
  __device__ __forceinline__ void load(const bf16* p) { q = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// ROWS rows per warp and iteration: their loads are issued back to back before any arithmetic, so a warp keeps ROWS x D
// elements in flight instead of D (one 1.5 KB row per warp left the ViT-B LayerNorm at 3.4 TB/s).
template <typename T, int ROWS>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, T* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int M, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int chunks = (D + 255) / 256;
  const int stride = gridDim.x * warps_per_block * ROWS;
  for (int row0 = (blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * ROWS; row0 < M; row0 += stride) {
    // rows stay in their STORAGE format (bf16: 4 registers per 8 elements) between the passes and are unpacked on use
    Raw8<T> raw[ROWS][LN_MAX_CHUNKS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = row0 + r;
#pragma unroll
      for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
        const int d = c * 256 + lane * 8;
        if (row < M && c < chunks && d < D) raw[r][c].load(x + (long long)row * D + d);
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = row0 + r;
      if (row >= M) break;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
        const int d = c * 256 + lane * 8;
        if (c < chunks && d < D) {
          float t8[8];
          raw[r][c].unpack(t8);
#pragma unroll
          for (int j = 0; j < 8; ++j) s += t8[j];
        }
      }
      const float mean = warp_sum(s) / (float)D;
      float q = 0.f;
#pragma unroll
      for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
        const int d = c * 256 + lane * 8;
        if (c < chunks && d < D) {
          float t8[8];
          raw[r][c].unpack(t8);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float t = t8[j] - mean; q += t * t; }
        }
      }
      const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
      for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
        const int d = c * 256 + lane * 8;
        if (c < chunks && d < D) {
          float g[8], b[8], o[8], t8[8];
          raw[r][c].unpack(t8);
          ld8(gamma + d, g);
          ld8(beta + d, b);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (t8[j] - mean) * rstd * g[j] + b[j];
          st8(y + (long long)row * D + d, o);
        }
      }
      if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma += sum dy*xhat; dbeta += sum dy.
// One warp per row.  The row's x and dy stay in registers in their STORAGE format between the statistics pass and
// the dx pass and are converted twice: with fp32 copies of xhat and g the kernel sat at the 128-register limit of
// two CTAs per SM and ptxas serialised the row's loads chunk by chunk (three DRAM round trips per row: 81 us for the
// ViT-B LayerNorms against 36 us of traffic).  gamma is read from shared memory.
// Threads per CTA / CTAs per SM by row width: the column accumulators (dgamma, dbeta, dx column sums: 3 x 8 floats per
// 256-column chunk) live in registers, which caps the CTA size for the wide rows (D = 768: 12 warps with <= 168 registers).
template <int C> struct LnBwdShape {
  static constexpr int kThreads = C == 1 ? 256 : (C == 2 ? 512 : 384);
  static constexpr int kCtasPerSm = C == 1 ? 2 : 1;
};

template <typename T, int C>
__global__ void __launch_bounds__(LnBwdShape<C>::kThreads, LnBwdShape<C>::kCtasPerSm) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, T* __restrict__ dx,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            const T* __restrict__ dres, float* __restrict__ dx_colsum,
                                                            int M, int D) {
  extern __shared__ float red[];  // [3][D] column partials (dgamma, dbeta, dx column sums), [D] gamma
  float* sg = red + 3 * D;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) sg[i] = gamma[i];
  __syncthreads();
  float ag[C][8], ab[C][8], ac[C][8];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[c][j] = ab[c][j] = ac[c][j] = 0.f;

  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
    Raw8<T> xr[C], dr[C], rr[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int d = c * 256 + lane * 8;
      if (d < D) {
        xr[c].load(x + (long long)row * D + d);
        dr[c].load(dy + (long long)row * D + d);
        if (dres != nullptr) rr[c].load(dres + (long long)row * D + d);   // residual-path gradient, same round trip
      }
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int d = c * 256 + lane * 8;
      if (d < D) {
        float xv[8], dv[8];
        xr[c].unpack(xv);
        dr[c].unpack(dv);
        const float4 g0 = *reinterpret_cast<const float4*>(sg + d), g1 = *reinterpret_cast<const float4*>(sg + d + 4);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = dv[j] * gm[j];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          ag[c][j] = fmaf(dv[j], xh, ag[c][j]);
          ab[c][j] += dv[j];
        }
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int d = c * 256 + lane * 8;
      if (d < D) {
        float xv[8], dv[8], o[8];
        xr[c].unpack(xv);
        dr[c].unpack(dv);
        const float4 g0 = *reinterpret_cast<const float4*>(sg + d), g1 = *reinterpret_cast<const float4*>(sg + d + 4);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          o[j] = rstd * (fmaf(dv[j], gm[j], -s1) - xh * s2);
        }
        if (dres != nullptr) {   // gradient that reaches x around the normalisation (the residual connection)
          float rv[8];
          rr[c].unpack(rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rv[j];
        }
        st8(dx + (long long)row * D + d, o);
        // column sums of the STORED dx (rounded to the storage type, as a separate pass over dx would see them): the bias
        // gradient of the Linear whose output fed this residual stream
#pragma unroll
        for (int j = 0; j < 8; ++j) ac[c][j] += to_f(from_f<T>(o[j]));
      }
    }
  }
  // block reduction of the per-warp column partials, then one atomic per column per block
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int d = c * 256 + lane * 8;
    if (d < D) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&red[d + j], ag[c][j]);
        atomicAdd(&red[D + d + j], ab[c][j]);
        if (dx_colsum != nullptr) atomicAdd(&red[2 * D + d + j], ac[c][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[D + i]);
    if (dx_colsum != nullptr) atomicAdd(dx_colsum + i, red[2 * D + i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Staged variant (bf16): the rows of x, dy and the residual gradient are contiguous in memory, so a producer thread
// streams them into a shared-memory ring with 1-D bulk copies (cp.async.bulk + mbarrier, LN_ROWS rows per tensor and
// stage) and the compute warps only ever wait on shared memory.  The register version above keeps one row per warp in
// flight -- 12 warps x 4.6 KB per SM against ~1.2 us of DRAM latency = 3.3 TB/s measured (the column accumulators take
// 72 registers per thread, which rules out a second row or more warps); the ring holds LN_STAGES x 12 rows per SM
// independently of the register file.  A warp copies its row from the ring into registers and releases the slot at
// once, so the ring is purely a latency-hiding FIFO.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LN_ROWS = 11;      // rows per stage = compute warps per CTA (+ the producer warp = 12 warps: 168 registers each)
constexpr int LN_STAGES = 3;

__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ln_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ln_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ln_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ln_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ln_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ln_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ln_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LN_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra LN_WAIT_%=;\n\t}" ::"r"(ln_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void ln_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ln_smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(ln_smem_u32(bar))
               : "memory");
}

template <int C>
__global__ void __launch_bounds__((LN_ROWS + 1) * 32, 1)
layernorm_bwd_staged_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in, bf16* __restrict__ dx,
                            float* __restrict__ dgamma, float* __restrict__ dbeta, const bf16* __restrict__ dres,
                            float* __restrict__ dx_colsum, int M, int D) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int ntens = dres != nullptr ? 3 : 2;
  const size_t tile_bytes = (size_t)LN_ROWS * D * sizeof(bf16);            // one tensor, one stage
  uint8_t* ring = ln_smem;                                                  // [LN_STAGES][3][LN_ROWS][D] bf16
  float* red = reinterpret_cast<float*>(ring + (size_t)LN_STAGES * 3 * tile_bytes);   // [3][D] column partials
  float* sg = red + 3 * D;                                                  // [D] gamma
  uint64_t* full = reinterpret_cast<uint64_t*>(sg + D);                     // [LN_STAGES]
  uint64_t* empty = full + LN_STAGES;                                       // [LN_STAGES]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) sg[i] = gamma[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < LN_STAGES; ++s) { ln_mbar_init(&full[s], 1); ln_mbar_init(&empty[s], LN_ROWS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ntiles = (M + LN_ROWS - 1) / LN_ROWS;

  if (warp == LN_ROWS) {
    // ---------------------------------------------------------------- producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        ln_mbar_wait(&empty[s], ph ^ 1u);
        const int r0 = t * LN_ROWS;
        const int rows = min(LN_ROWS, M - r0);
        const uint32_t bytes = (uint32_t)((size_t)rows * D * sizeof(bf16));
        ln_mbar_expect_tx(&full[s], bytes * (uint32_t)ntens);
        uint8_t* dst = ring + (size_t)s * 3 * tile_bytes;
        ln_bulk_load(dst, x + (size_t)r0 * D, bytes, &full[s]);
        ln_bulk_load(dst + tile_bytes, dy + (size_t)r0 * D, bytes, &full[s]);
        if (dres != nullptr) ln_bulk_load(dst + 2 * tile_bytes, dres + (size_t)r0 * D, bytes, &full[s]);
        if (++s == LN_STAGES) { s = 0; ph ^= 1u; }
      }
    }
    return;
  }

  // ------------------------------------------------------------------ compute warps: warp w owns row w of every tile
  float ag[C][8], ab[C][8], ac[C][8];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[c][j] = ab[c][j] = ac[c][j] = 0.f;
  int s = 0;
  uint32_t ph = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int row = t * LN_ROWS + warp;
    const bool live = row < M;
    float mean = 0.f, rstd = 0.f;
    if (live) { mean = mean_in[row]; rstd = rstd_in[row]; }
    ln_mbar_wait(&full[s], ph);
    Raw8<bf16> xr[C], dr[C], rr[C];
    const uint8_t* base = ring + (size_t)s * 3 * tile_bytes + (size_t)warp * D * sizeof(bf16);
    if (live) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int d = c * 256 + lane * 8;
        if (d < D) {
          xr[c].q = *reinterpret_cast<const uint4*>(base + (size_t)d * 2);
          dr[c].q = *reinterpret_cast<const uint4*>(base + tile_bytes + (size_t)d * 2);
          if (dres != nullptr) rr[c].q = *reinterpret_cast<const uint4*>(base + 2 * tile_bytes + (size_t)d * 2);
        }
      }
    }
    __syncwarp();
    if (lane == 0) ln_mbar_arrive(&empty[s]);        // the row is in registers: the slot may be refilled
    if (++s == LN_STAGES) { s = 0; ph ^= 1u; }
    if (!live) continue;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int d = c * 256 + lane * 8;
      if (d < D) {
        float xv[8], dv[8];
        xr[c].unpack(xv);
        dr[c].unpack(dv);
        const float4 g0 = *reinterpret_cast<const float4*>(sg + d), g1 = *reinterpret_cast<const float4*>(sg + d + 4);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = dv[j] * gm[j];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          ag[c][j] = fmaf(dv[j], xh, ag[c][j]);
          ab[c][j] += dv[j];
        }
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int d = c * 256 + lane * 8;
      if (d < D) {
        float xv[8], dv[8], o[8];
        xr[c].unpack(xv);
        dr[c].unpack(dv);
        const float4 g0 = *reinterpret_cast<const float4*>(sg + d), g1 = *reinterpret_cast<const float4*>(sg + d + 4);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          o[j] = rstd * (fmaf(dv[j], gm[j], -s1) - xh * s2);
        }
        if (dres != nullptr) {
          float rv[8];
          rr[c].unpack(rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rv[j];
        }
        st8(dx + (long long)row * D + d, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) ac[c][j] += to_f(from_f<bf16>(o[j]));
      }
    }
  }
  // block reduction of the per-warp column partials (compute warps only; the producer warp has left), then one atomic
  // per column per CTA
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int d = c * 256 + lane * 8;
    if (d < D) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&red[d + j], ag[c][j]);
        atomicAdd(&red[D + d + j], ab[c][j]);
        if (dx_colsum != nullptr) atomicAdd(&red[2 * D + d + j], ac[c][j]);
      }
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(LN_ROWS * 32) : "memory");
  for (int i = threadIdx.x; i < D; i += LN_ROWS * 32) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[D + i]);
    if (dx_colsum != nullptr) atomicAdd(dx_colsum + i, red[2 * D + i]);
  }
}

}  // namespace

extern "C" {

int egb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      int dtype, int M, int D, float eps, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(D % 8 == 0 && D <= 256 * LN_MAX_CHUNKS, "layernorm: D=%d must be a multiple of 8 and <= %d", D,
            256 * LN_MAX_CHUNKS);
  EGB_CHECK(M > 0, "layernorm: empty");
  // two rows per warp in flight measured SLOWER (35.8 vs 33.0 us average over the step's 39 launches): the kernel is not
  // bound by bytes in flight; one row per warp stays the default
  static const int rows2 = getenv("EGB_LN_FWD_ROWS") ? atoi(getenv("EGB_LN_FWD_ROWS")) : 1;
  const int R = (dtype == EGB_BF16 && rows2 == 2) ? 2 : 1;
  int blocks = (M + 8 * R - 1) / (8 * R);
  const int cap = egb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (dtype == EGB_BF16 && R == 2)
    layernorm_fwd_kernel<bf16, 2><<<blocks, 256, 0, st>>>((const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, M, D, eps);
  else if (dtype == EGB_BF16)
    layernorm_fwd_kernel<bf16, 1><<<blocks, 256, 0, st>>>((const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, M, D, eps);
  else
    layernorm_fwd_kernel<float, 1><<<blocks, 256, 0, st>>>((const float*)x, gamma, beta, (float*)y, mean, rstd, M, D, eps);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

/* dgamma / dbeta are ACCUMULATED into (caller zeroes them, or passes live .grad buffers).  dres (optional, same
   shape and dtype as dx) is added to dx: the gradient arriving at x through the residual connection, so that the
   sum of the two paths costs no extra pass. */
int egb_layernorm_bwd_res(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                          void* dx, float* dgamma, float* dbeta, const void* dres, int dtype, int M, int D, void* stream) {
  return egb_layernorm_bwd_ex(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, dres, nullptr, dtype, M, D, stream);
}

/* dx_colsum (optional, [D] fp32, ACCUMULATED): column sums of the stored dx -- the bias gradient of the Linear layer
   whose output was added into the stream this LayerNorm reads (timm `x = x + proj(..)` / `x + fc2(..)`, art.py:293-295),
   taken while dx is in registers instead of a second pass over dx. */
int egb_layernorm_bwd_ex(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                         void* dx, float* dgamma, float* dbeta, const void* dres, float* dx_colsum, int dtype, int M,
                         int D, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(D % 8 == 0 && D <= 256 * LN_MAX_CHUNKS, "layernorm_bwd: unsupported D=%d", D);
  // few, fat CTAs: every CTA ends with 3 D column atomics, and with several hundred CTAs those serialise per address in
  // L2 (a measurable tail)
  const int chunks = (D + 255) / 256;
  {
    // bf16 rows stream through the shared-memory ring (EGB_LN_STAGED=0 keeps the register version)
    static const int staged = getenv("EGB_LN_STAGED") ? atoi(getenv("EGB_LN_STAGED")) : 1;
    const size_t smem_s = (size_t)LN_STAGES * 3 * LN_ROWS * D * 2 + (size_t)4 * D * sizeof(float) + 2 * LN_STAGES * 8;
    const bool aligned = ((uintptr_t)dy % 16 == 0) && ((uintptr_t)x % 16 == 0) && (dres == nullptr || (uintptr_t)dres % 16 == 0);
    // (rows of <= 256 columns: 50 us staged vs 45 us from registers -- one 512-byte row per warp and stage is too little per
    //  barrier round trip; the ring pays off from two 256-column chunks up: ViT-B 85 -> 77 us)
    if (staged && dtype == EGB_BF16 && aligned && chunks >= 2 && M >= 4 * LN_ROWS && smem_s <= 200 * 1024) {
      int blocks = (M + LN_ROWS - 1) / LN_ROWS;
      if (blocks > egb_num_sms()) blocks = egb_num_sms();
#define EGB_LN_ST(CC)                                                                                                   \
  do {                                                                                                                  \
    static bool attr = false;                                                                                           \
    if (!attr) {                                                                                                        \
      EGB_CUDA(cudaFuncSetAttribute(layernorm_bwd_staged_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      attr = true;                                                                                                      \
    }                                                                                                                   \
    layernorm_bwd_staged_kernel<CC><<<blocks, (LN_ROWS + 1) * 32, smem_s, st>>>(                                        \
        (const bf16*)dy, (const bf16*)x, gamma, mean, rstd, (bf16*)dx, dgamma, dbeta, (const bf16*)dres, dx_colsum, M, D); \
  } while (0)
      switch (chunks) {
        case 1: EGB_LN_ST(1); break;
        case 2: EGB_LN_ST(2); break;
        case 3: EGB_LN_ST(3); break;
        default: EGB_LN_ST(4); break;
      }
#undef EGB_LN_ST
      egb_count_launch(1);
      EGB_LAUNCH_CHECK();
      return 0;
    }
  }
  const size_t smem = (size_t)4 * D * sizeof(float);
  const int threads = chunks == 1 ? LnBwdShape<1>::kThreads : (chunks == 2 ? LnBwdShape<2>::kThreads : LnBwdShape<3>::kThreads);
  int blocks = (M + threads / 32 - 1) / (threads / 32);
  const int cap = egb_num_sms() * (chunks == 1 ? LnBwdShape<1>::kCtasPerSm : 1);
  if (blocks > cap) blocks = cap;
#define EGB_LN_BWD(TT, CC)                                                                                         \
  layernorm_bwd_kernel<TT, CC><<<blocks, threads, smem, st>>>((const TT*)dy, (const TT*)x, gamma, mean, rstd, (TT*)dx, \
                                                         dgamma, dbeta, (const TT*)dres, dx_colsum, M, D)
  if (dtype == EGB_BF16) {
    switch (chunks) {
      case 1: EGB_LN_BWD(bf16, 1); break;
      case 2: EGB_LN_BWD(bf16, 2); break;
      case 3: EGB_LN_BWD(bf16, 3); break;
      default: EGB_LN_BWD(bf16, 4); break;
    }
  } else {
    switch (chunks) {
      case 1: EGB_LN_BWD(float, 1); break;
      case 2: EGB_LN_BWD(float, 2); break;
      case 3: EGB_LN_BWD(float, 3); break;
      default: EGB_LN_BWD(float, 4); break;
    }
  }
#undef EGB_LN_BWD
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                      float* dgamma, float* dbeta, int dtype, int M, int D, void* stream) {
  return egb_layernorm_bwd_res(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, nullptr, dtype, M, D, stream);
}

}  // extern "C"
