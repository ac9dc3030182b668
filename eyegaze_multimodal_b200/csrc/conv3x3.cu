// 3x3 convolution over zero-bordered channels-last images as a TRUE implicit GEMM on tcgen05 (64 -> 32 channels: the
// data gradient of the spectrogram CNN's second convolution, dual_eeg_transformer.py:81-86 / its autograd).
//
// The generic GEMM kernel runs this convolution over an overlapping-row VIEW of the image buffer (K = 3 row segments x
// 4 positions x 64 channels): every input element crosses L2 -> shared memory 12 times, 8.5 GB per launch, and the
// launch is bound by exactly that (1.5 ms at 0.13 of the tensor peak).  Here a CTA stages the input rows of a tile ONCE
// ([128 + 2 Wp + 2 positions][64 channels], one TMA box, 128-byte swizzle) and issues the nine taps as MMAs whose A
// descriptors simply START (a Wp + b) rows further down: the hardware applies the 128-byte swizzle to absolute
// shared-memory address bits, so a K-major operand may begin at any row of a swizzled tile (verified on the device by
// csrc/tests/shift_desc_test.cu: exact for every row offset with the descriptor's base-offset field left at 0).
//
//   y[p, c] = sum_{a, b < 3} sum_o x[p - (Wp + 1) + a Wp + b, o] * w[c][a][b][o]        p = flat padded position
//
// Persistent CTAs, 6 warps: warp 0 TMA producer (input tiles through a ring of four; the nine 32 x 64 weight tiles once),
// warp 1 MMA issuer (36 tcgen05.mma per tile: M = 128, N = 32, K = 16; fp32 accumulators double buffered in TMEM),
// warps 2..5 epilogue (a lane owns an output row = 32 channels = 64 contiguous bytes; the tile is one contiguous 8 KB).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);
int egb_tmap_rows64(CUtensorMap* out, const void* ptr, long long inner, long long rows, long long groups, long long rs,
                    long long gs, int box_rows);

namespace {

constexpr int CV_THREADS = 192;
constexpr int CV_M = 128;
constexpr int CV_XS = 4;        // input-tile ring (a tile is ~20 KB; a TMA round trip outlasts a tile's 36 MMAs)

struct ConvParams {
  bf16* y;            // [rows, 32]
  long long M;        // output rows m = 0 .. M-1, written at flat position m + out_shift
  long long out_shift;
  int Wp;             // padded image width (row shift of one kernel row)
  int box_rows;       // input rows per tile = 128 + 2 Wp + 2
  int tiles;
};

__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_c64_c32_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tile_bytes = ((p.box_rows * 128) + 1023) & ~1023;
  uint8_t* sX = smem;                                  // [CV_XS][box_rows][128 B]
  uint8_t* sW = smem + CV_XS * tile_bytes;             // [9][32 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 9 * 4096);
  uint64_t* full = bars;                     // [CV_XS] input tile landed
  uint64_t* empty = bars + CV_XS;            // [CV_XS] input tile consumed by the MMAs
  uint64_t* tfull = bars + 2 * CV_XS;        // [2] accumulator ready
  uint64_t* tempty = bars + 2 * CV_XS + 2;   // [2] accumulator drained
  uint64_t* wbar = bars + 2 * CV_XS + 4;     // weights landed
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2 * CV_XS + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmW);
    for (int i = 0; i < CV_XS; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 4);
    }
    ptx::mbar_init(wbar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<64>(slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(wbar, 9 * 4096);
      for (int t = 0; t < 9; ++t) ptx::tma_load_3d(sW + t * 4096, &tmW, wbar, (t / 3) * 256 + (t % 3) * 64, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        ptx::mbar_wait(&empty[s], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(p.box_rows * 128));
        ptx::tma_load_3d(sX + s * tile_bytes, &tmX, &full[s], 0, tile * CV_M, 0);   // rows past the buffer end: zero-filled
        if (++s == CV_XS) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc_bf16(CV_M, 32, 0, 0);
    ptx::mbar_wait(wbar, 0);
    int s = 0, ts = 0;
    uint32_t ph = 0, tph = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tempty[ts], tph ^ 1u);
      ptx::mbar_wait(&full[s], ph);
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t xa = ptx::smem_u32(sX + s * tile_bytes);
        const uint32_t td = tmem + (uint32_t)(32 * ts);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          // tap (a, b): the A operand starts (a Wp + b) rows into the tile (any row: the swizzle is on absolute address bits)
          const uint64_t ad = ptx::make_smem_desc(xa + (uint32_t)(((t / 3) * p.Wp + (t % 3)) * 128), 16u, 1024u);
          const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(sW + t * 4096), 16u, 1024u);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(td, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (t > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty[s]);     // the input tile may be refilled once these MMAs have retired
        ptx::umma_commit(&tfull[ts]);
      }
      __syncwarp();
      if (++s == CV_XS) { s = 0; ph ^= 1u; }
      if (++ts == 2) { ts = 0; tph ^= 1u; }
    }
  } else {
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access (warps 2..5 -> 2, 3, 0, 1)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tfull[s], ph);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(32 * s), v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_relaxed(&tempty[s]);
      const long long m = (long long)tile * CV_M + quad * 32 + lane;
      if (m < p.M) {
        uint4* dst = reinterpret_cast<uint4*>(p.y + (m + p.out_shift) * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]));
          __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3]));
          __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
          __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
          o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
          dst[g] = o;
        }
      }
      if (++s == 2) { s = 0; ph ^= 1u; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<64>(tmem);
  }
}

}  // namespace

extern "C" {

/* y[(m + out_shift) * 32 + c] = sum over the nine taps (a, b) and 64 input channels o of
       x[(m + a * Wp + b) * 64 + o] * w[c * 768 + a * 256 + b * 64 + o]          for m in [0, M)
   x: bf16 [x_rows, 64] (zero-bordered channels-last images, flat), w: bf16 [32, 768] (three 256-wide row segments of four
   64-wide slots, the fourth unused), y: bf16 rows of 32.  Reads x rows up to M + 2 Wp + 2 (rows >= x_rows read as 0). */
int egb_conv3x3_c64_c32(const void* x, long long x_rows, const void* w, void* y, long long M, long long out_shift, int Wp,
                        void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(M > 0 && Wp >= 3 && 128 + 2 * Wp + 2 <= 256, "conv3x3: unsupported geometry (M=%lld, Wp=%d)", M, Wp);
  EGB_CHECK(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)y % 16) == 0 && (out_shift * 64) % 16 == 0,
            "conv3x3: misaligned buffers");
  ConvParams p;
  p.y = (bf16*)y;
  p.M = M;
  p.out_shift = out_shift;
  p.Wp = Wp;
  p.box_rows = 128 + 2 * Wp + 2;
  p.tiles = (int)((M + CV_M - 1) / CV_M);
  CUtensorMap mx, mw;
  if (egb_tmap_rows64(&mx, x, 64, x_rows, 1, 64, 0, p.box_rows)) return 1;
  if (egb_tmap_rows64(&mw, w, 768, 32, 1, 768, 0, 32)) return 1;
  const int tile_bytes = ((p.box_rows * 128) + 1023) & ~1023;
  const size_t smem = (size_t)CV_XS * tile_bytes + 9 * 4096 + 256 + 1024;
  static bool attr = false;
  if (!attr) {
    EGB_CUDA(cudaFuncSetAttribute(conv3x3_c64_c32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const int grid = p.tiles < egb_num_sms() ? p.tiles : egb_num_sms();
  conv3x3_c64_c32_kernel<<<grid, CV_THREADS, smem, st>>>(mx, mw, p);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
