// 3x3 convolution over zero-bordered channels-last images as a TRUE implicit GEMM on tcgen05: the spectrogram CNN's second
// convolution (32 -> 64 channels, + bias; dual_eeg_transformer.py:81-86) and its data gradient (64 -> 32 channels).
//
// The generic GEMM kernel runs this convolution over an overlapping-row VIEW of the image buffer (K = 3 row segments x
// 4 positions x 64 channels): every input element crosses L2 -> shared memory 12 times, 8.5 GB per launch, and the
// launch is bound by exactly that (1.5 ms at 0.13 of the tensor peak).  Here a CTA stages the input rows of a tile ONCE
// ([128 + 2 Wp + 2 positions][64 channels], one TMA box, 128-byte swizzle) and issues the nine taps as MMAs whose A
// descriptors simply START (a Wp + b) rows further down: the hardware applies the 128-byte swizzle to absolute
// shared-memory address bits, so a K-major operand may begin at any row of a swizzled tile (verified on the device by
// csrc/tests/shift_desc_test.cu: exact for every row offset with the descriptor's base-offset field left at 0, for
// 128-byte rows / SWIZZLE_128B and for 64-byte rows / SWIZZLE_64B alike).
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
int egb_prof_enabled();
void egb_prof_begin(cudaStream_t st, double flops, double bytes, int kind);
void egb_prof_end(cudaStream_t st);
void egb_prof_tag(double a, double b, double c, double d);
int egb_tmap_2d(CUtensorMap* out, const void* ptr, long long inner, long long rows, long long rs, int box_inner, int box_rows);

namespace {

constexpr int CV_THREADS = 192;
constexpr int CV_M = 128;
constexpr int CV_XS = 3;        // input-tile ring (a tile is ~20 KB; a TMA round trip outlasts a tile's 36 MMAs)
constexpr int CV_CTAS_PER_SM = 2;   // two CTAs per SM: a tile's MMAs are issued by ONE thread (36-18 small instructions), two issuers keep the tensor pipe fed

struct ConvParams {
  bf16* y;            // [rows, COUT]
  const float* bias;  // [COUT] or NULL
  long long M;        // output rows m = 0 .. M-1, written at flat position m + out_shift
  long long out_shift;
  int Wp;             // padded image width (row shift of one kernel row)
  int box_rows;       // input rows per tile = 128 + 2 Wp + 2
  int tiles;
};

// K-major shared-memory operand descriptor for rows of RB bytes (128: SWIZZLE_128B, 64: SWIZZLE_64B); 8-row groups are
// 8 * RB bytes apart
template <int RB>
__device__ __forceinline__ uint64_t conv_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(((8u * RB) >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(RB == 128 ? 2 : 4) << 61;
  return d;
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(CV_THREADS, CV_CTAS_PER_SM)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  constexpr int RB = CIN * 2;                 // bytes per input row (position)
  constexpr int WTAP = COUT * RB;             // bytes of one tap's weight tile [COUT rows][CIN]
  constexpr int WSEG = 4 * CIN;               // columns of one kernel-row segment of the weight matrix (four position slots)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tile_bytes = ((p.box_rows * RB) + 1023) & ~1023;
  uint8_t* sX = smem;                                  // [CV_XS][box_rows][128 B]
  uint8_t* sW = smem + CV_XS * tile_bytes;             // [9][COUT rows][RB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 9 * WTAP);
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
  if (warp == 1) ptx::tmem_alloc<2 * COUT>(slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(wbar, 9 * WTAP);
      for (int t = 0; t < 9; ++t) ptx::tma_load_2d(sW + t * WTAP, &tmW, wbar, (t / 3) * WSEG + (t % 3) * CIN, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        ptx::mbar_wait(&empty[s], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(p.box_rows * RB));
        ptx::tma_load_2d(sX + s * tile_bytes, &tmX, &full[s], 0, tile * CV_M);      // rows past the buffer end: zero-filled
        if (++s == CV_XS) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc_bf16(CV_M, COUT, 0, 0);
    ptx::mbar_wait(wbar, 0);
    int s = 0, ts = 0;
    uint32_t ph = 0, tph = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tempty[ts], tph ^ 1u);
      ptx::mbar_wait(&full[s], ph);
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t xa = ptx::smem_u32(sX + s * tile_bytes);
        const uint32_t td = tmem + (uint32_t)(COUT * ts);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          // tap (a, b): the A operand starts (a Wp + b) rows into the tile (any row: the swizzle is on absolute address bits)
          const uint64_t ad = conv_desc<RB>(xa + (uint32_t)(((t / 3) * p.Wp + (t % 3)) * RB));
          const uint64_t bd = conv_desc<RB>(ptx::smem_u32(sW + t * WTAP));
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k) ptx::umma_bf16(td, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (t > 0 || k > 0) ? 1u : 0u);
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
      uint32_t v[COUT];
#pragma unroll
      for (int h = 0; h < COUT / 32; ++h)
        ptx::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(COUT * s + 32 * h), reinterpret_cast<uint32_t (&)[32]>(v[32 * h]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_relaxed(&tempty[s]);
      const long long m = (long long)tile * CV_M + quad * 32 + lane;
      if (m < p.M) {
        uint4* dst = reinterpret_cast<uint4*>(p.y + (m + p.out_shift) * COUT);   // this lane's row: COUT * 2 contiguous bytes
#pragma unroll
        for (int g = 0; g < COUT / 8; ++g) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[8 * g + i]);
          if (p.bias != nullptr) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + 8 * g));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + 8 * g + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          uint4 o;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
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
    ptx::tmem_dealloc<2 * COUT>(tmem);
  }
}

template <int CIN, int COUT>
int launch_conv(const void* x, long long x_rows, const void* w, const float* bias, void* y, long long M, long long out_shift,
                int Wp, cudaStream_t st) {
  EGB_CHECK(M > 0 && Wp >= 3 && 128 + 2 * Wp + 2 <= 256, "conv3x3: unsupported geometry (M=%lld, Wp=%d)", M, Wp);
  EGB_CHECK(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)y % 16) == 0 && (out_shift * COUT * 2) % 16 == 0 &&
                (bias == nullptr || ((uintptr_t)bias % 16) == 0),
            "conv3x3: misaligned buffers");
  ConvParams p;
  p.y = (bf16*)y;
  p.bias = bias;
  p.M = M;
  p.out_shift = out_shift;
  p.Wp = Wp;
  p.box_rows = 128 + 2 * Wp + 2;
  p.tiles = (int)((M + CV_M - 1) / CV_M);
  CUtensorMap mx, mw;
  if (egb_tmap_2d(&mx, x, CIN, x_rows, CIN, CIN, p.box_rows)) return 1;
  if (egb_tmap_2d(&mw, w, 12 * CIN, COUT, 12 * CIN, CIN, COUT)) return 1;
  const int tile_bytes = ((p.box_rows * CIN * 2) + 1023) & ~1023;
  const size_t smem = (size_t)CV_XS * tile_bytes + 9 * COUT * CIN * 2 + 256 + 1024;
  EGB_CHECK(smem <= 200 * 1024, "conv3x3: tile too large for shared memory (%zu bytes)", smem);
  static bool attr = false;
  if (!attr) {
    EGB_CUDA(cudaFuncSetAttribute(conv3x3_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  // two CTAs per SM while their tiles fit (the spectrogram geometries); one for very wide padded rows
  const int cap = (smem <= 112 * 1024 ? CV_CTAS_PER_SM : 1) * egb_num_sms();
  const int grid = p.tiles < cap ? p.tiles : cap;
  // bench.py's per-launch GEMM timing counts these launches in the tensor-core family (variant 4 = direct convolution)
  const bool prof = egb_prof_enabled() != 0;
  if (prof) {
    egb_prof_begin(st, 2.0 * (double)M * COUT * 9 * CIN, 2.0 * (double)M * (CIN + COUT), 0);
    egb_prof_tag((double)M, COUT, 9 * CIN, 4e10 + COUT * 1e7 + (bias != nullptr ? 1 : 0));
  }
  conv3x3_kernel<CIN, COUT><<<grid, CV_THREADS, smem, st>>>(mx, mw, p);
  if (prof) egb_prof_end(st);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

// ---- weight gradient --------------------------------------------------------------------------------------------------
//   dW^T[(a, j, c)][o] = sum over flat positions m of  x[m + a Wp + j][c] * dy[m + dy_shift][o]      a < 3, j < 4 (j = 3 unused)
// Both operands are MN-major (K = positions = tile rows).  The 32-channel activation tile [128 + 2 Wp + 3 rows][64 B]
// (SWIZZLE_64B) is read as an M = 128 operand whose four 32-element M atoms are ONE ROW (64 B) apart -- the descriptor's
// leading-dimension offset is just an address increment, so atom j is the same tile starting one position later: the four
// horizontal taps of a kernel row come out of ONE MMA (csrc/tests/shift_desc_test.cu, test_dw).  Per 128-position tile:
// 3 kernel rows x 8 K-steps = 24 MMAs (M = 128, N = 64, K = 16) into three accumulators that stay in TMEM for the whole
// CTA; every element of x and dy crosses L2 -> shared memory once.  At the end the CTA adds its partial sums to dW^T
// (fp32, red.global.add.v4).
struct ConvDwParams {
  float* dwt;          // [384, 64] fp32, zeroed by the caller
  long long dy_shift;
  int Wp;
  int box_rows;        // activation rows per tile = 128 + 2 Wp + 3
  int tiles;
};

constexpr int DW_STAGES = 3;

__device__ __forceinline__ void red_add4(float* addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
               "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}

__global__ void __launch_bounds__(CV_THREADS, CV_CTAS_PER_SM)
conv3x3_dw_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, const ConvDwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int xtile = ((p.box_rows * 64) + 1023) & ~1023;
  const int stage = xtile + CV_M * 128;               // activation rows, then the 128 gradient rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DW_STAGES * stage);
  uint64_t* full = bars;
  uint64_t* empty = bars + DW_STAGES;
  uint64_t* done = bars + 2 * DW_STAGES;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2 * DW_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmD);
    for (int i = 0; i < DW_STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<256>(slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        ptx::mbar_wait(&empty[s], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(p.box_rows * 64 + CV_M * 128));
        ptx::tma_load_2d(smem + s * stage, &tmX, &full[s], 0, tile * CV_M);
        ptx::tma_load_2d(smem + s * stage + xtile, &tmD, &full[s], 0, (int)p.dy_shift + tile * CV_M);   // rows past the end: zeros
        if (++s == DW_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc_bf16(CV_M, 64, 1, 1);
    int s = 0;
    uint32_t ph = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      ptx::mbar_wait(&full[s], ph);
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t xa = ptx::smem_u32(smem + s * stage);
        const uint64_t bd = ptx::make_smem_desc(xa + (uint32_t)xtile, 128u * 128u, 1024u);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          // MN-major, SWIZZLE_64B, M atoms 64 B (one position) apart, 8-position groups 512 B apart
          const uint32_t addr = xa + (uint32_t)(a * p.Wp * 64);
          uint64_t ad = (uint64_t)((addr & 0x3FFFFu) >> 4);
          ad |= (uint64_t)(64u >> 4) << 16;
          ad |= (uint64_t)(512u >> 4) << 32;
          ad |= (uint64_t)1 << 46;
          ad |= (uint64_t)4 << 61;
#pragma unroll
          for (int k = 0; k < CV_M / 16; ++k)
            ptx::umma_bf16(tmem + (uint32_t)(64 * a), ad + (uint64_t)(64 * k), bd + (uint64_t)(128 * k), idesc, (!first || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty[s]);
      }
      __syncwarp();
      first = false;
      if (++s == DW_STAGES) { s = 0; ph ^= 1u; }
    }
    if (lane == 0) ptx::umma_commit(done);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    ptx::mbar_wait(done, 0);
    ptx::tc_fence_after();
    if (quad < 3) {                                   // M rows 96..127 are the unused fourth position slot
      const int row = quad * 32 + lane;
#pragma unroll 1
      for (int a = 0; a < 3; ++a) {
        float* dst = p.dwt + ((size_t)(a * 128 + row)) * 64;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(64 * a + 32 * h), v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g) red_add4(dst + 32 * h + 4 * g, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<256>(tmem);
  }
}

int launch_conv_dw(const void* x, long long x_rows, const void* dy, long long dy_rows, float* dwt, long long M, long long dy_shift,
                   int Wp, cudaStream_t st) {
  EGB_CHECK(M > 0 && Wp >= 3 && 128 + 2 * Wp + 3 <= 256 && dy_shift >= 0 && dy_shift + M < (1ll << 31),
            "conv3x3_dw: unsupported geometry (M=%lld, Wp=%d)", M, Wp);
  EGB_CHECK(((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)dwt % 16) == 0, "conv3x3_dw: misaligned buffers");
  ConvDwParams p;
  p.dwt = dwt;
  p.dy_shift = dy_shift;
  p.Wp = Wp;
  p.box_rows = 128 + 2 * Wp + 3;
  p.tiles = (int)((M + CV_M - 1) / CV_M);
  CUtensorMap mx, md;
  if (egb_tmap_2d(&mx, x, 32, x_rows, 32, 32, p.box_rows)) return 1;
  const long long drows = dy_rows < dy_shift + M ? dy_rows : dy_shift + M;     // positions >= M contribute nothing
  if (egb_tmap_2d(&md, dy, 64, drows, 64, 64, CV_M)) return 1;
  const int xtile = ((p.box_rows * 64) + 1023) & ~1023;
  const size_t smem = (size_t)DW_STAGES * (xtile + CV_M * 128) + 256 + 1024;
  EGB_CHECK(smem <= 112 * 1024, "conv3x3_dw: tile too large for two CTAs per SM (%zu bytes)", smem);
  static bool attr = false;
  if (!attr) {
    EGB_CUDA(cudaFuncSetAttribute(conv3x3_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr = true;
  }
  const int cap = CV_CTAS_PER_SM * egb_num_sms();
  const int grid = p.tiles < cap ? p.tiles : cap;
  const bool prof = egb_prof_enabled() != 0;
  if (prof) {
    egb_prof_begin(st, 2.0 * (double)M * 64 * 288, 2.0 * (double)M * 96, 0);
    egb_prof_tag(288, 64, (double)M, 4e10 + 64 * 1e7 + 128);
  }
  conv3x3_dw_kernel<<<grid, CV_THREADS, smem, st>>>(mx, md, p);
  if (prof) egb_prof_end(st);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

/* y[(m + out_shift) * 32 + c] = sum over the nine taps (a, b) and 64 input channels o of
       x[(m + a * Wp + b) * 64 + o] * w[c * 768 + a * 256 + b * 64 + o]          for m in [0, M)
   x: bf16 [x_rows, 64] (zero-bordered channels-last images, flat), w: bf16 [32, 768] (three 256-wide row segments of four
   64-wide slots, the fourth unused), y: bf16 rows of 32.  Reads x rows up to M + 2 Wp + 2 (rows >= x_rows read as 0). */
int egb_conv3x3_c64_c32(const void* x, long long x_rows, const void* w, void* y, long long M, long long out_shift, int Wp,
                        void* stream) {
  return launch_conv<64, 32>(x, x_rows, w, nullptr, y, M, out_shift, Wp, (cudaStream_t)stream);
}

/* the forward convolution: y[(m + out_shift) * 64 + o] = bias[o] + sum over taps and 32 input channels c of
       x[(m + a * Wp + b) * 32 + c] * w[o * 384 + a * 128 + b * 32 + c];  x: bf16 [x_rows, 32], w: bf16 [64, 384], bias fp32 */
int egb_conv3x3_c32_c64(const void* x, long long x_rows, const void* w, const float* bias, void* y, long long M,
                        long long out_shift, int Wp, void* stream) {
  return launch_conv<32, 64>(x, x_rows, w, bias, y, M, out_shift, Wp, (cudaStream_t)stream);
}

/* weight gradient of the forward convolution, transposed: dwt[(a * 128 + j * 32 + c) * 64 + o] += sum over m in [0, M) of
       x[(m + a * Wp + j) * 32 + c] * dy[(m + dy_shift) * 64 + o]      a < 3, j < 3 (rows j = 3 of every 128 are not written)
   x: bf16 [x_rows, 32], dy: bf16 [dy_rows, 64] (zero at border positions), dwt: fp32 [384, 64], zeroed by the caller. */
int egb_conv3x3_dw_c32_c64(const void* x, long long x_rows, const void* dy, long long dy_rows, float* dwt, long long M,
                           long long dy_shift, int Wp, void* stream) {
  return launch_conv_dw(x, x_rows, dy, dy_rows, dwt, M, dy_shift, Wp, (cudaStream_t)stream);
}

}  // extern "C"
