// Shared device/host helpers for the sm_100a kernels of the gaze+EEG fusion hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI functions return 0 on success, non-zero otherwise; message via egb_last_error)
// ---------------------------------------------------------------------------------------------
void egb_set_error(const char* fmt, ...);

#define EGB_CHECK(cond, ...)                                    \
  do {                                                          \
    if (!(cond)) {                                              \
      egb_set_error(__VA_ARGS__);                               \
      return 1;                                                 \
    }                                                           \
  } while (0)

#define EGB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      egb_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), \
                    cudaGetErrorString(_e));                                             \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

#define EGB_LAUNCH_CHECK() EGB_CUDA(cudaGetLastError())

int egb_num_sms();

// ---------------------------------------------------------------------------------------------
// scalar conversion helpers: kernels are templated on the HBM storage type T in {float, bf16}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact (erf) GELU and its derivative: the reference uses nn.GELU() default.  erf through the Abramowitz-Stegun
// 7.1.26 rational form (|error| <= 1.5e-7, i.e. below fp32 rounding of erff itself for this use): one reciprocal,
// one exponential and five FMAs, no branches -- and the derivative reuses the same exponential, since with
// u = x / sqrt(2) the Gaussian factor exp(-u^2) IS the normal pdf's exp(-x^2 / 2).
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& gauss) {
  const float u = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, u, 1.0f)));
  gauss = __expf(-u * u);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, gauss, 1.0f);
  cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, g;
  gelu_parts(x, cdf, g);
  return x * cdf;
}
// both at once (the forward epilogue that also saves the derivative for the backward pass)
__device__ __forceinline__ void gelu_erf_both(float x, float& y, float& dy) {
  float cdf, g;
  gelu_parts(x, cdf, g);
  y = x * cdf;
  dy = fmaf(x * 0.3989422804014327f, g, cdf);
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float cdf, g;
  gelu_parts(x, cdf, g);
  return fmaf(x * 0.3989422804014327f, g, cdf);
}

// ---------------------------------------------------------------------------------------------------------------
// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two lanes per issued instruction).  The fma pipe retires one
// warp instruction per two cycles per scheduler, so an element-wise epilogue that has to keep pace with tcgen05
// (GELU and its derivative behind a K = 768 GEMM: ~17 fma-pipe operations per element against a budget of ~12) is
// issue-bound in scalar form; in packed form it fits.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 splat2(float c) { return make_float2(c, c); }

// gelu_erf_both() for two elements at once: same Abramowitz-Stegun 7.1.26 form, 14 packed fma-pipe operations,
// four MUFU (2 x rcp, 2 x ex2) and four sign/abs bit operations per pair.
__device__ __forceinline__ void gelu_erf_both2(float2 x, float2& y, float2& dy) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = fma2(ax, splat2(0.3275911f * 0.70710678118654752f), splat2(1.0f));
  float2 t, g;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  const float2 arg = mul2(mul2(x, x), splat2(-0.5f * 1.4426950408889634f));   // -u^2 log2(e), u = x / sqrt(2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g.x) : "f"(arg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g.y) : "f"(arg.y));
  float2 poly = fma2(t, splat2(1.061405429f), splat2(-1.453152027f));
  poly = fma2(t, poly, splat2(1.421413741f));
  poly = fma2(t, poly, splat2(-0.284496736f));
  poly = fma2(t, poly, splat2(0.254829592f));
  const float2 pg = mul2(mul2(poly, t), g);                                   // 1 - erf(|u|)
  const float2 e = fma2(pg, splat2(-1.0f), splat2(1.0f));                     // erf(|u|)
  const float2 sh = make_float2(copysignf(0.5f, x.x), copysignf(0.5f, x.y));
  const float2 cdf = fma2(sh, e, splat2(0.5f));
  y = mul2(x, cdf);
  dy = fma2(mul2(x, g), splat2(0.3989422804014327f), cdf);
}

// gelu_erf_both2 over NP independent pairs, written PHASE-MAJOR: every line is NP independent instructions, so the
// instruction stream itself interleaves the NP dependency chains.  With the pairs evaluated one after the other the
// two epilogue warps of a scheduler spent 28 % of their stall samples in fixed-latency dependency waits (ncu, fc1 + GELU
// GEMM: issue slots 40 % busy, tensor pipe 52 %).
template <int NP>
__device__ __forceinline__ void gelu_erf_both2n(const float2 (&x)[NP], float2 (&y)[NP], float2 (&dy)[NP]) {
  float2 t[NP], g[NP], poly[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const float2 ax = make_float2(fabsf(x[i].x), fabsf(x[i].y));
    t[i] = fma2(ax, splat2(0.3275911f * 0.70710678118654752f), splat2(1.0f));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t[i].x) : "f"(t[i].x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t[i].y) : "f"(t[i].y));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) g[i] = mul2(mul2(x[i], x[i]), splat2(-0.5f * 1.4426950408889634f));
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g[i].x) : "f"(g[i].x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g[i].y) : "f"(g[i].y));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) poly[i] = fma2(t[i], splat2(1.061405429f), splat2(-1.453152027f));
#pragma unroll
  for (int i = 0; i < NP; ++i) poly[i] = fma2(t[i], poly[i], splat2(1.421413741f));
#pragma unroll
  for (int i = 0; i < NP; ++i) poly[i] = fma2(t[i], poly[i], splat2(-0.284496736f));
#pragma unroll
  for (int i = 0; i < NP; ++i) poly[i] = fma2(t[i], poly[i], splat2(0.254829592f));
#pragma unroll
  for (int i = 0; i < NP; ++i) poly[i] = mul2(mul2(poly[i], t[i]), g[i]);               // 1 - erf(|u|)
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const float2 e = fma2(poly[i], splat2(-1.0f), splat2(1.0f));                        // erf(|u|)
    const float2 sh = make_float2(copysignf(0.5f, x[i].x), copysignf(0.5f, x[i].y));
    const float2 cdf = fma2(sh, e, splat2(0.5f));
    y[i] = mul2(x[i], cdf);
    dy[i] = fma2(mul2(x[i], g[i]), splat2(0.3989422804014327f), cdf);
  }
}

// Device-resident seed epoch (CUDA-graph replays): dropout seeds are launch ARGUMENTS, i.e. frozen into a captured graph.
// When the host has created the epoch word (egb_seed_epoch_enable), every kernel that draws a mask mixes the CURRENT
// value of that device word into its seed, and the graph advances the word once per replay (egb_seed_epoch_advance) --
// forward and backward of one replay see the same value, different replays draw different masks.  NULL = seeds as passed.
const unsigned long long* egb_seed_epoch_ptr();
__device__ __forceinline__ unsigned long long egb_mix_seed(unsigned long long seed, const unsigned long long* epoch) {
  if (epoch == nullptr) return seed;
  unsigned long long e = __ldg(epoch) * 0x9E3779B97F4A7C15ull;
  e ^= e >> 29;
  return seed ^ e;
}

// Counter-based dropout mask, stateless so that the backward pass regenerates it from (seed, element index) instead of
// storing it.  Elements are grouped into aligned RUNS of 8 consecutive indices: one multiply-xorshift hash of
// (seed, idx >> 3) seeds the run and the run's elements step a 32-bit LCG from there; an element is kept iff its state
// >= p * 2^32 (the comparison is decided by the state's high bits, the strong ones of an LCG).  Every kernel that touches
// 8 aligned elements at a time pays one hash and eight multiply-adds instead of eight hashes: the GEMM epilogues of the
// EEG encoder (dropout after every Linear) were bound by exactly that integer work.
__device__ __forceinline__ uint32_t drop_hash(uint64_t seed, uint64_t idx) {
  uint32_t h = ((uint32_t)idx ^ (uint32_t)seed) * 0x9E3779B1u + ((uint32_t)(idx >> 32) ^ (uint32_t)(seed >> 32)) * 0x85EBCA77u;
  h ^= h >> 16; h *= 0x21F0AAADu; h ^= h >> 15; h *= 0x735A2D97u; h ^= h >> 15;   // 2-round multiply-xorshift finaliser
  return h;
}
__device__ __forceinline__ uint32_t drop_step(uint32_t s) { return s * 0x2C9277B5u + 0xAC564B05u; }
// state of element `idx` (any index): hash of its run, stepped (idx & 7) times
__device__ __forceinline__ uint32_t drop_state(uint64_t seed, uint64_t idx) {
  uint32_t s = drop_hash(seed, idx >> 3);
  for (int k = (int)(idx & 7); k > 0; --k) s = drop_step(s);
  return s;
}
__device__ __forceinline__ bool drop_keep(uint64_t seed, uint64_t idx, uint32_t thresh) {
  return drop_state(seed, idx) >= thresh;
}
// i-fold composition of drop_step: s -> s * mul + add  (mod 2^32), as compile-time constants
struct DropLeap { uint32_t mul, add; };
__host__ __device__ constexpr DropLeap drop_leap(int i) {
  uint32_t m = 1u, a = 0u;
  for (int k = 0; k < i; ++k) { a = a * 0x2C9277B5u + 0xAC564B05u; m = m * 0x2C9277B5u; }
  return DropLeap{m, a};
}
// v[0..8) <- dropout of the aligned run starting at element index `base` (base % 8 == 0).  The eight states are taken
// by LEAPFROG from the run's hash (state_i = hash * A^i + C_i, the same values as i sequential steps): eight independent
// multiply-adds instead of a chain of eight -- the epilogue warps are bound by dependent-instruction latency.
__device__ __forceinline__ void drop_apply_run8(uint64_t seed, uint64_t base, uint32_t thresh, float scale, float (&v)[8]) {
  const uint32_t s0 = drop_hash(seed, base >> 3);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const DropLeap lp = drop_leap(i);
    const uint32_t s = s0 * lp.mul + lp.add;
    v[i] = s >= thresh ? v[i] * scale : 0.f;
  }
}
// Attention-probability dropout: the key columns of a query row are grouped into aligned runs of 32; ONE hash of
// (seed, row, run) seeds the run and column k of the run takes the k-fold LCG step of it (leapfrog constants, i.e. 32
// independent multiply-adds); keep iff state >= p * 2^32.  The softmax loops of the tensor-core attention kernels are
// bound by integer throughput once dropout is on: the first scheme (a full multiply-xorshift hash per key) cost the EEG
// layers 25 % of their time, one hash per PAIR of keys still ~10 integer instructions per pair; this one is one
// multiply-add and one compare per key.  `row_lin` = row index * Lk (element index of the row's first key).
__device__ __forceinline__ uint32_t drop_att_run_state(uint64_t seed, uint64_t row_lin, int col0) {
  return drop_hash(seed, row_lin + (uint64_t)col0);          // col0 = first column of the 32-column run
}
__device__ __forceinline__ bool drop_keep_att(uint64_t seed, uint64_t row_lin, int col, uint32_t thresh) {
  uint32_t s = drop_att_run_state(seed, row_lin, col & ~31);
  for (int k = col & 31; k > 0; --k) s = drop_step(s);
  return s >= thresh;
}
static inline uint32_t drop_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t < 0) t = 0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}

// 128-bit streaming load/store helpers
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// load/store 4 consecutive elements of T as floats (16 B for float, 8 B for bf16)
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
// 8 consecutive elements (32 B float / 16 B bf16)
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}
