// Shared GEMM epilogue: bias -> (pre-activation copy) -> activation -> dropout -> activation-backward
// mask -> residual -> store / atomic accumulate.  Used by both the tcgen05 bf16 kernel and the
// FP32 FMA kernel so the two precision modes have identical semantics.
#pragma once
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

struct EpiMat {
  char* ptr;
  int f32;        // 1: fp32 storage, 0: bf16
  int rpg;        // rows per group (>= 1)
  long long rs;   // row stride (elements)
  long long gs;   // group stride (elements)
  int vec_ok;     // 16-byte vector access is legal for runs starting at multiples of 8 columns
};

struct EpiParams {
  int M, N;
  EpiMat c, c_pre, res, aux;
  const float* bias;
  float alpha;
  int act, act_bwd;
  float aux_scale;
  float drop_scale;
  unsigned drop_thresh;  // 0 = no dropout
  unsigned long long seed;
  int accumulate;
};

static inline EpiMat make_epimat(const egb_matrix& m, int total_rows) {
  EpiMat e;
  e.ptr = (char*)m.ptr;
  e.f32 = (m.dtype == EGB_F32);
  e.rpg = m.rows_per_group > 0 ? m.rows_per_group : (total_rows > 0 ? total_rows : 1);
  e.rs = m.row_stride;
  e.gs = m.group_stride;
  const int esz = e.f32 ? 4 : 2;
  const int per16 = 16 / esz;
  e.vec_ok = m.ptr != nullptr && ((uintptr_t)m.ptr % 16 == 0) && (m.row_stride % per16 == 0) &&
             (m.group_stride % per16 == 0);
  return e;
}

__device__ __forceinline__ long long epi_row_offset(const EpiMat& m, int row) {
  const int g = row / m.rpg;
  const int r = row - g * m.rpg;
  return (long long)g * m.gs + (long long)r * m.rs;
}

// read W consecutive elements of row `off` starting at column n into v (as float)
template <int W>
__device__ __forceinline__ void epi_load(const EpiMat& m, long long off, int n, int ncols, float (&v)[W]) {
  if (m.f32) {
    const float* p = reinterpret_cast<const float*>(m.ptr) + off + n;
    if (m.vec_ok && ncols == W) {
#pragma unroll
      for (int i = 0; i < W; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(p + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = i < ncols ? p[i] : 0.f;
    }
  } else {
    const bf16* p = reinterpret_cast<const bf16*>(m.ptr) + off + n;
    if (m.vec_ok && ncols == W && (W % 8 == 0)) {
#pragma unroll
      for (int i = 0; i < W; i += 8) {
        float t[8];
        ld8(p + i, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i + j] = t[j];
      }
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = i < ncols ? __bfloat162float(p[i]) : 0.f;
    }
  }
}

template <int W>
__device__ __forceinline__ void epi_store(const EpiMat& m, long long off, int n, int ncols, const float (&v)[W]) {
  if (m.f32) {
    float* p = reinterpret_cast<float*>(m.ptr) + off + n;
    if (m.vec_ok && ncols == W) {
#pragma unroll
      for (int i = 0; i < W; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (i < ncols) p[i] = v[i];
    }
  } else {
    bf16* p = reinterpret_cast<bf16*>(m.ptr) + off + n;
    if (m.vec_ok && ncols == W && (W % 8 == 0)) {
#pragma unroll
      for (int i = 0; i < W; i += 8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = v[i + j];
        st8(p + i, t);
      }
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (i < ncols) p[i] = __float2bfloat16_rn(v[i]);
    }
  }
}

// v holds W accumulator values of logical row m, columns [n, n+W); columns >= N are ignored.
template <int W>
__device__ __forceinline__ void epi_apply_store(const EpiParams& p, int m, int n, float (&v)[W]) {
  if (m >= p.M || n >= p.N) return;
  const int ncols = min(W, p.N - n);
#pragma unroll
  for (int i = 0; i < W; ++i) v[i] *= p.alpha;
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < W; ++i)
      if (i < ncols) v[i] += __ldg(p.bias + n + i);
  }
  if (p.c_pre.ptr != nullptr) epi_store<W>(p.c_pre, epi_row_offset(p.c_pre, m), n, ncols, v);
  if (p.act == EGB_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (p.act == EGB_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = gelu_erf(v[i]);
  }
  if (p.drop_thresh != 0u) {
    const unsigned long long base = (unsigned long long)m * (unsigned long long)p.N + (unsigned long long)n;
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = drop_keep(p.seed, base + i, p.drop_thresh) ? v[i] * p.drop_scale : 0.f;
  }
  if (p.act_bwd != EGB_ACTBWD_NONE) {
    float a[W];
    epi_load<W>(p.aux, epi_row_offset(p.aux, m), n, ncols, a);
    if (p.act_bwd == EGB_ACTBWD_RELU_MASK) {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = (a[i] != 0.f) ? v[i] * p.aux_scale : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] *= gelu_erf_grad(a[i]);
    }
  }
  if (p.res.ptr != nullptr) {
    float r[W];
    epi_load<W>(p.res, epi_row_offset(p.res, m), n, ncols, r);
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] += r[i];
  }
  const long long off = epi_row_offset(p.c, m);
  if (p.accumulate) {
    float* c = reinterpret_cast<float*>(p.c.ptr) + off + n;
#pragma unroll
    for (int i = 0; i < W; ++i)
      if (i < ncols) atomicAdd(c + i, v[i]);
  } else {
    epi_store<W>(p.c, off, n, ncols, v);
  }
}
