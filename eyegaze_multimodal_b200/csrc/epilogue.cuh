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
  const unsigned long long* epoch;   // device seed epoch (egb_mix_seed), NULL when not enabled
  int accumulate;
  float* colsum;         // optional [N]: column sums of the stored output are accumulated here
  int exp;               // EGB_EPI_EXP experiment switch (0 = normal): 1 no residual / act' loads, 2 loads hit row 0 only, 3 no C store
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
  if (row < m.rpg) return (long long)row * m.rs;   // single group (the common case): no division
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
  if (p.act == EGB_ACT_GELU_DGRAD) {
    float dv[W];
#pragma unroll
    for (int i = 0; i < W; ++i) gelu_erf_both(v[i], v[i], dv[i]);
    if (p.c_pre.ptr != nullptr) epi_store<W>(p.c_pre, epi_row_offset(p.c_pre, m), n, ncols, dv);
  } else if (p.c_pre.ptr != nullptr) {
    epi_store<W>(p.c_pre, epi_row_offset(p.c_pre, m), n, ncols, v);
  }
  if (p.act == EGB_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (p.act == EGB_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = gelu_erf(v[i]);
  }
  if (p.drop_thresh != 0u) {
    const unsigned long long base = (unsigned long long)m * (unsigned long long)p.N + (unsigned long long)n;
    const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = drop_keep(seed_eff, base + i, p.drop_thresh) ? v[i] * p.drop_scale : 0.f;
  }
  if (p.act_bwd != EGB_ACTBWD_NONE) {
    float a[W];
    epi_load<W>(p.aux, epi_row_offset(p.aux, m), n, ncols, a);
    if (p.act_bwd == EGB_ACTBWD_RELU_MASK) {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = (a[i] != 0.f) ? v[i] * p.aux_scale : 0.f;
    } else if (p.act_bwd == EGB_ACTBWD_MUL) {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] *= a[i];
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

// ---------------------------------------------------------------------------------------------
// Row-hoisted variant for the tensor-core kernels: the per-row work (bounds, group/row offsets of every
// operand matrix -- integer divisions) is done once per tile row, not once per 32-column chunk.
// ---------------------------------------------------------------------------------------------
struct EpiRow {
  long long c, c_pre, res, aux;
  unsigned long long dbase;   // m * N: element index of the row's first column (dropout counter)
  bool ok;
};

__device__ __forceinline__ EpiRow epi_row_setup(const EpiParams& p, int m) {
  EpiRow r;
  r.ok = m < p.M;
  r.c = r.c_pre = r.res = r.aux = 0;
  r.dbase = (unsigned long long)m * (unsigned long long)p.N;
  if (r.ok) {
    r.c = epi_row_offset(p.c, m);
    if (p.c_pre.ptr != nullptr) r.c_pre = epi_row_offset(p.c_pre, m);
    if (p.res.ptr != nullptr) r.res = epi_row_offset(p.res, m);
    if (p.act_bwd != EGB_ACTBWD_NONE) r.aux = epi_row_offset(p.aux, m);
  }
  return r;
}

template <int W>
__device__ __forceinline__ void epi_apply_store_row(const EpiParams& p, const EpiRow& row, int m, int n, float (&v)[W]) {
  if (!row.ok || n >= p.N) return;
  const int ncols = min(W, p.N - n);
  if (p.alpha != 1.f) {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] *= p.alpha;
  }
  if (p.bias != nullptr) {
    if (ncols == W && ((reinterpret_cast<uintptr_t>(p.bias) & 15u) == 0) && (n & 3) == 0) {
#pragma unroll
      for (int i = 0; i < W; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (i < ncols) v[i] += __ldg(p.bias + n + i);
    }
  }
  if (p.act == EGB_ACT_GELU_DGRAD) {
    float dv[W];
#pragma unroll
    for (int i = 0; i < W; ++i) gelu_erf_both(v[i], v[i], dv[i]);
    if (p.c_pre.ptr != nullptr) epi_store<W>(p.c_pre, row.c_pre, n, ncols, dv);
  } else if (p.c_pre.ptr != nullptr) {
    epi_store<W>(p.c_pre, row.c_pre, n, ncols, v);
  }
  if (p.act == EGB_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (p.act == EGB_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = gelu_erf(v[i]);
  }
  if (p.drop_thresh != 0u) {
    const unsigned long long base = (unsigned long long)m * (unsigned long long)p.N + (unsigned long long)n;
    const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = drop_keep(seed_eff, base + i, p.drop_thresh) ? v[i] * p.drop_scale : 0.f;
  }
  if (p.act_bwd != EGB_ACTBWD_NONE) {
    float a[W];
    epi_load<W>(p.aux, row.aux, n, ncols, a);
    if (p.act_bwd == EGB_ACTBWD_RELU_MASK) {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = (a[i] != 0.f) ? v[i] * p.aux_scale : 0.f;
    } else if (p.act_bwd == EGB_ACTBWD_MUL) {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] *= a[i];
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] *= gelu_erf_grad(a[i]);
    }
  }
  if (p.res.ptr != nullptr) {
    float r[W];
    epi_load<W>(p.res, row.res, n, ncols, r);
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] += r[i];
  }
  if (p.accumulate) {
    float* c = reinterpret_cast<float*>(p.c.ptr) + row.c + n;
#pragma unroll
    for (int i = 0; i < W; ++i)
      if (i < ncols) atomicAdd(c + i, v[i]);
  } else {
    epi_store<W>(p.c, row.c, n, ncols, v);
  }
}

// ---------------------------------------------------------------------------------------------
// Compile-time specialised epilogue for the tensor-core kernels.  The generic function above decides everything
// at run time (per call: ~10 parameter loads and branches); in the transposed 8-column layout it is invoked four
// times per 32x32 chunk, and ncu shows the warps stalled on exactly those dependent branch chains.  The fast
// variants fix the operation list in a template mask; they require bf16 16-byte-aligned matrices (fp32 for the
// split-K accumulate), N % 8 == 0 and alpha == 1 (egb_epi_fast_mask() checks this and otherwise returns GENERIC).
// Dropout stays a (warp-uniform) run-time switch.
// ---------------------------------------------------------------------------------------------
enum : int {
  EF_BIAS = 1, EF_RELU = 2, EF_GELU = 4, EF_RES = 8, EF_PRE = 16, EF_ABWD_RELU = 32, EF_ABWD_GELU = 64, EF_ACC = 128,
  EF_DGELU = 256,     // with EF_GELU | EF_PRE: the saved tensor is gelu'(pre)
  EF_ABWD_MUL = 512,  // multiply by the saved derivative
  EF_COLSUM = 1024,   // accumulate the column sums of the stored tile (bias gradient) -- done by the chunk epilogue
  EF_GENERIC = 1 << 20
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// operands of one 8-column run that can be fetched ahead of the math (bias is shared by all rows of a chunk)
struct EpiPre8 {
  uint4 res, aux;
};

template <int F>
__device__ __forceinline__ void epi_prefetch8(const EpiParams& p, const EpiRow& row, int n, EpiPre8& pre) {
  if (!row.ok || n >= p.N) return;
  if (p.exp == 1) { pre.aux = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u); pre.res = pre.aux; return; }
  const long long ra = p.exp == 2 ? 0 : row.aux, rr = p.exp == 2 ? 0 : row.res;
  if (F & (EF_ABWD_RELU | EF_ABWD_GELU | EF_ABWD_MUL)) pre.aux = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.aux.ptr) + ra + n));
  if (F & EF_RES) pre.res = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.res.ptr) + rr + n));
}

__device__ __forceinline__ void unpack8(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}

// The arithmetic of one 8-column run (everything but the stores): v <- epilogue(v); dv <- gelu'(pre) for EF_DGELU.
// (m, n) locate the run for the dropout counter.
template <int F>
__device__ __forceinline__ void epi_math8(const EpiParams& p, unsigned long long seed_eff, unsigned long long row_base,
                                          int n, float (&v)[8], const float (&bias)[8], const EpiPre8& pre,
                                          float (&dv)[8]) {
  if (F & EF_BIAS) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float2 s2 = add2(make_float2(v[i], v[i + 1]), make_float2(bias[i], bias[i + 1]));
      v[i] = s2.x; v[i + 1] = s2.y;
    }
  }
  if (F & EF_DGELU) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {   // packed pairs: see gelu_erf_both2
      float2 y2, d2;
      gelu_erf_both2(make_float2(v[i], v[i + 1]), y2, d2);
      v[i] = y2.x; v[i + 1] = y2.y; dv[i] = d2.x; dv[i + 1] = d2.y;
    }
  } else {
    if (F & EF_PRE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dv[i] = v[i];
    }
    if (F & EF_RELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (F & EF_GELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    }
  }
  // n and N are multiples of 8 on this path: the run is aligned.  The effective seed (launch seed x device epoch) and
  // the row's first element index are computed once per kernel / per tile row by the caller: recomputed per 8-column run
  // (a 64-bit multiply each) they were a third of the instructions of the EEG encoder's dropout epilogues.
  if (p.drop_thresh != 0u) drop_apply_run8(seed_eff, row_base + (unsigned long long)n, p.drop_thresh, p.drop_scale, v);
  if (F & (EF_ABWD_RELU | EF_ABWD_GELU | EF_ABWD_MUL)) {
    float a[8];
    unpack8(pre.aux, a);
    if (F & EF_ABWD_RELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (a[i] != 0.f) ? v[i] * p.aux_scale : 0.f;
    } else if (F & EF_ABWD_MUL) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 m2 = mul2(make_float2(v[i], v[i + 1]), make_float2(a[i], a[i + 1]));
        v[i] = m2.x; v[i + 1] = m2.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= gelu_erf_grad(a[i]);
    }
  }
  if (F & EF_RES) {
    float r[8];
    unpack8(pre.res, r);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float2 s2 = add2(make_float2(v[i], v[i + 1]), make_float2(r[i], r[i + 1]));
      v[i] = s2.x; v[i + 1] = s2.y;
    }
  }
}

template <int F>
__device__ __forceinline__ void epi_fast8(const EpiParams& p, const EpiRow& row, int m, int n, float (&v)[8],
                                          const float (&bias)[8], const EpiPre8& pre, unsigned long long seed_eff) {
  if (!row.ok || n >= p.N) return;
  if (F & EF_ACC) {
    float* c = reinterpret_cast<float*>(p.c.ptr) + row.c + n;
    red_add_v4(c, v[0], v[1], v[2], v[3]);
    red_add_v4(c + 4, v[4], v[5], v[6], v[7]);
    return;
  }
  float dv[8];
  epi_math8<F>(p, seed_eff, row.dbase, n, v, bias, pre, dv);
  if (F & (EF_DGELU | EF_PRE)) st8(reinterpret_cast<bf16*>(p.c_pre.ptr) + row.c_pre + n, dv);
  if (p.exp != 3) st8(reinterpret_cast<bf16*>(p.c.ptr) + row.c + n, v);
}

// host: the specialisation that implements this call exactly, or EF_GENERIC
static inline int egb_epi_fast_mask(const EpiParams& e) {
  auto bf16_vec = [](const EpiMat& m) { return m.ptr != nullptr && !m.f32 && m.vec_ok; };
  if ((e.N % 8) != 0 || e.alpha != 1.f) return EF_GENERIC;
  if (e.bias != nullptr && ((uintptr_t)e.bias % 16) != 0) return EF_GENERIC;
  if (e.accumulate) return (e.c.f32 && e.c.vec_ok) ? EF_ACC : EF_GENERIC;
  if (!bf16_vec(e.c)) return EF_GENERIC;
  int f = 0;
  if (e.bias != nullptr) f |= EF_BIAS;
  if (e.act == EGB_ACT_RELU) f |= EF_RELU;
  if (e.act == EGB_ACT_GELU) f |= EF_GELU;
  if (e.act == EGB_ACT_GELU_DGRAD) { if (e.c_pre.ptr == nullptr) return EF_GENERIC; f |= EF_GELU | EF_DGELU; }
  if (e.c_pre.ptr != nullptr) { if (!bf16_vec(e.c_pre)) return EF_GENERIC; f |= EF_PRE; }
  if (e.res.ptr != nullptr) { if (!bf16_vec(e.res)) return EF_GENERIC; f |= EF_RES; }
  if (e.act_bwd != EGB_ACTBWD_NONE) {
    if (!bf16_vec(e.aux)) return EF_GENERIC;
    f |= (e.act_bwd == EGB_ACTBWD_RELU_MASK) ? EF_ABWD_RELU : (e.act_bwd == EGB_ACTBWD_MUL ? EF_ABWD_MUL : EF_ABWD_GELU);
  }
  if (e.colsum != nullptr) {
    if (((uintptr_t)e.colsum % 16) != 0) return EF_GENERIC;
    f |= EF_COLSUM;
  }
  switch (f) {   // the instantiated set (everything else runs the generic epilogue)
    case 0: case EF_BIAS: case EF_BIAS | EF_RES: case EF_BIAS | EF_RELU: case EF_BIAS | EF_GELU | EF_PRE | EF_DGELU:
    case EF_ABWD_RELU: case EF_ABWD_MUL: case EF_ABWD_RELU | EF_COLSUM: case EF_ABWD_MUL | EF_COLSUM:
      return f;
    default:
      return EF_GENERIC;
  }
}
