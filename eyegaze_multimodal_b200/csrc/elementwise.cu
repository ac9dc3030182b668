// Memory-bound glue kernels of the EEG path: dtype casts / weight re-layout, channels-last packing of the
// raw EEG, token-sequence assembly (+ positional embedding), pooling tail, bias-gradient column sums,
// dropout-mask regeneration and the fused cross-entropy.  All are coalesced, vectorised where the layout
// allows, and use warp shuffles for their reductions.
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);

namespace {

inline int grid_for(long long n, int block, int max_blocks = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------- casts
template <typename TO>
__global__ void cast_kernel(const float* __restrict__ src, TO* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float v[4];
      ld4(src + i, v);
      st4(dst + i, v);
    } else {
      for (long long j = i; j < n; ++j) dst[j] = from_f<TO>(src[j]);
    }
  }
}

template <typename TI>
__global__ void cast_to_f32_kernel(const TI* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = to_f(src[i]);
}

// dst[i0*d0+i1*d1+i2*d2+i3*d3] = src[i0*s0 + i1*s1 + i2*s2 + i3*s3]  (weight re-layouts, small copies)
template <typename TI, typename TO>
__global__ void copy_strided4_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int n0, int n1, int n2, int n3,
                                     long long s0, long long s1, long long s2, long long s3, long long d0, long long d1,
                                     long long d2, long long d3) {
  const long long n = (long long)n0 * n1 * n2 * n3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long r = i;
    const int i3 = (int)(r % n3); r /= n3;
    const int i2 = (int)(r % n2); r /= n2;
    const int i1 = (int)(r % n1); r /= n1;
    const int i0 = (int)r;
    dst[i0 * d0 + i1 * d1 + i2 * d2 + i3 * d3] = from_f<TO>(to_f(src[i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3]));
  }
}

// same-type copy whose innermost index is contiguous on both sides, 16 bytes per thread (all strides and both base
// addresses 16-byte aligned): the dense copy of dX[:, 1:, :] in the patch-embedding backward ran 212 us for 77 MB through
// the scalar kernel above (three 64-bit divisions per element)
__global__ void copy_strided4_vec16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n0, int n1, int n2,
                                           int n3v, long long s0, long long s1, long long s2, long long d0, long long d1,
                                           long long d2) {
  const long long n = (long long)n0 * n1 * n2 * n3v;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long r = i;
    const int i3 = (int)(r % n3v); r /= n3v;
    const int i2 = (int)(r % n2); r /= n2;
    const int i1 = (int)(r % n1); r /= n1;
    const int i0 = (int)r;
    dst[i0 * d0 + i1 * d1 + i2 * d2 + i3] = src[i0 * s0 + i1 * s1 + i2 * s2 + i3];
  }
}

// ---------------------------------------------------------------------------------------------- EEG packing
// (B,C,T) fp32 x2  ->  [2B, Tp, C] channels-last with `pad` zero rows in front and zero rows up to Tp behind.
// 32x32 shared-memory transpose so both the read (along T) and the write (along C) are coalesced.
template <typename TO>
__global__ void eeg_pack_kernel(const float* __restrict__ e1, const float* __restrict__ e2, TO* __restrict__ out, int B,
                                int C, int T, int pad, int Tp) {
  __shared__ float tile[32][33];
  const int s = blockIdx.z;
  const float* src = (s < B ? e1 + (long long)s * C * T : e2 + (long long)(s - B) * C * T);
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? src[(long long)c * T + t] : 0.f;
  }
  __syncthreads();
  TO* dst = out + (long long)s * Tp * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < C) dst[(long long)(t + pad) * C + c] = from_f<TO>(tile[threadIdx.x][i]);
  }
}

// ---------------------------------------------------------------------------------------------- sequence assembly
// X[s, l, :] = src(l) + pos[l, :]   with   l = 0: cls; 1..n_ibs: ibs tokens (shared by both streams);
// then C spectrogram tokens; then the T2 temporal-conv tokens.          (dual_eeg_transformer.py:1157-1179)
template <typename T>
__global__ void seq_assemble_kernel(const float* __restrict__ cls, const float* __restrict__ pos,
                                    const T* __restrict__ ibs, const T* __restrict__ spec, const T* __restrict__ h,
                                    T* __restrict__ out, int S, int B, int L, int D, int n_ibs, int n_spec, int n_h) {
  const int vec = D / 4;
  const long long total = (long long)S * L * vec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int d = (int)(i % vec) * 4;
    const long long row = i / vec;
    const int l = (int)(row % L);
    const int s = (int)(row / L);
    float v[4], p[4];
    ld4(pos + (long long)l * D + d, p);
    if (l == 0) {
      ld4(cls + d, v);
    } else if (l <= n_ibs) {
      ld4(ibs + ((long long)(s % B) * n_ibs + (l - 1)) * D + d, v);
    } else if (l <= n_ibs + n_spec) {
      ld4(spec + ((long long)s * n_spec + (l - 1 - n_ibs)) * D + d, v);
    } else {
      ld4(h + ((long long)s * n_h + (l - 1 - n_ibs - n_spec)) * D + d, v);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += p[j];
    st4(out + row * D + d, v);
  }
}

// dpos[l,:] += sum_s dX[s,l,:]  (cls and type-embedding gradients are rows of this sum);
// dibs[b,tok,:] = dX[b,1+tok,:] + dX[B+b,1+tok,:].
// grid = (L, S-splits): a CTA sums SEQ_BWD_SPLIT sequences of one position (eight independent 8-byte loads in flight
// per thread) and adds its partial row to dpos with fp32 atomics -- with one CTA per position walking all S sequences
// the kernel was a serial load chain on 139-197 CTAs (183 us for 40 MB).
constexpr int SEQ_BWD_SPLIT = 32;
template <typename T>
__global__ void seq_assemble_bwd_kernel(const T* __restrict__ dx, float* __restrict__ dpos, T* __restrict__ dibs, int S,
                                        int B, int L, int D, int n_ibs) {
  const int l = blockIdx.x;
  const int s0 = blockIdx.y * SEQ_BWD_SPLIT, s1 = min(S, s0 + SEQ_BWD_SPLIT);
  for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const T* src = dx + (long long)l * D + d;
    int s = s0;
    for (; s + 8 <= s1; s += 8) {
      float v[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u) ld4(src + (long long)(s + u) * L * D, v[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += v[u][j];
    }
    for (; s < s1; ++s) {
      float v[4];
      ld4(src + (long long)s * L * D, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += v[j];
    }
    float* o = dpos + (long long)l * D + d;
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(o + j, acc[j]);
  }
  if (dibs != nullptr && l >= 1 && l <= n_ibs) {
    for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < (long long)B * (D / 4); i += (long long)gridDim.y * blockDim.x) {
      const int b = (int)(i / (D / 4)), d = (int)(i % (D / 4)) * 4;
      float a[4], c[4];
      ld4(dx + ((long long)b * L + l) * D + d, a);
      ld4(dx + ((long long)(b + B) * L + l) * D + d, c);
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] += c[j];
      st4(dibs + ((long long)b * n_ibs + (l - 1)) * D + d, a);
    }
  }
}


// ---------------------------------------------------------------------------------------------- broadcast row add
// out[b, t, :] = x[b, t, :] + e[t, :]   (type embedding of the IBS tokens, dual_eeg_transformer.py:909)
template <typename T>
__global__ void add_rows_broadcast_kernel(const T* __restrict__ x, const float* __restrict__ e, T* __restrict__ out,
                                          long long rows, int NT, int D) {
  const int nv = D / 4;
  const long long total = rows * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int d = (int)(i % nv) * 4, t = (int)(r % NT);
    float a[4], b[4];
    ld4(x + r * D + d, a);
    ld4(e + (long long)t * D + d, b);
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] += b[j];
    st4(out + r * D + d, a);
  }
}

// ---------------------------------------------------------------------------------------------- pooling tail
// From Z [S=2B, L, D]: cls1/cls2, mean over tokens [offset, L) per stream, mean over the IBS tokens of
// stream 1, and the symmetric features [c1+c2, c1*c2, |c1-c2|].   (dual_eeg_transformer.py:1193-1225, 933-938)
template <typename T>
__global__ void tail_pool_kernel(const T* __restrict__ z, float* __restrict__ cls1, float* __restrict__ cls2,
                                 float* __restrict__ sym, float* __restrict__ zf, float* __restrict__ ibs_pool, int B,
                                 int L, int D, int n_ibs, int offset, int ibs_single) {
  const int b = blockIdx.x;
  const T* z1 = z + (long long)b * L * D;
  const T* z2 = z + (long long)(b + B) * L * D;
  const float inv_mp = 1.f / (float)(L - offset);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float c1 = to_f(z1[d]), c2 = to_f(z2[d]);
    cls1[(long long)b * D + d] = c1;
    cls2[(long long)b * D + d] = c2;
    float* sy = sym + (long long)b * 3 * D;
    sy[d] = c1 + c2;
    sy[D + d] = c1 * c2;
    sy[2 * D + d] = fabsf(c1 - c2);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll 8
    for (int l = offset; l < L; ++l) {          // (unrolled: sixteen loads in flight; the sums keep their order)
      m1 += to_f(z1[(long long)l * D + d]);
      m2 += to_f(z2[(long long)l * D + d]);
    }
    float* f = zf + (long long)b * 3 * D;
    f[D + d] = m1 * inv_mp;
    f[2 * D + d] = m2 * inv_mp;
    if (ibs_pool != nullptr) {
      float a = 0.f;
      if (ibs_single) {
        a = to_f(z1[(long long)D + d]);
      } else {
        for (int l = 1; l <= n_ibs; ++l) a += to_f(z1[(long long)l * D + d]);
        a /= (float)n_ibs;
      }
      ibs_pool[(long long)b * D + d] = a;
    }
  }
}

template <typename T>
__global__ void tail_pool_bwd_kernel(const T* __restrict__ z, const float* __restrict__ dcls1,
                                     const float* __restrict__ dcls2, const float* __restrict__ dsym,
                                     const float* __restrict__ dzf, const float* __restrict__ dibs_pool,
                                     T* __restrict__ dz, int B, int L, int D, int n_ibs, int offset, int ibs_single) {
  const int b = blockIdx.x;
  const T* z1 = z + (long long)b * L * D;
  const T* z2 = z + (long long)(b + B) * L * D;
  T* d1 = dz + (long long)b * L * D;
  T* d2 = dz + (long long)(b + B) * L * D;
  const float inv_mp = 1.f / (float)(L - offset);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float c1 = to_f(z1[d]), c2 = to_f(z2[d]);
    const float* ds = dsym + (long long)b * 3 * D;
    const float sg = (c1 > c2) ? 1.f : ((c1 < c2) ? -1.f : 0.f);
    float g1 = ds[d] + ds[D + d] * c2 + ds[2 * D + d] * sg;
    float g2 = ds[d] + ds[D + d] * c1 - ds[2 * D + d] * sg;
    if (dcls1 != nullptr) g1 += dcls1[(long long)b * D + d];
    if (dcls2 != nullptr) g2 += dcls2[(long long)b * D + d];
    d1[d] = from_f<T>(g1);
    d2[d] = from_f<T>(g2);
    const float gi = dibs_pool != nullptr ? dibs_pool[(long long)b * D + d] * (ibs_single ? 1.f : 1.f / (float)n_ibs) : 0.f;
    const float gm1 = dzf[(long long)b * 3 * D + D + d] * inv_mp;
    const float gm2 = dzf[(long long)b * 3 * D + 2 * D + d] * inv_mp;
    for (int l = 1; l < L; ++l) {
      const bool in_ibs = ibs_single ? (l == 1 && n_ibs > 0) : (l <= n_ibs);
      d1[(long long)l * D + d] = from_f<T>((in_ibs ? gi : 0.f) + (l >= offset ? gm1 : 0.f));
      d2[(long long)l * D + d] = from_f<T>(l >= offset ? gm2 : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------- bias gradient
// out[n] += sum_m X[m, n] over a strided 3-level row view.  Each CTA owns a 128-column panel and a
// slab of rows; partial sums go through shared memory, then one atomicAdd per column per CTA.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int M, int N, int rpg, long long rs,
                              long long gs, int rows_per_block) {
  __shared__ float part[8][128];
  const int n = blockIdx.x * 128 + (threadIdx.x & 127);
  const int ry = threadIdx.x >> 7;  // 0..1 (256 threads) -> 2 row lanes
  const int m0 = blockIdx.y * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float acc = 0.f;
  if (n < N) {
    for (int m = m0 + ry; m < m1; m += 2) {
      const int g = m / rpg;
      acc += to_f(x[(long long)g * gs + (long long)(m - g * rpg) * rs + n]);
    }
  }
  part[ry][threadIdx.x & 127] = acc;
  __syncthreads();
  if (ry == 0 && n < N) atomicAdd(out + n, part[0][threadIdx.x] + part[1][threadIdx.x]);
}


// Vectorised variant (rows 16-byte aligned): a thread owns V = 16/sizeof(T) consecutive columns, a warp spans 32*V
// columns of one row (512 contiguous bytes), the 8 warps of a CTA take every 8th row and keep 4 row loads in flight.
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, float* __restrict__ out, int M, int N,
                                                         int rpg, long long rs, long long gs, int rows_per_block) {
  constexpr int V = 16 / (int)sizeof(T);
  __shared__ float part[8][32 * V];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + cg) * V;
  const int m0 = blockIdx.y * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  if (n < N) {
    auto row_ptr = [&](int m) -> const T* {
      if (m < rpg) return x + (long long)m * rs + n;
      const int g = m / rpg;
      return x + (long long)g * gs + (long long)(m - g * rpg) * rs + n;
    };
    int m = m0 + rl;
    for (; m + 24 < m1; m += 32) {
      uint4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = __ldg(reinterpret_cast<const uint4*>(row_ptr(m + 8 * u)));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const T* e = reinterpret_cast<const T*>(&t[u]);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += to_f(e[i]);
      }
    }
    for (; m < m1; m += 8) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(row_ptr(m)));
      const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += to_f(e[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) part[rl][cg * V + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * V; c += 256) {
    const int col = blockIdx.x * 32 * V + c;
    if (col < N) {
      float sum = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) sum += part[r][c];
      atomicAdd(out + col, sum);
    }
  }
}

// ---------------------------------------------------------------------------------------------- dropout backward
// out[m,n] = dy[m,n] * keep(seed, m*N+n) / (1-p): regenerates the mask the GEMM epilogue applied.
template <typename T>
__global__ void dropout_bwd_kernel(const T* __restrict__ dy, T* __restrict__ out, long long n_elems,
                                   unsigned thresh, float scale, unsigned long long seed_in,
                                   const unsigned long long* epoch) {
  const unsigned long long seed = egb_mix_seed(seed_in, epoch);
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n_elems; i += stride) {
    float v[4];
    ld4(dy + i, v);
    unsigned st = drop_state(seed, (unsigned long long)i);     // i % 4 == 0: the four elements lie in one run
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = st >= thresh ? v[j] * scale : 0.f;
      st = drop_step(st);
    }
    st4(out + i, v);
  }
}


// Row-structured variant that also accumulates the column sums of its output (bias gradient): one warp per row, a lane
// owns 8 consecutive columns of every 256-column chunk, so its column sums stay in registers across rows.
template <typename T, int C>
__global__ void __launch_bounds__(256) dropout_bwd_colsum_kernel(const T* __restrict__ dy, T* __restrict__ out, int M, int N,
                                                                 unsigned thresh, float scale, unsigned long long seed_in,
                                                                 const unsigned long long* epoch, float* __restrict__ colsum) {
  const unsigned long long seed = egb_mix_seed(seed_in, epoch);
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float acc[C][8];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int d = c * 256 + lane * 8;
      if (d < N) {
        float v[8];
        const long long base = (long long)row * N + d;
        ld8(dy + base, v);
        drop_apply_run8(seed, (unsigned long long)base, thresh, scale, v);   // N % 8 == 0: base is run-aligned
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[c][j] += to_f(from_f<T>(v[j]));
        st8(out + base, v);
      }
    }
  }
  if (colsum != nullptr) {
    __shared__ float red[C * 256];
    for (int i = threadIdx.x; i < C * 256; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[c * 256 + j * 32 + lane], acc[c][j]);
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const int c = i >> 8, r = i & 255;
      atomicAdd(colsum + i, red[c * 256 + (r & 7) * 32 + (r >> 3)]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- activation backward
// out[m,n] = dy[m,n] * act'(aux[m,n]); dy is a strided 3-level row view, aux/out are dense [M,N].
//   mode 1: aux = forward output after relu (+dropout): factor = (aux != 0) * scale
//   mode 2: aux = forward pre-activation: factor = gelu'(aux) * scale
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ aux, T* __restrict__ out, int M, int N,
                               int rpg, long long rs, long long gs, int o_rpg, long long o_rs, long long o_gs, int mode,
                               float scale) {
  const int nv = N / 4;
  const long long total = (long long)M * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / nv), n = (int)(i % nv) * 4;
    const int g = m / rpg;
    float d[4], a[4];
    ld4(dy + (long long)g * gs + (long long)(m - g * rpg) * rs + n, d);
    ld4(aux + (long long)m * N + n, a);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = mode == 1 ? (a[j] != 0.f ? d[j] * scale : 0.f) : d[j] * gelu_erf_grad(a[j]) * scale;
    const int og = m / o_rpg;
    st4(out + (long long)og * o_gs + (long long)(m - og * o_rpg) * o_rs + n, d);
  }
}

// ---------------------------------------------------------------------------------------------- cross entropy
// loss = mean_b( logsumexp(x_b) - x_b[y_b] );  dlogits = (softmax - onehot) / B.  One warp per row.
__global__ void cross_entropy_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                     float* __restrict__ loss, float* __restrict__ dlogits, int B, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* x = logits + (long long)warp * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(x[c] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  const int y = (int)labels[warp];
  for (int c = lane; c < C; c += 32) dlogits[(long long)warp * C + c] = (expf(x[c] - lse) - (c == y ? 1.f : 0.f)) / (float)B;
  if (lane == 0) atomicAdd(loss, (lse - x[y]) / (float)B);
}

__global__ void scale_by_device_scalar_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                              float* __restrict__ out, long long n) {
  const float s = *g;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = x[i] * s;
}

}  // namespace

extern "C" {

int egb_cast_from_f32(const float* src, void* dst, int dtype, int64_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  const int g = grid_for((n + 3) / 4, 256);
  if (dtype == EGB_BF16)
    cast_kernel<bf16><<<g, 256, 0, st>>>(src, (bf16*)dst, n);
  else
    cast_kernel<float><<<g, 256, 0, st>>>(src, (float*)dst, n);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_cast_to_f32(const void* src, int dtype, float* dst, int64_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  const int g = grid_for(n, 256);
  if (dtype == EGB_BF16)
    cast_to_f32_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)src, dst, n);
  else
    cast_to_f32_kernel<float><<<g, 256, 0, st>>>((const float*)src, dst, n);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_copy_strided4(const void* src, int src_dtype, void* dst, int dst_dtype, const int32_t* sizes,
                      const int64_t* src_strides, const int64_t* dst_strides, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)sizes[0] * sizes[1] * sizes[2] * sizes[3];
  if (n <= 0) return 0;
  const int32_t* z = sizes;
  const int64_t *a = src_strides, *d = dst_strides;
  {
    const int per = src_dtype == EGB_F32 ? 4 : 8;      // elements per 16 bytes
    bool vec = src_dtype == dst_dtype && a[3] == 1 && d[3] == 1 && z[3] % per == 0 && (uintptr_t)src % 16 == 0 &&
               (uintptr_t)dst % 16 == 0 && n >= (1 << 16);
    for (int i = 0; i < 3; ++i) vec = vec && a[i] % per == 0 && d[i] % per == 0;
    if (vec) {
      const int gv = grid_for(n / per, 256);
      copy_strided4_vec16_kernel<<<gv, 256, 0, st>>>((const uint4*)src, (uint4*)dst, z[0], z[1], z[2], z[3] / per, a[0] / per,
                                                    a[1] / per, a[2] / per, d[0] / per, d[1] / per, d[2] / per);
      egb_count_launch(1);
      EGB_LAUNCH_CHECK();
      return 0;
    }
  }
  const int g = grid_for(n, 256);
#define EGB_CP(TI, TO)                                                                                              \
  copy_strided4_kernel<TI, TO><<<g, 256, 0, st>>>((const TI*)src, (TO*)dst, z[0], z[1], z[2], z[3], a[0], a[1], a[2], \
                                                  a[3], d[0], d[1], d[2], d[3])
  if (src_dtype == EGB_F32 && dst_dtype == EGB_F32) EGB_CP(float, float);
  else if (src_dtype == EGB_F32) EGB_CP(float, bf16);
  else if (dst_dtype == EGB_F32) EGB_CP(bf16, float);
  else EGB_CP(bf16, bf16);
#undef EGB_CP
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_zero(void* ptr, int64_t nbytes, void* stream) {
  if (nbytes <= 0) return 0;
  EGB_CUDA(cudaMemsetAsync(ptr, 0, (size_t)nbytes, (cudaStream_t)stream));
  return 0;
}

int egb_eeg_pack(const float* eeg1, const float* eeg2, void* out, int dtype, int B, int C, int T, int pad, int Tp,
                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(B > 0 && C > 0 && T > 0 && Tp >= T + pad, "eeg_pack: bad shape");
  const size_t esz = dtype == EGB_BF16 ? 2 : 4;
  EGB_CUDA(cudaMemsetAsync(out, 0, (size_t)2 * B * Tp * C * esz, st));
  dim3 grid((T + 31) / 32, (C + 31) / 32, 2 * B), block(32, 8);
  if (dtype == EGB_BF16)
    eeg_pack_kernel<bf16><<<grid, block, 0, st>>>(eeg1, eeg2, (bf16*)out, B, C, T, pad, Tp);
  else
    eeg_pack_kernel<float><<<grid, block, 0, st>>>(eeg1, eeg2, (float*)out, B, C, T, pad, Tp);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_seq_assemble_fwd(const float* cls, const float* pos, const void* ibs, const void* spec, const void* h,
                         void* out, int dtype, int S, int B, int L, int D, int n_ibs, int n_spec, int n_h,
                         void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(D % 4 == 0 && L == 1 + n_ibs + n_spec + n_h, "seq_assemble: L=%d != 1+%d+%d+%d or D%%4", L, n_ibs, n_spec,
            n_h);
  EGB_CHECK((n_ibs == 0 || ibs) && (n_spec == 0 || spec) && (n_h == 0 || h), "seq_assemble: missing source");
  const int g = grid_for((long long)S * L * (D / 4), 256);
  if (dtype == EGB_BF16)
    seq_assemble_kernel<bf16><<<g, 256, 0, st>>>(cls, pos, (const bf16*)ibs, (const bf16*)spec, (const bf16*)h,
                                                (bf16*)out, S, B, L, D, n_ibs, n_spec, n_h);
  else
    seq_assemble_kernel<float><<<g, 256, 0, st>>>(cls, pos, (const float*)ibs, (const float*)spec, (const float*)h,
                                                 (float*)out, S, B, L, D, n_ibs, n_spec, n_h);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_seq_assemble_bwd(const void* dx, float* dpos, void* dibs, int dtype, int S, int B, int L, int D, int n_ibs,
                         void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(D % 4 == 0, "seq_assemble_bwd: D%%4");
  const int threads = D / 4 < 64 ? 64 : (D / 4 > 256 ? 256 : D / 4);
  const dim3 grid(L, (S + SEQ_BWD_SPLIT - 1) / SEQ_BWD_SPLIT);
  if (dtype == EGB_BF16)
    seq_assemble_bwd_kernel<bf16><<<grid, threads, 0, st>>>((const bf16*)dx, dpos, (bf16*)dibs, S, B, L, D, n_ibs);
  else
    seq_assemble_bwd_kernel<float><<<grid, threads, 0, st>>>((const float*)dx, dpos, (float*)dibs, S, B, L, D, n_ibs);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_add_rows_broadcast(const void* x, const float* e, void* out, int dtype, int64_t rows, int NT, int D, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(D % 4 == 0 && NT > 0, "add_rows_broadcast: D%%4");
  const int g = grid_for(rows * (D / 4), 256);
  if (dtype == EGB_BF16)
    add_rows_broadcast_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, e, (bf16*)out, rows, NT, D);
  else
    add_rows_broadcast_kernel<float><<<g, 256, 0, st>>>((const float*)x, e, (float*)out, rows, NT, D);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_tail_pool_fwd(const void* z, int dtype, float* cls1, float* cls2, float* sym, float* zf, float* ibs_pool, int B,
                      int L, int D, int n_ibs, int offset, int ibs_single, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(offset < L, "tail_pool: no temporal tokens (offset %d >= L %d)", offset, L);
  const int threads = D < 256 ? ((D + 31) / 32) * 32 : 256;
  if (dtype == EGB_BF16)
    tail_pool_kernel<bf16><<<B, threads, 0, st>>>((const bf16*)z, cls1, cls2, sym, zf, ibs_pool, B, L, D, n_ibs, offset,
                                                  ibs_single);
  else
    tail_pool_kernel<float><<<B, threads, 0, st>>>((const float*)z, cls1, cls2, sym, zf, ibs_pool, B, L, D, n_ibs,
                                                   offset, ibs_single);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_tail_pool_bwd(const void* z, int dtype, const float* dcls1, const float* dcls2, const float* dsym,
                      const float* dzf, const float* dibs_pool, void* dz, int B, int L, int D, int n_ibs, int offset,
                      int ibs_single, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = D < 256 ? ((D + 31) / 32) * 32 : 256;
  if (dtype == EGB_BF16)
    tail_pool_bwd_kernel<bf16><<<B, threads, 0, st>>>((const bf16*)z, dcls1, dcls2, dsym, dzf, dibs_pool, (bf16*)dz, B,
                                                      L, D, n_ibs, offset, ibs_single);
  else
    tail_pool_bwd_kernel<float><<<B, threads, 0, st>>>((const float*)z, dcls1, dcls2, dsym, dzf, dibs_pool, (float*)dz,
                                                       B, L, D, n_ibs, offset, ibs_single);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_colsum(const egb_matrix* x, int M, int N, float* out, int zero_first, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(M > 0 && N > 0, "colsum: empty");
  if (zero_first) EGB_CUDA(cudaMemsetAsync(out, 0, (size_t)N * 4, st));
  const int rpg = x->rows_per_group > 0 ? x->rows_per_group : M;
  {
    const int V = x->dtype == EGB_BF16 ? 8 : 4;
    const bool vec = (N % V) == 0 && ((uintptr_t)x->ptr % 16) == 0 && (x->row_stride % V) == 0 &&
                     (rpg >= M || (x->group_stride % V) == 0);
    if (vec) {
      const int cb = (N + 32 * V - 1) / (32 * V);
      int rb = (6 * egb_num_sms() + cb - 1) / cb;
      if (rb > (M + 63) / 64) rb = (M + 63) / 64;
      if (rb < 1) rb = 1;
      const int rpb = ((M + rb - 1) / rb + 7) / 8 * 8;
      dim3 grid(cb, (M + rpb - 1) / rpb);
      if (x->dtype == EGB_BF16)
        colsum_vec_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x->ptr, out, M, N, rpg, x->row_stride, x->group_stride, rpb);
      else
        colsum_vec_kernel<float><<<grid, 256, 0, st>>>((const float*)x->ptr, out, M, N, rpg, x->row_stride, x->group_stride, rpb);
      egb_count_launch(1);
      EGB_LAUNCH_CHECK();
      return 0;
    }
  }
  const int col_blocks = (N + 127) / 128;
  int row_blocks = (4 * egb_num_sms() + col_blocks - 1) / col_blocks;
  if (row_blocks > (M + 63) / 64) row_blocks = (M + 63) / 64;
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = (M + row_blocks - 1) / row_blocks;
  dim3 grid(col_blocks, (M + rows_per_block - 1) / rows_per_block);
  if (x->dtype == EGB_BF16)
    colsum_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x->ptr, out, M, N, rpg, x->row_stride, x->group_stride,
                                              rows_per_block);
  else
    colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)x->ptr, out, M, N, rpg, x->row_stride, x->group_stride,
                                               rows_per_block);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_act_bwd(const egb_matrix* dy, const void* aux, const egb_matrix* out, int M, int N, int mode, float scale,
                void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(N % 4 == 0 && dy->row_stride % 4 == 0 && dy->group_stride % 4 == 0 && out->row_stride % 4 == 0 &&
                out->group_stride % 4 == 0,
            "act_bwd: N and strides must be multiples of 4");
  EGB_CHECK(mode == 1 || mode == 2, "act_bwd: bad mode");
  EGB_CHECK(dy->dtype == out->dtype, "act_bwd: dtype mismatch");
  const int rpg = dy->rows_per_group > 0 ? dy->rows_per_group : M;
  const int o_rpg = out->rows_per_group > 0 ? out->rows_per_group : M;
  const int g = grid_for((long long)M * (N / 4), 256);
  if (dy->dtype == EGB_BF16)
    act_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)dy->ptr, (const bf16*)aux, (bf16*)out->ptr, M, N, rpg,
                                           dy->row_stride, dy->group_stride, o_rpg, out->row_stride, out->group_stride,
                                           mode, scale);
  else
    act_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)dy->ptr, (const float*)aux, (float*)out->ptr, M, N, rpg,
                                            dy->row_stride, dy->group_stride, o_rpg, out->row_stride, out->group_stride,
                                            mode, scale);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_dropout_bwd(const void* dy, void* out, int dtype, int64_t n_elems, float p, uint64_t seed, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(n_elems % 4 == 0 && p > 0.f && p < 1.f, "dropout_bwd: bad arguments");
  const int g = grid_for(n_elems / 4, 256);
  if (dtype == EGB_BF16)
    dropout_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)dy, (bf16*)out, n_elems, drop_threshold(p), 1.f / (1.f - p),
                                               seed, egb_seed_epoch_ptr());
  else
    dropout_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)dy, (float*)out, n_elems, drop_threshold(p),
                                                1.f / (1.f - p), seed, egb_seed_epoch_ptr());
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_dropout_bwd_colsum(const void* dy, void* out, int dtype, int M, int N, float p, uint64_t seed, float* colsum,
                           void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(M > 0 && N > 0 && N % 8 == 0 && N <= 1024 && p > 0.f && p < 1.f, "dropout_bwd_colsum: bad arguments");
  const int chunks = (N + 255) / 256;
  int blocks = (M + 7) / 8;
  const int cap = egb_num_sms() * 4;       // few, fat CTAs: each ends with N column atomics
  if (blocks > cap) blocks = cap;
  const unsigned th = drop_threshold(p);
  const float sc = 1.f / (1.f - p);
  const unsigned long long* ep = egb_seed_epoch_ptr();
#define EGB_DBC(TT, CC) \
  dropout_bwd_colsum_kernel<TT, CC><<<blocks, 256, 0, st>>>((const TT*)dy, (TT*)out, M, N, th, sc, seed, ep, colsum)
  if (dtype == EGB_BF16) {
    switch (chunks) { case 1: EGB_DBC(bf16, 1); break; case 2: EGB_DBC(bf16, 2); break; case 3: EGB_DBC(bf16, 3); break; default: EGB_DBC(bf16, 4); }
  } else {
    switch (chunks) { case 1: EGB_DBC(float, 1); break; case 2: EGB_DBC(float, 2); break; case 3: EGB_DBC(float, 3); break; default: EGB_DBC(float, 4); }
  }
#undef EGB_DBC
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_cross_entropy(const float* logits, const int64_t* labels, float* loss, float* dlogits, int B, int C,
                      void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGB_CHECK(B > 0 && C > 0, "cross_entropy: empty");
  EGB_CUDA(cudaMemsetAsync(loss, 0, 4, st));
  const int threads = 128;
  cross_entropy_kernel<<<(B * 32 + threads - 1) / threads, threads, 0, st>>>(logits, (const long long*)labels, loss,
                                                                            dlogits, B, C);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_scale_by_device_scalar(const float* x, const float* g, float* out, int64_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  scale_by_device_scalar_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, g, out, n);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
