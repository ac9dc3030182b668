// egb_gemm dispatcher + the FP32-FMA GEMM (CUDA cores).
//
// Two precision modes share one descriptor and one epilogue:
//   in_dtype = EGB_BF16 : tcgen05 tensor-core kernel (gemm_tc.cu), fp32 accumulation in TMEM
//   in_dtype = EGB_F32  : this file's register-tiled FFMA kernel -- the fp32-parity mode whose
//                         logits must match the reference CPU model to <= 1e-4.
// bf16 problems the tensor-core kernel cannot tile (N < 16, strides not 16-byte aligned) are
// also routed here (bf16 in, fp32 math); they are the 3-class heads, a few kFLOP each.
#include <string.h>
#include <stdlib.h>
#include <cooperative_groups.h>
#include "common.cuh"
#include "epilogue.cuh"

extern void egb_count_launch(int n);
int egb_gemm_tc(const egb_gemm_desc* d, cudaStream_t stream);

namespace {

constexpr int FBM = 128, FBN = 128, FBK = 16, FTHREADS = 256;

struct FOperand {
  const char* ptr;
  int major;
  int rpg;
  long long rs, gs;
  int seg, shift;  // inner segmentation (0 = off)
  int vec;  // 4-element vector loads are legal
};

struct FParams {
  int M, N, K;
  FOperand a, b;
  int split_k, k_per_split;
  int cluster_k;       // 1: the split_k CTAs of a tile form a thread-block cluster (1, 1, split_k) and reduce through DSMEM
  EpiParams epi;
};

__device__ __forceinline__ long long op_row_offset(const FOperand& o, int row) {
  const int g = row / o.rpg;
  return (long long)g * o.gs + (long long)(row - g * o.rpg) * o.rs;
}

template <typename T>
__device__ __forceinline__ void load4_guard(const FOperand& o, long long off, int inner0, int inner_extent, bool row_ok,
                                            float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (!row_ok) return;
  const T* p = reinterpret_cast<const T*>(o.ptr) + off;
  if (o.seg > 0) {  // segment s of the inner index lives `shift` rows further down
    const int s = inner0 / o.seg;
    p += (long long)s * o.shift * o.rs + (inner0 - s * o.seg);
  } else {
    p += inner0;
  }
  if (o.vec && inner0 + 3 < inner_extent) {
    ld4(p, v);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (inner0 + i < inner_extent) v[i] = to_f(p[i]);
  }
}

// Loads this thread's share (2 x 4 elements) of a [128 mn x 16 k] operand tile.
template <typename T>
__device__ __forceinline__ void load_tile(const FOperand& o, int mn0, int extent_mn, int k0, int k_end,
                                          float (&r)[2][4]) {
  const int t = threadIdx.x;
  if (o.major == 0) {
    const int kq = (t & 3) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = mn0 + (t >> 2) + 64 * i;
      const bool ok = row < extent_mn;
      load4_guard<T>(o, ok ? op_row_offset(o, row) : 0, k0 + kq, k_end, ok, r[i]);
    }
  } else {
    const int mq = (t & 31) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = k0 + (t >> 5) + 8 * i;
      const bool ok = k < k_end;
      load4_guard<T>(o, ok ? op_row_offset(o, k) : 0, mn0 + mq, extent_mn, ok, r[i]);
    }
  }
}

__device__ __forceinline__ void store_tile(const FOperand& o, float (*s)[FBM + 4], const float (&r)[2][4]) {
  const int t = threadIdx.x;
  if (o.major == 0) {
    const int kq = (t & 3) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = (t >> 2) + 64 * i;
#pragma unroll
      for (int j = 0; j < 4; ++j) s[kq + j][row] = r[i][j];
    }
  } else {
    const int mq = (t & 31) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = (t >> 5) + 8 * i;
      *reinterpret_cast<float4*>(&s[k][mq]) = make_float4(r[i][0], r[i][1], r[i][2], r[i][3]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(FTHREADS) gemm_fma_kernel(const FParams p) {
  __shared__ __align__(16) float As[2][FBK][FBM + 4];
  __shared__ __align__(16) float Bs[2][FBK][FBN + 4];
  const int m0 = blockIdx.y * FBM;
  const int n0 = blockIdx.x * FBN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[2][4], rb[2][4];
  load_tile<T>(p.a, m0, p.M, kbeg, kend, ra);
  load_tile<T>(p.b, n0, p.N, kbeg, kend, rb);
  store_tile(p.a, As[0], ra);
  store_tile(p.b, Bs[0], rb);
  __syncthreads();

  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += FBK) {
    const bool has_next = k0 + FBK < kend;
    if (has_next) {
      load_tile<T>(p.a, m0, p.M, k0 + FBK, kend, ra);
      load_tile<T>(p.b, n0, p.N, k0 + FBK, kend, rb);
    }
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      store_tile(p.a, As[buf ^ 1], ra);
      store_tile(p.b, Bs[buf ^ 1], rb);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      epi_apply_store<4>(p.epi, m, n0 + h * 64 + tx * 4, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 64 x 64 tiles for SMALL problems (the fp32 heads / fusion tail: M = batch, N <= 768, K <= 768).  With 128 x 128
// tiles those launches occupy 4-12 SMs and every K step costs a thread 1024 FMAs (measured 170 us for 256x256x768);
// a quarter of the work per thread on four times as many CTAs brings the chain of dependent launches at the end
// of the forward / start of the backward pass down by ~3x.  Same operand views, same epilogue.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SBM = 64, SBN = 64;

// this thread's 4 elements of a [64 mn x 16 k] operand tile
template <typename T>
__device__ __forceinline__ void load_tile64(const FOperand& o, int mn0, int extent_mn, int k0, int k_end, float (&r)[4]) {
  const int t = threadIdx.x;
  if (o.major == 0) {
    const int kq = (t & 3) * 4;
    const int row = mn0 + (t >> 2);
    const bool ok = row < extent_mn;
    load4_guard<T>(o, ok ? op_row_offset(o, row) : 0, k0 + kq, k_end, ok, r);
  } else {
    const int mq = (t & 15) * 4;
    const int k = k0 + (t >> 4);
    const bool ok = k < k_end;
    load4_guard<T>(o, ok ? op_row_offset(o, k) : 0, mn0 + mq, extent_mn, ok, r);
  }
}
__device__ __forceinline__ void store_tile64(const FOperand& o, float (*s)[SBM + 4], const float (&r)[4]) {
  const int t = threadIdx.x;
  if (o.major == 0) {
    const int kq = (t & 3) * 4;
    const int row = t >> 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) s[kq + j][row] = r[j];
  } else {
    *reinterpret_cast<float4*>(&s[t >> 4][(t & 15) * 4]) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

template <typename T>
__global__ void __launch_bounds__(FTHREADS) gemm_fma_small_kernel(const FParams p) {
  __shared__ __align__(16) float As[2][FBK][SBM + 4];
  __shared__ __align__(16) float Bs[2][FBK][SBN + 4];
  __shared__ __align__(16) float red[SBM * SBN];   // partial tile of a cluster rank (cluster split-K)
  const int m0 = blockIdx.y * SBM;
  const int n0 = blockIdx.x * SBN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  load_tile64<T>(p.a, m0, p.M, kbeg, kend, ra);
  load_tile64<T>(p.b, n0, p.N, kbeg, kend, rb);
  store_tile64(p.a, As[0], ra);
  store_tile64(p.b, Bs[0], rb);
  __syncthreads();

  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += FBK) {
    const bool has_next = k0 + FBK < kend;
    if (has_next) {
      load_tile64<T>(p.a, m0, p.M, k0 + FBK, kend, ra);
      load_tile64<T>(p.b, n0, p.N, k0 + FBK, kend, rb);
    }
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[4] = {a0.x, a0.y, a0.z, a0.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      store_tile64(p.a, As[buf ^ 1], ra);
      store_tile64(p.b, Bs[buf ^ 1], rb);
      __syncthreads();
      buf ^= 1;
    }
  }
  if (p.cluster_k) {
    // K split over the CTAs of a cluster: ranks > 0 park their partial tile in their own shared memory, rank 0 adds them
    // in rank order over distributed shared memory and runs the epilogue (bias / activation / dropout need the full sum)
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    if (rank != 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(&red[(ty * 4 + i) * SBN + tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    cl.sync();
    if (rank == 0) {
      const unsigned nb = cl.num_blocks();
      for (unsigned r = 1; r < nb; ++r) {
        const float* peer = cl.map_shared_rank(red, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(&peer[(ty * 4 + i) * SBN + tx * 4]);
          acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
        }
      }
    }
    cl.sync();                                       // peers keep their shared memory alive until rank 0 has read it
    if (rank != 0) return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    epi_apply_store<4>(p.epi, m0 + ty * 4 + i, n0 + tx * 4, v);
  }
}

FOperand make_foperand(const egb_operand& o, int extent_mn, int K, int esz) {
  FOperand f;
  f.ptr = (const char*)o.ptr;
  f.major = o.major;
  const long long total_rows = o.major == 0 ? extent_mn : K;
  f.rpg = (int)(o.rows_per_group > 0 ? o.rows_per_group : (total_rows > 0 ? total_rows : 1));
  f.rs = o.row_stride;
  f.gs = o.group_stride;
  f.seg = o.seg_len;
  f.shift = o.seg_row_shift;
  f.vec = ((uintptr_t)o.ptr % (4 * esz) == 0) && (o.row_stride % 4 == 0) && (o.group_stride % 4 == 0);
  return f;
}

bool tc_compatible(const egb_gemm_desc* d) {
  // `span` = rows of the operand view one tile covers; groups must tile evenly into it
  auto ok = [&](const egb_operand& o, int extent_mn, int tile_rows) {
    if (((uintptr_t)o.ptr % 16) != 0 || (o.row_stride % 8) != 0 || o.row_stride <= 0) return false;
    const long long total = o.major == 0 ? extent_mn : d->K;
    if (o.rows_per_group > 0 && o.rows_per_group < total) {
      if (o.group_stride % 8 != 0 || o.seg_len > 0) return false;
      const int span = o.major == 0 ? tile_rows : 64;
      const int rpg = o.rows_per_group;
      if (rpg >= span ? (rpg % span != 0) : (span % rpg != 0)) return false;
    }
    if (o.seg_len > 0 && o.seg_len % 64 != 0) return false;
    return true;
  };
  return d->N >= 16 && d->K >= 16 && ok(d->a, d->M, 128) && ok(d->b, d->N, 256) && ok(d->b, d->N, 128) &&
         ok(d->b, d->N, 64);
}

}  // namespace

int egb_fill_epilogue(const egb_gemm_desc* d, EpiParams* e) {
  static const int epi_exp = getenv("EGB_EPI_EXP") ? atoi(getenv("EGB_EPI_EXP")) : 0;
  memset(e, 0, sizeof(*e));
  e->M = d->M;
  e->N = d->N;
  EGB_CHECK(d->c.ptr != nullptr, "gemm: null output");
  e->c = make_epimat(d->c, d->M);
  e->c_pre = make_epimat(d->c_pre, d->M);
  e->res = make_epimat(d->residual, d->M);
  e->aux = make_epimat(d->aux, d->M);
  e->bias = d->bias;
  e->alpha = d->alpha;
  e->act = d->act;
  e->act_bwd = d->act_bwd;
  e->exp = epi_exp;
  e->colsum = d->c_colsum;
  EGB_CHECK(d->act_bwd == EGB_ACTBWD_NONE || d->aux.ptr != nullptr, "gemm: act_bwd needs aux");
  EGB_CHECK(d->act != EGB_ACT_GELU_DGRAD || d->c_pre.ptr != nullptr, "gemm: EGB_ACT_GELU_DGRAD needs c_pre");
  e->aux_scale = d->aux_scale;
  if (d->dropout_p > 0.f) {
    EGB_CHECK(d->dropout_p < 1.f, "gemm: dropout_p must be < 1");
    e->drop_thresh = drop_threshold(d->dropout_p);
    e->drop_scale = 1.f / (1.f - d->dropout_p);
    e->seed = d->dropout_seed;
    e->epoch = egb_seed_epoch_ptr();
  }
  e->accumulate = d->accumulate;
  EGB_CHECK(!d->accumulate || d->c.dtype == EGB_F32, "gemm: accumulate requires fp32 output");
  EGB_CHECK(!d->accumulate || (d->act == 0 && d->act_bwd == 0 && d->residual.ptr == nullptr && d->bias == nullptr &&
                               d->c_pre.ptr == nullptr && d->dropout_p == 0.f),
            "gemm: accumulate mode supports only alpha scaling");
  return 0;
}

extern "C" int egb_gemm(const egb_gemm_desc* d, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EGB_CHECK(d != nullptr, "gemm: null descriptor");
  EGB_CHECK(d->M > 0 && d->N > 0 && d->K > 0, "gemm: empty problem %dx%dx%d", d->M, d->N, d->K);
  EGB_CHECK(d->a.ptr && d->b.ptr, "gemm: null operand");
  if (d->accumulate == 2) {  // zero-initialise the (dense or strided) fp32 output, then split-K accumulate
    EGB_CHECK(d->c.dtype == EGB_F32 && d->c.ptr, "gemm: accumulate requires an fp32 output");
    const int rpg = d->c.rows_per_group > 0 ? d->c.rows_per_group : d->M;
    if (rpg >= d->M && d->c.row_stride == d->N) {
      EGB_CUDA(cudaMemsetAsync(d->c.ptr, 0, (size_t)d->M * d->N * 4, stream));
    } else {
      EGB_CHECK(rpg >= d->M, "gemm: zero-init of grouped outputs is not supported");
      EGB_CUDA(cudaMemset2DAsync(d->c.ptr, (size_t)d->c.row_stride * 4, 0, (size_t)d->N * 4, (size_t)d->M, stream));
    }
  }
  if (d->c_colsum != nullptr) {
    // fused in the tensor-core kernels' specialised epilogues; every other route takes a second pass over C
    bool fused = false;
    if (d->in_dtype == EGB_BF16 && tc_compatible(d)) {
      EpiParams e;
      if (egb_fill_epilogue(d, &e)) return 1;
      const int mask = egb_epi_fast_mask(e);
      fused = mask != EF_GENERIC && (mask & EF_COLSUM) != 0;
    }
    if (!fused) {
      egb_gemm_desc d2 = *d;
      d2.c_colsum = nullptr;
      d2.accumulate = d->accumulate == 2 ? 1 : d->accumulate;   // the zero-init above already happened
      if (egb_gemm(&d2, stream_)) return 1;
      return egb_colsum(&d->c, d->M, d->N, d->c_colsum, 0, stream_);
    }
  }
  if (d->in_dtype == EGB_BF16 && tc_compatible(d)) return egb_gemm_tc(d, stream);

  FParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->M; p.N = d->N; p.K = d->K;
  const int esz = d->in_dtype == EGB_BF16 ? 2 : 4;
  p.a = make_foperand(d->a, d->M, d->K, esz);
  p.b = make_foperand(d->b, d->N, d->K, esz);
  if (egb_fill_epilogue(d, &p.epi)) return 1;
  // small problems (fewer 128 x 128 tiles than half the SMs): 64 x 64 tiles, four times the CTAs
  static const int small_ok = getenv("EGB_GEMM_FMA_SMALL") ? atoi(getenv("EGB_GEMM_FMA_SMALL")) : 1;
  const bool small = small_ok && ((d->M + FBM - 1) / FBM) * ((d->N + FBN - 1) / FBN) * 2 < egb_num_sms();
  const int bm = small ? SBM : FBM, bn = small ? SBN : FBN;
  const int mt = (d->M + bm - 1) / bm, nt = (d->N + bn - 1) / bn;
  int split = 1;
  if (d->accumulate) {
    split = d->split_k;
    if (split <= 0) {
      split = (2 * egb_num_sms() + mt * nt - 1) / (mt * nt);
      // the small launches are chains of K steps on a handful of SMs: split down to 64 per CTA
      static const int split64 = getenv("EGB_GEMM_FMA_SPLIT64") ? atoi(getenv("EGB_GEMM_FMA_SPLIT64")) : 1;
      const int max_split = (small && split64) ? (d->K + 63) / 64 : (d->K + 255) / 256;
      if (split > max_split) split = max_split;
      if (split < 1) split = 1;
    }
  } else if (small) {
    // no accumulation into C (bias / activation / dropout epilogues): split K over a thread-block cluster instead; the
    // fp32 heads of a step (M = batch, K <= 768) otherwise run 4-48 CTAs for 20-75 us each at the serial tail of the step
    static const int cluster_ok = getenv("EGB_GEMM_FMA_CLUSTER") ? atoi(getenv("EGB_GEMM_FMA_CLUSTER")) : 1;
    if (cluster_ok && d->split_k == 0)      // split_k = 1 on a non-accumulating launch: one CTA walks K in order (parity mode)
      while (split < 8 && d->K >= 128 * split && mt * nt * split * 2 <= 2 * egb_num_sms()) split *= 2;
    p.cluster_k = split > 1;
  }
  int kps = (d->K + split - 1) / split;
  kps = ((kps + FBK - 1) / FBK) * FBK;
  p.k_per_split = kps;
  p.split_k = (d->K + kps - 1) / kps;
  if (p.cluster_k && p.split_k != split) {            // (K not a multiple of 16 * split: fall back to one CTA per tile)
    p.cluster_k = 0;
    p.k_per_split = ((d->K + FBK - 1) / FBK) * FBK;
    p.split_k = 1;
  }
  dim3 grid(nt, mt, p.split_k);
  EGB_CHECK(mt <= 65535 && p.split_k <= 65535, "gemm: grid too large");
  if (small && p.cluster_k) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(FTHREADS, 1, 1);
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = (unsigned)p.split_k;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (d->in_dtype == EGB_BF16) EGB_CUDA(cudaLaunchKernelEx(&cfg, gemm_fma_small_kernel<bf16>, p));
    else EGB_CUDA(cudaLaunchKernelEx(&cfg, gemm_fma_small_kernel<float>, p));
  } else if (small) {
    if (d->in_dtype == EGB_BF16) gemm_fma_small_kernel<bf16><<<grid, FTHREADS, 0, stream>>>(p);
    else gemm_fma_small_kernel<float><<<grid, FTHREADS, 0, stream>>>(p);
  } else if (d->in_dtype == EGB_BF16)
    gemm_fma_kernel<bf16><<<grid, FTHREADS, 0, stream>>>(p);
  else
    gemm_fma_kernel<float><<<grid, FTHREADS, 0, stream>>>(p);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}
