// Fused multi-head attention, FP32-math path (CUDA cores):  O = dropout(softmax(Q K^T * scale)) V
// (art.py:203-213; timm Attention).  The (Lq x Lk) probability matrix never touches HBM: K/V of one
// (batch, head) live in shared memory, one warp owns one query row, softmax statistics come from warp
// shuffles, and only the row log-sum-exp is saved for the backward.  Used by the fp32-parity mode and
// for shapes the tensor-core kernel (attention_tc.cu) does not cover.
//
// Cross attention between the two players (dual_eeg_transformer.py:966-974) is the same kernel with
// kv_shift = B: query batch s reads keys/values of batch (s + kv_shift) % S, so both directions run
// as one launch over the stacked 2B batch.
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);
// tensor-core (tcgen05) path, attention_tc.cu
bool egb_attention_tc_supported(const egb_attention_desc* d, bool backward);
int egb_attention_tc_fwd(const egb_attention_desc* d, cudaStream_t st);
int egb_attention_tc_bwd(const egb_attention_desc* d, cudaStream_t st);

namespace {

constexpr int ATT_THREADS = 128;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int ATT_ROWS = 32;  // query (or key) rows per CTA
constexpr int ATT_MAXJ = 16;  // Lk (Lq in kernel B) <= 512

struct AttParams {
  const char *q, *k, *v;
  char* o;
  const char* d_o;
  char *dq, *dk, *dv;
  float* lse;    // [S,H,Lq]
  float* delta;  // [S,H,Lq]
  float* probs;  // optional [S,H,Lq,Lk]
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;  // elements
  long long dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs, do_bs, do_rs;
  int S, H, Lq, Lk, dk_dim, kv_shift;
  float scale;
  unsigned drop_thresh;
  float drop_scale;
  unsigned long long seed;
  const unsigned long long* epoch;   // device seed epoch (egb_mix_seed), NULL when not enabled
};

template <typename T>
__device__ __forceinline__ void load_rows_to_smem(float* dst, int ld, const T* src, long long rs, int rows, int dk) {
  const int per_row = dk / 4;
  for (int i = threadIdx.x; i < rows * per_row; i += blockDim.x) {
    const int r = i / per_row, d = (i % per_row) * 4;
    float v[4];
    ld4(src + (long long)r * rs + d, v);
    float* o = dst + r * ld + d;
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
  }
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attention_fwd_kernel(const AttParams p) {
  const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);   // dropout seed of THIS replay
  extern __shared__ float sm[];
  const int dk = p.dk_dim, ldk = dk + 1;
  float* Ks = sm;                       // [Lk][dk+1]
  float* Vs = Ks + p.Lk * ldk;          // [Lk][dk]
  float* qb = Vs + p.Lk * dk;           // [warps][dk]
  float* pb = qb + ATT_WARPS * dk;      // [warps][Lk]
  const int s = blockIdx.z, h = blockIdx.y;
  const int skv = (s + p.kv_shift) % p.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_rows_to_smem<T>(Ks, ldk, reinterpret_cast<const T*>(p.k) + skv * p.k_bs + h * dk, p.k_rs, p.Lk, dk);
  load_rows_to_smem<T>(Vs, dk, reinterpret_cast<const T*>(p.v) + skv * p.v_bs + h * dk, p.v_rs, p.Lk, dk);
  __syncthreads();
  const int i_end = min(p.Lq, (blockIdx.x + 1) * ATT_ROWS);
  for (int i = blockIdx.x * ATT_ROWS + warp; i < i_end; i += ATT_WARPS) {
    const T* qrow = reinterpret_cast<const T*>(p.q) + s * p.q_bs + (long long)i * p.q_rs + h * dk;
    for (int d = lane; d < dk; d += 32) qb[warp * dk + d] = to_f(qrow[d]) * p.scale;
    __syncwarp();
    float sc[ATT_MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXJ; ++jj) {
      const int j = jj * 32 + lane;
      float a = -INFINITY;
      if (j < p.Lk) {
        a = 0.f;
        const float* kr = Ks + j * ldk;
        for (int d = 0; d < dk; ++d) a = fmaf(qb[warp * dk + d], kr[d], a);
      }
      sc[jj] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXJ; ++jj) {
      const int j = jj * 32 + lane;
      sc[jj] = j < p.Lk ? __expf(sc[jj] - mx) : 0.f;
      sum += sc[jj];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    const long long row_id = ((long long)s * p.H + h) * p.Lq + i;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXJ; ++jj) {
      const int j = jj * 32 + lane;
      if (j < p.Lk) {
        float pr = sc[jj] * inv;
        if (p.probs != nullptr) p.probs[row_id * p.Lk + j] = pr;
        if (p.drop_thresh != 0u)
          pr = drop_keep_att(seed_eff, (unsigned long long)(row_id * p.Lk), j, p.drop_thresh) ? pr * p.drop_scale : 0.f;
        pb[warp * p.Lk + j] = pr;
      }
    }
    if (lane == 0 && p.lse != nullptr) p.lse[row_id] = mx + __logf(sum);
    __syncwarp();
    T* orow = reinterpret_cast<T*>(p.o) + s * p.o_bs + (long long)i * p.o_rs + h * dk;
    for (int d = lane; d < dk; d += 32) {
      float a = 0.f;
      for (int j = 0; j < p.Lk; ++j) a = fmaf(pb[warp * p.Lk + j], Vs[j * dk + d], a);
      orow[d] = from_f<T>(a);
    }
    __syncwarp();
  }
}

// Backward A: one warp per query row -> dQ row and delta_i = dO_i . O_i
template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attention_bwd_dq_kernel(const AttParams p) {
  const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);   // dropout seed of THIS replay
  extern __shared__ float sm[];
  const int dk = p.dk_dim, ldk = dk + 1;
  float* Ks = sm;                        // [Lk][dk+1]
  float* Vs = Ks + p.Lk * ldk;           // [Lk][dk+1]
  float* qb = Vs + p.Lk * ldk;           // [warps][dk]   scaled q
  float* gb = qb + ATT_WARPS * dk;       // [warps][dk]   dO row
  float* pb = gb + ATT_WARPS * dk;       // [warps][Lk]   dS row
  const int s = blockIdx.z, h = blockIdx.y;
  const int skv = (s + p.kv_shift) % p.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_rows_to_smem<T>(Ks, ldk, reinterpret_cast<const T*>(p.k) + skv * p.k_bs + h * dk, p.k_rs, p.Lk, dk);
  load_rows_to_smem<T>(Vs, ldk, reinterpret_cast<const T*>(p.v) + skv * p.v_bs + h * dk, p.v_rs, p.Lk, dk);
  __syncthreads();
  const int i_end = min(p.Lq, (blockIdx.x + 1) * ATT_ROWS);
  for (int i = blockIdx.x * ATT_ROWS + warp; i < i_end; i += ATT_WARPS) {
    const T* qrow = reinterpret_cast<const T*>(p.q) + s * p.q_bs + (long long)i * p.q_rs + h * dk;
    const T* grow = reinterpret_cast<const T*>(p.d_o) + s * p.do_bs + (long long)i * p.do_rs + h * dk;
    const T* orow = reinterpret_cast<const T*>(p.o) + s * p.o_bs + (long long)i * p.o_rs + h * dk;
    float dl = 0.f;
    for (int d = lane; d < dk; d += 32) {
      const float g = to_f(grow[d]);
      qb[warp * dk + d] = to_f(qrow[d]) * p.scale;
      gb[warp * dk + d] = g;
      dl += g * to_f(orow[d]);
    }
    dl = warp_sum(dl);
    __syncwarp();
    const long long row_id = ((long long)s * p.H + h) * p.Lq + i;
    const float lse = p.lse[row_id];
    if (lane == 0) p.delta[row_id] = dl;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXJ; ++jj) {
      const int j = jj * 32 + lane;
      if (j < p.Lk) {
        const float* kr = Ks + j * ldk;
        const float* vr = Vs + j * ldk;
        float a = 0.f, dp = 0.f;
        for (int d = 0; d < dk; ++d) {
          a = fmaf(qb[warp * dk + d], kr[d], a);
          dp = fmaf(gb[warp * dk + d], vr[d], dp);
        }
        const float pr = __expf(a - lse);
        if (p.drop_thresh != 0u)
          dp = drop_keep_att(seed_eff, (unsigned long long)(row_id * p.Lk), j, p.drop_thresh) ? dp * p.drop_scale : 0.f;
        pb[warp * p.Lk + j] = pr * (dp - dl);
      }
    }
    __syncwarp();
    T* dqrow = reinterpret_cast<T*>(p.dq) + s * p.dq_bs + (long long)i * p.dq_rs + h * dk;
    for (int d = lane; d < dk; d += 32) {
      float a = 0.f;
      for (int j = 0; j < p.Lk; ++j) a = fmaf(pb[warp * p.Lk + j], Ks[j * ldk + d], a);
      dqrow[d] = from_f<T>(a * p.scale);
    }
    __syncwarp();
  }
}

// Backward B: one warp per key row -> dK row and dV row (needs delta from kernel A)
template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attention_bwd_dkv_kernel(const AttParams p) {
  const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);   // dropout seed of THIS replay
  extern __shared__ float sm[];
  const int dk = p.dk_dim, ldk = dk + 1;
  float* Qs = sm;                        // [Lq][dk+1]  (unscaled)
  float* Gs = Qs + p.Lq * ldk;           // [Lq][dk+1]  dO
  float* kb = Gs + p.Lq * ldk;           // [warps][dk]
  float* vb = kb + ATT_WARPS * dk;       // [warps][dk]
  float* pb = vb + ATT_WARPS * dk;       // [warps][Lq]  p~ (dropped probs)
  float* sb = pb + ATT_WARPS * p.Lq;     // [warps][Lq]  dS
  const int s = blockIdx.z, h = blockIdx.y;
  const int skv = (s + p.kv_shift) % p.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_rows_to_smem<T>(Qs, ldk, reinterpret_cast<const T*>(p.q) + s * p.q_bs + h * dk, p.q_rs, p.Lq, dk);
  load_rows_to_smem<T>(Gs, ldk, reinterpret_cast<const T*>(p.d_o) + s * p.do_bs + h * dk, p.do_rs, p.Lq, dk);
  __syncthreads();
  const long long row_base = ((long long)s * p.H + h) * p.Lq;
  const int j_end = min(p.Lk, (blockIdx.x + 1) * ATT_ROWS);
  for (int j = blockIdx.x * ATT_ROWS + warp; j < j_end; j += ATT_WARPS) {
    const T* krow = reinterpret_cast<const T*>(p.k) + skv * p.k_bs + (long long)j * p.k_rs + h * dk;
    const T* vrow = reinterpret_cast<const T*>(p.v) + skv * p.v_bs + (long long)j * p.v_rs + h * dk;
    for (int d = lane; d < dk; d += 32) {
      kb[warp * dk + d] = to_f(krow[d]);
      vb[warp * dk + d] = to_f(vrow[d]);
    }
    __syncwarp();
#pragma unroll
    for (int ii = 0; ii < ATT_MAXJ; ++ii) {
      const int i = ii * 32 + lane;
      if (i < p.Lq) {
        const float* qr = Qs + i * ldk;
        const float* gr = Gs + i * ldk;
        float a = 0.f, dp = 0.f;
        for (int d = 0; d < dk; ++d) {
          a = fmaf(qr[d], kb[warp * dk + d], a);
          dp = fmaf(gr[d], vb[warp * dk + d], dp);
        }
        const float pr = __expf(a * p.scale - p.lse[row_base + i]);
        float pt = pr;
        if (p.drop_thresh != 0u) {
          const bool keep = drop_keep_att(seed_eff, (unsigned long long)((row_base + i) * p.Lk), j, p.drop_thresh);
          pt = keep ? pr * p.drop_scale : 0.f;
          dp = keep ? dp * p.drop_scale : 0.f;
        }
        pb[warp * p.Lq + i] = pt;
        sb[warp * p.Lq + i] = pr * (dp - p.delta[row_base + i]);
      }
    }
    __syncwarp();
    T* dkrow = reinterpret_cast<T*>(p.dk) + skv * p.dk_bs + (long long)j * p.dk_rs + h * dk;
    T* dvrow = reinterpret_cast<T*>(p.dv) + skv * p.dv_bs + (long long)j * p.dv_rs + h * dk;
    for (int d = lane; d < dk; d += 32) {
      float ak = 0.f, av = 0.f;
      for (int i = 0; i < p.Lq; ++i) {
        ak = fmaf(sb[warp * p.Lq + i], Qs[i * ldk + d], ak);
        av = fmaf(pb[warp * p.Lq + i], Gs[i * ldk + d], av);
      }
      dkrow[d] = from_f<T>(ak * p.scale);
      dvrow[d] = from_f<T>(av);
    }
    __syncwarp();
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  EGB_CHECK(bytes <= 227 * 1024, "attention: needs %zu bytes of shared memory (> 227 KB)", bytes);
  EGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace

extern "C" {

/* Layout: q/k/v/o and their gradients are addressed as base[b*bs + row*rs + head*dk + d] (elements). */


static int fill_att(const egb_attention_desc* d, AttParams* p) {
  EGB_CHECK(d->head_dim % 4 == 0 && d->head_dim <= 128, "attention: head_dim %d unsupported", d->head_dim);
  EGB_CHECK(d->Lq >= 1 && d->Lk >= 1 && d->Lq <= 32 * ATT_MAXJ && d->Lk <= 32 * ATT_MAXJ,
            "attention: sequence lengths %d/%d out of range (<= %d)", d->Lq, d->Lk, 32 * ATT_MAXJ);
  memset(p, 0, sizeof(*p));
  p->q = (const char*)d->q; p->k = (const char*)d->k; p->v = (const char*)d->v; p->o = (char*)d->o;
  p->d_o = (const char*)d->d_o; p->dq = (char*)d->dq; p->dk = (char*)d->dk; p->dv = (char*)d->dv;
  p->lse = d->lse; p->delta = d->delta; p->probs = d->probs;
  p->q_bs = d->q_bs; p->q_rs = d->q_rs; p->k_bs = d->k_bs; p->k_rs = d->k_rs; p->v_bs = d->v_bs; p->v_rs = d->v_rs;
  p->o_bs = d->o_bs; p->o_rs = d->o_rs; p->do_bs = d->do_bs; p->do_rs = d->do_rs;
  p->dq_bs = d->dq_bs; p->dq_rs = d->dq_rs; p->dk_bs = d->dk_bs; p->dk_rs = d->dk_rs; p->dv_bs = d->dv_bs; p->dv_rs = d->dv_rs;
  p->S = d->S; p->H = d->H; p->Lq = d->Lq; p->Lk = d->Lk; p->dk_dim = d->head_dim; p->kv_shift = d->kv_shift;
  p->scale = d->scale;
  if (d->dropout_p > 0.f) {
    p->drop_thresh = drop_threshold(d->dropout_p);
    p->drop_scale = 1.f / (1.f - d->dropout_p);
    p->seed = d->seed;
    p->epoch = egb_seed_epoch_ptr();
  }
  return 0;
}

// dq_colsum / dk_colsum / dv_colsum of the descriptor for the kernels that do not take them in-kernel: one column-sum
// pass per gradient tensor
int egb_attention_colsum_pass(const egb_attention_desc* d, cudaStream_t st) {
  if (d->dq_colsum == nullptr) return 0;
  EGB_CHECK(d->dk_colsum != nullptr && d->dv_colsum != nullptr, "attention_bwd: pass all three column-sum buffers or none");
  const int N = d->H * d->head_dim;
  egb_matrix m;
  m.dtype = d->dtype;
  m.ptr = d->dq; m.rows_per_group = d->Lq; m.row_stride = d->dq_rs; m.group_stride = d->dq_bs;
  if (egb_colsum(&m, d->S * d->Lq, N, d->dq_colsum, 0, st)) return 1;
  m.ptr = d->dk; m.rows_per_group = d->Lk; m.row_stride = d->dk_rs; m.group_stride = d->dk_bs;
  if (egb_colsum(&m, d->S * d->Lk, N, d->dk_colsum, 0, st)) return 1;
  m.ptr = d->dv; m.rows_per_group = d->Lk; m.row_stride = d->dv_rs; m.group_stride = d->dv_bs;
  return egb_colsum(&m, d->S * d->Lk, N, d->dv_colsum, 0, st);
}

int egb_attention_fwd(const egb_attention_desc* d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (egb_attention_tc_supported(d, false)) return egb_attention_tc_fwd(d, st);
  AttParams p;
  if (fill_att(d, &p)) return 1;
  const int dk = d->head_dim;
  const size_t smem = sizeof(float) * ((size_t)d->Lk * (dk + 1) + (size_t)d->Lk * dk + ATT_WARPS * dk + ATT_WARPS * d->Lk);
  dim3 grid((d->Lq + ATT_ROWS - 1) / ATT_ROWS, d->H, d->S);
  if (d->dtype == EGB_BF16) {
    if (set_smem(attention_fwd_kernel<bf16>, smem)) return 1;
    attention_fwd_kernel<bf16><<<grid, ATT_THREADS, smem, st>>>(p);
  } else {
    if (set_smem(attention_fwd_kernel<float>, smem)) return 1;
    attention_fwd_kernel<float><<<grid, ATT_THREADS, smem, st>>>(p);
  }
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_attention_bwd(const egb_attention_desc* d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (egb_attention_tc_supported(d, true)) return egb_attention_tc_bwd(d, st);
  AttParams p;
  if (fill_att(d, &p)) return 1;
  EGB_CHECK(d->lse && d->delta && d->d_o && d->dq && d->dk && d->dv, "attention_bwd: missing buffers");
  const int dk = d->head_dim;
  const size_t smem_a = sizeof(float) * ((size_t)2 * d->Lk * (dk + 1) + 2 * ATT_WARPS * dk + ATT_WARPS * d->Lk);
  const size_t smem_b = sizeof(float) * ((size_t)2 * d->Lq * (dk + 1) + 2 * ATT_WARPS * dk + 2 * ATT_WARPS * d->Lq);
  dim3 grid_a((d->Lq + ATT_ROWS - 1) / ATT_ROWS, d->H, d->S);
  dim3 grid_b((d->Lk + ATT_ROWS - 1) / ATT_ROWS, d->H, d->S);
  if (d->dtype == EGB_BF16) {
    if (set_smem(attention_bwd_dq_kernel<bf16>, smem_a) || set_smem(attention_bwd_dkv_kernel<bf16>, smem_b)) return 1;
    attention_bwd_dq_kernel<bf16><<<grid_a, ATT_THREADS, smem_a, st>>>(p);
    attention_bwd_dkv_kernel<bf16><<<grid_b, ATT_THREADS, smem_b, st>>>(p);
  } else {
    if (set_smem(attention_bwd_dq_kernel<float>, smem_a) || set_smem(attention_bwd_dkv_kernel<float>, smem_b)) return 1;
    attention_bwd_dq_kernel<float><<<grid_a, ATT_THREADS, smem_a, st>>>(p);
    attention_bwd_dkv_kernel<float><<<grid_b, ATT_THREADS, smem_b, st>>>(p);
  }
  egb_count_launch(2);
  EGB_LAUNCH_CHECK();
  return egb_attention_colsum_pass(d, st);
}

}  // extern "C"
