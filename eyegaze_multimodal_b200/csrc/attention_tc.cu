// Fused multi-head attention on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands, fp32 softmax.
//   forward : S = Q K^T (TMEM) -> row softmax in registers -> P (bf16) written back to TMEM ->
//             O = P V with the A operand read from TMEM (no shared-memory round trip for P)
//   backward: att_tc_bwd_pipe_kernel -- ONE kernel per (batch, head): 8 softmax warps + 2 MMA-issuing warps, 64-key
//             rounds with the score pair (S, dP = dO V^T) double-buffered in TMEM, P~ / dS staged in shared memory as
//             K-major tiles and re-read as MN-major A operands for dV = P~^T dO, dK = dS^T Q; dQ = dS K from the same
//             tiles.  (EGB_ATT_FUSED_BWD=1 / 0 select the two earlier designs kept below: the single-issuer fused kernel
//             and the dQ + dKV kernel pair with TMEM-resident P / dS.)
// Sequence lengths on this path are <= 256 (L = 33 / 139 / 235 for the EEG encoder, 197 for the ViT), so the
// whole key axis is ONE tcgen05 N tile (N <= 256) and no online-softmax rescaling across tiles is needed.
// One thread owns one TMEM lane = one query (or key) row, so row statistics need no shuffles at all.
//
// Operand tiles land directly in the 128-byte-swizzled UMMA layout -- by TMA (one bulk tensor copy per tile of the
// strided per-head slice of the packed [S, L, 3D] projection, head_dim 64) or by cp.async (head_dim 32); the same
// physical tile serves as a K-major operand (Q K^T, dO V^T) and as an MN-major operand (P V, dS K, dS^T Q, P^T dO).
// Results leave as whole row segments through a warp-private shared-memory transposition.
// art.py:203-213 / timm Attention; cross-brain attention (dual_eeg_transformer.py:966-974) via kv_shift.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/eyegaze_b200.h"
#include <cuda.h>
#include <stdlib.h>

extern void egb_count_launch(int n);
int egb_tmap_rows64(CUtensorMap* out, const void* ptr, long long inner, long long rows, long long groups, long long rs,
                    long long gs, int box_rows);
int egb_prof_enabled();
extern "C" int egb_attention_colsum_pass(const egb_attention_desc* d, cudaStream_t st);
void egb_prof_begin(cudaStream_t st, double flops, double bytes, int kind);
void egb_prof_end(cudaStream_t st);

namespace {

constexpr int TC_THREADS = 256;   // 8 warps: warp w owns TMEM lanes 32*(w%4).., column chunks of parity w/4
constexpr int TILE_ROWS = 128;
constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t ACC_COL = 128;  // accumulator columns of the second-stage MMAs (above the packed bf16 A operand)
constexpr uint32_t HALF_COL = 256; // second score matrix (dP) of the backward kernels

struct AttTcParams {
  const bf16 *q, *k, *v, *o, *d_o;
  bf16 *out, *dq, *dk, *dv;
  float* lse;    // [S,H,Lq]
  float* delta;  // [S,H,Lq]
  float* probs;  // optional [S,H,Lq,Lk]
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;
  long long dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs, do_bs, do_rs;
  int S, H, Lq, Lk, d, kv_shift, Lq_pad, Lk_pad;
  float scale;
  unsigned drop_thresh;
  float drop_scale;
  unsigned long long seed;
  const unsigned long long* epoch;   // device seed epoch (egb_mix_seed), NULL when not enabled
  long long* dbg;  // optional: phase timestamps (clock64) of CTA (0, 0, S/2), see egb_debug_attention_timing
  int use_tma;     // pipelined backward: operand tiles arrive by TMA (head_dim 64) instead of cp.async
  float *dq_cs, *dk_cs, *dv_cs;   // optional [H*d] fp32, accumulated: column sums of the stored dQ / dK / dV (bias gradients)
};
// tensor maps of the pipelined backward's operands ({H d, L, S} views, box {64, L_pad, 1}, 128-byte swizzle)
struct AttMaps {
  CUtensorMap q, k, v, g, o;
};

#define ATT_STAMP(slot)                                                                       \
  do {                                                                                        \
    if (p.dbg != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 &&         \
        blockIdx.z == gridDim.z / 2)                                                          \
      p.dbg[slot] = clock64();                                                                \
  } while (0)

// same, for a CTA in the middle of a (heads, sequences) grid: the first wave starts all its loads at once and is not
// representative
#define ATT_STAMP_MID(slot)                                                                   \
  do {                                                                                        \
    if (p.dbg != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2) \
      p.dbg[slot] = clock64();                                                                \
  } while (0)

#ifdef EGB_ATT_TIMING
#define ATT_CLK(v) const long long v = clock64()
#else
#define ATT_CLK(v) const long long v = 0
#endif

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// rows [0, rows_total) of a [rows][d] bf16 slice (row stride rs elements) -> 128-B pitched, 128-B swizzled tile.
// Rows >= rows_valid are zero-filled.  16-byte chunk c of row r lands at chunk (c ^ (r & 7)) of its 128-B line.
// The copies are asynchronous (cp.async / LDGSTS, zero-fill for the padding rows): a thread issues all of its
// chunks back to back; cp_async_wait_all() + a proxy fence + __syncthreads() publish the tiles to the tensor core.
__device__ __forceinline__ void load_rows_sw128(uint8_t* dst, const bf16* src, long long rs, int rows_valid,
                                                int rows_total, int d, int nthreads = 256) {
  const int sh = d == 64 ? 3 : 2;                 // 16-byte chunks per row = d / 8 (head_dim is 32 or 64)
  const int cmask = (1 << sh) - 1;
  const uint32_t dst0 = ptx::smem_u32(dst);
  const int total = rows_total << sh;
  for (int idx = threadIdx.x; idx < total; idx += nthreads) {
    const int r = idx >> sh, c = idx & cmask;
    const bool in = r < rows_valid;
    const bf16* g = in ? src + (long long)r * rs + c * 8 : src;
    const uint32_t sa = dst0 + (uint32_t)((r << 7) + ((c ^ (r & 7)) << 4));   // (r>>3)*1024 + (r&7)*128 == r*128
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "r"(in ? 16u : 0u) : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// D[tmem] (+)= A[tmem, bf16 packed 2 per column] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[NC]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&v)[32]) { ptx::tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 2^x for x <= 0 (one MUFU; results below the normal range flush to zero, which is what a softmax wants)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// acc[tmem_d] = A_tile[128 x d] . B_tile[n x d]^T : both operands K-major in shared memory (K step = 32 B)
__device__ __forceinline__ void mma_ss_kk(uint32_t tmem_d, const uint8_t* a, const uint8_t* b, int n, int d) {
  const uint32_t idesc = ptx::make_idesc_bf16(TILE_ROWS, n, 0, 0);
  const uint64_t ad = ptx::make_smem_desc(ptx::smem_u32(a), 16u, 1024u);
  const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(b), 16u, 1024u);
  for (int k = 0; k < d / 16; ++k)
    ptx::umma_bf16(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
}
// acc[tmem_d] = A[tmem: 128 x kdim, packed bf16] . B_tile[kdim rows][d] : B MN-major in shared memory
// (K step = 16 rows = 2048 B; the packed A operand advances 8 columns per step)
__device__ __forceinline__ void mma_ts_mn(uint32_t tmem_d, uint32_t tmem_a, const uint8_t* b, int kdim, int d,
                                          bool accumulate = false) {
  const uint32_t idesc = ptx::make_idesc_bf16(TILE_ROWS, d, 0, 1);
  const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(b), (uint32_t)kdim * 128u, 1024u);
  for (int ks = 0; ks < kdim / 16; ++ks)
    umma_bf16_ts(tmem_d, tmem_a + (uint32_t)(8 * ks), bd + (uint64_t)(128 * ks), idesc, (accumulate || ks > 0) ? 1u : 0u);
}

// accumulator columns [c_begin, c_end) (multiples of 32) of this thread's row -> scaled bf16 -> global
__device__ __forceinline__ void store_acc_row(uint32_t taddr, int c_begin, int c_end, float mul, bf16* row) {
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    uint32_t raw[32];
    ptx::tmem_ld32(taddr + (uint32_t)c0, raw);
    ptx::tmem_ld_wait();
    if (row != nullptr) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[g * 8 + i]) * mul;
        st8(row + c0 + g * 8, v);
      }
    }
  }
}

// This warp's 32 accumulator rows x ncols (32 or 64) fp32 TMEM columns -> * mul -> bf16 -> global memory, written as
// whole 64 / 128-byte row segments: the rows are staged in a warp-private 4 KB shared-memory area (16-byte slots
// XOR-swizzled by row) and each store instruction then covers 32 / (ncols / 8) complete rows.  One thread owns one
// TMEM lane = one row, so a direct store scatters 32 separate 16-byte pieces per instruction (measured: ~6 K cycles
// of LSU time per backward CTA).  g0 = first of the warp's rows (may be past the end when rows_valid <= 0).
// 8 packed bf16 values += into acc (column sums of the stored rows)
__device__ __forceinline__ void add_bf16x8(const uint4& v, float (&acc)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[2 * i] += __low2float(h[i]); acc[2 * i + 1] += __high2float(h[i]); }
}

// `cs` (optional, shared memory, PRIVATE to the calling warp): column sums of the rows this call stores are added to
// cs[0 .. ncols) with plain read-modify-writes (fp32 shared-memory atomics are compare-and-swap spin loops: with eight
// warps on the same 64 words they cost more than the column-sum pass they replace).
__device__ __forceinline__ void store_acc_rows_coalesced(uint32_t taddr, int ncols, float mul, uint8_t* stage, bf16* g0,
                                                         long long rs, int rows_valid, int lane, float* cs = nullptr) {
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t raw[32];
    ptx::tmem_ld32(taddr + (uint32_t)c0, raw);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 v;
      v.x = pack_bf16(__uint_as_float(raw[8 * g + 0]) * mul, __uint_as_float(raw[8 * g + 1]) * mul);
      v.y = pack_bf16(__uint_as_float(raw[8 * g + 2]) * mul, __uint_as_float(raw[8 * g + 3]) * mul);
      v.z = pack_bf16(__uint_as_float(raw[8 * g + 4]) * mul, __uint_as_float(raw[8 * g + 5]) * mul);
      v.w = pack_bf16(__uint_as_float(raw[8 * g + 6]) * mul, __uint_as_float(raw[8 * g + 7]) * mul);
      const int slot = (c0 >> 3) + g;
      *reinterpret_cast<uint4*>(stage + lane * 128 + ((slot ^ (lane & 7)) << 4)) = v;
    }
  }
  __syncwarp();
  const int sh = ncols == 64 ? 3 : 2;             // log2(16-byte pieces per row)
  const int ch = lane & ((1 << sh) - 1), r0 = lane >> sh, rpi = 32 >> sh;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int it = 0; it < (1 << sh); ++it) {
    const int rr = it * rpi + r0;
    const uint4 v = *reinterpret_cast<const uint4*>(stage + rr * 128 + ((ch ^ (rr & 7)) << 4));
    if (rr < rows_valid) {
      *reinterpret_cast<uint4*>(g0 + (long long)rr * rs + ch * 8) = v;
      if (cs != nullptr) add_bf16x8(v, acc);
    }
  }
  if (cs != nullptr) {                            // lanes that share `ch` hold different rows of the same 8 columns
#pragma unroll
    for (int j = 0; j < 8; ++j)
      for (int o = 16; o >= (1 << sh); o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    if (r0 == 0)
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[ch * 8 + j] += acc[j];
  }
  __syncwarp();
}

// Same for exactly 32 columns with a 2 KB staging area per warp (rows pitched 64 B; the slot swizzle keeps both the
// row-wise writes and the 8-rows-per-instruction read-out free of bank conflicts).
__device__ __forceinline__ void store_acc_rows32_coalesced(uint32_t taddr, float mul, uint8_t* stage, bf16* g0, long long rs,
                                                           int rows_valid, int lane, float* cs = nullptr) {
  uint32_t raw[32];
  ptx::tmem_ld32(taddr, raw);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 v;
    v.x = pack_bf16(__uint_as_float(raw[8 * g + 0]) * mul, __uint_as_float(raw[8 * g + 1]) * mul);
    v.y = pack_bf16(__uint_as_float(raw[8 * g + 2]) * mul, __uint_as_float(raw[8 * g + 3]) * mul);
    v.z = pack_bf16(__uint_as_float(raw[8 * g + 4]) * mul, __uint_as_float(raw[8 * g + 5]) * mul);
    v.w = pack_bf16(__uint_as_float(raw[8 * g + 6]) * mul, __uint_as_float(raw[8 * g + 7]) * mul);
    *reinterpret_cast<uint4*>(stage + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = v;
  }
  __syncwarp();
  const int ch = lane & 3, r0 = lane >> 2;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + r0;
    const uint4 v = *reinterpret_cast<const uint4*>(stage + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
    if (rr < rows_valid) {
      *reinterpret_cast<uint4*>(g0 + (long long)rr * rs + ch * 8) = v;
      if (cs != nullptr) add_bf16x8(v, acc);
    }
  }
  if (cs != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
    }
    if (r0 == 0)
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[ch * 8 + j] += acc[j];
  }
  __syncwarp();
}

// barrier init (thread 0) -- TMEM is allocated AFTER the tile copies have been issued, so a CTA that has to wait
// for the previous CTA's tensor memory stages its operands meanwhile
__device__ __forceinline__ void init_bars(uint64_t* bars) {
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
}
template <int COLS>
__device__ __forceinline__ void cta_epilogue(uint32_t tmem) {
  ptx::tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<COLS>(tmem);
  }
}

// 32 key columns of one query row of the forward pass: e = exp2(S sl2 - mb) (row sum returned), P~ = dropout(e) packed
// to bf16 pairs.  FAST = all 32 columns are real keys and the dropout counter does not carry into its high word (see
// softmax_bwd_cols32 for the reasoning; both produce the mask of drop_keep_att()).
template <bool DROP, bool FAST>
__device__ __forceinline__ float softmax_fwd_cols32(const uint32_t (&raw)[32], float sl2, float mb, int col0, int Lk,
                                                    unsigned long long seed, unsigned long long row_lin, uint32_t thresh,
                                                    float drop_scale, uint32_t (&pk)[16]) {
  const uint32_t s0 = DROP ? drop_att_run_state(seed, row_lin, col0) : 0u;   // col0 is a multiple of 32: one run
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float e0 = fast_exp2(fmaf(__uint_as_float(raw[2 * j]), sl2, -mb));
    float e1 = fast_exp2(fmaf(__uint_as_float(raw[2 * j + 1]), sl2, -mb));
    if (!FAST) {
      if (col0 + 2 * j >= Lk) e0 = 0.f;
      if (col0 + 2 * j + 1 >= Lk) e1 = 0.f;
    }
    sum0 += e0;
    sum1 += e1;
    if (DROP) {
      const DropLeap l0 = drop_leap(2 * j), l1 = drop_leap(2 * j + 1);
      e0 *= (s0 * l0.mul + l0.add) >= thresh ? drop_scale : 0.f;
      e1 *= (s0 * l1.mul + l1.add) >= thresh ? drop_scale : 0.f;
    }
    pk[j] = pack_bf16(e0, e1);
  }
  return sum0 + sum1;
}

// ======================================================================================= forward
template <bool DROP>
__global__ void __launch_bounds__(TC_THREADS, 3) att_tc_fwd_kernel(const __grid_constant__ AttTcParams p, const __grid_constant__ AttMaps maps) {
  const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);   // dropout seed of THIS replay
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sQ = align1024(smem_raw);
  uint8_t* sK = sQ + TILE_ROWS * 128;
  uint8_t* sV = sK + p.Lk_pad * 128;
  float* s_red = reinterpret_cast<float*>(sV + p.Lk_pad * 128);  // [2][128] cross-half row statistics
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 2 * TILE_ROWS);
  uint64_t* bar_ld = &bars[2];                   // TMA tile loads (head_dim 64)
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int q0 = blockIdx.x * TILE_ROWS, h = blockIdx.y, s = blockIdx.z;
  const int skv = (s + p.kv_shift) % p.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp >> 2, row = (warp & 3) * 32 + lane;
  const int d = p.d;

  ATT_STAMP(0);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::mbar_init(bar_ld, 1);
    ptx::fence_barrier_init();
    if (p.use_tma) {   // one bulk tensor copy per tile; rows past the sequence end are zero-filled by the TMA unit
      ptx::mbar_arrive_expect_tx(bar_ld, (uint32_t)((TILE_ROWS + 2 * p.Lk_pad) * 128));
      ptx::tma_load_3d(sQ, &maps.q, bar_ld, h * d, q0, s);
      ptx::tma_load_3d(sK, &maps.k, bar_ld, h * d, 0, skv);
      ptx::tma_load_3d(sV, &maps.v, bar_ld, h * d, 0, skv);
    }
  }
  if (!p.use_tma) {
    load_rows_sw128(sQ, p.q + s * p.q_bs + (long long)q0 * p.q_rs + h * d, p.q_rs, min(TILE_ROWS, p.Lq - q0), TILE_ROWS, d);
    load_rows_sw128(sK, p.k + skv * p.k_bs + h * d, p.k_rs, p.Lk, p.Lk_pad, d);
    load_rows_sw128(sV, p.v + skv * p.v_bs + h * d, p.v_rs, p.Lk, p.Lk_pad, d);
  }
  if (warp == 0) ptx::tmem_alloc<256>(slot);
  ATT_STAMP(1);
  cp_async_wait_all();
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (p.use_tma) ptx::mbar_wait(bar_ld, 0);
  ATT_STAMP(2);

  if (threadIdx.x == 0) {
    mma_ss_kk(tmem, sQ, sK, p.Lk_pad, d);
    ptx::umma_commit(&bars[0]);
  }
  const int i = q0 + row;
  const bool valid = i < p.Lq;
  const long long row_id = ((long long)s * p.H + h) * p.Lq + (valid ? i : 0);
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int nchunk = (p.Lk_pad + 31) / 32;
  ptx::mbar_wait(&bars[0], 0);
  ptx::tc_fence_after();
  ATT_STAMP(3);

  // Warps whose 32 query rows all lie past the sequence end (EEG: L = 139 leaves 11 rows in the second tile) skip the
  // softmax arithmetic; their P rows stay undefined, which only reaches O rows that are never stored.
  const bool warp_live = q0 + (warp & 3) * 32 < p.Lq;
  // pass 1: row maximum of the raw scores (this half's column chunks), combined across the two halves
  float mx = -INFINITY;
  for (int c = half; c < nchunk && warp_live; c += 2) {
    uint32_t raw[32];
    ptx::tmem_ld32(trow + (uint32_t)(c * 32), raw);
    ptx::tmem_ld_wait();
    if (c * 32 + 32 <= p.Lk) {   // all 32 columns are real keys: no per-column predicate (16 M of the ViT layer's 76 M
                                 // warp instructions were the ISETP / VIADD pairs of this loop), two independent chains
      float m0 = mx, m1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        m0 = fmaxf(m0, __uint_as_float(raw[j]));
        m1 = fmaxf(m1, __uint_as_float(raw[j + 1]));
      }
      mx = fmaxf(m0, m1);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < p.Lk) mx = fmaxf(mx, __uint_as_float(raw[j]));
    }
  }
  s_red[half * TILE_ROWS + row] = mx;
  __syncthreads();
  mx = fmaxf(s_red[row], s_red[TILE_ROWS + row]);
  const float sl2 = p.scale * LOG2E;
  const float mb = mx * sl2;
  if (p.probs != nullptr) {  // analysis hooks: export the normalised probabilities (extra passes, rare path)
    float sum0 = 0.f;
    for (int c = half; c < nchunk; c += 2) {
      uint32_t raw[32];
      ptx::tmem_ld32(trow + (uint32_t)(c * 32), raw);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < p.Lk) sum0 += fast_exp2(__uint_as_float(raw[j]) * sl2 - mb);
    }
    __syncthreads();                      // everyone has consumed the maxima
    s_red[half * TILE_ROWS + row] = sum0;
    __syncthreads();
    const float inv0 = 1.f / (s_red[row] + s_red[TILE_ROWS + row]);
    for (int c = half; c < nchunk; c += 2) {
      uint32_t raw[32];
      ptx::tmem_ld32(trow + (uint32_t)(c * 32), raw);
      ptx::tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < p.Lk) p.probs[row_id * p.Lk + c * 32 + j] = fast_exp2(__uint_as_float(raw[j]) * sl2 - mb) * inv0;
      }
    }
  }
  __syncthreads();                        // s_red is rewritten below
  // pass 2: exponentials, row sum, (dropout), bf16 probabilities written back to TMEM in place.  Round t handles
  // chunks 2t (half 0) and 2t+1 (half 1): their packed forms go to columns [32t, 32t+32) = chunk t, which both
  // halves have consumed once the round's barrier is passed (t = 0: this round; t >= 1: an earlier round).
  float sum = 0.f;
  const int nround = (nchunk + 1) / 2;
  for (int t = 0; t < nround; ++t) {
    const int c = 2 * t + half;
    uint32_t pk[16];
    const bool mine = c < nchunk && warp_live;
    if (mine) {
      uint32_t raw[32];
      ptx::tmem_ld32(trow + (uint32_t)(c * 32), raw);
      ptx::tmem_ld_wait();
      const unsigned long long row_lin = (unsigned long long)(row_id * p.Lk);
      const bool fast = c * 32 + 32 <= p.Lk;          // all 32 columns are real keys: no per-column predicates
      if (fast) sum += softmax_fwd_cols32<DROP, true>(raw, sl2, mb, c * 32, p.Lk, seed_eff, row_lin, p.drop_thresh, p.drop_scale, pk);
      else sum += softmax_fwd_cols32<DROP, false>(raw, sl2, mb, c * 32, p.Lk, seed_eff, row_lin, p.drop_thresh, p.drop_scale, pk);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (mine) tmem_st16(trow + (uint32_t)(c * 16), pk);
  }
  s_red[half * TILE_ROWS + row] = sum;
  ptx::tmem_st_wait();
  ATT_STAMP(4);
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::tc_fence_after();
    mma_ts_mn(tmem + ACC_COL, tmem, sV, p.Lk_pad, d);
    ptx::umma_commit(&bars[1]);
  }
  sum = s_red[row] + s_red[TILE_ROWS + row];
  if (half == 0 && valid && p.lse != nullptr) p.lse[row_id] = mx * p.scale + __logf(sum);
  ptx::mbar_wait(&bars[1], 0);
  ptx::tc_fence_after();
  ATT_STAMP(5);
  {
    // coalesced store through the Q tile's shared memory (dead since the score MMA retired)
    const int r0 = q0 + (warp & 3) * 32;
    const int cb = d >= 64 ? half * 32 : 0;
    if (half == 0 || d >= 64)
      store_acc_rows32_coalesced(trow + ACC_COL + (uint32_t)cb, 1.f / sum, sQ + warp * 2048,
                                 p.out + s * p.o_bs + (long long)r0 * p.o_rs + h * d + cb, p.o_rs, p.Lq - r0, lane);
  }
  ATT_STAMP(6);
  cta_epilogue<256>(tmem);
  ATT_STAMP(7);
}

// ======================================================================================= backward
// Both backward kernels walk the "other" sequence axis in chunks of 64 columns so that a CTA needs only 256 TMEM
// columns (S chunk | dP chunk | accumulators) and two CTAs share an SM: while one runs its softmax math the other's
// MMAs and tile staging proceed.  Per chunk: [S_c, dP_c MMAs] -> threads build bf16 dS_c (and P~_c) in place ->
// [accumulating MMAs with that chunk as the TMEM A operand] -> next chunk's score MMAs are issued right behind
// (tcgen05.mma executes in issue order, so they may overwrite the columns the accumulating MMAs just read).
constexpr uint32_t BW_DP = 64;    // dP chunk columns
constexpr uint32_t BW_ACC0 = 128; // first accumulator (dQ, or dV)
constexpr uint32_t BW_ACC1 = 192; // second accumulator (dK)

// -------------------------------------------------------------------------------------- dQ (+ delta)
template <bool DROP>
__global__ void __launch_bounds__(TC_THREADS) att_tc_bwd_dq_kernel(const AttTcParams p) {
  const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);   // dropout seed of THIS replay
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sQ = align1024(smem_raw);
  uint8_t* sG = sQ + TILE_ROWS * 128;  // dO tile
  uint8_t* sK = sG + TILE_ROWS * 128;
  uint8_t* sV = sK + p.Lk_pad * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + p.Lk_pad * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int q0 = blockIdx.x * TILE_ROWS, h = blockIdx.y, s = blockIdx.z;
  const int skv = (s + p.kv_shift) % p.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp >> 2, row = (warp & 3) * 32 + lane;
  const int d = p.d;
  const int rows = min(TILE_ROWS, p.Lq - q0);

  ATT_STAMP(0);
  init_bars(bars);
  load_rows_sw128(sQ, p.q + s * p.q_bs + (long long)q0 * p.q_rs + h * d, p.q_rs, rows, TILE_ROWS, d);
  load_rows_sw128(sG, p.d_o + s * p.do_bs + (long long)q0 * p.do_rs + h * d, p.do_rs, rows, TILE_ROWS, d);
  load_rows_sw128(sK, p.k + skv * p.k_bs + h * d, p.k_rs, p.Lk, p.Lk_pad, d);
  load_rows_sw128(sV, p.v + skv * p.v_bs + h * d, p.v_rs, p.Lk, p.Lk_pad, d);
  // delta_i = dO_i . O_i (both halves compute it; overlaps the copies and the TMEM wait)
  const int i = q0 + row;
  const bool valid = i < p.Lq;
  const long long row_id = ((long long)s * p.H + h) * p.Lq + (valid ? i : 0);
  float dl = 0.f, lse2 = 0.f;
  if (valid) {
    const bf16* orow = p.o + s * p.o_bs + (long long)i * p.o_rs + h * d;
    const bf16* grow = p.d_o + s * p.do_bs + (long long)i * p.do_rs + h * d;
    for (int c = 0; c < d; c += 8) {
      float a[8], b[8];
      ld8(orow + c, a);
      ld8(grow + c, b);
#pragma unroll
      for (int t = 0; t < 8; ++t) dl = fmaf(a[t], b[t], dl);
    }
    if (half == 0) p.delta[row_id] = dl;
    lse2 = p.lse[row_id] * LOG2E;
  }
  if (warp == 0) ptx::tmem_alloc<256>(slot);
  ATT_STAMP(1);
  cp_async_wait_all();
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  ATT_STAMP(2);

  const int nchunk = (p.Lk_pad + 63) / 64;
  if (threadIdx.x == 0) {
    const int w0 = min(64, p.Lk_pad);
    mma_ss_kk(tmem, sQ, sK, w0, d);           // S_0  = Q K_0^T
    mma_ss_kk(tmem + BW_DP, sG, sV, w0, d);   // dP_0 = dO V_0^T
    ptx::umma_commit(&bars[0]);
  }
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const float sl2 = p.scale * LOG2E;
  for (int c = 0; c < nchunk; ++c) {
    const int w = min(64, p.Lk_pad - 64 * c);          // chunk width (multiple of 16)
    const bool mine = 32 * half < w;
    ptx::mbar_wait(&bars[0], (uint32_t)(c & 1));
    ptx::tc_fence_after();
    if (c == 0) ATT_STAMP(3);
    uint32_t pk[16];
    if (mine) {
      uint32_t rs[32], rp[32];
      ptx::tmem_ld32(trow + (uint32_t)(32 * half), rs);
      ptx::tmem_ld32(trow + BW_DP + (uint32_t)(32 * half), rp);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float ds2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int col = 64 * c + 32 * half + 2 * j + u;
          const float pr = fast_exp2(__uint_as_float(rs[2 * j + u]) * sl2 - lse2);
          float dp = __uint_as_float(rp[2 * j + u]);
          if (DROP) dp = drop_keep_att(seed_eff, (unsigned long long)(row_id * p.Lk), col, p.drop_thresh) ? dp * p.drop_scale : 0.f;
          ds2[u] = col < p.Lk ? pr * (dp - dl) : 0.f;
        }
        pk[j] = pack_bf16(ds2[0], ds2[1]);
      }
    }
    ptx::tc_fence_before();
    __syncthreads();                                     // both halves have consumed the raw chunk
    ptx::tc_fence_after();
    if (mine) tmem_st16(trow + (uint32_t)(16 * half), pk);
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      ptx::tc_fence_after();
      mma_ts_mn(tmem + BW_ACC0, tmem, sK + (size_t)c * 64 * 128, w, d, c > 0);   // dQ += dS_c K_c
      if (c + 1 < nchunk) {
        const int w1 = min(64, p.Lk_pad - 64 * (c + 1));
        mma_ss_kk(tmem, sQ, sK + (size_t)(c + 1) * 64 * 128, w1, d);
        mma_ss_kk(tmem + BW_DP, sG, sV + (size_t)(c + 1) * 64 * 128, w1, d);
        ptx::umma_commit(&bars[0]);
      } else {
        ptx::umma_commit(&bars[1]);
      }
    }
  }
  ATT_STAMP(4);
  ptx::mbar_wait(&bars[1], 0);
  ptx::tc_fence_after();
  ATT_STAMP(5);
  {
    bf16* drow = valid ? p.dq + s * p.dq_bs + (long long)i * p.dq_rs + h * d : nullptr;
    if (d >= 64) store_acc_row(trow + BW_ACC0, half * (d / 2), half * (d / 2) + d / 2, p.scale, drow);
    else if (half == 0) store_acc_row(trow + BW_ACC0, 0, d, p.scale, drow);
  }
  ATT_STAMP(6);
  cta_epilogue<256>(tmem);
  ATT_STAMP(7);
}

// -------------------------------------------------------------------------------------- dK, dV
template <bool DROP>
__global__ void __launch_bounds__(TC_THREADS) att_tc_bwd_dkv_kernel(const AttTcParams p) {
  const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);   // dropout seed of THIS replay
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sK = align1024(smem_raw);   // 128 key rows of this tile
  uint8_t* sV = sK + TILE_ROWS * 128;
  uint8_t* sQ = sV + TILE_ROWS * 128;  // all query rows
  uint8_t* sG = sQ + p.Lq_pad * 128;   // dO, all query rows
  float* s_lse = reinterpret_cast<float*>(sG + p.Lq_pad * 128);
  float* s_del = s_lse + p.Lq_pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_del + p.Lq_pad);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int k0 = blockIdx.x * TILE_ROWS, h = blockIdx.y, s = blockIdx.z;
  const int skv = (s + p.kv_shift) % p.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp >> 2, row = (warp & 3) * 32 + lane;
  const int d = p.d;
  const int rows = min(TILE_ROWS, p.Lk - k0);
  const long long row_base = ((long long)s * p.H + h) * p.Lq;

  ATT_STAMP(0);
  init_bars(bars);
  load_rows_sw128(sK, p.k + skv * p.k_bs + (long long)k0 * p.k_rs + h * d, p.k_rs, rows, TILE_ROWS, d);
  load_rows_sw128(sV, p.v + skv * p.v_bs + (long long)k0 * p.v_rs + h * d, p.v_rs, rows, TILE_ROWS, d);
  load_rows_sw128(sQ, p.q + s * p.q_bs + h * d, p.q_rs, p.Lq, p.Lq_pad, d);
  load_rows_sw128(sG, p.d_o + s * p.do_bs + h * d, p.do_rs, p.Lq, p.Lq_pad, d);
  for (int t = threadIdx.x; t < p.Lq_pad; t += TC_THREADS) {
    s_lse[t] = t < p.Lq ? p.lse[row_base + t] * LOG2E : 0.f;
    s_del[t] = t < p.Lq ? p.delta[row_base + t] : 0.f;
  }
  if (warp == 0) ptx::tmem_alloc<256>(slot);
  ATT_STAMP(1);
  cp_async_wait_all();
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  ATT_STAMP(2);

  const int nchunk = (p.Lq_pad + 63) / 64;
  if (threadIdx.x == 0) {
    const int w0 = min(64, p.Lq_pad);
    mma_ss_kk(tmem, sK, sQ, w0, d);           // S^T_0  = K Q_0^T
    mma_ss_kk(tmem + BW_DP, sV, sG, w0, d);   // dP^T_0 = V dO_0^T
    ptx::umma_commit(&bars[0]);
  }
  const int j = k0 + row;
  const bool valid = j < p.Lk;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const float sl2 = p.scale * LOG2E;
  for (int c = 0; c < nchunk; ++c) {
    const int w = min(64, p.Lq_pad - 64 * c);
    const bool mine = 32 * half < w;
    ptx::mbar_wait(&bars[0], (uint32_t)(c & 1));
    ptx::tc_fence_after();
    if (c == 0) ATT_STAMP(3);
    uint32_t pkp[16], pks[16];
    if (mine) {
      uint32_t rs[32], rp[32];
      ptx::tmem_ld32(trow + (uint32_t)(32 * half), rs);
      ptx::tmem_ld32(trow + BW_DP + (uint32_t)(32 * half), rp);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        float pt2[2], ds2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int col = 64 * c + 32 * half + 2 * jj + u;   // query index
          const int ci = min(col, p.Lq_pad - 1);              // columns past Lq_pad are masked; keep the index in range
          const bool in = col < p.Lq && valid;
          const float pr = in ? fast_exp2(__uint_as_float(rs[2 * jj + u]) * sl2 - s_lse[ci]) : 0.f;
          float dp = __uint_as_float(rp[2 * jj + u]);
          float pt = pr;
          if (DROP) {
            const bool keep = drop_keep_att(seed_eff, (unsigned long long)((row_base + col) * p.Lk), j, p.drop_thresh);
            pt = keep ? pr * p.drop_scale : 0.f;
            dp = keep ? dp * p.drop_scale : 0.f;
          }
          pt2[u] = pt;
          ds2[u] = in ? pr * (dp - s_del[ci]) : 0.f;
        }
        pkp[jj] = pack_bf16(pt2[0], pt2[1]);
        pks[jj] = pack_bf16(ds2[0], ds2[1]);
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (mine) {
      tmem_st16(trow + (uint32_t)(16 * half), pkp);
      tmem_st16(trow + BW_DP + (uint32_t)(16 * half), pks);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      ptx::tc_fence_after();
      mma_ts_mn(tmem + BW_ACC0, tmem, sG + (size_t)c * 64 * 128, w, d, c > 0);          // dV += P~^T_c dO_c
      mma_ts_mn(tmem + BW_ACC1, tmem + BW_DP, sQ + (size_t)c * 64 * 128, w, d, c > 0);  // dK += dS^T_c Q_c
      if (c + 1 < nchunk) {
        const int w1 = min(64, p.Lq_pad - 64 * (c + 1));
        mma_ss_kk(tmem, sK, sQ + (size_t)(c + 1) * 64 * 128, w1, d);
        mma_ss_kk(tmem + BW_DP, sV, sG + (size_t)(c + 1) * 64 * 128, w1, d);
        ptx::umma_commit(&bars[0]);
      } else {
        ptx::umma_commit(&bars[1]);
      }
    }
  }
  ATT_STAMP(4);
  ptx::mbar_wait(&bars[1], 0);
  ptx::tc_fence_after();
  ATT_STAMP(5);
  if (half == 0)
    store_acc_row(trow + BW_ACC0, 0, d, 1.f, valid ? p.dv + skv * p.dv_bs + (long long)j * p.dv_rs + h * d : nullptr);
  else
    store_acc_row(trow + BW_ACC1, 0, d, p.scale, valid ? p.dk + skv * p.dk_bs + (long long)j * p.dk_rs + h * d : nullptr);
  ATT_STAMP(6);
  cta_epilogue<256>(tmem);
  ATT_STAMP(7);
}


// ======================================================================================= fused backward
// One CTA per (batch, head) computes dQ, dK and dV in a single pass -- the scores S and dP = dO V^T are produced
// ONCE (the two-kernel scheme above recomputes them and re-stages Q/K/V/dO per tile), rows = queries:
//   round (q-tile t, key chunk c of <= 128 keys):
//     S_c = Q_t K_c^T, dP_c = dO_t V_c^T          -> TMEM [0,128) / [128,256)
//     threads: P~ = dropout(exp(S - lse)), dS = P (dP~ - delta)   -> bf16 tiles sP, sDS[c] in shared memory, K-major
//     dV_c += P~^T dO_t,  dK_c += dS^T Q_t         -> the SAME shared tiles read as MN-major A operands (the UMMA
//                                                     descriptor's major-ness bit transposes for free); fp32 accumulators
//                                                     stay in TMEM [256,512) over all q-tiles
//   end of q-tile: dQ_t = dS_t K (A = sDS K-major over all key chunks) -> TMEM [0,d) -> scaled -> global.
// The next round's score MMAs are issued right behind the accumulating MMAs.  Nothing is recomputed, nothing is
// atomically added, and Q/K/V/dO are staged once per (batch, head); measured against the two-kernel scheme on B200:
// ViT-B (L = 197, d = 64) 742 -> 523 us per layer, EEG encoder (L = 139, d = 32, dropout) 926 -> 740 us.
constexpr uint32_t FB_S = 0, FB_DP = 128, FB_DV = 256, FB_DK = 384;

// acc[tmem_d] (+)= A^T-view[128 x 128 (k rows of the smem tile)] . B[k rows][d]: A tile stored K-major as
// [2 chunks of 64 m][128 k rows][128 B], read MN-major (M contiguous); B MN-major rows of 128 B
__device__ __forceinline__ void mma_ss_mnmn(uint32_t tmem_d, const uint8_t* a, const uint8_t* b, int ksteps, int d,
                                            bool accumulate) {
  const uint32_t idesc = ptx::make_idesc_bf16(TILE_ROWS, d, 1, 1);
  const uint64_t ad = ptx::make_smem_desc(ptx::smem_u32(a), 128u * 128u, 1024u);   // LBO = stride between 64-wide M chunks
  const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(b), 128u * 128u, 1024u);
  for (int ks = 0; ks < ksteps; ++ks)
    ptx::umma_bf16(tmem_d, ad + (uint64_t)(128 * ks), bd + (uint64_t)(128 * ks), idesc, (accumulate || ks > 0) ? 1u : 0u);
}
// acc[tmem_d] = A[128 x kdim] (K-major smem tile made of 64-wide chunks of [128 rows][128 B]) . B[kdim rows][d] (MN-major)
__device__ __forceinline__ void mma_ss_kmn(uint32_t tmem_d, const uint8_t* a, const uint8_t* b, int kdim, int d) {
  const uint32_t idesc = ptx::make_idesc_bf16(TILE_ROWS, d, 0, 1);
  const uint64_t ad = ptx::make_smem_desc(ptx::smem_u32(a), 16u, 1024u);
  const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(b), 128u * 128u, 1024u);
  for (int ks = 0; ks < kdim / 16; ++ks)
    ptx::umma_bf16(tmem_d, ad + (uint64_t)((ks >> 2) * (TILE_ROWS * 128 / 16) + (ks & 3) * 2), bd + (uint64_t)(128 * ks), idesc,
                   ks > 0 ? 1u : 0u);
}

template <bool DROP>
__global__ void __launch_bounds__(TC_THREADS, 1) att_tc_bwd_fused_kernel(const AttTcParams p) {
  const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);   // dropout seed of THIS replay
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sQ = align1024(smem_raw);
  uint8_t* sG = sQ + p.Lq_pad * 128;             // dO
  uint8_t* sK = sG + p.Lq_pad * 128;
  uint8_t* sV = sK + p.Lk_pad * 128;
  uint8_t* sP = sV + p.Lk_pad * 128;             // [2 x 64 keys][128 q rows][128 B]   P~ of the current round
  uint8_t* sDS = sP + 2 * TILE_ROWS * 128;       // [nk chunks][2 x 64 keys][128 q rows][128 B]  dS of the current q-tile
  const int nk = (p.Lk_pad + 127) / 128, nq = (p.Lq + 127) / 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + (size_t)nk * 2 * TILE_ROWS * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  uint64_t* bar_sp = &bars[0];
  uint64_t* bar_acc = &bars[1];
  uint64_t* bar_dq = &bars[2];
  const int h = blockIdx.x, s = blockIdx.y;
  const int skv = (s + p.kv_shift) % p.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp >> 2, row = (warp & 3) * 32 + lane;
  const int d = p.d;
  const long long row_base = ((long long)s * p.H + h) * p.Lq;

  ATT_STAMP(0);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_sp, 1);
    ptx::mbar_init(bar_acc, 1);
    ptx::mbar_init(bar_dq, 1);
    ptx::fence_barrier_init();
  }
  load_rows_sw128(sQ, p.q + s * p.q_bs + h * d, p.q_rs, p.Lq, p.Lq_pad, d);
  load_rows_sw128(sG, p.d_o + s * p.do_bs + h * d, p.do_rs, p.Lq, p.Lq_pad, d);
  load_rows_sw128(sK, p.k + skv * p.k_bs + h * d, p.k_rs, p.Lk, p.Lk_pad, d);
  load_rows_sw128(sV, p.v + skv * p.v_bs + h * d, p.v_rs, p.Lk, p.Lk_pad, d);
  // delta_i = dO_i . O_i and the row log-sum-exp: q-tile 0 while the tile copies are in flight, q-tile 1 later, behind
  // the first score MMAs
  float dl_t[2] = {0.f, 0.f}, lse_t[2] = {0.f, 0.f};
  auto row_stats = [&](int t) {
    const int i = t * TILE_ROWS + row;
    if (i < p.Lq) {
      const bf16* orow = p.o + s * p.o_bs + (long long)i * p.o_rs + h * d;
      const bf16* grow = p.d_o + s * p.do_bs + (long long)i * p.do_rs + h * d;
      uint4 ob[8], gb[8];
      const int n16 = d >> 3;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (u < n16) {
          ob[u] = __ldg(reinterpret_cast<const uint4*>(orow) + u);
          gb[u] = __ldg(reinterpret_cast<const uint4*>(grow) + u);
        }
      }
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (u < n16) {
          const __nv_bfloat162* oa = reinterpret_cast<const __nv_bfloat162*>(&ob[u]);
          const __nv_bfloat162* ga = reinterpret_cast<const __nv_bfloat162*>(&gb[u]);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            acc = fmaf(__low2float(oa[v]), __low2float(ga[v]), acc);
            acc = fmaf(__high2float(oa[v]), __high2float(ga[v]), acc);
          }
        }
      }
      dl_t[t] = acc;
      lse_t[t] = p.lse[row_base + i] * LOG2E;
    }
  };
  row_stats(0);
  if (warp == 0) ptx::tmem_alloc<512>(slot);
  ATT_STAMP(3);
  cp_async_wait_all();
  ATT_STAMP(4);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  ATT_STAMP(5);
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const float sl2 = p.scale * LOG2E;

  if (threadIdx.x == 0) {
    const int w0 = min(128, p.Lk_pad);
    mma_ss_kk(tmem + FB_S, sQ, sK, w0, d);
    mma_ss_kk(tmem + FB_DP, sG, sV, w0, d);
    ptx::umma_commit(bar_sp);
  }
  if (nq > 1) row_stats(1);
  int r = 0;                                      // round counter (barrier parities)
  for (int t = 0; t < nq; ++t) {
    const int i = t * TILE_ROWS + row;            // query row of this thread
    const bool valid = i < p.Lq;
    const long long row_id = row_base + (valid ? i : 0);
    const float dl = t == 0 ? dl_t[0] : dl_t[1], lse2 = t == 0 ? lse_t[0] : lse_t[1];
    for (int c = 0; c < nk; ++c, ++r) {
      const int w = min(128, p.Lk_pad - 128 * c);
      ptx::mbar_wait(bar_sp, (uint32_t)(r & 1));
      ptx::tc_fence_after();

      // this thread: 64 key columns [64 half, 64 half + 64) of the chunk, in two 32-column groups
      uint32_t pkp[2][16], pks[2][16];
      const bool mine = 64 * half < w;
      if (mine) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t rs[32], rp[32];
          ptx::tmem_ld32(trow + FB_S + (uint32_t)(64 * half + 32 * g), rs);
          ptx::tmem_ld32(trow + FB_DP + (uint32_t)(64 * half + 32 * g), rp);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float pt2[2], ds2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int col = 128 * c + 64 * half + 32 * g + 2 * j + u;   // key index
              const bool in = valid && col < p.Lk;
              const float pr = in ? fast_exp2(__uint_as_float(rs[2 * j + u]) * sl2 - lse2) : 0.f;
              float dp = __uint_as_float(rp[2 * j + u]);
              float pt = pr;
              if (DROP) {
                const bool keep = drop_keep_att(seed_eff, (unsigned long long)(row_id * p.Lk), col, p.drop_thresh);
                pt = keep ? pr * p.drop_scale : 0.f;
                dp = keep ? dp * p.drop_scale : 0.f;
              }
              pt2[u] = pt;
              ds2[u] = in ? pr * (dp - dl) : 0.f;
            }
            pkp[g][j] = pack_bf16(pt2[0], pt2[1]);
            pks[g][j] = pack_bf16(ds2[0], ds2[1]);
          }
        }
      }
      // the accumulating MMAs of the previous round read sP (and sDS): they must have retired before the overwrite
      if (r > 0) ptx::mbar_wait(bar_acc, (uint32_t)((r - 1) & 1));
      if (mine) {
        uint8_t* prow = sP + half * (TILE_ROWS * 128) + row * 128;
        uint8_t* srow = sDS + ((size_t)c * 2 + half) * (TILE_ROWS * 128) + row * 128;
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int slot16 = ((g * 4 + q4) ^ (row & 7)) << 4;
            *reinterpret_cast<uint4*>(prow + slot16) = make_uint4(pkp[g][4 * q4], pkp[g][4 * q4 + 1], pkp[g][4 * q4 + 2], pkp[g][4 * q4 + 3]);
            *reinterpret_cast<uint4*>(srow + slot16) = make_uint4(pks[g][4 * q4], pks[g][4 * q4 + 1], pks[g][4 * q4 + 2], pks[g][4 * q4 + 3]);
          }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncthreads();
      if (threadIdx.x == 0) {
        ptx::tc_fence_after();
        const uint8_t* gq = sG + (size_t)t * TILE_ROWS * 128;   // dO rows of this q-tile (MN-major B: k rows = queries)
        const uint8_t* qq = sQ + (size_t)t * TILE_ROWS * 128;
        mma_ss_mnmn(tmem + FB_DV + (uint32_t)(64 * c), sP, gq, TILE_ROWS / 16, d, t > 0);                                   // dV_c += P~^T dO_t
        mma_ss_mnmn(tmem + FB_DK + (uint32_t)(64 * c), sDS + (size_t)c * 2 * TILE_ROWS * 128, qq, TILE_ROWS / 16, d, t > 0);  // dK_c += dS^T Q_t
        ptx::umma_commit(bar_acc);
        if (c + 1 < nk) {
          const int w1 = min(128, p.Lk_pad - 128 * (c + 1));
          mma_ss_kk(tmem + FB_S, qq, sK + (size_t)(c + 1) * TILE_ROWS * 128, w1, d);
          mma_ss_kk(tmem + FB_DP, gq, sV + (size_t)(c + 1) * TILE_ROWS * 128, w1, d);
          ptx::umma_commit(bar_sp);
        } else {
          mma_ss_kmn(tmem + FB_S, sDS, sK, p.Lk_pad, d);          // dQ_t = dS_t K   (into the dead score columns)
          ptx::umma_commit(bar_dq);
        }
      }
    }
    ptx::mbar_wait(bar_dq, (uint32_t)(t & 1));
    ptx::tc_fence_after();

    {
      bf16* drow = valid ? p.dq + s * p.dq_bs + (long long)i * p.dq_rs + h * d : nullptr;
      if (d >= 64) store_acc_row(trow + FB_S, half * (d / 2), half * (d / 2) + d / 2, p.scale, drow);
      else if (half == 0) store_acc_row(trow + FB_S, 0, d, p.scale, drow);
    }
    if (t + 1 < nq) {
      ptx::tc_fence_before();
      __syncthreads();                            // dQ read out: the score columns may be overwritten
      if (threadIdx.x == 0) {
        ptx::tc_fence_after();
        const int w0 = min(128, p.Lk_pad);
        mma_ss_kk(tmem + FB_S, sQ + (size_t)(t + 1) * TILE_ROWS * 128, sK, w0, d);
        mma_ss_kk(tmem + FB_DP, sG + (size_t)(t + 1) * TILE_ROWS * 128, sV, w0, d);
        ptx::umma_commit(bar_sp);
      }
    }
  }
  ATT_STAMP(6);
  // bar_dq of the last q-tile covers every MMA issued before it: the dV / dK accumulators are final
  for (int c = 0; c < nk; ++c) {
    const int j = c * TILE_ROWS + row;            // key row of this thread
    const bool kv = j < p.Lk;
    if (half == 0)
      store_acc_row(trow + FB_DV + (uint32_t)(64 * c), 0, d, 1.f, kv ? p.dv + skv * p.dv_bs + (long long)j * p.dv_rs + h * d : nullptr);
    else
      store_acc_row(trow + FB_DK + (uint32_t)(64 * c), 0, d, p.scale, kv ? p.dk + skv * p.dk_bs + (long long)j * p.dk_rs + h * d : nullptr);
  }
  cta_epilogue<512>(tmem);
  ATT_STAMP(7);
}


// ======================================================================================= pipelined fused backward
// Same data flow as att_tc_bwd_fused_kernel, re-scheduled so that the tensor core and the softmax threads never wait
// for each other's latency (measured: in the kernel above every round pays the tcgen05 issue -> commit round trip
// plus the issuing thread's serial MMA stream on its critical path):
//   * a dedicated ninth warp issues every MMA; the 256 softmax threads never issue, they only signal mbarriers;
//   * rounds are 64 keys wide and the score pair (S, dP) is DOUBLE-BUFFERED in TMEM ([0,128) / [128,256)): the
//     scores of round r+1 are computed while the threads turn round r into P~ / dS;
//   * dV / dK accumulate once per PAIR of rounds (M = 128 keys) from the two 64-key halves of sP / sDS;
//   * dQ_t goes to spare TMEM columns when there are any (head_dim 32, or <= 128 keys), else into the score set the
//     q-tile's last round has just released.
// Barriers: bar_sp[set] (MMA -> threads: scores ready), bar_rd[set] (threads -> MMA: set consumed, tiles written),
// bar_acc (pair's accumulating MMAs retired: sP reusable), bar_dq (dQ_t and everything before it retired),
// bar_dqrd (threads -> MMA: dQ_t read out).
constexpr int PIPE_THREADS = TC_THREADS + 64;   // 8 softmax warps + 2 MMA-issuing warps

// 32 key columns of one query row of the backward pass: raw scores rs and dP~ = dO V^T rp (fp32 bit patterns from TMEM)
//   P = exp2(S sl2 - lse2),  P~ = dropout(P),  dS = P (dropout(dP~) - delta)      -> packed bf16 pairs pkp / pks.
// The loop is bound by instruction issue (two warps per scheduler), so the common case is stripped down: FAST = all
// 32 columns are real keys and the dropout counter does not carry into its high word -- no per-column predicates,
// one hash per column PAIR with the high-word product hoisted, keep-tests on the un-extracted halves
// (lo16 >= t  <=>  h << 16 >= t << 16;  hi16 >= t  <=>  h >= t << 16).  Rows past the sequence end come in with
// lse2 = +inf, which makes their P exactly 0 without a predicate.  Both paths produce the mask of drop_keep_att().
// NC (32 or 16) key columns starting at col0 (a multiple of NC; the dropout run is the aligned 32-column group)
template <bool DROP, bool FAST, int NC>
__device__ __forceinline__ void softmax_bwd_cols(const uint32_t (&rs)[NC], const uint32_t (&rp)[NC], float sl2, float lse2,
                                                 float dl, int col0, int Lk, unsigned long long seed,
                                                 unsigned long long row_lin, uint32_t thresh, float drop_scale,
                                                 uint32_t (&pkp)[NC / 2], uint32_t (&pks)[NC / 2]) {
  uint32_t s0 = 0u;
  if (DROP) {
    s0 = drop_att_run_state(seed, row_lin, col0 & ~31);
    if (NC == 16 && (col0 & 16)) {                       // second half of the run: 16 steps ahead
      const DropLeap l16 = drop_leap(16);
      s0 = s0 * l16.mul + l16.add;
    }
  }
#pragma unroll
  for (int j = 0; j < NC / 2; ++j) {
    float m0 = 1.f, m1 = 1.f;
    if (DROP) {
      const DropLeap l0 = drop_leap(2 * j), l1 = drop_leap(2 * j + 1);
      m0 = (s0 * l0.mul + l0.add) >= thresh ? drop_scale : 0.f;
      m1 = (s0 * l1.mul + l1.add) >= thresh ? drop_scale : 0.f;
    }
    float p0 = fast_exp2(fmaf(__uint_as_float(rs[2 * j]), sl2, -lse2));
    float p1 = fast_exp2(fmaf(__uint_as_float(rs[2 * j + 1]), sl2, -lse2));
    if (!FAST) {
      if (col0 + 2 * j >= Lk) p0 = 0.f;
      if (col0 + 2 * j + 1 >= Lk) p1 = 0.f;
    }
    const float d0 = __uint_as_float(rp[2 * j]), d1 = __uint_as_float(rp[2 * j + 1]);
    pkp[j] = pack_bf16(p0 * m0, p1 * m1);
    pks[j] = pack_bf16(p0 * (DROP ? fmaf(d0, m0, -dl) : d0 - dl), p1 * (DROP ? fmaf(d1, m1, -dl) : d1 - dl));
  }
}


template <bool DROP, int NSW>
__global__ void __launch_bounds__(NSW * 32 + 64, 1) att_tc_bwd_pipe_kernel(const __grid_constant__ AttTcParams p, const __grid_constant__ AttMaps maps) {
  const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);   // dropout seed of THIS replay
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sQ = align1024(smem_raw);
  uint8_t* sG = sQ + p.Lq_pad * 128;             // dO
  uint8_t* sK = sG + p.Lq_pad * 128;
  uint8_t* sV = sK + p.Lk_pad * 128;
  uint8_t* sP = sV + p.Lk_pad * 128;             // [2 x 64 keys][128 q rows][128 B]   P~ of the current pair of rounds
  uint8_t* sDS = sP + 2 * TILE_ROWS * 128;       // [nc x 64 keys][128 q rows][128 B]  dS of the current q-tile
  const int nk = (p.Lk_pad + 127) / 128, nc = (p.Lk_pad + 63) / 64, nq = (p.Lq + 127) / 128;
  const int R = nq * nc;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + (size_t)nk * 2 * TILE_ROWS * 128);
  uint64_t* bar_sp = &bars[0];                   // [2]
  uint64_t* bar_rd = &bars[2];                   // [2]
  uint64_t* bar_acc = &bars[4];
  uint64_t* bar_dq = &bars[5];
  uint64_t* bar_dqrd = &bars[6];
  uint64_t* bar_ld = &bars[7];                   // TMA tile loads
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  // [8 warps][3][64] column sums of this head's dQ | dK | dV rows, one private slot per softmax warp
  float* s_cs_all = reinterpret_cast<float*>(slot + 4);
  const bool want_cs = p.dq_cs != nullptr;
  if (want_cs)
    for (int i = threadIdx.x; i < 8 * 192; i += NSW * 32 + 64) s_cs_all[i] = 0.f;   // published by the __syncthreads below
  float* s_cs = s_cs_all + (threadIdx.x >> 5 & 7) * 192;   // (private per STORING warp: warps 0..7)
  const int h = blockIdx.x, s = blockIdx.y;
  const int skv = (s + p.kv_shift) % p.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int SM_THREADS = NSW * 32;            // softmax threads
  constexpr int CW = 64 / (NSW / 4);              // key columns of a 64-key round per thread (32 or 16)
  const int quarter = warp >> 2;                  // column share of this warp within a round
  const int half = quarter & 1, row = (warp & 3) * 32 + lane;   // `half`: column half in the store phases (quarter < 2)
  const bool storer = quarter < 2;                // warps that read out and store the results
  const int d = p.d;
  const long long row_base = ((long long)s * p.H + h) * p.Lq;
  const bool softmax_thread = warp < NSW;
  // TMEM columns: score sets [0,128) [128,256); dV_kc, dK_kc (d columns each per 128-key chunk), dQ
  const uint32_t DV0 = 256, DK0 = 256 + (uint32_t)(d * nk), DQ0 = 256 + (uint32_t)(2 * d * nk);
  const bool dq_alias = DQ0 + (uint32_t)d > 512u;

  ATT_STAMP_MID(0);
  long long w_sp = 0, w_acc = 0, w_dq = 0, w_el = 0, w_wr = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_sp[0], 1);
    ptx::mbar_init(&bar_sp[1], 1);
    ptx::mbar_init(&bar_rd[0], SM_THREADS);
    ptx::mbar_init(&bar_rd[1], SM_THREADS);
    ptx::mbar_init(bar_acc, 1);
    ptx::mbar_init(bar_dq, 1);
    ptx::mbar_init(bar_dqrd, SM_THREADS);
    ptx::mbar_init(bar_ld, 1);
    ptx::fence_barrier_init();
    if (p.use_tma) {
      // one bulk tensor copy per tile; rows past the sequence end are zero-filled by the TMA unit.  (The cp.async
      // path below is capped by the SM's outstanding-request budget at ~16 B/clk against DRAM latency: measured
      // 8 K cycles for the 132 KB of a ViT-B head, a quarter of the CTA's life.)
      ptx::mbar_arrive_expect_tx(bar_ld, (uint32_t)((3 * p.Lq_pad + 2 * p.Lk_pad) * 128));
      ptx::tma_load_3d(sQ, &maps.q, bar_ld, h * d, 0, s);
      ptx::tma_load_3d(sG, &maps.g, bar_ld, h * d, 0, s);
      ptx::tma_load_3d(sK, &maps.k, bar_ld, h * d, 0, skv);
      ptx::tma_load_3d(sV, &maps.v, bar_ld, h * d, 0, skv);
      ptx::tma_load_3d(sP, &maps.o, bar_ld, h * d, 0, s);
    }
  }
  float dl_t[2] = {0.f, 0.f}, lse_t[2] = {0.f, 0.f};
  const int n16 = d >> 3;
  if (softmax_thread) {
   if (!p.use_tma) {
    load_rows_sw128(sQ, p.q + s * p.q_bs + h * d, p.q_rs, p.Lq, p.Lq_pad, d, SM_THREADS);
    load_rows_sw128(sG, p.d_o + s * p.do_bs + h * d, p.do_rs, p.Lq, p.Lq_pad, d, SM_THREADS);
    load_rows_sw128(sK, p.k + skv * p.k_bs + h * d, p.k_rs, p.Lk, p.Lk_pad, d, SM_THREADS);
    load_rows_sw128(sV, p.v + skv * p.v_bs + h * d, p.v_rs, p.Lk, p.Lk_pad, d, SM_THREADS);
    // delta_i = dO_i . O_i: O is staged like the other tiles (coalesced; a per-thread read of its own row straight
    // from global memory costs 32 sectors per instruction) into the still idle sP area
    load_rows_sw128(sP, p.o + s * p.o_bs + h * d, p.o_rs, p.Lq, p.Lq_pad, d, SM_THREADS);
   }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = t * TILE_ROWS + row;
      if (i < p.Lq) lse_t[t] = p.lse[row_base + i];
    }
    ATT_STAMP_MID(1);
    cp_async_wait_all();
    ptx::fence_proxy_async_smem();
  } else if (warp == NSW) {
    ptx::tmem_alloc<512>(slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (p.use_tma) ptx::mbar_wait(bar_ld, 0);
  ATT_STAMP_MID(2);

  if (!softmax_thread) {
    // ------------------------------------------------------------------------------------------ MMA issuers
    // Two issuing threads (each tcgen05.commit tracks its own thread's MMAs): warp 8 feeds the score sets, warp 9
    // the accumulators and dQ -- a single thread's serial instruction stream (~40 cycles per MMA) put the 16
    // accumulating MMAs of a pair in front of the next scores.
    if (lane == 0 && warp == NSW) {
      auto issue_scores = [&](int r2) {
        const int t2 = r2 / nc, c2 = r2 - t2 * nc;
        const int w2 = min(64, p.Lk_pad - 64 * c2);
        const uint32_t sb = tmem + 128u * (uint32_t)(r2 & 1);
        mma_ss_kk(sb, sQ + (size_t)t2 * TILE_ROWS * 128, sK + (size_t)c2 * 64 * 128, w2, d);
        mma_ss_kk(sb + 64, sG + (size_t)t2 * TILE_ROWS * 128, sV + (size_t)c2 * 64 * 128, w2, d);
        ptx::umma_commit(&bar_sp[r2 & 1]);
      };
      issue_scores(0);
      if (R > 1) issue_scores(1);
      for (int r = 0, t = 0, c = 0; r + 2 < R; ++r) {
        ptx::mbar_wait(&bar_rd[r & 1], (uint32_t)((r >> 1) & 1));   // set r & 1 consumed
        if (c == nc - 1 && dq_alias) ptx::mbar_wait(bar_dqrd, (uint32_t)(t & 1));   // ... and dQ_t, parked there, read out
        ptx::tc_fence_after();
        issue_scores(r + 2);
        if (++c == nc) { c = 0; ++t; }
      }
    } else if (lane == 0 && warp == NSW + 1) {
      for (int r = 0, t = 0, c = 0; r < R; ++r) {
        const bool last_c = c == nc - 1;
        if ((c & 1) || last_c) {
          ptx::mbar_wait(&bar_rd[r & 1], (uint32_t)((r >> 1) & 1));   // P~ / dS tiles of the pair written
          ptx::tc_fence_after();
          const uint8_t* gq = sG + (size_t)t * TILE_ROWS * 128;   // dO rows of this q-tile (MN-major B: k rows = queries)
          const uint8_t* qq = sQ + (size_t)t * TILE_ROWS * 128;
          const int kc = c >> 1;
          mma_ss_mnmn(tmem + DV0 + (uint32_t)(d * kc), sP, gq, TILE_ROWS / 16, d, t > 0);                                     // dV_kc += P~^T dO_t
          mma_ss_mnmn(tmem + DK0 + (uint32_t)(d * kc), sDS + (size_t)kc * 2 * TILE_ROWS * 128, qq, TILE_ROWS / 16, d, t > 0); // dK_kc += dS^T Q_t
          ptx::umma_commit(bar_acc);
          if (last_c) {
            mma_ss_kmn(tmem + (dq_alias ? 128u * (uint32_t)(r & 1) : DQ0), sDS, sK, p.Lk_pad, d);   // dQ_t = dS_t K
            ptx::umma_commit(bar_dq);
          }
        }
        if (++c == nc) { c = 0; ++t; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------ softmax threads
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float sl2 = p.scale * LOG2E;
    uint8_t* stage = sP + (warp & 7) * 4096;      // warp-private staging rows of the coalesced result stores (warps 0..7)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = t * TILE_ROWS + row;
      if (i < p.Lq) {
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (u < n16) {
            const size_t off = (size_t)i * 128 + ((u ^ (i & 7)) << 4);
            const uint4 gb = *reinterpret_cast<const uint4*>(sG + off);
            const uint4 ob = *reinterpret_cast<const uint4*>(sP + off);
            const __nv_bfloat162* oa = reinterpret_cast<const __nv_bfloat162*>(&ob);
            const __nv_bfloat162* ga = reinterpret_cast<const __nv_bfloat162*>(&gb);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              acc = fmaf(__low2float(oa[v]), __low2float(ga[v]), acc);
              acc = fmaf(__high2float(oa[v]), __high2float(ga[v]), acc);
            }
          }
        }
        dl_t[t] = acc;
        lse_t[t] *= LOG2E;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");   // O rows consumed: sP may be written
    auto read_dq = [&](int tq) {
      ATT_CLK(c0);
      ptx::mbar_wait(bar_dq, (uint32_t)(tq & 1));
      ATT_CLK(c1);
      w_dq += c1 - c0;
      ptx::tc_fence_after();
      const uint32_t col = dq_alias ? 128u * (uint32_t)(((tq + 1) * nc - 1) & 1) : DQ0;
      // sP is idle here (bar_dq covers the accumulating MMAs that read it; this round's tiles are written later)
      const int r0 = tq * TILE_ROWS + (warp & 3) * 32;
      const int cb = d >= 64 ? half * 32 : 0;
      if (storer && (d >= 64 || half == 0))
        store_acc_rows_coalesced(trow + col + (uint32_t)cb, 32, p.scale, stage,
                                 p.dq + s * p.dq_bs + (long long)r0 * p.dq_rs + h * d + cb, p.dq_rs, p.Lq - r0, lane,
                                 want_cs ? s_cs + cb : nullptr);
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_dqrd);
      asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");   // every warp's staging rows are free again
    };
    int pairs = 0;                                // accumulating MMA groups issued so far
    for (int r = 0, t = 0, c = 0; r < R; ++r) {
      const int set = r & 1;
      const int w = min(64, p.Lk_pad - 64 * c);
      const int i = t * TILE_ROWS + row;          // query row of this thread
      const bool valid = i < p.Lq;
      const long long row_id = row_base + (valid ? i : 0);
      const float dl = t == 0 ? dl_t[0] : dl_t[1];
      const float lse2 = !valid ? INFINITY : t == 0 ? lse_t[0] : lse_t[1];   // +inf: P = 0 on the padding rows
      ATT_CLK(c0);
      ptx::mbar_wait(&bar_sp[set], (uint32_t)((r >> 1) & 1));
      ATT_CLK(c1);
      w_sp += c1 - c0;
      ptx::tc_fence_after();

      // this thread: CW key columns [CW quarter, CW quarter + CW) of the 64-key round
      uint32_t pkp[CW / 2], pks[CW / 2];
      const bool mine = CW * quarter < w;
      // a warp whose 32 query rows all lie past the sequence end (second q-tile: L = 139 leaves 11 live rows, L = 197
      // leaves 69) has P = dS = 0 on all of them: it writes the zeros without the TMEM reads and the softmax arithmetic
      const bool warp_live = t * TILE_ROWS + (warp & 3) * 32 < p.Lq;
      if (mine && !warp_live) {
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) pkp[j] = pks[j] = 0u;
      }
      if (mine && warp_live) {
        uint32_t rs[CW], rp[CW];
        tmem_ld_cols<CW>(trow + 128u * (uint32_t)set + (uint32_t)(CW * quarter), rs);
        tmem_ld_cols<CW>(trow + 128u * (uint32_t)set + 64u + (uint32_t)(CW * quarter), rp);
        ptx::tmem_ld_wait();
        const int col0 = 64 * c + CW * quarter;   // first key column of this thread
        const unsigned long long row_lin = (unsigned long long)(row_id * p.Lk);
        const bool fast = col0 + CW <= p.Lk;        // all columns are real keys: no per-column predicates
        if (fast) softmax_bwd_cols<DROP, true, CW>(rs, rp, sl2, lse2, dl, col0, p.Lk, seed_eff, row_lin, p.drop_thresh, p.drop_scale, pkp, pks);
        else softmax_bwd_cols<DROP, false, CW>(rs, rp, sl2, lse2, dl, col0, p.Lk, seed_eff, row_lin, p.drop_thresh, p.drop_scale, pkp, pks);
      }
      // previous q-tile's dQ: read it out behind this round's arithmetic; its barrier also covers the MMAs that read
      // the previous q-tile's sDS, which this q-tile now overwrites
      ATT_CLK(c2);
      w_el += c2 - c1;
      if (c == 0 && t > 0) read_dq(t - 1);
      // the accumulating MMAs of the previous pair read sP: they must have retired before the first overwrite
      ATT_CLK(c3);
      if ((c & 1) == 0 && pairs > 0) ptx::mbar_wait(bar_acc, (uint32_t)((pairs - 1) & 1));
      ATT_CLK(c4);
      w_acc += c4 - c3;
      if (mine) {
        uint8_t* prow = sP + (c & 1) * (TILE_ROWS * 128) + row * 128;
        uint8_t* srow = sDS + (size_t)c * (TILE_ROWS * 128) + row * 128;
#pragma unroll
        for (int q4 = 0; q4 < CW / 8; ++q4) {
          const int slot16 = ((quarter * (CW / 8) + q4) ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + slot16) = make_uint4(pkp[4 * q4], pkp[4 * q4 + 1], pkp[4 * q4 + 2], pkp[4 * q4 + 3]);
          *reinterpret_cast<uint4*>(srow + slot16) = make_uint4(pks[4 * q4], pks[4 * q4 + 1], pks[4 * q4 + 2], pks[4 * q4 + 3]);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bar_rd[set]);
      ATT_CLK(c5);
      w_wr += c5 - c4;
      if ((c & 1) || c == nc - 1) ++pairs;
      if (++c == nc) { c = 0; ++t; }
    }
    ATT_STAMP_MID(3);
    // Two key chunks: dV_0 / dK_0 have been final since the last q-tile's first pair retired (bar_acc, waited for in
    // round c = 2) -- store them while the last pair's MMAs and dQ drain.  Staging goes to sV, which only the score
    // MMAs read and those have all been consumed (sP / sDS / sQ / sG / sK are still being read).
    const bool early0 = nk == 2;
    if (early0 && storer) {
      const int j0 = (warp & 3) * 32;
      uint8_t* st2 = sV + warp * 2048;
      for (int c0 = 0; c0 < d; c0 += 32) {
        if (half == 0)
          store_acc_rows32_coalesced(trow + DV0 + (uint32_t)c0, 1.f, st2, p.dv + skv * p.dv_bs + (long long)j0 * p.dv_rs + h * d + c0,
                                     p.dv_rs, p.Lk - j0, lane, want_cs ? s_cs + 128 + c0 : nullptr);
        else
          store_acc_rows32_coalesced(trow + DK0 + (uint32_t)c0, p.scale, st2, p.dk + skv * p.dk_bs + (long long)j0 * p.dk_rs + h * d + c0,
                                     p.dk_rs, p.Lk - j0, lane, want_cs ? s_cs + 64 + c0 : nullptr);
      }
    }
    read_dq(nq - 1);                              // bar_dq of the last q-tile covers every MMA: dV / dK are final
    for (int kc = early0 ? 1 : 0; kc < nk && storer; ++kc) {
      const int j0 = kc * TILE_ROWS + (warp & 3) * 32;   // first key row of this warp
      if (half == 0)
        store_acc_rows_coalesced(trow + DV0 + (uint32_t)(d * kc), d, 1.f, stage,
                                 p.dv + skv * p.dv_bs + (long long)j0 * p.dv_rs + h * d, p.dv_rs, p.Lk - j0, lane,
                                 want_cs ? s_cs + 128 : nullptr);
      else
        store_acc_rows_coalesced(trow + DK0 + (uint32_t)(d * kc), d, p.scale, stage,
                                 p.dk + skv * p.dk_bs + (long long)j0 * p.dk_rs + h * d, p.dk_rs, p.Lk - j0, lane,
                                 want_cs ? s_cs + 64 : nullptr);
    }
  }
  ATT_STAMP_MID(4);
  if (p.dbg != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2) {
    p.dbg[8] = w_sp; p.dbg[9] = w_acc; p.dbg[10] = w_dq; p.dbg[11] = w_el; p.dbg[12] = w_wr;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ATT_STAMP_MID(5);
  if (warp == NSW) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
  if (want_cs && threadIdx.x < 3 * d) {          // one global atomic per column per CTA
    const int which = threadIdx.x / d, c = threadIdx.x - which * d;
    float* dst = which == 0 ? p.dq_cs : which == 1 ? p.dk_cs : p.dv_cs;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s_cs_all[w * 192 + which * 64 + c];
    atomicAdd(dst + h * d + c, v);
  }
}

// ======================================================================================= persistent pipelined backward
// att_tc_bwd_pipe_kernel with ONE CTA PER SM walking over the (batch, head) items: tensor memory and barriers are set up
// once, and the operand tiles of item i+1 are requested as soon as item i's last MMA has retired, i.e. they arrive
// while item i's dV / dK rows are still being stored.  O is staged in the (then idle) dS area instead of the P~ area,
// which doubles as the staging space of those stores.
template <bool DROP, int NSW>
__global__ void __launch_bounds__(NSW * 32 + 64, 1) att_tc_bwd_pers_kernel(const __grid_constant__ AttTcParams p, const __grid_constant__ AttMaps maps) {
  const unsigned long long seed_eff = egb_mix_seed(p.seed, p.epoch);   // dropout seed of THIS replay
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sQ = align1024(smem_raw);
  uint8_t* sG = sQ + p.Lq_pad * 128;             // dO
  uint8_t* sK = sG + p.Lq_pad * 128;
  uint8_t* sV = sK + p.Lk_pad * 128;
  uint8_t* sP = sV + p.Lk_pad * 128;             // [2 x 64 keys][128 q rows][128 B]   P~ of the current pair of rounds
  uint8_t* sDS = sP + 2 * TILE_ROWS * 128;       // [nc x 64 keys][128 q rows][128 B]  dS of the current q-tile; O at item start
  const int nk = (p.Lk_pad + 127) / 128, nc = (p.Lk_pad + 63) / 64, nq = (p.Lq + 127) / 128;
  const int R = nq * nc;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + (size_t)nk * 2 * TILE_ROWS * 128);
  uint64_t* bar_sp = &bars[0];                   // [2]
  uint64_t* bar_rd = &bars[2];                   // [2]
  uint64_t* bar_acc = &bars[4];
  uint64_t* bar_dq = &bars[5];
  uint64_t* bar_dqrd = &bars[6];
  uint64_t* bar_ld = &bars[7];                   // [2] TMA tile loads of even / odd items of this CTA
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 9);
  // [8 warps][3][64] column sums of this head's dQ | dK | dV rows, one private slot per softmax warp
  float* s_cs_all = reinterpret_cast<float*>(slot + 4);
  const bool want_cs = p.dq_cs != nullptr;
  float* s_cs = s_cs_all + (threadIdx.x >> 5 & 7) * 192;   // (private per STORING warp: warps 0..7)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int SM_THREADS = NSW * 32;            // softmax threads
  constexpr int CW = 64 / (NSW / 4);              // key columns of a 64-key round per thread (32 or 16)
  const int quarter = warp >> 2;                  // column share of this warp within a round
  const int half = quarter & 1, row = (warp & 3) * 32 + lane;   // `half`: column half in the store phases (quarter < 2)
  const bool storer = quarter < 2;                // warps that read out and store the results
  const int d = p.d;
  const bool softmax_thread = warp < NSW;
  // TMEM columns: score sets [0,128) [128,256); dV_kc, dK_kc (d columns each per 128-key chunk), dQ
  const uint32_t DV0 = 256, DK0 = 256 + (uint32_t)(d * nk), DQ0 = 256 + (uint32_t)(2 * d * nk);
  const bool dq_alias = DQ0 + (uint32_t)d > 512u;
  const int n_items = p.H * p.S;
  const int n16 = d >> 3;
  long long w_sp = 0, w_acc = 0, w_dq = 0, w_el = 0, w_wr = 0;

  auto init_round_bars = [&](bool again) {
    if (again) {                                 // an mbarrier object must be invalidated before it is initialised anew
      for (int i = 0; i < 7; ++i) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(&bars[i])) : "memory");
    }
    ptx::mbar_init(&bar_sp[0], 1);
    ptx::mbar_init(&bar_sp[1], 1);
    ptx::mbar_init(&bar_rd[0], SM_THREADS);
    ptx::mbar_init(&bar_rd[1], SM_THREADS);
    ptx::mbar_init(bar_acc, 1);
    ptx::mbar_init(bar_dq, 1);
    ptx::mbar_init(bar_dqrd, SM_THREADS);
  };
  // operand tiles of one item: Q, dO, K, V and O (O goes to the sDS area, idle between items).  TMA: thread 0 arms the
  // item's load barrier and issues five bulk tensor copies; cp.async: every softmax thread issues its share.
  auto issue_loads = [&](int item, int k) {
    const int hh = item % p.H, ss = item / p.H;
    const int skv2 = (ss + p.kv_shift) % p.S;
    if (p.use_tma) {
      if (threadIdx.x == 0) {
        uint64_t* bl = &bar_ld[k & 1];
        ptx::mbar_arrive_expect_tx(bl, (uint32_t)((3 * p.Lq_pad + 2 * p.Lk_pad) * 128));
        ptx::tma_load_3d(sQ, &maps.q, bl, hh * d, 0, ss);
        ptx::tma_load_3d(sG, &maps.g, bl, hh * d, 0, ss);
        ptx::tma_load_3d(sK, &maps.k, bl, hh * d, 0, skv2);
        ptx::tma_load_3d(sV, &maps.v, bl, hh * d, 0, skv2);
        ptx::tma_load_3d(sDS, &maps.o, bl, hh * d, 0, ss);
      }
    } else if (softmax_thread) {
      load_rows_sw128(sQ, p.q + ss * p.q_bs + hh * d, p.q_rs, p.Lq, p.Lq_pad, d, SM_THREADS);
      load_rows_sw128(sG, p.d_o + ss * p.do_bs + hh * d, p.do_rs, p.Lq, p.Lq_pad, d, SM_THREADS);
      load_rows_sw128(sK, p.k + skv2 * p.k_bs + hh * d, p.k_rs, p.Lk, p.Lk_pad, d, SM_THREADS);
      load_rows_sw128(sV, p.v + skv2 * p.v_bs + hh * d, p.v_rs, p.Lk, p.Lk_pad, d, SM_THREADS);
      load_rows_sw128(sDS, p.o + ss * p.o_bs + hh * d, p.o_rs, p.Lq, p.Lq_pad, d, SM_THREADS);
    }
  };

  if (threadIdx.x == 0) {
    init_round_bars(false);
    ptx::mbar_init(&bar_ld[0], 1);
    ptx::mbar_init(&bar_ld[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == NSW) ptx::tmem_alloc<512>(slot);
  issue_loads(blockIdx.x, 0);                    // (bar_ld is initialised by the issuing thread itself)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;

  int k_item = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k_item) {
  const int h = item % p.H, s = item / p.H;
  const int skv = (s + p.kv_shift) % p.S;
  const long long row_base = ((long long)s * p.H + h) * p.Lq;
  const int next_item = item + gridDim.x;
  float dl_t[2] = {0.f, 0.f}, lse_t[2] = {0.f, 0.f};
  if (want_cs)
    for (int i = threadIdx.x; i < 8 * 192; i += SM_THREADS + 64) s_cs_all[i] = 0.f;   // published by the __syncthreads below
  if (softmax_thread) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = t * TILE_ROWS + row;
      if (i < p.Lq) lse_t[t] = p.lse[row_base + i];
    }
    if (!p.use_tma) {
      cp_async_wait_all();
      ptx::fence_proxy_async_smem();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();                               // this item's tiles (cp.async) and the re-initialised barriers are visible
  ptx::tc_fence_after();
  if (p.use_tma) ptx::mbar_wait(&bar_ld[k_item & 1], (uint32_t)((k_item >> 1) & 1));

  if (!softmax_thread) {
    // ------------------------------------------------------------------------------------------ MMA issuers
    // Two issuing threads (each tcgen05.commit tracks its own thread's MMAs): warp 8 feeds the score sets, warp 9
    // the accumulators and dQ -- a single thread's serial instruction stream (~40 cycles per MMA) put the 16
    // accumulating MMAs of a pair in front of the next scores.
    if (lane == 0 && warp == NSW) {
      auto issue_scores = [&](int r2) {
        const int t2 = r2 / nc, c2 = r2 - t2 * nc;
        const int w2 = min(64, p.Lk_pad - 64 * c2);
        const uint32_t sb = tmem + 128u * (uint32_t)(r2 & 1);
        mma_ss_kk(sb, sQ + (size_t)t2 * TILE_ROWS * 128, sK + (size_t)c2 * 64 * 128, w2, d);
        mma_ss_kk(sb + 64, sG + (size_t)t2 * TILE_ROWS * 128, sV + (size_t)c2 * 64 * 128, w2, d);
        ptx::umma_commit(&bar_sp[r2 & 1]);
      };
      issue_scores(0);
      if (R > 1) issue_scores(1);
      for (int r = 0, t = 0, c = 0; r + 2 < R; ++r) {
        ptx::mbar_wait(&bar_rd[r & 1], (uint32_t)((r >> 1) & 1));   // set r & 1 consumed
        if (c == nc - 1 && dq_alias) ptx::mbar_wait(bar_dqrd, (uint32_t)(t & 1));   // ... and dQ_t, parked there, read out
        ptx::tc_fence_after();
        issue_scores(r + 2);
        if (++c == nc) { c = 0; ++t; }
      }
    } else if (lane == 0 && warp == NSW + 1) {
      for (int r = 0, t = 0, c = 0; r < R; ++r) {
        const bool last_c = c == nc - 1;
        if ((c & 1) || last_c) {
          ptx::mbar_wait(&bar_rd[r & 1], (uint32_t)((r >> 1) & 1));   // P~ / dS tiles of the pair written
          ptx::tc_fence_after();
          const uint8_t* gq = sG + (size_t)t * TILE_ROWS * 128;   // dO rows of this q-tile (MN-major B: k rows = queries)
          const uint8_t* qq = sQ + (size_t)t * TILE_ROWS * 128;
          const int kc = c >> 1;
          mma_ss_mnmn(tmem + DV0 + (uint32_t)(d * kc), sP, gq, TILE_ROWS / 16, d, t > 0);                                     // dV_kc += P~^T dO_t
          mma_ss_mnmn(tmem + DK0 + (uint32_t)(d * kc), sDS + (size_t)kc * 2 * TILE_ROWS * 128, qq, TILE_ROWS / 16, d, t > 0); // dK_kc += dS^T Q_t
          ptx::umma_commit(bar_acc);
          if (last_c) {
            mma_ss_kmn(tmem + (dq_alias ? 128u * (uint32_t)(r & 1) : DQ0), sDS, sK, p.Lk_pad, d);   // dQ_t = dS_t K
            ptx::umma_commit(bar_dq);
          }
        }
        if (++c == nc) { c = 0; ++t; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------ softmax threads
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float sl2 = p.scale * LOG2E;
    uint8_t* stage = sP + (warp & 7) * 4096;      // warp-private staging rows of the coalesced result stores (warps 0..7)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = t * TILE_ROWS + row;
      if (i < p.Lq) {
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (u < n16) {
            const size_t off = (size_t)i * 128 + ((u ^ (i & 7)) << 4);
            const uint4 gb = *reinterpret_cast<const uint4*>(sG + off);
            const uint4 ob = *reinterpret_cast<const uint4*>(sDS + off);
            const __nv_bfloat162* oa = reinterpret_cast<const __nv_bfloat162*>(&ob);
            const __nv_bfloat162* ga = reinterpret_cast<const __nv_bfloat162*>(&gb);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              acc = fmaf(__low2float(oa[v]), __low2float(ga[v]), acc);
              acc = fmaf(__high2float(oa[v]), __high2float(ga[v]), acc);
            }
          }
        }
        dl_t[t] = acc;
        lse_t[t] *= LOG2E;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");   // O rows consumed: sDS may be written
    auto read_dq = [&](int tq) {
      ATT_CLK(c0);
      ptx::mbar_wait(bar_dq, (uint32_t)(tq & 1));
      ATT_CLK(c1);
      w_dq += c1 - c0;
      ptx::tc_fence_after();
      const uint32_t col = dq_alias ? 128u * (uint32_t)(((tq + 1) * nc - 1) & 1) : DQ0;
      // sP is idle here (bar_dq covers the accumulating MMAs that read it; this round's tiles are written later)
      const int r0 = tq * TILE_ROWS + (warp & 3) * 32;
      const int cb = d >= 64 ? half * 32 : 0;
      if (storer && (d >= 64 || half == 0))
        store_acc_rows_coalesced(trow + col + (uint32_t)cb, 32, p.scale, stage,
                                 p.dq + s * p.dq_bs + (long long)r0 * p.dq_rs + h * d + cb, p.dq_rs, p.Lq - r0, lane,
                                 want_cs ? s_cs + cb : nullptr);
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_dqrd);
      asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory");   // every warp's staging rows are free again
    };
    int pairs = 0;                                // accumulating MMA groups issued so far
    for (int r = 0, t = 0, c = 0; r < R; ++r) {
      const int set = r & 1;
      const int w = min(64, p.Lk_pad - 64 * c);
      const int i = t * TILE_ROWS + row;          // query row of this thread
      const bool valid = i < p.Lq;
      const long long row_id = row_base + (valid ? i : 0);
      const float dl = t == 0 ? dl_t[0] : dl_t[1];
      const float lse2 = !valid ? INFINITY : t == 0 ? lse_t[0] : lse_t[1];   // +inf: P = 0 on the padding rows
      ATT_CLK(c0);
      ptx::mbar_wait(&bar_sp[set], (uint32_t)((r >> 1) & 1));
      ATT_CLK(c1);
      w_sp += c1 - c0;
      ptx::tc_fence_after();

      // this thread: CW key columns [CW quarter, CW quarter + CW) of the 64-key round
      uint32_t pkp[CW / 2], pks[CW / 2];
      const bool mine = CW * quarter < w;
      // a warp whose 32 query rows all lie past the sequence end (second q-tile: L = 139 leaves 11 live rows, L = 197
      // leaves 69) has P = dS = 0 on all of them: it writes the zeros without the TMEM reads and the softmax arithmetic
      const bool warp_live = t * TILE_ROWS + (warp & 3) * 32 < p.Lq;
      if (mine && !warp_live) {
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) pkp[j] = pks[j] = 0u;
      }
      if (mine && warp_live) {
        uint32_t rs[CW], rp[CW];
        tmem_ld_cols<CW>(trow + 128u * (uint32_t)set + (uint32_t)(CW * quarter), rs);
        tmem_ld_cols<CW>(trow + 128u * (uint32_t)set + 64u + (uint32_t)(CW * quarter), rp);
        ptx::tmem_ld_wait();
        const int col0 = 64 * c + CW * quarter;   // first key column of this thread
        const unsigned long long row_lin = (unsigned long long)(row_id * p.Lk);
        const bool fast = col0 + CW <= p.Lk;        // all columns are real keys: no per-column predicates
        if (fast) softmax_bwd_cols<DROP, true, CW>(rs, rp, sl2, lse2, dl, col0, p.Lk, seed_eff, row_lin, p.drop_thresh, p.drop_scale, pkp, pks);
        else softmax_bwd_cols<DROP, false, CW>(rs, rp, sl2, lse2, dl, col0, p.Lk, seed_eff, row_lin, p.drop_thresh, p.drop_scale, pkp, pks);
      }
      // previous q-tile's dQ: read it out behind this round's arithmetic; its barrier also covers the MMAs that read
      // the previous q-tile's sDS, which this q-tile now overwrites
      ATT_CLK(c2);
      w_el += c2 - c1;
      if (c == 0 && t > 0) read_dq(t - 1);
      // the accumulating MMAs of the previous pair read sP: they must have retired before the first overwrite
      ATT_CLK(c3);
      if ((c & 1) == 0 && pairs > 0) ptx::mbar_wait(bar_acc, (uint32_t)((pairs - 1) & 1));
      ATT_CLK(c4);
      w_acc += c4 - c3;
      if (mine) {
        uint8_t* prow = sP + (c & 1) * (TILE_ROWS * 128) + row * 128;
        uint8_t* srow = sDS + (size_t)c * (TILE_ROWS * 128) + row * 128;
#pragma unroll
        for (int q4 = 0; q4 < CW / 8; ++q4) {
          const int slot16 = ((quarter * (CW / 8) + q4) ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + slot16) = make_uint4(pkp[4 * q4], pkp[4 * q4 + 1], pkp[4 * q4 + 2], pkp[4 * q4 + 3]);
          *reinterpret_cast<uint4*>(srow + slot16) = make_uint4(pks[4 * q4], pks[4 * q4 + 1], pks[4 * q4 + 2], pks[4 * q4 + 3]);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bar_rd[set]);
      ATT_CLK(c5);
      w_wr += c5 - c4;
      if ((c & 1) || c == nc - 1) ++pairs;
      if (++c == nc) { c = 0; ++t; }
    }
    // Two key chunks: dV_0 / dK_0 have been final since the last q-tile's first pair retired (bar_acc, waited for in
    // round c = 2) -- store them while the last pair's MMAs and dQ drain.  Staging goes to sV, which only the score
    // MMAs read and those have all been consumed (sP / sDS / sQ / sG / sK are still being read).
    const bool early0 = nk == 2;
    if (early0 && storer) {
      const int j0 = (warp & 3) * 32;
      uint8_t* st2 = sV + warp * 2048;
      for (int c0 = 0; c0 < d; c0 += 32) {
        if (half == 0)
          store_acc_rows32_coalesced(trow + DV0 + (uint32_t)c0, 1.f, st2, p.dv + skv * p.dv_bs + (long long)j0 * p.dv_rs + h * d + c0,
                                     p.dv_rs, p.Lk - j0, lane, want_cs ? s_cs + 128 + c0 : nullptr);
        else
          store_acc_rows32_coalesced(trow + DK0 + (uint32_t)c0, p.scale, st2, p.dk + skv * p.dk_bs + (long long)j0 * p.dk_rs + h * d + c0,
                                     p.dk_rs, p.Lk - j0, lane, want_cs ? s_cs + 64 + c0 : nullptr);
      }
    }
    read_dq(nq - 1);                              // bar_dq of the last q-tile covers every MMA: dV / dK are final
    // Every MMA of this item has retired (bar_dq) and every warp is past its sV-staged stores (the bar.sync that ends
    // read_dq): Q / dO / K / V / sDS are dead -- the next item's tiles stream in under the remaining result stores, the
    // column-sum flush and the next item's prologue (measured on the one-item-per-CTA kernel: tile loads 4.8 K and
    // result stores 5 K of a 29 K-cycle CTA ran back to back).
    if (next_item < n_items) issue_loads(next_item, k_item + 1);
    for (int kc = early0 ? 1 : 0; kc < nk && storer; ++kc) {
      const int j0 = kc * TILE_ROWS + (warp & 3) * 32;   // first key row of this warp
      if (half == 0)
        store_acc_rows_coalesced(trow + DV0 + (uint32_t)(d * kc), d, 1.f, stage,
                                 p.dv + skv * p.dv_bs + (long long)j0 * p.dv_rs + h * d, p.dv_rs, p.Lk - j0, lane,
                                 want_cs ? s_cs + 128 : nullptr);
      else
        store_acc_rows_coalesced(trow + DK0 + (uint32_t)(d * kc), d, p.scale, stage,
                                 p.dk + skv * p.dk_bs + (long long)j0 * p.dk_rs + h * d, p.dk_rs, p.Lk - j0, lane,
                                 want_cs ? s_cs + 64 : nullptr);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();                               // every warp is done with this item's barriers, TMEM columns and s_cs
  if (want_cs && threadIdx.x < 3 * d) {          // one global atomic per column per item
    const int which = threadIdx.x / d, c = threadIdx.x - which * d;
    float* dst = which == 0 ? p.dq_cs : which == 1 ? p.dk_cs : p.dv_cs;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s_cs_all[w * 192 + which * 64 + c];
    atomicAdd(dst + h * d + c, v);
  }
  if (next_item < n_items) {
    if (threadIdx.x == 0) {                      // fresh phases for the next item's rounds (bar_ld keeps running)
      init_round_bars(true);
      ptx::fence_barrier_init();
    }
    __syncthreads();                             // s_cs flushed before it is zeroed again
  }
  }  // items
  if (p.dbg != nullptr && threadIdx.x == 0 && blockIdx.x == 0) {
    p.dbg[8] = w_sp; p.dbg[9] = w_acc; p.dbg[10] = w_dq; p.dbg[11] = w_el; p.dbg[12] = w_wr;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == NSW) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

template <typename K>
int set_smem_tc(K kernel, size_t bytes) {
  EGB_CHECK(bytes <= 227 * 1024, "attention_tc: needs %zu bytes of shared memory (> 227 KB)", bytes);
  EGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

long long* g_att_dbg = nullptr;

int fill_tc(const egb_attention_desc* d, AttTcParams* p) {
  memset(p, 0, sizeof(*p));
  p->q = (const bf16*)d->q; p->k = (const bf16*)d->k; p->v = (const bf16*)d->v; p->o = (const bf16*)d->o;
  p->out = (bf16*)d->o; p->d_o = (const bf16*)d->d_o;
  p->dq = (bf16*)d->dq; p->dk = (bf16*)d->dk; p->dv = (bf16*)d->dv;
  p->lse = d->lse; p->delta = d->delta; p->probs = d->probs;
  p->q_bs = d->q_bs; p->q_rs = d->q_rs; p->k_bs = d->k_bs; p->k_rs = d->k_rs; p->v_bs = d->v_bs; p->v_rs = d->v_rs;
  p->o_bs = d->o_bs; p->o_rs = d->o_rs; p->do_bs = d->do_bs; p->do_rs = d->do_rs;
  p->dq_bs = d->dq_bs; p->dq_rs = d->dq_rs; p->dk_bs = d->dk_bs; p->dk_rs = d->dk_rs; p->dv_bs = d->dv_bs; p->dv_rs = d->dv_rs;
  p->S = d->S; p->H = d->H; p->Lq = d->Lq; p->Lk = d->Lk; p->d = d->head_dim; p->kv_shift = d->kv_shift;
  p->Lq_pad = (d->Lq + 15) / 16 * 16;
  p->Lk_pad = (d->Lk + 15) / 16 * 16;
  p->scale = d->scale;
  if (d->dropout_p > 0.f) {
    p->drop_thresh = drop_threshold(d->dropout_p);
    p->drop_scale = 1.f / (1.f - d->dropout_p);
    p->seed = d->seed;
    p->epoch = egb_seed_epoch_ptr();
  }
  p->dbg = g_att_dbg;
  return 0;
}

bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

}  // namespace

// Debug aid: device buffer of 8 int64 receiving the clock64() phase stamps of one CTA of the next attention launches
// (0 start, 1 after TMEM alloc, 2 tiles staged, 3 first MMAs done, 4 softmax / dS written, 5 second MMAs done,
//  6 results stored, 7 TMEM released).  NULL disables.
extern "C" int egb_debug_attention_timing(long long* device_buf) {
  g_att_dbg = device_buf;
  return 0;
}

// The tensor-core path covers bf16 tensors with head_dim 32 or 64, sequence lengths <= 256 and 16-byte aligned
// rows; everything else (fp32 parity mode, other shapes) runs the CUDA-core kernels of attention.cu.
bool egb_attention_tc_supported(const egb_attention_desc* d, bool backward) {
  if (d->dtype != EGB_BF16) return false;
  if (d->head_dim != 32 && d->head_dim != 64) return false;
  if (d->Lq < 1 || d->Lk < 1 || d->Lq > 256 || d->Lk > 256) return false;
  const long long strides[] = {d->q_bs, d->q_rs, d->k_bs, d->k_rs, d->v_bs, d->v_rs, d->o_bs, d->o_rs};
  for (long long st : strides)
    if (st % 8 != 0) return false;
  if (!aligned16(d->q) || !aligned16(d->k) || !aligned16(d->v) || !aligned16(d->o)) return false;
  if (backward) {
    const long long bs[] = {d->do_bs, d->do_rs, d->dq_bs, d->dq_rs, d->dk_bs, d->dk_rs, d->dv_bs, d->dv_rs};
    for (long long st : bs)
      if (st % 8 != 0) return false;
    if (!aligned16(d->d_o) || !aligned16(d->dq) || !aligned16(d->dk) || !aligned16(d->dv)) return false;
  }
  return true;
}

int egb_attention_tc_fwd(const egb_attention_desc* d, cudaStream_t st) {
  AttTcParams p;
  if (fill_tc(d, &p)) return 1;
  const size_t smem = (size_t)(TILE_ROWS + 2 * p.Lk_pad) * 128 + 2 * TILE_ROWS * sizeof(float) + 1024 + 64;
  const bool drop = p.drop_thresh != 0u;
  if (set_smem_tc(att_tc_fwd_kernel<true>, smem) || set_smem_tc(att_tc_fwd_kernel<false>, smem)) return 1;
  dim3 grid((d->Lq + TILE_ROWS - 1) / TILE_ROWS, d->H, d->S);
  const bool prof = egb_prof_enabled() != 0;
  if (prof) egb_prof_begin(st, 4.0 * d->S * d->H * (double)d->Lq * d->Lk * d->head_dim,
                           2.0 * d->S * d->H * (double)d->head_dim * (2.0 * d->Lq + 2.0 * d->Lk), 2);
  static const int tma = getenv("EGB_ATT_TMA") ? atoi(getenv("EGB_ATT_TMA")) : 1;
  AttMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (tma && d->head_dim == 64) {
    const long long inner = (long long)d->H * d->head_dim;
    p.use_tma = !(egb_tmap_rows64(&maps.q, d->q, inner, d->Lq, d->S, d->q_rs, d->q_bs, TILE_ROWS) ||
                  egb_tmap_rows64(&maps.k, d->k, inner, d->Lk, d->S, d->k_rs, d->k_bs, p.Lk_pad) ||
                  egb_tmap_rows64(&maps.v, d->v, inner, d->Lk, d->S, d->v_rs, d->v_bs, p.Lk_pad));
  }
  if (drop) att_tc_fwd_kernel<true><<<grid, TC_THREADS, smem, st>>>(p, maps);
  else att_tc_fwd_kernel<false><<<grid, TC_THREADS, smem, st>>>(p, maps);
  if (prof) egb_prof_end(st);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_attention_tc_bwd(const egb_attention_desc* d, cudaStream_t st) {
  AttTcParams p;
  if (fill_tc(d, &p)) return 1;
  EGB_CHECK(d->lse && d->d_o && d->dq && d->dk && d->dv, "attention_bwd: missing buffers");
  const bool drop = p.drop_thresh != 0u;
  const bool prof = egb_prof_enabled() != 0;
  static const int fused = getenv("EGB_ATT_FUSED_BWD") ? atoi(getenv("EGB_ATT_FUSED_BWD")) : 2;   // 2 pipelined, 1 fused, 0 two kernels
  const int nk = (p.Lk_pad + 127) / 128;
  size_t smem_f = (size_t)(2 * p.Lq_pad + 2 * p.Lk_pad) * 128 + (size_t)(2 + 2 * nk) * TILE_ROWS * 128 + 1024 + 128;
  // the in-kernel column sums need 8 x 192 floats more; when they do not fit, the column-sum pass runs afterwards
  const bool cs_in_kernel = d->dq_colsum != nullptr && smem_f + 8 * 192 * 4 + 32 <= 227 * 1024;
  if (cs_in_kernel) smem_f += 8 * 192 * 4 + 32;
  if (fused && smem_f <= 227 * 1024) {
    if (set_smem_tc(att_tc_bwd_fused_kernel<true>, smem_f) || set_smem_tc(att_tc_bwd_fused_kernel<false>, smem_f)) return 1;
    dim3 grid(d->H, d->S);
    if (prof) egb_prof_begin(st, 10.0 * d->S * d->H * (double)d->Lq * d->Lk * d->head_dim,
                             2.0 * d->S * d->H * (double)d->head_dim * (4.0 * d->Lq + 4.0 * d->Lk), 3);
    if (fused >= 2) {
      static const int tma = getenv("EGB_ATT_TMA") ? atoi(getenv("EGB_ATT_TMA")) : 1;
      AttMaps maps;
      memset(&maps, 0, sizeof(maps));
      if (tma && d->head_dim == 64) {
        // a view the tensor-map encoder rejects (exotic strides) simply keeps the cp.async staging
        const long long inner = (long long)d->H * d->head_dim;
        p.use_tma = !(egb_tmap_rows64(&maps.q, d->q, inner, d->Lq, d->S, d->q_rs, d->q_bs, p.Lq_pad) ||
                      egb_tmap_rows64(&maps.g, d->d_o, inner, d->Lq, d->S, d->do_rs, d->do_bs, p.Lq_pad) ||
                      egb_tmap_rows64(&maps.o, d->o, inner, d->Lq, d->S, d->o_rs, d->o_bs, p.Lq_pad) ||
                      egb_tmap_rows64(&maps.k, d->k, inner, d->Lk, d->S, d->k_rs, d->k_bs, p.Lk_pad) ||
                      egb_tmap_rows64(&maps.v, d->v, inner, d->Lk, d->S, d->v_rs, d->v_bs, p.Lk_pad));
      }
      EGB_CHECK((d->dq_colsum == nullptr) == (d->dk_colsum == nullptr) && (d->dq_colsum == nullptr) == (d->dv_colsum == nullptr),
                "attention_bwd: pass all three column-sum buffers or none");
      if (cs_in_kernel) { p.dq_cs = d->dq_colsum; p.dk_cs = d->dk_colsum; p.dv_cs = d->dv_colsum; }
      // one persistent CTA per SM that prefetches its next (batch, head) item (EGB_ATT_PERSIST=0: one CTA per item)
      static const int persist = getenv("EGB_ATT_PERSIST") ? atoi(getenv("EGB_ATT_PERSIST")) : 1;
      // (head_dim 64 / TMA tiles only: 408 -> 389 us per ViT-B layer; with cp.async staging (head_dim 32) the prefetch
      //  competes with the result stores for the same LSU queue and the static item split costs more than it saves:
      //  483 -> 499 us per EEG layer)
      // eight softmax warps (32 key columns of a round per thread) or, EGB_ATT_NSW=16, sixteen (16 columns, 18-warp CTA,
      // 96 registers per thread).  Unlike the GEMM epilogue the backward does NOT profit from the extra warps: measured
      // 424 vs 388 us per ViT-B layer and 466 vs 450 us per EEG layer -- a round's critical path is the barrier / MMA
      // round trip, not the softmax arithmetic, and twice the threads arrive on every barrier.
      static const int nsw16 = getenv("EGB_ATT_NSW") ? (atoi(getenv("EGB_ATT_NSW")) == 16) : 0;
#define EGB_ATT_LAUNCH(KERNEL, GRID)                                                                          \
  do {                                                                                                        \
    if (nsw16) {                                                                                              \
      if (set_smem_tc(KERNEL<true, 16>, smem_f) || set_smem_tc(KERNEL<false, 16>, smem_f)) return 1;         \
      if (drop) KERNEL<true, 16><<<GRID, 16 * 32 + 64, smem_f, st>>>(p, maps);                                \
      else KERNEL<false, 16><<<GRID, 16 * 32 + 64, smem_f, st>>>(p, maps);                                    \
    } else {                                                                                                  \
      if (set_smem_tc(KERNEL<true, 8>, smem_f) || set_smem_tc(KERNEL<false, 8>, smem_f)) return 1;           \
      if (drop) KERNEL<true, 8><<<GRID, 8 * 32 + 64, smem_f, st>>>(p, maps);                                  \
      else KERNEL<false, 8><<<GRID, 8 * 32 + 64, smem_f, st>>>(p, maps);                                      \
    }                                                                                                         \
  } while (0)
      if (persist && p.use_tma) {
        const int items = d->H * d->S;
        const int ctas = items < egb_num_sms() ? items : egb_num_sms();
        EGB_ATT_LAUNCH(att_tc_bwd_pers_kernel, ctas);
      } else {
        EGB_ATT_LAUNCH(att_tc_bwd_pipe_kernel, grid);
      }
#undef EGB_ATT_LAUNCH
      if (prof) egb_prof_end(st);
      egb_count_launch(1);
      EGB_LAUNCH_CHECK();
      return cs_in_kernel ? 0 : egb_attention_colsum_pass(d, st);
    } else if (drop) att_tc_bwd_fused_kernel<true><<<grid, TC_THREADS, smem_f, st>>>(p);
    else att_tc_bwd_fused_kernel<false><<<grid, TC_THREADS, smem_f, st>>>(p);
    if (prof) egb_prof_end(st);
    egb_count_launch(1);
    EGB_LAUNCH_CHECK();
    return egb_attention_colsum_pass(d, st);
  }
  EGB_CHECK(d->delta != nullptr, "attention_bwd: missing delta scratch");
  const size_t smem_a = (size_t)(2 * TILE_ROWS + 2 * p.Lk_pad) * 128 + 1024 + 64;
  const size_t smem_b = (size_t)(2 * TILE_ROWS + 2 * p.Lq_pad) * 128 + (size_t)p.Lq_pad * 8 + 1024 + 64;
  if (set_smem_tc(att_tc_bwd_dq_kernel<true>, smem_a) || set_smem_tc(att_tc_bwd_dkv_kernel<true>, smem_b) ||
      set_smem_tc(att_tc_bwd_dq_kernel<false>, smem_a) || set_smem_tc(att_tc_bwd_dkv_kernel<false>, smem_b))
    return 1;
  dim3 grid_a((d->Lq + TILE_ROWS - 1) / TILE_ROWS, d->H, d->S);
  dim3 grid_b((d->Lk + TILE_ROWS - 1) / TILE_ROWS, d->H, d->S);
  if (prof) egb_prof_begin(st, 10.0 * d->S * d->H * (double)d->Lq * d->Lk * d->head_dim,
                           2.0 * d->S * d->H * (double)d->head_dim * (4.0 * d->Lq + 4.0 * d->Lk), 3);
  if (drop) {
    att_tc_bwd_dq_kernel<true><<<grid_a, TC_THREADS, smem_a, st>>>(p);
    att_tc_bwd_dkv_kernel<true><<<grid_b, TC_THREADS, smem_b, st>>>(p);
  } else {
    att_tc_bwd_dq_kernel<false><<<grid_a, TC_THREADS, smem_a, st>>>(p);
    att_tc_bwd_dkv_kernel<false><<<grid_b, TC_THREADS, smem_b, st>>>(p);
  }
  if (prof) egb_prof_end(st);
  egb_count_launch(2);
  EGB_LAUNCH_CHECK();
  return egb_attention_colsum_pass(d, st);
}
