// bf16 tensor-core GEMM for sm_100a: TMA (3-D tensor maps, 128-B swizzle) -> shared memory ->
// tcgen05.mma (fp32 accumulators in TMEM, double buffered) -> tcgen05.ld epilogue.
//
// One persistent CTA per SM, 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer (one lane) and
// TMEM owner, warps 2..5 = epilogue (one TMEM lane quadrant each).  Three mbarrier pipelines:
// smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
//
// Operands are 3-level strided views (see include/eyegaze_b200.h) in either K-major or MN-major
// form, so the same kernel runs y = x.W^T, dx = dy.W, dW = dy^T.x and the Conv1d/Conv2d layers
// of the reference as implicit GEMMs over overlapping-row views without an im2col buffer.
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "epilogue.cuh"

extern void egb_count_launch(int n);
int egb_prof_enabled();
void egb_prof_begin(cudaStream_t st, double flops, double bytes, int kind);
void egb_prof_end(cudaStream_t st);
void egb_prof_tag(double a, double b, double c, double d);
int egb_fill_epilogue(const egb_gemm_desc* d, EpiParams* e);

#ifndef EGB_LEAN_RES
#define EGB_LEAN_RES 0      // experiment switch: bias + residual epilogues on the lean 16-warp path as well
#endif

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int NUM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int EPI_WARPS = 8;
constexpr int STG_PITCH = 32;  // floats per staged epilogue row; 16-byte chunks are XOR-swizzled by (row & 7)
constexpr int STG_BYTES_PER_WARP = 32 * STG_PITCH * 4;

struct TcParams {
  int M, N, K;
  int a_major, b_major;
  int a_rpg, b_rpg;
  int a_seg, a_shift, b_seg, b_shift;  // inner segmentation (0 = off)
  int m_tiles, n_tiles, k_blocks, split_k, kb_per_split;
  EpiParams epi;
  long long* dbg;  // optional [gridDim.x][8] cycle counters (egb_debug_gemm_timing)
  int dbg_skip;    // experiment switch (EGB_GEMM_SKIPB=1): see the pair producer
  int bkt;         // K extent of one pipeline stage of the pair kernel (64 or 128)
  int row_epi;     // 1: row-layout epilogue with TMA stores (tmC / tmP are valid)
};

// cycles spent inside a barrier wait, accumulated into *acc when profiling is on
#define TIMED_WAIT(acc, stmt)            \
  do {                                   \
    if (p.dbg != nullptr) {              \
      const long long _t0 = clock64();   \
      stmt;                              \
      acc += clock64() - _t0;            \
    } else {                             \
      stmt;                              \
    }                                    \
  } while (0)

template <int BN, int BKT = 64>
struct TcConfig {
  static constexpr int A_BYTES = BM * BKT * 2;
  static constexpr int B_BYTES = BN * BKT * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BKT == 128) ? ((BN == 256) ? 2 : (BN == 128 ? 3 : 4)) : ((BN == 256) ? 4 : (BN == 128 ? 5 : 7));
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES_PER_WARP + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Issues the TMA loads of one operand tile: TILE rows of the M/N index x BK of the reduction index.
//   K-major : one {64 k, TILE rows} box;  MN-major: TILE/64 boxes of {64 mn, BK reduction rows}.
// `seg`/`shift` implement the inner-index segmentation of egb_operand (single-group operands).
template <int TILE>
__device__ __forceinline__ void load_operand_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int major,
                                                  int rpg, int seg, int shift, int tile, int kb) {
  if (major == 1 && rpg < 0) {
    // chunked MN-major map {64 mn, k rows, 64-wide chunks}: ONE box {64, BK, TILE/64} lands as TILE/64 consecutive
    // [BK rows][128 B] chunks -- exactly the MN-major UMMA layout -- with a single TMA instruction
    ptx::tma_load_3d(dst, tm, bar, 0, kb * BK, tile * (TILE / 64));
  } else if (major == 0) {
    int inner0 = kb * BK, row0 = tile * TILE;
    if (seg > 0) { row0 += (inner0 / seg) * shift; inner0 %= seg; }
    ptx::tma_load_3d(dst, tm, bar, inner0, row0 % rpg, row0 / rpg);
  } else {
    const int red0 = kb * BK;
#pragma unroll
    for (int j = 0; j < TILE / 64; ++j) {
      int inner0 = tile * TILE + j * 64, row0 = red0;
      if (seg > 0) { row0 += (inner0 / seg) * shift; inner0 %= seg; }
      ptx::tma_load_3d(dst + j * (BK * 128), tm, bar, inner0, row0 % rpg, row0 / rpg);
    }
  }
}


// same as load_operand_tile, for a CTA pair: the bytes are signalled on the leader CTA's barrier
template <int TILE>
__device__ __forceinline__ void load_operand_tile_2sm(uint8_t* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr,
                                                      int major, int rpg, int seg, int shift, int tile, int kb) {
  if (major == 1 && rpg < 0) {
    ptx::tma_load_3d_2sm(dst, tm, bar_cluster_addr, 0, kb * BK, tile * (TILE / 64));
  } else if (major == 0) {
    int inner0 = kb * BK, row0 = tile * TILE;
    if (seg > 0) { row0 += (inner0 / seg) * shift; inner0 %= seg; }
    ptx::tma_load_3d_2sm(dst, tm, bar_cluster_addr, inner0, row0 % rpg, row0 / rpg);
  } else {
    const int red0 = kb * BK;
#pragma unroll
    for (int j = 0; j < TILE / 64; ++j) {
      int inner0 = tile * TILE + j * 64, row0 = red0;
      if (seg > 0) { row0 += (inner0 / seg) * shift; inner0 %= seg; }
      ptx::tma_load_3d_2sm(dst + j * (BK * 128), tm, bar_cluster_addr, inner0, row0 % rpg, row0 / rpg);
    }
  }
}


// Epilogue of one accumulator tile for one warp: TMEM lane quadrant (32 rows m_base..m_base+31), columns
// [col0, col0 + ncol).  A tcgen05.ld gives every lane ONE ROW (32 consecutive columns); storing from that layout
// would scatter each warp-wide store over 32 different 128-byte lines.  Each 32x32 fp32 chunk is therefore
// transposed through a warp-private shared-memory tile so that a lane owns 8 consecutive columns of a row and
// 4 lanes cover a 64-byte row segment: every global access of the epilogue (output, pre-activation copy, residual,
// activation-backward operand, split-K atomics) becomes sector-exact and 4x denser per instruction.
// TMEM loads are double buffered: chunk i+1 is in flight while chunk i goes through the math and stores.


// One TMA instruction per operand tile and 128-wide K stage, through a "chunk map" {64 elements, rows, 64-wide chunks}
// (chunk stride 128 B).  K-major: rows = M/N index, chunks run along K (inside one segment for the conv views, whose
// segments start `shift` rows further down); MN-major: rows = K index, chunks run along M/N.
template <int TILE, int BKT>
__device__ __forceinline__ void load_stage_chunked(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int major, int seg,
                                                   int shift, int tile, int kb) {
  if (major == 0) {
    int inner0 = kb * BKT, row0 = tile * TILE;
    if (seg > 0) { row0 += (inner0 / seg) * shift; inner0 %= seg; }
    ptx::tma_load_3d(dst, tm, bar, 0, row0, inner0 >> 6);
  } else {
    int mn0 = tile * TILE, row0 = kb * BKT;
    if (seg > 0) { row0 += (mn0 / seg) * shift; mn0 %= seg; }
    ptx::tma_load_3d(dst, tm, bar, 0, row0, mn0 >> 6);
  }
}

// Global operands of one 32x32 chunk in the transposed (8-column run) layout: bias once, residual / act' operand per
// row run.  They are fetched a whole chunk AHEAD of their use -- chunk 0 even before the accumulator is ready --
// because a DRAM round trip per chunk was what bounded the residual / act' epilogues (K = 768: 12.5 us per tile
// against 5.5 us of MMA).
struct EpiChunkOps {
  float bias[8];
  EpiPre8 run[4];
};
template <int EF>
__device__ __forceinline__ void epilogue_prefetch(const EpiParams& epi, const EpiRow (&rows)[4], int n, int lane,
                                                  EpiChunkOps& ops) {
  const int cg = (lane & 3) * 8;
  if (EF != EF_GENERIC) {
    if ((EF & EF_BIAS) && n + cg < epi.N) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + n + cg));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + n + cg + 4));
      ops.bias[0] = b0.x; ops.bias[1] = b0.y; ops.bias[2] = b0.z; ops.bias[3] = b0.w;
      ops.bias[4] = b1.x; ops.bias[5] = b1.y; ops.bias[6] = b1.z; ops.bias[7] = b1.w;
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) epi_prefetch8<EF>(epi, rows[it], n + cg, ops.run[it]);
  }
}

template <int EF>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& epi, const EpiRow (&rows)[4], int m_base, int n,
                                               const uint32_t (&r)[32], float* stage, int lane, const EpiChunkOps& ops,
                                               unsigned long long seed_eff) {
  const int sub = lane >> 2, cg = (lane & 3) * 8;
  __syncwarp();                                                 // previous chunk's readers are done
  // Explicit shared-space accesses: through a generic pointer the compiler emits generic LD/ST and has to assume
  // that the global stores of one 8-column run alias the staged loads of the next, which serialises the runs.
  // All eight staged loads are issued before the math.
  const uint32_t sbase = ptx::smem_u32(stage);
  const uint32_t mine = sbase + (uint32_t)(lane * STG_PITCH * 4);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + (uint32_t)((i ^ (lane & 7)) << 4)), "r"(r[4 * i]),
                 "r"(r[4 * i + 1]), "r"(r[4 * i + 2]), "r"(r[4 * i + 3])
                 : "memory");
  __syncwarp();
  float v[4][8];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + sub;
    const uint32_t src = sbase + (uint32_t)(rr * STG_PITCH * 4);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[it][0]), "=f"(v[it][1]), "=f"(v[it][2]), "=f"(v[it][3])
                 : "r"(src + (uint32_t)((((lane & 3) * 2) ^ (rr & 7)) << 4))
                 : "memory");
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[it][4]), "=f"(v[it][5]), "=f"(v[it][6]), "=f"(v[it][7])
                 : "r"(src + (uint32_t)((((lane & 3) * 2 + 1) ^ (rr & 7)) << 4))
                 : "memory");
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if (EF == EF_GENERIC) epi_apply_store_row<8>(epi, rows[it], m_base + it * 8 + sub, n + cg, v[it]);
    else epi_fast8<EF>(epi, rows[it], m_base + it * 8 + sub, n + cg, v[it], ops.bias, ops.run[it], seed_eff);
  }
  if (EF != EF_GENERIC && (EF & EF_COLSUM)) {
    // column sums of the 32 x 32 chunk as stored: this lane's four rows, then the eight lanes that share its columns
    float cs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[i] = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (rows[it].ok && n + cg < epi.N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) cs[i] += v[it][i];
      }
    }
    // butterfly over the eight lanes that share these columns, halving the columns a lane keeps in every round
    // (4 + 2 + 1 shuffles instead of 3 x 8): lane (sub, q) ends up with the sum of column cg + sub
    float a4[4], b2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float recv = __shfl_xor_sync(0xffffffffu, (sub & 4) ? cs[i] : cs[i + 4], 16);
      a4[i] = ((sub & 4) ? cs[i + 4] : cs[i]) + recv;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float recv = __shfl_xor_sync(0xffffffffu, (sub & 2) ? a4[i] : a4[i + 2], 8);
      b2[i] = ((sub & 2) ? a4[i + 2] : a4[i]) + recv;
    }
    const float recv = __shfl_xor_sync(0xffffffffu, (sub & 1) ? b2[0] : b2[1], 4);
    const float tot = ((sub & 1) ? b2[1] : b2[0]) + recv;
    if (n + cg < epi.N) atomicAdd(epi.colsum + n + cg + sub, tot);   // result unused: compiles to RED
  }
}

// L2 prefetch of the row-layout operand (residual / activation-backward factor) of a tile this CTA will drain LATER:
// ncu put 27 % of the dX-through-GELU' kernel's stall samples on the first use of those loads -- fetched one 32 x 32 chunk
// ahead (~500 cycles) they still arrive from DRAM.  Issued a whole tile ahead they are L2 hits by the time the epilogue
// asks for them.  `t` = index of the calling thread among the CTA's 256 epilogue threads; rows [m0, m0 + 128), columns
// [n0, n0 + bn).
template <int EF>
__device__ __forceinline__ void epilogue_prefetch_l2(const EpiParams& epi, int m0, int n0, int bn, int t) {
  // (activation-backward operands only: a residual is the layer's own input, still L2-resident from the forward read --
  //  prefetching it cost the proj + residual GEMMs 8-13 %)
  if ((EF & (EF_ABWD_RELU | EF_ABWD_GELU | EF_ABWD_MUL)) == 0 || EF == EF_GENERIC) return;
  const EpiMat& mat = epi.aux;
  int cols = epi.N - n0;
  if (cols > bn) cols = bn;
  if (cols <= 0) return;
  const int lines_per_row = (cols * 2 + 127) >> 7;
  for (int idx = t; idx < 128 * lines_per_row; idx += 256) {
    const int r = idx / lines_per_row, l = idx - r * lines_per_row;
    const int m = m0 + r;
    if (m < epi.M) {
      const char* ptr = mat.ptr + (epi_row_offset(mat, m) + n0) * 2 + (long long)l * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    }
  }
}

// before the accumulator is ready: row bookkeeping and the first chunk's operands
template <int EF>
__device__ __forceinline__ void epilogue_tile_begin(const EpiParams& epi, int m_base, int n0, int col0, int lane,
                                                    EpiRow (&rows)[4], EpiChunkOps& ops0) {
#pragma unroll
  for (int it = 0; it < 4; ++it) rows[it] = epi_row_setup(epi, m_base + it * 8 + (lane >> 2));
  epilogue_prefetch<EF>(epi, rows, n0 + col0, lane, ops0);
}

template <int EF>
__device__ __forceinline__ void epilogue_tile(const EpiParams& epi, uint32_t taddr, int m_base, int n0, int col0, int ncol,
                                              float* stage, int lane, const EpiRow (&rows)[4], EpiChunkOps& opsA) {
  // One accumulator chunk in registers at a time (a second one pushed the residual / act' variants over the 168
  // registers a 10-warp CTA allows and the row offsets into local memory); the operand prefetch of the next chunk
  // is what sits between the TMEM load and its wait.
  uint32_t ra[32];
  EpiChunkOps opsB;
  const unsigned long long seed_eff = epi.drop_thresh != 0u ? egb_mix_seed(epi.seed, epi.epoch) : 0ull;
#pragma unroll 1
  for (int c = 0; c < ncol; c += 64) {
    ptx::tmem_ld32(taddr + (uint32_t)(col0 + c), ra);
    if (c + 32 < ncol) epilogue_prefetch<EF>(epi, rows, n0 + col0 + c + 32, lane, opsB);
    ptx::tmem_ld_wait();
    epilogue_chunk<EF>(epi, rows, m_base, n0 + col0 + c, ra, stage, lane, opsA, seed_eff);
    if (c + 32 < ncol) {
      ptx::tmem_ld32(taddr + (uint32_t)(col0 + c + 32), ra);
      if (c + 64 < ncol) epilogue_prefetch<EF>(epi, rows, n0 + col0 + c + 64, lane, opsA);
      ptx::tmem_ld_wait();
      epilogue_chunk<EF>(epi, rows, m_base, n0 + col0 + c + 32, ra, stage, lane, opsB, seed_eff);
    }
  }
}

// ================================================================================================
// Row-layout epilogue with TMA stores.
// A tcgen05.ld hands every lane ONE ROW of the accumulator (32 consecutive columns).  Instead of transposing each
// 32 x 32 chunk through shared memory so that the lanes can issue coalesced global stores themselves (8 STS + 8 LDS +
// 4 predicated STG per chunk and a dependent STS -> LDS round trip: with K = 256 the drain of a tile took three times
// as long as its MMAs), the lane converts its row segment to bf16, writes the 64 bytes into a warp-private
// [32 rows][64 B] tile in the 64-byte-swizzled layout (conflict-free 16-byte stores) and ONE lane hands the tile to the
// TMA unit (cp.async.bulk.tensor store, box {32 columns, 32 rows}).  The TMA unit clips rows >= M / columns >= N, so
// the store path has no predicates; the two 2 KB tiles of a warp alternate, guarded by cp.async.bulk.wait_group.read.
// Row-layout operands (residual, activation-backward factor) are read by the lane straight from its own row
// (4 x 16 B, prefetched a chunk ahead); the bias is a warp-uniform broadcast load.  The accumulator is handed back to
// the MMA warp as soon as the LAST tcgen05.ld of the tile has completed, before that chunk's arithmetic and stores.
// Column sums (EF_COLSUM) are taken with a halving butterfly over the 32 rows (31 shuffles per chunk).
// ================================================================================================
struct RowOps {
  uint4 x[4];      // residual or activation-backward operand of this lane's row, 32 columns
};

template <int EF>
__device__ __forceinline__ void row_prefetch(const bf16* row_ptr, int n, int N, RowOps& o) {
  if (EF & (EF_RES | EF_ABWD_RELU | EF_ABWD_GELU | EF_ABWD_MUL)) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (n + 8 * g < N) o.x[g] = __ldg(reinterpret_cast<const uint4*>(row_ptr + n + 8 * g));
  }
}

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// one 32 x 32 chunk: this lane's 32 accumulator values r of row m, columns [n, n + 32)
template <int EF>
__device__ __forceinline__ void row_chunk(const EpiParams& epi, const CUtensorMap* tmC, const CUtensorMap* tmP,
                                          unsigned long long seed_eff, unsigned long long row_base, bool row_ok, int m_warp,
                                          int n, const uint32_t (&r)[32], const RowOps& o, uint8_t* stage, int lane,
                                          uint32_t& nstores, int wide = -1) {
  constexpr bool TWO = (EF & (EF_PRE | EF_DGELU)) != 0;      // second output (pre-activation / gelu')
  // WIDE mode (single-output variants, `wide` = 0 / 1: first / second 32 columns of a 64-column box): the warp's whole
  // 4 KB staging area is ONE [32 rows][128 B] tile in the 128-byte-swizzled layout and one TMA store covers 64 columns --
  // the TMA unit's cost is per instruction, not per byte (32 stores of 2 KB per tile held the K = 256 GEMMs at ~8 K
  // cycles per tile against 2.8 K of MMA).
  // the tile(s) this chunk writes must have been read out by the TMA unit: with one output the two 2 KB tiles
  // alternate (one store may stay in flight), with two outputs -- or the single wide tile -- everything is rewritten
  if (wide != 1) {
    if (lane == 0) {
      if (TWO || wide == 0 || wide == -2) ptx::bulk_wait_read<0>(); else ptx::bulk_wait_read<1>();   // (-2: a narrow chunk after wide ones)
    }
    __syncwarp();
  }
  uint8_t* bufC = stage + ((TWO || wide >= 0) ? 0 : (nstores & 1u) * 2048u);
  uint8_t* bufP = stage + 2048;
  const uint32_t rowC = ptx::smem_u32(bufC) + (uint32_t)(lane * (wide >= 0 ? 128 : 64));
  const uint32_t rowP = ptx::smem_u32(bufP) + (uint32_t)(lane * 64);
  const uint32_t sw = wide >= 0 ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
  const uint32_t goff = wide == 1 ? 4u : 0u;
  float cs[(EF & EF_COLSUM) ? 32 : 1];
  if (EF == (EF_BIAS | EF_GELU | EF_PRE | EF_DGELU)) {
    // fc1 + GELU + saved GELU': 16 columns (8 packed pairs) at a time through the phase-major evaluation
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int nh = n + 16 * hh;
      float2 xv[8], yv[8], dv2[8];
      if (nh < epi.N) {                                   // N % 16 == 0 on this path (checked by the host)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(epi.bias + nh + 4 * q));
          xv[2 * q] = add2(make_float2(__uint_as_float(r[16 * hh + 4 * q]), __uint_as_float(r[16 * hh + 4 * q + 1])), make_float2(b.x, b.y));
          xv[2 * q + 1] = add2(make_float2(__uint_as_float(r[16 * hh + 4 * q + 2]), __uint_as_float(r[16 * hh + 4 * q + 3])), make_float2(b.z, b.w));
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) xv[q] = make_float2(0.f, 0.f);
      }
      gelu_erf_both2n<8>(xv, yv, dv2);
      if (epi.drop_thresh != 0u) {
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) { v[2 * i] = yv[4 * gg + i].x; v[2 * i + 1] = yv[4 * gg + i].y; }
          drop_apply_run8(seed_eff, row_base + (unsigned long long)(nh + 8 * gg), epi.drop_thresh, epi.drop_scale, v);
#pragma unroll
          for (int i = 0; i < 4; ++i) yv[4 * gg + i] = make_float2(v[2 * i], v[2 * i + 1]);
        }
      }
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        const uint32_t g = (uint32_t)(2 * hh + gg);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowC + ((g ^ sw) << 4)),
                     "r"(pack2_bf16(yv[4 * gg].x, yv[4 * gg].y)), "r"(pack2_bf16(yv[4 * gg + 1].x, yv[4 * gg + 1].y)),
                     "r"(pack2_bf16(yv[4 * gg + 2].x, yv[4 * gg + 2].y)), "r"(pack2_bf16(yv[4 * gg + 3].x, yv[4 * gg + 3].y))
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowP + ((g ^ sw) << 4)),
                     "r"(pack2_bf16(dv2[4 * gg].x, dv2[4 * gg].y)), "r"(pack2_bf16(dv2[4 * gg + 1].x, dv2[4 * gg + 1].y)),
                     "r"(pack2_bf16(dv2[4 * gg + 2].x, dv2[4 * gg + 2].y)), "r"(pack2_bf16(dv2[4 * gg + 3].x, dv2[4 * gg + 3].y))
                     : "memory");
      }
    }
  } else
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float v[8], dv[8], bias[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * g + i]);
    if ((EF & EF_BIAS) && n + 8 * g < epi.N) {   // warp-uniform address: a broadcast load
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + n + 8 * g));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + n + 8 * g + 4));
      bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
      bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
    }
    EpiPre8 pre;
    pre.res = o.x[g];
    pre.aux = o.x[g];
    epi_math8<EF>(epi, seed_eff, row_base, n + 8 * g, v, bias, pre, dv);
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[i] = pack2_bf16(v[2 * i], v[2 * i + 1]);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowC + ((((uint32_t)g + goff) ^ sw) << 4)), "r"(pk[0]), "r"(pk[1]),
                 "r"(pk[2]), "r"(pk[3])
                 : "memory");
    if (TWO) {
      uint32_t pd[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pd[i] = pack2_bf16(dv[2 * i], dv[2 * i + 1]);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowP + (((uint32_t)g ^ sw) << 4)), "r"(pd[0]), "r"(pd[1]),
                   "r"(pd[2]), "r"(pd[3])
                   : "memory");
    }
    if (EF & EF_COLSUM) {
      // column sums of the values AS STORED (rounded to bf16), rows past M and columns past N contribute nothing
      const bool ok = row_ok && n + 8 * g < epi.N;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&pk[i]);
        cs[8 * g + 2 * i] = ok ? __low2float(h) : 0.f;
        cs[8 * g + 2 * i + 1] = ok ? __high2float(h) : 0.f;
      }
    }
  }
  if (wide != 0) {
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (wide == 1) ptx::tma_store_2d(tmP, bufC, n - 32, m_warp);     // tmP = the 64-column store map of C here
      else ptx::tma_store_2d(tmC, bufC, n, m_warp);
      if (TWO) ptx::tma_store_2d(tmP, bufP, n, m_warp);
      ptx::bulk_commit();
    }
    ++nstores;
  }
  if (EF & EF_COLSUM) {
    // halving butterfly over the warp's 32 rows: after the round with lane bit b a lane keeps the half of its columns
    // selected by that bit, summed with its partner's; lane l ends up with the sum of column l
    float a16[16], a8[8], a4[4], a2[2];
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float recv = __shfl_xor_sync(0xffffffffu, b16 ? cs[i] : cs[i + 16], 16);
      a16[i] = (b16 ? cs[i + 16] : cs[i]) + recv;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float recv = __shfl_xor_sync(0xffffffffu, b8 ? a16[i] : a16[i + 8], 8);
      a8[i] = (b8 ? a16[i + 8] : a16[i]) + recv;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float recv = __shfl_xor_sync(0xffffffffu, b4 ? a8[i] : a8[i + 4], 4);
      a4[i] = (b4 ? a8[i + 4] : a8[i]) + recv;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float recv = __shfl_xor_sync(0xffffffffu, b2 ? a4[i] : a4[i + 2], 2);
      a2[i] = (b2 ? a4[i + 2] : a4[i]) + recv;
    }
    const float recv = __shfl_xor_sync(0xffffffffu, b1 ? a2[0] : a2[1], 1);
    const float tot = (b1 ? a2[1] : a2[0]) + recv;
    if (n + lane < epi.N) atomicAdd(epi.colsum + n + lane, tot);   // result unused: compiles to RED
  }
}

// One accumulator tile of one warp: TMEM lane quadrant at taddr (column 0 of the warp's share), rows m_warp + lane,
// global columns [n_base, n_base + ncol).  `release` hands the accumulator back to the MMA warp.
template <int EF, typename Release>
__device__ __forceinline__ void epilogue_tile_row(const EpiParams& epi, const CUtensorMap* tmC, const CUtensorMap* tmP,
                                                  uint32_t taddr, int m_warp, int n_base, int ncol, uint8_t* stage,
                                                  int lane, uint32_t& nstores, bool wide_ok, Release release) {
  const int m = m_warp + lane;
  const bool row_ok = m < epi.M;
  int nvalid = epi.N - n_base;
  if (nvalid > ncol) nvalid = ncol;
  const int nch = m_warp < epi.M ? (nvalid + 31) / 32 : 0;      // warp-uniform
  if (nch <= 0) { release(); return; }
  const bf16* row_ptr = nullptr;
  if (EF & EF_RES) row_ptr = reinterpret_cast<const bf16*>(epi.res.ptr) + epi_row_offset(epi.res, row_ok ? m : 0);
  if (EF & (EF_ABWD_RELU | EF_ABWD_GELU | EF_ABWD_MUL))
    row_ptr = reinterpret_cast<const bf16*>(epi.aux.ptr) + epi_row_offset(epi.aux, row_ok ? m : 0);
  uint32_t ra[32], rb[32];
  RowOps oa, ob;
  const unsigned long long seed_eff = epi.drop_thresh != 0u ? egb_mix_seed(epi.seed, epi.epoch) : 0ull;
  const unsigned long long row_base = (unsigned long long)m * (unsigned long long)epi.N;
  ptx::tmem_ld32(taddr, ra);
  row_prefetch<EF>(row_ptr, n_base, epi.N, oa);
  // 64-column boxes when this warp's share is a whole number of them (tmP then carries the wide map of C)
  const bool wide = (EF & (EF_PRE | EF_DGELU)) == 0 && wide_ok && (ncol & 63) == 0;
#pragma unroll 1
  for (int c = 0; c < nch; c += 2) {
    ptx::tmem_ld_wait();
    if (c + 1 < nch) {
      ptx::tmem_ld32(taddr + (uint32_t)(32 * (c + 1)), rb);
      row_prefetch<EF>(row_ptr, n_base + 32 * (c + 1), epi.N, ob);
    } else {
      release();
    }
    const bool pair = wide && c + 1 < nch;
    row_chunk<EF>(epi, tmC, tmP, seed_eff, row_base, row_ok, m_warp, n_base + 32 * c, ra, oa, stage, lane, nstores, pair ? 0 : (wide ? -2 : -1));
    if (c + 1 < nch) {
      ptx::tmem_ld_wait();
      if (c + 2 < nch) {
        ptx::tmem_ld32(taddr + (uint32_t)(32 * (c + 2)), ra);
        row_prefetch<EF>(row_ptr, n_base + 32 * (c + 2), epi.N, oa);
      } else {
        release();
      }
      row_chunk<EF>(epi, tmC, tmP, seed_eff, row_base, row_ok, m_warp, n_base + 32 * (c + 1), rb, ob, stage, lane, nstores,
                    pair ? 1 : (wide ? -2 : -1));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LEAN form of the row-layout epilogue for CTAs with SIXTEEN epilogue warps (four per TMEM lane quadrant, 64 columns of
// a 256-wide tile each).  ncu on the fc1 + GELU and the EEG ffn1 GEMMs: the two epilogue warps a scheduler gets in the
// 10-warp CTA issue one instruction per ~4.5 cycles (fixed-latency dependency stalls 28-35 %, issue slots 40-45 % busy,
// tensor pipe 23-52 %) whatever the instruction-level parallelism of the source -- the drain needs more warps, and 18
// warps leave 112 registers per thread.  Hence: one accumulator chunk in registers, ONE 2 KB staging tile per warp (the
// second output of the GELU variant goes through the same tile after the first store has been read out), bias fetched
// per 8-column run.
// ------------------------------------------------------------------------------------------------
// Activation-backward operand of this lane's row, 32 columns (64 bytes) per chunk.  The warp's share is 64 columns = ONE
// 128-byte line per row: chunk 0's loads are issued before the accumulator is waited for (DRAM latency under the MMAs),
// chunk 1's right after chunk 0's accumulator values have arrived -- they hit the line chunk 0 brought into L1.
template <int EF>
__device__ __forceinline__ void row_lean_load(const EpiParams& epi, int m_warp, int n, int lane, RowOps& o) {
  if (EF & (EF_RES | EF_ABWD_RELU | EF_ABWD_GELU | EF_ABWD_MUL)) {
    const int m = m_warp + lane;
    const EpiMat& mat = (EF & EF_RES) ? epi.res : epi.aux;
    const bf16* row = reinterpret_cast<const bf16*>(mat.ptr) + epi_row_offset(mat, m < epi.M ? m : 0) + n;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (n + 8 * g < epi.N) o.x[g] = __ldg(reinterpret_cast<const uint4*>(row + 8 * g));
  }
}

template <int EF, int CHUNK>
__device__ __forceinline__ void row_chunk_lean(const EpiParams& epi, const CUtensorMap* tmC, const CUtensorMap* tmP,
                                               unsigned long long seed_eff, unsigned long long row_base, int m_warp, int n,
                                               const uint32_t (&r)[32], uint8_t* stage, int lane, const RowOps& ops,
                                               bool row_ok) {
  constexpr bool TWO = (EF & (EF_PRE | EF_DGELU)) != 0;
  const uint32_t rowC = ptx::smem_u32(stage) + (uint32_t)(lane * 64);
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);
  uint32_t pd[TWO ? 16 : 1];
  if (lane == 0) ptx::bulk_wait_read<0>();
  __syncwarp();
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float v[8], dv[8], bias[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * g + i]);
    if ((EF & EF_BIAS) && n + 8 * g < epi.N) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + n + 8 * g));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + n + 8 * g + 4));
      bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
      bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
    }
    EpiPre8 pre;
    pre.res = (EF & (EF_RES | EF_ABWD_RELU | EF_ABWD_GELU | EF_ABWD_MUL)) ? ops.x[g] : make_uint4(0u, 0u, 0u, 0u);
    pre.aux = pre.res;
    epi_math8<EF>(epi, seed_eff, row_base, n + 8 * g, v, bias, pre, dv);
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[i] = pack2_bf16(v[2 * i], v[2 * i + 1]);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowC + (((uint32_t)g ^ sw) << 4)), "r"(pk[0]), "r"(pk[1]),
                 "r"(pk[2]), "r"(pk[3])
                 : "memory");
    if (TWO) {
#pragma unroll
      for (int i = 0; i < 4; ++i) pd[4 * g + i] = pack2_bf16(dv[2 * i], dv[2 * i + 1]);
    }
  }
  if (EF & EF_COLSUM) {
    // Column sums of the tile AS STORED, read back from the staging tile: lane l walks column l down the 32 rows (one
    // 2-byte load per row; a row's 32 columns are one conflict-free 64-byte access).  Rows past M hold exact zeros (their
    // A rows are zero-filled by TMA), so no row mask is needed.  The register butterfly of the 8-warp build needs ~100
    // live registers; this CTA has 96 per thread.
    __syncwarp();
    const uint32_t base = ptx::smem_u32(stage) + (uint32_t)((lane & 7) * 2);
    float tot = 0.f;
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      unsigned short h;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(base + (uint32_t)(rr * 64) + ((((uint32_t)lane >> 3) ^ (((uint32_t)rr >> 1) & 3u)) << 4)));
      tot += __uint_as_float((uint32_t)h << 16);
    }
    if (n + lane < epi.N) atomicAdd(epi.colsum + n + lane, tot);   // result unused: compiles to RED
  }
  ptx::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    ptx::tma_store_2d(tmC, stage, n, m_warp);
    ptx::bulk_commit();
  }
  if (TWO) {
    if (lane == 0) ptx::bulk_wait_read<0>();
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowC + (((uint32_t)g ^ sw) << 4)), "r"(pd[4 * g]),
                   "r"(pd[4 * g + 1]), "r"(pd[4 * g + 2]), "r"(pd[4 * g + 3])
                   : "memory");
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_2d(tmP, stage, n, m_warp);
      ptx::bulk_commit();
    }
  }
}

template <int EF, typename Release>
__device__ __forceinline__ void epilogue_tile_row_lean(const EpiParams& epi, const CUtensorMap* tmC, const CUtensorMap* tmP,
                                                       uint32_t taddr, int m_warp, int n_base, int ncol, uint8_t* stage,
                                                       int lane, const RowOps& ops0, Release release) {
  int nvalid = epi.N - n_base;
  if (nvalid > ncol) nvalid = ncol;
  const int nch = m_warp < epi.M ? (nvalid + 31) / 32 : 0;      // warp-uniform
  if (nch <= 0) { release(); return; }
  const unsigned long long seed_eff = epi.drop_thresh != 0u ? egb_mix_seed(epi.seed, epi.epoch) : 0ull;
  const unsigned long long row_base = (unsigned long long)(m_warp + lane) * (unsigned long long)epi.N;
  // the warp's share is 64 columns = at most two chunks (compile-time chunk index: the operand registers are indexed by it)
  const bool row_ok = m_warp + lane < epi.M;
  RowOps ops1;
  {
    uint32_t ra[32];
    ptx::tmem_ld32(taddr, ra);
    ptx::tmem_ld_wait();
    if (nch == 1) release();
    else row_lean_load<EF>(epi, m_warp, n_base + 32, lane, ops1);
    row_chunk_lean<EF, 0>(epi, tmC, tmP, seed_eff, row_base, m_warp, n_base, ra, stage, lane, ops0, row_ok);
  }
  if (nch > 1) {
    uint32_t ra[32];
    ptx::tmem_ld32(taddr + 32u, ra);
    ptx::tmem_ld_wait();
    release();
    row_chunk_lean<EF, 1>(epi, tmC, tmP, seed_eff, row_base, m_warp, n_base + 32, ra, stage, lane, ops1, row_ok);
  }
}

// Epilogue variants that run the row-layout path: the ones WITHOUT a row-layout operand.  (Measured with the residual /
// activation-backward operand read by each lane from its own row -- 32 different 128-byte lines per load instruction --
// those variants lost 25-80 %: proj + residual 68 -> 121 us, dX-through-GELU' 297 -> 409 us; they keep the transposed
// epilogue, whose 8-column runs make those reads sector-exact.)
template <int EF>
struct RowEpi {
  static constexpr bool ok = EF != EF_GENERIC &&
                             (EF & (EF_ACC | EF_RES | EF_ABWD_RELU | EF_ABWD_GELU | EF_ABWD_MUL | EF_COLSUM)) == 0;
};

template <int BN, int EF, int BKT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP, const TcParams p) {
  using Cfg = TcConfig<BN, BKT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* stage_all = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + EPI_WARPS * STG_BYTES_PER_WARP);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long t_wait0 = 0, t_wait1 = 0;
  const long long t_start = p.dbg != nullptr ? clock64() : 0;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles * p.split_k;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ks = tile % p.split_k;
        const int mn = tile / p.split_k;
        const int nt = mn % p.n_tiles;
        const int mt = mn / p.n_tiles;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          TIMED_WAIT(t_wait0, ptx::mbar_wait(&empty_bar[stage], phase ^ 1u));
          if (p.dbg_skip == 2) {   // EXPERIMENT: no loads at all
            ptx::mbar_arrive(&full_bar[stage]);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
          ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = tiles + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (BKT == 128) {
            load_stage_chunked<BM, BKT>(sa, &tmA, &full_bar[stage], p.a_major, p.a_seg, p.a_shift, mt, kb);
            load_stage_chunked<BN, BKT>(sb, &tmB, &full_bar[stage], p.b_major, p.b_seg, p.b_shift, nt, kb);
          } else {
            load_operand_tile<BM>(sa, &tmA, &full_bar[stage], p.a_major, p.a_rpg, p.a_seg, p.a_shift, mt, kb);
            load_operand_tile<BN>(sb, &tmB, &full_bar[stage], p.b_major, p.b_rpg, p.b_seg, p.b_shift, nt, kb);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = ptx::make_idesc_bf16(BM, BN, p.a_major, p.b_major);
    // per-UMMA_K (16 elements) descriptor advance, in 16-byte units
    const uint32_t a_adv = p.a_major == 0 ? 2u : (16u * 128u) >> 4;
    const uint32_t b_adv = p.b_major == 0 ? 2u : (16u * 128u) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ks = tile % p.split_k;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(tiles + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = ptx::make_smem_desc(sa, p.a_major == 0 ? 16u : (uint32_t)(BKT * 128), 1024u);
          const uint64_t bdesc = ptx::make_smem_desc(sb, p.b_major == 0 ? 16u : (uint32_t)(BKT * 128), 1024u);
#pragma unroll
          for (int k = 0; k < BKT / 16; ++k) {
            const uint32_t ao = p.a_major == 0 ? (uint32_t)((k >> 2) * (BM * 128 / 16) + (k & 3) * 2) : (uint32_t)(k * a_adv);
            const uint32_t bo = p.b_major == 0 ? (uint32_t)((k >> 2) * (BN * 128 / 16) + (k & 3) * 2) : (uint32_t)(k * b_adv);
            ptx::umma_bf16(tmem_d, adesc + (uint64_t)ao, bdesc + (uint64_t)bo, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (kb == kb1 - 1) ptx::umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int chalf = (warp - 2) >> 2;     // which half of the tile's columns this warp drains
    constexpr int NCOL = BN >= 64 ? BN / 2 : BN;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t nstores = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mn = tile / p.split_k;
      const int nt = mn % p.n_tiles;
      const int mt = mn / p.n_tiles;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      if (RowEpi<EF>::ok && p.row_epi) {
        TIMED_WAIT(t_wait0, ptx::mbar_wait(&tfull_bar[acc], acc_phase));
        ptx::tc_fence_after();
        uint64_t* tb = &tempty_bar[acc];
        epilogue_tile_row<EF>(p.epi, &tmC, &tmP, taddr + (uint32_t)(chalf * NCOL), mt * BM + quad * 32, nt * BN + chalf * NCOL,
                              NCOL, reinterpret_cast<uint8_t*>(stage_all + (warp - 2) * (32 * STG_PITCH)), lane, nstores,
                              p.row_epi == 2, [&]() {
                                ptx::tc_fence_before();
                                __syncwarp();
                                if (lane == 0) ptx::mbar_arrive_relaxed(tb);
                              });
      } else {
        {
          const int tn = tile + gridDim.x;              // the next tile of this CTA
          if (tn < total_tiles) {
            const int mn2 = tn / p.split_k;
            epilogue_prefetch_l2<EF>(p.epi, (mn2 / p.n_tiles) * BM, (mn2 % p.n_tiles) * BN, BN, (warp - 2) * 32 + lane);
          }
        }
        EpiRow rows[4];
        EpiChunkOps ops0;
        epilogue_tile_begin<EF>(p.epi, mt * BM + quad * 32, nt * BN, chalf * NCOL, lane, rows, ops0);
        TIMED_WAIT(t_wait0, ptx::mbar_wait(&tfull_bar[acc], acc_phase));
        ptx::tc_fence_after();
        epilogue_tile<EF>(p.epi, taddr, mt * BM + quad * 32, nt * BN, chalf * NCOL, NCOL,
                      stage_all + (warp - 2) * (32 * STG_PITCH), lane, rows, ops0);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_relaxed(&tempty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) ptx::bulk_wait<0>();     // this warp's TMA stores have completed before the CTA's smem goes away
  }

  if (p.dbg != nullptr && lane == 0 && warp < 3) {  // 0: producer (empty wait) 1: MMA (full, tempty) 2: epilogue (tfull)
    long long* o = p.dbg + (long long)blockIdx.x * 8;
    if (warp == 0) o[0] = t_wait0;
    if (warp == 1) { o[1] = t_wait0; o[2] = t_wait1; o[4] = clock64() - t_start; }
    if (warp == 2) o[3] = t_wait0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}


// ================================================================================================
// CTA-pair variant (cta_group::2): one 256 x BN output tile per pair of CTAs on neighbouring SMs.  Each CTA stages
// its own 128 rows of A and only HALF of the B tile (BN/2 rows); the pair's tensor cores read both halves, so the
// L2 -> SM operand traffic per FLOP drops by a third against the single-CTA 128 x 256 tile (32 KB instead of 48 KB
// per 128x256x64 step) -- the single-CTA kernel is bound by exactly that traffic on the big projection GEMMs.
// Only the leader CTA (cluster rank 0) issues MMAs; TMA completions of both CTAs land on the leader's "full"
// barrier; MMA completion is multicast to both CTAs' "empty" / "accumulator full" barriers; the epilogue warps of
// both CTAs release the accumulator on the leader's barrier.
// ================================================================================================
template <int BN, int BKT>
struct Tc2Config {
  static constexpr int A_BYTES = BM * BKT * 2;
  static constexpr int B_BYTES = (BN / 2) * BKT * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BKT == 128) ? ((BN == 256) ? 3 : 4) : ((BN == 256) ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES_PER_WARP + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Operand tile of one pipeline stage for the pair kernel.  BKT = 128 ("super-stages") is used when both operands have
// chunked tensor maps: ONE TMA instruction then brings a whole [TILE x 128] tile (two 64-wide K chunks for a K-major
// operand, TILE/64 MN chunks of 128 k-rows for an MN-major one).  The per-instruction cost of TMA, not bytes, was what
// starved the MMA warp (measured: MN-major dW GEMMs 760 -> 1218 TFLOP/s when 4 boxes per stage became 2).
template <int TILE, int BKT>
__device__ __forceinline__ void load_stage_operand_2sm(uint8_t* dst, const CUtensorMap* tm, uint32_t bar, int major, int rpg,
                                                       int seg, int shift, int tile, int kb) {
  if (BKT == 64) {
    load_operand_tile_2sm<TILE>(dst, tm, bar, major, rpg, seg, shift, tile, kb);
  } else if (major == 0) {   // chunked K-major map {64 k, rows, K/64 chunks}: box {64, TILE, 2}
    ptx::tma_load_3d_2sm(dst, tm, bar, 0, tile * TILE, kb * (BKT / 64));
  } else {                   // chunked MN-major map {64 mn, k rows, extent/64 chunks}: box {64, BKT, TILE/64}
    ptx::tma_load_3d_2sm(dst, tm, bar, 0, kb * BKT, tile * (TILE / 64));
  }
}

template <int BN, int EF, int BKT, int NEPI = EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * NEPI, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP, const TcParams p) {
  using Cfg = Tc2Config<BN, BKT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* stage_all = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + EPI_WARPS * STG_BYTES_PER_WARP);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long t_wait0 = 0, t_wait1 = 0;
  const long long t_start = p.dbg != nullptr ? clock64() : 0;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], 2 * NEPI);  // epilogue warps of both CTAs (used in the leader only)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles * p.split_k;  // m_tiles counts 256-row pair tiles here

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int ks = tile % p.split_k;
        const int mn = tile / p.split_k;
        const int nt = mn % p.n_tiles;
        const int mt = mn / p.n_tiles;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          TIMED_WAIT(t_wait0, ptx::mbar_wait(&empty_bar[stage], phase ^ 1u));
          // EXPERIMENT (p.split_k < 0 encodes "skip B on odd k-blocks"): results are wrong, timing shows traffic sensitivity
          const bool skip_b = p.dbg_skip == 1 && (kb & 1);
          if (p.dbg_skip == 2) {   // EXPERIMENT: no loads at all (pure MMA + epilogue rate on stale shared memory)
            if (rank == 0) ptx::mbar_arrive(&full_bar[stage]);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], skip_b ? 2 * Cfg::A_BYTES : 2 * Cfg::STAGE_BYTES);
          const uint32_t lead_bar = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = tiles + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          load_stage_operand_2sm<BM, BKT>(sa, &tmA, lead_bar, p.a_major, p.a_rpg, p.a_seg, p.a_shift, mt * 2 + (int)rank, kb);
          if (!skip_b)
            load_stage_operand_2sm<BN / 2, BKT>(sb, &tmB, lead_bar, p.b_major, p.b_rpg, p.b_seg, p.b_shift, nt * 2 + (int)rank, kb);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(2 * BM, BN, p.a_major, p.b_major);
      const uint32_t a_adv = p.a_major == 0 ? 2u : (16u * 128u) >> 4;
      const uint32_t b_adv = p.b_major == 0 ? 2u : (16u * 128u) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int ks = tile % p.split_k;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        TIMED_WAIT(t_wait1, ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u));
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          TIMED_WAIT(t_wait0, ptx::mbar_wait(&full_bar[stage], phase));
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = ptx::smem_u32(tiles + stage * Cfg::STAGE_BYTES);
            const uint32_t sb = sa + Cfg::A_BYTES;
            const uint64_t adesc = ptx::make_smem_desc(sa, p.a_major == 0 ? 16u : (uint32_t)(BKT * 128), 1024u);
            const uint64_t bdesc = ptx::make_smem_desc(sb, p.b_major == 0 ? 16u : (uint32_t)(BKT * 128), 1024u);
#pragma unroll
            for (int k = 0; k < BKT / 16; ++k) {
              // K-major: 32 B per UMMA_K inside a 64-wide chunk, chunks of [rows][128 B] back to back;
              // MN-major: 16 k-rows = 2048 B per UMMA_K
              const uint32_t ao = p.a_major == 0 ? (uint32_t)((k >> 2) * (BM * 128 / 16) + (k & 3) * 2) : (uint32_t)(k * a_adv);
              const uint32_t bo = p.b_major == 0 ? (uint32_t)((k >> 2) * ((BN / 2) * 128 / 16) + (k & 3) * 2) : (uint32_t)(k * b_adv);
              ptx::umma_bf16_2sm(tmem_d, adesc + (uint64_t)ao, bdesc + (uint64_t)bo, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit_2sm(&empty_bar[stage], 3);  // both CTAs' smem slots reusable once these MMAs retire
            if (kb == kb1 - 1) ptx::umma_commit_2sm(&tfull_bar[acc], 3);
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows)
    const int quad = warp & 3;
    const int chalf = (warp - 2) >> 2;          // column share of this warp: NEPI / 4 warps per TMEM lane quadrant
    constexpr int NCOL = BN / (NEPI / 4);
    static_assert(NEPI == EPI_WARPS || NCOL == 64, "the lean epilogue drains 64 columns per warp");
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t nstores = 0;
    const uint32_t lead_tempty0 = ptx::mapa_u32(ptx::smem_u32(&tempty_bar[0]), 0);
    const uint32_t lead_tempty1 = ptx::mapa_u32(ptx::smem_u32(&tempty_bar[1]), 0);
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int mn = tile / p.split_k;
      const int nt = mn % p.n_tiles;
      const int mt = mn / p.n_tiles;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      const int m_warp = (mt * 2 + (int)rank) * BM + quad * 32;
      if (NEPI > EPI_WARPS) {
        // sixteen epilogue warps: always the lean row-layout epilogue (the host launches this variant only for it)
        RowOps rops;
        row_lean_load<EF>(p.epi, m_warp, nt * BN + chalf * NCOL, lane, rops);    // row operand: in flight under the MMAs
        TIMED_WAIT(t_wait0, ptx::mbar_wait(&tfull_bar[acc], acc_phase));
        ptx::tc_fence_after();
        const uint32_t tb = acc == 0 ? lead_tempty0 : lead_tempty1;
        epilogue_tile_row_lean<EF>(p.epi, &tmC, &tmP, taddr + (uint32_t)(chalf * NCOL), m_warp, nt * BN + chalf * NCOL, NCOL,
                                   reinterpret_cast<uint8_t*>(stage_all) + (warp - 2) * 2048, lane, rops, [&]() {
                                     ptx::tc_fence_before();
                                     __syncwarp();
                                     if (lane == 0) ptx::mbar_arrive_cluster_relaxed(tb);
                                   });
      } else if (RowEpi<EF>::ok && p.row_epi) {
        TIMED_WAIT(t_wait0, ptx::mbar_wait(&tfull_bar[acc], acc_phase));
        ptx::tc_fence_after();
        const uint32_t tb = acc == 0 ? lead_tempty0 : lead_tempty1;
        epilogue_tile_row<EF>(p.epi, &tmC, &tmP, taddr + (uint32_t)(chalf * NCOL), m_warp, nt * BN + chalf * NCOL, NCOL,
                              reinterpret_cast<uint8_t*>(stage_all + (warp - 2) * (32 * STG_PITCH)), lane, nstores,
                              p.row_epi == 2, [&]() {
                                ptx::tc_fence_before();
                                __syncwarp();
                                if (lane == 0) ptx::mbar_arrive_cluster_relaxed(tb);
                              });
      } else {
        {
          const int tn = tile + num_clusters;           // the next tile of this CTA pair
          if (tn < total_tiles) {
            const int mn2 = tn / p.split_k;
            epilogue_prefetch_l2<EF>(p.epi, ((mn2 / p.n_tiles) * 2 + (int)rank) * BM, (mn2 % p.n_tiles) * BN, BN,
                                     (warp - 2) * 32 + lane);
          }
        }
        EpiRow rows[4];
        EpiChunkOps ops0;
        epilogue_tile_begin<EF>(p.epi, m_warp, nt * BN, chalf * NCOL, lane, rows, ops0);
        TIMED_WAIT(t_wait0, ptx::mbar_wait(&tfull_bar[acc], acc_phase));
        ptx::tc_fence_after();
        epilogue_tile<EF>(p.epi, taddr, m_warp, nt * BN, chalf * NCOL, NCOL,
                      stage_all + (warp - 2) * (32 * STG_PITCH), lane, rows, ops0);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster_relaxed(acc == 0 ? lead_tempty0 : lead_tempty1);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) ptx::bulk_wait<0>();     // this warp's TMA stores have completed before the CTA's smem goes away
  }

  if (p.dbg != nullptr && lane == 0 && warp < 3) {
    long long* o = p.dbg + (long long)blockIdx.x * 8;
    if (warp == 0) o[0] = t_wait0;
    if (warp == 1) { o[1] = t_wait0; o[2] = t_wait1; o[4] = clock64() - t_start; }
    if (warp == 2) o[3] = t_wait0;
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point fetched at run time: no -lcuda link)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  long long inner, rows, groups, rs, gs;
  int b0, b1, b2;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
long long* g_gemm_dbg = nullptr;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

// 3-D bf16 tensor map {inner, rows_per_group, groups} with a {64, box_rows, box_groups} box
int make_map(CUtensorMap* out, const void* ptr, long long inner, long long rows, long long groups, long long rs,
             long long gs, int box_rows, int box_groups) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.inner = inner; key.rows = rows; key.groups = groups; key.rs = rs; key.gs = gs;
  key.b0 = 64; key.b1 = box_rows; key.b2 = box_groups;
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode_fn();
  EGB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  EGB_CHECK(((uintptr_t)ptr % 16) == 0, "TMA operand base %p is not 16-byte aligned", ptr);
  EGB_CHECK((rs % 8) == 0 && rs > 0, "TMA operand row stride %lld must be a positive multiple of 8 bf16", rs);
  EGB_CHECK(groups == 1 || ((gs % 8) == 0 && gs > 0), "TMA operand group stride %lld must be a multiple of 8", gs);
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)groups};
  cuuint64_t strides[2] = {(cuuint64_t)rs * 2, (cuuint64_t)(groups == 1 ? rs : gs) * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, (cuuint32_t)box_groups};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGB_CHECK(r == CUDA_SUCCESS,
            "cuTensorMapEncodeTiled failed (%d): inner=%lld rows=%lld groups=%lld rs=%lld gs=%lld box=%d,%d", (int)r,
            inner, rows, groups, rs, gs, box_rows, box_groups);
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    if (g_maps.size() > 8192) g_maps.clear();
    g_maps[key] = *out;
  }
  return 0;
}

// 2-D bf16 map of a dense output matrix {N columns, M rows} (row stride ld elements) for the TMA stores of the row-layout
// epilogue: box {32 columns, 32 rows}, 64-byte swizzle (the layout row_chunk() writes)
int make_store_map(CUtensorMap* out, const void* ptr, long long N, long long M, long long ld, int box_cols = 32) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.inner = N; key.rows = M; key.groups = 1; key.rs = ld; key.gs = -3232;
  key.b0 = box_cols; key.b1 = 32; key.b2 = 1;
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode_fn();
  EGB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (store map) failed (%d): N=%lld M=%lld ld=%lld", (int)r, N, M, ld);
  std::lock_guard<std::mutex> lk(g_maps_mu);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps[key] = *out;
  return 0;
}

// Decides whether this launch runs the row-layout epilogue and builds its store maps (mc: output, mp: second output).
int setup_row_epilogue(TcParams* p, CUtensorMap* mc, CUtensorMap* mp) {
  static const int enabled = getenv("EGB_GEMM_ROWEPI") ? atoi(getenv("EGB_GEMM_ROWEPI")) : 1;
  memset(mc, 0, sizeof(*mc));
  memset(mp, 0, sizeof(*mp));
  p->row_epi = 0;
  if (!enabled) return 0;
  const int mask = egb_epi_fast_mask(p->epi);
  // (variants with a row operand or column sums run the row layout only in the 16-epilogue-warp pair kernel; the other
  //  kernels check RowEpi<EF> and keep the transposed epilogue for them)
  if (mask == EF_GENERIC || (mask & (EF_ACC | EF_ABWD_GELU)) || p->N < 32) return 0;
  if ((mask & EF_RES) && !EGB_LEAN_RES) return 0;
  if ((mask & (EF_RES | EF_ABWD_RELU | EF_ABWD_MUL | EF_COLSUM)) && (p->N % 64) != 0) return 0;
  auto dense = [&](const EpiMat& m) {
    return m.ptr != nullptr && !m.f32 && m.vec_ok && m.rpg >= p->M && m.rs >= p->N;
  };
  if (!dense(p->epi.c)) return 0;
  if ((mask & EF_PRE) && (!dense(p->epi.c_pre) || (p->N % 16) != 0)) return 0;
  if (make_store_map(mc, p->epi.c.ptr, p->N, p->M, p->epi.c.rs)) return 1;
  if ((mask & EF_PRE) && make_store_map(mp, p->epi.c_pre.ptr, p->N, p->M, p->epi.c_pre.rs)) return 1;
  p->row_epi = 1;
  // 64-column store boxes (half the TMA store instructions): measured neutral on every shape of the bench (EEG qkv
  // 53.2 vs 56.6 us, fc1 + GELU unaffected), i.e. the K = 256 tiles are NOT held up by the number of store instructions;
  // kept as an experiment switch, off by default
  static const int wide = getenv("EGB_GEMM_ROWEPI_WIDE") ? atoi(getenv("EGB_GEMM_ROWEPI_WIDE")) : 0;
  if (wide && !(mask & EF_PRE) && p->N >= 64) {      // single output: the second map slot carries C with 64-column boxes
    if (make_store_map(mp, p->epi.c.ptr, p->N, p->M, p->epi.c.rs, 64)) return 1;
    p->row_epi = 2;
  }
  return 0;
}

// builds the map of one operand. `tile_rows` = BM or BN; `extent_mn` = M or N; K = reduction length
// true when the operand can use a chunked map (single group, no segmentation, 64-aligned extents)
bool operand_chunkable(const egb_operand& o, int extent_mn, int K) {
  if (o.seg_len != 0) return false;
  const long long total_rows = o.major == 0 ? extent_mn : K;
  if (o.rows_per_group > 0 && o.rows_per_group < total_rows) return false;
  if (o.major == 0) return (K % 64) == 0;
  return (extent_mn % 64) == 0;
}

// chunked K-major map for BK = 128 stages: {64 k, rows, K/64 chunks}, box {64, tile_rows, 2}
int make_kmajor_chunked_map(CUtensorMap* out, const egb_operand& o, int extent_mn, int K, int tile_rows) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = o.ptr; key.inner = 64; key.rows = extent_mn; key.groups = K / 64; key.rs = o.row_stride; key.gs = -128;
  key.b0 = 64; key.b1 = tile_rows; key.b2 = 2;
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode_fn();
  EGB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  EGB_CHECK(((uintptr_t)o.ptr % 16) == 0 && (o.row_stride % 8) == 0 && o.row_stride > 0, "TMA operand misaligned");
  cuuint64_t dims[3] = {64u, (cuuint64_t)extent_mn, (cuuint64_t)(K / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)o.row_stride * 2, 128u};
  cuuint32_t box[3] = {64u, (cuuint32_t)tile_rows, 2u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(o.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (chunked K-major) failed (%d): K=%d extent=%d rs=%lld", (int)r, K,
            extent_mn, (long long)o.row_stride);
  std::lock_guard<std::mutex> lk(g_maps_mu);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps[key] = *out;
  return 0;
}


// chunk map {64, rows, chunks} with a {64, box_rows, box_chunks} box (see load_stage_chunked)
int make_chunk_map(CUtensorMap* out, const void* ptr, long long rows, long long chunks, long long rs, int box_rows,
                   int box_chunks) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.inner = 64; key.rows = rows; key.groups = chunks; key.rs = rs; key.gs = -777;
  key.b0 = 64; key.b1 = box_rows; key.b2 = box_chunks;
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode_fn();
  EGB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  EGB_CHECK(((uintptr_t)ptr % 16) == 0 && (rs % 8) == 0 && rs > 0, "TMA operand misaligned");
  cuuint64_t dims[3] = {64u, (cuuint64_t)rows, (cuuint64_t)chunks};
  cuuint64_t strides[2] = {(cuuint64_t)rs * 2, 128u};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, (cuuint32_t)box_chunks};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (chunk map) failed (%d): rows=%lld chunks=%lld rs=%lld box=%d,%d", (int)r,
            rows, chunks, rs, box_rows, box_chunks);
  std::lock_guard<std::mutex> lk(g_maps_mu);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps[key] = *out;
  return 0;
}

// Can this operand be fetched through a chunk map with 128-wide K stages by the single-CTA kernel?
bool operand_chunkable128(const egb_operand& o, int extent_mn, int K, int tile) {
  const long long total_rows = o.major == 0 ? extent_mn : K;
  if (o.rows_per_group > 0 && o.rows_per_group < total_rows) return false;   // grouped views keep the 64-wide path
  if (o.major == 0) {
    if (o.seg_len > 0) return (o.seg_len % 128) == 0 && (K % o.seg_len) == 0;
    return (K % 64) == 0;
  }
  if (o.seg_len > 0) return (o.seg_len % tile) == 0 && (extent_mn % o.seg_len) == 0 && tile >= 64;
  return (extent_mn % 64) == 0 && tile >= 64;
}

int make_operand_chunk_map(CUtensorMap* out, const egb_operand& o, int extent_mn, int K, int tile) {
  if (o.major == 0) {
    long long rows = extent_mn, chunks = K / 64;
    if (o.seg_len > 0) { rows = extent_mn + (long long)(K / o.seg_len - 1) * o.seg_row_shift; chunks = o.seg_len / 64; }
    return make_chunk_map(out, o.ptr, rows, chunks, o.row_stride, tile, 2);
  }
  long long rows = K, chunks = extent_mn / 64;
  if (o.seg_len > 0) { rows = K + (long long)(extent_mn / o.seg_len - 1) * o.seg_row_shift; chunks = o.seg_len / 64; }
  return make_chunk_map(out, o.ptr, rows, chunks, o.row_stride, 128, tile / 64);
}

int make_operand_map(CUtensorMap* out, const egb_operand& o, int extent_mn, int K, int tile_rows, int* rpg_out,
                     int bk = BK) {
  if (o.major == 1 && o.seg_len == 0 && (o.rows_per_group <= 0 || o.rows_per_group >= K) && (extent_mn % 64) == 0 &&
      tile_rows >= 64) {
    // MN-major, single group, MN extent a multiple of 64: 3-D map {64, K rows, extent/64 chunks} (chunk stride 128 B)
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = o.ptr; key.inner = 64; key.rows = K; key.groups = extent_mn / 64; key.rs = o.row_stride; key.gs = -64;
    key.b0 = 64; key.b1 = bk; key.b2 = tile_rows / 64;
    *rpg_out = -1;
    {
      std::lock_guard<std::mutex> lk(g_maps_mu);
      auto it = g_maps.find(key);
      if (it != g_maps.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = get_encode_fn();
    EGB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    EGB_CHECK(((uintptr_t)o.ptr % 16) == 0 && (o.row_stride % 8) == 0 && o.row_stride > 0, "TMA operand misaligned");
    cuuint64_t dims[3] = {64u, (cuuint64_t)K, (cuuint64_t)(extent_mn / 64)};
    cuuint64_t strides[2] = {(cuuint64_t)o.row_stride * 2, 128u};
    cuuint32_t box[3] = {64u, (cuuint32_t)bk, (cuuint32_t)(tile_rows / 64)};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(o.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EGB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (chunked MN-major) failed (%d): K=%d extent=%d rs=%lld", (int)r, K,
              extent_mn, (long long)o.row_stride);
    std::lock_guard<std::mutex> lk(g_maps_mu);
    if (g_maps.size() > 8192) g_maps.clear();
    g_maps[key] = *out;
    return 0;
  }
  const long long total_rows = o.major == 0 ? extent_mn : K;
  long long rpg = o.rows_per_group > 0 ? o.rows_per_group : total_rows;
  long long groups = (total_rows + rpg - 1) / rpg;
  const int span = o.major == 0 ? tile_rows : BK;  // rows of the operand view covered by one tile
  int box_rows, box_groups;
  if (rpg >= span || groups == 1) {
    // single group: rows past the extent are zero-filled by TMA, so the box may overhang
    EGB_CHECK(groups == 1 || rpg % span == 0, "rows_per_group %lld must be a multiple of the tile span %d", rpg, span);
    box_rows = span;
    box_groups = 1;
  } else {
    EGB_CHECK(span % rpg == 0, "tile span %d must be a multiple of rows_per_group %lld", span, rpg);
    box_rows = (int)rpg;
    box_groups = span / (int)rpg;
  }
  long long inner = o.major == 0 ? K : extent_mn;
  long long phys_rows = rpg;
  if (o.seg_len > 0) {
    EGB_CHECK(groups == 1 && o.seg_len % 64 == 0, "segmented operands must be single-group with seg_len %% 64 == 0");
    const long long nseg = (inner + o.seg_len - 1) / o.seg_len;
    phys_rows = rpg + (nseg - 1) * o.seg_row_shift;
    inner = o.seg_len;
    rpg = phys_rows;  // single group: tile coordinates never wrap
  }
  *rpg_out = (int)rpg;
  return make_map(out, o.ptr, inner, phys_rows, groups, o.row_stride, o.group_stride, box_rows, box_groups);
}

template <int BN, int EF, int BKT>
int launch_tc_ef_bk(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mp, const TcParams& p, cudaStream_t stream) {
  using Cfg = TcConfig<BN, BKT>;
  static bool attr_set = false;
  if (!attr_set) {
    EGB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EF, BKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.split_k;
  const int grid = total < egb_num_sms() ? total : egb_num_sms();
  const bool prof = egb_prof_enabled() != 0;
  if (prof) {
    const double out_b = p.epi.c.f32 ? 4.0 : 2.0;
    egb_prof_begin(stream, 2.0 * p.M * (double)p.N * p.K, 2.0 * ((double)p.M * p.K + (double)p.N * p.K) + out_b * p.M * p.N, 0);
    egb_prof_tag(p.M, p.N, p.K, 1e10 + BN * 1e7 + EF);      // kernel variant 1 = single CTA
  }
  gemm_tc_kernel<BN, EF, BKT><<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(ma, mb, mc, mp, p);
  if (prof) egb_prof_end(stream);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

// run-time epilogue mask -> compile-time instantiation (FAST = the specialised set is built for this tile shape)
#define EGB_EF_SWITCH(FN, BN_, FAST)                                                                            \
  if (FAST) {                                                                                                   \
    switch (egb_epi_fast_mask(p.epi)) {                                                                         \
      case 0: return FN<BN_, 0>(ma, mb, mc, mp, p, stream);                                                             \
      case EF_BIAS: return FN<BN_, EF_BIAS>(ma, mb, mc, mp, p, stream);                                                 \
      case EF_BIAS | EF_RES: return FN<BN_, EF_BIAS | EF_RES>(ma, mb, mc, mp, p, stream);                               \
      case EF_BIAS | EF_RELU: return FN<BN_, EF_BIAS | EF_RELU>(ma, mb, mc, mp, p, stream);                             \
      case EF_BIAS | EF_GELU | EF_PRE | EF_DGELU:                                                               \
        return FN<BN_, EF_BIAS | EF_GELU | EF_PRE | EF_DGELU>(ma, mb, mc, mp, p, stream);                               \
      case EF_ABWD_RELU: return FN<BN_, EF_ABWD_RELU>(ma, mb, mc, mp, p, stream);                                       \
      case EF_ABWD_MUL: return FN<BN_, EF_ABWD_MUL>(ma, mb, mc, mp, p, stream);                                         \
      case EF_ABWD_RELU | EF_COLSUM: return FN<BN_, EF_ABWD_RELU | EF_COLSUM>(ma, mb, mc, mp, p, stream);               \
      case EF_ABWD_MUL | EF_COLSUM: return FN<BN_, EF_ABWD_MUL | EF_COLSUM>(ma, mb, mc, mp, p, stream);                 \
      case EF_ACC: return FN<BN_, EF_ACC>(ma, mb, mc, mp, p, stream);                                                   \
      default: break;                                                                                           \
    }                                                                                                           \
  }                                                                                                             \
  return FN<BN_, EF_GENERIC>(ma, mb, mc, mp, p, stream);

template <int BN, int EF>
int launch_tc_ef(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mp, const TcParams& p, cudaStream_t stream) {
  if (p.bkt == 128 && BN <= 128) return launch_tc_ef_bk<BN, EF, 128>(ma, mb, mc, mp, p, stream);
  return launch_tc_ef_bk<BN, EF, 64>(ma, mb, mc, mp, p, stream);
}

template <int BN>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mp, const TcParams& p, cudaStream_t stream) {
  EGB_EF_SWITCH(launch_tc_ef, BN, true)
}

// the 16-epilogue-warp build of the pair kernel (lean row-layout epilogue): plain / bias / bias+ReLU / bias+GELU+GELU'
template <int EF>
struct Wide16 {
  static constexpr bool ok = EF == 0 || EF == EF_BIAS || EF == (EF_BIAS | EF_RELU) || EF == (EF_BIAS | EF_GELU | EF_PRE | EF_DGELU) ||
                             EF == EF_ABWD_MUL || EF == EF_ABWD_RELU || (EF == (EF_BIAS | EF_RES) && EGB_LEAN_RES) ||
                             EF == (EF_ABWD_MUL | EF_COLSUM) || EF == (EF_ABWD_RELU | EF_COLSUM);
  // (bias + residual was measured on this path too and lost: proj + residual 69 -> 96 us, fc2 + residual 168 -> 176 us, EEG
  //  40 -> 53 us; the residual is L2-resident and the transposed epilogue's sector-exact reads serve it better.  The
  //  activation-backward variants: EEG dX-through-ReLU 145 -> 129 us, ViT dX-through-GELU' 291 -> 289 us.)
};

template <int BN, int EF, int BKT>
int launch_tc2_ef_bk16(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mp, const TcParams& p, cudaStream_t stream) {
  using Cfg = Tc2Config<BN, BKT>;
  static bool attr_set = false;
  if (!attr_set) {
    EGB_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN, EF, BKT, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.split_k;
  const int pairs = egb_num_sms() / 2;
  const int grid = 2 * (total < pairs ? total : pairs);
  const bool prof = egb_prof_enabled() != 0;
  if (prof) {
    const double out_b = p.epi.c.f32 ? 4.0 : 2.0;
    egb_prof_begin(stream, 2.0 * p.M * (double)p.N * p.K, 2.0 * ((double)p.M * p.K + (double)p.N * p.K) + out_b * p.M * p.N, 0);
    egb_prof_tag(p.M, p.N, p.K, 3e10 + BN * 1e7 + EF);      // 3 = CTA pair, 16 epilogue warps
  }
  gemm_tc2_kernel<BN, EF, BKT, 16><<<grid, 64 + 32 * 16, Cfg::SMEM_BYTES, stream>>>(ma, mb, mc, mp, p);
  if (prof) egb_prof_end(stream);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

template <int BN, int EF, int BKT>
int launch_tc2_ef_bk(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mp, const TcParams& p, cudaStream_t stream) {
  using Cfg = Tc2Config<BN, BKT>;
  if constexpr (Wide16<EF>::ok && BN == 256) {
    static const int w16 = getenv("EGB_GEMM_EPI16") ? atoi(getenv("EGB_GEMM_EPI16")) : 1;
    if (w16 && p.row_epi) return launch_tc2_ef_bk16<BN, EF, BKT>(ma, mb, mc, mp, p, stream);
  }
  static bool attr_set = false;
  if (!attr_set) {
    EGB_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN, EF, BKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.split_k;
  const int pairs = egb_num_sms() / 2;
  const int grid = 2 * (total < pairs ? total : pairs);
  const bool prof = egb_prof_enabled() != 0;
  if (prof) {
    const double out_b = p.epi.c.f32 ? 4.0 : 2.0;
    egb_prof_begin(stream, 2.0 * p.M * (double)p.N * p.K, 2.0 * ((double)p.M * p.K + (double)p.N * p.K) + out_b * p.M * p.N, 0);
    egb_prof_tag(p.M, p.N, p.K, 2e10 + BN * 1e7 + EF);      // 2 = CTA pair, 8 epilogue warps
  }
  gemm_tc2_kernel<BN, EF, BKT><<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(ma, mb, mc, mp, p);
  if (prof) egb_prof_end(stream);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

template <int BN, int EF>
int launch_tc2_ef(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mp, const TcParams& p, cudaStream_t stream) {
  if (p.bkt == 128) return launch_tc2_ef_bk<BN, EF, 128>(ma, mb, mc, mp, p, stream);
  return launch_tc2_ef_bk<BN, EF, 64>(ma, mb, mc, mp, p, stream);
}

template <int BN>
int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mp, const TcParams& p, cudaStream_t stream) {
  EGB_EF_SWITCH(launch_tc2_ef, BN, true)
}

// CTA-pair path: 256 x BN tiles.  Chosen for problems with at least one full wave of pair tiles.
int gemm_tc_pair(const egb_gemm_desc* d, cudaStream_t stream, int BN) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.a_major = d->a.major; p.b_major = d->b.major;
  p.a_seg = d->a.seg_len; p.a_shift = d->a.seg_row_shift;
  p.b_seg = d->b.seg_len; p.b_shift = d->b.seg_row_shift;
  CUtensorMap ma, mb;
  static const int bk128 = getenv("EGB_GEMM_BK128") ? atoi(getenv("EGB_GEMM_BK128")) : 1;
  p.bkt = (bk128 && operand_chunkable(d->a, d->M, d->K) && operand_chunkable(d->b, d->N, d->K)) ? 128 : 64;
  if (p.bkt == 128) {
    if (d->a.major == 0) { if (make_kmajor_chunked_map(&ma, d->a, d->M, d->K, BM)) return 1; p.a_rpg = -2; }
    else if (make_operand_map(&ma, d->a, d->M, d->K, BM, &p.a_rpg, 128)) return 1;
    if (d->b.major == 0) { if (make_kmajor_chunked_map(&mb, d->b, d->N, d->K, BN / 2)) return 1; p.b_rpg = -2; }
    else if (make_operand_map(&mb, d->b, d->N, d->K, BN / 2, &p.b_rpg, 128)) return 1;
  } else {
    if (make_operand_map(&ma, d->a, d->M, d->K, BM, &p.a_rpg)) return 1;
    if (make_operand_map(&mb, d->b, d->N, d->K, BN / 2, &p.b_rpg)) return 1;
  }
  p.m_tiles = (d->M + 2 * BM - 1) / (2 * BM);
  p.n_tiles = (d->N + BN - 1) / BN;
  p.k_blocks = (d->K + p.bkt - 1) / p.bkt;
  int split = 1;
  if (d->accumulate) {
    split = d->split_k;
    if (split <= 0) {
      // pick the split whose work-item count fills whole waves of CTA pairs best (each item keeps >= 16 k-blocks)
      const int tiles = p.m_tiles * p.n_tiles;
      const int pairs = egb_num_sms() / 2;
      int max_split = p.k_blocks / (p.bkt == 128 ? 8 : 16);
      if (max_split > 64) max_split = 64;
      if (max_split < 1) max_split = 1;
      double best = -1.0;
      split = 1;
      for (int sp = 1; sp <= max_split; ++sp) {
        const int kps = (p.k_blocks + sp - 1) / sp;
        const int items = tiles * ((p.k_blocks + kps - 1) / kps);
        const int waves = (items + pairs - 1) / pairs;
        // time ~ waves * (k-blocks per item + fixed per-item cost of the accumulate epilogue, ~12 k-blocks)
        const double cost = (double)waves * (kps + (p.bkt == 128 ? 6 : 12));
        const double score = 1.0 / cost;
        if (score > best * 1.0001) { best = score; split = sp; }
      }
    }
    if (split > p.k_blocks) split = p.k_blocks;
  }
  p.kb_per_split = (p.k_blocks + split - 1) / split;
  p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  if (egb_fill_epilogue(d, &p.epi)) return 1;
  p.dbg = g_gemm_dbg;
  static const int skipb = getenv("EGB_GEMM_SKIPB") ? atoi(getenv("EGB_GEMM_SKIPB")) : 0;
  p.dbg_skip = skipb;
  CUtensorMap mc, mp;
  if (setup_row_epilogue(&p, &mc, &mp)) return 1;
  return BN == 256 ? launch_tc2<256>(ma, mb, mc, mp, p, stream) : launch_tc2<128>(ma, mb, mc, mp, p, stream);
}

}  // namespace


// Debug aid: device buffer of [grid][8] int64 receiving per-CTA barrier-wait cycle totals of the following tensor-core
// GEMM launches: [0] producer waiting for a free stage, [1] MMA warp waiting for operands, [2] MMA warp waiting for
// a drained accumulator, [3] epilogue warp waiting for an accumulator, [4] kernel cycles.  NULL disables.
// 3-D bf16 tensor map {inner, rows, groups} with a {64 elements, box_rows, 1} box, 128-byte swizzle (cached); used by
// the attention kernels to stage per-head row tiles with one TMA instruction each
int egb_tmap_rows64(CUtensorMap* out, const void* ptr, long long inner, long long rows, long long groups, long long rs,
                    long long gs, int box_rows) {
  return make_map(out, ptr, inner, rows, groups, rs, gs, box_rows, 1);
}

// 2-D bf16 tensor map {inner elements, rows} (row stride rs elements) with a {box_inner, box_rows} box and a 64- or
// 128-byte swizzle (box_inner * 2 bytes must equal the swizzle span); cached.  Used by the implicit-GEMM convolution.
int egb_tmap_2d(CUtensorMap* out, const void* ptr, long long inner, long long rows, long long rs, int box_inner, int box_rows) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.inner = inner; key.rows = rows; key.groups = 1; key.rs = rs; key.gs = -2222;
  key.b0 = box_inner; key.b1 = box_rows; key.b2 = 1;
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode_fn();
  EGB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  EGB_CHECK(box_inner == 32 || box_inner == 64, "tmap_2d: box of 32 or 64 bf16 elements");
  EGB_CHECK(((uintptr_t)ptr % 16) == 0 && (rs % 8) == 0 && rs > 0 && box_rows <= 256, "tmap_2d: misaligned operand");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)rs * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (2-D) failed (%d): inner=%lld rows=%lld rs=%lld box=%d,%d", (int)r, inner,
            rows, rs, box_inner, box_rows);
  std::lock_guard<std::mutex> lk(g_maps_mu);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps[key] = *out;
  return 0;
}

extern "C" int egb_debug_gemm_timing(long long* device_buf) {
  g_gemm_dbg = device_buf;
  return 0;
}

// Called from egb_gemm for in_dtype == EGB_BF16.
int egb_gemm_tc(const egb_gemm_desc* d, cudaStream_t stream) {
  EGB_CHECK(d->M > 0 && d->N > 0 && d->K > 0, "gemm: empty problem %dx%dx%d", d->M, d->N, d->K);
  {
    // CTA pairs when the problem offers enough 256-row tiles and N fills a 256- or 128-wide pair tile
    // (EGB_GEMM_PAIR=0 disables the path, =2 forces it for every eligible shape: used by the device tests)
    static const int pair_mode = getenv("EGB_GEMM_PAIR") ? atoi(getenv("EGB_GEMM_PAIR")) : 1;
    const int bn2 = (d->N % 256 == 0 || d->N >= 1024) ? 256 : ((d->N % 128 == 0) ? 128 : 0);
    if (pair_mode && bn2 != 0 && d->M >= 256) {
      const long long tiles = (long long)((d->M + 255) / 256) * ((d->N + bn2 - 1) / bn2);
      if (pair_mode == 2 || tiles >= egb_num_sms() / 2 || d->accumulate) return gemm_tc_pair(d, stream, bn2);
    }
  }
  int BN = d->N > 128 ? 256 : (d->N > 64 ? 128 : 64);
  // a 256-wide tile wastes MMA work when N is just above a multiple of 128
  if (BN == 256 && (d->N % 256) != 0 && (d->N % 256) <= 128 && d->N < 1024) BN = 128;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.a_major = d->a.major; p.b_major = d->b.major;
  p.a_seg = d->a.seg_len; p.a_shift = d->a.seg_row_shift;
  p.b_seg = d->b.seg_len; p.b_shift = d->b.seg_row_shift;
  CUtensorMap ma, mb;
  static const int bk128s = getenv("EGB_GEMM_BK128") ? atoi(getenv("EGB_GEMM_BK128")) : 1;
  p.bkt = (bk128s && BN <= 128 && operand_chunkable128(d->a, d->M, d->K, BM) && operand_chunkable128(d->b, d->N, d->K, BN))
              ? 128 : 64;
  if (p.bkt == 128) {
    if (make_operand_chunk_map(&ma, d->a, d->M, d->K, BM)) return 1;
    if (make_operand_chunk_map(&mb, d->b, d->N, d->K, BN)) return 1;
  } else {
    if (make_operand_map(&ma, d->a, d->M, d->K, BM, &p.a_rpg)) return 1;
    if (make_operand_map(&mb, d->b, d->N, d->K, BN, &p.b_rpg)) return 1;
  }
  p.m_tiles = (d->M + BM - 1) / BM;
  p.n_tiles = (d->N + BN - 1) / BN;
  p.k_blocks = (d->K + p.bkt - 1) / p.bkt;
  int split = 1;
  if (d->accumulate) {
    EGB_CHECK(d->c.dtype == EGB_F32, "gemm: accumulate requires an fp32 output");
    split = d->split_k;
    if (split <= 0) {
      const int tiles = p.m_tiles * p.n_tiles;
      split = (2 * egb_num_sms() + tiles - 1) / tiles;
      const int max_split = (p.k_blocks + 7) / 8;  // keep >= 8 k-stages per split
      if (split > max_split) split = max_split;
      if (split < 1) split = 1;
    }
    if (split > p.k_blocks) split = p.k_blocks;
  }
  p.kb_per_split = (p.k_blocks + split - 1) / split;
  p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  if (egb_fill_epilogue(d, &p.epi)) return 1;
  p.dbg = g_gemm_dbg;
  {
    static const int skipb = getenv("EGB_GEMM_SKIPB") ? atoi(getenv("EGB_GEMM_SKIPB")) : 0;
    p.dbg_skip = skipb;
  }
  CUtensorMap mc, mp;
  if (setup_row_epilogue(&p, &mc, &mp)) return 1;
  switch (BN) {
    case 256: return launch_tc<256>(ma, mb, mc, mp, p, stream);
    case 128: return launch_tc<128>(ma, mb, mc, mp, p, stream);
    default: return launch_tc<64>(ma, mb, mc, mp, p, stream);
  }
}
