// Experiment: can a K-major, 128-byte-swizzled UMMA operand start at an ARBITRARY row of a shared-memory tile (start
// address not 1024-byte aligned) when the descriptor's matrix-base-offset field carries (address >> 7) & 7?
// A: [R rows][64 bf16] staged with the usual 128-B swizzle relative to a 1024-aligned base; B: [32][64].
// D[m][c] = sum_k A[m + shift][k] * B[c][k] for m < 128, c < 32, computed by four tcgen05.mma (K = 16 each).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../ptx.cuh"

using bf16 = __nv_bfloat16;

__global__ void __launch_bounds__(128) shift_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ out,
                                                    int R, int shift, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = sA + ((R * 128 + 1023) & ~1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int idx = threadIdx.x; idx < R * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
  }
  for (int idx = threadIdx.x; idx < 32 * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<32>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, 32, 0, 0);
    const uint32_t a_addr = ptx::smem_u32(sA) + (uint32_t)shift * 128u;
    uint64_t ad = ptx::make_smem_desc(a_addr, 16u, 1024u);
    if (mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;          // matrix base offset
    const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(sB), 16u, 1024u);
    for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  uint32_t v[32];
  ptx::tmem_ld32(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
  ptx::tmem_ld_wait();
  for (int c = 0; c < 32; ++c) out[threadIdx.x * 32 + c] = __uint_as_float(v[c]);
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<32>(tmem);
}

// Same question for 64-BYTE rows (32 bf16 = the 32-channel activation of the forward convolution): tile staged with the
// 64-byte swizzle (16-byte chunk index ^= (row >> 1) & 3), descriptor swizzle mode 4 (SWIZZLE_64B), SBO = 512 B.
// D[m][c] = sum_{k < 32} A[m + shift][k] * B[c][k], c < 64: two tcgen05.mma (K = 16 each).
__global__ void __launch_bounds__(128) shift64_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ out,
                                                      int R, int shift) {
  extern __shared__ uint8_t raw[];
  uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = sA + ((R * 64 + 1023) & ~1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int idx = threadIdx.x; idx < R * 4; idx += blockDim.x) {
    const int r = idx >> 2, c = idx & 3;
    *reinterpret_cast<uint4*>(sA + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 32 + c * 8);
  }
  for (int idx = threadIdx.x; idx < 64 * 4; idx += blockDim.x) {
    const int r = idx >> 2, c = idx & 3;
    *reinterpret_cast<uint4*>(sB + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 32 + c * 8);
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<64>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 0, 0);
    auto desc64 = [](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
      d |= (uint64_t)((16u >> 4) & 0x3FFFu) << 16;
      d |= (uint64_t)((512u >> 4) & 0x3FFFu) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)4 << 61;   // SWIZZLE_64B
      return d;
    };
    const uint64_t ad = desc64(ptx::smem_u32(sA) + (uint32_t)shift * 64u);
    const uint64_t bd = desc64(ptx::smem_u32(sB));
    for (int k = 0; k < 2; ++k) ptx::umma_bf16(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    ptx::tmem_ld32(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + (uint32_t)(32 * h), v);
    ptx::tmem_ld_wait();
    for (int c = 0; c < 32; ++c) out[threadIdx.x * 64 + 32 * h + c] = __uint_as_float(v[c]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<64>(tmem);
}

int test64() {
  const int R = 160;
  std::vector<bf16> hA(R * 32), hB(64 * 32);
  std::vector<float> fA(R * 32), fB(64 * 32);
  srand(2);
  for (int i = 0; i < R * 32; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < 64 * 32; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
  bf16 *dA, *dB; float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = ((R * 64 + 1023) & ~1023) + 64 * 64 + 2048;
  std::vector<float> ho(128 * 64);
  for (int shift : {0, 8, 16, 1, 2, 3, 5, 7, 10, 11, 21}) {
    cudaMemset(dO, 0, 128 * 64 * 4);
    shift64_kernel<<<1, 128, smem>>>(dA, dB, dO, R, shift);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("64B shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho.data(), dO, ho.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < 128; ++m)
      for (int c = 0; c < 64; ++c) {
        double ref = 0;
        for (int k = 0; k < 32; ++k) ref += (double)fA[(m + shift) * 32 + k] * fB[c * 32 + k];
        worst = fmax(worst, fabs(ref - ho[m * 64 + c]));
      }
    printf("64-byte rows, SWIZZLE_64B  shift %2d rows: max abs err %.3e  %s\n", shift, worst, worst < 1e-3 ? "OK" : "WRONG");
  }
  return 0;
}

// Weight-gradient form: BOTH operands MN-major (K = tile rows).  A = the 32-channel activation tile [R rows][32 bf16]
// (64-byte rows, SWIZZLE_64B) read as M = 128 = (j = 0..3, c): the descriptor's leading-dimension byte offset -- the
// distance between consecutive 32-element M atoms -- is set to ONE ROW (64 B), so atom j is the same tile one row
// further down: M index 32 j + c of K index k addresses row (k + shift + j), channel c.  B = [KR rows][64 bf16]
// (128-byte rows, SWIZZLE_128B), N = 64.  D[32 j + c][o] = sum_{k < 32} A[k + shift + j][c] * B[k][o].
__global__ void __launch_bounds__(128) dw_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ out,
                                                 int R, int shift, int lbo_bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = sA + ((R * 64 + 1023) & ~1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int idx = threadIdx.x; idx < R * 4; idx += blockDim.x) {
    const int r = idx >> 2, c = idx & 3;
    *reinterpret_cast<uint4*>(sA + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 32 + c * 8);
  }
  for (int idx = threadIdx.x; idx < 32 * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<64>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 1, 1);
    uint64_t ad = 0;
    const uint32_t a_addr = ptx::smem_u32(sA) + (uint32_t)shift * 64u;
    ad |= (uint64_t)((a_addr & 0x3FFFFu) >> 4);
    ad |= (uint64_t)(((uint32_t)lbo_bytes >> 4) & 0x3FFFu) << 16;
    ad |= (uint64_t)((512u >> 4) & 0x3FFFu) << 32;
    ad |= (uint64_t)1 << 46;
    ad |= (uint64_t)4 << 61;   // SWIZZLE_64B
    const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(sB), 32u * 128u, 1024u);
    for (int k = 0; k < 2; ++k) ptx::umma_bf16(tmem, ad + (uint64_t)(64 * k), bd + (uint64_t)(128 * k), idesc, k > 0 ? 1u : 0u);
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    ptx::tmem_ld32(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + (uint32_t)(32 * h), v);
    ptx::tmem_ld_wait();
    for (int c = 0; c < 32; ++c) out[threadIdx.x * 64 + 32 * h + c] = __uint_as_float(v[c]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<64>(tmem);
}

int test_dw() {
  const int R = 96;
  std::vector<bf16> hA(R * 32), hB(32 * 64);
  std::vector<float> fA(R * 32), fB(32 * 64);
  srand(3);
  for (int i = 0; i < R * 32; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < 32 * 64; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
  bf16 *dA, *dB; float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = ((R * 64 + 1023) & ~1023) + 32 * 128 + 2048;
  std::vector<float> ho(128 * 64);
  for (int shift : {0, 8, 1, 2, 3, 5, 7, 10, 11, 21, 34, 35}) {
    cudaMemset(dO, 0, 128 * 64 * 4);
    dw_kernel<<<1, 128, smem>>>(dA, dB, dO, R, shift, 64);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("dW shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho.data(), dO, ho.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, worst_j0 = 0;
    for (int m = 0; m < 128; ++m)
      for (int o = 0; o < 64; ++o) {
        const int j = m >> 5, c = m & 31;
        double ref = 0;
        for (int k = 0; k < 32; ++k) ref += (double)fA[(k + shift + j) * 32 + c] * fB[k * 64 + o];
        worst = fmax(worst, fabs(ref - ho[m * 64 + o]));
        if (j == 0) worst_j0 = fmax(worst_j0, fabs(ref - ho[m * 64 + o]));
      }
    printf("MN-major x MN-major, M atoms one row apart  shift %2d rows: max abs err %.3e (first atom alone %.3e)  %s\n", shift, worst,
           worst_j0, worst < 1e-3 ? "OK" : "WRONG");
  }
  return 0;
}

int main() {
  if (test_dw()) return 1;
  if (test64()) return 1;

  const int R = 160;
  std::vector<bf16> hA(R * 64), hB(32 * 64);
  std::vector<float> fA(R * 64), fB(32 * 64);
  srand(1);
  for (int i = 0; i < R * 64; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < 32 * 64; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
  bf16 *dA, *dB; float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 32 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = ((R * 128 + 1023) & ~1023) + 32 * 128 + 2048;
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> ho(128 * 32);
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 8, 1, 2, 3, 5, 7, 10, 11, 21}) {
      cudaMemset(dO, 0, 128 * 32 * 4);
      shift_kernel<<<1, 128, smem>>>(dA, dB, dO, R, shift, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(ho.data(), dO, ho.size() * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int c = 0; c < 32; ++c) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)fA[(m + shift) * 64 + k] * fB[c * 64 + k];
          worst = fmax(worst, fabs(ref - ho[m * 32 + c]));
        }
      printf("base_offset %s  shift %2d rows: max abs err %.3e  %s\n", mode ? "set  " : "unset", shift, worst, worst < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
