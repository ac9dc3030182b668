// Standalone device test for egb_gemm (no torch): checks the tcgen05 kernel and the FFMA kernel
// against a double-precision host reference over K-major / MN-major operands, grouped
// (conv-style, overlapping-row) views, every epilogue, split-K accumulation, and times the
// ViT-sized problems.  Run on a B200:  ./gemm_test [quick]
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../../include/eyegaze_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct HostOperand {
  std::vector<float> data;  // values (already bf16-rounded when dtype is bf16)
  int major, rpg;
  long long rs, gs;
  long long base;  // element offset of the view inside data
};

static double op_at(const HostOperand& o, int mn, int k) {
  int row = o.major == 0 ? mn : k;
  int inner = o.major == 0 ? k : mn;
  int g = row / o.rpg, r = row % o.rpg;
  return o.data[o.base + (long long)g * o.gs + (long long)r * o.rs + inner];
}

static void* upload(const std::vector<float>& v, int dtype) {
  void* d;
  if (dtype == EGB_F32) {
    CK(cudaMalloc(&d, v.size() * 4));
    CK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  } else {
    std::vector<__nv_bfloat16> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16_rn(v[i]);
    CK(cudaMalloc(&d, v.size() * 2));
    CK(cudaMemcpy(d, h.data(), v.size() * 2, cudaMemcpyHostToDevice));
  }
  return d;
}

static double gelu_h(double x) { return 0.5 * x * (1.0 + erf(x * 0.7071067811865475)); }
static double gelu_grad_h(double x) {
  return 0.5 * (1.0 + erf(x * 0.7071067811865475)) + x * 0.3989422804014327 * exp(-0.5 * x * x);
}

struct Case {
  const char* name;
  int M, N, K;
  int a_major, b_major;
  int a_rpg, b_rpg;          // 0 = ungrouped
  long long a_rs, a_gs, b_rs, b_gs;  // 0 = dense default
  int c_rpg; long long c_rs, c_gs;   // output grouping (0 = dense)
  int act, act_bwd, use_bias, use_res, use_pre, c_f32, accumulate;
  float alpha;
};

static int run_case(const Case& cs, int in_dtype, bool check_full, int timing_iters) {
  const int M = cs.M, N = cs.N, K = cs.K;
  HostOperand A, B;
  auto setup = [&](HostOperand& o, int major, int extent_mn, int rpg, long long rs, long long gs) {
    o.major = major;
    long long rows = major == 0 ? extent_mn : K;
    long long inner = major == 0 ? K : extent_mn;
    o.rpg = rpg > 0 ? rpg : (int)rows;
    o.rs = rs > 0 ? rs : inner;
    long long groups = (rows + o.rpg - 1) / o.rpg;
    o.gs = gs > 0 ? gs : o.rs * o.rpg;
    o.base = 0;
    long long need = (groups - 1) * o.gs + (long long)(o.rpg - 1) * o.rs + inner;
    o.data.resize(need + 64);
    for (auto& x : o.data) { x = frand(); if (in_dtype == EGB_BF16) x = bf16_round(x); }
  };
  setup(A, cs.a_major, M, cs.a_rpg, cs.a_rs, cs.a_gs);
  setup(B, cs.b_major, N, cs.b_rpg, cs.b_rs, cs.b_gs);

  const int c_rpg = cs.c_rpg > 0 ? cs.c_rpg : M;
  const long long c_rs = cs.c_rs > 0 ? cs.c_rs : N;
  const long long c_gs = cs.c_gs > 0 ? cs.c_gs : c_rs * c_rpg;
  const long long c_groups = (M + c_rpg - 1) / c_rpg;
  const long long c_elems = (c_groups - 1) * c_gs + (long long)(c_rpg - 1) * c_rs + N + 16;
  auto c_off = [&](int m) { return (long long)(m / c_rpg) * c_gs + (long long)(m % c_rpg) * c_rs; };

  std::vector<float> bias(N), res(c_elems), aux(c_elems), cinit(c_elems);
  for (auto& x : bias) x = frand();
  const int c_dtype = cs.c_f32 ? EGB_F32 : EGB_BF16;
  for (auto& x : res) { x = frand(); if (!cs.c_f32) x = bf16_round(x); }
  for (auto& x : aux) { x = frand(); if (!cs.c_f32) x = bf16_round(x); if (cs.act_bwd == 1 && x < 0) x = 0; }
  for (auto& x : cinit) x = cs.accumulate ? frand() : -777.f;

  void* dA = upload(A.data, in_dtype);
  void* dB = upload(B.data, in_dtype);
  void* dC = upload(cinit, c_dtype);
  void* dPre = upload(cinit, c_dtype);
  void* dRes = upload(res, c_dtype);
  void* dAux = upload(aux, c_dtype);
  float* dBias;
  CK(cudaMalloc(&dBias, N * 4));
  CK(cudaMemcpy(dBias, bias.data(), N * 4, cudaMemcpyHostToDevice));

  egb_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = M; d.N = N; d.K = K; d.in_dtype = in_dtype;
  d.a = {dA, A.major, cs.a_rpg > 0 ? cs.a_rpg : 0, A.rs, A.gs};
  d.b = {dB, B.major, cs.b_rpg > 0 ? cs.b_rpg : 0, B.rs, B.gs};
  d.c = {dC, c_dtype, cs.c_rpg > 0 ? cs.c_rpg : 0, c_rs, c_gs};
  if (cs.use_pre) d.c_pre = {dPre, c_dtype, cs.c_rpg > 0 ? cs.c_rpg : 0, c_rs, c_gs};
  if (cs.use_res) d.residual = {dRes, c_dtype, cs.c_rpg > 0 ? cs.c_rpg : 0, c_rs, c_gs};
  if (cs.act_bwd) d.aux = {dAux, c_dtype, cs.c_rpg > 0 ? cs.c_rpg : 0, c_rs, c_gs};
  d.bias = cs.use_bias ? dBias : nullptr;
  d.alpha = cs.alpha; d.act = cs.act; d.act_bwd = cs.act_bwd; d.aux_scale = 1.25f;
  d.accumulate = cs.accumulate;

  if (egb_gemm(&d, 0) != 0) {
    printf("[FAIL] %-28s dtype=%d : egb_gemm error: %s\n", cs.name, in_dtype, egb_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[FAIL] %-28s dtype=%d : kernel error %s\n", cs.name, in_dtype, cudaGetErrorString(e));
    exit(3);  // context is dead
  }

  std::vector<float> hc(c_elems), hpre(c_elems);
  auto download = [&](void* dptr, std::vector<float>& out) {
    if (c_dtype == EGB_F32) {
      CK(cudaMemcpy(out.data(), dptr, c_elems * 4, cudaMemcpyDeviceToHost));
    } else {
      std::vector<__nv_bfloat16> t(c_elems);
      CK(cudaMemcpy(t.data(), dptr, c_elems * 2, cudaMemcpyDeviceToHost));
      for (long long i = 0; i < c_elems; ++i) out[i] = __bfloat162float(t[i]);
    }
  };
  download(dC, hc);
  if (cs.use_pre) download(dPre, hpre);

  double max_err = 0, max_ref = 0;
  long long checked = 0, bad = 0;
  const double tol_rel = (c_dtype == EGB_BF16) ? 1.2e-2 : (in_dtype == EGB_BF16 ? 2e-4 : 2e-5);
  auto check_one = [&](int m, int n) {
    double acc = 0;
    for (int k = 0; k < K; ++k) acc += op_at(A, m, k) * op_at(B, n, k);
    double x = acc * cs.alpha;
    if (cs.use_bias) x += bias[n];
    const double pre = x;
    if (cs.act == 1) x = x > 0 ? x : 0;
    if (cs.act == 2 || cs.act == 3) x = gelu_h(x);
    const long long off = c_off(m) + n;
    if (cs.act_bwd == 1) x = aux[off] != 0 ? x * 1.25 : 0;
    if (cs.act_bwd == 2) x *= gelu_grad_h(aux[off]);
    if (cs.act_bwd == 3) x *= aux[off];
    if (cs.use_res) x += res[off];
    if (cs.accumulate) x += cinit[off];
    const double scale = sqrt((double)K) * 0.1 + fabs(x);
    const double err = fabs(hc[off] - x);
    if (err > max_err) max_err = err;
    if (fabs(x) > max_ref) max_ref = fabs(x);
    if (err > tol_rel * scale) ++bad;
    if (cs.use_pre) {
      const double pre_want = cs.act == 3 ? gelu_grad_h(pre) : pre;   // act 3 saves gelu'(pre)
      const double ep = fabs(hpre[off] - pre_want);
      if (ep > tol_rel * (sqrt((double)K) * 0.1 + fabs(pre_want))) ++bad;
    }
    ++checked;
  };
  if (check_full) {
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) check_one(m, n);
  } else {
    for (int i = 0; i < 4000; ++i) {
      rng_state = rng_state * 1664525u + 1013904223u;
      int m = (rng_state >> 4) % M;
      rng_state = rng_state * 1664525u + 1013904223u;
      int n = (rng_state >> 4) % N;
      check_one(m, n);
    }
    for (int n = 0; n < N; n += 7) { check_one(0, n); check_one(M - 1, n); }
    for (int m = 0; m < M; m += 61) { check_one(m, 0); check_one(m, N - 1); }
  }
  // untouched padding must stay untouched (grouped outputs write only their own rows)
  double ms = 0;
  if (timing_iters > 0 && !cs.accumulate) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) egb_gemm(&d, 0);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < timing_iters; ++i) egb_gemm(&d, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t;
    CK(cudaEventElapsedTime(&t, e0, e1));
    ms = t / timing_iters;
  }
  printf("[%s] %-28s dtype=%s M=%d N=%d K=%d maj=%d%d checked=%lld bad=%lld max_err=%.3e max_ref=%.2f", bad ? "FAIL" : " ok ",
         cs.name, in_dtype == EGB_BF16 ? "bf16" : "f32 ", M, N, K, cs.a_major, cs.b_major, checked, bad, max_err, max_ref);
  if (ms > 0) printf("  %.3f ms  %.1f TFLOP/s", ms, 2.0 * M * N * K / ms * 1e-9);
  printf("\n");
  fflush(stdout);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dPre); cudaFree(dRes); cudaFree(dAux); cudaFree(dBias);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && strcmp(argv[1], "quick") == 0;
  const bool bigonly = argc > 1 && strcmp(argv[1], "big") == 0;   // profiling runs: large shapes only, 2 timed launches
  int fails = 0;
  // name, M,N,K, a_major,b_major, a_rpg,b_rpg, a_rs,a_gs,b_rs,b_gs, c_rpg,c_rs,c_gs, act,act_bwd,bias,res,pre,c_f32,acc, alpha
  std::vector<Case> small = {
      {"nt_basic", 128, 128, 64, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1.f},
      {"nt_k256", 128, 128, 256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1.f},
      {"nt_tails", 200, 72, 136, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 1.f},
      {"nt_bn256", 300, 256, 320, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1.f},
      {"nt_multi_tile", 1000, 768, 256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 1, 1, 1, 0, 0, 0.5f},
      {"nn_dx (B mn-major)", 256, 192, 128, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1.f},
      {"nn_dx_relu_mask", 333, 256, 1024, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"nn_dx_gelu_grad", 333, 384, 512, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 1.f},
      {"nn_dx_mul_saved_grad", 333, 384, 512, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 0, 0, 0, 0, 0, 1.f},
      {"nt_gelu_save_dgrad", 1000, 768, 256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 0, 1, 0, 1, 0, 0, 1.f},
      {"tn_dw (A,B mn-major)", 256, 128, 512, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1.f},
      {"tn_dw_splitk_acc", 256, 320, 5000, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1.f},
      {"tn_dw_a_only", 192, 128, 304, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1.f},
      // conv1d-style: A rows overlap (row stride 4*C, K = 25*C), groups = batch items, padded group stride
      {"conv1_view C=32", 4 * 256, 256, 800, 0, 0, 256, 0, 128, 1048 * 32, 0, 0, 256, 256, 280 * 256, 1, 0, 1, 0, 0, 0, 0, 1.f},
      {"conv2_view rpg=64", 6 * 64, 256, 6400, 0, 0, 64, 0, 1024, 280 * 256, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1.f},
      // conv dW: reduction over (batch, t) rows; B = overlapping view as mn-major, A = dY mn-major
      {"conv_dw_view", 256, 800, 4 * 256, 1, 1, 0, 256, 0, 0, 128, 1048 * 32, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1.f},
      {"small_n3 (fma route)", 256, 3, 256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 1.f},
  };
  for (auto& c : small) {
    if (bigonly) break;
    fails += run_case(c, EGB_BF16, true, 0);
    fails += run_case(c, EGB_F32, true, 0);
  }
  if (!quick) {
    std::vector<Case> big = {
        {"vit_qkv", 50432, 2304, 768, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1.f},
        {"vit_proj_res", 50432, 768, 768, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 1.f},
        {"vit_fc1_gelu", 50432, 3072, 768, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 1, 0, 1, 0, 0, 1.f},
        {"vit_proj_nores", 50432, 768, 768, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1.f},
        {"vit_fc1_gelu_dgrad", 50432, 3072, 768, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 0, 1, 0, 1, 0, 0, 1.f},
        {"vit_dx_fc2_mul", 50432, 3072, 768, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 0, 0, 0, 0, 0, 1.f},
        {"eeg_proj_res", 71168, 256, 256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 1.f},
        {"vit_fc2", 50432, 768, 3072, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 1.f},
        {"vit_dx_fc2", 50432, 3072, 768, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 1.f},
        {"vit_dw_fc1", 3072, 768, 50432, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1.f},
        {"eeg_ffn1", 71168, 1024, 256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1.f},
        {"eeg_dw_qproj", 256, 256, 71168, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1.f},
    };
    const char* only = argc > 2 ? argv[2] : nullptr;   // ./gemm_test big <substring>: just the matching big cases, 10 timed launches
    for (auto& c : big) {
      if (only != nullptr && strstr(c.name, only) == nullptr) continue;
      fails += run_case(c, EGB_BF16, false, c.accumulate ? 0 : (bigonly && only == nullptr ? 2 : 10));
    }
    // fp32 FFMA kernel throughput on one mid-size problem
    Case f = {"f32_ffn1", 8192, 1024, 256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 1, 0, 1.f};
    fails += run_case(f, EGB_F32, false, 10);
  }
  printf("gemm_test: %d failing case(s); launches=%lld\n", fails, (long long)egb_launch_count());
  return fails ? 1 : 0;
}
