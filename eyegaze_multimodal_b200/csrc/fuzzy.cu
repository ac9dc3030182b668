// FuzzyGatingFusion (3_Models/fusion/fuzzy_gating_fusion.py:297-390) forward and backward as one kernel
// each: the reference issues ~40 ATen micro-kernels on (B,3)/(B,)/(B,4) tensors; here one thread owns one
// trial and the 13 scalar parameter gradients are block-reduced with warp shuffles.
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

extern void egb_count_launch(int n);

namespace {

constexpr int FZ_MAXC = 16;  // num_classes <= 16

struct FuzzyParams {
  const float *tau_img, *tau_eeg, *c_rel, *c_unrel_img, *c_unrel_eeg;
  const float *ls_rel_img, *ls_rel_eeg, *ls_unrel_img, *ls_unrel_eeg, *beta;
  int mode;  // 0 full, 1 no_temperature, 2 no_fuzzification, 3 fixed_weights
  int B, C;
  float eps_temp, eps_log, eps_div, max_entropy;
};

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

struct RowState {
  float z_i[FZ_MAXC], z_e[FZ_MAXC], p_i[FZ_MAXC], p_e[FZ_MAXC];
  float T_i, T_e, H_i, H_e;
  float mu[4];      // img_rel, img_unrel, eeg_rel, eeg_unrel
  float w[4], theta[4];
  float num, den, alpha_raw, alpha;
  float c_i, c_e;   // no_fuzzification confidences (pre-clamp)
};

__device__ __forceinline__ float entropy_of(const float* z, float* p, int C, float eps_log) {
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, z[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) { p[c] = expf(z[c] - mx); s += p[c]; }
  float h = 0.f;
  for (int c = 0; c < C; ++c) { p[c] /= s; h -= p[c] * logf(p[c] + eps_log); }
  return h;
}

__device__ __forceinline__ void fuzzy_row(const FuzzyParams& P, const float* xi, const float* xe, RowState& r) {
  const bool temp = (P.mode == 0 || P.mode == 2);
  r.T_i = temp ? softplus_f(*P.tau_img) + P.eps_temp : 1.f;
  r.T_e = temp ? softplus_f(*P.tau_eeg) + P.eps_temp : 1.f;
  for (int c = 0; c < P.C; ++c) { r.z_i[c] = xi[c] / r.T_i; r.z_e[c] = xe[c] / r.T_e; }
  r.H_i = entropy_of(r.z_i, r.p_i, P.C, P.eps_log);
  r.H_e = entropy_of(r.z_e, r.p_e, P.C, P.eps_log);
  if (P.mode == 3) {
    r.alpha = r.alpha_raw = 0.5f;
  } else if (P.mode == 2) {
    r.c_i = 1.f - r.H_i / (P.max_entropy + P.eps_div);
    r.c_e = 1.f - r.H_e / (P.max_entropy + P.eps_div);
    const float ci = fmaxf(r.c_i, 0.f), ce = fmaxf(r.c_e, 0.f);
    r.alpha_raw = ci / (ci + ce + P.eps_div);
    r.alpha = fminf(fmaxf(r.alpha_raw, 0.f), 1.f);
  } else {
    const float cen[4] = {*P.c_rel, *P.c_unrel_img, *P.c_rel, *P.c_unrel_eeg};
    const float ls[4] = {*P.ls_rel_img, *P.ls_unrel_img, *P.ls_rel_eeg, *P.ls_unrel_eeg};
    for (int k = 0; k < 4; ++k) {
      const float h = k < 2 ? r.H_i : r.H_e;
      const float sg = expf(ls[k]);
      r.mu[k] = expf(-((h - cen[k]) * (h - cen[k])) / (2.f * sg * sg + P.eps_div));
    }
    r.w[0] = r.mu[0] * r.mu[3];  // img rel  & eeg unrel
    r.w[1] = r.mu[1] * r.mu[2];  // img unrel & eeg rel
    r.w[2] = r.mu[0] * r.mu[2];  // both rel
    r.w[3] = r.mu[1] * r.mu[3];  // both unrel
    r.num = 0.f;
    r.den = P.eps_div;
    for (int k = 0; k < 4; ++k) {
      r.theta[k] = 1.f / (1.f + expf(-P.beta[k]));
      r.num += r.w[k] * r.theta[k];
      r.den += r.w[k];
    }
    r.alpha_raw = r.num / r.den;
    r.alpha = fminf(fmaxf(r.alpha_raw, 0.f), 1.f);
  }
}

// aux layout per row (16 floats): H_img, H_eeg, mu[4], w[4], alpha, T_img, T_eeg, pad
__global__ void fuzzy_fwd_kernel(const FuzzyParams P, const float* __restrict__ img, const float* __restrict__ eeg,
                                 float* __restrict__ fused, float* __restrict__ alpha, float* __restrict__ aux) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  RowState r;
  fuzzy_row(P, img + (long long)b * P.C, eeg + (long long)b * P.C, r);
  for (int c = 0; c < P.C; ++c) fused[(long long)b * P.C + c] = r.alpha * r.z_i[c] + (1.f - r.alpha) * r.z_e[c];
  alpha[b] = r.alpha;
  float* a = aux + (long long)b * 16;
  a[0] = r.H_i; a[1] = r.H_e;
  for (int k = 0; k < 4; ++k) { a[2 + k] = (P.mode < 2) ? r.mu[k] : 0.f; a[6 + k] = (P.mode < 2) ? r.w[k] : 0.f; }
  a[10] = r.alpha; a[11] = r.T_i; a[12] = r.T_e;
  if (b == 0) {  // parameter-derived analysis values (aux_info['fuzz_params'] / ['consequents']), once per call
    float* q = aux + (long long)P.B * 16;
    q[0] = expf(*P.ls_rel_img); q[1] = expf(*P.ls_rel_eeg); q[2] = expf(*P.ls_unrel_img); q[3] = expf(*P.ls_unrel_eeg);
    for (int k = 0; k < 4; ++k) q[4 + k] = 1.f / (1.f + expf(-P.beta[k]));
    q[8] = r.T_i; q[9] = r.T_e;
  }
}

// dparams (16 floats, accumulated): tau_img, tau_eeg, c_unrel_img, c_unrel_eeg, ls_rel_img, ls_rel_eeg,
// ls_unrel_img, ls_unrel_eeg, beta[4]
__global__ void fuzzy_bwd_kernel(const FuzzyParams P, const float* __restrict__ img, const float* __restrict__ eeg,
                                 const float* __restrict__ g_fused, const float* __restrict__ g_alpha,
                                 float* __restrict__ d_img, float* __restrict__ d_eeg, float* __restrict__ dparams) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float dp[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) dp[k] = 0.f;
  if (b < P.B) {
    RowState r;
    const float* xi = img + (long long)b * P.C;
    const float* xe = eeg + (long long)b * P.C;
    fuzzy_row(P, xi, xe, r);
    const float* g = g_fused + (long long)b * P.C;
    float dalpha = g_alpha != nullptr ? g_alpha[b] : 0.f;
    float dz_i[FZ_MAXC], dz_e[FZ_MAXC];
    for (int c = 0; c < P.C; ++c) {
      dalpha += g[c] * (r.z_i[c] - r.z_e[c]);
      dz_i[c] = r.alpha * g[c];
      dz_e[c] = (1.f - r.alpha) * g[c];
    }
    float dH_i = 0.f, dH_e = 0.f;
    const bool pass = (r.alpha_raw >= 0.f && r.alpha_raw <= 1.f);
    if (P.mode == 2 && pass) {
      const float ci = fmaxf(r.c_i, 0.f), ce = fmaxf(r.c_e, 0.f);
      const float s = ci + ce + P.eps_div;
      const float dci = dalpha * (ce + P.eps_div) / (s * s);
      const float dce = -dalpha * ci / (s * s);
      if (r.c_i >= 0.f) dH_i = -dci / (P.max_entropy + P.eps_div);
      if (r.c_e >= 0.f) dH_e = -dce / (P.max_entropy + P.eps_div);
    } else if (P.mode < 2 && pass) {
      const float dnum = dalpha / r.den;
      const float dden = -dalpha * r.num / (r.den * r.den);
      float dw[4];
      for (int k = 0; k < 4; ++k) {
        dw[k] = dnum * r.theta[k] + dden;
        dp[8 + k] = dnum * r.w[k] * r.theta[k] * (1.f - r.theta[k]);
      }
      float dmu[4];
      dmu[0] = dw[0] * r.mu[3] + dw[2] * r.mu[2];
      dmu[1] = dw[1] * r.mu[2] + dw[3] * r.mu[3];
      dmu[2] = dw[1] * r.mu[1] + dw[2] * r.mu[0];
      dmu[3] = dw[0] * r.mu[0] + dw[3] * r.mu[1];
      const float cen[4] = {*P.c_rel, *P.c_unrel_img, *P.c_rel, *P.c_unrel_eeg};
      const float ls[4] = {*P.ls_rel_img, *P.ls_unrel_img, *P.ls_rel_eeg, *P.ls_unrel_eeg};
      for (int k = 0; k < 4; ++k) {
        const float h = k < 2 ? r.H_i : r.H_e;
        const float sg2 = expf(2.f * ls[k]);
        const float D = 2.f * sg2 + P.eps_div;
        const float diff = h - cen[k];
        const float gmu = dmu[k] * r.mu[k];
        const float dh = gmu * (-2.f * diff / D);
        if (k < 2) dH_i += dh; else dH_e += dh;
        const float dc = gmu * (2.f * diff / D);
        const float dls = gmu * (diff * diff / (D * D)) * 4.f * sg2;
        // parameter slots: c_unrel_img=2, c_unrel_eeg=3, ls_rel_img=4, ls_rel_eeg=5, ls_unrel_img=6, ls_unrel_eeg=7
        if (k == 1) dp[2] += dc;
        if (k == 3) dp[3] += dc;
        if (k == 0) dp[4] += dls;
        if (k == 2) dp[5] += dls;
        if (k == 1) dp[6] += dls;
        if (k == 3) dp[7] += dls;
      }
    }
    // entropy -> softmax -> logits
    for (int pass_i = 0; pass_i < 2; ++pass_i) {
      const float* p = pass_i == 0 ? r.p_i : r.p_e;
      float* dz = pass_i == 0 ? dz_i : dz_e;
      const float dH = pass_i == 0 ? dH_i : dH_e;
      if (dH != 0.f) {
        float dpv[FZ_MAXC], dot = 0.f;
        for (int c = 0; c < P.C; ++c) {
          dpv[c] = -dH * (logf(p[c] + P.eps_log) + p[c] / (p[c] + P.eps_log));
          dot += p[c] * dpv[c];
        }
        for (int c = 0; c < P.C; ++c) dz[c] += p[c] * (dpv[c] - dot);
      }
    }
    const bool temp = (P.mode == 0 || P.mode == 2);
    float dT_i = 0.f, dT_e = 0.f;
    for (int c = 0; c < P.C; ++c) {
      d_img[(long long)b * P.C + c] = dz_i[c] / r.T_i;
      d_eeg[(long long)b * P.C + c] = dz_e[c] / r.T_e;
      dT_i -= dz_i[c] * r.z_i[c] / r.T_i;
      dT_e -= dz_e[c] * r.z_e[c] / r.T_e;
    }
    if (temp) {
      dp[0] = dT_i / (1.f + expf(-*P.tau_img));
      dp[1] = dT_e / (1.f + expf(-*P.tau_eeg));
    }
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float v = warp_sum(dp[k]);
    if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dparams + k, v);
  }
}

}  // namespace

extern "C" {



static int fill_fuzzy(const egb_fuzzy_desc* d, FuzzyParams* p) {
  EGB_CHECK(d->num_classes >= 1 && d->num_classes <= FZ_MAXC, "fuzzy: num_classes %d unsupported", d->num_classes);
  EGB_CHECK(d->mode >= 0 && d->mode <= 3, "fuzzy: bad mode %d", d->mode);
  p->tau_img = d->tau_img; p->tau_eeg = d->tau_eeg; p->c_rel = d->c_reliable;
  p->c_unrel_img = d->c_unreliable_img; p->c_unrel_eeg = d->c_unreliable_eeg;
  p->ls_rel_img = d->log_sigma_reliable_img; p->ls_rel_eeg = d->log_sigma_reliable_eeg;
  p->ls_unrel_img = d->log_sigma_unreliable_img; p->ls_unrel_eeg = d->log_sigma_unreliable_eeg;
  p->beta = d->beta; p->mode = d->mode; p->B = d->B; p->C = d->num_classes;
  p->eps_temp = d->eps_temp; p->eps_log = d->eps_log; p->eps_div = d->eps_div;
  p->max_entropy = logf((float)d->num_classes);
  return 0;
}

int egb_fuzzy_fwd(const egb_fuzzy_desc* d, const float* img, const float* eeg, float* fused, float* alpha, float* aux,
                  void* stream) {
  FuzzyParams p;
  if (fill_fuzzy(d, &p)) return 1;
  if (d->B == 0) return 0;
  fuzzy_fwd_kernel<<<(d->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p, img, eeg, fused, alpha, aux);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

int egb_fuzzy_bwd(const egb_fuzzy_desc* d, const float* img, const float* eeg, const float* g_fused,
                  const float* g_alpha, float* d_img, float* d_eeg, float* dparams, void* stream) {
  FuzzyParams p;
  if (fill_fuzzy(d, &p)) return 1;
  if (d->B == 0) return 0;
  fuzzy_bwd_kernel<<<(d->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p, img, eeg, g_fused, g_alpha, d_img, d_eeg,
                                                                       dparams);
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
