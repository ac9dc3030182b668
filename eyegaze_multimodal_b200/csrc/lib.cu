// Library-wide state: error string, launch counter, device properties.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "../../include/eyegaze_b200.h"

static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};

void egb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void egb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// seed epoch word, one per device (one process drives one GPU; the table covers a multi-device process too)
static unsigned long long* g_epoch[64] = {nullptr};
const unsigned long long* egb_seed_epoch_ptr() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  return g_epoch[dev];
}
static __global__ void seed_epoch_advance_kernel(unsigned long long* e) { *e += 1ull; }

int egb_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

// ---------------------------------------------------------------------------------------------
// optional per-launch timing of the dominant (GEMM) kernel with CUDA events on the launching stream
// (bench.py's roofline block).  Disabled by default: zero overhead on the normal path.
// ---------------------------------------------------------------------------------------------
#include <vector>
struct ProfRec { cudaEvent_t e0, e1; double flops; double bytes; int kind; double tag[4]; };
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static std::mutex g_prof_mu;
static int g_prof_on = 0;

int egb_prof_enabled() { return g_prof_on; }

static cudaEvent_t prof_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
// called by the GEMM launcher right before / after its kernel launch
void egb_prof_begin(cudaStream_t st, double flops, double bytes, int kind) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.e0 = prof_event(); r.e1 = prof_event(); r.flops = flops; r.bytes = bytes; r.kind = kind;
  r.tag[0] = r.tag[1] = r.tag[2] = r.tag[3] = 0.0;
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
}
// shape / variant of the launch just begun (per-launch table of egb_prof_dump)
void egb_prof_tag(double a, double b, double c, double d) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) { double* t = g_prof.back().tag; t[0] = a; t[1] = b; t[2] = c; t[3] = d; }
}
void egb_prof_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().e1, st);
}

extern "C" {
/* kind: 0 = tcgen05 GEMM, 1 = FFMA GEMM.  out[0..3] = launches, total ms, total flops, total algorithmic bytes */
int egb_prof_enable(int on) { g_prof_on = on; return 0; }
/* per-launch records of `kind` (in launch order): out[i*8 ..] = {ms, flops, bytes, tag0..tag3, 0}; *n_out = records written */
int egb_prof_dump(int kind, double* out, int max_records, int* n_out) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (auto& r : g_prof) {
    if (r.kind != kind || n >= max_records) continue;
    float t = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
      double* o = out + (size_t)n * 8;
      o[0] = t; o[1] = r.flops; o[2] = r.bytes; o[3] = r.tag[0]; o[4] = r.tag[1]; o[5] = r.tag[2]; o[6] = r.tag[3]; o[7] = 0.0;
      ++n;
    }
  }
  *n_out = n;
  return 0;
}
int egb_prof_read(int kind, double* out, int reset) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double n = 0, ms = 0, fl = 0, by = 0;
  for (auto& r : g_prof) {
    if (r.kind != kind) continue;
    float t = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
      n += 1; ms += t; fl += r.flops; by += r.bytes;
    }
  }
  out[0] = n; out[1] = ms; out[2] = fl; out[3] = by;
  if (reset) {
    for (auto& r : g_prof) { g_event_pool.push_back(r.e0); g_event_pool.push_back(r.e1); }
    g_prof.clear();
  }
  return 0;
}
/* creates (once per device) the device word that all dropout seeds are mixed with; returns its address in *out */
int egb_seed_epoch_enable(void** out) {
  int dev = 0;
  EGB_CUDA(cudaGetDevice(&dev));
  EGB_CHECK(dev >= 0 && dev < 64, "seed_epoch: device index out of range");
  if (g_epoch[dev] == nullptr) {
    unsigned long long* p = nullptr;
    EGB_CUDA(cudaMalloc(&p, sizeof(unsigned long long)));
    EGB_CUDA(cudaMemset(p, 0, sizeof(unsigned long long)));
    g_epoch[dev] = p;
  }
  if (out != nullptr) *out = g_epoch[dev];
  return 0;
}
int egb_seed_epoch_advance(void* stream) {
  const unsigned long long* p = egb_seed_epoch_ptr();
  EGB_CHECK(p != nullptr, "seed_epoch_advance: call egb_seed_epoch_enable first");
  seed_epoch_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(const_cast<unsigned long long*>(p));
  egb_count_launch(1);
  EGB_LAUNCH_CHECK();
  return 0;
}
const char* egb_last_error(void) { return g_err; }
int egb_version(void) { return 1; }
int64_t egb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
}
